// kernel_stats.cuh -- k_stats: record gate + scalar counters (src/bamqualcheck.cpp:313-434), QualityCheck
// (src/QualityCheck.hpp:111-271) and the TripletCounting walk (src/TripletCounting.hpp:136-236).
// Included by kernels.cuh (inside namespace bqc).
//
// One record per lane.  Control flow is organised in warp-uniform phases separated by __syncwarp():
//   A  decode, CIGAR summary, aux walk, gates           (short, divergent per lane)
//   B  triplet walk of the eligible lanes               (rolling 3-base windows over read and reference)
//   C  per-cycle base / quality pass                    (common trip count = longest read of the warp)
//   D  per-read histogram bumps, main-chromosome block
// All tables of the CTA live in shared memory (u32) and are flushed with 64-bit REDs at the end.
#pragma once

namespace bqc {

// per-base stream over the 4-bit SEQ field: 16 bases per 8-byte chunk, next base in the low nibble
struct NibStream {
    const uint8_t* p;
    uint64_t w;
    __device__ __forceinline__ void seek(const uint8_t* seq, uint32_t i) {  // position on base i
        p = seq + ((i >> 4) << 3);
        w = swap_nibbles(ldu64(p)) >> (4 * (i & 15u));
    }
    static __device__ __forceinline__ uint64_t swap_nibbles(uint64_t x) {
        return ((x & 0x0F0F0F0F0F0F0F0FULL) << 4) | ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL);
    }
    // base i (the caller passes i so that chunk boundaries are detected without extra state)
    __device__ __forceinline__ uint32_t get(uint32_t i) {
        if ((i & 15u) == 0) { w = swap_nibbles(ldu64(p)); }
        uint32_t v = (uint32_t)w & 15u;
        w >>= 4;
        if ((i & 15u) == 15u) p += 8;
        return v;
    }
};

__global__ void __launch_bounds__(kStatsThreads) k_stats(EngineView E, BatchView B, uint32_t lane) {
    extern __shared__ uint32_t sm[];
    const StatsSmem S = stats_smem_layout(B.cycb, E.insert_smem);
    for (uint32_t i = threadIdx.x; i < S.total; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    const Layout& L = E.L;
    uint64_t* G = E.counters + (uint64_t)lane * L.lane_stride;
    const uint32_t cycb = B.cycb;
    const uint32_t lane_id = threadIdx.x & 31u;
    // Dna5 ordinal per BAM nibble, 4 bits per entry: forward view and reverse-complement view (R5, R7)
    const uint64_t LUT_FWD = 0x4444444344424104ULL;  // nib 1->0, 2->1, 4->2, 8->3, else 4
    const uint64_t LUT_REV = 0x4444444044414234ULL;  // nib 1->3, 2->2, 4->1, 8->0, else 4

    for (uint32_t r0 = blockIdx.x * blockDim.x + threadIdx.x - lane_id; r0 < B.n_records; r0 += gridDim.x * blockDim.x) {
        // ---------------------------------------------------------------- phase A
        const uint32_t rec = r0 + lane_id;
        const uint64_t grec = B.first_record + rec;
        RecHdr h;
        bool live = rec < B.n_records && !(B.rec_lane && B.rec_lane[rec] != lane);
        if (live) {
            const uint32_t off = B.offsets[rec];
            if (!decode_hdr(B.bytes + off, B.offsets[rec + 1] - off, h)) { report_error(E, grec, 4); live = false; }
        }
        uint32_t flag = 0, Ls = 0, mate = 0, delc = 0, insc = 0, first_op = 0, last_op = 0;
        bool isfirst = false, rc = false, mapped = false, inmain = false, do_trip = false;
        uint64_t* GM = G;
        if (live) {
            flag = h.flag;
            Ls = (uint32_t)h.lseq;
            const bool primary = !(flag & 0x900u);
            const bool hasmate = (flag & 0xC0u) != 0;
            isfirst = (flag & 0x40u) != 0;
            mate = isfirst ? 0u : 1u;
            rc = (flag & 0x10u) != 0;
            mapped = !(flag & 0x4u);
            inmain = h.rid >= 0 && h.rid < E.n_ref && E.main_chrom[h.rid];
            GM = G + L.o_mate0 + mate * L.mate_stride;
            // CIGAR summary (needed by mis_match, cigar_count and the triplet filter)
            uint32_t clipped = 0;
            if (primary) {
                for (uint32_t i = 0; i < h.ncig; ++i) {
                    uint32_t c = ldu32(h.p + h.o_cig + 4 * i);
                    uint32_t op = c & 15u, n = c >> 4;
                    if (op == 2) delc += n;
                    else if (op == 1) insc += n;
                    else if (op == 4 || op == 5) clipped += n;
                    if (i == 0) first_op = c;
                    if (i == h.ncig - 1) last_op = c;
                }
            }
            const bool do_cig = primary && hasmate && inmain && mapped;  // src/bamqualcheck.cpp:392-428
            AuxInfo ai = aux_walk(h, [&](uint32_t nm) {
                if (do_cig) bump(sm + S.mm + mate * kHS, kHS, GM + L.m_mismatch, L.mmcap, nm - delc - insc, E, grec);
            });
            // getLane(): src/bamqualcheck.cpp:72-100, then the gate :318-335
            if (ai.rg == 2) { report_error(E, grec, 1); live = false; }
            else if (ai.rg == 0) { report_error(E, grec, 5); live = false; }
            else if (flag & 0x800u) { atomicAdd(sm + S.sc + S_SUPPLEMENTARY, 1u); live = false; }
            else if (flag & 0x100u) { atomicAdd(sm + S.sc + S_NOT_PRIMARY, 1u); live = false; }
            else {
                if (flag & 0x400u) atomicAdd(sm + S.sc + S_DUPLICATES, 1u);
                if (flag & 0x200u) atomicAdd(sm + S.sc + S_QCFAILED, 1u);
                if (Ls > L.cyc || Ls > cycb) { report_error(E, grec, 16); live = false; }
            }
            // TripletCounting filter (src/TripletCounting.hpp:136-168)
            if (live && !(flag & 0x600u)) {
                bool elig = (flag & 0x1u) && (flag & 0x2u) && mapped && !(flag & 0x8u) && h.mapq >= 60u;
                if (elig) {
                    if (ai.as_state != 1 || ai.as_value < 0) { report_error(E, grec, 3); live = false; elig = false; }
                    else elig = ai.as_value >= 50 && clipped == 0;
                }
                do_trip = elig && E.ref && h.rid >= 0 && h.rid < E.n_ref && E.ref[h.rid] != nullptr && h.ncig > 0 && Ls >= 3 && h.pos >= 0;
            }
            if (live && !hasmate) { report_error(E, grec, 2); live = false; do_trip = false; }  // :385-389
        }
        __syncwarp();
        // ---------------------------------------------------------------- phase B: triplet walk (:195-236)
        if (do_trip) {
            const uint32_t* __restrict__ ref = E.ref[h.rid];
            const uint64_t reflen = E.ref_len[h.rid];
            const uint32_t refmax = (uint32_t)((reflen + 15) / 16) + 3u;  // last allocated word (reads that overhang the contig)
            uint32_t* tri = sm + S.tri + ((rc ? 2u : 0u) + (isfirst ? 0u : 1u)) * 4u;
            const uint8_t* seqp = h.p + h.o_seq;
            const uint8_t* qualp = h.p + h.o_qual;
            uint32_t it = 0;
            uint32_t cc = (first_op >> 4) - 1u;      // wraps for a zero count exactly like the size_t in the reference
            uint32_t chromPos = (uint32_t)h.pos + 1u;
            uint32_t readPos = 1;
            const uint32_t last = Ls - 1;
            bool reload = true, ok = true;
            NibStream ns;                             // positioned on readPos + 1
            uint32_t nprev = 0, ncur = 0, rprev = 0, rcur = 0;
            uint64_t qw = 0;                          // QUAL chunk, current byte in the low 8 bits
            uint32_t rw = 0;                          // reference chunk (16 bases), next base in the low 2 bits
            for (; readPos < last; ++readPos, ++chromPos, --cc) {
                if (cc == 0) {
                    do {
                        ++it;
                        if (it >= h.ncig) { ok = false; break; }  // the reference reads past the CIGAR here (undefined)
                        uint32_t c = ldu32(h.p + h.o_cig + 4 * it);
                        uint32_t op = c & 15u, n = c >> 4;
                        if (op == 2 || op == 3 || op == 5 || op == 6) chromPos += n;
                        else if (op == 4 || op == 1) readPos += n;
                        else cc = n;
                    } while (cc == 0);
                    if (!ok || readPos >= last) break;
                    reload = true;
                }
                if (reload) {  // (re)position the three streams: read bases, qualities, reference bases
                    ns.seek(seqp, readPos - 1);
                    nprev = ns.get(readPos - 1);
                    ncur = ns.get(readPos);
                    qw = ldu64(qualp + (readPos & ~7u)) >> (8 * (readPos & 7u));
                    const uint32_t cp = chromPos - 1;
                    rw = __ldg(ref + min(cp >> 4, refmax)) >> (2 * (cp & 15u));
                    rprev = rw & 3u;
                    rw >>= 2;
                    if (((cp + 1) & 15u) == 0) rw = __ldg(ref + min((cp + 1) >> 4, refmax));
                    rcur = rw & 3u;
                    rw >>= 2;
                    reload = false;
                } else if ((readPos & 7u) == 0) {
                    qw = ldu64(qualp + readPos);
                }
                if (((chromPos + 1) & 15u) == 0) rw = __ldg(ref + min((chromPos + 1) >> 4, refmax));
                const uint32_t nnext = ns.get(readPos + 1);
                const uint32_t rnext = rw & 3u;
                rw >>= 2;
                const uint32_t q = (uint32_t)qw & 255u;
                qw >>= 8;
                // quality >= '5' as signed chars, base and both flanks A/C/G/T, flanks equal to the reference
                // context (char compare), context inside the contig
                const uint32_t base = (uint32_t)(LUT_FWD >> (4 * ncur)) & 7u;
                bool cnt = (int8_t)(q + 33u) >= (int8_t)53 && base != 4u && nprev == (1u << rprev) && nnext == (1u << rnext) &&
                           (uint64_t)chromPos + 2 <= reflen;
                if (cnt) atomicAdd(tri + ((rprev << 4) + (rcur << 2) + rnext) * 16u + base, 1u);
                nprev = ncur;
                ncur = nnext;
                rprev = rcur;
                rcur = rnext;
            }
        }
        __syncwarp();
        // ---------------------------------------------------------------- phase C: get_count (:111-176)
        const uint32_t myL = live ? Ls : 0u;
        const uint32_t maxL = __reduce_max_sync(0xFFFFFFFFu, myL);
        uint32_t cntN = 0, cntGC = 0, sumQ = 0;
        {
            uint32_t* pc = sm + S.pc + mate * PC_ROWS * cycb;
            const uint8_t* seqp = live ? h.p + h.o_seq : B.bytes;
            const uint8_t* qualp = live ? h.p + h.o_qual : B.bytes;
            const uint64_t lut = rc ? LUT_REV : LUT_FWD;
            uint32_t cyc = rc ? Ls - 1u : 0u;
            const uint32_t step = rc ? 0xFFFFFFFFu : 1u;
            uint64_t seqw = 0, qualw = 0;
            for (uint32_t i = 0; i < maxL; ++i) {
                if (i < myL) {
                    if ((i & 15u) == 0) seqw = NibStream::swap_nibbles(ldu64(seqp + (i >> 1)));
                    if ((i & 7u) == 0) qualw = ldu64(qualp + i);
                    const uint32_t nb = (uint32_t)seqw & 15u;
                    seqw >>= 4;
                    const uint32_t q = (uint32_t)qualw & 255u;
                    qualw >>= 8;
                    const uint32_t d = (uint32_t)(lut >> (4 * nb)) & 7u;
                    atomicAdd(pc + d * cycb + cyc, 1u);
                    atomicAdd(pc + PC_QUAL * cycb + cyc, q);
                    cntN += (nb + 1u) >> 4;
                    cntGC += (0x14u >> nb) & 1u;
                    sumQ += q;
                    cyc += step;
                }
            }
        }
        // ---------------------------------------------------------------- phase D
        if (live) {
            atomicAdd(sm + S.sc + S_READCOUNT, 1u);
            atomicAdd(sm + S.sc + S_TOTALBPS, Ls);
            atomicAdd((unsigned long long*)(GM + L.m_readnr), 1ULL);
            atomicAdd(sm + S.nc + mate * (cycb + 8) + cntN, 1u);
            atomicAdd(sm + S.gc + mate * (cycb + 8) + cntGC, 1u);
            if (Ls > 0) {
                uint32_t rnd = (2u * sumQ + Ls) / (2u * Ls);   // round(double(S)/L), exact
                uint32_t cel = (sumQ + Ls - 1u) / Ls;          // ceil(double(S)/L)
                bump(sm + S.aq + mate * kQS, kQS, GM + L.m_avgq, kQCap, rnd, E, grec);
                bump(sm + S.cq + mate * kQS, kQS, GM + L.m_ceilq, kQCap, cel, E, grec);
            }
            atomicAdd(sm + S.rl + mate * (cycb + 8) + Ls, 1u);
            if (isfirst) {  // src/bamqualcheck.cpp:359-374
                if (!mapped) {
                    atomicAdd(sm + S.sc + S_FIRSTUNMAPPED, 1u);
                    if (flag & 0x8u) atomicAdd(sm + S.sc + S_BOTHUNMAPPED, 1u);
                }
                if (flag & 0x2u) {
                    atomicAdd(sm + S.sc + S_PROPERPAIR, 1u);
                    if (((flag >> 4) & 1u) == ((flag >> 5) & 1u)) atomicAdd(sm + S.sc + S_FF_RR, 1u);
                }
            } else if (!mapped) {
                atomicAdd(sm + S.sc + S_SECONDUNMAPPED, 1u);
            }
            if (inmain) {  // main chromosomes only: src/bamqualcheck.cpp:392-434
                if (mapped) {
                    // cigar_count (src/QualityCheck.hpp:222-271) on the read-oriented CIGAR
                    if (h.ncig == 0) {
                        report_error(E, grec, 16);
                    } else {
                        uint32_t fo = rc ? last_op : first_op, lo = rc ? first_op : last_op;
                        uint32_t* pc = sm + S.pc + mate * PC_ROWS * cycb;
                        if ((fo & 15u) == 4u) {
                            uint32_t n = min(fo >> 4, cycb);
                            for (uint32_t j = 0; j < n; ++j) atomicAdd(pc + PC_SC5 * cycb + j, 1u);
                        } else if ((lo & 15u) == 4u) {
                            uint32_t n = lo >> 4;
                            for (uint32_t j = (n <= Ls ? Ls - n : Ls); j < Ls; ++j) atomicAdd(pc + PC_SC3 * cycb + j, 1u);
                        }
                        bump(sm + S.dl + mate * kHS, kHS, GM + L.m_del, L.delcap, delc, E, grec);
                        bump(sm + S.in + mate * kHS, kHS, GM + L.m_ins, L.mmcap, insc, E, grec);
                        atomicAdd(sm + S.mq + mate * kMapqCap + h.mapq, 1u);  // map_Q :178-185
                        if (isfirst && !(flag & 0x8u) && h.nrid >= 0 && h.nrid < E.n_ref && E.main_chrom[h.nrid]) {
                            uint32_t idx = (uint32_t)(h.tlen < 0 ? -(int64_t)h.tlen : (int64_t)h.tlen);  // insert_size :187-196
                            if (idx >= L.isize1) idx = L.isize1 - 1;
                            if (idx < E.insert_smem) atomicAdd(sm + S.isz + idx, 1u);
                            else atomicAdd((unsigned long long*)(G + L.o_insert + idx), 1ULL);
                        }
                    }
                }
                if (isfirst) {
                    if ((mapped || !(flag & 0x8u)) && !(flag & 0x400u)) atomicAdd(sm + S.sc + S_FIRST_AND_OR_SECOND_MAPPED, 1u);
                    if ((flag & 0x2u) && !(flag & 0x400u)) atomicAdd(sm + S.sc + S_AUTO_PROPERPAIR, 1u);
                }
                // OverallNumbers::coverage (:430-433) is handled by k_cov_scatter + k_cov_flush.
            }
        }
        __syncwarp();
    }
    __syncthreads();
    // ---- flush the CTA-private tables (skip zeros) ---------------------------------------------------
    auto flush = [&](uint32_t smo, uint32_t n, uint64_t* g) {
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            uint32_t v = sm[smo + i];
            if (v) atomicAdd((unsigned long long*)(g + i), (unsigned long long)v);
        }
    };
    for (uint32_t m = 0; m < 2; ++m) {
        uint64_t* GMm = G + L.o_mate0 + m * L.mate_stride;
        for (uint32_t r = 0; r < PC_ROWS; ++r) flush(S.pc + (m * PC_ROWS + r) * cycb, cycb, GMm + L.m_pc + r * pad8(L.cyc));
        flush(S.rl + m * (cycb + 8), cycb + 1, GMm + L.m_readlen);
        flush(S.nc + m * (cycb + 8), cycb + 1, GMm + L.m_ncount);
        flush(S.gc + m * (cycb + 8), cycb + 1, GMm + L.m_gccount);
        flush(S.aq + m * kQS, kQS, GMm + L.m_avgq);
        flush(S.cq + m * kQS, kQS, GMm + L.m_ceilq);
        flush(S.mq + m * kMapqCap, kMapqCap, GMm + L.m_mapq);
        flush(S.mm + m * kHS, kHS, GMm + L.m_mismatch);
        flush(S.dl + m * kHS, kHS, GMm + L.m_del);
        flush(S.in + m * kHS, kHS, GMm + L.m_ins);
    }
    flush(S.isz, E.insert_smem, G + L.o_insert);
    flush(S.tri, kTriplet, G + L.o_triplet);
    flush(S.sc, S_COUNT, G + L.o_scalars);
}

}  // namespace bqc
