// kernel_stats.cuh -- k_stats: record gate + scalar counters (src/bamqualcheck.cpp:313-434), QualityCheck
// (src/QualityCheck.hpp:111-271) and the TripletCounting walk (src/TripletCounting.hpp:136-236).
// Included by kernels.cuh (inside namespace bqc).
//
// One record per lane.  Control flow is organised in warp-uniform phases separated by __syncwarp():
//   A  decode, CIGAR summary, aux walk, gates           (short, divergent per lane)
//   B  triplet walk of the eligible lanes               (CIGAR runs, 14 positions per 16-nibble window: swar.h)
//   C  per-cycle base / quality pass                    (8 cycles per step, common trip count per warp)
//   D  per-read histogram bumps, main-chromosome block
// All tables of the CTA live in shared memory (u32) and are flushed with 64-bit REDs at the end; the record bytes of
// a warp are staged in shared memory with cp.async (k_stats<true>) or read from global memory (k_stats<false>).
#pragma once

namespace bqc {

__device__ __forceinline__ void red_shared_inc(uint32_t saddr) {
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(saddr));
}
__device__ __forceinline__ void red_shared_add(uint32_t saddr, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(saddr), "r"(v));
}

// Phases A-D for the (up to) 32 records of one warp.  `present` lanes hold a record whose bytes start at recp (in the
// warp's staging area: M = SMem, or in global memory: M = GMem) and span `avail` bytes.
template <class M>
__device__ __forceinline__ void stats_records(const EngineView& E, const BatchView& B, uint32_t lane, uint32_t* sm, const StatsSmem& S,
                                              uint32_t rec, bool present, const uint8_t* recp, uint32_t avail) {
    const Layout& L = E.L;
    uint64_t* G = E.counters + (uint64_t)lane * L.lane_stride;
    const uint32_t cycb = B.cycb;
    {
        // ---------------------------------------------------------------- phase A
        const uint64_t grec = B.first_record + rec;
        RecHdr h;
        h.p = recp;
        bool live = present && lane_match(B, rec, lane);
        if (live) {
            if (!decode_hdr<M>(recp, avail, h)) { report_error(E, grec, 4); live = false; }
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
                    uint32_t c = ldu32<M>(h.p + h.o_cig + 4 * i);
                    uint32_t op = c & 15u, n = c >> 4;
                    if (op == 2) delc += n;
                    else if (op == 1) insc += n;
                    else if (op == 4 || op == 5) clipped += n;
                    if (i == 0) first_op = c;
                    if (i == h.ncig - 1) last_op = c;
                }
            }
            const bool do_cig = primary && hasmate && inmain && mapped;  // src/bamqualcheck.cpp:392-428
            AuxInfo ai = aux_walk<M>(h, [&](uint32_t nm) {
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
            // The reference's loop visits the read positions of every match-like CIGAR run; the test at a position
            // only looks at read bases p-1..p+1, QUAL[p] and reference bases c-1..c+1 with c - p constant inside a
            // run.  Runs are enumerated exactly like the loop does (:201-212); inside a run the tests are evaluated
            // 14 positions at a time on a 16-nibble window (swar.h), leaving one predicated shared atomic per position.
            const uint32_t* __restrict__ ref = E.ref[h.rid];
            const uint64_t reflen = E.ref_len[h.rid];
            const uint32_t refmax = (uint32_t)((reflen + 15) / 16) + 3u;  // last allocated word (reads that overhang the contig)
            const uint32_t tri_s = (uint32_t)__cvta_generic_to_shared(sm + S.tri + ((rc ? 2u : 0u) + (isfirst ? 0u : 1u)) * 4u);
            const uint8_t* seqp = h.p + h.o_seq;
            const uint8_t* qualp = h.p + h.o_qual;
            uint32_t it = 0;
            uint32_t cc = (first_op >> 4) - 1u;      // wraps for a zero count exactly like the size_t in the reference
            uint32_t chromPos = (uint32_t)h.pos + 1u;
            uint32_t readPos = 1;
            const uint32_t last = Ls - 1;
            bool ok = true;
            while (readPos < last) {
                if (cc == 0) {
                    do {
                        ++it;
                        if (it >= h.ncig) { ok = false; break; }  // the reference reads past the CIGAR here (undefined)
                        uint32_t c = ldu32<M>(h.p + h.o_cig + 4 * it);
                        uint32_t op = c & 15u, n = c >> 4;
                        if (op == 2 || op == 3 || op == 5 || op == 6) chromPos += n;
                        else if (op == 4 || op == 1) readPos += n;
                        else cc = n;
                    } while (cc == 0);
                    if (!ok || readPos >= last) break;
                }
                const uint32_t run = min(cc, last - readPos);
                // positions [readPos, end) of this run lie inside the contig: chromPos + 2 <= reflen (:217-224)
                const int64_t lim = (int64_t)reflen - 1 - (int64_t)chromPos + (int64_t)readPos;  // first p that fails
                uint32_t end = readPos + run;
                if (lim < (int64_t)end) end = lim > (int64_t)readPos ? (uint32_t)lim : readPos;
                const uint32_t delta = chromPos - readPos;
                for (uint32_t p = readPos; p < end;) {
                    const uint32_t g = (p - 1u) & ~1u;                       // window = read positions g..g+15
                    const uint64_t R = swar_swap_nibbles(ldu64<M>(seqp + (g >> 1)));
                    const uint64_t q0 = ldu64<M>(qualp + g), q1 = ldu64<M>(qualp + g + 8);
                    const int32_t cg = (int32_t)(delta + g);                 // reference position of window base 0 (>= -1)
                    const uint32_t cgc = cg < 0 ? 0u : (uint32_t)cg;
                    const uint32_t wi = cgc >> 4;
                    uint32_t F = __funnelshift_r(__ldg(ref + min(wi, refmax)), __ldg(ref + min(wi + 1u, refmax)), 2u * (cgc & 15u));
                    if (cg < 0) F <<= 2;
                    uint64_t T4, C4;
                    swar_triplet_masks(R, F, T4, C4);
                    const uint32_t jlo = p - g, jhi = min(end - g, 15u);      // centres jlo..jhi-1 of the window
                    C4 &= ((1ULL << (4u * jhi)) - 1ULL) & ~((1ULL << (4u * jlo)) - 1ULL);
                    const uint32_t qk[4] = {swar_q20((uint32_t)q0), swar_q20((uint32_t)(q0 >> 32)), swar_q20((uint32_t)q1), swar_q20((uint32_t)(q1 >> 32))};
                    const uint32_t c_lo = (uint32_t)C4, c_hi = (uint32_t)(C4 >> 32), t_lo = (uint32_t)T4, t_hi = (uint32_t)(T4 >> 32);
#pragma unroll
                    for (uint32_t j = 1; j < 15; ++j) {
                        const uint32_t cw = j < 8 ? c_lo : c_hi, tw = j < 8 ? t_lo : t_hi;
                        const uint32_t cnt = (cw >> (4u * (j & 7u))) & (qk[j >> 2] >> (8u * (j & 3u) + 7u)) & 1u;
                        const uint32_t ctx = (F >> (2u * (j - 1u))) & 63u;
                        const uint32_t base = (tw >> (4u * (j & 7u))) & 3u;
                        red_shared_add(tri_s + (ctx * 16u + base) * 4u, cnt);  // adds 0 where the position does not count: no branch
                    }
                    p = g + jhi;
                }
                readPos += run;
                chromPos += run;
                cc -= run;
            }
        }
        __syncwarp();
        // ---------------------------------------------------------------- phase C: get_count (:111-176)
        const uint32_t myL = live ? Ls : 0u;
        const uint32_t maxL = __reduce_max_sync(0xFFFFFFFFu, myL);
        uint32_t cntN = 0, cntGC = 0, sumQ = 0;
        {
            // Eight CYCLES per step, for both orientations: a forward read takes stored bases 8s..8s+7, a reverse read
            // the stored bases L-8-8s..L-1-8s (an unaligned nibble window, possibly starting before base 0 in the
            // last step) and brings them into cycle order: __brev of the nibble word reverses the base order AND
            // the bits of every nibble, which is the complement for the one-hot codes (R7); the quality bytes are
            // byte-reversed.  Per word: Dna5 ordinals, N / GC counts and the quality sum come from swar.h; per base
            // one increment of dnacount[d][cycle] and one add to qualcount[cycle].  Bases past the end of the read
            // (last step) are sent to a dump row with quality 0, so every lane runs the same instructions.
            // Shared-memory conflicts: the rows are padded by one word per 8 cycles (cycle c lives at c + c/8), so
            // lanes working on different steps hit different banks, and lanes visit their steps in rotated order
            // (lane l starts at step l mod nsteps, reverse reads half a turn further).  Before, lanes of one mate and
            // orientation added to the same qualcount word in the same instruction (ncu: 8 wavefronts per add) and
            // the shared atomic pipe, not instruction issue, limited the kernel.
            const uint32_t rowp = stats_row_words(cycb), rowb = rowp * 4u;
            const uint32_t pc_s = (uint32_t)__cvta_generic_to_shared(sm + S.pc + mate * (PC_ROWS + 1u) * rowp);
            const uint8_t* seqp = live ? h.p + h.o_seq : recp;
            const uint8_t* qualp = live ? h.p + h.o_qual : recp;
            const uint32_t mysteps = (myL + 7u) >> 3;
            const uint32_t rot = mysteps ? ((threadIdx.x & 31u) + (rc ? (mysteps >> 1) : 0u)) % mysteps : 0u;
            for (uint32_t it = 0; it < ((maxL + 7u) >> 3); ++it) {
                if (it < mysteps) {
                    uint32_t step = it + rot;
                    if (step >= mysteps) step -= mysteps;
                    const uint32_t v = min(8u, myL - 8u * step);          // cycles of this step that exist
                    const int32_t n0 = rc ? (int32_t)Ls - 8 - (int32_t)(8u * step) : (int32_t)(8u * step);  // first stored base
                    const uint64_t X = swar_swap_nibbles(ldu64<M>(seqp + (n0 >> 1)));
                    uint32_t W = (uint32_t)(X >> (4u * (uint32_t)(n0 & 1)));
                    const uint64_t Q = ldu64<M>(qualp + n0);
                    uint32_t qlo = (uint32_t)Q, qhi = (uint32_t)(Q >> 32);
                    if (rc) {
                        W = __brev(W);
                        const uint32_t t = __byte_perm(qhi, 0u, 0x0123u);
                        qhi = __byte_perm(qlo, 0u, 0x0123u);
                        qlo = t;
                    }
                    const uint32_t nm = v == 8u ? 0xFFFFFFFFu : (1u << (4u * v)) - 1u;                 // nibbles that exist
                    const uint64_t qm = v == 8u ? ~0ULL : (1ULL << (8u * v)) - 1ULL;
                    qlo &= (uint32_t)qm;
                    qhi &= (uint32_t)(qm >> 32);
                    uint32_t pop4;
                    const uint32_t oh = swar_onehot8(W, pop4);            // bit 4j: base j is A/C/G/T; pop4: bit count per nibble
                    const uint32_t D = (swar_dna5_8(W, oh) & nm) | (~nm & 0x88888888u);  // Dna5 ordinal per nibble (src/QualityCheck.hpp:135); 8 = dump row
                    cntN += __popc(pop4 & 0x44444444u & nm);              // literal N: all four bits (:137)
                    cntGC += __popc(((W >> 1) | (W >> 2)) & oh & nm);     // literal C or G (:141)
                    sumQ = __dp4a(qlo, 0x01010101u, __dp4a(qhi, 0x01010101u, sumQ));
                    const uint32_t ca = pc_s + 36u * step;                // row 0 at cycle 8 * step (9 words per 8 cycles)
#pragma unroll
                    for (uint32_t j = 0; j < 8u; ++j) {
                        const uint32_t d = (D >> (4u * j)) & 15u;
                        const uint32_t q = ((j < 4u ? qlo : qhi) >> (8u * (j & 3u))) & 255u;
                        // ONE shared atomic per base: the count of (base, cycle) in the low 12 bits of the cell, the quality of the
                        // same base in the 20 bits above (qualcount[cycle] is the sum over the five base rows).  The cells are
                        // emptied every kStatsUnpackEvery CTA iterations, before either field can run over (k_stats, unpack).
                        red_shared_add(ca + d * rowb + 4u * j, (q << 12) | 1u);
                    }
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
                        const uint32_t rowp = stats_row_words(cycb);
                        uint32_t* pc = sm + S.pc + mate * (PC_ROWS + 1u) * rowp;
                        if ((fo & 15u) == 4u) {
                            uint32_t n = min(fo >> 4, cycb);
                            for (uint32_t j = 0; j < n; ++j) atomicAdd(pc + PC_SC5 * rowp + j + (j >> 3), 1u);
                        } else if ((lo & 15u) == 4u) {
                            uint32_t n = lo >> 4;
                            for (uint32_t j = (n <= Ls ? Ls - n : Ls); j < Ls; ++j) atomicAdd(pc + PC_SC3 * rowp + j + (j >> 3), 1u);
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
    }
}

// 16-byte asynchronous global -> shared copy (LDGSTS, bypasses L1)
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }

// Bulk asynchronous copy global -> shared (the TMA unit's 1-D mode: cp.async.bulk, SASS UBLKCP) that signals an mbarrier
// with the number of bytes it delivered.  One lane issues the copy of the warp's whole record span; it replaces ~580
// 16-byte cp.async (LDGSTS) per span, 18 per lane plus their address arithmetic.  (STAGE == 2; measured SLOWER than the
// LDGSTS form here, 4.7 vs 4.15 ms per 10 M records: the span cannot be double-buffered within the shared memory that
// two CTAs per SM leave, and one bulk copy per warp has a longer latency than 18 independent 16-byte copies per lane.)
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

// Packed per-cycle cells (count : 12 | quality sum : 20): a CTA looks at 256 records per iteration of its loop, so a cell
// gains at most 256 counts and 256 x 255 of quality per iteration; 15 iterations stay below 4096 and 2^20.
static const uint32_t kStatsUnpackEvery = 4095u / kStatsThreads;
static const uint32_t kStatsStage = 10240;   // bytes of staging per warp: 32 standard 2x150 bp records are 9280 bytes
static const uint32_t kStatsStageTail = 64;  // the decoders read up to a few words past the end of a record
static const uint32_t kStatsMbarBytes = (kStatsThreads / 32) * 8;   // one mbarrier per warp, behind the staging areas

// The 32 records of a warp are one contiguous byte span of the batch.  The warp copies the span into its staging area
// -- STAGE 1: coalesced 16-byte async copies (LDGSTS, ~18 in flight per lane, one round trip to HBM); STAGE 2: one bulk
// copy by the TMA unit (cp.async.bulk + mbarrier) -- and every phase then reads record bytes from shared memory;
// without staging (STAGE 0) the kernel waits on dependent, unaligned global loads (ncu: half of all stall samples).  A
// span larger than the stage is processed in pieces; a single record larger than the stage is read straight from
// global memory.
template <int STAGE>
__global__ void __launch_bounds__(kStatsThreads, kStatsThreads > 256 ? 1 : (STAGE ? 2 : 4)) __maxnreg__(kStatsThreads > 256 ? 96 : (STAGE ? 96 : 64)) k_stats(EngineView E, BatchView B, uint32_t lane) {
    extern __shared__ __align__(16) uint32_t sm[];
    const StatsSmem S = stats_smem_layout(B.cycb, E.insert_smem);
    for (uint32_t i = threadIdx.x; i < S.total; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    const Layout& L = E.L;
    uint64_t* G = E.counters + (uint64_t)lane * L.lane_stride;
    const uint32_t cycb = B.cycb;
    const uint32_t lane_id = threadIdx.x & 31u;
    uint32_t phase = 0;   // parity of this warp's mbarrier
    if (STAGE == 2) {
        if (lane_id == 0) mbar_init((uint32_t)__cvta_generic_to_shared((uint8_t*)(sm + S.stage) + (kStatsThreads / 32) * kStatsStage) + (threadIdx.x >> 5) * 8u, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncthreads();
    }

    // the packed per-cycle cells -> counts to the global rows, quality to the CTA's 32-bit quality row; cells zeroed
    auto unpack = [&]() {
        const uint32_t rowp = stats_row_words(cycb);
        for (uint32_t m = 0; m < 2; ++m) {
            uint64_t* GMm = G + L.o_mate0 + m * L.mate_stride;
            uint32_t* base = sm + S.pc + m * (PC_ROWS + 1u) * rowp;
            for (uint32_t c = threadIdx.x; c < cycb; c += blockDim.x) {
                const uint32_t cc = c + (c >> 3);
                uint32_t qs = 0;
#pragma unroll
                for (uint32_t r = PC_A; r <= PC_N; ++r) {
                    const uint32_t v = base[r * rowp + cc];
                    if (v) {
                        base[r * rowp + cc] = 0;
                        qs += v >> 12;
                        atomicAdd((unsigned long long*)(GMm + L.m_pc + r * pad8(L.cyc) + c), (unsigned long long)(v & 0xFFFu));
                    }
                }
                if (qs) base[PC_QUAL * rowp + cc] += qs;   // this thread owns the column
            }
        }
    };
    const LaneRecords LR = lane_records(B, lane);
    // Epochs of at most kStatsUnpackEvery - 1 warp iterations (32 records each) per warp, then the CTA empties its packed
    // cells.  A warp gets its records from a ticket (B.tickets) or, without one, from the static grid-stride split.
    uint32_t cta0 = blockIdx.x * blockDim.x;
    for (;;) {
    bool more = true;
    for (uint32_t it = 0; it + 1u < kStatsUnpackEvery; ++it) {
        uint32_t r0;
        if (B.tickets) r0 = warp_take32(B.tickets, lane_id);
        else { r0 = cta0 + threadIdx.x - lane_id; cta0 += gridDim.x * blockDim.x; }
        if (r0 >= LR.n) { more = false; break; }
        const uint32_t n_here = min(32u, LR.n - r0);
        uint32_t off = 0, end = 0;
        const uint32_t myrec = lane_id < n_here ? LR[r0 + lane_id] : 0u;   // (staged launches never use an index list: myrec = r0 + lane_id)
        if (lane_id < n_here) { off = B.offsets[myrec]; end = B.offsets[myrec + 1]; }
        if (!STAGE) {
            stats_records<GMem>(E, B, lane, sm, S, myrec, lane_id < n_here, B.bytes + off, end - off);
            continue;
        }
        uint8_t* stage = (uint8_t*)(sm + S.stage) + (threadIdx.x >> 5) * kStatsStage;
        const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(stage);
        const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared((uint8_t*)(sm + S.stage) + (kStatsThreads / 32) * kStatsStage) + (threadIdx.x >> 5) * 8u;
        uint32_t sub = 0;
        while (sub < n_here) {
            const uint32_t lo = __shfl_sync(0xFFFFFFFFu, off, sub) & ~15u;
            const bool fits = lane_id >= sub && lane_id < n_here && end >= lo && end - lo + kStatsStageTail <= kStatsStage;
            const uint32_t m = __popc(__ballot_sync(0xFFFFFFFFu, fits));  // offsets ascend: the fitting lanes are sub..sub+m-1
            if (m == 0) {
                stats_records<GMem>(E, B, lane, sm, S, r0 + lane_id, lane_id == sub, B.bytes + off, end - off);
                sub += 1;
                continue;
            }
            const uint32_t hi = __shfl_sync(0xFFFFFFFFu, end, sub + m - 1);
            const uint32_t nvec = (hi + 48u - lo + 15u) >> 4;
            if (STAGE == 2) {
                if (lane_id == 0) {
                    mbar_expect_tx(bar_s, nvec << 4);
                    bulk_copy_g2s(stage_s, B.bytes + lo, nvec << 4, bar_s);
                }
                mbar_wait(bar_s, phase);
                phase ^= 1u;
            } else {
                for (uint32_t v = lane_id; v < nvec; v += 32u) cp_async16(stage_s + 16u * v, B.bytes + lo + 16u * v);
                cp_async_wait_all();
                __syncwarp();
            }
            stats_records<SMem>(E, B, lane, sm, S, r0 + lane_id, fits, stage + (off - lo), end - off);
            __syncwarp();  // every lane is done with the stage before it is refilled
            sub += m;
        }
    }
    const int any = __syncthreads_or(more ? 1 : 0);
    unpack();
    __syncthreads();
    if (!any) break;
    }
    // ---- flush the CTA-private tables (skip zeros) ---------------------------------------------------
    auto flush = [&](uint32_t smo, uint32_t n, uint64_t* g) {
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            uint32_t v = sm[smo + i];
            if (v) atomicAdd((unsigned long long*)(g + i), (unsigned long long)v);
        }
    };
    for (uint32_t m = 0; m < 2; ++m) {
        uint64_t* GMm = G + L.o_mate0 + m * L.mate_stride;
        const uint32_t rowp = stats_row_words(cycb);
        for (uint32_t r = 0; r < PC_ROWS; ++r)  // padded rows: cycle c lives at c + c/8
            for (uint32_t c = threadIdx.x; c < cycb; c += blockDim.x) {
                uint32_t v = sm[S.pc + (m * (PC_ROWS + 1u) + r) * rowp + c + (c >> 3)];
                if (v) atomicAdd((unsigned long long*)(GMm + L.m_pc + r * pad8(L.cyc) + c), (unsigned long long)v);
            }
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
    for (uint32_t i = threadIdx.x; i < kTriplet; i += blockDim.x) {  // smem context order is (next, cur, prev): swar.h
        uint32_t v = sm[S.tri + i];
        if (v) atomicAdd((unsigned long long*)(G + L.o_triplet + triplet_ctx_to_result(i >> 4) * 16u + (i & 15u)), (unsigned long long)v);
    }
    flush(S.sc, S_COUNT, G + L.o_scalars);
}

}  // namespace bqc
