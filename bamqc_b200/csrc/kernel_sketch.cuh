// kernel_sketch.cuh -- k_sketch32: the default-configuration (k = 32) fast path of the KmerStream sketch
// update: ReadQualityHasher::operator() (src/ReadQualityHasher.hpp:30-68) -> RepHash::update
// (src/kmerstream/RepHash.hpp:101-114) -> StreamCounter::operator() (src/kmerstream/StreamCounter.hpp:67-93).
// Same arithmetic as the generic k_sketch in kernels.cuh; the differences are mechanical:
//   * 16 bases (one 8-byte SEQ chunk) per outer iteration, fully unrolled, next chunks prefetched one
//     iteration ahead so the loads overlap the hash updates of the current chunk;
//   * with k = 32 the base leaving the window sits at the same nibble of the chunk two iterations back,
//     so the lagging SEQ stream needs no loads at all;
//   * the four 128-bit table lookups per base collapse into one 32-byte row indexed by (in, out):
//     row = { H[in] ^ rotl_k(H[out]),  rotl_k(Ht[in]) ^ Ht[out] }.
#pragma once

namespace bqc {

// Only the low 64 bits of both 128-bit states are ever used (hash = h.lo ^ ht.lo) and a base stays in the window
// for k = 32 steps, so each state can be carried in 96 bits without changing a single hash:
//   h  = XOR_j rotl^j(H[s_j]), j = 0..31: its low word takes at most the top 31 bits of H.hi, bits of H.lo that rotate
//        up into hi would need 64 more steps to come back -> keep (lo, top 32 bits of hi); the 32-bit hi part simply
//        shifts left and a base's contribution leaves it by itself after 32 steps;
//   ht = XOR_m rotl^m(Ht[s]), m = 31..0, rotated RIGHT each step: bits come back into lo from the LOW end of hi ->
//        keep (lo, low 32 bits of hi) with bits 0..32 of Ht.hi cleared in the table (they never reach lo in 32 steps).
// One 32-byte row per (strand, in, out): {a_lo, b_lo, a_hi32, b_hi32} with a = H[in] ^ rotl_k(H[out]) and
// b = rotl_k(Ht''[in]) ^ Ht''[out]; the generic k_sketch keeps the full 128-bit arithmetic of the reference.
struct HashPairRow {
    uint64_t a_lo, b_lo;
    uint32_t a_hi, b_hi;
    uint32_t pad[2];
};
struct HashPairTable {  // [strand][in 0..15][out 0..16]
    HashPairRow row[2][16][17];
};

// Valid k-mer hashes are rare on low-quality stretches (one base under the cutoff silences the next k windows), so
// probing the sketch straight from the per-base loop would run the ~70 instructions of a probe for a handful of
// active lanes.  Instead each lane appends its hash to a per-warp queue in shared memory (ballot + rank) and the
// warp probes 32 queued hashes at a time with every lane busy.  The order of the updates does not matter: F2 and
// sumCount are sums, the sketch counters saturate at 15.
static const uint32_t kSketchQueue = 64;  // entries per warp (at most 31 waiting + 32 new)

struct SketchProbe {  // one in-flight probe per lane: the word was loaded, the 4-bit increment is still to be done
    uint32_t word;       // index of the sketch word (kNone: nothing pending)
    uint32_t sh, old;
    __device__ __forceinline__ void finish(uint32_t* sk) {
        if (word != kNone) {
            while (((old >> sh) & 15u) != 15u) {  // 4-bit saturating increment
                const uint32_t assumed = old;
                old = atomicCAS(sk + word, assumed, assumed + (1u << sh));
                if (old == assumed) break;
            }
            word = kNone;
        }
    }
};

__global__ void __launch_bounds__(kSketchThreads, 1) k_sketch32(EngineView E, BatchView B, uint32_t lane, SketchParams SP, const HashTables* __restrict__ HT) {
    extern __shared__ uint32_t sm[];                 // F2 table (f2size u32), the pair table, the per-warp hash queues
    const Layout& L = E.L;
    HashPairTable* PT = reinterpret_cast<HashPairTable*>(sm + L.f2size);
    uint64_t* queue = reinterpret_cast<uint64_t*>(PT + 1) + (threadIdx.x >> 5) * kSketchQueue;
    for (uint32_t i = threadIdx.x; i < L.f2size; i += blockDim.x) sm[i] = 0;
    for (uint32_t i = threadIdx.x; i < 2 * 16 * 17; i += blockDim.x) {
        const uint32_t s = i / (16 * 17), in = (i / 17) % 16, out = i % 17;
        // HT->t[s][which][n][0] = hi, [1] = lo; which: 0 H[in], 1 rotl_k(H[out]), 2 rotl_k(Ht[in]), 3 Ht[out]
        // (rotl_32 swaps and mixes the halves: rotl_32(X).lo = X.lo << 32 | X.hi >> 32, .hi = X.hi << 32 | X.lo >> 32)
        HashPairRow r;
        r.a_lo = HT->t[s][0][in][1] ^ HT->t[s][1][out][1];
        r.a_hi = (uint32_t)(HT->t[s][0][in][0] >> 32);                 // top 32 bits of H[in].hi; the out term has none
        // Ht'' = Ht with hi bits 0..32 cleared: rotl_32(Ht'')[in].lo differs from the stored rotl_32(Ht)[in].lo in bit 0
        r.b_lo = (HT->t[s][2][in][1] & ~1ULL) ^ HT->t[s][3][out][1];
        r.b_hi = (uint32_t)HT->t[s][2][in][0];                         // low 32 bits of rotl_32(Ht[in]).hi = Ht[in].lo >> 32
        r.pad[0] = r.pad[1] = 0;
        PT->row[s][in][out] = r;
    }
    __syncthreads();
    uint64_t* G = E.counters + (uint64_t)lane * L.lane_stride + L.o_qk + (uint64_t)SP.qk * L.qk_stride;
    uint32_t* sk = E.sketch + ((uint64_t)lane * L.n_qk + SP.qk) * 32ull * (L.sk_size * 2ull);
    const uint32_t words_per_level = L.sk_size * 2u;
    const uint64_t idx_mask = (uint64_t)L.sk_size * 16ull - 1ull;
    const uint32_t f2mask = L.f2size - 1u;
    const int32_t q_thresh = SP.q_thresh;
    uint32_t warp_count = 0;                         // k-mers hashed by this warp (same value in every lane; < 2^32 per launch)
    const uint32_t lane_id = threadIdx.x & 31u;
    const uint32_t lt_mask = (1u << lane_id) - 1u;
    uint32_t qn = 0;                                 // queued hashes of this warp (uniform)
    SketchProbe probe = {kNone, 0u, 0u};

    // StreamCounter::operator() (src/kmerstream/StreamCounter.hpp:67-93) for one hash per participating lane
    auto probe_issue = [&](uint64_t hv) {
        probe.finish(sk);
        atomicAdd(sm + ((uint32_t)hv & f2mask), 1u);
        uint32_t w = hv ? (uint32_t)(__ffsll((long long)hv) - 1) : 63u;  // bitScanForward, 63 for 0
        if (w > 31u) w = 31u;
        const uint64_t index = (hv >> (w + 1u)) & idx_mask;
        probe.word = w * words_per_level + (uint32_t)(index >> 3);
        probe.sh = ((uint32_t)index & 7u) * 4u;
        probe.old = __ldcg(sk + probe.word);   // tested one drain later: the L2 latency overlaps the hashing in between
    };

    for (uint32_t r0 = blockIdx.x * blockDim.x + threadIdx.x - lane_id; r0 < B.n_records; r0 += gridDim.x * blockDim.x) {
        const uint32_t rec = r0 + lane_id;
        RecHdr h;
        bool act = rec < B.n_records && !(B.rec_lane && B.rec_lane[rec] != lane);
        if (act) {
            const uint32_t off = B.offsets[rec];
            act = decode_hdr(B.bytes + off, B.offsets[rec + 1] - off, h);
        }
        act = act && !(h.flag & 0xF00u) && (h.flag & 0xC0u) && (uint32_t)h.lseq >= 32u;
        const uint32_t Ls = act ? (uint32_t)h.lseq : 0u;
        const uint32_t maxL = __reduce_max_sync(0xFFFFFFFFu, Ls);
        const uint32_t s = act ? ((h.flag >> 4) & 1u) : 0u;
        const uint8_t* seqp = act ? h.p + h.o_seq : B.bytes;
        const uint8_t* qualp = act ? h.p + h.o_qual : B.bytes;
        const HashPairRow* rows = &PT->row[s][0][0];
        uint64_t hlo = 0, tlo = 0;
        uint32_t hhi = 0, thi = 0;   // top 32 bits of h.hi, low 32 bits of ht.hi
        uint64_t seq_cur = ldu64(seqp), q_lo = ldu64(qualp), q_hi = ldu64(qualp + 8);
        uint64_t hist1 = 0, hist2 = 0;
        uint32_t run = 0;
        const uint32_t nchunks = (maxL + 15u) >> 4;
        for (uint32_t c = 0; c < nchunks; ++c) {
            // prefetch the next chunk of both streams (reads past a short record stay inside the padded batch)
            const uint64_t seq_next = ldu64(seqp + 8 * (c + 1));
            const uint64_t q_lo_next = ldu64(qualp + 16 * (c + 1)), q_hi_next = ldu64(qualp + 16 * (c + 1) + 8);
            const bool has_out = c >= 2;
#pragma unroll
            for (uint32_t j = 0; j < 16; ++j) {
                const uint32_t i = c * 16 + j;
                bool emit = false;
                if (i < Ls) {
                    const uint32_t nin = (uint32_t)(seq_cur >> (4 * (j ^ 1))) & 15u;
                    const uint32_t nout = has_out ? ((uint32_t)(hist2 >> (4 * (j ^ 1))) & 15u) : 16u;
                    const uint32_t q = (uint32_t)((j < 8 ? q_lo : q_hi) >> (8 * (j & 7))) & 255u;
                    const uint32_t ridx = nin * 17u + nout;
                    const ulonglong2 ab = *reinterpret_cast<const ulonglong2*>(rows + ridx);        // a_lo, b_lo
                    const uint2 abh = *reinterpret_cast<const uint2*>(&rows[ridx].a_hi);            // a_hi32, b_hi32
                    // h = rotl1(h) ^ H[in] ^ rotl_k(H[out])
                    hlo = ((hlo << 1) | (uint64_t)(hhi >> 31)) ^ ab.x;
                    hhi = (hhi << 1) ^ abh.x;
                    // ht = rotr1(ht ^ rotl_k(Ht[in]) ^ Ht[out])
                    const uint64_t xl = tlo ^ ab.y;
                    const uint32_t xh = thi ^ abh.y;
                    tlo = (xl >> 1) | ((uint64_t)xh << 63);
                    thi = xh >> 1;
                    const bool valid = (nin != 15u) && ((int8_t)(q + 33u) >= (int8_t)q_thresh);
                    run = valid ? run + 1u : 0u;
                    emit = run >= 32u;
                }
                const uint32_t em = __ballot_sync(0xFFFFFFFFu, emit);
                if (em) {  // warp-uniform
                    if (emit) queue[qn + __popc(em & lt_mask)] = hlo ^ tlo;
                    const uint32_t n = __popc(em);
                    qn += n;
                    warp_count += n;
                    if (qn >= 32u) {
                        __syncwarp();
                        qn -= 32u;
                        probe_issue(queue[qn + lane_id]);
                        __syncwarp();
                    }
                }
            }
            hist2 = hist1;
            hist1 = seq_cur;
            seq_cur = seq_next;
            q_lo = q_lo_next;
            q_hi = q_hi_next;
        }
    }
    __syncwarp();
    if (lane_id < qn) probe_issue(queue[lane_id]);  // what is left in the queue
    probe.finish(sk);
    if (lane_id == 0 && warp_count) atomicAdd((unsigned long long*)G, (unsigned long long)warp_count);
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < L.f2size; i += blockDim.x) {
        uint32_t v = sm[i];
        if (v) atomicAdd((unsigned long long*)(G + 8 + i), (unsigned long long)v);
    }
}

}  // namespace bqc
