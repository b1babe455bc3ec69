// kernels.cuh -- hand-written CUDA (sm_100a) for the per-record statistics pass of bamqualcheck.
//
// Execution model: one THREAD per BAM record, records dealt to persistent CTAs in
// an interleaved grid-stride order so neighbouring threads read neighbouring records.  Each statistic
// family gets its own kernel so that its hot table can be privatised in shared memory:
//   k_stats  : gate + scalars, QualityCheck per-cycle tables and histograms, TripletCounting walk
//              (~21 KB of smem tables + 8 x 10 KB of per-warp record staging per CTA)
//   k_eightmer: OverallNumbers::count8mers                 (65536 x 16-bit counters = 128 KB smem per CTA)
//   k_sketch : ReadQualityHasher + RepHash + StreamCounter (F2 table = 128 KB smem per CTA; the 8 MiB
//              4-bit sketch stays in L2 and is updated with load-test-CAS)
//   k_cov_*  : coverage windows entirely on the device (kernel_cov.cuh): anchor recurrence as a scan, tiled
//              shared-memory depth histogram
// Measured on B200 (profiles/ubench): shared-memory atomics sustain ~1700 G/s chip-wide, global REDs to
// an L2-resident table ~150 G/s and collapse to <14 G/s on hot bins, hence the privatisation.
//
// Per-base tests are evaluated 8 or 16 bases at a time on packed words (swar.h); only the table update is per base.
// All arithmetic is integer; every table update is a commutative add (or a saturating 4-bit add for
// the sketch), so results do not depend on scheduling and are bit-exact against the oracle.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "layout.h"
#include "swar.h"

namespace bqc {

// ------------------------------------------------------------------------------------------------
// views passed to the kernels
// ------------------------------------------------------------------------------------------------
struct BatchView {
    const uint8_t* bytes;      // inflated BAM records, padded with >= 64 readable bytes
    const uint32_t* offsets;   // n_records + 1 byte offsets
    const uint8_t* rec_lane;   // per record lane (multi-lane runs) or NULL
    const uint32_t* index;     // multi-lane runs: the record indices grouped by lane (file order inside a lane), or NULL:
    const uint32_t* index_range;  // ... lane l owns index[index_range[l] .. index_range[l + 1]), so that a lane's pass only
                               // touches its own records instead of filtering the whole batch
    uint32_t n_records;
    uint32_t cycb;             // per-cycle smem capacity for this batch (>= max l_seq, multiple of 8)
    uint64_t first_record;     // global index of record 0 (error reporting)
    uint32_t* tickets = nullptr;  // zeroed before every lane pass: [0] k_stats, [1] k_eightmer, [2 + qk] sketch kernels -- the
                               // persistent table kernels take warps of 32 records from these counters instead of a
                               // static round-robin split (a CTA that becomes resident late, e.g. behind a coverage CTA,
                               // then simply takes fewer records); NULL = static split
};

// next 32 records of a warp: a ticket (all lanes get the same value), or the static grid-stride position
__device__ __forceinline__ uint32_t warp_take32(uint32_t* ticket, uint32_t lane_id) {
    uint32_t r0 = 0;
    if (lane_id == 0) r0 = atomicAdd(ticket, 32u);
    return __shfl_sync(0xFFFFFFFFu, r0, 0);
}

struct EngineView {
    Layout L;
    uint64_t* counters;              // [n_lanes][lane_stride]
    uint32_t* sketch;                // [n_lanes][n_qk][32][sk_size*2] 4-bit counters, 8 per word
    const uint32_t* const* ref;      // [n_ref] 2-bit packed contigs (16 bases per word) or NULL
    const uint64_t* ref_len;         // [n_ref]
    const uint8_t* main_chrom;       // [n_ref]
    int32_t n_ref;
    unsigned long long* error;       // min over (record << 8 | code), ~0 if none
    uint32_t insert_smem;            // insert-size bins kept in shared memory
};

struct SketchParams {
    uint32_t k;            // k-mer size (1..63)
    int32_t q_thresh;      // (char)(q_base + q_cutoff) as signed char
    uint32_t qk;           // pair index
};

// hash tables for one k (built on host, see engine.cu build_hash_tables): [strand][which][17] x 128 bit
struct HashTables {
    uint64_t t[2][4][17][2];  // which: 0 = h 'in', 1 = h 'out', 2 = ht 'in', 3 = ht 'out'; [..][0]=hi [1]=lo
};

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
// Where record bytes are read from: global memory through the read-only path, or a warp's staging area in shared
// memory (k_stats copies the contiguous byte span of its 32 records there with 16-byte async copies).
struct GMem {
    static __device__ __forceinline__ uint32_t ld8(const uint8_t* p) { return __ldg(p); }
    static __device__ __forceinline__ uint32_t ld32(const uint32_t* p) { return __ldg(p); }
    static __device__ __forceinline__ uint64_t ld64(const uint64_t* p) { return __ldg(p); }
};
struct SMem {
    // Explicit ld.shared: with plain dereferences the compiler folded the "two aligned loads + funnel shift" of
    // ldu32/ldu64 back into ONE load at the unaligned address, which faults in shared memory (misaligned address).
    static __device__ __forceinline__ uint32_t ld8(const uint8_t* p) {
        uint32_t v;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)));
        return v;
    }
    static __device__ __forceinline__ uint32_t ld32(const uint32_t* p) {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)));
        return v;
    }
    static __device__ __forceinline__ uint64_t ld64(const uint64_t* p) {
        uint64_t v;
        asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)));
        return v;
    }
};
template <class M = GMem>
__device__ __forceinline__ uint32_t ldg8(const uint8_t* p) { return M::ld8(p); }
template <class M = GMem>
__device__ __forceinline__ uint32_t ldu32(const uint8_t* p) {  // unaligned little-endian 32-bit load
    uintptr_t a = (uintptr_t)p;
    const uint32_t* w = (const uint32_t*)(a & ~(uintptr_t)3);
    uint32_t sh = (uint32_t)(a & 3) * 8;
    uint32_t lo = M::ld32(w);
    if (sh == 0) return lo;
    return __funnelshift_r(lo, M::ld32(w + 1), sh);
}
template <class M = GMem>
__device__ __forceinline__ uint64_t ldu64(const uint8_t* p) {  // unaligned little-endian 64-bit load
    uintptr_t a = (uintptr_t)p;
    const uint64_t* w = (const uint64_t*)(a & ~(uintptr_t)7);
    uint32_t sh = (uint32_t)(a & 7) * 8;
    uint64_t lo = M::ld64(w);
    if (sh == 0) return lo;
    return (lo >> sh) | (M::ld64(w + 1) << (64 - sh));
}
// nibble i (0..15) of a 64-bit SEQ chunk: byte i/2, high nibble first
__device__ __forceinline__ uint32_t nib_of(uint64_t w, uint32_t i) { return (uint32_t)(w >> (4 * (i ^ 1))) & 15u; }

// the records a kernel launched for `lane` iterates over: all of the batch (filtered by rec_lane, if any), or the lane's
// own index list
struct LaneRecords {
    uint32_t n;
    const uint32_t* idx;
    __device__ __forceinline__ uint32_t operator[](uint32_t i) const { return idx ? idx[i] : i; }
};
__device__ __forceinline__ LaneRecords lane_records(const BatchView& B, uint32_t lane) {
    LaneRecords R;
    if (B.index) { const uint32_t a = B.index_range[lane]; R.n = B.index_range[lane + 1] - a; R.idx = B.index + a; }
    else { R.n = B.n_records; R.idx = nullptr; }
    return R;
}
__device__ __forceinline__ bool lane_match(const BatchView& B, uint32_t rec, uint32_t lane) { return B.index || !B.rec_lane || B.rec_lane[rec] == lane; }

__device__ __forceinline__ void report_error(const EngineView& E, uint64_t rec, uint32_t code) {
    atomicMin(E.error, (unsigned long long)((rec << 8) | code));
}

struct RecHdr {
    const uint8_t* p;
    uint32_t bs;
    int32_t rid, pos;
    uint32_t lname, mapq, ncig, flag;
    int32_t lseq, nrid, tlen;
    uint32_t o_cig, o_seq, o_qual, o_aux, o_end;
};

// BAM record fixed fields (SAM/BAM spec).  Returns false if the record is malformed.
template <class M = GMem>
__device__ __forceinline__ bool decode_hdr(const uint8_t* p, uint32_t avail, RecHdr& h) {
    uintptr_t a = (uintptr_t)p;
    const uint32_t* w = (const uint32_t*)(a & ~(uintptr_t)3);
    uint32_t sh = (uint32_t)(a & 3) * 8;
    uint32_t v[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) v[i] = M::ld32(w + i);
#define BQC_F(i) (sh ? __funnelshift_r(v[i], v[i + 1], sh) : v[i])
    h.p = p;
    h.bs = BQC_F(0);
    h.rid = (int32_t)BQC_F(1);
    h.pos = (int32_t)BQC_F(2);
    uint32_t x = BQC_F(3);
    h.lname = x & 255u;
    h.mapq = (x >> 8) & 255u;
    uint32_t y = BQC_F(4);
    h.ncig = y & 0xFFFFu;
    h.flag = y >> 16;
    h.lseq = (int32_t)BQC_F(5);
    h.nrid = (int32_t)BQC_F(6);
    h.tlen = (int32_t)BQC_F(8);
#undef BQC_F
    if (h.bs + 4u != avail || h.bs < 32u || h.lseq < 0) return false;
    h.o_cig = 36u + h.lname;
    h.o_seq = h.o_cig + 4u * h.ncig;
    h.o_qual = h.o_seq + (((uint32_t)h.lseq + 1u) >> 1);
    h.o_aux = h.o_qual + (uint32_t)h.lseq;
    h.o_end = avail;
    return h.o_aux <= h.o_end && h.o_aux >= h.o_qual;
}

// BAM 4-bit code -> Dna5 ordinal (A0 C1 G2 T3 else 4) and its complement (SURVEY Appendix F LUTs)
__device__ __forceinline__ uint32_t dna5_of(uint32_t nib) {
    // nibbles 1,2,4,8 -> 0..3; packed 3-bit LUT would not fit 16 entries in 32 bits, so: popc test + ffs
    return (__popc(nib) == 1) ? (uint32_t)(__ffs((int)nib) - 1) : 4u;
}

struct AuxInfo {
    uint32_t rg;        // 0 = no RG tag, 1 = RG:Z, 2 = RG of another type
    uint32_t as_state;  // 0 = no AS tag, 1 = readable, 2 = unreadable type
    int32_t as_value;
};

template <class M = GMem>
__device__ __forceinline__ uint32_t aux_value_size(uint32_t type, const uint8_t* p, uint32_t pos, uint32_t end) {
    // size of the value that starts at pos for a tag of `type`; 0xFFFFFFFF if malformed
    switch (type) {
        case 'A': case 'c': case 'C': return 1;
        case 's': case 'S': return 2;
        case 'i': case 'I': case 'f': return 4;
        case 'Z': case 'H': {
            uint32_t q = pos;
            while (q < end && ldg8<M>(p + q) != 0) ++q;
            return q - pos + 1;
        }
        case 'B': {
            if (pos + 5 > end) return 0xFFFFFFFFu;
            uint32_t sub = ldg8<M>(p + pos);
            uint32_t cnt = ldu32<M>(p + pos + 1);
            uint32_t es = (sub == 'c' || sub == 'C') ? 1u : (sub == 's' || sub == 'S') ? 2u : 4u;
            return 5u + cnt * es;
        }
        default: return 0xFFFFFFFFu;
    }
}
// SeqAn extractTagValue into an integer (R15)
template <class M = GMem>
__device__ __forceinline__ bool aux_int(uint32_t type, const uint8_t* p, long long& out) {
    switch (type) {
        case 'c': out = (int8_t)ldg8<M>(p); return true;
        case 'C': case 'A': out = (long long)ldg8<M>(p); return true;
        case 's': out = (int16_t)(ldg8<M>(p) | (ldg8<M>(p + 1) << 8)); return true;
        case 'S': out = (long long)(ldg8<M>(p) | (ldg8<M>(p + 1) << 8)); return true;
        case 'i': out = (int32_t)ldu32<M>(p); return true;
        case 'I': out = (long long)ldu32<M>(p); return true;
        case 'f': out = (long long)__uint_as_float(ldu32<M>(p)); return true;
        default: return false;
    }
}

// One walk over the aux block: RG type (src/bamqualcheck.cpp:77-99), first AS (src/TripletCounting.hpp:
// 113-129) and every integer-typed NM (src/QualityCheck.hpp:201-218, via on_nm).
template <class M = GMem, typename OnNM>
__device__ __forceinline__ AuxInfo aux_walk(const RecHdr& h, OnNM on_nm) {
    AuxInfo ai = {0u, 0u, 0};
    uint32_t pos = h.o_aux;
    const uint8_t* p = h.p;
    while (pos + 3 <= h.o_end) {
        uint32_t k0 = ldg8<M>(p + pos), k1 = ldg8<M>(p + pos + 1), ty = ldg8<M>(p + pos + 2);
        pos += 3;
        uint32_t sz = aux_value_size<M>(ty, p, pos, h.o_end);
        if (sz == 0xFFFFFFFFu) break;
        if (k0 == 'R' && k1 == 'G') {
            if (ai.rg == 0) ai.rg = (ty == 'Z') ? 1u : 2u;
        } else if (k0 == 'A' && k1 == 'S') {
            if (ai.as_state == 0) {
                long long v;
                if (aux_int<M>(ty, p + pos, v)) { ai.as_state = 1; ai.as_value = (int32_t)v; }
                else ai.as_state = 2;
            }
        } else if (k0 == 'N' && k1 == 'M') {
            if (ty == 'c' || ty == 'C' || ty == 'i' || ty == 'I' || ty == 's' || ty == 'S') {
                long long v = 0;
                aux_int<M>(ty, p + pos, v);
                on_nm((uint32_t)v);
            }
        }
        pos += sz;
    }
    return ai;
}

// ------------------------------------------------------------------------------------------------
// k_stats
// ------------------------------------------------------------------------------------------------
static const uint32_t kQS = 64;    // smem bins of the average-quality histograms (rest -> global)
static const uint32_t kHS = 32;    // smem bins of mismatch / del / ins histograms
#ifndef BQC_STATS_THREADS
#define BQC_STATS_THREADS 256
#endif
static const uint32_t kStatsThreads = BQC_STATS_THREADS;

struct StatsSmem {  // word offsets into the dynamic shared array
    uint32_t pc, rl, nc, gc, aq, cq, mq, mm, dl, in, isz, tri, sc, total, stage;
};
// words per per-cycle row in k_stats' shared memory: one pad word per 8 cycles (cycle c lives at c + c/8)
__host__ __device__ inline uint32_t stats_row_words(uint32_t cycb) { return cycb + (cycb >> 3) + 1u; }
__host__ __device__ inline StatsSmem stats_smem_layout(uint32_t cycb, uint32_t insert_smem) {
    StatsSmem s;
    uint32_t o = 0;
    s.pc = o;  o += 2 * (PC_ROWS + 1) * stats_row_words(cycb);  // padded per-cycle rows + one dump row per mate
    s.rl = o;  o += 2 * (cycb + 8);
    s.nc = o;  o += 2 * (cycb + 8);
    s.gc = o;  o += 2 * (cycb + 8);
    s.aq = o;  o += 2 * kQS;
    s.cq = o;  o += 2 * kQS;
    s.mq = o;  o += 2 * kMapqCap;
    s.mm = o;  o += 2 * kHS;
    s.dl = o;  o += 2 * kHS;
    s.in = o;  o += 2 * kHS;
    s.isz = o; o += insert_smem;
    s.tri = o; o += kTriplet;
    s.sc = o;  o += S_COUNT;
    s.total = o;
    s.stage = (o + 3u) & ~3u;  // 16-byte aligned start of the per-warp staging areas (k_stats)
    return s;
}

__device__ __forceinline__ void bump(uint32_t* sm, uint32_t smcap, uint64_t* g, uint32_t gcap, uint32_t idx, const EngineView& E, uint64_t rec) {
    if (idx < smcap) atomicAdd(sm + idx, 1u);
    else if (idx < gcap) atomicAdd((unsigned long long*)(g + idx), 1ULL);
    else report_error(E, rec, 16 /*BQC_ERR_UNSUPPORTED*/);
}

}  // namespace bqc
#include "kernel_stats.cuh"
namespace bqc {

// ------------------------------------------------------------------------------------------------
// k_eightmer: OverallNumbers::count8mers (src/OverallNumbers.hpp:137-168)
// The whole 65536-bin table of a CTA lives in shared memory as 16-bit counters, two per 32-bit word (128 KB), so
// every record is read once.  Overflow is exact: an increment is a 32-bit atomicAdd whose return value shows a
// field that was 0xFFFF; the wrap is credited to the 64-bit global bin at once (+65536) and, for the low field,
// the carry that leaked into the neighbouring high field is taken back from ITS global bin (-1 mod 2^64).  The
// final flush adds the 16-bit residues.  All of it is commutative modular arithmetic, so the sums are exact.
// ------------------------------------------------------------------------------------------------
static const uint32_t kEightThreads = 1024;
__global__ void __launch_bounds__(kEightThreads, 1) k_eightmer(EngineView E, BatchView B, uint32_t lane) {
    extern __shared__ uint32_t sm[];  // 32768 words = 65536 x u16
    for (uint32_t i = threadIdx.x; i < 32768u; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    unsigned long long* g = (unsigned long long*)(E.counters + (uint64_t)lane * E.L.lane_stride + E.L.o_eightmer);
    // Sixteen bases per step (swar.h): 2-bit codes of the whole chunk at once -- forward C(2)->1 G(4)->2 T(8)->3 else 0;
    // reverse reads see the complemented base A(1)->3 C(2)->2 G(4)->1 else 0 (char->Dna conversion, R5/R7) -- packed
    // into a 32-bit word; the 8-mer that ends at base e is a 16-bit field of (previous word : this word), taken with one
    // funnel shift.  Forward reads pack big-endian (first base most significant, :144-156), reverse reads roll the
    // reverse-complement code from the other end, which is the little-endian packing of the complemented codes.
    // A window counts iff it holds no literal N (:146-166): a nibble-stride mask of "N among the last eight bases",
    // with the distance to the last N carried between chunks (starting at 0, which also rules out the first 7 ends).
    const LaneRecords LR = lane_records(B, lane);
    const uint32_t lane_id = threadIdx.x & 31u;
    uint32_t w0 = blockIdx.x * blockDim.x + threadIdx.x - lane_id;   // first record of this warp's 32
    for (;;) {
        if (B.tickets) w0 = warp_take32(B.tickets + 1, lane_id);
        if (w0 >= LR.n) break;
        const uint32_t ri = w0 + lane_id;
        if (!B.tickets) w0 += gridDim.x * blockDim.x;
        if (ri >= LR.n) continue;
        const uint32_t rec = LR[ri];
        if (!lane_match(B, rec, lane)) continue;
        const uint32_t off = B.offsets[rec];
        RecHdr h;
        if (!decode_hdr(B.bytes + off, B.offsets[rec + 1] - off, h)) continue;
        if ((h.flag & 0x900u) || !(h.flag & 0xC0u)) continue;
        const uint32_t Ls = (uint32_t)h.lseq;
        if (Ls < 8u) continue;
        const bool rc = (h.flag & 0x10u) != 0;
        const uint8_t* seqp = h.p + h.o_seq;
        uint32_t prev = 0;        // packed codes of the previous chunk
        uint32_t since_n = 0;     // bases since the last N (or the start of the read) at the start of the chunk, saturating
        const uint32_t nch = (Ls + 15u) >> 4;
        // SEQ as a stream of aligned 8-byte words, each loaded once, one chunk ahead of its use, realigned in registers
        // (an unaligned ldu64 per chunk is two loads that both wait right before their first use)
        const uint64_t* sA = reinterpret_cast<const uint64_t*>((uintptr_t)seqp & ~(uintptr_t)7);
        const uint32_t ssh = ((uint32_t)(uintptr_t)seqp & 7u) * 8u;
        uint64_t A0 = __ldg(sA), A1 = __ldg(sA + 1);
        for (uint32_t c = 0; c < nch; ++c) {
            const uint64_t A2 = __ldg(sA + c + 2u);                       // (past a short record: still inside the padded batch)
            const uint64_t R = swar_swap_nibbles((A0 >> ssh) | ((A1 << 1) << (63u - ssh)));   // base j of the chunk in nibble j
            A0 = A1;
            A1 = A2;
            const uint32_t rem = Ls - 16u * c;                             // bases of the read in this chunk (>= 1)
            const uint64_t inr = rem >= 16u ? ~0ULL : (1ULL << (4u * rem)) - 1ULL;
            const uint64_t s2 = (R & 0x5555555555555555ULL) + ((R >> 1) & 0x5555555555555555ULL);
            const uint64_t t = (s2 & 0x3333333333333333ULL) + ((s2 >> 2) & 0x3333333333333333ULL);  // set bits per nibble
            const uint64_t u = t ^ kNib1;
            const uint64_t oh = ~(u | (u >> 1) | (u >> 2)) & kNib1;        // A/C/G/T
            const uint64_t isn = (t >> 2) & kNib1 & inr;                   // literal N inside the read
            uint64_t code = swar_code4(R);
            if (rc) code ^= 0x3333333333333333ULL;                         // complement: 3 - c
            code &= oh * 3ULL;                                             // everything else counts as 0 (R5)
            // 16 codes -> 32 bits, base j at bits 2j
            uint64_t x = (code | (code >> 2)) & 0x0F0F0F0F0F0F0F0FULL;
            x = (x | (x >> 4)) & 0x00FF00FF00FF00FFULL;
            x = (x | (x >> 8)) & 0x0000FFFF0000FFFFULL;
            const uint32_t ple = (uint32_t)x | ((uint32_t)(x >> 32) << 16);
            uint32_t lo, hi, cur;
            if (rc) {   // field of base e: bits 18 + 2e .. of (cur : prev)  ->  shift (cur : prev) >> 18 by 2e
                cur = ple;
                lo = (prev >> 18) | (cur << 14);
                hi = cur >> 18;
            } else {    // big-endian: base j at bits 30 - 2j; field of base e: bits 30 - 2e .. of (prev : cur)
                const uint32_t br = __brev(ple);
                cur = ((br & 0x55555555u) << 1) | ((br >> 1) & 0x55555555u);
                lo = cur;
                hi = prev;
            }
            // ends whose window holds an N or starts before the read / ends after it
            uint64_t bad = isn;
            bad |= bad << 4;
            bad |= bad << 8;
            bad |= bad << 16;
            if (since_n < 7u) bad |= (1ULL << (4u * (7u - since_n))) - 1ULL;
            bad |= ~inr;
            since_n = isn ? (uint32_t)__clzll((long long)isn) >> 2 : min(since_n + 16u, 64u);
            const uint32_t bad_lo = (uint32_t)bad, bad_hi = (uint32_t)(bad >> 32);
#pragma unroll
            for (uint32_t e = 0; e < 16u; ++e) {
                const uint32_t code16 = __funnelshift_r(lo, hi, rc ? 2u * e : 30u - 2u * e) & 0xFFFFu;
                if (!(((e < 8u ? bad_lo : bad_hi) >> (4u * (e & 7u))) & 1u)) {
                    const uint32_t sh = (code16 & 1u) << 4;
                    const uint32_t old = atomicAdd(sm + (code16 >> 1), 1u << sh);
                    if (((old >> sh) & 0xFFFFu) == 0xFFFFu) {  // this increment wrapped the 16-bit field
                        atomicAdd(g + code16, 65536ULL);
                        if (sh == 0u) {                       // the carry went into the high field (bin code + 1)
                            atomicAdd(g + code16 + 1u, ~0ULL);
                            if (old == 0xFFFFFFFFu) atomicAdd(g + code16 + 1u, 65536ULL);  // and wrapped that one too
                        }
                    }
                }
            }
            prev = cur;
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < 32768u; i += blockDim.x) {
        const uint32_t v = sm[i];
        if (v & 0xFFFFu) atomicAdd(g + 2u * i, (unsigned long long)(v & 0xFFFFu));
        if (v >> 16) atomicAdd(g + 2u * i + 1u, (unsigned long long)(v >> 16));
    }
}

// ------------------------------------------------------------------------------------------------
// k_sketch: ReadQualityHasher::operator() (src/ReadQualityHasher.hpp:30-68) -> RepHash
// (src/kmerstream/RepHash.hpp:85-114) -> StreamCounter::operator() (src/kmerstream/StreamCounter.hpp:67-93)
// ------------------------------------------------------------------------------------------------
static const uint32_t kSketchThreads = 1024;
template <bool F2_IN_SMEM>
__global__ void __launch_bounds__(kSketchThreads, 1) k_sketch(EngineView E, BatchView B, uint32_t lane, SketchParams SP, const HashTables* __restrict__ HT) {
    extern __shared__ uint32_t sm[];
    __shared__ ulonglong2 tab[2][4][17];
    const Layout& L = E.L;
    const uint32_t f2n = F2_IN_SMEM ? L.f2size : 0u;
    for (uint32_t i = threadIdx.x; i < f2n; i += blockDim.x) sm[i] = 0;
    for (uint32_t i = threadIdx.x; i < 2 * 4 * 17; i += blockDim.x) {
        const uint64_t* src = &HT->t[0][0][0][0] + 2 * i;
        (&tab[0][0][0])[i] = make_ulonglong2(src[1], src[0]);  // .x = lo, .y = hi
    }
    __syncthreads();
    uint64_t* G = E.counters + (uint64_t)lane * L.lane_stride + L.o_qk + (uint64_t)SP.qk * L.qk_stride;
    uint32_t* sk = E.sketch + ((uint64_t)lane * L.n_qk + SP.qk) * 32ull * (L.sk_size * 2ull);
    const uint32_t words_per_level = L.sk_size * 2u;
    const uint64_t idx_mask = (uint64_t)L.sk_size * 16ull - 1ull;
    const uint32_t f2mask = L.f2size - 1u;
    const uint32_t k = SP.k;
    unsigned long long my_count = 0;

    // Control flow is kept warp-uniform (one record per lane, common trip count, __syncwarp per base):
    // with independent thread scheduling the lanes otherwise drift apart after the first CAS loop and the
    // per-base loop runs one lane at a time (measured: 2 active threads per instruction).
    const uint32_t lane_id = threadIdx.x & 31u;
    const LaneRecords LR = lane_records(B, lane);
    for (uint32_t r0 = blockIdx.x * blockDim.x + threadIdx.x - lane_id; r0 < LR.n; r0 += gridDim.x * blockDim.x) {
        const uint32_t ri = r0 + lane_id;
        const uint32_t rec = ri < LR.n ? LR[ri] : 0u;
        RecHdr h;
        bool act = ri < LR.n && lane_match(B, rec, lane);
        if (act) {
            const uint32_t off = B.offsets[rec];
            act = decode_hdr(B.bytes + off, B.offsets[rec + 1] - off, h);
        }
        // primary, not QC-fail, not duplicate (src/bamqualcheck.cpp:439-442); l >= k (ReadQualityHasher.hpp:35)
        act = act && !(h.flag & 0xF00u) && (h.flag & 0xC0u) && (uint32_t)h.lseq >= k;
        const uint32_t Ls = act ? (uint32_t)h.lseq : 0u;
        const uint32_t maxL = __reduce_max_sync(0xFFFFFFFFu, Ls);
        const uint32_t s = act ? ((h.flag >> 4) & 1u) : 0u;
        const uint8_t* seqp = act ? h.p + h.o_seq : B.bytes;
        const uint8_t* qualp = act ? h.p + h.o_qual : B.bytes;
        uint64_t hlo = 0, hhi = 0, tlo = 0, thi = 0;
        uint64_t seqw = 0, qualw = 0, lagw = 0;
        uint32_t run = 0;
        bool pend = false;
        uint32_t* pwp = nullptr;
        uint32_t psh = 0, pold = 0;
        for (uint32_t i = 0; i < maxL; ++i, __syncwarp()) {
            if (i >= Ls) continue;
            if ((i & 15u) == 0) seqw = ldu64(seqp + (i >> 1));
            if ((i & 7u) == 0) qualw = ldu64(qualp + i);
            uint32_t nin = nib_of(seqw, i & 15u);
            uint32_t nout = 16u;
            if (i >= k) {
                uint32_t j = i - k;
                if ((j & 15u) == 0 || i == k) lagw = ldu64(seqp + ((j >> 4) << 3));
                nout = nib_of(lagw, j & 15u);
            }
            uint32_t q = (uint32_t)(qualw >> (8 * (i & 7u))) & 255u;
            // h = rotl1(h) ^ rotl_k(H[out]) ^ H[in]
            ulonglong2 a = tab[s][0][nin], b = tab[s][1][nout];
            uint64_t nh = (hhi << 1) | (hlo >> 63);
            hlo = ((hlo << 1) | (hhi >> 63)) ^ a.x ^ b.x;
            hhi = nh ^ a.y ^ b.y;
            // ht = rotr1(ht ^ H[twin[out]] ^ rotl_k(H[twin[in]]))
            ulonglong2 c = tab[s][2][nin], d = tab[s][3][nout];
            uint64_t xl = tlo ^ c.x ^ d.x, xh = thi ^ c.y ^ d.y;
            tlo = (xl >> 1) | (xh << 63);
            thi = (xh >> 1) | (xl << 63);
            bool valid = (nin != 15u) && ((int8_t)(q + 33u) >= (int8_t)SP.q_thresh);
            run = valid ? run + 1u : 0u;
            // The sketch probe is software-pipelined: the word is loaded in the iteration that produces the
            // hash and tested (CAS only while the nibble is < 15) one iteration later, so the L2 latency
            // overlaps the next base's hash update.
            if (pend) {
                while (((pold >> psh) & 15u) != 15u) {  // 4-bit saturating increment
                    uint32_t assumed = pold;
                    pold = atomicCAS(pwp, assumed, assumed + (1u << psh));
                    if (pold == assumed) break;
                }
                pend = false;
            }
            if (run >= k) {
                uint64_t hv = hlo ^ tlo;
                ++my_count;
                if (F2_IN_SMEM) atomicAdd(sm + ((uint32_t)hv & f2mask), 1u);
                else atomicAdd((unsigned long long*)(G + 8 + ((uint32_t)hv & f2mask)), 1ULL);
                uint32_t w = hv ? (uint32_t)(__ffsll((long long)hv) - 1) : 63u;  // bitScanForward, 63 for 0
                if (w > 31u) w = 31u;
                uint64_t index = (hv >> (w + 1u)) & idx_mask;
                pwp = sk + w * words_per_level + (uint32_t)(index >> 3);
                psh = ((uint32_t)index & 7u) * 4u;
                pold = __ldcg(pwp);
                pend = true;
            }
        }
        if (pend) {
            while (((pold >> psh) & 15u) != 15u) {
                uint32_t assumed = pold;
                pold = atomicCAS(pwp, assumed, assumed + (1u << psh));
                if (pold == assumed) break;
            }
            pend = false;
        }
        __syncwarp();
    }
    // sumCount: warp reduce then one atomic per warp
    for (int o = 16; o > 0; o >>= 1) my_count += __shfl_xor_sync(0xFFFFFFFFu, my_count, o);
    if ((threadIdx.x & 31u) == 0 && my_count) atomicAdd((unsigned long long*)G, my_count);
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < f2n; i += blockDim.x) {
        uint32_t v = sm[i];
        if (v) atomicAdd((unsigned long long*)(G + 8 + i), (unsigned long long)v);
    }
}

// ------------------------------------------------------------------------------------------------
// sketch export / import / merge (StreamCounter::join, src/kmerstream/StreamCounter.hpp:95-112)
// ------------------------------------------------------------------------------------------------
__global__ void k_sketch_export_u8(const uint32_t* sk, uint64_t nwords, uint8_t* out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nwords; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t w = sk[i];
        uint2 o;
        o.x = (w & 15u) | (((w >> 4) & 15u) << 8) | (((w >> 8) & 15u) << 16) | (((w >> 12) & 15u) << 24);
        o.y = ((w >> 16) & 15u) | (((w >> 20) & 15u) << 8) | (((w >> 24) & 15u) << 16) | (((w >> 28) & 15u) << 24);
        reinterpret_cast<uint2*>(out)[i] = o;
    }
}
__global__ void k_sketch_import_u8(uint32_t* sk, uint64_t nwords, const uint8_t* in) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nwords; i += (uint64_t)gridDim.x * blockDim.x) {
        uint2 v = reinterpret_cast<const uint2*>(in)[i];
        uint32_t w = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            w |= min((v.x >> (8 * j)) & 255u, 15u) << (4 * j);
            w |= min((v.y >> (8 * j)) & 255u, 15u) << (16 + 4 * j);
        }
        sk[i] = w;
    }
}
__global__ void k_sketch_merge(uint32_t* dst, const uint32_t* src, uint64_t nwords) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nwords; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t a = dst[i], b = src[i], w = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) w |= min(((a >> (4 * j)) & 15u) + ((b >> (4 * j)) & 15u), 15u) << (4 * j);
        dst[i] = w;
    }
}
// move the tables of a result block to a layout with a larger per-cycle capacity: segs = {src offset, dst offset, length}
__global__ void k_relayout(const unsigned long long* src, unsigned long long* dst, const uint3* __restrict__ segs, uint32_t nseg, uint64_t stride_src, uint64_t stride_dst, uint32_t n_lanes) {
    for (uint32_t lane = 0; lane < n_lanes; ++lane)
        for (uint32_t sgi = blockIdx.y; sgi < nseg; sgi += gridDim.y) {
            const uint3 sg = segs[sgi];
            for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < sg.z; i += gridDim.x * blockDim.x)
                dst[lane * stride_dst + sg.y + i] = src[lane * stride_src + sg.x + i];
        }
}
__global__ void k_counters_add(unsigned long long* dst, const unsigned long long* src, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) dst[i] += src[i];
}

}  // namespace bqc
#include "kernel_sketch.cuh"
#include "kernel_cov.cuh"
