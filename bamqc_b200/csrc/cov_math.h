// cov_math.h -- the arithmetic of the coverage window anchor (OverallNumbers::coverage, src/OverallNumbers.hpp:79-110)
// in a form that can be evaluated in parallel.  Host/device: the kernels in kernel_cov.cuh use these functions and
// tests/cov_selftest.cpp checks them (and the block decomposition built on them) against a sequential restatement of
// the reference loop.
//
// Reference state: {first, id, shift}.  Per qualifying record with begin b (unsigned, arithmetic mod 2^32 like the
// reference's `beginpos - shift`):
//     first           -> id = rid, shift = b                                   (:84-89)
//     id != rid or b - shift > 2000 -> flush two windows, shift = b   "reset"  (:91-100)
//     pos = b - shift; 1000 < pos < 2000 -> flush one window, shift += 1000 "roll" (:104-110)
// After every record id == rid, so the state that matters is p = b - shift of the record just processed:
// p in [0, 1000] or p == 2000 (the one value that neither rolls nor resets).  Everything else follows from p and the
// gaps g = b_next - b between consecutive qualifying records:
//     x = p + g;  x > 2000 (or another contig) -> reset, p' = 0;  1000 < x < 2000 -> p' = x - 1000;  else p' = x.
// Virtual coordinates: every window the reference ever holds gets 1000 consecutive positions, X = 1000 * v1 + p.
// A record that does not reset advances X by g (a roll moves 1000 from p to the window index); a reset jumps to the
// start of the window after next: X' = X - p + 2000.  So X is a prefix sum once the resets are known.
//
// Parallel form: a record is a CANDIDATE if it changes contig or its gap is >= 1000 (mod 2^32, so backward steps
// are candidates too).  Between candidates no reset can happen and p has the closed form cov_posf(); only
// candidates need the state.  A candidate that changes contig or whose gap is in [2001, 2^32 - 2001] resets from
// every state ("definite"); the others depend on p, and a block of them is a function on the 1002 possible states
// (all-states simulation, kernel_cov.cuh).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define BQC_CM __host__ __device__ __forceinline__
#else
#define BQC_CM inline
#endif

namespace bqc {

static const uint32_t kCovV = 1000;          // vsize (src/OverallNumbers.hpp:51)
static const uint32_t kCovEdge = 2000;       // p == 2 * vsize: no roll, no reset; every write of the record is lost (R9)
static const uint32_t kCovStates = 1002;     // p = 0..1000 and 2000

BQC_CM uint32_t cov_state_index(uint32_t p) { return p == kCovEdge ? 1001u : p; }
BQC_CM uint32_t cov_state_value(uint32_t s) { return s == 1001u ? kCovEdge : s; }

// p after walking a distance d from a window anchor without resets (every step < 1000): 0 stays 0, otherwise the
// anchor rolls so that p ends in 1..1000
BQC_CM uint32_t cov_posf(uint32_t d) { return d == 0u ? 0u : ((d - 1u) % kCovV) + 1u; }

// One record in state p (of the previous qualifying record) with gap g; `other` = contig changed.
BQC_CM uint32_t cov_step(uint32_t p, uint32_t g, bool other, bool& reset) {
    const uint32_t x = p + g;  // mod 2^32, as (beginpos - shift)
    reset = other || x > 2u * kCovV;
    if (reset) return 0u;
    if (x > kCovV && x < 2u * kCovV) return x - kCovV;
    return x;
}

// A gap that resets whatever the state is: p + g > 2000 for every p in [0, 2000] without wrapping.
BQC_CM bool cov_gap_definite(uint32_t g) { return (uint32_t)(g - 2001u) <= 0xFFFFFFFFu - 2000u - 2001u; }

// State of a record in the stretch that follows a candidate (all gaps in the stretch are < 1000 on the same contig):
// q = state right after the candidate, dd = b_record - b_candidate, ddj = b_record - b_j* where j* is the first
// record of the stretch whose begin differs from the candidate's (only used when q is the edge state: records at
// the candidate's own position stay at 2000, j* resets).
BQC_CM uint32_t cov_stretch(uint32_t q, uint32_t dd, uint32_t ddj) {
    if (q == kCovEdge) return dd == 0u ? kCovEdge : cov_posf(ddj);
    return cov_posf(q + dd);
}

// Advance of the virtual coordinate at a record: g if it does not reset, otherwise to the window after next.
BQC_CM uint32_t cov_dx(uint32_t p_before, uint32_t g, bool reset) { return reset ? 2u * kCovV - p_before : g; }

// Packed covered interval of one record relative to its begin (read-oriented CIGAR, S lengths added to the offset,
// M and D cover; src/OverallNumbers.hpp:112-134): c0 = offset of the first covered position, len = covered length
// (both saturated at 2047: everything from 2000 - p on is lost anyway), bit 31 = the record covers several
// separate intervals (an S between two M/D runs) and has to be re-walked.
static const uint32_t kCovComplex = 0x80000000u;
BQC_CM uint32_t cov_pack_iv(uint32_t c0, uint32_t len, bool complex) {
    return (c0 > 2047u ? 2047u : c0) | ((len > 2047u ? 2047u : len) << 11) | (complex ? kCovComplex : 0u);
}


// ------------------------------------------------------------------------------------------------
// Shards: one coordinate-ordered record stream cut at arbitrary records, every piece processed by its own engine
// (SURVEY 8e "the exception").  A shard that does not know what came before it works in its own virtual
// coordinates (origin = start of the window v1 that was open when it began) with empty windows, and reports
//   n, first/last (rid, begin) of its qualifying records,
//   F      its anchor recurrence as a function of the state at its entry (1002 values),
//   span   virtual coordinate of the window start of its last record,
//   head   difference array of its own depth over [0, 2000] (the two windows that were open at its entry),
//   tail   difference array of its own depth over [span, span + 2000] (the two windows open at its end).
// cov_shards_combine() then adds what the open windows of the shards before contribute to a shard's first 2000
// positions and flushes the last two windows (src/bamqualcheck.cpp:447-453).  Host only.
// ------------------------------------------------------------------------------------------------
struct CovShardPiece {
    uint64_t n;           // qualifying records (0: the shard leaves state and windows untouched)
    uint64_t span;
    const int32_t* head;  // [2001]
    const int32_t* tail;  // [2001]
};

// delta[101]: to be added to the sum of the shards' own poscov histograms
inline void cov_shards_combine(const CovShardPiece* sh, int n_shards, long long* delta) {
    for (int i = 0; i <= 100; ++i) delta[i] = 0;
    long long carry[2001];
    for (int i = 0; i <= 2000; ++i) carry[i] = 0;
    for (int k = 0; k < n_shards; ++k) {
        const CovShardPiece& S = sh[k];
        if (S.n == 0) continue;
        // the first min(2000, span) positions of the shard were counted with the shard's own depth only
        const uint64_t L = S.span < 2000 ? S.span : 2000;
        long long own = 0, both = 0;
        for (uint64_t i = 0; i < L; ++i) {
            own += S.head[i];
            both += S.head[i] + carry[i];
            delta[own > 100 ? 100 : own] -= 1;
            delta[both > 100 ? 100 : both] += 1;
        }
        // windows open after the shard: its own tail plus what is left of the incoming windows beyond span
        long long next[2001];
        for (int i = 0; i <= 2000; ++i) next[i] = S.tail[i];
        for (uint64_t d = 0; d <= 2000; ++d) {
            if (!carry[d]) continue;
            if (d <= S.span) next[0] += carry[d];
            else next[d - S.span] += carry[d];
        }
        for (int i = 0; i <= 2000; ++i) carry[i] = next[i];
    }
    long long depth = 0;   // end of the run: the two open windows
    for (int i = 0; i < 2000; ++i) { depth += carry[i]; delta[depth > 100 ? 100 : depth] += 1; }
}

}  // namespace bqc
