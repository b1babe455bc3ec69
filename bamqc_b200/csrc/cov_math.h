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
// candidates need the state, and a block of candidates is a function on the 1002 possible states (all-states
// simulation, kernel_cov.cuh).
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

}  // namespace bqc
