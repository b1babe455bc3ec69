// swar.h -- word-parallel (SIMD-within-a-register) forms of the per-base tests of the statistics pass.
// Host/device: the same functions are exercised on the CPU by tests/test_swar_cpu.py (via swar_selftest.cpp)
// against per-base restatements, and used by the kernels in kernel_stats.cuh.
//
// Conventions: a "nibble window" is 16 BAM 4-bit bases in a 64-bit word, base j in bits 4j..4j+3 (i.e. the bytes of
// SEQ with their two nibbles swapped); BAM codes are one-hot for A/C/G/T (1/2/4/8), 15 for N, other values are IUPAC
// ambiguity codes.  A "2-bit window" is 16 reference bases, base j in bits 2j..2j+1 (A0 C1 G2 T3).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define BQC_SW __host__ __device__ __forceinline__
#else
#define BQC_SW inline
#endif

namespace bqc {

static const uint64_t kNib1 = 0x1111111111111111ULL;

BQC_SW uint64_t swar_swap_nibbles(uint64_t x) { return ((x & 0x0F0F0F0F0F0F0F0FULL) << 4) | ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL); }

// Dna ordinal (A0 C1 G2 T3) in the low two bits of every nibble that is one-hot: code = (n >> 1) - (n >> 3).
// No borrow can leave a nibble (a nibble with bit 3 set has (n >> 1) >= 4), other nibbles hold garbage.
BQC_SW uint64_t swar_code4(uint64_t R) { return ((R >> 1) & 0x7777777777777777ULL) - ((R >> 3) & kNib1); }

// bit 4j set iff nibble j has exactly one bit set (the base is A, C, G or T: Dna5 ordinal != 4)
BQC_SW uint64_t swar_onehot4(uint64_t R) {
    const uint64_t s = (R & 0x5555555555555555ULL) + ((R >> 1) & 0x5555555555555555ULL);
    const uint64_t t = (s & 0x3333333333333333ULL) + ((s >> 2) & 0x3333333333333333ULL);  // popcount per nibble, 0..4
    const uint64_t u = t ^ kNib1;
    return ~(u | (u >> 1) | (u >> 2)) & kNib1;
}

// 8 two-bit codes (16 bits) -> low two bits of 8 nibbles
BQC_SW uint32_t swar_spread8(uint32_t v) {
    v = (v | (v << 8)) & 0x00FF00FFu;
    v = (v | (v << 4)) & 0x0F0F0F0Fu;
    v = (v | (v << 2)) & 0x33333333u;
    return v;
}
BQC_SW uint64_t swar_spread16(uint32_t F) { return (uint64_t)swar_spread8(F & 0xFFFFu) | ((uint64_t)swar_spread8(F >> 16) << 32); }

// TripletCounting::countBasesInTriplets (src/TripletCounting.hpp:213-232) for 16 consecutive read positions against
// the 16 reference bases they are aligned to.  T4: Dna ordinal per nibble; C4 bit 4j set iff read base j is A/C/G/T
// and read bases j-1 and j+1 both equal their reference bases (j = 1..14; bits 0 and 60 are never set).
BQC_SW void swar_triplet_masks(uint64_t R, uint32_t F, uint64_t& T4, uint64_t& C4) {
    T4 = swar_code4(R);
    const uint64_t oh = swar_onehot4(R);
    const uint64_t X = T4 ^ swar_spread16(F);
    const uint64_t m = oh & ~(X | (X >> 1));  // base j is exactly the reference base
    C4 = oh & (m << 4) & (m >> 4);
}

// ---- 8-base (32-bit) forms used by the per-cycle pass (QualityCheck::read_counts, src/QualityCheck.hpp:111-176)
BQC_SW uint32_t swar_swap_nibbles32(uint32_t x) { return ((x & 0x0F0F0F0Fu) << 4) | ((x >> 4) & 0x0F0F0F0Fu); }
// bit 4j set iff nibble j is one-hot; pop4 = number of set bits per nibble (0..4)
BQC_SW uint32_t swar_onehot8(uint32_t W, uint32_t& pop4) {
    const uint32_t s = (W & 0x55555555u) + ((W >> 1) & 0x55555555u);
    pop4 = (s & 0x33333333u) + ((s >> 2) & 0x33333333u);
    const uint32_t u = pop4 ^ 0x11111111u;
    return ~(u | (u >> 1) | (u >> 2)) & 0x11111111u;
}
// Dna5 ordinal per nibble: A0 C1 G2 T3 for one-hot nibbles, 4 for everything else (N, IUPAC codes, '=')
BQC_SW uint32_t swar_dna5_8(uint32_t W, uint32_t oh) {
    const uint32_t code = ((W >> 1) & 0x77777777u) - ((W >> 3) & 0x11111111u);
    return (code & (oh * 3u)) | ((~oh & 0x11111111u) << 2);
}

// bit 7 of every byte whose value q satisfies (signed char)(q + 33) >= '5', i.e. 20 <= q <= 94 (src/TripletCounting.hpp:214)
BQC_SW uint32_t swar_q20(uint32_t w) {
    const uint32_t l = w & 0x7F7F7F7Fu;
    return (l + 0x6C6C6C6Cu) & ~(l + 0x21212121u) & ~w & 0x80808080u;
}

// smem triplet index used by the kernel: context bits are (prev, cur, next) from low to high, the result tables
// order them (prev, cur, next) from high to low (src/TripletCounting.hpp:50-54)
BQC_SW uint32_t triplet_ctx_to_result(uint32_t c) { return ((c & 3u) << 4) | (c & 12u) | (c >> 4); }

}  // namespace bqc
