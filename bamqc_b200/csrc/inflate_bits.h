// inflate_bits.h -- the bit reader of k_inflate (kernel_inflate.cuh), host/device: tests/inflate_selftest.cpp runs the
// same code on the CPU against a plain LSB-first bit reader (RFC 1951 3.1.1).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define BQC_IB __host__ __device__ __forceinline__
#else
#define BQC_IB inline
#endif

namespace bqc {

BQC_IB uint32_t ib_load(const uint32_t* p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}
BQC_IB uint32_t ib_funnel_r(uint32_t lo, uint32_t hi, uint32_t shift) {   // (hi:lo) >> (shift mod 32), low word
#ifdef __CUDA_ARCH__
    return __funnelshift_r(lo, hi, shift);
#else
    return (uint32_t)((((uint64_t)hi << 32) | lo) >> (shift & 31u));
#endif
}

// The input as three consecutive aligned 32-bit words and a bit position: a 32-bit window is one funnel shift,
// dropping bits is one add, and a whole trip (up to four literals, or a length and a distance with their extra
// bits) is decoded between two refills.  Identical in every lane.
struct BitWin {
    const uint32_t* base;   // aligned word that holds the first payload byte
    uint32_t wi;            // index of the word in nx
    uint32_t wlim;          // last word index that may be loaded (payload + one refill of slack)
    uint32_t lo, hi, nx;
    uint32_t bp;            // bits of `lo` consumed; < 32 after norm(), < 96 always
    uint32_t bit0;          // misalignment of the payload in bits
    BQC_IB uint32_t load(uint32_t w) const { return ib_load(base + (w < wlim ? w : wlim)); }
    BQC_IB void seek(uint32_t byte) {
        const uint32_t bits = bit0 + byte * 8u, w = bits >> 5;
        bp = bits & 31u;
        lo = load(w);
        hi = load(w + 1u);
        nx = load(w + 2u);
        wi = w + 2u;
    }
    BQC_IB void init(const uint8_t* p, uint32_t clen) {
        const uintptr_t a = (uintptr_t)p;
        base = (const uint32_t*)(a & ~(uintptr_t)3);
        bit0 = (uint32_t)(a & 3) * 8u;
        wlim = (bit0 + clen * 8u + 31u) / 32u + 2u;
        seek(0);
    }
    BQC_IB void shift() {
        lo = hi;
        hi = nx;
        ++wi;
        nx = load(wi);
        bp -= 32u;
    }
    BQC_IB void norm() {
        if (bp >= 32u) {
            shift();
            if (bp >= 32u) shift();   // a match with long codes and many extra bits (up to 48 bits in one trip)
        }
    }
    BQC_IB uint32_t win() const { return ib_funnel_r(lo, hi, bp); }   // 32 valid bits after norm()
    // 32 bits from bp for 32 <= bp < 64 as well (the distance code right after a length, without a refill)
    BQC_IB uint32_t win2() const {
        const bool up = bp >= 32u;
        return ib_funnel_r(up ? hi : lo, up ? nx : hi, bp);   // the shift amount wraps at 32
    }
    BQC_IB uint32_t take(uint32_t n) {  // n <= 16 (header fields)
        norm();
        const uint32_t v = win() & ((1u << n) - 1u);
        bp += n;
        return v;
    }
    BQC_IB uint32_t bits_used() const { return (wi - 2u) * 32u + bp - bit0; }
    BQC_IB uint32_t bytes_used() const { return (bits_used() + 7u) >> 3; }   // a partial byte counts
};

}  // namespace bqc
