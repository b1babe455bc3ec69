// layout.h -- device/host shared description of the per-lane result block (all counters uint64).
//
// The block mirrors `struct Counts` of the reference (src/bamqualcheck.cpp:14-38): OverallNumbers
// (src/OverallNumbers.hpp:12-27), two QualityCheck instances (src/QualityCheck.hpp:12-25,42-46), the
// 64 TripletCounts (src/TripletCounting.hpp:29-46) and, per (q,k), StreamCounter::sumCount / F2table
// (src/kmerstream/StreamCounter.hpp:67-71).  The 4-bit sketch tables live in a separate array because
// they merge by saturating add, not by sum.  Fixed capacities replace the reference's growing
// String<>s; bqc_result_table() trims to the reference's lengths.
#pragma once
#include <stdint.h>

namespace bqc {

enum Scalar {
    S_SUPPLEMENTARY = 0, S_DUPLICATES, S_QCFAILED, S_NOT_PRIMARY, S_READCOUNT, S_TOTALBPS, S_BOTHUNMAPPED,
    S_FIRSTUNMAPPED, S_SECONDUNMAPPED, S_FIRST_AND_OR_SECOND_MAPPED, S_FF_RR, S_PROPERPAIR, S_AUTO_PROPERPAIR,
    S_COUNT = 16
};
enum PerCycleRow { PC_A = 0, PC_C, PC_G, PC_T, PC_N, PC_QUAL, PC_SC5, PC_SC3, PC_ROWS = 8 };

static const uint32_t kQCap = 256;     // average-quality histograms (raw phred byte <= 255)
static const uint32_t kMapqCap = 256;
static const uint32_t kPoscov = 104;   // 101 used
static const uint32_t kEightmer = 65536;
static const uint32_t kTriplet = 1024;
static const uint32_t kNone = 0xFFFFFFFFu;

struct Layout {
    uint32_t cyc;        // per-cycle capacity (max read length)
    uint32_t isize1;     // isize + 1
    uint32_t mmcap;      // mismatch / insertion histogram capacity
    uint32_t delcap;     // deletion histogram capacity
    uint32_t n_qk;       // number of (q,k) pairs
    uint32_t f2size;     // F2 table entries per sketch
    uint32_t sk_size;    // uint64 words per sketch level (StreamCounter::size)
    // offsets in uint64 units inside one lane block
    uint32_t o_scalars, o_poscov, o_insert, o_eightmer, o_triplet, o_mate0, mate_stride, o_qk, qk_stride;
    // offsets inside one mate block
    uint32_t m_readlen, m_ncount, m_gccount, m_avgq, m_ceilq, m_mapq, m_mismatch, m_del, m_ins, m_pc /* PC_ROWS x cyc */, m_readnr;
    uint64_t lane_stride;
};

#ifdef __CUDACC__
#define BQC_HD __host__ __device__
#else
#define BQC_HD
#endif
BQC_HD inline uint32_t pad8(uint32_t x) { return (x + 7u) & ~7u; }

inline Layout make_layout(uint32_t cyc, uint32_t isize, uint32_t n_qk, uint32_t f2size, uint32_t sk_size) {
    Layout L;
    L.cyc = cyc;
    L.isize1 = isize + 1;
    L.mmcap = pad8(cyc + 1);
    L.delcap = 4096;
    L.n_qk = n_qk;
    L.f2size = f2size;
    L.sk_size = sk_size;
    uint32_t m = 0;
    L.m_readlen = m;  m += pad8(cyc + 1);
    L.m_ncount = m;   m += pad8(cyc + 1);
    L.m_gccount = m;  m += pad8(cyc + 1);
    L.m_avgq = m;     m += kQCap;
    L.m_ceilq = m;    m += kQCap;
    L.m_mapq = m;     m += kMapqCap;
    L.m_mismatch = m; m += L.mmcap;
    L.m_del = m;      m += L.delcap;
    L.m_ins = m;      m += L.mmcap;
    L.m_pc = m;       m += PC_ROWS * pad8(cyc);
    L.m_readnr = m;   m += 8;
    L.mate_stride = m;
    uint32_t o = 0;
    L.o_scalars = o;  o += S_COUNT;
    L.o_poscov = o;   o += kPoscov;
    L.o_insert = o;   o += pad8(L.isize1);
    L.o_eightmer = o; o += kEightmer;
    L.o_triplet = o;  o += kTriplet;
    L.o_mate0 = o;    o += 2 * L.mate_stride;
    L.o_qk = o;
    L.qk_stride = 8 + f2size;  // [0] = sumCount, [8..] = F2 table
    o += n_qk * L.qk_stride;
    L.lane_stride = o;
    return L;
}

}  // namespace bqc
