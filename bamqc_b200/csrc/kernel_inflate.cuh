// kernel_inflate.cuh -- BGZF inflate on the device: the step right before the statistics pass (readRecord's hidden
// cost at src/bamqualcheck.cpp:306).  A BGZF file is a sequence of independent raw DEFLATE
// streams of at most 64 KiB of output each (SAM/BAM specification section 4.1; RFC 1951), so one WARP inflates one
// BGZF block and thousands of blocks are in flight:
//   * every lane holds the same bit window and decodes the same symbol (warp-uniform control flow, no
//     divergence); the Huffman tables of the current DEFLATE block live in shared memory (9-bit primary table
//     for literal/length codes, 7-bit for distances, canonical bit-by-bit search for the rare longer codes);
//   * literals are stored up to four at a time by the first lanes; a match is copied by all 32 lanes (overlapping
//     matches index the source modulo the distance, which reproduces the byte-by-byte semantics of LZ77);
//   * table construction from the code lengths is spread over the lanes (counts with shared-memory atomics,
//     canonical codes from the sorted symbol list, bit-reversed fan-out into the primary table).
// Blocks are handed out through an atomic ticket (compressed sizes vary).  Errors (bad block type, distance
// before the start of the block, output size different from ISIZE, input overrun) set a flag per BGZF block; the
// engine reports them as the reference's "Could not read record" failure.
//
// What bounds it (profiles/r2): a warp needs ~6 ms for a 64 KiB block whatever else runs, the symbol stream of BAM
// data is match-dominated, and every instruction of the symbol loop is one issue slot for one symbol, so throughput =
// resident warps x issue share / instructions per symbol.  Round 2 therefore (a) halved the tables (4.3 KB per warp
// -> 36 warps per SM instead of 30, CTAs of two warps so that finished blocks free their slots early), (b) lets the
// launches of consecutive submissions overlap (engine.cu, two inflate streams) because one 256 MB submission has
// fewer blocks than the GPU has warp slots, and (c) cut the instructions of a match: 32-bit window out of three
// resident input words (no refill between the length and the distance code), field extraction with bfe, table
// entries laid out in bytes, the modulo only for overlapping matches.
#pragma once
#include "inflate_bits.h"

namespace bqc {

struct InflateBlock {   // one BGZF block, filled by the host from the block headers (18 + XLEN bytes) and ISIZE
    uint32_t cbeg;      // offset of the raw DEFLATE payload in the compressed buffer
    uint32_t clen;      // payload bytes
    uint32_t obeg;      // offset of the block's output in the inflated buffer
    uint32_t isize;     // inflated size (BGZF trailer)
};

#ifndef BQC_INFLATE_LITBITS
#define BQC_INFLATE_LITBITS 9
#endif
#ifndef BQC_INFLATE_DISTBITS
#define BQC_INFLATE_DISTBITS 7
#endif
#ifndef BQC_INFLATE_WARPS
#define BQC_INFLATE_WARPS 2
#endif
#ifndef BQC_INFLATE_MINBLOCKS
#define BQC_INFLATE_MINBLOCKS 18
#endif
static const uint32_t kInflateWarps = BQC_INFLATE_WARPS;          // warps per CTA = BGZF blocks in flight per CTA
static const uint32_t kInflateStreams = kInflateWarps;
static const uint32_t kLitBits = BQC_INFLATE_LITBITS, kDistBits = BQC_INFLATE_DISTBITS, kClBits = 7;

// Table entries are packed so that one shared-memory load yields everything a symbol needs, each field in its own
// byte (one PRMT / LOP to extract):
//   byte 0: bits 0-3 code length (0 = the code is longer than the primary index: canonical search), bit 4 the value
//           is a base with extra bits (length / distance), bit 5 end of block (bits 4+5: a symbol that must not
//           occur), bit 6 literal;
//   byte 1: number of extra bits;   bytes 2-3: value (literal byte, length base, distance base; the symbol itself
//           for the code-length alphabet).
static const uint32_t kIsBase = 1u << 4, kIsEob = 1u << 5, kIsInvalid = kIsBase | kIsEob, kIsLiteral = 1u << 6;

struct alignas(16) InflateTabs {                   // per warp, shared memory
    uint32_t lit[1u << kLitBits];
    uint32_t dist[1u << kDistBits];
    uint32_t cl[1u << kClBits];
    uint16_t lit_sorted[288], dist_sorted[32], cl_sorted[20];   // symbols ordered by (length, symbol)
    uint16_t lit_count[16], dist_count[16], cl_count[16];       // symbols per code length
    uint16_t first[16], off0[16], offs[16];        // builder scratch: first canonical code / sorted offset per length
    uint16_t slow_first[4], slow_index[4];         // state of the canonical search after the primary bits: [0] lit [1] dist [2] cl
    uint8_t lens[320];                             // code lengths of the literal/length + distance alphabets
    uint8_t cl_lens[32];                           // code lengths of the code-length alphabet
};

__constant__ uint16_t c_len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t c_len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t c_dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__constant__ uint8_t c_dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__constant__ uint8_t c_cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
// floor(65536 / d) + 1: (x * c_rcp[d]) >> 16 == x / d for x <= 32, d in 2..31
__constant__ uint16_t c_rcp[32] = {0, 0, 32769, 21846, 16385, 13108, 10923, 9363, 8193, 7282, 6554, 5958, 5462, 5042, 4682, 4370,
                                   4097, 3856, 3641, 3450, 3277, 3121, 2979, 2850, 2731, 2622, 2521, 2428, 2341, 2260, 2185, 2115};

// packed entry of a symbol (without the code length)
__device__ __forceinline__ uint32_t inflate_lit_entry(uint32_t s) {
    if (s < 256u) return (s << 16) | kIsLiteral;
    if (s == 256u) return kIsEob;
    if (s > 285u) return kIsInvalid;
    return kIsBase | ((uint32_t)c_len_extra[s - 257u] << 8) | ((uint32_t)c_len_base[s - 257u] << 16);
}
__device__ __forceinline__ uint32_t inflate_dist_entry(uint32_t s) {
    if (s > 29u) return kIsInvalid;
    return kIsBase | ((uint32_t)c_dist_extra[s] << 8) | ((uint32_t)c_dist_base[s] << 16);
}
__device__ __forceinline__ uint32_t inflate_cl_entry(uint32_t s) { return s << 16; }

__device__ __forceinline__ uint32_t bfe32(uint32_t x, uint32_t pos, uint32_t len) {   // len bits of x from bit pos (pos, len < 32)
    uint32_t m;
    asm("bmsk.clamp.b32 %0, %1, %2;" : "=r"(m) : "r"(0u), "r"(len));   // (bfe.u32 costs three more PRMTs: it truncates pos and len to 8 bits)
    return (x >> pos) & m;
}
// table[w & mask] of a table at shared address table_saddr.  The tables are rewritten (plain stores) only between
// __syncwarp()s outside the symbol loop, so the loads need no memory clobber.
__device__ __forceinline__ uint32_t inflate_lds(uint32_t table_saddr, uint32_t w, uint32_t mask) {
    uint32_t a, e;
    asm("and.b32 %0, %1, %2;\n\tmad.lo.u32 %0, %0, 4, %3;" : "=&r"(a) : "r"(w), "r"(mask), "r"(table_saddr));
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"(a));
    return e;
}

// BitWin, the bit reader (three resident input words and a bit position), lives in inflate_bits.h: host/device code,
// checked on the CPU by tests/inflate_selftest.cpp.

// Build the decoding tables of one alphabet from lens[0..n): count[], sorted[], the primary table (PB index bits)
// and the entry state of the canonical search for longer codes.  WHICH: 0 literal/length, 1 distance, 2 code
// lengths.  Returns false for an over-subscribed code.  Called by all lanes.
template <uint32_t PB, uint32_t WHICH>
__device__ __forceinline__ bool inflate_build(InflateTabs& T, const uint8_t* lens, uint32_t n, uint16_t* count, uint16_t* sorted, uint32_t* tab, uint32_t lane) {
    __syncwarp();
    if (lane < 8u) reinterpret_cast<uint32_t*>(count)[lane] = 0;
    for (uint32_t i = lane; i < (1u << PB); i += 32u) tab[i] = 0;
    __syncwarp();
    for (uint32_t s = lane; s < n; s += 32u) {  // 16-bit counters updated through their 32-bit word
        const uint32_t l = lens[s];
        atomicAdd(reinterpret_cast<uint32_t*>(count) + (l >> 1), 1u << (16 * (l & 1)));
    }
    __syncwarp();
    if (lane == 0) {
        uint32_t code = 0, off = 0, left = 1, bad = 0;
        for (uint32_t l = 1; l < 16; ++l) {
            code <<= 1;
            left <<= 1;
            const uint32_t c = count[l];
            if (c > left) { bad = 1; left = 0; } else left -= c;
            T.first[l] = (uint16_t)code;
            T.off0[l] = (uint16_t)off;
            T.offs[l] = (uint16_t)off;
            code += c;
            off += c;
            if (l == PB) { T.slow_first[WHICH] = (uint16_t)(code << 1); T.slow_index[WHICH] = (uint16_t)off; }
        }
        T.offs[0] = (uint16_t)bad;
        T.off0[0] = (uint16_t)off;  // number of coded symbols
        if (!bad)
            for (uint32_t s = 0; s < n; ++s) {  // symbols of one length keep their order
                const uint32_t l = lens[s];
                if (l) sorted[T.offs[l]++] = (uint16_t)s;
            }
    }
    __syncwarp();
    if (T.offs[0]) return false;
    const uint32_t total = T.off0[0];
    for (uint32_t i = lane; i < total; i += 32u) {
        const uint32_t s = sorted[i];
        const uint32_t l = lens[s];
        if (l <= PB) {
            const uint32_t code = (uint32_t)T.first[l] + (i - (uint32_t)T.off0[l]);
            const uint32_t r = __brev(code) >> (32u - l);   // Huffman codes are packed starting from their MSB
            const uint32_t entry = l | (WHICH == 0 ? inflate_lit_entry(s) : WHICH == 1 ? inflate_dist_entry(s) : inflate_cl_entry(s));
            for (uint32_t k = r; k < (1u << PB); k += (1u << l)) tab[k] = entry;
        }
    }
    __syncwarp();
    return true;
}

// Codes longer than the primary index: canonical search one bit at a time over the 32-bit window `w`, entered with
// the state it has after PB bits (no shorter code matched, or the primary entry would exist).  Returns the packed
// entry with the full code length, or 0 if no code matches.
template <uint32_t PB, uint32_t WHICH>
__device__ __noinline__ uint32_t inflate_decode_slow(const InflateTabs& T, uint32_t w, const uint16_t* count, const uint16_t* sorted) {
    uint32_t bits = w >> PB;
    int code = (int)((__brev(w) >> (32u - PB)) << 1);
    int first = T.slow_first[WHICH], index = T.slow_index[WHICH];
    for (uint32_t len = PB + 1; len < 16; ++len) {
        code |= (int)(bits & 1u);
        bits >>= 1;
        const int c = count[len];
        if (code >= first && code - c < first) {  // (code >= first always holds for a valid code set)
            const uint32_t s = sorted[index + (code - first)];
            return len | (WHICH == 0 ? inflate_lit_entry(s) : WHICH == 1 ? inflate_dist_entry(s) : inflate_cl_entry(s));
        }
        index += c;
        first += c;
        first <<= 1;
        code <<= 1;
    }
    return 0;
}

// ctl[0] = ticket, ctl[1] = 1 + index of the first BGZF block that failed to inflate (0 = none; atomicMax on the
// bitwise complement so that a zeroed word means "none")
__global__ void __launch_bounds__(kInflateWarps * 32, BQC_INFLATE_MINBLOCKS) k_inflate(const uint8_t* __restrict__ cin, const InflateBlock* __restrict__ blocks, uint32_t n_blocks,
                                                                                        uint8_t* out, uint32_t* ctl) {
    extern __shared__ __align__(16) uint8_t inflate_smem[];
    InflateTabs* tabs = reinterpret_cast<InflateTabs*>(inflate_smem);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lane8 = (lane & 3u) * 8u;
    InflateTabs& T = tabs[threadIdx.x / 32u];
    const uint32_t lit_sa = (uint32_t)__cvta_generic_to_shared(T.lit), dist_sa = (uint32_t)__cvta_generic_to_shared(T.dist);
    const uint32_t kLitMask = (1u << kLitBits) - 1u, kDistMask = (1u << kDistBits) - 1u;
    for (;;) {
        uint32_t b = 0;
        if (lane == 0) b = atomicAdd(ctl, 1u);
        b = __shfl_sync(0xFFFFFFFFu, b, 0);
        if (b >= n_blocks) break;
        const InflateBlock blk = blocks[b];
        uint8_t* o = out + blk.obeg;
        uint8_t* ol = o + lane;    // this lane's column of the output: lane j of a step at offset p writes ol[p]
        const uint32_t isize = blk.isize;
        uint32_t pos = 0;
        BitWin br;
        br.init(cin + blk.cbeg, blk.clen);
        bool ok = true;
        uint32_t last = 0;
        uint32_t pend = kNone;     // deferred store of the last step of the previous match (per lane): offset in ol
        uint32_t pend_val = 0;
        while (ok && !last) {
            const uint32_t hdr = br.take(3);
            last = hdr & 1u;
            const uint32_t type = hdr >> 1;
            if (type == 0u) {  // stored (RFC 1951 3.2.4)
                br.bp = (br.bp + 7u) & ~7u;   // bit0 and the word size are multiples of 8
                const uint32_t len = br.take(16);
                const uint32_t nlen = br.take(16);
                const uint32_t p = br.bytes_used();
                if (len != (~nlen & 0xFFFFu) || pos + len > isize || p + len > blk.clen) { ok = false; break; }
                const uint8_t* src = cin + blk.cbeg + p;
                for (uint32_t j = lane; j < len; j += 32u) o[pos + j] = ldg8(src + j);
                pos += len;
                br.seek(p + len);
                continue;
            }
            if (type == 3u) { ok = false; break; }
            uint32_t nlit = 288, ndist = 30;
            if (type == 1u) {  // fixed codes (3.2.6)
                for (uint32_t s = lane; s < 288; s += 32u) T.lens[s] = (uint8_t)(s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8);
                if (lane < 30u) T.lens[288 + lane] = 5;
            } else {           // dynamic codes (3.2.7)
                const uint32_t h = br.take(14);
                nlit = (h & 31u) + 257u;
                ndist = ((h >> 5) & 31u) + 1u;
                const uint32_t ncl = (h >> 10) + 4u;
                if (nlit > 286 || ndist > 30) { ok = false; break; }
                if (lane < 19u) T.cl_lens[lane] = 0;
                __syncwarp();
                for (uint32_t i = 0; i < ncl; ++i) {
                    const uint32_t v = br.take(3);
                    if (lane == 0) T.cl_lens[c_cl_order[i]] = (uint8_t)v;
                }
                if (!inflate_build<kClBits, 2>(T, T.cl_lens, 19, T.cl_count, T.cl_sorted, T.cl, lane)) { ok = false; break; }
                const uint32_t total = nlit + ndist;
                uint32_t i = 0;
                while (i < total) {
                    br.norm();
                    const uint32_t w = br.win();
                    const uint32_t ce = T.cl[w & ((1u << kClBits) - 1u)];
                    if (!ce) { ok = false; break; }  // code-length codes are at most 7 bits: every valid code is in the table
                    const uint32_t cl = ce & 15u;
                    const uint32_t sym = ce >> 16;
                    if (sym < 16u) {
                        br.bp += cl;
                        if (lane == 0) T.lens[i] = (uint8_t)sym;
                        ++i;
                        __syncwarp();
                        continue;
                    }
                    const uint32_t xb = sym == 16u ? 2u : sym == 17u ? 3u : 7u;
                    const uint32_t x = (w >> cl) & ((1u << xb) - 1u);
                    br.bp += cl + xb;
                    uint32_t val = 0;
                    const uint32_t rep = (sym == 18u ? 11u : 3u) + x;
                    if (sym == 16u) {
                        if (i == 0) { ok = false; break; }
                        val = T.lens[i - 1];
                    }
                    if (i + rep > total) { ok = false; break; }
                    for (uint32_t j = lane; j < rep; j += 32u) T.lens[i + j] = (uint8_t)val;
                    i += rep;
                    __syncwarp();
                }
                if (!ok) break;
                if (T.lens[256] == 0) { ok = false; break; }  // no end-of-block code
            }
            if (!inflate_build<kLitBits, 0>(T, T.lens, nlit, T.lit_count, T.lit_sorted, T.lit, lane)) { ok = false; break; }
            if (!inflate_build<kDistBits, 1>(T, T.lens + nlit, ndist, T.dist_count, T.dist_sorted, T.dist, lane)) { ok = false; break; }
            // ---- symbols of this block ------------------------------------------------------------------
            for (;;) {
                br.norm();
                uint32_t w = br.win();
                uint32_t e = inflate_lds(lit_sa, w, kLitMask);
                if (e & kIsLiteral) {
                    // Literals come in runs: the window holds 32 bits, enough for three codes of the primary table
                    // and usually a fourth.
                    uint32_t l = e & 15u, used = l, b0 = e >> 16, nl = 1;
                    w >>= l;
                    e = inflate_lds(lit_sa, w, kLitMask);
                    if (e & kIsLiteral) {
                        l = e & 15u;
                        used += l;
                        w >>= l;
                        b0 |= (e >> 8) & 0xFF00u;
                        nl = 2;
                        e = inflate_lds(lit_sa, w, kLitMask);
                        if (e & kIsLiteral) {
                            l = e & 15u;
                            used += l;
                            w >>= l;
                            b0 |= e & 0xFF0000u;
                            nl = 3;
                            e = inflate_lds(lit_sa, w, kLitMask);   // zeros above the window: valid only if the code fits
                            if ((e & kIsLiteral) && used + (e & 15u) <= 32u) {
                                used += e & 15u;
                                b0 |= (e << 8) & 0xFF000000u;
                                nl = 4;
                            }
                        }
                    }
                    br.bp += used;
                    if (pos + nl > isize) { ok = false; break; }
                    if (lane < nl) ol[pos] = (uint8_t)(b0 >> lane8);
                    pos += nl;
                    continue;
                }
                uint32_t l = e & 15u;
                if (!l) {
                    e = inflate_decode_slow<kLitBits, 0>(T, w, T.lit_count, T.lit_sorted);
                    if (!e) { ok = false; break; }
                    l = e & 15u;
                    if (e & kIsLiteral) {   // a literal with a long code
                        br.bp += l;
                        if (pos >= isize) { ok = false; break; }
                        if (lane == 0) o[pos] = (uint8_t)(e >> 16);
                        pos += 1;
                        continue;
                    }
                }
                if (e & kIsEob) {     // end of block, or a symbol that must not occur
                    br.bp += l;
                    if (e & kIsBase) ok = false;
                    break;
                }
                const uint32_t eb = __byte_perm(e, 0, 0x4441);
                const uint32_t len = (e >> 16) + bfe32(w, l, eb);
                br.bp += l + eb;          // <= 15 + 5 bits: bp < 52, the distance code is still inside lo:hi:nx
                w = br.win2();
                uint32_t d = inflate_lds(dist_sa, w, kDistMask);
                uint32_t dl = d & 15u;
                if (!dl) {
                    d = inflate_decode_slow<kDistBits, 1>(T, w, T.dist_count, T.dist_sorted);
                    if (!d || (d & kIsEob)) { ok = false; break; }   // (symbols 30/31 never reach the primary table: ndist <= 30)
                    dl = d & 15u;
                }
                const uint32_t deb = __byte_perm(d, 0, 0x4441);
                const uint32_t dist = (d >> 16) + bfe32(w, dl, deb);
                br.bp += dl + deb;        // <= 15 + 13 bits: bp < 80
                if (dist > pos || pos + len > isize) { ok = false; break; }
                // The copy is software-pipelined: the last step of a match is loaded now and stored when the
                // next match arrives (or at the end of the BGZF block), so the L2 round trip of the load overlaps
                // the decoding of the following symbols instead of stalling the warp at the store.
                if (pend != kNone) { ol[pend] = (uint8_t)pend_val; pend = kNone; }
                __syncwarp();  // the bytes the match refers to were stored by other lanes
                const uint32_t sp = pos - dist;
                if (len <= 32u && dist >= len) {   // the common case: one step, source and destination apart
                    if (lane < len) { pend_val = ol[sp]; pend = pos; }
                } else {
                    // an overlapping match repeats with period dist: lane j reads byte j mod dist
                    const uint32_t rc = dist < 32u ? (uint32_t)c_rcp[dist] : 0u;
                    uint32_t m = dist == 1u ? 0u : lane - dist * ((lane * rc) >> 16);   // lane % dist (lane itself when dist >= 32)
                    if (len <= 32u) {
                        if (lane < len) { pend_val = o[sp + m]; pend = pos; }
                    } else {
                        uint32_t j = lane;
                        if (dist >= len) {
                            for (; j + 32u < len + lane; j += 32u) o[pos + j] = o[sp + j];   // all but the last step (uniform trip count)
                            if (j < len) { pend_val = o[sp + j]; pend = pos + j - lane; }
                        } else if (dist < 32u) {
                            const uint32_t step = dist == 1u ? 0u : 32u - dist * ((32u * rc) >> 16);   // 32 % dist
                            for (; j + 32u < len + lane; j += 32u) {
                                o[pos + j] = o[sp + m];
                                m += step;
                                if (m >= dist) m -= dist;
                            }
                            if (j < len) { pend_val = o[sp + m]; pend = pos + j - lane; }
                        } else {
                            for (; j + 32u < len + lane; j += 32u) o[pos + j] = o[sp + j % dist];
                            if (j < len) { pend_val = o[sp + j % dist]; pend = pos + j - lane; }
                        }
                    }
                }
                pos += len;
            }
            if (br.bytes_used() > blk.clen + 8u) ok = false;  // ran past the payload
        }
        if (pend != kNone) ol[pend] = (uint8_t)pend_val;
        if (ok && (pos != isize || br.bytes_used() > blk.clen)) ok = false;
        if (!ok && lane == 0) atomicMax(ctl + 1, 0xFFFFFFFFu - b);  // largest complement = smallest index
        __syncwarp();
    }
}

}  // namespace bqc
