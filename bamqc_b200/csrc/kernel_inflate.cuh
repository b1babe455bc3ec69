// kernel_inflate.cuh -- BGZF inflate on the device: the step right before the statistics pass (readRecord's hidden
// cost at src/bamqualcheck.cpp:306).  A BGZF file is a sequence of independent raw DEFLATE
// streams of at most 64 KiB of output each (SAM/BAM specification section 4.1; RFC 1951), so one WARP inflates one
// BGZF block and thousands of blocks are in flight:
//   * every lane holds the same bit buffer and decodes the same symbol (warp-uniform control flow, no
//     divergence); the Huffman tables of the current DEFLATE block live in shared memory (10-bit primary table
//     for literal/length codes, 8-bit for distances, canonical bit-by-bit search for the rare longer codes);
//   * a literal is stored by one lane; a match is copied by all 32 lanes (overlapping matches index the source
//     modulo the distance, which reproduces the byte-by-byte semantics of LZ77);
//   * table construction from the code lengths is spread over the lanes (counts with shared-memory atomics,
//     canonical codes from the sorted symbol list, bit-reversed fan-out into the primary table).
// Blocks are handed out through an atomic ticket (compressed sizes vary).  Errors (bad block type, distance
// before the start of the block, output size different from ISIZE, input overrun) set a flag per BGZF block; the
// engine reports them as the reference's "Could not read record" failure.
#pragma once

namespace bqc {

struct InflateBlock {   // one BGZF block, filled by the host from the block headers (18 + XLEN bytes) and ISIZE
    uint32_t cbeg;      // offset of the raw DEFLATE payload in the compressed buffer
    uint32_t clen;      // payload bytes
    uint32_t obeg;      // offset of the block's output in the inflated buffer
    uint32_t isize;     // inflated size (BGZF trailer)
};

static const uint32_t kInflateWarps = 8;          // warps per CTA
static const uint32_t kLitBits = 10, kDistBits = 8, kClBits = 7;

struct alignas(16) InflateTabs {                   // per warp, shared memory
    uint16_t lit[1u << kLitBits];                  // (code length << 9) | symbol; 0 = longer than kLitBits
    uint16_t dist[1u << kDistBits];                // (code length << 5) | symbol
    uint16_t cl[1u << kClBits];                    // code-length alphabet
    uint16_t lit_sorted[288], dist_sorted[32], cl_sorted[20];   // symbols ordered by (length, symbol)
    uint16_t lit_count[16], dist_count[16], cl_count[16];       // symbols per code length
    uint16_t first[16], off0[16], offs[16];        // builder scratch: first canonical code / sorted offset per length
    uint8_t lens[320];                             // code lengths of the literal/length + distance alphabets
    uint8_t cl_lens[32];                           // code lengths of the code-length alphabet
};

__constant__ uint16_t c_len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t c_len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t c_dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__constant__ uint8_t c_dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__constant__ uint8_t c_cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

struct BitReader {  // LSB-first bit stream (RFC 1951 section 3.1.1); identical in every lane
    const uint8_t* in;
    uint32_t pos;     // next byte to load
    uint64_t buf;
    uint32_t cnt;     // valid bits in buf
    __device__ __forceinline__ void init(const uint8_t* p) { in = p; pos = 0; buf = 0; cnt = 0; }
    __device__ __forceinline__ void refill() {  // afterwards cnt >= 32
        if (cnt < 32u) {
            buf |= (uint64_t)ldu32(in + pos) << cnt;
            pos += 4;
            cnt += 32;
        }
    }
    __device__ __forceinline__ uint32_t peek(uint32_t n) const { return (uint32_t)buf & ((1u << n) - 1u); }
    __device__ __forceinline__ void drop(uint32_t n) { buf >>= n; cnt -= n; }
    __device__ __forceinline__ uint32_t take(uint32_t n) { uint32_t v = peek(n); drop(n); return v; }
    __device__ __forceinline__ uint32_t bytes_used() const { return pos - (cnt >> 3); }  // bytes consumed (partial byte counts)
};

// Build the decoding tables of one alphabet from lens[0..n): count[], sorted[] and the primary table (PB index
// bits, entry = length << SB | symbol).  Returns false for an over-subscribed code.  Called by all lanes.
template <uint32_t PB, uint32_t SB>
__device__ __forceinline__ bool inflate_build(InflateTabs& T, const uint8_t* lens, uint32_t n, uint16_t* count, uint16_t* sorted, uint16_t* tab, uint32_t lane) {
    __syncwarp();
    if (lane < 8) reinterpret_cast<uint32_t*>(count)[lane] = 0;
    for (uint32_t i = lane; i < (1u << PB) / 2; i += 32) reinterpret_cast<uint32_t*>(tab)[i] = 0;
    __syncwarp();
    for (uint32_t s = lane; s < n; s += 32) {  // 16-bit counters updated through their 32-bit word
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
    for (uint32_t i = lane; i < total; i += 32) {
        const uint32_t s = sorted[i];
        const uint32_t l = lens[s];
        if (l <= PB) {
            const uint32_t code = (uint32_t)T.first[l] + (i - (uint32_t)T.off0[l]);
            const uint32_t r = __brev(code) >> (32u - l);   // Huffman codes are packed starting from their MSB
            const uint16_t entry = (uint16_t)((l << SB) | s);
            for (uint32_t k = r; k < (1u << PB); k += (1u << l)) tab[k] = entry;
        }
    }
    __syncwarp();
    return true;
}

// canonical decode, one bit at a time (codes longer than the primary table); needs >= 15 bits in the buffer
__device__ __forceinline__ int inflate_decode_slow(BitReader& br, const uint16_t* count, const uint16_t* sorted) {
    uint32_t bits = (uint32_t)br.buf;
    int code = 0, first = 0, index = 0;
    for (uint32_t len = 1; len < 16; ++len) {
        code |= (int)(bits & 1u);
        bits >>= 1;
        const int c = count[len];
        if (code - c < first) {
            br.drop(len);
            return sorted[index + (code - first)];
        }
        index += c;
        first += c;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}
template <uint32_t PB, uint32_t SB>
__device__ __forceinline__ int inflate_decode(BitReader& br, const uint16_t* tab, const uint16_t* count, const uint16_t* sorted) {
    const uint32_t e = tab[br.peek(PB)];
    if (e) {
        br.drop(e >> SB);
        return (int)(e & ((1u << SB) - 1u));
    }
    return inflate_decode_slow(br, count, sorted);
}

// ctl[0] = ticket, ctl[1] = 1 + index of the first BGZF block that failed to inflate (0 = none; atomicMin on the
// bitwise complement so that a zeroed word means "none")
__global__ void __launch_bounds__(kInflateWarps * 32) k_inflate(const uint8_t* __restrict__ cin, const InflateBlock* __restrict__ blocks, uint32_t n_blocks,
                                                                 uint8_t* out, uint32_t* ctl) {
    __shared__ InflateTabs tabs[kInflateWarps];
    const uint32_t lane = threadIdx.x & 31u;
    InflateTabs& T = tabs[threadIdx.x >> 5];
    for (;;) {
        uint32_t b = 0;
        if (lane == 0) b = atomicAdd(ctl, 1u);
        b = __shfl_sync(0xFFFFFFFFu, b, 0);
        if (b >= n_blocks) break;
        const InflateBlock blk = blocks[b];
        uint8_t* o = out + blk.obeg;
        const uint32_t isize = blk.isize;
        uint32_t pos = 0;
        BitReader br;
        br.init(cin + blk.cbeg);
        bool ok = true;
        uint32_t last = 0;
        while (ok && !last) {
            br.refill();
            last = br.take(1);
            const uint32_t type = br.take(2);
            if (type == 0u) {  // stored (RFC 1951 3.2.4)
                br.drop(br.cnt & 7u);
                br.refill();
                const uint32_t len = br.take(16);
                br.refill();
                const uint32_t nlen = br.take(16);
                const uint32_t p = br.bytes_used();
                if (len != (~nlen & 0xFFFFu) || pos + len > isize || p + len > blk.clen) { ok = false; break; }
                for (uint32_t j = lane; j < len; j += 32) o[pos + j] = ldg8(br.in + p + j);
                pos += len;
                br.pos = p + len;
                br.buf = 0;
                br.cnt = 0;
                continue;
            }
            if (type == 3u) { ok = false; break; }
            uint32_t nlit = 288, ndist = 30;
            if (type == 1u) {  // fixed codes (3.2.6)
                for (uint32_t s = lane; s < 288; s += 32) T.lens[s] = (uint8_t)(s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8);
                if (lane < 30) T.lens[288 + lane] = 5;
            } else {           // dynamic codes (3.2.7)
                br.refill();
                nlit = br.take(5) + 257;
                ndist = br.take(5) + 1;
                const uint32_t ncl = br.take(4) + 4;
                if (nlit > 286 || ndist > 30) { ok = false; break; }
                if (lane < 19) T.cl_lens[lane] = 0;
                __syncwarp();
                for (uint32_t i = 0; i < ncl; ++i) {
                    br.refill();
                    const uint32_t v = br.take(3);
                    if (lane == 0) T.cl_lens[c_cl_order[i]] = (uint8_t)v;
                }
                if (!inflate_build<kClBits, 5>(T, T.cl_lens, 19, T.cl_count, T.cl_sorted, T.cl, lane)) { ok = false; break; }
                const uint32_t total = nlit + ndist;
                uint32_t i = 0;
                while (i < total) {
                    br.refill();
                    const int sym = inflate_decode<kClBits, 5>(br, T.cl, T.cl_count, T.cl_sorted);
                    if (sym < 0) { ok = false; break; }
                    if (sym < 16) {
                        if (lane == 0) T.lens[i] = (uint8_t)sym;
                        ++i;
                        __syncwarp();
                        continue;
                    }
                    uint32_t val = 0, rep;
                    if (sym == 16) {
                        if (i == 0) { ok = false; break; }
                        val = T.lens[i - 1];
                        rep = 3 + br.take(2);
                    } else if (sym == 17) rep = 3 + br.take(3);
                    else rep = 11 + br.take(7);
                    if (i + rep > total) { ok = false; break; }
                    for (uint32_t j = lane; j < rep; j += 32) T.lens[i + j] = (uint8_t)val;
                    i += rep;
                    __syncwarp();
                }
                if (!ok) break;
                if (T.lens[256] == 0) { ok = false; break; }  // no end-of-block code
            }
            if (!inflate_build<kLitBits, 9>(T, T.lens, nlit, T.lit_count, T.lit_sorted, T.lit, lane)) { ok = false; break; }
            if (!inflate_build<kDistBits, 5>(T, T.lens + nlit, ndist, T.dist_count, T.dist_sorted, T.dist, lane)) { ok = false; break; }
            // ---- symbols of this block ------------------------------------------------------------------
            for (;;) {
                br.refill();
                int sym = inflate_decode<kLitBits, 9>(br, T.lit, T.lit_count, T.lit_sorted);
                if (sym < 0) { ok = false; break; }
                if (sym < 256) {
                    if (pos >= isize) { ok = false; break; }
                    if (lane == 0) o[pos] = (uint8_t)sym;
                    ++pos;
                    continue;
                }
                if (sym == 256) break;
                sym -= 257;
                if (sym >= 29) { ok = false; break; }
                const uint32_t len = c_len_base[sym] + br.take(c_len_extra[sym]);
                br.refill();
                const int dsym = inflate_decode<kDistBits, 5>(br, T.dist, T.dist_count, T.dist_sorted);
                if (dsym < 0 || dsym >= 30) { ok = false; break; }
                const uint32_t dist = c_dist_base[dsym] + br.take(c_dist_extra[dsym]);
                if (dist > pos || pos + len > isize) { ok = false; break; }
                __syncwarp();  // the bytes the match refers to were stored by other lanes
                const uint8_t* src = o + pos - dist;
                if (dist >= len) {
                    for (uint32_t j = lane; j < len; j += 32) o[pos + j] = src[j];
                } else {       // overlapping match: byte j repeats with period dist
                    for (uint32_t j = lane; j < len; j += 32) o[pos + j] = src[j % dist];
                }
                __syncwarp();
                pos += len;
            }
            if (br.bytes_used() > blk.clen + 8u) ok = false;  // ran past the payload
        }
        if (ok && (pos != isize || br.bytes_used() > blk.clen)) ok = false;
        if (!ok && lane == 0) atomicMax(ctl + 1, 0xFFFFFFFFu - b);  // largest complement = smallest index
        __syncwarp();
    }
}

}  // namespace bqc
