// kernel_frame.cuh -- device-side record framing of an inflated BAM byte stream (what SeqAn's readRecord does
// one record at a time at src/bamqualcheck.cpp:306: block_size -> next record) plus the per-record pre-pass of
// the coverage statistic (which records take part in OverallNumbers::coverage, src/bamqualcheck.cpp:430-433).
//
// The record chain p -> p + 4 + block_size is a dependent load per record, so it is cut into windows that are
// framed concurrently from SPECULATED starts and then VERIFIED:
//   k_frame_speculate  one thread per 4 KiB window: the first position whose next six hops all look like BAM
//                      records (same plausibility test as the host framer) is taken as the window's start; the
//                      chain is walked to the first record that starts beyond the window (count, exit);
//   k_frame_relax      a few rounds in which a window that its predecessor's chain enters somewhere else than at
//                      the speculated start is re-walked from there (repairs false starts and missed starts);
//   k_frame_verify     the result is accepted only if every window's start is exactly the exit of the
//                      previous window that holds a start and the first window starts at the true stream
//                      position -- by induction the accepted chain is the sequential one.  Anything else sets
//                      `bad` and k_frame_repair re-frames the buffer sequentially (exact either way).
//                      Also the exclusive scan of the per-block record counts;
//   k_frame_emit       walks each window again and writes the dense offsets, the (rid, pos) pairs of the records
//                      that take part in the coverage statistic and the longest read.
//   k_frame_tail       moves the partial record at the end of a buffer in front of the next buffer
//                      (stream submissions need not end on record boundaries).
#pragma once

namespace bqc {

static const uint32_t kFrameWindow = 4096;    // bytes per speculation window
static const uint32_t kFrameThreads = 128;    // windows per CTA
static const uint32_t kFrameHops = 6;
static const uint32_t kFrameHead = 1u << 20;  // room in front of a device buffer for the carried partial record

struct FrameResult {
    uint32_t n_records;
    uint32_t max_lseq;
    uint32_t bad;        // 0 = verified; otherwise 1 + index of the first inconsistent window
    uint32_t start;      // stream position of the first record of this buffer (kFrameHead - carried bytes)
    uint32_t end;        // end of the last whole record
    uint32_t total;      // end of the data in the buffer
    uint32_t tail_overflow;  // 1: the partial record at the end does not fit the head room of the next buffer; 2: more records than the offset arrays hold
    uint32_t skip_emit;  // k_frame_emit has nothing to do (repaired, or overflow)
    uint32_t repaired;   // the speculation failed verification (1 + first inconsistent window) and k_frame_repair framed the buffer
    uint32_t inflate_bad; // device inflate: 1 + index of the first BGZF block of this buffer that failed (0 = none)
    uint32_t pad[2];
};

struct FrameMeta { int32_t rid; uint32_t pos; };  // rid < 0: the record does not take part in the coverage statistic

// Plausibility of "a record starts at p" (SAM/BAM spec field ranges); speculation only.  The caller has checked
// p + 36 <= n.  Returns block_size, or 0 if the header is implausible.
__device__ __forceinline__ uint32_t frame_plausible(const uint8_t* d, uint64_t p, int32_t n_ref) {
    const uint32_t bs = ldu32(d + p);
    if (bs < 34u || bs > (1u << 28)) return 0;
    const int32_t rid = (int32_t)ldu32(d + p + 4);
    if (rid < -1 || rid >= n_ref) return 0;
    const int32_t pos = (int32_t)ldu32(d + p + 8), nrid = (int32_t)ldu32(d + p + 24), npos = (int32_t)ldu32(d + p + 28);
    if (nrid < -1 || nrid >= n_ref || pos < -1 || npos < -1) return 0;
    const uint32_t x = ldu32(d + p + 12), y = ldu32(d + p + 16);
    const uint32_t lname = x & 255u, ncig = y & 0xFFFFu;
    const int32_t lseq = (int32_t)ldu32(d + p + 20);
    if (lname < 1u || lseq < 0) return 0;
    const uint64_t need = 32ull + lname + 4ull * ncig + ((uint64_t)lseq + 1) / 2 + (uint64_t)lseq;
    if (need > bs) return 0;
    return bs;
}
// "six records in a row start at p" -- or fewer when the chain runs into the end of the data (a partial record at
// the end of a stream buffer is normal)
__device__ __forceinline__ bool frame_chain_plausible(const uint8_t* d, uint64_t n, uint64_t p, int32_t n_ref) {
    uint64_t q = p;
    for (uint32_t hops = 0; hops < kFrameHops; ++hops) {
        if (q + 36 > n) return hops >= 1;
        const uint32_t bs = frame_plausible(d, q, n_ref);
        if (!bs) return false;
        if (q + 4 + (uint64_t)bs > n) return hops >= 1;
        if (q + 36 + (ldu32(d + q + 12) & 255u) <= n && ldg8(d + q + 36 + (ldu32(d + q + 12) & 255u) - 1) != 0) return false;  // read_name is NUL terminated
        q += 4 + (uint64_t)bs;
    }
    return true;
}

// the records that start in window [lo, hi): count and the position of the first record beyond the window
__device__ __forceinline__ void frame_walk(const uint8_t* d, uint64_t n, uint64_t s, uint64_t hi, uint32_t& count, uint32_t& exit_pos) {
    uint64_t p = s;
    uint32_t c = 0;
    while (p < hi && p + 4 <= n) {  // a record belongs to the window it starts in
        const uint32_t bs = ldu32(d + p);
        if (bs < 32u || p + 4 + (uint64_t)bs > n) break;
        ++c;
        p += 4 + (uint64_t)bs;
    }
    count = c;
    exit_pos = (uint32_t)p;
}

// A stream that begins somewhere inside a record (a later piece of a file cut at BGZF block boundaries): the first
// position from which six records in a row look plausible becomes the stream position.  The caller verifies the guess
// from the other side: the piece before it must end exactly there (include/bamqc_b200.h, bqc_stream_unknown_start).
static const uint32_t kFrameSeekSpan = 1u << 20;   // a record is smaller than the head room (1 MiB)
__global__ void __launch_bounds__(256) k_frame_seek(const uint8_t* __restrict__ d, FrameResult* fr, int32_t n_ref) {
    __shared__ uint32_t s_best;
    const uint64_t start = fr->start, n = fr->total;
    if (threadIdx.x == 0) s_best = kNone;
    __syncthreads();
    for (uint64_t base = start; base < n && base < start + kFrameSeekSpan; base += blockDim.x) {
        const uint64_t p = base + threadIdx.x;
        if (p + 36 <= n && frame_chain_plausible(d, n, p, n_ref)) atomicMin(&s_best, (uint32_t)p);
        __syncthreads();
        if (s_best != kNone) break;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (s_best != kNone) fr->start = s_best;
        else { fr->start = (uint32_t)n; fr->tail_overflow = 1u; }   // no record boundary in sight: reported as an unreadable stream
    }
}

__global__ void __launch_bounds__(kFrameThreads) k_frame_speculate(const uint8_t* __restrict__ d, const FrameResult* __restrict__ fr, int32_t n_ref, uint32_t nwin,
                                                                   uint32_t* __restrict__ ws, uint32_t* __restrict__ we, uint32_t* __restrict__ wc) {
    const uint32_t w = blockIdx.x * kFrameThreads + threadIdx.x;
    if (w >= nwin) return;
    const uint64_t start = fr->start, n = fr->total;
    const uint64_t lo = (uint64_t)w * kFrameWindow, hi = min(n, lo + kFrameWindow);
    uint64_t s = ~0ull;
    if (start >= lo && start < lo + kFrameWindow) {
        if (start < n) s = start;  // the true position of the stream
    } else if (lo > start) {
        for (uint64_t p = lo; p < hi; ++p)
            if (frame_chain_plausible(d, n, p, n_ref)) { s = p; break; }
    }
    uint32_t c = 0, e = kNone;
    if (s != ~0ull) frame_walk(d, n, s, hi, c, e);
    ws[w] = (uint32_t)s;  // ~0 -> kNone
    we[w] = e;
    wc[w] = c;
}

// One relaxation round over the windows (Jacobi: reads the previous state, writes the next).  A window whose
// nearest predecessor with a start exits INTO this window at a position other than the window's own start is
// re-entered there and walked again.  This repairs the two failure modes of the speculation -- a false start
// (with records of one size a wrong block_size lands on a record boundary once in ~300 tries) and a window whose
// true first record did not look plausible -- in as many rounds as there are adjacent bad windows.
__global__ void __launch_bounds__(kFrameThreads) k_frame_relax(const uint8_t* __restrict__ d, const FrameResult* __restrict__ fr, uint32_t nwin,
                                                               const uint32_t* __restrict__ ws, const uint32_t* __restrict__ we, const uint32_t* __restrict__ wc,
                                                               uint32_t* __restrict__ ws2, uint32_t* __restrict__ we2, uint32_t* __restrict__ wc2) {
    const uint32_t w = blockIdx.x * kFrameThreads + threadIdx.x;
    if (w >= nwin) return;
    const uint32_t start = fr->start;
    const uint64_t n = fr->total;
    const uint32_t w0 = start / kFrameWindow;
    uint32_t s = ws[w], e = we[w], c = wc[w];
    if (w > w0) {
        uint32_t u = w - 1;
        while (u > w0 && ws[u] == kNone) --u;
        if (ws[u] != kNone) {
            const uint32_t x = we[u];
            if (x / kFrameWindow == w && x != s && x < n) {
                s = x;
                frame_walk(d, n, s, min(n, (uint64_t)(w + 1) * kFrameWindow), c, e);
            }
        }
    }
    ws2[w] = s;
    we2[w] = e;
    wc2[w] = c;
}

__global__ void __launch_bounds__(kFrameThreads) k_frame_blocksum(uint32_t nwin, const uint32_t* __restrict__ wc, uint32_t* __restrict__ block_sum) {
    __shared__ uint32_t red[kFrameThreads / 32];
    const uint32_t w = blockIdx.x * kFrameThreads + threadIdx.x;
    uint32_t t = w < nwin ? wc[w] : 0u;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xFFFFFFFFu, t, o);
    if ((threadIdx.x & 31u) == 0) red[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t s = 0;
        for (uint32_t i = 0; i < kFrameThreads / 32; ++i) s += red[i];
        block_sum[blockIdx.x] = s;
    }
}

// one CTA of 1024 threads
__global__ void __launch_bounds__(1024) k_frame_verify(FrameResult* fr, uint32_t nwin, uint32_t nblk, const uint32_t* __restrict__ ws, const uint32_t* __restrict__ we,
                                                       const uint32_t* __restrict__ block_sum, uint32_t* __restrict__ block_base, uint32_t* __restrict__ offsets, uint32_t rec_cap,
                                                       uint32_t force_bad, const uint32_t* __restrict__ inflate_ctl) {
    __shared__ uint32_t s_bad, s_end, s_carry;
    __shared__ uint32_t wsum[32];
    if (threadIdx.x == 0) { s_bad = 0xFFFFFFFFu; s_end = 0; s_carry = 0; }
    __syncthreads();
    const uint32_t start = fr->start, total = fr->total;
    const uint32_t w0 = start / kFrameWindow;
    uint32_t my_bad = 0xFFFFFFFFu, my_end = 0;
    for (uint32_t w = threadIdx.x; w < nwin; w += blockDim.x) {
        const uint32_t s = ws[w];
        if (s == kNone) continue;
        my_end = max(my_end, we[w]);
        if (w <= w0) {
            if (w < w0 || s != start) my_bad = min(my_bad, w);
            continue;
        }
        uint32_t u = w - 1;
        while (u > w0 && ws[u] == kNone) --u;
        if (ws[u] == kNone || we[u] != s) my_bad = min(my_bad, w);
    }
    if (my_bad != 0xFFFFFFFFu) atomicMin(&s_bad, my_bad);
    if (my_end) atomicMax(&s_end, my_end);
    // exclusive scan of the per-block record counts
    for (uint32_t b0 = 0; b0 < nblk; b0 += blockDim.x) {
        const uint32_t b = b0 + threadIdx.x;
        const uint32_t v = b < nblk ? block_sum[b] : 0u;
        uint32_t incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if ((threadIdx.x & 31u) >= (uint32_t)o) incl += t;
        }
        if ((threadIdx.x & 31u) == 31u) wsum[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t x = wsum[threadIdx.x], xi = x;
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xFFFFFFFFu, xi, o);
                if (threadIdx.x >= (uint32_t)o) xi += t;
            }
            wsum[threadIdx.x] = xi - x;
        }
        __syncthreads();
        const uint32_t excl = s_carry + wsum[threadIdx.x >> 5] + incl - v;
        if (b < nblk) block_base[b] = excl;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const uint32_t end = s_end ? s_end : min(start, total);  // no record at all: everything is tail
        const bool bad = s_bad != 0xFFFFFFFFu || force_bad;
        fr->inflate_bad = (inflate_ctl && inflate_ctl[1]) ? (0xFFFFFFFFu - inflate_ctl[1]) + 1u : 0u;
        fr->n_records = s_carry;
        fr->bad = bad ? (s_bad == 0xFFFFFFFFu ? 1u : s_bad + 1u) : 0u;
        fr->end = end;
        fr->tail_overflow = (total - end > kFrameHead) ? 1u : 0u;
        if (!bad && s_carry > rec_cap) { fr->tail_overflow = 2u; fr->n_records = 0; fr->skip_emit = 1u; }
        else if (!bad) offsets[s_carry] = end;
    }
}

__device__ __forceinline__ FrameMeta frame_meta_of(const uint8_t* d, uint64_t p, int32_t n_ref, const uint8_t* __restrict__ main_chrom, uint32_t& mx) {
    const int32_t rid = (int32_t)ldu32(d + p + 4);
    const uint32_t pos = ldu32(d + p + 8);
    const uint32_t flag = ldu32(d + p + 16) >> 16;
    const int32_t lseq = (int32_t)ldu32(d + p + 20);
    if (lseq > 0) mx = max(mx, (uint32_t)lseq);
    // the records OverallNumbers::coverage sees: src/bamqualcheck.cpp:318-335 (primary), :392-433 (main
    // chromosome, mapped, not duplicate), with a first/second flag (:385-389 is fatal otherwise)
    const bool q = !(flag & 0x900u) && (flag & 0xC0u) && !(flag & 0x4u) && !(flag & 0x400u) && rid >= 0 && rid < n_ref && main_chrom[rid];
    FrameMeta m;
    m.rid = q ? rid : -1;
    m.pos = q ? pos : 0u;
    return m;
}

// The verification failed (never seen on real data; a corrupt or adversarial stream can do it): one thread walks
// the true chain from the stream position.  Slow (a dependent load per record) but exact, and everything that
// follows on the stream sees the same FrameResult as after a verified speculation.
__global__ void k_frame_repair(const uint8_t* __restrict__ d, FrameResult* fr, int32_t n_ref, const uint8_t* __restrict__ main_chrom, uint32_t* __restrict__ offsets,
                               FrameMeta* __restrict__ meta, uint32_t rec_cap) {
    if (!fr->bad || threadIdx.x || blockIdx.x) return;
    const uint64_t n = fr->total;
    uint64_t p = fr->start;
    uint32_t cnt = 0, mx = 0;
    bool over = false;
    while (p + 4 <= n) {
        const uint32_t bs = ldu32(d + p);
        if (bs < 32u || p + 4 + (uint64_t)bs > n) break;
        if (cnt >= rec_cap) { over = true; break; }
        offsets[cnt] = (uint32_t)p;
        { const FrameMeta m = frame_meta_of(d, p, n_ref, main_chrom, mx); if (meta) meta[cnt] = m; }
        ++cnt;
        p += 4 + (uint64_t)bs;
    }
    fr->repaired = fr->bad;
    fr->bad = 0;
    fr->skip_emit = 1u;
    fr->end = (uint32_t)p;
    fr->max_lseq = mx;
    fr->tail_overflow = over ? 2u : ((n - p > kFrameHead) ? 1u : 0u);
    fr->n_records = over ? 0u : cnt;
    if (!over) offsets[cnt] = (uint32_t)p;
}

__global__ void __launch_bounds__(kFrameThreads) k_frame_emit(const uint8_t* __restrict__ d, FrameResult* fr, int32_t n_ref, const uint8_t* __restrict__ main_chrom, uint32_t nwin,
                                                              const uint32_t* __restrict__ ws, const uint32_t* __restrict__ wc, const uint32_t* __restrict__ block_base,
                                                              uint32_t* __restrict__ offsets, FrameMeta* __restrict__ meta) {
    __shared__ uint32_t wsum[kFrameThreads / 32];
    __shared__ uint32_t s_max;
    if (fr->skip_emit) return;  // uniform: repaired, or more records than the offset arrays hold
    if (threadIdx.x == 0) s_max = 0;
    const uint32_t w = blockIdx.x * kFrameThreads + threadIdx.x;
    const uint32_t c = w < nwin ? wc[w] : 0u;
    uint32_t incl = c;
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if ((threadIdx.x & 31u) >= (uint32_t)o) incl += t;
    }
    if ((threadIdx.x & 31u) == 31u) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t base = block_base[blockIdx.x] + incl - c;
    for (uint32_t i = 0; i < (threadIdx.x >> 5); ++i) base += wsum[i];
    uint32_t mx = 0;
    if (c) {
        uint64_t p = ws[w];
        for (uint32_t i = 0; i < c; ++i) {
            const uint32_t bs = ldu32(d + p);
            offsets[base + i] = (uint32_t)p;
            { const FrameMeta m = frame_meta_of(d, p, n_ref, main_chrom, mx); if (meta) meta[base + i] = m; }
            p += 4 + (uint64_t)bs;
        }
    }
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
    if ((threadIdx.x & 31u) == 0 && mx) atomicMax(&s_max, mx);
    __syncthreads();
    if (threadIdx.x == 0 && s_max) atomicMax(&fr->max_lseq, s_max);
}

// Lane (read group) of every framed record: the value of the first RG tag is looked up in the table of @RG IDs
// (getLane, src/bamqualcheck.cpp:72-100; a value that is not in the header maps to lane 0, like the reference's
// laneNames[...] which default-inserts 0; a record without RG:Z keeps lane 0 and k_stats reports it).  names: the lane
// ids back to back, name_off[l] .. name_off[l + 1] delimits id l; later duplicates win like the map assignment.
__global__ void __launch_bounds__(256) k_frame_lanes(const uint8_t* __restrict__ d, const FrameResult* __restrict__ fr, const uint32_t* __restrict__ offsets,
                                                     const char* __restrict__ names, const uint32_t* __restrict__ name_off, uint32_t n_lanes, uint8_t* __restrict__ lane_out) {
    const uint32_t n = fr->n_records;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        const uint32_t off = offsets[r], end = offsets[r + 1];
        const uint8_t* p = d + off;
        uint32_t lane = 0;
        if (end - off >= 36u) {
            const uint32_t lname = ldu32(p + 12) & 255u, ncig = ldu32(p + 16) & 0xFFFFu;
            const int32_t lseq = (int32_t)ldu32(p + 20);
            const uint32_t ls = lseq > 0 ? (uint32_t)lseq : 0u;
            uint64_t q = 36ull + lname + 4ull * ncig + (ls + 1u) / 2u + ls;
            const uint64_t avail = end - off;
            while (q + 3 <= avail) {
                const uint32_t k0 = ldg8(p + q), k1 = ldg8(p + q + 1), ty = ldg8(p + q + 2);
                q += 3;
                uint64_t sz;
                if (ty == 'A' || ty == 'c' || ty == 'C') sz = 1;
                else if (ty == 's' || ty == 'S') sz = 2;
                else if (ty == 'i' || ty == 'I' || ty == 'f') sz = 4;
                else if (ty == 'Z' || ty == 'H') { uint64_t z = q; while (z < avail && ldg8(p + z) != 0) ++z; sz = z - q + 1; }
                else if (ty == 'B') {
                    if (q + 5 > avail) break;
                    const uint32_t sub = ldg8(p + q), cnt = ldu32(p + q + 1);
                    sz = 5ull + (uint64_t)cnt * ((sub == 'c' || sub == 'C') ? 1u : (sub == 's' || sub == 'S') ? 2u : 4u);
                } else break;
                if (k0 == 'R' && k1 == 'G') {
                    if (ty == 'Z') {
                        const uint32_t vlen = (uint32_t)(sz - 1);   // without the NUL
                        for (uint32_t l = 0; l < n_lanes; ++l) {
                            const uint32_t a = name_off[l], m = name_off[l + 1] - a;
                            if (m != vlen) continue;
                            uint32_t i = 0;
                            while (i < m && (uint32_t)(uint8_t)names[a + i] == ldg8(p + q + i)) ++i;
                            if (i == m) lane = l;
                        }
                    }
                    break;   // the first RG tag decides (:86)
                }
                q += sz;
            }
        }
        lane_out[r] = (uint8_t)lane;
    }
}

// Several read groups: the record indices of one lane, in file order (stable compaction, decoupled look-back), so that
// the lane's pass of every table kernel touches its own records only.  Launched once per lane in lane order:
// range[lane] is where the lane's list starts in `index`, range[lane + 1] is written by the last tile.
// n_ptr: the record count on the device (FrameResult::n_records of a device-framed buffer) or NULL (then n).
__global__ void __launch_bounds__(1024) k_lane_partition(const uint8_t* __restrict__ rec_lane, const uint32_t* __restrict__ n_ptr, uint32_t n, uint32_t lane,
                                                         uint32_t* __restrict__ index, uint32_t* range, unsigned long long* lb, uint32_t* ticket) {
    __shared__ uint32_t ws[33];
    __shared__ uint32_t s_tile;
    __shared__ unsigned long long s_base;
    if (n_ptr) n = *n_ptr;
    const uint32_t ntile = (n + 1023u) / 1024u;
    const uint32_t start = lane ? range[lane] : 0u;
    if (ntile == 0) { if (blockIdx.x == 0 && threadIdx.x == 0) { if (!lane) range[0] = 0; range[lane + 1] = start; } return; }
    for (;;) {
        if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= ntile) break;
        const uint32_t r = tile * 1024u + threadIdx.x;
        const uint32_t f = (r < n && rec_lane[r] == lane) ? 1u : 0u;
        uint32_t total;
        const uint32_t excl = cov_block_scan(f, ws, total) - f;
        if (threadIdx.x < 32) {
            const unsigned long long pre = cov_lookback(lb, tile, total, (unsigned long long)start);
            if (threadIdx.x == 0) {
                s_base = pre;
                if (tile == ntile - 1u) { if (!lane) range[0] = 0; range[lane + 1] = (uint32_t)(pre + total); }
            }
        }
        __syncthreads();
        if (f) index[(uint32_t)s_base + excl] = r;
        __syncthreads();
    }
}

// Prepare the frame header of the next buffer: carry the partial record [end, total) of the previous buffer (if
// any) in front of kFrameHead and set start/total.  One CTA.
// `skip`: bytes at the front of the new data that are not records (the BAM header in the first buffer of a file).
__global__ void __launch_bounds__(256) k_frame_tail(const uint8_t* __restrict__ prev_bytes, const FrameResult* __restrict__ prev, uint8_t* __restrict__ bytes, FrameResult* fr, uint32_t n_new,
                                                    uint32_t skip) {
    uint32_t tail = 0;
    if (prev) {
        tail = prev->total - prev->end;
        if (prev->tail_overflow) tail = 0;  // fatal for the run; the host reports it
    }
    const uint8_t* src = prev_bytes + (prev ? prev->end : 0u);
    uint8_t* dst = bytes + kFrameHead - tail;
    for (uint32_t i = threadIdx.x; i < tail; i += blockDim.x) dst[i] = src[i];
    if (threadIdx.x == 0) {
        fr->n_records = 0;
        fr->max_lseq = 0;
        fr->bad = 0;
        fr->start = kFrameHead - tail + skip;
        fr->end = 0;
        fr->total = kFrameHead + n_new;
        fr->tail_overflow = 0;
        fr->skip_emit = 0;
        fr->repaired = 0;
        fr->inflate_bad = 0;
    }
}

}  // namespace bqc
