// swar_selftest.cpp -- CPU check of bamqc_b200/csrc/swar.h (test infrastructure; built and run by tests/test_swar_cpu.py).
// Random reads (A/C/G/T/N and IUPAC nibbles, qualities over the whole byte range, CIGARs with I/D/N/S/H/P ops and
// zero-length ops, reads that overhang the contig) go through
//   (1) a per-position walk that restates src/TripletCounting.hpp:195-236, and
//   (2) the run enumeration + 16-nibble windows used by k_stats (same statements as kernel_stats.cuh phase B),
// and the 1024 triplet counters must be equal.  Also checks the per-cycle word forms used by k_cycles.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>

#include "../bamqc_b200/csrc/swar.h"

using namespace bqc;

static uint64_t rng_state = 0x9E3779B97F4A7C15ULL;
static uint64_t rnd() {
    uint64_t z = (rng_state += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

struct Read {
    std::vector<uint8_t> seq;   // packed 4-bit, high nibble first, padded
    std::vector<uint8_t> qual;  // padded
    std::vector<uint32_t> cigar;
    uint32_t L;
    int32_t pos;
};

static uint64_t ld64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static uint32_t nib(const Read& r, uint32_t i) { return (r.seq[i >> 1] >> ((i & 1) ? 0 : 4)) & 15u; }
static uint32_t refbase(const std::vector<uint32_t>& ref, uint32_t refmax, uint32_t c) {
    uint32_t w = c >> 4; if (w > refmax) w = refmax;
    return (ref[w] >> (2 * (c & 15))) & 3u;
}

static void walk_naive(const Read& r, const std::vector<uint32_t>& ref, uint64_t reflen, uint32_t refmax, uint32_t* tri) {
    uint32_t it = 0, cc = (r.cigar[0] >> 4) - 1u, chromPos = (uint32_t)r.pos + 1u, readPos = 1;
    const uint32_t last = r.L - 1;
    for (; readPos < last; ++readPos, ++chromPos, --cc) {
        if (cc == 0) {
            bool ok = true;
            do {
                ++it;
                if (it >= r.cigar.size()) { ok = false; break; }
                uint32_t c = r.cigar[it], op = c & 15u, n = c >> 4;
                if (op == 2 || op == 3 || op == 5 || op == 6) chromPos += n;
                else if (op == 4 || op == 1) readPos += n;
                else cc = n;
            } while (cc == 0);
            if (!ok || readPos >= last) break;
        }
        const uint32_t q = r.qual[readPos];
        const uint32_t np = nib(r, readPos - 1), nc = nib(r, readPos), nn = nib(r, readPos + 1);
        const uint32_t rp = refbase(ref, refmax, chromPos - 1), rcur = refbase(ref, refmax, chromPos), rn = refbase(ref, refmax, chromPos + 1);
        const bool onehot = nc == 1 || nc == 2 || nc == 4 || nc == 8;
        const uint32_t base = nc == 1 ? 0 : nc == 2 ? 1 : nc == 4 ? 2 : 3;
        if ((int8_t)(q + 33u) >= (int8_t)53 && onehot && np == (1u << rp) && nn == (1u << rn) && (uint64_t)chromPos + 2 <= reflen)
            tri[((rp << 4) + (rcur << 2) + rn) * 16u + base]++;
    }
}

static void walk_swar(const Read& r, const std::vector<uint32_t>& ref, uint64_t reflen, uint32_t refmax, uint32_t* tri_s) {
    const uint8_t* seqp = r.seq.data();
    const uint8_t* qualp = r.qual.data();
    uint32_t it = 0, cc = (r.cigar[0] >> 4) - 1u, chromPos = (uint32_t)r.pos + 1u, readPos = 1;
    const uint32_t last = r.L - 1;
    bool ok = true;
    while (readPos < last) {
        if (cc == 0) {
            do {
                ++it;
                if (it >= r.cigar.size()) { ok = false; break; }
                uint32_t c = r.cigar[it], op = c & 15u, n = c >> 4;
                if (op == 2 || op == 3 || op == 5 || op == 6) chromPos += n;
                else if (op == 4 || op == 1) readPos += n;
                else cc = n;
            } while (cc == 0);
            if (!ok || readPos >= last) break;
        }
        const uint32_t run = cc < last - readPos ? cc : last - readPos;
        const int64_t lim = (int64_t)reflen - 1 - (int64_t)chromPos + (int64_t)readPos;
        uint32_t end = readPos + run;
        if (lim < (int64_t)end) end = lim > (int64_t)readPos ? (uint32_t)lim : readPos;
        const uint32_t delta = chromPos - readPos;
        for (uint32_t p = readPos; p < end;) {
            const uint32_t g = (p - 1u) & ~1u;
            const uint64_t R = swar_swap_nibbles(ld64(seqp + (g >> 1)));
            const uint64_t q0 = ld64(qualp + g), q1 = ld64(qualp + g + 8);
            const int32_t cg = (int32_t)(delta + g);
            const uint32_t cgc = cg < 0 ? 0u : (uint32_t)cg;
            const uint32_t wi = cgc >> 4;
            const uint64_t pair = (uint64_t)ref[wi < refmax ? wi : refmax] | ((uint64_t)ref[wi + 1u < refmax ? wi + 1u : refmax] << 32);
            uint32_t F = (uint32_t)(pair >> (2u * (cgc & 15u)));
            if (cg < 0) F <<= 2;
            uint64_t T4, C4;
            swar_triplet_masks(R, F, T4, C4);
            const uint32_t jlo = p - g, jhi = (end - g) < 15u ? (end - g) : 15u;
            C4 &= ((1ULL << (4u * jhi)) - 1ULL) & ~((1ULL << (4u * jlo)) - 1ULL);
            const uint32_t qk[4] = {swar_q20((uint32_t)q0), swar_q20((uint32_t)(q0 >> 32)), swar_q20((uint32_t)q1), swar_q20((uint32_t)(q1 >> 32))};
            const uint32_t c_lo = (uint32_t)C4, c_hi = (uint32_t)(C4 >> 32), t_lo = (uint32_t)T4, t_hi = (uint32_t)(T4 >> 32);
            for (uint32_t j = 1; j < 15; ++j) {
                const uint32_t cw = j < 8 ? c_lo : c_hi, tw = j < 8 ? t_lo : t_hi;
                const uint32_t cnt = (cw >> (4u * (j & 7u))) & (qk[j >> 2] >> (8u * (j & 3u) + 7u)) & 1u;
                const uint32_t ctx = (F >> (2u * (j - 1u))) & 63u;
                const uint32_t base = (tw >> (4u * (j & 7u))) & 3u;
                tri_s[ctx * 16u + base] += cnt;
            }
            p = g + jhi;
        }
        readPos += run;
        chromPos += run;
        cc -= run;
    }
}

static int test_triplets(int n_reads) {
    const uint64_t reflen = 5000;
    const uint32_t refmax = (uint32_t)((reflen + 15) / 16) + 3u;
    std::vector<uint32_t> ref(refmax + 2);
    for (auto& w : ref) w = (uint32_t)rnd();
    int bad = 0;
    for (int t = 0; t < n_reads; ++t) {
        Read r;
        r.L = 3 + (uint32_t)(rnd() % 260);
        if (rnd() % 4 == 0) r.L = 150;
        r.seq.assign(r.L / 2 + 80, 0);
        r.qual.assign(r.L + 80, 0);
        // position: mostly inside, sometimes overhanging the contig end, sometimes 0
        uint32_t m = (uint32_t)(rnd() % 10);
        r.pos = m == 0 ? 0 : m == 1 ? (int32_t)(reflen - rnd() % 200) : (int32_t)(rnd() % (reflen - 300));
        // CIGAR
        uint32_t nops = 1 + (rnd() % 3 == 0 ? (uint32_t)(rnd() % 5) : 0);
        uint32_t left = r.L;
        for (uint32_t i = 0; i < nops; ++i) {
            static const uint32_t ops[] = {0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8};
            uint32_t op = i == 0 && nops == 1 ? 0u : ops[rnd() % 11];
            uint32_t n = (rnd() % 8 == 0) ? 0u : (i + 1 == nops ? left : 1 + (uint32_t)(rnd() % (left ? left : 1)));
            if (op == 2 || op == 3 || op == 5 || op == 6) n = (uint32_t)(rnd() % 40);
            else if (n <= left) left -= n;
            r.cigar.push_back(n << 4 | op);
        }
        if (rnd() % 16 == 0) r.cigar[0] = (uint32_t)(rnd() % 3) << 4 | (uint32_t)(rnd() % 9);
        // bases: follow the reference along a plain diagonal (so flanks match often), with noise
        for (uint32_t i = 0; i < r.L; ++i) {
            uint32_t c = refbase(ref, refmax, (uint32_t)r.pos + i);
            uint32_t nb = 1u << c;
            uint32_t e = (uint32_t)(rnd() % 100);
            if (e < 3) nb = 1u << (rnd() % 4);
            else if (e < 5) nb = 15;
            else if (e < 6) nb = (uint32_t)(rnd() % 16);
            r.seq[i >> 1] |= (uint8_t)(nb << ((i & 1) ? 0 : 4));
            uint32_t qe = (uint32_t)(rnd() % 100);
            r.qual[i] = qe < 70 ? 37 : qe < 80 ? 19 : qe < 85 ? 20 : qe < 88 ? 94 : qe < 90 ? 95 : (uint8_t)(rnd() % 256);
        }
        for (size_t i = (r.L + 1) / 2; i < r.seq.size(); ++i) r.seq[i] = (uint8_t)rnd();  // bytes after SEQ are QUAL etc.
        for (size_t i = r.L; i < r.qual.size(); ++i) r.qual[i] = (uint8_t)rnd();
        if (r.L & 1) r.seq[r.L >> 1] = (uint8_t)((r.seq[r.L >> 1] & 0xF0) | (rnd() & 15));
        uint32_t a[1024] = {0}, b[1024] = {0}, bp[1024] = {0};
        walk_naive(r, ref, reflen, refmax, a);
        walk_swar(r, ref, reflen, refmax, b);
        for (uint32_t i = 0; i < 1024; ++i) bp[triplet_ctx_to_result(i >> 4) * 16u + (i & 15u)] += b[i];
        if (memcmp(a, bp, sizeof a) != 0) {
            if (bad < 5) {
                fprintf(stderr, "triplet mismatch: read %d L=%u pos=%d cigar:", t, r.L, r.pos);
                for (auto c : r.cigar) fprintf(stderr, " %u%c", c >> 4, "MIDNSHP=XXXXXXXX"[c & 15]);
                fprintf(stderr, "\n");
            }
            ++bad;
        }
    }
    return bad;
}

static uint32_t brev32(uint32_t x) { uint32_t r = 0; for (int i = 0; i < 32; ++i) r |= ((x >> i) & 1u) << (31 - i); return r; }

// the 8-base word forms of the per-cycle pass against the per-base tables of the reference semantics
// (A/a 0, C/c 1, G/g 2, T/t 3, else 4; reverse reads see the complemented base: SURVEY Appendix F, R5/R7)
static int test_cycles(int n) {
    const uint64_t LUT_FWD = 0x4444444344424104ULL, LUT_REV = 0x4444444044414234ULL;
    int bad = 0;
    for (int t = 0; t < n; ++t) {
        uint32_t W = 0;
        for (int j = 0; j < 8; ++j) {
            uint32_t e = (uint32_t)(rnd() % 10), nb = e < 7 ? 1u << (rnd() % 4) : e < 8 ? 15u : (uint32_t)(rnd() % 16);
            W |= nb << (4 * j);
        }
        uint32_t pop4, oh = swar_onehot8(W, pop4), D = swar_dna5_8(W, oh);
        uint32_t Wr = brev32(W), pop4r, ohr = swar_onehot8(Wr, pop4r), Dr = swar_dna5_8(Wr, ohr);
        uint32_t nN = 0, nGC = 0;
        for (int j = 0; j < 8; ++j) {
            uint32_t nb = (W >> (4 * j)) & 15u;
            if (((D >> (4 * j)) & 15u) != ((LUT_FWD >> (4 * nb)) & 7u)) ++bad;
            if (((Dr >> (4 * (7 - j))) & 15u) != ((LUT_REV >> (4 * nb)) & 7u)) ++bad;
            nN += nb == 15u;
            nGC += nb == 2u || nb == 4u;
        }
        if ((uint32_t)__builtin_popcount(pop4 & 0x44444444u) != nN || (uint32_t)__builtin_popcount(pop4r & 0x44444444u) != nN) ++bad;
        if ((uint32_t)__builtin_popcount(((W >> 1) | (W >> 2)) & oh) != nGC || (uint32_t)__builtin_popcount(((Wr >> 1) | (Wr >> 2)) & ohr) != nGC) ++bad;
        uint32_t x = (uint32_t)rnd();
        if (swar_swap_nibbles32(x) != (uint32_t)swar_swap_nibbles(x)) ++bad;
        uint32_t q = swar_q20(x);
        for (int j = 0; j < 4; ++j) {
            uint32_t b = (x >> (8 * j)) & 255u;
            if (((q >> (8 * j + 7)) & 1u) != (uint32_t)((int8_t)(b + 33u) >= (int8_t)53)) ++bad;
        }
    }
    return bad;
}

// the per-cycle pass of k_stats (kernel_stats.cuh phase C: eight cycles per step, reverse reads through an unaligned
// nibble window + brev, padded rows, dump row) against src/QualityCheck.hpp:111-176 done base by base
static uint32_t bperm_rev(uint32_t x) { return (x >> 24) | ((x >> 8) & 0xFF00u) | ((x << 8) & 0xFF0000u) | (x << 24); }
static int test_cycle_pass(int n_reads) {
    const uint64_t LUT_FWD = 0x4444444344424104ULL, LUT_REV = 0x4444444044414234ULL;
    int bad = 0;
    for (int t = 0; t < n_reads; ++t) {
        const uint32_t L = 1 + (uint32_t)(rnd() % 300), cycb = (L + 7u) & ~7u, rowp = cycb + (cycb >> 3) + 1u;
        const bool rc = rnd() & 1;
        std::vector<uint8_t> buf(64 + L / 2 + L + 64);
        for (auto& b : buf) b = (uint8_t)rnd();
        uint8_t* seqp = buf.data() + 64;
        uint8_t* qualp = seqp + (L + 1) / 2;
        for (uint32_t i = 0; i < L; ++i) {
            uint32_t e = (uint32_t)(rnd() % 10), nb = e < 7 ? 1u << (rnd() % 4) : e < 8 ? 15u : (uint32_t)(rnd() % 16);
            seqp[i >> 1] = (uint8_t)((i & 1) ? (seqp[i >> 1] & 0xF0) | nb : (seqp[i >> 1] & 0x0F) | (nb << 4));
        }
        std::vector<uint32_t> want(6 * cycb, 0), got(9 * rowp, 0);
        uint32_t wN = 0, wGC = 0, wQ = 0, cntN = 0, cntGC = 0, sumQ = 0;
        for (uint32_t i = 0; i < L; ++i) {
            const uint32_t nb = (seqp[i >> 1] >> ((i & 1) ? 0 : 4)) & 15u, q = qualp[i], cyc = rc ? L - 1 - i : i;
            want[(uint32_t)(((rc ? LUT_REV : LUT_FWD) >> (4 * nb)) & 7u) * cycb + cyc]++;
            want[5 * cycb + cyc] += q;
            wN += nb == 15u; wGC += nb == 2u || nb == 4u; wQ += q;
        }
        const uint32_t mysteps = (L + 7u) >> 3;
        for (uint32_t step = 0; step < mysteps; ++step) {
            const uint32_t v = 8u < L - 8u * step ? 8u : L - 8u * step;
            const int32_t n0 = rc ? (int32_t)L - 8 - (int32_t)(8u * step) : (int32_t)(8u * step);
            const uint64_t X = swar_swap_nibbles(ld64(seqp + (n0 >> 1)));
            uint32_t W = (uint32_t)(X >> (4u * (uint32_t)(n0 & 1)));
            const uint64_t Q = ld64(qualp + n0);
            uint32_t qlo = (uint32_t)Q, qhi = (uint32_t)(Q >> 32);
            if (rc) { W = brev32(W); const uint32_t tq = bperm_rev(qhi); qhi = bperm_rev(qlo); qlo = tq; }
            const uint32_t nm = v == 8u ? 0xFFFFFFFFu : (1u << (4u * v)) - 1u;
            const uint64_t qm = v == 8u ? ~0ULL : (1ULL << (8u * v)) - 1ULL;
            qlo &= (uint32_t)qm; qhi &= (uint32_t)(qm >> 32);
            uint32_t pop4;
            const uint32_t oh = swar_onehot8(W, pop4);
            const uint32_t D = (swar_dna5_8(W, oh) & nm) | (~nm & 0x88888888u);
            cntN += (uint32_t)__builtin_popcount(pop4 & 0x44444444u & nm);
            cntGC += (uint32_t)__builtin_popcount(((W >> 1) | (W >> 2)) & oh & nm);
            for (uint32_t j = 0; j < 8; ++j) {
                const uint32_t d = (D >> (4u * j)) & 15u, q = ((j < 4u ? qlo : qhi) >> (8u * (j & 3u))) & 255u;
                sumQ += q;
                got[d * rowp + 9u * step + j]++;
                got[5 * rowp + 9u * step + j] += q;
            }
        }
        bool ok = cntN == wN && cntGC == wGC && sumQ == wQ;
        for (uint32_t r = 0; r < 6 && ok; ++r)
            for (uint32_t c = 0; c < cycb; ++c)
                if (got[r * rowp + c + (c >> 3)] != want[r * cycb + c] + ((r < 5 && false) ? 1u : 0u)) { ok = false; break; }
        if (!ok) { if (bad < 5) fprintf(stderr, "cycle pass mismatch: L=%u rc=%d\n", L, (int)rc); ++bad; }
    }
    return bad;
}

// the 16-base steps of k_eightmer (kernels.cuh) against OverallNumbers::count8mers (src/OverallNumbers.hpp:137-168)
// done base by base on the read-oriented sequence
static int test_eightmers(int n_reads) {
    int bad = 0;
    for (int t = 0; t < n_reads; ++t) {
        const uint32_t Ls = 8 + (uint32_t)(rnd() % 200);
        const bool rc = rnd() & 1;
        std::vector<uint8_t> seq(Ls / 2 + 40);
        for (auto& b : seq) b = (uint8_t)rnd();
        std::vector<uint32_t> nibs(Ls);
        for (uint32_t i = 0; i < Ls; ++i) {
            uint32_t e = (uint32_t)(rnd() % 40), nb = e < 36 ? 1u << (rnd() % 4) : e < 38 ? 15u : (uint32_t)(rnd() % 16);
            nibs[i] = nb;
            seq[i >> 1] = (uint8_t)((i & 1) ? (seq[i >> 1] & 0xF0) | nb : (seq[i >> 1] & 0x0F) | (nb << 4));
        }
        std::vector<uint32_t> want, got;
        {   // read-oriented bases: reverse reads are reverse-complemented first (nibble bit reversal)
            std::vector<uint32_t> o(Ls);
            for (uint32_t i = 0; i < Ls; ++i) {
                uint32_t nb = nibs[rc ? Ls - 1 - i : i];
                if (rc) nb = ((nb & 1) << 3) | ((nb & 2) << 1) | ((nb & 4) >> 1) | ((nb & 8) >> 3);
                o[i] = nb;
            }
            for (uint32_t w = 0; w + 8 <= Ls; ++w) {
                bool hasn = false; uint32_t code = 0;
                for (uint32_t i = 0; i < 8; ++i) {
                    const uint32_t nb = o[w + i];
                    hasn |= nb == 15u;
                    code = code * 4 + (nb == 2 ? 1 : nb == 4 ? 2 : nb == 8 ? 3 : 0);
                }
                if (!hasn) want.push_back(code);
            }
        }
        uint32_t prev = 0, since_n = 0;
        const uint8_t* seqp = seq.data();
        const uint32_t nch = (Ls + 15u) >> 4;
        for (uint32_t c = 0; c < nch; ++c) {
            const uint64_t R = swar_swap_nibbles(ld64(seqp + 8u * c));
            const uint32_t rem = Ls - 16u * c;
            const uint64_t inr = rem >= 16u ? ~0ULL : (1ULL << (4u * rem)) - 1ULL;
            const uint64_t s2 = (R & 0x5555555555555555ULL) + ((R >> 1) & 0x5555555555555555ULL);
            const uint64_t tt = (s2 & 0x3333333333333333ULL) + ((s2 >> 2) & 0x3333333333333333ULL);
            const uint64_t u = tt ^ kNib1;
            const uint64_t oh = ~(u | (u >> 1) | (u >> 2)) & kNib1;
            const uint64_t isn = (tt >> 2) & kNib1 & inr;
            uint64_t code = swar_code4(R);
            if (rc) code ^= 0x3333333333333333ULL;
            code &= oh * 3ULL;
            uint64_t x = (code | (code >> 2)) & 0x0F0F0F0F0F0F0F0FULL;
            x = (x | (x >> 4)) & 0x00FF00FF00FF00FFULL;
            x = (x | (x >> 8)) & 0x0000FFFF0000FFFFULL;
            const uint32_t ple = (uint32_t)x | ((uint32_t)(x >> 32) << 16);
            uint32_t lo, hi, cur;
            if (rc) { cur = ple; lo = (prev >> 18) | (cur << 14); hi = cur >> 18; }
            else { const uint32_t br = brev32(ple); cur = ((br & 0x55555555u) << 1) | ((br >> 1) & 0x55555555u); lo = cur; hi = prev; }
            uint64_t badm = isn;
            badm |= badm << 4; badm |= badm << 8; badm |= badm << 16;
            if (since_n < 7u) badm |= (1ULL << (4u * (7u - since_n))) - 1ULL;
            badm |= ~inr;
            since_n = isn ? (uint32_t)__builtin_clzll(isn) >> 2 : (since_n + 16u < 64u ? since_n + 16u : 64u);
            for (uint32_t e = 0; e < 16u; ++e) {
                const uint32_t sh = rc ? 2u * e : 30u - 2u * e;
                const uint32_t code16 = (uint32_t)((((uint64_t)hi << 32) | lo) >> sh) & 0xFFFFu;
                if (!((badm >> (4u * e)) & 1u)) got.push_back(code16);
            }
            prev = cur;
        }
        // forward reads emit windows in read order; reverse reads emit the read-oriented windows in reverse order
        if (rc) std::reverse(got.begin(), got.end());
        if (got != want) { if (bad < 5) fprintf(stderr, "8-mer mismatch: L=%u rc=%d got %zu want %zu\n", Ls, (int)rc, got.size(), want.size()); ++bad; }
    }
    return bad;
}

int main(int argc, char** argv) {
    int n = argc > 1 ? atoi(argv[1]) : 200000;
    int bad = test_triplets(n);
    printf("triplets: %d reads, %d mismatching\n", n, bad);
    int bad2 = test_cycles(n);
    printf("cycle words: %d words, %d mismatching\n", n, bad2);
    int bad3 = test_cycle_pass(n / 4);
    printf("cycle pass: %d reads, %d mismatching\n", n / 4, bad3);
    int bad4 = test_eightmers(n / 4);
    printf("8-mers: %d reads, %d mismatching\n", n / 4, bad4);
    return bad || bad2 || bad3 || bad4 ? 1 : 0;
}
