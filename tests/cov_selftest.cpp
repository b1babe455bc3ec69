// cov_selftest.cpp -- CPU check of the parallel form of the coverage statistic (bamqc_b200/csrc/cov_math.h and the
// block decomposition of kernel_cov.cuh) against a sequential restatement of OverallNumbers::coverage /
// update_coverage / update_vectors (/root/reference/src/OverallNumbers.hpp:59-135) and the end-of-run flush
// (/root/reference/src/bamqualcheck.cpp:447-453).  Run by tests/test_cov_cpu.py.  Test infrastructure only.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../bamqc_b200/csrc/cov_math.h"

using namespace bqc;

struct Rec { int32_t rid; uint32_t b; uint32_t c0, len; };  // covered interval [c0, c0 + len) relative to b

// ---- sequential restatement (src/OverallNumbers.hpp:79-135) ---------------------------------------------------
struct SeqModel {
    bool first = true;
    int id = 0;
    int shift = 0;
    std::vector<unsigned> v1, v2;
    std::vector<uint64_t> poscov;
    std::vector<uint32_t> pstate;  // b - shift after each record (for the state check)
    SeqModel() : v1(1000, 0), v2(1000, 0), poscov(101, 0) {}
    void update_coverage() { for (unsigned i = 0; i < 1000; ++i) poscov[v1[i] > 100 ? 100 : v1[i]] += 1; }
    void update_vectors() { v1.assign(1000, 0); std::swap(v1, v2); }
    void coverage(const Rec& r) {
        unsigned beginpos = r.b;
        if (first) { first = false; id = r.rid; shift = (int)beginpos; }
        if (id != r.rid || ((beginpos - (unsigned)shift) > 2000u)) {
            id = r.rid;
            update_coverage(); update_vectors(); update_coverage();
            v1.assign(1000, 0);
            shift = (int)beginpos;
        }
        unsigned pos = beginpos - (unsigned)shift;
        if (pos > 1000 && pos < 2000) { update_coverage(); update_vectors(); shift += 1000; pos = beginpos - (unsigned)shift; }
        pstate.push_back(pos);
        for (unsigned j = r.c0; j < r.c0 + r.len; ++j) {
            if (pos + j < 1000) v1[pos + j] += 1;
            else if (pos - 1000 + j < 1000) v2[pos - 1000 + j] += 1;  // beyond: the reference writes out of bounds (lost, R9)
        }
    }
    void finish() { update_coverage(); update_vectors(); update_coverage(); }
};

// ---- parallel form: blocks of RB records, candidates, closed-form stretches, all-states tables ------------------
struct Carry { bool first = true; int32_t rid_prev = 0; uint32_t b_prev = 0, p_prev = 0; uint64_t xc = 0; std::vector<int64_t> D = std::vector<int64_t>(2048, 0); };

struct Block {
    uint32_t n = 0, nc = 0;
    bool isfirst = false;
    std::vector<uint32_t> sb; std::vector<int32_t> srid;
    std::vector<uint32_t> cg, cGs, cGj, cord, cpos, cjs; std::vector<uint8_t> cdef;
};
static void analyse(Block& K, const std::vector<Rec>& q, uint32_t r0, uint32_t n, const Carry& cy, bool blk0) {
    K.n = n; K.isfirst = blk0 && cy.first;
    K.sb.assign(n + 1, 0); K.srid.assign(n + 1, 0);
    for (uint32_t i = 0; i < n; ++i) { K.sb[i + 1] = q[r0 + i].b; K.srid[i + 1] = q[r0 + i].rid; }
    if (!blk0) { K.sb[0] = q[r0 - 1].b; K.srid[0] = q[r0 - 1].rid; }
    else if (K.isfirst) { K.sb[0] = q[0].b; K.srid[0] = q[0].rid; }
    else { K.sb[0] = cy.b_prev; K.srid[0] = cy.rid_prev; }
    K.cg.assign(n + 2, 0); K.cGs.assign(n + 2, 0); K.cGj.assign(n + 2, 0); K.cord.assign(n + 2, 0); K.cpos.assign(n + 2, 0); K.cjs.assign(n + 2, 0); K.cdef.assign(n + 2, 0);
    uint32_t ord = 0;
    for (uint32_t i = 1; i <= n; ++i) {
        bool other = K.srid[i] != K.srid[i - 1] || (K.isfirst && i == 1);
        if (other || (uint32_t)(K.sb[i] - K.sb[i - 1]) >= kCovV) { ++ord; K.cpos[ord] = i; K.cg[ord] = K.sb[i] - K.sb[i - 1]; K.cdef[ord] = other; }
        K.cord[i] = ord;
    }
    K.nc = ord;
    K.cpos[ord + 1] = n + 1;
    for (uint32_t i = 1; i <= n; ++i) {
        uint32_t c = K.cord[i];
        if (K.cpos[c] != i && K.sb[i] != K.sb[i - 1] && K.sb[i - 1] == K.sb[K.cpos[c]]) K.cjs[c] = i;
    }
    for (uint32_t c = 0; c <= K.nc; ++c) {
        uint32_t last = K.cpos[c + 1] - 1;
        K.cGs[c] = K.sb[last] - K.sb[K.cpos[c]];
        K.cGj[c] = K.cjs[c] ? K.sb[last] - K.sb[K.cjs[c]] : 0;
    }
}
static uint32_t apply(const Block& K, uint32_t c, uint32_t p) {
    uint32_t q = p;
    if (c) { bool reset; q = cov_step(p, K.cg[c], K.cdef[c] != 0, reset); }
    return cov_stretch(q, K.cGs[c], K.cGj[c]);
}

struct ParModel {
    Carry cy;
    std::vector<uint64_t> poscov = std::vector<uint64_t>(101, 0);
    std::vector<uint32_t> pstate;
    uint32_t RB;
    explicit ParModel(uint32_t rb) : RB(rb) {}
    void batch(const std::vector<Rec>& q) {
        const uint32_t nq = (uint32_t)q.size();
        if (!nq) return;
        const uint32_t nblk = (nq + RB - 1) / RB;
        // tables
        std::vector<int> type(nblk); std::vector<uint32_t> da(nblk), db(nblk); std::vector<std::vector<uint16_t>> tab(nblk);
        for (uint32_t blk = 0; blk < nblk; ++blk) {
            Block K; analyse(K, q, blk * RB, std::min(RB, nq - blk * RB), cy, blk == 0);
            uint32_t lastdef = 0;
            for (uint32_t c = 1; c <= K.nc; ++c) if (K.cdef[c]) lastdef = c;
            if (K.nc == 0) { type[blk] = 0; da[blk] = K.cGs[0]; db[blk] = K.cGj[0]; }
            else if (lastdef) { uint32_t p = 0; for (uint32_t c = lastdef; c <= K.nc; ++c) p = apply(K, c, p); type[blk] = 1; da[blk] = p; }
            else { type[blk] = 2; tab[blk].resize(kCovStates); for (uint32_t s = 0; s < kCovStates; ++s) { uint32_t p = cov_state_value(s); for (uint32_t c = 0; c <= K.nc; ++c) p = apply(K, c, p); tab[blk][s] = (uint16_t)cov_state_index(p); } }
        }
        // link
        std::vector<uint32_t> state_in(nblk);
        uint32_t p = cy.p_prev;
        for (uint32_t blk = 0; blk < nblk; ++blk) {
            state_in[blk] = p;
            if (type[blk] == 0) p = cov_stretch(p, da[blk], db[blk]);
            else if (type[blk] == 1) p = da[blk];
            else p = cov_state_value(tab[blk][cov_state_index(p)]);
        }
        // codes
        std::vector<uint64_t> base(nq); std::vector<uint32_t> A(nq), Bv(nq);
        uint64_t X = cy.xc + cy.p_prev;  // virtual coordinate of the last record before the batch
        uint32_t plast = 0;
        for (uint32_t blk = 0; blk < nblk; ++blk) {
            Block K; analyse(K, q, blk * RB, std::min(RB, nq - blk * RB), cy, blk == 0);
            uint32_t pp = state_in[blk];
            for (uint32_t c = 0; c <= K.nc; ++c) {
                uint32_t qq = pp, dx = 0;
                if (c) { bool reset; qq = cov_step(pp, K.cg[c], K.cdef[c] != 0, reset); dx = (K.isfirst && c == 1) ? 0 : cov_dx(pp, K.cg[c], reset); }
                pp = cov_stretch(qq, K.cGs[c], K.cGj[c]);
                K.cGs[c] = qq; K.cg[c] = dx;
            }
            for (uint32_t i = 1; i <= K.n; ++i) {
                uint32_t c = K.cord[i], cp = K.cpos[c], qq = K.cGs[c], ps, dx;
                if (c && cp == i) { ps = qq; dx = K.cg[c]; }
                else {
                    uint32_t js = K.cjs[c], g = K.sb[i] - K.sb[i - 1];
                    if (qq == kCovEdge && i == js) { ps = 0; dx = 0; }
                    else { ps = cov_stretch(qq, K.sb[i] - K.sb[cp], js ? K.sb[i] - K.sb[js] : 0); dx = g; }
                }
                X += (uint64_t)(int64_t)(int32_t)dx;  // a backward step inside the windows moves X back
                uint32_t r = blk * RB + i - 1;
                uint32_t iv = cov_pack_iv(q[r].c0, q[r].len, false);
                uint32_t c0 = iv & 2047u, len = (iv >> 11) & 2047u, lim = 2000u - ps;
                base[r] = X; A[r] = std::min(c0, lim); Bv[r] = std::min(c0 + len, lim);
                if ((X - ps) % 1000 != 0) { printf("V not aligned: r %u i %u c %u cp %u qq %u ps %u dx %d X %llu js %u sb[i] %u sb[i-1] %u sb[cp] %u nc %u\n", r, i, c, cp, qq, ps, (int)dx, (unsigned long long)X, K.cjs[c], K.sb[i], K.sb[i-1], K.sb[cp], K.nc); exit(3); }
                pstate.push_back(ps);
                plast = ps;
            }
        }
        // everything below the start of the first live window is final: histogram of [xc, xl), xl = X - p of the last
        // record, and the new carry (= the reference's v1 | v2 as a difference array) over [xl, xl + 2000]
        const uint64_t xc = cy.xc, xl = X - plast;
        if (xl < xc || xl - xc > (1ull << 32)) { printf("bad range: xc %llu xl %llu X %llu plast %u\n", (unsigned long long)xc, (unsigned long long)xl, (unsigned long long)X, plast); exit(2); }
        std::vector<int64_t> diff((size_t)(xl - xc) + 1, 0), dn(2048, 0);
        auto add = [&](uint64_t a, uint64_t b) {
            if (a >= b) return;
            if (b > xc && a < xl) { diff[(size_t)(std::max(a, xc) - xc)] += 1; if (b < xl) diff[(size_t)(b - xc)] -= 1; }
            if (b > xl) { dn[(size_t)(std::max(a, xl) - xl)] += 1; dn[(size_t)(b - xl)] -= 1; }
        };
        for (uint32_t r = 0; r < nq; ++r) add(base[r] + A[r], base[r] + Bv[r]);
        for (uint32_t d = 0; d <= 2000; ++d) {
            int64_t v = cy.D[d];
            if (!v) continue;
            uint64_t x = xc + d;
            if (x < xl) diff[(size_t)(x - xc)] += v;
            if (x <= xl) dn[0] += v; else dn[(size_t)(x - xl)] += v;
        }
        int64_t depth = 0;
        for (uint64_t x = xc; x < xl; ++x) { depth += diff[(size_t)(x - xc)]; poscov[depth > 100 ? 100 : depth] += 1; }
        cy.D = dn;
        cy.first = false; cy.rid_prev = q[nq - 1].rid; cy.b_prev = q[nq - 1].b; cy.p_prev = plast; cy.xc = xl;
    }
    void finish() {
        int64_t depth = 0;
        for (uint32_t d = 0; d < 2000u; ++d) { depth += cy.D[d]; poscov[depth > 100 ? 100 : depth] += 1; }
    }
};

static uint64_t g_edges = 0, g_records = 0, g_backward = 0;
static int run_case(const std::vector<Rec>& recs, const std::vector<size_t>& cuts, uint32_t rb, const char* name) {
    SeqModel S;
    for (const Rec& r : recs) S.coverage(r);
    S.finish();
    for (uint32_t p : S.pstate) g_edges += p == 2000;
    g_records += recs.size();
    for (size_t i = 1; i < recs.size(); ++i) g_backward += recs[i].rid == recs[i - 1].rid && recs[i].b < recs[i - 1].b;
    ParModel P(rb);
    size_t lo = 0;
    for (size_t c : cuts) { P.batch(std::vector<Rec>(recs.begin() + lo, recs.begin() + c)); lo = c; }
    P.batch(std::vector<Rec>(recs.begin() + lo, recs.end()));
    P.finish();
    int bad = 0;
    if (S.pstate != P.pstate) {
        for (size_t i = 0; i < S.pstate.size() && bad < 5; ++i)
            if (S.pstate[i] != P.pstate[i]) { printf("%s: state of record %zu: sequential %u parallel %u\n", name, i, S.pstate[i], P.pstate[i]); ++bad; }
        ++bad;
    }
    for (int i = 0; i <= 100; ++i)
        if (S.poscov[i] != P.poscov[i]) { if (bad < 10) printf("%s: poscov[%d]: sequential %llu parallel %llu\n", name, i, (unsigned long long)S.poscov[i], (unsigned long long)P.poscov[i]); ++bad; }
    return bad;
}

int main() {
    std::mt19937_64 rng(12345);
    int bad = 0, cases = 0;
    for (int it = 0; it < 400; ++it) {
        const int mode = it % 8;
        const size_t n = 1 + rng() % (mode == 7 ? 6000 : 1500);
        std::vector<Rec> recs;
        uint32_t b = (uint32_t)(rng() % 5000);
        int32_t rid = 0;
        for (size_t i = 0; i < n; ++i) {
            uint32_t g;
            const uint32_t u = (uint32_t)(rng() % 1000);
            switch (mode) {
                case 0: g = (uint32_t)(rng() % 40); break;                                      // dense
                case 1: g = u < 900 ? (uint32_t)(rng() % 600) : 900 + (uint32_t)(rng() % 1300); break;  // gaps around 1000..2200
                case 2: g = u < 500 ? 0 : (u < 800 ? 1000 : (u < 900 ? 2000 : (uint32_t)(rng() % 3000))); break;  // exact edges, ties
                case 3: g = 990 + (uint32_t)(rng() % 1020); break;                             // every record a candidate, no definite resets
                case 4: g = u < 950 ? (uint32_t)(rng() % 300) : (uint32_t)(0u - rng() % 3000); break;   // unsorted: backward steps
                case 5: g = u < 30 ? 2000u - (uint32_t)(rng() % 3) : (u < 500 ? 0 : (uint32_t)(rng() % 1001)); break;
                case 6: g = (uint32_t)(rng() % 2500); break;
                default: g = u < 970 ? (uint32_t)(rng() % 350) : 1000 + (uint32_t)(rng() % 1001); break;  // long ambiguous chains
            }
            b += g;
            if (rng() % (mode == 3 ? 100000 : 400) == 0) { rid = (int32_t)(rng() % 3); if (rng() % 2) b = (uint32_t)(rng() % 100000); }
            Rec r;
            r.rid = rid; r.b = b;
            r.c0 = rng() % 10 == 0 ? (uint32_t)(rng() % 60) : 0;
            r.len = rng() % 50 == 0 ? (uint32_t)(rng() % 2600) : 100 + (uint32_t)(rng() % 60);
            recs.push_back(r);
        }
        std::vector<size_t> cuts;
        const int ncut = (int)(rng() % 4);
        for (int c = 0; c < ncut; ++c) cuts.push_back(rng() % (n + 1));
        if (it % 16 == 0) cuts.push_back(0);  // an empty first batch
        std::sort(cuts.begin(), cuts.end());
        const uint32_t rb = (it % 3 == 0) ? 2048u : (it % 3 == 1 ? 64u : 7u);
        char name[64];
        snprintf(name, sizeof(name), "case %d (mode %d, n %zu, rb %u)", it, mode, n, rb);
        bad += run_case(recs, cuts, rb, name);
        ++cases;
    }
    {   // no record at all: two empty windows
        bad += run_case(std::vector<Rec>(), std::vector<size_t>(), 2048, "empty");
        ++cases;
    }
    if (bad) { printf("cov_selftest: %d mismatches in %d cases\n", bad, cases); return 1; }
    printf("cov_selftest ok: %d cases, %llu records, %llu in the edge state (pos == 2000), %llu backward steps\n", cases, (unsigned long long)g_records, (unsigned long long)g_edges, (unsigned long long)g_backward);
    return 0;
}
