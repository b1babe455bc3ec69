// cov_selftest.cpp -- CPU check of the parallel form of the coverage statistic (bamqc_b200/csrc/cov_math.h and the
// block decomposition of kernel_cov.cuh) against a sequential restatement of OverallNumbers::coverage /
// update_coverage / update_vectors (/root/reference/src/OverallNumbers.hpp:59-135) and the end-of-run flush
// (/root/reference/src/bamqualcheck.cpp:447-453).  Run by tests/test_cov_cpu.py.  Test infrastructure only.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
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
        const uint32_t g = K.sb[i] - K.sb[i - 1];
        if (other || g >= kCovV) { ++ord; K.cpos[ord] = i; K.cg[ord] = g; K.cdef[ord] = other || cov_gap_definite(g); }
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
    std::vector<int64_t> head = std::vector<int64_t>(2048, 0);  // own depth over [0, 2000] of the first batch (shard mode)
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
        if (xc == 0)
            for (uint32_t r = 0; r < nq; ++r) {
                const uint64_t a = base[r] + A[r], b = base[r] + Bv[r];
                if (a < b && a <= 2000) { head[(size_t)a] += 1; if (b <= 2000) head[(size_t)b] -= 1; }
            }
        for (uint32_t d = 0; d <= 2000; ++d) {
            int64_t v = cy.D[d];
            if (!v) continue;
            uint64_t x = xc + d;
            if (x < xl) diff[(size_t)(x - xc)] += v;
            if (x <= xl) dn[0] += v; else dn[(size_t)(x - xl)] += v;
        }
        int64_t depth = 0;
        std::vector<uint64_t> direct(101, 0);
        for (uint64_t x = xc; x < xl; ++x) { depth += diff[(size_t)(x - xc)]; direct[depth > 100 ? 100 : depth] += 1; }
        // the same histogram the way k_cov_tiles does it: tiles of T positions, record ranges from first_rec (first record
        // whose window start lies at or beyond the tile), look-back of two tiles, carry entries in tiles 0 and 1
        {
            const uint64_t T = 1024;
            std::vector<uint64_t> V(nq);
            for (uint32_t r = 0; r < nq; ++r) V[r] = base[r] - pstate[pstate.size() - nq + r];
            const uint32_t vt_last = (uint32_t)((V[nq - 1] - xc) / T);
            const uint32_t ntiles = (uint32_t)((V[nq - 1] - xc + T - 1) / T);
            std::vector<uint32_t> first_rec(vt_last + 2, 0xFFFFFFFFu);
            for (uint32_t r = 0; r < nq; ++r) {
                const uint32_t t1 = (uint32_t)((V[r] - xc) / T);
                for (uint32_t tt = r ? (uint32_t)((V[r - 1] - xc) / T) + 1 : 0; tt <= t1; ++tt) first_rec[tt] = r;
            }
            std::vector<uint64_t> tiled(101, 0);
            for (uint32_t t = 0; t < ntiles; ++t) {
                const uint64_t T0 = xc + (uint64_t)t * T, T1 = std::min(T0 + T, xl);
                const uint32_t len = (uint32_t)(T1 - T0);
                std::vector<int64_t> d(T + 1, 0);
                const uint32_t jlo = t >= 2 ? first_rec[t - 2] : 0, jhi = t + 1 <= vt_last ? first_rec[t + 1] : nq;
                if (jlo == 0xFFFFFFFFu || jhi == 0xFFFFFFFFu) { printf("first_rec hole at tile %u\n", t); exit(4); }
                for (uint32_t j = jlo; j < jhi; ++j) {
                    const uint64_t a = base[j] + A[j], b = base[j] + Bv[j];
                    if (b <= T0 || a >= T1 || a >= b) continue;
                    d[(size_t)(std::max(a, T0) - T0)] += 1;
                    if (b < T1) d[(size_t)(b - T0)] -= 1;
                }
                if (t < 2)
                    for (uint32_t i = 0; i < len && t * T + i <= 2000; ++i) d[i] += cy.D[t * T + i];
                int64_t dep = 0;
                if (t == 1) for (uint32_t i = 0; i < T; ++i) dep += cy.D[i];   // depth the open windows carry into tile 1
                for (uint32_t i = 0; i < len; ++i) { dep += d[i]; tiled[dep > 100 ? 100 : dep] += 1; }
            }
            if (tiled != direct) { printf("tiled histogram differs from the direct one (nq %u, ntiles %u)\n", nq, ntiles); exit(5); }
            // k_cov_carry: only the records from first_rec[vt_last - 2] on can reach beyond xl
            std::vector<int64_t> dn2(2048, 0);
            const uint32_t jlo = vt_last >= 2 ? first_rec[vt_last - 2] : 0;
            for (uint32_t j = jlo; j < nq; ++j) {
                const uint64_t a = base[j] + A[j], b = base[j] + Bv[j];
                if (a < b && b > xl) { dn2[(size_t)(std::max(a, xl) - xl)] += 1; dn2[(size_t)(b - xl)] -= 1; }
            }
            for (uint32_t d = 0; d <= 2000; ++d) {
                int64_t v = cy.D[d];
                if (!v) continue;
                uint64_t x = xc + d;
                if (x <= xl) dn2[0] += v; else dn2[(size_t)(x - xl)] += v;
            }
            if (dn2 != dn) { printf("carry from the tail records differs (nq %u, vt_last %u, jlo %u)\n", nq, vt_last, jlo); exit(6); }
        }
        for (int i = 0; i <= 100; ++i) poscov[i] += direct[i];
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

// One stream cut into shards, each processed on its own from an unknown state (cov_math.h "Shards")
static int run_sharded(const std::vector<Rec>& recs, std::vector<size_t> cuts, uint32_t rb, const char* name, FILE* dump = nullptr) {
    SeqModel S;
    for (const Rec& r : recs) S.coverage(r);
    S.finish();
    cuts.push_back(recs.size());
    const int K = (int)cuts.size();
    std::vector<std::vector<Rec>> part((size_t)K);
    size_t lo = 0;
    for (int k = 0; k < K; ++k) { part[k].assign(recs.begin() + lo, recs.begin() + cuts[k]); lo = cuts[k]; }
    // A: boundaries; B: what each shard needs to know about the record before it
    std::vector<int> have_prev(K, 0); std::vector<int32_t> prid(K, 0); std::vector<uint32_t> pb(K, 0);
    {
        bool have = false; int32_t rid = 0; uint32_t b = 0;
        for (int k = 0; k < K; ++k) {
            have_prev[k] = have; prid[k] = rid; pb[k] = b;
            if (!part[k].empty()) { have = true; rid = part[k].back().rid; b = part[k].back().b; }
        }
    }
    // C: shard functions (state at entry -> state at exit); D: chain them
    std::vector<uint32_t> p_in(K, 0);
    std::vector<std::vector<uint16_t>> Fs((size_t)K);
    uint32_t p = 0;
    for (int k = 0; k < K; ++k) {
        p_in[k] = p;
        Fs[k].resize(kCovStates);
        for (uint32_t s = 0; s < kCovStates; ++s) Fs[k][s] = (uint16_t)s;
        if (part[k].empty()) continue;
        std::vector<uint16_t>& F = Fs[k];
        for (uint32_t s = 0; s < kCovStates; ++s) {
            uint32_t q = cov_state_value(s);
            int32_t rid = prid[k]; uint32_t b = pb[k];
            bool first = !have_prev[k];
            for (const Rec& r : part[k]) {
                bool reset;
                if (first) { q = 0; first = false; }
                else q = cov_step(q, r.b - b, r.rid != rid, reset);
                rid = r.rid; b = r.b;
            }
            F[s] = (uint16_t)cov_state_index(q);
        }
        p = cov_state_value(F[cov_state_index(p)]);
    }
    // E: every shard on its own; F: combine
    std::vector<uint64_t> total(101, 0);
    std::vector<std::vector<uint64_t>> own((size_t)K);
    std::vector<std::vector<int32_t>> heads((size_t)K, std::vector<int32_t>(2001, 0)), tails((size_t)K, std::vector<int32_t>(2001, 0));
    std::vector<CovShardPiece> pieces((size_t)K);
    for (int k = 0; k < K; ++k) {
        ParModel P(rb);
        P.cy.first = !have_prev[k]; P.cy.rid_prev = prid[k]; P.cy.b_prev = pb[k]; P.cy.p_prev = p_in[k]; P.cy.xc = 0;
        P.batch(part[k]);
        for (int i = 0; i <= 100; ++i) total[i] += P.poscov[i];
        own[k] = P.poscov;
        for (int i = 0; i <= 2000; ++i) { heads[k][i] = (int32_t)P.head[i]; tails[k][i] = (int32_t)P.cy.D[i]; }
        pieces[k].n = part[k].size(); pieces[k].span = P.cy.xc; pieces[k].head = heads[k].data(); pieces[k].tail = tails[k].data();
    }
    long long delta[101];
    cov_shards_combine(pieces.data(), K, delta);
    if (dump) {  // the pieces as JSON (tests/test_dist_gloo.py replays the exchange protocol with them)
        fprintf(dump, "{\"expected\": [");
        for (int i = 0; i <= 100; ++i) fprintf(dump, "%s%llu", i ? "," : "", (unsigned long long)S.poscov[i]);
        fprintf(dump, "], \"pieces\": [");
        for (int k = 0; k < K; ++k) {
            fprintf(dump, "%s{\"n\": %zu, \"first_rid\": %d, \"first_b\": %u, \"last_rid\": %d, \"last_b\": %u, \"have_prev\": %d, \"prev_rid\": %d, \"prev_b\": %u, \"p_in\": %u, \"span\": %llu, ",
                    k ? "," : "", part[k].size(), part[k].empty() ? 0 : part[k].front().rid, part[k].empty() ? 0u : part[k].front().b,
                    part[k].empty() ? 0 : part[k].back().rid, part[k].empty() ? 0u : part[k].back().b, have_prev[k], prid[k], pb[k], p_in[k], (unsigned long long)pieces[k].span);
            auto arr = [&](const char* key, auto& v, size_t n, const char* end) {
                fprintf(dump, "\"%s\": [", key);
                for (size_t i = 0; i < n; ++i) fprintf(dump, "%s%lld", i ? "," : "", (long long)v[i]);
                fprintf(dump, "]%s", end);
            };
            arr("table", Fs[k], kCovStates, ", "); arr("head", heads[k], 2001, ", "); arr("tail", tails[k], 2001, ", "); arr("poscov", own[k], 101, "}");
        }
        fprintf(dump, "]}\n");
    }
    int bad = 0;
    for (int i = 0; i <= 100; ++i)
        if ((long long)S.poscov[i] != (long long)total[i] + delta[i]) {
            if (bad < 6) printf("%s sharded x%d: poscov[%d]: sequential %llu, shards %lld\n", name, K, i, (unsigned long long)S.poscov[i], (long long)total[i] + delta[i]);
            ++bad;
        }
    return bad;
}

int main(int argc, char** argv) {
    std::mt19937_64 rng(12345);
    if (argc >= 4 && std::string(argv[1]) == "--dump-shards") {  // --dump-shards K out.json: one random stream in K pieces
        const int K = atoi(argv[2]);
        std::vector<Rec> recs;
        uint32_t b = 1000;
        int32_t rid = 0;
        for (int i = 0; i < 6000; ++i) {
            const uint32_t u = (uint32_t)(rng() % 1000);
            b += u < 900 ? (uint32_t)(rng() % 400) : 900 + (uint32_t)(rng() % 1300);
            if (rng() % 1500 == 0) { ++rid; b = (uint32_t)(rng() % 5000); }
            Rec r; r.rid = rid; r.b = b; r.c0 = rng() % 10 == 0 ? (uint32_t)(rng() % 60) : 0; r.len = 100 + (uint32_t)(rng() % 60);
            recs.push_back(r);
        }
        std::vector<size_t> cuts;
        for (int k = 1; k < K; ++k) cuts.push_back(k == 2 ? cuts.back() : recs.size() * k / K + rng() % 50);  // piece 2 (if any) is empty
        FILE* f = fopen(argv[3], "w");
        const int bad = run_sharded(recs, cuts, 2048, "dump", f);
        fclose(f);
        return bad ? 1 : 0;
    }
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
        bad += run_sharded(recs, cuts, rb, name);
        cases += 2;
    }
    {   // no record at all: two empty windows
        bad += run_case(std::vector<Rec>(), std::vector<size_t>(), 2048, "empty");
        bad += run_sharded(std::vector<Rec>(), std::vector<size_t>(1, 0), 2048, "empty");
        cases += 2;
    }
    if (bad) { printf("cov_selftest: %d mismatches in %d cases\n", bad, cases); return 1; }
    printf("cov_selftest ok: %d cases, %llu records, %llu in the edge state (pos == 2000), %llu backward steps\n", cases, (unsigned long long)g_records, (unsigned long long)g_edges, (unsigned long long)g_backward);
    return 0;
}
