// TEST INFRASTRUCTURE.  C wrapper around the REFERENCE's own kmerstream sources, compiled where they
// lie under /root/reference (never copied): src/kmerstream/RepHash.{hpp,cpp}, lsb.cpp,
// StreamCounter.hpp, mersennetwister.h and src/ReadQualityHasher.hpp.  Output: oracle/_ref/libkmerstream_ref.so
// (git-ignored, travels to the GPU box).  Used by tests to pin the oracle's restatement and by
// tests/golden/make_kmerstream_golden.py to generate the committed golden vectors.
//
// ReadQualityHasher.hpp includes CommandLineParser.hpp (SeqAn ArgumentParser, absent here): its include
// guard is pre-defined and the two things the header needs from it are declared below.
#include <cstring>
#include <cmath>
#include <sstream>
#include <cassert>
#include <time.h>
#include <string>
#include <vector>
#include <stdint.h>
#include <stddef.h>

#define COMMAND_LINE_PARSER_H_
struct ProgramOptions {  // fields as in src/CommandLineParser.hpp:13-41
    std::vector<int> klist;
    double e;
    std::vector<size_t> q_cutoff;
    size_t q_base;
    int seed;
    int isize;
    ProgramOptions() : e(0.01), q_base(33), seed(0), isize(1000) {}
};
namespace seqan {  // only what RunBamStream (ReadQualityHasher.hpp:113-122) touches
typedef std::string CharString;
inline const char* toCString(CharString& s) { return s.c_str(); }
inline size_t length(CharString& s) { return s.size(); }
}  // namespace seqan

#define private public  // expose StreamCounter::table/F2table and RepHash::hvals for dumps
#include "ReadQualityHasher.hpp"
#undef private

extern "C" {
int ref_rephash_windows(int seed, int k, const char* s, int l, uint64_t* out) {
    RepHash hf;
    hf.seed(seed);
    hf.init(k);
    if (l < k) return 0;
    hf.init(s);
    out[0] = hf.hash();
    for (int i = 1; i + k <= l; ++i) {
        hf.update((unsigned char)s[i - 1], (unsigned char)s[i + k - 1]);
        out[i] = hf.hash();
    }
    return l - k + 1;
}
// direct re-initialisation at every window (no rolling)
int ref_rephash_windows_init(int seed, int k, const char* s, int l, uint64_t* out) {
    RepHash hf;
    hf.seed(seed);
    hf.init(k);
    for (int i = 0; i + k <= l; ++i) {
        hf.init(s + i);
        out[i] = hf.hash();
    }
    return l < k ? 0 : l - k + 1;
}
void ref_rephash_hvals(int seed, uint64_t* out64) {
    RepHash hf;
    hf.seed(seed);
    for (int i = 0; i < 32; ++i) {
        out64[2 * i] = hf.hvals[i].hi;
        out64[2 * i + 1] = hf.hvals[i].lo;
    }
}
void ref_streamcounter(double e, int seed, const uint64_t* hashes, uint64_t n, uint64_t* out6, uint64_t* table_out, uint64_t* f2_out) {
    StreamCounter sc(e, seed);
    for (uint64_t i = 0; i < n; ++i) sc(hashes[i]);
    out6[0] = sc.get_sumCount();
    out6[1] = sc.F0();
    out6[2] = sc.f1();
    out6[3] = sc.F2();
    out6[4] = sc.size;
    out6[5] = sc.F2size;
    if (table_out) memcpy(table_out, sc.table, sc.size * sc.MAX_TABLE * 8);
    if (f2_out) memcpy(f2_out, sc.F2table, sc.F2size * 8);
}
void ref_hasher(double e, int seed, int q, int k, const char* seqs, const char* quals, const int* lens, int nreads, uint64_t* out4, uint64_t* table_out, uint64_t* f2_out) {
    ProgramOptions opt;
    opt.e = e;
    opt.seed = seed;
    std::vector<std::vector<ReadQualityHasher> > sps(1, std::vector<ReadQualityHasher>(1, ReadQualityHasher(opt)));
    sps[0][0].setQualityCutoff((size_t)q);
    sps[0][0].setK((size_t)k);
    size_t off = 0;
    for (int i = 0; i < nreads; ++i) {
        seqan::CharString s(seqs + off, seqs + off + lens[i]), qq(quals + off, quals + off + lens[i]);
        RunBamStream(sps, s, qq);
        off += (size_t)lens[i];
    }
    ReadQualityHasher& h = sps[0][0];
    out4[0] = h.get_sumCount();
    out4[1] = h.F0();
    out4[2] = h.f1();
    out4[3] = h.F2();
    if (table_out) memcpy(table_out, h.sc.table, h.sc.size * h.sc.MAX_TABLE * 8);
    if (f2_out) memcpy(f2_out, h.sc.F2table, h.sc.F2size * 8);
}
uint64_t ref_bitscan(uint64_t v) { return bitScanForward(v); }
}
