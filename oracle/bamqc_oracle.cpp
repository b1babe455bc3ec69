// =============================================================================
// bamqc_oracle.cpp  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of DecodeGenetics/BamQC `bamqualcheck` (the per-record
// statistics pass and everything around it that shapes the `.bamqc` output).
// It exists only so that tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs have something to compare the CUDA path
// against.  Nothing under bamqc_b200/ may include, link or execute this file.
//
// Each function cites the reference file:line it follows (paths relative to
// /root/reference).  The reference depends on SeqAn 1.4.2 (README.md:7-8,
// Makefile:14), which is not vendored and not available here; decoding and
// alphabet conversions that live inside SeqAn are restated from the SAM/BAM
// specification and the behaviours listed in SURVEY.md Appendix C (R1-R15).
//
// Parity pinning: the kmerstream part (RepHash, StreamCounter, ReadQualityHasher)
// is pinned against the reference's own sources compiled into oracle/_ref/
// (see oracle/Makefile, tests/test_oracle_kmerstream.py, tests/golden/).  The
// statistics classes are additionally pinned against the reference's own
// headers compiled over a minimal SeqAn stand-in (oracle/miniseqan/, see
// oracle/Makefile target _ref/bamqualcheck_ref) when /root/reference is present.
// The SeqAn boundary itself (BAM decode, alphabet tables) is "parity unpinned".
//
// Build: see oracle/Makefile.   Usage (same flags as the reference CLI,
// src/CommandLineParser.hpp:59-85):
//   bamqualcheck_oracle -r ref.fa [-c LIST] [-i INT] [-k LIST] [-q LIST]
//                       [-e F] [-s INT] -o out.bamqc [--dump out.dump] in.bam
// `in.bam` may be BGZF-compressed or a raw (already inflated) BAM byte stream.
// =============================================================================
#include <zlib.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <random>
#include <set>
#include <sstream>
#include <string>
#include <vector>

namespace oracle {

// -----------------------------------------------------------------------------
// Options  (src/CommandLineParser.hpp:13-41, defaults :64-81)
// -----------------------------------------------------------------------------
struct Options {
    std::string bamFile;
    std::string referenceFile = "genome.fa";
    std::string outputFile;
    std::string dumpFile;  // oracle-only: raw tables for the parity tests
    std::string chroms =
        "chr1,chr2,chr3,chr4,chr5,chr6,chr7,chr8,chr9,chr10,chr11,chr12,chr13,chr14,chr15,chr16,"
        "chr17,chr18,chr19,chr20,chr21,chr22";
    std::vector<int> klist;
    double e = 0.01;
    std::vector<size_t> q_cutoff;
    size_t q_base = 33;
    int seed = 1;
    int isize = 1000;
    bool quiet = false;
    long maxRecords = -1;  // oracle-only: stop after this many records (bench sample)
};

// -----------------------------------------------------------------------------
// kmerstream: RepHash  (src/kmerstream/RepHash.hpp:8-123, RepHash.cpp:4-17)
// -----------------------------------------------------------------------------
static const unsigned char kTwin[32] = {0,  20, 2,  7,  4,  5,  6,  3,  8,  9,  10, 11, 12, 13, 14, 15,
                                        16, 17, 18, 19, 1,  21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31};

struct State128 {
    uint64_t hi = 0, lo = 0;
};

class RepHashO {
   public:
    // RepHash.cpp:4-17.  MTRand(seed).randInt() is MT19937 (mersennetwister.h); the two randInt()
    // calls of one word are evaluated left to right by g++ (golden vectors in tests/golden).
    void seed(int s) {
        std::mt19937 mtr((uint32_t)s);
        for (int i = 0; i < 32; i++) {
            uint64_t a = mtr(), b = mtr(), c = mtr(), d = mtr();
            hvals[i].hi = (a << 32) | b;
            hvals[i].lo = (c << 32) | d;
        }
        h = State128();
        ht = State128();
    }
    // RepHash.hpp:45-51
    void init(int k_) {
        k = (size_t)k_;
        last1mask = (1ULL << 63);
        lastkmask = ((1ULL << k) - 1) << (64 - k);
        firstkmask = (1ULL << k) - 1;
    }
    // RepHash.hpp:85-99
    void init(const char* s_) {
        h = State128();
        ht = State128();
        const unsigned char* s = (const unsigned char*)s_;
        for (size_t i = 0; i < k; i++) {
            shl1(h);
            x(h, hvals[s[i] & 31]);
            shl1(ht);
            x(ht, hvals[kTwin[s[k - 1 - i] & 31]]);
        }
    }
    // RepHash.hpp:101-114
    void update(unsigned char out, unsigned char in) {
        State128 z = hvals[out & 31];
        shlk(z);
        shl1(h);
        x(h, z);
        x(h, hvals[in & 31]);
        State128 zt = hvals[kTwin[in & 31]];
        shlk(zt);
        x(ht, hvals[kTwin[out & 31]]);
        x(ht, zt);
        shr1(ht);
    }
    uint64_t hash() const { return h.lo ^ ht.lo; }  // RepHash.hpp:81-83
    State128 hvals[32];

   private:
    static void x(State128& a, const State128& b) {
        a.hi ^= b.hi;
        a.lo ^= b.lo;
    }
    void shlk(State128& v) const {  // RepHash.hpp:57-61
        uint64_t upper = v.hi & lastkmask;
        v.hi = (v.hi << k) | ((v.lo & lastkmask) >> (64 - k));
        v.lo = (v.lo << k) | (upper >> (64 - k));
    }
    void shl1(State128& v) const {  // RepHash.hpp:69-73
        uint64_t last1 = v.hi & last1mask;
        v.hi = (v.hi << 1) | ((v.lo & last1mask) >> 63);
        v.lo = (v.lo << 1) | (last1 >> 63);
    }
    void shr1(State128& v) const {  // RepHash.hpp:75-79
        uint64_t first1 = v.hi & 1ULL;
        v.hi = (v.hi >> 1) | ((v.lo & 1ULL) << 63);
        v.lo = (v.lo >> 1) | (first1 << 63);
    }
    size_t k = 0;
    uint64_t last1mask = 0, lastkmask = 0, firstkmask = 0;
    State128 h, ht;
};

// -----------------------------------------------------------------------------
// kmerstream: StreamCounter  (src/kmerstream/StreamCounter.hpp:23-357, lsb.cpp:26-29)
// -----------------------------------------------------------------------------
static size_t roundUpPowerOfTwo(size_t size) {  // StreamCounter.hpp:11-21
    size--;
    size |= size >> 1;
    size |= size >> 2;
    size |= size >> 4;
    size |= size >> 8;
    size |= size >> 16;
    size |= size >> 32;
    size++;
    return size;
}

// lsb.cpp:26-29: index of the least significant one bit, 63 for 0 (de Bruijn table lookup of
// (bb & -bb) * debruijn >> 58; index64[0] == 63).
static inline uint64_t bitScanForward(uint64_t bb) { return bb ? (uint64_t)__builtin_ctzll(bb) : 63; }

class StreamCounterO {
   public:
    static const size_t MAX_TABLE = 32, countsPerLong = 16, countWidth = 4;
    static const uint64_t maxVal = 15;
    StreamCounterO(double e, int /*seed*/) {  // StreamCounter.hpp:25-43
        size_t numcounts = (size_t)(48.0 / (e * e) + 1);
        F2size = roundUpPowerOfTwo((size_t)(2.0 / (e * e) + 1));
        F2table.assign(F2size, 0);
        if (numcounts < 8192) numcounts = 8192;
        size = (numcounts + countsPerLong - 1) / countsPerLong;
        size = roundUpPowerOfTwo(size);
        mask = (size * countsPerLong) - 1;
        M.assign(MAX_TABLE, 0);
        table.assign(size * MAX_TABLE, 0);
    }
    void operator()(uint64_t hashval) {  // StreamCounter.hpp:67-93
        sumCount++;
        ++F2table[hashval & (F2size - 1)];
        size_t w = bitScanForward(hashval);
        if (w >= MAX_TABLE) w = MAX_TABLE - 1;
        if (M[w] == size * countsPerLong * maxVal) return;
        uint64_t hval = hashval >> (w + 1);
        uint64_t index = hval & mask;
        uint64_t val = getVal(index, w);
        if (val != maxVal) {
            setVal(index, w, val + 1);
            M[w]++;
        }
    }
    size_t F0() const {  // StreamCounter.hpp:114-140
        size_t R = size * countsPerLong;
        double sum = 0;
        int n = 0;
        double limit = 0.2;
        while (n == 0 && limit > 1e-8) {
            for (size_t i = 0; i < MAX_TABLE; i++) {
                size_t ts = 0;
                for (size_t j = 0; j < R; j++)
                    if (getVal(j, i) > 0) ts++;
                if (ts <= (1 - limit) * R && ts >= limit * R) {
                    double est = (log(1.0 - ts / ((double)R)) / log(1.0 - 1.0 / R)) * pow(2.0, i + 1);
                    sum += est;
                    n++;
                    break;
                }
            }
            limit = limit / 1.5;
        }
        return nanToSize(sum / n);
    }
    size_t f1() const {  // StreamCounter.hpp:142-172
        size_t R = size * countsPerLong;
        double sum = 0;
        int n = 0;
        double limit = 0.2;
        while (n == 0 && limit > 1e-8) {
            for (size_t i = 0; i < MAX_TABLE; i++) {
                size_t r1 = 0, r0 = 0;
                for (size_t j = 0; j < R; j++) {
                    uint64_t val = getVal(j, i);
                    if (val == 0) r0++;
                    if (val == 1) r1++;
                }
                if ((r0 <= (1 - limit) * R) && (r0 >= limit * R)) {
                    sum += (R - 1) * (r1 / ((double)r0)) * pow(2.0, i + 1);
                    n++;
                    break;
                }
            }
            limit = limit / 1.5;
        }
        return nanToSize(sum / n);
    }
    size_t F2() const {  // StreamCounter.hpp:308-317 (sequential summation order kept)
        double sum = 0, sqsum = 0;
        for (size_t i = 0; i < F2size; i++) {
            double c = (double)F2table[i];
            sum += c;
            sqsum += c * c;
        }
        return (size_t)(sqsum + (sqsum - sum * sum) / F2size);
    }
    size_t get_sumCount() const { return sumCount; }  // StreamCounter.hpp:319-321
    uint64_t getVal(size_t index, size_t w) const {    // StreamCounter.hpp:325-330
        size_t wordindex = w * size + (index / countsPerLong);
        size_t bitindex = index & (countsPerLong - 1);
        uint64_t bitmask = maxVal << (countWidth * bitindex);
        return (table[wordindex] & bitmask) >> (countWidth * bitindex);
    }
    void setVal(size_t index, size_t w, uint64_t val) {  // StreamCounter.hpp:332-340
        if (val > maxVal) val = maxVal;
        size_t wordindex = w * size + (index / countsPerLong);
        size_t bitindex = index & (countsPerLong - 1);
        uint64_t bitmask = maxVal << (countWidth * bitindex);
        table[wordindex] = (((val & maxVal) << (countWidth * bitindex)) & bitmask) | (table[wordindex] & ~bitmask);
    }
    std::vector<uint64_t> table, F2table;
    std::vector<size_t> M;
    size_t F2size = 0, size = 0, sumCount = 0;
    uint64_t mask = 0;

   private:
    // (size_t)(NaN) as compiled by g++ on x86-64 (cvttsd2si path) yields 2^63; measured with the
    // reference's own StreamCounter on an empty sketch (tests/golden/kmerstream_golden.json).
    static size_t nanToSize(double v) {
        if (std::isnan(v)) return 9223372036854775808ULL;
        return (size_t)v;
    }
};

// -----------------------------------------------------------------------------
// ReadQualityHasher  (src/ReadQualityHasher.hpp:13-122)
// -----------------------------------------------------------------------------
class ReadQualityHasherO {
   public:
    ReadQualityHasherO(const Options& opt) : q_cutoff(0), q_base(opt.q_base), k(0), sc(opt.e, opt.seed) {
        if (opt.seed != 0) hf.seed(opt.seed);  // :15-19 (seed 0 = time based in the reference; unsupported)
    }
    void setQualityCutoff(size_t q) { q_cutoff = q; }
    void setK(size_t k_) {
        k = k_;
        hf.init((int)k);
    }
    void operator()(const char* s, size_t l, const char* q, size_t /*ql*/) {  // :30-68
        size_t i = 0, j = 0;
        bool last_valid = false;
        if (l < k) return;
        while (j < l) {
            char c = s[j];
            if (c != 'N' && c != 'n' && (q[j] >= (char)(q_base + q_cutoff))) {
                if (last_valid) {
                    hf.update(s[i], s[j]);
                    i++;
                    j++;
                } else {
                    if (i + k - 1 == j) {
                        hf.init(s + i);
                        last_valid = true;
                        j++;
                    } else {
                        j++;
                    }
                }
            } else {
                j++;
                i = j;
                last_valid = false;
            }
            if (last_valid) sc(hf.hash());
        }
    }
    size_t q_cutoff, q_base;
    RepHashO hf;
    size_t k;
    StreamCounterO sc;
};

// -----------------------------------------------------------------------------
// Decoded BAM record (what SeqAn's readRecord hands to the loop)
// -----------------------------------------------------------------------------
struct CigarEl {
    char operation;
    uint32_t count;
};
struct Record {
    std::string qName;
    uint16_t flag = 0;
    int32_t rID = -1, beginPos = -1, rNextId = -1, pNext = -1, tLen = 0;
    uint8_t mapQ = 0;
    std::vector<CigarEl> cigar;
    std::string seq, qual, tags;
};
struct Tag {  // one entry of SeqAn's BamTagsDict index
    char key[2];
    char type;
    size_t valueBegin, valueEnd;  // raw value bytes inside Record::tags
};

static bool buildTagsDict(const std::string& tags, std::vector<Tag>& dict) {
    dict.clear();
    size_t p = 0, n = tags.size();
    while (p + 3 <= n) {
        Tag t;
        t.key[0] = tags[p];
        t.key[1] = tags[p + 1];
        t.type = tags[p + 2];
        p += 3;
        size_t len = 0;
        switch (t.type) {
            case 'A': case 'c': case 'C': len = 1; break;
            case 's': case 'S': len = 2; break;
            case 'i': case 'I': case 'f': len = 4; break;
            case 'Z': case 'H': {
                size_t q = p;
                while (q < n && tags[q] != '\0') ++q;
                len = q - p + 1;
                break;
            }
            case 'B': {
                if (p + 5 > n) return false;
                char sub = tags[p];
                uint32_t cnt;
                memcpy(&cnt, &tags[p + 1], 4);
                size_t es = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
                len = 5 + (size_t)cnt * es;
                break;
            }
            default: return false;
        }
        t.valueBegin = p;
        t.valueEnd = std::min(n, p + len);
        p += len;
        dict.push_back(t);
    }
    return true;
}

// SeqAn extractTagValue: any integer tag type converted to the destination type (R15).
static bool extractInt(const std::string& tags, const Tag& t, long long& out) {
    const char* p = tags.data() + t.valueBegin;
    switch (t.type) {
        case 'c': out = (int8_t)p[0]; return true;
        case 'C': out = (uint8_t)p[0]; return true;
        case 'A': out = (char)p[0]; return true;
        case 's': { int16_t v; memcpy(&v, p, 2); out = v; return true; }
        case 'S': { uint16_t v; memcpy(&v, p, 2); out = v; return true; }
        case 'i': { int32_t v; memcpy(&v, p, 4); out = v; return true; }
        case 'I': { uint32_t v; memcpy(&v, p, 4); out = v; return true; }
        case 'f': { float v; memcpy(&v, p, 4); out = (long long)v; return true; }
        default: return false;
    }
}

// char -> Dna5 ordinal (R5): ACGT/acgt -> 0..3, everything else 4.
static inline int dna5(char c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 4;
    }
}
// char -> Dna ordinal (R5): C->1 G->2 T/U->3 everything else 0.
static inline int dna4(char c) {
    switch (c) {
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': case 'U': case 'u': return 3;
        default: return 0;
    }
}
// IUPAC complement used by reverseComplement(CharString) (R7).
static inline char complementChar(char c) {
    switch (c) {
        case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
        case 'M': return 'K'; case 'K': return 'M'; case 'R': return 'Y'; case 'Y': return 'R';
        case 'V': return 'B'; case 'B': return 'V'; case 'H': return 'D'; case 'D': return 'H';
        default: return c;  // W, S, N, '=' are their own complements
    }
}

// -----------------------------------------------------------------------------
// OverallNumbers  (src/OverallNumbers.hpp:8-216)
// -----------------------------------------------------------------------------
class OverallNumbersO {
   public:
    unsigned supplementary = 0, duplicates = 0, QCfailed = 0, not_primary_alignment = 0, readcount = 0;
    uint64_t totalbps = 0;
    unsigned bothunmapped = 0, firstunmapped = 0, secondunmapped = 0, first_and_or_second_mapped = 0,
             FF_RR_orientation = 0, properpair_count = 0, auto_properpair_count = 0;
    std::vector<unsigned> poscov;
    std::vector<uint64_t> eightmercount;
    uint64_t covLost = 0;  // oracle-only: increments the reference writes out of bounds (R9)

    OverallNumbersO() : first(true), vsize(1000), covsize(100), shift(0), id(0) {  // :50-57
        eightmercount.assign(65536, 0);
        poscov.assign(covsize + 1, 0);
        v1.assign(vsize, 0);
        v2.assign(vsize, 0);
    }
    void update_vectors() {  // :59-64
        v1.assign(vsize, 0);
        std::swap(v1, v2);
    }
    void update_coverage() {  // :66-77
        for (unsigned i = 0; i < vsize; ++i) {
            if (v1[i] > covsize) poscov[covsize] += 1;
            else poscov[v1[i]] += 1;
        }
    }
    void coverage(const Record& record) {  // :79-135
        unsigned beginpos = (unsigned)record.beginPos;
        if (first) {
            first = false;
            id = record.rID;
            shift = (int)beginpos;
        }
        if (id != record.rID || ((beginpos - (unsigned)shift) > 2 * vsize)) {
            id = record.rID;
            update_coverage();
            update_vectors();
            update_coverage();
            v1.assign(vsize, 0);
            shift = (int)beginpos;
        }
        unsigned pos = beginpos - (unsigned)shift;
        if ((pos > vsize) && (pos < 2 * vsize)) {
            update_coverage();
            update_vectors();
            shift += vsize;
            pos = beginpos - (unsigned)shift;
        }
        int c = 0;
        for (unsigned i = 0; i < record.cigar.size(); ++i) {
            if (record.cigar[i].operation == 'S') c += record.cigar[i].count;
            if (record.cigar[i].operation == 'M' || record.cigar[i].operation == 'D') {
                for (unsigned j = c; j < record.cigar[i].count + c; ++j) {
                    if (pos + j < vsize) {
                        v1[pos + j] += 1;
                    } else {
                        // :129 writes v2[pos - vsize + j]; beyond vsize that is an out-of-bounds write
                        // into String<> slack in the reference (R9): modelled as lost.
                        unsigned idx = pos - vsize + j;
                        if (idx < vsize) v2[idx] += 1;
                        else ++covLost;
                    }
                }
                c += record.cigar[i].count;
            }
        }
    }
    void count8mers(const std::string& seq) {  // :137-168
        // DnaString text = seq; hashInit/hashNext with an ungapped 8-shape = big-endian base-4 code (R8).
        size_t L = seq.size();
        if (L < 8) return;  // reference reads past the end here (hazard); never in test data
        unsigned skip = 0;
        size_t itSeq = 0;
        for (unsigned i = 0; i < 7; ++i, ++itSeq) {
            if (seq[itSeq] == 'N') skip = 8;
            if (skip > 0) --skip;
        }
        uint32_t code = 0;
        for (unsigned i = 0; i < 7; ++i) code = (code << 2) | (uint32_t)dna4(seq[i]);
        for (size_t it = 0; it + 8 <= L; ++it, ++itSeq) {
            code = ((code << 2) | (uint32_t)dna4(seq[it + 7])) & 0xFFFFu;
            if (seq[itSeq] == 'N') skip = 8;
            if (skip > 0) --skip;
            else ++eightmercount[code];
        }
    }
    // :170-216.  Top 10 counts in descending order; equal counts take ascending table index (posSet).
    void ten_most_abundant_kmers(std::ostream& out) const {
        std::vector<uint64_t> v(eightmercount);
        std::partial_sort(v.begin(), v.begin() + 10, v.end(), std::greater<uint64_t>());
        std::set<int> posSet;
        for (int i = 0; i < 10; ++i) {
            int pos = (int)(std::find(eightmercount.begin(), eightmercount.end(), v[i]) - eightmercount.begin());
            while (posSet.count(pos) != 0)
                pos = (int)(std::find(eightmercount.begin() + pos + 1, eightmercount.end(), v[i]) - eightmercount.begin());
            posSet.insert(pos);
            char kmer[9];
            for (int b = 0; b < 8; ++b) kmer[b] = "ACGT"[(pos >> (2 * (7 - b))) & 3];
            kmer[8] = 0;
            out << "nr_" << i + 1 << "_most_abundant_8mer " << kmer << " " << v[i] << std::endl;
        }
    }

   private:
    bool first;
    unsigned vsize, covsize;
    int shift, id;
    std::vector<unsigned> v1, v2;
};

// -----------------------------------------------------------------------------
// QualityCheck  (src/QualityCheck.hpp:8-279)
// -----------------------------------------------------------------------------
class QualityCheckO {
   public:
    std::vector<double> avgqualcount;
    std::vector<unsigned> scposcount_5prime, scposcount_3prime;
    std::vector<uint64_t> dnacount[5];
    std::vector<unsigned> averageQual, Ncount;
    std::vector<uint64_t> GCcount;
    std::vector<unsigned> insertSize, mapQ, readLength, mismatch, delhist, inshist;
    std::vector<uint64_t> qualcount;
    unsigned delcount = 0, inscount = 0, qualcount_readnr = 0;

    explicit QualityCheckO(int isize) { insertSize.assign(isize + 1, 0); }  // :60-64

    int check_read_len(const std::string& seq, const std::string& qual) {  // :70-79
        if (seq.size() != qual.size()) {
            std::cerr << "ERROR: length of sequence and quality is not the same" << "\n";
            return 1;
        }
        return 0;
    }
    void resize_strings(const std::string& dnaseq) {  // :85-105
        if (qualcount.size() < dnaseq.size()) {
            size_t L = dnaseq.size();
            qualcount.resize(L, 0);
            avgqualcount.resize(L, 0);
            Ncount.resize(L + 1, 0);
            GCcount.resize(L + 1, 0);
            scposcount_5prime.resize(L, 0);
            scposcount_3prime.resize(L, 0);
            for (unsigned j = 0; j < 5; ++j) dnacount[j].resize(L, 0);
        }
    }
    void get_count(const std::string& seq, const std::string& qual) {  // :111-116
        resize_strings(seq);
        read_counts(seq, qual);
        read_length(seq);
    }
    void read_counts(const std::string& seq, const std::string& qual) {  // :122-166
        unsigned cntN = 0, cntGC = 0, avgQual = 0;
        qualcount_readnr += 1;
        unsigned j = 0;
        for (size_t i = 0; i < seq.size(); ++i) {
            dnacount[dna5(seq[i])][j] += 1;
            if (seq[i] == 'N') cntN += 1;
            if (seq[i] == 'C' || seq[i] == 'G') cntGC += 1;
            ++j;
        }
        j = 0;
        for (size_t i = 0; i < qual.size(); ++i) {
            // the reference indexes qualcount[j] unchecked; qual longer than seq is a hazard
            if (j < qualcount.size()) qualcount[j] += (int)(unsigned char)qual[i] - 33;
            avgQual += (int)(unsigned char)qual[i] - 33;
            ++j;
        }
        Ncount[cntN] += 1;
        GCcount[cntGC] += 1;
        if (averageQual.size() <= ceil(double(avgQual) / seq.size()))
            averageQual.resize((size_t)(ceil(double(avgQual) / seq.size()) + 1), 0);
        averageQual[(int)round(double(avgQual) / seq.size())] += 1;
    }
    void read_length(const std::string& seq) {  // :168-176
        unsigned lseq = (unsigned)seq.size();
        if (readLength.size() <= lseq) readLength.resize(lseq + 1, 0);
        readLength[lseq] += 1;
    }
    void map_Q(uint8_t mapq) {  // :178-185
        if (mapQ.size() <= mapq) mapQ.resize(mapq + 1, 0);
        mapQ[mapq] += 1;
    }
    void insert_size(int tlen) {  // :187-196
        unsigned index = (unsigned)abs(tlen);
        if (index >= insertSize.size()) index = (unsigned)insertSize.size() - 1;
        insertSize[index] += 1;
    }
    void mis_match(const std::string& tags, const std::vector<Tag>& dict) {  // :198-220
        for (unsigned tagid = 0; tagid < dict.size(); ++tagid) {
            if (dict[tagid].key[0] == 'N' && dict[tagid].key[1] == 'M') {
                char tagType = dict[tagid].type;
                if (tagType == 'c' || tagType == 'C' || tagType == 'i' || tagType == 'I' || tagType == 's' ||
                    tagType == 'S') {
                    long long v = 0;
                    extractInt(tags, dict[tagid], v);
                    unsigned x = (unsigned)v;
                    unsigned mmcount = x - delcount - inscount;
                    if (mmcount > (1u << 26)) {  // the reference would resize to ~4G entries here (hazard)
                        std::cerr << "ORACLE: NM smaller than indel count (hazard, SURVEY App. C)\n";
                        exit(3);
                    }
                    if (mismatch.size() <= mmcount) mismatch.resize(mmcount + 1, 0);
                    mismatch[mmcount] += 1;
                }
            }
        }
    }
    void cigar_count(const Record& record) {  // :222-271
        int cigarlength = (int)record.cigar.size();
        delcount = 0;
        inscount = 0;
        if (cigarlength == 0) {  // the reference dereferences cigar[0] (hazard); never in test data
            std::cerr << "ORACLE: mapped read with empty CIGAR (hazard)\n";
            exit(3);
        }
        if (record.cigar[0].operation == 'S') {
            for (unsigned j = 0; j < record.cigar[0].count; ++j)
                if (j < scposcount_5prime.size()) scposcount_5prime[j] += 1;
        } else if (record.cigar[cigarlength - 1].operation == 'S') {
            for (unsigned j = (unsigned)record.seq.size() - record.cigar[cigarlength - 1].count; j < record.seq.size(); ++j)
                if (j < scposcount_3prime.size()) scposcount_3prime[j] += 1;
        }
        for (int i = 0; i < cigarlength; ++i) {
            if (record.cigar[i].operation == 'D') delcount += record.cigar[i].count;
            else if (record.cigar[i].operation == 'I') inscount += record.cigar[i].count;
        }
        if (delhist.size() <= delcount) delhist.resize(delcount + 1, 0);
        delhist[delcount] += 1;
        if (inshist.size() <= inscount) inshist.resize(inscount + 1, 0);
        inshist[inscount] += 1;
    }
    void avgQualPerPos() {  // :273-279
        for (unsigned i = 0; i < qualcount.size(); ++i) avgqualcount[i] = (qualcount[i] / (double)qualcount_readnr);
    }
};

// -----------------------------------------------------------------------------
// TripletCounting  (src/TripletCounting.hpp:10-301)
// -----------------------------------------------------------------------------
struct TripletCountsO {
    size_t forwardFirst[4] = {0, 0, 0, 0}, forwardSecond[4] = {0, 0, 0, 0}, reverseFirst[4] = {0, 0, 0, 0},
           reverseSecond[4] = {0, 0, 0, 0};
};
struct FastaRecord {
    std::string id, seq;
};
struct GenomeO {  // :60-67; the reference streams the FASTA forward-only, one contig resident at a time
    std::string filename;
    std::vector<FastaRecord> records;  // whole file (R10: id = header line, seq = non-whitespace chars)
    size_t next = 0;
    std::string chromName;
    const std::string* chrom = nullptr;
    unsigned count = 0;
};

static bool loadFasta(GenomeO& g) {
    std::ifstream in(g.filename.c_str(), std::ios::binary);
    if (!in.good()) {
        std::cerr << "ERROR: Could not open fasta file " << g.filename << std::endl;  // :75-79
        return false;
    }
    std::string line;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (!line.empty() && line[0] == '>') {
            g.records.push_back(FastaRecord());
            g.records.back().id = line.substr(1);
        } else if (!g.records.empty()) {
            for (char c : line)
                if (!isspace((unsigned char)c)) g.records.back().seq.push_back(c);
        }
    }
    return true;
}
static bool readFastaRecord(GenomeO& g, const std::string& recordChrom) {  // :85-104
    if (g.next >= g.records.size()) {
        std::cout << "ERROR: Could not read fasta record at line: " << g.count + 1 << " from " << g.filename << std::endl;
        std::cout << "recordChrom: " << recordChrom << std::endl;
        std::cout << "Last chromName was: " << g.chromName << std::endl;
        return true;
    }
    const FastaRecord& r = g.records[g.next++];
    g.count += 1;
    g.chrom = &r.seq;
    std::string s = r.id.substr(0, r.id.find(' '));
    g.chromName = s.substr(0, s.find('\t'));
    return false;
}

static const int kMinBaseQAscii = 53, kMinMapQ = 60, kMaxClipped = 0, kMinAlignmentScore = 50;  // :20-22

static int alignmentScore(const Record& read) {  // :110-130
    std::vector<Tag> dict;
    buildTagsDict(read.tags, dict);
    const Tag* t = nullptr;
    for (auto& d : dict)
        if (d.key[0] == 'A' && d.key[1] == 'S') {
            t = &d;
            break;
        }
    if (!t) {
        std::cerr << "ERROR: Read " << read.qName << " has no AS tag." << std::endl;
        return -1;
    }
    long long v = 0;
    if (!extractInt(read.tags, *t, v)) {
        std::cerr << "ERROR: Could not read AS tag for read " << read.qName << std::endl;
        return -1;
    }
    return (int)v;
}
static int checkFlagsAndQuality(const Record& read) {  // :136-168, flags "1100xxxx000x"
    if (!(read.flag & 0x1)) return 0;
    if (!(read.flag & 0x2)) return 0;
    if (read.flag & 0x4) return 0;
    if (read.flag & 0x8) return 0;
    if (read.flag & 0x100) return 0;
    if (read.mapQ < kMinMapQ) return 0;
    int as = alignmentScore(read);
    if (as < 0) return -1;
    if (as < kMinAlignmentScore) return 0;
    uint32_t clipped = 0;
    for (auto& c : read.cigar)
        if (c.operation == 'S' || c.operation == 'H') clipped += c.count;
    if (clipped > (uint32_t)kMaxClipped) return 0;
    return 1;
}
static void countPosition(TripletCountsO& c, int base, const Record& read) {  // :174-189
    if (read.flag & 0x10) {
        if (read.flag & 0x40) c.reverseFirst[base] += 1;
        else c.reverseSecond[base] += 1;
    } else {
        if (read.flag & 0x40) c.forwardFirst[base] += 1;
        else c.forwardSecond[base] += 1;
    }
}
static void countBasesInTriplets(std::vector<TripletCountsO>& counts, const Record& read, const std::string& chrom) {  // :195-236
    size_t it = 0;
    if (read.cigar.empty() || read.seq.size() < 2) return;  // hazards in the reference
    size_t cigarCount = read.cigar[0].count - 1;
    size_t chromPos = (size_t)read.beginPos + 1;
    const size_t L = read.seq.size();
    for (size_t readPos = 1; readPos < L - 1; ++readPos, ++chromPos, --cigarCount) {
        while (cigarCount == 0) {
            ++it;
            if (it >= read.cigar.size()) {  // SEQAN_ASSERT compiled out in the reference: undefined read (hazard)
                std::cerr << "ORACLE: CIGAR exhausted inside triplet walk (hazard)\n";
                exit(3);
            }
            char op = read.cigar[it].operation;
            if (op == 'D' || op == 'N' || op == 'H' || op == 'P') chromPos += read.cigar[it].count;
            else if (op == 'S' || op == 'I') readPos += read.cigar[it].count;
            else cigarCount = read.cigar[it].count;
        }
        if (readPos >= L - 1) break;
        if (read.qual[readPos] < (char)kMinBaseQAscii) continue;
        int base = dna5(read.seq[readPos]);
        if (base == 4 || read.seq[readPos - 1] == 'N' || read.seq[readPos + 1] == 'N') continue;
        if (chromPos + 2 > chrom.size()) continue;  // reference reads past the contig end (hazard)
        // infix(chrom, chromPos-1, chromPos+2) as Dna5String -> DnaString: N -> A (R6)
        int c0 = dna5(chrom[chromPos - 1]) & 3, c1 = dna5(chrom[chromPos]) & 3, c2 = dna5(chrom[chromPos + 1]) & 3;
        // char vs Dna compares in char space (SeqAn CompareType): read flank must be exactly A/C/G/T
        if (read.seq[readPos - 1] != "ACGT"[c0]) continue;
        if (read.seq[readPos + 1] != "ACGT"[c2]) continue;
        countPosition(counts[(c0 << 4) + (c1 << 2) + c2], base, read);
    }
}
// :242-265; returns 1 on fatal error
static int tripletCounting(std::vector<TripletCountsO>& counts, const Record& record, const std::vector<std::string>& nameStore, GenomeO& genome) {
    int res = checkFlagsAndQuality(record);
    if (res == -1) return 1;
    if (res == 0) return 0;
    if (record.rID < 0 || (size_t)record.rID >= nameStore.size()) return 0;  // hazard
    const std::string& recordChrom = nameStore[record.rID];
    while (recordChrom != genome.chromName)
        if (readFastaRecord(genome, recordChrom)) return 1;
    countBasesInTriplets(counts, record, *genome.chrom);
    return 0;
}

// -----------------------------------------------------------------------------
// Counts  (src/bamqualcheck.cpp:14-38)
// -----------------------------------------------------------------------------
struct CountsO {
    OverallNumbersO all;
    QualityCheckO r1, r2;
    std::vector<TripletCountsO> tripletCounts;
    std::vector<std::vector<ReadQualityHasherO>> sps;
    explicit CountsO(const Options& opt) : r1(opt.isize), r2(opt.isize), tripletCounts(64) {
        sps.assign(opt.q_cutoff.size(), std::vector<ReadQualityHasherO>(opt.klist.size(), ReadQualityHasherO(opt)));
        for (size_t i = 0; i < sps.size(); i++)
            for (size_t j = 0; j < sps[i].size(); j++) {
                sps[i][j].setQualityCutoff(opt.q_cutoff[i]);
                sps[i][j].setK(opt.klist[j]);
            }
    }
};

// -----------------------------------------------------------------------------
// BGZF / BAM input (SeqAn BamStream stand-in)
// -----------------------------------------------------------------------------
static bool readFile(const std::string& path, std::vector<uint8_t>& out) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize((size_t)n);
    size_t got = n ? fread(out.data(), 1, (size_t)n, f) : 0;
    fclose(f);
    return got == (size_t)n;
}
static bool inflateBgzf(const std::vector<uint8_t>& in, std::vector<uint8_t>& out) {
    size_t p = 0;
    out.clear();
    while (p + 18 <= in.size()) {
        if (in[p] != 0x1f || in[p + 1] != 0x8b || in[p + 2] != 8 || !(in[p + 3] & 4)) return false;
        uint16_t xlen;
        memcpy(&xlen, &in[p + 10], 2);
        size_t x = p + 12, xend = x + xlen;
        int bsize = -1;
        while (x + 4 <= xend) {
            uint16_t slen;
            memcpy(&slen, &in[x + 2], 2);
            if (in[x] == 'B' && in[x + 1] == 'C' && slen == 2) {
                uint16_t b;
                memcpy(&b, &in[x + 4], 2);
                bsize = b;
            }
            x += 4 + slen;
        }
        if (bsize < 0) return false;
        size_t blockLen = (size_t)bsize + 1;
        if (p + blockLen > in.size()) return false;
        uint32_t isize;
        memcpy(&isize, &in[p + blockLen - 4], 4);
        size_t cbeg = p + 12 + xlen, clen = blockLen - 12 - xlen - 8;
        size_t o = out.size();
        out.resize(o + isize);
        if (isize) {
            z_stream zs;
            memset(&zs, 0, sizeof(zs));
            if (inflateInit2(&zs, -15) != Z_OK) return false;
            zs.next_in = (Bytef*)&in[cbeg];
            zs.avail_in = (uInt)clen;
            zs.next_out = &out[o];
            zs.avail_out = isize;
            int rc = inflate(&zs, Z_FINISH);
            inflateEnd(&zs);
            if (rc != Z_STREAM_END) return false;
        }
        p += blockLen;
    }
    return p == in.size();
}

struct BamInput {
    std::vector<uint8_t> data;  // inflated BAM stream
    size_t pos = 0;
    std::string headerText;
    std::vector<std::string> nameStore;
    bool atEnd() const { return pos >= data.size(); }
};
static inline int32_t rdI32(const uint8_t* p) {
    int32_t v;
    memcpy(&v, p, 4);
    return v;
}
// SAM text (`bamqualcheck ... -` or a .sam file: src/bamqualcheck.cpp:252-260, src/CommandLineParser.hpp:88-107).  The
// reference hands the text to SeqAn's BamStream, which fills the same BamAlignmentRecord as for BAM; this reader is an
// independent restatement of that (SAM specification sections 1.3-1.5): the text is turned into the inflated BAM
// stream that readRecord() below decodes, so both formats meet the statistics through one decoder.
template <typename T> static void putLE(std::vector<uint8_t>& v, T x) {
    uint8_t b[sizeof(T)];
    memcpy(b, &x, sizeof(T));
    v.insert(v.end(), b, b + sizeof(T));
}
static bool samToBam(std::istream& in, std::vector<uint8_t>& out) {
    std::string text, line;
    std::vector<std::string> names;
    std::vector<int32_t> lens;
    std::vector<std::string> body;
    while (std::getline(in, line)) {
        if (!line.empty() && line[line.size() - 1] == '\r') line.erase(line.size() - 1);
        if (line.empty()) continue;
        if (line[0] == '@' && body.empty()) {
            text += line + "\n";
            if (line.compare(0, 3, "@SQ") == 0) {
                std::string sn;
                long ln = 0;
                std::istringstream ls(line);
                std::string f;
                while (std::getline(ls, f, '\t')) {
                    if (f.compare(0, 3, "SN:") == 0) sn = f.substr(3);
                    if (f.compare(0, 3, "LN:") == 0) ln = atol(f.c_str() + 3);
                }
                names.push_back(sn);
                lens.push_back((int32_t)ln);
            }
        } else body.push_back(line);
    }
    out.clear();
    out.insert(out.end(), {'B', 'A', 'M', 1});
    putLE<int32_t>(out, (int32_t)text.size());
    out.insert(out.end(), text.begin(), text.end());
    putLE<int32_t>(out, (int32_t)names.size());
    for (size_t i = 0; i < names.size(); ++i) {
        putLE<int32_t>(out, (int32_t)names[i].size() + 1);
        out.insert(out.end(), names[i].begin(), names[i].end());
        out.push_back(0);
        putLE<int32_t>(out, lens[i]);
    }
    auto refId = [&](const std::string& n) -> int32_t {
        for (size_t i = 0; i < names.size(); ++i)
            if (names[i] == n) return (int32_t)i;
        return -1;
    };
    for (const std::string& l : body) {
        std::vector<std::string> f;
        size_t p = 0;
        for (;;) {
            size_t q = l.find('\t', p);
            f.push_back(l.substr(p, q == std::string::npos ? std::string::npos : q - p));
            if (q == std::string::npos) break;
            p = q + 1;
        }
        if (f.size() < 11) return false;
        const int32_t rid = f[2] == "*" ? -1 : refId(f[2]);
        const int32_t nrid = f[6] == "*" ? -1 : (f[6] == "=" ? rid : refId(f[6]));
        std::vector<uint32_t> cig;
        if (f[5] != "*") {
            uint32_t n = 0;
            for (char ch : f[5]) {
                if (ch >= '0' && ch <= '9') { n = n * 10 + (uint32_t)(ch - '0'); continue; }
                const char* ops = "MIDNSHP=X";
                const char* w = strchr(ops, ch);
                if (!w) return false;
                cig.push_back((n << 4) | (uint32_t)(w - ops));
                n = 0;
            }
        }
        const std::string seq = f[9] == "*" ? std::string() : f[9];
        std::vector<uint8_t> rec;
        putLE<int32_t>(rec, rid);
        putLE<int32_t>(rec, atoi(f[3].c_str()) - 1);
        rec.push_back((uint8_t)(f[0].size() + 1));
        rec.push_back((uint8_t)atoi(f[4].c_str()));
        putLE<uint16_t>(rec, 4680);
        putLE<uint16_t>(rec, (uint16_t)cig.size());
        putLE<uint16_t>(rec, (uint16_t)atoi(f[1].c_str()));
        putLE<int32_t>(rec, (int32_t)seq.size());
        putLE<int32_t>(rec, nrid);
        putLE<int32_t>(rec, atoi(f[7].c_str()) - 1);
        putLE<int32_t>(rec, atoi(f[8].c_str()));
        rec.insert(rec.end(), f[0].begin(), f[0].end());
        rec.push_back(0);
        for (uint32_t c : cig) putLE<uint32_t>(rec, c);
        static const char nt16[] = "=ACMGRSVTWYHKDBN";
        for (size_t i = 0; i < seq.size(); i += 2) {
            auto code = [&](char ch) -> uint8_t {
                const char* w = strchr(nt16, toupper((unsigned char)ch));
                return (uint8_t)(w && *w ? w - nt16 : 15);
            };
            rec.push_back((uint8_t)((code(seq[i]) << 4) | (i + 1 < seq.size() ? code(seq[i + 1]) : 0)));
        }
        if (f[10] == "*") rec.insert(rec.end(), seq.size(), 0xFF);
        else {
            if (f[10].size() != seq.size()) return false;
            for (char ch : f[10]) rec.push_back((uint8_t)(ch - 33));
        }
        for (size_t t = 11; t < f.size(); ++t) {
            const std::string& g = f[t];
            if (g.size() < 5 || g[2] != ':' || g[4] != ':') return false;
            const char ty = g[3];
            const std::string val = g.substr(5);
            rec.push_back((uint8_t)g[0]);
            rec.push_back((uint8_t)g[1]);
            if (ty == 'A') { rec.push_back('A'); rec.push_back((uint8_t)(val.empty() ? 0 : val[0])); }
            else if (ty == 'i') {  // the smallest BAM integer type that holds the value
                const long long v = atoll(val.c_str());
                if (v >= 0) {
                    if (v <= 255) { rec.push_back('C'); rec.push_back((uint8_t)v); }
                    else if (v <= 65535) { rec.push_back('S'); putLE<uint16_t>(rec, (uint16_t)v); }
                    else { rec.push_back('I'); putLE<uint32_t>(rec, (uint32_t)v); }
                } else {
                    if (v >= -128) { rec.push_back('c'); rec.push_back((uint8_t)(int8_t)v); }
                    else if (v >= -32768) { rec.push_back('s'); putLE<int16_t>(rec, (int16_t)v); }
                    else { rec.push_back('i'); putLE<int32_t>(rec, (int32_t)v); }
                }
            } else if (ty == 'f') { rec.push_back('f'); putLE<float>(rec, (float)atof(val.c_str())); }
            else if (ty == 'Z' || ty == 'H') { rec.push_back((uint8_t)ty); rec.insert(rec.end(), val.begin(), val.end()); rec.push_back(0); }
            else if (ty == 'B') {
                if (val.empty()) return false;
                const char sub = val[0];
                std::vector<std::string> items;
                std::istringstream vs(val.size() > 2 ? val.substr(2) : std::string());
                std::string it;
                while (std::getline(vs, it, ',')) items.push_back(it);
                rec.push_back('B');
                rec.push_back((uint8_t)sub);
                putLE<int32_t>(rec, (int32_t)items.size());
                for (const std::string& x : items) {
                    switch (sub) {
                        case 'c': rec.push_back((uint8_t)(int8_t)atoi(x.c_str())); break;
                        case 'C': rec.push_back((uint8_t)atoi(x.c_str())); break;
                        case 's': putLE<int16_t>(rec, (int16_t)atoi(x.c_str())); break;
                        case 'S': putLE<uint16_t>(rec, (uint16_t)atoi(x.c_str())); break;
                        case 'i': putLE<int32_t>(rec, (int32_t)atoll(x.c_str())); break;
                        case 'I': putLE<uint32_t>(rec, (uint32_t)atoll(x.c_str())); break;
                        case 'f': putLE<float>(rec, (float)atof(x.c_str())); break;
                        default: return false;
                    }
                }
            } else return false;
        }
        putLE<int32_t>(out, (int32_t)rec.size());
        out.insert(out.end(), rec.begin(), rec.end());
    }
    return true;
}

static bool openBam(const std::string& path, BamInput& b) {
    std::vector<uint8_t> raw;
    const bool sam = path == "-" || (path.size() > 4 && path.compare(path.size() - 4, 4, ".sam") == 0);
    if (sam) {
        if (path == "-") { std::cerr << "Reading from stdin" << std::endl; if (!samToBam(std::cin, raw)) return false; }
        else { std::ifstream f(path.c_str()); if (!f.good() || !samToBam(f, raw)) return false; }
    } else if (!readFile(path, raw)) return false;
    if (raw.size() >= 4 && memcmp(raw.data(), "BAM\1", 4) == 0) b.data.swap(raw);
    else if (!inflateBgzf(raw, b.data)) return false;
    const std::vector<uint8_t>& d = b.data;
    if (d.size() < 12 || memcmp(d.data(), "BAM\1", 4) != 0) return false;
    int32_t l_text = rdI32(&d[4]);
    b.headerText.assign((const char*)&d[8], (size_t)l_text);
    size_t p = 8 + (size_t)l_text;
    int32_t n_ref = rdI32(&d[p]);
    p += 4;
    for (int i = 0; i < n_ref; ++i) {
        int32_t l_name = rdI32(&d[p]);
        p += 4;
        b.nameStore.push_back(std::string((const char*)&d[p], (size_t)(l_name > 0 ? l_name - 1 : 0)));
        p += (size_t)l_name + 4;
    }
    b.pos = p;
    return true;
}
// readRecord(record, BamStream): BAM record decode (R1-R3)
static int readRecord(Record& r, BamInput& b) {
    const std::vector<uint8_t>& d = b.data;
    if (b.pos + 4 > d.size()) return 1;
    int32_t block_size = rdI32(&d[b.pos]);
    if (block_size < 32 || b.pos + 4 + (size_t)block_size > d.size()) return 1;
    const uint8_t* p = &d[b.pos + 4];
    b.pos += 4 + (size_t)block_size;
    r.rID = rdI32(p);
    r.beginPos = rdI32(p + 4);
    uint8_t l_read_name = p[8];
    r.mapQ = p[9];
    uint16_t n_cigar, flag;
    memcpy(&n_cigar, p + 12, 2);
    memcpy(&flag, p + 14, 2);
    r.flag = flag;
    int32_t l_seq = rdI32(p + 16);
    r.rNextId = rdI32(p + 20);
    r.pNext = rdI32(p + 24);
    r.tLen = rdI32(p + 28);
    size_t need = 32 + (size_t)l_read_name + 4 * (size_t)n_cigar + ((size_t)l_seq + 1) / 2 + (size_t)l_seq;
    if (l_seq < 0 || need > (size_t)block_size) return 1;
    const uint8_t* q = p + 32;
    r.qName.assign((const char*)q, l_read_name ? l_read_name - 1 : 0);
    q += l_read_name;
    r.cigar.resize(n_cigar);
    static const char ops[] = "MIDNSHP=X";
    for (unsigned i = 0; i < n_cigar; ++i) {
        uint32_t v;
        memcpy(&v, q + 4 * i, 4);
        unsigned op = v & 15;
        r.cigar[i].operation = op < 9 ? ops[op] : '?';
        r.cigar[i].count = v >> 4;
    }
    q += 4 * (size_t)n_cigar;
    static const char nt16[] = "=ACMGRSVTWYHKDBN";
    r.seq.resize((size_t)l_seq);
    for (int i = 0; i < l_seq; ++i) r.seq[i] = nt16[(q[i >> 1] >> ((~i & 1) << 2)) & 15];
    q += ((size_t)l_seq + 1) / 2;
    r.qual.resize((size_t)l_seq);
    for (int i = 0; i < l_seq; ++i) r.qual[i] = (char)(q[i] + 33);
    q += l_seq;
    r.tags.assign((const char*)q, (size_t)block_size - need);
    return 0;
}

// -----------------------------------------------------------------------------
// main-loop helpers  (src/bamqualcheck.cpp:44-123)
// -----------------------------------------------------------------------------
static void getSampleIdAndLaneNames(std::string& id, std::map<std::string, unsigned>& laneNames, const std::string& headerText) {  // :44-66
    std::istringstream hs(headerText);
    std::string line;
    while (std::getline(hs, line)) {
        if (line.compare(0, 3, "@RG") != 0) continue;
        std::istringstream ls(line);
        std::string field;
        bool firstField = true;
        while (std::getline(ls, field, '\t')) {
            if (firstField) {
                firstField = false;
                continue;
            }
            if (field.size() < 3 || field[2] != ':') continue;
            std::string key = field.substr(0, 2), value = field.substr(3);
            if (key == "ID") {
                unsigned l = (unsigned)laneNames.size();
                laneNames[value] = l;
            }
            if (key == "SM") id = value;
        }
    }
}
static std::set<int> initChroms(const Options& opt, const std::vector<std::string>& nameStore) {  // :106-123
    std::set<int> chrIdset;
    std::istringstream cs(opt.chroms);
    std::string name;
    while (std::getline(cs, name, ',')) {
        for (size_t i = 0; i < nameStore.size(); ++i)
            if (nameStore[i] == name) {
                chrIdset.insert((int)i);
                break;
            }
    }
    return chrIdset;
}
// :72-100.  Returns -1 on error, else lane index.  A record without RG falls off the end of the
// reference function (undefined return value); the oracle treats it as fatal (-2).
static long getLane(const Record& record, const std::vector<Tag>& dict, std::map<std::string, unsigned>& laneNames) {
    for (unsigned tagid = 0; tagid < dict.size(); ++tagid) {
        if (dict[tagid].key[0] == 'R' && dict[tagid].key[1] == 'G') {
            if (dict[tagid].type == 'Z') {
                std::string readlane(record.tags.data() + dict[tagid].valueBegin, record.tags.data() + dict[tagid].valueEnd);
                if (!readlane.empty() && readlane.back() == '\0') readlane.pop_back();
                return laneNames[readlane];
            } else {
                std::cout << "Read does not have Z" << "\n";
                return -1;
            }
        }
    }
    return -2;
}

template <typename T>
static void printString(const std::vector<T>& v, std::ostream& out) {  // :130-139
    for (auto& x : v) out << " " << x;
    out << std::endl;
}

static void writeTripletCounts(std::ostream& out, const std::vector<TripletCountsO>& counts) {  // TripletCounting.hpp:271-301
    const char bases[] = {'A', 'C', 'G', 'T'};
    for (size_t index = 0; index < 4; ++index) {
        char base = bases[index];
        out << "triplet_counts_" << base << "_1st_FW";
        for (auto& c : counts) out << " " << c.forwardFirst[index];
        out << std::endl;
        out << "triplet_counts_" << base << "_1st_RC";
        for (auto& c : counts) out << " " << c.reverseFirst[index];
        out << std::endl;
        out << "triplet_counts_" << base << "_2nd_FW";
        for (auto& c : counts) out << " " << c.forwardSecond[index];
        out << std::endl;
        out << "triplet_counts_" << base << "_2nd_RC";
        for (auto& c : counts) out << " " << c.reverseSecond[index];
        out << std::endl;
    }
}

static void writeOutput(std::ostream& outFile, const std::string& sampleId, std::map<std::string, unsigned>& laneNames, std::vector<CountsO>& counts, const std::vector<size_t>& q_cutoff, const std::vector<int>& klist) {  // bamqualcheck.cpp:156-233
    for (auto it = laneNames.begin(); it != laneNames.end(); ++it) {
        outFile << "sample_id " << sampleId << std::endl;
        outFile << "lane " << it->first << std::endl;
        CountsO& c = counts[it->second];
        c.r1.avgQualPerPos();
        c.r2.avgQualPerPos();
        outFile << "total_read_pairs " << c.all.readcount / 2 << std::endl;
        outFile << "total_bps " << c.all.totalbps << std::endl;
        outFile << "supplementary_alignments " << c.all.supplementary << std::endl;
        outFile << "marked_duplicate " << c.all.duplicates << std::endl;
        outFile << "QC_failed " << c.all.QCfailed << std::endl;
        outFile << "not_primary_alignment " << c.all.not_primary_alignment << std::endl;
        outFile << "both_reads_unmapped " << c.all.bothunmapped << std::endl;
        outFile << "first_read_unmapped " << c.all.firstunmapped << std::endl;
        outFile << "second_read_unmapped " << c.all.secondunmapped << std::endl;
        outFile << "first_and_or_second_read_mapped " << c.all.first_and_or_second_mapped << std::endl;
        outFile << "FF_RR_oriented_pairs " << c.all.FF_RR_orientation << std::endl;
        outFile << "total_proper_pairs " << c.all.properpair_count << std::endl;
        outFile << "total_proper_pairs_autosome " << c.all.auto_properpair_count << std::endl;
        outFile << "genome_coverage_histogram"; printString(c.all.poscov, outFile);
        outFile << "insert_size_histogram"; printString(c.r1.insertSize, outFile);
        outFile << "read_length_histogram_first"; printString(c.r1.readLength, outFile);
        outFile << "read_length_histogram_second"; printString(c.r2.readLength, outFile);
        outFile << "N_count_histogram_first"; printString(c.r1.Ncount, outFile);
        outFile << "N_count_histogram_second"; printString(c.r2.Ncount, outFile);
        outFile << "GC_content_histogram_first"; printString(c.r1.GCcount, outFile);
        outFile << "GC_content_histogram_second"; printString(c.r2.GCcount, outFile);
        outFile << "average_base_qual_histogram_first"; printString(c.r1.averageQual, outFile);
        outFile << "average_base_qual_histogram_second"; printString(c.r2.averageQual, outFile);
        outFile << "mapping_qual_histogram_first"; printString(c.r1.mapQ, outFile);
        outFile << "mapping_qual_histogram_second"; printString(c.r2.mapQ, outFile);
        outFile << "mismatch_count_histogram_first"; printString(c.r1.mismatch, outFile);
        outFile << "mismatch_count_histogram_second"; printString(c.r2.mismatch, outFile);
        outFile << "deletion_count_histogram_first"; printString(c.r1.delhist, outFile);
        outFile << "deletion_count_histogram_second"; printString(c.r2.delhist, outFile);
        outFile << "insertion_count_histogram_first"; printString(c.r1.inshist, outFile);
        outFile << "insertion_count_histogram_second"; printString(c.r2.inshist, outFile);
        outFile << "Ns_by_position_first"; printString(c.r1.dnacount[4], outFile);
        outFile << "Ns_by_position_second"; printString(c.r2.dnacount[4], outFile);
        outFile << "As_by_position_first"; printString(c.r1.dnacount[0], outFile);
        outFile << "As_by_position_second"; printString(c.r2.dnacount[0], outFile);
        outFile << "Cs_by_position_first"; printString(c.r1.dnacount[1], outFile);
        outFile << "Cs_by_position_second"; printString(c.r2.dnacount[1], outFile);
        outFile << "Gs_by_position_first"; printString(c.r1.dnacount[2], outFile);
        outFile << "Gs_by_position_second"; printString(c.r2.dnacount[2], outFile);
        outFile << "Ts_by_position_first"; printString(c.r1.dnacount[3], outFile);
        outFile << "Ts_by_position_second"; printString(c.r2.dnacount[3], outFile);
        outFile << "average_base_qual_by_position_first"; printString(c.r1.avgqualcount, outFile);
        outFile << "average_base_qual_by_position_second"; printString(c.r2.avgqualcount, outFile);
        outFile << "soft_clipping_5_prime_by_position_first"; printString(c.r1.scposcount_5prime, outFile);
        outFile << "soft_clipping_3_prime_by_position_first"; printString(c.r1.scposcount_3prime, outFile);
        outFile << "soft_clipping_5_prime_by_position_second"; printString(c.r2.scposcount_5prime, outFile);
        outFile << "soft_clipping_3_prime_by_position_second"; printString(c.r2.scposcount_3prime, outFile);
        c.all.ten_most_abundant_kmers(outFile);
        outFile << "8mer_count"; printString(c.all.eightmercount, outFile);
        for (size_t i = 0; i < c.sps.size(); i++)
            for (size_t j = 0; j < c.sps[i].size(); j++) {
                outFile << klist[j] << "mer_count_after_qual_clipping_" << q_cutoff[i] << " " << c.sps[i][j].sc.get_sumCount() << std::endl;
                outFile << "distinct_" << klist[j] << "mer_count_after_qual_clipping_" << q_cutoff[i] << " " << c.sps[i][j].sc.F0() << std::endl;
                outFile << "unique_" << klist[j] << "mer_count_after_qual_clipping_" << q_cutoff[i] << " " << c.sps[i][j].sc.f1() << std::endl;
                outFile << klist[j] << "mer_F2_after_qual_clipping_" << q_cutoff[i] << " " << c.sps[i][j].sc.F2() << std::endl;
            }
        writeTripletCounts(outFile, c.tripletCounts);
    }
}

// oracle-only raw dump: text lines "name lane n v0 v1 ..." plus "<dump>.sketch" (binary u64:
// for each lane, q, k: sketch table then F2 table).
static void writeDump(const Options& opt, std::map<std::string, unsigned>& laneNames, std::vector<CountsO>& counts) {
    std::ofstream d(opt.dumpFile.c_str());
    std::ofstream sk((opt.dumpFile + ".sketch").c_str(), std::ios::binary);
    for (auto it = laneNames.begin(); it != laneNames.end(); ++it) {
        unsigned l = it->second;
        CountsO& c = counts[l];
        d << "cov_lost " << l << " 1 " << c.all.covLost << "\n";
        d << "qualcount_first " << l << " " << c.r1.qualcount.size();
        for (auto v : c.r1.qualcount) d << " " << v;
        d << "\nqualcount_second " << l << " " << c.r2.qualcount.size();
        for (auto v : c.r2.qualcount) d << " " << v;
        d << "\nqualcount_readnr " << l << " 2 " << c.r1.qualcount_readnr << " " << c.r2.qualcount_readnr << "\n";
        for (size_t i = 0; i < c.sps.size(); i++)
            for (size_t j = 0; j < c.sps[i].size(); j++) {
                const StreamCounterO& s = c.sps[i][j].sc;
                d << "sketch_geometry " << l << " 3 " << s.size << " " << s.F2size << " " << s.sumCount << "\n";
                sk.write((const char*)s.table.data(), (std::streamsize)(s.table.size() * 8));
                sk.write((const char*)s.F2table.data(), (std::streamsize)(s.F2table.size() * 8));
            }
    }
}

// -----------------------------------------------------------------------------
// run(): main  (src/bamqualcheck.cpp:239-457)
// -----------------------------------------------------------------------------
static int run(const Options& opt, double* seconds_loop = nullptr, uint64_t* n_records = nullptr) {
    BamInput inStream;
    if (!openBam(opt.bamFile, inStream)) {
        std::cerr << "ERROR: Could not open " << opt.bamFile << " for reading.\n";
        return 1;
    }
    std::ofstream outFile(opt.outputFile.c_str(), std::ios::out | std::ios::binary);
    if (!outFile.good()) {
        std::cerr << "ERROR: Could not open output file " << opt.outputFile << '\n';
        return 1;
    }
    std::string sampleId;
    std::map<std::string, unsigned> laneNames;
    getSampleIdAndLaneNames(sampleId, laneNames, inStream.headerText);

    GenomeO genome;
    genome.filename = opt.referenceFile;
    loadFasta(genome);  // the reference ignores openFastaFile's result (:291) and fails later on first use
    std::set<int> chrIdset = initChroms(opt, inStream.nameStore);

    unsigned lanecount = (unsigned)laneNames.size();
    std::vector<CountsO> counts(lanecount, CountsO(opt));

    auto t0 = std::chrono::steady_clock::now();
    uint64_t nrec = 0;
    Record record;
    std::vector<Tag> tagsDict;
    while (!inStream.atEnd()) {
        if (opt.maxRecords >= 0 && (long)nrec >= opt.maxRecords) break;
        if (readRecord(record, inStream) != 0) {
            std::cerr << "ERROR: Could not read record from BAM File " << opt.bamFile << "\n";
            return 1;
        }
        ++nrec;
        buildTagsDict(record.tags, tagsDict);
        long ll = getLane(record, tagsDict, laneNames);
        if (ll < 0) return 1;
        unsigned l = (unsigned)ll;
        if (l >= counts.size()) {  // RG not declared in the header: reference aliases/overruns (hazard)
            std::cerr << "ORACLE: read group not in header (hazard)\n";
            return 3;
        }
        CountsO& C = counts[l];
        if (record.flag & 0x800) { C.all.supplementary += 1; continue; }
        if (record.flag & 0x100) { C.all.not_primary_alignment += 1; continue; }
        if (record.flag & 0x400) C.all.duplicates += 1;
        if (record.flag & 0x200) C.all.QCfailed += 1;

        if (!(record.flag & 0x400) && !(record.flag & 0x200))
            if (tripletCounting(C.tripletCounts, record, inStream.nameStore, genome) != 0) return 1;

        if (record.flag & 0x10) {  // :345-350
            std::reverse(record.seq.begin(), record.seq.end());
            for (auto& ch : record.seq) ch = complementChar(ch);
            std::reverse(record.qual.begin(), record.qual.end());
            std::reverse(record.cigar.begin(), record.cigar.end());
        }
        C.all.readcount += 1;
        C.all.totalbps += record.seq.size();
        if (record.flag & 0x40) {
            C.r1.check_read_len(record.seq, record.qual);
            C.r1.get_count(record.seq, record.qual);
            if (record.flag & 0x4) {
                C.all.firstunmapped += 1;
                if (record.flag & 0x8) C.all.bothunmapped += 1;
            }
            if (record.flag & 0x2) {
                C.all.properpair_count += 1;
                bool rc = record.flag & 0x10, nrc = record.flag & 0x20;
                if ((!rc && !nrc) || (rc && nrc)) C.all.FF_RR_orientation += 1;
            }
        } else if (record.flag & 0x80) {
            C.r2.check_read_len(record.seq, record.qual);
            C.r2.get_count(record.seq, record.qual);
            if (record.flag & 0x4) C.all.secondunmapped += 1;
        } else {
            std::cerr << "ERROR: No first or second flag in read in:  " << opt.bamFile << "\n";
            return 1;
        }
        if (chrIdset.count(record.rID) != 0) {
            if (record.flag & 0x40) {
                if (!(record.flag & 0x4)) {
                    C.r1.cigar_count(record);
                    C.r1.map_Q(record.mapQ);
                    C.r1.mis_match(record.tags, tagsDict);
                    if (!(record.flag & 0x8))
                        if (chrIdset.count(record.rNextId) != 0) C.r1.insert_size(record.tLen);
                }
                if ((!(record.flag & 0x4) || !(record.flag & 0x8)) && !(record.flag & 0x400)) C.all.first_and_or_second_mapped += 1;
                if ((record.flag & 0x2) && !(record.flag & 0x400)) C.all.auto_properpair_count += 1;
            } else if (record.flag & 0x80) {
                if (!(record.flag & 0x4)) {
                    C.r2.cigar_count(record);
                    C.r2.map_Q(record.mapQ);
                    C.r2.mis_match(record.tags, tagsDict);
                }
            }
            if (!(record.flag & 0x4) && !(record.flag & 0x400)) C.all.coverage(record);
        }
        C.all.count8mers(record.seq);
        if (!(record.flag & 0x200) && !(record.flag & 0x400))
            for (auto& row : C.sps)
                for (auto& h : row) h(record.seq.c_str(), record.seq.size(), record.qual.c_str(), record.qual.size());
    }
    auto t1 = std::chrono::steady_clock::now();
    if (seconds_loop) *seconds_loop = std::chrono::duration<double>(t1 - t0).count();
    if (n_records) *n_records = nrec;

    for (auto it = laneNames.begin(); it != laneNames.end(); ++it) {  // :447-453
        unsigned lid = it->second;
        counts[lid].all.update_coverage();
        counts[lid].all.update_vectors();
        counts[lid].all.update_coverage();
    }
    writeOutput(outFile, sampleId, laneNames, counts, opt.q_cutoff, opt.klist);
    if (!opt.dumpFile.empty()) writeDump(opt, laneNames, counts);
    return 0;
}

static void parseList(const std::string& s, std::vector<int>& out) {  // CommandLineParser.hpp:132-141
    std::stringstream sk(s);
    int i;
    while (sk >> i) {
        out.push_back(i);
        if (sk.peek() == ',') sk.ignore();
    }
}
static void parseListQ(const std::string& s, std::vector<size_t>& out) {  // :121-130
    std::stringstream sq(s);
    size_t j;
    while (sq >> j) {
        out.push_back(j);
        if (sq.peek() == ',') sq.ignore();
    }
}

}  // namespace oracle

// -----------------------------------------------------------------------------
// C entry points used by tests (ctypes) for the kmerstream restatement
// -----------------------------------------------------------------------------
extern "C" {
// hash of every k-window of s[0..l) computed by init + rolling update
int oracle_rephash_windows(int seed, int k, const char* s, int l, uint64_t* out) {
    oracle::RepHashO hf;
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
void oracle_rephash_hvals(int seed, uint64_t* out64) {
    oracle::RepHashO hf;
    hf.seed(seed);
    for (int i = 0; i < 32; ++i) {
        out64[2 * i] = hf.hvals[i].hi;
        out64[2 * i + 1] = hf.hvals[i].lo;
    }
}
// feed n hashes into a StreamCounter(e, seed) and report {sumCount, F0, f1, F2, size, F2size}
void oracle_streamcounter(double e, int seed, const uint64_t* hashes, uint64_t n, uint64_t* out6, uint64_t* table_out, uint64_t* f2_out) {
    oracle::StreamCounterO sc(e, seed);
    for (uint64_t i = 0; i < n; ++i) sc(hashes[i]);
    out6[0] = sc.get_sumCount();
    out6[1] = sc.F0();
    out6[2] = sc.f1();
    out6[3] = sc.F2();
    out6[4] = sc.size;
    out6[5] = sc.F2size;
    if (table_out) memcpy(table_out, sc.table.data(), sc.table.size() * 8);
    if (f2_out) memcpy(f2_out, sc.F2table.data(), sc.F2table.size() * 8);
}
// ReadQualityHasher over a set of reads (concatenated, with lengths)
void oracle_hasher(double e, int seed, int q, int k, const char* seqs, const char* quals, const int* lens, int nreads, uint64_t* out4, uint64_t* table_out, uint64_t* f2_out) {
    oracle::Options opt;
    opt.e = e;
    opt.seed = seed;
    oracle::ReadQualityHasherO h(opt);
    h.setQualityCutoff((size_t)q);
    h.setK((size_t)k);
    size_t off = 0;
    for (int i = 0; i < nreads; ++i) {
        h(seqs + off, (size_t)lens[i], quals + off, (size_t)lens[i]);
        off += (size_t)lens[i];
    }
    out4[0] = h.sc.get_sumCount();
    out4[1] = h.sc.F0();
    out4[2] = h.sc.f1();
    out4[3] = h.sc.F2();
    if (table_out) memcpy(table_out, h.sc.table.data(), h.sc.table.size() * 8);
    if (f2_out) memcpy(f2_out, h.sc.F2table.data(), h.sc.F2table.size() * 8);
}
uint64_t oracle_bitscan(uint64_t v) { return oracle::bitScanForward(v); }
}

#ifndef ORACLE_NO_MAIN
int main(int argc, char** argv) {
    oracle::Options opt;
    std::string kmer = "32", qcut = "17";
    bool haveR = false, haveO = false, timing = false;
    std::vector<std::string> positional;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto need = [&](const char* name) -> std::string {
            if (i + 1 >= argc) {
                std::cerr << "bamqualcheck: option requires an argument -- " << name << std::endl;
                exit(1);
            }
            return argv[++i];
        };
        if (a == "-r" || a == "--reference") { opt.referenceFile = need("r"); haveR = true; }
        else if (a == "-i" || a == "--insert-size") opt.isize = atoi(need("i").c_str());
        else if (a == "-c" || a == "--chromosomes") opt.chroms = need("c");
        else if (a == "-o" || a == "--output-file") { opt.outputFile = need("o"); haveO = true; }
        else if (a == "-k" || a == "--kmer-size") kmer = need("k");
        else if (a == "-q" || a == "--quality-cutoff") qcut = need("q");
        else if (a == "-e" || a == "--error-rate") opt.e = atof(need("e").c_str());
        else if (a == "-s" || a == "--seed") opt.seed = atoi(need("s").c_str());
        else if (a == "--dump") opt.dumpFile = need("dump");
        else if (a == "--max-records") opt.maxRecords = atol(need("max-records").c_str());
        else if (a == "--timing") timing = true;
        else positional.push_back(a);
    }
    if (!haveR || !haveO || positional.size() != 1) {
        std::cerr << "bamqualcheck: -r, -o and one BAMFILE are required" << std::endl;
        return 1;
    }
    opt.bamFile = positional[0];
    oracle::parseListQ(qcut, opt.q_cutoff);
    oracle::parseList(kmer, opt.klist);
    double secs = 0;
    uint64_t nrec = 0;
    int rc = oracle::run(opt, &secs, &nrec);
    if (timing) fprintf(stderr, "ORACLE_TIMING records=%llu loop_seconds=%.6f\n", (unsigned long long)nrec, secs);
    return rc;
}
#endif
