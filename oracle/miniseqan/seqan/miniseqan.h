// miniseqan.h -- TEST INFRASTRUCTURE.  A minimal stand-in for the parts of SeqAn 1.4.2 that
// DecodeGenetics/BamQC uses, so that the reference's OWN sources (src/bamqualcheck.cpp, OverallNumbers.hpp,
// QualityCheck.hpp, TripletCounting.hpp, ReadQualityHasher.hpp, CommandLineParser.hpp, kmerstream/*) can be
// compiled unmodified into oracle/_ref/bamqualcheck_ref (see oracle/Makefile).  SeqAn itself is not
// vendored by the reference and not available offline.
//
// What this file is NOT: it is not SeqAn.  Containers, alphabets, BAM/FASTA readers and the argument parser
// are re-implemented from the SAM/BAM specification and from the SeqAn behaviours listed in SURVEY.md
// Appendix C (R1-R15); every statistic, gate and output line comes from the reference's own code.  The
// resulting binary pins the oracle's restatement of that code; the SeqAn boundary stays "parity unpinned".
#ifndef MINISEQAN_H_
#define MINISEQAN_H_

#include <zlib.h>

#include <math.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <vector>
#include <stdint.h>

typedef unsigned char __uint8;
typedef unsigned int __uint32;
typedef int64_t __int64;
typedef uint64_t __uint64;

#define SEQAN_ASSERT_NEQ(a, b) ((void)0)

namespace seqan {

// ---------------------------------------------------------------------------------------------- tags
struct Standard {};
struct Rooted {};
struct Sam {};
struct Bam {};

// ------------------------------------------------------------------------------------------ alphabets
struct Dna5;
struct Dna {  // R5: char -> Dna: C 1, G 2, T/U 3, everything else 0 (A)
    unsigned char value;
    Dna() : value(0) {}
    Dna(char c) : value(fromChar(c)) {}
    inline Dna(Dna5 const& o);
    static unsigned char fromChar(char c) {
        switch (c) {
            case 'C': case 'c': return 1;
            case 'G': case 'g': return 2;
            case 'T': case 't': case 'U': case 'u': return 3;
            default: return 0;
        }
    }
    operator int() const { return value; }  // ordinal (contextToIndex, array subscripts)
    char letter() const { return "ACGT"[value & 3]; }
};
struct Dna5 {  // R5: char -> Dna5: ACGT/acgt 0-3, everything else 4 (N)
    unsigned char value;
    Dna5() : value(0) {}
    Dna5(char c) : value(fromChar(c)) {}
    Dna5(Dna const& o) : value(o.value) {}
    static unsigned char fromChar(char c) {
        switch (c) {
            case 'A': case 'a': return 0;
            case 'C': case 'c': return 1;
            case 'G': case 'g': return 2;
            case 'T': case 't': return 3;
            default: return 4;
        }
    }
    operator int() const { return value; }
    char letter() const { return "ACGTN"[value > 4 ? 4 : value]; }
};
inline Dna::Dna(Dna5 const& o) : value(o.value & 3) {}  // R6: Dna5 -> Dna keeps the low two bits, N -> A
// comparisons between a SimpleType and a char happen in char space (SeqAn CompareType<SimpleType, T> = T)
inline bool operator==(Dna5 const& a, char b) { return a.letter() == b; }
inline bool operator!=(Dna5 const& a, char b) { return a.letter() != b; }
inline bool operator==(char a, Dna const& b) { return a == b.letter(); }
inline bool operator!=(char a, Dna const& b) { return a != b.letter(); }
inline std::ostream& operator<<(std::ostream& o, Dna const& d) { return o << d.letter(); }
inline std::ostream& operator<<(std::ostream& o, Dna5 const& d) { return o << d.letter(); }
template <typename T> inline unsigned ordValue(T const& c) { return (unsigned)(unsigned char)c; }
inline unsigned ordValue(Dna const& c) { return c.value; }
inline unsigned ordValue(Dna5 const& c) { return c.value; }

// -------------------------------------------------------------------------------------------- String
template <typename T>
class String {
   public:
    std::vector<T> d;
    String() {}
    String(const char* s) { for (; *s; ++s) d.push_back(T(*s)); }
    String(std::string const& s) { for (size_t i = 0; i < s.size(); ++i) d.push_back(T(s[i])); }
    template <typename U> String(String<U> const& o) { d.reserve(o.d.size()); for (size_t i = 0; i < o.d.size(); ++i) d.push_back(T(o.d[i])); }
    String& operator=(const char* s) { d.clear(); for (; *s; ++s) d.push_back(T(*s)); return *this; }
    String& operator=(std::string const& s) { d.clear(); for (size_t i = 0; i < s.size(); ++i) d.push_back(T(s[i])); return *this; }
    String& operator=(char c) { d.assign(1, T(c)); return *this; }
    template <typename U> String& operator=(String<U> const& o) { d.clear(); d.reserve(o.d.size()); for (size_t i = 0; i < o.d.size(); ++i) d.push_back(T(o.d[i])); return *this; }
    // unchecked, like SeqAn; capacity is kept generous (R9) so that the reference's one-past-the-window write
    // in OverallNumbers::coverage lands in slack exactly as it does with SeqAn's Generous allocation
    T& operator[](size_t i) { return d.data()[i]; }
    T const& operator[](size_t i) const { return d.data()[i]; }
    mutable std::string cstr_;
};
typedef String<char> CharString;
typedef String<Dna> DnaString;
typedef String<Dna5> Dna5String;

template <typename T> inline size_t length(String<T> const& s) { return s.d.size(); }
inline size_t length(std::string const& s) { return s.size(); }
template <typename T> inline void generous(String<T>& s, size_t n) { size_t cap = n < 32 ? 32 : n + n / 2; if (s.d.capacity() < cap) s.d.reserve(cap); }
// growth by copy construction only (SeqAn's valueConstruct): element types need not be assignable
template <typename T> inline void resize(String<T>& s, size_t n) { generous(s, n); while (s.d.size() > n) s.d.pop_back(); while (s.d.size() < n) s.d.push_back(T()); }
template <typename T, typename V> inline void resize(String<T>& s, size_t n, V const& v) { generous(s, n); while (s.d.size() > n) s.d.pop_back(); while (s.d.size() < n) s.d.push_back(T(v)); }
template <typename T> inline void clear(String<T>& s) { s.d.clear(); }
template <typename T> inline bool empty(String<T> const& s) { return s.d.empty(); }
template <typename T> inline void swap(String<T>& a, String<T>& b) { a.d.swap(b.d); }
template <typename T> inline void reverse(String<T>& s) { std::reverse(s.d.begin(), s.d.end()); }
template <typename T, typename V> inline void appendValue(String<T>& s, V const& v) { s.d.push_back(T(v)); }
inline const char* toCString(CharString const& s) { s.cstr_.assign(s.d.begin(), s.d.end()); return s.cstr_.c_str(); }
inline bool operator==(CharString const& a, CharString const& b) { return a.d == b.d; }
inline bool operator!=(CharString const& a, CharString const& b) { return a.d != b.d; }
inline bool operator<(CharString const& a, CharString const& b) { return a.d < b.d; }  // R12: lexicographic
inline bool operator==(CharString const& a, const char* b) { size_t n = strlen(b); return a.d.size() == n && memcmp(a.d.data(), b, n) == 0; }
inline bool operator!=(CharString const& a, const char* b) { return !(a == b); }
template <typename T> inline std::ostream& operator<<(std::ostream& o, String<T> const& s) { for (size_t i = 0; i < s.d.size(); ++i) o << s.d[i]; return o; }
template <typename T> inline String<T> infix(String<T> const& s, size_t b, size_t e) { String<T> r; if (e > s.d.size()) e = s.d.size(); if (b < e) r.d.assign(s.d.begin() + b, s.d.begin() + e); return r; }
template <typename A, typename B> inline bool isEqual(A const& a, B const& b) { return a == b; }
inline bool isEqual(const char* a, const char* b) { return strcmp(a, b) == 0; }
inline bool isEqual(std::string const& a, const char* b) { return a == b; }
template <typename A, typename B> inline bool isNotEqual(A const& a, B const& b) { return a != b; }

template <typename T> struct RootedIter {  // a class type so that goNext() is found by argument-dependent lookup
    T* p;
    RootedIter(T* q = 0) : p(q) {}
    T& operator*() const { return *p; }
    bool operator!=(RootedIter const& o) const { return p != o.p; }
    bool operator==(RootedIter const& o) const { return p == o.p; }
};
template <typename T> inline void goNext(RootedIter<T>& it) { ++it.p; }
template <typename TString, typename TSpec = Standard> struct Iterator;
template <typename T, typename TSpec> struct Iterator<String<T>, TSpec> { typedef T* Type; };
template <typename T> struct Iterator<String<T>, Rooted> { typedef RootedIter<T> Type; };
template <typename T> inline T* begin(String<T>& s) { return s.d.data(); }
template <typename T> inline T* end(String<T>& s) { return s.d.data() + s.d.size(); }
template <typename T, typename Tag> inline T* begin(String<T>& s, Tag) { return s.d.data(); }
template <typename T, typename Tag> inline T* end(String<T>& s, Tag) { return s.d.data() + s.d.size(); }
template <typename T> inline void goNext(T*& it) { ++it; }

inline char complementIupac(char c) {  // R7
    switch (c) {
        case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
        case 'M': return 'K'; case 'K': return 'M'; case 'R': return 'Y'; case 'Y': return 'R';
        case 'V': return 'B'; case 'B': return 'V'; case 'H': return 'D'; case 'D': return 'H';
        default: return c;
    }
}
inline void reverseComplement(CharString& s) { std::reverse(s.d.begin(), s.d.end()); for (size_t i = 0; i < s.d.size(); ++i) s.d[i] = complementIupac(s.d[i]); }

// ----------------------------------------------------------------------------------------- StringSet
template <typename TString>
class StringSet {
   public:
    std::vector<TString> d;
    TString& operator[](size_t i) { return d[i]; }
    TString const& operator[](size_t i) const { return d[i]; }
};
template <typename T> inline size_t length(StringSet<T> const& s) { return s.d.size(); }
template <typename T> inline void resize(StringSet<T>& s, size_t n) { s.d.resize(n); }
template <typename T, typename TSpec> struct Iterator<StringSet<T>, TSpec> { typedef T* Type; };
template <typename T, typename Tag> inline T* begin(StringSet<T>& s, Tag) { return s.d.data(); }
template <typename T, typename Tag> inline T* end(StringSet<T>& s, Tag) { return s.d.data() + s.d.size(); }
inline void strSplit(StringSet<CharString>& out, CharString const& s, char sep) {  // R13
    out.d.clear();
    CharString cur;
    for (size_t i = 0; i < s.d.size(); ++i) {
        if (s.d[i] == sep) { out.d.push_back(cur); cur.d.clear(); }
        else cur.d.push_back(s.d[i]);
    }
    out.d.push_back(cur);
}
template <typename TId> inline bool getIdByName(StringSet<CharString> const& store, CharString const& name, TId& id) {
    for (size_t i = 0; i < store.d.size(); ++i)
        if (store.d[i] == name) { id = (TId)i; return true; }
    return false;
}

// --------------------------------------------------------------------------------------------- Pair
template <typename A, typename B> struct Pair { A i1; B i2; };
template <typename A, typename B> inline A& getValueI1(Pair<A, B>& p) { return p.i1; }
template <typename A, typename B> inline B& getValueI2(Pair<A, B>& p) { return p.i2; }

// -------------------------------------------------------------------------------------------- Shape
template <unsigned Q> struct UngappedShape {};
template <typename TValue, typename TSpec> struct Shape;
template <typename TValue, unsigned Q>
struct Shape<TValue, UngappedShape<Q> > {  // R8: big-endian base-4 hash of an ungapped q-gram
    uint64_t hValue, leftFactor;
    unsigned leftChar;
    Shape() : hValue(0), leftFactor(1), leftChar(0) { for (unsigned i = 1; i < Q; ++i) leftFactor *= 4; }
};
template <typename TValue, unsigned Q> inline unsigned length(Shape<TValue, UngappedShape<Q> > const&) { return Q; }
template <typename TValue, unsigned Q, typename TIter>
inline void hashInit(Shape<TValue, UngappedShape<Q> >& me, TIter it) {
    me.hValue = 0;
    me.leftChar = 0;
    for (unsigned i = 0; i + 1 < Q; ++i) me.hValue = me.hValue * 4 + ordValue(it[i]);
}
template <typename TValue, unsigned Q, typename TIter>
inline uint64_t hashNext(Shape<TValue, UngappedShape<Q> >& me, TIter it) {
    me.hValue = (me.hValue - (uint64_t)me.leftChar * me.leftFactor) * 4 + ordValue(it[Q - 1]);
    me.leftChar = ordValue(*it);
    return me.hValue;
}
inline void unhash(DnaString& result, uint64_t hash, unsigned q) {
    result.d.assign(q, Dna());
    for (unsigned i = q; i-- > 0;) { result.d[i].value = (unsigned char)(hash & 3); hash >>= 2; }
}

// ---------------------------------------------------------------------------------------------- BAM
template <typename T = char, typename C = unsigned> struct CigarElement { T operation; C count; CigarElement() : operation(0), count(0) {} };
enum BamHeaderRecordType { BAM_HEADER_FIRST, BAM_HEADER_REFERENCE, BAM_HEADER_READ_GROUP, BAM_HEADER_PROGRAM, BAM_HEADER_COMMENT };
struct BamHeaderRecord { BamHeaderRecordType type; String<Pair<CharString, CharString> > tags; };
struct BamHeader { String<BamHeaderRecord> records; };
inline void clear(BamHeader& h) { h.records.d.clear(); }

struct BamAlignmentRecord {
    CharString qName;
    __uint32 flag;
    int rID, beginPos;
    __uint8 mapQ;
    unsigned bin;
    String<CigarElement<> > cigar;
    int rNextId, pNext, tLen;
    CharString seq, qual, tags;
    BamAlignmentRecord() : flag(0), rID(-1), beginPos(-1), mapQ(0), bin(0), rNextId(-1), pNext(-1), tLen(0) {}
};
inline void clear(BamAlignmentRecord& r) { r = BamAlignmentRecord(); }
inline bool hasFlagMultiple(BamAlignmentRecord const& r) { return (r.flag & 0x1) != 0; }
inline bool hasFlagAllProper(BamAlignmentRecord const& r) { return (r.flag & 0x2) != 0; }
inline bool hasFlagUnmapped(BamAlignmentRecord const& r) { return (r.flag & 0x4) != 0; }
inline bool hasFlagNextUnmapped(BamAlignmentRecord const& r) { return (r.flag & 0x8) != 0; }
inline bool hasFlagRC(BamAlignmentRecord const& r) { return (r.flag & 0x10) != 0; }
inline bool hasFlagNextRC(BamAlignmentRecord const& r) { return (r.flag & 0x20) != 0; }
inline bool hasFlagFirst(BamAlignmentRecord const& r) { return (r.flag & 0x40) != 0; }
inline bool hasFlagLast(BamAlignmentRecord const& r) { return (r.flag & 0x80) != 0; }
inline bool hasFlagSecondary(BamAlignmentRecord const& r) { return (r.flag & 0x100) != 0; }
inline bool hasFlagQCNoPass(BamAlignmentRecord const& r) { return (r.flag & 0x200) != 0; }
inline bool hasFlagDuplicate(BamAlignmentRecord const& r) { return (r.flag & 0x400) != 0; }
inline bool hasFlagSupplementary(BamAlignmentRecord const& r) { return (r.flag & 0x800) != 0; }

template <typename TNameStore> struct BamIOContext { TNameStore* names; BamIOContext() : names(0) {} };

class BamTagsDict {  // index over the raw aux bytes (SAM/BAM spec section 4.2.4)
   public:
    struct Entry { size_t key, value, end; char type; };
    CharString const* host;
    std::vector<Entry> e;
    explicit BamTagsDict(CharString const& tags) : host(&tags) {
        std::vector<char> const& t = tags.d;
        size_t p = 0, n = t.size();
        while (p + 3 <= n) {
            Entry x;
            x.key = p;
            x.type = t[p + 2];
            p += 3;
            size_t len = 0;
            switch (x.type) {
                case 'A': case 'c': case 'C': len = 1; break;
                case 's': case 'S': len = 2; break;
                case 'i': case 'I': case 'f': len = 4; break;
                case 'Z': case 'H': { size_t q = p; while (q < n && t[q] != '\0') ++q; len = q - p + 1; break; }
                case 'B': {
                    if (p + 5 > n) { len = n - p; break; }
                    char sub = t[p];
                    uint32_t cnt;
                    memcpy(&cnt, &t[p + 1], 4);
                    len = 5 + (size_t)cnt * ((sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4);
                    break;
                }
                default: len = n - p; break;
            }
            x.value = p;
            x.end = std::min(n, p + len);
            p += len;
            e.push_back(x);
        }
    }
};
inline size_t length(BamTagsDict const& d) { return d.e.size(); }
inline CharString getTagKey(BamTagsDict const& d, size_t id) { CharString k; k.d.assign(d.host->d.begin() + d.e[id].key, d.host->d.begin() + d.e[id].key + 2); return k; }
inline char getTagType(BamTagsDict const& d, size_t id) { return d.e[id].type; }
inline CharString getTagValue(BamTagsDict const& d, size_t id) {  // R4: type char followed by the raw value bytes
    CharString v;
    v.d.push_back(d.e[id].type);
    v.d.insert(v.d.end(), d.host->d.begin() + d.e[id].value, d.host->d.begin() + d.e[id].end);
    return v;
}
inline bool findTagKey(unsigned& idx, BamTagsDict const& d, const char* key) {
    for (size_t i = 0; i < d.e.size(); ++i)
        if (d.host->d[d.e[i].key] == key[0] && d.host->d[d.e[i].key + 1] == key[1]) { idx = (unsigned)i; return true; }
    return false;
}
template <typename TDest> inline bool extractTagValue(TDest& dest, BamTagsDict const& d, size_t id) {  // R15
    const char* p = d.host->d.data() + d.e[id].value;
    switch (d.e[id].type) {
        case 'c': dest = (TDest)(int8_t)p[0]; return true;
        case 'C': dest = (TDest)(uint8_t)p[0]; return true;
        case 'A': dest = (TDest)p[0]; return true;
        case 's': { int16_t v; memcpy(&v, p, 2); dest = (TDest)v; return true; }
        case 'S': { uint16_t v; memcpy(&v, p, 2); dest = (TDest)v; return true; }
        case 'i': { int32_t v; memcpy(&v, p, 4); dest = (TDest)v; return true; }
        case 'I': { uint32_t v; memcpy(&v, p, 4); dest = (TDest)v; return true; }
        case 'f': { float v; memcpy(&v, p, 4); dest = (TDest)v; return true; }
        default: return false;
    }
}

class BamStream {
   public:
    enum Format { SAM, BAM };
    enum Mode { READ, WRITE };
    BamHeader header;
    StringSet<CharString> _nameStore;
    BamIOContext<StringSet<CharString> > bamIOContext;
    std::vector<uint8_t> data;
    size_t pos;
    bool good;
    BamStream(const char* path, Mode, Format format) : pos(0), good(false) {
        bamIOContext.names = &_nameStore;
        if (format != BAM) return;  // SAM on stdin is outside the hot path
        std::vector<uint8_t> raw;
        FILE* f = fopen(path, "rb");
        if (!f) return;
        fseek(f, 0, SEEK_END);
        long n = ftell(f);
        fseek(f, 0, SEEK_SET);
        raw.resize((size_t)n);
        size_t got = n ? fread(raw.data(), 1, (size_t)n, f) : 0;
        fclose(f);
        if (got != (size_t)n) return;
        if (raw.size() >= 4 && memcmp(raw.data(), "BAM\1", 4) == 0) data.swap(raw);
        else if (!inflateAll(raw)) return;
        good = parseHeader();
    }
    bool inflateAll(std::vector<uint8_t> const& in) {
        size_t p = 0;
        while (p + 18 <= in.size()) {
            if (in[p] != 0x1f || in[p + 1] != 0x8b) return false;
            uint16_t xlen; memcpy(&xlen, &in[p + 10], 2);
            size_t x = p + 12, xend = x + xlen;
            int bsize = -1;
            while (x + 4 <= xend) {
                uint16_t slen; memcpy(&slen, &in[x + 2], 2);
                if (in[x] == 'B' && in[x + 1] == 'C' && slen == 2) { uint16_t b; memcpy(&b, &in[x + 4], 2); bsize = b; }
                x += 4 + slen;
            }
            if (bsize < 0) return false;
            size_t blen = (size_t)bsize + 1;
            if (p + blen > in.size()) return false;
            uint32_t isize; memcpy(&isize, &in[p + blen - 4], 4);
            size_t o = data.size();
            data.resize(o + isize);
            if (isize) {
                z_stream zs; memset(&zs, 0, sizeof(zs));
                if (inflateInit2(&zs, -15) != Z_OK) return false;
                zs.next_in = (Bytef*)&in[p + 12 + xlen];
                zs.avail_in = (uInt)(blen - 12 - xlen - 8);
                zs.next_out = &data[o];
                zs.avail_out = isize;
                int rc = inflate(&zs, Z_FINISH);
                inflateEnd(&zs);
                if (rc != Z_STREAM_END) return false;
            }
            p += blen;
        }
        return p == in.size();
    }
    bool parseHeader() {
        if (data.size() < 12 || memcmp(data.data(), "BAM\1", 4) != 0) return false;
        int32_t l_text; memcpy(&l_text, &data[4], 4);
        std::string text((const char*)&data[8], (size_t)l_text);
        size_t p = 8 + (size_t)l_text;
        int32_t n_ref; memcpy(&n_ref, &data[p], 4); p += 4;
        for (int i = 0; i < n_ref; ++i) {
            int32_t l_name; memcpy(&l_name, &data[p], 4); p += 4;
            _nameStore.d.push_back(CharString(std::string((const char*)&data[p], (size_t)(l_name > 0 ? l_name - 1 : 0))));  // R11
            p += (size_t)l_name + 4;
        }
        pos = p;
        std::istringstream hs(text);
        std::string line;
        while (std::getline(hs, line)) {
            if (line.size() < 3 || line[0] != '@') continue;
            BamHeaderRecord rec;
            std::string ty = line.substr(1, 2);
            rec.type = ty == "HD" ? BAM_HEADER_FIRST : ty == "SQ" ? BAM_HEADER_REFERENCE : ty == "RG" ? BAM_HEADER_READ_GROUP : ty == "PG" ? BAM_HEADER_PROGRAM : BAM_HEADER_COMMENT;
            if (rec.type != BAM_HEADER_COMMENT) {
                std::istringstream ls(line);
                std::string field;
                bool firstField = true;
                while (std::getline(ls, field, '\t')) {
                    if (firstField) { firstField = false; continue; }
                    if (field.size() < 3 || field[2] != ':') continue;
                    Pair<CharString, CharString> kv;
                    kv.i1 = field.substr(0, 2);
                    kv.i2 = field.substr(3);
                    rec.tags.d.push_back(kv);
                }
            }
            header.records.d.push_back(rec);
        }
        return true;
    }
};
inline bool isGood(BamStream const& s) { return s.good; }
inline bool atEnd(BamStream const& s) { return s.pos >= s.data.size(); }
inline int readRecord(BamAlignmentRecord& r, BamStream& s) {  // R1-R3
    std::vector<uint8_t> const& d = s.data;
    if (s.pos + 4 > d.size()) return 1;
    int32_t bs; memcpy(&bs, &d[s.pos], 4);
    if (bs < 32 || s.pos + 4 + (size_t)bs > d.size()) return 1;
    const uint8_t* p = &d[s.pos + 4];
    s.pos += 4 + (size_t)bs;
    memcpy(&r.rID, p, 4);
    memcpy(&r.beginPos, p + 4, 4);
    uint8_t l_name = p[8];
    r.mapQ = p[9];
    uint16_t bin, ncig, flag;
    memcpy(&bin, p + 10, 2); memcpy(&ncig, p + 12, 2); memcpy(&flag, p + 14, 2);
    r.bin = bin;
    r.flag = flag;
    int32_t l_seq; memcpy(&l_seq, p + 16, 4);
    memcpy(&r.rNextId, p + 20, 4); memcpy(&r.pNext, p + 24, 4); memcpy(&r.tLen, p + 28, 4);
    size_t need = 32 + (size_t)l_name + 4 * (size_t)ncig + ((size_t)l_seq + 1) / 2 + (size_t)l_seq;
    if (l_seq < 0 || need > (size_t)bs) return 1;
    const uint8_t* q = p + 32;
    r.qName.d.assign((const char*)q, (const char*)q + (l_name ? l_name - 1 : 0));
    q += l_name;
    r.cigar.d.resize(ncig);
    static const char ops[] = "MIDNSHP=X";
    for (unsigned i = 0; i < ncig; ++i) {
        uint32_t v; memcpy(&v, q + 4 * i, 4);
        r.cigar.d[i].operation = (v & 15) < 9 ? ops[v & 15] : '?';
        r.cigar.d[i].count = v >> 4;
    }
    q += 4 * (size_t)ncig;
    static const char nt16[] = "=ACMGRSVTWYHKDBN";
    r.seq.d.resize((size_t)l_seq);
    for (int i = 0; i < l_seq; ++i) r.seq.d[i] = nt16[(q[i >> 1] >> ((~i & 1) << 2)) & 15];
    q += ((size_t)l_seq + 1) / 2;
    r.qual.d.resize((size_t)l_seq);
    for (int i = 0; i < l_seq; ++i) r.qual.d[i] = (char)(q[i] + 33);
    q += l_seq;
    r.tags.d.assign((const char*)q, (const char*)q + ((size_t)bs - need));
    return 0;
}
template <typename TStream, typename TContext> inline int write2(TStream& out, BamAlignmentRecord const& r, TContext const&, Sam) {
    out << r.qName << "\t" << r.flag << "\t" << r.rID << "\t" << r.beginPos + 1 << "\n";
    return 0;
}

// -------------------------------------------------------------------------------------- SequenceStream
class SequenceStream {
   public:
    enum Mode { READ, WRITE };
    enum Format { FASTA, FASTQ };
    std::ifstream in;
    std::string pending;
    bool havePending;
    SequenceStream() : havePending(false) {}
};
inline void open(SequenceStream& s, const char* path, SequenceStream::Mode, SequenceStream::Format) { s.in.open(path, std::ios::binary); }
inline bool isGood(SequenceStream const& s) { return s.in.is_open() && !s.in.bad(); }
inline int readRecord(CharString& id, CharString& seq, SequenceStream& s) {  // R10
    std::string line;
    if (!s.havePending) {
        while (std::getline(s.in, line)) {
            if (!line.empty() && line[0] == '>') { s.pending = line; s.havePending = true; break; }
        }
    }
    if (!s.havePending) return 1;
    std::string hdr = s.pending.substr(1);
    if (!hdr.empty() && hdr[hdr.size() - 1] == '\r') hdr.erase(hdr.size() - 1);
    s.havePending = false;
    id = hdr;
    seq.d.clear();
    while (std::getline(s.in, line)) {
        if (!line.empty() && line[0] == '>') { s.pending = line; s.havePending = true; break; }
        for (size_t i = 0; i < line.size(); ++i)
            if (!isspace((unsigned char)line[i])) seq.d.push_back(line[i]);
    }
    return 0;
}

// -------------------------------------------------------------------------------------- ArgumentParser
struct ArgParseArgument {
    enum ArgumentType { STRING, INTEGER, INT64, DOUBLE, INPUTFILE, OUTPUTFILE };
    ArgumentType type;
    std::string label;
    ArgParseArgument(ArgumentType t, const char* l = "", bool = false, unsigned = 1) : type(t), label(l) {}
    ArgParseArgument(ArgumentType t, const char* l, const char*, unsigned) : type(t), label(l) {}
};
struct ArgParseOption {
    std::string shortName, longName, help, label, value, def;
    ArgParseArgument::ArgumentType type;
    bool required, isSet, hasDefault;
    ArgParseOption(const char* s, const char* l, const char* h, ArgParseArgument::ArgumentType t, const char* lab = "")
        : shortName(s), longName(l), help(h), label(lab), type(t), required(false), isSet(false), hasDefault(false) {}
};
class ArgumentParser {
   public:
    enum ParseResult { PARSE_OK, PARSE_ERROR, PARSE_HELP, PARSE_VERSION, PARSE_WRITE_CTD, PARSE_EXPORT_HELP };
    std::string name, version, date, shortDesc;
    std::vector<ArgParseOption> options;
    std::vector<std::string> validSuffixes, positional;
    explicit ArgumentParser(const char* n) : name(n) {}
    ArgParseOption* find(std::string const& key) {
        for (size_t i = 0; i < options.size(); ++i)
            if (options[i].shortName == key || options[i].longName == key) return &options[i];
        return 0;
    }
};
inline void setShortDescription(ArgumentParser& p, const char* s) { p.shortDesc = s; }
inline void setDate(ArgumentParser& p, const char* s) { p.date = s; }
inline void setVersion(ArgumentParser& p, const char* s) { p.version = s; }
inline void addUsageLine(ArgumentParser&, const char*) {}
inline void addDescription(ArgumentParser&, const char*) {}
inline void addSection(ArgumentParser&, const char*) {}
inline void addArgument(ArgumentParser&, ArgParseArgument const&) {}
inline void addOption(ArgumentParser& p, ArgParseOption const& o) { p.options.push_back(o); }
inline void setValidValues(ArgumentParser& p, unsigned, const char* values) { std::istringstream s(values); std::string v; while (s >> v) p.validSuffixes.push_back(v); }
inline void setDefaultValue(ArgumentParser& p, const char* key, const char* v) { ArgParseOption* o = p.find(key); if (o) { o->def = v; o->hasDefault = true; } }
inline void setMinValue(ArgumentParser&, const char*, const char*) {}
inline void setRequired(ArgumentParser& p, const char* key) { ArgParseOption* o = p.find(key); if (o) o->required = true; }
inline ArgumentParser::ParseResult parse(ArgumentParser& p, int argc, char const** argv, std::ostream& out, std::ostream& err) {
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "-h" || a == "--help") { out << p.name << " - " << p.shortDesc << "\n"; return ArgumentParser::PARSE_HELP; }
        if (a == "--version") { out << p.name << " version: " << p.version << "\n"; return ArgumentParser::PARSE_VERSION; }
        if (a == "-") { err << p.name << ": illegal option -- -\n"; return ArgumentParser::PARSE_ERROR; }
        if (a.size() > 1 && a[0] == '-') {
            std::string key = a.substr(a[1] == '-' ? 2 : 1), val;
            size_t eq = key.find('=');
            bool haveVal = false;
            if (eq != std::string::npos) { val = key.substr(eq + 1); key = key.substr(0, eq); haveVal = true; }
            ArgParseOption* o = p.find(key);
            if (!o) { err << p.name << ": illegal option -- " << key << "\n"; return ArgumentParser::PARSE_ERROR; }
            if (!haveVal) {
                if (i + 1 >= argc) { err << p.name << ": option requires an argument -- " << key << "\n"; return ArgumentParser::PARSE_ERROR; }
                val = argv[++i];
            }
            o->value = val;
            o->isSet = true;
        } else {
            p.positional.push_back(a);
        }
    }
    for (size_t i = 0; i < p.options.size(); ++i)
        if (p.options[i].required && !p.options[i].isSet) { err << p.name << ": option requires an argument -- " << p.options[i].shortName << "\n"; return ArgumentParser::PARSE_ERROR; }
    if (p.positional.size() != 1) { err << p.name << ": wrong number of arguments\n"; return ArgumentParser::PARSE_ERROR; }
    bool okSuffix = p.validSuffixes.empty();
    for (size_t i = 0; i < p.validSuffixes.size(); ++i) {
        std::string const& s = p.validSuffixes[i];
        if (p.positional[0].size() >= s.size() && p.positional[0].compare(p.positional[0].size() - s.size(), s.size(), s) == 0) okSuffix = true;
    }
    if (!okSuffix) { err << p.name << ": the given value '" << p.positional[0] << "' is not in the list of allowed file extensions\n"; return ArgumentParser::PARSE_ERROR; }
    return ArgumentParser::PARSE_OK;
}
inline std::string optionText(ArgumentParser& p, const char* key) { ArgParseOption* o = p.find(key); if (!o) return ""; return o->isSet ? o->value : o->def; }
inline bool getOptionValue(CharString& d, ArgumentParser& p, const char* key) { ArgParseOption* o = p.find(key); if (o && (o->isSet || o->hasDefault)) d = optionText(p, key); return true; }
inline bool getOptionValue(std::string& d, ArgumentParser& p, const char* key) { d = optionText(p, key); return true; }
inline bool getOptionValue(int& d, ArgumentParser& p, const char* key) { d = atoi(optionText(p, key).c_str()); return true; }
inline bool getOptionValue(double& d, ArgumentParser& p, const char* key) { d = atof(optionText(p, key).c_str()); return true; }
inline bool getArgumentValue(CharString& d, ArgumentParser& p, unsigned i) { if (i < p.positional.size()) d = p.positional[i]; return true; }

}  // namespace seqan
#endif  // MINISEQAN_H_
