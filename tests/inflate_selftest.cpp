// inflate_selftest.cpp -- CPU check of bamqc_b200/csrc/inflate_bits.h (BitWin, the bit reader of k_inflate) against a plain
// LSB-first bit reader (RFC 1951 3.1.1) on random payloads at every byte misalignment, with the access pattern of the
// symbol loop: header fields through take(), trips of up to 32 bits out of win() after norm(), a second window through
// win2() without a refill in between (the distance code), byte alignment and seek() as for stored blocks.
// Test infrastructure.  usage: inflate_selftest [rounds]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../bamqc_b200/csrc/inflate_bits.h"

using namespace bqc;

struct PlainBits {   // bit i of the stream = bit (i & 7) of byte i >> 3
    const uint8_t* p;
    uint64_t pos = 0;
    uint32_t peek(uint32_t n) const {
        uint32_t v = 0;
        for (uint32_t i = 0; i < n; ++i) v |= (uint32_t)((p[(pos + i) >> 3] >> ((pos + i) & 7)) & 1u) << i;
        return v;
    }
};

int main(int argc, char** argv) {
    const long rounds = argc > 1 ? atol(argv[1]) : 2000;
    std::mt19937_64 rng(20261018);
    long bad = 0, checks = 0;
    for (long r = 0; r < rounds && bad < 10; ++r) {
        const uint32_t clen = 64 + (uint32_t)(rng() % 3000);
        const uint32_t mis = (uint32_t)(rng() % 4);
        std::vector<uint32_t> store((clen + mis) / 4 + 16, 0);   // aligned storage, slack for the reader's look-ahead
        uint8_t* payload = reinterpret_cast<uint8_t*>(store.data()) + mis;
        for (uint32_t i = 0; i < clen; ++i) payload[i] = (uint8_t)rng();
        BitWin br;
        br.init(payload, clen);
        PlainBits ref{payload};
        auto expect = [&](uint32_t got, uint32_t want, const char* what) {
            ++checks;
            if (got != want && bad < 10) { ++bad; printf("round %ld %s: got %08x want %08x at bit %llu (mis %u)\n", r, what, got, want, (unsigned long long)ref.pos, mis); }
        };
        while (ref.pos + 96 < (uint64_t)clen * 8 && bad < 10) {
            switch (rng() % 5) {
                case 0: {   // header field
                    const uint32_t n = 1 + (uint32_t)(rng() % 16);
                    expect(br.take(n), ref.peek(n), "take");
                    ref.pos += n;
                    break;
                }
                case 1: {   // literal trip: one window, up to 32 bits in up to four pieces
                    br.norm();
                    uint32_t w = br.win(), used = 0;
                    expect(w, ref.peek(32), "win");
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t l = 1 + (uint32_t)(rng() % 9);
                        if (used + l > 32) break;
                        expect(w & ((1u << l) - 1u), PlainBits{payload, ref.pos + used}.peek(l), "piece");
                        w >>= l;
                        used += l;
                    }
                    br.bp += used;
                    ref.pos += used;
                    break;
                }
                case 2: {   // match trip: length (<= 20 bits) from win(), distance (<= 28 bits) from win2() without a refill
                    br.norm();
                    expect(br.win(), ref.peek(32), "win");
                    const uint32_t a = 1 + (uint32_t)(rng() % 20);
                    br.bp += a;
                    ref.pos += a;
                    expect(br.win2(), ref.peek(32), "win2");
                    const uint32_t b = 1 + (uint32_t)(rng() % 28);
                    br.bp += b;
                    ref.pos += b;
                    break;
                }
                case 3: {   // stored block: align to a byte, 16 + 16 bits, jump
                    br.norm();
                    br.bp = (br.bp + 7u) & ~7u;
                    ref.pos = (ref.pos + 7) & ~7ull;
                    expect(br.take(16), ref.peek(16), "len");
                    ref.pos += 16;
                    expect(br.take(16), ref.peek(16), "nlen");
                    ref.pos += 16;
                    expect(br.bytes_used(), (uint32_t)(ref.pos >> 3), "bytes_used");
                    const uint32_t skip = (uint32_t)(rng() % 40);
                    if (ref.pos / 8 + skip + 16 < clen) {
                        br.seek((uint32_t)(ref.pos >> 3) + skip);
                        ref.pos += 8ull * skip;
                    }
                    break;
                }
                default:
                    expect(br.bits_used(), (uint32_t)ref.pos, "bits_used");
                    expect(br.bytes_used(), (uint32_t)((ref.pos + 7) >> 3), "bytes_used");
            }
        }
    }
    printf("%ld checks, %ld mismatching\n", checks, bad);
    return bad ? 1 : 0;
}
