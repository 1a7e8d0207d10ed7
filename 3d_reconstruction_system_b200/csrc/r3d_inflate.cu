// r3d_inflate.cu -- zlib-stream (RFC 1950 / 1951) inflate for the PNG decoder (a1), host code only.
//
// The frame decode that feeds the GPU path is inflate-bound (7.6 of the 10.7 ms that one core spends on a 1242x375 16-bit
// depth PNG go to zlib's inflate, 123 MB/s of output on noisy depth data), so the decoder carries its own: whole-buffer
// in, whole-buffer out (a PNG's size is known from its header), a 64-bit bit buffer refilled with one unaligned load,
// 11-bit / 8-bit two-level decode tables whose entries carry base value, extra-bit count and code length, literals
// decoded back to back, matches copied eight bytes at a time.  Every stream it accepts produces the same bytes as zlib
// (tests/test_inflate_cpu.py: zlib levels 0-9, stored / fixed / dynamic blocks, overlapping matches, truncated and
// corrupted streams); the Adler-32 trailer is checked like zlib's uncompress() does.
#include <stdint.h>
#include <string.h>

#include "r3d_common.cuh"

namespace r3d {

namespace {

constexpr int kLitBits = 11, kDistBits = 8;
constexpr int kMaxLitSyms = 288, kMaxDistSyms = 32, kMaxCodeLen = 15;
// table sizes (primary + all sub-tables).  A sub-table of width w belongs to a complete prefix subtree of depth w, which
// has at least w + 1 leaves, so sub-tables cost at most 2^w / (w + 1) entries per symbol: 16 / 5 for the 286 literal /
// length symbols under an 11-bit root (2048 + 915), 128 / 8 for the 30 distance symbols under an 8-bit root (256 + 480).
// build_table() checks the capacity anyway.
constexpr int kLitTableSize = 3072, kDistTableSize = 768;

// table entry: [4:0] bits to consume, [8:5] extra bits (or sub-table width), [11:9] kind, [30:16] value, [31] literal.
// A literal entry of the literal/length root table may carry TWO literals when both codewords fit the root width
// ([30] set, second literal in [15:8], [4:0] the sum of the two lengths): literal-heavy data -- noisy depth images --
// is bound by the lookup -> shift -> lookup latency chain, and this halves the chain per byte.
enum : uint32_t { KIND_LITERAL = 0, KIND_BASE = 1, KIND_END = 2, KIND_SUB = 3, KIND_BAD = 4 };
constexpr uint32_t kLiteralFlag = 0x80000000u, kDoubleFlag = 0x40000000u;
inline uint32_t entry(uint32_t bits, uint32_t extra, uint32_t kind, uint32_t value) {
    return bits | (extra << 5) | (kind << 9) | (value << 16) | (kind == KIND_LITERAL ? kLiteralFlag : 0u);
}
inline uint32_t value_of(uint32_t e) { return (e >> 16) & 0x7fffu; }

const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

inline uint32_t bit_reverse(uint32_t v, int n) {
    uint32_t r = 0;
    for (int i = 0; i < n; ++i) { r = (r << 1) | (v & 1u); v >>= 1; }
    return r;
}

// Canonical Huffman code lengths -> two-level table.  `mode` selects what a symbol means.  Returns false for an
// over-subscribed code or an incomplete one, with zlib's exception: a literal/length or distance code whose longest
// codeword has one bit may be incomplete (the unused codeword decodes to an error).
enum { MODE_LITLEN = 0, MODE_DIST = 1, MODE_CODELEN = 2 };
bool build_table(const uint8_t* lens, int n_syms, int mode, int root, uint32_t* table, int table_cap) {
    const bool lit = mode == MODE_LITLEN;
    int count[kMaxCodeLen + 1] = {0};
    for (int s = 0; s < n_syms; ++s) count[lens[s]]++;
    if (count[0] == n_syms) {
        // no code at all: legal for the distance alphabet of a block made of literals only
        if (mode != MODE_DIST) return false;
        for (int i = 0; i < (1 << root); ++i) table[i] = entry(1, 0, KIND_BAD, 0);
        return true;
    }
    int left = 1, max_len = 0;
    for (int l = 1; l <= kMaxCodeLen; ++l) {
        left = (left << 1) - count[l];
        if (left < 0) return false;                       // over-subscribed
        if (count[l]) max_len = l;
    }
    if (left > 0 && (mode == MODE_CODELEN || max_len != 1)) return false;   // incomplete
    uint32_t next_code[kMaxCodeLen + 2];
    {
        uint32_t code = 0;
        count[0] = 0;
        for (int l = 1; l <= kMaxCodeLen; ++l) { code = (code + (uint32_t)count[l - 1]) << 1; next_code[l] = code; }
    }
    for (int i = 0; i < (1 << root); ++i) table[i] = entry(1, 0, KIND_BAD, 0);
    // sub-tables: one per distinct `root`-bit prefix of the long codes, sized for the longest code under that prefix
    int used = 1 << root;
    // first pass: width of each sub-table
    uint8_t sub_width[1 << kLitBits];
    memset(sub_width, 0, sizeof sub_width);
    if (max_len > root) {
        uint32_t code_of[kMaxLitSyms];
        uint32_t nc[kMaxCodeLen + 2];
        memcpy(nc, next_code, sizeof nc);
        for (int s = 0; s < n_syms; ++s) {
            const int l = lens[s];
            if (!l) continue;
            code_of[s] = nc[l]++;
            if (l > root) {
                const uint32_t prefix = bit_reverse(code_of[s] >> (l - root), root);
                if (l - root > sub_width[prefix]) sub_width[prefix] = (uint8_t)(l - root);
            }
        }
        for (int p = 0; p < (1 << root); ++p) {
            if (!sub_width[p]) continue;
            if (used + (1 << sub_width[p]) > table_cap) return false;
            table[p] = entry((uint32_t)root, sub_width[p], KIND_SUB, (uint32_t)used);
            for (int i = 0; i < (1 << sub_width[p]); ++i) table[used + i] = entry(1, 0, KIND_BAD, 0);
            used += 1 << sub_width[p];
        }
    }
    for (int s = 0; s < n_syms; ++s) {
        const int l = lens[s];
        if (!l) continue;
        const uint32_t code = next_code[l]++;
        uint32_t e;
        if (mode == MODE_CODELEN) {
            e = entry(0, 0, KIND_LITERAL, (uint32_t)s);
        } else if (lit) {
            if (s < 256) e = entry(0, 0, KIND_LITERAL, (uint32_t)s);
            else if (s == 256) e = entry(0, 0, KIND_END, 0);
            else if (s < 286) e = entry(0, kLenExtra[s - 257], KIND_BASE, kLenBase[s - 257]);
            else e = entry(0, 0, KIND_BAD, 0);
        } else {
            e = s < 30 ? entry(0, kDistExtra[s], KIND_BASE, kDistBase[s]) : entry(0, 0, KIND_BAD, 0);
        }
        if (l <= root) {
            const uint32_t rev = bit_reverse(code, l);
            for (uint32_t i = rev; i < (1u << root); i += 1u << l) table[i] = e | (uint32_t)l;
        } else {
            const uint32_t prefix = bit_reverse(code >> (l - root), root);
            const uint32_t sub = value_of(table[prefix]), width = (table[prefix] >> 5) & 15u;
            const int rest = l - root;
            const uint32_t rev = bit_reverse(code & ((1u << rest) - 1u), rest);
            for (uint32_t i = rev; i < (1u << width); i += 1u << rest) table[sub + i] = e | (uint32_t)rest;
        }
    }
    if (lit) {
        // pair up literals: entry i decodes codeword 1 from its low bits; when the bits above it hold a complete second
        // literal codeword, one lookup yields both
        uint32_t single[1 << kLitBits];
        memcpy(single, table, sizeof(uint32_t) << root);
        for (uint32_t i = 0; i < (1u << root); ++i) {
            const uint32_t e1 = single[i];
            if ((int32_t)e1 >= 0) continue;
            const uint32_t l1 = e1 & 31u;
            const uint32_t e2 = single[i >> l1];
            if ((int32_t)e2 >= 0 || l1 + (e2 & 31u) > (uint32_t)root) continue;
            table[i] = (e1 & 0x80ff0000u) | kDoubleFlag | (((e2 >> 16) & 0xffu) << 8) | (l1 + (e2 & 31u));
        }
    }
    return true;
}

struct BitReader {
    const uint8_t* in;
    const uint8_t* end;
    uint64_t buf = 0;
    unsigned cnt = 0;       // valid bits in buf
    unsigned over = 0;      // zero bytes supplied past the end of the input

    // at least 56 valid bits afterwards (zeros past the end of the input; `over` counts them)
    inline void refill() {
        if (end - in >= 8) {
            uint64_t v;
            memcpy(&v, in, 8);
            buf |= v << cnt;
            in += (63 - cnt) >> 3;
            cnt |= 56;
        } else {
            while (cnt <= 56) {
                if (in < end) buf |= (uint64_t)*in++ << cnt;
                else ++over;
                cnt += 8;
            }
        }
    }
    inline uint32_t peek(unsigned n) const { return (uint32_t)(buf & ((1ull << n) - 1ull)); }
    inline void drop(unsigned n) { buf >>= n; cnt -= n; }
    inline uint32_t take(unsigned n) { const uint32_t v = peek(n); drop(n); return v; }
    // bits really available (not the zero padding)
    inline bool overrun() const { return over * 8u > cnt; }
};

struct Tables {
    uint32_t lit[kLitTableSize];
    uint32_t dist[kDistTableSize];
};

bool fixed_tables(Tables& t) {
    uint8_t lens[kMaxLitSyms];
    for (int i = 0; i < 144; ++i) lens[i] = 8;
    for (int i = 144; i < 256; ++i) lens[i] = 9;
    for (int i = 256; i < 280; ++i) lens[i] = 7;
    for (int i = 280; i < 288; ++i) lens[i] = 8;
    if (!build_table(lens, 288, MODE_LITLEN, kLitBits, t.lit, kLitTableSize)) return false;
    uint8_t dl[32];
    for (int i = 0; i < 32; ++i) dl[i] = 5;
    return build_table(dl, 32, MODE_DIST, kDistBits, t.dist, kDistTableSize);
}

bool dynamic_tables(BitReader& br, Tables& t) {
    br.refill();
    const unsigned hlit = br.take(5) + 257, hdist = br.take(5) + 1, hclen = br.take(4) + 4;
    if (hlit > 286 || hdist > 30) return false;
    static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    uint8_t cl[19] = {0};
    for (unsigned i = 0; i < hclen; ++i) {
        if (br.cnt < 3) br.refill();
        cl[order[i]] = (uint8_t)br.take(3);
    }
    uint32_t cltab[1 << 7];
    if (!build_table(cl, 19, MODE_CODELEN, 7, cltab, 1 << 7)) return false;   // longest code 7 bits: no sub-tables
    uint8_t lens[kMaxLitSyms + kMaxDistSyms];
    unsigned n = 0;
    while (n < hlit + hdist) {
        br.refill();
        const uint32_t e = cltab[br.peek(7)];
        if (((e >> 9) & 7u) != KIND_LITERAL) return false;
        br.drop(e & 31u);
        const unsigned sym = value_of(e);
        if (sym < 16) { lens[n++] = (uint8_t)sym; continue; }
        unsigned rep, val = 0;
        if (sym == 16) {
            if (n == 0) return false;
            val = lens[n - 1];
            rep = 3 + br.take(2);
        } else if (sym == 17) rep = 3 + br.take(3);
        else rep = 11 + br.take(7);
        if (n + rep > hlit + hdist) return false;
        memset(lens + n, (int)val, rep);
        n += rep;
    }
    if (br.overrun()) return false;
    if (lens[256] == 0) return false;                    // no end-of-block code
    if (!build_table(lens, (int)hlit, MODE_LITLEN, kLitBits, t.lit, kLitTableSize)) return false;
    return build_table(lens + hlit, (int)hdist, MODE_DIST, kDistBits, t.dist, kDistTableSize);
}

// Adler-32 with the per-byte dependency chain (a += p; b += a) broken up: over a run of k = 16 B bytes
// a' = a + sum p[i] and b' = b + k a + sum (k - i) p[i].  Byte i = 16 q + j has weight 16 (B - q) - j, so with sixteen
// column accumulators (c1[j] += p[16 q + j]; c2[j] += c1[j], plain vector adds) the weighted sum is
// 16 sum_j c2[j] - sum_j j c1[j].
__attribute__((optimize("O3", "tree-vectorize"))) uint32_t adler32(const uint8_t* p, size_t n) {
    uint32_t a = 1, b = 0;
    while (n >= 16) {
        const size_t blocks = (n >> 4) < 256 ? (n >> 4) : 256;      // c2 <= 255 * 256 * 257 / 2: no overflow
        uint32_t c1[16] = {0}, c2[16] = {0};
        for (size_t q = 0; q < blocks; ++q) {
            for (int j = 0; j < 16; ++j) { c1[j] += p[16 * q + j]; c2[j] += c1[j]; }
        }
        uint32_t s1 = 0, s2 = 0, sj = 0;
        for (int j = 0; j < 16; ++j) { s1 += c1[j]; s2 += c2[j]; sj += (uint32_t)j * c1[j]; }
        const uint32_t k = (uint32_t)blocks * 16u;
        b = (b + k * a + 16u * s2 - sj) % 65521u;
        a = (a + s1) % 65521u;
        p += k;
        n -= k;
    }
    while (n--) { a += *p++; b += a; }
    return ((b % 65521u) << 16) | (a % 65521u);
}

}  // namespace

// Inflates one zlib stream of exactly known output size.  Returns 0 on success, a negative code otherwise:
// -1 bad header, -2 bad block / code tables, -3 input ends early, -4 output does not fit or a distance reaches before the
// start, -5 output shorter than `dst_len`, -6 Adler-32 mismatch.
int inflate_zlib(const uint8_t* src, size_t src_len, uint8_t* dst, size_t dst_len) {
    if (src_len < 6) return -1;
    const unsigned cmf = src[0], flg = src[1];
    if ((cmf & 15u) != 8 || (cmf >> 4) > 7 || ((cmf << 8) | flg) % 31u != 0 || (flg & 32u)) return -1;
    BitReader br{src + 2, src + src_len};
    uint8_t* out = dst;
    uint8_t* const out_end = dst + dst_len;
    Tables* tabs = new Tables;
    struct Free { Tables* t; ~Free() { delete t; } } free_tabs{tabs};
    bool fixed_ready = false, last = false;
    Tables* fixed = nullptr;
    struct FreeFixed { Tables** t; ~FreeFixed() { delete *t; } } free_fixed{&fixed};
    while (!last) {
        br.refill();
        if (br.overrun()) return -3;
        last = br.take(1) != 0;
        const unsigned type = br.take(2);
        if (type == 0) {
            // stored: skip to the byte boundary, LEN / NLEN, raw bytes
            br.drop(br.cnt & 7u);
            br.refill();
            const unsigned len = br.take(16), nlen = br.take(16);
            if ((len ^ nlen) != 0xffffu) return -2;
            if (br.overrun()) return -3;
            // bytes still in the bit buffer first, then straight from the input
            unsigned n = len;
            while (n && br.cnt >= 8) {
                if (br.over * 8u >= br.cnt) return -3;
                if (out >= out_end) return -4;
                *out++ = (uint8_t)br.take(8);
                --n;
            }
            if (n) {
                br.buf = 0;                                  // (bit buffer drained: forget the look-ahead bits above cnt)
                br.cnt = 0;
                if ((size_t)(br.end - br.in) < n) return -3;
                if ((size_t)(out_end - out) < n) return -4;
                memcpy(out, br.in, n);
                br.in += n;
                out += n;
            }
            continue;
        }
        const Tables* t;
        if (type == 1) {
            if (!fixed_ready) {
                fixed = new Tables;
                if (!fixed_tables(*fixed)) return -2;
                fixed_ready = true;
            }
            t = fixed;
        } else if (type == 2) {
            if (!dynamic_tables(br, *tabs)) return br.overrun() ? -3 : -2;
            t = tabs;
        } else {
            return -2;
        }
        // ---- fast loop: while 16 input bytes and a maximum-length match + copy slack are in reach, nothing is checked
        // against the buffer ends and the bit buffer lives in registers
        bool block_done = false;
        if (br.end - br.in >= 16 && out_end - out >= 274) {
            const uint8_t* in = br.in;
            const uint8_t* const in_safe = br.end - 16;
            uint8_t* const out_safe = out_end - 274;
            uint64_t buf = br.buf;
            unsigned cnt = br.cnt;
            const uint32_t* const lit = t->lit;
            const uint32_t* const dtab = t->dist;
            int bad = 0;
#define R3D_REFILL()  do { uint64_t v_; memcpy(&v_, in, 8); buf |= v_ << cnt; in += (63 - cnt) >> 3; cnt |= 56; } while (0)
            while (in <= in_safe && out <= out_safe) {
                R3D_REFILL();
                uint32_t e = lit[buf & ((1u << kLitBits) - 1u)];
#define R3D_LITERALS()  do { buf >>= (e & 31u); cnt -= (e & 31u); out[0] = (uint8_t)(e >> 16); out[1] = (uint8_t)(e >> 8); \
                             out += 1u + ((e >> 30) & 1u); } while (0)
                if ((int32_t)e < 0) {                       // up to three lookups (<= 11 bits, one or two literals each) per refill
                    R3D_LITERALS();
                    e = lit[buf & ((1u << kLitBits) - 1u)];
                    if ((int32_t)e < 0) {
                        R3D_LITERALS();
                        e = lit[buf & ((1u << kLitBits) - 1u)];
                        if ((int32_t)e < 0) {
                            R3D_LITERALS();
                            continue;
                        }
                    }
                    R3D_REFILL();                            // a length + distance pair needs up to 48 bits
                }
                unsigned kind = (e >> 9) & 7u;
                if (kind == KIND_SUB) {
                    buf >>= kLitBits; cnt -= kLitBits;
                    e = lit[value_of(e) + (uint32_t)(buf & ((1u << ((e >> 5) & 15u)) - 1u))];
                    kind = (e >> 9) & 7u;
                    if (kind == KIND_LITERAL) {
                        buf >>= (e & 31u); cnt -= (e & 31u);
                        *out++ = (uint8_t)(e >> 16);
                        continue;
                    }
                }
                buf >>= (e & 31u); cnt -= (e & 31u);
                if (kind != KIND_BASE) {
                    if (kind == KIND_END) block_done = true; else bad = 1;
                    break;
                }
                const unsigned xl = (e >> 5) & 15u;
                const unsigned length = value_of(e) + (unsigned)(buf & ((1u << xl) - 1u));
                buf >>= xl; cnt -= xl;
                uint32_t d = dtab[buf & ((1u << kDistBits) - 1u)];
                if (((d >> 9) & 7u) == KIND_SUB) {
                    buf >>= kDistBits; cnt -= kDistBits;
                    d = dtab[value_of(d) + (uint32_t)(buf & ((1u << ((d >> 5) & 15u)) - 1u))];
                }
                if (((d >> 9) & 7u) != KIND_BASE) { bad = 1; break; }
                buf >>= (d & 31u); cnt -= (d & 31u);
                const unsigned xd = (d >> 5) & 15u;
                const unsigned dist = value_of(d) + (unsigned)(buf & ((1u << xd) - 1u));
                buf >>= xd; cnt -= xd;
                if (dist > (size_t)(out - dst)) { bad = 2; break; }
                const uint8_t* from = out - dist;
                uint8_t* const stop = out + length;
                if (dist >= 8) {
                    do {
                        uint64_t v;
                        memcpy(&v, from, 8);
                        memcpy(out, &v, 8);
                        from += 8; out += 8;
                    } while (out < stop);
                } else if (dist == 1) {
                    memset(out, *from, length);
                } else {
                    do { *out++ = *from++; } while (out < stop);
                }
                out = stop;
            }
#undef R3D_REFILL
#undef R3D_LITERALS
            br.in = in; br.buf = buf; br.cnt = cnt;
            if (bad) return bad == 2 ? -4 : -2;
        }
        // ---- careful loop: the rest of the block (and whole blocks near the ends of the buffers)
        while (!block_done) {
            br.refill();
            uint32_t e = t->lit[br.peek(kLitBits)];
            // (the literal flag first: a two-literal entry keeps its second literal where the other kinds keep theirs)
            if ((int32_t)e >= 0 && ((e >> 9) & 7u) == KIND_SUB) {
                br.drop(kLitBits);
                e = t->lit[value_of(e) + br.peek((e >> 5) & 15u)];
            }
            const unsigned kind = (int32_t)e < 0 ? (unsigned)KIND_LITERAL : ((e >> 9) & 7u);
            if (kind == KIND_LITERAL) {
                const unsigned n_lit = 1u + ((e >> 30) & 1u);
                if ((size_t)(out_end - out) < n_lit) return -4;
                br.drop(e & 31u);
                *out++ = (uint8_t)(e >> 16);
                if (n_lit == 2) *out++ = (uint8_t)(e >> 8);
                if (br.overrun()) return -3;
                continue;
            }
            br.drop(e & 31u);
            if (kind == KIND_END) break;
            if (kind != KIND_BASE) return br.overrun() ? -3 : -2;
            // (after a refill: <= 15 + 5 bits gone so far, 36 left at least; the distance needs <= 15 + 13)
            const unsigned length = value_of(e) + br.take((e >> 5) & 15u);
            uint32_t d = t->dist[br.peek(kDistBits)];
            if (((d >> 9) & 7u) == KIND_SUB) {
                br.drop(kDistBits);
                d = t->dist[value_of(d) + br.peek((d >> 5) & 15u)];
            }
            if (((d >> 9) & 7u) != KIND_BASE) return br.overrun() ? -3 : -2;
            br.drop(d & 31u);
            const unsigned dist = value_of(d) + br.take((d >> 5) & 15u);
            if (br.overrun()) return -3;
            if (dist > (size_t)(out - dst)) return -4;
            if ((size_t)(out_end - out) < length) return -4;
            const uint8_t* from = out - dist;
            if (dist >= 8 && (size_t)(out_end - out) >= length + 8) {
                // eight bytes at a time; may write up to 7 bytes past the match, inside the output buffer
                uint8_t* o = out;
                uint8_t* const stop = out + length;
                do {
                    uint64_t v;
                    memcpy(&v, from, 8);
                    memcpy(o, &v, 8);
                    from += 8;
                    o += 8;
                } while (o < stop);
                out = stop;
            } else if (dist == 1) {
                memset(out, *from, length);
                out += length;
            } else {
                for (unsigned i = 0; i < length; ++i) out[i] = from[i];
                out += length;
            }
        }
    }
    if (br.overrun()) return -3;
    if (out != out_end) return -5;
    // the Adler-32 of the output follows the last block at the next byte boundary
    const uint8_t* tr = br.in - ((br.cnt >> 3) - br.over);     // whole bytes still in the bit buffer, minus the padding ones
    if ((size_t)(src + src_len - tr) < 4) return -3;
    const uint32_t want = ((uint32_t)tr[0] << 24) | ((uint32_t)tr[1] << 16) | ((uint32_t)tr[2] << 8) | tr[3];
    if (adler32(dst, dst_len) != want) return -6;
    return 0;
}

}  // namespace r3d

// Exposed for the CPU tests (tests/test_inflate_cpu.py): the PNG decoder's inflate on a caller's buffers.
extern "C" int r3d_inflate(const void* src, size_t src_len, void* dst, size_t dst_len) {
    if (!src || (!dst && dst_len)) return r3d::set_error(nullptr, R3D_ERR_ARG, "r3d_inflate: null buffer");
    const int rc = r3d::inflate_zlib((const uint8_t*)src, src_len, (uint8_t*)dst, dst_len);
    if (rc != 0) return r3d::set_error(nullptr, R3D_ERR_ARG, "r3d_inflate: malformed zlib stream (code %d)", rc);
    return R3D_OK;
}
