// r3d_strtod.cuh -- decimal text -> double for the point-text readers (a8), host code only.
//
// float() in the reference's txt_read (octomap/txt_transfer_octomap.py:22-24) is correctly rounded; so is strtod, which
// the readers used for every field at ~100 ns apiece.  parse_double below takes the plain decimal fields that make up
// these files -- [sign] digits [. digits] [e [sign] digits], at most 19 significant digits -- and rounds them itself:
// Clinger's exact case (significand <= 2^53, |exponent| <= 22: one IEEE multiplication or division), otherwise Eisel and
// Lemire's method (the top bits of a 64 x 128-bit product with a tabulated power of five decide the rounding; Lemire,
// "Number parsing at a gigabyte per second", SPE 2021).  Anything else -- more digits, subnormal or overflowing results,
// inf / nan spellings, whatever strtod accepts beyond that -- returns false and the caller asks strtod, so the result is
// the correctly rounded double in every case (tests/test_textin_cpu.py compares with Python's float()).
#pragma once
#include <stdint.h>
#include <string.h>

#include "r3d_pow10_tables.inc"

namespace r3d {

namespace strtod_detail {

struct U128 { uint64_t hi, lo; };
static const U128 kPow5_128[651] = {R3D_POW5_128_TABLE};

inline U128 mul64(uint64_t a, uint64_t b) {
    const unsigned __int128 p = (unsigned __int128)a * b;
    return U128{(uint64_t)(p >> 64), (uint64_t)p};
}

// w * 10^q, w != 0, -342 <= q <= 308, rounded to nearest even.  false: the result is subnormal, zero or infinite (the
// caller's fallback decides).
inline bool eisel_lemire(uint64_t w, int q, double* out) {
    const int lz = __builtin_clzll(w);
    w <<= lz;
    const U128& t = kPow5_128[q + 342];
    U128 prod = mul64(w, t.hi);
    if ((prod.hi & 0x1ffull) == 0x1ffull) {                  // the low part of the power can still change the rounding bits
        const U128 second = mul64(w, t.lo);
        prod.lo += second.hi;
        if (second.hi > prod.lo) ++prod.hi;
    }
    const int upperbit = (int)(prod.hi >> 63);
    uint64_t mantissa = prod.hi >> (upperbit + 9);           // 54 bits: 53 + a rounding bit
    const int power2 = ((217706 * q) >> 16) + 63 + upperbit - lz + 1023;
    if (power2 <= 0 || power2 >= 0x7ff) return false;
    // exactly half way between two doubles (only possible for small |q|): round to even, not up
    if (prod.lo <= 1 && q >= -4 && q <= 23 && (mantissa & 3ull) == 1ull && (mantissa << (upperbit + 9)) == prod.hi) mantissa &= ~1ull;
    mantissa += mantissa & 1ull;
    mantissa >>= 1;
    int e2 = power2;
    if (mantissa >= (2ull << 52)) { mantissa = 1ull << 52; ++e2; }
    if (e2 >= 0x7ff) return false;
    const uint64_t bits = (mantissa & ~(1ull << 52)) | ((uint64_t)e2 << 52);
    memcpy(out, &bits, 8);
    return true;
}

static const double kExact10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

}  // namespace strtod_detail

// [p, e): one field without surrounding blanks.  true: *out is the correctly rounded value of a plain decimal literal.
inline bool parse_double(const char* p, const char* e, double* out) {
    using namespace strtod_detail;
    if (p >= e) return false;
    bool neg = false;
    if (*p == '-' || *p == '+') { neg = *p == '-'; ++p; }
    uint64_t w = 0;
    int digits = 0;          // significant digits accumulated in w (leading zeros not counted)
    int exp10 = 0;
    bool any = false;
    const char* s = p;
    while (s < e && (unsigned)(*s - '0') < 10u) {
        if (w || *s != '0') {
            if (digits >= 19) return false;
            w = w * 10u + (unsigned)(*s - '0');
            ++digits;
        }
        any = true;
        ++s;
    }
    if (s < e && *s == '.') {
        ++s;
        while (s < e && (unsigned)(*s - '0') < 10u) {
            if (w || *s != '0') {
                if (digits >= 19) return false;
                w = w * 10u + (unsigned)(*s - '0');
                ++digits;
            }
            --exp10;
            any = true;
            ++s;
        }
    }
    if (!any) return false;
    if (s < e && (*s == 'e' || *s == 'E')) {
        ++s;
        bool eneg = false;
        if (s < e && (*s == '-' || *s == '+')) { eneg = *s == '-'; ++s; }
        if (s >= e) return false;
        int x = 0;
        while (s < e && (unsigned)(*s - '0') < 10u) {
            if (x < 100000) x = x * 10 + (*s - '0');
            ++s;
        }
        exp10 += eneg ? -x : x;
    }
    if (s != e) return false;                                // trailing characters: not a plain decimal literal
    double v;
    if (w == 0) v = 0.0;
    else if (w <= (1ull << 53) && exp10 >= -22 && exp10 <= 22) {
        v = (double)w;
        if (exp10 < 0) v /= kExact10[-exp10]; else v *= kExact10[exp10];
    } else {
        if (exp10 < -342 || exp10 > 308) return false;
        if (!eisel_lemire(w, exp10, &v)) return false;
    }
    *out = neg ? -v : v;
    return true;
}

}  // namespace r3d
