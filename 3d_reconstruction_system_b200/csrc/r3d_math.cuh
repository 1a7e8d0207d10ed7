// r3d_math.cuh -- scalar arithmetic shared by every kernel, written so the SAME source also compiles for
// the host (tests/hostmath builds it with g++ to check the formulas against the oracle without a GPU;
// that build is test infrastructure, the product only ever runs the device code).
//
// Every floating-point operation whose rounding is part of the parity contract goes through the
// explicit round-to-nearest wrappers below so that no FMA contraction can change a result
// (device: __dmul_rn & co; host: plain operators, built with -ffp-contract=off).
#pragma once
#include <math.h>
#include <stdint.h>
#include <float.h>
#include <string.h>

#if defined(__CUDACC__)
#define R3D_HD __host__ __device__ __forceinline__
#else
#define R3D_HD inline
#endif

namespace r3d {

// ---------------------------------------------------------------- rounding-explicit ops
R3D_HD double dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
R3D_HD double dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
R3D_HD double dsub(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}
R3D_HD double ddiv(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}
R3D_HD float fmul(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
R3D_HD float fadd(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
R3D_HD float fsub(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fsub_rn(a, b);
#else
    return a - b;
#endif
}
R3D_HD float fdiv(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}
R3D_HD double dsqrt(double a) {
#if defined(__CUDA_ARCH__)
    return __dsqrt_rn(a);
#else
    return sqrt(a);
#endif
}

// ---------------------------------------------------------------- K1: back-projection + pose
// point_camera (transfer/camera_to_world.py:57-59): w = Rinv . (p - t), rows evaluated left to right.
// rt = Rinv row-major (9) followed by t (3).
struct Pose {
    double r00, r01, r02, r10, r11, r12, r20, r21, r22, t0, t1, t2;
};
R3D_HD void pose_load(const double* __restrict__ rt, Pose& p) {
    p.r00 = rt[0]; p.r01 = rt[1]; p.r02 = rt[2];
    p.r10 = rt[3]; p.r11 = rt[4]; p.r12 = rt[5];
    p.r20 = rt[6]; p.r21 = rt[7]; p.r22 = rt[8];
    p.t0 = rt[9]; p.t1 = rt[10]; p.t2 = rt[11];
}
R3D_HD void pose_apply(const Pose& p, double X, double Y, double Z, double& wx, double& wy, double& wz) {
    const double d0 = dsub(X, p.t0), d1 = dsub(Y, p.t1), d2 = dsub(Z, p.t2);
    wx = dadd(dadd(dmul(p.r00, d0), dmul(p.r01, d1)), dmul(p.r02, d2));
    wy = dadd(dadd(dmul(p.r10, d0), dmul(p.r11, d1)), dmul(p.r12, d2));
    wz = dadd(dadd(dmul(p.r20, d0), dmul(p.r21, d1)), dmul(p.r22, d2));
}
// The reference's sum is np.dot (BLAS): its accumulators start at +0.0, so a sum of signed zeros is +0.0, never -0.0
// (e.g. the world x of a Z = 0 pixel under an axis-aligned pose: "0.0000" in the reference's PLY, not "-0.0000").
// w + 0.0 is exact for every other value.  Applied where the float64 value is what the caller sees (text writers);
// float32 records feed voxel keys, where the sign of zero is immaterial.
R3D_HD double pose_canon(double w) { return dadd(w, 0.0); }
// gentxtcord tables (transfer/camera_to_world.py:77-78): ((i - cx) / fx) and ((j - cy) / fy)
R3D_HD double pixel_coeff(int i, double c, double f) { return ddiv(dsub((double)i, c), f); }

// depth decode: a1/a2 (Z = raw * depth_scale) and a3 (Z = fB / (raw*depth_scale), d <= 0 -> 0)
R3D_HD double decode_z(double raw, int mode, double depth_scale, double fB, bool& valid) {
    const double d = dmul(raw, depth_scale);
    valid = (d > 0.0) && (d <= DBL_MAX);
    if (mode == 0) return d;
    return (d > 0.0) ? ddiv(fB, d) : 0.0;
}

// ---------------------------------------------------------------- OctoMap keys (a10)
constexpr int kTreeDepth = 16;
constexpr int kTreeMaxVal = 32768;

// coordToKeyChecked(double coordinate): (int)floor(resolution_factor * coordinate) + tree_max_val in [0, 65536)
R3D_HD bool coord_to_key(double res_factor, float x, uint16_t& key) {
    const double f = floor(dmul(res_factor, (double)x));
    if (!(f >= -32768.0 && f < 32768.0)) return false;  // also rejects NaN / inf like the x86 cast does
    key = (uint16_t)((int)f + kTreeMaxVal);
    return true;
}
R3D_HD bool coord_to_key3(double res_factor, float x, float y, float z, uint16_t& kx, uint16_t& ky, uint16_t& kz) {
    return coord_to_key(res_factor, x, kx) & coord_to_key(res_factor, y, ky) & coord_to_key(res_factor, z, kz);
}
// keyToCoord(key) = (double(int(key) - tree_max_val) + 0.5) * resolution
R3D_HD double key_to_coord(double res, uint16_t key) { return dmul(dadd((double)((int)key - kTreeMaxVal), 0.5), res); }

// updateNodeLogOdds (float add, clamp to [min, max])
R3D_HD float clamped_add(float v, float upd, float cmin, float cmax) {
    v = fadd(v, upd);
    if (v < cmin) return cmin;
    if (v > cmax) return cmax;
    return v;
}

// ---------------------------------------------------------------- bricks
// A brick is the 8x8x8 voxel block under one depth-13 node.  brick key = bx | by<<13 | bz<<26.
// Voxel index inside a brick = 9-bit Morton code, child-index convention (x lowest bit per level), so that
// the 8 children of a depth-15 node are 8 consecutive voxels and a depth-14 node is 64 consecutive voxels.
R3D_HD uint32_t spread3(uint32_t v) { return (v | (v << 2) | (v << 4)) & 0x49u; }
R3D_HD uint32_t brick_voxel_index(uint32_t kx, uint32_t ky, uint32_t kz) {
    return spread3(kx & 7u) | (spread3(ky & 7u) << 1) | (spread3(kz & 7u) << 2);
}
R3D_HD uint32_t compact3(uint32_t m) { m &= 0x49u; return (m | (m >> 2) | (m >> 4)) & 7u; }
R3D_HD void brick_voxel_coords(uint32_t idx, uint32_t& x, uint32_t& y, uint32_t& z) {
    x = compact3(idx); y = compact3(idx >> 1); z = compact3(idx >> 2);
}
R3D_HD uint64_t brick_key(uint32_t kx, uint32_t ky, uint32_t kz) {
    return (uint64_t)(kx >> 3) | ((uint64_t)(ky >> 3) << 13) | ((uint64_t)(kz >> 3) << 26);
}
R3D_HD void brick_key_unpack(uint64_t bk, uint32_t& bx, uint32_t& by, uint32_t& bz) {
    bx = (uint32_t)(bk & 0x1fffu); by = (uint32_t)((bk >> 13) & 0x1fffu); bz = (uint32_t)((bk >> 26) & 0x1fffu);
}
// 39-bit Morton code of the brick coordinates (13 levels, x lowest): pre-order position of the depth-13 node
R3D_HD uint64_t spread13(uint64_t v) {
    uint64_t r = 0;
    for (int i = 0; i < 13; ++i) r |= ((v >> i) & 1ull) << (3 * i);
    return r;
}
R3D_HD uint64_t brick_morton(uint64_t bk) {
    uint32_t bx, by, bz;
    brick_key_unpack(bk, bx, by, bz);
    return spread13(bx) | (spread13(by) << 1) | (spread13(bz) << 2);
}
R3D_HD uint64_t hash64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

// which of `nparts` GPUs owns a brick (multi-GPU apply); high hash bits, the table slots use the low ones
R3D_HD uint32_t brick_owner(uint64_t bk, uint32_t nparts) { return (uint32_t)((hash64(bk) >> 32) % nparts); }

// ---------------------------------------------------------------- 3-D DDA (a11: computeRayKeys)
// Mixed float / double exactly as upstream: direction, length, origin are float; tMax, tDelta and the voxel
// border are double; the half-voxel offset is rounded through float.
struct Ray {
    int kx, ky, kz;        // current key
    int ex, ey, ez;        // end key
    int sx, sy, sz;        // step
    double tmx, tmy, tmz;  // tMax
    double tdx, tdy, tdz;  // tDelta
    float length;
};
// returns: -1 origin or end out of bounds (no free cells), 0 same cell (empty ray), 1 walk (first key = origin key)
R3D_HD int ray_setup(double res, double res_factor, float ox, float oy, float oz, float ex, float ey, float ez, Ray& r) {
    uint16_t a, b, c, d, e, f;
    if (!coord_to_key3(res_factor, ox, oy, oz, a, b, c) || !coord_to_key3(res_factor, ex, ey, ez, d, e, f)) return -1;
    r.kx = a; r.ky = b; r.kz = c; r.ex = d; r.ey = e; r.ez = f;
    if (a == d && b == e && c == f) return 0;
    float dx = fsub(ex, ox), dy = fsub(ey, oy), dz = fsub(ez, oz);
    const float nsq = fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz));
    r.length = (float)dsqrt((double)nsq);
    dx = fdiv(dx, r.length); dy = fdiv(dy, r.length); dz = fdiv(dz, r.length);
#define R3D_AXIS(dir, step, key, org, tmax, tdelta)                                        \
    if (dir > 0.0f) step = 1; else if (dir < 0.0f) step = -1; else step = 0;                \
    if (step != 0) {                                                                        \
        double border = key_to_coord(res, (uint16_t)key);                                   \
        border = dadd(border, (double)(float)dmul(dmul((double)step, res), 0.5));           \
        tmax = ddiv(dsub(border, (double)org), (double)dir);                                \
        tdelta = ddiv(res, fabs((double)dir));                                              \
    } else { tmax = DBL_MAX; tdelta = DBL_MAX; }
    R3D_AXIS(dx, r.sx, r.kx, ox, r.tmx, r.tdx)
    R3D_AXIS(dy, r.sy, r.ky, oy, r.tmy, r.tdy)
    R3D_AXIS(dz, r.sz, r.kz, oz, r.tmz, r.tdz)
#undef R3D_AXIS
    return 1;
}
// One DDA step.  Returns true when the new current key is a free cell to record, false when the ray is done.
R3D_HD bool ray_step(Ray& r) {
    if (r.tmx < r.tmy) {
        if (r.tmx < r.tmz) { r.kx += r.sx; r.tmx = dadd(r.tmx, r.tdx); }
        else               { r.kz += r.sz; r.tmz = dadd(r.tmz, r.tdz); }
    } else {
        if (r.tmy < r.tmz) { r.ky += r.sy; r.tmy = dadd(r.tmy, r.tdy); }
        else               { r.kz += r.sz; r.tmz = dadd(r.tmz, r.tdz); }
    }
    r.kx &= 0xffff; r.ky &= 0xffff; r.kz &= 0xffff;   // key_type is uint16: wraps like upstream
    if (r.kx == r.ex && r.ky == r.ey && r.kz == r.ez) return false;
    const double m01 = r.tmx < r.tmy ? r.tmx : r.tmy;
    const double dist = m01 < r.tmz ? m01 : r.tmz;
    if (dist > (double)r.length) return false;
    return true;
}
// The same walk in branch-free form for the ray-casting kernel (lanes of a warp pick different axes every step, so
// per-axis branches would serialise).  ray_select picks the axis of the next step with upstream's comparison chain and
// returns tMax of that axis, which IS min(tMax) -- the value upstream compares with the ray length after each step.
//   first:  a = ray_select(r, t)                      (after ray_setup returned 1 and the origin key was recorded)
//   loop:   ray_advance(r, a); if (ray_at_end(r)) stop; a = ray_select(r, t); if (t > length) stop; record key
R3D_HD int ray_select(const Ray& r, double& tsel) {
    const bool xy = r.tmx < r.tmy, xz = r.tmx < r.tmz, yz = r.tmy < r.tmz;
    const bool selx = xy & xz, sely = (!xy) & yz;
    tsel = selx ? r.tmx : (sely ? r.tmy : r.tmz);
    return selx ? 0 : (sely ? 1 : 2);
}
R3D_HD void ray_advance(Ray& r, int axis) {
    const double nx = dadd(r.tmx, r.tdx), ny = dadd(r.tmy, r.tdy), nz = dadd(r.tmz, r.tdz);
    r.kx = (r.kx + (axis == 0 ? r.sx : 0)) & 0xffff;   // key_type is uint16: wraps like upstream
    r.ky = (r.ky + (axis == 1 ? r.sy : 0)) & 0xffff;
    r.kz = (r.kz + (axis == 2 ? r.sz : 0)) & 0xffff;
    r.tmx = axis == 0 ? nx : r.tmx;
    r.tmy = axis == 1 ? ny : r.tmy;
    r.tmz = axis == 2 ? nz : r.tmz;
}
R3D_HD bool ray_at_end(const Ray& r) { return ((r.kx ^ r.ex) | (r.ky ^ r.ey) | (r.kz ^ r.ez)) == 0; }

// computeUpdate's per-point prologue: decides whether the endpoint is an occupied cell and where the ray ends.
// maxrange < 0: unlimited.  Returns true when the endpoint (px,py,pz) is within range (occupied candidate);
// (ex,ey,ez) receives the ray end (the point itself, or origin + dir * maxrange).
R3D_HD bool scan_point_end(float ox, float oy, float oz, float px, float py, float pz, double maxrange,
                           float& ex, float& ey, float& ez) {
    ex = px; ey = py; ez = pz;
    if (maxrange < 0.0) return true;
    float dx = fsub(px, ox), dy = fsub(py, oy), dz = fsub(pz, oz);
    const float nsq = fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz));
    const double len = dsqrt((double)nsq);
    if (len <= maxrange) return true;
    if (len > 0) { const float fl = (float)len; dx = fdiv(dx, fl); dy = fdiv(dy, fl); dz = fdiv(dz, fl); }
    const float mr = (float)maxrange;
    ex = fadd(ox, fmul(dx, mr)); ey = fadd(oy, fmul(dy, mr)); ez = fadd(oz, fmul(dz, mr));
    return false;
}

}  // namespace r3d

// ---------------------------------------------------------------- K6: "%.4f" of a double, byte-exact (a7)
// genply's float_formatter (transfer/camera_to_world.py:117, pixel_to_camera.py:68,108): C / Python "%.4f", i.e. the
// exact binary value rounded half-to-even to 4 decimals.  x = m * 2^e exactly, so x * 10^4 = (m * 625) * 2^(e+4) with
// m * 625 < 2^63: one 64-bit product and a shift with an exact remainder give the correctly rounded integer q of
// 1e-4 units whenever |x| < 2^63 / 10^4; beyond that (x is then an integer multiple of 1/8 at least, no rounding
// happens) a base-10^9 bignum carries the digits.  inf -> "inf" / "-inf", nan -> "nan".
namespace r3d {

R3D_HD uint64_t double_bits(double x) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u; memcpy(&u, &x, 8); return u;
#endif
}

struct Fixed4 {
    uint64_t q;      // |x| in 1e-4 units, correctly rounded (valid when kind == 0)
    int kind;        // 0 finite & small, 1 finite & needs the bignum, 2 inf, 3 nan
    int neg;
    uint64_t v;      // m * 625 (kind 1)
    int sh;          // left shift (kind 1)
};

R3D_HD Fixed4 fixed4_decompose(double x) {
    Fixed4 f;
    const uint64_t b = double_bits(x);
    f.neg = (int)(b >> 63);
    const int ex = (int)((b >> 52) & 0x7ffu);
    const uint64_t frac = b & 0xfffffffffffffull;
    f.q = 0; f.v = 0; f.sh = 0; f.kind = 0;
    if (ex == 0x7ff) { f.kind = frac ? 3 : 2; if (frac) f.neg = 0; return f; }
    const uint64_t m = ex ? (frac | (1ull << 52)) : frac;
    const int e = ex ? ex - 1075 : -1074;
    const uint64_t v = m * 625ull;              // < 2^63
    const int sh = e + 4;
    if (sh >= 0) {
        int lz = 0;
        { uint64_t t = v; while (t && !(t >> 63)) { t <<= 1; ++lz; } if (!v) lz = 64; }
        if (sh < lz || v == 0) { f.q = v << sh; return f; }   // still < 2^64
        f.kind = 1; f.v = v; f.sh = sh;
        return f;
    }
    const int s = -sh;
    if (s >= 64) { f.q = 0; return f; }
    uint64_t q = v >> s;
    const uint64_t r = v & ((1ull << s) - 1ull), half = 1ull << (s - 1);
    if (r > half || (r == half && (q & 1ull))) ++q;
    f.q = q;
    return f;
}

// The same count of 1e-4 units for |x| < 429 496 (it fits 32 bits) as ONE fused multiply-add: fma(|x|, 1e4, 2^52) rounds
// the EXACT product to an integer, ties to even -- the rule above -- and the integer is the low word of the sum's mantissa.
constexpr double kFixed4Two52 = 4503599627370496.0;
constexpr double kFixed4FastLimit = 429496.0;
R3D_HD uint32_t fixed4_units_fma(double a) {   // a = |x| < kFixed4FastLimit
#ifdef __CUDA_ARCH__
    return (uint32_t)__double2loint(__fma_rn(a, 1e4, kFixed4Two52));
#else
    return (uint32_t)(double_bits(fma(a, 1e4, kFixed4Two52)) & 0xffffffffull);
#endif
}

R3D_HD int dec_digits_u64(uint64_t q) {
    int n = 1;
    while (q >= 10ull) { q /= 10ull; ++n; }
    return n;
}

// bignum: value = v << sh in base 1e9 limbs (little endian); returns the limb count
constexpr int kBigLimbs = 40;
R3D_HD int fixed4_big(uint64_t v, int sh, uint32_t* limb) {
    int n = 0;
    while (v) { limb[n++] = (uint32_t)(v % 1000000000ull); v /= 1000000000ull; }
    while (sh > 0) {
        const int step = sh > 29 ? 29 : sh;     // limb * 2^29 + carry < 2^64
        uint64_t carry = 0;
        for (int i = 0; i < n; ++i) {
            const uint64_t t = ((uint64_t)limb[i] << step) + carry;
            limb[i] = (uint32_t)(t % 1000000000ull);
            carry = t / 1000000000ull;
        }
        while (carry) { limb[n++] = (uint32_t)(carry % 1000000000ull); carry /= 1000000000ull; }
        sh -= step;
    }
    return n;
}

// number of characters of "%.4f" % x
R3D_HD int fixed4_len(double x) {
    const Fixed4 f = fixed4_decompose(x);
    if (f.kind == 3) return 3;
    if (f.kind == 2) return 3 + f.neg;
    if (f.kind == 0) {
        const int d = dec_digits_u64(f.q);
        return f.neg + (d > 4 ? d : 5) + 1;   // at least "0" before the point, 4 digits after it, the point
    }
    uint32_t limb[kBigLimbs];
    const int n = fixed4_big(f.v, f.sh, limb);
    const int d = (n - 1) * 9 + dec_digits_u64(limb[n - 1]);
    return f.neg + d + 1;
}

// writes exactly fixed4_len(x) characters at dst
R3D_HD void fixed4_write(double x, char* dst, int len) {
    const Fixed4 f = fixed4_decompose(x);
    if (f.kind == 3) { dst[0] = 'n'; dst[1] = 'a'; dst[2] = 'n'; return; }
    if (f.kind == 2) { if (f.neg) *dst++ = '-'; dst[0] = 'i'; dst[1] = 'n'; dst[2] = 'f'; return; }
    char* p = dst + len;                        // fill from the last digit backwards
    if (f.kind == 0) {
        uint64_t q = f.q;
        for (int i = 0; i < 4; ++i) { *--p = (char)('0' + (int)(q % 10ull)); q /= 10ull; }
        *--p = '.';
        do { *--p = (char)('0' + (int)(q % 10ull)); q /= 10ull; } while (q);
    } else {
        uint32_t limb[kBigLimbs];
        const int n = fixed4_big(f.v, f.sh, limb);
        int produced = 0;
        for (int i = 0; i < n; ++i) {
            uint32_t w = limb[i];
            const int cnt = (i == n - 1) ? dec_digits_u64(w) : 9;
            for (int k = 0; k < cnt; ++k) {
                *--p = (char)('0' + (int)(w % 10u)); w /= 10u;
                if (++produced == 4) *--p = '.';
            }
        }
    }
    if (f.neg) *--p = '-';
}

// "%.4f" of a value whose 1e-4 units fit 32 bits (|x| < 429 496 -- every coordinate of a metric map), K6's fast path: the
// units come from one fused multiply-add (fixed4_units_fma), the digits from 32-bit multiply-shifts.  tests/hostmath runs
// exactly this code on the host against Python's formatter.

// returns whether the fast path applies; q = units, neg = sign bit, len = characters of "%.4f"
R3D_HD bool fast4_measure(double v, uint32_t& q, uint32_t& neg, unsigned& len) {
    const double a = fabs(v);
    neg = (uint32_t)(double_bits(v) >> 63);
    q = fixed4_units_fma(a);
    len = neg + 6u + (q >= 100000u) + (q >= 1000000u) + (q >= 10000000u) + (q >= 100000000u) + (q >= 1000000000u);
    return a < kFixed4FastLimit;      // false for NaN
}
// writes the `len` characters at p and the separator after them; returns the position after the separator
R3D_HD char* fast4_write(uint32_t q, uint32_t neg, unsigned len, char* p, char sep) {
    char* const e = p + len;
    const uint32_t ip = q / 10000u;
    const uint32_t fr = q - ip * 10000u;
    const uint32_t hi = (fr * 5243u) >> 19, lo = fr - hi * 100u;          // fr / 100, fr % 100 (exact below 43 699)
    const uint32_t t1 = (hi * 205u) >> 11, t0 = (lo * 205u) >> 11;        // x / 10 for x < 1 029
    e[0] = sep;
    e[-1] = (char)('0' + lo - t0 * 10u);
    e[-2] = (char)('0' + t0);
    e[-3] = (char)('0' + hi - t1 * 10u);
    e[-4] = (char)('0' + t1);
    e[-5] = '.';
    if (q < 1000000u) {              // |x| < 100: one or two digits before the point
        const uint32_t t = (ip * 205u) >> 11;
        e[-6] = (char)('0' + ip - t * 10u);
        if (q >= 100000u) e[-7] = (char)('0' + t);
    } else {                         // three to six: all of them, stored where the number has them
        const uint32_t top = ip / 10000u;                                        // (ip < 429 497: top < 43)
        const uint32_t b = ip - top * 10000u;
        const uint32_t bh = (b * 5243u) >> 19, bl = b - bh * 100u;
        const uint32_t b3 = (bh * 205u) >> 11, b1 = (bl * 205u) >> 11;
        const uint32_t c1 = (top * 205u) >> 11;
        e[-6] = (char)('0' + bl - b1 * 10u);
        e[-7] = (char)('0' + b1);
        e[-8] = (char)('0' + bh - b3 * 10u);
        if (q >= 10000000u) e[-9] = (char)('0' + b3);
        if (q >= 100000000u) e[-10] = (char)('0' + top - c1 * 10u);
        if (q >= 1000000000u) e[-11] = (char)('0' + c1);
    }
    if (neg) *p = '-';
    return e + 1;
}

// one PLY row: "%.4f %.4f %.4f \n" (camera_to_world.py:118-119) or "%.4f %.4f %.4f %d %d %d 0\n" (pixel_to_camera.py:70-73)
R3D_HD int u8_len(unsigned v) { return v >= 100u ? 3 : (v >= 10u ? 2 : 1); }
R3D_HD int ply_row_len(double x, double y, double z, bool rgb, unsigned r, unsigned g, unsigned b) {
    const int n = fixed4_len(x) + fixed4_len(y) + fixed4_len(z);
    return rgb ? n + 3 + u8_len(r) + u8_len(g) + u8_len(b) + 3 + 2 : n + 4;
}
R3D_HD char* u8_write(char* p, unsigned v) {
    if (v >= 100u) *p++ = (char)('0' + v / 100u);
    if (v >= 10u) *p++ = (char)('0' + (v / 10u) % 10u);
    *p++ = (char)('0' + v % 10u);
    return p;
}
R3D_HD void ply_row_write(char* p, double x, double y, double z, bool rgb, unsigned r, unsigned g, unsigned b) {
    int n = fixed4_len(x); fixed4_write(x, p, n); p += n; *p++ = ' ';
    n = fixed4_len(y); fixed4_write(y, p, n); p += n; *p++ = ' ';
    n = fixed4_len(z); fixed4_write(z, p, n); p += n; *p++ = ' ';
    if (rgb) {
        p = u8_write(p, r); *p++ = ' ';
        p = u8_write(p, g); *p++ = ' ';
        p = u8_write(p, b); *p++ = ' ';
        *p++ = '0';
    }
    *p++ = '\n';
}

}  // namespace r3d
