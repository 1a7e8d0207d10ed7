// Host build of 3d_reconstruction_system_b200/csrc/r3d_math.cuh -- TEST INFRASTRUCTURE ONLY.
// The kernels' scalar arithmetic is written once for host and device; this wrapper lets the CPU test-suite check
// those formulas against the oracle where no GPU exists.  It is never loaded by the product package.
#include "../../3d_reconstruction_system_b200/csrc/r3d_math.cuh"
#include "../../3d_reconstruction_system_b200/csrc/r3d_repr.cuh"
#include <cstddef>
using namespace r3d;
extern "C" {
int hm_coord_to_key(double res_factor, float x, float y, float z, uint16_t* k) {
    return coord_to_key3(res_factor, x, y, z, k[0], k[1], k[2]) ? 1 : 0;
}
// returns -1 (out of bounds), else number of free keys written (origin key first), like computeRayKeys
long hm_ray_keys(double res, const float* o, const float* e, uint16_t* out, long cap) {
    Ray r;
    const int st = ray_setup(res, 1.0 / res, o[0], o[1], o[2], e[0], e[1], e[2], r);
    if (st < 0) return -1;
    if (st == 0) return 0;
    long n = 0;
    do {
        if (n < cap) { out[3 * n] = (uint16_t)r.kx; out[3 * n + 1] = (uint16_t)r.ky; out[3 * n + 2] = (uint16_t)r.kz; }
        ++n;
    } while (ray_step(r));
    return n;
}
// the branch-free walk the ray-casting kernel uses (ray_select / ray_advance): must list the same keys
long hm_ray_keys_predicated(double res, const float* o, const float* e, uint16_t* out, long cap) {
    Ray r;
    const int st = ray_setup(res, 1.0 / res, o[0], o[1], o[2], e[0], e[1], e[2], r);
    if (st < 0) return -1;
    if (st == 0) return 0;
    long n = 0;
    if (n < cap) { out[0] = (uint16_t)r.kx; out[1] = (uint16_t)r.ky; out[2] = (uint16_t)r.kz; }
    ++n;
    double t;
    int a = ray_select(r, t);
    for (;;) {
        ray_advance(r, a);
        if (ray_at_end(r)) break;
        a = ray_select(r, t);
        if (t > (double)r.length) break;
        if (n < cap) { out[3 * n] = (uint16_t)r.kx; out[3 * n + 1] = (uint16_t)r.ky; out[3 * n + 2] = (uint16_t)r.kz; }
        ++n;
    }
    return n;
}
// the loop order the dense kernel uses: choose the axis first, then decide about the key the previous advance reached
long hm_ray_keys_pending(double res, const float* o, const float* e, uint16_t* out, long cap) {
    Ray r;
    const int st = ray_setup(res, 1.0 / res, o[0], o[1], o[2], e[0], e[1], e[2], r);
    if (st < 0) return -1;
    if (st == 0) return 0;
    long n = 0;
    if (n < cap) { out[0] = (uint16_t)r.kx; out[1] = (uint16_t)r.ky; out[2] = (uint16_t)r.kz; }
    ++n;
    bool pending = false;
    for (;;) {
        double t;
        const int a = ray_select(r, t);
        if (pending) {
            if (ray_at_end(r) || t > (double)r.length) break;
            if (n < cap) { out[3 * n] = (uint16_t)r.kx; out[3 * n + 1] = (uint16_t)r.ky; out[3 * n + 2] = (uint16_t)r.kz; }
            ++n;
        }
        ray_advance(r, a);
        pending = true;
    }
    return n;
}
int hm_scan_point_end(const float* o, const float* p, double maxrange, float* e) {
    return scan_point_end(o[0], o[1], o[2], p[0], p[1], p[2], maxrange, e[0], e[1], e[2]) ? 1 : 0;
}
void hm_pose_apply(const double* rt, const double* p, double* w) {
    Pose ps; pose_load(rt, ps);
    pose_apply(ps, p[0], p[1], p[2], w[0], w[1], w[2]);
    for (int k = 0; k < 3; ++k) w[k] = pose_canon(w[k]);
}
double hm_pixel_coeff(int i, double c, double f) { return pixel_coeff(i, c, f); }
double hm_decode_z(double raw, int mode, double scale, double fB, int* valid) { bool v; double z = decode_z(raw, mode, scale, fB, v); *valid = v; return z; }
float hm_clamped_add(float v, float u, float lo, float hi) { return clamped_add(v, u, lo, hi); }
uint32_t hm_brick_voxel_index(uint32_t x, uint32_t y, uint32_t z) { return brick_voxel_index(x, y, z); }
void hm_brick_voxel_coords(uint32_t i, uint32_t* xyz) { brick_voxel_coords(i, xyz[0], xyz[1], xyz[2]); }
uint64_t hm_brick_key(uint32_t x, uint32_t y, uint32_t z) { return brick_key(x, y, z); }
uint64_t hm_brick_morton(uint64_t bk) { return brick_morton(bk); }
// K6: "%.4f" rows.  Formats n rows into out (rows back to back), returns the total length.
// K6: "X,Y,Z\n" rows with str(float64) fields (z_int: Z printed as an integer).  Returns the total length.
long hm_txt_rows(const double* xyz, long n, int z_int, char* out) {
    long off = 0;
    for (long i = 0; i < n; ++i) off += txt_row_write(out + off, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], z_int != 0);
    return off;
}
int hm_repr(double x, char* out) { return repr_double(x, out); }
// K6 fast path: 1e-4 units of |x| by one fused multiply-add, beside the integer statement of the rule.  Returns how many
// of the n values are in the fast range; mismatches are counted into *bad.
long hm_fixed4_units(const double* v, long n, long* bad) {
    long in_range = 0;
    *bad = 0;
    for (long i = 0; i < n; ++i) {
        const double a = fabs(v[i]);
        if (!(a < kFixed4FastLimit)) continue;
        ++in_range;
        const Fixed4 f = fixed4_decompose(v[i]);
        if (f.kind != 0 || f.q != (uint64_t)fixed4_units_fma(a)) ++*bad;
    }
    return in_range;
}
// K6's PLY row as the kernel writes it: the 32-bit fast path when all three coordinates are in range, else the exact path.
// *fast_rows counts the rows that took the fast path.
long hm_ply_rows_fast(const double* xyz, long n, char* out, long* fast_rows) {
    long off = 0;
    *fast_rows = 0;
    for (long i = 0; i < n; ++i) {
        const double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
        uint32_t q[3], ng[3];
        unsigned ln[3];
        const bool fx = fast4_measure(x, q[0], ng[0], ln[0]), fy = fast4_measure(y, q[1], ng[1], ln[1]), fz = fast4_measure(z, q[2], ng[2], ln[2]);
        char* p = out + off;
        if (fx && fy && fz) {
            ++*fast_rows;
            p = fast4_write(q[0], ng[0], ln[0], p, ' ');
            p = fast4_write(q[1], ng[1], ln[1], p, ' ');
            p = fast4_write(q[2], ng[2], ln[2], p, ' ');
            *p++ = '\n';
            off = p - out;
        } else {
            const int len = ply_row_len(x, y, z, false, 0, 0, 0);
            ply_row_write(p, x, y, z, false, 0, 0, 0);
            off += len;
        }
    }
    return off;
}
long hm_ply_rows(const double* xyz, long n, const unsigned char* rgb, char* out) {
    long off = 0;
    for (long i = 0; i < n; ++i) {
        const unsigned r = rgb ? rgb[3 * i] : 0, g = rgb ? rgb[3 * i + 1] : 0, b = rgb ? rgb[3 * i + 2] : 0;
        const int len = ply_row_len(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], rgb != nullptr, r, g, b);
        ply_row_write(out + off, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], rgb != nullptr, r, g, b);
        off += len;
    }
    return off;
}
}
