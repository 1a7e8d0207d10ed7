// r3d_text.cu -- K6: the ASCII side of the path on the GPU.
//
// genply / genply_RGB / genply_noRGB (transfer/camera_to_world.py:112-134, transfer/pixel_to_camera.py:55-124) format
// every point with "%.4f" in a Python loop (1.2 s per 1242x375 frame in the reference); here a thread formats three
// consecutive rows (one for the str(float64) rows of the txt files), byte for byte what Python produces: the count of 1e-4
// units of a coordinate is one fused multiply-add (r3d_math.cuh::fixed4_units_fma -- the exact product rounded to an
// integer, ties to even, which is "%.4f"'s rule), the digits 32-bit multiply-shifts; values beyond 429 496 m, infinities
// and NaN take the exact integer / bignum path (r3d_math.cuh::fixed4_*).  Rows have different lengths, so the text of a
// tile (768 PLY rows) can only be placed once the length of everything before it is known.  ONE pass: a tile measures its
// rows, publishes its byte count, obtains its offset by a decoupled look-back over the tiles before it (Merrill & Garland:
// one 64-bit word of {status, value} per tile, tiles handed out by an atomic counter so that every predecessor is already
// running), formats each row ONCE into a shared-memory staging buffer at its in-tile offset, and copies the buffer out
// with coalesced 16-byte stores.  HBM-bound in principle (24 B of coordinates in, ~27 B of text out per point); measured
// at 0.44 of the HBM peak, bound by issue slots and the barrier behind the look-back.
#include <cub/block/block_scan.cuh>

#include "r3d_common.cuh"
#include "r3d_repr.cuh"

namespace r3d {

#ifndef K6_THREADS
#define K6_THREADS 256
#endif
#ifndef K6_PLY_ROWS
#define K6_PLY_ROWS 3      // PLY rows a thread formats per tile (measured on 29.8 M rows: 1 -> 40, 2 -> 49, 3 -> 53 G rows/s)
#endif
constexpr int K6_STAGE_TXT = 24 * 1024;       // a tile whose text is longer (astronomic coordinates) is written row by row
#ifndef K6_STAGE_PLY_ROW
#define K6_STAGE_PLY_ROW 48
#endif
// K6_STAGE_PLY_ROW: staged bytes per PLY row of a tile (rows average 27 bytes, 41 with colours)

struct TextArgs {
    const double *x, *y, *z;       // coordinate i at x[i * stride] ...
    unsigned long long stride;
    const unsigned char* rgb;      // n x 3 or nullptr
    int z_int;                     // txt rows: Z printed as an integer
    unsigned long long n;
    unsigned long long* tile_state;  // [n_tiles]: status << 62 | byte count (aggregate of the tile, or inclusive prefix)
    unsigned* tile_counter;          // next tile to hand out
    unsigned long long* total;       // bytes of all rows (written by the last tile)
    char* out;
    unsigned long long cap;          // bytes `out` holds; a tile that would write past it writes nothing
};

constexpr unsigned long long kTileAggregate = 1ull << 62, kTilePrefix = 2ull << 62, kTileValue = (1ull << 62) - 1ull;

// (fast4_measure / fast4_write, the 32-bit "%.4f" path: r3d_math.cuh, shared with the host test harness)

// kTxt: "X,Y,Z\n" rows with str(float64) fields (r3d_repr.cuh) instead of PLY rows.  A thread formats kRows consecutive rows.
template <bool kTxt, int kRows>
__global__ void __launch_bounds__(K6_THREADS) k6_rows(const TextArgs a) {
    typedef cub::BlockScan<unsigned, K6_THREADS> Scan;
    constexpr unsigned kTile = K6_THREADS * kRows;
    constexpr int K6_STAGE_BYTES = kTxt ? K6_STAGE_TXT : K6_STAGE_PLY_ROW * (int)kTile;
    __shared__ typename Scan::TempStorage tmp;
    __shared__ __align__(16) char stage[K6_STAGE_BYTES + 16];
    __shared__ unsigned s_tile;
    __shared__ unsigned long long s_base;
    const unsigned long long n_tiles = (a.n + kTile - 1) / kTile;
    const bool rgb = a.rgb != nullptr;
    const unsigned lane = threadIdx.x & 31u;
    for (;;) {
        if (threadIdx.x == 0) s_tile = atomicAdd(a.tile_counter, 1u);
        __syncthreads();
        const unsigned long long t = s_tile;
        if (t >= n_tiles) break;
        const unsigned long long i0 = t * kTile + (unsigned long long)threadIdx.x * kRows;
        unsigned len[kRows];
        unsigned mine = 0;
        // txt rows (one per thread): the shortest digits of the three fields are kept (Ryu is too long to run twice), the
        // characters are generated once, straight into the staging buffer
        ReprDec rd[3];
        // PLY rows: units, signs and lengths of the three coordinates (fast path), colour bytes
        uint32_t q[kRows][3], sg[kRows], ln[kRows], col[kRows];
        bool fast[kRows];
#pragma unroll
        for (int k = 0; k < kRows; ++k) {
            len[k] = 0; fast[k] = true; sg[k] = 0; ln[k] = 0; col[k] = 0;
            q[k][0] = q[k][1] = q[k][2] = 0;
            const unsigned long long i = i0 + k;
            if (i < a.n) {
                const double x = a.x[i * a.stride], y = a.y[i * a.stride], z = a.z[i * a.stride];
                if (kTxt) {
                    rd[0] = repr_decompose(x); rd[1] = repr_decompose(y);
                    rd[2] = a.z_int ? repr_decompose_integer(z) : repr_decompose(z);
                    len[k] = (unsigned)(repr_len(rd[0]) + repr_len(rd[1]) + repr_len(rd[2])) + 3u;
                } else {
                    uint32_t nx, ny, nz;
                    unsigned lx, ly, lz;
                    const bool fx = fast4_measure(x, q[k][0], nx, lx), fy = fast4_measure(y, q[k][1], ny, ly), fz = fast4_measure(z, q[k][2], nz, lz);
                    fast[k] = fx && fy && fz;
                    sg[k] = nx | (ny << 1) | (nz << 2);
                    ln[k] = lx | (ly << 8) | (lz << 16);
                    len[k] = lx + ly + lz + 4u;
                    if (!fast[k]) len[k] = (unsigned)(fixed4_len(x) + fixed4_len(y) + fixed4_len(z)) + 4u;
                    if (rgb) {
                        const unsigned r = a.rgb[3 * i], g = a.rgb[3 * i + 1], b = a.rgb[3 * i + 2];
                        col[k] = r | (g << 8) | (b << 16);
                        len[k] += (unsigned)(u8_len(r) + u8_len(g) + u8_len(b)) + 4u;     // "r g b 0" after the third blank
                    }
                }
            }
            mine += len[k];
        }
        unsigned off, total;
        Scan(tmp).ExclusiveSum(mine, off, total);
        // ---- decoupled look-back (warp 0): publish this tile's byte count, add up the tiles before it
        if (threadIdx.x < 32) {
            if (lane == 0) atomicExch(a.tile_state + t, (t == 0 ? kTilePrefix : kTileAggregate) | (unsigned long long)total);
            unsigned long long excl = 0;
            long long look = (long long)t - 1;
            while (look >= 0) {
                const long long idx = look - (long long)lane;
                unsigned long long v = kTilePrefix;                       // before the first tile: an empty prefix
                if (idx >= 0) {
                    do { v = *reinterpret_cast<volatile unsigned long long*>(a.tile_state + idx); } while ((v >> 62) == 0ull);
                }
                const unsigned pm = __ballot_sync(0xffffffffu, (v >> 62) == 2ull);
                const unsigned upto = pm ? (unsigned)(__ffs(pm) - 1) : 31u;   // nearest tile that knows its prefix
                unsigned long long c = lane <= upto ? (v & kTileValue) : 0ull;
                for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
                excl += c;
                if (pm) break;
                look -= 32;
            }
            if (lane == 0) {
                s_base = excl;
                if (t != 0) atomicExch(a.tile_state + t, kTilePrefix | (excl + total));
                if (t + 1 == n_tiles) *a.total = excl + total;
            }
        }
        __syncthreads();
        const unsigned long long base = s_base;
        const bool fits = a.out != nullptr && base + total <= a.cap;
        const bool staged = fits && total <= (unsigned)K6_STAGE_BYTES;
        // stage at (base % 16) so that the copy below moves aligned 16-byte words; a tile whose text is longer than the
        // stage (astronomic coordinates) writes its rows straight to the output
        const unsigned skew = (unsigned)(base & 15ull);
        if (staged) {
            // (a pointer into the stage alone: the stores below are shared-memory stores with 32-bit addresses)
            char* p = stage + skew + off;
#pragma unroll
            for (int k = 0; k < kRows; ++k) {
                if (len[k] == 0) continue;
                if (kTxt) {
                    repr_write(rd[0], p); p += repr_len(rd[0]); *p++ = ',';
                    repr_write(rd[1], p); p += repr_len(rd[1]); *p++ = ',';
                    repr_write(rd[2], p); p += repr_len(rd[2]); *p++ = '\n';
                } else if (fast[k]) {
                    p = fast4_write(q[k][0], sg[k] & 1u, ln[k] & 255u, p, ' ');
                    p = fast4_write(q[k][1], (sg[k] >> 1) & 1u, (ln[k] >> 8) & 255u, p, ' ');
                    p = fast4_write(q[k][2], sg[k] >> 2, ln[k] >> 16, p, ' ');
                    if (rgb) {
                        p = u8_write(p, col[k] & 255u); *p++ = ' ';
                        p = u8_write(p, (col[k] >> 8) & 255u); *p++ = ' ';
                        p = u8_write(p, col[k] >> 16); *p++ = ' ';
                        *p++ = '0';
                    }
                    *p++ = '\n';
                } else {
                    const unsigned long long i = i0 + k;
                    ply_row_write(p, a.x[i * a.stride], a.y[i * a.stride], a.z[i * a.stride], rgb, col[k] & 255u, (col[k] >> 8) & 255u, col[k] >> 16);
                    p += len[k];
                }
            }
        } else if (fits) {
            char* p = a.out + base + off;
#pragma unroll 1
            for (int k = 0; k < kRows; ++k) {
                if (len[k] == 0) continue;
                if (kTxt) {
                    char* w = p;
                    repr_write(rd[0], w); w += repr_len(rd[0]); *w++ = ',';
                    repr_write(rd[1], w); w += repr_len(rd[1]); *w++ = ',';
                    repr_write(rd[2], w); w += repr_len(rd[2]); *w++ = '\n';
                } else {
                    const unsigned long long i = i0 + k;
                    ply_row_write(p, a.x[i * a.stride], a.y[i * a.stride], a.z[i * a.stride], rgb, col[k] & 255u, (col[k] >> 8) & 255u, col[k] >> 16);
                }
                p += len[k];
            }
        }
        if (staged) {
            __syncthreads();
            char* dst = a.out + (base - skew);
            const unsigned span = skew + total;
            const unsigned full = span / 16u;
            for (unsigned w = threadIdx.x; w < full; w += K6_THREADS) {
                if (w == 0 && skew) {   // first word is shared with the previous tile: byte stores
                    for (unsigned k = skew; k < 16u; ++k) dst[k] = stage[k];
                } else {
                    *reinterpret_cast<uint4*>(dst + 16u * w) = *reinterpret_cast<const uint4*>(stage + 16u * w);
                }
            }
            // tail (shared with the next tile), and the head when the tile is shorter than one word
            for (unsigned k = full * 16u + threadIdx.x; k < span; k += K6_THREADS)
                if (k >= skew) dst[k] = stage[k];
        }
        __syncthreads();
    }
}

}  // namespace r3d

using namespace r3d;

static int format_rows(r3d_ctx* ctx, bool txt, int z_int, const double* x, const double* y, const double* z, size_t stride, uint64_t n,
                       const uint8_t* rgb, char* out, size_t cap, size_t* len) {
    if (!ctx) return set_error(nullptr, R3D_ERR_ARG, "null context");
    if (!len) return set_error(ctx, R3D_ERR_ARG, "null length pointer");
    *len = 0;
    if (n == 0) return R3D_OK;
    if (!x || !y || !z || stride == 0) return set_error(ctx, R3D_ERR_ARG, "r3d_format_*_rows: null coordinates");
    DeviceSetter ds(ctx->device);
    // coordinates: device pointers are used in place; host arrays are staged as one block [min(x,y,z), max + span)
    TextArgs a;
    memset(&a, 0, sizeof a);
    a.n = n; a.stride = stride; a.z_int = z_int;
    const bool dev = is_device_ptr(x);
    if (dev != is_device_ptr(y) || dev != is_device_ptr(z)) return set_error(ctx, R3D_ERR_ARG, "coordinate arrays must live on the same side");
    if (dev) { a.x = x; a.y = y; a.z = z; }
    else {
        const size_t span = ((size_t)(n - 1) * stride + 1) * sizeof(double);
        const double* lo = x < y ? (x < z ? x : z) : (y < z ? y : z);
        const double* hi = x > y ? (x > z ? x : z) : (y > z ? y : z);
        const size_t whole = (size_t)((const char*)hi - (const char*)lo) + span;
        if (whole <= 3 * span + 64) {   // interleaved or adjacent arrays: one copy
            R3D_TRY(scratch_reserve(ctx, SCR_IN0, whole));
            R3D_CUDA_OK(ctx, cudaMemcpyAsync(ctx->scratch[SCR_IN0], lo, whole, cudaMemcpyHostToDevice, ctx->stream));
            const char* d = (const char*)ctx->scratch[SCR_IN0];
            a.x = (const double*)(d + ((const char*)x - (const char*)lo));
            a.y = (const double*)(d + ((const char*)y - (const char*)lo));
            a.z = (const double*)(d + ((const char*)z - (const char*)lo));
        } else {
            R3D_TRY(scratch_reserve(ctx, SCR_IN0, 3 * span));
            char* d = (char*)ctx->scratch[SCR_IN0];
            R3D_CUDA_OK(ctx, cudaMemcpyAsync(d, x, span, cudaMemcpyHostToDevice, ctx->stream));
            R3D_CUDA_OK(ctx, cudaMemcpyAsync(d + span, y, span, cudaMemcpyHostToDevice, ctx->stream));
            R3D_CUDA_OK(ctx, cudaMemcpyAsync(d + 2 * span, z, span, cudaMemcpyHostToDevice, ctx->stream));
            a.x = (const double*)d; a.y = (const double*)(d + span); a.z = (const double*)(d + 2 * span);
        }
    }
    if (rgb) {
        if (is_device_ptr(rgb)) a.rgb = rgb;
        else {
            R3D_TRY(scratch_reserve(ctx, SCR_IN1, (size_t)n * 3));
            R3D_CUDA_OK(ctx, cudaMemcpyAsync(ctx->scratch[SCR_IN1], rgb, (size_t)n * 3, cudaMemcpyHostToDevice, ctx->stream));
            a.rgb = (const unsigned char*)ctx->scratch[SCR_IN1];
        }
    }
    const unsigned long long tile_rows = (unsigned long long)K6_THREADS * (txt ? 1 : K6_PLY_ROWS);
    const unsigned long long n_tiles = (n + tile_rows - 1) / tile_rows;
    R3D_TRY(scratch_reserve(ctx, SCR_TILE, (size_t)(n_tiles + 4) * 8 + 256));
    unsigned long long* state = (unsigned long long*)ctx->scratch[SCR_TILE];
    a.tile_state = state + 2;
    a.total = state;
    a.tile_counter = (unsigned*)(state + 1);
    R3D_CUDA_OK(ctx, cudaMemsetAsync(state, 0, (size_t)(n_tiles + 4) * 8, ctx->stream));
    // where the rows go: the caller's device buffer, or (host buffer) device scratch of the same capacity, clipped to the
    // longest the rows can be
    const size_t row_max = txt ? (size_t)kTxtRowMax : (size_t)(rgb ? 1000 : 990);   // "%.4f" of 1.8e308 has 314 characters
    size_t room = cap;
    if ((double)n * (double)row_max < (double)room) room = (size_t)n * row_max;
    const bool out_dev = out && is_device_ptr(out);
    char* d_out = out;
    if (out && !out_dev) {
        R3D_TRY(scratch_reserve(ctx, SCR_OUT0, room + 64));
        d_out = (char*)ctx->scratch[SCR_OUT0];
    } else if (out && ((uintptr_t)out & 15u) != 0) {
        return set_error(ctx, R3D_ERR_ARG, "device text buffer must be 16-byte aligned");
    }
    a.out = d_out;
    a.cap = out ? room : 0;
    int per_sm = 0;
    if (txt) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k6_rows<true, 1>, K6_THREADS, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k6_rows<false, K6_PLY_ROWS>, K6_THREADS, 0);
    if (per_sm < 1) per_sm = 1;
    // persistent CTAs, all resident (the look-back spins on tiles handed out earlier: they must be running)
    unsigned long long grid = (unsigned long long)ctx->sm_count * per_sm;
    if (grid > n_tiles) grid = n_tiles;
    if (txt) k6_rows<true, 1><<<(unsigned)grid, K6_THREADS, 0, ctx->stream>>>(a);
    else k6_rows<false, K6_PLY_ROWS><<<(unsigned)grid, K6_THREADS, 0, ctx->stream>>>(a);
    ctx->launches++;
    R3D_CUDA_OK(ctx, cudaGetLastError());
    unsigned long long total = 0;
    R3D_CUDA_OK(ctx, cudaMemcpyAsync(ctx->pinned, a.total, 8, cudaMemcpyDeviceToHost, ctx->stream));
    R3D_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(&total, ctx->pinned, 8);
    *len = (size_t)total;
    if (!out || cap < total) return R3D_OK;      // size query (or buffer too small): *len tells what is needed
    if (!out_dev) {
        R3D_CUDA_OK(ctx, cudaMemcpyAsync(out, d_out, (size_t)total, cudaMemcpyDeviceToHost, ctx->stream));
        R3D_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return finish(ctx);
}

extern "C" int r3d_format_ply_rows(r3d_ctx* ctx, const double* x, const double* y, const double* z, size_t stride, uint64_t n,
                                   const uint8_t* rgb, char* out, size_t cap, size_t* len) {
    return format_rows(ctx, false, 0, x, y, z, stride, n, rgb, out, cap, len);
}

extern "C" int r3d_format_txt_rows(r3d_ctx* ctx, const double* x, const double* y, const double* z, size_t stride, uint64_t n,
                                   int z_is_integer, char* out, size_t cap, size_t* len) {
    return format_rows(ctx, true, z_is_integer ? 1 : 0, x, y, z, stride, n, nullptr, out, cap, len);
}
