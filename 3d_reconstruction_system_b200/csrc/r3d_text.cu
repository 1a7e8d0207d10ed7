// r3d_text.cu -- K6: the ASCII side of the path on the GPU.
//
// genply / genply_RGB / genply_noRGB (transfer/camera_to_world.py:112-134, transfer/pixel_to_camera.py:55-124) format
// every point with "%.4f" in a Python loop (1.2 s per 1242x375 frame in the reference); here one thread formats one
// row, byte for byte what Python's "%.4f" produces (r3d_math.cuh::fixed4_*: exact integer arithmetic, round half to
// even on the exact binary value).  Rows have different lengths, so the text of a tile of 256 rows can only be placed
// once the length of everything before it is known.  ONE pass: a tile measures its rows, publishes its byte count,
// obtains its offset by a decoupled look-back over the tiles before it (Merrill & Garland: one 64-bit word of
// {status, value} per tile, tiles handed out by an atomic counter so that every predecessor is already running),
// formats each row ONCE into a shared-memory staging buffer at its in-tile offset, and copies the buffer out with
// coalesced 16-byte stores.  HBM-bound in principle: 24 B of coordinates in, ~27 B of text out per point.
#include <cub/block/block_scan.cuh>

#include "r3d_common.cuh"
#include "r3d_repr.cuh"

namespace r3d {

constexpr int K6_THREADS = 256;
constexpr int K6_STAGE_BYTES = 24 * 1024;   // a tile whose text is longer (astronomic coordinates) is written row by row

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

// "%.4f" of a value whose 1e-4 units fit 32 bits (|x| < 429 496 -- every coordinate of a metric map): digits with 32-bit
// arithmetic.  q = |x| in 1e-4 units, correctly rounded (fixed4_decompose).  Writes backwards from `end`.
__device__ __forceinline__ int fixed4_len_u32(uint32_t q, int neg) {
    const uint32_t ip = q / 10000u;
    return neg + 5 + (ip >= 10000u ? (ip >= 100000u ? 6 : 5) : (ip >= 100u ? (ip >= 1000u ? 4 : 3) : (ip >= 10u ? 2 : 1)));
}
__device__ __forceinline__ void fixed4_write_u32(uint32_t q, int neg, char* end) {
    uint32_t ip = q / 10000u, fr = q - ip * 10000u;
    char* p = end;
#pragma unroll
    for (int i = 0; i < 4; ++i) { const uint32_t t = fr / 10u; *--p = (char)('0' + (fr - t * 10u)); fr = t; }
    *--p = '.';
    do { const uint32_t t = ip / 10u; *--p = (char)('0' + (ip - t * 10u)); ip = t; } while (ip);
    if (neg) *--p = '-';
}

struct Num4 {       // one coordinate, measured
    Fixed4 f;
    int len;
    bool fast;
};
__device__ __forceinline__ Num4 num4_measure(double v) {
    Num4 r;
    r.f = fixed4_decompose(v);
    r.fast = r.f.kind == 0 && r.f.q < 4294960000ull;
    r.len = r.fast ? fixed4_len_u32((uint32_t)r.f.q, r.f.neg) : fixed4_len(v);
    return r;
}
__device__ __forceinline__ char* num4_write(const Num4& m, double v, char* p) {
    if (m.fast) fixed4_write_u32((uint32_t)m.f.q, m.f.neg, p + m.len);
    else fixed4_write(v, p, m.len);
    return p + m.len;
}

// kTxt: "X,Y,Z\n" rows with str(float64) fields (r3d_repr.cuh) instead of PLY rows
template <bool kTxt>
__global__ void __launch_bounds__(K6_THREADS) k6_rows(const TextArgs a) {
    typedef cub::BlockScan<unsigned, K6_THREADS> Scan;
    __shared__ typename Scan::TempStorage tmp;
    __shared__ __align__(16) char stage[K6_STAGE_BYTES + 16];
    __shared__ unsigned s_tile;
    __shared__ unsigned long long s_base;
    const unsigned long long n_tiles = (a.n + K6_THREADS - 1) / K6_THREADS;
    const bool rgb = a.rgb != nullptr;
    const unsigned lane = threadIdx.x & 31u;
    for (;;) {
        if (threadIdx.x == 0) s_tile = atomicAdd(a.tile_counter, 1u);
        __syncthreads();
        const unsigned long long t = s_tile;
        if (t >= n_tiles) break;
        const unsigned long long i = t * K6_THREADS + threadIdx.x;
        double x = 0, y = 0, z = 0;
        unsigned r = 0, g = 0, b = 0, len = 0;
        char row[kTxt ? kTxtRowMax : 4];
        Num4 mx, my, mz;
        if (i < a.n) {
            x = a.x[i * a.stride]; y = a.y[i * a.stride]; z = a.z[i * a.stride];
            if (kTxt) {
                len = (unsigned)txt_row_write(row, x, y, z, a.z_int != 0);
            } else {
                mx = num4_measure(x); my = num4_measure(y); mz = num4_measure(z);
                len = (unsigned)(mx.len + my.len + mz.len) + 4u;
                if (rgb) {
                    r = a.rgb[3 * i]; g = a.rgb[3 * i + 1]; b = a.rgb[3 * i + 2];
                    len += (unsigned)(u8_len(r) + u8_len(g) + u8_len(b)) + 4u;       // "r g b 0" after the third blank
                }
            }
        }
        unsigned off, total;
        Scan(tmp).ExclusiveSum(len, off, total);
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
        if (fits && total <= (unsigned)K6_STAGE_BYTES) {
            // stage at (base % 16) so that the copy below moves aligned 16-byte words
            const unsigned skew = (unsigned)(base & 15ull);
            if (len) {
                char* p = stage + skew + off;
                if (kTxt) { for (unsigned k = 0; k < len; ++k) p[k] = row[k]; }
                else {
                    p = num4_write(mx, x, p); *p++ = ' ';
                    p = num4_write(my, y, p); *p++ = ' ';
                    p = num4_write(mz, z, p); *p++ = ' ';
                    if (rgb) {
                        p = u8_write(p, r); *p++ = ' ';
                        p = u8_write(p, g); *p++ = ' ';
                        p = u8_write(p, b); *p++ = ' ';
                        *p++ = '0';
                    }
                    *p++ = '\n';
                }
            }
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
        } else if (fits && len) {
            if (kTxt) { for (unsigned k = 0; k < len; ++k) a.out[base + off + k] = row[k]; }
            else ply_row_write(a.out + base + off, x, y, z, rgb, r, g, b);
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
    const unsigned long long n_tiles = (n + K6_THREADS - 1) / K6_THREADS;
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
    if (txt) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k6_rows<true>, K6_THREADS, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k6_rows<false>, K6_THREADS, 0);
    if (per_sm < 1) per_sm = 1;
    // persistent CTAs, all resident (the look-back spins on tiles handed out earlier: they must be running)
    unsigned long long grid = (unsigned long long)ctx->sm_count * per_sm;
    if (grid > n_tiles) grid = n_tiles;
    if (txt) k6_rows<true><<<(unsigned)grid, K6_THREADS, 0, ctx->stream>>>(a);
    else k6_rows<false><<<(unsigned)grid, K6_THREADS, 0, ctx->stream>>>(a);
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
