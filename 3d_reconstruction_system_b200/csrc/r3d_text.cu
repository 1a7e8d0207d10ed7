// r3d_text.cu -- K6: the ASCII side of the path on the GPU.
//
// genply / genply_RGB / genply_noRGB (transfer/camera_to_world.py:112-134, transfer/pixel_to_camera.py:55-124) format
// every point with "%.4f" in a Python loop (1.2 s per 1242x375 frame in the reference); here one thread formats one
// row, byte for byte what Python's "%.4f" produces (r3d_math.cuh::fixed4_*: exact integer arithmetic, round half to
// even on the exact binary value).  Rows have different lengths, so the text is produced in two passes over tiles of
// 256 rows: (1) row lengths summed per tile, exclusive scan over the tiles (CUB), (2) rows written into a
// shared-memory staging buffer at their in-tile offsets and copied out with coalesced 16-byte stores.
// HBM-bound in principle: 24 B of coordinates in, ~27 B of text out per point.
#include <cub/block/block_reduce.cuh>
#include <cub/block/block_scan.cuh>
#include <cub/device/device_scan.cuh>

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
    unsigned long long* tile_len;  // [n_tiles] -> exclusive offsets after the scan
    char* out;
};

__device__ __forceinline__ int k6_row_len(const TextArgs& a, unsigned long long i) {
    const bool rgb = a.rgb != nullptr;
    return ply_row_len(a.x[i * a.stride], a.y[i * a.stride], a.z[i * a.stride], rgb, rgb ? a.rgb[3 * i] : 0u, rgb ? a.rgb[3 * i + 1] : 0u,
                       rgb ? a.rgb[3 * i + 2] : 0u);
}

// kTxt: "X,Y,Z\n" rows with str(float64) fields (r3d_repr.cuh) instead of PLY rows
template <bool kTxt>
__global__ void __launch_bounds__(K6_THREADS) k6_count(const TextArgs a) {
    typedef cub::BlockReduce<unsigned, K6_THREADS> Reduce;
    __shared__ typename Reduce::TempStorage tmp;
    const unsigned long long n_tiles = (a.n + K6_THREADS - 1) / K6_THREADS;
    for (unsigned long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const unsigned long long i = t * K6_THREADS + threadIdx.x;
        unsigned len = 0;
        if (i < a.n) {
            if (kTxt) {
                char row[kTxtRowMax];
                len = (unsigned)txt_row_write(row, a.x[i * a.stride], a.y[i * a.stride], a.z[i * a.stride], a.z_int != 0);
            } else {
                len = (unsigned)k6_row_len(a, i);
            }
        }
        const unsigned sum = Reduce(tmp).Sum(len);
        if (threadIdx.x == 0) a.tile_len[t] = sum;
        __syncthreads();
    }
}

template <bool kTxt>
__global__ void __launch_bounds__(K6_THREADS) k6_write(const TextArgs a) {
    typedef cub::BlockScan<unsigned, K6_THREADS> Scan;
    __shared__ typename Scan::TempStorage tmp;
    __shared__ __align__(16) char stage[K6_STAGE_BYTES + 16];
    const unsigned long long n_tiles = (a.n + K6_THREADS - 1) / K6_THREADS;
    const bool rgb = a.rgb != nullptr;
    for (unsigned long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const unsigned long long i = t * K6_THREADS + threadIdx.x;
        double x = 0, y = 0, z = 0;
        unsigned r = 0, g = 0, b = 0, len = 0;
        char row[kTxt ? kTxtRowMax : 4];
        if (i < a.n) {
            x = a.x[i * a.stride]; y = a.y[i * a.stride]; z = a.z[i * a.stride];
            if (kTxt) {
                len = (unsigned)txt_row_write(row, x, y, z, a.z_int != 0);
            } else {
                if (rgb) { r = a.rgb[3 * i]; g = a.rgb[3 * i + 1]; b = a.rgb[3 * i + 2]; }
                len = (unsigned)ply_row_len(x, y, z, rgb, r, g, b);
            }
        }
        unsigned off, total;
        Scan(tmp).ExclusiveSum(len, off, total);
        const unsigned long long base = a.tile_len[t];   // exclusive offset of the tile
        if (total <= (unsigned)K6_STAGE_BYTES) {
            // stage at (base % 16) so that the copy below moves aligned 16-byte words
            const unsigned skew = (unsigned)(base & 15ull);
            if (len) {
                if (kTxt) { for (unsigned k = 0; k < len; ++k) stage[skew + off + k] = row[k]; }
                else ply_row_write(stage + skew + off, x, y, z, rgb, r, g, b);
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
        } else if (len) {
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
    R3D_TRY(scratch_reserve(ctx, SCR_TILE, (size_t)(n_tiles + 1) * 16 + 256));
    unsigned long long* lens = (unsigned long long*)ctx->scratch[SCR_TILE];
    unsigned long long* offs = lens + n_tiles + 1;
    a.tile_len = lens;
    const unsigned grid = (unsigned)(n_tiles < (unsigned long long)ctx->sm_count * 8 ? n_tiles : (unsigned long long)ctx->sm_count * 8);
    R3D_CUDA_OK(ctx, cudaMemsetAsync(lens + n_tiles, 0, 8, ctx->stream));
    if (txt) k6_count<true><<<grid, K6_THREADS, 0, ctx->stream>>>(a);
    else k6_count<false><<<grid, K6_THREADS, 0, ctx->stream>>>(a);
    ctx->launches++;
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, lens, offs, (int)(n_tiles + 1), ctx->stream);
    R3D_TRY(scratch_reserve(ctx, SCR_CUBTMP, tmp_bytes + 256));
    R3D_CUDA_OK(ctx, cub::DeviceScan::ExclusiveSum(ctx->scratch[SCR_CUBTMP], tmp_bytes, lens, offs, (int)(n_tiles + 1), ctx->stream));
    ctx->launches++;
    unsigned long long total = 0;
    R3D_CUDA_OK(ctx, cudaMemcpyAsync(ctx->pinned, offs + n_tiles, 8, cudaMemcpyDeviceToHost, ctx->stream));
    R3D_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(&total, ctx->pinned, 8);
    *len = (size_t)total;
    if (!out || cap < total) return R3D_OK;      // size query (or buffer too small): *len tells what is needed
    const bool out_dev = is_device_ptr(out);
    char* d_out = out;
    if (!out_dev) {
        R3D_TRY(scratch_reserve(ctx, SCR_OUT0, (size_t)total + 64));
        d_out = (char*)ctx->scratch[SCR_OUT0];
    } else if (((uintptr_t)out & 15u) != 0) {
        return set_error(ctx, R3D_ERR_ARG, "device text buffer must be 16-byte aligned");
    }
    a.tile_len = offs;
    a.out = d_out;
    if (txt) k6_write<true><<<grid, K6_THREADS, 0, ctx->stream>>>(a);
    else k6_write<false><<<grid, K6_THREADS, 0, ctx->stream>>>(a);
    ctx->launches++;
    R3D_CUDA_OK(ctx, cudaGetLastError());
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
