// r3d_backproject.cu -- K1: fused depth decode -> pinhole back-projection -> pose transform -> xyz records.
//
// Replaces gentxtcord + get_pointdata/point_camera of the reference (transfer/camera_to_world.py:57-59,
// 67-105; transfer/pixel_to_camera.py:24-44).  HBM-bound: sizeof(depth sample) + 12 B per pixel.
//
// Main kernel (k1_bulk_vec): persistent CTAs, each owning a contiguous run of 2048-pixel tiles of the
// flattened frame batch.  Depth tiles arrive in shared memory through a 2-stage cp.async.bulk (UBLKCP)
// ring signalled by mbarriers; xyz records are assembled in shared memory and leave through
// cp.async.bulk shared->global stores (bulk groups), so the SM's LSU only sees conflict-free LDS/STS.
// All arithmetic is fp64 with separately rounded products/sums (r3d_math.cuh) and is cast once.
#include <cub/device/device_scan.cuh>

#include <type_traits>

#include "r3d_common.cuh"

namespace r3d {

constexpr int K1_THREADS = 256;
constexpr int K1_PPT = 4;                       // pixels per thread per tile
constexpr int K1_TILE = K1_THREADS * K1_PPT;    // 1024 pixels
constexpr int K1_STAGES = 4;

struct K1Args {
    const void* depth;
    void* out;
    const double* rt;              // n_frames x 12 or nullptr (camera frame)
    unsigned long long px_begin;   // first flat pixel handled by this launch (generic kernels)
    unsigned long long px_count;   // pixels handled by this launch
    unsigned long long n_tiles;    // bulk: full tiles
    unsigned W, H, WH, n_frames;
    unsigned long long pitch;      // bytes per row
    double fx, fy, cx, cy, depth_scale, fB;
    int mode;
    int valid_fast;                // validity is "sample > 0 (and finite)": depth_scale cannot under- or overflow a sample
    // compaction
    const unsigned long long* tile_offsets;  // exclusive scan of valid counts per tile
    unsigned long long* tile_counts;
    unsigned long long* frame_counts;
    unsigned long long* frame_ends;          // fast compaction: records written up to and including each frame's last pixel
    // disparity mode, integer samples: Z of every possible sample value (exactly decode_z's), or nullptr.  One cached 8-byte
    // load instead of an fp64 division per pixel (the division made C4's 640x480 disparity frames fp64-pipe bound: 0.58 of
    // the HBM peak against 0.94 for depth frames)
    const double* ztab;
};

// ------------------------------------------------------------------ PTX helpers (mbarrier + bulk async copy)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global, tracked by bulk async-groups
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_all() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ------------------------------------------------------------------ per-pixel arithmetic
template <typename DepthT>
__device__ __forceinline__ double raw_to_double(DepthT v) { return (double)v; }

// Validity of a sample (decode_z's rule: d = raw * depth_scale, d > 0 and finite).  With depth_scale in [1e-250, 1e250]
// no sample of these types can under- or overflow the product, and the rule is a compare on the sample itself.
template <typename DepthT>
__device__ __forceinline__ bool sample_valid(const K1Args& a, DepthT raw, int mode) {
    if (a.valid_fast) {
        if constexpr (sizeof(DepthT) == 4) return raw > 0.0f && raw <= 3.402823466e+38f;
        else return raw != 0;
    }
    bool valid;
    decode_z(raw_to_double(raw), mode, a.depth_scale, a.fB, valid);
    return valid;
}

template <typename OutT>
__device__ __forceinline__ OutT out_cast(double v);
template <>
__device__ __forceinline__ float out_cast<float>(double v) { return __double2float_rn(v); }
template <>
__device__ __forceinline__ double out_cast<double>(double v) { return v; }
// world coordinates: float64 records get the reference's +0.0 for sums of signed zeros (pose_canon)
template <typename OutT>
__device__ __forceinline__ OutT world_cast(double v) {
    if constexpr (sizeof(OutT) == 8) return pose_canon(v);
    else return __double2float_rn(v);
}

// One pixel: raw sample + table entries -> record.  `pose_frame` caches which frame `pose` holds.
template <typename OutT, bool kWorld, int kMode = -1>
__device__ __forceinline__ bool k1_pixel(const K1Args& a, double raw, double au, double bv, unsigned frame,
                                         unsigned& pose_frame, Pose& pose, OutT& ox, OutT& oy, OutT& oz) {
    bool valid;
    const double Z = decode_z(raw, kMode < 0 ? a.mode : kMode, a.depth_scale, a.fB, valid);
    const double X = dmul(au, Z);
    const double Y = dmul(bv, Z);
    if (kWorld) {
        if (frame != pose_frame) {
            pose_load(a.rt + (size_t)frame * 12, pose);
            pose_frame = frame;
        }
        double wx, wy, wz;
        pose_apply(pose, X, Y, Z, wx, wy, wz);
        ox = world_cast<OutT>(wx); oy = world_cast<OutT>(wy); oz = world_cast<OutT>(wz);
    } else {
        ox = out_cast<OutT>(X); oy = out_cast<OutT>(Y); oz = out_cast<OutT>(Z);
    }
    return valid;
}

// Same with the pose already in registers (hot kernel: the reload sits on the rare frame-wrap path).
template <typename OutT, bool kWorld, int kMode>
__device__ __forceinline__ void k1_pixel_pose(const K1Args& a, double raw, double au, double bv, const Pose& pose, OutT& ox,
                                              OutT& oy, OutT& oz) {
    bool valid;
    double Z;
    if (kMode == 1 && a.ztab != nullptr) Z = __ldg(a.ztab + __double2int_rn(raw));
    else Z = decode_z(raw, kMode, a.depth_scale, a.fB, valid);
    const double X = dmul(au, Z);
    const double Y = dmul(bv, Z);
    if (kWorld) {
        double wx, wy, wz;
        pose_apply(pose, X, Y, Z, wx, wy, wz);
        ox = world_cast<OutT>(wx); oy = world_cast<OutT>(wy); oz = world_cast<OutT>(wz);
    } else {
        ox = out_cast<OutT>(X); oy = out_cast<OutT>(Y); oz = out_cast<OutT>(Z);
    }
}

__device__ __forceinline__ void k1_tables(const K1Args& a, double* col, double* row) {
    for (unsigned i = threadIdx.x; i < a.W; i += blockDim.x) col[i] = pixel_coeff((int)i, a.cx, a.fx);
    for (unsigned j = threadIdx.x; j < a.H; j += blockDim.x) row[j] = pixel_coeff((int)j, a.cy, a.fy);
}

// ------------------------------------------------------------------ k1_bulk_vec: the hot kernel
// Requires W >= K1_THREADS and H >= 8 (smaller images take the generic kernel) so that stepping a pixel index by
// 1024 wraps the column at most (1 + 1024/W) times and the row at most once.
//
// A thread owns 4 CONSECUTIVE pixels of each 1024-pixel group: one vector load of the samples, the column
// coefficients as two 16-byte loads, one row coefficient, and three 16-byte stores of the 12 output floats (48-byte
// thread stride: conflict-free per quarter warp).  A tile is kGroups groups, so the per-tile work (mbarrier wait,
// proxy fence, CTA barrier, bulk issue) is paid once per 4 * kGroups pixels per thread.  The first version of this
// kernel (one pixel per thread per 256-pixel row of the tile, 4-stage ring, 4 CTAs per SM) issued 67 warp instructions
// per pixel, 28 of them the fp64 arithmetic the parity contract fixes, and ran at 70 % issue-slot utilisation and
// 0.83 of the HBM copy rate; this layout issues about 47 and reaches 0.94.  A warp in which some thread's 4 pixels
// cross a row end (about one warp-group in ten at W = 1242) or need another frame's pose takes the per-pixel path.
//
// Measured on B200 (C2, 4 500 frames per launch; tools/k1_probe.py, profiles/r1_k1_vec_sweep.jsonl), fraction of the
// measured HBM copy rate: groups x stages x CTAs/SM = 2x2x3 0.94 (default) | 3x2x2 0.94 | 2x4x2 0.93 | 1x4x3 0.91 |
// 2x3x3 0.88 | 1x2x4 0.86 | 1x2x3 0.85 | 2x1x3 0.71; tiles interleaved across CTAs instead of one contiguous run per
// CTA 0.81; three output buffers 0.85.
#ifndef K1V_GROUPS
#define K1V_GROUPS 2
#endif
#ifndef K1V_STAGES
#define K1V_STAGES 2
#endif
#ifndef K1V_MINB
#define K1V_MINB 3
#endif
constexpr int K1V_GROUP = K1_THREADS * 4;       // 1024 pixels

template <typename DepthT> struct SampleVec;
template <> struct SampleVec<unsigned char> { using type = unsigned; };
template <> struct SampleVec<unsigned short> { using type = uint2; };
template <> struct SampleVec<float> { using type = uint4; };

__device__ __forceinline__ void store_records4(float* dst, const float (&o)[12]) {
    float4* d = reinterpret_cast<float4*>(dst);
    d[0] = make_float4(o[0], o[1], o[2], o[3]);
    d[1] = make_float4(o[4], o[5], o[6], o[7]);
    d[2] = make_float4(o[8], o[9], o[10], o[11]);
}
__device__ __forceinline__ void store_records4(double*, const double (&)[3]) {}

// The 4 consecutive pixels of one thread: (u0, v0, f0) is the first one.  emit(j, x, y, z) receives the records in
// pixel order.  Returns the index of the pixel that is the last one of its frame (f0's), or -1.
template <typename DepthT, typename OutT, bool kWorld, int kMode, typename Emit>
__device__ __forceinline__ int k1_group4(const K1Args& a, const double* col, const double* row, const DepthT (&raw)[4], unsigned u0,
                                         unsigned v0, unsigned f0, unsigned& pose_frame, Pose& pose, Emit&& emit) {
    const unsigned W = a.W, H = a.H;
    int last = -1;
    if (__all_sync(0xffffffffu, (u0 + 3u < W) & (!kWorld | (f0 == pose_frame)))) {
        // no thread of the warp crosses a row end or needs another pose: one row coefficient, column coefficients by
        // pairs.  (The pose reload sits on the other path because the compiler predicates its 12 loads instead of
        // branching around them: 12 issue slots per group when it is in line.)
        const double bv = row[v0];
        double au[4];
        if ((u0 & 1u) == 0) {
            const double2 c01 = *reinterpret_cast<const double2*>(col + u0);
            const double2 c23 = *reinterpret_cast<const double2*>(col + u0 + 2);
            au[0] = c01.x; au[1] = c01.y; au[2] = c23.x; au[3] = c23.y;
        } else {
            au[0] = col[u0]; au[1] = col[u0 + 1]; au[2] = col[u0 + 2]; au[3] = col[u0 + 3];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            OutT x, y, z;
            k1_pixel_pose<OutT, kWorld, kMode>(a, raw_to_double(raw[j]), au[j], bv, pose, x, y, z);
            emit(j, x, y, z);
        }
        if ((u0 + 4u == W) & (v0 + 1u == H)) last = 3;
    } else {
        if (kWorld && f0 != pose_frame) { pose_frame = f0; pose_load(a.rt + (size_t)f0 * 12, pose); }
        unsigned u = u0, v = v0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            OutT x, y, z;
            k1_pixel_pose<OutT, kWorld, kMode>(a, raw_to_double(raw[j]), col[u], row[v], pose, x, y, z);
            emit(j, x, y, z);
            if (++u == W) {
                u = 0;
                if (++v == H) {                // next frame (rare): switch pose
                    v = 0;
                    last = j;
                    if (kWorld && ++pose_frame < a.n_frames) pose_load(a.rt + (size_t)pose_frame * 12, pose);
                }
            }
        }
    }
    return last;
}

template <typename DepthT>
__device__ __forceinline__ void load_samples4(const DepthT* p, DepthT (&raw)[4]) {
    *reinterpret_cast<typename SampleVec<DepthT>::type*>(raw) = *reinterpret_cast<const typename SampleVec<DepthT>::type*>(p);
}

// Shared-memory carve-up of the two bulk kernels: [mbarriers + small arrays | col | row | input ring | output buffers]
__device__ __forceinline__ size_t k1v_tables(const K1Args& a, unsigned char* smem, size_t head, double*& col, double*& row) {
    col = reinterpret_cast<double*>(smem + head);
    row = col + ((a.W + 1u) & ~1u);                                     // 16-byte aligned
    return head + ((size_t)(a.W + 1 + a.H) * 8 + 127) / 128 * 128;
}

template <typename DepthT, typename OutT, bool kWorld, int kMode, int kGroups>
__global__ void __launch_bounds__(K1_THREADS, K1V_MINB) k1_bulk_vec(const K1Args a) {
    constexpr int kStages = K1V_STAGES, kOutBufs = 2;
    constexpr int kTile = K1V_GROUP * kGroups;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);                 // kStages mbarriers
    double *col, *row;
    size_t off = k1v_tables(a, smem, 128, col, row);
    DepthT* in_s = reinterpret_cast<DepthT*>(smem + off);               // kStages x kTile
    off += (size_t)kStages * kTile * sizeof(DepthT);
    OutT* out_s = reinterpret_cast<OutT*>(smem + off);                  // kOutBufs x kTile x 3

    const unsigned tid = threadIdx.x;
    constexpr uint32_t kInBytes = kTile * sizeof(DepthT);
    constexpr uint32_t kOutBytes = kTile * 3 * sizeof(OutT);

    // contiguous run of tiles for this CTA (tiles interleaved across CTAs measured 0.81 against 0.94 of the copy rate)
    const unsigned long long per = (a.n_tiles + gridDim.x - 1) / gridDim.x;
    const unsigned long long t0 = (unsigned long long)blockIdx.x * per;
    unsigned long long t1 = t0 + per;
    if (t1 > a.n_tiles) t1 = a.n_tiles;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    k1_tables(a, col, row);
    __syncthreads();
    if (t0 >= t1) return;

    const DepthT* gin = reinterpret_cast<const DepthT*>(a.depth);
    OutT* gout = reinterpret_cast<OutT*>(a.out);
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            if (t0 + s < t1) {
                mbar_arrive_expect_tx(&full[s], kInBytes);
                bulk_load(in_s + (size_t)s * kTile, gin + (t0 + s) * kTile, kInBytes, &full[s]);
            }
        }
    }

    // (frame, row, column) of the first of this thread's 4 pixels in the current group, advanced by 1024 per group
    const unsigned long long px0 = t0 * kTile + tid * 4u;
    unsigned f0 = (unsigned)(px0 / a.WH);
    const unsigned r0 = (unsigned)(px0 - (unsigned long long)f0 * a.WH);
    unsigned v0 = r0 / a.W, u0 = r0 - v0 * a.W;
    const unsigned W = a.W, H = a.H;
    const unsigned q_grp = K1V_GROUP / W, r_grp = K1V_GROUP - q_grp * W;
    unsigned pose_frame = 0xffffffffu;
    Pose pose;

    unsigned stage = 0, parity = 0, ob = 0;
    for (unsigned long long t = t0; t < t1; ++t) {
        mbar_wait(&full[stage], parity);
        const DepthT* tin = in_s + (size_t)stage * kTile + tid * 4u;
        OutT* tout = out_s + (size_t)ob * kTile * 3 + tid * 12u;
#pragma unroll 1
        for (int g = 0; g < kGroups; ++g) {
            DepthT raw[4];
            load_samples4(tin + g * K1V_GROUP, raw);
            OutT* const gdst = tout + g * K1V_GROUP * 3;
            OutT o[sizeof(OutT) == 4 ? 12 : 3];                         // float: 12 values leave as three 16-byte stores
            k1_group4<DepthT, OutT, kWorld, kMode>(a, col, row, raw, u0, v0, f0, pose_frame, pose, [&](int j, OutT x, OutT y, OutT z) {
                if constexpr (sizeof(OutT) == 4) { o[3 * j] = x; o[3 * j + 1] = y; o[3 * j + 2] = z; }
                else { gdst[3 * j] = x; gdst[3 * j + 1] = y; gdst[3 * j + 2] = z; }   // double: pixel by pixel (registers)
            });
            if constexpr (sizeof(OutT) == 4) store_records4(gdst, o);
            u0 += r_grp; v0 += q_grp;
            if (u0 >= W) { u0 -= W; ++v0; }
            if (v0 >= H) { v0 -= H; ++f0; }
        }
        // the bulk store issued kOutBufs-1 tiles ago must have finished reading the buffer the NEXT tile writes
        if (tid == 0) bulk_wait_read<kOutBufs - 2>();
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            bulk_store(gout + t * (unsigned long long)kTile * 3, out_s + (size_t)ob * kTile * 3, kOutBytes);
            bulk_commit();
            const unsigned long long tn = t + kStages;
            if (tn < t1) {   // every thread is past its reads of this stage (barrier above): refill it
                mbar_arrive_expect_tx(&full[stage], kInBytes);
                bulk_load(in_s + (size_t)stage * kTile, gin + tn * kTile, kInBytes, &full[stage]);
            }
        }
        if (++stage == kStages) { stage = 0; parity ^= 1u; }
        if (++ob == kOutBufs) ob = 0;
    }
    if (tid == 0) bulk_wait_all<0>();
}

// ------------------------------------------------------------------ generic kernels (tails, pitched / unaligned input)
template <typename DepthT>
__device__ __forceinline__ double load_raw(const K1Args& a, unsigned f, unsigned v, unsigned u) {
    const unsigned char* p = reinterpret_cast<const unsigned char*>(a.depth) +
                             ((size_t)f * a.H + v) * a.pitch + (size_t)u * sizeof(DepthT);
    return raw_to_double(*reinterpret_cast<const DepthT*>(p));
}

template <typename DepthT, typename OutT, bool kWorld>
__global__ void __launch_bounds__(K1_THREADS) k1_generic(const K1Args a) {
    extern __shared__ __align__(128) unsigned char smem[];
    double* col = reinterpret_cast<double*>(smem);
    double* row = col + a.W;
    k1_tables(a, col, row);
    __syncthreads();
    OutT* gout = reinterpret_cast<OutT*>(a.out);
    unsigned pose_frame = 0xffffffffu;
    Pose pose;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.px_count; i += stride) {
        const unsigned long long p = a.px_begin + i;
        const unsigned f = (unsigned)(p / a.WH);
        const unsigned r = (unsigned)(p - (unsigned long long)f * a.WH);
        const unsigned v = r / a.W, u = r - v * a.W;
        OutT x, y, z;
        k1_pixel<OutT, kWorld>(a, load_raw<DepthT>(a, f, v, u), col[u], row[v], f, pose_frame, pose, x, y, z);
        gout[p * 3 + 0] = x; gout[p * 3 + 1] = y; gout[p * 3 + 2] = z;
    }
}

// compaction pass 1: valid pixels per 1024-pixel tile and per frame
template <typename DepthT>
__global__ void __launch_bounds__(K1_THREADS) k1_count(const K1Args a) {
    __shared__ unsigned warp_cnt[K1_THREADS / 32];
    const unsigned long long n_tiles = (a.px_count + K1_TILE - 1) / K1_TILE;
    for (unsigned long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        unsigned mine = 0;
#pragma unroll
        for (int j = 0; j < K1_PPT; ++j) {
            const unsigned long long p = t * K1_TILE + threadIdx.x + j * K1_THREADS;
            bool valid = false;
            unsigned f = 0;
            if (p < a.px_count) {
                f = (unsigned)(p / a.WH);
                const unsigned r = (unsigned)(p - (unsigned long long)f * a.WH);
                const unsigned v = r / a.W, u = r - v * a.W;
                decode_z(load_raw<DepthT>(a, f, v, u), a.mode, a.depth_scale, a.fB, valid);
            }
            mine += valid ? 1u : 0u;
            // per-frame counts, one atomic per (warp, frame)
            const unsigned key = valid ? f : 0xffffffffu;
            const unsigned peers = __match_any_sync(0xffffffffu, key);
            if (valid && (threadIdx.x & 31u) == (unsigned)(__ffs(peers) - 1))
                atomicAdd(&a.frame_counts[f], (unsigned long long)__popc(peers));
        }
        for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
        if ((threadIdx.x & 31u) == 0) warp_cnt[threadIdx.x >> 5] = mine;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned s = 0;
            for (int w = 0; w < K1_THREADS / 32; ++w) s += warp_cnt[w];
            a.tile_counts[t] = s;
        }
        __syncthreads();
    }
}

// fast path, pass 1 for the partial last tile: same count, no per-frame atomics (frame counts come from frame_ends)
template <typename DepthT>
__global__ void __launch_bounds__(K1_THREADS) k1_count_tail(const K1Args a) {
    __shared__ unsigned warp_cnt[K1_THREADS / 32];
    const unsigned long long n_tiles = (a.px_count + K1_TILE - 1) / K1_TILE;
    for (unsigned long long t = a.n_tiles; t < n_tiles; ++t) {
        unsigned mine = 0;
        for (int j = 0; j < K1_PPT; ++j) {
            const unsigned long long p = t * K1_TILE + threadIdx.x + j * K1_THREADS;
            bool valid = false;
            if (p < a.px_count) {
                const unsigned f = (unsigned)(p / a.WH);
                const unsigned r = (unsigned)(p - (unsigned long long)f * a.WH);
                const unsigned v = r / a.W, u = r - v * a.W;
                decode_z(load_raw<DepthT>(a, f, v, u), a.mode, a.depth_scale, a.fB, valid);
            }
            mine += valid ? 1u : 0u;
        }
        for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
        if ((threadIdx.x & 31u) == 0) warp_cnt[threadIdx.x >> 5] = mine;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned sum = 0;
            for (int w = 0; w < K1_THREADS / 32; ++w) sum += warp_cnt[w];
            a.tile_counts[t] = sum;
        }
        __syncthreads();
    }
}

// compaction pass 2: ordered write of the valid records
template <typename DepthT, typename OutT, bool kWorld>
__global__ void __launch_bounds__(K1_THREADS) k1_compact(const K1Args a) {
    extern __shared__ __align__(128) unsigned char smem[];
    double* col = reinterpret_cast<double*>(smem);
    double* row = col + a.W;
    __shared__ unsigned warp_cnt[K1_PPT][K1_THREADS / 32];
    k1_tables(a, col, row);
    __syncthreads();
    OutT* gout = reinterpret_cast<OutT*>(a.out);
    unsigned pose_frame = 0xffffffffu;
    Pose pose;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned long long n_tiles = (a.px_count + K1_TILE - 1) / K1_TILE;
    // (fast path: a.n_tiles full tiles were written by k1_bulk_compact; this kernel then only handles the partial last tile)
    for (unsigned long long t = (a.frame_ends ? a.n_tiles : 0ull) + blockIdx.x; t < n_tiles; t += gridDim.x) {
        OutT x[K1_PPT], y[K1_PPT], z[K1_PPT];
        unsigned rank[K1_PPT];
        bool ok[K1_PPT];
#pragma unroll
        for (int j = 0; j < K1_PPT; ++j) {
            const unsigned long long p = t * K1_TILE + threadIdx.x + j * K1_THREADS;
            ok[j] = false;
            if (p < a.px_count) {
                const unsigned f = (unsigned)(p / a.WH);
                const unsigned r = (unsigned)(p - (unsigned long long)f * a.WH);
                const unsigned v = r / a.W, u = r - v * a.W;
                ok[j] = k1_pixel<OutT, kWorld>(a, load_raw<DepthT>(a, f, v, u), col[u], row[v], f, pose_frame, pose,
                                               x[j], y[j], z[j]);
            }
            const unsigned b = __ballot_sync(0xffffffffu, ok[j]);
            rank[j] = __popc(b & ((1u << lane) - 1u));
            if (lane == 0) warp_cnt[j][warp] = __popc(b);
        }
        __syncthreads();
        const unsigned long long base = a.tile_offsets[t];
#pragma unroll
        for (int j = 0; j < K1_PPT; ++j) {
            unsigned before = 0;
            for (int jj = 0; jj < K1_PPT; ++jj)
                for (int w = 0; w < K1_THREADS / 32; ++w)
                    if (jj < j || (jj == j && w < (int)warp)) before += warp_cnt[jj][w];
            if (ok[j]) {
                const unsigned long long q = base + before + rank[j];
                gout[q * 3 + 0] = x[j]; gout[q * 3 + 1] = y[j]; gout[q * 3 + 2] = z[j];
            }
            if (a.frame_ends) {
                const unsigned long long p = t * K1_TILE + threadIdx.x + j * K1_THREADS;
                if (p < a.px_count && (p + 1) % a.WH == 0) a.frame_ends[p / a.WH] = base + before + rank[j] + (ok[j] ? 1u : 0u);
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ fast compaction (bulk-eligible input)
// pass 1: valid pixels per full 1024-pixel tile.  One warp per tile, 16-byte loads, no block-level reduction.
template <typename DepthT>
__global__ void __launch_bounds__(K1_THREADS) k1_count_tiles(const K1Args a) {
    constexpr int kPer = 16 / (int)sizeof(DepthT);                      // samples per 16-byte load
    constexpr int kLoads = K1_TILE / (32 * kPer);                       // loads per lane per tile
    const DepthT* gin = reinterpret_cast<const DepthT*>(a.depth);
    const unsigned lane = threadIdx.x & 31u;
    const unsigned long long warps = (unsigned long long)gridDim.x * (K1_THREADS / 32);
    for (unsigned long long t = (unsigned long long)blockIdx.x * (K1_THREADS / 32) + (threadIdx.x >> 5); t < a.n_tiles; t += warps) {
        const uint4* p = reinterpret_cast<const uint4*>(gin + t * K1_TILE) + lane;
        uint4 v[kLoads];
#pragma unroll
        for (int k = 0; k < kLoads; ++k) v[k] = __ldcs(p + k * 32);
        unsigned mine = 0;
#pragma unroll
        for (int k = 0; k < kLoads; ++k) {
            const DepthT* e = reinterpret_cast<const DepthT*>(&v[k]);
#pragma unroll
            for (int j = 0; j < kPer; ++j) mine += sample_valid(a, e[j], a.mode) ? 1u : 0u;
        }
        mine = __reduce_add_sync(0xffffffffu, mine);
        if (lane == 0) a.tile_counts[t] = mine;
    }
}

// 12 consecutive floats of one lane to an address that is S (1..3) words past a 16-byte boundary, when the lanes of the
// warp hold consecutive 12-word blocks: every 16-byte group that straddles two lanes is assembled with one shuffle per
// word and written by the lower lane; lane 0 writes its first 4 - S words and lane 31 its last S words as scalars.
template <int S>
__device__ __forceinline__ void store_records4_shifted(float* dst, const float (&o)[12], unsigned lane) {
    float nb[4 - S];
#pragma unroll
    for (int k = 0; k < 4 - S; ++k) nb[k] = __shfl_down_sync(0xffffffffu, o[k], 1);
    float4* g = reinterpret_cast<float4*>(dst - S);
    g[1] = make_float4(o[4 - S], o[5 - S], o[6 - S], o[7 - S]);
    g[2] = make_float4(o[8 - S], o[9 - S], o[10 - S], o[11 - S]);
    if (lane != 31u) {
        float w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) w[k] = k < S ? o[12 - S + k] : nb[k - S < 0 ? 0 : k - S];
        g[3] = make_float4(w[0], w[1], w[2], w[3]);
    } else {
#pragma unroll
        for (int k = 0; k < S; ++k) dst[12 - S + k] = o[12 - S + k];
    }
    if (lane == 0u) {
#pragma unroll
        for (int k = 0; k < 4 - S; ++k) dst[k] = o[k];
    }
}

// pass 2: k1_bulk_vec with in-tile compaction.  A tile is kGroups of the 1024-pixel tiles pass 1 counted, so the
// output offset of the tile comes from the scan of those counts.  Phase A: every thread decodes the validity of its 4
// consecutive pixels of each group, a warp scan + a 32-entry scan over the (group, warp) totals give every thread the
// position of its first record.  Phase B: the records are computed as in k1_bulk_vec and packed in shared memory in
// pixel order at the 16-byte phase of their destination; a warp whose 128 pixels are all valid (the common case) writes
// its 384 words with 16-byte stores whatever their alignment (store_records4_shifted), other warps store word by word.
// The 16-byte aligned middle of the tile leaves through one cp.async.bulk store, the (<= 3 float) head and tail
// through scalar stores.
template <typename DepthT, typename OutT, bool kWorld, int kMode, int kGroups>
__global__ void __launch_bounds__(K1_THREADS, K1V_MINB) k1_bulk_compact(const K1Args a) {
    constexpr int kStages = K1V_STAGES, kOutBufs = 2;
    constexpr int kTile = K1V_GROUP * kGroups;
    constexpr int kPerWord = 16 / (int)sizeof(OutT);                    // output elements per 16 bytes
    static_assert(kGroups * (K1_THREADS / 32) <= 32, "one warp scans the (group, warp) totals");
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    unsigned* part = reinterpret_cast<unsigned*>(smem + 64);            // [kGroups][8] valid pixels per (group, warp)
    double *col, *row;
    size_t off = k1v_tables(a, smem, 256, col, row);
    DepthT* in_s = reinterpret_cast<DepthT*>(smem + off);
    off += (size_t)kStages * kTile * sizeof(DepthT);
    OutT* out_s = reinterpret_cast<OutT*>(smem + off);                  // kOutBufs x (kTile * 3 + kPerWord)
    constexpr size_t kOutStride = (size_t)kTile * 3 + kPerWord;

    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    constexpr uint32_t kInBytes = kTile * sizeof(DepthT);
    const unsigned long long per = (a.n_tiles + gridDim.x - 1) / gridDim.x;
    const unsigned long long t0 = (unsigned long long)blockIdx.x * per;
    unsigned long long t1 = t0 + per;
    if (t1 > a.n_tiles) t1 = a.n_tiles;
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    if (tid < 32) part[tid] = 0;
    k1_tables(a, col, row);
    __syncthreads();
    if (t0 >= t1) return;
    const DepthT* gin = reinterpret_cast<const DepthT*>(a.depth);
    OutT* gout = reinterpret_cast<OutT*>(a.out);
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            if (t0 + s < t1) {
                mbar_arrive_expect_tx(&full[s], kInBytes);
                bulk_load(in_s + (size_t)s * kTile, gin + (t0 + s) * kTile, kInBytes, &full[s]);
            }
        }
    }
    const unsigned long long px0 = t0 * kTile + tid * 4u;
    unsigned f0 = (unsigned)(px0 / a.WH);
    const unsigned r0 = (unsigned)(px0 - (unsigned long long)f0 * a.WH);
    unsigned v0 = r0 / a.W, u0 = r0 - v0 * a.W;
    const unsigned W = a.W, H = a.H;
    const unsigned q_grp = K1V_GROUP / W, r_grp = K1V_GROUP - q_grp * W;
    unsigned pose_frame = 0xffffffffu;
    Pose pose;
    unsigned stage = 0, parity = 0, ob = 0;
    unsigned nxt_valid = 0;
#pragma unroll
    for (int g = 0; g < kGroups; ++g) nxt_valid += (unsigned)a.tile_counts[t0 * kGroups + g];
    unsigned long long nxt_base = a.tile_offsets[t0 * kGroups];
    for (unsigned long long t = t0; t < t1; ++t) {
        mbar_wait(&full[stage], parity);
        const DepthT* tin = in_s + (size_t)stage * kTile + tid * 4u;
        // what pass 1 counted in this tile (read one tile ahead: the loads have a whole tile to arrive): all-valid and
        // all-invalid tiles skip the position bookkeeping
        const unsigned tile_valid = nxt_valid;
        const unsigned long long base = nxt_base;                       // records before this tile
        if (t + 1 < t1) {
            nxt_valid = 0;
#pragma unroll
            for (int g = 0; g < kGroups; ++g) nxt_valid += (unsigned)a.tile_counts[(t + 1) * kGroups + g];
            nxt_base = a.tile_offsets[(t + 1) * kGroups];
        }
        const unsigned skew = (unsigned)((base * 3ull) % (unsigned)kPerWord);
        OutT* buf = out_s + (size_t)ob * kOutStride + skew;
        unsigned long long meta = 0;
        unsigned excl = 0, total = tile_valid;
        const bool mixed = tile_valid != 0u && tile_valid != (unsigned)kTile;
        if (mixed) {
            // ---- phase A: validity nibble and warp-exclusive record count of this thread, per group (16 bits each)
#pragma unroll
            for (int g = 0; g < kGroups; ++g) {
                DepthT raw[4];
                load_samples4(tin + g * K1V_GROUP, raw);
                unsigned m = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) m |= sample_valid(a, raw[j], kMode) ? (1u << j) : 0u;
                const unsigned c = __popc(m);
                unsigned incl = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const unsigned n = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (unsigned)o) incl += n; }
                if (lane == 31u) part[g * 8 + warp] = incl;
                meta |= (unsigned long long)(m | ((incl - c) << 4)) << (16 * g);
            }
            __syncthreads();                                            // part[] complete
            // exclusive scan of the (group, warp) totals, redundantly in every warp
            const unsigned c = part[lane];
            unsigned incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned n = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (unsigned)o) incl += n; }
            excl = incl - c;
        }
        // ---- phase B: records, packed
        bool packed = false;
        if constexpr (sizeof(OutT) == 4) {
            if (tile_valid == (unsigned)kTile) {
                // every pixel of the tile is valid (the common case): the k1_bulk_vec loop with the records shifted to
                // the 16-byte phase of their destination -- the phase is the same for every group of the tile, so the
                // store variant is chosen once per tile
                auto run = [&](auto phase) {
                    constexpr int kS = decltype(phase)::value;
#pragma unroll 1
                    for (int g = 0; g < kGroups; ++g) {
                        DepthT raw[4];
                        load_samples4(tin + g * K1V_GROUP, raw);
                        const unsigned pos0 = (unsigned)g * K1V_GROUP + tid * 4u;
                        OutT* const dst = buf + pos0 * 3u;
                        OutT o[12];
                        const int last = k1_group4<DepthT, OutT, kWorld, kMode>(a, col, row, raw, u0, v0, f0, pose_frame, pose, [&](int j, OutT x, OutT y, OutT z) {
                            o[3 * j] = x; o[3 * j + 1] = y; o[3 * j + 2] = z;
                        });
                        if constexpr (kS == 0) store_records4(dst, o);
                        else store_records4_shifted<kS>(dst, o, lane);
                        if (last >= 0) a.frame_ends[f0] = base + pos0 + (unsigned)last + 1u;
                        u0 += r_grp; v0 += q_grp;
                        if (u0 >= W) { u0 -= W; ++v0; }
                        if (v0 >= H) { v0 -= H; ++f0; }
                    }
                };
                switch (skew & 3u) {
                    case 0: run(std::integral_constant<int, 0>()); break;
                    case 1: run(std::integral_constant<int, 1>()); break;
                    case 2: run(std::integral_constant<int, 2>()); break;
                    default: run(std::integral_constant<int, 3>()); break;
                }
                packed = true;
            }
        }
        if (!packed && tile_valid != 0u) {
#pragma unroll 1
            for (int g = 0; g < kGroups; ++g) {
                unsigned m = 15u, pos0 = (unsigned)g * K1V_GROUP + tid * 4u;   // all-valid tile: every pixel keeps its place
                if (mixed) {
                    const unsigned mg = (unsigned)(meta >> (16 * g)) & 0xffffu;
                    m = mg & 15u;
                    pos0 = __shfl_sync(0xffffffffu, excl, g * 8 + (int)warp) + (mg >> 4);   // first record of this thread
                }
                if (mixed && __all_sync(0xffffffffu, m == 0u)) {
                    // 128 invalid pixels in a row (sky): no record to compute, only a frame end to report
                    if (v0 + 1u == H && W - u0 <= 4u) a.frame_ends[f0] = base + pos0;
                    u0 += r_grp; v0 += q_grp;
                    if (u0 >= W) { u0 -= W; ++v0; }
                    if (v0 >= H) { v0 -= H; ++f0; }
                    continue;
                }
                DepthT raw[4];
                load_samples4(tin + g * K1V_GROUP, raw);
                OutT* const dst = buf + pos0 * 3u;
                OutT o[sizeof(OutT) == 4 ? 12 : 3];
                unsigned rank = 0;
                const int last = k1_group4<DepthT, OutT, kWorld, kMode>(a, col, row, raw, u0, v0, f0, pose_frame, pose, [&](int j, OutT x, OutT y, OutT z) {
                    if constexpr (sizeof(OutT) == 4) { o[3 * j] = x; o[3 * j + 1] = y; o[3 * j + 2] = z; }
                    else {                                              // double: word by word
                        if ((m >> j) & 1u) { dst[3 * rank] = x; dst[3 * rank + 1] = y; dst[3 * rank + 2] = z; ++rank; }
                    }
                });
                if constexpr (sizeof(OutT) == 4) {
                    if (__all_sync(0xffffffffu, m == 15u)) {
                        // the warp's 384 words are consecutive: 16-byte stores at any alignment
                        switch ((unsigned)((uintptr_t)dst >> 2) & 3u) {
                            case 0: store_records4(dst, o); break;
                            case 1: store_records4_shifted<1>(dst, o, lane); break;
                            case 2: store_records4_shifted<2>(dst, o, lane); break;
                            default: store_records4_shifted<3>(dst, o, lane); break;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if ((m >> j) & 1u) { dst[3 * rank] = o[3 * j]; dst[3 * rank + 1] = o[3 * j + 1]; dst[3 * rank + 2] = o[3 * j + 2]; ++rank; }
                    }
                }
                if (last >= 0) a.frame_ends[f0] = base + pos0 + __popc(m & ((2u << last) - 1u));
                u0 += r_grp; v0 += q_grp;
                if (u0 >= W) { u0 -= W; ++v0; }
                if (v0 >= H) { v0 -= H; ++f0; }
            }
        } else if (!packed) {
            // no valid pixel in the tile: nothing to compute; frames that end inside it still report their record count
#pragma unroll 1
            for (int g = 0; g < kGroups; ++g) {
                const unsigned to_row_end = W - u0;                     // pixels from this thread's first one to the end of its row
                if (v0 + 1u == H && to_row_end <= 4u) a.frame_ends[f0] = base;
                u0 += r_grp; v0 += q_grp;
                if (u0 >= W) { u0 -= W; ++v0; }
                if (v0 >= H) { v0 -= H; ++f0; }
            }
        }
        // the bulk store issued from the other buffer must have finished reading it before the next tile fills it
        if (tid == 0) bulk_wait_read<kOutBufs - 2>();
        fence_proxy_async_smem();
        __syncthreads();                                                // records packed
        // split [skew, skew + 3 total) into head | 16-byte aligned middle | tail (element indices in the staging buffer)
        OutT* const buf0 = buf - skew;
        const unsigned first = skew, end = skew + total * 3u;
        unsigned mid0 = (first + kPerWord - 1) / kPerWord * kPerWord, mid1 = end / kPerWord * kPerWord;
        if (mid1 < mid0) { mid0 = end; mid1 = end; }
        OutT* gdst = gout + base * 3ull - skew;                         // 16-byte aligned
        if (tid == 0) {
            if (mid1 > mid0) bulk_store(gdst + mid0, buf0 + mid0, (mid1 - mid0) * (unsigned)sizeof(OutT));
            bulk_commit();
            const unsigned long long tn = t + kStages;
            if (tn < t1) {
                mbar_arrive_expect_tx(&full[stage], kInBytes);
                bulk_load(in_s + (size_t)stage * kTile, gin + tn * kTile, kInBytes, &full[stage]);
            }
        }
        if (tid >= 32 && tid < 32 + (unsigned)kPerWord) {               // head and tail: a few scalar stores by warp 1
            const unsigned k = tid - 32;
            if (first + k < mid0 && first + k < end) gdst[first + k] = buf0[first + k];
            if (mid1 + k < end && mid1 >= mid0 && mid1 + k >= mid0) gdst[mid1 + k] = buf0[mid1 + k];
        }
        if (++stage == kStages) { stage = 0; parity ^= 1u; }
        if (++ob == kOutBufs) ob = 0;
    }
    if (tid == 0) bulk_wait_all<0>();
}

// frame_ends (records written up to each frame's end) -> per-frame counts
__global__ void k1_frame_counts(const unsigned long long* __restrict__ ends, unsigned n, unsigned long long* counts) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) counts[i] = ends[i] - (i ? ends[i - 1] : 0ull);
}

// T . [x y z 1]^T (other_tools/transfer_T_icp.py:10-12), rows left to right
__global__ void k_transform_points(const double* __restrict__ in, double* __restrict__ out, unsigned long long n,
                                   const double* __restrict__ T) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double x = in[3 * i], y = in[3 * i + 1], z = in[3 * i + 2];
#pragma unroll
        for (int k = 0; k < 3; ++k)
            out[3 * i + k] = pose_canon(dadd(dadd(dadd(dmul(T[4 * k], x), dmul(T[4 * k + 1], y)), dmul(T[4 * k + 2], z)), T[4 * k + 3]));
    }
}

// point_camera on explicit points: Rinv . (p - t)
__global__ void k_pose_apply_points(const double* __restrict__ in, double* __restrict__ out, unsigned long long n,
                                    const double* __restrict__ rt) {
    Pose pose;
    pose_load(rt, pose);
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double wx, wy, wz;
        pose_apply(pose, in[3 * i], in[3 * i + 1], in[3 * i + 2], wx, wy, wz);
        out[3 * i] = pose_canon(wx); out[3 * i + 1] = pose_canon(wy); out[3 * i + 2] = pose_canon(wz);
    }
}

__global__ void k1_build_ztab(double* tab, int n, int mode, double depth_scale, double fB) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        bool valid;
        tab[i] = decode_z((double)i, mode, depth_scale, fB, valid);
    }
}

// ------------------------------------------------------------------ launch plumbing
static size_t elem_size(int dtype) { return dtype == R3D_U8 ? 1 : (dtype == R3D_U16 ? 2 : 4); }

template <typename DepthT, typename OutT, bool kWorld>
static int launch_k1_typed(r3d_ctx* ctx, cudaStream_t st, K1Args a, bool bulk_ok, int compact, unsigned long long total) {
    const size_t table_bytes = ((size_t)(a.W + a.H) * 8 + 127) / 128 * 128;
    if (compact) {
        const unsigned long long n_tiles = (total + K1_TILE - 1) / K1_TILE;
        R3D_TRY(scratch_reserve(ctx, SCR_TILE, (size_t)n_tiles * 16 + (size_t)a.n_frames * 8 + 512));
        unsigned long long* counts = (unsigned long long*)ctx->scratch[SCR_TILE];
        unsigned long long* offsets = counts + n_tiles;
        a.tile_counts = counts;
        a.tile_offsets = offsets;
        a.px_begin = 0;
        a.px_count = total;
        const int grid = (int)((n_tiles < (unsigned long long)ctx->sm_count * 8) ? n_tiles : (unsigned long long)ctx->sm_count * 8);
        const bool fast = bulk_ok && total >= K1_TILE && a.frame_counts != nullptr;
        if (fast) {
            // pass 1: full tiles with vector loads, the partial last tile (if any) with the generic counter
            a.n_tiles = total / K1_TILE;
            a.frame_ends = offsets + n_tiles;
            k1_count_tiles<DepthT><<<(unsigned)((a.n_tiles < (unsigned long long)ctx->sm_count * 16) ? a.n_tiles : (unsigned long long)ctx->sm_count * 16), K1_THREADS, 0, st>>>(a);
            ctx->launches++;
            if (a.n_tiles < n_tiles) {
                K1Args tail = a;
                tail.frame_counts = nullptr;
                tail.n_tiles = a.n_tiles;      // first tile of the tail
                k1_count_tail<DepthT><<<1, K1_THREADS, 0, st>>>(tail);
                ctx->launches++;
            }
        } else {
            k1_count<DepthT><<<grid, K1_THREADS, 0, st>>>(a);
            ctx->launches++;
        }
        size_t tmp_bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, counts, offsets, (int)n_tiles, st);
        R3D_TRY(scratch_reserve(ctx, SCR_CUBTMP, tmp_bytes + 256));
        R3D_CUDA_OK(ctx, cub::DeviceScan::ExclusiveSum(ctx->scratch[SCR_CUBTMP], tmp_bytes, counts, offsets, (int)n_tiles, st));
        ctx->launches++;
        if (fast) {
            constexpr int kGroups = sizeof(OutT) == 4 ? K1V_GROUPS : 1;
            const unsigned long long n_full = a.n_tiles;                 // full 1024-pixel tiles (counted by pass 1)
            const unsigned long long n_super = n_full / kGroups;
            if (n_super) {
                K1Args m = a;
                m.n_tiles = n_super;
                const size_t tables_v = ((size_t)(a.W + 1 + a.H) * 8 + 127) / 128 * 128;
                const size_t smem = 256 + tables_v + (size_t)K1V_STAGES * K1V_GROUP * kGroups * sizeof(DepthT) +
                                    2 * ((size_t)K1V_GROUP * kGroups * 3 + 16 / sizeof(OutT)) * sizeof(OutT);
                auto kern = a.mode == R3D_MODE_DEPTH ? k1_bulk_compact<DepthT, OutT, kWorld, 0, kGroups> : k1_bulk_compact<DepthT, OutT, kWorld, 1, kGroups>;
                R3D_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                int per_sm = 0;
                R3D_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, K1_THREADS, smem));
                if (per_sm < 1) return set_error(ctx, R3D_ERR_UNSUPPORTED, "image %ux%u needs %zu B of shared memory per CTA", a.W, a.H, smem);
                unsigned long long g2 = (unsigned long long)ctx->sm_count * per_sm;
                if (g2 > n_super) g2 = n_super;
                kern<<<(unsigned)g2, K1_THREADS, smem, st>>>(m);
                ctx->launches++;
            }
            if (n_super * kGroups < n_tiles) {
                // the 1024-pixel tiles after the last whole bulk tile, and the partial last one
                K1Args tail = a;
                tail.n_tiles = n_super * kGroups;
                R3D_CUDA_OK(ctx, cudaFuncSetAttribute(k1_compact<DepthT, OutT, kWorld>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)table_bytes));
                k1_compact<DepthT, OutT, kWorld><<<1, K1_THREADS, table_bytes, st>>>(tail);
                ctx->launches++;
            }
            k1_frame_counts<<<(a.n_frames + 255) / 256, 256, 0, st>>>(a.frame_ends, a.n_frames, a.frame_counts);
            ctx->launches++;
            R3D_CUDA_OK(ctx, cudaGetLastError());
            return R3D_OK;
        }
        R3D_CUDA_OK(ctx, cudaFuncSetAttribute(k1_compact<DepthT, OutT, kWorld>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)table_bytes));
        k1_compact<DepthT, OutT, kWorld><<<grid, K1_THREADS, table_bytes, st>>>(a);
        ctx->launches++;
        R3D_CUDA_OK(ctx, cudaGetLastError());
        return R3D_OK;
    }
    unsigned long long done = 0;
    constexpr int kGroups = sizeof(OutT) == 4 ? K1V_GROUPS : 1;      // 24 KB (float) / 24 KB (double) of records per tile
    constexpr int kTileV = K1V_GROUP * kGroups;
    if (bulk_ok && total >= (unsigned long long)kTileV) {
        a.n_tiles = total / kTileV;
        const size_t tables_v = ((size_t)(a.W + 1 + a.H) * 8 + 127) / 128 * 128;
        const size_t smem = 128 + tables_v + (size_t)K1V_STAGES * kTileV * sizeof(DepthT) + (size_t)2 * kTileV * 3 * sizeof(OutT);
        auto kern = a.mode == R3D_MODE_DEPTH ? k1_bulk_vec<DepthT, OutT, kWorld, 0, kGroups> : k1_bulk_vec<DepthT, OutT, kWorld, 1, kGroups>;
        R3D_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        R3D_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, K1_THREADS, smem));
        if (per_sm < 1) return set_error(ctx, R3D_ERR_UNSUPPORTED, "image %ux%u needs %zu B of shared memory per CTA", a.W, a.H, smem);
        unsigned long long grid = (unsigned long long)ctx->sm_count * per_sm;
        if (grid > a.n_tiles) grid = a.n_tiles;
        kern<<<(unsigned)grid, K1_THREADS, smem, st>>>(a);
        ctx->launches++;
        done = a.n_tiles * kTileV;
    }
    if (done < total) {
        a.px_begin = done;
        a.px_count = total - done;
        unsigned long long blocks = (a.px_count + K1_THREADS - 1) / K1_THREADS;
        const unsigned long long cap = (unsigned long long)ctx->sm_count * 16;
        if (blocks > cap) blocks = cap;
        R3D_CUDA_OK(ctx, cudaFuncSetAttribute(k1_generic<DepthT, OutT, kWorld>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)table_bytes));
        k1_generic<DepthT, OutT, kWorld><<<(unsigned)blocks, K1_THREADS, table_bytes, st>>>(a);
        ctx->launches++;
    }
    R3D_CUDA_OK(ctx, cudaGetLastError());
    return R3D_OK;
}

template <typename DepthT>
static int launch_k1_depth(r3d_ctx* ctx, cudaStream_t st, const K1Args& a, bool bulk_ok, int compact, int out_dtype,
                           unsigned long long total) {
    const bool world = a.rt != nullptr;
    if (out_dtype == R3D_OUT_F32)
        return world ? launch_k1_typed<DepthT, float, true>(ctx, st, a, bulk_ok, compact, total)
                     : launch_k1_typed<DepthT, float, false>(ctx, st, a, bulk_ok, compact, total);
    return world ? launch_k1_typed<DepthT, double, true>(ctx, st, a, bulk_ok, compact, total)
                 : launch_k1_typed<DepthT, double, false>(ctx, st, a, bulk_ok, compact, total);
}

// Everything on the device already: depth/out/rt are device pointers.
static int launch_k1(r3d_ctx* ctx, cudaStream_t st, const void* d_depth, int dtype, int W, int H, size_t pitch, int n_frames,
                     const double intr[4], const double* d_rt, int mode, double depth_scale, double fB, int compact,
                     int out_dtype, void* d_out, unsigned long long* d_frame_counts) {
    K1Args a;
    memset(&a, 0, sizeof a);
    a.depth = d_depth; a.out = d_out; a.rt = d_rt;
    a.W = (unsigned)W; a.H = (unsigned)H; a.WH = (unsigned)W * (unsigned)H; a.n_frames = (unsigned)n_frames;
    a.pitch = pitch;
    a.fx = intr[0]; a.fy = intr[1]; a.cx = intr[2]; a.cy = intr[3];
    a.depth_scale = depth_scale; a.fB = fB; a.mode = mode;
    a.valid_fast = (depth_scale >= 1e-250 && depth_scale <= 1e250) ? 1 : 0;
    a.frame_counts = d_frame_counts;
    a.ztab = nullptr;
    if (mode == R3D_MODE_DISPARITY && dtype != R3D_F32) {
        const int n = dtype == R3D_U8 ? 256 : 65536;
        if (!ctx->ztab) R3D_CUDA_OK(ctx, cudaMalloc(&ctx->ztab, 65536 * sizeof(double)));
        if (ctx->ztab_n != n || ctx->ztab_scale != depth_scale || ctx->ztab_fB != fB) {
            k1_build_ztab<<<(n + 255) / 256, 256, 0, st>>>(ctx->ztab, n, mode, depth_scale, fB);
            ctx->launches++;
            ctx->ztab_n = n; ctx->ztab_scale = depth_scale; ctx->ztab_fB = fB;
        }
        a.ztab = ctx->ztab;
    }
    const unsigned long long total = (unsigned long long)n_frames * a.WH;
    const size_t es = elem_size(dtype);
    const size_t osz = out_dtype == R3D_OUT_F32 ? 4 : 8;
    const bool bulk_ok = W >= K1_THREADS && H >= 8 && pitch == (size_t)W * es && ((uintptr_t)d_depth % 16 == 0) && ((uintptr_t)d_out % 16 == 0) &&
                         ((size_t)(W + H) * 8 + (size_t)K1_STAGES * K1_TILE * es + 3 * (size_t)K1_TILE * 3 * osz < 200 * 1024) &&
                         (size_t)(W + H) * 8 < 100 * 1024;
    switch (dtype) {
        case R3D_U8: return launch_k1_depth<unsigned char>(ctx, st, a, bulk_ok, compact, out_dtype, total);
        case R3D_U16: return launch_k1_depth<unsigned short>(ctx, st, a, bulk_ok, compact, out_dtype, total);
        case R3D_F32: return launch_k1_depth<float>(ctx, st, a, bulk_ok, compact, out_dtype, total);
    }
    return set_error(ctx, R3D_ERR_ARG, "unknown depth dtype %d", dtype);
}

}  // namespace r3d

using namespace r3d;

// ------------------------------------------------------------------ C ABI
extern "C" int r3d_pose_to_rt(const double* poses, int n, double t_scale, double* rt) {
    if (!poses || !rt || n < 0) return set_error(nullptr, R3D_ERR_ARG, "r3d_pose_to_rt: null argument");
    for (int i = 0; i < n; ++i) {
        const double* q = poses + (size_t)i * 7;
        double* o = rt + (size_t)i * 12;
        // scipy Rotation.from_quat: normalise; as_matrix: the formula below (scalar-last)
        const double nrm = sqrt(((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3]);
        if (!(nrm > 0.0)) return set_error(nullptr, R3D_ERR_ARG, "Found zero norm quaternions in `quat` (frame %d)", i);
        const double x = q[0] / nrm, y = q[1] / nrm, z = q[2] / nrm, w = q[3] / nrm;
        const double x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w;
        const double xy = x * y, zw = z * w, xz = x * z, yw = y * w, yz = y * z, xw = x * w;
        const double m00 = ((x2 - y2) - z2) + w2, m01 = 2.0 * (xy - zw), m02 = 2.0 * (xz + yw);
        const double m10 = 2.0 * (xy + zw), m11 = ((-x2 + y2) - z2) + w2, m12 = 2.0 * (yz - xw);
        const double m20 = 2.0 * (xz - yw), m21 = 2.0 * (yz + xw), m22 = ((-x2 - y2) + z2) + w2;
        // np.matrix(...).I: general inverse (cofactor form, fixed order)
        const double c00 = m11 * m22 - m12 * m21, c01 = m12 * m20 - m10 * m22, c02 = m10 * m21 - m11 * m20;
        const double det = (m00 * c00 + m01 * c01) + m02 * c02;
        o[0] = c00 / det; o[1] = (m02 * m21 - m01 * m22) / det; o[2] = (m01 * m12 - m02 * m11) / det;
        o[3] = c01 / det; o[4] = (m00 * m22 - m02 * m20) / det; o[5] = (m02 * m10 - m00 * m12) / det;
        o[6] = c02 / det; o[7] = (m01 * m20 - m00 * m21) / det; o[8] = (m00 * m11 - m01 * m10) / det;
        o[9] = t_scale * q[4]; o[10] = t_scale * q[5]; o[11] = t_scale * q[6];
    }
    return R3D_OK;
}

extern "C" int r3d_backproject_rt(r3d_ctx* ctx, const void* depth, int dtype, int W, int H, size_t pitch, int n_frames,
                                  const double intr[4], const double* rt, int mode, double depth_scale, double fB,
                                  int compact, int out_dtype, void* out_xyz, uint64_t* out_counts) {
    if (!ctx) return set_error(nullptr, R3D_ERR_ARG, "null context");
    if (!depth || !out_xyz || !intr) return set_error(ctx, R3D_ERR_ARG, "r3d_backproject: null buffer");
    if (W <= 0 || H <= 0 || n_frames < 0 || W > 65535 || H > 65535) return set_error(ctx, R3D_ERR_ARG, "bad image shape %dx%dx%d", n_frames, H, W);
    if (dtype < R3D_U8 || dtype > R3D_F32) return set_error(ctx, R3D_ERR_ARG, "bad depth dtype %d", dtype);
    if (mode != R3D_MODE_DEPTH && mode != R3D_MODE_DISPARITY) return set_error(ctx, R3D_ERR_ARG, "bad mode %d", mode);
    if (out_dtype != R3D_OUT_F32 && out_dtype != R3D_OUT_F64) return set_error(ctx, R3D_ERR_ARG, "bad out dtype %d", out_dtype);
    const size_t es = elem_size(dtype);
    if (pitch == 0) pitch = (size_t)W * es;
    if (pitch < (size_t)W * es) return set_error(ctx, R3D_ERR_ARG, "pitch %zu smaller than a row", pitch);
    if (dtype != R3D_U8 && pitch % es) return set_error(ctx, R3D_ERR_ARG, "pitch %zu not a multiple of the sample size", pitch);
    DeviceSetter ds(ctx->device);
    if (n_frames == 0) return R3D_OK;
    const size_t osz = out_dtype == R3D_OUT_F32 ? 4 : 8;
    const size_t frame_in = (size_t)H * pitch, frame_out = (size_t)W * H * 3 * osz;
    const bool depth_dev = is_device_ptr(depth), out_dev = is_device_ptr(out_xyz);
    const bool counts_dev = out_counts && is_device_ptr(out_counts);

    // pose table on the device
    const double* d_rt = nullptr;
    if (rt) {
        if (is_device_ptr(rt)) d_rt = rt;
        else {
            R3D_TRY(scratch_reserve(ctx, SCR_POSE, (size_t)n_frames * 96));
            R3D_CUDA_OK(ctx, cudaMemcpyAsync(ctx->scratch[SCR_POSE], rt, (size_t)n_frames * 96, cudaMemcpyHostToDevice, ctx->stream));
            // the staging streams below must see the table
            R3D_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
            d_rt = (const double*)ctx->scratch[SCR_POSE];
        }
    }
    // per-frame counters (compaction): written straight into the caller's array when that is device memory
    unsigned long long* d_counts = nullptr;
    cudaError_t ce;
    if (compact) {
        if (counts_dev) d_counts = (unsigned long long*)out_counts;
        else {
            R3D_TRY(scratch_reserve(ctx, SCR_MISC, (size_t)n_frames * 8 + 64));
            d_counts = (unsigned long long*)ctx->scratch[SCR_MISC];
        }
        R3D_CUDA_OK(ctx, cudaMemsetAsync(d_counts, 0, (size_t)n_frames * 8, ctx->stream));
    }
    int rc = R3D_OK;
    if (depth_dev && out_dev) {
        cudaEventRecord(ctx->ev_a, ctx->stream);
        rc = launch_k1(ctx, ctx->stream, depth, dtype, W, H, pitch, n_frames, intr, d_rt, mode, depth_scale, fB, compact,
                       out_dtype, out_xyz, d_counts);
        cudaEventRecord(ctx->ev_b, ctx->stream);
    } else if (compact) {
        // Compaction with a host buffer on either side, through the same ring of device slots as the plain mode: a chunk of
        // frames is uploaded, packed into the slot's record buffer (or straight into the caller's device buffer) and its
        // per-frame counts are read back; the records of chunk c leave for the host (at the running offset the counts of
        // the chunks before give) while chunk c + 1 is uploaded and packed.  Device memory is bounded by the ring, never
        // by the batch.
        const size_t per_frame = (depth_dev ? 0 : frame_in) + (out_dev ? 0 : frame_out);
        int chunk = (int)(ctx->stage_chunk_bytes / (per_frame ? per_frame : 1));
        if (chunk < 1) chunk = 1;
        if (chunk > n_frames) chunk = n_frames;
        int slots = ctx->stage_slots < 2 ? 2 : ctx->stage_slots;
        const size_t slot_in = (frame_in * chunk + 255) / 256 * 256, slot_out = (frame_out * chunk + 255) / 256 * 256;
        if (!depth_dev) rc = scratch_reserve(ctx, SCR_IN0, slot_in * slots);
        if (rc == R3D_OK && !out_dev) rc = scratch_reserve(ctx, SCR_OUT0, slot_out * slots);
        unsigned long long* h_counts = nullptr;     // pinned mirror of the per-frame counts
        if (rc == R3D_OK && cudaHostAlloc((void**)&h_counts, (size_t)n_frames * 8, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            rc = set_error(ctx, R3D_ERR_OOM, "cudaHostAlloc(per-frame counts) failed");
        }
        cudaStream_t s_in = ctx->copy_stream[0], s_out = ctx->copy_stream[1], s_k = ctx->stream;
        unsigned long long run_off = 0;             // records written before the chunk being finished
        const int n_chunks = (n_frames + chunk - 1) / chunk;
        // finish chunk c: wait for its counts, then (host output) send its records on their way
        auto finish_chunk = [&](int c) -> int {
            const int s = c % slots, f0 = c * chunk;
            const int nf = (n_frames - f0 < chunk) ? n_frames - f0 : chunk;
            R3D_CUDA_OK(ctx, cudaEventSynchronize(ctx->ev_k[s]));
            unsigned long long tot = 0;
            for (int i = 0; i < nf; ++i) tot += h_counts[f0 + i];
            if (!out_dev && tot) {
                R3D_CUDA_OK(ctx, cudaMemcpyAsync((char*)out_xyz + (size_t)run_off * 3 * osz, (char*)ctx->scratch[SCR_OUT0] + (size_t)s * slot_out,
                                                 (size_t)tot * 3 * osz, cudaMemcpyDeviceToHost, s_out));
            }
            R3D_CUDA_OK(ctx, cudaEventRecord(ctx->ev_out[s], s_out));
            run_off += tot;
            return R3D_OK;
        };
        for (int c = 0; c < n_chunks && rc == R3D_OK; ++c) {
            const int s = c % slots, f0 = c * chunk;
            const int nf = (n_frames - f0 < chunk) ? n_frames - f0 : chunk;
            const void* din = depth_dev ? (const void*)((const char*)depth + (size_t)f0 * frame_in)
                                        : (const void*)((const char*)ctx->scratch[SCR_IN0] + (size_t)s * slot_in);
            if (!depth_dev) {
                if (c >= slots) cudaStreamWaitEvent(s_in, ctx->ev_k[s], 0);      // the kernel that read this slot is done
                cudaMemcpyAsync((void*)din, (const char*)depth + (size_t)f0 * frame_in, frame_in * nf, cudaMemcpyHostToDevice, s_in);
                cudaEventRecord(ctx->ev_in[s], s_in);
                cudaStreamWaitEvent(s_k, ctx->ev_in[s], 0);
            }
            // a device output is packed in place: the chunk starts where the chunks before it ended, so their counts must be in
            if (out_dev && c > 0) rc = finish_chunk(c - 1);
            if (rc != R3D_OK) break;
            void* dout = out_dev ? (void*)((char*)out_xyz + (size_t)run_off * 3 * osz) : (void*)((char*)ctx->scratch[SCR_OUT0] + (size_t)s * slot_out);
            if (!out_dev && c >= slots) cudaStreamWaitEvent(s_k, ctx->ev_out[s], 0);   // the slot's records have been read back
            // (launch_k1 takes the generic kernel for a chunk whose first record is not 16-byte aligned)
            rc = launch_k1(ctx, s_k, din, dtype, W, H, pitch, nf, intr, d_rt ? d_rt + (size_t)f0 * 12 : nullptr, mode, depth_scale, fB, 1, out_dtype,
                           dout, d_counts + f0);
            cudaMemcpyAsync(h_counts + f0, d_counts + f0, (size_t)nf * 8, cudaMemcpyDeviceToHost, s_k);
            cudaEventRecord(ctx->ev_k[s], s_k);
            if (!out_dev && c > 0 && rc == R3D_OK) rc = finish_chunk(c - 1);
        }
        if (rc == R3D_OK && n_chunks > 0) rc = finish_chunk(n_chunks - 1);
        cudaStreamSynchronize(s_in);
        cudaStreamSynchronize(s_k);
        cudaStreamSynchronize(s_out);
        if (h_counts) cudaFreeHost(h_counts);
    } else {
        // host-side buffers: a ring of device slots; uploads, kernels and read-backs run on three streams chained by
        // events, so the read-back engine (the bound: 12 of the 14 bytes per pixel) never waits for an upload
        const size_t per_frame = (depth_dev ? 0 : frame_in) + (out_dev ? 0 : frame_out);
        int chunk = (int)(ctx->stage_chunk_bytes / (per_frame ? per_frame : 1));
        if (chunk < 1) chunk = 1;
        if (chunk > n_frames) chunk = n_frames;
        int slots = ctx->stage_slots;
        if ((long long)slots * chunk > (long long)n_frames + chunk - 1) slots = (n_frames + chunk - 1) / chunk;
        // slot strides rounded up to 256 bytes: every slot starts 16-byte aligned, which the bulk-copy kernel needs
        const size_t slot_in = (frame_in * chunk + 255) / 256 * 256, slot_out = (frame_out * chunk + 255) / 256 * 256;
        if (!depth_dev) rc = scratch_reserve(ctx, SCR_IN0, slot_in * slots);
        if (rc == R3D_OK && !out_dev) rc = scratch_reserve(ctx, SCR_OUT0, slot_out * slots);
        cudaStream_t s_in = ctx->copy_stream[0], s_out = ctx->copy_stream[1], s_k = ctx->stream;
        for (int f0 = 0, c = 0; f0 < n_frames && rc == R3D_OK; f0 += chunk, ++c) {
            const int s = c % slots;
            const int nf = (n_frames - f0 < chunk) ? n_frames - f0 : chunk;
            const void* din = depth_dev ? (const void*)((const char*)depth + (size_t)f0 * frame_in)
                                        : (const void*)((const char*)ctx->scratch[SCR_IN0] + (size_t)s * slot_in);
            void* dout = out_dev ? (void*)((char*)out_xyz + (size_t)f0 * frame_out) : (void*)((char*)ctx->scratch[SCR_OUT0] + (size_t)s * slot_out);
            if (!depth_dev) {
                if (c >= slots) cudaStreamWaitEvent(s_in, ctx->ev_k[s], 0);      // the kernel that read this slot is done
                cudaMemcpyAsync((void*)din, (const char*)depth + (size_t)f0 * frame_in, frame_in * nf, cudaMemcpyHostToDevice, s_in);
                cudaEventRecord(ctx->ev_in[s], s_in);
                cudaStreamWaitEvent(s_k, ctx->ev_in[s], 0);
            }
            if (!out_dev && c >= slots) cudaStreamWaitEvent(s_k, ctx->ev_out[s], 0);   // the slot's records have been read back
            rc = launch_k1(ctx, s_k, din, dtype, W, H, pitch, nf, intr, d_rt ? d_rt + (size_t)f0 * 12 : nullptr, mode, depth_scale,
                           fB, 0, out_dtype, dout, nullptr);
            cudaEventRecord(ctx->ev_k[s], s_k);
            if (rc == R3D_OK && !out_dev) {
                cudaStreamWaitEvent(s_out, ctx->ev_k[s], 0);
                cudaMemcpyAsync((char*)out_xyz + (size_t)f0 * frame_out, dout, frame_out * nf, cudaMemcpyDeviceToHost, s_out);
                cudaEventRecord(ctx->ev_out[s], s_out);
            }
        }
        cudaStreamSynchronize(s_in);
        cudaStreamSynchronize(s_k);
        cudaStreamSynchronize(s_out);
    }
    if (rc != R3D_OK) return rc;
    if (out_counts) {
        if (compact) {
            if (!counts_dev) {
                cudaMemcpyAsync(out_counts, d_counts, (size_t)n_frames * 8, cudaMemcpyDeviceToHost, ctx->stream);
                cudaStreamSynchronize(ctx->stream);
            }
        } else if (!counts_dev) {
            for (int i = 0; i < n_frames; ++i) out_counts[i] = (uint64_t)W * H;
        } else {
            uint64_t* h = (uint64_t*)malloc((size_t)n_frames * 8);
            for (int i = 0; i < n_frames; ++i) h[i] = (uint64_t)W * H;
            cudaMemcpy(out_counts, h, (size_t)n_frames * 8, cudaMemcpyHostToDevice);
            free(h);
        }
    }
    rc = finish(ctx);
    if (rc == R3D_OK && depth_dev && out_dev && ctx->blocking) cudaEventElapsedTime(&ctx->last_kernel_ms, ctx->ev_a, ctx->ev_b);
    return rc;
}

extern "C" int r3d_backproject(r3d_ctx* ctx, const void* depth, int dtype, int W, int H, size_t pitch, int n_frames,
                               const double intr[4], const double* poses, int mode, double depth_scale, double fB,
                               int compact, float* out_xyz, uint64_t* out_counts) {
    if (!ctx) return set_error(nullptr, R3D_ERR_ARG, "null context");
    if (n_frames < 0) return set_error(ctx, R3D_ERR_ARG, "negative frame count");
    double* rt = nullptr;
    if (poses) {
        if (is_device_ptr(poses)) return set_error(ctx, R3D_ERR_ARG, "poses must be host memory (converted per frame on the host)");
        rt = (double*)malloc((size_t)(n_frames ? n_frames : 1) * 96);
        int rc = r3d_pose_to_rt(poses, n_frames, 1.0, rt);
        if (rc != R3D_OK) { free(rt); strncpy(ctx->err, g_last_error, sizeof ctx->err - 1); return rc; }
    }
    int rc = r3d_backproject_rt(ctx, depth, dtype, W, H, pitch, n_frames, intr, rt, mode, depth_scale, fB, compact,
                                R3D_OUT_F32, out_xyz, out_counts);
    free(rt);
    return rc;
}

static int points_affine(r3d_ctx* ctx, const double* xyz, uint64_t n, const double* coeffs, int n_coeffs, double* out_xyz,
                         bool homogeneous) {
    if (!ctx) return set_error(nullptr, R3D_ERR_ARG, "null context");
    if ((!xyz || !out_xyz) && n) return set_error(ctx, R3D_ERR_ARG, "null buffer");
    if (!coeffs) return set_error(ctx, R3D_ERR_ARG, "null transform");
    if (n == 0) return R3D_OK;
    DeviceSetter ds(ctx->device);
    const bool in_dev = is_device_ptr(xyz), out_dev = is_device_ptr(out_xyz);
    const size_t bytes = (size_t)n * 24;
    R3D_TRY(scratch_reserve(ctx, SCR_POSE, 128));
    R3D_CUDA_OK(ctx, cudaMemcpyAsync(ctx->scratch[SCR_POSE], coeffs, (size_t)n_coeffs * 8, cudaMemcpyHostToDevice, ctx->stream));
    const double* din = xyz;
    double* dout = out_xyz;
    if (!in_dev) {
        R3D_TRY(scratch_reserve(ctx, SCR_IN0, bytes));
        R3D_CUDA_OK(ctx, cudaMemcpyAsync(ctx->scratch[SCR_IN0], xyz, bytes, cudaMemcpyHostToDevice, ctx->stream));
        din = (const double*)ctx->scratch[SCR_IN0];
    }
    if (!out_dev) {
        R3D_TRY(scratch_reserve(ctx, SCR_OUT0, bytes));
        dout = (double*)ctx->scratch[SCR_OUT0];
    }
    unsigned long long blocks = (n + 255) / 256;
    if (blocks > (unsigned long long)ctx->sm_count * 16) blocks = (unsigned long long)ctx->sm_count * 16;
    if (homogeneous) k_transform_points<<<(unsigned)blocks, 256, 0, ctx->stream>>>(din, dout, n, (const double*)ctx->scratch[SCR_POSE]);
    else k_pose_apply_points<<<(unsigned)blocks, 256, 0, ctx->stream>>>(din, dout, n, (const double*)ctx->scratch[SCR_POSE]);
    ctx->launches++;
    R3D_CUDA_OK(ctx, cudaGetLastError());
    if (!out_dev) {
        R3D_CUDA_OK(ctx, cudaMemcpyAsync(out_xyz, dout, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        R3D_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return finish(ctx);
}

extern "C" int r3d_transform_points(r3d_ctx* ctx, const double* xyz, uint64_t n, const double T[16], double* out_xyz) {
    return points_affine(ctx, xyz, n, T, 16, out_xyz, true);
}

extern "C" int r3d_pose_apply_points(r3d_ctx* ctx, const double* xyz, uint64_t n, const double rt[12], double* out_xyz) {
    return points_affine(ctx, xyz, n, rt, 12, out_xyz, false);
}
