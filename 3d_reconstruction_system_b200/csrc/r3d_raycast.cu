// r3d_raycast.cu -- K3 for a bounded range: insertPointCloud's computeUpdate (free cells along every ray, endpoint
// occupied) for BATCHES of scans, two batches in flight.
//
// Replaces, with r3d_octree.cu's apply kernel, the `insertPointCloud(points, origin, maxrange)` of the un-vendored
// `octomap` extension (named by BASELINE.json; upstream OccupancyOcTreeBase::computeUpdate / OcTreeBaseImpl::computeRayKeys).
//
//   k_scan_prepare   one thread per point: the per-ray set-up of computeRayKeys (fp64 divisions, a square root) is done
//                    ONCE, at full lane efficiency; rays with at least one free cell become 80-byte records, appended per
//                    scan, and the AABB of the origin key and every ray-end key is reduced into the scan's geometry.
//   k_scan_walk      persistent warps; a lane walks one ray at a time and idle lanes are re-filled from the record list
//                    as soon as K3_REFILL_MIN of them are idle -- a re-fill is five 16-byte loads, so the threshold can
//                    be small and nearly every lane of every warp instruction does useful work (rays of one image walk
//                    between 1 and ~1000 cells).  All scans of a batch share one ray counter: the drain at the end of a
//                    launch is paid once per batch, and the next batch's launch (other slot, other stream) fills it.
//   k_cells_list     touched brick cells of each scan's cube -> list;  k_cells_emit: list -> 136-byte records (free minus
//                    occupied), cube cleared.
//
// The per-scan delta is direct-mapped: a cube of brick cells (32 mask words each) sized from the scan's AABB, not from
// maxrange -- a forward-looking frustum touches a small corner of the (2 maxrange)^3 cube (KITTI-shape street scan at
// 0.1 m / 80 m: ~4 x 10^4 cells = 5 MB instead of 205^3 cells = 1.1 GB), so the masks of a whole batch live in L2.
// Free cells are collected per 4x4x4 sub-block in a 64-bit register mask (bit = x + 4y + 16z inside the sub-block: one
// IDP.4A on the packed local coordinates) and published with one red.or when the ray leaves the sub-block, unless the
// word fetched on entering it shows every bit already set; k_cells_emit permutes the words into the Morton order of
// the record format.
#include <vector>

#include <limits.h>
#include <stdlib.h>
#include <time.h>

#include "r3d_octree.cuh"

namespace r3d {

constexpr int K3_THREADS = 256;
constexpr int K3_MAX_BATCH = 8;
#ifndef K3_REFILL_MIN
#define K3_REFILL_MIN 8
#endif
#ifndef K3_STAGES
#define K3_STAGES 2      // shared-memory slots of the sub-block word fetch = steps between request and use
#endif
#ifndef K3_CHUNK
#define K3_CHUNK 512     // most rays a warp claims from the batch's counter at a time; halved until every resident warp gets a
                         // chunk (8 scans of 465 750 rays: 512, 4 scans: 256).  Long runs of consecutive rays per warp are
                         // what pays: measured per scan with 4 scans per batch, 64: 0.65, 128: 0.59, 192: 0.55, 256: 0.53,
                         // 384: 0.55, 512: 0.58 ms; with 8 per batch, 256: 0.498, 384: 0.487, 512 / 1024: 0.486 ms; claims
                         // that shrink towards the end of the batch (guided self-scheduling) were slower (0.54-0.60)
#endif

// One ray with at least one free cell, as computeRayKeys sets it up (r3d_math.cuh::ray_setup).
struct __align__(16) RayRec {
    double tm[3], td[3];   // tMax, tDelta
    float len;             // ray length (float, as upstream)
    uint32_t kxy;          // origin key: kx | ky << 16
    uint32_t kzs;          // kz | (sx + 1) << 16 | (sy + 1) << 18 | (sz + 1) << 20
    uint32_t exy;          // end key
    uint32_t ezf;          // ez | endpoint in range (occupied) << 16
    uint32_t pad[3];
};
static_assert(sizeof(RayRec) == 80, "five 16-byte loads per re-fill");

struct ScanDesc {
    const float* xyz;
    uint32_t n;
    float ox, oy, oz;
};

struct BatchArgs {
    ScanDesc scan[K3_MAX_BATCH];
    int n_scans;
    double maxrange, res, res_factor;
    int* geom;              // [n_scans][8]: min key x y z, then MINUS max key x y z (all atomicMin targets), 2 spare
    uint32_t* counters;     // [n_scans][CNT_COUNT]
    RayRec* rays;           // [n_scans][ray_stride]
    uint64_t ray_stride;
    uint32_t* cmasks;       // this slot's cubes: [n_scans][cube_cells][32]
    uint8_t* ctouched;      // [n_scans][cube_cells]
    uint32_t cube_cells;    // capacity of one cube (multiple of 4)
    uint32_t* lists;        // [n_scans][rec_cap]
    DeltaRecord* recs;      // [n_scans][rec_cap]
    uint32_t rec_cap;
};

// Brick grid of a scan: the AABB of its keys, one brick of margin on every side (the walk may overshoot the end key by
// one voxel before the length test stops it).  A pure function of the reduced geometry, so every kernel derives the same.
struct Grid {
    int bx0, by0, bz0;
    uint32_t dx, dy, dz, ncells, need;
    bool ok;
};
__device__ __forceinline__ Grid grid_of(const int* g, uint32_t cap) {
    Grid r;
    const int x0 = g[0], y0 = g[1], z0 = g[2], x1 = -g[3], y1 = -g[4], z1 = -g[5];
    r.bx0 = (x0 >> 3) - 1; r.by0 = (y0 >> 3) - 1; r.bz0 = (z0 >> 3) - 1;
    const int bx1 = (x1 >> 3) + 1, by1 = (y1 >> 3) + 1, bz1 = (z1 >> 3) + 1;
    // (keys must not reach the border of the key space: upstream's uint16 keys would wrap there, the hash path mirrors that)
    const bool valid = x0 <= x1 && y0 <= y1 && z0 <= z1 && r.bx0 >= 0 && r.by0 >= 0 && r.bz0 >= 0 && bx1 <= 8191 && by1 <= 8191 && bz1 <= 8191;
    r.dx = valid ? (uint32_t)(bx1 - r.bx0 + 1) : 0u;
    r.dy = valid ? (uint32_t)(by1 - r.by0 + 1) : 0u;
    r.dz = valid ? (uint32_t)(bz1 - r.bz0 + 1) : 0u;
    const unsigned long long cells = (unsigned long long)r.dx * r.dy * r.dz;
    r.need = !valid ? 0xffffffffu : (cells > 0xfffffff0ull ? 0xfffffff0u : (uint32_t)cells);
    r.ok = valid && cells <= cap;
    r.ncells = r.ok ? (uint32_t)cells : 0u;
    return r;
}

// ------------------------------------------------------------------ prepare
__global__ void __launch_bounds__(256) k_scan_prepare(const BatchArgs a) {
    const int s = blockIdx.y;
    const float* __restrict__ xyz = a.scan[s].xyz;
    const uint32_t n = a.scan[s].n;
    const float ox = a.scan[s].ox, oy = a.scan[s].oy, oz = a.scan[s].oz;
    int* geom = a.geom + s * 8;
    uint32_t* cnt = a.counters + s * CNT_COUNT;
    RayRec* rays = a.rays + (size_t)s * a.ray_stride;
    const unsigned lane = threadIdx.x & 31u;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        uint16_t kx, ky, kz;
        if (coord_to_key3(a.res_factor, ox, oy, oz, kx, ky, kz)) {   // (the host only sends scans whose origin has a key)
            atomicMin(geom + 0, (int)kx); atomicMin(geom + 1, (int)ky); atomicMin(geom + 2, (int)kz);
            atomicMin(geom + 3, -(int)kx); atomicMin(geom + 4, -(int)ky); atomicMin(geom + 5, -(int)kz);
        }
    }
    const uint32_t n_pad = (n + 31u) & ~31u;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += stride) {
        int rc = -2;
        bool in_range = false;
        Ray r;
        r.kx = r.ky = r.kz = r.ex = r.ey = r.ez = r.sx = r.sy = r.sz = 0;
        r.tmx = r.tmy = r.tmz = r.tdx = r.tdy = r.tdz = 0.0; r.length = 0.f;
        if (i < n) {
            const float px = xyz[3 * (size_t)i], py = xyz[3 * (size_t)i + 1], pz = xyz[3 * (size_t)i + 2];
            float fx, fy, fz;
            in_range = scan_point_end(ox, oy, oz, px, py, pz, a.maxrange, fx, fy, fz);
            rc = ray_setup(a.res, a.res_factor, ox, oy, oz, fx, fy, fz, r);
            // an endpoint inside the sensor's own voxel: no ray, but that voxel is occupied.  (rc < 0 with a valid origin
            // means the endpoint has no key: upstream ignores the point.)
            if (rc == 0 && in_range) cnt[CNT_ORIGIN_OCC] = 1u;
        }
        const bool walk = rc == 1;
        const unsigned wm = __ballot_sync(0xffffffffu, walk);
        if (wm == 0u) continue;
        const int mnx = __reduce_min_sync(0xffffffffu, walk ? r.ex : INT_MAX), mny = __reduce_min_sync(0xffffffffu, walk ? r.ey : INT_MAX),
                  mnz = __reduce_min_sync(0xffffffffu, walk ? r.ez : INT_MAX);
        const int mxx = __reduce_max_sync(0xffffffffu, walk ? r.ex : INT_MIN), mxy = __reduce_max_sync(0xffffffffu, walk ? r.ey : INT_MIN),
                  mxz = __reduce_max_sync(0xffffffffu, walk ? r.ez : INT_MIN);
        uint32_t base = 0;
        if (lane == 0) {
            base = atomicAdd(&cnt[CNT_NRAYS], (uint32_t)__popc(wm));
            if (mnx < __ldcg(geom + 0)) atomicMin(geom + 0, mnx);
            if (mny < __ldcg(geom + 1)) atomicMin(geom + 1, mny);
            if (mnz < __ldcg(geom + 2)) atomicMin(geom + 2, mnz);
            if (-mxx < __ldcg(geom + 3)) atomicMin(geom + 3, -mxx);
            if (-mxy < __ldcg(geom + 4)) atomicMin(geom + 4, -mxy);
            if (-mxz < __ldcg(geom + 5)) atomicMin(geom + 5, -mxz);
        }
        base = __shfl_sync(0xffffffffu, base, 0);
        if (walk) {
            RayRec rec;
            rec.tm[0] = r.tmx; rec.tm[1] = r.tmy; rec.tm[2] = r.tmz;
            rec.td[0] = r.tdx; rec.td[1] = r.tdy; rec.td[2] = r.tdz;
            rec.len = r.length;
            rec.kxy = (uint32_t)r.kx | ((uint32_t)r.ky << 16);
            rec.kzs = (uint32_t)r.kz | ((uint32_t)(r.sx + 1) << 16) | ((uint32_t)(r.sy + 1) << 18) | ((uint32_t)(r.sz + 1) << 20);
            rec.exy = (uint32_t)r.ex | ((uint32_t)r.ey << 16);
            rec.ezf = (uint32_t)r.ez | (in_range ? 0x10000u : 0u);
            rec.pad[0] = rec.pad[1] = rec.pad[2] = 0u;
            uint4* dst = reinterpret_cast<uint4*>(rays + base + __popc(wm & ((1u << lane) - 1u)));
            const uint4* src = reinterpret_cast<const uint4*>(&rec);
#pragma unroll
            for (int q = 0; q < 5; ++q) dst[q] = src[q];
        }
    }
}

// ------------------------------------------------------------------ walk
// Per-lane state.  The position inside the current brick is ONE register: P = lx | ly << 8 | lz << 16 with every field
// biased by 8 (values 8..15), so that a step of +-1 along an axis is one add and leaving the brick shows as a cleared bit 3
// of that field (15 + 1 = 16, 8 - 1 = 7) without a borrow into the next field.
struct WalkLane {
    double tmx, tmy, tmz, tdx, tdy, tdz, len;
    uint32_t P, eP;        // position / end position inside the brick
    uint32_t cell, ecell;  // brick cell (index into the slot's cubes) / end cell
    int dPx, dPy, dPz;     // P increment of one step along each axis
    int csx, csy, csz;     // cell increment of one brick step along each axis
    uint32_t widx;         // index of the current sub-block's free word in the 64-bit view of the masks
    uint64_t mask;         // cells of that sub-block visited by this ray (bit = x + 4y + 16z)
    uint64_t seen;         // what the word held when the ray entered the sub-block; 0 = not known (yet)
    int epos;              // position (of the unrolled step sequence) at which the word of the current sub-block was requested; -1: none pending
    unsigned steps;        // statistics: free cells recorded by this lane
};

// The word of the sub-block a ray enters is fetched with an ASYNCHRONOUS copy into a per-lane shared-memory slot
// (cp.async / LDGSTS): no destination register, so no scoreboard for the warp to wait on -- with a register destination
// ptxas waits for every fetch in flight at the head of the step loop (ncu: 35 % of all stall samples on that one wait,
// whatever the number of staging registers).  One commit group per step; a step first waits for the group committed
// K3_STAGES steps earlier (normally long complete), then reads its slot.
__device__ __forceinline__ void fetch_word_async(uint32_t smem_slot, const uint64_t* gptr, bool pred) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q cp.async.ca.shared.global [%0], [%1], 8;\n\t}" :: "r"(smem_slot), "l"(gptr), "r"((unsigned)pred) : "memory");
}
__device__ __forceinline__ void fetch_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void fetch_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(kPending) : "memory"); }
__device__ __forceinline__ uint64_t lds_u64(uint32_t smem_slot) {
    uint64_t v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(smem_slot) : "memory");
    return v;
}
// bit of the current cell in its sub-block's word, and the word's index
__device__ __forceinline__ uint64_t lane_bit(uint32_t P) { return 1ull << __dp4a(P & 0x00030303u, 0x00100401u, 0u); }
__device__ __forceinline__ uint32_t lane_widx(uint32_t cell, uint32_t P) { return cell * 16u + __dp4a((P >> 2) & 0x00010101u, 0x00040201u, 8u); }

// One DDA step of computeRayKeys for an active lane: advance along L.axis, pick the next axis with upstream's comparison
// chain (x if tMax.x < tMax.y and tMax.x < tMax.z; else y if tMax.y < tMax.z; else z), stop at the end key or when
// min(tMax) > length -- evaluated as "every tMax > length", the minimum itself is never needed.  The three mutually
// exclusive updates are predicated instructions (as C++ conditionals they compile to divergent branches).
// `slot` is the shared-memory slot (of this lane, for this position of the unrolled step sequence) the word fetched on
// entering a sub-block lands in; the word is moved into `seen` K3_STAGES steps after it was requested; a ray that leaves
// a sub-block earlier publishes without knowing the word (a redundant red.or, never a wrong bit).
template <int kPos>
__device__ __forceinline__ bool walk_step(uint64_t* masks64, uint8_t* touched, uint32_t total_cells, uint32_t* miss, WalkLane& L, uint32_t slot) {
    // the word requested K3_STAGES steps ago (same position of the sequence) has landed: it is what the sub-block held
    fetch_wait<K3_STAGES - 1>();
    if (L.epos == kPos) { L.seen = lds_u64(slot); L.epos = -1; }
    const uint32_t Pold = L.P;
    int cs;
    // axis of this step (upstream's comparison chain), then tMax += tDelta and position += step on that axis only.  The
    // three mutually exclusive updates are written without selects on the (busy) integer pipe: tMax through an FMA with a
    // 1.0 / 0.0 multiplier on the (idle) fp64 pipe -- fma(1, d, t) = rn(t + d) and fma(0, d, t) = t are exact -- and the
    // position through predicated adds.
    asm("{\n\t.reg .pred pxy, psx, psy, psz;\n\t.reg .b32 hx, hy, hz, zero;\n\t.reg .f64 mx, my, mz;\n\t"
        "setp.lt.f64 pxy, %2, %3;\n\tsetp.lt.and.f64 psx, %2, %4, pxy;\n\tsetp.lt.and.f64 psy, %3, %4, !pxy;\n\t"
        "or.pred psz, psx, psy;\n\tnot.pred psz, psz;\n\t"
        "selp.b32 hx, 0x3FF00000, 0, psx;\n\tselp.b32 hy, 0x3FF00000, 0, psy;\n\tselp.b32 hz, 0x3FF00000, 0, psz;\n\t"
        "mov.b32 zero, 0;\n\tmov.b64 mx, {zero, hx};\n\tmov.b64 my, {zero, hy};\n\tmov.b64 mz, {zero, hz};\n\t"
        "fma.rn.f64 %2, mx, %8, %2;\n\tfma.rn.f64 %3, my, %9, %3;\n\tfma.rn.f64 %4, mz, %10, %4;\n\t"
        "@psx add.s32 %0, %0, %5;\n\t@psy add.s32 %0, %0, %6;\n\t@psz add.s32 %0, %0, %7;\n\t"
        "mov.s32 %1, %13;\n\t@psx mov.s32 %1, %11;\n\t@psy mov.s32 %1, %12;\n\t}"
        : "+r"(L.P), "=&r"(cs), "+d"(L.tmx), "+d"(L.tmy), "+d"(L.tmz)
        : "r"(L.dPx), "r"(L.dPy), "r"(L.dPz), "d"(L.tdx), "d"(L.tdy), "d"(L.tdz), "r"(L.csx), "r"(L.csy), "r"(L.csz));
    const bool past = (L.tmx > L.len) & (L.tmy > L.len) & (L.tmz > L.len);
    // brick / sub-block bookkeeping on the packed position
    const bool new_brick = (~L.P & 0x00080808u) != 0u;
    L.P = (L.P & 0x00070707u) | 0x00080808u;
    L.cell += new_brick ? (uint32_t)cs : 0u;
    const bool new_sub = ((L.P ^ Pold) & 0x00040404u) != 0u;
    const bool done = ((L.cell == L.ecell) & (L.P == L.eP)) | past;
    // ---- leaving the sub-block (or the ray): publish its cells unless all of them are known to be set
    if (done | new_sub) {
#ifndef K3_NO_STATS
        L.steps += (unsigned)__popcll(L.mask);      // statistics: free cells recorded (DDA steps/s in the bench)
#endif
        if ((L.mask & ~L.seen) != 0ull) {
            atomicOr(reinterpret_cast<unsigned long long*>(masks64 + L.widx), (unsigned long long)L.mask);
            touched[L.widx >> 4] = 1;
        }
    }
    // ---- entering the next one
    const bool enter = new_sub & !done;
    if (enter) {
        if (L.cell >= total_cells) {   // memory-safety guard; the grid is sized so that it never trips
            *miss = 1u;
            L.cell = 0;
        }
        L.widx = lane_widx(L.cell, L.P);
        L.mask = 0;
        L.seen = 0;
        L.epos = kPos;
    }
    fetch_word_async(slot, masks64 + L.widx, enter);
    fetch_commit();
    L.mask |= lane_bit(L.P);
    return !done;
}

template <int kPos>
__device__ __forceinline__ void walk_sequence(uint64_t* masks64, uint8_t* touched, uint32_t total_cells, uint32_t* miss, WalkLane& L, uint32_t slot0, bool& active) {
    if (active) active = walk_step<kPos>(masks64, touched, total_cells, miss, L, slot0 + (uint32_t)kPos * K3_THREADS * 8u);
    if constexpr (kPos + 1 < K3_STAGES) walk_sequence<kPos + 1>(masks64, touched, total_cells, miss, L, slot0, active);
}

struct WalkGrid {
    int bx0, by0, bz0;
    uint32_t dx, dxy, cube_off;
};

#ifndef K3_MIN_CTAS
#define K3_MIN_CTAS 3
#endif

__global__ void __launch_bounds__(K3_THREADS, K3_MIN_CTAS) k_scan_walk(const BatchArgs a) {
    __shared__ WalkGrid sg[K3_MAX_BATCH];
    __shared__ uint32_t s_prefix[K3_MAX_BATCH + 1];
    __shared__ __align__(8) uint64_t s_words[K3_STAGES][K3_THREADS];
    if (threadIdx.x < (unsigned)a.n_scans) {
        const int s = threadIdx.x;
        const Grid g = grid_of(a.geom + s * 8, a.cube_cells);
        sg[s].bx0 = g.bx0; sg[s].by0 = g.by0; sg[s].bz0 = g.bz0;
        sg[s].dx = g.dx; sg[s].dxy = g.dx * g.dy; sg[s].cube_off = (uint32_t)s * a.cube_cells;
        s_prefix[s + 1] = g.ok ? a.counters[s * CNT_COUNT + CNT_NRAYS] : 0u;   // a scan whose cube does not fit is skipped (k_cells_list reports it)
        // the sensor's own voxel as an endpoint
        if (blockIdx.x == 0 && g.ok && a.counters[s * CNT_COUNT + CNT_ORIGIN_OCC]) {
            uint16_t kx, ky, kz;
            if (coord_to_key3(a.res_factor, a.scan[s].ox, a.scan[s].oy, a.scan[s].oz, kx, ky, kz)) {
                const uint32_t c = sg[s].cube_off + (uint32_t)((kx >> 3) - g.bx0) + g.dx * (uint32_t)((ky >> 3) - g.by0) + sg[s].dxy * (uint32_t)((kz >> 3) - g.bz0);
                const unsigned vox = brick_voxel_index(kx, ky, kz);
                atomicOr(a.cmasks + (size_t)c * 32 + (vox >> 5), 1u << (vox & 31u));
                a.ctouched[c] = 1;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        s_prefix[0] = 0;
        for (int s = 0; s < a.n_scans; ++s) { acc += s_prefix[s + 1]; s_prefix[s + 1] = acc; }
    }
    __syncthreads();
    const uint32_t total = s_prefix[a.n_scans];
    if (total == 0u) return;
    const unsigned lane = threadIdx.x & 31u;
    uint64_t* const masks64 = reinterpret_cast<uint64_t*>(a.cmasks);
    uint8_t* const touched = a.ctouched;
    const uint32_t total_cells = (uint32_t)a.n_scans * a.cube_cells;
    uint32_t* const miss = a.counters + CNT_GRID_MISS;
    unsigned long long* const ray_counter = reinterpret_cast<unsigned long long*>(a.counters + CNT_RAY_LO);
    bool active = false, exhausted = false;
    WalkLane L;
    memset(&L, 0, sizeof L);
    L.epos = -1;
    const uint32_t slot0 = (uint32_t)__cvta_generic_to_shared(&s_words[0][threadIdx.x]);
    // rays are claimed from the batch's counter a chunk at a time per warp (one atomic round trip per chunk, not per re-fill);
    // a small batch is cut finer so that every resident warp gets a share
    uint32_t chunk = K3_CHUNK;
    while (chunk > 32u && (unsigned long long)chunk * gridDim.x * (K3_THREADS / 32) > total) chunk >>= 1;
    uint32_t my_next = 0, my_end = 0;
    for (;;) {
        const unsigned act = __ballot_sync(0xffffffffu, active);
        const unsigned idle = ~act;
        if (!exhausted && __popc(idle) >= K3_REFILL_MIN) {
            if (my_next == my_end) {
                unsigned long long base = 0;
                const uint32_t take_n = chunk;
                if (lane == 0) base = atomicAdd(ray_counter, (unsigned long long)chunk);
                base = __shfl_sync(0xffffffffu, base, 0);
                my_next = base < total ? (uint32_t)base : total;
                my_end = base + take_n < total ? (uint32_t)base + take_n : total;
                if (my_next == my_end) { exhausted = true; continue; }
            }
            const uint32_t i = my_next + (uint32_t)__popc(idle & ((1u << lane) - 1u));
            const bool take = !active && i < my_end;
            my_next = my_next + (uint32_t)__popc(idle) < my_end ? my_next + (uint32_t)__popc(idle) : my_end;
            if (take) {
                int s = 0;
                while (i >= s_prefix[s + 1]) ++s;
                const uint4* rp = reinterpret_cast<const uint4*>(a.rays + (size_t)s * a.ray_stride + (i - s_prefix[s]));
                const uint4 q0 = __ldg(rp), q1 = __ldg(rp + 1), q2 = __ldg(rp + 2), q3 = __ldg(rp + 3), q4 = __ldg(rp + 4);
                L.tmx = __hiloint2double((int)q0.y, (int)q0.x); L.tmy = __hiloint2double((int)q0.w, (int)q0.z);
                L.tmz = __hiloint2double((int)q1.y, (int)q1.x); L.tdx = __hiloint2double((int)q1.w, (int)q1.z);
                L.tdy = __hiloint2double((int)q2.y, (int)q2.x); L.tdz = __hiloint2double((int)q2.w, (int)q2.z);
                L.len = (double)__uint_as_float(q3.x);
                const uint32_t kxy = q3.y, kzs = q3.z, exy = q3.w, ezf = q4.x;
                const int kx = (int)(kxy & 0xffffu), ky = (int)(kxy >> 16), kz = (int)(kzs & 0xffffu);
                const int ex = (int)(exy & 0xffffu), ey = (int)(exy >> 16), ez = (int)(ezf & 0xffffu);
                const int sx = (int)((kzs >> 16) & 3u) - 1, sy = (int)((kzs >> 18) & 3u) - 1, sz = (int)((kzs >> 20) & 3u) - 1;
                const WalkGrid g = sg[s];
                L.cell = g.cube_off + (uint32_t)((kx >> 3) - g.bx0) + g.dx * (uint32_t)((ky >> 3) - g.by0) + g.dxy * (uint32_t)((kz >> 3) - g.bz0);
                L.ecell = g.cube_off + (uint32_t)((ex >> 3) - g.bx0) + g.dx * (uint32_t)((ey >> 3) - g.by0) + g.dxy * (uint32_t)((ez >> 3) - g.bz0);
                L.P = ((uint32_t)(kx & 7) | ((uint32_t)(ky & 7) << 8) | ((uint32_t)(kz & 7) << 16)) | 0x00080808u;
                L.eP = ((uint32_t)(ex & 7) | ((uint32_t)(ey & 7) << 8) | ((uint32_t)(ez & 7) << 16)) | 0x00080808u;
                L.dPx = sx; L.dPy = sy * 256; L.dPz = sz * 65536;
                L.csx = sx; L.csy = sy * (int)g.dx; L.csz = sz * (int)g.dxy;
                if (ezf & 0x10000u) {   // the endpoint is an occupied cell (Morton order, like the record format); fire and forget
                    const unsigned vox = brick_voxel_index((uint32_t)ex, (uint32_t)ey, (uint32_t)ez);
                    atomicOr(a.cmasks + (size_t)L.ecell * 32 + (vox >> 5), 1u << (vox & 31u));
                    touched[L.ecell] = 1;
                }
                active = true;
                // the origin cell is the first free cell; its word is requested like any other sub-block's, into the last slot
                // (a fetch of the lane's previous ray may still be on its way there: fetch_wait<0> below)
                L.widx = lane_widx(L.cell, L.P);
                L.seen = 0;
                L.epos = K3_STAGES - 1;
                L.mask = lane_bit(L.P);
            }
            // (a dropped fetch of this lane may still be on its way into the same slot: let it land first, once per re-fill)
            fetch_wait<0>();
            fetch_word_async(slot0 + (K3_STAGES - 1) * K3_THREADS * 8u, masks64 + L.widx, take);
            fetch_commit();
            continue;
        }
        if (act == 0) break;   // no ray left anywhere in this warp
        const int keep_going = exhausted ? 0 : 32 - K3_REFILL_MIN;
        do {
            walk_sequence<0>(masks64, touched, total_cells, miss, L, slot0, active);
        } while (__popc(__ballot_sync(0xffffffffu, active)) > keep_going);
    }
    // statistics only: free-cell visits of this batch
    unsigned long long steps = L.steps;
    for (int o = 16; o > 0; o >>= 1) steps += __shfl_xor_sync(0xffffffffu, steps, o);
    if (lane == 0 && steps) atomicAdd(reinterpret_cast<unsigned long long*>(a.counters + CNT_STEPS_LO), steps);
}

// ------------------------------------------------------------------ list / emit
// touched cells of each scan's cube -> list (one byte per cell, four cells per thread and load)
__global__ void __launch_bounds__(256) k_cells_list(const BatchArgs a) {
    const int s = blockIdx.y;
    uint32_t* cnt = a.counters + s * CNT_COUNT;
    const Grid g = grid_of(a.geom + s * 8, a.cube_cells);
    if (!g.ok) {
        if (blockIdx.x == 0 && threadIdx.x == 0) { cnt[CNT_GRID_MISS] = 1u; cnt[CNT_GRID_NEED] = g.need; }
        return;
    }
    const uint32_t* touched = reinterpret_cast<const uint32_t*>(a.ctouched + (size_t)s * a.cube_cells);
    uint32_t* list = a.lists + (size_t)s * a.rec_cap;
    const uint32_t n_words = (g.ncells + 3u) / 4u;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += stride) {
        const uint32_t w = touched[i];
        if (!w) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            if ((w >> (8 * b)) & 0xffu) {
                const uint32_t q = atomicAdd(&cnt[CNT_DELTA], 1u);
                if (q < a.rec_cap) list[q] = i * 4u + (uint32_t)b;
            }
        }
    }
}

// free word in walk order (bit = x0 | x1 << 1 | y0 << 2 | y1 << 3 | z0 << 4 within a 32-bit half, z1 selects the half)
// -> Morton order (x0 | y0 << 1 | z0 << 2 | x1 << 3 | y1 << 4) of the record format
__device__ __forceinline__ uint32_t walk_to_morton(uint32_t w) {
    uint32_t r = 0;
#pragma unroll
    for (int m = 0; m < 32; ++m) {
        const int l = (m & 1) | ((m >> 3) & 1) << 1 | ((m >> 1) & 1) << 2 | ((m >> 4) & 1) << 3 | ((m >> 2) & 1) << 4;
        r |= ((w >> l) & 1u) << m;
    }
    return r;
}

// one warp per listed cell -> record (free already minus occupied); clears the cell and its byte.  Does nothing for a
// scan whose list overflowed (the host grows the buffers and runs the batch again).
__global__ void __launch_bounds__(256) k_cells_emit(const BatchArgs a) {
    const int s = blockIdx.y;
    const uint32_t* cnt = a.counters + s * CNT_COUNT;
    const Grid g = grid_of(a.geom + s * 8, a.cube_cells);
    if (!g.ok) return;
    const uint32_t n = cnt[CNT_DELTA];
    if (n > a.rec_cap) return;
    uint32_t* cmasks = a.cmasks + (size_t)s * a.cube_cells * 32;
    uint8_t* touched = a.ctouched + (size_t)s * a.cube_cells;
    const uint32_t* list = a.lists + (size_t)s * a.rec_cap;
    DeltaRecord* out = a.recs + (size_t)s * a.rec_cap;
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < n; q += warps) {
        const uint32_t cell = list[q];
        uint32_t w = cmasks[(size_t)cell * 32 + lane];
        if (lane >= 16) w = walk_to_morton(w);
        const uint32_t occ_w = __shfl_sync(0xffffffffu, w, lane & 15u);
        if (lane >= 16) w &= ~occ_w;   // occupied wins
        if (lane == 0) {
            const uint32_t cx = cell % g.dx, cy = (cell / g.dx) % g.dy, cz = cell / (g.dx * g.dy);
            out[q].key = (uint64_t)(uint32_t)(g.bx0 + (int)cx) | ((uint64_t)(uint32_t)(g.by0 + (int)cy) << 13) | ((uint64_t)(uint32_t)(g.bz0 + (int)cz) << 26);
            touched[cell] = 0;
        }
        out[q].mask[lane] = w;
        cmasks[(size_t)cell * 32 + lane] = 0;
    }
}

// ------------------------------------------------------------------ host side: the pipeline
struct ScanPipe {
    int B = 0;                       // scans per batch
    uint64_t cube_cells = 0;         // capacity of one cube (cells)
    uint64_t rec_cap = 0;            // records per scan
    uint64_t ray_cap = 0;            // rays per scan
    int* geom = nullptr;             // [2][B][8]
    uint32_t* counters = nullptr;    // [2][B][CNT_COUNT]
    uint32_t* lists = nullptr;       // [2][B][rec_cap]
    DeltaRecord* recs = nullptr;     // [2][B][rec_cap]
    RayRec* rays = nullptr;          // [2][B][ray_cap]
    uint32_t* cmasks = nullptr;      // [2][B][cube_cells][32]
    uint8_t* ctouched = nullptr;     // [2][B][cube_cells]
    uint32_t* mail = nullptr;        // pinned: [2][B][CNT_COUNT] counters + [2] pool cursor snapshots
    // pool cursor of the tree as of a point of the context stream (after a batch's applies), read back asynchronously: keeps
    // the host's upper bound of the cursor tight without ever waiting for it
    cudaEvent_t snap_ev[2] = {nullptr, nullptr};
    uint64_t snap_after[2] = {0, 0};   // records of the applies queued after the snapshot
    bool snap_valid[2] = {false, false};
    cudaStream_t rc_stream[2] = {nullptr, nullptr};
    cudaEvent_t cast_done[2] = {nullptr, nullptr}, done[2] = {nullptr, nullptr}, start = nullptr, t0 = nullptr, t1 = nullptr;
    cudaEvent_t start_ev = nullptr;  // what the slots' streams wait for before a call's first batches: `start`, or the caller's gate
    int walk_blocks_per_sm = 0;
    int overlap = 1;
    bool dirty = false;              // a batch was abandoned half-way: cubes must be cleared before the next use
};

void scan_pipe_destroy(r3d_ctx* ctx) {
    ScanPipe* p = ctx->scan_pipe;
    if (!p) return;
    cudaFree(p->geom); cudaFree(p->counters); cudaFree(p->lists); cudaFree(p->recs); cudaFree(p->rays); cudaFree(p->cmasks); cudaFree(p->ctouched);
    if (p->mail) cudaFreeHost(p->mail);
    for (int i = 0; i < 2; ++i) {
        if (p->rc_stream[i]) cudaStreamDestroy(p->rc_stream[i]);
        if (p->cast_done[i]) cudaEventDestroy(p->cast_done[i]);
        if (p->done[i]) cudaEventDestroy(p->done[i]);
        if (p->snap_ev[i]) cudaEventDestroy(p->snap_ev[i]);
    }
    if (p->start) cudaEventDestroy(p->start);
    if (p->t0) cudaEventDestroy(p->t0);
    if (p->t1) cudaEventDestroy(p->t1);
    delete p;
    ctx->scan_pipe = nullptr;
}

static int pipe_sync_all(r3d_ctx* ctx, ScanPipe* p) {
    R3D_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < 2; ++i)
        if (p->rc_stream[i]) R3D_CUDA_OK(ctx, cudaStreamSynchronize(p->rc_stream[i]));
    R3D_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));   // work the side streams handed back
    return R3D_OK;
}

template <typename T>
static int pipe_realloc(r3d_ctx* ctx, T** ptr, size_t count, const char* what) {
    cudaFree(*ptr);
    *ptr = nullptr;
    cudaError_t e = cudaMalloc(ptr, count * sizeof(T));
    if (e != cudaSuccess) {
        cudaGetLastError();
        *ptr = nullptr;
        return set_error(ctx, R3D_ERR_OOM, "cudaMalloc(%zu MB, %s) failed: %s", count * sizeof(T) >> 20, what, cudaGetErrorString(e));
    }
    return R3D_OK;
}

// Context-level, created on first use and kept: a fresh tree allocates nothing on its way through the pipeline.
static int pipe_get(r3d_ctx* ctx, ScanPipe** out) {
    if (!ctx->scan_pipe) {
        ScanPipe* p = new ScanPipe();
        ctx->scan_pipe = p;
        for (int i = 0; i < 2; ++i) {
            // the ray casts run on side streams of the lowest priority: the short list / emit / apply kernels of the batch
            // before (context stream, highest priority) run beside them, not behind them
            R3D_CUDA_OK(ctx, cudaStreamCreateWithFlags(&p->rc_stream[i], cudaStreamNonBlocking));
            R3D_CUDA_OK(ctx, cudaEventCreateWithFlags(&p->cast_done[i], cudaEventDisableTiming));
            R3D_CUDA_OK(ctx, cudaEventCreateWithFlags(&p->done[i], cudaEventDisableTiming));
            R3D_CUDA_OK(ctx, cudaEventCreateWithFlags(&p->snap_ev[i], cudaEventDisableTiming));
        }
        R3D_CUDA_OK(ctx, cudaEventCreateWithFlags(&p->start, cudaEventDisableTiming));
        R3D_CUDA_OK(ctx, cudaEventCreate(&p->t0));
        R3D_CUDA_OK(ctx, cudaEventCreate(&p->t1));
        R3D_CUDA_OK(ctx, cudaHostAlloc((void**)&p->mail, (2 * K3_MAX_BATCH * CNT_COUNT + 8) * sizeof(uint32_t), cudaHostAllocDefault));
        if (const char* v = getenv("R3D_PIPE_OVERLAP")) p->overlap = atoi(v) != 0;
        int per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_scan_walk, K3_THREADS, 0);
        p->walk_blocks_per_sm = per_sm < 1 ? 1 : per_sm;
    }
    *out = ctx->scan_pipe;
    return R3D_OK;
}

// (Re)shape the pipeline's buffers.  Grow-only per dimension; changing B or the cube size re-lays the cubes out.
static int pipe_reserve(r3d_ctx* ctx, ScanPipe* p, int B, uint64_t cube_cells, uint64_t rec_cap, uint64_t ray_cap, bool* shaped) {
    cube_cells = (cube_cells + 3) / 4 * 4;
    *shaped = false;
    if (rec_cap < p->rec_cap) rec_cap = p->rec_cap;
    if (ray_cap < p->ray_cap) ray_cap = p->ray_cap;
    const bool same_cubes = B == p->B && cube_cells == p->cube_cells;
    if (same_cubes && rec_cap == p->rec_cap && ray_cap == p->ray_cap && !p->dirty) return R3D_OK;
    *shaped = true;
    R3D_TRY(pipe_sync_all(ctx, p));
    if (B != p->B) {
        R3D_TRY(pipe_realloc(ctx, &p->geom, (size_t)2 * B * 8, "scan geometry"));
        R3D_TRY(pipe_realloc(ctx, &p->counters, (size_t)2 * B * CNT_COUNT, "scan counters"));
    }
    if (B != p->B || rec_cap != p->rec_cap) {
        R3D_TRY(pipe_realloc(ctx, &p->lists, (size_t)2 * B * rec_cap, "cell lists"));
        R3D_TRY(pipe_realloc(ctx, &p->recs, (size_t)2 * B * rec_cap, "delta records"));
    }
    if (B != p->B || ray_cap != p->ray_cap) R3D_TRY(pipe_realloc(ctx, &p->rays, (size_t)2 * B * ray_cap, "ray records"));
    if (!same_cubes) {
        p->B = 0; p->cube_cells = 0;
        R3D_TRY(pipe_realloc(ctx, &p->cmasks, (size_t)2 * B * cube_cells * 32, "scan cubes"));
        R3D_TRY(pipe_realloc(ctx, &p->ctouched, (size_t)2 * B * cube_cells, "scan cube byte maps"));
        p->dirty = true;
    }
    if (p->dirty) {   // fresh memory, or a batch that was abandoned half-way
        R3D_CUDA_OK(ctx, cudaMemsetAsync(p->cmasks, 0, (size_t)2 * B * cube_cells * 128, ctx->stream));
        R3D_CUDA_OK(ctx, cudaMemsetAsync(p->ctouched, 0, (size_t)2 * B * cube_cells, ctx->stream));
        p->dirty = false;
    }
    p->B = B; p->cube_cells = cube_cells; p->rec_cap = rec_cap; p->ray_cap = ray_cap;
    return R3D_OK;
}

// largest batch size (<= want) whose two slots of cubes fit the scratch budget; 0 when not even one cube per slot fits
static int pipe_fit_batch(r3d_ctx* ctx, int want, uint64_t cube_cells) {
    cube_cells = (cube_cells + 3) / 4 * 4;
    for (int B = want; B >= 1; --B)
        if ((uint64_t)2 * B * cube_cells * 129ull <= ctx->cell_budget_bytes) return B;
    return 0;
}

struct Batch {
    uint32_t first = 0, count = 0;
    bool timed = false;
};

static int pipe_enqueue(r3d_tree* t, ScanPipe* p, const float* d_xyz, const uint64_t* offsets, const uint64_t* n_points, const float* origins,
                        double maxrange, const Batch& b, int slot) {
    r3d_ctx* ctx = t->ctx;
    BatchArgs a;
    memset(&a, 0, sizeof a);
    uint32_t n_max = 0;
    for (uint32_t j = 0; j < b.count; ++j) {
        const uint32_t s = b.first + j;
        a.scan[j].xyz = d_xyz + offsets[s] * 3;
        a.scan[j].n = (uint32_t)n_points[s];
        a.scan[j].ox = origins[3 * s]; a.scan[j].oy = origins[3 * s + 1]; a.scan[j].oz = origins[3 * s + 2];
        if (a.scan[j].n > n_max) n_max = a.scan[j].n;
    }
    a.n_scans = (int)b.count;
    a.maxrange = maxrange; a.res = t->res; a.res_factor = t->res_factor;
    a.geom = p->geom + (size_t)slot * p->B * 8;
    a.counters = p->counters + (size_t)slot * p->B * CNT_COUNT;
    a.rays = p->rays + (size_t)slot * p->B * p->ray_cap;
    a.ray_stride = p->ray_cap;
    a.cmasks = p->cmasks + (size_t)slot * p->B * p->cube_cells * 32;
    a.ctouched = p->ctouched + (size_t)slot * p->B * p->cube_cells;
    a.cube_cells = (uint32_t)p->cube_cells;
    a.lists = p->lists + (size_t)slot * p->B * p->rec_cap;
    a.recs = p->recs + (size_t)slot * p->B * p->rec_cap;
    a.rec_cap = (uint32_t)p->rec_cap;
    cudaStream_t rs = p->overlap ? p->rc_stream[slot] : ctx->stream;
    if (p->overlap) {
        // the slot's stream starts after the call's set-up and after the slot's previous batch has been emitted and its
        // counters read back (that batch used the same cubes, counters, ray records and record buffers)
        R3D_CUDA_OK(ctx, cudaStreamWaitEvent(rs, p->start_ev, 0));
        R3D_CUDA_OK(ctx, cudaStreamWaitEvent(rs, p->done[slot], 0));
    }
    R3D_CUDA_OK(ctx, cudaMemsetAsync(a.geom, 0x7f, (size_t)b.count * 8 * sizeof(int), rs));
    R3D_CUDA_OK(ctx, cudaMemsetAsync(a.counters, 0, (size_t)b.count * CNT_COUNT * sizeof(uint32_t), rs));
    {
        unsigned gx = (n_max + 255u) / 256u;
        const unsigned cap = (unsigned)ctx->sm_count * 8u;
        if (gx > cap) gx = cap;
        if (gx < 1u) gx = 1u;      // an empty scan still has a geometry (its origin)
        k_scan_prepare<<<dim3(gx, b.count), 256, 0, rs>>>(a);
        if (b.timed) cudaEventRecord(p->t0, rs);
        k_scan_walk<<<(unsigned)(ctx->sm_count * p->walk_blocks_per_sm), K3_THREADS, 0, rs>>>(a);
        if (b.timed) cudaEventRecord(p->t1, rs);
        ctx->launches += 2;
    }
    if (p->overlap) {
        R3D_CUDA_OK(ctx, cudaEventRecord(p->cast_done[slot], rs));
        R3D_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->stream, p->cast_done[slot], 0));
    }
    k_cells_list<<<dim3((unsigned)ctx->sm_count * 2u, b.count), 256, 0, ctx->stream>>>(a);
    k_cells_emit<<<dim3((unsigned)ctx->sm_count * 4u, b.count), 256, 0, ctx->stream>>>(a);
    ctx->launches += 2;
    R3D_CUDA_OK(ctx, cudaGetLastError());
    uint32_t* mail = p->mail + (size_t)slot * K3_MAX_BATCH * CNT_COUNT;
    R3D_CUDA_OK(ctx, cudaMemcpyAsync(mail, a.counters, (size_t)b.count * CNT_COUNT * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    R3D_CUDA_OK(ctx, cudaEventRecord(p->done[slot], ctx->stream));
    return R3D_OK;
}

static bool origin_has_key(const r3d_tree* t, const float* o) {
    for (int i = 0; i < 3; ++i) {
        const double f = floor(t->res_factor * (double)o[i]);
        if (!(f >= -32768.0 && f < 32768.0)) return false;
    }
    return true;
}

int dense_scans_run(r3d_tree* t, const float* d_xyz, const uint64_t* n_points, const float* origins, uint32_t n_scans, double maxrange,
                    ScanSink* sink, uint32_t* done_out, uint64_t* rays_out, uint64_t* steps_out) {
    r3d_ctx* ctx = t->ctx;
    *done_out = 0;
    if (n_scans == 0 || !(maxrange >= 0.0)) return R3D_OK;
    timespec ts_enter;
    clock_gettime(CLOCK_MONOTONIC, &ts_enter);
    // scans the pipeline can take: a bounded range (so the cube is bounded) and an origin inside the key space
    uint32_t n_ok = 0;
    uint64_t n_max = 0;
    std::vector<uint64_t> offsets(n_scans + 1, 0);
    for (uint32_t s = 0; s < n_scans; ++s) offsets[s + 1] = offsets[s] + n_points[s];
    for (; n_ok < n_scans; ++n_ok) {
        if (n_points[n_ok] > 0xfffffff0ull || !origin_has_key(t, origins + 3 * (size_t)n_ok)) break;
        if (n_points[n_ok] > n_max) n_max = n_points[n_ok];
    }
    if (n_ok == 0) return R3D_OK;
    ScanPipe* p = nullptr;
    R3D_TRY(pipe_get(ctx, &p));
    int cfg_B = 8;
    if (const char* v = getenv("R3D_SCAN_BATCH")) { const int b = atoi(v); if (b >= 1 && b <= K3_MAX_BATCH) cfg_B = b; }
    int want_B = (uint32_t)cfg_B > n_ok ? (int)n_ok : cfg_B;
    // first guess of the cube: 2^20 cells (128 MB) or the whole (2 reach + 1)^3 cube when that is smaller; a scan that
    // needs more reports it and the pipeline is re-shaped
    uint64_t cube = p->cube_cells ? p->cube_cells : (1ull << 20);
    {
        const double reach = ceil(maxrange * t->res_factor / 8.0) + 3.0;
        const double full = (2.0 * reach + 1.0) * (2.0 * reach + 1.0) * (2.0 * reach + 1.0);
        if (!p->cube_cells && full < (double)cube) cube = (uint64_t)full;
    }
    // `lay` cubes per slot are laid out (kept from call to call: a single-scan call after a batch call re-shapes nothing),
    // `B` of them are used per batch
    int lay = pipe_fit_batch(ctx, p->B > cfg_B ? p->B : cfg_B, cube);
    if (lay == 0) return R3D_OK;   // scratch budget too small for direct mapping: hash path
    int B = want_B < lay ? want_B : lay;
    uint64_t rec_cap = p->rec_cap ? p->rec_cap : (1ull << 16);
    uint64_t ray_cap = (n_max + 255) / 256 * 256;
    auto now_ns = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec; };
    static const bool trace = getenv("R3D_PIPE_TRACE") != nullptr;   // host-side turnarounds above 1 ms, to stderr
    t->pipe_wait_ns = t->pipe_work_ns = t->pipe_max_turn_ns = t->pipe_scans = 0;
    uint32_t next = 0;          // first scan not yet consumed
    float kernel_ms = 0.f;
    uint64_t pre_ns = 0;
    for (int attempt = 0; attempt < 24 && next < n_ok; ++attempt) {
        bool shaped = false;
        {
            const uint64_t t_r = trace ? now_ns() : 0;
            R3D_TRY(pipe_reserve(ctx, p, lay, cube, rec_cap, ray_cap, &shaped));
            if (trace && now_ns() - t_r > 1000000ull)
                fprintf(stderr, "[r3d pipe] shaping the pipeline took %.2f ms (%d cubes per slot of %llu cells, %llu records, %llu rays per scan)\n", (now_ns() - t_r) * 1e-6,
                        lay, (unsigned long long)cube, (unsigned long long)rec_cap, (unsigned long long)ray_cap);
        }
        if (p->overlap) {
            // A caller that queued work the ray casting does not depend on (the multi-GPU merge's sorted apply of the round
            // before, r3d_scan_deltas_compute) passes the point of the context stream BEFORE that work as a gate: the first
            // batches then start beside it instead of behind it.  Only when nothing was re-shaped and both slots are idle.
            const bool gated = t->cast_gate_armed && attempt == 0 && !shaped && cudaEventQuery(p->done[0]) == cudaSuccess &&
                               cudaEventQuery(p->done[1]) == cudaSuccess;
            t->cast_gate_armed = false;     // one use: whatever this call queues later (staging copies of host scans) is not before the gate
            if (gated) {
                p->start_ev = t->cast_gate;
            } else {
                R3D_CUDA_OK(ctx, cudaEventRecord(p->start, ctx->stream));   // buffers shaped, cubes clear, scans resident
                // nothing of an earlier call is pending on the slots
                R3D_CUDA_OK(ctx, cudaEventRecord(p->done[0], ctx->stream));
                R3D_CUDA_OK(ctx, cudaEventRecord(p->done[1], ctx->stream));
                p->start_ev = p->start;
            }
        }
        p->dirty = true;        // until every queued batch has been emitted
        uint64_t t_mark = now_ns();
        if (trace && attempt == 0) pre_ns = t_mark - ((uint64_t)ts_enter.tv_sec * 1000000000ull + (uint64_t)ts_enter.tv_nsec);
        Batch cur, nxt;
        cur.first = next; cur.count = (uint32_t)B < n_ok - next ? (uint32_t)B : n_ok - next;
        cur.timed = cur.first + cur.count == n_ok;
        int slot = 0;
        R3D_TRY(pipe_enqueue(t, p, d_xyz, offsets.data(), n_points, origins, maxrange, cur, slot));
        p->snap_valid[0] = p->snap_valid[1] = false;
        bool reshape = false;
        while (cur.count) {
            nxt.first = cur.first + cur.count;
            nxt.count = nxt.first < n_ok ? ((uint32_t)B < n_ok - nxt.first ? (uint32_t)B : n_ok - nxt.first) : 0u;
            nxt.timed = nxt.count && nxt.first + nxt.count == n_ok;
            if (nxt.count) R3D_TRY(pipe_enqueue(t, p, d_xyz, offsets.data(), n_points, origins, maxrange, nxt, slot ^ 1));
            const uint64_t t_wait = now_ns();
            R3D_CUDA_OK(ctx, cudaEventSynchronize(p->done[slot]));
            const uint64_t t_got = now_ns();
            t->pipe_work_ns += t_wait - t_mark;
            if (t_wait - t_mark > t->pipe_max_turn_ns) t->pipe_max_turn_ns = t_wait - t_mark;
            t->pipe_wait_ns += t_got - t_wait;
            t_mark = t_got;
            const uint32_t* mail = p->mail + (size_t)slot * K3_MAX_BATCH * CNT_COUNT;
            if (cur.timed) cudaEventElapsedTime(&kernel_ms, p->t0, p->t1), kernel_ms /= (float)cur.count;
            if (sink->mode == ScanSink::APPLY) {
                // the newest cursor snapshot that has arrived (the applies of the batch before last finished long ago)
                for (int k = 0; k < 2; ++k) {
                    const int q = slot ^ 1 ^ k;   // the other slot's snapshot is the newer one
                    if (p->snap_valid[q] && cudaEventQuery(p->snap_ev[q]) == cudaSuccess) {
                        const uint64_t tight = (uint64_t)p->mail[2 * K3_MAX_BATCH * CNT_COUNT + q] + p->snap_after[q];
                        if (tight < t->pool_bound) t->pool_bound = tight;
                        break;
                    }
                }
            }
            if (mail[CNT_GRID_MISS] && !mail[CNT_GRID_NEED]) return set_error(ctx, R3D_ERR_STATE, "a ray left its scan's cube (internal sizing error)");
            for (uint32_t j = 0; j < cur.count; ++j) {
                const uint32_t* c = mail + (size_t)j * CNT_COUNT;
                const uint32_t s = cur.first + j;
                if (c[CNT_GRID_MISS] && c[CNT_GRID_NEED]) {
                    // the scan's cube does not fit: re-shape the pipeline for it, or hand the scan to the hash path
                    R3D_TRY(pipe_sync_all(ctx, p));
                    const uint64_t need = c[CNT_GRID_NEED] == 0xffffffffu ? 0 : (uint64_t)c[CNT_GRID_NEED] + c[CNT_GRID_NEED] / 8;
                    const int nb = need ? pipe_fit_batch(ctx, cfg_B, need) : 0;
                    if (nb == 0) { *done_out = s; ctx->last_kernel_ms = kernel_ms; return R3D_OK; }
                    cube = need; lay = nb; B = want_B < lay ? want_B : lay;
                    reshape = true;
                    break;
                }
                if (c[CNT_DELTA] > p->rec_cap) {
                    R3D_TRY(pipe_sync_all(ctx, p));
                    rec_cap = (uint64_t)c[CNT_DELTA] * 2;
                    reshape = true;
                    break;
                }
                const DeltaRecord* recs = p->recs + ((size_t)slot * p->B + j) * p->rec_cap;
                const uint64_t n_rec = c[CNT_DELTA];
                if (sink->mode == ScanSink::APPLY) {
                    const uint64_t t_a = trace ? now_ns() : 0;
                    const uint64_t cap0 = t->pool_cap, tcap0 = t->tcap;
                    R3D_TRY(apply_delta_impl(t, recs, n_rec));
                    if (trace && now_ns() - t_a > 1000000ull)
                        fprintf(stderr, "[r3d pipe] apply of scan %u took %.2f ms on the host (pool %llu -> %llu bricks, table %llu -> %llu, bound %llu)\n", s,
                                (now_ns() - t_a) * 1e-6, (unsigned long long)cap0, (unsigned long long)t->pool_cap, (unsigned long long)tcap0,
                                (unsigned long long)t->tcap, (unsigned long long)t->pool_bound);
                    p->snap_after[0] += n_rec; p->snap_after[1] += n_rec;
                } else if (sink->mode == ScanSink::EXPORT_USER) {
                    sink->counts[s] = n_rec;
                    if (sink->used + n_rec > sink->capacity) {
                        R3D_TRY(pipe_sync_all(ctx, p));
                        for (uint32_t r = s + 1; r < n_scans; ++r) sink->counts[r] = 0;
                        return set_error(ctx, R3D_ERR_OOM, "record buffer holds %llu records, scan %u needs %llu in total so far",
                                         (unsigned long long)sink->capacity, s, (unsigned long long)(sink->used + n_rec));
                    }
                    if (n_rec)
                        R3D_CUDA_OK(ctx, cudaMemcpyAsync((char*)sink->records + sink->used * sizeof(DeltaRecord), recs, n_rec * sizeof(DeltaRecord), cudaMemcpyDefault, ctx->stream));
                    sink->used += n_rec;
                } else {
                    if (n_rec > t->delta_cap) R3D_TRY(tree_reserve_delta(t, n_rec + n_rec / 4 + 1024));
                    if (n_rec) R3D_CUDA_OK(ctx, cudaMemcpyAsync(t->delta, recs, n_rec * sizeof(DeltaRecord), cudaMemcpyDeviceToDevice, ctx->stream));
                }
                // (applied records stay in the pipeline's buffers: a batch insert leaves no delta to export -- copying the
                // last one out cost a fresh tree a synchronising allocation at the end of every call)
                t->delta_n = sink->mode == ScanSink::APPLY ? 0 : n_rec;
                t->last_batch_records = n_rec;
                *rays_out += n_points[s];
                next = s + 1;
                t->pipe_scans = next;
            }
            if (reshape) break;
            if (sink->mode == ScanSink::APPLY) {
                R3D_CUDA_OK(ctx, cudaMemcpyAsync(p->mail + 2 * K3_MAX_BATCH * CNT_COUNT + slot, t->counters + CNT_POOL_USED, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
                R3D_CUDA_OK(ctx, cudaEventRecord(p->snap_ev[slot], ctx->stream));
                p->snap_after[slot] = 0;
                p->snap_valid[slot] = true;
            }
            *steps_out += (uint64_t)mail[CNT_STEPS_LO] | ((uint64_t)mail[CNT_STEPS_HI] << 32);
            cur = nxt;
            slot ^= 1;
        }
        if (!reshape) p->dirty = false;
    }
    if (next < n_ok) return set_error(ctx, R3D_ERR_OOM, "scan pipeline could not be shaped for scan %u", next);
    *done_out = next;
    ctx->last_kernel_ms = kernel_ms;
    if (trace)
        fprintf(stderr, "[r3d pipe] %u scans: %.2f ms in the call (%.2f before the first batch), %.2f waiting for batches, %.2f of host work between them, longest turnaround %.2f\n", next,
                (now_ns() - ((uint64_t)ts_enter.tv_sec * 1000000000ull + (uint64_t)ts_enter.tv_nsec)) * 1e-6, pre_ns * 1e-6, t->pipe_wait_ns * 1e-6, t->pipe_work_ns * 1e-6,
                t->pipe_max_turn_ns * 1e-6);
    return R3D_OK;
}

}  // namespace r3d
