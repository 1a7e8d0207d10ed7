// r3d_octree.cu -- occupancy half of the path: OcTree(res), updateNode, insertPointCloud, writeBinary.
//
// Replaces the un-vendored `octomap` extension the reference imports (octomap/txt_transfer_octomap.py:2,25,33-36;
// octomap/ply_transfer_octomap.py:2,33,45-48).  Data layout in HBM:
//
//   * the map is a flat voxel store, not a pointer tree: a depth-16 OcTreeKey (3 x uint16) is split into a BRICK
//     (key >> 3 per axis: the depth-13 node) and a 9-bit Morton voxel index inside it.  Bricks live in a pool
//     (512 float32 log-odds + 512 "known" bits each) addressed through an open-addressing hash table
//     brick key -> pool index.  Leaf log-odds of the flat store equal the leaf log-odds of upstream's tree
//     (update-time pruning / expansion never changes a leaf value); the tree shape needed for .bt and size() is
//     derived from the values on demand (r3d_bt.cu).
//   * one scan's update (insertPointCloud) is a DELTA: per touched brick a 512-bit occupied mask and a 512-bit free
//     mask, filled by the ray-casting kernel (K3) with red.or.  With a bounded range the masks are direct-mapped (a
//     cube of brick cells around the sensor origin, context-level scratch, byte map of touched cells); with an
//     unbounded one they sit in a per-scan hash table.  Either way they are read back as 136-byte records and applied
//     to the store by the clamped log-odds kernel (K4).  Records are what multi-GPU runs exchange.
#include <vector>

#include <time.h>

#include "r3d_octree.cuh"

namespace r3d {

// ------------------------------------------------------------------ hash table primitives
__device__ __forceinline__ uint64_t ld_cg_u64(const uint64_t* p) { return __ldcg(reinterpret_cast<const unsigned long long*>(p)); }

// find-or-insert of a brick key; returns the slot, or kNoSlot when the table is full
__device__ __forceinline__ uint64_t table_find_or_insert(uint64_t* keys, uint64_t cap, uint64_t bk, bool& inserted) {
    const uint64_t mask = cap - 1;
    uint64_t slot = hash64(bk) & mask;
    inserted = false;
    for (uint64_t probe = 0; probe < cap; ++probe) {
        const uint64_t k = ld_cg_u64(keys + slot);
        if (k == bk) return slot;
        if (k == kEmptyKey) {
            const unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(keys + slot), kEmptyKey, bk);
            if (old == kEmptyKey) { inserted = true; return slot; }
            if (old == bk) return slot;
        }
        slot = (slot + 1) & mask;
    }
    return kNoSlot;
}
__device__ __forceinline__ uint64_t table_find(const uint64_t* keys, uint64_t cap, uint64_t bk) {
    const uint64_t mask = cap - 1;
    uint64_t slot = hash64(bk) & mask;
    for (uint64_t probe = 0; probe < cap; ++probe) {
        const uint64_t k = ld_cg_u64(keys + slot);
        if (k == bk) return slot;
        if (k == kEmptyKey) return kNoSlot;
        slot = (slot + 1) & mask;
    }
    return kNoSlot;
}

template <typename T>
__device__ __forceinline__ void load_point(const T* xyz, unsigned long long i, float& x, float& y, float& z) {
    x = (float)xyz[3 * i]; y = (float)xyz[3 * i + 1]; z = (float)xyz[3 * i + 2];   // binding: point3d(float) cast
}

// ------------------------------------------------------------------ updateNode(point, ...) batches (a10)
// pass 1: make sure every touched brick exists in the table and has a pool index
template <typename T>
__global__ void k_points_ensure(const T* __restrict__ xyz, unsigned long long n, double res_factor, uint64_t* tkeys,
                                uint32_t* tvals, uint64_t tcap, uint32_t* counters) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float x, y, z;
        load_point(xyz, i, x, y, z);
        uint16_t kx, ky, kz;
        if (!coord_to_key3(res_factor, x, y, z, kx, ky, kz)) { atomicAdd(&counters[CNT_DROPPED], 1u); continue; }
        bool inserted;
        const uint64_t slot = table_find_or_insert(tkeys, tcap, brick_key(kx, ky, kz), inserted);
        if (slot == kNoSlot) { counters[CNT_OVERFLOW] = 1; continue; }
        if (inserted) tvals[slot] = atomicAdd(&counters[CNT_POOL_USED], 1u);
    }
}

// pass 2: clamped log-odds update; equal (brick, voxel) targets inside a warp are merged so that one lane applies
// the update count times (the update function is the same for every point of the batch, so order is immaterial)
template <typename T>
__global__ void k_points_update(const T* __restrict__ xyz, unsigned long long n, double res_factor, const uint64_t* tkeys,
                                const uint32_t* tvals, uint64_t tcap, float* values, uint32_t* known, float upd, float cmin,
                                float cmax) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned long long n_pad = (n + 31ull) & ~31ull;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += stride) {
        unsigned long long target = 0xffffffffffffff00ull + lane;   // unique per lane when invalid
        bool valid = false;
        if (i < n) {
            float x, y, z;
            load_point(xyz, i, x, y, z);
            uint16_t kx, ky, kz;
            if (coord_to_key3(res_factor, x, y, z, kx, ky, kz)) {
                const uint64_t slot = table_find(tkeys, tcap, brick_key(kx, ky, kz));
                if (slot != kNoSlot) {
                    target = (unsigned long long)tvals[slot] * kBrickVoxels + brick_voxel_index(kx, ky, kz);
                    valid = true;
                }
            }
        }
        const unsigned peers = __match_any_sync(0xffffffffu, target);
        if (valid && lane == (unsigned)(__ffs(peers) - 1)) {
            const int count = __popc(peers);
            int* p = reinterpret_cast<int*>(values + target);
            int old = *reinterpret_cast<volatile int*>(p);
            for (;;) {
                float v = __int_as_float(old);
                for (int c = 0; c < count; ++c) {
                    const float nv = clamped_add(v, upd, cmin, cmax);
                    if (nv == v) break;   // saturated (upstream's early abort) or zero update
                    v = nv;
                }
                if (__float_as_int(v) == old) break;
                const int prev = atomicCAS(p, old, __float_as_int(v));
                if (prev == old) break;
                old = prev;
            }
            const unsigned vox = (unsigned)(target % kBrickVoxels);
            uint32_t* kw = known + (target / kBrickVoxels) * 16 + (vox >> 5);
            const uint32_t bit = 1u << (vox & 31u);
            if (!(*reinterpret_cast<volatile uint32_t*>(kw) & bit)) atomicOr(kw, bit);
        }
    }
}

// ------------------------------------------------------------------ K3: per-scan ray casting into the scan delta
struct ScanArgs {
    const float* xyz;
    unsigned long long n;
    float ox, oy, oz;
    double maxrange, res, res_factor;
    // hash mode (unbounded reach): open-addressing table of brick keys; the masks of a brick sit at its table position
    uint64_t* skeys;
    uint32_t* smasks;   // per slot: 16 words occupied, 16 words free
    uint64_t scap;
    uint32_t* counters;
    // dense mode (bounded reach): the masks are direct-mapped -- cell (cx, cy, cz) of a gdim^3 grid of bricks centred on
    // the sensor origin owns words [32 * cell, 32 * cell + 32) of `cmasks`; `ctouched` has one bit per cell that holds
    // anything, so that only touched cells are read back.  No lookup structure, no allocation, no dependent loads.
    uint32_t* cmasks;
    uint32_t* ctouched;
    int gx0, gy0, gz0;  // brick coordinates of cell (0,0,0)
    uint32_t gdim;      // cells per axis
    uint32_t gcells;    // gdim^3
};

// computeUpdate (OccupancyOcTreeBase): per point, free cells along the ray, endpoint occupied when in range.
//
// Persistent warps; every lane walks one ray at a time and idle lanes are re-filled from a global ray counter as soon
// as K3_REFILL_MIN of them are idle (rays of one image differ in length by orders of magnitude: sky pixels walk 0
// cells, far ground pixels ~1000).  The walk is the branch-free form of computeRayKeys (r3d_math.cuh).  Free cells are
// collected in a 64-bit register mask per 4x4x4 sub-block (the depth-14 node: 64 consecutive Morton voxels = one
// aligned 64-bit word of the brick's free mask) and written with ONE red.or when the ray leaves the sub-block; the
// word's current value is fetched (L2) when the ray enters the sub-block and only consulted when it leaves, so the load
// latency overlaps the walk, and the atomic is skipped when every bit is already set (128 M visits -> 10 M distinct
// cells per scan).  Different lanes cross sub-block and brick borders at different steps, so that bookkeeping is
// written without divergent slow paths: some lane needs it on almost every iteration.
constexpr int K3_THREADS = 256;
#ifdef K3_REFILL_MIN_OVERRIDE
constexpr int K3_REFILL_MIN = K3_REFILL_MIN_OVERRIDE;
#else
constexpr int K3_REFILL_MIN = 8;
#endif

__device__ __forceinline__ uint64_t ldcg_u64_if(const uint64_t* p, uint64_t otherwise, bool pred) {
    uint64_t v = otherwise;
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q ld.global.ca.u64 %0, [%1];\n\t}" : "+l"(v) : "l"(p), "r"((unsigned)pred) : "memory");
    return v;
}
__device__ __forceinline__ unsigned sub_index(int kx, int ky, int kz) {
    return (unsigned)((kx >> 2) & 1) | ((unsigned)((ky >> 2) & 1) << 1) | ((unsigned)((kz >> 2) & 1) << 2);
}
__device__ __forceinline__ uint64_t sub_bit(int kx, int ky, int kz) {
    const unsigned x = kx & 3, y = ky & 3, z = kz & 3;
    const unsigned bit = (x & 1u) | ((y & 1u) << 1) | ((z & 1u) << 2) | ((x & 2u) << 2) | ((y & 2u) << 3) | ((z & 2u) << 4);
    return 1ull << bit;
}

// ---- dense mode
__device__ __forceinline__ bool cell_of_key(const ScanArgs& a, int kx, int ky, int kz, uint32_t& cell) {
    const uint32_t ux = (uint32_t)((kx >> 3) - a.gx0), uy = (uint32_t)((ky >> 3) - a.gy0), uz = (uint32_t)((kz >> 3) - a.gz0);
    cell = (uz * a.gdim + uy) * a.gdim + ux;
    return (ux < a.gdim) & (uy < a.gdim) & (uz < a.gdim);
}

#ifndef K3_VARIANT
#define K3_VARIANT 1
#endif
// key += step and tMax += tDelta (round to nearest, no contraction) when axis == which, as predicated instructions: written
// as C++ conditionals the compiler turns the three mutually exclusive updates into a chain of divergent branches
__device__ __forceinline__ void step_axis_if(int axis, int which, int& k, int s, double& tm, double td) {
    asm("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %2, %3;\n\t@p add.s32 %0, %0, %4;\n\t@p add.rn.f64 %1, %1, %5;\n\t}"
        : "+r"(k), "+d"(tm) : "r"(axis), "r"(which), "r"(s), "d"(td));
}
// Per-lane state of the dense walker.
struct DenseLane {
    int kx, ky, kz, ex, ey, ez, sx, sy, sz;
    double tmx, tmy, tmz, tdx, tdy, tdz, length;
    int csx, csy, csz;   // cell-index change of one brick step along each axis
    int axis;
    int mx, my, mz;      // (K3_VARIANT 1) the axis of the next step as 0/1 integers
    uint32_t cell;       // grid cell of the current brick
    uint32_t widx;       // index of the current sub-block's free word in the 64-bit view of cmasks
    uint64_t mask;       // cells of that sub-block visited by this ray
    uint64_t seen;       // what the word held when the ray entered the sub-block; 0 = not known (yet)
    unsigned age;        // iterations since the ray entered the sub-block (saturates at 3)
    unsigned steps;
};

// One iteration of the walk for an active lane.  `stage` is the register the word fetched on entering a sub-block
// lands in; the caller alternates between two of them, and a fetched word is moved into `seen` two iterations after
// it was requested.  The warp therefore never waits on the load it has just issued (a scoreboard wait is warp-wide:
// with ~28 active lanes some lane enters a sub-block on nearly every iteration), only on the one from two iterations
// ago.  A ray that leaves a sub-block earlier publishes its cells without knowing the word (a redundant red.or).
__device__ __forceinline__ bool dense_step(const ScanArgs& a, uint64_t* masks64, uint8_t* touched, DenseLane& L, uint64_t& stage) {
    if (L.age == 1) L.seen = stage;      // requested two iterations ago, into this same register
    // ---- one DDA step along the chosen axis (ray_advance), then pick the next axis (ray_select).
#if K3_VARIANT == 2
    // (candidate, not measured yet: built with -DK3_VARIANT=2)  The axis is carried as an integer; the three updates are
    // written as conditional statements so that they compile to predicated DADD / IADD (no selects, no multiplies), and
    // "min(tMax) > length" is evaluated as "every tMax > length": three chained compares on the fp64 pipe instead of a
    // select tree on the ALU pipe (the minimum itself is never needed).
    const bool ax = L.axis == 0, ay = L.axis == 1;
    const int okx = L.kx, oky = L.ky, okz = L.kz;
    step_axis_if(L.axis, 0, L.kx, L.sx, L.tmx, L.tdx);
    step_axis_if(L.axis, 1, L.ky, L.sy, L.tmy, L.tdy);
    step_axis_if(L.axis, 2, L.kz, L.sz, L.tmz, L.tdz);
    const int cstep = ax ? L.csx : (ay ? L.csy : L.csz);
    const bool xy = L.tmx < L.tmy, xz = L.tmx < L.tmz, yz = L.tmy < L.tmz;
    const bool selx = xy & xz, sely = (!xy) & yz;
    L.axis = selx ? 0 : (sely ? 1 : 2);
    const bool past = (L.tmx > L.length) & (L.tmy > L.length) & (L.tmz > L.length);
    const bool done = (((L.kx ^ L.ex) | (L.ky ^ L.ey) | (L.kz ^ L.ez)) == 0) | past;
    const int diff = (L.kx ^ okx) | (L.ky ^ oky) | (L.kz ^ okz);
#elif K3_VARIANT & 1
    // The axis is carried as three 0/1 integers and applied by multiplication: selects run on the integer / select pipe,
    // which bounds this kernel, multiplies on the FMA and fp64 pipes.  x * 1.0 and t + 0.0 are exact, so the tMax update
    // tm + m * td is bit-identical to "tm + td on the chosen axis, untouched elsewhere".
    const int okx = L.kx, oky = L.ky, okz = L.kz;
    L.kx += L.mx * L.sx; L.ky += L.my * L.sy; L.kz += L.mz * L.sz;
    const int cstep = L.mx * L.csx + L.my * L.csy + L.mz * L.csz;
    L.tmx = dadd(L.tmx, dmul((double)L.mx, L.tdx));
    L.tmy = dadd(L.tmy, dmul((double)L.my, L.tdy));
    L.tmz = dadd(L.tmz, dmul((double)L.mz, L.tdz));
    const bool xy = L.tmx < L.tmy, xz = L.tmx < L.tmz, yz = L.tmy < L.tmz;
    const bool selx = xy & xz, sely = (!xy) & yz;
    const double tsel = selx ? L.tmx : (sely ? L.tmy : L.tmz);
    L.mx = selx ? 1 : 0; L.my = sely ? 1 : 0; L.mz = 1 - L.mx - L.my;
    const bool done = (((L.kx ^ L.ex) | (L.ky ^ L.ey) | (L.kz ^ L.ez)) == 0) | (tsel > L.length);
    const int diff = (L.kx ^ okx) | (L.ky ^ oky) | (L.kz ^ okz);
#else
    const bool ax = L.axis == 0, ay = L.axis == 1, az = L.axis == 2;
    const double nx = dadd(L.tmx, L.tdx), ny = dadd(L.tmy, L.tdy), nz = dadd(L.tmz, L.tdz);
    const int kold = ax ? L.kx : (ay ? L.ky : L.kz);
    const int knew = kold + (ax ? L.sx : (ay ? L.sy : L.sz));
    L.kx = ax ? knew : L.kx; L.ky = ay ? knew : L.ky; L.kz = az ? knew : L.kz;
    L.tmx = ax ? nx : L.tmx; L.tmy = ay ? ny : L.tmy; L.tmz = az ? nz : L.tmz;
    const bool xy = L.tmx < L.tmy, xz = L.tmx < L.tmz, yz = L.tmy < L.tmz;
    const bool selx = xy & xz, sely = (!xy) & yz;
    const double tsel = selx ? L.tmx : (sely ? L.tmy : L.tmz);
    const int cstep = ax ? L.csx : (ay ? L.csy : L.csz);
    L.axis = selx ? 0 : (sely ? 1 : 2);
    const bool done = (((L.kx ^ L.ex) | (L.ky ^ L.ey) | (L.kz ^ L.ez)) == 0) | (tsel > L.length);
    const int diff = knew ^ kold;            // only one coordinate moved
#endif
    const bool new_sub = (diff >> 2) != 0;
    // ---- leaving the sub-block (or the ray): publish its cells unless all of them are known to be set
    if ((done | new_sub) && (L.mask & ~L.seen) != 0) {
        atomicOr(reinterpret_cast<unsigned long long*>(masks64 + L.widx), (unsigned long long)L.mask);
        touched[L.widx >> 4] = 1;
    }
    // ---- entering the next one
    const bool enter = new_sub & !done;
    L.cell += (enter & ((diff >> 3) != 0)) ? (uint32_t)cstep : 0u;
    if (enter) {
        if (L.cell >= a.gcells) {          // memory-safety guard; the grid is sized so that it never trips
            a.counters[CNT_GRID_MISS] = 1;
            L.cell = 0;
        }
        L.widx = L.cell * 16u + 8u + sub_index(L.kx, L.ky, L.kz);
        L.mask = 0;
        L.seen = 0;
    }
    stage = ldcg_u64_if(masks64 + L.widx, stage, enter);
    L.age = enter ? 0u : (L.age < 3u ? L.age + 1u : 3u);
    L.mask |= sub_bit(L.kx, L.ky, L.kz);
    L.steps += done ? 0u : 1u;
    return !done;
}

__global__ void __launch_bounds__(K3_THREADS, 3) k_scan_raycast_dense(const ScanArgs a, unsigned long long* ray_counter, const uint32_t* abort_flag) {
    if (__ldcg(abort_flag)) return;   // an earlier scan of the pipeline is waiting for the host: leave the scratch alone
    const unsigned lane = threadIdx.x & 31u;
    uint64_t* const masks64 = reinterpret_cast<uint64_t*>(a.cmasks);
    uint8_t* const touched = reinterpret_cast<uint8_t*>(a.ctouched);
    bool active = false, exhausted = false;
    // (keys never wrap here: the host only picks this kernel when the whole grid lies inside the key range)
    DenseLane L;
    memset(&L, 0, sizeof L);
    L.age = 3;
    uint64_t stage_a = 0, stage_b = 0;
    for (;;) {
        const unsigned act = __ballot_sync(0xffffffffu, active);
        const unsigned idle = ~act;
        if (!exhausted && __popc(idle) >= K3_REFILL_MIN) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(ray_counter, (unsigned long long)__popc(idle));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base + __popc(idle) >= a.n) exhausted = true;
            const unsigned long long i = base + __popc(idle & ((1u << lane) - 1u));
            if (!active && i < a.n) {
                const float px = a.xyz[3 * i], py = a.xyz[3 * i + 1], pz = a.xyz[3 * i + 2];
                float fx, fy, fz;
                const bool in_range = scan_point_end(a.ox, a.oy, a.oz, px, py, pz, a.maxrange, fx, fy, fz);
                if (in_range) {
                    uint16_t qx, qy, qz;
                    if (coord_to_key3(a.res_factor, px, py, pz, qx, qy, qz)) {
                        uint32_t c;
                        if (cell_of_key(a, qx, qy, qz, c)) {
                            const unsigned vox = brick_voxel_index(qx, qy, qz);
                            uint32_t* w = a.cmasks + (size_t)c * 32 + (vox >> 5);
                            const uint32_t bit = 1u << (vox & 31u);
                            if (!(__ldcg(w) & bit)) {
                                atomicOr(w, bit);
                                touched[c] = 1;
                            }
                        } else {
                            a.counters[CNT_GRID_MISS] = 1;
                        }
                    }
                }
                Ray r;
                if (ray_setup(a.res, a.res_factor, a.ox, a.oy, a.oz, fx, fy, fz, r) == 1) {
                    L.kx = r.kx; L.ky = r.ky; L.kz = r.kz; L.ex = r.ex; L.ey = r.ey; L.ez = r.ez; L.sx = r.sx; L.sy = r.sy; L.sz = r.sz;
                    L.tmx = r.tmx; L.tmy = r.tmy; L.tmz = r.tmz; L.tdx = r.tdx; L.tdy = r.tdy; L.tdz = r.tdz;
                    L.length = (double)r.length;
                    L.csx = r.sx; L.csy = r.sy * (int)a.gdim; L.csz = r.sz * (int)(a.gdim * a.gdim);
                    const bool ok = cell_of_key(a, L.kx, L.ky, L.kz, L.cell);
                    if (!ok) a.counters[CNT_GRID_MISS] = 1;
                    active = ok;
                    // the origin cell is the first free cell; its word is read here (this path is long anyway)
                    L.widx = (ok ? L.cell : 0u) * 16u + 8u + sub_index(L.kx, L.ky, L.kz);
                    L.seen = ld_cg_u64(masks64 + L.widx);
                    L.age = 3;
                    L.mask = sub_bit(L.kx, L.ky, L.kz);
                    ++L.steps;
                    double t;
                    L.axis = ray_select(r, t);
                    L.mx = L.axis == 0; L.my = L.axis == 1; L.mz = L.axis == 2;
                }
            }
            continue;
        }
        if (act == 0) break;   // no ray left anywhere in this warp
        const int keep_going = exhausted ? 0 : 32 - K3_REFILL_MIN;
        do {
            if (active) active = dense_step(a, masks64, touched, L, stage_a);
            if (active) active = dense_step(a, masks64, touched, L, stage_b);
        } while (__popc(__ballot_sync(0xffffffffu, active)) > keep_going);
        // fetched words still on their way are dropped (their rays publish without them): the next round may start
        // with either staging register
        if (L.age < 2) L.age = 3;
    }
    // statistics only: free-cell visits of this scan
    unsigned long long total = L.steps;
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    if (lane == 0 && total) atomicAdd(reinterpret_cast<unsigned long long*>(&a.counters[CNT_STEPS_LO]), total);
}

// dense mode read-back, step 1: list the touched cells (one byte per cell, four cells per thread and load)
__global__ void k_cells_list(const uint32_t* __restrict__ touched, uint32_t n_words, uint32_t* list, uint32_t cap, uint32_t* counters,
                             const uint32_t* abort_flag) {
    if (__ldcg(abort_flag)) return;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += stride) {
        const uint32_t w = touched[i];
        if (!w) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            if ((w >> (8 * b)) & 0xffu) {
                const uint32_t q = atomicAdd(&counters[CNT_DELTA], 1u);
                if (q < cap) list[q] = i * 4u + (uint32_t)b;
            }
        }
    }
}
// step 2: one warp per listed cell -> record (free already minus occupied); clears the cell and its bit.  Does nothing
// when the list overflowed (the host grows the buffers and runs both steps again: the masks are still intact).
__global__ void __launch_bounds__(256) k_cells_emit(uint32_t* cmasks, uint8_t* touched, const uint32_t* __restrict__ list, uint32_t cap,
                                                    const uint32_t* counters, DeltaRecord* out, int gx0, int gy0, int gz0, uint32_t gdim,
                                                    uint32_t* abort_flag, int set_abort) {
    if (__ldcg(abort_flag)) return;
    const uint32_t n = counters[CNT_DELTA];
    if (n > cap) {   // pipelined use: stop every queued scan kernel until the host has grown the buffers and listed again
        if (set_abort && blockIdx.x == 0 && threadIdx.x == 0) atomicExch(abort_flag, 1u);
        return;
    }
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < n; q += warps) {
        const uint32_t cell = list[q];
        uint32_t w = cmasks[(size_t)cell * 32 + lane];
        const uint32_t occ_w = __shfl_sync(0xffffffffu, w, lane & 15u);
        if (lane >= 16) w &= ~occ_w;   // occupied wins
        if (lane == 0) {
            const uint32_t cx = cell % gdim, cy = (cell / gdim) % gdim, cz = cell / (gdim * gdim);
            out[q].key = (uint64_t)(uint32_t)(gx0 + (int)cx) | ((uint64_t)(uint32_t)(gy0 + (int)cy) << 13) | ((uint64_t)(uint32_t)(gz0 + (int)cz) << 26);
            touched[cell] = 0;
        }
        out[q].mask[lane] = w;
        cmasks[(size_t)cell * 32 + lane] = 0;
    }
}

// ---- hash mode
// slot of a brick in the scratch table (bounded probe length; a too-full table is grown by the host)
// (arguments by value: taking the address of the kernel-parameter struct would spill it to local memory)
__device__ __noinline__ uint32_t scratch_slot_hash(uint64_t* skeys, uint64_t scap, uint32_t* counters, uint64_t bk) {
    const uint64_t mask = scap - 1;
    uint64_t slot = hash64(bk) & mask;
    for (int probe = 0; probe < 128; ++probe) {
        const uint64_t k = ld_cg_u64(skeys + slot);
        if (k == bk) return (uint32_t)slot;
        if (k == kEmptyKey) {
            const unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(skeys + slot), kEmptyKey, bk);
            if (old == kEmptyKey) {
                if (atomicAdd(&counters[CNT_SCRATCH_USED], 1u) + 1u > (uint32_t)(scap / 2)) counters[CNT_OVERFLOW] = 1;
                return (uint32_t)slot;
            }
            if (old == bk) return (uint32_t)slot;
        }
        slot = (slot + 1) & mask;
    }
    counters[CNT_OVERFLOW] = 1;
    return 0xffffffffu;
}

__global__ void __launch_bounds__(K3_THREADS) k_scan_raycast_hash(const ScanArgs a, unsigned long long* ray_counter) {
    const unsigned lane = threadIdx.x & 31u;
    bool active = false, exhausted = false;
    Ray r;
    int axis = 0;
    double length = 0.0;
    uint64_t* word = nullptr;      // 64-bit free-mask word of the current sub-block
    uint64_t seen = ~0ull, mask = 0;
    uint32_t slot = 0xffffffffu;
    unsigned long long steps = 0;
    for (;;) {
        const unsigned act = __ballot_sync(0xffffffffu, active);
        const unsigned idle = ~act;
        if (!exhausted && __popc(idle) >= K3_REFILL_MIN) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(ray_counter, (unsigned long long)__popc(idle));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base + __popc(idle) >= a.n) exhausted = true;
            const unsigned long long i = base + __popc(idle & ((1u << lane) - 1u));
            if (!active && i < a.n) {
                const float px = a.xyz[3 * i], py = a.xyz[3 * i + 1], pz = a.xyz[3 * i + 2];
                float fx, fy, fz;
                const bool in_range = scan_point_end(a.ox, a.oy, a.oz, px, py, pz, a.maxrange, fx, fy, fz);
                if (in_range) {
                    uint16_t qx, qy, qz;
                    if (coord_to_key3(a.res_factor, px, py, pz, qx, qy, qz)) {
                        const uint32_t s = scratch_slot_hash(a.skeys, a.scap, a.counters, brick_key(qx, qy, qz));
                        if (s != 0xffffffffu) {
                            const unsigned vox = brick_voxel_index(qx, qy, qz);
                            uint32_t* w = a.smasks + (size_t)s * 32 + (vox >> 5);
                            const uint32_t bit = 1u << (vox & 31u);
                            if (!(__ldcg(w) & bit)) atomicOr(w, bit);
                        }
                    }
                }
                if (ray_setup(a.res, a.res_factor, a.ox, a.oy, a.oz, fx, fy, fz, r) == 1) {
                    active = true;
                    length = (double)r.length;
                    slot = scratch_slot_hash(a.skeys, a.scap, a.counters, brick_key((uint32_t)r.kx, (uint32_t)r.ky, (uint32_t)r.kz));
                    const bool ok = slot != 0xffffffffu;
                    word = reinterpret_cast<uint64_t*>(a.smasks + (size_t)(ok ? slot : 0u) * 32 + 16) + sub_index(r.kx, r.ky, r.kz);
                    seen = ldcg_u64_if(word, ~0ull, ok);
                    mask = sub_bit(r.kx, r.ky, r.kz);
                    ++steps;
                    double t;
                    axis = ray_select(r, t);
                }
            }
            continue;
        }
        if (act == 0) break;
        const int keep_going = exhausted ? 0 : 32 - K3_REFILL_MIN;
        do {
            if (active) {
                const int px = r.kx, py = r.ky, pz = r.kz;
                ray_advance(r, axis);
                double t;
                axis = ray_select(r, t);
                const bool done = ray_at_end(r) | (t > length);
                const int diff = (r.kx ^ px) | (r.ky ^ py) | (r.kz ^ pz);
                const bool new_sub = (diff >> 2) != 0;
                if ((done | new_sub) && (mask & ~seen) != 0) atomicOr(reinterpret_cast<unsigned long long*>(word), (unsigned long long)mask);
                const bool enter = new_sub & !done;
                if (enter) {
                    if ((diff >> 3) != 0) slot = scratch_slot_hash(a.skeys, a.scap, a.counters, brick_key((uint32_t)r.kx, (uint32_t)r.ky, (uint32_t)r.kz));
                    const bool ok = slot != 0xffffffffu;
                    word = reinterpret_cast<uint64_t*>(a.smasks + (size_t)(ok ? slot : 0u) * 32 + 16) + sub_index(r.kx, r.ky, r.kz);
                    seen = ldcg_u64_if(word, ~0ull, ok);
                    mask = 0;
                }
                mask |= sub_bit(r.kx, r.ky, r.kz);
                steps += done ? 0u : 1u;
                active = !done;
            }
        } while (__popc(__ballot_sync(0xffffffffu, active)) > keep_going);
    }
    for (int o = 16; o > 0; o >>= 1) steps += __shfl_xor_sync(0xffffffffu, steps, o);
    if (lane == 0 && steps) atomicAdd(reinterpret_cast<unsigned long long*>(&a.counters[CNT_STEPS_LO]), steps);
}

// computeDiscreteUpdate's pre-pass: keep one voxel-centre point per distinct endpoint key
__global__ void k_scan_discretize(const ScanArgs a, float* out_xyz) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
        uint16_t kx, ky, kz;
        if (!coord_to_key3(a.res_factor, a.xyz[3 * i], a.xyz[3 * i + 1], a.xyz[3 * i + 2], kx, ky, kz)) continue;
        bool inserted;
        const uint64_t slot = table_find_or_insert(a.skeys, a.scap, brick_key(kx, ky, kz), inserted);
        if (slot == kNoSlot) { a.counters[CNT_OVERFLOW] = 1; continue; }
        if (inserted && atomicAdd(&a.counters[CNT_SCRATCH_USED], 1u) + 1u > (uint32_t)(a.scap / 2)) a.counters[CNT_OVERFLOW] = 1;
        const unsigned vox = brick_voxel_index(kx, ky, kz);
        const uint32_t bit = 1u << (vox & 31u);
        const uint32_t old = atomicOr(a.smasks + slot * 32 + (vox >> 5), bit);
        if (!(old & bit)) {
            const uint32_t q = atomicAdd(&a.counters[CNT_DISCRETE], 1u);
            out_xyz[3 * q + 0] = (float)key_to_coord(a.res, kx);
            out_xyz[3 * q + 1] = (float)key_to_coord(a.res, ky);
            out_xyz[3 * q + 2] = (float)key_to_coord(a.res, kz);
        }
    }
}

// scratch table -> compact records (free already minus occupied); resets the slots it consumes.  One warp per slot.
__global__ void __launch_bounds__(256) k_scan_compact(uint64_t* skeys, uint32_t* smasks, uint64_t scap, DeltaRecord* out,
                                                      uint32_t* counters, uint32_t out_cap) {
    const unsigned lane = threadIdx.x & 31u;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t s = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < scap; s += warps) {
        const uint64_t k = skeys[s];
        if (k == kEmptyKey) continue;
        uint32_t w = smasks[s * 32 + lane];
        const uint32_t occ_w = __shfl_sync(0xffffffffu, w, lane & 15u);
        if (lane >= 16) w &= ~occ_w;   // occupied wins
        uint32_t q = 0;
        if (lane == 0) q = atomicAdd(&counters[CNT_DELTA], 1u);
        q = __shfl_sync(0xffffffffu, q, 0);
        if (q < out_cap) {
            if (lane == 0) out[q].key = k;
            out[q].mask[lane] = w;
        }
        smasks[s * 32 + lane] = 0;
        if (lane == 0) skeys[s] = kEmptyKey;
    }
}

// ------------------------------------------------------------------ K4: clamped log-odds apply of one scan's delta
// One warp per record.  Every brick appears once per delta, so the warp owns the brick's values for this launch.
__global__ void __launch_bounds__(256) k_apply_delta(const DeltaRecord* __restrict__ recs, uint32_t n, uint64_t* tkeys,
                                                     uint32_t* tvals, uint64_t tcap, float* values, uint32_t* known,
                                                     uint32_t* counters, float hit, float miss, float cmin, float cmax,
                                                     uint32_t part, uint32_t nparts) {
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += warps) {
        if (nparts > 1 && brick_owner(recs[r].key, nparts) != part) continue;   // another GPU owns this brick
        uint32_t idx = 0;
        if (lane == 0) {
            bool inserted;
            const uint64_t slot = table_find_or_insert(tkeys, tcap, recs[r].key, inserted);
            if (slot == kNoSlot) { counters[CNT_APPLY_OVERFLOW] = 1; idx = 0xffffffffu; }
            else if (inserted) { idx = atomicAdd(&counters[CNT_POOL_USED], 1u); tvals[slot] = idx; }
            else idx = tvals[slot];
        }
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx == 0xffffffffu) continue;
        // lane handles voxels [16*lane, 16*lane+16): half of mask word lane/2
        const uint32_t sh = (lane & 1u) * 16u;
        const uint32_t occ = (recs[r].mask[lane >> 1] >> sh) & 0xffffu;
        const uint32_t fre = (recs[r].mask[16 + (lane >> 1)] >> sh) & 0xffffu;
        const uint32_t any = occ | fre;
        if (any) {
            float4* v4 = reinterpret_cast<float4*>(values + (size_t)idx * kBrickVoxels + lane * 16);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t m = (any >> (4 * q)) & 0xfu;
                if (!m) continue;
                float4 v = v4[q];
                float* e = reinterpret_cast<float*>(&v);
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const uint32_t bit = 1u << (4 * q + b);
                    if (occ & bit) e[b] = clamped_add(e[b], hit, cmin, cmax);
                    else if (fre & bit) e[b] = clamped_add(e[b], miss, cmin, cmax);
                }
                v4[q] = v;
            }
        }
        const uint32_t other = __shfl_xor_sync(0xffffffffu, any, 1);
        if (!(lane & 1u)) {
            const uint32_t word = any | (other << 16);
            if (word) known[(size_t)idx * 16 + (lane >> 1)] |= word;
        }
    }
}

// ------------------------------------------------------------------ queries
__global__ void k_search(const uint16_t* __restrict__ keys, unsigned long long n, const uint64_t* tkeys, const uint32_t* tvals,
                         uint64_t tcap, const float* values, const uint32_t* known, float* out_v, uint8_t* out_f) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t kx = keys[3 * i], ky = keys[3 * i + 1], kz = keys[3 * i + 2];
        const uint64_t slot = table_find(tkeys, tcap, brick_key(kx, ky, kz));
        float v = 0.f;
        uint8_t f = 0;
        if (slot != kNoSlot) {
            const uint32_t idx = tvals[slot];
            const unsigned vox = brick_voxel_index(kx, ky, kz);
            if (known[(size_t)idx * 16 + (vox >> 5)] & (1u << (vox & 31u))) { f = 1; v = values[(size_t)idx * kBrickVoxels + vox]; }
        }
        if (out_v) out_v[i] = v;
        if (out_f) out_f[i] = f;
    }
}

__global__ void k_coord_to_key(const float* __restrict__ xyz, unsigned long long n, double res_factor, uint16_t* keys, uint8_t* valid) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint16_t kx = 0, ky = 0, kz = 0;
        const bool ok = coord_to_key3(res_factor, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], kx, ky, kz);
        keys[3 * i] = kx; keys[3 * i + 1] = ky; keys[3 * i + 2] = kz;
        if (valid) valid[i] = ok ? 1 : 0;
    }
}

// table -> pool_keys[pool index] = brick key
__global__ void k_table_to_pool_keys(const uint64_t* tkeys, const uint32_t* tvals, uint64_t tcap, uint64_t* pool_keys) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < tcap; s += stride) {
        const uint64_t k = tkeys[s];
        if (k != kEmptyKey) pool_keys[tvals[s]] = k;
    }
}

__global__ void k_rehash(const uint64_t* okeys, const uint32_t* ovals, uint64_t ocap, uint64_t* nkeys, uint32_t* nvals, uint64_t ncap) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < ocap; s += stride) {
        const uint64_t k = okeys[s];
        if (k == kEmptyKey) continue;
        bool inserted;
        const uint64_t slot = table_find_or_insert(nkeys, ncap, k, inserted);
        nvals[slot] = ovals[s];
    }
}

__global__ void k_count_known(const uint32_t* known, uint64_t n_words, unsigned long long* out) {
    unsigned long long c = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += stride) c += __popc(known[i]);
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31u) == 0 && c) atomicAdd(out, c);
}

// every known voxel -> (key, value); order unspecified
__global__ void k_export_voxels(const uint64_t* pool_keys, const float* values, const uint32_t* known, uint32_t n_bricks,
                                uint16_t* out_keys, float* out_vals, unsigned long long cap, unsigned long long* counter) {
    const uint64_t total = (uint64_t)n_bricks * kBrickVoxels;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const uint32_t b = (uint32_t)(i / kBrickVoxels), vox = (uint32_t)(i % kBrickVoxels);
        if (!(known[(size_t)b * 16 + (vox >> 5)] & (1u << (vox & 31u)))) continue;
        const unsigned long long q = atomicAdd(counter, 1ull);
        if (q >= cap) continue;
        uint32_t bx, by, bz, x, y, z;
        brick_key_unpack(pool_keys[b], bx, by, bz);
        brick_voxel_coords(vox, x, y, z);
        if (out_keys) { out_keys[3 * q] = (uint16_t)(bx * 8 + x); out_keys[3 * q + 1] = (uint16_t)(by * 8 + y); out_keys[3 * q + 2] = (uint16_t)(bz * 8 + z); }
        if (out_vals) out_vals[q] = values[i];
    }
}

__global__ void k_to_max_likelihood(float* values, const uint32_t* known, uint64_t total, float thres, float cmin, float cmax) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const uint32_t vox = (uint32_t)(i % kBrickVoxels);
        if (known[(i / kBrickVoxels) * 16 + (vox >> 5)] & (1u << (vox & 31u))) values[i] = values[i] >= thres ? cmax : cmin;
    }
}

// pool -> brick records (key, 512 log-odds, 16 known words); one warp per brick
__global__ void __launch_bounds__(256) k_export_bricks(const uint64_t* pool_keys, const float* values, const uint32_t* known,
                                                       uint32_t n_bricks, BrickRecord* out) {
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < n_bricks; b += warps) {
        if (lane == 0) out[b].key = pool_keys[b];
        if (lane < 16) out[b].known[lane] = known[(size_t)b * 16 + lane];
        for (uint32_t v = lane; v < kBrickVoxels; v += 32) out[b].value[v] = values[(size_t)b * kBrickVoxels + v];
    }
}

// brick records -> store (known voxels of the record overwrite / create the local ones); one warp per record
__global__ void __launch_bounds__(256) k_import_bricks(const BrickRecord* __restrict__ recs, uint32_t n, uint64_t* tkeys,
                                                       uint32_t* tvals, uint64_t tcap, float* values, uint32_t* known,
                                                       uint32_t* counters) {
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += warps) {
        uint32_t idx = 0;
        if (lane == 0) {
            bool inserted;
            const uint64_t slot = table_find_or_insert(tkeys, tcap, recs[r].key, inserted);
            if (slot == kNoSlot) { counters[CNT_OVERFLOW] = 1; idx = 0xffffffffu; }
            else if (inserted) { idx = atomicAdd(&counters[CNT_POOL_USED], 1u); tvals[slot] = idx; }
            else idx = tvals[slot];
        }
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx == 0xffffffffu) continue;
        for (uint32_t v = lane; v < kBrickVoxels; v += 32)
            if (recs[r].known[v >> 5] & (1u << (v & 31u))) values[(size_t)idx * kBrickVoxels + v] = recs[r].value[v];
        if (lane < 16) known[(size_t)idx * 16 + lane] |= recs[r].known[lane];
    }
}

// ------------------------------------------------------------------ host side: memory management
static unsigned grid_for(r3d_ctx* ctx, unsigned long long items, int block = 256, int per_sm = 8) {
    unsigned long long b = (items + block - 1) / block;
    const unsigned long long cap = (unsigned long long)ctx->sm_count * per_sm;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

int tree_sync_counters(r3d_tree* t) {
    r3d_ctx* ctx = t->ctx;
    R3D_CUDA_OK(ctx, cudaMemcpyAsync(ctx->pinned, t->counters, CNT_COUNT * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    R3D_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(t->h_counters, ctx->pinned, CNT_COUNT * sizeof(uint32_t));
    // the read-back is ordered after every queued kernel of the stream, so the mirror is exact again
    t->pool_used = t->h_counters[CNT_POOL_USED];
    t->pool_bound = t->pool_used;
    t->pool_dirty = false;
    if (t->h_counters[CNT_APPLY_OVERFLOW]) return set_error(ctx, R3D_ERR_STATE, "brick table overflow while applying a delta (internal sizing error)");
    return R3D_OK;
}

int tree_settle(r3d_tree* t) {
    if (!t->pool_dirty) return R3D_OK;
    return tree_sync_counters(t);
}

// zero every per-scan counter with one memset (everything but the pool cursor and the sticky apply-overflow flag)
static int tree_reset_scan_counters(r3d_tree* t) {
    r3d_ctx* ctx = t->ctx;
    // (includes the pipeline's abort flag: the serial path must never inherit one from a failed pipelined batch)
    R3D_CUDA_OK(ctx, cudaMemsetAsync(t->counters + 1, 0, (CNT_APPLY_OVERFLOW - 1) * sizeof(uint32_t), ctx->stream));
    return R3D_OK;
}

static int tree_set_counter(r3d_tree* t, int which, uint32_t v) {
    r3d_ctx* ctx = t->ctx;
    R3D_CUDA_OK(ctx, cudaMemcpyAsync(t->counters + which, &v, sizeof v, cudaMemcpyHostToDevice, ctx->stream));
    // v lives on the stack: the copy of 4 pageable bytes is staged before the call returns
    return R3D_OK;
}

// hash table with at least `want_entries` * 2 slots
static int tree_reserve_table(r3d_tree* t, uint64_t want_entries) {
    r3d_ctx* ctx = t->ctx;
    uint64_t need = 1024;
    while (need < want_entries * 2) need <<= 1;
    if (need <= t->tcap) return R3D_OK;
    uint64_t* nk = nullptr;
    uint32_t* nv = nullptr;
    R3D_CUDA_OK(ctx, cudaMalloc(&nk, need * sizeof(uint64_t)));
    cudaError_t e = cudaMalloc(&nv, need * sizeof(uint32_t));
    if (e != cudaSuccess) { cudaFree(nk); return set_error(ctx, R3D_ERR_OOM, "cudaMalloc(hash values) failed: %s", cudaGetErrorString(e)); }
    R3D_CUDA_OK(ctx, cudaMemsetAsync(nk, 0xff, need * sizeof(uint64_t), ctx->stream));
    R3D_CUDA_OK(ctx, cudaMemsetAsync(nv, 0, need * sizeof(uint32_t), ctx->stream));
    if (t->tcap) {
        k_rehash<<<grid_for(ctx, t->tcap), 256, 0, ctx->stream>>>(t->tkeys, t->tvals, t->tcap, nk, nv, need);
        ctx->launches++;
        R3D_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(t->tkeys);
        cudaFree(t->tvals);
    }
    t->tkeys = nk; t->tvals = nv; t->tcap = need;
    return R3D_OK;
}

// brick pool with room for `want` bricks; new bricks are zero (log-odds 0 = a freshly created node, nothing known)
static int tree_reserve_pool(r3d_tree* t, uint64_t want) {
    r3d_ctx* ctx = t->ctx;
    if (want <= t->pool_cap) return R3D_OK;
    // geometric growth from 4096 bricks (8.6 MB): a growth step costs a device allocation, a copy and a sync
    uint64_t ncap = t->pool_cap ? t->pool_cap : 4096;
    while (ncap < want) ncap *= 2;
    if (ncap > 0xfffffff0ull) return set_error(ctx, R3D_ERR_OOM, "brick pool would exceed 2^32 bricks");
    float* nv = nullptr;
    uint32_t* nk = nullptr;
    cudaError_t e = cudaMalloc(&nv, ncap * kBrickVoxels * sizeof(float));
    if (e != cudaSuccess) return set_error(ctx, R3D_ERR_OOM, "cudaMalloc(%llu bricks of log-odds) failed: %s", (unsigned long long)ncap, cudaGetErrorString(e));
    e = cudaMalloc(&nk, ncap * 16 * sizeof(uint32_t));
    if (e != cudaSuccess) { cudaFree(nv); return set_error(ctx, R3D_ERR_OOM, "cudaMalloc(known masks) failed: %s", cudaGetErrorString(e)); }
    const uint64_t used = t->pool_cap;   // everything below the old capacity may hold data
    if (used) {
        R3D_CUDA_OK(ctx, cudaMemcpyAsync(nv, t->values, used * kBrickVoxels * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
        R3D_CUDA_OK(ctx, cudaMemcpyAsync(nk, t->known, used * 16 * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    R3D_CUDA_OK(ctx, cudaMemsetAsync(nv + used * kBrickVoxels, 0, (ncap - used) * kBrickVoxels * sizeof(float), ctx->stream));
    R3D_CUDA_OK(ctx, cudaMemsetAsync(nk + used * 16, 0, (ncap - used) * 16 * sizeof(uint32_t), ctx->stream));
    R3D_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(t->values);
    cudaFree(t->known);
    t->values = nv; t->known = nk; t->pool_cap = ncap;
    return R3D_OK;
}

static int tree_reserve_scratch(r3d_tree* t, uint64_t want_slots) {
    r3d_ctx* ctx = t->ctx;
    uint64_t need = 1ull << 17;
    while (need < want_slots) need <<= 1;
    if (need <= t->scap) return R3D_OK;
    cudaFree(t->skeys); cudaFree(t->smasks); cudaFree(t->delta);
    t->skeys = nullptr; t->smasks = nullptr; t->delta = nullptr; t->scap = 0;
    R3D_CUDA_OK(ctx, cudaMalloc(&t->skeys, need * sizeof(uint64_t)));
    R3D_CUDA_OK(ctx, cudaMalloc(&t->smasks, need * 32 * sizeof(uint32_t)));
    R3D_CUDA_OK(ctx, cudaMalloc(&t->delta, (need / 2 + 1) * sizeof(DeltaRecord)));
    R3D_CUDA_OK(ctx, cudaMemsetAsync(t->skeys, 0xff, need * sizeof(uint64_t), ctx->stream));
    R3D_CUDA_OK(ctx, cudaMemsetAsync(t->smasks, 0, need * 32 * sizeof(uint32_t), ctx->stream));
    t->scap = need;
    t->delta_cap = need / 2 + 1;
    return R3D_OK;
}

static int tree_reset_scratch(r3d_tree* t) {
    r3d_ctx* ctx = t->ctx;
    R3D_CUDA_OK(ctx, cudaMemsetAsync(t->skeys, 0xff, t->scap * sizeof(uint64_t), ctx->stream));
    R3D_CUDA_OK(ctx, cudaMemsetAsync(t->smasks, 0, t->scap * 32 * sizeof(uint32_t), ctx->stream));
    return R3D_OK;
}

// device copy of a host array (or the pointer itself when it already is device memory)
template <typename T>
static int stage_in(r3d_ctx* ctx, int slot, const T* p, size_t count, const T** out) {
    if (is_device_ptr(p)) { *out = p; return R3D_OK; }
    R3D_TRY(scratch_reserve(ctx, slot, count * sizeof(T) + 16));
    R3D_CUDA_OK(ctx, cudaMemcpyAsync(ctx->scratch[slot], p, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    *out = reinterpret_cast<const T*>(ctx->scratch[slot]);
    return R3D_OK;
}

int tree_refresh_pool_keys(r3d_tree* t) {
    r3d_ctx* ctx = t->ctx;
    if (t->pool_keys_cap < t->pool_cap) {
        cudaFree(t->pool_keys);
        t->pool_keys = nullptr;
        R3D_CUDA_OK(ctx, cudaMalloc(&t->pool_keys, (t->pool_cap + 1) * sizeof(uint64_t)));
        t->pool_keys_cap = t->pool_cap;
    }
    if (t->tcap) {
        k_table_to_pool_keys<<<grid_for(ctx, t->tcap), 256, 0, ctx->stream>>>(t->tkeys, t->tvals, t->tcap, t->pool_keys);
        ctx->launches++;
    }
    R3D_CUDA_OK(ctx, cudaGetLastError());
    return R3D_OK;
}

template <typename T>
static int update_points_impl(r3d_tree* t, const T* xyz, uint64_t n, float upd, uint64_t* n_dropped) {
    if (!t) return set_error(nullptr, R3D_ERR_ARG, "null tree");
    r3d_ctx* ctx = t->ctx;
    if (!xyz && n) return set_error(ctx, R3D_ERR_ARG, "null points");
    DeviceSetter ds(ctx->device);
    R3D_TRY(tree_settle(t));
    uint64_t dropped_total = 0;
    const uint64_t chunk = 1ull << 22;
    const bool dev = n ? is_device_ptr(xyz) : true;
    for (uint64_t off = 0; off < n; off += chunk) {
        const uint64_t m = (n - off < chunk) ? n - off : chunk;
        const T* d = xyz + off * 3;
        if (!dev) R3D_TRY(stage_in(ctx, SCR_IN0, xyz + off * 3, (size_t)m * 3, &d));
        R3D_TRY(tree_reserve_table(t, (uint64_t)t->pool_used + m));
        R3D_TRY(tree_set_counter(t, CNT_DROPPED, 0));
        R3D_TRY(tree_set_counter(t, CNT_OVERFLOW, 0));
        k_points_ensure<T><<<grid_for(ctx, m), 256, 0, ctx->stream>>>(d, m, t->res_factor, t->tkeys, t->tvals, t->tcap, t->counters);
        ctx->launches++;
        R3D_TRY(tree_sync_counters(t));
        if (t->h_counters[CNT_OVERFLOW]) return set_error(ctx, R3D_ERR_STATE, "brick table overflow (internal sizing error)");
        t->pool_used = t->h_counters[CNT_POOL_USED];
        dropped_total += t->h_counters[CNT_DROPPED];
        R3D_TRY(tree_reserve_pool(t, t->pool_used));
        k_points_update<T><<<grid_for(ctx, m), 256, 0, ctx->stream>>>(d, m, t->res_factor, t->tkeys, t->tvals, t->tcap, t->values,
                                                                     t->known, upd, t->cmin, t->cmax);
        ctx->launches++;
        R3D_CUDA_OK(ctx, cudaGetLastError());
        if (!dev) R3D_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));   // staging buffer is reused by the next chunk
    }
    if (n_dropped) *n_dropped = dropped_total;
    return finish(ctx);
}

// Grid geometry of the dense mode: a cube of (2*reach+1)^3 bricks centred on the origin's brick, reach = maxrange plus
// a margin (every visited voxel contains a point of the ray no farther than the ray length from the origin; the margin
// is 4 voxels + 2 bricks).  Returns non-zero when the scan cannot use it -- unbounded range, scratch over budget, origin
// not finite, or the cube not completely inside the key range (keys would wrap like upstream's uint16) -- and the hash
// table serves those.
static int dense_grid_geometry(r3d_tree* t, const float origin[3], double maxrange, int* gx0, int* gy0, int* gz0, uint32_t* gdim) {
    if (!(maxrange >= 0.0)) return 1;
    const double reach_vox = ceil(maxrange * t->res_factor) + 4.0;
    if (!(reach_vox < 8.0 * 4000.0)) return 1;
    const int reach = (int)(reach_vox / 8.0) + 2;
    const uint64_t g = (uint64_t)(2 * reach + 1);
    if (g * g * g * 128ull > t->ctx->cell_budget_bytes || g * g * g > (1ull << 27)) return 1;
    int o[3];
    for (int i = 0; i < 3; ++i) {
        const double f = floor(t->res_factor * (double)origin[i]);
        if (!(f >= -32768.0 && f < 32768.0)) return 1;
        o[i] = (((int)f + r3d::kTreeMaxVal) >> 3) - reach;
        if (o[i] < 0 || o[i] + (int)g > 8192) return 1;
    }
    *gx0 = o[0]; *gy0 = o[1]; *gz0 = o[2];
    *gdim = (uint32_t)g;
    return 0;
}

// direct-mapped scan scratch of the context: `cells` x (32 mask words) + touched bitmap, all zero between scans
static int ctx_reserve_cells(r3d_ctx* ctx, uint64_t cells) {
    if (cells <= ctx->cell_cap && !ctx->cells_dirty) return R3D_OK;
    if (cells > ctx->cell_cap) {
        R3D_CUDA_OK(ctx, cudaDeviceSynchronize());
        cudaFree(ctx->cell_masks); cudaFree(ctx->cell_touched);
        ctx->cell_masks = nullptr; ctx->cell_touched = nullptr; ctx->cell_cap = 0;
        cudaError_t e = cudaMalloc(&ctx->cell_masks, cells * 128);
        if (e == cudaSuccess) e = cudaMalloc(&ctx->cell_touched, (cells / 4 + 1) * 4);   // one byte per cell
        if (e != cudaSuccess) {
            cudaGetLastError();
            cudaFree(ctx->cell_masks); ctx->cell_masks = nullptr;
            return set_error(ctx, R3D_ERR_OOM, "cudaMalloc(%llu MB of scan scratch) failed: %s", (unsigned long long)(cells * 128 >> 20), cudaGetErrorString(e));
        }
        ctx->cell_cap = cells;
        ctx->cells_dirty = true;
    }
    if (ctx->cells_dirty) {   // fresh memory, or a scan that was abandoned half-way
        R3D_CUDA_OK(ctx, cudaMemsetAsync(ctx->cell_masks, 0, ctx->cell_cap * 128, ctx->stream));
        R3D_CUDA_OK(ctx, cudaMemsetAsync(ctx->cell_touched, 0, (ctx->cell_cap / 4 + 1) * 4, ctx->stream));
        ctx->cells_dirty = false;
    }
    return R3D_OK;
}

static int tree_reserve_delta(r3d_tree* t, uint64_t want) {
    r3d_ctx* ctx = t->ctx;
    if (want <= t->delta_cap) return R3D_OK;
    R3D_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(t->delta);
    t->delta = nullptr; t->delta_cap = 0;
    R3D_CUDA_OK(ctx, cudaMalloc(&t->delta, want * sizeof(DeltaRecord)));
    t->delta_cap = want;
    return R3D_OK;
}

static unsigned raycast_blocks(r3d_tree* t, unsigned long long n_rays) {
    r3d_ctx* ctx = t->ctx;
    if (t->raycast_blocks_per_sm == 0) {
        int a = 0, b = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_scan_raycast_dense, K3_THREADS, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_scan_raycast_hash, K3_THREADS, 0);
        t->raycast_blocks_per_sm = a < 1 ? 1 : a;
        t->raycast_blocks_per_sm_hash = b < 1 ? 1 : b;
    }
    (void)ctx;
    return 0;
}

// dense mode: rays -> direct-mapped masks -> records.  Returns R3D_OK with *fallback = true when the hash path must
// take over (a ray left the grid: cannot happen by construction, kept as a safety net).
static int scan_delta_dense(r3d_tree* t, const ScanArgs& a0, bool* fallback) {
    r3d_ctx* ctx = t->ctx;
    ScanArgs a = a0;
    *fallback = false;
    R3D_TRY(ctx_reserve_cells(ctx, a.gcells));
    a.cmasks = ctx->cell_masks;
    a.ctouched = ctx->cell_touched;
    if (t->delta_cap < (1u << 16)) R3D_TRY(tree_reserve_delta(t, 1u << 16));
    ctx->cells_dirty = true;     // until the read-back below has cleaned up
    if (a.n) {
        unsigned long long* ray_counter = reinterpret_cast<unsigned long long*>(&t->counters[CNT_RAY_LO]);   // zeroed by the caller
        raycast_blocks(t, a.n);
        unsigned long long blocks = (unsigned long long)ctx->sm_count * t->raycast_blocks_per_sm;
        const unsigned long long need = (a.n + K3_THREADS - 1) / K3_THREADS;
        if (blocks > need) blocks = need;
        cudaEventRecord(ctx->ev_a, ctx->stream);
        k_scan_raycast_dense<<<(unsigned)blocks, K3_THREADS, 0, ctx->stream>>>(a, ray_counter, t->counters + CNT_ABORT);
        cudaEventRecord(ctx->ev_b, ctx->stream);
        ctx->launches++;
    }
    const uint32_t n_words = a.gcells / 4 + 1;
    for (int attempt = 0; attempt < 8; ++attempt) {
        R3D_TRY(scratch_reserve(ctx, SCR_TILE, (size_t)t->delta_cap * 4 + 256));
        if (attempt) R3D_TRY(tree_set_counter(t, CNT_DELTA, 0));
        k_cells_list<<<grid_for(ctx, n_words, 256, 8), 256, 0, ctx->stream>>>(a.ctouched, n_words, (uint32_t*)ctx->scratch[SCR_TILE],
                                                                              (uint32_t)t->delta_cap, t->counters, t->counters + CNT_ABORT);
        k_cells_emit<<<grid_for(ctx, (uint64_t)t->delta_cap * 32, 256, 8), 256, 0, ctx->stream>>>(
            a.cmasks, reinterpret_cast<uint8_t*>(a.ctouched), (const uint32_t*)ctx->scratch[SCR_TILE], (uint32_t)t->delta_cap, t->counters, t->delta, a.gx0, a.gy0, a.gz0, a.gdim,
            t->counters + CNT_ABORT, 0);
        ctx->launches += 2;
        R3D_CUDA_OK(ctx, cudaGetLastError());
        R3D_TRY(tree_sync_counters(t));
        if (t->h_counters[CNT_DELTA] <= t->delta_cap) {
            ctx->cells_dirty = false;
            if (a.n && attempt == 0) cudaEventElapsedTime(&ctx->last_kernel_ms, ctx->ev_a, ctx->ev_b);   // the read-back above fenced both
            if (t->h_counters[CNT_GRID_MISS]) { *fallback = true; return R3D_OK; }   // records discarded, scratch clean
            t->delta_n = t->h_counters[CNT_DELTA];
            return R3D_OK;
        }
        R3D_TRY(tree_reserve_delta(t, (uint64_t)t->h_counters[CNT_DELTA] * 2));   // list overflow: masks untouched, list again
    }
    return set_error(ctx, R3D_ERR_OOM, "scan delta does not fit the record buffer");
}

static int apply_delta_impl(r3d_tree* t, const DeltaRecord* d_recs, uint64_t n, uint32_t part = 0, uint32_t nparts = 1);

// ---- two-deep scan pipeline (dense mode, device-resident scans): scan s+1 is queued -- ray cast, cell list, records,
// counter read-back -- before the host waits for scan s's counters and queues its apply, so the GPU never idles on the
// host's turnaround.  Each slot has its own counters, cell list and record buffer; the cell scratch is shared (scan
// s's emit clears it before scan s+1's ray cast starts, in stream order).  If a scan's records do not fit, its emit
// sets the sticky abort flag instead of touching anything and every later queued kernel skips; the host then grows the
// buffers, clears the flag, lists / emits again and re-queues what was skipped.
static int pipe_reserve(r3d_tree* t, uint64_t cap) {
    r3d_ctx* ctx = t->ctx;
    if (!t->pipe_counters) {
        R3D_CUDA_OK(ctx, cudaMalloc(&t->pipe_counters, 2 * CNT_COUNT * sizeof(uint32_t)));
        R3D_CUDA_OK(ctx, cudaMemsetAsync(t->pipe_counters, 0, 2 * CNT_COUNT * sizeof(uint32_t), ctx->stream));
        for (int i = 0; i < 2; ++i) {
            R3D_CUDA_OK(ctx, cudaEventCreateWithFlags(&t->pipe_done[i], cudaEventDisableTiming));
            R3D_CUDA_OK(ctx, cudaEventCreateWithFlags(&t->rc_done[i], cudaEventDisableTiming));
        }
        R3D_CUDA_OK(ctx, cudaEventCreateWithFlags(&t->pipe_start, cudaEventDisableTiming));
        if (const char* v = getenv("R3D_PIPE_OVERLAP")) t->pipe_overlap = atoi(v) != 0;
    }
    // the ray-cast streams belong to the context (creating a stream costs milliseconds; trees come and go)
    for (int i = 0; i < 2; ++i)
        if (!ctx->rc_stream[i]) R3D_CUDA_OK(ctx, cudaStreamCreateWithFlags(&ctx->rc_stream[i], cudaStreamNonBlocking));
    if (t->delta_cap < cap) R3D_TRY(tree_reserve_delta(t, cap));
    if (t->delta_b_cap < t->delta_cap) {
        R3D_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(t->delta_b);
        t->delta_b = nullptr; t->delta_b_cap = 0;
        R3D_CUDA_OK(ctx, cudaMalloc(&t->delta_b, t->delta_cap * sizeof(DeltaRecord)));
        t->delta_b_cap = t->delta_cap;
    }
    if (t->pipe_list_cap < t->delta_cap) {
        R3D_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(t->pipe_list);
        t->pipe_list = nullptr; t->pipe_list_cap = 0;
        R3D_CUDA_OK(ctx, cudaMalloc(&t->pipe_list, 2 * t->delta_cap * sizeof(uint32_t)));
        t->pipe_list_cap = t->delta_cap;
    }
    return R3D_OK;
}

struct PipeScan {
    ScanArgs a;
    bool timed;
    bool overlap;    // ray cast on the slot's own stream (the scan has its own cell cube)
};

static int pipe_enqueue(r3d_tree* t, const PipeScan& ps, int slot, bool cast) {
    r3d_ctx* ctx = t->ctx;
    uint32_t* cnt = t->pipe_counters + slot * CNT_COUNT;
    ScanArgs a = ps.a;
    a.counters = cnt;
    uint32_t* abort_flag = t->counters + CNT_ABORT;
    if (cast) {
        cudaStream_t rs = ctx->stream;
        if (ps.overlap) {
            // the slot's stream starts after the batch's set-up and after the slot's previous scan has been emitted and its
            // counters read back (that scan used the same cube, counters and mailbox)
            rs = ctx->rc_stream[slot];
            R3D_CUDA_OK(ctx, cudaStreamWaitEvent(rs, t->pipe_start, 0));
            if (t->pipe_done_valid[slot]) R3D_CUDA_OK(ctx, cudaStreamWaitEvent(rs, t->pipe_done[slot], 0));
        }
        R3D_CUDA_OK(ctx, cudaMemsetAsync(cnt, 0, CNT_COUNT * sizeof(uint32_t), rs));
        if (a.n) {
            raycast_blocks(t, a.n);
            unsigned long long blocks = (unsigned long long)ctx->sm_count * t->raycast_blocks_per_sm;
            const unsigned long long need = (a.n + K3_THREADS - 1) / K3_THREADS;
            if (blocks > need) blocks = need;
            if (ps.timed) cudaEventRecord(ctx->ev_a, rs);
            k_scan_raycast_dense<<<(unsigned)blocks, K3_THREADS, 0, rs>>>(a, reinterpret_cast<unsigned long long*>(cnt + CNT_RAY_LO), abort_flag);
            if (ps.timed) cudaEventRecord(ctx->ev_b, rs);
            ctx->launches++;
        }
        if (ps.overlap) {
            R3D_CUDA_OK(ctx, cudaEventRecord(t->rc_done[slot], rs));
            R3D_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->stream, t->rc_done[slot], 0));
        }
    } else {
        R3D_CUDA_OK(ctx, cudaMemsetAsync(cnt + CNT_DELTA, 0, sizeof(uint32_t), ctx->stream));
    }
    const uint32_t n_words = a.gcells / 4 + 1;
    uint32_t* list = t->pipe_list + (size_t)slot * t->pipe_list_cap;
    DeltaRecord* out = slot ? t->delta_b : t->delta;
    k_cells_list<<<grid_for(ctx, n_words, 256, 8), 256, 0, ctx->stream>>>(a.ctouched, n_words, list, (uint32_t)t->delta_cap, cnt, abort_flag);
    k_cells_emit<<<grid_for(ctx, (uint64_t)t->delta_cap * 32, 256, 8), 256, 0, ctx->stream>>>(
        a.cmasks, reinterpret_cast<uint8_t*>(a.ctouched), list, (uint32_t)t->delta_cap, cnt, out, a.gx0, a.gy0, a.gz0, a.gdim, abort_flag, 1);
    ctx->launches += 2;
    R3D_CUDA_OK(ctx, cudaGetLastError());
    R3D_CUDA_OK(ctx, cudaMemcpyAsync((char*)ctx->pinned + 1024 + 256 * slot, cnt, CNT_COUNT * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    // the pool cursor as of this point of the stream (= after every apply queued before this scan): keeps the host's
    // upper bound of it tight without a synchronising read-back
    R3D_CUDA_OK(ctx, cudaMemcpyAsync((char*)ctx->pinned + 1024 + 256 * slot + 128, t->counters + CNT_POOL_USED, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    R3D_CUDA_OK(ctx, cudaEventRecord(t->pipe_done[slot], ctx->stream));
    t->pipe_done_valid[slot] = true;
    return R3D_OK;
}

// Returns the number of scans fully inserted; fewer than n_scans means "continue with the serial path from there".
static int insert_scans_pipelined(r3d_tree* t, const float* d_xyz, const uint64_t* n_points, const float* origins, uint32_t n_scans,
                                  double maxrange, uint32_t* done_out, uint64_t* rays_out, uint64_t* steps_out) {
    r3d_ctx* ctx = t->ctx;
    *done_out = 0;
    std::vector<PipeScan> scans(n_scans);
    uint64_t off = 0;
    uint32_t gcells_max = 0;
    for (uint32_t s = 0; s < n_scans; ++s) {
        ScanArgs& a = scans[s].a;
        memset(&a, 0, sizeof a);
        int gx0, gy0, gz0;
        uint32_t gdim;
        if (n_points[s] > 0xfffffff0ull || dense_grid_geometry(t, origins + 3 * (size_t)s, maxrange, &gx0, &gy0, &gz0, &gdim) != 0) return R3D_OK;   // serial path decides
        a.xyz = d_xyz + off * 3; a.n = n_points[s];
        a.ox = origins[3 * s]; a.oy = origins[3 * s + 1]; a.oz = origins[3 * s + 2];
        a.maxrange = maxrange; a.res = t->res; a.res_factor = t->res_factor;
        a.gx0 = gx0; a.gy0 = gy0; a.gz0 = gz0; a.gdim = gdim; a.gcells = gdim * gdim * gdim;
        if (a.gcells > gcells_max) gcells_max = a.gcells;
        scans[s].timed = s + 1 == n_scans;
        off += n_points[s];
    }
    R3D_TRY(pipe_reserve(t, t->delta_cap < (1u << 16) ? (1u << 16) : t->delta_cap));
    // two cell cubes (one per pipeline slot) when the scratch budget allows: the next scan's ray cast then overlaps this
    // scan's tail, list, emit and apply
    const uint64_t cube_cells = ((uint64_t)gcells_max / 4 + 1) * 4;
    const bool overlap = t->pipe_overlap && n_scans > 1 && 2 * cube_cells * 128ull <= ctx->cell_budget_bytes;
    R3D_TRY(ctx_reserve_cells(ctx, overlap ? 2 * cube_cells : gcells_max));
    R3D_CUDA_OK(ctx, cudaMemsetAsync(t->counters + CNT_ABORT, 0, sizeof(uint32_t), ctx->stream));
    for (uint32_t s = 0; s < n_scans; ++s) {
        const uint64_t cube = overlap ? (uint64_t)(s & 1u) * cube_cells : 0;
        scans[s].a.cmasks = ctx->cell_masks + cube * 32;
        scans[s].a.ctouched = ctx->cell_touched + cube / 4;
        scans[s].overlap = overlap;
    }
    ctx->cells_dirty = true;
    if (overlap) {
        R3D_CUDA_OK(ctx, cudaEventRecord(t->pipe_start, ctx->stream));   // cubes cleared, abort flag reset, scans resident
        t->pipe_done_valid[0] = t->pipe_done_valid[1] = false;
    }
    auto now_ns = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec; };
    t->pipe_wait_ns = t->pipe_work_ns = t->pipe_max_turn_ns = t->pipe_scans = 0;
    uint64_t t_mark = now_ns();
    R3D_TRY(pipe_enqueue(t, scans[0], 0, true));
    uint64_t prev_records = 0;      // records of the apply queued last (not yet reflected in the cursor read back below)
    for (uint32_t s = 0; s < n_scans; ++s) {
        const int slot = (int)(s & 1u);
        if (s + 1 < n_scans) R3D_TRY(pipe_enqueue(t, scans[s + 1], slot ^ 1, true));
        uint32_t hc[CNT_COUNT];
        for (int attempt = 0;; ++attempt) {
            const uint64_t t_wait = now_ns();
            R3D_CUDA_OK(ctx, cudaEventSynchronize(t->pipe_done[slot]));
            const uint64_t t_got = now_ns();
            t->pipe_work_ns += t_wait - t_mark;
            if (t_wait - t_mark > t->pipe_max_turn_ns) t->pipe_max_turn_ns = t_wait - t_mark;
            t->pipe_wait_ns += t_got - t_wait;
            t_mark = t_got;
            memcpy(hc, (char*)ctx->pinned + 1024 + 256 * slot, sizeof hc);
            if (hc[CNT_DELTA] <= t->delta_cap) break;
            if (attempt >= 8) return set_error(ctx, R3D_ERR_OOM, "scan delta does not fit the record buffer");
            // records of scan s did not fit: nothing after its list kernel has run (abort flag).  Grow, clear, list again,
            // and queue scan s+1 again.
            R3D_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
            if (overlap) {   // a ray cast that started before the flag was raised may still be running
                R3D_CUDA_OK(ctx, cudaStreamSynchronize(ctx->rc_stream[0]));
                R3D_CUDA_OK(ctx, cudaStreamSynchronize(ctx->rc_stream[1]));
            }
            R3D_TRY(pipe_reserve(t, (uint64_t)hc[CNT_DELTA] * 2));   // the other slot's records were applied already
            R3D_CUDA_OK(ctx, cudaMemsetAsync(t->counters + CNT_ABORT, 0, sizeof(uint32_t), ctx->stream));
            if (overlap) R3D_CUDA_OK(ctx, cudaEventRecord(t->pipe_start, ctx->stream));   // ray casts queued from here on see the cleared flag
            R3D_TRY(pipe_enqueue(t, scans[s], slot, false));
            if (s + 1 < n_scans) R3D_TRY(pipe_enqueue(t, scans[s + 1], slot ^ 1, true));
        }
        if (hc[CNT_GRID_MISS]) {
            // cannot happen by construction; hand the rest (from this scan on) to the serial path, which falls back to the hash table
            R3D_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
            ctx->cells_dirty = false;   // both queued scans have been emitted: the scratch is clean
            return R3D_OK;
        }
        {   // scan s's read-back was queued after apply(s-2) and before apply(s-1)
            uint32_t cursor;
            memcpy(&cursor, (char*)ctx->pinned + 1024 + 256 * slot + 128, sizeof cursor);
            const uint64_t tight = (uint64_t)cursor + prev_records;
            if (tight < t->pool_bound) t->pool_bound = tight;
        }
        R3D_TRY(apply_delta_impl(t, slot ? t->delta_b : t->delta, hc[CNT_DELTA]));
        prev_records = hc[CNT_DELTA];
        *rays_out += scans[s].a.n;
        *steps_out += (uint64_t)hc[CNT_STEPS_LO] | ((uint64_t)hc[CNT_STEPS_HI] << 32);
        t->delta_n = hc[CNT_DELTA];
        *done_out = s + 1;
        t->pipe_scans = s + 1;
        if (scans[s].timed && scans[s].a.n) cudaEventElapsedTime(&ctx->last_kernel_ms, ctx->ev_a, ctx->ev_b);
    }
    if ((n_scans & 1u) == 0u) {   // the last scan used slot 1: keep t->delta = "records of the last scan"
        DeltaRecord* tmp = t->delta; t->delta = t->delta_b; t->delta_b = tmp;
        const uint64_t c = t->delta_cap; t->delta_cap = t->delta_b_cap; t->delta_b_cap = c;
    }
    ctx->cells_dirty = false;
    return R3D_OK;
}

// ray-cast one scan into the scratch table and compact it into t->delta (t->delta_n records)
static int scan_delta_impl(r3d_tree* t, const float* xyz, uint64_t n, const float origin[3], double maxrange, int discretize) {
    r3d_ctx* ctx = t->ctx;
    if ((!xyz && n) || !origin) return set_error(ctx, R3D_ERR_ARG, "null scan buffer");
    if (n > 0xfffffff0ull) return set_error(ctx, R3D_ERR_ARG, "scan too large");
    const float* d = xyz;
    if (n) R3D_TRY(stage_in(ctx, SCR_IN0, xyz, (size_t)n * 3, &d));
    ScanArgs a;
    memset(&a, 0, sizeof a);
    a.xyz = d; a.n = n;
    a.ox = origin[0]; a.oy = origin[1]; a.oz = origin[2];
    a.maxrange = maxrange; a.res = t->res; a.res_factor = t->res_factor;
    a.counters = t->counters;
    int gx0 = 0, gy0 = 0, gz0 = 0;
    uint32_t gdim = 0;
    bool dense = dense_grid_geometry(t, origin, maxrange, &gx0, &gy0, &gz0, &gdim) == 0;
    if (!dense || discretize) R3D_TRY(tree_reserve_scratch(t, t->scap ? t->scap : (1ull << 18)));
    for (int attempt = 0; attempt < 14; ++attempt) {
        a.xyz = d; a.n = n;
        a.skeys = t->skeys; a.smasks = t->smasks; a.scap = t->scap;
        R3D_TRY(tree_reset_scan_counters(t));
        bool overflow = false;
        if (discretize && n) {
            R3D_TRY(scratch_reserve(ctx, SCR_IN1, (size_t)n * 12 + 16));
            k_scan_discretize<<<grid_for(ctx, n), 256, 0, ctx->stream>>>(a, (float*)ctx->scratch[SCR_IN1]);
            ctx->launches++;
            R3D_TRY(tree_sync_counters(t));
            overflow = t->h_counters[CNT_OVERFLOW] != 0;
            R3D_TRY(tree_reset_scratch(t));
            R3D_TRY(tree_set_counter(t, CNT_SCRATCH_USED, 0));
            a.xyz = (const float*)ctx->scratch[SCR_IN1];
            a.n = t->h_counters[CNT_DISCRETE];
        }
        if (!overflow && dense) {
            a.gx0 = gx0; a.gy0 = gy0; a.gz0 = gz0; a.gdim = gdim; a.gcells = gdim * gdim * gdim;
            bool fallback = false;
            R3D_TRY(scan_delta_dense(t, a, &fallback));
            if (!fallback) {
                t->last_scan_rays = a.n;
                t->last_scan_steps = (uint64_t)t->h_counters[CNT_STEPS_LO] | ((uint64_t)t->h_counters[CNT_STEPS_HI] << 32);
                return R3D_OK;
            }
            dense = false;
            R3D_TRY(tree_reserve_scratch(t, t->scap ? t->scap : (1ull << 18)));
            continue;
        }
        if (!overflow && a.n) {
            // persistent warps pulling rays from a counter (slot CNT_RAY_LO/HI of the tree's counter block, zeroed above)
            unsigned long long* ray_counter = reinterpret_cast<unsigned long long*>(&t->counters[CNT_RAY_LO]);
            raycast_blocks(t, a.n);
            unsigned long long blocks = (unsigned long long)ctx->sm_count * t->raycast_blocks_per_sm_hash;
            const unsigned long long need = (a.n + K3_THREADS - 1) / K3_THREADS;
            if (blocks > need) blocks = need;
            k_scan_raycast_hash<<<(unsigned)blocks, K3_THREADS, 0, ctx->stream>>>(a, ray_counter);
            ctx->launches++;
        }
        if (!overflow) {
            k_scan_compact<<<grid_for(ctx, t->scap * 32, 256, 8), 256, 0, ctx->stream>>>(t->skeys, t->smasks, t->scap, t->delta, t->counters,
                                                                                      (uint32_t)t->delta_cap);
            ctx->launches++;
            R3D_CUDA_OK(ctx, cudaGetLastError());
            R3D_TRY(tree_sync_counters(t));
            overflow = t->h_counters[CNT_OVERFLOW] != 0 || t->h_counters[CNT_DELTA] > t->delta_cap;
        }
        if (!overflow) {
            t->delta_n = t->h_counters[CNT_DELTA];
            t->last_scan_rays = a.n;
            t->last_scan_steps = (uint64_t)t->h_counters[CNT_STEPS_LO] | ((uint64_t)t->h_counters[CNT_STEPS_HI] << 32);
            return R3D_OK;
        }
        // table too small for this scan: grow, wipe, cast again (ray casting is a pure function of the scan)
        R3D_TRY(tree_reserve_scratch(t, t->scap * 4));
        R3D_TRY(tree_reset_scratch(t));
    }
    return set_error(ctx, R3D_ERR_OOM, "scan delta does not fit the scratch table");
}

static int apply_delta_impl(r3d_tree* t, const DeltaRecord* d_recs, uint64_t n, uint32_t part, uint32_t nparts) {
    r3d_ctx* ctx = t->ctx;
    if (n == 0) return R3D_OK;
    // sized from the host-side upper bound of the pool cursor: no read-back between a scan's apply and the next scan.
    // When the BOUND (not necessarily the pool) would outgrow the capacity, read the exact cursor back first: several
    // applies in a row (multi-GPU rounds) inflate the bound by every record, most of which hit existing bricks.
    if (t->pool_dirty && t->pool_bound + n > t->pool_cap) R3D_TRY(tree_settle(t));
    R3D_TRY(tree_reserve_table(t, t->pool_bound + n));
    R3D_TRY(tree_reserve_pool(t, t->pool_bound + n));
    k_apply_delta<<<grid_for(ctx, n * 32, 256, 8), 256, 0, ctx->stream>>>(d_recs, (uint32_t)n, t->tkeys, t->tvals, t->tcap, t->values, t->known,
                                                                          t->counters, t->hit, t->miss, t->cmin, t->cmax, part, nparts);
    ctx->launches++;
    R3D_CUDA_OK(ctx, cudaGetLastError());
    t->pool_bound += n;
    t->pool_dirty = true;
    return R3D_OK;
}

}  // namespace r3d

using namespace r3d;

// ------------------------------------------------------------------ C ABI
static float logodds_f(double p) { return (float)log(p / (1 - p)); }

extern "C" int r3d_tree_create(r3d_ctx* ctx, double resolution, r3d_tree** tree) {
    if (!ctx) return set_error(nullptr, R3D_ERR_ARG, "null context");
    if (!tree) return set_error(ctx, R3D_ERR_ARG, "null out pointer");
    if (!(resolution > 0)) return set_error(ctx, R3D_ERR_ARG, "resolution must be positive");
    DeviceSetter ds(ctx->device);
    r3d_tree* t = new r3d_tree();
    t->ctx = ctx;
    t->res = resolution;
    t->res_factor = 1.0 / resolution;
    t->hit = logodds_f(0.7); t->miss = logodds_f(0.4);
    t->cmin = logodds_f(0.1192); t->cmax = logodds_f(0.971);
    t->occ_thres = logodds_f(0.5);
    cudaError_t e = cudaMalloc(&t->counters, CNT_COUNT * sizeof(uint32_t));
    if (e != cudaSuccess) { delete t; return set_error(ctx, R3D_ERR_OOM, "cudaMalloc(counters): %s", cudaGetErrorString(e)); }
    cudaMemsetAsync(t->counters, 0, CNT_COUNT * sizeof(uint32_t), ctx->stream);
    int rc = tree_reserve_table(t, 1024);
    if (rc == R3D_OK) rc = tree_reserve_pool(t, 1024);
    if (rc != R3D_OK) { r3d_tree_destroy(t); return rc; }
    *tree = t;
    return finish(ctx);
}

extern "C" void r3d_tree_destroy(r3d_tree* t) {
    if (!t) return;
    DeviceSetter ds(t->ctx->device);
    cudaStreamSynchronize(t->ctx->stream);
    cudaFree(t->tkeys); cudaFree(t->tvals); cudaFree(t->values); cudaFree(t->known); cudaFree(t->pool_keys);
    cudaFree(t->skeys); cudaFree(t->smasks); cudaFree(t->delta); cudaFree(t->counters);
    cudaFree(t->delta_b); cudaFree(t->pipe_counters); cudaFree(t->pipe_list);
    for (int i = 0; i < 2; ++i) {
        if (t->ctx->rc_stream[i]) cudaStreamSynchronize(t->ctx->rc_stream[i]);
        if (t->pipe_done[i]) cudaEventDestroy(t->pipe_done[i]);
        if (t->rc_done[i]) cudaEventDestroy(t->rc_done[i]);
    }
    if (t->pipe_start) cudaEventDestroy(t->pipe_start);
    delete t;
}

extern "C" int r3d_tree_clear(r3d_tree* t) {
    if (!t) return set_error(nullptr, R3D_ERR_ARG, "null tree");
    r3d_ctx* ctx = t->ctx;
    DeviceSetter ds(ctx->device);
    R3D_CUDA_OK(ctx, cudaMemsetAsync(t->tkeys, 0xff, t->tcap * sizeof(uint64_t), ctx->stream));
    R3D_CUDA_OK(ctx, cudaMemsetAsync(t->values, 0, t->pool_cap * kBrickVoxels * sizeof(float), ctx->stream));
    R3D_CUDA_OK(ctx, cudaMemsetAsync(t->known, 0, t->pool_cap * 16 * sizeof(uint32_t), ctx->stream));
    R3D_CUDA_OK(ctx, cudaMemsetAsync(t->counters, 0, CNT_COUNT * sizeof(uint32_t), ctx->stream));
    t->pool_used = 0;
    t->pool_bound = 0;
    t->pool_dirty = false;
    t->delta_n = 0;
    return finish(ctx);
}

extern "C" int r3d_tree_reserve(r3d_tree* t, uint64_t n_bricks) {
    if (!t) return set_error(nullptr, R3D_ERR_ARG, "null tree");
    if (n_bricks > 0xfffffff0ull) return set_error(t->ctx, R3D_ERR_ARG, "too many bricks");
    DeviceSetter ds(t->ctx->device);
    R3D_TRY(tree_reserve_table(t, n_bricks));
    R3D_TRY(tree_reserve_pool(t, n_bricks));
    return finish(t->ctx);
}

extern "C" int r3d_tree_resolution(r3d_tree* t, double* res) {
    if (!t || !res) return set_error(t ? t->ctx : nullptr, R3D_ERR_ARG, "null argument");
    *res = t->res;
    return R3D_OK;
}

extern "C" int r3d_tree_params(r3d_tree* t, float out[5]) {
    if (!t || !out) return set_error(t ? t->ctx : nullptr, R3D_ERR_ARG, "null argument");
    out[0] = t->hit; out[1] = t->miss; out[2] = t->cmin; out[3] = t->cmax; out[4] = t->occ_thres;
    return R3D_OK;
}

extern "C" int r3d_tree_update_points(r3d_tree* t, const float* xyz, uint64_t n, int occupied, uint64_t* n_dropped) {
    return update_points_impl<float>(t, xyz, n, t ? (occupied ? t->hit : t->miss) : 0.f, n_dropped);
}
extern "C" int r3d_tree_update_points_f64(r3d_tree* t, const double* xyz, uint64_t n, int occupied, uint64_t* n_dropped) {
    return update_points_impl<double>(t, xyz, n, t ? (occupied ? t->hit : t->miss) : 0.f, n_dropped);
}
extern "C" int r3d_tree_update_points_logodds(r3d_tree* t, const float* xyz, uint64_t n, float upd, uint64_t* n_dropped) {
    return update_points_impl<float>(t, xyz, n, upd, n_dropped);
}

extern "C" int r3d_scan_delta_compute(r3d_tree* t, const float* xyz, uint64_t n, const float origin[3], double maxrange,
                                      int discretize, uint64_t* n_records) {
    if (!t) return set_error(nullptr, R3D_ERR_ARG, "null tree");
    DeviceSetter ds(t->ctx->device);
    R3D_TRY(scan_delta_impl(t, xyz, n, origin, maxrange, discretize));
    if (n_records) *n_records = t->delta_n;
    return finish(t->ctx);
}

extern "C" int r3d_scan_delta_export(r3d_tree* t, void* records, uint64_t capacity_records, uint64_t* n_records) {
    if (!t) return set_error(nullptr, R3D_ERR_ARG, "null tree");
    r3d_ctx* ctx = t->ctx;
    if (n_records) *n_records = t->delta_n;
    if (t->delta_n > capacity_records) return set_error(ctx, R3D_ERR_ARG, "delta has %llu records, buffer holds %llu", (unsigned long long)t->delta_n, (unsigned long long)capacity_records);
    if (t->delta_n == 0) return R3D_OK;
    if (!records) return set_error(ctx, R3D_ERR_ARG, "null record buffer");
    DeviceSetter ds(ctx->device);
    R3D_CUDA_OK(ctx, cudaMemcpyAsync(records, t->delta, t->delta_n * sizeof(DeltaRecord), cudaMemcpyDefault, ctx->stream));
    R3D_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return R3D_OK;
}

extern "C" int r3d_tree_apply_delta(r3d_tree* t, const void* records, uint64_t n_records) {
    if (!t) return set_error(nullptr, R3D_ERR_ARG, "null tree");
    r3d_ctx* ctx = t->ctx;
    if (!records && n_records) return set_error(ctx, R3D_ERR_ARG, "null records");
    if (n_records > 0xfffffff0ull) return set_error(ctx, R3D_ERR_ARG, "too many records");
    DeviceSetter ds(ctx->device);
    const DeltaRecord* d = reinterpret_cast<const DeltaRecord*>(records);
    if (n_records) R3D_TRY(stage_in(ctx, SCR_OUT0, reinterpret_cast<const DeltaRecord*>(records), (size_t)n_records, &d));
    R3D_TRY(apply_delta_impl(t, d, n_records));
    return finish(ctx);
}

extern "C" int r3d_tree_apply_delta_owned(r3d_tree* t, const void* records, uint64_t n_records, uint32_t part, uint32_t nparts) {
    if (!t) return set_error(nullptr, R3D_ERR_ARG, "null tree");
    r3d_ctx* ctx = t->ctx;
    if (!records && n_records) return set_error(ctx, R3D_ERR_ARG, "null records");
    if (n_records > 0xfffffff0ull) return set_error(ctx, R3D_ERR_ARG, "too many records");
    if (nparts == 0 || part >= nparts) return set_error(ctx, R3D_ERR_ARG, "bad partition %u of %u", part, nparts);
    DeviceSetter ds(ctx->device);
    const DeltaRecord* d = reinterpret_cast<const DeltaRecord*>(records);
    if (n_records) R3D_TRY(stage_in(ctx, SCR_OUT0, reinterpret_cast<const DeltaRecord*>(records), (size_t)n_records, &d));
    R3D_TRY(apply_delta_impl(t, d, n_records, part, nparts));
    return finish(ctx);
}

extern "C" int r3d_tree_num_bricks(r3d_tree* t, uint64_t* n) {
    if (!t || !n) return set_error(t ? t->ctx : nullptr, R3D_ERR_ARG, "null argument");
    DeviceSetter ds(t->ctx->device);
    R3D_TRY(tree_settle(t));
    *n = t->pool_used;
    return R3D_OK;
}

extern "C" int r3d_tree_export_bricks(r3d_tree* t, void* records, uint64_t capacity, uint64_t* n) {
    if (!t) return set_error(nullptr, R3D_ERR_ARG, "null tree");
    r3d_ctx* ctx = t->ctx;
    { DeviceSetter ds0(ctx->device); R3D_TRY(tree_settle(t)); }
    if (n) *n = t->pool_used;
    if (t->pool_used > capacity) return set_error(ctx, R3D_ERR_ARG, "map has %u bricks, buffer holds %llu", t->pool_used, (unsigned long long)capacity);
    if (t->pool_used == 0) return R3D_OK;
    if (!records) return set_error(ctx, R3D_ERR_ARG, "null brick buffer");
    DeviceSetter ds(ctx->device);
    R3D_TRY(tree_refresh_pool_keys(t));
    const bool dev = is_device_ptr(records);
    BrickRecord* d = reinterpret_cast<BrickRecord*>(records);
    if (!dev) { R3D_TRY(scratch_reserve(ctx, SCR_OUT0, (size_t)t->pool_used * sizeof(BrickRecord))); d = (BrickRecord*)ctx->scratch[SCR_OUT0]; }
    k_export_bricks<<<grid_for(ctx, (uint64_t)t->pool_used * 32, 256, 8), 256, 0, ctx->stream>>>(t->pool_keys, t->values, t->known, t->pool_used, d);
    ctx->launches++;
    R3D_CUDA_OK(ctx, cudaGetLastError());
    if (!dev) R3D_CUDA_OK(ctx, cudaMemcpyAsync(records, d, (size_t)t->pool_used * sizeof(BrickRecord), cudaMemcpyDeviceToHost, ctx->stream));
    R3D_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return R3D_OK;
}

extern "C" int r3d_tree_import_bricks(r3d_tree* t, const void* records, uint64_t n_records) {
    if (!t) return set_error(nullptr, R3D_ERR_ARG, "null tree");
    r3d_ctx* ctx = t->ctx;
    if (!records && n_records) return set_error(ctx, R3D_ERR_ARG, "null records");
    if (n_records > 0xfffffff0ull) return set_error(ctx, R3D_ERR_ARG, "too many records");
    if (n_records == 0) return R3D_OK;
    DeviceSetter ds(ctx->device);
    const BrickRecord* d = reinterpret_cast<const BrickRecord*>(records);
    R3D_TRY(tree_settle(t));
    R3D_TRY(stage_in(ctx, SCR_OUT0, reinterpret_cast<const BrickRecord*>(records), (size_t)n_records, &d));
    R3D_TRY(tree_reserve_table(t, (uint64_t)t->pool_used + n_records));
    R3D_TRY(tree_reserve_pool(t, (uint64_t)t->pool_used + n_records));
    R3D_TRY(tree_set_counter(t, CNT_OVERFLOW, 0));
    k_import_bricks<<<grid_for(ctx, n_records * 32, 256, 8), 256, 0, ctx->stream>>>(d, (uint32_t)n_records, t->tkeys, t->tvals, t->tcap, t->values,
                                                                               t->known, t->counters);
    ctx->launches++;
    R3D_CUDA_OK(ctx, cudaGetLastError());
    R3D_TRY(tree_sync_counters(t));
    if (t->h_counters[CNT_OVERFLOW]) return set_error(ctx, R3D_ERR_STATE, "brick table overflow while importing bricks");
    t->pool_used = t->h_counters[CNT_POOL_USED];
    return finish(ctx);
}

extern "C" int r3d_tree_insert_scan(r3d_tree* t, const float* xyz, uint64_t n, const float origin[3], double maxrange, int discretize) {
    if (!t) return set_error(nullptr, R3D_ERR_ARG, "null tree");
    DeviceSetter ds(t->ctx->device);
    R3D_TRY(scan_delta_impl(t, xyz, n, origin, maxrange, discretize));
    R3D_TRY(apply_delta_impl(t, t->delta, t->delta_n));
    return finish(t->ctx);
}

extern "C" int r3d_tree_insert_scans(r3d_tree* t, const float* xyz, const uint64_t* n_points, const float* origins, uint32_t n_scans,
                                     double maxrange, int discretize) {
    if (!t) return set_error(nullptr, R3D_ERR_ARG, "null tree");
    r3d_ctx* ctx = t->ctx;
    if (n_scans && (!n_points || !origins)) return set_error(ctx, R3D_ERR_ARG, "null scan table");
    if (is_device_ptr(n_points) || is_device_ptr(origins)) return set_error(ctx, R3D_ERR_ARG, "n_points / origins must be host arrays");
    DeviceSetter ds(ctx->device);
    uint64_t off = 0, rays = 0, steps = 0;
    uint32_t first = 0;
    if (n_scans > 1 && !discretize && maxrange >= 0.0 && xyz && is_device_ptr(xyz)) {
        R3D_TRY(insert_scans_pipelined(t, xyz, n_points, origins, n_scans, maxrange, &first, &rays, &steps));
        for (uint32_t s = 0; s < first; ++s) off += n_points[s];
    }
    for (uint32_t s = first; s < n_scans; ++s) {
        R3D_TRY(scan_delta_impl(t, xyz ? xyz + off * 3 : nullptr, n_points[s], origins + 3 * (size_t)s, maxrange, discretize));
        R3D_TRY(apply_delta_impl(t, t->delta, t->delta_n));
        off += n_points[s];
        rays += t->last_scan_rays;
        steps += t->last_scan_steps;
    }
    t->last_scan_rays = rays;       // totals of the batch
    t->last_scan_steps = steps;
    return finish(ctx);
}

extern "C" int r3d_scan_deltas_compute(r3d_tree* t, const float* xyz, const uint64_t* n_points, const float* origins, uint32_t n_scans,
                                       double maxrange, int discretize, void* records, uint64_t capacity_records, uint64_t* counts) {
    if (!t) return set_error(nullptr, R3D_ERR_ARG, "null tree");
    r3d_ctx* ctx = t->ctx;
    if (n_scans && (!n_points || !origins || !counts)) return set_error(ctx, R3D_ERR_ARG, "null scan table");
    DeviceSetter ds(ctx->device);
    uint64_t off = 0, used = 0;
    for (uint32_t s = 0; s < n_scans; ++s) {
        R3D_TRY(scan_delta_impl(t, xyz ? xyz + off * 3 : nullptr, n_points[s], origins + 3 * (size_t)s, maxrange, discretize));
        off += n_points[s];
        counts[s] = t->delta_n;
        if (used + t->delta_n > capacity_records) {
            // report what is needed so far; the caller grows the buffer and calls again from scan s
            for (uint32_t r = s + 1; r < n_scans; ++r) counts[r] = 0;
            return set_error(ctx, R3D_ERR_OOM, "record buffer holds %llu records, scan %u needs %llu in total so far",
                             (unsigned long long)capacity_records, s, (unsigned long long)(used + t->delta_n));
        }
        if (t->delta_n)
            R3D_CUDA_OK(ctx, cudaMemcpyAsync((char*)records + used * sizeof(DeltaRecord), t->delta, t->delta_n * sizeof(DeltaRecord), cudaMemcpyDefault, ctx->stream));
        used += t->delta_n;
    }
    R3D_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return R3D_OK;
}

extern "C" int r3d_tree_apply_deltas_owned(r3d_tree* t, const void* records, const uint64_t* counts, uint32_t n_scans, uint32_t part, uint32_t nparts) {
    if (!t) return set_error(nullptr, R3D_ERR_ARG, "null tree");
    r3d_ctx* ctx = t->ctx;
    if (n_scans && (!records || !counts)) return set_error(ctx, R3D_ERR_ARG, "null argument");
    if (nparts == 0 || part >= nparts) return set_error(ctx, R3D_ERR_ARG, "bad partition %u of %u", part, nparts);
    if (n_scans && !is_device_ptr(records)) return set_error(ctx, R3D_ERR_ARG, "r3d_tree_apply_deltas_owned expects device records");
    DeviceSetter ds(ctx->device);
    const DeltaRecord* d = reinterpret_cast<const DeltaRecord*>(records);
    for (uint32_t s = 0; s < n_scans; ++s) {
        if (counts[s] > 0xfffffff0ull) return set_error(ctx, R3D_ERR_ARG, "too many records");
        R3D_TRY(apply_delta_impl(t, d, counts[s], part, nparts));
        d += counts[s];
    }
    return finish(ctx);
}

extern "C" int r3d_delta_expand_keys(const void* records_host, uint64_t n_records, uint16_t* free_keys, uint64_t free_cap,
                                     uint64_t* n_free, uint16_t* occ_keys, uint64_t occ_cap, uint64_t* n_occ) {
    if (!records_host && n_records) return set_error(nullptr, R3D_ERR_ARG, "null records");
    if (is_device_ptr(records_host)) return set_error(nullptr, R3D_ERR_ARG, "r3d_delta_expand_keys expects host memory");
    const DeltaRecord* recs = reinterpret_cast<const DeltaRecord*>(records_host);
    uint64_t nf = 0, no = 0;
    for (uint64_t r = 0; r < n_records; ++r) {
        uint32_t bx, by, bz;
        brick_key_unpack(recs[r].key, bx, by, bz);
        for (int plane = 0; plane < 2; ++plane) {
            for (uint32_t vox = 0; vox < 512; ++vox) {
                if (!(recs[r].mask[plane * 16 + (vox >> 5)] & (1u << (vox & 31u)))) continue;
                uint32_t x, y, z;
                brick_voxel_coords(vox, x, y, z);
                uint16_t* dst = plane ? free_keys : occ_keys;
                uint64_t& cnt = plane ? nf : no;
                const uint64_t cap = plane ? free_cap : occ_cap;
                if (dst && cnt < cap) { dst[3 * cnt] = (uint16_t)(bx * 8 + x); dst[3 * cnt + 1] = (uint16_t)(by * 8 + y); dst[3 * cnt + 2] = (uint16_t)(bz * 8 + z); }
                cnt++;
            }
        }
    }
    if (n_free) *n_free = nf;
    if (n_occ) *n_occ = no;
    return R3D_OK;
}

extern "C" int r3d_tree_last_scan_stats(r3d_tree* t, uint64_t out[4]) {
    if (!t || !out) return set_error(t ? t->ctx : nullptr, R3D_ERR_ARG, "null argument");
    out[0] = t->last_scan_rays; out[1] = t->last_scan_steps; out[2] = t->delta_n; out[3] = t->pool_used;
    return R3D_OK;
}

extern "C" int r3d_tree_pipeline_stats(r3d_tree* t, uint64_t out[4]) {
    if (!t || !out) return set_error(t ? t->ctx : nullptr, R3D_ERR_ARG, "null argument");
    out[0] = t->pipe_wait_ns; out[1] = t->pipe_work_ns; out[2] = t->pipe_max_turn_ns; out[3] = t->pipe_scans;
    return R3D_OK;
}

extern "C" int r3d_tree_update_inner_occupancy(r3d_tree* t) {
    if (!t) return set_error(nullptr, R3D_ERR_ARG, "null tree");
    return R3D_OK;   // inner-node values are derived from the leaves when the tree shape is needed (a12)
}

extern "C" int r3d_tree_to_max_likelihood(r3d_tree* t) {
    if (!t) return set_error(nullptr, R3D_ERR_ARG, "null tree");
    r3d_ctx* ctx = t->ctx;
    DeviceSetter ds(ctx->device);
    R3D_TRY(tree_settle(t));
    const uint64_t total = (uint64_t)t->pool_used * kBrickVoxels;
    if (total) {
        k_to_max_likelihood<<<grid_for(ctx, total), 256, 0, ctx->stream>>>(t->values, t->known, total, t->occ_thres, t->cmin, t->cmax);
        ctx->launches++;
    }
    R3D_CUDA_OK(ctx, cudaGetLastError());
    return finish(ctx);
}

extern "C" int r3d_tree_num_voxels(r3d_tree* t, uint64_t* n) {
    if (!t || !n) return set_error(t ? t->ctx : nullptr, R3D_ERR_ARG, "null argument");
    r3d_ctx* ctx = t->ctx;
    DeviceSetter ds(ctx->device);
    R3D_TRY(tree_settle(t));
    R3D_TRY(scratch_reserve(ctx, SCR_MISC, 64));
    R3D_CUDA_OK(ctx, cudaMemsetAsync(ctx->scratch[SCR_MISC], 0, 8, ctx->stream));
    const uint64_t words = (uint64_t)t->pool_used * 16;
    if (words) {
        k_count_known<<<grid_for(ctx, words), 256, 0, ctx->stream>>>(t->known, words, (unsigned long long*)ctx->scratch[SCR_MISC]);
        ctx->launches++;
    }
    unsigned long long h = 0;
    R3D_CUDA_OK(ctx, cudaMemcpyAsync(&h, ctx->scratch[SCR_MISC], 8, cudaMemcpyDeviceToHost, ctx->stream));
    R3D_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    *n = h;
    return R3D_OK;
}

extern "C" int r3d_tree_search(r3d_tree* t, const uint16_t* keys, uint64_t n, float* values, uint8_t* found) {
    if (!t) return set_error(nullptr, R3D_ERR_ARG, "null tree");
    r3d_ctx* ctx = t->ctx;
    if (!keys && n) return set_error(ctx, R3D_ERR_ARG, "null keys");
    if (n == 0) return R3D_OK;
    DeviceSetter ds(ctx->device);
    const uint16_t* dk = keys;
    R3D_TRY(stage_in(ctx, SCR_IN0, keys, (size_t)n * 3, &dk));
    const bool vdev = values && is_device_ptr(values), fdev = found && is_device_ptr(found);
    R3D_TRY(scratch_reserve(ctx, SCR_OUT0, (size_t)n * 4 + 16));
    R3D_TRY(scratch_reserve(ctx, SCR_OUT1, (size_t)n + 16));
    float* dv = vdev ? values : (float*)ctx->scratch[SCR_OUT0];
    uint8_t* df = fdev ? found : (uint8_t*)ctx->scratch[SCR_OUT1];
    k_search<<<grid_for(ctx, n), 256, 0, ctx->stream>>>(dk, n, t->tkeys, t->tvals, t->tcap, t->values, t->known, dv, df);
    ctx->launches++;
    R3D_CUDA_OK(ctx, cudaGetLastError());
    if (values && !vdev) R3D_CUDA_OK(ctx, cudaMemcpyAsync(values, dv, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (found && !fdev) R3D_CUDA_OK(ctx, cudaMemcpyAsync(found, df, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    R3D_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return R3D_OK;
}

extern "C" int r3d_tree_export_voxels(r3d_tree* t, uint16_t* keys, float* values, uint64_t cap, uint64_t* n) {
    if (!t) return set_error(nullptr, R3D_ERR_ARG, "null tree");
    r3d_ctx* ctx = t->ctx;
    DeviceSetter ds(ctx->device);
    R3D_TRY(tree_settle(t));
    R3D_TRY(tree_refresh_pool_keys(t));
    R3D_TRY(scratch_reserve(ctx, SCR_MISC, 64));
    R3D_CUDA_OK(ctx, cudaMemsetAsync(ctx->scratch[SCR_MISC], 0, 8, ctx->stream));
    uint16_t* dk = nullptr;
    float* dv = nullptr;
    if (keys && cap) { R3D_TRY(scratch_reserve(ctx, SCR_OUT0, (size_t)cap * 6 + 16)); dk = (uint16_t*)ctx->scratch[SCR_OUT0]; }
    if (values && cap) { R3D_TRY(scratch_reserve(ctx, SCR_OUT1, (size_t)cap * 4 + 16)); dv = (float*)ctx->scratch[SCR_OUT1]; }
    if (t->pool_used) {
        k_export_voxels<<<grid_for(ctx, (uint64_t)t->pool_used * kBrickVoxels), 256, 0, ctx->stream>>>(
            t->pool_keys, t->values, t->known, t->pool_used, dk, dv, cap, (unsigned long long*)ctx->scratch[SCR_MISC]);
        ctx->launches++;
    }
    unsigned long long h = 0;
    R3D_CUDA_OK(ctx, cudaMemcpyAsync(&h, ctx->scratch[SCR_MISC], 8, cudaMemcpyDeviceToHost, ctx->stream));
    R3D_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    const uint64_t m = h < cap ? h : cap;
    if (dk && m) R3D_CUDA_OK(ctx, cudaMemcpy(keys, dk, (size_t)m * 6, cudaMemcpyDefault));
    if (dv && m) R3D_CUDA_OK(ctx, cudaMemcpy(values, dv, (size_t)m * 4, cudaMemcpyDefault));
    if (n) *n = h;
    return R3D_OK;
}

extern "C" int r3d_coord_to_key(r3d_tree* t, const float* xyz, uint64_t n, uint16_t* keys, uint8_t* valid) {
    if (!t) return set_error(nullptr, R3D_ERR_ARG, "null tree");
    r3d_ctx* ctx = t->ctx;
    if ((!xyz || !keys) && n) return set_error(ctx, R3D_ERR_ARG, "null buffer");
    if (n == 0) return R3D_OK;
    DeviceSetter ds(ctx->device);
    const float* d = xyz;
    R3D_TRY(stage_in(ctx, SCR_IN0, xyz, (size_t)n * 3, &d));
    const bool kdev = is_device_ptr(keys), vdev = valid && is_device_ptr(valid);
    R3D_TRY(scratch_reserve(ctx, SCR_OUT0, (size_t)n * 6 + 16));
    R3D_TRY(scratch_reserve(ctx, SCR_OUT1, (size_t)n + 16));
    uint16_t* dk = kdev ? keys : (uint16_t*)ctx->scratch[SCR_OUT0];
    uint8_t* dv = vdev ? valid : (uint8_t*)ctx->scratch[SCR_OUT1];
    k_coord_to_key<<<grid_for(ctx, n), 256, 0, ctx->stream>>>(d, n, t->res_factor, dk, dv);
    ctx->launches++;
    R3D_CUDA_OK(ctx, cudaGetLastError());
    if (!kdev) R3D_CUDA_OK(ctx, cudaMemcpyAsync(keys, dk, (size_t)n * 6, cudaMemcpyDeviceToHost, ctx->stream));
    if (valid && !vdev) R3D_CUDA_OK(ctx, cudaMemcpyAsync(valid, dv, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    R3D_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return R3D_OK;
}
