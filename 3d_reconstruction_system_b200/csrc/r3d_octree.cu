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

#include <stdlib.h>
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
};

// computeUpdate (OccupancyOcTreeBase) for an UNBOUNDED range (maxrange < 0), or a scan whose cube cannot be direct-mapped:
// the masks of a brick sit in a per-scan hash table.  Bounded ranges -- every BASELINE configuration -- take the batched
// direct-mapped pipeline of r3d_raycast.cu.  Persistent warps; every lane walks one ray at a time and idle lanes are
// re-filled from a global ray counter as soon as K3H_REFILL_MIN of them are idle.  The walk is the branch-free form of
// computeRayKeys (r3d_math.cuh); free cells are collected in a 64-bit register mask per 4x4x4 sub-block (64 consecutive
// Morton voxels = one aligned 64-bit word of the brick's free mask) and written with ONE red.or when the ray leaves it.
constexpr int K3_THREADS = 256;
constexpr int K3H_REFILL_MIN = 8;

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
        if (!exhausted && __popc(idle) >= K3H_REFILL_MIN) {
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
        const int keep_going = exhausted ? 0 : 32 - K3H_REFILL_MIN;
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
unsigned grid_for(r3d_ctx* ctx, unsigned long long items, int block, int per_sm) {
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

static int tree_reserve_table(r3d_tree* t, uint64_t want_entries);
static int tree_reserve_pool(r3d_tree* t, uint64_t want);

int tree_reserve(r3d_tree* t, uint64_t n_bricks) {
    R3D_TRY(tree_reserve_table(t, n_bricks));
    return tree_reserve_pool(t, n_bricks);
}

int tree_flush_deferred(r3d_tree* t) {
    if (t->deferred.empty()) return R3D_OK;
    std::vector<r3d_tree::Deferred> jobs;
    jobs.swap(t->deferred);          // (apply_delta_impl may come back here through tree_settle)
    size_t n_scans = 0;
    for (const auto& j : jobs) n_scans += j.counts.size();
    static const bool sorted = !getenv("R3D_ROUND_SORTED") || atoi(getenv("R3D_ROUND_SORTED")) != 0;
    if (sorted && n_scans >= 4) return apply_round_sorted(t, jobs);   // one sorted, scan-ordered pass (r3d_round.cu)
    for (const auto& j : jobs) {
        const DeltaRecord* d = j.recs;
        for (uint64_t c : j.counts) {
            R3D_TRY(apply_delta_impl(t, d, c, j.part, j.nparts));
            d += c;
        }
    }
    return R3D_OK;
}

int tree_settle(r3d_tree* t) {
    R3D_TRY(tree_flush_deferred(t));
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
    if (t->tcap && need < (1ull << 32)) need <<= 1;   // regrowing: one doubling ahead (12 bytes per slot against 2 112 per brick)
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
        // No wait, no cudaFree (a device-wide synchronisation) here: kernels queued earlier hold the old table, everything
        // queued from now on is ordered after the re-hash by the stream.  The old arrays are retired and freed with the tree
        // (together they are smaller than the new table).
        t->retired.push_back(t->tkeys);
        t->retired.push_back(t->tvals);
        t->n_table_grow++;
    }
    t->tkeys = nk; t->tvals = nv; t->tcap = need;
    return R3D_OK;
}

// brick pool with room for `want` bricks; new bricks are zero (log-odds 0 = a freshly created node, nothing known).
// With virtual memory management the pool grows IN PLACE: one physical allocation per growth step (a multiple of 32 768
// bricks = 64 MB of log-odds + 2 MB of masks, geometric steps) is mapped behind a reserved address range; nothing is copied or
// freed and the pointers kernels hold stay valid.
constexpr uint64_t kPoolChunkBricks = 32768;
constexpr uint64_t kPoolMaxBricks = 1ull << 26;      // 137 GB of log-odds: more than a B200 holds

static int tree_reserve_pool(r3d_tree* t, uint64_t want) {
    r3d_ctx* ctx = t->ctx;
    if (want <= t->pool_cap) return R3D_OK;
    if (want > 0xfffffff0ull) return set_error(ctx, R3D_ERR_OOM, "brick pool would exceed 2^32 bricks");
    if (t->pool_cap == 0 && !t->pool_vmm && !getenv("R3D_POOL_MALLOC") && vmm_supported(ctx->device)) {
        // (a step of 32 768 bricks is 64 MB of log-odds and 2 MB of masks: both multiples of the 2 MB mapping granularity)
        if (vmm_reserve(&t->vm_values, ctx->device, kPoolMaxBricks * kBrickVoxels * sizeof(float)) &&
            vmm_reserve(&t->vm_known, ctx->device, kPoolMaxBricks * 16 * sizeof(uint32_t)) &&
            (kPoolChunkBricks * 16 * sizeof(uint32_t)) % t->vm_known.chunk == 0 && (kPoolChunkBricks * kBrickVoxels * sizeof(float)) % t->vm_values.chunk == 0) {
            t->pool_vmm = true;
            t->values = reinterpret_cast<float*>(t->vm_values.base);
            t->known = reinterpret_cast<uint32_t*>(t->vm_known.base);
        } else {
            vmm_release(&t->vm_values);
            vmm_release(&t->vm_known);
        }
    }
    if (t->pool_vmm) {
        // geometric steps (doubling) keep the number of driver calls logarithmic: each may wait for running kernels
        uint64_t ncap = 2 * t->pool_cap;
        if (ncap < want) ncap = want;
        ncap = (ncap + kPoolChunkBricks - 1) / kPoolChunkBricks * kPoolChunkBricks;
        if (ncap > kPoolMaxBricks) ncap = kPoolMaxBricks;
        if (ncap < want) return set_error(ctx, R3D_ERR_OOM, "brick pool would exceed %llu bricks", (unsigned long long)kPoolMaxBricks);
        if (!vmm_grow(&t->vm_values, ncap * kBrickVoxels * sizeof(float)) || !vmm_grow(&t->vm_known, ncap * 16 * sizeof(uint32_t))) {
            // out of device memory at the geometric step: take exactly what is asked for
            ncap = (want + kPoolChunkBricks - 1) / kPoolChunkBricks * kPoolChunkBricks;
            if (!vmm_grow(&t->vm_values, ncap * kBrickVoxels * sizeof(float)) || !vmm_grow(&t->vm_known, ncap * 16 * sizeof(uint32_t)))
                return set_error(ctx, R3D_ERR_OOM, "out of device memory growing the brick pool to %llu bricks", (unsigned long long)ncap);
        }
        const uint64_t have = t->vm_values.mapped / (kBrickVoxels * sizeof(float));
        const uint64_t used = t->pool_cap;
        R3D_CUDA_OK(ctx, cudaMemsetAsync(t->values + used * kBrickVoxels, 0, (have - used) * kBrickVoxels * sizeof(float), ctx->stream));
        R3D_CUDA_OK(ctx, cudaMemsetAsync(t->known + used * 16, 0, (have - used) * 16 * sizeof(uint32_t), ctx->stream));
        if (used) t->n_pool_grow++;
        t->pool_cap = have;
        return R3D_OK;
    }
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
    if (used) t->n_pool_grow++;
    t->values = nv; t->known = nk; t->pool_cap = ncap;
    return R3D_OK;
}

static int tree_reserve_scratch(r3d_tree* t, uint64_t want_slots) {
    r3d_ctx* ctx = t->ctx;
    uint64_t need = 1ull << 17;
    while (need < want_slots) need <<= 1;
    if (need <= t->scap) return R3D_OK;
    R3D_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(t->skeys); cudaFree(t->smasks);
    t->skeys = nullptr; t->smasks = nullptr; t->scap = 0;
    R3D_CUDA_OK(ctx, cudaMalloc(&t->skeys, need * sizeof(uint64_t)));
    R3D_CUDA_OK(ctx, cudaMalloc(&t->smasks, need * 32 * sizeof(uint32_t)));
    R3D_CUDA_OK(ctx, cudaMemsetAsync(t->skeys, 0xff, need * sizeof(uint64_t), ctx->stream));
    R3D_CUDA_OK(ctx, cudaMemsetAsync(t->smasks, 0, need * 32 * sizeof(uint32_t), ctx->stream));
    t->scap = need;
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

int tree_reserve_delta(r3d_tree* t, uint64_t want) {
    r3d_ctx* ctx = t->ctx;
    if (want <= t->delta_cap) return R3D_OK;
    R3D_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(t->delta);
    t->delta = nullptr; t->delta_cap = 0;
    R3D_CUDA_OK(ctx, cudaMalloc(&t->delta, want * sizeof(DeltaRecord)));
    t->delta_cap = want;
    return R3D_OK;
}

static unsigned raycast_blocks(r3d_tree* t) {
    if (t->raycast_blocks_per_sm_hash == 0) {
        int b = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_scan_raycast_hash, K3_THREADS, 0);
        t->raycast_blocks_per_sm_hash = b < 1 ? 1 : b;
    }
    return (unsigned)t->raycast_blocks_per_sm_hash;
}

// ray-cast one scan into t->delta (t->delta_n records).  Bounded range: the direct-mapped pipeline (a batch of one);
// unbounded range, or a cube that cannot be direct-mapped: the per-scan hash table.
static int scan_delta_impl(r3d_tree* t, const float* xyz, uint64_t n, const float origin[3], double maxrange, int discretize, bool allow_dense = true) {
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
    bool dense = allow_dense && maxrange >= 0.0;
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
            ScanSink sink;
            sink.mode = ScanSink::EXPORT_TREE;
            uint32_t done = 0;
            uint64_t rays = 0, steps = 0, n1 = a.n;
            R3D_TRY(dense_scans_run(t, a.xyz, &n1, origin, 1, maxrange, &sink, &done, &rays, &steps));
            if (done == 1) {
                R3D_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));   // the staging buffers of a host scan are reused by the next call
                t->last_scan_rays = a.n;
                t->last_scan_steps = steps;
                return R3D_OK;
            }
            dense = false;   // cannot be direct-mapped (key-space border, scratch budget): the hash table serves it
            R3D_TRY(tree_reserve_scratch(t, t->scap ? t->scap : (1ull << 18)));
            continue;
        }
        if (!overflow && a.n) {
            // persistent warps pulling rays from a counter (slot CNT_RAY_LO/HI of the tree's counter block, zeroed above)
            unsigned long long* ray_counter = reinterpret_cast<unsigned long long*>(&t->counters[CNT_RAY_LO]);
            unsigned long long blocks = (unsigned long long)ctx->sm_count * raycast_blocks(t);
            const unsigned long long need = (a.n + K3_THREADS - 1) / K3_THREADS;
            if (blocks > need) blocks = need;
            k_scan_raycast_hash<<<(unsigned)blocks, K3_THREADS, 0, ctx->stream>>>(a, ray_counter);
            ctx->launches++;
        }
        if (!overflow) {
            if (t->delta_cap < t->scap / 2 + 1) R3D_TRY(tree_reserve_delta(t, t->scap / 2 + 1));
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

int apply_delta_impl(r3d_tree* t, const DeltaRecord* d_recs, uint64_t n, uint32_t part, uint32_t nparts) {
    r3d_ctx* ctx = t->ctx;
    R3D_TRY(tree_flush_deferred(t));     // order: whatever was deferred comes first
    if (n == 0) return R3D_OK;
    // sized from the host-side upper bound of the pool cursor: no read-back between a scan's apply and the next scan.
    // When the BOUND (not necessarily the pool) would outgrow the capacity, read the exact cursor back first: several
    // applies in a row (multi-GPU rounds) inflate the bound by every record, most of which hit existing bricks.
    if (t->pool_dirty && t->pool_bound + n > t->pool_cap) {
        R3D_TRY(tree_settle(t));
        // the exact cursor is known now; leave room for several deltas of this size, so that the next applies neither wait
        // for a read-back nor regrow (a regrowth copies the pool)
        R3D_TRY(tree_reserve_table(t, t->pool_bound + 8 * n));
        R3D_TRY(tree_reserve_pool(t, t->pool_bound + 8 * n));
    }
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
    cudaFree(t->tkeys); cudaFree(t->tvals); cudaFree(t->pool_keys);
    for (void* q : t->retired) cudaFree(q);
    if (t->pool_vmm) { vmm_release(&t->vm_values); vmm_release(&t->vm_known); }
    else { cudaFree(t->values); cudaFree(t->known); }
    cudaFree(t->skeys); cudaFree(t->smasks); cudaFree(t->delta); cudaFree(t->counters);
    if (t->cast_gate) cudaEventDestroy(t->cast_gate);
    delete t;
}

extern "C" int r3d_tree_clear(r3d_tree* t) {
    if (!t) return set_error(nullptr, R3D_ERR_ARG, "null tree");
    r3d_ctx* ctx = t->ctx;
    t->deferred.clear();
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

// Serial remainder of a batch call: scans the pipeline cannot take (unbounded range, host buffers, discretize, a cube that
// cannot be direct-mapped), interleaved with pipelined runs of the scans it can.
static int scans_run(r3d_tree* t, const float* xyz, const uint64_t* n_points, const float* origins, uint32_t n_scans, double maxrange,
                     int discretize, ScanSink* sink, uint64_t* rays_out, uint64_t* steps_out) {
    r3d_ctx* ctx = t->ctx;
    const bool pipelined = !discretize && maxrange >= 0.0 && xyz && is_device_ptr(xyz);
    uint64_t off = 0;
    uint32_t s = 0;
    while (s < n_scans) {
        if (pipelined) {
            uint32_t done = 0;
            ScanSink sub = *sink;
            if (sub.counts) sub.counts += s;
            R3D_TRY(dense_scans_run(t, xyz + off * 3, n_points + s, origins + 3 * (size_t)s, n_scans - s, maxrange, &sub, &done, rays_out, steps_out));
            sink->used = sub.used;
            for (uint32_t k = 0; k < done; ++k) off += n_points[s + k];
            s += done;
            if (s >= n_scans) break;
        }
        // one scan through the serial path (hash table when it must be)
        R3D_TRY(scan_delta_impl(t, xyz ? xyz + off * 3 : nullptr, n_points[s], origins + 3 * (size_t)s, maxrange, discretize, !pipelined));   // (the pipeline has just declined this scan: hash table)
        if (sink->mode == ScanSink::APPLY) {
            R3D_TRY(apply_delta_impl(t, t->delta, t->delta_n));
        } else if (sink->mode == ScanSink::EXPORT_USER) {
            sink->counts[s] = t->delta_n;
            if (sink->used + t->delta_n > sink->capacity) {
                for (uint32_t r = s + 1; r < n_scans; ++r) sink->counts[r] = 0;
                return set_error(ctx, R3D_ERR_OOM, "record buffer holds %llu records, scan %u needs %llu in total so far",
                                 (unsigned long long)sink->capacity, s, (unsigned long long)(sink->used + t->delta_n));
            }
            if (t->delta_n)
                R3D_CUDA_OK(ctx, cudaMemcpyAsync((char*)sink->records + sink->used * sizeof(DeltaRecord), t->delta, t->delta_n * sizeof(DeltaRecord), cudaMemcpyDefault, ctx->stream));
            sink->used += t->delta_n;
        }
        *rays_out += t->last_scan_rays;
        *steps_out += t->last_scan_steps;
        off += n_points[s];
        ++s;
    }
    return R3D_OK;
}

extern "C" int r3d_tree_insert_scans(r3d_tree* t, const float* xyz, const uint64_t* n_points, const float* origins, uint32_t n_scans,
                                     double maxrange, int discretize) {
    if (!t) return set_error(nullptr, R3D_ERR_ARG, "null tree");
    r3d_ctx* ctx = t->ctx;
    if (n_scans && (!n_points || !origins)) return set_error(ctx, R3D_ERR_ARG, "null scan table");
    if (is_device_ptr(n_points) || is_device_ptr(origins)) return set_error(ctx, R3D_ERR_ARG, "n_points / origins must be host arrays");
    DeviceSetter ds(ctx->device);
    uint64_t rays = 0, steps = 0;
    ScanSink sink;
    sink.mode = ScanSink::APPLY;
    R3D_TRY(scans_run(t, xyz, n_points, origins, n_scans, maxrange, discretize, &sink, &rays, &steps));
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
    // the deltas noted by r3d_tree_defer_deltas_owned (the round before this one, every rank's share): ONE sorted,
    // scan-ordered pass, queued now that the stream is empty (its two small read-backs cost nothing here)
    // -- and the ray casting of THIS round does not depend on it (it reads the scans, not the tree): the slots' streams wait
    // for the point of the stream before the pass (cast_gate), so the walkers start beside the apply kernel
    // (device-resident scans only: a host scan is staged on the context stream AFTER this point)
    if (!t->deferred.empty() && xyz && is_device_ptr(xyz) && !discretize) {
        if (!t->cast_gate) R3D_CUDA_OK(ctx, cudaEventCreateWithFlags(&t->cast_gate, cudaEventDisableTiming));
        R3D_CUDA_OK(ctx, cudaEventRecord(t->cast_gate, ctx->stream));
        static const bool gate_on = !getenv("R3D_CAST_GATE") || atoi(getenv("R3D_CAST_GATE")) != 0;
        t->cast_gate_armed = gate_on;
    }
    int rc = tree_flush_deferred(t);
    uint64_t rays = 0, steps = 0;
    ScanSink sink;
    sink.mode = ScanSink::EXPORT_USER;
    sink.records = records; sink.capacity = capacity_records; sink.counts = counts;
    if (rc == R3D_OK) rc = scans_run(t, xyz, n_points, origins, n_scans, maxrange, discretize, &sink, &rays, &steps);
    t->cast_gate_armed = false;
    R3D_TRY(rc);
    t->last_scan_rays = rays;
    t->last_scan_steps = steps;
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

extern "C" int r3d_tree_defer_deltas_owned(r3d_tree* t, const void* records, const uint64_t* counts, uint32_t n_scans, uint32_t part, uint32_t nparts) {
    if (!t) return set_error(nullptr, R3D_ERR_ARG, "null tree");
    r3d_ctx* ctx = t->ctx;
    if (n_scans && (!records || !counts)) return set_error(ctx, R3D_ERR_ARG, "null argument");
    if (nparts == 0 || part >= nparts) return set_error(ctx, R3D_ERR_ARG, "bad partition %u of %u", part, nparts);
    if (n_scans && !is_device_ptr(records)) return set_error(ctx, R3D_ERR_ARG, "r3d_tree_defer_deltas_owned expects device records");
    r3d_tree::Deferred job;
    job.recs = reinterpret_cast<const DeltaRecord*>(records);
    job.part = part; job.nparts = nparts;
    for (uint32_t s = 0; s < n_scans; ++s) {
        if (counts[s] > 0xfffffff0ull) return set_error(ctx, R3D_ERR_ARG, "too many records");
        job.counts.push_back(counts[s]);
    }
    if (n_scans) t->deferred.push_back(std::move(job));
    return R3D_OK;
}

extern "C" int r3d_tree_flush_deferred(r3d_tree* t) {
    if (!t) return set_error(nullptr, R3D_ERR_ARG, "null tree");
    DeviceSetter ds(t->ctx->device);
    R3D_TRY(tree_flush_deferred(t));
    return finish(t->ctx);
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
    out[0] = t->last_scan_rays; out[1] = t->last_scan_steps; out[2] = t->delta_n ? t->delta_n : t->last_batch_records; out[3] = t->pool_used;
    return R3D_OK;
}

extern "C" int r3d_tree_pipeline_stats(r3d_tree* t, uint64_t out[4]) {
    if (!t || !out) return set_error(t ? t->ctx : nullptr, R3D_ERR_ARG, "null argument");
    out[0] = t->pipe_wait_ns; out[1] = t->pipe_work_ns; out[2] = t->pipe_max_turn_ns; out[3] = t->pipe_scans;
    return R3D_OK;
}

extern "C" int r3d_tree_growth_stats(r3d_tree* t, uint64_t out[4]) {
    if (!t || !out) return set_error(t ? t->ctx : nullptr, R3D_ERR_ARG, "null argument");
    out[0] = t->n_pool_grow; out[1] = t->n_table_grow; out[2] = t->pool_cap; out[3] = t->tcap;
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
    R3D_TRY(tree_flush_deferred(t));
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
