// r3d_bt.cu -- K5: tree shape, size() and the .bt byte stream, derived on the GPU from the flat voxel store.
//
// Replaces OcTree.writeBinary / size() of the `octomap` extension (octomap/txt_transfer_octomap.py:36,
// octomap/ply_transfer_octomap.py:48): toMaxLikelihood -> prune -> header -> pre-order 2-bits-per-child stream.
//
// Upstream's tree after any sequence of non-lazy updateNode calls is maximally pruned under EXACT float equality
// of sibling leaves (every update re-tests all ancestors), so its shape is a pure function of the leaf values:
// "S0" below.  writeBinary then thresholds every node and runs prune(): passes over depth 15, 14, ... 1 that
// collapse nodes whose 8 children are leaves with the same occupancy, STOPPING at the first pass that prunes
// nothing.  Both rules are reproduced level by level: three levels inside each brick (one warp per brick), then
// one small kernel per level above over the Morton-sorted brick list.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <string>
#include <utility>
#include <vector>

#include "r3d_octree.cuh"

namespace r3d {

struct UpNode {            // a node at depth <= 13
    uint64_t prefix;       // Morton code of the node (3 bits per level, child-index convention)
    uint64_t node_cnt;     // nodes in the subtree, this one included
    uint64_t offset;       // pre-order index among INNER nodes (~0: not emitted)
    float s0val;           // common log-odds when S0-collapsed
    uint32_t inner_cnt;    // inner nodes in the subtree, this one included (0 for a leaf)
    uint32_t first_child;  // index of the first child in the next level's array
    uint16_t mask;         // the 2 bytes writeBinaryNode emits for this node
    uint8_t flags;         // bit0 leaf, bit1 occupied, bit2 S0 leaf
    uint8_t nchild;
};
enum { F_LEAF = 1, F_OCC = 2, F_S0 = 4 };

struct BtCounters {
    unsigned long long pruned[17];   // nodes collapsed by the max-likelihood pass at each depth
};

__device__ __forceinline__ uint32_t spread4(uint32_t x) { return (x & 1u) | ((x & 2u) << 1) | ((x & 4u) << 2) | ((x & 8u) << 3); }
// children c = 0..7 of 14-node k live in lanes 4k + c/2, half c&1
__device__ __forceinline__ uint32_t gather8(uint32_t b0, uint32_t b1, uint32_t k) {
    return spread4((b0 >> (4 * k)) & 0xfu) | (spread4((b1 >> (4 * k)) & 0xfu) << 1);
}

// Per-brick structure, computed cooperatively by one warp.  ran15 / ran14 / ran13: whether the max-likelihood prune
// passes at depth 15 / 14 / 13 run (all false: the value-pruned shape only, for size()).
struct BrickShape {
    // per lane: its two depth-15 nodes
    uint32_t kn[2], oc[2];         // known / occupied bits of the 8 voxels
    bool exists15[2], leaf15[2], occ15[2], s0leaf15[2], mlc15[2];
    // per lane, about the depth-14 node k = lane / 4 (identical in the 4 lanes of a group)
    bool exists14, leaf14, occ14, s0leaf14, mlc14;
    uint32_t ex15_8, leaf15_8, occ15_8;   // 8-bit child summaries of node k
    // brick root (identical in all lanes)
    bool leaf13, occ13, s0leaf13, mlc13;
    uint32_t ex14_8, leaf14_8, occ14_8;
    float s0val;                  // lane 0's first value (the S0 value when s0leaf13)
};

__device__ __forceinline__ void brick_shape(const float* __restrict__ values, const uint32_t* __restrict__ known, float thres,
                                            bool ran15, bool ran14, bool ran13, BrickShape& s) {
    const unsigned lane = threadIdx.x & 31u;
    const float4* v4 = reinterpret_cast<const float4*>(values + lane * 16);
    float v[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 t = v4[q];
        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
    const uint32_t kn16 = (known[lane >> 1] >> ((lane & 1u) * 16u)) & 0xffffu;
    float first[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint32_t kn = (kn16 >> (8 * h)) & 0xffu;
        uint32_t oc = 0;
        bool alleq = true;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            if (v[8 * h + b] >= thres) oc |= 1u << b;
            alleq = alleq && (v[8 * h + b] == v[8 * h]);
        }
        oc &= kn;
        s.kn[h] = kn; s.oc[h] = oc;
        s.exists15[h] = kn != 0;
        const bool full = kn == 0xffu;
        s.s0leaf15[h] = full && alleq;
        s.mlc15[h] = ran15 && !s.s0leaf15[h] && full && (oc == 0xffu || oc == 0u);
        s.leaf15[h] = s.s0leaf15[h] || s.mlc15[h];
        s.occ15[h] = (oc & 1u) != 0;
        first[h] = v[8 * h];
    }
    const uint32_t k = lane >> 2;
    const uint32_t ex0 = __ballot_sync(0xffffffffu, s.exists15[0]), ex1 = __ballot_sync(0xffffffffu, s.exists15[1]);
    const uint32_t lf0 = __ballot_sync(0xffffffffu, s.leaf15[0]), lf1 = __ballot_sync(0xffffffffu, s.leaf15[1]);
    const uint32_t oc0 = __ballot_sync(0xffffffffu, s.occ15[0]), oc1 = __ballot_sync(0xffffffffu, s.occ15[1]);
    const uint32_t s00 = __ballot_sync(0xffffffffu, s.s0leaf15[0]), s01 = __ballot_sync(0xffffffffu, s.s0leaf15[1]);
    const float ref14 = __shfl_sync(0xffffffffu, first[0], lane & ~3u);
    const uint32_t eq0 = __ballot_sync(0xffffffffu, first[0] == ref14), eq1 = __ballot_sync(0xffffffffu, first[1] == ref14);
    s.ex15_8 = gather8(ex0, ex1, k);
    s.leaf15_8 = gather8(lf0, lf1, k);
    s.occ15_8 = gather8(oc0, oc1, k);
    const uint32_t s0_8 = gather8(s00, s01, k), eq_8 = gather8(eq0, eq1, k);
    s.exists14 = s.ex15_8 != 0;
    s.s0leaf14 = (s0_8 == 0xffu) && (eq_8 == 0xffu);
    s.mlc14 = ran14 && !s.s0leaf14 && (s.ex15_8 == 0xffu) && (s.leaf15_8 == 0xffu) && (s.occ15_8 == 0xffu || s.occ15_8 == 0u);
    s.leaf14 = s.s0leaf14 || s.mlc14;
    s.occ14 = (s.occ15_8 & 1u) != 0;
    // brick root: one representative lane per group (lane % 4 == 0) votes
    const bool rep = (lane & 3u) == 0;
    const float ref13 = __shfl_sync(0xffffffffu, first[0], 0);
    auto pick8 = [&](bool p) {   // bit k = predicate of group k
        const uint32_t b = __ballot_sync(0xffffffffu, rep && p);
        uint32_t r = 0;
#pragma unroll
        for (int g = 0; g < 8; ++g) r |= ((b >> (4 * g)) & 1u) << g;
        return r;
    };
    s.ex14_8 = pick8(s.exists14);
    s.leaf14_8 = pick8(s.leaf14);
    s.occ14_8 = pick8(s.occ14);
    const uint32_t s014_8 = pick8(s.s0leaf14), eq14_8 = pick8(ref14 == ref13);
    s.s0leaf13 = (s014_8 == 0xffu) && (eq14_8 == 0xffu);
    s.mlc13 = ran13 && !s.s0leaf13 && (s.ex14_8 == 0xffu) && (s.leaf14_8 == 0xffu) && (s.occ14_8 == 0xffu || s.occ14_8 == 0u);
    s.leaf13 = s.s0leaf13 || s.mlc13;
    s.occ13 = (s.occ14_8 & 1u) != 0;
    s.s0val = ref13;
}

__device__ __forceinline__ uint32_t group4_sum(uint32_t x) {
    x += __shfl_xor_sync(0xffffffffu, x, 1);
    x += __shfl_xor_sync(0xffffffffu, x, 2);
    return x;
}
__device__ __forceinline__ uint32_t warp_sum(uint32_t x) {
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
__device__ __forceinline__ uint16_t child_codes(uint32_t exists8, uint32_t leaf8, uint32_t occ8) {
    uint32_t m = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        if (!((exists8 >> c) & 1u)) continue;
        const uint32_t code = ((leaf8 >> c) & 1u) ? (((occ8 >> c) & 1u) ? 2u : 1u) : 3u;
        m |= code << (2 * c);
    }
    return (uint16_t)m;
}

// pass A: how many nodes would the max-likelihood passes at depth 15 / 14 / 13 collapse (assuming each runs)
__global__ void __launch_bounds__(256) k_bt_count(const float* values, const uint32_t* known, uint32_t n_bricks, float thres,
                                                  BtCounters* cnt) {
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    uint32_t c15 = 0, c14 = 0, c13 = 0;
    for (uint32_t b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < n_bricks; b += warps) {
        BrickShape s;
        brick_shape(values + (size_t)b * kBrickVoxels, known + (size_t)b * 16, thres, true, true, true, s);
        c15 += (s.mlc15[0] ? 1u : 0u) + (s.mlc15[1] ? 1u : 0u);
        if ((lane & 3u) == 0 && s.mlc14) c14++;
        if (lane == 0 && s.mlc13) c13++;
    }
    c15 = warp_sum(c15); c14 = warp_sum(c14); c13 = warp_sum(c13);
    if (lane == 0) {
        if (c15) atomicAdd(&cnt->pruned[15], (unsigned long long)c15);
        if (c14) atomicAdd(&cnt->pruned[14], (unsigned long long)c14);
        if (c13) atomicAdd(&cnt->pruned[13], (unsigned long long)c13);
    }
}

// pass B: the depth-13 node of every brick, in Morton order
__global__ void __launch_bounds__(256) k_bt_bricks(const float* values, const uint32_t* known, const uint32_t* order,
                                                   const uint64_t* sorted_morton, uint32_t n_bricks, float thres, int ran15, int ran14,
                                                   int ran13, UpNode* out) {
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n_bricks; i += warps) {
        const uint32_t b = order[i];
        BrickShape s;
        brick_shape(values + (size_t)b * kBrickVoxels, known + (size_t)b * 16, thres, ran15 != 0, ran14 != 0, ran13 != 0, s);
        uint32_t nodes15 = 0, inner15 = 0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (!s.exists15[h]) continue;
            nodes15 += 1u + (s.leaf15[h] ? 0u : (uint32_t)__popc(s.kn[h]));
            inner15 += s.leaf15[h] ? 0u : 1u;
        }
        const uint32_t nodes_g = group4_sum(nodes15), inner_g = group4_sum(inner15);
        uint32_t nodes14 = 0, inner14 = 0;
        if ((lane & 3u) == 0 && s.exists14) {
            nodes14 = 1u + (s.leaf14 ? 0u : nodes_g);
            inner14 = s.leaf14 ? 0u : 1u + inner_g;
        }
        const uint32_t nodes_b = warp_sum(nodes14), inner_b = warp_sum(inner14);
        if (lane == 0) {
            UpNode n;
            n.prefix = sorted_morton[i];
            n.node_cnt = 1ull + (s.leaf13 ? 0ull : (unsigned long long)nodes_b);
            n.inner_cnt = s.leaf13 ? 0u : 1u + inner_b;
            n.offset = ~0ull;
            n.s0val = s.s0val;
            n.first_child = b;   // pool index of the brick
            n.mask = child_codes(s.ex14_8, s.leaf14_8, s.occ14_8);
            n.flags = (uint8_t)((s.leaf13 ? F_LEAF : 0) | (s.occ13 ? F_OCC : 0) | (s.s0leaf13 ? F_S0 : 0));
            n.nchild = (uint8_t)__popc(s.ex14_8);
            out[i] = n;
        }
    }
}

// pass C: bytes of the inner nodes inside every emitted brick
__global__ void __launch_bounds__(256) k_bt_emit_bricks(const float* values, const uint32_t* known, const UpNode* nodes,
                                                        uint32_t n_bricks, float thres, int ran15, int ran14, int ran13, uint16_t* out) {
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n_bricks; i += warps) {
        const UpNode nd = nodes[i];
        if (nd.offset == ~0ull || (nd.flags & F_LEAF)) continue;
        const uint32_t b = nd.first_child;
        BrickShape s;
        brick_shape(values + (size_t)b * kBrickVoxels, known + (size_t)b * 16, thres, ran15 != 0, ran14 != 0, ran13 != 0, s);
        if (lane == 0) out[nd.offset] = nd.mask;
        // inner nodes per 14-node subtree (1 + inner 15-children), exclusive prefix over k
        const uint32_t in15 = ((s.exists15[0] && !s.leaf15[0]) ? 1u : 0u) + ((s.exists15[1] && !s.leaf15[1]) ? 1u : 0u);
        const uint32_t in15_g = group4_sum(in15);
        const bool inner14 = s.exists14 && !s.leaf14;
        const uint32_t sub14 = inner14 ? 1u + in15_g : 0u;       // same in the 4 lanes of the group
        const uint32_t k = lane >> 2;
        uint32_t before = 0;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const uint32_t sg = __shfl_sync(0xffffffffu, sub14, 4 * g);
            if ((uint32_t)g < k) before += sg;
        }
        const unsigned long long off14 = nd.offset + 1ull + before;
        if (inner14 && (lane & 3u) == 0) out[off14] = child_codes(s.ex15_8, s.leaf15_8, s.occ15_8);
        // inner 15-nodes of this group, in child order c = 2*(lane%4) + h
        const uint32_t inner_bits = s.ex15_8 & ~s.leaf15_8;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (!(inner14 && s.exists15[h] && !s.leaf15[h])) continue;
            const uint32_t c = 2u * (lane & 3u) + h;
            const unsigned long long off15 = off14 + 1ull + __popc(inner_bits & ((1u << c) - 1u));
            uint32_t m = 0;
#pragma unroll
            for (int v = 0; v < 8; ++v)
                if ((s.kn[h] >> v) & 1u) m |= (((s.oc[h] >> v) & 1u) ? 2u : 1u) << (2 * v);
            out[off15] = (uint16_t)m;
        }
    }
}

// heads of sibling groups in a sorted level
__global__ void k_bt_heads(const UpNode* child, uint32_t n, uint32_t* heads) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    heads[i] = (i == 0 || (child[i].prefix >> 3) != (child[i - 1].prefix >> 3)) ? 1u : 0u;
}

// one parent per sibling group
__global__ void k_bt_parents(const UpNode* child, uint32_t n, const uint32_t* heads, const uint32_t* incl, UpNode* parent, int depth,
                             int ran, BtCounters* cnt) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !heads[i]) return;
    const uint64_t pp = child[i].prefix >> 3;
    UpNode p;
    p.prefix = pp;
    p.first_child = i;
    p.offset = ~0ull;
    uint32_t exists8 = 0, leaf8 = 0, occ8 = 0, s08 = 0, eq8 = 0, nch = 0;
    unsigned long long nodes = 0;
    uint32_t inner = 0;
    const float ref = child[i].s0val;
    for (uint32_t c = i; c < n && (child[c].prefix >> 3) == pp && nch < 8; ++c, ++nch) {
        const UpNode ch = child[c];
        const uint32_t ci = (uint32_t)(ch.prefix & 7u);
        exists8 |= 1u << ci;
        if (ch.flags & F_LEAF) leaf8 |= 1u << ci;
        if (ch.flags & F_OCC) occ8 |= 1u << ci;
        if (ch.flags & F_S0) s08 |= 1u << ci;
        if (ch.s0val == ref) eq8 |= 1u << ci;
        nodes += ch.node_cnt;
        inner += ch.inner_cnt;
    }
    const bool s0leaf = (s08 == 0xffu) && (eq8 == 0xffu);
    const bool mlc = ran && depth > 0 && !s0leaf && exists8 == 0xffu && leaf8 == 0xffu && (occ8 == 0xffu || occ8 == 0u);
    const bool leaf = s0leaf || mlc;
    if (mlc) atomicAdd(&cnt->pruned[depth], 1ull);
    const bool occ = leaf ? ((occ8 >> (child[i].prefix & 7u)) & 1u) != 0 : occ8 != 0;
    p.flags = (uint8_t)((leaf ? F_LEAF : 0) | (occ ? F_OCC : 0) | (s0leaf ? F_S0 : 0));
    p.s0val = ref;
    p.nchild = (uint8_t)nch;
    p.node_cnt = 1ull + (leaf ? 0ull : nodes);
    p.inner_cnt = leaf ? 0u : 1u + inner;
    p.mask = child_codes(exists8, leaf8, occ8);
    parent[incl[i] - 1] = p;
}

// pre-order offsets of the children of every emitted inner node
__global__ void k_bt_offsets(const UpNode* parent, uint32_t n_parent, UpNode* child, uint16_t* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_parent) return;
    const UpNode p = parent[i];
    if (p.offset == ~0ull || (p.flags & F_LEAF)) return;
    out[p.offset] = p.mask;
    unsigned long long run = p.offset + 1ull;
    for (uint32_t c = 0; c < p.nchild; ++c) {
        UpNode& ch = child[p.first_child + c];
        if (!(ch.flags & F_LEAF)) { ch.offset = run; run += ch.inner_cnt; }
    }
}

__global__ void k_bt_morton(const uint64_t* pool_keys, uint32_t n, uint64_t* morton, uint32_t* idx) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    morton[i] = brick_morton(pool_keys[i]);
    idx[i] = i;
}

__global__ void k_bt_set_root_offset(UpNode* root) { root->offset = 0; }

// The serialiser's temporaries come out of one context-level block (SCR_BT, grow-only): two dozen cudaMalloc / cudaFree
// pairs per call, each a device-wide synchronisation, were most of the 80 ms a 5.8 MB .bt took (and the odd full second).
// A piece that does not fit the block falls back to its own allocation.
struct BtArena {
    char* base = nullptr;
    size_t cap = 0, used = 0;
};
struct DevBuf {
    void* p = nullptr;
    bool owned = false;
    BtArena* arena = nullptr;
    ~DevBuf() { if (p && owned) cudaFree(p); }
    template <typename T> T* as() { return reinterpret_cast<T*>(p); }
    cudaError_t alloc(size_t bytes) {
        bytes = (bytes ? bytes : 16) + 255 & ~(size_t)255;
        if (arena && arena->used + bytes <= arena->cap) {
            p = arena->base + arena->used;
            arena->used += bytes;
            return cudaSuccess;
        }
        owned = true;
        return cudaMalloc(&p, bytes);
    }
};

static unsigned blocks_for(uint64_t n, int block = 256) { return (unsigned)((n + block - 1) / block > 0 ? (n + block - 1) / block : 1); }

// Derives the tree shape.  ml = false: value-pruned shape only (size()).  payload may be null.
static int tree_shape(r3d_tree* t, bool ml, uint64_t* n_nodes, std::vector<uint8_t>* payload) {
    r3d_ctx* ctx = t->ctx;
    *n_nodes = 0;
    if (payload) payload->clear();
    R3D_TRY(tree_settle(t));
    const uint32_t nb = t->pool_used;
    if (nb == 0) return R3D_OK;
    R3D_TRY(tree_refresh_pool_keys(t));
    cudaStream_t st = ctx->stream;
    size_t tmp_sort = 0, tmp_scan = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)nb, 0, 39, st);
    cub::DeviceScan::InclusiveSum(nullptr, tmp_scan, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)nb, st);
    BtArena arena;
    {
        // sort keys / indices / heads / sums (32 B per brick), level 13 and the (geometrically shrinking) levels above
        // it (40 B per node), the CUB scratch, and room for a typical payload (2 B per inner node, ~9 B per brick measured)
        const size_t want = (size_t)nb * (32 + 60 + 16) + tmp_sort + tmp_scan + (1u << 20);
        if (scratch_reserve(ctx, SCR_BT, want) == R3D_OK) {
            arena.base = (char*)ctx->scratch[SCR_BT];
            arena.cap = ctx->scratch_bytes[SCR_BT];
        }
    }
    DevBuf morton_in, morton_out, idx_in, idx_out, cub_tmp, counters, heads, incl;
    for (DevBuf* b : {&morton_in, &morton_out, &idx_in, &idx_out, &cub_tmp, &counters, &heads, &incl}) b->arena = &arena;
    R3D_CUDA_OK(ctx, morton_in.alloc((size_t)nb * 8));
    R3D_CUDA_OK(ctx, morton_out.alloc((size_t)nb * 8));
    R3D_CUDA_OK(ctx, idx_in.alloc((size_t)nb * 4));
    R3D_CUDA_OK(ctx, idx_out.alloc((size_t)nb * 4));
    R3D_CUDA_OK(ctx, counters.alloc(sizeof(BtCounters)));
    R3D_CUDA_OK(ctx, heads.alloc((size_t)nb * 4));
    R3D_CUDA_OK(ctx, incl.alloc((size_t)nb * 4));
    R3D_CUDA_OK(ctx, cudaMemsetAsync(counters.p, 0, sizeof(BtCounters), st));
    k_bt_morton<<<blocks_for(nb), 256, 0, st>>>(t->pool_keys, nb, morton_in.as<uint64_t>(), idx_in.as<uint32_t>());
    ctx->launches++;
    R3D_CUDA_OK(ctx, cub_tmp.alloc((tmp_sort > tmp_scan ? tmp_sort : tmp_scan) + 256));
    R3D_CUDA_OK(ctx, cub::DeviceRadixSort::SortPairs(cub_tmp.p, tmp_sort, morton_in.as<uint64_t>(), morton_out.as<uint64_t>(),
                                                     idx_in.as<uint32_t>(), idx_out.as<uint32_t>(), (int)nb, 0, 39, st));
    ctx->launches++;

    const float thres = ml ? t->occ_thres : 0.f;
    BtCounters hc;
    memset(&hc, 0, sizeof hc);
    bool ran[17] = {false};
    const unsigned brick_grid = blocks_for((uint64_t)nb * 32) < (unsigned)ctx->sm_count * 8 ? blocks_for((uint64_t)nb * 32) : (unsigned)ctx->sm_count * 8;
    if (ml) {
        k_bt_count<<<brick_grid, 256, 0, st>>>(t->values, t->known, nb, thres, counters.as<BtCounters>());
        ctx->launches++;
        R3D_CUDA_OK(ctx, cudaMemcpyAsync(&hc, counters.p, sizeof hc, cudaMemcpyDeviceToHost, st));
        R3D_CUDA_OK(ctx, cudaStreamSynchronize(st));
        ran[15] = true;
        ran[14] = hc.pruned[15] > 0;
        ran[13] = ran[14] && hc.pruned[14] > 0;
        ran[12] = ran[13] && hc.pruned[13] > 0;
    }
    // levels 13 .. 0
    std::vector<DevBuf> level(14);
    for (DevBuf& b : level) b.arena = &arena;
    std::vector<uint32_t> count(14, 0);
    count[13] = nb;
    R3D_CUDA_OK(ctx, level[13].alloc((size_t)nb * sizeof(UpNode)));
    if (ml) {
        k_bt_bricks<<<brick_grid, 256, 0, st>>>(t->values, t->known, idx_out.as<uint32_t>(), morton_out.as<uint64_t>(), nb, thres, 1,
                                                ran[14] ? 1 : 0, ran[13] ? 1 : 0, level[13].as<UpNode>());
    } else {
        // size(): exact-equality leaves only, no max-likelihood pass at any depth
        k_bt_bricks<<<brick_grid, 256, 0, st>>>(t->values, t->known, idx_out.as<uint32_t>(), morton_out.as<uint64_t>(), nb, thres, 0, 0, 0,
                                                level[13].as<UpNode>());
    }
    ctx->launches++;
    for (int d = 12; d >= 0; --d) {
        const uint32_t nc = count[d + 1];
        k_bt_heads<<<blocks_for(nc), 256, 0, st>>>(level[d + 1].as<UpNode>(), nc, heads.as<uint32_t>());
        R3D_CUDA_OK(ctx, cub::DeviceScan::InclusiveSum(cub_tmp.p, tmp_scan, heads.as<uint32_t>(), incl.as<uint32_t>(), (int)nc, st));
        uint32_t np = 0;
        R3D_CUDA_OK(ctx, cudaMemcpyAsync(&np, incl.as<uint32_t>() + (nc - 1), 4, cudaMemcpyDeviceToHost, st));
        R3D_CUDA_OK(ctx, cudaStreamSynchronize(st));
        count[d] = np;
        R3D_CUDA_OK(ctx, level[d].alloc((size_t)np * sizeof(UpNode)));
        k_bt_parents<<<blocks_for(nc), 256, 0, st>>>(level[d + 1].as<UpNode>(), nc, heads.as<uint32_t>(), incl.as<uint32_t>(),
                                                     level[d].as<UpNode>(), d, ran[d] ? 1 : 0, counters.as<BtCounters>());
        ctx->launches += 3;
        if (ml && d > 0) {
            R3D_CUDA_OK(ctx, cudaMemcpyAsync(&hc, counters.p, sizeof hc, cudaMemcpyDeviceToHost, st));
            R3D_CUDA_OK(ctx, cudaStreamSynchronize(st));
            ran[d - 1] = ran[d] && hc.pruned[d] > 0;
        }
    }
    if (count[0] != 1) return set_error(ctx, R3D_ERR_STATE, "tree shape: %u roots", count[0]);
    UpNode root;
    R3D_CUDA_OK(ctx, cudaMemcpyAsync(&root, level[0].p, sizeof root, cudaMemcpyDeviceToHost, st));
    R3D_CUDA_OK(ctx, cudaStreamSynchronize(st));
    *n_nodes = root.node_cnt;
    if (!payload) return R3D_OK;
    if (root.flags & F_LEAF) {   // a fully collapsed root still writes its (empty) child bytes
        payload->assign(2, 0);
        return R3D_OK;
    }
    const uint64_t n_inner = root.inner_cnt;
    DevBuf out;
    out.arena = &arena;
    R3D_CUDA_OK(ctx, out.alloc((size_t)n_inner * 2));
    k_bt_set_root_offset<<<1, 1, 0, st>>>(level[0].as<UpNode>());
    for (int d = 0; d <= 12; ++d) {
        k_bt_offsets<<<blocks_for(count[d]), 256, 0, st>>>(level[d].as<UpNode>(), count[d], level[d + 1].as<UpNode>(), out.as<uint16_t>());
        ctx->launches++;
    }
    k_bt_emit_bricks<<<brick_grid, 256, 0, st>>>(t->values, t->known, level[13].as<UpNode>(), nb, thres, ml ? 1 : 0, ran[14] ? 1 : 0,
                                                 ran[13] ? 1 : 0, out.as<uint16_t>());
    ctx->launches++;
    R3D_CUDA_OK(ctx, cudaGetLastError());
    payload->resize((size_t)n_inner * 2);
    R3D_CUDA_OK(ctx, cudaMemcpyAsync(payload->data(), out.p, (size_t)n_inner * 2, cudaMemcpyDeviceToHost, st));
    R3D_CUDA_OK(ctx, cudaStreamSynchronize(st));
    return R3D_OK;
}

static std::string bt_header(uint64_t n_nodes, double res) {
    char buf[512];
    snprintf(buf, sizeof buf,
             "# Octomap OcTree binary file\n# (feel free to add / change comments, but leave the first line as it is!)\n#\n"
             "id OcTree\nsize %llu\nres %g\ndata\n",
             (unsigned long long)n_nodes, res);
    return std::string(buf);
}

}  // namespace r3d

using namespace r3d;

extern "C" int r3d_tree_size(r3d_tree* t, uint64_t* n_nodes) {
    if (!t || !n_nodes) return set_error(t ? t->ctx : nullptr, R3D_ERR_ARG, "null argument");
    DeviceSetter ds(t->ctx->device);
    return tree_shape(t, false, n_nodes, nullptr);
}

extern "C" int r3d_tree_write_bt_mem(r3d_tree* t, uint8_t* buf, size_t cap, size_t* len) {
    if (!t || !len) return set_error(t ? t->ctx : nullptr, R3D_ERR_ARG, "null argument");
    DeviceSetter ds(t->ctx->device);
    uint64_t n_nodes = 0;
    std::vector<uint8_t> payload;
    R3D_TRY(tree_shape(t, true, &n_nodes, &payload));
    const std::string hdr = bt_header(n_nodes, t->res);
    *len = hdr.size() + payload.size();
    if (buf && cap >= *len) {
        memcpy(buf, hdr.data(), hdr.size());
        if (!payload.empty()) memcpy(buf + hdr.size(), payload.data(), payload.size());
    }
    return R3D_OK;
}

extern "C" int r3d_tree_write_bt(r3d_tree* t, const char* path) {
    if (!t || !path) return set_error(t ? t->ctx : nullptr, R3D_ERR_ARG, "null argument");
    DeviceSetter ds(t->ctx->device);
    uint64_t n_nodes = 0;
    std::vector<uint8_t> payload;
    R3D_TRY(tree_shape(t, true, &n_nodes, &payload));
    const std::string hdr = bt_header(n_nodes, t->res);
    FILE* f = fopen(path, "wb");
    if (!f) return set_error(t->ctx, R3D_ERR_IO, "cannot open %s for writing", path);
    bool ok = fwrite(hdr.data(), 1, hdr.size(), f) == hdr.size();
    if (ok && !payload.empty()) ok = fwrite(payload.data(), 1, payload.size(), f) == payload.size();
    ok = (fclose(f) == 0) && ok;
    if (!ok) return set_error(t->ctx, R3D_ERR_IO, "short write to %s", path);
    return R3D_OK;
}

// ------------------------------------------------------------------ .bt reader (section 8f: the step after the path)
// AbstractOcTree::readBinary / OcTreeBase::readBinaryData of the `octomap` extension: header text, then the pre-order
// 2-bits-per-child stream; an occupied leaf gets the upper clamping value, a free leaf the lower one.  The stream is
// walked on the host (it is a few hundred KB) into whole-brick records -- a leaf at depth d <= 13 covers 8^(13-d)
// bricks -- which the GPU store imports (k_import_bricks).
namespace r3d {

struct BtReader {
    const uint8_t* p;
    const uint8_t* end;
    float occ, fre;
    std::vector<BrickRecord> bricks;           // depth-13 nodes in stream order; the last one is under construction
    bool overflow = false;
    uint64_t nodes = 1;                        // the root

    // fill voxels [first, first+count) of the brick being built
    static void fill(BrickRecord& b, uint32_t first, uint32_t count, float v) {
        for (uint32_t i = first; i < first + count; ++i) { b.value[i] = v; b.known[i >> 5] |= 1u << (i & 31u); }
    }
    // a leaf at depth <= 13: every brick below it is completely v.  prefix = Morton code of the node (3 bits per level)
    void leaf_above_bricks(uint64_t prefix, int depth, float v) {
        const uint64_t n = 1ull << (3 * (13 - depth));
        if (bricks.size() + n > (1ull << 24)) { overflow = true; return; }
        for (uint64_t k = 0; k < n; ++k) {
            BrickRecord b;
            memset(&b, 0, sizeof b);
            b.key = morton_to_brick_key((prefix << (3 * (13 - depth))) | k);
            fill(b, 0, 512, v);
            bricks.push_back(b);
        }
    }
    static uint64_t morton_to_brick_key(uint64_t m) {
        uint64_t x = 0, y = 0, z = 0;
        for (int i = 0; i < 13; ++i) {
            x |= ((m >> (3 * i)) & 1ull) << i;
            y |= ((m >> (3 * i + 1)) & 1ull) << i;
            z |= ((m >> (3 * i + 2)) & 1ull) << i;
        }
        return x | (y << 13) | (z << 26);
    }
    // inner node at `depth` (0 = root) with Morton prefix; below depth 13 vox_first is the node's first voxel in its brick
    bool node(uint64_t prefix, int depth, uint32_t vox_first) {
        if (p + 2 > end) return false;
        const unsigned bits = (unsigned)p[0] | ((unsigned)p[1] << 8);
        p += 2;
        for (int c = 0; c < 8; ++c) {
            // child c: bit 2c set -> "free" flag, bit 2c+1 -> "occupied" flag; both -> inner (writeBinaryNode's encoding)
            const unsigned code = (bits >> (2 * c)) & 3u;
            if (code == 0u) continue;
            ++nodes;
            const int cd = depth + 1;
            const uint64_t cp = (prefix << 3) | (uint64_t)c;
            if (cd <= 13) {
                if (code == 3u) {
                    if (cd == 13) {
                        if (bricks.size() >= (1ull << 24)) { overflow = true; return false; }
                        BrickRecord b;
                        memset(&b, 0, sizeof b);
                        b.key = morton_to_brick_key(cp);
                        bricks.push_back(b);
                        if (!node(cp, cd, 0)) return false;
                    } else if (!node(cp, cd, 0)) return false;
                } else {
                    leaf_above_bricks(cp, cd, code == 2u ? occ : fre);
                    if (overflow) return false;
                }
            } else {
                // inside the last brick: voxel range of the child, Morton order = child-index order
                const uint32_t span = 1u << (3 * (16 - cd));
                const uint32_t first = vox_first + (uint32_t)c * span;
                if (code == 3u) { if (!node(cp, cd, first)) return false; }
                else fill(bricks.back(), first, span, code == 2u ? occ : fre);
            }
        }
        return true;
    }
};

}  // namespace r3d

extern "C" int r3d_tree_read_bt_mem(r3d_tree* t, const uint8_t* data, size_t len) {
    if (!t || (!data && len)) return set_error(t ? t->ctx : nullptr, R3D_ERR_ARG, "null argument");
    r3d_ctx* ctx = t->ctx;
    // header: lines until "data\n"
    const char* s = reinterpret_cast<const char*>(data);
    size_t pos = 0;
    auto next_line = [&](std::string& line) {
        if (pos >= len) return false;
        size_t e = pos;
        while (e < len && s[e] != '\n') ++e;
        line.assign(s + pos, e - pos);
        pos = e < len ? e + 1 : e;
        return true;
    };
    std::string line;
    if (!next_line(line) || line.rfind("# Octomap OcTree binary file", 0) != 0)
        return set_error(ctx, R3D_ERR_UNSUPPORTED, "not an OctoMap binary (.bt) file: first line is not '# Octomap OcTree binary file'");
    std::string id;
    double res = 0.0;
    unsigned long long size = 0;
    bool have_data = false;
    while (next_line(line)) {
        if (line.empty() || line[0] == '#') continue;
        if (line == "data") { have_data = true; break; }
        char key[32] = {0};
        char val[128] = {0};
        if (sscanf(line.c_str(), "%31s %127s", key, val) == 2) {
            if (!strcmp(key, "id")) id = val;
            else if (!strcmp(key, "res")) res = atof(val);
            else if (!strcmp(key, "size")) size = strtoull(val, nullptr, 10);
        }
    }
    if (!have_data) return set_error(ctx, R3D_ERR_UNSUPPORTED, ".bt header has no 'data' line");
    if (id != "OcTree") return set_error(ctx, R3D_ERR_UNSUPPORTED, ".bt holds a tree of type '%s', only OcTree is supported", id.c_str());
    if (!(res > 0.0)) return set_error(ctx, R3D_ERR_UNSUPPORTED, ".bt header has no valid resolution");
    DeviceSetter ds(ctx->device);
    R3D_TRY(r3d_tree_clear(t));
    t->res = res;
    t->res_factor = 1.0 / res;
    if (size == 0) return R3D_OK;
    BtReader rd;
    rd.p = data + pos;
    rd.end = data + len;
    rd.occ = t->cmax;
    rd.fre = t->cmin;
    if (!rd.node(0, 0, 0)) {
        if (rd.overflow) return set_error(ctx, R3D_ERR_OOM, ".bt expands to more than 2^24 bricks");
        return set_error(ctx, R3D_ERR_IO, ".bt payload is truncated");
    }
    if (rd.nodes != size) return set_error(ctx, R3D_ERR_IO, ".bt header says %llu nodes, the stream holds %llu", size, (unsigned long long)rd.nodes);
    if (!rd.bricks.empty()) R3D_TRY(r3d_tree_import_bricks(t, rd.bricks.data(), rd.bricks.size()));
    return R3D_OK;
}

extern "C" int r3d_tree_read_bt(r3d_tree* t, const char* path) {
    if (!t || !path) return set_error(t ? t->ctx : nullptr, R3D_ERR_ARG, "null argument");
    FILE* f = fopen(path, "rb");
    if (!f) return set_error(t->ctx, R3D_ERR_IO, "cannot open %s", path);
    std::vector<uint8_t> buf;
    uint8_t chunk[1 << 16];
    size_t n;
    while ((n = fread(chunk, 1, sizeof chunk, f)) > 0) buf.insert(buf.end(), chunk, chunk + n);
    fclose(f);
    return r3d_tree_read_bt_mem(t, buf.data(), buf.size());
}

// ------------------------------------------------------------------ .ot writer / reader (section 8f: full tree, log-odds kept)
// AbstractOcTree::write / OcTreeBaseImpl::writeData of the `octomap` extension: header ("# Octomap OcTree file"), then
// pre-order per node: float32 value, one byte of child-exists bits, the children.  The tree written is upstream's tree
// as it stands after non-lazy updates: maximally pruned under exact equality of sibling leaf values, inner values =
// max of the children (the same shape r3d_tree_size counts).  Rarely used next to .bt, so it is assembled on the host
// from the exported bricks (sorted by Morton code; voxels inside a brick already are in child order).
namespace r3d {

struct OtWriter {
    std::vector<uint8_t> bytes;
    const BrickRecord* rec = nullptr;
    const std::vector<std::pair<uint64_t, uint32_t>>* order = nullptr;   // (brick Morton code, record index), sorted

    struct Res { float value; bool leaf; };

    size_t open_node() { const size_t pos = bytes.size(); bytes.resize(pos + 5); return pos; }
    void close_node(size_t pos, float v, uint8_t mask) { memcpy(&bytes[pos], &v, 4); bytes[pos + 4] = mask; }

    // children results -> this node (pruning when all 8 are equal-valued leaves)
    Res finish(size_t pos, const Res* child, uint8_t mask) {
        float mx = 0.f;
        bool first = true, all_leaf = true, all_eq = true;
        float v0 = 0.f;
        for (int c = 0; c < 8; ++c) {
            if (!((mask >> c) & 1)) continue;
            if (first) { mx = child[c].value; v0 = child[c].value; first = false; }
            else { if (child[c].value > mx) mx = child[c].value; if (!(child[c].value == v0)) all_eq = false; }
            if (!child[c].leaf) all_leaf = false;
        }
        if (mask == 0xff && all_leaf && all_eq) {
            bytes.resize(pos + 5);
            close_node(pos, v0, 0);
            return {v0, true};
        }
        close_node(pos, mx, mask);
        return {mx, false};
    }

    // node inside a brick: level 13 (span 512) .. 16 (span 1); exists iff any voxel of its range is known
    Res brick_node(const BrickRecord& b, uint32_t first, uint32_t span) {
        const size_t pos = open_node();
        if (span == 1) { close_node(pos, b.value[first], 0); return {b.value[first], true}; }
        Res child[8];
        uint8_t mask = 0;
        const uint32_t cs = span / 8;
        for (int c = 0; c < 8; ++c) {
            const uint32_t f = first + (uint32_t)c * cs;
            bool any = false;
            for (uint32_t i = f; i < f + cs && !any; ++i) any = (b.known[i >> 5] >> (i & 31u)) & 1u;
            if (!any) continue;
            child[c] = brick_node(b, f, cs);
            mask |= (uint8_t)(1u << c);
        }
        return finish(pos, child, mask);
    }

    // node above brick level: bricks [lo, hi) of the sorted order share the Morton prefix of this node
    Res upper_node(int level, size_t lo, size_t hi) {
        if (level == 13) return brick_node(rec[(*order)[lo].second], 0, 512);
        const size_t pos = open_node();
        Res child[8];
        uint8_t mask = 0;
        const int sh = 3 * (12 - level);
        size_t i = lo;
        while (i < hi) {
            const unsigned c = (unsigned)(((*order)[i].first >> sh) & 7u);
            size_t j = i + 1;
            while (j < hi && (unsigned)(((*order)[j].first >> sh) & 7u) == c) ++j;
            child[c] = upper_node(level + 1, i, j);
            mask |= (uint8_t)(1u << c);
            i = j;
        }
        return finish(pos, child, mask);
    }
};

static std::string ot_header(unsigned long long n_nodes, double res) {
    std::string h = bt_header(n_nodes, res);
    const std::string from = "# Octomap OcTree binary file";
    h.replace(0, from.size(), "# Octomap OcTree file");
    return h;
}

static int tree_serialise_ot(r3d_tree* t, std::vector<uint8_t>& out) {
    r3d_ctx* ctx = t->ctx;
    uint64_t nb = 0;
    R3D_TRY(r3d_tree_num_bricks(t, &nb));
    std::vector<BrickRecord> recs(nb);
    if (nb) R3D_TRY(r3d_tree_export_bricks(t, recs.data(), nb, &nb));
    std::vector<std::pair<uint64_t, uint32_t>> order;
    order.reserve(nb);
    for (uint32_t i = 0; i < nb; ++i) {
        bool any = false;
        for (int w = 0; w < 16 && !any; ++w) any = recs[i].known[w] != 0;
        if (any) order.emplace_back(brick_morton(recs[i].key), i);
    }
    std::sort(order.begin(), order.end());
    OtWriter w;
    w.rec = recs.data();
    w.order = &order;
    if (!order.empty()) w.upper_node(0, 0, order.size());
    const std::string hdr = ot_header(w.bytes.size() / 5, t->res);
    out.assign(hdr.begin(), hdr.end());
    out.insert(out.end(), w.bytes.begin(), w.bytes.end());
    (void)ctx;
    return R3D_OK;
}

// reader: leaves at any depth carry their own value
struct OtReader {
    const uint8_t* p;
    const uint8_t* end;
    std::vector<BrickRecord> bricks;
    bool overflow = false;
    uint64_t nodes = 0;

    void leaf_above_bricks(uint64_t prefix, int depth, float v) {
        const uint64_t n = 1ull << (3 * (13 - depth));
        if (bricks.size() + n > (1ull << 24)) { overflow = true; return; }
        for (uint64_t k = 0; k < n; ++k) {
            BrickRecord b;
            memset(&b, 0, sizeof b);
            b.key = BtReader::morton_to_brick_key((prefix << (3 * (13 - depth))) | k);
            BtReader::fill(b, 0, 512, v);
            bricks.push_back(b);
        }
    }
    bool node(uint64_t prefix, int depth, uint32_t vox_first) {
        if (p + 5 > end) return false;
        float v;
        memcpy(&v, p, 4);
        const unsigned mask = p[4];
        p += 5;
        ++nodes;
        if (mask == 0) {   // leaf
            if (depth <= 13) { leaf_above_bricks(prefix, depth, v); return !overflow; }
            BtReader::fill(bricks.back(), vox_first, 1u << (3 * (16 - depth)), v);
            return true;
        }
        if (depth >= 16) return false;   // a depth-16 node cannot have children
        if (depth == 13) {
            if (bricks.size() >= (1ull << 24)) { overflow = true; return false; }
            BrickRecord b;
            memset(&b, 0, sizeof b);
            b.key = BtReader::morton_to_brick_key(prefix);
            bricks.push_back(b);
        }
        for (int c = 0; c < 8; ++c) {
            if (!((mask >> c) & 1u)) continue;
            const int cd = depth + 1;
            uint32_t first = 0;
            if (cd > 13) first = vox_first + (uint32_t)c * (1u << (3 * (16 - cd)));
            if (!node((prefix << 3) | (uint64_t)c, cd, first)) return false;
        }
        return true;
    }
};

}  // namespace r3d

extern "C" int r3d_tree_write_ot_mem(r3d_tree* t, uint8_t* buf, size_t cap, size_t* len) {
    if (!t || !len) return set_error(t ? t->ctx : nullptr, R3D_ERR_ARG, "null argument");
    DeviceSetter ds(t->ctx->device);
    std::vector<uint8_t> out;
    R3D_TRY(tree_serialise_ot(t, out));
    *len = out.size();
    if (buf && cap >= out.size()) memcpy(buf, out.data(), out.size());
    return R3D_OK;
}

extern "C" int r3d_tree_write_ot(r3d_tree* t, const char* path) {
    if (!t || !path) return set_error(t ? t->ctx : nullptr, R3D_ERR_ARG, "null argument");
    DeviceSetter ds(t->ctx->device);
    std::vector<uint8_t> out;
    R3D_TRY(tree_serialise_ot(t, out));
    FILE* f = fopen(path, "wb");
    if (!f) return set_error(t->ctx, R3D_ERR_IO, "cannot open %s for writing", path);
    const bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
    if (fclose(f) != 0 || !ok) return set_error(t->ctx, R3D_ERR_IO, "short write to %s", path);
    return R3D_OK;
}

extern "C" int r3d_tree_read_ot_mem(r3d_tree* t, const uint8_t* data, size_t len) {
    if (!t || (!data && len)) return set_error(t ? t->ctx : nullptr, R3D_ERR_ARG, "null argument");
    r3d_ctx* ctx = t->ctx;
    const char* s = reinterpret_cast<const char*>(data);
    size_t pos = 0;
    auto next_line = [&](std::string& line) {
        if (pos >= len) return false;
        size_t e = pos;
        while (e < len && s[e] != '\n') ++e;
        line.assign(s + pos, e - pos);
        pos = e < len ? e + 1 : e;
        return true;
    };
    std::string line;
    if (!next_line(line) || line.rfind("# Octomap OcTree file", 0) != 0)
        return set_error(ctx, R3D_ERR_UNSUPPORTED, "not an OctoMap .ot file: first line is not '# Octomap OcTree file'");
    std::string id;
    double res = 0.0;
    unsigned long long size = 0;
    bool have_data = false;
    while (next_line(line)) {
        if (line.empty() || line[0] == '#') continue;
        if (line == "data") { have_data = true; break; }
        char key[32] = {0};
        char val[128] = {0};
        if (sscanf(line.c_str(), "%31s %127s", key, val) == 2) {
            if (!strcmp(key, "id")) id = val;
            else if (!strcmp(key, "res")) res = atof(val);
            else if (!strcmp(key, "size")) size = strtoull(val, nullptr, 10);
        }
    }
    if (!have_data) return set_error(ctx, R3D_ERR_UNSUPPORTED, ".ot header has no 'data' line");
    if (id != "OcTree") return set_error(ctx, R3D_ERR_UNSUPPORTED, ".ot holds a tree of type '%s', only OcTree is supported", id.c_str());
    if (!(res > 0.0)) return set_error(ctx, R3D_ERR_UNSUPPORTED, ".ot header has no valid resolution");
    DeviceSetter ds(ctx->device);
    R3D_TRY(r3d_tree_clear(t));
    t->res = res;
    t->res_factor = 1.0 / res;
    if (size == 0) return R3D_OK;
    OtReader rd;
    rd.p = data + pos;
    rd.end = data + len;
    if (!rd.node(0, 0, 0)) {
        if (rd.overflow) return set_error(ctx, R3D_ERR_OOM, ".ot expands to more than 2^24 bricks");
        return set_error(ctx, R3D_ERR_IO, ".ot payload is truncated or malformed");
    }
    if (rd.nodes != size) return set_error(ctx, R3D_ERR_IO, ".ot header says %llu nodes, the stream holds %llu", size, (unsigned long long)rd.nodes);
    if (!rd.bricks.empty()) R3D_TRY(r3d_tree_import_bricks(t, rd.bricks.data(), rd.bricks.size()));
    return R3D_OK;
}

extern "C" int r3d_tree_read_ot(r3d_tree* t, const char* path) {
    if (!t || !path) return set_error(t ? t->ctx : nullptr, R3D_ERR_ARG, "null argument");
    FILE* f = fopen(path, "rb");
    if (!f) return set_error(t->ctx, R3D_ERR_IO, "cannot open %s", path);
    std::vector<uint8_t> buf;
    uint8_t chunk[1 << 16];
    size_t n;
    while ((n = fread(chunk, 1, sizeof chunk, f)) > 0) buf.insert(buf.end(), chunk, chunk + n);
    fclose(f);
    return r3d_tree_read_ot_mem(t, buf.data(), buf.size());
}
