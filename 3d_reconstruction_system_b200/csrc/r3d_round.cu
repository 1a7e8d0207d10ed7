// r3d_round.cu -- K4 for a whole ROUND of the multi-GPU merge: the deltas of many scans (every rank's share of a round,
// in global scan order) applied with three launches instead of one per scan and rank.
//
// north_star: "deltas are merged ... followed by a sorted, scan-ordered apply, so the final tree is independent of GPU
// count".  Clamped float32 adds are not associative, so a voxel must see its updates in scan order; bricks are independent of
// each other.  (1) k_round_index keeps the records of the bricks this rank owns and emits (brick key << 8 | scan order,
// record address); (2) a radix sort groups them by brick, scan order inside a brick; (3) k_round_apply gives every brick
// to one warp, which finds / creates the brick ONCE and applies its records in scan order.  At 8 GPUs a rank applies 8
// scans for every scan it ray-casts; one launch per scan made that 8 x ~15 us of launch-bound kernels per 0.7 ms ray cast.
#include <cub/device/device_radix_sort.cuh>

#include "r3d_octree.cuh"

namespace r3d {

struct RoundScan {
    const DeltaRecord* recs;
    uint32_t n;
    uint32_t order;      // position of the scan in the round (< 256)
};

__global__ void __launch_bounds__(256) k_round_index(const RoundScan* __restrict__ scans, uint32_t part, uint32_t nparts, uint64_t* keys, uint64_t* vals,
                                                     uint32_t* counter, uint32_t cap) {
    const RoundScan sc = scans[blockIdx.y];
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t n_pad = (sc.n + 31u) & ~31u;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_pad; r += gridDim.x * blockDim.x) {
        uint64_t key = 0;
        bool mine = false;
        if (r < sc.n) {
            key = sc.recs[r].key;
            mine = nparts <= 1 || brick_owner(key, nparts) == part;
        }
        const unsigned m = __ballot_sync(0xffffffffu, mine);
        if (!m) continue;
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(counter, (uint32_t)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (mine) {
            const uint32_t pos = base + __popc(m & ((1u << lane) - 1u));
            if (pos < cap) {
                keys[pos] = (key << 8) | (uint64_t)sc.order;
                vals[pos] = (uint64_t)(uintptr_t)(sc.recs + r);
            }
        }
    }
}

// number of distinct bricks among the sorted entries (= the most bricks the apply can create)
__global__ void k_round_count_runs(const uint64_t* __restrict__ keys, uint32_t n, uint32_t* out) {
    uint32_t c = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        c += (i == 0 || (keys[i] >> 8) != (keys[i - 1] >> 8)) ? 1u : 0u;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31u) == 0 && c) atomicAdd(out, c);
}

// one warp per run of equal brick keys: the brick is found / created once, its records are applied in scan order
__global__ void __launch_bounds__(256) k_round_apply(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ vals, uint32_t n, uint64_t* tkeys,
                                                     uint32_t* tvals, uint64_t tcap, float* values, uint32_t* known, uint32_t* counters, float hit, float miss,
                                                     float cmin, float cmax) {
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
        const uint64_t bk = keys[i] >> 8;
        if (i != 0 && (keys[i - 1] >> 8) == bk) continue;       // not the first record of its brick
        uint32_t idx = 0;
        if (lane == 0) {
            // (same find-or-insert as k_apply_delta; r3d_octree.cu)
            const uint64_t mask = tcap - 1;
            uint64_t slot = hash64(bk) & mask;
            idx = 0xffffffffu;
            for (uint64_t probe = 0; probe < tcap; ++probe) {
                const uint64_t k = __ldcg(reinterpret_cast<const unsigned long long*>(tkeys + slot));
                if (k == bk) { idx = tvals[slot]; break; }
                if (k == kEmptyKey) {
                    const unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(tkeys + slot), kEmptyKey, bk);
                    if (old == kEmptyKey) { idx = atomicAdd(&counters[CNT_POOL_USED], 1u); tvals[slot] = idx; break; }
                    if (old == bk) { idx = tvals[slot]; break; }
                }
                slot = (slot + 1) & mask;
            }
            if (idx == 0xffffffffu) counters[CNT_APPLY_OVERFLOW] = 1;
        }
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx == 0xffffffffu) continue;
        // lane handles voxels [16*lane, 16*lane+16): half of mask word lane/2.  The 16 log-odds stay in registers over the run.
        float4* v4 = reinterpret_cast<float4*>(values + (size_t)idx * kBrickVoxels + lane * 16);
        float4 v[4] = {v4[0], v4[1], v4[2], v4[3]};
        float* e = reinterpret_cast<float*>(v);
        uint32_t touched = 0;
        const uint32_t sh = (lane & 1u) * 16u;
        for (uint32_t j = i; j < n && (keys[j] >> 8) == bk; ++j) {
            const DeltaRecord* rec = reinterpret_cast<const DeltaRecord*>((uintptr_t)vals[j]);
            const uint32_t occ = (rec->mask[lane >> 1] >> sh) & 0xffffu;
            const uint32_t fre = (rec->mask[16 + (lane >> 1)] >> sh) & 0xffffu;
            touched |= occ | fre;
#pragma unroll
            for (int b = 0; b < 16; ++b) {
                const uint32_t bit = 1u << b;
                if (occ & bit) e[b] = clamped_add(e[b], hit, cmin, cmax);
                else if (fre & bit) e[b] = clamped_add(e[b], miss, cmin, cmax);
            }
        }
        if (touched) { v4[0] = v[0]; v4[1] = v[1]; v4[2] = v[2]; v4[3] = v[3]; }
        const uint32_t other = __shfl_xor_sync(0xffffffffu, touched, 1);
        if (!(lane & 1u)) {
            const uint32_t word = touched | (other << 16);
            if (word) known[(size_t)idx * 16 + (lane >> 1)] |= word;
        }
    }
}

// `jobs`: deltas in apply order (each a run of scans back to back in device memory).  At most 256 scans per sorted pass.
int apply_round_sorted(r3d_tree* t, const std::vector<r3d_tree::Deferred>& jobs) {
    r3d_ctx* ctx = t->ctx;
    std::vector<RoundScan> scans;
    uint32_t part = 0, nparts = 1;
    size_t first_job = 0;
    auto run_pass = [&](const std::vector<RoundScan>& sc) -> int {
        uint64_t total = 0;
        for (const auto& s : sc) total += s.n;
        if (total == 0) return R3D_OK;
        if (total > 0xfffffff0ull) return set_error(ctx, R3D_ERR_ARG, "too many records in one round");
        // scratch: scan table | counters | keys x2 | values x2
        const size_t tab_bytes = (sc.size() * sizeof(RoundScan) + 255) / 256 * 256;
        const size_t arr = ((size_t)total * 8 + 255) / 256 * 256;
        R3D_TRY(scratch_reserve(ctx, SCR_OUT1, tab_bytes + 256 + 4 * arr));
        char* base = (char*)ctx->scratch[SCR_OUT1];
        RoundScan* d_tab = (RoundScan*)base;
        uint32_t* d_cnt = (uint32_t*)(base + tab_bytes);
        uint64_t* k0 = (uint64_t*)(base + tab_bytes + 256);
        uint64_t* k1 = k0 + arr / 8;
        uint64_t* v0 = k1 + arr / 8;
        uint64_t* v1 = v0 + arr / 8;
        R3D_CUDA_OK(ctx, cudaMemcpyAsync(d_tab, sc.data(), sc.size() * sizeof(RoundScan), cudaMemcpyHostToDevice, ctx->stream));
        R3D_CUDA_OK(ctx, cudaMemsetAsync(d_cnt, 0, 64, ctx->stream));
        uint32_t n_max = 0;
        for (const auto& s : sc) n_max = s.n > n_max ? s.n : n_max;
        unsigned gx = (n_max + 255u) / 256u;
        if (gx > 64u) gx = 64u;
        if (gx < 1u) gx = 1u;
        k_round_index<<<dim3(gx, (unsigned)sc.size()), 256, 0, ctx->stream>>>(d_tab, part, nparts, k0, v0, d_cnt, (uint32_t)total);
        ctx->launches++;
        uint32_t n_own = 0;
        R3D_CUDA_OK(ctx, cudaMemcpyAsync(ctx->pinned, d_cnt, 4, cudaMemcpyDeviceToHost, ctx->stream));
        R3D_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));     // (the scan table on the stack has been uploaded too)
        memcpy(&n_own, ctx->pinned, 4);
        if (n_own == 0) return R3D_OK;
        size_t tmp = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp, k0, k1, v0, v1, (int)n_own, 0, 48, ctx->stream);
        R3D_TRY(scratch_reserve(ctx, SCR_CUBTMP, tmp + 256));
        R3D_CUDA_OK(ctx, cub::DeviceRadixSort::SortPairs(ctx->scratch[SCR_CUBTMP], tmp, k0, k1, v0, v1, (int)n_own, 0, 48, ctx->stream));
        k_round_count_runs<<<grid_for(ctx, n_own), 256, 0, ctx->stream>>>(k1, n_own, d_cnt + 1);
        ctx->launches += 2;
        uint32_t n_runs = 0;
        R3D_CUDA_OK(ctx, cudaMemcpyAsync(ctx->pinned, d_cnt + 1, 4, cudaMemcpyDeviceToHost, ctx->stream));
        R3D_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        memcpy(&n_runs, ctx->pinned, 4);
        // room for every brick the pass can create (same bookkeeping as apply_delta_impl)
        if (t->pool_dirty && t->pool_bound + n_runs > t->pool_cap) {
            R3D_TRY(tree_sync_counters(t));
            R3D_TRY(tree_reserve(t, t->pool_bound + 4ull * n_runs));
        }
        R3D_TRY(tree_reserve(t, t->pool_bound + n_runs));
        k_round_apply<<<grid_for(ctx, (uint64_t)n_own * 32, 256, 8), 256, 0, ctx->stream>>>(k1, v1, n_own, t->tkeys, t->tvals, t->tcap, t->values, t->known, t->counters,
                                                                                          t->hit, t->miss, t->cmin, t->cmax);
        ctx->launches++;
        R3D_CUDA_OK(ctx, cudaGetLastError());
        t->pool_bound += n_runs;
        t->pool_dirty = true;
        return R3D_OK;
    };
    for (size_t j = 0; j < jobs.size(); ++j) {
        const auto& job = jobs[j];
        // a pass holds jobs of one partition and at most 256 scans
        if (!scans.empty() && (job.part != part || job.nparts != nparts || scans.size() + job.counts.size() > 256)) {
            R3D_TRY(run_pass(scans));
            scans.clear();
        }
        if (scans.empty()) { part = job.part; nparts = job.nparts; first_job = j; }
        (void)first_job;
        const DeltaRecord* d = job.recs;
        for (uint64_t c : job.counts) {
            if (scans.size() == 256) { R3D_TRY(run_pass(scans)); scans.clear(); }
            RoundScan s;
            s.recs = d; s.n = (uint32_t)c; s.order = (uint32_t)scans.size();
            scans.push_back(s);
            d += c;
        }
    }
    if (!scans.empty()) R3D_TRY(run_pass(scans));
    return R3D_OK;
}

}  // namespace r3d
