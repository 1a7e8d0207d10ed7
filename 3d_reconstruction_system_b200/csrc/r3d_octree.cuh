// r3d_octree.cuh -- the voxel store shared by the update kernels (r3d_octree.cu) and the .bt serialiser (r3d_bt.cu).
#pragma once
#include <vector>

#include "r3d_common.cuh"
#include "r3d_vmm.cuh"

namespace r3d {

constexpr uint64_t kEmptyKey = ~0ull;
constexpr uint64_t kNoSlot = ~0ull;
constexpr uint32_t kBrickVoxels = 512;

// one brick of a scan delta: see R3D_DELTA_RECORD_BYTES in r3d.h
struct DeltaRecord {
    uint64_t key;
    uint32_t mask[32];   // [0,16): occupied, [16,32): free (already minus occupied)
};
static_assert(sizeof(DeltaRecord) == R3D_DELTA_RECORD_BYTES, "record layout is part of the ABI");

// one whole brick of the map: see R3D_BRICK_RECORD_BYTES in r3d.h
struct BrickRecord {
    uint64_t key;
    float value[512];
    uint32_t known[16];
};
static_assert(sizeof(BrickRecord) == R3D_BRICK_RECORD_BYTES, "record layout is part of the ABI");

// device counters of a tree
enum { CNT_POOL_USED = 0, CNT_OVERFLOW, CNT_DROPPED, CNT_SCRATCH_USED, CNT_DELTA, CNT_DISCRETE, CNT_STEPS_LO, CNT_STEPS_HI, CNT_RAY_LO, CNT_RAY_HI, CNT_GRID_MISS,
       CNT_GRID_NEED = 11 /* cells the scan's cube would need (0xffffffff: cannot be direct-mapped at all) */,
       CNT_NRAYS = 12 /* rays with at least one free cell, appended by k_scan_prepare */,
       CNT_ORIGIN_OCC = 13 /* an in-range endpoint lies in the sensor's own voxel */,
       CNT_APPLY_OVERFLOW = 15 /* sticky */, CNT_COUNT = 16 };

}  // namespace r3d

struct r3d_tree {
    r3d_ctx* ctx = nullptr;
    double res = 0.1, res_factor = 10.0;
    float hit = 0, miss = 0, cmin = 0, cmax = 0, occ_thres = 0;
    // persistent store: hash (brick key -> pool index) + brick pool
    uint64_t* tkeys = nullptr;
    uint32_t* tvals = nullptr;
    uint64_t tcap = 0;
    std::vector<void*> retired;   // tables replaced by a re-hash: freed with the tree (no device-wide sync while mapping)
    float* values = nullptr;      // [pool_cap][512] log-odds, Morton order inside the brick
    uint32_t* known = nullptr;    // [pool_cap][16]  voxel was updated at least once (node exists)
    uint64_t pool_cap = 0;
    // the pool grows in place (r3d_vmm.cuh) where the driver offers virtual memory management; values / known then point into these
    r3d::VmmRegion vm_values, vm_known;
    bool pool_vmm = false;
    uint32_t pool_used = 0;       // host mirror of counters[CNT_POOL_USED] as of the last counter read-back
    uint64_t pool_bound = 0;      // upper bound of the device value once every queued apply has run
    bool pool_dirty = false;      // applies were queued since the last read-back: call tree_settle before using pool_used
    int raycast_blocks_per_sm = 0, raycast_blocks_per_sm_hash = 0;
    uint64_t* pool_keys = nullptr; // [pool_cap] brick key per pool entry, rebuilt from the table on demand
    uint64_t pool_keys_cap = 0;
    // per-scan scratch table + compacted delta
    uint64_t* skeys = nullptr;
    uint32_t* smasks = nullptr;
    uint64_t scap = 0;
    r3d::DeltaRecord* delta = nullptr;
    uint64_t delta_cap = 0, delta_n = 0;
    uint64_t pipe_wait_ns = 0, pipe_work_ns = 0, pipe_max_turn_ns = 0, pipe_scans = 0;   // host clock of the last pipelined batch
    // applies handed in with r3d_tree_defer_deltas_owned: queued by the next scan batch once its first ray casts are in flight
    // (so that they run BESIDE the ray casting instead of ahead of it), or by whatever touches the map first
    struct Deferred { const r3d::DeltaRecord* recs; std::vector<uint64_t> counts; uint32_t part, nparts; };
    std::vector<Deferred> deferred;
    uint64_t n_pool_grow = 0, n_table_grow = 0;   // regrowth events since the tree was created (each copies / re-hashes)
    uint64_t last_scan_rays = 0, last_scan_steps = 0;
    cudaEvent_t cast_gate = nullptr;   // r3d_scan_deltas_compute: the context stream before the deferred applies were queued
    bool cast_gate_armed = false;
    uint64_t last_batch_records = 0;   // records of the last scan a batch insert applied (statistics; its delta is not kept)
    uint32_t* counters = nullptr;
    uint32_t h_counters[r3d::CNT_COUNT] = {0};
};

namespace r3d {
// r3d_raycast.cu: bounded-range insertPointCloud batches (K3 dense pipeline).  Where the records of each scan go:
struct ScanSink {
    enum Mode { APPLY, EXPORT_USER, EXPORT_TREE } mode = APPLY;
    void* records = nullptr;        // EXPORT_USER: caller's buffer (host or device), records back to back
    uint64_t capacity = 0, used = 0;
    uint64_t* counts = nullptr;     // EXPORT_USER: per-scan record counts
};
// Runs scans [0, n_scans) of a device-resident batch through the dense pipeline.  *done < n_scans: scan *done cannot be
// direct-mapped (the caller takes it through the hash path and calls again for the rest).
int dense_scans_run(r3d_tree* t, const float* d_xyz, const uint64_t* n_points, const float* origins, uint32_t n_scans, double maxrange,
                    ScanSink* sink, uint32_t* done, uint64_t* rays, uint64_t* steps);
void scan_pipe_destroy(r3d_ctx* ctx);
int apply_delta_impl(r3d_tree* t, const DeltaRecord* d_recs, uint64_t n, uint32_t part = 0, uint32_t nparts = 1);
int tree_reserve_delta(r3d_tree* t, uint64_t want);
unsigned grid_for(r3d_ctx* ctx, unsigned long long items, int block = 256, int per_sm = 8);
int tree_sync_counters(r3d_tree* t);
int tree_settle(r3d_tree* t);   // deferred applies queued, pool_used exact again (reads the counters back if applies are pending)
int tree_flush_deferred(r3d_tree* t);
int tree_reserve(r3d_tree* t, uint64_t n_bricks);   // table + pool with room for n_bricks
// r3d_round.cu: many scans' deltas in one sorted, scan-ordered pass
int apply_round_sorted(r3d_tree* t, const std::vector<r3d_tree::Deferred>& jobs);
int tree_refresh_pool_keys(r3d_tree* t);
}  // namespace r3d
