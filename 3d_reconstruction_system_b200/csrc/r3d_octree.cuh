// r3d_octree.cuh -- the voxel store shared by the update kernels (r3d_octree.cu) and the .bt serialiser (r3d_bt.cu).
#pragma once
#include "r3d_common.cuh"

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
       CNT_ABORT = 14 /* sticky: a pipelined scan could not be read back; every queued scan kernel skips until the host clears it */,
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
    float* values = nullptr;      // [pool_cap][512] log-odds, Morton order inside the brick
    uint32_t* known = nullptr;    // [pool_cap][16]  voxel was updated at least once (node exists)
    uint64_t pool_cap = 0;
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
    // second slot of the two-deep scan pipeline (r3d_tree_insert_scans): record buffer, per-slot counters, cell lists, events
    r3d::DeltaRecord* delta_b = nullptr;
    uint64_t delta_b_cap = 0;
    uint32_t* pipe_counters = nullptr;   // [2][CNT_COUNT]
    uint32_t* pipe_list = nullptr;       // [2][pipe_list_cap]
    uint64_t pipe_list_cap = 0;
    cudaEvent_t pipe_done[2] = {nullptr, nullptr};
    // overlapped ray casting (two cell cubes): scan s+1's ray cast runs on its own stream beside scan s's tail, list,
    // emit and apply.  R3D_PIPE_OVERLAP=0 keeps everything on the context stream with one cube.
    cudaEvent_t rc_done[2] = {nullptr, nullptr};
    cudaEvent_t pipe_start = nullptr;
    bool pipe_done_valid[2] = {false, false};
    int pipe_overlap = 1;
    uint64_t pipe_wait_ns = 0, pipe_work_ns = 0, pipe_max_turn_ns = 0, pipe_scans = 0;   // host clock of the last batch
    uint64_t last_scan_rays = 0, last_scan_steps = 0;
    uint32_t* counters = nullptr;
    uint32_t h_counters[r3d::CNT_COUNT] = {0};
};

namespace r3d {
int tree_sync_counters(r3d_tree* t);
int tree_settle(r3d_tree* t);   // pool_used exact again (reads the counters back if applies are pending)
int tree_refresh_pool_keys(r3d_tree* t);
}  // namespace r3d
