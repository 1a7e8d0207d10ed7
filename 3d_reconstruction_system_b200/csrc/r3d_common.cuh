// r3d_common.cuh -- context object, error plumbing and small device helpers shared by the translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/r3d.h"
#include "r3d_math.cuh"

namespace r3d { struct ScanPipe; }

struct r3d_ctx {
    int device = 0;
    int sm_count = 148;
    int blocking = 1;
    cudaStream_t stream = nullptr;   // every kernel of this context
    cudaStream_t copy_stream[2] = {nullptr, nullptr};  // host-pointer staging pipeline
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    r3d::ScanPipe* scan_pipe = nullptr;   // r3d_raycast.cu: streams, cubes and buffers of the scan pipeline (created on first use)
    // staging ring of r3d_backproject_rt with host buffers: per slot "input uploaded", "kernel done", "output read back"
    static constexpr int kMaxStageSlots = 4;
    cudaEvent_t ev_in[kMaxStageSlots] = {}, ev_k[kMaxStageSlots] = {}, ev_out[kMaxStageSlots] = {};
    int stage_slots = 3;             // R3D_STAGE_SLOTS
    size_t stage_chunk_bytes = 192u << 20;   // R3D_STAGE_CHUNK_MB: host bytes (in + out) per chunk
    float last_kernel_ms = 0.f;
    uint64_t launches = 0;
    char err[1024] = {0};
    // reusable device scratch (grown on demand, freed in r3d_destroy)
    void* scratch[9] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // SCR_COUNT slots
    size_t scratch_bytes[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    uint64_t cell_budget_bytes = 16ull << 30;   // cap of the ray caster's direct-mapped scratch (R3D_SCAN_SCRATCH_GB overrides)
    // K1, disparity mode with integer samples: Z = fB / (raw * depth_scale) for every possible sample value (r3d_backproject.cu)
    double* ztab = nullptr;
    double ztab_scale = 0.0, ztab_fB = 0.0;
    int ztab_n = 0;
    void* pinned = nullptr;          // small pinned mailbox for counters read back from the device
    size_t pinned_bytes = 0;
};

namespace r3d {

// scratch slots
enum { SCR_POSE = 0, SCR_IN0 = 1, SCR_IN1 = 2, SCR_OUT0 = 3, SCR_OUT1 = 4, SCR_TILE = 5, SCR_CUBTMP = 6, SCR_MISC = 7, SCR_BT = 8, SCR_COUNT = 9 };
static_assert(SCR_COUNT == 9, "r3d_ctx::scratch holds SCR_COUNT slots");

extern char g_last_error[1024];

inline int set_error(r3d_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    strncpy(g_last_error, buf, sizeof g_last_error - 1);
    if (ctx) strncpy(ctx->err, buf, sizeof ctx->err - 1);
    return code;
}

#define R3D_CUDA_OK(ctx, expr)                                                                          \
    do {                                                                                                \
        cudaError_t e__ = (expr);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return r3d::set_error(ctx, e__ == cudaErrorMemoryAllocation ? R3D_ERR_OOM : R3D_ERR_CUDA,   \
                                  "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

#define R3D_TRY(expr)                   \
    do {                                \
        int rc__ = (expr);              \
        if (rc__ != R3D_OK) return rc__; \
    } while (0)

// Grows ctx->scratch[slot] to at least `bytes` (contents are NOT preserved).
int scratch_reserve(r3d_ctx* ctx, int slot, size_t bytes);
// true when p is device (or managed) memory; false for any host pointer.  *pinned = page-locked host.
bool is_device_ptr(const void* p, bool* pinned = nullptr);
int device_set(r3d_ctx* ctx);
int finish(r3d_ctx* ctx);   // sync if blocking, surface async errors

// r3d_inflate.cu: one zlib stream of exactly known output size; 0 on success, negative on a malformed stream
int inflate_zlib(const uint8_t* src, size_t src_len, uint8_t* dst, size_t dst_len);

struct DeviceSetter {
    int prev = -1;
    explicit DeviceSetter(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceSetter() { if (prev >= 0) cudaSetDevice(prev); }
};

}  // namespace r3d
