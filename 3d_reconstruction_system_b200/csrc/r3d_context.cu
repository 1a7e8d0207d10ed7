// r3d_context.cu -- context lifetime, error text, scratch memory, pointer classification.
#include <stdlib.h>

#include "r3d_octree.cuh"

namespace r3d {

char g_last_error[1024] = {0};

int scratch_reserve(r3d_ctx* ctx, int slot, size_t bytes) {
    if (slot < 0 || slot >= SCR_COUNT) return set_error(ctx, R3D_ERR_ARG, "bad scratch slot %d", slot);
    if (bytes <= ctx->scratch_bytes[slot]) return R3D_OK;
    if (ctx->scratch[slot]) {
        // other streams may still be using the old block
        cudaDeviceSynchronize();
        cudaFree(ctx->scratch[slot]);
        ctx->scratch[slot] = nullptr;
        ctx->scratch_bytes[slot] = 0;
    }
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&ctx->scratch[slot], want);
    if (e != cudaSuccess) {
        want = bytes;
        e = cudaMalloc(&ctx->scratch[slot], want);
    }
    if (e != cudaSuccess) {
        ctx->scratch[slot] = nullptr;
        return set_error(ctx, R3D_ERR_OOM, "cudaMalloc(%zu B scratch) failed: %s", want, cudaGetErrorString(e));
    }
    ctx->scratch_bytes[slot] = want;
    return R3D_OK;
}

bool is_device_ptr(const void* p, bool* pinned) {
    cudaPointerAttributes at;
    memset(&at, 0, sizeof at);
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        if (pinned) *pinned = false;
        return false;
    }
    if (pinned) *pinned = at.type == cudaMemoryTypeHost;
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

int device_set(r3d_ctx* ctx) {
    R3D_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    return R3D_OK;
}

int finish(r3d_ctx* ctx) {
    if (ctx->blocking) {
        R3D_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    R3D_CUDA_OK(ctx, cudaGetLastError());
    return R3D_OK;
}

}  // namespace r3d

using namespace r3d;

extern "C" const char* r3d_version(void) { return "r3d_b200 0.1 (sm_100a)"; }

extern "C" int r3d_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" r3d_ctx* r3d_create(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error(nullptr, R3D_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        cudaGetLastError();
        return nullptr;
    }
    if (device < 0 || device >= n) {
        set_error(nullptr, R3D_ERR_ARG, "device %d out of range (have %d)", device, n);
        return nullptr;
    }
    if ((e = cudaSetDevice(device)) != cudaSuccess) {
        set_error(nullptr, R3D_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
        return nullptr;
    }
    r3d_ctx* ctx = new r3d_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    // the context stream gets the highest priority: the scan pipeline runs its ray casts on side streams (default,
    // lowest priority) and the short list / emit / apply kernels of the previous scan must not queue behind them
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    bool ok = cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_hi) == cudaSuccess;
    for (int s = 0; s < 2 && ok; ++s) {
        ok = cudaStreamCreateWithFlags(&ctx->copy_stream[s], cudaStreamNonBlocking) == cudaSuccess;
    }
    for (int s = 0; s < r3d_ctx::kMaxStageSlots && ok; ++s) {
        ok = cudaEventCreateWithFlags(&ctx->ev_in[s], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->ev_k[s], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->ev_out[s], cudaEventDisableTiming) == cudaSuccess;
    }
    if (const char* v = getenv("R3D_STAGE_SLOTS")) {
        const int n = atoi(v);
        if (n >= 1 && n <= r3d_ctx::kMaxStageSlots) ctx->stage_slots = n;
    }
    if (const char* v = getenv("R3D_STAGE_CHUNK_MB")) {
        const long mb = atol(v);
        if (mb >= 1 && mb <= 8192) ctx->stage_chunk_bytes = (size_t)mb << 20;
    }
    ok = ok && cudaEventCreate(&ctx->ev_a) == cudaSuccess && cudaEventCreate(&ctx->ev_b) == cudaSuccess;
    if (const char* gb = getenv("R3D_SCAN_SCRATCH_GB")) {
        const double v = atof(gb);
        if (v >= 0) ctx->cell_budget_bytes = (uint64_t)(v * (double)(1ull << 30));
    }
    ctx->pinned_bytes = 4096;
    ok = ok && cudaHostAlloc(&ctx->pinned, ctx->pinned_bytes, cudaHostAllocDefault) == cudaSuccess;
    if (!ok) {
        set_error(nullptr, R3D_ERR_CUDA, "context resources: %s", cudaGetErrorString(cudaGetLastError()));
        r3d_destroy(ctx);
        return nullptr;
    }
    return ctx;
}

extern "C" void r3d_destroy(r3d_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < SCR_COUNT; ++i) if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
    scan_pipe_destroy(ctx);
    if (ctx->ztab) cudaFree(ctx->ztab);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->ev_a) cudaEventDestroy(ctx->ev_a);
    if (ctx->ev_b) cudaEventDestroy(ctx->ev_b);
    for (int s = 0; s < 2; ++s)
        if (ctx->copy_stream[s]) cudaStreamDestroy(ctx->copy_stream[s]);
    for (int s = 0; s < r3d_ctx::kMaxStageSlots; ++s) {
        if (ctx->ev_in[s]) cudaEventDestroy(ctx->ev_in[s]);
        if (ctx->ev_k[s]) cudaEventDestroy(ctx->ev_k[s]);
        if (ctx->ev_out[s]) cudaEventDestroy(ctx->ev_out[s]);
    }
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char* r3d_last_error(r3d_ctx* ctx) { return ctx ? ctx->err : g_last_error; }

extern "C" int r3d_set_blocking(r3d_ctx* ctx, int blocking) {
    if (!ctx) return set_error(nullptr, R3D_ERR_ARG, "null context");
    ctx->blocking = blocking ? 1 : 0;
    return R3D_OK;
}

extern "C" int r3d_synchronize(r3d_ctx* ctx) {
    if (!ctx) return set_error(nullptr, R3D_ERR_ARG, "null context");
    DeviceSetter ds(ctx->device);
    R3D_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    R3D_CUDA_OK(ctx, cudaStreamSynchronize(ctx->copy_stream[0]));
    R3D_CUDA_OK(ctx, cudaStreamSynchronize(ctx->copy_stream[1]));
    R3D_CUDA_OK(ctx, cudaGetLastError());
    return R3D_OK;
}

extern "C" void* r3d_stream(r3d_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
extern "C" uint64_t r3d_launch_count(r3d_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" float r3d_last_kernel_ms(r3d_ctx* ctx) { return ctx ? ctx->last_kernel_ms : 0.f; }

extern "C" void* r3d_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        set_error(nullptr, R3D_ERR_OOM, "cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    return p;
}
extern "C" void r3d_host_free(void* p) {
    if (p) cudaFreeHost(p);
}


extern "C" void* r3d_device_alloc(r3d_ctx* ctx, size_t bytes) {
    if (!ctx) { set_error(nullptr, R3D_ERR_ARG, "null context"); return nullptr; }
    DeviceSetter ds(ctx->device);
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error(ctx, R3D_ERR_OOM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        return nullptr;
    }
    return p;
}
extern "C" void r3d_device_free(r3d_ctx* ctx, void* p) {
    if (!ctx || !p) return;
    DeviceSetter ds(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(p);
}
extern "C" int r3d_memcpy(r3d_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx) return set_error(nullptr, R3D_ERR_ARG, "null context");
    if ((!dst || !src) && bytes) return set_error(ctx, R3D_ERR_ARG, "r3d_memcpy: null buffer");
    if (!bytes) return R3D_OK;
    DeviceSetter ds(ctx->device);
    R3D_CUDA_OK(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, ctx->stream));
    R3D_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return R3D_OK;
}
