// r3d_vmm.cu -- see r3d_vmm.cuh.
#include "r3d_vmm.cuh"

namespace r3d {

namespace {
struct Api {
    CUresult (*memAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*memAddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*memCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*memRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*memMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*memUnmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*memSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    CUresult (*memGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
    bool ok = false;
};

const Api& api() {
    static Api a = [] {
        Api x;
        auto get = [](const char* name, void** fn) {
            cudaDriverEntryPointQueryResult st;
            return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &st) == cudaSuccess && st == cudaDriverEntryPointSuccess && *fn != nullptr;
        };
        x.ok = get("cuMemAddressReserve", (void**)&x.memAddressReserve) && get("cuMemAddressFree", (void**)&x.memAddressFree) &&
               get("cuMemCreate", (void**)&x.memCreate) && get("cuMemRelease", (void**)&x.memRelease) && get("cuMemMap", (void**)&x.memMap) &&
               get("cuMemUnmap", (void**)&x.memUnmap) && get("cuMemSetAccess", (void**)&x.memSetAccess) &&
               get("cuMemGetAllocationGranularity", (void**)&x.memGetAllocationGranularity);
        if (!x.ok) cudaGetLastError();
        return x;
    }();
    return a;
}

CUmemAllocationProp prop_for(int device) {
    CUmemAllocationProp p = {};
    p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    p.location.id = device;
    return p;
}
}  // namespace

bool vmm_supported(int device) {
    if (!api().ok) return false;
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMemoryPoolsSupported, device) != cudaSuccess) { cudaGetLastError(); return false; }
    size_t gran = 0;
    const CUmemAllocationProp p = prop_for(device);
    return api().memGetAllocationGranularity(&gran, &p, CU_MEM_ALLOC_GRANULARITY_MINIMUM) == CUDA_SUCCESS && gran > 0;
}

bool vmm_reserve(VmmRegion* r, int device, size_t max_bytes) {
    if (!api().ok) return false;
    size_t gran = 0;
    const CUmemAllocationProp p = prop_for(device);
    if (api().memGetAllocationGranularity(&gran, &p, CU_MEM_ALLOC_GRANULARITY_MINIMUM) != CUDA_SUCCESS || gran == 0) return false;
    max_bytes = (max_bytes + gran - 1) / gran * gran;
    CUdeviceptr base = 0;
    if (api().memAddressReserve(&base, max_bytes, 0, 0, 0) != CUDA_SUCCESS) return false;
    r->base = base; r->reserved = max_bytes; r->mapped = 0; r->chunk = gran; r->device = device;
    r->handles.clear();
    return true;
}

bool vmm_grow(VmmRegion* r, size_t want_bytes) {
    if (want_bytes <= r->mapped) return true;
    if (want_bytes > r->reserved) return false;
    const CUmemAllocationProp p = prop_for(r->device);
    CUmemAccessDesc acc = {};
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = r->device;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    size_t size = (want_bytes - r->mapped + r->chunk - 1) / r->chunk * r->chunk;
    if (r->mapped + size > r->reserved) size = r->reserved - r->mapped;
    CUmemGenericAllocationHandle h;
    if (api().memCreate(&h, size, &p, 0) != CUDA_SUCCESS) return false;
    if (api().memMap(r->base + r->mapped, size, 0, h, 0) != CUDA_SUCCESS) { api().memRelease(h); return false; }
    if (api().memSetAccess(r->base + r->mapped, size, &acc, 1) != CUDA_SUCCESS) {
        api().memUnmap(r->base + r->mapped, size);
        api().memRelease(h);
        return false;
    }
    r->handles.push_back(h);
    r->mapped += size;
    return true;
}

void vmm_release(VmmRegion* r) {
    if (!r->base) return;
    if (r->mapped) api().memUnmap(r->base, r->mapped);
    for (auto h : r->handles) api().memRelease(h);
    api().memAddressFree(r->base, r->reserved);
    r->handles.clear();
    r->base = 0; r->reserved = r->mapped = 0;
}

}  // namespace r3d
