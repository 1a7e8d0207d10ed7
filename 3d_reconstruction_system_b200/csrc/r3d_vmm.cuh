// r3d_vmm.cuh -- a device buffer that GROWS IN PLACE: a virtual address range is reserved up front (the map's brick pool may
// take most of a B200's 180 GB) and physical memory is mapped behind it chunk by chunk as the map grows.  Growing copies
// nothing, frees nothing and synchronises nothing, and the base pointer never changes, so kernels in flight on any stream
// stay valid -- a cudaMalloc'd pool had to be re-allocated, copied (GBs) and freed on every doubling, which cost up to
// 250 ms of host time inside a 4 500-scan run.  Driver entry points are fetched with cudaGetDriverEntryPoint (no link
// dependency on libcuda: the library still loads where there is no driver, e.g. for the CPU-side ABI tests).
#pragma once
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>

namespace r3d {

struct VmmRegion {
    CUdeviceptr base = 0;
    size_t reserved = 0;     // bytes of address space
    size_t mapped = 0;       // bytes backed by memory, always a prefix
    size_t chunk = 0;        // allocation granularity: every mapping is a multiple of it
    int device = 0;
    std::vector<CUmemGenericAllocationHandle> handles;   // ONE physical allocation per growth step (a step is three driver calls,
                                                          // each of which may wait for running kernels: few, large steps)
};

bool vmm_supported(int device);
// reserve `max_bytes` of address space; nothing is mapped yet
bool vmm_reserve(VmmRegion* r, int device, size_t max_bytes);
// make at least `want_bytes` usable; false when the device (or the reservation) has no more room
bool vmm_grow(VmmRegion* r, size_t want_bytes);
void vmm_release(VmmRegion* r);

}  // namespace r3d
