// vm_alloc.cpp -- see vm_alloc.hpp.  Driver entry points are fetched through the
// runtime (cudaGetDriverEntryPoint), so the library has no link dependency on libcuda.
#include "vm_alloc.hpp"

#include <cuda.h>
#include <cuda_runtime.h>

namespace mb2 {

namespace {

struct Driver {
    CUresult (*getGranularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*addressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*addressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*create)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long) = nullptr;
    CUresult (*release)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*unmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*setAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
    bool ok = false;
};

template <typename Fn>
bool fetch(const char *name, Fn *out)
{
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return false;
    *out = (Fn)fn;
    return true;
}

Driver &driver()
{
    static Driver d = [] {
        Driver x;
        x.ok = fetch("cuMemGetAllocationGranularity", &x.getGranularity) &&
               fetch("cuMemAddressReserve", &x.addressReserve) && fetch("cuMemAddressFree", &x.addressFree) &&
               fetch("cuMemCreate", &x.create) && fetch("cuMemRelease", &x.release) &&
               fetch("cuMemMap", &x.map) && fetch("cuMemUnmap", &x.unmap) && fetch("cuMemSetAccess", &x.setAccess);
        return x;
    }();
    return d;
}

CUmemAllocationProp propFor(int gpu)
{
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = gpu;
    return prop;
}

size_t granularity(int gpu)
{
    static size_t cached[64] = {};
    if (gpu >= 0 && gpu < 64 && cached[gpu]) return cached[gpu];
    size_t g = 2u << 20;
    CUmemAllocationProp prop = propFor(gpu);
    if (driver().ok) driver().getGranularity(&g, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM);
    if (gpu >= 0 && gpu < 64) cached[gpu] = g;
    return g;
}

size_t roundUp(size_t v, size_t g) { return (v + g - 1) / g * g; }

}

bool vmReserve(int gpu, VMRange *r, size_t reserve_bytes, size_t initial_bytes, std::string *err)
{
    Driver &d = driver();
    if (!d.ok) {
        *err = "CUDA virtual memory management entry points are unavailable";
        return false;
    }
    const size_t g = granularity(gpu);
    reserve_bytes = roundUp(std::max(reserve_bytes, initial_bytes), g);
    CUdeviceptr base = 0;
    if (d.addressReserve(&base, reserve_bytes, 0, 0, 0) != CUDA_SUCCESS) {
        *err = "cuMemAddressReserve failed (" + std::to_string(reserve_bytes) + " bytes)";
        return false;
    }
    r->base = (void *)base;
    r->reserved = reserve_bytes;
    r->mapped = 0;
    return vmGrow(gpu, r, initial_bytes, err);
}

bool vmGrow(int gpu, VMRange *r, size_t new_bytes, std::string *err)
{
    if (new_bytes <= r->mapped) return true;
    Driver &d = driver();
    const size_t g = granularity(gpu);
    const size_t target = roundUp(new_bytes, g);
    if (target > r->reserved) {
        *err = "allocation outgrew its reserved address range (" + std::to_string(r->reserved) + " bytes)";
        return false;
    }
    const size_t delta = target - r->mapped;
    CUmemAllocationProp prop = propFor(gpu);
    CUmemGenericAllocationHandle h = 0;
    if (d.create(&h, delta, &prop, 0) != CUDA_SUCCESS) {
        *err = "cuMemCreate failed: out of device memory (" + std::to_string(delta) + " bytes)";
        return false;
    }
    const CUdeviceptr at = (CUdeviceptr)r->base + r->mapped;
    if (d.map(at, delta, 0, h, 0) != CUDA_SUCCESS) {
        d.release(h);
        *err = "cuMemMap failed";
        return false;
    }
    CUmemAccessDesc access = {};
    access.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    access.location.id = gpu;
    access.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    if (d.setAccess(at, delta, &access, 1) != CUDA_SUCCESS) {
        *err = "cuMemSetAccess failed";
        return false;
    }
    r->handles.push_back(h);
    r->handleBytes.push_back(delta);
    cudaMemset((void *)at, 0, delta);
    r->mapped = target;
    return true;
}

void vmRelease(VMRange *r)
{
    Driver &d = driver();
    if (!r->base || !d.ok) return;
    size_t off = 0;
    for (size_t i = 0; i < r->handles.size(); i++) {
        d.unmap((CUdeviceptr)r->base + off, r->handleBytes[i]);
        d.release(r->handles[i]);
        off += r->handleBytes[i];
    }
    d.addressFree((CUdeviceptr)r->base, r->reserved);
    r->base = nullptr;
    r->handles.clear();
    r->handleBytes.clear();
    r->mapped = r->reserved = 0;
}

}
