// mpm_vmm.cu -- exportable device memory for the renderer hand-off.
//
// The reference never leaves the GPU: G2P writes an rgba32f storage image that the MultiMesh shader samples
// (MLSMPM3DFluidMultithreadGPU.cs:340-355, 402-412; g2p.glsl:149-150).  A cudaMalloc pointer cannot be imported by
// Vulkan / Godot, so the (x, y, z, |v|) array lives in an allocation made with the driver's virtual memory management
// API (cuMemCreate with a POSIX-file-descriptor handle type): mpm_export_positions hands out a file descriptor that
// another API or process imports -- Vulkan: VkImportMemoryFdInfoKHR with VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT;
// CUDA: cuMemImportFromShareableHandle / cudaImportExternalMemory -- and reads the array the solver keeps refreshing,
// without a host round trip.  The driver entry points are resolved through cudaGetDriverEntryPoint: the library does not
// link against libcuda.
#include <cuda.h>
#include <cuda_runtime.h>

#include <string>

#include "mpm_solver.h"

namespace mpm {

namespace {
struct Drv {
    CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
    CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
    bool ok = false;
};

template <class F>
bool entry(const char* name, F* fn)
{
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || !p || q != cudaDriverEntryPointSuccess) { cudaGetLastError(); return false; }
    *fn = reinterpret_cast<F>(p);
    return true;
}

const Drv& drv()
{
    static Drv d;
    static bool tried = false;
    if (!tried) {
        tried = true;
        d.ok = entry("cuMemGetAllocationGranularity", &d.MemGetAllocationGranularity) && entry("cuMemCreate", &d.MemCreate) &&
               entry("cuMemAddressReserve", &d.MemAddressReserve) && entry("cuMemMap", &d.MemMap) && entry("cuMemSetAccess", &d.MemSetAccess) &&
               entry("cuMemExportToShareableHandle", &d.MemExportToShareableHandle) && entry("cuMemUnmap", &d.MemUnmap) &&
               entry("cuMemRelease", &d.MemRelease) && entry("cuMemAddressFree", &d.MemAddressFree);
    }
    return d;
}
}  // namespace

// exportable allocation of at least `bytes` on `device`; false if the driver / device cannot (the caller falls back to cudaMalloc)
bool vmm_alloc(int device, size_t bytes, ExportableAlloc* out)
{
    const Drv& d = drv();
    if (!d.ok) return false;
    cudaFree(nullptr);  // (make sure the primary context exists and is current)
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = device;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t gran = 0;
    if (d.MemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM) != CUDA_SUCCESS || gran == 0) return false;
    const size_t size = (bytes + gran - 1) / gran * gran;
    CUmemGenericAllocationHandle h = 0;
    if (d.MemCreate(&h, size, &prop, 0) != CUDA_SUCCESS) return false;
    CUdeviceptr va = 0;
    if (d.MemAddressReserve(&va, size, 0, 0, 0) != CUDA_SUCCESS) { d.MemRelease(h); return false; }
    if (d.MemMap(va, size, 0, h, 0) != CUDA_SUCCESS) { d.MemAddressFree(va, size); d.MemRelease(h); return false; }
    CUmemAccessDesc acc = {};
    acc.location = prop.location;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    if (d.MemSetAccess(va, size, &acc, 1) != CUDA_SUCCESS) { d.MemUnmap(va, size); d.MemAddressFree(va, size); d.MemRelease(h); return false; }
    out->ptr = reinterpret_cast<void*>(va);
    out->bytes = size;
    out->handle = (unsigned long long)h;
    return true;
}

void vmm_free(ExportableAlloc* a)
{
    if (!a->ptr) return;
    const Drv& d = drv();
    d.MemUnmap((CUdeviceptr)a->ptr, a->bytes);
    d.MemAddressFree((CUdeviceptr)a->ptr, a->bytes);
    d.MemRelease((CUmemGenericAllocationHandle)a->handle);
    a->ptr = nullptr; a->bytes = 0; a->handle = 0;
}

// a new file descriptor for the allocation (the caller owns and closes it)
bool vmm_export_fd(const ExportableAlloc& a, int* fd)
{
    const Drv& d = drv();
    if (!d.ok || !a.ptr) return false;
    int out = -1;
    if (d.MemExportToShareableHandle(&out, (CUmemGenericAllocationHandle)a.handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) != CUDA_SUCCESS) return false;
    *fd = out;
    return true;
}

}  // namespace mpm
