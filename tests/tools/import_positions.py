"""Child process of test_zero_copy_hand_off: imports the solver's position array from an inherited file descriptor with the
CUDA driver API (what a renderer does once with Vulkan's VK_KHR_external_memory_fd) and prints a SHA-256 of its bytes."""
import hashlib
import sys

import numpy as np
from cuda.bindings import driver as cu


def ck(res):
    err = res[0]
    if err != cu.CUresult.CUDA_SUCCESS:
        raise RuntimeError(f"CUDA driver error {err}")
    return res[1] if len(res) == 2 else res[1:]


fd, nbytes, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ck(cu.cuInit(0))
dev = ck(cu.cuDeviceGet(0))
ctx = ck(cu.cuDevicePrimaryCtxRetain(dev))
ck(cu.cuCtxSetCurrent(ctx))
h = ck(cu.cuMemImportFromShareableHandle(fd, cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR))
va = ck(cu.cuMemAddressReserve(nbytes, 0, 0, 0))
ck(cu.cuMemMap(va, nbytes, 0, h, 0))
acc = cu.CUmemAccessDesc()
acc.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
acc.location.id = 0
acc.flags = cu.CUmemAccess_flags.CU_MEM_ACCESS_FLAGS_PROT_READ
ck(cu.cuMemSetAccess(va, nbytes, [acc], 1))
out = np.zeros((n, 4), np.float32)
ck(cu.cuMemcpyDtoH(out.ctypes.data, va, 16 * n))
print("SHA256", hashlib.sha256(out.tobytes()).hexdigest(), flush=True)
ck(cu.cuMemUnmap(va, nbytes))
ck(cu.cuMemAddressFree(va, nbytes))
ck(cu.cuMemRelease(h))
