"""D2H bandwidth into pinned host memory allocated while the process is bound to each NUMA node's CPUs.
Run on the GPU box: python profiles/tools/numa_d2h_probe.py   (prints one line per node)."""
import glob
import os
import subprocess
import sys

if len(sys.argv) == 1:
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout)
    for d in glob.glob("/sys/bus/pci/devices/*/numa_node"):
        try:
            cls = open(os.path.dirname(d) + "/class").read().strip()
            if cls.startswith("0x0302") or cls.startswith("0x0300"):
                print(os.path.dirname(d), "numa_node", open(d).read().strip())
        except OSError:
            pass
    for node in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
        cpus = open(node + "/cpulist").read().strip()
        out = subprocess.run(["taskset", "-c", cpus, sys.executable, __file__, "child"], capture_output=True, text=True)
        print(os.path.basename(node), "cpus", cpus, "->", out.stdout.strip(), out.stderr.strip()[-200:])
else:
    import time
    import torch
    n = 512 << 20
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    host.copy_(dev); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        host.copy_(dev, non_blocking=True)
    torch.cuda.synchronize()
    print("D2H %.1f GB/s" % (5 * n / (time.perf_counter() - t0) / 1e9))
