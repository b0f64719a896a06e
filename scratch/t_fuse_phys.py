import sys, os, subprocess, json
sys.path[:0] = ['/root/repo', '/root/repo/mls-mpm-godot_b200', '/root/repo/tests']
import numpy as np
import mpm_b200
mode = sys.argv[1]
grid = (128, 128, 128)
if mode == "strict":
    p = mpm_b200.default_params("3d_gpu", grid=grid, interaction=0, kernel_path=2, math_mode=0)
else:
    p = mpm_b200.default_params("3d_gpu", grid=grid, interaction=0, kernel_path=3, math_mode=1)
with mpm_b200.Solver(p, 4096000) as s:
    s.initialise_sim((4, 4, 4), (84, 84, 84), 0.5)
    for chunk in range(4):
        s.step(50)
        pos, vel, C, m = s.download()
        cells = np.unique((pos.astype(np.int32) * np.array([1 << 20, 1 << 10, 1])).sum(1)).size
        blocks = np.unique(((pos.astype(np.int32) >> 3) * np.array([1 << 20, 1 << 10, 1])).sum(1)).size
        print(mode, os.environ.get("MPM_NO_FUSED_UPDATE", "-"), "step", 50 * (chunk + 1), "com", pos.mean(0).round(3), "ke", float((vel.astype(np.float64) ** 2).sum()) / 2,
              "max|v|", float(np.abs(vel).max()), "cells", cells, "blocks", blocks, "ymax", float(pos[:, 1].max()), flush=True)
