import sys, os
ROOT = __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
sys.path[:0] = [ROOT, ROOT + '/mls-mpm-godot_b200', ROOT + '/tests']
import numpy as np
import mpm_b200
grid = (256, 256, 256)
p = mpm_b200.default_params("3d_gpu", grid=grid, interaction=0, kernel_path=3, math_mode=1)
with mpm_b200.Solver(p, 32768000) as s:
    s.initialise_sim((4, 4, 4), (164, 164, 164), 0.5)
    for steps in (100, 100):
        s.step(steps)
        pos, vel, C, m = s.download()
        c = pos.astype(np.int32)
        key = (c[:, 0].astype(np.int64) << 20) | (c[:, 1].astype(np.int64) << 10) | c[:, 2]
        u, cnt = np.unique(key, return_counts=True)
        hist = np.bincount(np.minimum(cnt, 40))
        print(os.environ.get("MPM_NO_FUSED_UPDATE", "fused"), "cells", u.size, "max/cell", cnt.max(), "top", np.sort(cnt)[-5:], "mean", cnt.mean().round(2),
              "hist[1,2,4,8,16,32,40+]", [int(hist[k]) if k < hist.size else 0 for k in (1, 2, 4, 8, 16, 32, 40)],
              "ke", float((vel.astype(np.float64) ** 2).sum() / 2), "max|v|", float(np.abs(vel).max()),
              "bbox", pos.min(0).round(2), pos.max(0).round(2), flush=True)
