"""Development aid: run the stable-binning scenario with MPM_DEBUG_BIN=1 and report which side of the verification is off."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "mls-mpm-godot_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
os.environ["MPM_DEBUG_BIN"] = "1"
import helpers, mpm_b200
from oracle import orc

for grid, n, crowd in ((32, 20000, False), (32, 200001, True), (96, 200001, True)):
    op = orc.variant("3d_gpu", grid); op.interaction = 0
    pos, vel, Cm, mass = helpers.random_cloud(op, n, seed=33, vel_sigma=0.8)
    if crowd:
        pos[: n // 2] = pos[: n // 2] * 0.25 + 4.0
        mass[: n // 2] *= 0.05
    with mpm_b200.Solver(helpers.mpm_params_from_orc(op, kernel_path=3, math_mode=1), n) as s:
        s.upload(pos, vel, Cm, mass)
        for rnd in range(3):
            if rnd:
                s.step(2)
            s.run_phase(5)
            try:
                keys, perm = s.last_sort()
                ok = np.array_equal(perm.astype(np.int64), np.argsort(keys, kind="stable"))
                print(grid, n, crowd, "round", rnd, "perm == stable sort:", ok, flush=True)
            except mpm_b200.MpmError as e:
                print(grid, n, crowd, "round", rnd, "ERROR", e, flush=True)
            s.download()
