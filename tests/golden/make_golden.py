"""Generates tests/golden/*.npz -- run once, by hand, from the repo root: python tests/golden/make_golden.py

PARITY UNPINNED: the reference (C#/GLSL, Godot 4.5) cannot run in this image and ships no vectors, so
these fixtures are produced by the *NumPy* restatement (oracle/oracle_np.py), which is independent of the
C oracle and of the CUDA solver; tests then require the C oracle (CPU tests) and the CUDA solver (-m gpu)
to reproduce them.  Inputs are seeded random clouds so every term (C, mass, interaction, walls) is live.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import orc, oracle_np as onp  # noqa: E402
import helpers  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CASES = {  # name: (variant, grid, n, steps, seed)
    "2d_st": ("2d_st", (32, 32, 1), 600, 4, 11),
    "2d_mt": ("2d_mt", (32, 32, 1), 600, 4, 12),
    "3d_float": ("3d_float", (16, 16, 16), 700, 3, 13),
    "3d_fixed": ("3d_fixed", (16, 16, 16), 700, 3, 14),
    "3d_gpu": ("3d_gpu", (24, 16, 16), 900, 3, 15),  # non-cubic on purpose
}


def case_params(name):
    variant, grid, n, steps, seed = CASES[name]
    p = orc.variant(variant, grid)
    if p.interaction in (1, 2):
        p.sphere_pos[:] = [grid[0] * 0.3, grid[1] * 0.5, grid[2] * 0.5]
        p.sphere_radius = 4.0
    return p, n, steps, seed


def main():
    for name in CASES:
        p, n, steps, seed = case_params(name)
        pos, vel, Cm, mass = helpers.random_cloud(p, n, seed)
        G = orc.num_cells(p)
        grid = np.zeros((G, 4), np.int32)
        q = dict(pos=pos.copy(), vel=vel.copy(), C=Cm.copy())
        onp.step(p, q["pos"], q["vel"], q["C"], mass, grid, steps)
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), pos0=pos, vel0=vel, C0=Cm, mass=mass,
                            pos=q["pos"], vel=q["vel"], C=q["C"], grid=grid, steps=steps)
        print(name, n, "particles", steps, "steps", os.path.getsize(os.path.join(OUT, f"{name}.npz")), "bytes")


if __name__ == "__main__":
    main()
