"""Child process of test_programmatic_dependent_launch_changes_no_bit: steps one scene on the FAST cell path and prints a
SHA-256 of the particle state and the grid (the environment of the child decides how the kernels are launched)."""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mls-mpm-godot_b200"), os.path.join(ROOT, "tests")]
import mpm_b200  # noqa: E402

grid, steps = int(sys.argv[1]), int(sys.argv[2])
lo, hi = (grid // 8, grid // 4, grid // 8), (grid // 2, grid * 3 // 4, grid // 2)
p = mpm_b200.default_params("3d_gpu", grid=(grid, grid, grid), interaction=0, kernel_path=3, math_mode=1)
h = hashlib.sha256()
with mpm_b200.Solver(p, 1 << 22) as s:
    s.initialise_sim(lo, hi, 0.5)
    s.step(steps)
    for a in s.download():
        h.update(a.tobytes())
    h.update(s.download_grid().tobytes())
    st = s.stats()
print(st.unordered_binnings, st.kernel_launches, h.hexdigest())
