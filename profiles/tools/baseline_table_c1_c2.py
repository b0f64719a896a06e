"""GPU ms/step (CUDA events via the solver's own timing) and the CPU port beside it."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "mls-mpm-godot_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import mpm_b200
from oracle import orc
import helpers

out = {}
# config 1: 2D dam-break 128^2, block [4,68)^2 at spacing 0.5 = 16384 particles, parameters of D (reference-shaped path, float grid)
op = orc.variant("2d_st", (128, 128, 1))
pos = orc.init_block(2, (4, 4), (68, 68), 0.5)
with mpm_b200.Solver(helpers.mpm_params_from_orc(op), pos.shape[0]) as s:
    s.initialise_sim((4, 4), (68, 68), 0.5)
    s.step(50); s.sync(); s.set_timing(True); s.step(400); s.sync()
    st = s.stats()
ref = orc.State(op, pos); ref.step(5)
t0 = time.perf_counter(); ref.step(200); dt = time.perf_counter() - t0
out["c1"] = dict(n=int(pos.shape[0]), gpu_ms=st.ms_step, gpu_rate=pos.shape[0] / (st.ms_step * 1e-3), cpu_rate=pos.shape[0] * 200 / dt, cpu_threads=1,
                 path=int(st.kernel_path))
# config 2: 3D dam-break 64^3, 262144 particles, parameters of H; CPU: all-core fixed-point shape
op = orc.variant("3d_gpu", 64); op.interaction = 0
pos = orc.init_block(3, (4, 4, 4), (36, 36, 36), 0.5)
ref = orc.State(op, pos); ref.step_mt(3)
t0 = time.perf_counter(); ref.step_mt(200); dt = time.perf_counter() - t0
out["c2"] = dict(n=int(pos.shape[0]), cpu_rate=pos.shape[0] * 200 / dt, cpu_threads=os.cpu_count())
print(json.dumps(out))
