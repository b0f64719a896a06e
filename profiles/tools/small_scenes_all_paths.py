import sys
ROOT = __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
sys.path[:0] = [ROOT, ROOT + '/mls-mpm-godot_b200', ROOT + '/tests']
import numpy as np, helpers, mpm_b200
from oracle import orc
# small scenes through every kernel family: reference-shaped, tiled strict, tiled fast, cell (incl. a pile-up and 2 slabs)
op = orc.variant("3d_gpu", 32); op.interaction = 0
pos, vel, Cm, mass = helpers.random_cloud(op, 6000, seed=1)
for kp, mm in ((1, 0), (2, 0), (2, 1), (3, 1)):
    with mpm_b200.Solver(helpers.mpm_params_from_orc(op, kernel_path=kp, math_mode=mm), 8000) as s:
        s.upload(pos, vel, Cm, mass); s.step(3); s.download(); s.positions(); s.download_grid()
# pile-up + ragged on the cell path (virtual cells)
rng = np.random.default_rng(2)
same = (np.array([[12.0, 13.0, 14.0]], np.float32) + rng.uniform(0.01, 0.99, (1500, 3)).astype(np.float32))
with mpm_b200.Solver(helpers.mpm_params_from_orc(op, kernel_path=3, math_mode=1, rest_density=40.0), 2000) as s:
    s.upload(same, mass=np.full(1500, 0.01, np.float32)); s.step(2); s.download()
# B = 8 blocks
op8 = orc.variant("3d_gpu", 96); op8.interaction = 0
p8, v8, c8, m8 = helpers.random_cloud(op8, 20000, seed=3)
with mpm_b200.Solver(helpers.mpm_params_from_orc(op8, kernel_path=3, math_mode=1), 20000) as s:
    s.upload(p8, v8, c8, m8); s.step(3); s.download()
print("sanitize scenes done")
