import sys, os, threading, time
sys.path[:0] = ['/root/repo', '/root/repo/mls-mpm-godot_b200', '/root/repo/tests']
import numpy as np, helpers, mpm_b200
from oracle import orc
op = orc.variant("3d_gpu", (256, 96, 96)); op.interaction = 0
n = 400000
pos, vel, Cm, mass = helpers.random_cloud(op, n, seed=3, vel_sigma=1.0)
world = 2
hub = mpm_b200.LocalHub(world)
out = [None]*world
def work(r):
    try:
        with mpm_b200.Solver(helpers.mpm_params_from_orc(op, kernel_path=3, math_mode=1), n) as s:
            s.comm_init_local(hub, r, world)
            s.upload(pos, vel, Cm, mass)
            print(r, "slab", s.slab(), s.stats().local_particles, flush=True)
            for k in range(6):
                s.step(1); s.sync()
                print(r, "step", k, s.stats().local_particles, flush=True)
            out[r] = s.download_ids().shape
    except Exception as e:
        print(r, "ERR", e, flush=True)
th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
[t.start() for t in th]; [t.join(100) for t in th]
print(out)
