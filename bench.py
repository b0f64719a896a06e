#!/usr/bin/env python
"""bench.py -- particle-steps/s of the MLS-MPM fluid step (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps 20 --warmup 3            # this solver (libmpm_b200.so through its C ABI)
    python bench.py --impl reference --steps 2 --warmup 1     # the reference algorithm on the host cores

A "step" is one pass of the hot path (bin -> clear -> P2G_1 -> P2G_2 -> grid update -> G2P) over the whole
particle set.  Workload at any N: BASELINE config 4 -- 3D dam-break, 256^3 grid, 32 768 000 particles (block
[4,164)^3 at spacing 0.5), parameters of the reference's shipping GPU scene
(MLSMPM3DFluidMultithreadGPU.cs:54-84), int32 x 1e7 fixed-point grid.  Arithmetic: --math fast (default; FMA and
re-association, parity within the tolerances stated in tests/test_parity_gpu.py) or --math strict (bit-exact against
the reference algorithm; the tiled kernels).  N > 1 slab-shards that same scene (strong scaling).  `value` is
device-timed with state resident in HBM; `e2e` is the same metric through the host-facing call sequence of one reference
frame (_Process, MLSMPM3DFluidMultithreadGPU.cs:234-251): parameter block in (set_sphere), step, positions (x,y,z,|v|) out
to pinned host memory (the particle_pos_tex hand-off).

Beside the headline (steps W..W+K of the collapsing block) the line carries what the headline flatters:
  evolved   the same scene after 100 steps (fluid spread out, pile-ups at the walls): ms/step, per-phase times, P2G+G2P fraction
  configs   BASELINE configs 2 and 3 (64^3 / 262 144 and 128^3 / 4 096 000 particles) on this GPU (N = 1)
  weak      weak scaling: a scene of N x 14.2 M particles (uniform pool, 64 x-planes per GPU), "scaling": "weak" inside it
  cpu_baseline / same_config_pair   the CPU port on the SAME config 4 scene (and on config 3 beside our config 3 number)
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "mls-mpm-godot_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

WORKLOADS = {  # name: (grid, block lo, block hi, spacing)  -- SURVEY.md 8d synthetic inputs
    "c2": ((64, 64, 64), (4, 4, 4), (36, 36, 36), 0.5),          # 262 144 particles
    "c3": ((128, 128, 128), (24, 24, 24), (104, 104, 104), 0.5),  # 4 096 000
    "c4": ((256, 256, 256), (4, 4, 4), (164, 164, 164), 0.5),     # 32 768 000
    # BASELINE config 5 (splash, deliberately non-uniform in x): the first block is listed here, the others in EXTRA_BLOCKS
    "c5": ((512, 256, 256), (4, 4, 4), (508, 68, 252), 0.5),      # pool: 63 995 904
}


def weak_workload(world):
    """Weak scaling scene: 64 x-planes per GPU, a pool filling x in [4, 64 N - 4), y in [4, 132), z in [4, 252) at spacing
    0.5: 112 x 256 x 496 = 14.2 M particles on one GPU, (128 N - 16) x 256 x 496 on N."""
    return (64 * world, 256, 256), (4, 4, 4), (64 * world - 4, 132, 252), 0.5
EXTRA_BLOCKS = {  # more lattice blocks appended to the scene (lo, hi, spacing)
    "c5": [((96, 90, 48), (256, 250, 208), 0.5),    # falling block, 32 768 000
           ((320, 90, 64), (448, 218, 192), 0.4)],  # dense block,   32 768 000
}
WORKLOAD_DESC = {
    "c2": "3D dam-break 64^3 grid, 262144 particles",
    "c3": "3D block-drop 128^3 grid, 4096000 particles",
    "c4": "3D dam-break 256^3 grid, 32768000 particles",
    "c5": "3D splash 512x256x256 grid, 129.5M particles (pool + falling block + dense block)",
}


def lattice_count(lo, hi, sp):
    """Points of `for (float i = lo; i < hi; i += sp)` per axis, multiplied (fp32 accumulation like the reference)."""
    import numpy as np
    n = 1
    for a, b in zip(lo, hi):
        k, x = 0, np.float32(a)
        while x < np.float32(b):
            k += 1
            x = np.float32(x + np.float32(sp))
        n *= k
    return n


def ncu_traffic(workload):
    """DRAM bytes per launch of each kernel from the committed `ncu --set full` capture of this round (profiles/r2/), else r1."""
    for rnd in ("r2", "r1"):
        try:
            with open(os.path.join(ROOT, "profiles", rnd, f"traffic_{workload}.json")) as f:
                d = json.load(f)
            return {k: v["dram_bytes"] for k, v in d["kernels"].items()}, d["source"]
        except Exception:
            continue
    return {}, None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index
        self.t_begin = self.t_end = None

    def start(self):
        """nvidia-smi needs ~1 s to attach, so it is started before the warm-up; samples are windowed afterwards."""
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def mark_end(self):
        self.t_end = time.perf_counter()

    def stop(self):
        if self.proc:
            self.proc.terminate()
        inside = [r for t, r in self.rows if self.t_begin is not None and self.t_begin - 0.02 <= t <= self.t_end + 0.02]
        window = "timed region"
        if not inside:  # region shorter than the sampling period: fall back to every sample taken under load
            inside, window = [r for _, r in self.rows], "warm-up + timed region + e2e"
        sm = sorted(int(r[0]) for r in inside if r and r[0].isdigit())
        mx = [int(r[1]) for r in inside if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in inside if len(r) >= 6 for k in range(4) if r[2 + k].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "window": window}


def scene_params(name):
    """CPU arms only (oracle parameter block of the same scene)."""
    from oracle import orc
    grid, lo, hi, sp = WORKLOADS[name]
    op = orc.variant("3d_gpu", grid)
    op.interaction = 0  # sphere disabled
    return op, lo, hi, sp


def cpu_scene(workload, sample):
    """Oracle state of the bench scene: the whole configured scene (every lattice block), or -- sample=True -- the bounded
    stand-in used when the whole scene would take too long on the host: the corner sub-block [4,84)^3 of the first lattice
    (4 096 000 particles) in a grid of at most 128^3."""
    import numpy as np
    from oracle import orc
    grid, lo, hi, sp = WORKLOADS[workload]
    if sample:
        hi = tuple(min(h, 84) for h in hi)
        grid = tuple(min(g, 128) for g in grid)
        blocks = [(lo, hi, sp)]
        what = f"sub-block {lo}-{hi} of the {WORKLOAD_DESC[workload]} lattice in a {grid} grid"
    else:
        blocks = [(lo, hi, sp)] + EXTRA_BLOCKS.get(workload, [])
        what = f"the whole scene: {WORKLOAD_DESC[workload]}"
    op = orc.variant("3d_gpu", grid)
    op.interaction = 0  # sphere disabled
    pos = np.concatenate([orc.init_block(3, b[0], b[1], b[2]) for b in blocks])
    return orc.State(op, pos), pos.shape[0], what


def run_reference(args):
    """The reference algorithm on the host cores: the C restatement (oracle, kind "port") in the reference's own
    threading shape (fixed-point variant: every phase across all cores with atomic int adds,
    MLSMPM3DFluidMultithreadNew.cs:277-288).  The .NET solver itself cannot run here (no dotnet/godot in the image).
    It steps the SAME scene as the GPU arm (config 4: 32.8 M particles, ~3 s per step on 16 cores) unless a probe step says the
    W + K steps would take longer than MPM_REF_BUDGET_S (default 300 s): then the bounded corner sample, and the line says so."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    budget = float(os.environ.get("MPM_REF_BUDGET_S", "300"))
    # probe on the small stand-in: the whole scene costs about (particles ratio) x as much per step
    probe, n_probe, _ = cpu_scene(args.workload, True)
    probe.step_mt(1, cores)
    t0 = time.perf_counter(); probe.step_mt(1, cores); t_probe = time.perf_counter() - t0
    grid, lo, hi, sp = WORKLOADS[args.workload]
    n_full = sum(lattice_count(*b) for b in [(lo, hi, sp)] + EXTRA_BLOCKS.get(args.workload, []))
    whole = t_probe * n_full / n_probe * (args.steps + args.warmup) <= budget
    del probe
    st, n, what = cpu_scene(args.workload, not whole)
    for _ in range(args.warmup):
        st.step_mt(1, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st.step_mt(1, cores)
    dt = time.perf_counter() - t0
    val = n * args.steps / dt
    sample = f"{what}: {n} particles, {args.steps} steps after {args.warmup}"
    line = {"impl": "reference", "metric": "particle-steps/s", "value": val, "unit": "particle-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32+int32-fixed-point", "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC[args.workload], "same_scene_as_gpu_arm": bool(whole), "sample": sample},
            "cpu_baseline": {"value": val, "unit": "particle-steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline(workload, budget_s=25.0):
    """The CPU port on the bench scene itself (one warm-up step, then as many steps as fit the budget, at least one); the
    corner sample instead if a single step of the whole scene would not fit."""
    cores = os.cpu_count() or 1
    st, n, what = cpu_scene(workload, True)
    st.step_mt(1, cores)
    t0 = time.perf_counter(); st.step_mt(1, cores); t_probe = time.perf_counter() - t0
    grid, lo, hi, sp = WORKLOADS[workload]
    n_full = sum(lattice_count(*b) for b in [(lo, hi, sp)] + EXTRA_BLOCKS.get(workload, []))
    whole = 3.0 * t_probe * n_full / n <= budget_s  # (a warm-up step and at least two timed ones)
    if whole:
        del st
        st, n, what = cpu_scene(workload, False)
        st.step_mt(1, cores)
    steps, t0 = 0, time.perf_counter()
    while steps < 8 and (steps == 0 or (time.perf_counter() - t0) * (steps + 1) / steps < budget_s):
        st.step_mt(1, cores)
        steps += 1
    dt = time.perf_counter() - t0
    return {"value": n * steps / dt, "unit": "particle-steps/s", "cores": cores, "kind": "port", "same_config": bool(whole),
            "sample": f"{what}: {n} particles, {steps} steps, oracle C restatement, all-core atomic fixed-point shape of "
                      f"MLSMPM3DFluidMultithreadNew.cs"}


def timed_steps(solver, steps, sync_ranks=lambda: None):
    """`steps` steps with two CUDA events around the whole mpm_step() call (timing level 2: the number that is reported),
    then `steps` more with events around every phase of every step (level 1: the breakdown; the 12-18 event records of a
    step cost a few percent of a sub-millisecond step, so that pass's own ms_step is reported beside, never instead)."""
    solver.set_timing(2)
    sync_ranks()
    solver.step(steps); solver.sync()
    sync_ranks()
    whole_ms = solver.stats().ms_step
    solver.set_timing(1)
    solver.step(steps); solver.sync()
    sync_ranks()
    st = solver.stats()
    solver.set_timing(0)
    return whole_ms, st


def time_scene(mpm_b200, grid, blocks, local_rank, math_mode, path, warmup, steps, presteps=0):
    """One more scene on this GPU (single solver, no communicator): device-timed ms/step (two events around the K steps)
    and, from a second copy of the scene over the same steps, the per-phase times."""
    params = mpm_b200.default_params("3d_gpu", grid=grid, interaction=0, kernel_path=path, math_mode=math_mode)
    n = sum(lattice_count(*b) for b in blocks)
    out = {}
    for level in (1, 2):
        with mpm_b200.Solver(params, n, device=local_rank) as sv:
            for k, (blo, bhi, bsp) in enumerate(blocks):
                sv.initialise_sim(blo, bhi, bsp, append=k > 0)
            if presteps:
                sv.step(presteps)
            sv.step(warmup); sv.sync()
            sv.set_timing(level)
            sv.step(steps); sv.sync()
            out[level] = sv.stats()
    st, whole_ms = out[1], out[2].ms_step
    G = grid[0] * grid[1] * grid[2]
    t3 = st.ms_p2g1 + st.ms_p2g2 + st.ms_g2p
    return {"particles": n, "grid": list(grid), "ms_per_step": whole_ms, "value": n / (whole_ms * 1e-3), "ms_per_step_phase_pass": st.ms_step,
            "phase_ms": {"sort": st.ms_sort, "clear": st.ms_clear, "p2g1": st.ms_p2g1, "p2g2": st.ms_p2g2, "update": st.ms_update, "g2p": st.ms_g2p},
            "p2g_g2p_gbs": (188 * n + 60 * G) / (t3 * 1e-3) / 1e9 if t3 > 0 else 0.0, "steps": steps, "warmup": warmup, "presteps": presteps}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--sort-interval", type=int, default=0)
    ap.add_argument("--path", type=int, default=0, help="0 auto, 1 reference-shaped, 2 tiled, 3 cell")
    ap.add_argument("--math", default="fast", choices=["strict", "fast"],
                    help="strict = bit-exact vs the reference algorithm; fast = FMA/hoisted (tolerance in tests)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--presteps", type=int, default=0, help="advance the scene this many untimed steps first (evolved-scene numbers)")
    ap.add_argument("--evolved-at", type=int, default=100, help="also time K steps from this step on (0 = off); reported under `evolved`")
    ap.add_argument("--no-extras", dest="extras", action="store_false", help="skip the config 2 / config 3 / weak-scaling measurements")
    ap.add_argument("--rebalance", type=int, default=-1,
                    help="multi-GPU: before the warm-up, this many rounds of (2 steps, 4 timed steps, mpm_comm_rebalance_weighted with the rank's "
                         "measured compute ms per million particles); default 0 (measured on config 5: profiles/r2/README.md)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import mpm_b200

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    grid, lo, hi, sp = WORKLOADS[args.workload]
    # the reference's shipping GPU scene constants (MLSMPM3DFluidMultithreadGPU.cs:54-84), sphere disabled
    params = mpm_b200.default_params("3d_gpu", grid=grid, interaction=0, kernel_path=args.path,
                                     sort_interval=args.sort_interval, math_mode=1 if args.math == "fast" else 0)
    blocks = [(lo, hi, sp)] + EXTRA_BLOCKS.get(args.workload, [])
    n_total = sum(lattice_count(*b) for b in blocks)
    G = grid[0] * grid[1] * grid[2]
    n_rebalance = max(args.rebalance, 0)

    def build_scene():
        sv = mpm_b200.Solver(params, n_total, device=local_rank)
        if world > 1:
            uid = [mpm_b200.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            sv.comm_init(uid[0], rank, world)
        # every rank generates the same lattice on its device; with a communicator the library cuts equal-count
        # x-slabs from the particle histogram and keeps the particles of its own slab
        for k, (blo, bhi, bsp) in enumerate(blocks):
            sv.initialise_sim(blo, bhi, bsp, append=k > 0)
        assert sv.num_particles == n_total, (sv.num_particles, n_total)
        for _ in range(n_rebalance if world > 1 else 0):  # cuts by measured time: what a host does between frames
            sv.step(2)          # (the first binning after an upload or a re-cut is a cold start: not what a step costs)
            sv.set_timing(1)
            sv.step(4); sv.sync()
            t = sv.stats()
            sv.set_timing(0)
            compute_ms = t.ms_sort + t.ms_clear + t.ms_p2g1 + t.ms_p2g2 + t.ms_update + t.ms_g2p
            sv.comm_rebalance(3, cost_per_particle=compute_ms / max(t.local_particles, 1) * 1e6)
        return sv

    # ---- pass A, the breakdown: the same scene, the same steps as the timed region below, with CUDA events around every
    # phase of every step (12-18 event records per step: a few percent of a sub-millisecond step, which is why the headline
    # is NOT taken from this pass).  Kernel durations of the roofline block come from here.
    solver = build_scene()
    if args.presteps > 0:
        solver.step(args.presteps)
    solver.step(max(args.warmup, 3))
    solver.sync()
    solver.set_timing(1)
    barrier()
    solver.step(args.steps)
    solver.sync()
    barrier()
    st = solver.stats()
    solver.set_timing(0)
    if os.environ.get("MPM_BENCH_ALLRANKS"):  # per-rank phase times (load balance / exchange skew), to stderr
        print(f"[rank {rank}] n_local={st.local_particles} cells={st.num_cells} ms_step={st.ms_step:.3f} sort={st.ms_sort:.3f} "
              f"p2g1={st.ms_p2g1:.3f} p2g2={st.ms_p2g2:.3f} update={st.ms_update:.3f} g2p={st.ms_g2p:.3f} exchange={st.ms_exchange:.3f} "
              f"(mass {st.ms_halo_mass:.3f} momentum {st.ms_halo_momentum:.3f} migration {st.ms_migration:.3f})",
              file=sys.stderr, flush=True)
    evolved_phases = None
    if args.evolved_at > 0 and args.workload == "c4":
        done = solver.stats().steps
        if done < args.evolved_at:
            solver.step(int(args.evolved_at - done))
        solver.sync()
        solver.set_timing(1)
        barrier()
        solver.step(args.steps); solver.sync()
        barrier()
        evolved_phases = solver.stats()
        solver.set_timing(0)
    solver.close()

    # ---- pass B, the number: a fresh copy of the scene (the stable binning makes it the same bits on one GPU)
    solver = build_scene()
    n_local = solver.stats().local_particles

    # ---- warm-up, then the timed region: K steps, device-timed on the solver's stream
    sampler = ClockSampler(local_rank)
    sampler.start()
    if args.presteps > 0:
        solver.step(args.presteps)
    solver.step(max(args.warmup, 3))
    solver.sync()
    solver.set_timing(2)  # two events around the K steps
    launches0 = solver.stats().kernel_launches
    barrier()
    sampler.mark_begin()
    t0 = time.perf_counter()
    solver.step(args.steps)   # mpm_step records CUDA events on its own stream around the K steps
    solver.sync()
    wall_ms = (time.perf_counter() - t0) * 1e3
    sampler.mark_end()
    barrier()
    st_whole = solver.stats()
    launches = st_whole.kernel_launches - launches0
    solver.set_timing(0)
    dev_ms = st_whole.ms_step * args.steps
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = n_total * args.steps / (total_ms * 1e-3)

    # ---- e2e: one reference frame per step through the C ABI with host buffers
    # (multi-GPU: migration changes the local count every step, so the pinned buffer is sized for the whole scene)
    # Two pinned buffers: every step's (x,y,z,|v|) array lands on the host, and the transfer of step k overlaps the
    # compute of step k+1 (mpm_get_positions_async), the way a renderer double-buffers the hand-off.
    pinned = [mpm_b200.host_alloc(16 * n_total) for _ in range(2)]
    e2e_steps = max(4, min(args.steps, 20))
    for k in range(2):  # untimed: creates the copy stream / second device array and touches both pinned buffers
        solver.step(1)
        solver.positions_into_async(pinned[k], n_total)
    solver.wait_positions()
    solver.sync()
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        solver.set_sphere((-21.648403 + 0.01 * k, 0.0, 31.707275))  # HandleMouseInteraction: params H2D each frame
        solver.step(1)
        solver.positions_into_async(pinned[k & 1], n_total)           # particle_pos_tex hand-off: 16 B/particle D2H
    solver.wait_positions()                                           # the last step's array is on the host too
    solver.sync()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = n_total * e2e_steps / float(t.item())
    # the same frame loop with the quantised hand-off (4 x uint16 per particle: half the bytes over PCIe); reported beside
    # e2e, never instead of it -- e2e is the reference-shaped float4 array
    solver.positions_q16_into_async(pinned[0], n_total); solver.wait_positions(); solver.sync()
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        solver.set_sphere((-21.648403 + 0.01 * k, 0.0, 31.707275))
        solver.step(1)
        solver.positions_q16_into_async(pinned[k & 1], n_total)
    solver.wait_positions()
    solver.sync()
    barrier()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_q16_val = n_total * e2e_steps / float(t.item())
    for ptr in pinned:
        mpm_b200.host_free(ptr)
    clocks = sampler.stop()

    # ---- the numbers the headline flatters (device-timed, same solver / same GPU)
    evolved = None
    if args.evolved_at > 0 and args.workload == "c4":
        done = solver.stats().steps
        if done < args.evolved_at:
            solver.step(int(args.evolved_at - done))
        solver.sync()
        solver.set_timing(2)
        barrier()
        solver.step(args.steps); solver.sync()
        barrier()
        e_whole = solver.stats().ms_step
        solver.set_timing(0)
        se = evolved_phases
        t = torch.tensor([e_whole], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ems = float(t.item())
        t3 = se.ms_p2g1 + se.ms_p2g2 + se.ms_g2p
        evolved = {"from_step": int(se.steps - args.steps), "steps": args.steps, "ms_per_step": ems, "ms_per_step_phase_pass": se.ms_step, "value": n_total / (ems * 1e-3),
                   "phase_ms": {"sort": se.ms_sort, "clear": se.ms_clear, "p2g1": se.ms_p2g1, "p2g2": se.ms_p2g2, "update": se.ms_update,
                                "g2p": se.ms_g2p, "exchange": se.ms_exchange},
                   "p2g_g2p_gbs": (188 * se.local_particles + 60 * se.num_cells) / (t3 * 1e-3) / 1e9 if t3 > 0 else 0.0,
                   "far_movers_last_binning": int(se.far_movers), "unordered_binnings": int(se.unordered_binnings)}
    st_final = solver.stats()
    sort_interval = solver.params.sort_interval or 1
    solver.close()
    solver = None
    configs, weak = None, None
    if args.extras and args.workload == "c4":
        mm = 1 if args.math == "fast" else 0
        if world == 1:
            configs = {}
            for name, k in (("c2", 50), ("c3", 30)):
                g, lo2, hi2, sp2 = WORKLOADS[name]
                configs[name] = dict(time_scene(mpm_b200, g, [(lo2, hi2, sp2)], local_rank, mm, args.path, 5, k), workload=WORKLOAD_DESC[name])
        # weak scaling: N x 14.2 M particles, 64 x-planes per GPU
        wg, wlo, whi, wsp = weak_workload(world)
        wparams = mpm_b200.default_params("3d_gpu", grid=wg, interaction=0, kernel_path=args.path, math_mode=mm)
        wn = lattice_count(wlo, whi, wsp)
        ws = mpm_b200.Solver(wparams, wn, device=local_rank)
        if world > 1:
            uid = [mpm_b200.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            ws.comm_init(uid[0], rank, world)
        ws.initialise_sim(wlo, whi, wsp)
        ws.step(5); ws.sync()
        w_whole, wst = timed_steps(ws, args.steps, barrier)
        ws.close()
        t = torch.tensor([w_whole], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wms = float(t.item())
        weak = {"scaling": "weak", "workload": f"pool x in [4, {wg[0] - 4}), y in [4, 132), z in [4, 252) at spacing 0.5 in a {wg[0]}x256x256 grid: "
                                                f"64 x-planes and ~14.2 M particles per GPU", "particles": wn, "n_gpus": world,
                "ms_per_step": wms, "value": wn / (wms * 1e-3), "steps": args.steps, "warmup": 5, "exchange_ms_rank0": wst.ms_exchange}

    if rank == 0:
        peak, peak_kind = peaks()
        phases = {"sort": st.ms_sort, "clear": st.ms_clear, "p2g1": st.ms_p2g1, "p2g2": st.ms_p2g2, "update": st.ms_update,
                  "g2p": st.ms_g2p, "exchange": st.ms_exchange}
        N, Gc = n_local, st.num_cells
        alg_bytes = {"p2g1": 64 * N + 16 * Gc, "p2g2": 52 * N + 28 * Gc, "g2p": 72 * N + 16 * Gc}
        dom = max(alg_bytes, key=lambda k: phases[k])
        achieved = alg_bytes[dom] / (phases[dom] * 1e-3) / 1e9 if phases[dom] > 0 else 0.0
        t3 = phases["p2g1"] + phases["p2g2"] + phases["g2p"]
        headline = (188 * N + 60 * Gc) / (t3 * 1e-3) / 1e9 if t3 > 0 else 0.0
        traffic, traffic_src = ncu_traffic(args.workload) if world == 1 else ({}, None)
        kname = {"p2g1": "k_p2g1_cell", "p2g2": "k_p2g2_cell", "g2p": "k_g2p_cell"} if st.kernel_path == 3 else {}
        per_kernel = {k: {"ms": phases[k], "algorithmic_bytes": alg_bytes[k],
                          "achieved_gbs": alg_bytes[k] / (phases[k] * 1e-3) / 1e9 if phases[k] > 0 else 0.0,
                          "frac": alg_bytes[k] / (phases[k] * 1e-3) / 1e9 / peak if phases[k] > 0 else 0.0,
                          "traffic": traffic.get(kname.get(k))} for k in alg_bytes}
        line = {
            "metric": "particle-steps/s", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32+int32-fixed-point", "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC[args.workload], "grid": list(grid), "particles": n_total, "variant": "3d_gpu (H)",
                       "grid_mode": "fixed 1e7", "math": args.math, "kernel_path": {1: "reference-shaped", 2: "tiled", 3: "cell"}[st.kernel_path],
                       "sort_interval": sort_interval, "parallelism": f"x-slab x{world}",
                       "presteps": args.presteps,
                       "rebalance": (f"{n_rebalance} rounds of 2 + 4 timed steps + mpm_comm_rebalance_weighted(3 planes, measured compute ms per "
                                     f"million particles) before the warm-up" if (world > 1 and n_rebalance) else "none (equal-count cuts from the upload)"),
                       "l2": f"inputs ({64e-9 * n_total / world:.1f} GB of particle planes per GPU) exceed the 126 MB L2; no flush needed",
                       "timing": "value / ms_per_step: two CUDA events on the solver stream around the K timed steps (mpm_set_timing 2), max over "
                                 "ranks; phase_ms / kernels / roofline: a second copy of the scene over the same steps with events around every "
                                 "phase of every step (mpm_set_timing 1; that pass's own ms/step is ms_per_step_phase_pass); wall-clock "
                                 "cross-check in wall_ms_per_step"},
            "wall_ms_per_step": wall_ms / args.steps,
            "ms_per_step_phase_pass": st.ms_step,
            "phase_ms": phases,
            "p2g_g2p_gbs": headline, "p2g_g2p_frac": headline / peak,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_kind": peak_kind, "traffic": traffic.get(kname.get(dom)), "traffic_source": traffic_src,
                         "algorithmic_bytes": alg_bytes[dom],
                         "note": "algorithmic bytes = SURVEY 8d per-unit figures x (local particles, local cells); duration = CUDA "
                                 "events around the kernel on the solver's stream, averaged over the same K steps in the per-phase pass"},
            "kernels": per_kernel,
            "e2e": {"value": e2e_val, "unit": "particle-steps/s", "h2d_bytes_per_step": 140, "d2h_bytes_per_step": 16 * n_local,
                    "what": "per step: mpm_set_sphere (140-B parameter block), mpm_step(1), mpm_get_positions_async -> pinned host "
                            "(two buffers: the D2H copy of step k overlaps step k+1; every step's array reaches the host)"},
            "e2e_q16": {"value": e2e_q16_val, "unit": "particle-steps/s", "d2h_bytes_per_step": 8 * n_local,
                        "what": "the same frame loop with mpm_get_positions_q16_async (x, y, z as uint16 fractions of the domain, |v| as binary16)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        line["binning"] = {"ranking": "atomic cursor" if (os.environ.get("MPM_ATOMIC_BINNING") or world > 1) else "stable (== std::stable_sort)",
                           "far_movers_last_binning": int(st_final.far_movers), "unordered_binnings": int(st_final.unordered_binnings)}
        if evolved is not None:
            evolved["p2g_g2p_frac"] = evolved["p2g_g2p_gbs"] / peak
            line["evolved"] = evolved
        if configs is not None:
            for c in configs.values():
                c["p2g_g2p_frac"] = c["p2g_g2p_gbs"] / peak
            line["configs"] = configs
        if weak is not None:
            line["weak"] = weak
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(args.workload)
            if configs is not None and args.workload != "c3":  # a second pair on a scene the CPU finishes quickly: config 3 on both sides
                cb3 = cpu_baseline("c3", budget_s=12.0)
                line["same_config_pair"] = {"workload": WORKLOAD_DESC["c3"], "ours": configs["c3"]["value"], "cpu_port": cb3["value"],
                                            "cores": cb3["cores"], "cpu_same_config": cb3["same_config"], "sample": cb3["sample"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    try:
        main()
    except BaseException:  # a failed rank must not sit in NCCL teardown while its peers wait in a collective
        import traceback
        traceback.print_exc()
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(1)
