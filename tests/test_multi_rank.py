"""Multi-GPU x-slabs.  CPU part: slab cuts (host logic) alone and across two gloo ranks.
GPU part (-m gpu, one device): k solvers in one process on the LOCAL transport -- the same halo exchange-add and
migration code the NCCL transport runs -- must reproduce the single-solver result bit for bit in fixed-point mode."""
import os
import socket
import threading

import numpy as np
import pytest

import helpers
from oracle import orc


# ---------------------------------------------------------------- CPU: slab cuts
def test_slab_cuts_balance_and_constraints(lib):
    import mpm_b200
    rng = np.random.default_rng(3)
    hist = np.zeros(256, np.int64)
    hist[4:164] = rng.integers(150_000, 250_000, 160)           # dam-break block occupies x in [4, 164)
    for world in (1, 2, 3, 4, 8):
        cuts = mpm_b200.slab_cuts(hist, world)
        assert cuts[0] == 0 and cuts[-1] == 256 and len(cuts) == world + 1
        assert all(b - a >= 4 for a, b in zip(cuts, cuts[1:]))
        per = [hist[a:b].sum() for a, b in zip(cuts, cuts[1:])]
        assert sum(per) == hist.sum()
        assert max(per) - min(per) <= 2 * hist.max(), (world, per)   # equal counts up to one plane's worth
    # degenerate inputs: everything in one plane, empty histogram, too many ranks
    one = np.zeros(64, np.int64); one[10] = 1000
    cuts = mpm_b200.slab_cuts(one, 4)
    assert cuts[0] == 0 and cuts[-1] == 64 and all(b - a >= 4 for a, b in zip(cuts, cuts[1:]))
    cuts = mpm_b200.slab_cuts(np.zeros(64, np.int64), 4)
    assert cuts[-1] == 64 and all(b - a >= 4 for a, b in zip(cuts, cuts[1:]))
    with pytest.raises(mpm_b200.MpmError):
        mpm_b200.slab_cuts(np.ones(8, np.int64), 4)


def test_slab_cuts_of_a_cost_weighted_histogram(lib):
    """What mpm_comm_rebalance_weighted hands to the cut routine: every rank's planes scaled by its cost per particle
    (quantised to 1/4096).  Two slabs of equal particle density, the right half three times as expensive per particle:
    the cut that equalises the summed cost sits where the cheap side holds 3/4 of the cost-free count."""
    import mpm_b200
    hist = np.zeros(128, np.int64)
    hist[4:124] = 100_000
    wq = np.where(np.arange(128) < 64, 4096, 3 * 4096).astype(np.int64)   # planes owned by rank 0 / rank 1 before the re-cut
    cuts = mpm_b200.slab_cuts(hist * wq, 2)
    # total cost = 60 planes x 1 + 60 planes x 3 = 240 plane-units; half = 120 = 60 cheap planes + 20 expensive ones
    assert abs(cuts[1] - 84) <= 1, cuts
    assert mpm_b200.slab_cuts(hist * 4096, 2)[1] == mpm_b200.slab_cuts(hist, 2)[1] == 64   # equal costs == equal counts


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    import mpm_b200
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank holds a different part of the scene; the cuts must come out identical everywhere
        rng = np.random.default_rng(100 + rank)
        x = rng.uniform(4, 100 if rank == 0 else 60, 50_000).astype(np.float32)
        local = np.bincount(x.astype(np.int32), minlength=128).astype(np.int64)
        t = torch.from_numpy(local.copy())
        dist.all_reduce(t)
        cuts = mpm_b200.slab_cuts(t.numpy(), world)
        gathered = [None] * world
        dist.all_gather_object(gathered, cuts)
        # what the NCCL bootstrap does with the 128-byte id: rank 0 makes it, everyone gets the same bytes
        uid = [bytes(range(128)) if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        q.put((rank, cuts, gathered, int(t.numpy().sum()), uid[0]))
    finally:
        dist.destroy_process_group()


def test_slab_cuts_agree_across_two_gloo_ranks(lib):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (_, c0, g0, tot0, u0), (_, c1, g1, tot1, u1) = out
    assert c0 == c1 and g0 == [c0, c0] and g1 == g0 and tot0 == tot1 == 100_000
    assert c0[0] == 0 and c0[2] == 128 and 4 <= c0[1] <= 124
    assert u0 == u1 == bytes(range(128))


# ---------------------------------------------------------------- GPU: k slabs == 1 slab, bit for bit
def _run_ranks(op, world, pos, vel, Cm, mass, steps, **over):
    """k solvers on device 0, one thread each (the LOCAL transport blocks until its neighbours arrive)."""
    import mpm_b200
    hub = mpm_b200.LocalHub(world)
    out, errs = [None] * world, []

    def work(r):
        try:
            with mpm_b200.Solver(helpers.mpm_params_from_orc(op, **over), pos.shape[0]) as s:
                s.comm_init_local(hub, r, world)
                s.upload(pos, vel, Cm, mass)          # the global set on every rank; each keeps its slab
                s.step(steps)
                gp, gv, gc, gm = s.download()
                out[r] = dict(pos=gp, vel=gv, C=gc, mass=gm, ids=s.download_ids(), grid=s.download_grid(), slab=s.slab(),
                              stats=s.stats())
        except Exception as e:  # noqa: BLE001
            errs.append((r, e))

    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(300)
    hub.close()
    assert not errs, errs
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("path", [1, 2], ids=["reference_path", "tiled_path"])
@pytest.mark.parametrize("world", [2, 3])
def test_k_slabs_bit_identical_to_one(lib, world, path):
    op = orc.variant("3d_gpu", (48, 32, 32))
    op.interaction = 0
    n = 60000
    pos, vel, Cm, mass = helpers.random_cloud(op, n, seed=7, vel_sigma=1.5)   # fast particles: real migration traffic
    steps = 8
    ref = orc.State(op, pos, vel, Cm, mass)
    ref.step(steps)
    ranks = _run_ranks(op, world, pos, vel, Cm, mass, steps, kernel_path=path)
    ids = np.concatenate([r["ids"] for r in ranks])
    assert np.array_equal(np.sort(ids), np.arange(n, dtype=np.uint32)), "particles lost or duplicated by migration"
    for what in ("pos", "vel", "C", "mass"):
        full = np.zeros_like(getattr(ref, what))
        for r in ranks:
            full[r["ids"]] = r[what]
        helpers.assert_bit_equal(full, getattr(ref, what), f"{what} ({world} slabs)")
    # each rank's particles sit in its own slab, and its owned grid planes equal the single-domain grid
    Ry, Rz = 32, 32
    gref = ref.grid.reshape(48, Ry, Rz, 4)
    cover = []
    for r in ranks:
        x0, x1, gx0, nxl = r["slab"]
        cover.append((x0, x1))
        cx = r["pos"][:, 0].astype(np.int32)
        assert ((cx >= x0) & (cx < x1)).all()
        g = r["grid"].reshape(nxl, Ry, Rz, 4)
        helpers.assert_bit_equal(g[x0 - gx0: x1 - gx0], gref[x0:x1], f"owned planes of rank {r['stats'].rank}")
    assert cover[0][0] == 0 and cover[-1][1] == 48 and all(a[1] == b[0] for a, b in zip(cover, cover[1:]))
    assert sum(r["stats"].local_particles for r in ranks) == n


@pytest.mark.gpu
def test_migration_burst_takes_the_overflow_round(lib):
    """A migration message is sized from the previous step's count over the same slab boundary (4096 records at
    first): 12000 particles crossing in one step must go through the overflow round and still be bit-exact."""
    op = orc.variant("3d_gpu", (48, 32, 32))
    op.interaction = 0
    rng = np.random.default_rng(5)
    n_bg, n_jet = 30000, 24000
    pos, vel, Cm, mass = helpers.random_cloud(op, n_bg + n_jet, seed=9, vel_sigma=0.2)
    # a dense sheet that straddles the middle of the domain and moves in +x at 0.9 cells per step (dt = 0.2)
    pos[n_bg:, 0] = rng.uniform(23.2, 24.8, n_jet).astype(np.float32)
    vel[n_bg:] = (4.5, 0.0, 0.0)
    Cm[n_bg:] = 0.0
    mass[n_bg:] = 0.05
    steps = 3
    ref = orc.State(op, pos, vel, Cm, mass)
    ref.step(steps)
    ranks = _run_ranks(op, 2, pos, vel, Cm, mass, steps, kernel_path=2)
    moved = (ref.pos[:, 0].astype(np.int32) >= ranks[0]["slab"][1]) & (pos[:, 0].astype(np.int32) < ranks[0]["slab"][1])
    assert moved.sum() > 8192, "the scene does not produce a burst"
    for what in ("pos", "vel", "C"):
        full = np.zeros_like(getattr(ref, what))
        for r in ranks:
            full[r["ids"]] = r[what]
        helpers.assert_bit_equal(full, getattr(ref, what), f"{what} after a migration burst")


@pytest.mark.gpu
@pytest.mark.parametrize("path,math", [(2, 0), (3, 1)], ids=["tiled_strict", "cell_fast"])
def test_particle_jumping_more_than_one_slab_is_held_back(lib, path, math):
    """|v| dt larger than a whole neighbouring slab (numerical outliers of violent scenes): the particle is parked in
    the neighbour's far plane for that step and counted; nothing is lost and nothing fails."""
    op = orc.variant("3d_gpu", (64, 32, 32))
    op.interaction = 0
    n = 40000
    pos, vel, Cm, mass = helpers.random_cloud(op, n, seed=13, vel_sigma=0.2)
    pos[:, 0] = 20.0 + (pos[:, 0] - 3.5) * (24.0 / 57.0)   # the bulk sits in x in [20, 44]: the slab cuts fall inside it
    fast = np.arange(0, 40)   # 40 isolated outliers in the empty regions: 200 cells per unit time = 40 cells per step
    k = np.arange(20)
    for grp, x, vx in ((fast[:20], 6.5, 200.0), (fast[20:], 57.5, -200.0)):
        pos[grp, 0], pos[grp, 1], pos[grp, 2] = x, 5.5 + 5.0 * (k % 5), 5.5 + 5.0 * (k // 5)
        vel[grp] = (vx, 0.0, 0.0)
    Cm[fast] = 0.0
    mass[fast] = 1e-3
    ranks = _run_ranks(op, 4, pos, vel, Cm, mass, 2, kernel_path=path, math_mode=math)
    ids = np.concatenate([r["ids"] for r in ranks])
    assert np.array_equal(np.sort(ids), np.arange(n, dtype=np.uint32)), "particles lost or duplicated"
    assert sum(r["stats"].slab_jump_clamps for r in ranks) >= 20
    for r in ranks:
        x0, x1, _, _ = r["slab"]
        cx = r["pos"][:, 0].astype(np.int32)
        assert ((cx >= x0) & (cx < x1)).all(), "a particle sits outside its rank's slab"


@pytest.mark.gpu
def test_rebalance_moves_the_cuts_and_keeps_the_bits(lib):
    """A jet leaves the first slab: mpm_comm_rebalance every 5 steps must shift the cuts after it, hand the particles to
    their new owners, and leave the physics untouched (bit-identical to the oracle on the strict path)."""
    import mpm_b200
    op = orc.variant("3d_gpu", (64, 32, 32))
    op.interaction = 0
    n = 50000
    pos, vel, Cm, mass = helpers.random_cloud(op, n, seed=17, vel_sigma=0.2)
    pos[:, 0] = 6.0 + (pos[:, 0] - 3.5) * (20.0 / 57.0)   # everything starts in x in [6, 26] ...
    vel[:, 0] += 2.5                                         # ... and drifts to +x, half a cell per step
    steps, world = 20, 3
    ref = orc.State(op, pos, vel, Cm, mass)
    ref.step(steps)
    hub = mpm_b200.LocalHub(world)
    out, errs = [None] * world, []

    def work(r):
        try:
            with mpm_b200.Solver(helpers.mpm_params_from_orc(op, kernel_path=2), n) as s:
                s.comm_init_local(hub, r, world)
                s.upload(pos, vel, Cm, mass)
                slab0 = None
                for k in range(steps // 5):
                    s.step(5)
                    if slab0 is None:
                        slab0 = s.slab()
                    s.comm_rebalance(3)
                gp, gv, gc, gm = s.download()
                out[r] = dict(pos=gp, vel=gv, C=gc, ids=s.download_ids(), slab0=slab0, slab=s.slab(), stats=s.stats())
        except Exception as e:  # noqa: BLE001
            errs.append((r, e))

    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [t.start() for t in th]
    [t.join(300) for t in th]
    hub.close()
    assert not errs, errs
    assert any(o["slab"][:2] != o["slab0"][:2] for o in out), "the cuts never moved"
    cover = [o["slab"][:2] for o in out]
    assert cover[0][0] == 0 and cover[-1][1] == 64 and all(a[1] == b[0] for a, b in zip(cover, cover[1:]))
    ids = np.concatenate([o["ids"] for o in out])
    assert np.array_equal(np.sort(ids), np.arange(n, dtype=np.uint32))
    for what in ("pos", "vel", "C"):
        full = np.zeros_like(getattr(ref, what))
        for o in out:
            full[o["ids"]] = o[what]
        helpers.assert_bit_equal(full, getattr(ref, what), f"{what} with re-cut slabs")
    for o in out:
        cx = o["pos"][:, 0].astype(np.int32)
        assert ((cx >= o["slab"][0]) & (cx < o["slab"][1])).all()
    counts = [o["stats"].local_particles for o in out]
    assert max(counts) - min(counts) < 0.35 * n / world, counts   # still roughly balanced although the fluid moved 10 cells


@pytest.mark.gpu
def test_weighted_rebalance_cuts_by_cost(lib):
    """mpm_comm_rebalance_weighted: a rank that reports three times the cost per particle ends up with about a third of
    the particles of the others (the cuts equalise count x cost), nothing is lost, and the physics is untouched
    (bit-identical to the oracle on the strict path).  Equal costs behave like mpm_comm_rebalance."""
    import mpm_b200
    op = orc.variant("3d_gpu", (96, 32, 32))
    op.interaction = 0
    n = 60000
    pos, vel, Cm, mass = helpers.random_cloud(op, n, seed=23, vel_sigma=0.1)   # uniform in x over [3.5, 92.5)
    steps, world = 12, 3
    ref = orc.State(op, pos, vel, Cm, mass)
    ref.step(steps)
    res = {}
    for tag, costs in (("weighted", (3.0, 1.0, 1.0)), ("equal", (0.7, 0.7, 0.7))):
        hub = mpm_b200.LocalHub(world)
        out, errs = [None] * world, []

        def work(r):
            try:
                with mpm_b200.Solver(helpers.mpm_params_from_orc(op, kernel_path=2), n) as s:
                    s.comm_init_local(hub, r, world)
                    s.upload(pos, vel, Cm, mass)
                    for k in range(steps // 2):
                        s.step(2)
                        s.comm_rebalance(3, cost_per_particle=costs[r])
                    gp, gv, gc, gm = s.download()
                    out[r] = dict(pos=gp, vel=gv, C=gc, ids=s.download_ids(), slab=s.slab(), stats=s.stats())
            except Exception as e:  # noqa: BLE001
                errs.append((r, e))

        th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
        [t.start() for t in th]
        [t.join(300) for t in th]
        hub.close()
        assert not errs, errs
        ids = np.concatenate([o["ids"] for o in out])
        assert np.array_equal(np.sort(ids), np.arange(n, dtype=np.uint32)), "particles lost or duplicated"
        for what in ("pos", "vel", "C"):
            full = np.zeros_like(getattr(ref, what))
            for o in out:
                full[o["ids"]] = o[what]
            helpers.assert_bit_equal(full, getattr(ref, what), f"{what} with cost-weighted cuts ({tag})")
        res[tag] = [o["stats"].local_particles for o in out]
    cw, ce = res["weighted"], res["equal"]
    # equal costs: equal counts (to the width of a plane: 60000 / 89 planes = 674 particles)
    assert max(ce) - min(ce) < 2 * 700, ce
    # cost 3 : 1 : 1 -> counts 1/7 : 3/7 : 3/7 of n, reached within the 6 re-cuts of <= 3 planes each (the first cut has to
    # move from plane 32 to plane ~16)
    assert abs(cw[0] - n / 7) < 0.05 * n and abs(cw[1] - 3 * n / 7) < 0.05 * n and abs(cw[2] - 3 * n / 7) < 0.05 * n, cw


@pytest.mark.gpu
def test_k_slabs_cell_path_within_fast_tolerance(lib):
    """The cell path (FAST math) on 3 slabs: float accumulation order differs from the 1-slab run, so the bar is the
    FAST tolerance against the strict oracle, plus exact particle bookkeeping."""
    op = orc.variant("3d_gpu", (48, 32, 32))
    op.interaction = 0
    n, steps = 60000, 4
    pos, vel, Cm, mass = helpers.random_cloud(op, n, seed=7, vel_sigma=1.5)
    ref = orc.State(op, pos, vel, Cm, mass)
    ref.step(steps)
    ranks = _run_ranks(op, 3, pos, vel, Cm, mass, steps, kernel_path=3, math_mode=1)
    ids = np.concatenate([r["ids"] for r in ranks])
    assert np.array_equal(np.sort(ids), np.arange(n, dtype=np.uint32))
    fp, fv = np.zeros_like(ref.pos), np.zeros_like(ref.vel)
    for r in ranks:
        fp[r["ids"]], fv[r["ids"]] = r["pos"], r["vel"]
    assert np.abs(fp - ref.pos).max() < 2e-4 and helpers.rel_err(fv, ref.vel) < 2e-4
    # the halo planes went by direct peer stores (k_halo_push / k_halo_wait_add), as between the processes of a multi-GPU run
    assert all(r["stats"].halo_peer_exchanges == 2 * steps for r in ranks), [r["stats"].halo_peer_exchanges for r in ranks]


@pytest.mark.gpu
def test_halo_transports_agree_bit_for_bit(lib, monkeypatch):
    """The same 3-slab run with the halo planes sent through the transport (MPM_NO_P2P=1: copies + add kernels) and by
    peer stores (flags, no copies): identical bits, and identical to the oracle (strict path)."""
    op = orc.variant("3d_gpu", (48, 32, 32))
    op.interaction = 0
    n, steps = 40000, 5
    pos, vel, Cm, mass = helpers.random_cloud(op, n, seed=23, vel_sigma=1.0)
    ref = orc.State(op, pos, vel, Cm, mass); ref.step(steps)
    res = {}
    for mode in ("peer", "copy"):
        if mode == "copy":
            monkeypatch.setenv("MPM_NO_P2P", "1")
        ranks = _run_ranks(op, 3, pos, vel, Cm, mass, steps, kernel_path=2)
        assert all((r["stats"].halo_peer_exchanges > 0) == (mode == "peer") for r in ranks)
        full = np.zeros_like(ref.pos)
        for r in ranks:
            full[r["ids"]] = r["pos"]
        res[mode] = full
        helpers.assert_bit_equal(full, ref.pos, f"pos, halos by {mode}")
    helpers.assert_bit_equal(res["peer"], res["copy"], "peer-store halos vs copied halos")


@pytest.mark.gpu
def test_per_rank_checkpoint_resumes_on_any_world_size(lib, tmp_path):
    """mpm_save_state under a communicator writes one file per rank (records + global indices); mpm_load_state finds them,
    restores the original particle order, and the run continues -- here on ONE solver -- bit-identically to the oracle."""
    import mpm_b200
    op = orc.variant("3d_gpu", (48, 32, 32))
    op.interaction = 0
    n = 30000
    pos, vel, Cm, mass = helpers.random_cloud(op, n, seed=29, vel_sigma=1.0)
    ref = orc.State(op, pos, vel, Cm, mass); ref.step(7)
    path = str(tmp_path / "state.mpm")
    world = 3
    hub = mpm_b200.LocalHub(world)
    errs = []

    def work(r):
        try:
            with mpm_b200.Solver(helpers.mpm_params_from_orc(op, kernel_path=2), n) as s:
                s.comm_init_local(hub, r, world)
                s.upload(pos, vel, Cm, mass)
                s.step(4)
                s.save_state(path)
        except Exception as e:  # noqa: BLE001
            errs.append((r, e))

    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [t.start() for t in th]
    [t.join(300) for t in th]
    hub.close()
    assert not errs, errs
    assert all(os.path.exists(f"{path}.rank{r}of{world}") for r in range(world)) and not os.path.exists(path)
    with mpm_b200.Solver(helpers.mpm_params_from_orc(op, kernel_path=2), n) as s:
        s.load_state(path)
        assert s.num_particles == n and s.stats().steps == 4
        s.step(3)
        gp, gv, gc, gm = s.download()
    helpers.assert_bit_equal(gp, ref.pos, "pos"); helpers.assert_bit_equal(gv, ref.vel, "vel"); helpers.assert_bit_equal(gm, ref.mass, "mass")


@pytest.mark.gpu
def test_slab_dam_break_trajectory_and_balance(lib):
    """Reduced dam-break on 2 slabs, 30 steps, lattice generated on the device on every rank (mpm_init_block)."""
    import mpm_b200
    op = orc.variant("3d_gpu", 32)
    op.interaction = 0
    lo, hi = (4, 4, 4), (20, 20, 20)
    pos = orc.init_block(3, lo, hi, 0.5)
    ref = orc.State(op, pos)
    ref.step(30)
    world = 2
    hub = mpm_b200.LocalHub(world)
    out, errs = [None] * world, []

    def work(r):
        try:
            with mpm_b200.Solver(helpers.mpm_params_from_orc(op, kernel_path=2), pos.shape[0]) as s:
                s.comm_init_local(hub, r, world)
                s.initialise_sim(lo, hi, 0.5)
                n0 = s.stats().local_particles
                s.step(30)
                out[r] = (s.download(), s.download_ids(), n0, s.slab())
        except Exception as e:  # noqa: BLE001
            errs.append((r, e))

    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [t.start() for t in th]
    [t.join(300) for t in th]
    hub.close()
    assert not errs, errs
    assert abs(out[0][2] - out[1][2]) <= 2 * 32 * 32 * 4, "initial slabs are not balanced"   # within two lattice planes
    full = np.zeros_like(ref.pos)
    fullv = np.zeros_like(ref.vel)
    for (gp, gv, gc, gm), ids, _, _ in out:
        full[ids], fullv[ids] = gp, gv
    helpers.assert_bit_equal(full, ref.pos, "pos after 30 steps on 2 slabs")
    helpers.assert_bit_equal(fullv, ref.vel, "vel after 30 steps on 2 slabs")


# ---------------------------------------------------------------- GPU x2: the NCCL transport, one process per GPU
def _nccl_worker(rank, world, uid, op_bytes, arrs, steps, q, path=2, math=0):
    import ctypes
    import mpm_b200
    op = orc.OrcParams.from_buffer_copy(op_bytes)
    pos, vel, Cm, mass = arrs
    try:
        with mpm_b200.Solver(helpers.mpm_params_from_orc(op, kernel_path=path, math_mode=math), pos.shape[0], device=rank) as s:
            s.comm_init(uid, rank, world)
            s.upload(pos, vel, Cm, mass)
            s.step(steps)
            gp, gv, gc, gm = s.download()
            q.put((rank, gp, gv, gc, s.download_ids(), None, s.stats().halo_peer_exchanges))
    except Exception as e:  # noqa: BLE001
        q.put((rank, None, None, None, None, repr(e), 0))


@pytest.mark.gpu
def test_nccl_two_gpus_bit_identical_to_oracle(lib):
    import mpm_b200
    if lib.mpm_device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    import torch.multiprocessing as mp
    op = orc.variant("3d_gpu", (48, 32, 32))
    op.interaction = 0
    n, steps = 60000, 8
    pos, vel, Cm, mass = helpers.random_cloud(op, n, seed=11, vel_sigma=1.5)
    ref = orc.State(op, pos, vel, Cm, mass)
    ref.step(steps)
    uid = mpm_b200.comm_unique_id()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, uid, bytes(op), (pos, vel, Cm, mass), steps, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=300) for _ in procs]
    [p.join(60) for p in procs]
    assert all(r[5] is None for r in res), [r[5] for r in res]
    fp, fv, fc = np.zeros_like(ref.pos), np.zeros_like(ref.vel), np.zeros_like(ref.C)
    for _, gp, gv, gc, ids, _, _ in res:
        fp[ids], fv[ids], fc[ids] = gp, gv, gc
    helpers.assert_bit_equal(fp, ref.pos, "pos (2 GPUs, NCCL)")
    helpers.assert_bit_equal(fv, ref.vel, "vel (2 GPUs, NCCL)")
    helpers.assert_bit_equal(fc, ref.C, "C (2 GPUs, NCCL)")


@pytest.mark.gpu
def test_two_gpus_cell_path_ipc_halos_within_fast_tolerance(lib):
    """What the scaling benchmark runs: one process per GPU, cell kernels (FAST), halo planes by peer stores over CUDA IPC,
    migration through NCCL.  Against the strict oracle within the FAST tolerance; exact particle bookkeeping."""
    import mpm_b200
    if lib.mpm_device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    import torch.multiprocessing as mp
    op = orc.variant("3d_gpu", (96, 64, 64))
    op.interaction = 0
    n, steps = 200000, 6
    pos, vel, Cm, mass = helpers.random_cloud(op, n, seed=31, vel_sigma=1.5)
    ref = orc.State(op, pos, vel, Cm, mass)
    ref.step(steps)
    uid = mpm_b200.comm_unique_id()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, uid, bytes(op), (pos, vel, Cm, mass), steps, q, 3, 1)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=300) for _ in procs]
    [p.join(60) for p in procs]
    assert all(r[5] is None for r in res), [r[5] for r in res]
    assert all(r[6] == 2 * steps for r in res), "the halos did not go by peer stores"
    ids = np.concatenate([r[4] for r in res])
    assert np.array_equal(np.sort(ids), np.arange(n, dtype=np.uint32))
    fp, fv = np.zeros_like(ref.pos), np.zeros_like(ref.vel)
    for _, gp, gv, gc, i, _, _ in res:
        fp[i], fv[i] = gp, gv
    assert np.abs(fp - ref.pos).max() < 2e-4 and helpers.rel_err(fv, ref.vel) < 2e-4
