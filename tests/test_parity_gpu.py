"""GPU parity tests (run on the B200 box with -m gpu).  Everything goes through the C ABI.

Bar: int32 fixed-point grid mode (the reference's deterministic model: MLSMPM3DFluidMultithreadNew.cs and the
GLSL shaders) -> grid words, particle floats, cell keys and the binning permutation are BIT-EXACT against the
oracle.  Float grid mode -> atomic order differs from the reference's serial order, so the tolerance is
calibrated from the oracle's own sensitivity to particle order and written in the test.
PARITY UNPINNED: the oracle is a restatement (no reference golden vectors exist)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import helpers
from oracle import orc

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import make_golden  # noqa: E402

pytestmark = pytest.mark.gpu

PHASES = ["clear_grid", "p2g1", "p2g2", "update_grid", "g2p"]


def make_solver(op, n, **over):
    import mpm_b200
    return mpm_b200.Solver(helpers.mpm_params_from_orc(op, **over), max(n, 1))


def sphere_into_cloud(op):
    if op.interaction in (1, 2):
        op.sphere_pos[:] = [op.grid[0] * 0.3, op.grid[1] * 0.5, op.grid[2] * 0.5]
        op.sphere_radius = 4.0


# ---------------------------------------------------------------- fixed-point mode: bit-exact
@pytest.mark.parametrize("path", [1, 2], ids=["reference_path", "tiled_path"])
@pytest.mark.parametrize("variant,grid", [("3d_fixed", 32), ("3d_gpu", (40, 32, 24)), ("3d_gpu", 96)])
def test_each_phase_bit_exact_fixed(lib, variant, grid, path):
    op = orc.variant(variant, grid)
    sphere_into_cloud(op)
    n = 20000
    pos, vel, Cm, mass = helpers.random_cloud(op, n, seed=21)
    ref = orc.State(op, pos, vel, Cm, mass)
    with make_solver(op, n, kernel_path=path) as s:
        s.upload(pos, vel, Cm, mass)
        for k, ph in enumerate(PHASES):
            getattr(ref, ph)()
            s.run_phase(k)
            helpers.assert_bit_equal(s.download_grid(), ref.grid, f"grid after {ph}")
        gp, gv, gc, gm = s.download()
        helpers.assert_bit_equal(gp, ref.pos, "pos"); helpers.assert_bit_equal(gv, ref.vel, "vel")
        helpers.assert_bit_equal(gc, ref.C, "C"); helpers.assert_bit_equal(gm, ref.mass, "mass")
        helpers.assert_bit_equal(s.positions(), ref.positions(), "positions (x,y,z,|v|)")
        assert s.stats().kernel_path == path


@pytest.mark.parametrize("sort_interval", [1, 4])
@pytest.mark.parametrize("path", [1, 2], ids=["reference_path", "tiled_path"])
def test_dam_break_trajectory_bit_exact_fixed(lib, path, sort_interval):
    """Reference GPU scene family (MLSMPM3DFluidMultithreadGPU.cs:654-707) at reduced size, 30 steps."""
    op = orc.variant("3d_gpu", 32)
    op.interaction = 0
    pos = orc.init_block(3, (4, 4, 4), (20, 20, 20), 0.5)  # 32^3 = 32768 particles against a corner
    ref = orc.State(op, pos)
    with make_solver(op, pos.shape[0], kernel_path=path, sort_interval=sort_interval) as s:
        assert s.initialise_sim((4, 4, 4), (20, 20, 20), 0.5) == pos.shape[0]
        for chunk in range(3):
            ref.step(10)
            s.step(10)
            gp, gv, gc, gm = s.download()
            helpers.assert_bit_equal(gp, ref.pos, f"pos after {10 * (chunk + 1)} steps")
            helpers.assert_bit_equal(gv, ref.vel, "vel"); helpers.assert_bit_equal(gc, ref.C, "C")
        helpers.assert_bit_equal(s.download_grid(), ref.grid, "grid")
        assert np.abs(gv).max() > 0.5  # the block really collapsed


@pytest.mark.parametrize("name", ["3d_fixed", "3d_gpu"])
@pytest.mark.parametrize("path", [1, 2], ids=["reference_path", "tiled_path"])
def test_golden_fixed(lib, name, path):
    op, n, steps, seed = make_golden.case_params(name)
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", f"{name}.npz"))
    with make_solver(op, n, kernel_path=path) as s:
        s.upload(z["pos0"], z["vel0"], z["C0"], z["mass"])
        s.step(int(z["steps"]))
        gp, gv, gc, _ = s.download()
        helpers.assert_bit_equal(s.download_grid(), z["grid"], "grid")
        helpers.assert_bit_equal(gp, z["pos"], "pos"); helpers.assert_bit_equal(gv, z["vel"], "vel")
        helpers.assert_bit_equal(gc, z["C"], "C")


# ---------------------------------------------------------------- MPM_MATH_FAST: stated tolerance
# FAST re-associates the arithmetic (FMA, hoisting, sum factorisation, reciprocal multiplies): every particle /
# grid quantity differs from the strict result by rounding only.  Tolerances (relative to the largest magnitude
# of the compared array, per step):  grid mass/momentum 2e-6, particle vel / C 2e-5 (C is a difference of O(1)
# terms), positions 1e-6 of the domain.  Measured on B200: profiles/r2/fast_math_errors.txt (the lines these tests print).
FAST_TOL = {"grid": 2e-6, "vel": 2e-5, "C": 2e-5, "pos": 1e-6}


@pytest.mark.parametrize("path", [2, 3], ids=["tiled_path", "cell_path"])
@pytest.mark.parametrize("variant,grid", [("3d_fixed", 32), ("3d_gpu", (40, 32, 24)), ("3d_gpu", 96)])
def test_fast_math_each_phase_within_tolerance(lib, variant, grid, path):
    op = orc.variant(variant, grid)
    sphere_into_cloud(op)
    n = 20000
    pos, vel, Cm, mass = helpers.random_cloud(op, n, seed=21)
    ref = orc.State(op, pos, vel, Cm, mass)
    errs = {}
    with make_solver(op, n, kernel_path=path, math_mode=1) as s:
        s.upload(pos, vel, Cm, mass)
        for k, ph in enumerate(PHASES):
            getattr(ref, ph)()
            s.run_phase(k)
            g = s.download_grid().astype(np.float64) / 1e7
            errs[f"grid after {ph}"] = helpers.rel_err(g, ref.grid.astype(np.float64) / 1e7)
            # keep the two in lock-step so each phase is judged on identical inputs
            if ph in ("p2g1", "p2g2"):
                assert errs[f"grid after {ph}"] <= FAST_TOL["grid"], (ph, errs)
            # after update_grid the cells hold v = momentum / mass: a one-unit (1e-7) difference in the momentum of a
            # node with almost no mass is a large relative change THERE (that is all the plain max-norm above sees: up to
            # 4e-2 on the cell path, which truncates once per (cell, node)), but such a node carries no weight in G2P.
            # So the velocity grid is judged by what a node's velocity stands for -- |v - v_ref| * mass against the
            # largest momentum -- at twice the P2G tolerance (the division adds one rounding).
            if ph == "update_grid":
                gr = ref.grid.astype(np.float64) / 1e7
                errs["update_grid, mass-weighted"] = float(np.abs((g[:, :3] - gr[:, :3]) * gr[:, 3:4]).max() / np.abs(gr[:, :3] * gr[:, 3:4]).max())
                assert errs["update_grid, mass-weighted"] <= 2 * FAST_TOL["grid"], (ph, errs)
                assert helpers.rel_err(g[:, 3], gr[:, 3]) <= FAST_TOL["grid"], (ph, errs)
        gp, gv, gc, gm = s.download()
        assert s.stats().kernel_path == path
    errs["pos"] = float(np.abs(gp.astype(np.float64) - ref.pos).max() / max(op.grid))
    errs["vel"], errs["C"] = helpers.rel_err(gv, ref.vel), helpers.rel_err(gc, ref.C)
    print("FAST_MATH_ERRORS", variant, grid, path, {k: f"{v:.3g}" for k, v in errs.items()})
    assert errs["pos"] <= FAST_TOL["pos"] and errs["vel"] <= FAST_TOL["vel"] and errs["C"] <= FAST_TOL["C"], errs
    helpers.assert_bit_equal(gm, ref.mass, "mass")


def test_fast_tiled_kernels_with_a_stale_binning(lib):
    """MPM_MATH_FAST on the tiled path with sort_interval = 4: between two bin phases particles drift out of their block
    and take the kernels' slow path (global atomics / direct gathers).  Same per-particle bar as a fresh binning every step."""
    op = orc.variant("3d_gpu", 32)
    op.interaction = 0
    pos = orc.init_block(3, (4, 4, 4), (20, 20, 20), 0.5)
    ref = orc.State(op, pos); ref.step(12)
    out = {}
    for si in (1, 4):
        with make_solver(op, pos.shape[0], kernel_path=2, math_mode=1, sort_interval=si) as s:
            s.initialise_sim((4, 4, 4), (20, 20, 20), 0.5)
            s.step(12)
            out[si] = s.download()
    for si in (1, 4):
        gp, gv = out[si][0], out[si][1]
        assert np.abs(gp.astype(np.float64) - ref.pos).max() / 32 <= 12 * FAST_TOL["pos"], si
        assert helpers.rel_err(gv, ref.vel) <= 12 * FAST_TOL["vel"], si
    # the slow path computes the same sums in another order: the two runs agree to rounding, not to the bit
    assert np.abs(out[1][0] - out[4][0]).max() < 1e-4


@pytest.mark.parametrize("path", [2, 3], ids=["tiled_path", "cell_path"])
def test_fast_math_dam_break_drift(lib, path):
    """30 steps of the reduced dam-break in FAST mode vs the strict oracle: aggregates.
    Individual trajectories decorrelate (chaotic), so bounds are on centre of mass, kinetic energy and bounds."""
    op = orc.variant("3d_gpu", 32)
    op.interaction = 0
    pos = orc.init_block(3, (4, 4, 4), (20, 20, 20), 0.5)
    ref = orc.State(op, pos); ref.step(30)
    with make_solver(op, pos.shape[0], kernel_path=path, math_mode=1) as s:
        s.initialise_sim((4, 4, 4), (20, 20, 20), 0.5)
        s.step(30)
        gp, gv, _, _ = s.download()
    com = np.abs(gp.mean(0) - ref.pos.mean(0)).max()
    ke, ke_ref = (gv.astype(np.float64) ** 2).sum(), (ref.vel.astype(np.float64) ** 2).sum()
    med = float(np.median(np.abs(gp - ref.pos).max(1)))
    print("FAST_MATH_DRIFT", path, dict(com=float(com), ke_rel=float(abs(ke - ke_ref) / ke_ref), median_particle_dev=med))
    assert com < 1e-3 and abs(ke - ke_ref) / ke_ref < 1e-3 and med < 1e-3
    assert gp.min() >= 2.0 and gp.max() <= 30.0


# ---------------------------------------------------------------- binning: bit-exact
def block_key(op, pos, B):
    """Restatement of the bin key: id of the BxBxB block holding the base cell, (bx*NBy + by)*NBz + bz
    (the reference's index formula, MLSMPM3DFluidMultithread.cs:282, on block coordinates)."""
    c = pos.astype(np.int32)  # (Vector3I)p.pos: truncation (MLSMPM3DFluidMultithread.cs:259)
    nby, nbz = -(-op.grid[1] // B), -(-op.grid[2] // B)
    b = c // B
    return ((b[:, 0] * nby + b[:, 1]) * nbz + b[:, 2]).astype(np.uint32)


@pytest.mark.parametrize("grid,B", [(32, 4), (96, 8), ((128, 96, 96), 8)])
def test_cell_keys_and_sort_permutation_bit_exact(lib, grid, B):
    op = orc.variant("3d_gpu", grid)
    op.interaction = 0
    n = 300000
    pos, vel, Cm, mass = helpers.random_cloud(op, n, seed=33)
    pos[: n // 2] = pos[: n // 2] * 0.25 + 4.0  # half the particles crowd one corner: many equal keys
    with make_solver(op, n, kernel_path=2) as s:
        s.upload(pos, vel, Cm, mass)
        s.run_phase(5)
        keys, perm = s.last_sort()
        assert np.array_equal(keys, block_key(op, pos, B)), "cell keys"
        expect = orc.stable_sort_perm(keys)  # == std::stable_sort order
        assert np.array_equal(perm.astype(np.int32), expect), "binning permutation is not the stable sort"
        assert np.array_equal(perm.astype(np.int64), np.argsort(keys, kind="stable"))
        # binning must be invisible from outside: downloads stay in original index order
        gp, gv, gc, gm = s.download()
        helpers.assert_bit_equal(gp, pos, "pos order"); helpers.assert_bit_equal(gm, mass, "mass order")


def test_sort_edge_cases(lib):
    op = orc.variant("3d_gpu", 32)
    op.interaction = 0
    with make_solver(op, 5000, kernel_path=2) as s:
        s.step(2)  # empty particle set: nothing to do, nothing to crash
        assert s.num_particles == 0 and not s.download_grid().any()
        one = np.array([[10.5, 10.5, 10.5]], np.float32)
        s.upload(one)
        ref = orc.State(op, one)
        s.step(3); ref.step(3)
        helpers.assert_bit_equal(s.download()[0], ref.pos, "single particle")
        # all particles in one cell (maximum key collision), ragged count (not a multiple of any tile)
        rng = np.random.default_rng(2)
        same = (np.array([[12.0, 13.0, 14.0]], np.float32) + rng.uniform(0.01, 0.99, (4097, 3)).astype(np.float32))
        light = np.full(4097, 0.01, np.float32)  # keeps the int32 x 1e7 accumulators (|v| < 214.7) in range
        s.params.rest_density = 40.0; s.update_push_constants(); op.rest_density = 40.0
        s.upload(same, mass=light)
        ref = orc.State(op, same, mass=light)
        s.step(2); ref.step(2)
        helpers.assert_bit_equal(s.download()[0], ref.pos, "one-cell pile-up")
        helpers.assert_bit_equal(s.download_grid(), ref.grid, "one-cell pile-up grid")


# ---------------------------------------------------------------- float grid mode: calibrated tolerance
def shuffle_sensitivity(op, pos, vel, Cm, mass, steps):
    """max relative deviation of the oracle's own result between natural and shuffled particle order."""
    a = orc.State(op, pos, vel, Cm, mass); a.step(steps)
    perm = np.random.default_rng(0).permutation(pos.shape[0])
    b = orc.State(op, pos[perm], vel[perm], Cm[perm], mass[perm]); b.step(steps)
    inv = np.argsort(perm)
    return max(helpers.rel_err(b.pos[inv], a.pos), helpers.rel_err(b.vel[inv], a.vel), helpers.rel_err(b.C[inv], a.C),
               helpers.rel_err(b.grid_f(), a.grid_f())), a


@pytest.mark.parametrize("name", ["2d_st", "2d_mt", "3d_float"])
def test_float_grid_within_calibrated_tolerance(lib, name):
    op, n, steps, seed = make_golden.case_params(name)
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", f"{name}.npz"))
    sens, ref = shuffle_sensitivity(op, z["pos0"], z["vel0"], z["C0"], z["mass"], int(z["steps"]))
    tol = max(4.0 * sens, 2e-6)  # kappa = 4; floor = a few fp32 ulps
    assert tol < 1e-3, f"order sensitivity unexpectedly large: {sens}"
    with make_solver(op, n) as s:
        s.upload(z["pos0"], z["vel0"], z["C0"], z["mass"])
        s.step(int(z["steps"]))
        gp, gv, gc, _ = s.download()
        g = s.download_grid().view(np.float32)
    for what, a, b in (("pos", gp, ref.pos), ("vel", gv, ref.vel), ("C", gc, ref.C), ("grid", g, ref.grid_f())):
        e = helpers.rel_err(a, b)
        assert e <= tol, f"{name} {what}: rel err {e:.3g} > tol {tol:.3g} (oracle order sensitivity {sens:.3g})"
    # and against the committed golden vectors (made by the NumPy restatement)
    assert helpers.rel_err(gp, z["pos"]) <= tol and helpers.rel_err(gv, z["vel"]) <= tol


def test_2d_dam_break_config1_drift(lib):
    """BASELINE config 1 (2D dam-break 128^2, 16384 particles, reference 2D params), 100 steps: trajectories
    decorrelate chaotically in float mode, so compare aggregates."""
    op = orc.variant("2d_st", (128, 128, 1))
    pos = orc.init_block(2, (4, 4), (68, 68), 0.5)
    assert pos.shape[0] == 16384
    ref = orc.State(op, pos); ref.step(100)
    with make_solver(op, pos.shape[0]) as s:
        s.initialise_sim((4, 4), (68, 68), 0.5)
        s.step(100)
        gp, gv, _, _ = s.download()
    com_err = np.abs(gp.mean(0) - ref.pos.mean(0)).max()
    ke, ke_ref = 0.5 * (gv.astype(np.float64) ** 2).sum(), 0.5 * (ref.vel.astype(np.float64) ** 2).sum()
    assert com_err < 0.05, f"centre of mass drift {com_err}"          # cells, after 100 steps
    assert abs(ke - ke_ref) / ke_ref < 0.05, (ke, ke_ref)              # kinetic energy within 5 %
    assert gp[:, :2].min() >= 1.0 and gp[:, :2].max() <= 126.0         # clamp bounds hold


# ---------------------------------------------------------------- buffers in the reference's layouts
def test_aos80_round_trip_and_position_texture(lib):
    import mpm_b200
    op = orc.variant("3d_gpu", 32)
    n = 12345
    pos, vel, Cm, mass = helpers.random_cloud(op, n, seed=44)
    rec = np.zeros(n, mpm_b200.PARTICLE80)
    rec["pos"], rec["vel"], rec["mass"] = pos, vel, mass
    rec["C_x"], rec["C_y"], rec["C_z"] = Cm[:, 0:3], Cm[:, 3:6], Cm[:, 6:9]
    with make_solver(op, n) as s:
        s.upload_aos80(rec)
        back = s.download_aos80()
        for f in ("pos", "vel", "mass", "C_x", "C_y", "C_z"):
            helpers.assert_bit_equal(back[f], rec[f], f)
        p4 = s.positions()
        helpers.assert_bit_equal(p4[:, :3], pos, "positions before any step")
        dev_ptr, width = s.positions_device()
        assert dev_ptr and width == int(np.sqrt(np.float32(n))) + 1  # MLSMPM3DFluidMultithreadGPU.cs:196


def test_params_setters(lib):
    op = orc.variant("3d_gpu", 32)
    pos, vel, Cm, mass = helpers.random_cloud(op, 3000, seed=5)
    with make_solver(op, 3000) as s:
        s.upload(pos, vel, Cm, mass)
        s.params.dt = 0.9  # Dt setter clamps to [0, 0.4] (MLSMPM3DFluidMultithreadGPU.cs:64)
        s.params.gravity = -0.5  # the UI's startup value (main_ui.tscn:162)
        s.update_push_constants()
        op.dt, op.gravity = 0.4, -0.5
        s.set_sphere((9.0, 9.0, 9.0)); op.sphere_pos[:] = [9.0, 9.0, 9.0]
        ref = orc.State(op, pos, vel, Cm, mass)
        s.step(2); ref.step(2)
        helpers.assert_bit_equal(s.download()[1], ref.vel, "vel after parameter change")


# ---------------------------------------------------------------- full-size, size-independent properties
def test_full_size_block_drop_properties(lib):
    """BASELINE config 3 (128^3, 4 096 000 particles): too big for the oracle in seconds, so check
    (1) tiled path == reference-shaped path bit for bit (two independent kernel sets),
    (2) grid mass == sum of the particles' encoded weights within the truncation bound (partition of unity),
    (3) download order is the upload order."""
    op = orc.variant("3d_gpu", 128)
    op.interaction = 0
    lo, hi = (24, 24, 24), (104, 104, 104)
    res = {}
    for path in (1, 2):
        with make_solver(op, 4096000, kernel_path=path) as s:
            assert s.initialise_sim(lo, hi, 0.5) == 4096000
            s.run_phase(0); s.run_phase(1)
            g1 = s.download_grid()
            total = g1[:, 3].astype(np.int64).sum()
            n = 4096000
            assert 0 <= n * 10_000_000 - total <= 27 * n, "mass not conserved within the truncation bound"
            s.run_phase(2); s.run_phase(3); s.run_phase(4)
            s.step(2)
            res[path] = (s.download_grid(), s.download())
    helpers.assert_bit_equal(res[1][0], res[2][0], "grid: reference-shaped vs tiled path")
    for k, what in enumerate(("pos", "vel", "C", "mass")):
        helpers.assert_bit_equal(res[1][1][k], res[2][1][k], what)
    xm = res[2][1][0][:, 0].reshape(160, -1).mean(1)  # lattice order: index = (ix*160 + iy)*160 + iz
    assert np.all(np.diff(xm) > 0.25), "original (lattice) order lost"


# ---------------------------------------------------------------- cell path (counting-sort binning, one thread per cell)
def test_cell_path_auto_selection_and_edge_cases(lib):
    import mpm_b200
    op = orc.variant("3d_gpu", 32)
    op.interaction = 0
    with make_solver(op, 5000, math_mode=1) as s:           # AUTO + FAST -> cell path
        assert s.stats().kernel_path == mpm_b200.PATH_CELL
        s.step(2)                                            # empty particle set
        assert s.num_particles == 0 and not s.download_grid().any()
        one = np.array([[10.5, 10.5, 10.5]], np.float32)
        s.upload(one)
        ref = orc.State(op, one)
        s.step(3); ref.step(3)
        assert np.abs(s.download()[0] - ref.pos).max() < 1e-5, "single particle"
        # every particle in one cell (one thread walks 4097 particles), ragged count
        rng = np.random.default_rng(2)
        same = (np.array([[12.0, 13.0, 14.0]], np.float32) + rng.uniform(0.01, 0.99, (4097, 3)).astype(np.float32))
        light = np.full(4097, 0.01, np.float32)
        s.params.rest_density = 40.0; s.update_push_constants(); op.rest_density = 40.0
        s.upload(same, mass=light)
        ref = orc.State(op, same, mass=light)
        s.step(1); ref.step(1)
        gp, gv, _, _ = s.download()
        # 4097 fp32 accumulations per node before the single truncation: rounding grows with the count (~1e-7 * sqrt(n)
        # relative on the node sums) and the gamma = 7 EOS amplifies density errors; the bar here is "same physics"
        e_pos, e_vel = float(np.abs(gp - ref.pos).max()), helpers.rel_err(gv, ref.vel)
        print("CELL_PILE_UP_ERRORS", e_pos, e_vel)
        assert e_pos < 1e-3 and e_vel < 5e-3, "one-cell pile-up"
    with pytest.raises(mpm_b200.MpmError):                   # the cell path has no strict arithmetic
        make_solver(op, 100, kernel_path=3, math_mode=0)


def test_cell_binning_is_a_valid_cell_sort(lib):
    """Counting-sort binning: after a bin phase every particle is still there exactly once (download order is
    unchanged) and a following step matches the strict oracle within the FAST tolerance on a crowded, ragged cloud."""
    op = orc.variant("3d_gpu", (128, 96, 96))
    op.interaction = 0
    n = 300001
    pos, vel, Cm, mass = helpers.random_cloud(op, n, seed=33)
    pos[: n // 2] = pos[: n // 2] * 0.25 + 4.0   # half the particles crowd one corner (many particles per cell)
    mass[: n // 2] *= 0.05                        # keep the fixed-point accumulators in range there
    with make_solver(op, n, kernel_path=3, math_mode=1) as s:
        s.upload(pos, vel, Cm, mass)
        s.run_phase(5)
        gp, gv, gc, gm = s.download()
        helpers.assert_bit_equal(gp, pos, "pos order"); helpers.assert_bit_equal(gm, mass, "mass order")
        helpers.assert_bit_equal(gc, Cm, "C order")
        ref = orc.State(op, pos, vel, Cm, mass)
        ref.clear_grid(); ref.p2g1()
        s.run_phase(0); s.run_phase(1)
        g = s.download_grid().astype(np.float64)
        # The oracle truncates every particle's addend toward zero; the cell path truncates once per (cell, node).  So a
        # node may differ by up to one fixed-point unit per particle that touches it (particles in the 3x3x3 cells
        # around it), plus fp32 rounding of the accumulated value.
        R = op.grid
        cells = np.bincount(((pos[:, 0].astype(np.int64) * R[1] + pos[:, 1].astype(np.int64)) * R[2] + pos[:, 2].astype(np.int64)),
                            minlength=R[0] * R[1] * R[2]).reshape(R[0], R[1], R[2]).astype(np.float64)
        touch = sum(np.roll(cells, (dx, dy, dz), (0, 1, 2)) for dx in (-1, 0, 1) for dy in (-1, 0, 1) for dz in (-1, 0, 1)).reshape(-1, 1)
        gr = ref.grid.astype(np.float64)
        assert (np.abs(g - gr) <= touch + 2.0 + 4e-7 * np.abs(gr)).all(), float(np.abs(g - gr).max())
        # mass is conserved up to one truncation per (cell, node)
        total, want = g[:, 3].sum(), (mass.astype(np.float64) * 1e7).sum()
        assert abs(total - want) <= 27.0 * n + 1e-7 * want


def test_cell_path_phases_with_downloads_in_between(lib):
    """On the cell path the particle state lives in 64-byte records that the P2G kernels read through the binning's
    index, and P2G_1 leaves position / mass planes + slot-order ids for G2P.  Running the phases one by one with
    downloads in between (each download converts records -> planes and invalidates the binning, so every particle phase
    re-bins; G2P then runs without a P2G_1 since its binning and must rebuild its inputs) has to give the same step as
    mpm_step, up to the order of the fp32 per-cell sums, and must keep the download order."""
    op = orc.variant("3d_gpu", (64, 64, 64))
    op.interaction = 0
    n = 150000
    pos, vel, Cm, mass = helpers.random_cloud(op, n, seed=91)
    res = []
    for piecewise in (False, True):
        with make_solver(op, n, kernel_path=3, math_mode=1) as s:
            s.upload(pos, vel, Cm, mass)
            s.step(1)                                  # records become the live state
            if not piecewise:
                s.step(1)
            else:
                s.run_phase(5); s.run_phase(0); s.run_phase(1)
                mid = s.download()                      # records -> planes; the binning is void again
                s.run_phase(2)
                s.download_grid()
                s.run_phase(3)
                mid2 = s.download()
                for a, b in zip(mid, mid2):
                    helpers.assert_bit_equal(a, b, "particles must not change before G2P")
                s.run_phase(4)                          # fresh binning, no P2G_1 since: G2P rebuilds its inputs
            res.append(s.download())
            helpers.assert_bit_equal(res[-1][3], mass, "mass / download order")
    for what, a, b in zip(("pos", "vel", "C"), res[0], res[1]):
        e = helpers.rel_err(a, b)
        assert e < 1e-4, f"{what}: piecewise vs mpm_step rel err {e:.3g}"  # (rounding order of the per-cell fp32 sums)


def test_cell_path_box_sweeps_leave_nothing_behind(lib):
    """The cell path clears and updates only the bounding box of the occupied grid blocks (+ apron).  A block of fluid
    thrown across the grid makes that box move and shrink every step: after each step the grid must be exactly zero
    wherever the strict path (dense sweeps) has nothing within two cells, and centre of mass / kinetic energy must agree
    at the end."""
    op = orc.variant("3d_gpu", (96, 64, 64))
    op.interaction = 0
    pos = orc.init_block(3, (6, 30, 6), (26, 50, 26), 0.5)
    vel = np.zeros_like(pos); vel[:, 0] = 1.5; vel[:, 2] = 0.8        # cells per unit time, dt = 0.2: a block per ~25 steps
    n = pos.shape[0]
    with make_solver(op, n, kernel_path=3, math_mode=1) as a, make_solver(op, n, kernel_path=2, math_mode=0) as b:
        a.upload(pos, vel); b.upload(pos, vel)
        for step in range(30):
            a.step(1); b.step(1)
            if step % 3 == 2 or step < 3:
                ga = a.download_grid().reshape(96, 64, 64, 4)
                gb = b.download_grid().reshape(96, 64, 64, 4)
                near = (gb != 0).any(-1)
                for ax in range(3):
                    near = near | np.roll(near, 1, ax) | np.roll(near, -1, ax)
                    near = near | np.roll(near, 1, ax) | np.roll(near, -1, ax)
                assert not ga[~near].any(), f"step {step}: stale grid cells outside the fluid's support"
                assert (ga[..., 3] > 0).sum() > 0.9 * (gb[..., 3] > 0).sum()
        pa, va = a.download()[:2]; pb, vb = b.download()[:2]
    # 30 steps of a splashing block decorrelate single particles between FAST and STRICT arithmetic: compare aggregates
    assert np.abs(pa.mean(0) - pb.mean(0)).max() < 2e-3, np.abs(pa.mean(0) - pb.mean(0)).max()
    ka, kb = (va.astype(np.float64) ** 2).sum(), (vb.astype(np.float64) ** 2).sum()
    assert abs(ka - kb) / kb < 2e-3, (ka, kb)


def test_cell_path_full_size_block_drop(lib):
    """BASELINE config 3 (128^3, 4 096 000 particles) on the cell path vs the strict reference-shaped path after 3
    steps: FAST tolerance per particle (no chaos yet), exact particle count and order."""
    op = orc.variant("3d_gpu", 128)
    op.interaction = 0
    lo, hi = (24, 24, 24), (104, 104, 104)
    res = {}
    for path, math in ((1, 0), (3, 1)):
        with make_solver(op, 4096000, kernel_path=path, math_mode=math) as s:
            assert s.initialise_sim(lo, hi, 0.5) == 4096000
            s.step(3)
            res[path] = s.download()
    gp, gv, gc, gm = res[3]
    rp, rv, rc_, rm = res[1]
    assert np.abs(gp.astype(np.float64) - rp).max() / 128 <= 3 * FAST_TOL["pos"]
    assert helpers.rel_err(gv, rv) <= 5 * FAST_TOL["vel"] and helpers.rel_err(gc, rc_) <= 5 * FAST_TOL["C"]  # 3 steps compound
    helpers.assert_bit_equal(gm, rm, "mass")


def test_pipelined_position_hand_off_matches_the_synchronous_one(lib):
    """mpm_get_positions_async (double-buffered, copy stream) must deliver exactly what mpm_get_positions delivers,
    for consecutive steps, on both binned paths."""
    import ctypes as C
    import mpm_b200
    op = orc.variant("3d_gpu", 32)
    op.interaction = 0
    pos = orc.init_block(3, (4, 4, 4), (20, 20, 20), 0.5)
    n = pos.shape[0]
    for path, math in ((2, 0), (3, 1)):
        bufs = [mpm_b200.host_alloc(16 * n) for _ in range(2)]
        views = [np.ctypeslib.as_array((C.c_float * (4 * n)).from_address(b)).reshape(n, 4) for b in bufs]
        with make_solver(op, n, kernel_path=path, math_mode=math) as a, make_solver(op, n, kernel_path=path, math_mode=math) as b:
            a.initialise_sim((4, 4, 4), (20, 20, 20), 0.5)
            b.initialise_sim((4, 4, 4), (20, 20, 20), 0.5)
            got = []
            for k in range(4):
                a.step(1)
                a.positions_into_async(bufs[k & 1], n)
                if k >= 1:  # the previous snapshot must already be intact while this one is in flight
                    pass
                a.wait_positions()
                got.append(views[k & 1].copy())
            for k in range(4):
                b.step(1)
                want = b.positions()
                if path == 2:
                    helpers.assert_bit_equal(got[k], want, f"async hand-off, step {k}")
                else:  # cell path: fp32 accumulation order varies run to run
                    assert np.abs(got[k] - want).max() < 1e-4
        for p in bufs:
            mpm_b200.host_free(p)


def test_reference_shipping_scene_bit_exact(lib):
    """The scene the reference actually ships (MLSMPM3DFluidMultithreadGPU.cs:654-707): 64^3 grid, a centred 32^3 box at
    spacing 0.6 -> 54^3 = 157 464 particles, the GPU variant's constants, the sphere repulsor at the scene's default
    position, the UI's start-up gravity -0.5 (main_ui.tscn:162).  20 steps, strict arithmetic, both binned-or-not paths."""
    op = orc.variant("3d_gpu", 64)
    op.gravity = -0.5
    lo, hi = (16, 16, 16), (48, 48, 48)
    pos = orc.init_block(3, lo, hi, 0.6)
    assert pos.shape[0] == 157464
    ref = orc.State(op, pos)
    ref.step(20)
    for path in (1, 2):
        with make_solver(op, pos.shape[0], kernel_path=path) as s:
            assert s.initialise_sim(lo, hi, 0.6) == 157464
            s.process()            # one frame = sim_iterations (2) steps, as _Process does
            s.step(18)
            gp, gv, gc, _ = s.download()
            helpers.assert_bit_equal(gp, ref.pos, "pos"); helpers.assert_bit_equal(gv, ref.vel, "vel")
            helpers.assert_bit_equal(gc, ref.C, "C")
            helpers.assert_bit_equal(s.positions(), ref.positions(), "particle_pos_tex contents")
            _, width = s.positions_device()
            assert width == 397     # (uint)sqrt(157464) + 1, MLSMPM3DFluidMultithreadGPU.cs:196


def test_bench_line_has_the_contract_keys(lib):
    """bench.py on the small C2 scene: one JSON line with the keys the driver reads."""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--workload", "c2", "--steps", "5", "--warmup", "3",
                          "--no-cpu-baseline"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks", "kernels"):
        assert k in line, k
    assert line["value"] > 0 and line["gpu_launches"] > 0 and line["config"]["workload"].startswith("3D dam-break 64^3")
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(line["roofline"])
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(line["e2e"])


def test_sphere_list_colliders(lib):
    """a list of sphere repulsors (mpm_set_colliders) follows the rule of the reference's single sphere
    (g2p.glsl:122-129); strict paths bit-exact against the oracle, cell path within the FAST tolerance."""
    op = orc.variant("3d_gpu", 32)
    op.sphere_pos[:] = [10.0, 14.0, 16.0]
    op.sphere_radius = 4.0
    extra = [(20.0, 12.0, 16.0, 5.0), (16.0, 20.0, 10.0, 3.5), (16.0, 16.0, 22.0, 6.0)]
    op.n_extra_spheres = len(extra)
    for k, e in enumerate(extra):
        op.extra_spheres[k][:] = list(e)
    n = 30000
    pos, vel, Cm, mass = helpers.random_cloud(op, n, seed=71)
    ref = orc.State(op, pos, vel, Cm, mass)
    ref.step(4)
    plain = orc.variant("3d_gpu", 32)
    plain.sphere_pos[:] = [10.0, 14.0, 16.0]; plain.sphere_radius = 4.0
    ref1 = orc.State(plain, pos, vel, Cm, mass); ref1.step(4)
    assert np.abs(ref.vel - ref1.vel).max() > 0.5, "the extra spheres touch nothing in this scene"
    for path, math in ((1, 0), (2, 0), (3, 1)):
        with make_solver(op, n, kernel_path=path, math_mode=math) as s:
            s.set_colliders(extra)
            s.upload(pos, vel, Cm, mass)
            s.step(4)
            gp, gv, gc, _ = s.download()
        if math == 0:
            helpers.assert_bit_equal(gp, ref.pos, "pos"); helpers.assert_bit_equal(gv, ref.vel, "vel"); helpers.assert_bit_equal(gc, ref.C, "C")
        else:
            # a particle within rounding of a sphere's surface may take or miss the unit push: compare the bulk
            close = np.abs(gv - ref.vel).max(1) < 1e-3
            assert close.mean() > 0.999 and np.abs(gp - ref.pos)[close].max() < 1e-3


def test_checkpoint_resume_is_bit_exact(lib, tmp_path):
    """save after 5 steps, load into a fresh solver, run 5 more: identical to 10 uninterrupted steps."""
    import mpm_b200
    op = orc.variant("3d_gpu", 32)
    op.interaction = 0
    pos, vel, Cm, mass = helpers.random_cloud(op, 20000, seed=8)
    ref = orc.State(op, pos, vel, Cm, mass)
    ref.step(10)
    path = str(tmp_path / "state.mpm")
    with make_solver(op, 20000, kernel_path=2) as a:
        a.upload(pos, vel, Cm, mass)
        a.step(5)
        a.save_state(path)
    import ctypes as C
    assert os.path.getsize(path) == 64 + (C.sizeof(mpm_b200.MpmParams) + 4 + 4 * 4 * 7) + 80 * 20000  # header | parameters | records
    with make_solver(op, 20000, kernel_path=2) as b:
        b.load_state(path)
        assert b.num_particles == 20000 and b.stats().steps == 5
        b.step(5)
        gp, gv, gc, gm = b.download()
    helpers.assert_bit_equal(gp, ref.pos, "pos"); helpers.assert_bit_equal(gv, ref.vel, "vel")
    helpers.assert_bit_equal(gc, ref.C, "C"); helpers.assert_bit_equal(gm, ref.mass, "mass")
    other = orc.variant("3d_gpu", 40)
    with make_solver(other, 20000) as c:
        with pytest.raises(mpm_b200.MpmError):
            c.load_state(path)     # written for another grid
    stiffer = orc.variant("3d_gpu", 32)
    stiffer.interaction = 0
    stiffer.eos_stiffness = 2.0
    with make_solver(stiffer, 20000, kernel_path=2) as d:
        with pytest.raises(mpm_b200.MpmError, match="eos_stiffness"):
            d.load_state(path)     # the file carries the parameters it was produced with: different physics is refused


# ---------------------------------------------------------------- cell path: the binning is the stable sort, and reproducible
@pytest.mark.parametrize("grid,B", [(32, 4), (96, 8), ((128, 96, 96), 8)])
def test_cell_binning_permutation_is_the_stable_sort(lib, grid, B):
    """North-star subsystem 1 on the BENCHMARKED path: after every bin phase of the cell path the permutation
    (cell-major rank -> record index) equals std::stable_sort of the records by cell key, bit for bit, and the keys are the
    oracle's cell index (orc_cell_keys) of the positions the records hold.  The cloud is crowded (many equal keys) and a
    clump of ~2000 particles flies 4-5 cells per step, so the cold binning (radix sort after an upload), the warm
    binning (stable ranks from the previous layout) and its far-mover list are all exercised.  mpm_debug_last_sort also
    verifies on the device that the layout in place (src_of, ids) is the one the ranks describe."""
    op = orc.variant("3d_gpu", grid)
    op.interaction = 0
    n = 200001 if B == 8 else 40001                        # (4^3 blocks: a gentler scene, their apron is a quarter of the block)
    pos, vel, Cm, mass = helpers.random_cloud(op, n, seed=33, vel_sigma=0.5 if B == 8 else 0.2)
    crowd = 0.25 if B == 8 else 0.6                        # part of the cloud squeezed into a corner: many particles per cell
    pos[: n // 2] = pos[: n // 2] * crowd + 4.0
    mass[: n // 2] *= crowd ** 3
    R = np.array(op.grid[:3], np.float32)
    centre = R * np.array([0.3, 0.75, 0.75], np.float32)
    d2 = ((pos - centre) ** 2).sum(1)
    clump = np.argsort(d2)[:2000 if B == 8 else 500]        # the particles nearest to `centre` move together ...
    vel[clump] = np.array([22.0, 0.0, 0.0], np.float32)     # ... 4.4 cells per step: far movers, fewer than the list holds
    with make_solver(op, n, kernel_path=3, math_mode=1) as s:
        s.upload(pos, vel, Cm, mass)
        far_seen = 0
        for rnd in range(4):
            if rnd:
                s.step(1)                                   # a warm binning inside
            s.run_phase(5)                                  # cold (round 0) / warm binning of the current state
            st = s.stats()
            assert st.unordered_binnings == 0, "far-mover list overflowed: the scene is more violent than intended"
            far_seen += st.far_movers
            keys, perm = s.last_sort()
            ids = s.record_ids()
            assert np.array_equal(np.sort(ids), np.arange(n, dtype=np.uint32)), "record ids are not a permutation"
            expect = orc.stable_sort_perm(keys)             # == std::stable_sort order
            assert np.array_equal(perm.astype(np.int32), expect), f"round {rnd}: binning permutation is not the stable sort"
            gp = s.download()[0]                            # original index order; ends the binning's validity
            ref = orc.State(op, gp)
            want = helpers.block_major_key(op, ref.cell_keys(), B)
            assert np.array_equal(keys, want[ids]), f"round {rnd}: cell keys differ from orc_cell_keys"
            if rnd == 0:
                helpers.assert_bit_equal(gp, pos, "binning must not change or reorder the particles")
        print("STABLE_BINNING far movers seen over the warm rounds:", far_seen)
        assert far_seen > 100, "the scene produced no far movers: that branch went untested"


def test_cell_binning_far_movers_at_scale(lib):
    """A block of fluid thrown 12 cells per step: EVERY particle leaves its block's apron in every step.  The binning still
    equals the stable sort (far arrivals are placed behind a cell's regular ones, then merged into slot order), the steps
    match the oracle, and nothing is booked as unordered."""
    op = orc.variant("3d_gpu", 96)
    op.interaction = 0
    pos = orc.init_block(3, (10, 40, 40), (22, 52, 52), 0.5)  # 24^3 = 13824 particles
    n = pos.shape[0]
    vel = np.zeros_like(pos); vel[:, 0] = 60.0
    mass = np.full(n, 0.01, np.float32)                      # keeps mass * |v| inside the int32 x 1e7 range
    with make_solver(op, n, kernel_path=3, math_mode=1) as s:
        s.upload(pos, vel, mass=mass)
        s.step(3)
        s.run_phase(5)                                       # the binning of the state after 3 steps: all of it far movers
        st = s.stats()
        assert st.far_movers > 10000 and st.unordered_binnings == 0, (st.far_movers, st.unordered_binnings)
        keys, perm = s.last_sort()                           # (verifies the layout in place on the device)
        assert np.array_equal(perm.astype(np.int32), orc.stable_sort_perm(keys))
        gp, gv, gc, gm = s.download()
    ref = orc.State(op, pos, vel, mass=mass); ref.step(3)
    # (FAST against strict at |v| = 60: 36 cells travelled, a relative 1e-4 of that)
    assert np.abs(gp - ref.pos).max() < 4e-3 and helpers.rel_err(gv, ref.vel) < 1e-4, (np.abs(gp - ref.pos).max(), helpers.rel_err(gv, ref.vel))
    helpers.assert_bit_equal(gm, mass, "mass / order")


def test_cell_binning_too_many_far_arrivals_in_one_cell_is_counted(lib):
    """More than 32 far arrivals in ONE cell (200 particles of one cell thrown together): that cell is left in atomic order
    -- still a valid cell sort, the step matches the oracle -- and the binning is booked in MpmStats.unordered_binnings."""
    op = orc.variant("3d_gpu", 96)
    op.interaction = 0
    rng = np.random.default_rng(9)
    pos = (np.array([[20.0, 45.0, 45.0]], np.float32) + rng.uniform(0.05, 0.95, (200, 3)).astype(np.float32))
    vel = np.zeros_like(pos); vel[:, 0] = 60.0
    mass = np.full(200, 0.01, np.float32)
    with make_solver(op, 200, kernel_path=3, math_mode=1) as s:
        s.upload(pos, vel, mass=mass)
        s.step(3)
        assert s.stats().unordered_binnings >= 1
        gp, gv, _, gm = s.download()
    ref = orc.State(op, pos, vel, mass=mass); ref.step(3)
    # (200 fp32 addends per node before the one truncation, |v| = 60, 36 cells travelled: a relative 5e-4 of that)
    assert np.abs(gp - ref.pos).max() < 2e-2 and helpers.rel_err(gv, ref.vel) < 1e-3
    helpers.assert_bit_equal(gm, mass, "mass / order")


@pytest.mark.parametrize("scene", ["dam_break_32", "c3_block_drop_128"])
def test_cell_path_is_reproducible_bit_for_bit(lib, scene):
    """Two runs of the FAST cell path give identical bits (stable in-cell order => fixed fp32 accumulation order).  The
    large case is BASELINE config 3 (4 096 000 particles, 128^3), 12 steps."""
    if scene == "dam_break_32":
        op, lo, hi, steps = orc.variant("3d_gpu", 32), (4, 4, 4), (20, 20, 20), 40
    else:
        op, lo, hi, steps = orc.variant("3d_gpu", 128), (24, 24, 24), (104, 104, 104), 12
    op.interaction = 0
    runs = []
    for _ in range(2):
        with make_solver(op, 4096000 if scene != "dam_break_32" else 32768, kernel_path=3, math_mode=1) as s:
            s.initialise_sim(lo, hi, 0.5)
            s.step(steps)
            runs.append((s.download(), s.download_grid(), s.stats().unordered_binnings))
    for k, what in enumerate(("pos", "vel", "C", "mass")):
        helpers.assert_bit_equal(runs[0][0][k], runs[1][0][k], f"{scene}: {what} differs between two runs")
    helpers.assert_bit_equal(runs[0][1], runs[1][1], "grid differs between two runs")
    assert runs[0][2] == 0 and runs[1][2] == 0


def test_quantised_position_hand_off_matches_the_float_one(lib):
    """mpm_get_positions_q16_async: 4 x uint16 per particle (x, y, z as fractions of the domain, |v| as binary16) in original
    index order, through the same double-buffered copy as the float4 hand-off.  Decoded positions are within half a code
    step (grid_size / 65535 / 2, plus the fp32 rounding of the scaling) of mpm_get_positions, |v| within binary16 rounding."""
    import mpm_b200
    op = orc.variant("3d_gpu", 64)
    op.interaction = 0
    for path, math in ((3, 1), (2, 0)):
        with make_solver(op, 262144, kernel_path=path, math_mode=math) as s:
            n = s.initialise_sim((4, 4, 4), (36, 36, 36), 0.5)
            s.step(7)
            ref = s.positions().copy()
            buf = mpm_b200.host_alloc(8 * n)
            try:
                s.positions_q16_into_async(buf, n)
                s.wait_positions()
                q = np.ctypeslib.as_array((C.c_uint16 * (4 * n)).from_address(buf)).reshape(n, 4).copy()
            finally:
                mpm_b200.host_free(buf)
        dec = q[:, :3].astype(np.float64) * 64.0 / 65535.0
        assert np.abs(dec - ref[:, :3]).max() <= 0.505 * 64.0 / 65535.0  # (half a code step, and the fp32 rounding of p * scale)
        speed = q[:, 3].copy().view(np.float16).astype(np.float64)
        assert np.abs(speed - ref[:, 3]).max() <= 1e-3 * max(1.0, float(ref[:, 3].max()))


def test_programmatic_dependent_launch_changes_no_bit(lib):
    """The per-step kernels are launched with the programmatic-stream-serialization attribute and start with
    griddepcontrol.wait (mpm_common.cuh: pdl_prologue): the next grid's CTAs are scheduled while the current grid drains.
    That must only remove idle time.  The same scene stepped in a child process with the attribute (default) and without
    (MPM_NO_PDL=1) ends in the same bits, for a grid with 4^3-cell blocks and one with 8^3-cell blocks."""
    import subprocess
    tool = os.path.join(os.path.dirname(__file__), "tools", "state_hash.py")
    for grid, steps in ((64, 30), (128, 12)):
        outs = []
        for extra in ({}, {"MPM_NO_PDL": "1"}):
            env = dict(os.environ, **extra)
            env.pop("MPM_PDL_MASK", None)
            r = subprocess.run([sys.executable, tool, str(grid), str(steps)], capture_output=True, text=True, timeout=300, env=env)
            assert r.returncode == 0, r.stderr[-2000:]
            outs.append(r.stdout.split()[-3:])
        assert outs[0] == outs[1], (grid, outs)
        assert outs[0][0] == "0"  # every binning in the stable order: the runs are comparable bit for bit


def test_timing_levels_agree(lib):
    """mpm_set_timing(2) brackets the whole mpm_step() call with two events; level 1 adds events around every phase.  Both
    report the same kind of number: the per-phase sum of level 1 is within its own ms_step, and the level-2 ms_step is not
    larger than level 1's (the events themselves cost time) beyond noise; the per-phase fields read 0 at level 2."""
    op = orc.variant("3d_gpu", 64)
    op.interaction = 0
    with make_solver(op, 262144, kernel_path=3, math_mode=1) as s:
        s.initialise_sim((4, 4, 4), (36, 36, 36), 0.5)
        s.step(5); s.sync()
        s.set_timing(2); s.step(40); s.sync()
        st2 = s.stats()
        s.set_timing(1); s.step(40); s.sync()
        st1 = s.stats()
    assert st2.ms_step > 0 and st2.ms_p2g1 == 0 and st2.ms_sort == 0
    phases = st1.ms_sort + st1.ms_clear + st1.ms_p2g1 + st1.ms_p2g2 + st1.ms_update + st1.ms_g2p
    assert 0 < phases <= st1.ms_step * 1.001
    assert st2.ms_step <= st1.ms_step * 1.10, (st2.ms_step, st1.ms_step)


# ---------------------------------------------------------------- cell path against the ORACLE on the BASELINE configs
def _particle_report(tag, gp, gv, gc, ref):
    dp = np.abs(gp.astype(np.float64) - ref.pos).max(1)
    dv = np.abs(gv.astype(np.float64) - ref.vel).max(1)
    rep = dict(pos_med=float(np.median(dp)), pos_p999=float(np.quantile(dp, 0.999)), pos_max=float(dp.max()),
               vel_med=float(np.median(dv)), vel_p999=float(np.quantile(dv, 0.999)), vel_max=float(dv.max()),
               vel_scale=float(np.abs(ref.vel).max()), C_rel=helpers.rel_err(gc, ref.C),
               com=float(np.abs(gp.mean(0) - ref.pos.mean(0)).max()),
               ke_rel=float(abs((gv.astype(np.float64) ** 2).sum() - (ref.vel.astype(np.float64) ** 2).sum()) / max((ref.vel.astype(np.float64) ** 2).sum(), 1e-30)))
    print("CELL_VS_ORACLE", tag, {k: f"{v:.3g}" for k, v in rep.items()})
    return rep


def _assert_drift(rep, op, steps):
    """Stated N-step bound of the FAST cell path against the strict oracle, PER PARTICLE: the per-step FAST tolerances
    (positions 1e-6 of the domain, velocities 2e-5 of the largest velocity) accumulated linearly over the steps for the
    worst particle, a fifth of that for the median one; centre of mass and kinetic energy to 1e-5.  (Measured on B200,
    20 steps: worst position error 1e-4 cells against the 1.3e-3 allowed, worst velocity error 7e-5 of the scale.)"""
    R = max(op.grid[:3])
    pos_tol, vel_tol = steps * FAST_TOL["pos"] * R, steps * FAST_TOL["vel"] * rep["vel_scale"]
    assert rep["pos_max"] <= pos_tol and rep["pos_med"] <= 0.2 * pos_tol, (rep, pos_tol)
    assert rep["vel_max"] <= vel_tol and rep["vel_med"] <= 0.2 * vel_tol, (rep, vel_tol)
    assert rep["com"] < 1e-5 and rep["ke_rel"] < 1e-5, rep


def test_cell_path_config2_dam_break_vs_oracle_per_particle(lib):
    """BASELINE config 2 (the reference's own scene size): 64^3 grid, dam-break block [4,36)^3 at spacing 0.5 = 262 144
    particles, GPU-variant constants, 20 steps, FAST cell path against the strict oracle, PER PARTICLE, within the stated
    20-step bound (_assert_drift).  For scale the oracle's own answer to a 1e-6-cell perturbation of the input is printed."""
    op = orc.variant("3d_gpu", 64)
    op.interaction = 0
    lo, hi = (4, 4, 4), (36, 36, 36)
    pos = orc.init_block(3, lo, hi, 0.5)
    assert pos.shape[0] == 262144
    ref = orc.State(op, pos); ref.step_mt(20)
    rng = np.random.default_rng(1)
    pert = orc.State(op, pos + rng.uniform(-1e-6, 1e-6, pos.shape).astype(np.float32)); pert.step_mt(20)
    sens_p = np.abs(pert.pos.astype(np.float64) - ref.pos).max(1)
    sens_v = np.abs(pert.vel.astype(np.float64) - ref.vel).max(1)
    with make_solver(op, pos.shape[0], kernel_path=3, math_mode=1) as s:
        assert s.initialise_sim(lo, hi, 0.5) == 262144
        s.step(20)
        gp, gv, gc, gm = s.download()
        assert s.stats().kernel_path == 3
    rep = _particle_report("c2", gp, gv, gc, ref)
    print("ORACLE_SENSITIVITY c2 (the oracle's own answer to a 1e-6 perturbation of the input, for scale)",
          dict(pos_med=float(np.median(sens_p)), pos_max=float(sens_p.max()), vel_med=float(np.median(sens_v)), vel_max=float(sens_v.max())))
    _assert_drift(rep, op, 20)
    helpers.assert_bit_equal(gm, np.ones(262144, np.float32), "mass / count")


def test_cell_path_shipping_scene_vs_oracle(lib):
    """The scene the reference ships (H:654-707): 64^3, centred 32^3 box at spacing 0.6 = 157 464 particles, sphere
    repulsor at the scene's default position, the UI's gravity -0.5, one frame = 2 steps (_Process) + 18 more, on the FAST
    cell path -- the path a maintainer gets by default -- against the oracle, per particle, same bar as config 2."""
    op = orc.variant("3d_gpu", 64)
    op.gravity = -0.5
    lo, hi = (16, 16, 16), (48, 48, 48)
    pos = orc.init_block(3, lo, hi, 0.6)
    ref = orc.State(op, pos); ref.step_mt(20)
    rng = np.random.default_rng(1)
    pert = orc.State(op, pos + rng.uniform(-1e-6, 1e-6, pos.shape).astype(np.float32)); pert.step_mt(20)
    sens_p = np.abs(pert.pos.astype(np.float64) - ref.pos).max(1)
    sens_v = np.abs(pert.vel.astype(np.float64) - ref.vel).max(1)
    with make_solver(op, pos.shape[0], kernel_path=3, math_mode=1) as s:
        assert s.initialise_sim(lo, hi, 0.6) == 157464
        s.process(); s.step(18)
        gp, gv, gc, _ = s.download()
        p4 = s.positions()
        _, width = s.positions_device()
    rep = _particle_report("shipping", gp, gv, gc, ref)
    _assert_drift(rep, op, 20)
    assert width == 397
    helpers.assert_bit_equal(p4[:, :3], gp, "particle_pos_tex xyz == particle positions")
    assert np.abs(p4[:, 3] - np.sqrt((gv.astype(np.float64) ** 2).sum(1))).max() < 1e-5


def test_cell_path_config3_one_step_vs_oracle(lib):
    """BASELINE config 3 (128^3, 4 096 000 particles): one full step of the FAST cell path against the oracle (all host
    cores, fixed-point atomics: bit-identical to the serial oracle), per phase grid and per particle, FAST tolerance."""
    op = orc.variant("3d_gpu", 128)
    op.interaction = 0
    lo, hi = (24, 24, 24), (104, 104, 104)
    pos = orc.init_block(3, lo, hi, 0.5)
    assert pos.shape[0] == 4096000
    rng = np.random.default_rng(5)
    vel = rng.normal(0, 0.3, pos.shape).astype(np.float32)   # a resting lattice would leave most terms zero
    Cm = rng.normal(0, 0.05, (pos.shape[0], 9)).astype(np.float32)
    ref = orc.State(op, pos, vel, Cm)
    ref.step_mt(1)
    with make_solver(op, pos.shape[0], kernel_path=3, math_mode=1) as s:
        s.upload(pos, vel, Cm)
        s.step(1)
        g = s.download_grid().astype(np.float64) / 1e7
        gp, gv, gc, gm = s.download()
    gr = ref.grid.astype(np.float64) / 1e7
    # after UpdateGrid the cells hold velocity and mass: judge the velocity by the momentum it stands for
    mom_err = float(np.abs((g[:, :3] - gr[:, :3]) * gr[:, 3:4]).max() / np.abs(gr[:, :3] * gr[:, 3:4]).max())
    mass_err = helpers.rel_err(g[:, 3], gr[:, 3])
    e = dict(mom=mom_err, mass=mass_err, pos=float(np.abs(gp.astype(np.float64) - ref.pos).max() / 128),
             vel=helpers.rel_err(gv, ref.vel), C=helpers.rel_err(gc, ref.C), vel_elem=helpers.elem_err(gv, ref.vel, 1e-2))
    print("CELL_C3_ONE_STEP", {k: f"{v:.3g}" for k, v in e.items()})
    assert e["mom"] <= FAST_TOL["grid"] and e["mass"] <= FAST_TOL["grid"]
    assert e["pos"] <= FAST_TOL["pos"] and e["vel"] <= FAST_TOL["vel"] and e["C"] <= FAST_TOL["C"] and e["vel_elem"] <= 2e-4


def test_cell_path_config4_invariants(lib):
    """BASELINE config 4, the benchmarked workload (256^3, 32 768 000 particles): size-independent properties.
    (1) after P2G_1 the grid mass equals the particles' encoded mass up to one truncation per (cell, node);
    (2) P2G_2 adds no net momentum (internal forces cancel) beyond truncation; (3) three steps keep every particle, in
    the original (lattice) order, finite and inside the clamp box; (4) total momentum changes by gravity alone."""
    op = orc.variant("3d_gpu", 256)
    op.interaction = 0
    n = 32768000
    with make_solver(op, n, kernel_path=3, math_mode=1) as s:
        assert s.initialise_sim((4, 4, 4), (164, 164, 164), 0.5) == n
        s.run_phase(5); s.run_phase(0); s.run_phase(1)
        g1 = s.download_grid().astype(np.int64)
        total = int(g1[:, 3].sum())
        occupied_nodes = int((g1[:, 3] != 0).sum())
        assert abs(n * 10_000_000 - total) <= 27 * occupied_nodes + 1e-7 * n * 1e7, (n * 10_000_000 - total, occupied_nodes)
        assert not g1[:, :3].any(), "a resting lattice with v = 0, C = 0 carries no momentum"
        s.run_phase(2)
        g2 = s.download_grid().astype(np.int64)
        net = np.abs(g2[:, :3].sum(0))
        assert (net <= 27 * occupied_nodes).all(), net
        del g1, g2
        s.step(3)
        gp, gv, gc, gm = s.download()
        assert s.stats().unordered_binnings == 0
    assert gp.shape[0] == n and np.isfinite(gp).all() and np.isfinite(gv).all() and np.isfinite(gc).all()
    assert gp.min() >= 2.0 and gp.max() <= 254.0
    helpers.assert_bit_equal(gm, np.ones(n, np.float32), "mass")
    xm = gp[:, 0].reshape(320, -1).mean(1)                 # lattice order: index = (ix * 320 + iy) * 320 + iz
    assert np.all(np.diff(xm) > 0.25), "original (lattice) order lost"
    # three steps of free fall at most (the block rests on the floor and against two walls, which take some of it back)
    py = gv[:, 1].astype(np.float64).mean()
    assert 1.05 * 3 * op.dt * op.gravity < py < 0.0, py


def test_cell_path_update_grid_mass_weighted(lib):
    """UpdateGrid on the cell path, judged by what a node's velocity stands for: |v - v_ref| * mass against the largest
    momentum.  (The plain max-norm of the velocity grid is dominated by nodes with almost no mass, where one fixed-point
    unit of momentum is a large velocity change that no particle ever sees.)"""
    for variant, grid in (("3d_fixed", 32), ("3d_gpu", (40, 32, 24)), ("3d_gpu", 96)):
        op = orc.variant(variant, grid)
        sphere_into_cloud(op)
        n = 20000
        pos, vel, Cm, mass = helpers.random_cloud(op, n, seed=21)
        ref = orc.State(op, pos, vel, Cm, mass)
        ref.clear_grid(); ref.p2g1(); ref.p2g2(); ref.update_grid()
        with make_solver(op, n, kernel_path=3, math_mode=1) as s:
            s.upload(pos, vel, Cm, mass)
            for k in range(4):
                s.run_phase(k)
            g = s.download_grid().astype(np.float64) / 1e7
        gr = ref.grid.astype(np.float64) / 1e7
        err = float(np.abs((g[:, :3] - gr[:, :3]) * gr[:, 3:4]).max() / np.abs(gr[:, :3] * gr[:, 3:4]).max())
        print("CELL_UPDATE_GRID_MASS_WEIGHTED", variant, grid, f"{err:.3g}")
        assert err <= FAST_TOL["grid"], (variant, grid, err)
        assert helpers.rel_err(g[:, 3], gr[:, 3]) <= FAST_TOL["grid"]


# ---------------------------------------------------------------- 2D mouse radial push (D:381-406)
@pytest.mark.parametrize("variant", ["2d_st", "2d_mt"])
def test_mouse_radial_push_2d(lib, variant):
    """MPM_INTERACT_MOUSE_2D: the reference's 2D solvers push particles radially away from the held-down mouse
    (MLSMPM2DFluid.cs:381-406).  Float grid (atomic order differs from the serial oracle): calibrated tolerance as in
    test_float_grid_within_calibrated_tolerance; the push itself must be visible."""
    op = orc.variant(variant, (64, 64, 1))
    n = 4000
    pos, vel, Cm, mass = helpers.random_cloud(op, n, seed=12, margin=8.0)
    op.interaction = orc.INTERACT_MOUSE_2D
    op.mouse_pos[:] = [30.0, 34.0]
    op.mouse_radius = 10.0
    off = orc.variant(variant, (64, 64, 1))
    sens, ref = shuffle_sensitivity(op, pos, vel, Cm, mass, 5)
    quiet = orc.State(off, pos, vel, Cm, mass); quiet.step(5)
    assert np.abs(ref.vel - quiet.vel).max() > 0.05, "the mouse pushes nothing in this scene"
    tol = max(4.0 * sens, 2e-6)
    with make_solver(op, n) as s:
        s.upload(pos, vel, Cm, mass)
        s.step(5)
        gp, gv, gc, _ = s.download()
    for what, a, b in (("pos", gp, ref.pos), ("vel", gv, ref.vel), ("C", gc, ref.C)):
        e = helpers.rel_err(a, b)
        assert e <= tol, f"{variant} mouse push {what}: rel err {e:.3g} > tol {tol:.3g}"


# ---------------------------------------------------------------- guards: positions outside the grid, fixed-point overflow
def test_bad_positions_are_rejected_and_never_index_outside_the_grid(lib):
    """The reference indexes the grid with (int)pos unchecked (F:281-283: IndexOutOfRangeException).  Here an upload with
    a non-finite position or one whose stencil leaves the grid is refused (MPM_ERR_DOMAIN), on every kernel path."""
    import mpm_b200
    op = orc.variant("3d_gpu", 32)
    op.interaction = 0
    pos, vel, Cm, mass = helpers.random_cloud(op, 5000, seed=3)
    for path, math in ((1, 0), (2, 0), (3, 1)):
        with make_solver(op, 5000, kernel_path=path, math_mode=math) as s:
            for bad in (np.nan, np.inf, 0.5, 31.5, -3.0, 1e9):
                q = pos.copy(); q[1234, 1] = bad
                with pytest.raises(mpm_b200.MpmError) as e:
                    s.upload(q, vel, Cm, mass)
                assert e.value.code == mpm_b200.ERR_DOMAIN and s.num_particles == 0
            s.upload(pos, vel, Cm, mass)       # the clean set still works afterwards
            s.step(2); s.sync()
            assert np.isfinite(s.download()[0]).all()


def test_overflow_detector(lib):
    """MpmParams.overflow_check: a node accumulating more than 2^31 / 1e7 = 214.7 mass units wraps the reference's int32 grid
    silently; with the detector on, mpm_sync returns MPM_ERR_OVERFLOW and MpmStats.overflow is set.  AUTO + FAST keeps off
    the cell path when the detector is requested, and asking for both explicitly is refused."""
    import mpm_b200
    op = orc.variant("3d_gpu", 32)
    op.interaction = 0
    rng = np.random.default_rng(5)
    pos = (np.array([[12.0, 13.0, 14.0]], np.float32) + rng.uniform(0.01, 0.99, (3000, 3)).astype(np.float32))
    heavy = np.full(3000, 1.0, np.float32)      # 3000 mass units on 27 nodes: far beyond 214.7
    light = np.full(3000, 0.001, np.float32)
    for path, math in ((1, 0), (2, 0), (2, 1), (0, 1)):
        with make_solver(op, 3000, kernel_path=path, math_mode=math, overflow_check=1) as s:
            assert s.stats().kernel_path in (1, 2)
            s.upload(pos, mass=light)
            s.run_phase(0); s.run_phase(1); s.sync()
            assert s.stats().overflow == 0
            s.upload(pos, mass=heavy)
            s.run_phase(0); s.run_phase(1)
            with pytest.raises(mpm_b200.MpmError) as e:
                s.sync()
            assert e.value.code == mpm_b200.ERR_OVERFLOW and s.stats().overflow == 1
    with pytest.raises(mpm_b200.MpmError):
        make_solver(op, 100, kernel_path=3, math_mode=1, overflow_check=1)


# ---------------------------------------------------------------- zero-copy hand-off
def test_zero_copy_hand_off_to_another_process(lib):
    """The reference's positions never leave the GPU (G2P writes the texture the MultiMesh shader samples, H:340-355,
    402-412).  mpm_export_positions gives a file descriptor of the allocation that holds the (x, y, z, |v|) array; a second
    PROCESS imports it through the CUDA driver API (as a renderer would through VK_KHR_external_memory_fd) and reads, with no
    copy through the exporting process, exactly the bytes mpm_get_positions returns -- also after further steps, without
    exporting again."""
    import hashlib
    import subprocess
    op = orc.variant("3d_gpu", 32)
    lo, hi = (4, 4, 4), (20, 20, 20)
    tool = os.path.join(os.path.dirname(__file__), "tools", "import_positions.py")
    for path, math in ((3, 1), (2, 0)):
        with make_solver(op, 32768, kernel_path=path, math_mode=math) as s:
            n = s.initialise_sim(lo, hi, 0.5)
            s.step(3)
            fd, nbytes, width = s.export_positions()
            assert fd >= 0 and nbytes >= 16 * n and width == int(np.sqrt(np.float32(n))) + 1
            try:
                for more in (0, 2):
                    if more:
                        s.step(more)
                    s.refresh_positions(); s.sync()          # device-side only: nothing is copied to the host here
                    out = subprocess.run([sys.executable, tool, str(fd), str(nbytes), str(n)], pass_fds=(fd,), capture_output=True,
                                         text=True, timeout=120)
                    assert out.returncode == 0, out.stderr[-2000:]
                    want = hashlib.sha256(s.positions().tobytes()).hexdigest()
                    assert out.stdout.split()[-1] == want, "the importing process sees other bytes than mpm_get_positions"
            finally:
                os.close(fd)
