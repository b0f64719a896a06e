"""CPU tests of the oracle itself (it is the checker, so it is checked first).

PARITY UNPINNED: no reference golden vectors exist; the oracle is pinned by (1) physical invariants of
the algorithm, (2) bit-for-bit agreement between the C restatement and the independent NumPy one, and
(3) the committed fixtures in tests/golden (made by the NumPy restatement)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import helpers
from oracle import orc, oracle_np as onp

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import make_golden  # noqa: E402


@pytest.mark.parametrize("name", helpers.VARIANT_NAMES)
def test_c_oracle_matches_numpy_restatement(name):
    p = orc.variant(name, 16 if name.startswith("3d") else 32)
    if p.interaction in (1, 2):
        p.sphere_pos[:] = [5.0, 8.0, 8.0]
        p.sphere_radius = 4.0
    pos, vel, Cm, mass = helpers.random_cloud(p, 500, seed=3)
    s = orc.State(p, pos, vel, Cm, mass)
    q = dict(pos=pos.copy(), vel=vel.copy(), C=Cm.copy(), grid=np.zeros_like(s.grid))
    for phase in ("clear_grid", "p2g1", "p2g2", "update_grid", "g2p"):
        getattr(s, phase)()
        if phase == "clear_grid":
            onp.clear_grid(p, q["grid"])
        elif phase == "p2g1":
            onp.p2g1(p, q["pos"], q["vel"], q["C"], mass, q["grid"])
        elif phase == "p2g2":
            onp.p2g2(p, q["pos"], q["C"], mass, q["grid"])
        elif phase == "update_grid":
            onp.update_grid(p, q["grid"])
        else:
            onp.g2p(p, q["pos"], q["vel"], q["C"], q["grid"])
        helpers.assert_bit_equal(s.grid, q["grid"], f"{name} grid after {phase}")
    for k, a in (("pos", s.pos), ("vel", s.vel), ("C", s.C)):
        helpers.assert_bit_equal(a, q[k], f"{name} {k}")


@pytest.mark.parametrize("name", helpers.VARIANT_NAMES)
def test_c_oracle_reproduces_golden(name):
    p, n, steps, seed = make_golden.case_params(name)
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", f"{name}.npz"))
    s = orc.State(p, z["pos0"], z["vel0"], z["C0"], z["mass"])
    s.step(int(z["steps"]))
    helpers.assert_bit_equal(s.grid, z["grid"], "grid")
    helpers.assert_bit_equal(s.pos, z["pos"], "pos")
    helpers.assert_bit_equal(s.vel, z["vel"], "vel")
    helpers.assert_bit_equal(s.C, z["C"], "C")


@pytest.mark.parametrize("name", ["3d_float", "2d_st"])
def test_partition_of_unity_and_momentum(name):
    """P2G_1 conserves mass and momentum (sum w = 1, sum w*d = 0); P2G_2 adds none."""
    p = orc.variant(name, 16 if name.startswith("3d") else 32)
    pos, vel, Cm, mass = helpers.random_cloud(p, 400, seed=5)
    s = orc.State(p, pos, vel, Cm, mass)
    s.clear_grid(); s.p2g1()
    g = s.grid_f().astype(np.float64)
    assert abs(g[:, 3].sum() - mass.astype(np.float64).sum()) < 1e-3
    mom = (mass[:, None].astype(np.float64) * vel.astype(np.float64)).sum(0)
    # sum_i w_i (v + C d_i) = v + C * sum w d = v only for quadratic weights' first moment: sum w*d = 0
    assert np.abs(g[:, :3].sum(0) - mom)[: p.dim].max() < 2e-3
    before = g[:, :3].sum(0)
    s.p2g2()
    after = s.grid_f().astype(np.float64)[:, :3].sum(0)
    assert np.abs(after - before).max() < 2e-3


def test_lattice_at_rest_stays_at_rest_without_gravity():
    """interior of a uniform lattice at rest density feels no net force."""
    p = orc.variant("3d_float", 24)
    p.gravity = 0.0
    pos = orc.init_block(3, (6, 6, 6), (18, 18, 18), 0.5)  # commensurate lattice: uniform density inside
    s = orc.State(p, pos)
    s.step(1)
    centre = np.all(np.abs(pos - 11.75) < 2.0, axis=1)  # > 3 cells (two stencil reaches) from the surface
    assert centre.sum() == 512 and np.abs(s.vel[centre]).max() < 1e-6
    assert np.abs(s.vel[~centre]).max() > 1e-2  # while the free surface does accelerate


def test_fixed_point_codec():
    L = orc.lib()
    assert L.orc_encode_fixed(C.c_float(1.23456789), 10_000_000) == int(np.float32(1.23456789) * np.float32(1e7))
    assert L.orc_encode_fixed(C.c_float(-0.9e-7), 10_000_000) == 0  # truncation toward zero
    assert L.orc_encode_fixed(C.c_float(-2.5e-7), 10_000_000) == -2
    assert L.orc_decode_fixed(12345678, 10_000_000) == np.float32(12345678) / np.float32(1e7)


def test_pow_semantic_within_one_ulp_of_libm():
    """The parity pow (binary64, rounded once) and this host's powf never differ by more than 1 ulp."""
    L = orc.lib()
    rng = np.random.default_rng(0)
    xs = rng.uniform(0.05, 3.0, 20000).astype(np.float32)
    worst = 0
    for y in (4.0, 7.0, 2.5):
        a = np.array([L.orc_pow(0, C.c_float(x), C.c_float(y)) for x in xs[:4000]], np.float32)
        b = np.array([L.orc_pow(1, C.c_float(x), C.c_float(y)) for x in xs[:4000]], np.float32)
        ulp = np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64)).max()
        worst = max(worst, int(ulp))
    assert worst <= 1


def test_init_block_matches_reference_counts():
    # MLSMPM3DFluidMultithreadGPU.cs:658-671: centred 32^3 box at spacing 0.6 in a 64^3 grid -> 54^3
    pos = orc.init_block(3, (16, 16, 16), (48, 48, 48), 0.6)
    assert pos.shape[0] == 54 ** 3 == 157464
    # MLSMPM3DFluidMultithread.cs:133-146: 16^3 box, spacing 0.5 -> 32768; 2D (MLSMPM2DFluid.cs:130-140): 1024
    assert orc.init_block(3, (8, 8, 8), (24, 24, 24), 0.5).shape[0] == 32768
    assert orc.init_block(2, (16, 16), (48, 48), 1.0).shape[0] == 1024
    # fp32 accumulation: the 0.6 lattice is NOT lo + k*0.6 exactly
    x = np.unique(pos[:, 0])
    acc = np.float32(16.0)
    for k in range(len(x)):
        assert x[k] == acc
        acc = np.float32(acc + np.float32(0.6))


def test_mt_step_equals_serial_in_fixed_mode():
    """Integer atomics commute: the all-core baseline path gives the serial result bit for bit."""
    p = orc.variant("3d_fixed", 16)
    p.interaction = 0
    pos, vel, Cm, mass = helpers.random_cloud(p, 5000, seed=9)
    a = orc.State(p, pos, vel, Cm, mass); b = orc.State(p, pos, vel, Cm, mass)
    a.step(2); b.step_mt(2, 4)
    helpers.assert_bit_equal(a.grid, b.grid, "grid"); helpers.assert_bit_equal(a.pos, b.pos, "pos")


def test_stable_sort_perm_is_stable():
    rng = np.random.default_rng(1)
    keys = rng.integers(0, 50, 10000).astype(np.uint32)
    perm = orc.stable_sort_perm(keys)
    assert np.array_equal(perm, np.argsort(keys, kind="stable").astype(np.int32))


@pytest.mark.parametrize("name", ["2d_st", "2d_mt"])
def test_mouse_radial_push_c_oracle_matches_numpy(name):
    """MPM_INTERACT_MOUSE_2D (MLSMPM2DFluid.cs:381-406): particles inside the mouse radius are pushed radially, with a
    strength that grows towards the centre; the C restatement and the NumPy one agree bit for bit, and the push is there."""
    p = orc.variant(name, 32)
    p.interaction = orc.INTERACT_MOUSE_2D
    p.mouse_pos[:] = [15.0, 17.0]
    p.mouse_radius = 6.0
    pos, vel, Cm, mass = helpers.random_cloud(p, 800, seed=4)
    pos[0, :2] = [15.0, 17.0]                 # one particle exactly on the mouse: the zero-length normal (D:396-399)
    s = orc.State(p, pos, vel, Cm, mass)
    quiet_p = orc.variant(name, 32)
    quiet = orc.State(quiet_p, pos, vel, Cm, mass)
    q = dict(pos=pos.copy(), vel=vel.copy(), C=Cm.copy(), grid=np.zeros_like(s.grid))
    for _ in range(3):
        s.step(1); quiet.step(1)
        onp.clear_grid(p, q["grid"]); onp.p2g1(p, q["pos"], q["vel"], q["C"], mass, q["grid"])
        onp.p2g2(p, q["pos"], q["C"], mass, q["grid"]); onp.update_grid(p, q["grid"])
        onp.g2p(p, q["pos"], q["vel"], q["C"], q["grid"])
    helpers.assert_bit_equal(s.pos, q["pos"], "pos"); helpers.assert_bit_equal(s.vel, q["vel"], "vel")
    assert np.isfinite(s.vel).all()
    d = np.linalg.norm(pos[:, :2] - np.array([15.0, 17.0], np.float32), axis=1)
    pushed = np.abs(s.vel - quiet.vel).max(1)
    assert pushed[(d > 0.5) & (d < 4.0)].min() > 1e-3 and pushed[d > 9.0].max() < 0.05
