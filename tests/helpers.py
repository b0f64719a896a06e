"""Shared test helpers: scenes, parameter translation oracle <-> C ABI, comparisons."""
import numpy as np

from oracle import orc

VARIANT_NAMES = ["2d_st", "2d_mt", "3d_float", "3d_fixed", "3d_gpu"]

# fields that exist with the same meaning in OrcParams and MpmParams
_SHARED = ["dt", "gravity", "rest_density", "dynamic_viscosity", "eos_stiffness", "eos_power", "grid_mode",
           "fixed_point_mult", "stress_form", "eq16_order", "bc_mode", "bc_hi_off", "bc_friction", "clamp_min",
           "clamp_max_off", "wall_min", "wall_max_off", "wall_gain", "interaction", "sphere_radius", "mouse_radius"]


def mpm_params_from_orc(op, **overrides):
    """MpmParams (C ABI) carrying exactly the constants of an OrcParams."""
    import mpm_b200
    p = mpm_b200.MpmParams()
    p.struct_size = __import__("ctypes").sizeof(mpm_b200.MpmParams)
    p.dim = op.dim
    p.grid_size[:] = [op.grid[0], op.grid[1], op.grid[2] if op.dim == 3 else 1]
    for f in _SHARED:
        setattr(p, f, getattr(op, f))
    p.sphere_pos[:] = list(op.sphere_pos)
    p.mouse_pos[:] = list(op.mouse_pos)
    for k, v in overrides.items():
        setattr(p, k, v)
    return p


def random_cloud(op, n, seed=0, margin=3.5, vel_sigma=0.3, c_sigma=0.1):
    """Random particle cloud well inside the walls, random vel / C / mass (exercises every term)."""
    rng = np.random.default_rng(seed)
    R = np.array([op.grid[0], op.grid[1], op.grid[2] if op.dim == 3 else 1], np.float32)
    pos = np.zeros((n, 3), np.float32)
    for a in range(op.dim):
        pos[:, a] = rng.uniform(margin, R[a] - margin, n).astype(np.float32)
    vel = rng.normal(0, vel_sigma, (n, 3)).astype(np.float32)
    Cm = rng.normal(0, c_sigma, (n, 9)).astype(np.float32)
    mass = rng.uniform(0.5, 1.5, n).astype(np.float32)
    if op.dim == 2:
        vel[:, 2] = 0
        Cm[:, [2, 5, 6, 7, 8]] = 0
    return pos, vel, Cm, mass


def dam_break_block(op, lo, hi, spacing):
    return orc.init_block(op.dim, lo, hi, spacing)


def bits(a):
    return np.ascontiguousarray(a).view(np.int32)


def assert_bit_equal(a, b, what):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    same = bits(a) == bits(b)
    if not same.all():
        bad = np.argwhere(~same)
        i = tuple(bad[0])
        raise AssertionError(f"{what}: {bad.shape[0]} of {same.size} words differ; first at {i}: {a[i]!r} vs {b[i]!r}")


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def elem_err(a, b, atol):
    """Element-wise error: max over elements of |a - b| / (atol + |b|).  Unlike rel_err (a norm-wise bound: max |a - b| over
    max |b|) a small entry cannot hide behind a large one; `atol` is the magnitude below which an entry is noise."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float((np.abs(a - b) / (atol + np.abs(b))).max())


def block_major_key(op, cell_index, B):
    """The solver's bin key of the oracle's cell index (orc_cell_keys: (cx * Ry + cy) * Rz + cz, F:259 + F:282): the same
    formula applied to BLOCK coordinates, then to the cell inside the B^3 block -- key = block << 3 log2(B) | cell-in-block."""
    Ry, Rz = op.grid[1], op.grid[2]
    ci = np.asarray(cell_index, np.int64)
    cx, cy, cz = ci // (Ry * Rz), (ci // Rz) % Ry, ci % Rz
    nby, nbz = -(-Ry // B), -(-Rz // B)
    lb = {4: 2, 8: 3}[B]
    blk = ((cx // B) * nby + cy // B) * nbz + cz // B
    return ((blk << (3 * lb)) | ((cx % B) << (2 * lb)) | ((cy % B) << lb) | (cz % B)).astype(np.uint32)
