"""ctypes wrapper around oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
PARITY UNPINNED (see mpm_oracle.h): the reference ships no golden vectors and cannot run in this image.

Variant presets restate the constants table of SURVEY.md 8a (citations in mpm_oracle.h).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

GRID_FLOAT, GRID_FIXED = 0, 1
STRESS_3D, STRESS_2D_TRACE = 0, 1
EQ16_VOL4_DT, EQ16_DTVOL_4 = 0, 1
BC_SLIP, BC_FRICTION = 0, 1
INTERACT_NONE, INTERACT_SPHERE_POST, INTERACT_SPHERE_PRE, INTERACT_MOUSE_2D = 0, 1, 2, 3
POW_F64_ROUNDED, POW_LIBM_POWF = 0, 1


class OrcParams(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("grid", C.c_int32 * 3),
        ("dt", C.c_float), ("gravity", C.c_float), ("rest_density", C.c_float),
        ("dynamic_viscosity", C.c_float), ("eos_stiffness", C.c_float), ("eos_power", C.c_float),
        ("grid_mode", C.c_int32), ("fixed_point_mult", C.c_int32),
        ("stress_form", C.c_int32), ("eq16_order", C.c_int32),
        ("bc_mode", C.c_int32), ("bc_hi_off", C.c_int32), ("bc_friction", C.c_float),
        ("clamp_min", C.c_float), ("clamp_max_off", C.c_float),
        ("wall_min", C.c_float), ("wall_max_off", C.c_float), ("wall_gain", C.c_float),
        ("interaction", C.c_int32), ("sphere_pos", C.c_float * 3), ("sphere_radius", C.c_float),
        ("mouse_pos", C.c_float * 2), ("mouse_radius", C.c_float),
        ("pow_mode", C.c_int32),
        ("n_extra_spheres", C.c_int32), ("extra_spheres", (C.c_float * 4) * 7),
    ]


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = [os.path.join(_HERE, f) for f in ("mpm_oracle.c", "mpm_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        fp = C.POINTER(C.c_float)
        vp = C.c_void_p
        PP = C.POINTER(OrcParams)
        L.orc_init_block.restype = C.c_int64
        L.orc_init_block.argtypes = [C.c_int32, fp, fp, C.c_float, fp, C.c_int64]
        L.orc_clear_grid.argtypes = [PP, vp]
        L.orc_p2g1.argtypes = [PP, C.c_int64, fp, fp, fp, fp, vp]
        L.orc_p2g2.argtypes = [PP, C.c_int64, fp, fp, fp, vp]
        L.orc_update_grid.argtypes = [PP, vp]
        L.orc_g2p.argtypes = [PP, C.c_int64, fp, fp, fp, vp]
        L.orc_step.argtypes = [PP, C.c_int64, fp, fp, fp, fp, vp, C.c_int32]
        L.orc_step_mt.restype = C.c_int32
        L.orc_step_mt.argtypes = [PP, C.c_int64, fp, fp, fp, fp, vp, C.c_int32, C.c_int32]
        L.orc_positions.argtypes = [C.c_int64, fp, fp, fp]
        L.orc_cell_keys.argtypes = [PP, C.c_int64, fp, C.POINTER(C.c_int32)]
        L.orc_stable_sort_perm.argtypes = [C.c_int64, C.POINTER(C.c_uint32), C.POINTER(C.c_int32)]
        L.orc_pow.restype = C.c_float
        L.orc_pow.argtypes = [C.c_int32, C.c_float, C.c_float]
        L.orc_encode_fixed.restype = C.c_int32
        L.orc_encode_fixed.argtypes = [C.c_float, C.c_int32]
        L.orc_decode_fixed.restype = C.c_float
        L.orc_decode_fixed.argtypes = [C.c_int32, C.c_int32]
        _LIB = L
    return _LIB


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def variant(name, grid=None):
    """Parameter block of one of the reference's five solver copies (SURVEY.md 8a 'variant constants')."""
    p = OrcParams()
    p.dt, p.rest_density, p.dynamic_viscosity = 0.2, 4.0, 0.1
    p.fixed_point_mult = 10_000_000
    p.bc_friction, p.sphere_radius, p.mouse_radius = 0.5, 15.0, 10.0
    p.pow_mode = POW_F64_ROUNDED
    if name == "2d_st":      # D
        p.dim, R = 2, 64
        p.gravity, p.eos_stiffness, p.eos_power = 0.3, 10.0, 7.0
        p.grid_mode, p.stress_form, p.eq16_order = GRID_FLOAT, STRESS_2D_TRACE, EQ16_DTVOL_4
        p.bc_mode, p.bc_hi_off = BC_SLIP, 3
        p.clamp_min, p.clamp_max_off = 1.0, 2.0
        p.wall_min, p.wall_max_off, p.wall_gain = 2.0, 3.0, 0.5
    elif name == "2d_mt":    # M
        p.dim, R = 2, 64
        p.gravity, p.eos_stiffness, p.eos_power = 0.3, 10.0, 4.0
        p.grid_mode, p.stress_form, p.eq16_order = GRID_FLOAT, STRESS_2D_TRACE, EQ16_VOL4_DT
        p.bc_mode, p.bc_hi_off = BC_FRICTION, 4
        p.clamp_min, p.clamp_max_off = 1.0, 1.0
        p.wall_min, p.wall_max_off, p.wall_gain = 2.0, 3.0, 0.5
    elif name == "3d_float":  # F
        p.dim, R = 3, 32
        p.gravity, p.eos_stiffness, p.eos_power = -0.3, 10.0, 4.0
        p.grid_mode, p.stress_form, p.eq16_order = GRID_FLOAT, STRESS_3D, EQ16_VOL4_DT
        p.bc_mode, p.bc_hi_off = BC_SLIP, 3
        p.clamp_min, p.clamp_max_off = 1.0, 2.0
        p.wall_min, p.wall_max_off, p.wall_gain = 3.0, 4.0, 1.0
    elif name == "3d_fixed":  # X
        p.dim, R = 3, 32
        p.gravity, p.eos_stiffness, p.eos_power = -0.3, 10.0, 4.0
        p.grid_mode, p.stress_form, p.eq16_order = GRID_FIXED, STRESS_3D, EQ16_VOL4_DT
        p.bc_mode, p.bc_hi_off = BC_SLIP, 3
        p.clamp_min, p.clamp_max_off = 1.0, 2.0
        p.wall_min, p.wall_max_off, p.wall_gain = 3.0, 4.0, 1.0
        p.interaction = INTERACT_SPHERE_POST
        p.sphere_pos[:] = [0.0, 0.0, 31.707275]   # MLSMPM3DFluidMultithreadNew.tscn:40
    elif name == "3d_gpu":    # H + GLSL
        p.dim, R = 3, 64
        p.gravity, p.eos_stiffness, p.eos_power = -0.3, 1.0, 7.0
        p.grid_mode, p.stress_form, p.eq16_order = GRID_FIXED, STRESS_3D, EQ16_VOL4_DT
        p.bc_mode, p.bc_hi_off = BC_SLIP, 3
        p.clamp_min, p.clamp_max_off = 2.0, 2.0
        p.wall_min, p.wall_max_off, p.wall_gain = 3.0, 3.0, 1.0
        p.interaction = INTERACT_SPHERE_PRE
        p.sphere_pos[:] = [-21.648403, 0.0, 31.707275]  # MLSMPM3DFluidMultithreadGPU.tscn:48
    else:
        raise ValueError(name)
    if grid is None:
        grid = (R, R, R if p.dim == 3 else 1)
    if isinstance(grid, int):
        grid = (grid, grid, grid if p.dim == 3 else 1)
    p.grid[:] = list(grid)
    return p


def num_cells(p):
    return p.grid[0] * p.grid[1] * (p.grid[2] if p.dim == 3 else 1)


def init_block(dim, lo, hi, spacing):
    lo3 = np.array(list(lo) + [0.0] * (3 - len(lo)), np.float32)
    hi3 = np.array(list(hi) + [0.0] * (3 - len(hi)), np.float32)
    n = lib().orc_init_block(dim, _fp(lo3), _fp(hi3), np.float32(spacing), None, 0)
    pos = np.zeros((n, 3), np.float32)
    lib().orc_init_block(dim, _fp(lo3), _fp(hi3), np.float32(spacing), _fp(pos), n)
    return pos


class State:
    """Particle + grid state in the oracle's layout."""

    def __init__(self, p, pos, vel=None, Cm=None, mass=None):
        n = pos.shape[0]
        self.p = p
        self.n = n
        self.pos = np.ascontiguousarray(pos, np.float32).copy()
        self.vel = np.zeros((n, 3), np.float32) if vel is None else np.ascontiguousarray(vel, np.float32).copy()
        self.C = np.zeros((n, 9), np.float32) if Cm is None else np.ascontiguousarray(Cm, np.float32).copy()
        self.mass = np.ones(n, np.float32) if mass is None else np.ascontiguousarray(mass, np.float32).copy()
        G = num_cells(p)
        self.grid = np.zeros((G, 4), np.int32)  # raw 32-bit words; view as float32 in float mode

    def grid_f(self):
        return self.grid.view(np.float32)

    def _g(self):
        return self.grid.ctypes.data_as(C.c_void_p)

    def clear_grid(self):
        lib().orc_clear_grid(C.byref(self.p), self._g())

    def p2g1(self):
        lib().orc_p2g1(C.byref(self.p), self.n, _fp(self.pos), _fp(self.vel), _fp(self.C), _fp(self.mass), self._g())

    def p2g2(self):
        lib().orc_p2g2(C.byref(self.p), self.n, _fp(self.pos), _fp(self.C), _fp(self.mass), self._g())

    def update_grid(self):
        lib().orc_update_grid(C.byref(self.p), self._g())

    def g2p(self):
        lib().orc_g2p(C.byref(self.p), self.n, _fp(self.pos), _fp(self.vel), _fp(self.C), self._g())

    def step(self, iterations=1):
        lib().orc_step(C.byref(self.p), self.n, _fp(self.pos), _fp(self.vel), _fp(self.C), _fp(self.mass),
                       self._g(), iterations)

    def step_mt(self, iterations=1, nthreads=0):
        return lib().orc_step_mt(C.byref(self.p), self.n, _fp(self.pos), _fp(self.vel), _fp(self.C),
                                 _fp(self.mass), self._g(), iterations, nthreads)

    def positions(self):
        out = np.zeros((self.n, 4), np.float32)
        lib().orc_positions(self.n, _fp(self.pos), _fp(self.vel), _fp(out))
        return out

    def cell_keys(self):
        k = np.zeros(self.n, np.int32)
        lib().orc_cell_keys(C.byref(self.p), self.n, _fp(self.pos), k.ctypes.data_as(C.POINTER(C.c_int32)))
        return k


def stable_sort_perm(keys):
    keys = np.ascontiguousarray(keys, np.uint32)
    perm = np.zeros(keys.shape[0], np.int32)
    lib().orc_stable_sort_perm(keys.shape[0], keys.ctypes.data_as(C.POINTER(C.c_uint32)),
                               perm.ctypes.data_as(C.POINTER(C.c_int32)))
    return perm
