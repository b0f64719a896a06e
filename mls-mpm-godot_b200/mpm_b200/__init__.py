"""ctypes binding of libmpm_b200.so (the C ABI in include/mpm_b200.h).

This is the Python twin of the C# P/Invoke shim (host/MpmB200.cs): the same exports, the same blittable
structs (MpmParams, the reference's 80-byte Particle and 16-byte Cell).  It holds no solver logic and has
no CPU fallback -- if the shared library or a CUDA device is missing, calls raise MpmError.

`Solver` mirrors the surface of the reference's GPU solver node
(mls-mpm/3d/fluid_multithread_gpu/MLSMPM3DFluidMultithreadGPU.cs): parameters as attributes
(dt, gravity, rest_density, dynamic_viscosity, eos_stiffness, eos_power, sphere_pos), initialise_sim()
(InitialiseSim, :654), process() (_Process, :234: sim_iterations x the five dispatches), positions()
(particle_pos_tex, :196/:342).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MPM_B200_LIB") or os.path.join(os.path.dirname(_HERE), "libmpm_b200.so")  # (override: A/B builds)

OK, ERR_INVALID, ERR_CUDA, ERR_STATE, ERR_OVERFLOW, ERR_COMM, ERR_DOMAIN = 0, 1, 2, 3, 4, 5, 6
GRID_FLOAT, GRID_FIXED = 0, 1
MATH_STRICT, MATH_FAST = 0, 1
PATH_AUTO, PATH_REFERENCE, PATH_TILED, PATH_CELL = 0, 1, 2, 3
VARIANT_2D_ST, VARIANT_2D_MT, VARIANT_3D_FLOAT, VARIANT_3D_FIXED, VARIANT_3D_GPU = range(5)
VARIANTS = {"2d_st": 0, "2d_mt": 1, "3d_float": 2, "3d_fixed": 3, "3d_gpu": 4}
PHASE_CLEAR, PHASE_P2G1, PHASE_P2G2, PHASE_UPDATE, PHASE_G2P, PHASE_SORT = range(6)
COMM_ID_BYTES = 128


class MpmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"mpm_b200 error {code}: {msg}")
        self.code = code


class MpmParams(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("dim", C.c_int32), ("grid_size", C.c_int32 * 3),
        ("dt", C.c_float), ("gravity", C.c_float), ("rest_density", C.c_float),
        ("dynamic_viscosity", C.c_float), ("eos_stiffness", C.c_float), ("eos_power", C.c_float),
        ("grid_mode", C.c_int32), ("fixed_point_mult", C.c_int32), ("stress_form", C.c_int32),
        ("eq16_order", C.c_int32), ("bc_mode", C.c_int32), ("bc_hi_off", C.c_int32),
        ("bc_friction", C.c_float), ("clamp_min", C.c_float), ("clamp_max_off", C.c_float),
        ("wall_min", C.c_float), ("wall_max_off", C.c_float), ("wall_gain", C.c_float),
        ("interaction", C.c_int32), ("sphere_pos", C.c_float * 3), ("sphere_radius", C.c_float),
        ("mouse_pos", C.c_float * 2), ("mouse_radius", C.c_float),
        ("math_mode", C.c_int32), ("kernel_path", C.c_int32), ("sort_interval", C.c_int32),
        ("overflow_check", C.c_int32),
    ]


class MpmStats(C.Structure):
    _fields_ = [
        ("num_particles", C.c_int64), ("num_cells", C.c_int64), ("steps", C.c_int64),
        ("kernel_launches", C.c_int64),
        ("ms_sort", C.c_float), ("ms_clear", C.c_float), ("ms_p2g1", C.c_float), ("ms_p2g2", C.c_float),
        ("ms_update", C.c_float), ("ms_g2p", C.c_float), ("ms_exchange", C.c_float), ("ms_step", C.c_float),
        ("kernel_path", C.c_int32), ("overflow", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32),
        ("local_particles", C.c_int64), ("migrated", C.c_int64), ("slab_jump_clamps", C.c_int64), ("unordered_binnings", C.c_int64), ("far_movers", C.c_int64), ("halo_peer_exchanges", C.c_int64),
        ("ms_halo_mass", C.c_float), ("ms_halo_momentum", C.c_float), ("ms_migration", C.c_float), ("reserved0", C.c_int32),
    ]


# the reference's blittable records (MLSMPM3DFluidMultithreadGPU.cs:8-33)
PARTICLE80 = np.dtype([("pos", "<f4", 3), ("pad_pos", "<f4"), ("vel", "<f4", 3), ("mass", "<f4"),
                       ("C_x", "<f4", 3), ("pad_cx", "<f4"), ("C_y", "<f4", 3), ("pad_cy", "<f4"),
                       ("C_z", "<f4", 3), ("pad_cz", "<f4")])
assert PARTICLE80.itemsize == 80

EXPORTS = [
    "mpm_abi_version", "mpm_device_count", "mpm_default_params", "mpm_create", "mpm_destroy",
    "mpm_last_error", "mpm_set_params", "mpm_get_params", "mpm_set_sphere", "mpm_init_block",
    "mpm_add_block", "mpm_upload_particles", "mpm_upload_particles_soa", "mpm_download_particles",
    "mpm_download_particles_soa", "mpm_download_grid", "mpm_step", "mpm_sync", "mpm_run_phase",
    "mpm_get_positions", "mpm_num_particles", "mpm_set_timing", "mpm_get_stats", "mpm_debug_last_sort",
    "mpm_get_stream", "mpm_host_alloc", "mpm_host_free", "mpm_comm_unique_id", "mpm_comm_init",
    "mpm_local_hub_create", "mpm_local_hub_destroy", "mpm_comm_init_local", "mpm_comm_slab", "mpm_download_ids",
    "mpm_slab_cuts", "mpm_get_positions_async", "mpm_get_positions_q16_async", "mpm_wait_positions", "mpm_comm_rebalance", "mpm_comm_rebalance_weighted", "mpm_set_colliders",
    "mpm_save_state", "mpm_load_state", "mpm_export_positions",
]

_lib = None


def load():
    """Load libmpm_b200.so (built in-tree by `make -C mls-mpm-godot_b200`); raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MpmError(ERR_STATE, f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    fp, PP = C.POINTER(C.c_float), C.POINTER(MpmParams)
    sig = {
        "mpm_abi_version": (i32, []),
        "mpm_device_count": (i32, []),
        "mpm_default_params": (i32, [i32, PP]),
        "mpm_create": (i32, [PP, i64, i32, C.POINTER(vp)]),
        "mpm_destroy": (i32, [vp]),
        "mpm_last_error": (C.c_char_p, [vp]),
        "mpm_set_params": (i32, [vp, PP]),
        "mpm_get_params": (i32, [vp, PP]),
        "mpm_set_sphere": (i32, [vp, fp]),
        "mpm_init_block": (i32, [vp, fp, fp, C.c_float]),
        "mpm_add_block": (i32, [vp, fp, fp, C.c_float]),
        "mpm_upload_particles": (i32, [vp, vp, i64]),
        "mpm_upload_particles_soa": (i32, [vp, fp, fp, fp, fp, i64]),
        "mpm_download_particles": (i32, [vp, vp, i64]),
        "mpm_download_particles_soa": (i32, [vp, fp, fp, fp, fp, i64]),
        "mpm_download_grid": (i32, [vp, vp, i64]),
        "mpm_step": (i32, [vp, i32]),
        "mpm_sync": (i32, [vp]),
        "mpm_run_phase": (i32, [vp, i32]),
        "mpm_get_positions": (i32, [vp, vp, i64, C.POINTER(vp), C.POINTER(C.c_uint32)]),
        "mpm_num_particles": (i32, [vp, C.POINTER(i64)]),
        "mpm_set_timing": (i32, [vp, i32]),
        "mpm_get_stats": (i32, [vp, C.POINTER(MpmStats)]),
        "mpm_debug_last_sort": (i32, [vp, vp, vp, i64]),
        "mpm_get_stream": (i32, [vp, C.POINTER(vp)]),
        "mpm_host_alloc": (i32, [i64, C.POINTER(vp)]),
        "mpm_host_free": (i32, [vp]),
        "mpm_comm_unique_id": (i32, [vp]),
        "mpm_comm_init": (i32, [vp, vp, i32, i32]),
        "mpm_local_hub_create": (i32, [i32, C.POINTER(vp)]),
        "mpm_local_hub_destroy": (i32, [vp]),
        "mpm_comm_init_local": (i32, [vp, vp, i32, i32]),
        "mpm_comm_slab": (i32, [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]),
        "mpm_download_ids": (i32, [vp, vp, i64]),
        "mpm_slab_cuts": (i32, [C.POINTER(i64), i32, i32, i32, C.POINTER(i32)]),
        "mpm_get_positions_async": (i32, [vp, vp, i64]),
        "mpm_get_positions_q16_async": (i32, [vp, vp, i64]),
        "mpm_wait_positions": (i32, [vp]),
        "mpm_comm_rebalance": (i32, [vp, i32]),
        "mpm_comm_rebalance_weighted": (i32, [vp, i32, C.c_float]),
        "mpm_set_colliders": (i32, [vp, fp, i32]),
        "mpm_save_state": (i32, [vp, C.c_char_p]),
        "mpm_load_state": (i32, [vp, C.c_char_p]),
        "mpm_export_positions": (i32, [vp, C.POINTER(i32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]),
    }
    for name, (res, args) in sig.items():
        try:
            fn = getattr(L, name)
        except AttributeError:
            if os.environ.get("MPM_B200_LIB"):  # an older build selected for an A/B measurement may lack newer entry points
                continue
            raise
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


def default_params(variant, grid=None, **overrides):
    """MpmParams of one of the reference's five solver copies; `grid` (int or 3-tuple) overrides its size."""
    L = load()
    p = MpmParams()
    v = VARIANTS[variant] if isinstance(variant, str) else variant
    rc = L.mpm_default_params(v, C.byref(p))
    if rc:
        raise MpmError(rc, "bad variant")
    if grid is not None:
        if isinstance(grid, int):
            grid = (grid, grid, grid if p.dim == 3 else 1)
        p.grid_size[:] = list(grid)
    for k, val in overrides.items():
        if k in ("sphere_pos", "mouse_pos"):
            getattr(p, k)[:] = list(val)
        else:
            setattr(p, k, val)
    return p


def _fp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


class Solver:
    """Python mirror of the reference's solver node surface over the C ABI."""

    def __init__(self, params, max_particles, device=0):
        self._L = load()
        self._h = C.c_void_p()
        self.params = params
        rc = self._L.mpm_create(C.byref(params), int(max_particles), int(device), C.byref(self._h))
        if rc:
            raise MpmError(rc, (self._L.mpm_last_error(None) or b"").decode())
        self.sim_iterations = 2  # MLSMPM3DFluidMultithreadGPU.cs:69

    # -- plumbing
    def _ck(self, rc):
        if rc:
            raise MpmError(rc, (self._L.mpm_last_error(self._h) or b"").decode())

    def close(self):
        if self._h:
            self._L.mpm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- parameters (UpdatePushConstants, :444-503)
    def update_push_constants(self):
        self._ck(self._L.mpm_set_params(self._h, C.byref(self.params)))

    def set_sphere(self, pos):
        a = (C.c_float * 3)(*pos)
        self._ck(self._L.mpm_set_sphere(self._h, a))

    def save_state(self, path):
        self._ck(self._L.mpm_save_state(self._h, os.fsencode(path)))

    def load_state(self, path):
        self._ck(self._L.mpm_load_state(self._h, os.fsencode(path)))

    def set_colliders(self, spheres):
        """Further sphere repulsors [(x, y, z, r), ...] (at most 7) applied after `sphere_pos`."""
        a = np.ascontiguousarray(spheres, np.float32).reshape(-1, 4)
        self._ck(self._L.mpm_set_colliders(self._h, _fp(a), a.shape[0]))

    # -- scene (InitialiseSim, :654-707)
    def initialise_sim(self, lo, hi, spacing, append=False):
        lo3 = (C.c_float * 3)(*(list(lo) + [0.0] * (3 - len(lo))))
        hi3 = (C.c_float * 3)(*(list(hi) + [0.0] * (3 - len(hi))))
        fn = self._L.mpm_add_block if append else self._L.mpm_init_block
        self._ck(fn(self._h, lo3, hi3, C.c_float(spacing)))
        return self.num_particles

    def upload(self, pos, vel=None, Cm=None, mass=None):
        arrs = [None if a is None else np.ascontiguousarray(a, np.float32) for a in (pos, vel, Cm, mass)]
        n = arrs[0].shape[0]
        self._ck(self._L.mpm_upload_particles_soa(self._h, _fp(arrs[0]), _fp(arrs[1]), _fp(arrs[2]), _fp(arrs[3]), n))

    def upload_aos80(self, rec):
        rec = np.ascontiguousarray(rec, PARTICLE80)
        self._ck(self._L.mpm_upload_particles(self._h, rec.ctypes.data_as(C.c_void_p), rec.shape[0]))

    def download(self):
        n = self.stats().local_particles  # (multi-GPU: triggers the pending slab partition first)
        pos, vel = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
        Cm, mass = np.zeros((n, 9), np.float32), np.zeros(n, np.float32)
        self._ck(self._L.mpm_download_particles_soa(self._h, _fp(pos), _fp(vel), _fp(Cm), _fp(mass), n))
        return pos, vel, Cm, mass

    def download_aos80(self):
        n = self.num_particles
        rec = np.zeros(n, PARTICLE80)
        self._ck(self._L.mpm_download_particles(self._h, rec.ctypes.data_as(C.c_void_p), n))
        return rec

    def download_grid(self):
        st = self.stats()
        g = np.zeros((st.num_cells, 4), np.int32)
        self._ck(self._L.mpm_download_grid(self._h, g.ctypes.data_as(C.c_void_p), st.num_cells))
        return g

    # -- stepping (_Process, :234-251)
    def step(self, iterations=1):
        self._ck(self._L.mpm_step(self._h, int(iterations)))

    def process(self):
        self.step(self.sim_iterations)

    def sync(self):
        self._ck(self._L.mpm_sync(self._h))

    def run_phase(self, phase):
        self._ck(self._L.mpm_run_phase(self._h, int(phase)))

    def positions(self, out=None):
        """(x, y, z, |v|) per particle, original index order (particle_pos_tex)."""
        n = self.num_particles
        if out is None:
            out = np.zeros((n, 4), np.float32)
        w = C.c_uint32()
        self._ck(self._L.mpm_get_positions(self._h, out.ctypes.data_as(C.c_void_p), n, None, C.byref(w)))
        return out

    def positions_into(self, host_ptr, cap):
        self._ck(self._L.mpm_get_positions(self._h, C.c_void_p(host_ptr), int(cap), None, None))

    def positions_into_async(self, host_ptr, cap):
        """Pipelined hand-off: returns at once; the pinned buffer is complete after wait_positions()."""
        self._ck(self._L.mpm_get_positions_async(self._h, C.c_void_p(host_ptr), int(cap)))

    def positions_q16_into_async(self, host_ptr, cap):
        """The same at 8 bytes per particle: 4 x uint16 = x, y, z as fractions of the domain (code * grid_size / 65535), |v| as binary16."""
        self._ck(self._L.mpm_get_positions_q16_async(self._h, C.c_void_p(host_ptr), int(cap)))

    def wait_positions(self):
        self._ck(self._L.mpm_wait_positions(self._h))

    def export_positions(self):
        """(fd, bytes, tex_width): a POSIX file descriptor of the allocation that holds the (x, y, z, |v|) array, for a
        renderer / another process to import (zero-copy hand-off); the caller closes the fd."""
        fd, nbytes, w = C.c_int32(-1), C.c_uint64(0), C.c_uint32(0)
        self._ck(self._L.mpm_export_positions(self._h, C.byref(fd), C.byref(nbytes), C.byref(w)))
        return fd.value, nbytes.value, w.value

    def refresh_positions(self):
        """Bring the device-side (x, y, z, |v|) array up to date (no host copy)."""
        self._ck(self._L.mpm_get_positions(self._h, None, 0, None, None))

    def positions_device(self):
        dp, w = C.c_void_p(), C.c_uint32()
        self._ck(self._L.mpm_get_positions(self._h, None, 0, C.byref(dp), C.byref(w)))
        return dp.value, w.value

    @property
    def num_particles(self):
        n = C.c_int64()
        self._ck(self._L.mpm_num_particles(self._h, C.byref(n)))
        return n.value

    def set_timing(self, on=True):
        self._ck(self._L.mpm_set_timing(self._h, int(on)))  # True / 1: per phase, 2: whole call only

    def stats(self):
        st = MpmStats()
        self._ck(self._L.mpm_get_stats(self._h, C.byref(st)))
        return st

    def last_sort(self):
        n = self.num_particles
        keys, perm = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        self._ck(self._L.mpm_debug_last_sort(self._h, keys.ctypes.data_as(C.c_void_p), perm.ctypes.data_as(C.c_void_p), n))
        return keys, perm

    def record_ids(self):
        """Original index of the particle in each record, in the order `last_sort()` reports its keys in."""
        n = self.num_particles
        ids = np.zeros(n, np.uint32)
        self._ck(self._L.mpm_download_ids(self._h, ids.ctypes.data_as(C.c_void_p), n))
        return ids

    def stream(self):
        sp = C.c_void_p()
        self._ck(self._L.mpm_get_stream(self._h, C.byref(sp)))
        return sp.value or 0

    # -- multi-GPU x-slabs (no reference counterpart)
    def comm_init(self, unique_id: bytes, rank, world):
        """NCCL transport: one process per GPU; unique_id comes from comm_unique_id() on rank 0."""
        buf = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(unique_id)
        self._ck(self._L.mpm_comm_init(self._h, buf, rank, world))

    def comm_init_local(self, hub, rank, world):
        """LOCAL transport: k solvers of one process (one host thread each) share `hub`."""
        self._hub = hub  # keep it alive as long as the solver
        self._ck(self._L.mpm_comm_init_local(self._h, hub._h, rank, world))

    def comm_rebalance(self, max_shift=2, cost_per_particle=None):
        """Re-cut the slabs towards equal particle counts -- or, with this rank's measured cost per particle (a unit all
        ranks share, numbers around 1), towards equal summed cost (collective; between steps)."""
        if cost_per_particle is None:
            self._ck(self._L.mpm_comm_rebalance(self._h, int(max_shift)))
        else:
            self._ck(self._L.mpm_comm_rebalance_weighted(self._h, int(max_shift), C.c_float(cost_per_particle)))

    def slab(self):
        """(x0, x1, gx0, nxl): owned planes [x0, x1), stored planes [gx0, gx0 + nxl)."""
        v = [C.c_int32() for _ in range(4)]
        self._ck(self._L.mpm_comm_slab(self._h, *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)

    def download_ids(self):
        n = self.stats().local_particles
        ids = np.zeros(n, np.uint32)
        self._ck(self._L.mpm_download_ids(self._h, ids.ctypes.data_as(C.c_void_p), n))
        return ids


class LocalHub:
    """Rendezvous object of the LOCAL transport (mpm_local_hub_create)."""

    def __init__(self, world):
        self._L = load()
        self._h = C.c_void_p()
        rc = self._L.mpm_local_hub_create(int(world), C.byref(self._h))
        if rc:
            raise MpmError(rc, "mpm_local_hub_create failed")
        self.world = world

    def close(self):
        if self._h:
            self._L.mpm_local_hub_destroy(self._h)
            self._h = C.c_void_p()


def slab_cuts(hist, world, min_width=4):
    """Equal-count x-slab cuts from an x-plane particle histogram (host-only; mpm_slab_cuts)."""
    L = load()
    h = np.ascontiguousarray(hist, np.int64)
    cuts = (C.c_int32 * (world + 1))()
    rc = L.mpm_slab_cuts(h.ctypes.data_as(C.POINTER(C.c_int64)), h.shape[0], int(world), int(min_width), cuts)
    if rc:
        raise MpmError(rc, "mpm_slab_cuts: invalid arguments (grid too narrow for this many ranks?)")
    return list(cuts)


def comm_unique_id():
    L = load()
    buf = (C.c_uint8 * COMM_ID_BYTES)()
    rc = L.mpm_comm_unique_id(buf)
    if rc:
        raise MpmError(rc, "mpm_comm_unique_id failed")
    return bytes(buf)


def host_alloc(nbytes):
    L = load()
    p = C.c_void_p()
    rc = L.mpm_host_alloc(int(nbytes), C.byref(p))
    if rc:
        raise MpmError(rc, "mpm_host_alloc failed")
    return p.value


def host_free(ptr):
    load().mpm_host_free(C.c_void_p(ptr))
