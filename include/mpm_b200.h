/*
 * mpm_b200.h -- C ABI of libmpm_b200.so: a B200-native (sm_100a) MLS-MPM fluid solver that sits behind
 * the solver surface of Miotismon/mls-mpm-godot.
 *
 * The reference has no FFI/plugin layer; its "API" is the fields and methods of the GPU solver node
 *   H = mls-mpm/3d/fluid_multithread_gpu/MLSMPM3DFluidMultithreadGPU.cs
 * (and the CPU copies F = 3d/fluid_multithread/..., X = 3d/fluid_multithread_fixed_point/...,
 * D = 2d/fluid/..., M = 2d/fluid_multithread/...).  Every entry point below names the reference member
 * it replaces.  A C# host binds these with [DllImport("mpm_b200", CallingConvention = Cdecl)]; see
 * INTEGRATION.md and mls-mpm-godot_b200/host/MpmB200.cs.
 *
 * Conventions: plain C, cdecl, no exceptions cross the boundary.  Every call returns an int32 status
 * (MPM_OK = 0); mpm_last_error() gives a UTF-8 message owned by the library.  Host buffers are
 * caller-owned and only touched during the call.  Device memory is library-owned.  A handle is used from
 * one thread at a time (the reference calls everything from Godot's main thread).  mpm_step() is
 * asynchronous with respect to the GPU; mpm_sync() or any download waits for it.
 * There is no CPU fallback: without a CUDA device every call that needs one fails with MPM_ERR_CUDA.
 */
#ifndef MPM_B200_H
#define MPM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define MPM_API __declspec(dllexport)
#else
#define MPM_API __attribute__((visibility("default")))
#endif

#define MPM_ABI_VERSION 1

/* ---- status codes ---- */
#define MPM_OK 0
#define MPM_ERR_INVALID 1  /* bad argument / parameter combination */
#define MPM_ERR_CUDA 2     /* CUDA runtime error (message has the CUDA string) */
#define MPM_ERR_STATE 3    /* call not valid in the current state (e.g. step before upload) */
#define MPM_ERR_OVERFLOW 4 /* fixed-point accumulator left the int32 range (debug detector) */
#define MPM_ERR_COMM 5     /* multi-GPU exchange failed */
#define MPM_ERR_DOMAIN 6   /* particle positions that are non-finite or whose 3x3(x3) stencil leaves the grid, i.e. outside
                              [1, R-1) on some axis: rejected at upload / load; if a running simulation produces one
                              (NaN from a blown-up step) the particle is skipped and mpm_sync reports it.  The reference
                              indexes the grid with (int)pos unchecked and dies with IndexOutOfRangeException (F:281-283) */

/* ---- enums (int32 in the struct) ---- */
#define MPM_GRID_FLOAT 0 /* Cell{Vector3 vel; float mass}            F:16-20, D:16-20 */
#define MPM_GRID_FIXED 1 /* Cell{int vel_x, vel_y, vel_z, mass} x1e7 X:18-24, H:25-32 */

#define MPM_STRESS_3D 0       /* strain = C + C^T                     F:340-345, p2g_2.glsl:108-112 */
#define MPM_STRESS_2D_TRACE 1 /* off-diagonals summed, diagonal kept  D:276-283 */

#define MPM_EQ16_VOL4_DT 0 /* ((-volume*4)*stress)*dt  F:347, X:407, M:307, p2g_2.glsl:115 */
#define MPM_EQ16_DTVOL_4 1 /* ((-dt*volume)*stress)*4  D:285 */

#define MPM_BC_SLIP 0     /* F:402-404, X:479-481, update_grid.glsl:64-66 */
#define MPM_BC_FRICTION 1 /* M:366-368 */

#define MPM_INTERACT_NONE 0
#define MPM_INTERACT_SPHERE_POST 1 /* X:570-576 */
#define MPM_INTERACT_SPHERE_PRE 2  /* g2p.glsl:122-129 */
#define MPM_INTERACT_MOUSE_2D 3    /* D:381-406 (mouse held down) */

/* arithmetic mode */
#define MPM_MATH_STRICT 0 /* every fp32 op in the C# source order, no FMA: with MPM_GRID_FIXED the grid
                             ints and particle floats are bit-identical to the reference algorithm */
#define MPM_MATH_FAST 1   /* FMA contraction + per-axis hoisting; results within float tolerance */

/* kernel path */
#define MPM_PATH_AUTO 0
#define MPM_PATH_REFERENCE 1 /* one thread per particle, global atomics: the shape of p2g_1.glsl etc. */
#define MPM_PATH_TILED 2     /* block-binned particles (stable radix sort), one thread per particle, shared-memory grid tiles;
                                runs MPM_MATH_STRICT bit-exactly (and MPM_MATH_FAST) */
#define MPM_PATH_CELL 3      /* cell-binned particles (counting sort), one thread per grid cell with the 27-node stencil
                                in registers: the fast B200 path; MPM_MATH_FAST only.  AUTO picks it for 3D fixed + FAST */

/* variant presets for mpm_default_params() */
#define MPM_VARIANT_2D_ST 0    /* D */
#define MPM_VARIANT_2D_MT 1    /* M */
#define MPM_VARIANT_3D_FLOAT 2 /* F */
#define MPM_VARIANT_3D_FIXED 3 /* X */
#define MPM_VARIANT_3D_GPU 4   /* H + GLSL: the shipping scene */

/*
 * Parameter block.  It is the union of the reference's push-constant blocks (H:444-503:
 * {fixed_point_mult, grid_size}, {.., dt, rest_density, dynamic_viscosity, eos_stiffness, eos_power},
 * {.., dt, gravity}, {.., dt, sphere_pos, tex_width}) plus the compile-time constants in which the five
 * solver copies differ (SURVEY.md 8a "variant constants").  All members are 4 bytes; no padding.
 */
typedef struct MpmParams {
    int32_t struct_size;      /* = sizeof(MpmParams); checked */
    int32_t dim;              /* 2 or 3 */
    int32_t grid_size[3];     /* Rx, Ry, Rz; the reference is cubic (H:43), this is a superset */
    float dt;                 /* H:57-67: clamped to [0, 0.4] like the Dt setter */
    float gravity;            /* H:71, applied on y (F:396, D:319) */
    float rest_density;       /* H:76 */
    float dynamic_viscosity;  /* H:78 */
    float eos_stiffness;      /* H:82 */
    float eos_power;          /* H:84 */
    int32_t grid_mode;        /* MPM_GRID_* */
    int32_t fixed_point_mult; /* H:98 */
    int32_t stress_form;      /* MPM_STRESS_* */
    int32_t eq16_order;       /* MPM_EQ16_* */
    int32_t bc_mode;          /* MPM_BC_* */
    int32_t bc_hi_off;        /* BC where idx < 2 || idx > R - bc_hi_off */
    float bc_friction;        /* M:366 */
    float clamp_min;          /* F:476 = 1, g2p.glsl:115 = 2 */
    float clamp_max_off;      /* clamp max = R - clamp_max_off */
    float wall_min;           /* F:507 */
    float wall_max_off;       /* wall max = R - wall_max_off */
    float wall_gain;          /* F:509 = 1, D:410 = 0.5 */
    int32_t interaction;      /* MPM_INTERACT_* */
    float sphere_pos[3];      /* H:93, patched per frame by HandleMouseInteraction H:618-642 */
    float sphere_radius;      /* H:94 */
    float mouse_pos[2];       /* D:54 */
    float mouse_radius;       /* D:52 */
    /* ---- B200 solver controls (no reference counterpart) ---- */
    int32_t math_mode;        /* MPM_MATH_* */
    int32_t kernel_path;      /* MPM_PATH_* */
    int32_t sort_interval;    /* re-bin particles every k steps (tiled path); 0 = library default */
    int32_t overflow_check;   /* 1: debug detector for the int32 x fixed_point_mult grid: every encode (saturation / NaN) and every
                                 integer add (wrap-around) is checked, mpm_sync returns MPM_ERR_OVERFLOW and MpmStats.overflow is
                                 set.  Slower (atomics return values); runs on the reference-shaped and tiled paths -- AUTO avoids the
                                 cell path when it is set, MPM_PATH_CELL with it is rejected.  Values are identical either way:
                                 bit-exact results hold for in-range data (|value| < 2^31 / fixed_point_mult = 214.7 at 1e7) */
} MpmParams;

/* Particle record of the reference's GPU buffer: 80 bytes, std430 (H:8-22, p2g_1.glsl:4-9). */
typedef struct MpmParticle80 {
    float pos[3];
    float pad_pos;
    float vel[3];
    float mass;
    float C_x[3]; /* column 0 of C (Basis.X / GLSL C[0]) */
    float pad_cx;
    float C_y[3];
    float pad_cy;
    float C_z[3];
    float pad_cz;
} MpmParticle80;

/* Grid cell of the reference (H:25-32): four 32-bit words (vel_x, vel_y, vel_z, mass); int32 in
 * MPM_GRID_FIXED, float in MPM_GRID_FLOAT (the float copies order it {vel, mass} too: F:16-20). */
typedef struct MpmCell16 {
    int32_t w[4];
} MpmCell16;

typedef struct MpmStats {
    int64_t num_particles;
    int64_t num_cells;
    int64_t steps;            /* steps executed since create */
    int64_t kernel_launches;  /* this library's kernels launched since create */
    /* average device time per step of each phase over the last mpm_step() call that had timing on (ms) */
    float ms_sort, ms_clear, ms_p2g1, ms_p2g2, ms_update, ms_g2p, ms_exchange, ms_step;
    int32_t kernel_path;      /* the path actually used (MPM_PATH_*) */
    int32_t overflow;         /* sticky overflow flag */
    int32_t rank, world;
    int64_t local_particles;  /* particles owned by this rank (multi-GPU) */
    int64_t migrated;         /* particles this rank has sent to its neighbours since the communicator was attached */
    int64_t slab_jump_clamps; /* multi-GPU: particles that would have crossed more than one slab in a single step (|v| dt
                                 larger than the neighbouring slab is wide) and were held back in that slab's far plane */
    int64_t unordered_binnings; /* cell path: bin phases that left cells in atomic order instead of the stable one: more than
                                   131072 particles jumped out of their grid block's one-cell apron in that step (bulk motion of
                                   more than a cell per step), or one cell received more than 32 such particles.  0 means every
                                   binning so far equals std::stable_sort */
    int64_t far_movers;         /* cell path: particles of the most recent bin phase that had left their grid block's one-cell
                                   apron (put in order by the fix-up pass; a lower bound once a binning has given up) */
    int64_t halo_peer_exchanges; /* multi-GPU: halo exchanges done by direct peer stores (k_halo_push / k_halo_wait_add over
                                   CUDA IPC or in-process peer pointers) rather than through the transport */
    /* multi-GPU: ms_exchange split into its three parts (same averaging): the mass halo after P2G_1, the momentum halo after
       P2G_2, the particle migration after G2P.  Each includes the wait for the neighbours. */
    float ms_halo_mass, ms_halo_momentum, ms_migration;
    int32_t reserved0;
} MpmStats;

typedef struct MpmSolver MpmSolver; /* opaque */

/* Version / capability probe; usable without a GPU. */
MPM_API int32_t mpm_abi_version(void);
/* Number of visible CUDA devices (0 without a driver); never fails. */
MPM_API int32_t mpm_device_count(void);

/* Fill *p with the constants of one of the reference's five solver copies (variant constants table). */
MPM_API int32_t mpm_default_params(int32_t variant, MpmParams* p);

/* _Ready + InitGPU (H:158-207, 265-435): allocate device state for up to max_particles on `device`. */
MPM_API int32_t mpm_create(const MpmParams* p, int64_t max_particles, int32_t device, MpmSolver** out);
/* CleanupGpu on NotificationPredelete (H:546-616, 709-715). */
MPM_API int32_t mpm_destroy(MpmSolver* s);
MPM_API const char* mpm_last_error(const MpmSolver* s); /* s may be NULL: error of the last failed create */

/* UpdatePushConstants (H:444-503) and the UI setters (main_ui.tscn:70-72).  Grid size, dim and grid_mode
 * are fixed at create time. */
MPM_API int32_t mpm_set_params(MpmSolver* s, const MpmParams* p);
MPM_API int32_t mpm_get_params(const MpmSolver* s, MpmParams* p);
/* HandleMouseInteraction (H:618-642): patch only the sphere position. */
MPM_API int32_t mpm_set_sphere(MpmSolver* s, const float pos[3]);

/* Sphere list (colliders beyond the single hard-coded sphere, g2p.glsl:122-129 / X:570-576): up to 7 further
 * spheres (x, y, z, radius each), applied after MpmParams.sphere_pos in order with the same test and push, on the
 * pre- or post-advection position as MpmParams.interaction says.  count = 0 removes them. */
#define MPM_MAX_EXTRA_SPHERES 7
MPM_API int32_t mpm_set_colliders(MpmSolver* s, const float* xyzr, int32_t count);

/* InitialiseSim (H:654-707, F:129-183): lattice of points in [lo, hi) with the reference's
 * float-accumulating loops, vel = 0, C = 0, mass = 1, grid zeroed.  Replaces the particle set. */
MPM_API int32_t mpm_init_block(MpmSolver* s, const float lo[3], const float hi[3], float spacing);
/* As above but appended to the current set (multi-block scenes). */
MPM_API int32_t mpm_add_block(MpmSolver* s, const float lo[3], const float hi[3], float spacing);

/* StorageBufferCreate(particle_bytes) (H:293-322): upload n particles in the reference's 80-byte layout. */
MPM_API int32_t mpm_upload_particles(MpmSolver* s, const MpmParticle80* ps, int64_t n);
/* SoA overload: pos[3n], vel[3n], C[9n] column-major, mass[n]; vel/C/mass may be NULL (0, 0, 1). */
MPM_API int32_t mpm_upload_particles_soa(MpmSolver* s, const float* pos, const float* vel, const float* C,
                                         const float* mass, int64_t n);
/* BufferGetData(particle_buffer) (H:210-228, commented out in the reference): particles in ORIGINAL
 * index order (particle i keeps index i), whatever the internal binning. */
MPM_API int32_t mpm_download_particles(MpmSolver* s, MpmParticle80* ps, int64_t cap);
MPM_API int32_t mpm_download_particles_soa(MpmSolver* s, float* pos, float* vel, float* C, float* mass,
                                           int64_t cap);
/* BufferGetData(grid_buffer): cells in reference order x*Ry*Rz + y*Rz + z (F:282), field order of H:25-32. */
MPM_API int32_t mpm_download_grid(MpmSolver* s, MpmCell16* cells, int64_t cap);

/* Checkpoint / resume (the reference has none: state is re-created in _Ready).  Raw little-endian
 * file: 64-byte header {"MPMB200\0", u32 version = 1, i32 dim, i32 grid[3], i64 n, i64 steps, pad} followed by n
 * particle records in the reference's 80-byte layout (H:8-22), original index order.  Loading replaces the particle
 * set (like mpm_upload_particles) and restores the step counter; parameters are not stored (the host owns them). */
MPM_API int32_t mpm_save_state(MpmSolver* s, const char* path);
MPM_API int32_t mpm_load_state(MpmSolver* s, const char* path);

/* _Process -> sim_iterations x SetComputeLists (H:234-251, 505-544): enqueue `iterations` steps
 * (clear, P2G_1, P2G_2, update, G2P) on the solver's stream and return. */
MPM_API int32_t mpm_step(MpmSolver* s, int32_t iterations);
MPM_API int32_t mpm_sync(MpmSolver* s);

/* The five phases of Simulate() one by one (F:185-220), for per-kernel parity tests and profiling.
 * phase: 0 clear, 1 p2g_1, 2 p2g_2, 3 update_grid, 4 g2p, 5 bin/sort (no reference counterpart). */
MPM_API int32_t mpm_run_phase(MpmSolver* s, int32_t phase);

/* particle_pos_tex (H:196, 342-355; g2p.glsl:149-150): float4 (x, y, z, |v|) per particle in original
 * index order; texel (i % width, i / width) with width = (uint)sqrt(N) + 1 is element i of this array.
 * dst may be NULL to only query.  *device_ptr (optional) receives the device address of the array so a
 * renderer can consume it without a host round trip. */
MPM_API int32_t mpm_get_positions(MpmSolver* s, float* dst4, int64_t cap, void** device_ptr,
                                  uint32_t* tex_width);

/* Zero-copy hand-off (H:340-355, 402-412: the reference's positions never leave the GPU -- G2P writes a storage image the
 * MultiMesh shader samples).  The (x, y, z, |v|) array lives in an allocation made with cuMemCreate; *fd receives a new POSIX
 * file descriptor for it (the caller owns and closes it) that a renderer imports once -- Vulkan: VkImportMemoryFdInfoKHR with
 * VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT; CUDA in another process: cuMemImportFromShareableHandle /
 * cudaImportExternalMemory -- *bytes the allocation's size (the array starts at offset 0, 16 bytes per particle, index i at
 * texel (i % width, i / width)).  After every step, mpm_get_positions(s, NULL, 0, NULL, NULL) refreshes the array on the
 * solver's stream and mpm_sync() makes it visible: no byte crosses PCIe.  MPM_ERR_STATE if the driver cannot export. */
MPM_API int32_t mpm_export_positions(MpmSolver* s, int32_t* fd, uint64_t* bytes, uint32_t* tex_width);

/* Pipelined form of the hand-off for hosts that render while the next step runs (double buffering): enqueues the
 * hand-off array and its copy into dst4 (pinned host memory from mpm_host_alloc) on a separate copy stream and returns
 * at once, so the device-to-host transfer overlaps the following mpm_step.  dst4 is complete after
 * mpm_wait_positions(); the caller alternates between two host buffers.  Two device arrays are kept internally. */
MPM_API int32_t mpm_get_positions_async(MpmSolver* s, float* dst4, int64_t cap);
MPM_API int32_t mpm_wait_positions(MpmSolver* s);
/* The same pipelined hand-off at half the bytes for hosts that must copy (the transfer over PCIe is what bounds them: 16 bytes
 * per particle at ~56 GB/s): 4 x uint16 per particle -- x, y, z as fractions of the domain, code = rint(p / grid_size * 65535)
 * (decode p = code * grid_size / 65535: a step of 0.004 cell on a 256-cell axis), and |v| as an IEEE binary16.  A renderer
 * uploads it as an RGBA16 texture.  No reference counterpart (the reference never copies). */
MPM_API int32_t mpm_get_positions_q16_async(MpmSolver* s, uint16_t* dst4, int64_t cap);

MPM_API int32_t mpm_num_particles(const MpmSolver* s, int64_t* n);
/* Per-phase timing (Time.GetTicksUsec around each phase, F:190-219) is off by default.  enabled = 1: CUDA events around every
 * phase of every step (MpmStats.ms_sort ... ms_step); enabled = 2: two events around the whole mpm_step() call only (ms_step;
 * the per-phase fields read 0) -- the 12 to 18 event records of a step cost a few percent of a sub-millisecond step. */
MPM_API int32_t mpm_set_timing(MpmSolver* s, int32_t enabled);
MPM_API int32_t mpm_get_stats(MpmSolver* s, MpmStats* st);

/* Binning introspection (tiled path): the cell keys and the permutation (sorted rank -> index before the
 * sort) of the most recent bin phase, for the bit-exact sort parity test. */
MPM_API int32_t mpm_debug_last_sort(MpmSolver* s, uint32_t* keys_before, uint32_t* perm, int64_t cap);

/* The CUDA stream the solver enqueues on (cudaStream_t as void*), so callers can time with events. */
MPM_API int32_t mpm_get_stream(MpmSolver* s, void** stream);

/* Pinned host memory for fast mpm_get_positions()/downloads (a C# host keeps it as IntPtr). */
MPM_API int32_t mpm_host_alloc(int64_t bytes, void** out);
MPM_API int32_t mpm_host_free(void* p);

/* ---- multi-GPU (no reference counterpart: the reference is single-device) -----------------------------------
 * x-slab decomposition (contiguous in the reference's cell order x*Ry*Rz + y*Rz + z, F:282).  A rank owns the
 * planes [x0, x1) and the particles whose base cell x lies there, stores [x0-1, x1+1).  Per step: exchange-add of
 * the two overlap planes with each neighbour after P2G_1 and after P2G_2, redundant grid update on the ghost
 * planes, particle migration after G2P.  int32 adds commute, so in MPM_GRID_FIXED the k-rank result is
 * bit-identical to the 1-rank result.  Only MPM_GRID_FIXED, dim = 3 is supported with a communicator.
 *
 * Two transports:
 *   NCCL  - one process per GPU (torchrun / mpirun): rank 0 calls mpm_comm_unique_id, the host broadcasts the
 *           128 bytes by any means, every rank calls mpm_comm_init.  libnccl.so.2 is resolved with dlopen at that
 *           moment (the copy already loaded in the process, e.g. torch's, else the system one).
 *   LOCAL - k solvers inside one process (one host thread per solver, any mix of devices, peer copies +
 *           events): mpm_local_hub_create, then mpm_comm_init_local on each.  Every rank must be inside
 *           mpm_step / an upload concurrently, as with NCCL.
 * After either init, mpm_init_block / mpm_upload_particles* take the GLOBAL particle set on every rank (the
 * same data everywhere); slab cuts are chosen from its x-plane histogram (equal counts) and each rank keeps
 * its own slab.  Downloads then return the rank's LOCAL particles in slot order; mpm_download_ids gives their
 * global indices. */
#define MPM_COMM_ID_BYTES 128
MPM_API int32_t mpm_comm_unique_id(uint8_t id[MPM_COMM_ID_BYTES]);
MPM_API int32_t mpm_comm_init(MpmSolver* s, const uint8_t id[MPM_COMM_ID_BYTES], int32_t rank, int32_t world);

typedef struct MpmLocalHub MpmLocalHub; /* opaque rendezvous object shared by the k solvers of one process */
MPM_API int32_t mpm_local_hub_create(int32_t world, MpmLocalHub** hub);
MPM_API int32_t mpm_local_hub_destroy(MpmLocalHub* hub); /* after every attached solver is destroyed */
MPM_API int32_t mpm_comm_init_local(MpmSolver* s, MpmLocalHub* hub, int32_t rank, int32_t world);

/* Load balance for long runs (fluid flows from slab to slab): collective over all ranks, between steps.  Builds the
 * global x-plane particle histogram, moves every cut towards its equal-count position by at most max_shift planes
 * (1..3), migrates the particles that changed owner and re-creates the rank's local grid.  Results are unaffected
 * (ownership is not physics): bit-identical in MPM_GRID_FIXED. */
MPM_API int32_t mpm_comm_rebalance(MpmSolver* s, int32_t max_shift);
/* The same with the ranks' particles weighted: cost_per_particle is THIS rank's measured cost of a particle in a unit all
 * ranks share and that keeps the numbers around 1 (e.g. its compute milliseconds per step and million local particles,
 * from mpm_set_timing(1) + mpm_get_stats; quantised to 1/4096; <= 0 or NaN counts as 1).  The cuts then equalise the
 * summed cost instead of the counts -- for scenes whose ranks need different times for the same number of particles
 * (BASELINE config 5: bodies of different density).  Equal costs on all ranks == mpm_comm_rebalance. */
MPM_API int32_t mpm_comm_rebalance_weighted(MpmSolver* s, int32_t max_shift, float cost_per_particle);
/* Slab of this rank: owned planes [x0, x1), stored planes [gx0, gx0 + nxl) (what mpm_download_grid returns). */
MPM_API int32_t mpm_comm_slab(const MpmSolver* s, int32_t* x0, int32_t* x1, int32_t* gx0, int32_t* nxl);
/* Global (original) index of each local particle, in the slot order of mpm_download_particles_soa. */
MPM_API int32_t mpm_download_ids(MpmSolver* s, uint32_t* ids, int64_t cap);
/* Host-only helper (usable without a GPU): equal-count slab cuts from an x-plane particle histogram.
 * cuts[0] = 0 <= cuts[1] <= ... <= cuts[world] = rx, every slab at least min_width planes wide. */
MPM_API int32_t mpm_slab_cuts(const int64_t* hist, int32_t rx, int32_t world, int32_t min_width, int32_t* cuts);

#ifdef __cplusplus
}
#endif
#endif /* MPM_B200_H */
