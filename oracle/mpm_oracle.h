/*
 * mpm_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, strict IEEE binary32, no FMA contraction) of the MLS-MPM fluid step of
 * Miotismon/mls-mpm-godot.  It is the parity checker for the CUDA solver and the "port" CPU baseline of
 * bench.py.  Nothing in the product path (mls-mpm-godot_b200/, include/) may include, link or call it.
 *
 * PARITY UNPINNED: the reference ships no golden vectors, no tests and cannot be executed here (no
 * dotnet / godot / glslang in the image), so this oracle is pinned only by (1) its own invariants
 * (tests/test_oracle_invariants.py), (2) an independent NumPy restatement (oracle/oracle_np.py) that must
 * agree bit-for-bit, and (3) line-by-line citation of the C# sources below.
 *
 * Reference files followed (paths relative to /root/reference/mls-mpm):
 *   F = 3d/fluid_multithread/MLSMPM3DFluidMultithread.cs            (3D, float grid, serial P2G)
 *   X = 3d/fluid_multithread_fixed_point/MLSMPM3DFluidMultithreadNew.cs (3D, int32 x1e7 grid, atomic P2G)
 *   D = 2d/fluid/MLSMPM2DFluid.cs                                   (2D, float grid, serial)
 *   M = 2d/fluid_multithread/MLSMPM2DFluidMultithread.cs            (2D, friction BC variant)
 *   H = 3d/fluid_multithread_gpu/MLSMPM3DFluidMultithreadGPU.cs + compute_shaders/{clear_grid,p2g_1,p2g_2,update_grid,g2p}.glsl (GPU variant)
 */
#ifndef MPM_ORACLE_H
#define MPM_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* grid numeric model */
#define ORC_GRID_FLOAT 0 /* Cell{Vector3 vel; float mass}  F:16-20, D:16-20 */
#define ORC_GRID_FIXED 1 /* Cell{int vel_x,vel_y,vel_z,mass}  X:18-24, H:25-32 */

/* stress form */
#define ORC_STRESS_3D 0      /* strain = C + C^T                         F:340-345 */
#define ORC_STRESS_2D_TRACE 1 /* only off-diagonals summed, diagonal kept D:276-283 */

/* order of the scalar factors of eq. 16 */
#define ORC_EQ16_VOL4_DT 0 /* ((-volume*4)*stress)*dt   F:347, X:407, M:307, p2g_2.glsl:115 */
#define ORC_EQ16_DTVOL_4 1 /* ((-dt*volume)*stress)*4   D:285 */

/* grid boundary condition */
#define ORC_BC_SLIP 0     /* zero normal component, idx<2 || idx>R-3   F:402-404, X:479-481, D:324-325 */
#define ORC_BC_FRICTION 1 /* slip + friction 0.5, idx<2 || idx>R-4     M:366-368 */

/* interaction */
#define ORC_INTERACT_NONE 0
#define ORC_INTERACT_SPHERE_POST 1 /* sphere repulsor on post-advection pos  X:570-576 */
#define ORC_INTERACT_SPHERE_PRE 2  /* sphere repulsor on pre-advection pos   g2p.glsl:122-129 */
#define ORC_INTERACT_MOUSE_2D 3    /* radial mouse push                      D:381-406 */

/* pow semantic for the EOS (F:331 calls Mathf.Pow -> MathF.Pow -> the platform CRT powf, which differs
 * between platforms in the last ulp).  Mode 0 is the platform-independent restatement used for parity:
 * the value is computed in binary64 and rounded once to binary32 (integer exponents by left-to-right
 * repeated multiplication, others by pow()).  Mode 1 calls this host's libm powf. */
#define ORC_POW_F64_ROUNDED 0
#define ORC_POW_LIBM_POWF 1

typedef struct OrcParams {
    int32_t dim;          /* 2 or 3 */
    int32_t grid[3];      /* Rx, Ry, Rz (Rz = 1 in 2D).  Reference is cubic only (H:43). */
    float dt;             /* F:30 */
    float gravity;        /* F:33 (-0.3, y), D:33 (+0.3, y) */
    float rest_density;   /* F:36 */
    float dynamic_viscosity; /* F:37 */
    float eos_stiffness;  /* F:39, H:82 */
    float eos_power;      /* F:40, D:40, H:84 */
    int32_t grid_mode;    /* ORC_GRID_* */
    int32_t fixed_point_mult; /* X:53 */
    int32_t stress_form;  /* ORC_STRESS_* */
    int32_t eq16_order;   /* ORC_EQ16_* */
    int32_t bc_mode;      /* ORC_BC_* */
    int32_t bc_hi_off;    /* BC applies for idx > R - bc_hi_off (3: F:402, 4: M:367) */
    float bc_friction;    /* M:366 */
    float clamp_min;      /* F:476 (1), g2p.glsl:115 (2) */
    float clamp_max_off;  /* clamp max = R - clamp_max_off  (2: F:476, 1: M:434) */
    float wall_min;       /* F:507 (3), D:411 (2) */
    float wall_max_off;   /* wall_max = R - wall_max_off (4: F:508, 3: D:412 and g2p.glsl:134) */
    float wall_gain;      /* 1 (F:509), 0.5 (D:410) */
    int32_t interaction;  /* ORC_INTERACT_* */
    float sphere_pos[3];  /* X:62 */
    float sphere_radius;  /* X:63 (15) */
    float mouse_pos[2];   /* D:54 */
    float mouse_radius;   /* D:52 (10) */
    int32_t pow_mode;     /* ORC_POW_* */
    /* more sphere repulsors of the same kind ("sphere list with radius as a parameter"); applied
     * after the first one, in order, each with the test and push of X:570-576 / g2p.glsl:122-129 */
    int32_t n_extra_spheres;   /* 0..7 */
    float extra_spheres[7][4]; /* x, y, z, radius */
} OrcParams;

/* Particle state is SoA: pos[3N], vel[3N], C[9N] (column-major like Godot's Basis: C[3*col+row], so
 * C[0..2] is Basis.X / GLSL C[0]), mass[N].  2D uses x,y and the upper-left 2x2 of C.
 * Grid is G cells of four 32-bit words in reference index order x*Ry*Rz + y*Rz + z (F:282):
 * (vel_x, vel_y, vel_z, mass) -- floats in ORC_GRID_FLOAT, int32 in ORC_GRID_FIXED (X:18-24). */

/* Scene generator: lattice block, fp32 accumulating loops (F:136-146, H:661-671).  Returns the count;
 * writes at most cap positions (3 floats each) if pos != NULL. */
int64_t orc_init_block(int32_t dim, const float lo[3], const float hi[3], float spacing, float* pos,
                       int64_t cap);

void orc_clear_grid(const OrcParams* P, void* grid);
void orc_p2g1(const OrcParams* P, int64_t n, const float* pos, const float* vel, const float* C,
              const float* mass, void* grid);
void orc_p2g2(const OrcParams* P, int64_t n, const float* pos, const float* C, const float* mass,
              void* grid);
void orc_update_grid(const OrcParams* P, void* grid);
void orc_g2p(const OrcParams* P, int64_t n, float* pos, float* vel, float* C, const void* grid);
void orc_step(const OrcParams* P, int64_t n, float* pos, float* vel, float* C, const float* mass,
              void* grid, int32_t iterations);

/* Same step in the reference's threading shape, for the CPU baseline: float grid = serial P2G, all-core
 * clear/update/G2P (F:222-232,252-373,412-422); fixed grid = every phase across all cores with atomic
 * int adds (X:277-288,336-339).  Returns the thread count used. */
int32_t orc_step_mt(const OrcParams* P, int64_t n, float* pos, float* vel, float* C, const float* mass,
                    void* grid, int32_t iterations, int32_t nthreads);

/* Position hand-off of g2p.glsl:149-150: out[i] = (x, y, z, |v|). */
void orc_positions(int64_t n, const float* pos, const float* vel, float* out4);

/* Binning reference: key of the base cell (F:259 + F:282), and the stable permutation that
 * std::stable_sort by key would give. */
void orc_cell_keys(const OrcParams* P, int64_t n, const float* pos, int32_t* keys);
void orc_stable_sort_perm(int64_t n, const uint32_t* keys, int32_t* perm);

float orc_pow(int32_t mode, float x, float y);
int32_t orc_encode_fixed(float f, int32_t mult); /* X:151-154 */
float orc_decode_fixed(int32_t i, int32_t mult); /* X:156-159 */

#ifdef __cplusplus
}
#endif
#endif
