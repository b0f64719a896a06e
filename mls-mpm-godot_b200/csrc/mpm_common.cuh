// mpm_common.cuh -- device-side parameter block, particle/grid layout and exact-arithmetic helpers shared
// by every kernel of libmpm_b200.so.
//
// Layout in HBM
//   particles : array-of-structures-of-arrays.  Slots are grouped by 32; a group is 16 "planes" of 32 floats
//               (2 KB, contiguous): planes 0-2 pos, 3-5 vel, 6 mass, 7-15 C column-major (7-9 = Basis.X).
//               Field k of slot i lives at  (i / 32) * 512 + k * 32 + (i % 32).  A warp reading field k of 32
//               consecutive slots still reads one (or two) full 128-B lines, exactly as with plain SoA planes,
//               but all 16 fields of a slot sit at CONSTANT offsets (k * 128 B) from one address: a kernel
//               computes one pointer per particle and the other 15 accesses use immediate offsets (the
//               per-plane 64-bit address arithmetic was ~10 % of the issued instructions with plain planes).
//   grid      : array of 16-B cells (vel_x, vel_y, vel_z, mass) in the reference's index order
//               x*Ry*Rz + y*Rz + z (MLSMPM3DFluidMultithread.cs:282), int32 x 1e7 or float.
//               A multi-GPU rank stores only its x-slab: local plane lx = x - gx0, nxl planes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <utility>

namespace mpm {

// host side of pdl_prologue(): an ordinary launch with the programmatic-stream-serialization attribute (MPM_NO_PDL=1: without)
enum { PDL_BIN = 1, PDL_RANK = 2, PDL_SWEEP = 4, PDL_CELL = 8, PDL_HALO = 16, PDL_MIG = 32 };  // kernel groups (MPM_PDL_MASK: which of them get the attribute)
template <int GROUP = 63, class... KA, class... A>
inline cudaError_t launch_pdl(void (*kern)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, A&&... args)
{
    static const int mask = getenv("MPM_NO_PDL") ? 0 : (getenv("MPM_PDL_MASK") ? atoi(getenv("MPM_PDL_MASK")) : 63);
    const bool off = (mask & GROUP) == 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = off ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<A>(args)...);  // (coerces the arguments to the kernel's parameter types)
}

enum Plane { PX = 0, PY, PZ, VX, VY, VZ, PM, C0, C1, C2, C3, C4, C5, C6, C7, C8, NPLANES };
static_assert(NPLANES == 16, "the group layout below assumes 16 fields");

struct DevParams {
    int dim;
    int Rx, Ry, Rz;  // global grid resolution (Rz = 1 in 2D)
    int gx0, nxl;    // first global x plane stored locally, number of local planes (incl. ghosts)
    float dt, gravity, rest_density, visc, eos_k, eos_p;
    int eos_pi;      // eos_p if it is an integer in [1,64], else 0
    int grid_mode;   // 0 float, 1 fixed
    float fmult;     // (float)fixed_point_mult
    int stress_form, eq16_order, bc_mode, bc_hi_off;
    float bc_friction;
    float clamp_min, clamp_max_off, wall_min, wall_max_off, wall_gain;
    int interaction;
    float sphere[3], sphere_r, mouse[2], mouse_r;
    int overflow_check;
    int32_t* flags;      // device: [0] sticky fixed-point overflow flag (overflow_check), [1] particles skipped because their
                         // position was non-finite or outside the grid (the reference throws IndexOutOfRangeException there)
    int n_extra;         // further sphere repulsors (mpm_set_colliders), applied after the first one
    float extra[7][4];   // x, y, z, radius
};

constexpr int GROUP = 32;                 // slots per group
constexpr int GROUP_FLOATS = GROUP * 16;  // floats per group (16 planes)

struct ParticleView {
    float* base;
    int64_t pitch;  // capacity in slots (multiple of 128)
    // pointer to field 0 of slot i; field k is rec(i)[k * GROUP]
    __host__ __device__ __forceinline__ float* rec(int64_t i) const { return base + ((i >> 5) << 9) + (i & 31); }
    __host__ __device__ __forceinline__ float& at(int k, int64_t i) const { return base[((i >> 5) << 9) + (k << 5) + (i & 31)]; }
};

// 64-byte particle records (cell path): the 16 fields of a slot, contiguous, as four 16-byte quads
//     (px py pz m) (vx vy vz c2) (c0 c1 c3 c4) (c6 c7 c5 c8)
// The order is chosen for the packed fp32 instructions of sm_100 (FFMA2 / FMUL2 / FADD2 take aligned register pairs,
// and an LDS.128 / LDG.128 lands a quad in four consecutive registers): (px, py), (vx, vy) and the (x, y) rows of the
// three columns of C -- (c0, c1), (c3, c4), (c6, c7) -- each sit on an even offset, so the pairs the P2G kernels
// multiply with need no register moves (with the plane order they cost ~60 MOVs per particle in P2G_1).
// rec_pos(k) = offset of plane field k inside the record.
__host__ __device__ __forceinline__ constexpr int rec_pos(int k)
{
    constexpr int pos[16] = {/*PX*/ 0, /*PY*/ 1, /*PZ*/ 2, /*VX*/ 4, /*VY*/ 5, /*VZ*/ 6, /*PM*/ 3,
                             /*C0*/ 8, /*C1*/ 9, /*C2*/ 7, /*C3*/ 10, /*C4*/ 11, /*C5*/ 14, /*C6*/ 12, /*C7*/ 13, /*C8*/ 15};
    return pos[k];
}
struct RecView {
    float* base;
    __host__ __device__ __forceinline__ float& at(int k, int64_t i) const { return base[i * 16 + rec_pos(k)]; }
};

// ---- strict IEEE binary32 operators: the _rn intrinsics are never contracted into FMA by nvcc, so the
// results do not depend on -fmad and equal the C# / C oracle operation by operation.
__device__ __forceinline__ float sadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float ssub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float smul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float sdiv(float a, float b) { return __fdiv_rn(a, b); }

// EncodeFixedPoint / DecodeFixedPoint (MLSMPM3DFluidMultithreadNew.cs:151-159)
__device__ __forceinline__ int encode_fixed(float f, float fmult) { return __float2int_rz(__fmul_rn(f, fmult)); }
__device__ __forceinline__ float decode_fixed(int i, float fmult) { return __fdiv_rn(__int2float_rn(i), fmult); }

// Quadratic B-spline weights of one axis (MLSMPM3DFluidMultithread.cs:259-263), exact op order.
__device__ __forceinline__ int axis_weights(float p, float w[3])
{
    int c = __float2int_rz(p);
    float cd = ssub(ssub(p, __int2float_rn(c)), 0.5f);
    float a = ssub(0.5f, cd), b = sadd(0.5f, cd);
    w[0] = smul(0.5f, smul(a, a));
    w[1] = ssub(0.75f, smul(cd, cd));
    w[2] = smul(0.5f, smul(b, b));
    return c;
}

// EOS power (Mathf.Pow, MLSMPM3DFluidMultithread.cs:331).  The reference value is the platform CRT's
// powf (<= 1 ulp, platform-dependent); the solver evaluates in binary64 and rounds once, which is
// reproducible on any IEEE machine (integer exponents: left-to-right repeated multiplication).
__device__ __forceinline__ float eos_pow(float x, const DevParams& P)
{
    if (P.eos_pi > 0) {
        double xd = (double)x, r = xd;
        for (int k = 1; k < P.eos_pi; ++k) r = __dmul_rn(r, xd);
        return __double2float_rn(r);
    }
    return __double2float_rn(pow((double)x, (double)P.eos_p));
}

// ---- programmatic dependent launch (sm_90+).  The kernels of a step are short (3 - 700 us) and strictly ordered on one
// stream; between two of them the GPU would sit idle for the launch latency and the ramp-up of the next grid (~2-4 us, 12 to
// 25 times per step: a tenth of a 0.2 - 0.6 ms step).  A kernel launched with launch_pdl() may have its CTAs scheduled as
// soon as every CTA of the kernel before it has executed launch_dependents (or exited); it then parks in griddepcontrol.wait
// until that earlier grid has COMPLETED and its writes are visible: the ordering is the stream's, the idle gap is gone.
// Every kernel launched that way must call pdl_prologue() (or pdl_wait()) before it touches global memory; in a kernel
// launched the ordinary way both instructions are no-ops.
//
// Order: WAIT FIRST, then release the successor.  The other order (release, then wait) lets three generations of grids be
// resident at once -- kernel N still running, N + 1 and N + 2 both parked -- and that is not safe: with the three small
// migration kernels (push -> fill -> pull, all resident together) k_mig_pull then ran before k_mig_fill had finished and
// particles were lost (tests/test_multi_rank.py; bisected kernel by kernel on a B200).  With wait-then-release a grid is
// only ever released by a grid whose own prerequisite has completed: two generations in flight, as the feature is specified.
//
// pdl_wait() alone leaves the release to the CTA's exit: for kernels whose parked successors would cost them something --
// the ranking kernels lean on L1 for their table look-ups, and a successor's shared memory comes out of the same 256 KB
// (C4: +20 us on the binning with an early release, -10 us without it; C2 gains either way).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue()
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");
}

// (fmaxf / fminf return the other operand for a NaN: a non-finite position is pulled back into the clamp box instead of
// travelling on as NaN; for every other value this is v < lo ? lo : (v > hi ? hi : v))
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// ---- guards.  The reference indexes the grid with whatever (int)pos gives and lets the managed runtime throw; here a
// particle whose 3x3(x3) stencil would leave the (local) grid is skipped and counted, and mpm_sync reports it.
__device__ __forceinline__ bool stencil_in_grid(const DevParams& P, int cx, int cy, int cz)
{
    return cx - 1 >= P.gx0 && cx + 1 < P.gx0 + P.nxl && cy >= 1 && cy + 1 < P.Ry && (P.dim == 2 || (cz >= 1 && cz + 1 < P.Rz));
}
__device__ __forceinline__ void flag_bad_particle(const DevParams& P) { atomicAdd(P.flags + 1, 1); }

// ---- overflow detector (MpmParams.overflow_check): EncodeFixedPoint saturates / the int32 accumulators wrap silently in
// the reference too; with the detector on, every encode and every integer add is checked and a sticky flag is raised.
__device__ __forceinline__ int encode_fixed_checked(float f, const DevParams& P)
{
    if (P.overflow_check && !(fabsf(__fmul_rn(f, P.fmult)) < 2147483648.0f)) atomicExch(P.flags, 1);  // (also NaN)
    return encode_fixed(f, P.fmult);
}
// truncation of a value that is already in fixed-point units (MPM_MATH_FAST pre-scales by the multiplier)
__device__ __forceinline__ int f2i_checked(float x, const DevParams& P)
{
    if (P.overflow_check && !(fabsf(x) < 2147483648.0f)) atomicExch(P.flags, 1);
    return __float2int_rz(x);
}
__device__ __forceinline__ void int_add_checked(int* p, int v, const DevParams& P)
{
    if (P.overflow_check) {
        const int old = atomicAdd(p, v);
        const int sum = (int)((unsigned)old + (unsigned)v);
        if (((old ^ sum) & (v ^ sum)) < 0) atomicExch(P.flags, 1);  // operands of one sign, result of the other
    } else {
        atomicAdd(p, v);
    }
}

// local cell index of global node (nx, ny, nz)
__device__ __forceinline__ int64_t cell_index(const DevParams& P, int nx, int ny, int nz)
{
    return ((int64_t)(nx - P.gx0) * P.Ry + ny) * P.Rz + nz;
}

}  // namespace mpm
