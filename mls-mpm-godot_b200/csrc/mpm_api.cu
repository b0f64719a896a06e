// mpm_api.cu -- implementation of the C ABI in include/mpm_b200.h: handle lifetime, parameter block,
// buffer upload/download in the reference's layouts, the step driver and statistics.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>

#include "mpm_bin.h"
#include "mpm_kernels.h"
#include "mpm_solver.h"

using namespace mpm;

static thread_local std::string g_create_error;

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            s->err = std::string(#call) + ": " + cudaGetErrorString(e_);                      \
            return MPM_ERR_CUDA;                                                              \
        }                                                                                     \
    } while (0)

static int fail(MpmSolver* s, int code, const std::string& msg)
{
    if (s) s->err = msg; else g_create_error = msg;
    return code;
}

// ---------------------------------------------------------------- parameters

static int validate_params(const MpmParams* p, std::string& why)
{
    if (!p) { why = "params is NULL"; return MPM_ERR_INVALID; }
    if (p->struct_size != (int32_t)sizeof(MpmParams)) { why = "MpmParams.struct_size mismatch (ABI version?)"; return MPM_ERR_INVALID; }
    if (p->dim != 2 && p->dim != 3) { why = "dim must be 2 or 3"; return MPM_ERR_INVALID; }
    for (int a = 0; a < p->dim; ++a)
        if (p->grid_size[a] < 8 || p->grid_size[a] > 4096) { why = "grid_size must be in [8, 4096] per axis"; return MPM_ERR_INVALID; }
    if (p->grid_mode != MPM_GRID_FLOAT && p->grid_mode != MPM_GRID_FIXED) { why = "grid_mode"; return MPM_ERR_INVALID; }
    if (p->grid_mode == MPM_GRID_FIXED && p->fixed_point_mult <= 0) { why = "fixed_point_mult must be > 0"; return MPM_ERR_INVALID; }
    if (p->grid_mode == MPM_GRID_FIXED && p->bc_mode == MPM_BC_FRICTION) { why = "friction BC exists only for the float grid (reference M)"; return MPM_ERR_INVALID; }
    if (p->stress_form != MPM_STRESS_3D && p->stress_form != MPM_STRESS_2D_TRACE) { why = "stress_form"; return MPM_ERR_INVALID; }
    if ((p->dim == 2) != (p->stress_form == MPM_STRESS_2D_TRACE)) { why = "stress_form must match dim"; return MPM_ERR_INVALID; }
    if (p->eq16_order != 0 && p->eq16_order != 1) { why = "eq16_order"; return MPM_ERR_INVALID; }
    if (p->dim == 3 && p->eq16_order != MPM_EQ16_VOL4_DT) { why = "eq16_order DTVOL_4 exists only in 2D (reference D)"; return MPM_ERR_INVALID; }
    if (p->bc_mode != 0 && p->bc_mode != 1) { why = "bc_mode"; return MPM_ERR_INVALID; }
    if (p->interaction < 0 || p->interaction > 3) { why = "interaction"; return MPM_ERR_INVALID; }
    if (p->interaction == MPM_INTERACT_MOUSE_2D && p->dim != 2) { why = "mouse interaction is 2D only"; return MPM_ERR_INVALID; }
    if (p->math_mode != MPM_MATH_STRICT && p->math_mode != MPM_MATH_FAST) { why = "math_mode"; return MPM_ERR_INVALID; }
    if (p->kernel_path < 0 || p->kernel_path > 3) { why = "kernel_path"; return MPM_ERR_INVALID; }
    if (p->sort_interval < 0) { why = "sort_interval"; return MPM_ERR_INVALID; }
    if (!(p->rest_density > 0.0f)) { why = "rest_density must be > 0"; return MPM_ERR_INVALID; }
    return MPM_OK;
}

static void to_dev(const MpmParams& h, DevParams& d, int gx0, int nxl)
{   // (the collider list in d.n_extra / d.extra is set by mpm_set_colliders and left alone here)
    d.dim = h.dim;
    d.Rx = h.grid_size[0]; d.Ry = h.grid_size[1]; d.Rz = (h.dim == 3) ? h.grid_size[2] : 1;
    d.gx0 = gx0; d.nxl = nxl;
    d.dt = std::min(std::max(h.dt, 0.0f), 0.4f);  // Dt setter, MLSMPM3DFluidMultithreadGPU.cs:64
    d.gravity = h.gravity; d.rest_density = h.rest_density; d.visc = h.dynamic_viscosity;
    d.eos_k = h.eos_stiffness; d.eos_p = h.eos_power;
    d.eos_pi = (h.eos_power == truncf(h.eos_power) && h.eos_power >= 1.0f && h.eos_power <= 64.0f) ? (int)h.eos_power : 0;
    d.grid_mode = h.grid_mode; d.fmult = (float)h.fixed_point_mult;
    d.stress_form = h.stress_form; d.eq16_order = h.eq16_order; d.bc_mode = h.bc_mode; d.bc_hi_off = h.bc_hi_off;
    d.bc_friction = h.bc_friction;
    d.clamp_min = h.clamp_min; d.clamp_max_off = h.clamp_max_off;
    d.wall_min = h.wall_min; d.wall_max_off = h.wall_max_off; d.wall_gain = h.wall_gain;
    d.interaction = h.interaction;
    for (int a = 0; a < 3; ++a) d.sphere[a] = h.sphere_pos[a];
    d.sphere_r = h.sphere_radius; d.mouse[0] = h.mouse_pos[0]; d.mouse[1] = h.mouse_pos[1]; d.mouse_r = h.mouse_radius;
    d.overflow_check = h.overflow_check;
}

extern "C" int32_t mpm_abi_version(void) { return MPM_ABI_VERSION; }

extern "C" int32_t mpm_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int32_t mpm_default_params(int32_t variant, MpmParams* p)
{
    if (!p || variant < 0 || variant > 4) return MPM_ERR_INVALID;
    memset(p, 0, sizeof(*p));
    p->struct_size = (int32_t)sizeof(MpmParams);
    p->dt = 0.2f; p->rest_density = 4.0f; p->dynamic_viscosity = 0.1f;
    p->fixed_point_mult = 10000000;
    p->bc_friction = 0.5f; p->sphere_radius = 15.0f; p->mouse_radius = 10.0f;
    p->bc_mode = MPM_BC_SLIP; p->bc_hi_off = 3;
    p->math_mode = MPM_MATH_STRICT; p->kernel_path = MPM_PATH_AUTO; p->sort_interval = 0;
    int R = 64;
    switch (variant) {
        case MPM_VARIANT_2D_ST:
            p->dim = 2; R = 64; p->gravity = 0.3f; p->eos_stiffness = 10.0f; p->eos_power = 7.0f;
            p->grid_mode = MPM_GRID_FLOAT; p->stress_form = MPM_STRESS_2D_TRACE; p->eq16_order = MPM_EQ16_DTVOL_4;
            p->clamp_min = 1.0f; p->clamp_max_off = 2.0f; p->wall_min = 2.0f; p->wall_max_off = 3.0f; p->wall_gain = 0.5f;
            break;
        case MPM_VARIANT_2D_MT:
            p->dim = 2; R = 64; p->gravity = 0.3f; p->eos_stiffness = 10.0f; p->eos_power = 4.0f;
            p->grid_mode = MPM_GRID_FLOAT; p->stress_form = MPM_STRESS_2D_TRACE; p->eq16_order = MPM_EQ16_VOL4_DT;
            p->bc_mode = MPM_BC_FRICTION; p->bc_hi_off = 4;
            p->clamp_min = 1.0f; p->clamp_max_off = 1.0f; p->wall_min = 2.0f; p->wall_max_off = 3.0f; p->wall_gain = 0.5f;
            break;
        case MPM_VARIANT_3D_FLOAT:
            p->dim = 3; R = 32; p->gravity = -0.3f; p->eos_stiffness = 10.0f; p->eos_power = 4.0f;
            p->grid_mode = MPM_GRID_FLOAT; p->stress_form = MPM_STRESS_3D; p->eq16_order = MPM_EQ16_VOL4_DT;
            p->clamp_min = 1.0f; p->clamp_max_off = 2.0f; p->wall_min = 3.0f; p->wall_max_off = 4.0f; p->wall_gain = 1.0f;
            break;
        case MPM_VARIANT_3D_FIXED:
            p->dim = 3; R = 32; p->gravity = -0.3f; p->eos_stiffness = 10.0f; p->eos_power = 4.0f;
            p->grid_mode = MPM_GRID_FIXED; p->stress_form = MPM_STRESS_3D; p->eq16_order = MPM_EQ16_VOL4_DT;
            p->clamp_min = 1.0f; p->clamp_max_off = 2.0f; p->wall_min = 3.0f; p->wall_max_off = 4.0f; p->wall_gain = 1.0f;
            p->interaction = MPM_INTERACT_SPHERE_POST;
            p->sphere_pos[0] = 0.0f; p->sphere_pos[1] = 0.0f; p->sphere_pos[2] = 31.707275f;
            break;
        case MPM_VARIANT_3D_GPU:
            p->dim = 3; R = 64; p->gravity = -0.3f; p->eos_stiffness = 1.0f; p->eos_power = 7.0f;
            p->grid_mode = MPM_GRID_FIXED; p->stress_form = MPM_STRESS_3D; p->eq16_order = MPM_EQ16_VOL4_DT;
            p->clamp_min = 2.0f; p->clamp_max_off = 2.0f; p->wall_min = 3.0f; p->wall_max_off = 3.0f; p->wall_gain = 1.0f;
            p->interaction = MPM_INTERACT_SPHERE_PRE;
            p->sphere_pos[0] = -21.648403f; p->sphere_pos[1] = 0.0f; p->sphere_pos[2] = 31.707275f;
            break;
    }
    p->grid_size[0] = R; p->grid_size[1] = R; p->grid_size[2] = (p->dim == 3) ? R : 1;
    return MPM_OK;
}

// ---------------------------------------------------------------- lifetime

static int resolve_path(const MpmParams& p)
{
    if (p.kernel_path != MPM_PATH_AUTO) return p.kernel_path;
    // (the overflow detector lives in the reference-shaped and tiled kernels: with it on, FAST runs the tiled kernels)
    if (p.dim == 3 && p.grid_mode == MPM_GRID_FIXED) return (p.math_mode == MPM_MATH_FAST && !p.overflow_check) ? MPM_PATH_CELL : MPM_PATH_TILED;
    return MPM_PATH_REFERENCE;
}

extern "C" int32_t mpm_create(const MpmParams* p, int64_t max_particles, int32_t device, MpmSolver** out)
{
    if (!out) return fail(nullptr, MPM_ERR_INVALID, "out is NULL");
    *out = nullptr;
    std::string why;
    int rc = validate_params(p, why);
    if (rc) return fail(nullptr, rc, why);
    if (max_particles <= 0 || max_particles > (int64_t)0x7fffff00) return fail(nullptr, MPM_ERR_INVALID, "max_particles out of range");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, MPM_ERR_CUDA, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                                               " (libmpm_b200 has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fail(nullptr, MPM_ERR_INVALID, "device index out of range");
    MpmSolver* s = new (std::nothrow) MpmSolver();
    if (!s) return fail(nullptr, MPM_ERR_INVALID, "out of host memory");
    s->hp = *p;
    if (s->hp.dim == 2) s->hp.grid_size[2] = 1;
    s->device = device;
    auto bail = [&](int code) { g_create_error = s->err; mpm_destroy(s); return code; };
#define CKC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { s->err = std::string(#call) + ": " + cudaGetErrorString(e_); return bail(MPM_ERR_CUDA); } } while (0)
    CKC(cudaSetDevice(device));
    CKC(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    s->cap = max_particles;
    s->pitch = (max_particles + 127) / 128 * 128;
    to_dev(s->hp, s->dp, 0, s->hp.grid_size[0]);
    s->ncells = (int64_t)s->dp.nxl * s->dp.Ry * s->dp.Rz;
    CKC(cudaMalloc(&s->part, sizeof(float) * NPLANES * s->pitch));
    CKC(cudaMalloc(&s->orig_id, sizeof(uint32_t) * s->pitch));
    CKC(cudaMalloc(&s->grid, 16 * s->ncells));
    // the hand-off array: exportable (file descriptor for Vulkan / another process) where the driver allows, else plain
    if (vmm_alloc(device, sizeof(float4) * s->pitch, &s->positions_mem)) s->positions = reinterpret_cast<float4*>(s->positions_mem.ptr);
    else CKC(cudaMalloc(&s->positions, sizeof(float4) * s->pitch));
    CKC(cudaMalloc(&s->overflow_flag, 2 * sizeof(int32_t)));  // [0] overflow detector, [1] particles skipped for their position
    s->dp.flags = s->overflow_flag;
    CKC(cudaMemsetAsync(s->grid, 0, 16 * s->ncells, s->stream));
    CKC(cudaMemsetAsync(s->overflow_flag, 0, 2 * sizeof(int32_t), s->stream));
    s->path = resolve_path(s->hp);
    if (s->path == MPM_PATH_TILED && !(s->hp.dim == 3 && s->hp.grid_mode == MPM_GRID_FIXED)) {
        s->err = "MPM_PATH_TILED implements the 3D fixed-point grid; use MPM_PATH_AUTO or MPM_PATH_REFERENCE";
        return bail(MPM_ERR_INVALID);
    }
    if (s->path == MPM_PATH_CELL && !(s->hp.dim == 3 && s->hp.grid_mode == MPM_GRID_FIXED && s->hp.math_mode == MPM_MATH_FAST)) {
        s->err = "MPM_PATH_CELL implements dim = 3, MPM_GRID_FIXED, MPM_MATH_FAST; use MPM_PATH_AUTO";
        return bail(MPM_ERR_INVALID);
    }
    if (s->path == MPM_PATH_CELL && s->hp.overflow_check) {
        s->err = "overflow_check is a debug detector of the reference-shaped and tiled kernels; use MPM_PATH_AUTO or MPM_PATH_TILED with it";
        return bail(MPM_ERR_INVALID);
    }
    s->sort_interval = s->hp.sort_interval > 0 ? s->hp.sort_interval : 1;
    if (s->path == MPM_PATH_CELL) CKC(cudaMalloc(&s->rec, sizeof(float) * 16 * s->pitch));
    if (s->path == MPM_PATH_TILED || s->path == MPM_PATH_CELL) {
        CKC(cudaMalloc(&s->part_alt, sizeof(float) * NPLANES * s->pitch));
        CKC(cudaMalloc(&s->orig_id_alt, sizeof(uint32_t) * s->pitch));
        rc = (s->path == MPM_PATH_TILED) ? sort_create(s) : bin_create(s);
        if (rc) return bail(rc);
    }
    CKC(cudaStreamSynchronize(s->stream));
#undef CKC
    *out = s;
    return MPM_OK;
}

extern "C" int32_t mpm_destroy(MpmSolver* s)
{
    if (!s) return MPM_OK;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    comm_destroy(s);
    sort_destroy(s);
    bin_destroy(s);
    for (cudaEvent_t ev : s->ev) cudaEventDestroy(ev);
    cudaFree(s->part); cudaFree(s->part_alt); cudaFree(s->rec); cudaFree(s->orig_id); cudaFree(s->orig_id_alt);
    if (s->positions_mem.ptr) vmm_free(&s->positions_mem); else cudaFree(s->positions);
    cudaFree(s->grid); cudaFree(s->positions_b); cudaFree(s->overflow_flag); cudaFree(s->stage);
    if (s->copy_stream) { cudaStreamSynchronize(s->copy_stream); cudaStreamDestroy(s->copy_stream); }
    for (int k = 0; k < 2; ++k) { if (s->pos_ready[k]) cudaEventDestroy(s->pos_ready[k]); if (s->pos_copied[k]) cudaEventDestroy(s->pos_copied[k]); }
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
    return MPM_OK;
}

extern "C" const char* mpm_last_error(const MpmSolver* s) { return s ? s->err.c_str() : g_create_error.c_str(); }

extern "C" int32_t mpm_set_params(MpmSolver* s, const MpmParams* p)
{
    if (!s) return MPM_ERR_INVALID;
    std::string why;
    int rc = validate_params(p, why);
    if (rc) return fail(s, rc, why);
    if (p->dim != s->hp.dim || p->grid_mode != s->hp.grid_mode || p->grid_size[0] != s->hp.grid_size[0] ||
        p->grid_size[1] != s->hp.grid_size[1] || (p->dim == 3 && p->grid_size[2] != s->hp.grid_size[2]))
        return fail(s, MPM_ERR_INVALID, "dim, grid_size and grid_mode are fixed at mpm_create");
    if (resolve_path(*p) != s->path) return fail(s, MPM_ERR_INVALID, "kernel_path is fixed at mpm_create");
    s->hp = *p;
    if (s->hp.dim == 2) s->hp.grid_size[2] = 1;
    to_dev(s->hp, s->dp, s->dp.gx0, s->dp.nxl);
    s->sort_interval = s->hp.sort_interval > 0 ? s->hp.sort_interval : 1;
    return MPM_OK;
}

extern "C" int32_t mpm_get_params(const MpmSolver* s, MpmParams* p)
{
    if (!s || !p) return MPM_ERR_INVALID;
    *p = s->hp;
    p->dt = s->dp.dt;
    return MPM_OK;
}

extern "C" int32_t mpm_set_sphere(MpmSolver* s, const float pos[3])
{
    if (!s || !pos) return MPM_ERR_INVALID;
    for (int a = 0; a < 3; ++a) { s->hp.sphere_pos[a] = pos[a]; s->dp.sphere[a] = pos[a]; }
    return MPM_OK;
}

extern "C" int32_t mpm_set_colliders(MpmSolver* s, const float* xyzr, int32_t count)
{
    if (!s || count < 0 || count > MPM_MAX_EXTRA_SPHERES || (count > 0 && !xyzr)) return MPM_ERR_INVALID;
    s->dp.n_extra = count;
    for (int k = 0; k < count; ++k)
        for (int a = 0; a < 4; ++a) s->dp.extra[k][a] = xyzr[4 * k + a];
    return MPM_OK;
}

// ---------------------------------------------------------------- particle set

// multi-GPU (mpm_comm.cu): the GLOBAL set is written to the planes by init/add/upload; the slab partition
// (histogram -> cuts -> keep own particles) runs lazily before the next step / download.
namespace mpm {
int comm_partition(MpmSolver* s);          // no-op unless a partition is pending (also catches up on a pending migration's counts)
int comm_partition_for_step(MpmSolver* s); // ... without the catching up (mpm_step)
int comm_halo_planes(const MpmSolver* s, int side);  // stored planes at the low (0) / high (1) end that neighbours add to: 0 or 2
bool comm_partitioned(const MpmSolver* s);  // the planes hold this rank's local particles
void comm_mark_global(MpmSolver* s);       // the planes now hold the global set again
}

// Uploaded / loaded positions must be finite and leave room for the 3x3(x3) stencil: [1, R-1) per axis (the
// reference clamps to [1, R-2] / [2, R-2] after every step, so nothing it produces is outside).
__global__ void __launch_bounds__(256) k_count_bad_positions(mpm::ParticleView pv, int64_t first, int64_t n, int dim, float rx, float ry, float rz, int* bad)
{
    const int64_t i = first + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= first + n) return;
    const float x = pv.at(mpm::PX, i), y = pv.at(mpm::PY, i), z = pv.at(mpm::PZ, i);
    bool ok = x >= 1.0f && x < rx - 1.0f && y >= 1.0f && y < ry - 1.0f;  // (false for NaN)
    if (dim == 3) ok = ok && z >= 1.0f && z < rz - 1.0f;
    if (!ok) atomicAdd(bad, 1);
}

static int validate_positions(MpmSolver* s, int64_t first, int64_t n)
{
    if (n <= 0) return MPM_OK;
    int* d_bad = nullptr;
    int bad = 0;
    CK(cudaMalloc(&d_bad, sizeof(int)));
    cudaMemsetAsync(d_bad, 0, sizeof(int), s->stream);
    k_count_bad_positions<<<(unsigned)((n + 255) / 256), 256, 0, s->stream>>>(s->view(), first, n, s->hp.dim, (float)s->hp.grid_size[0], (float)s->hp.grid_size[1],
                                                                           (float)s->hp.grid_size[2], d_bad);
    cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, s->stream);
    cudaError_t e = cudaStreamSynchronize(s->stream);
    cudaFree(d_bad);
    if (e != cudaSuccess) return fail(s, MPM_ERR_CUDA, cudaGetErrorString(e));
    if (bad) {
        s->n = 0;  // the set is rejected as a whole
        return fail(s, MPM_ERR_DOMAIN, std::to_string(bad) + " of " + std::to_string(n) + " particle positions are non-finite or outside [1, R-1)");
    }
    return MPM_OK;
}

// grow-only device staging buffer for uploads / downloads (they used to cudaMalloc + cudaFree on every call)
static int stage_buffer(MpmSolver* s, size_t bytes, float** out)
{
    if (bytes > s->stage_bytes) {
        cudaFree(s->stage);
        s->stage = nullptr; s->stage_bytes = 0;
        CK(cudaMalloc(&s->stage, bytes));
        s->stage_bytes = bytes;
    }
    *out = reinterpret_cast<float*>(s->stage);
    return MPM_OK;
}

static void particles_changed(MpmSolver* s)
{
    s->positions_valid = false;
    s->sorted_valid = false;
    s->steps_since_sort = 0;
    s->fresh_particles = true;
    s->in_rec = false;  // the planes were just (re)written
    if (s->bin) s->bin->next_valid = false;
    if (s->comm) comm_mark_global(s);
}

static int lattice_axis(float lo, float hi, float spacing, std::vector<float>& out)
{
    if (!(spacing > 0.0f)) return MPM_ERR_INVALID;
    // for (float i = lo; i < hi; i += spacing): fp32 accumulation (MLSMPM3DFluidMultithreadGPU.cs:661)
    for (float i = lo; i < hi; i += spacing) {
        out.push_back(i);
        if (out.size() > (size_t)1 << 20) return MPM_ERR_INVALID;
    }
    return MPM_OK;
}

static int add_block(MpmSolver* s, const float lo[3], const float hi[3], float spacing, bool replace)
{
    if (!s || !lo || !hi) return MPM_ERR_INVALID;
    CK(cudaSetDevice(s->device));
    std::vector<float> ax[3];
    for (int a = 0; a < s->hp.dim; ++a)
        if (lattice_axis(lo[a], hi[a], spacing, ax[a])) return fail(s, MPM_ERR_INVALID, "bad lattice extent/spacing");
    if (s->hp.dim == 2) ax[2].push_back(0.0f);
    const int64_t cnt = (int64_t)ax[0].size() * ax[1].size() * ax[2].size();
    const int64_t base = replace ? 0 : s->n;
    if (base + cnt > s->cap) return fail(s, MPM_ERR_INVALID, "lattice exceeds max_particles");
    if (s->comm && !replace && comm_partitioned(s))
        return fail(s, MPM_ERR_STATE, "multi-GPU: mpm_add_block after the slab partition (a step or download) is not supported");
    if (!replace) ensure_planes(s);
    float* d_ax = nullptr;
    const size_t tot = ax[0].size() + ax[1].size() + ax[2].size();
    CK(cudaMalloc(&d_ax, sizeof(float) * tot));
    float* dx = d_ax; float* dy = dx + ax[0].size(); float* dz = dy + ax[1].size();
    cudaMemcpyAsync(dx, ax[0].data(), sizeof(float) * ax[0].size(), cudaMemcpyHostToDevice, s->stream);
    cudaMemcpyAsync(dy, ax[1].data(), sizeof(float) * ax[1].size(), cudaMemcpyHostToDevice, s->stream);
    cudaMemcpyAsync(dz, ax[2].data(), sizeof(float) * ax[2].size(), cudaMemcpyHostToDevice, s->stream);
    launch_lattice(dx, (int)ax[0].size(), dy, (int)ax[1].size(), dz, (int)ax[2].size(), s->view(), base, s->stream);
    launch_iota(s->orig_id + base, (uint32_t)base, cnt, s->stream);
    s->launches += 2;
    if (replace) cudaMemsetAsync(s->grid, 0, 16 * s->ncells, s->stream);
    cudaError_t e = cudaStreamSynchronize(s->stream);
    cudaFree(d_ax);
    if (e != cudaSuccess) return fail(s, MPM_ERR_CUDA, cudaGetErrorString(e));
    s->n = base + cnt;
    particles_changed(s);
    return validate_positions(s, base, cnt);
}

extern "C" int32_t mpm_init_block(MpmSolver* s, const float lo[3], const float hi[3], float spacing)
{
    return add_block(s, lo, hi, spacing, true);
}
extern "C" int32_t mpm_add_block(MpmSolver* s, const float lo[3], const float hi[3], float spacing)
{
    return add_block(s, lo, hi, spacing, false);
}


extern "C" int32_t mpm_upload_particles(MpmSolver* s, const MpmParticle80* ps, int64_t n)
{
    if (!s || (!ps && n > 0) || n < 0) return MPM_ERR_INVALID;
    if (n > s->cap) return fail(s, MPM_ERR_INVALID, "n exceeds max_particles");
    CK(cudaSetDevice(s->device));
    if (n > 0) {
        float* stage = nullptr;
        { int rc_ = stage_buffer(s, sizeof(MpmParticle80) * n, &stage); if (rc_) return rc_; }
        cudaMemcpyAsync(stage, ps, sizeof(MpmParticle80) * n, cudaMemcpyHostToDevice, s->stream);
        launch_aos80_to_soa(stage, s->view(), 0, n, s->stream);
        launch_iota(s->orig_id, 0, n, s->stream);
        s->launches += 2;
        cudaError_t e = cudaStreamSynchronize(s->stream);
        if (e != cudaSuccess) return fail(s, MPM_ERR_CUDA, cudaGetErrorString(e));
    }
    s->n = n;
    particles_changed(s);
    return validate_positions(s, 0, n);
}

extern "C" int32_t mpm_upload_particles_soa(MpmSolver* s, const float* pos, const float* vel, const float* C,
                                            const float* mass, int64_t n)
{
    if (!s || (!pos && n > 0) || n < 0) return MPM_ERR_INVALID;
    if (n > s->cap) return fail(s, MPM_ERR_INVALID, "n exceeds max_particles");
    CK(cudaSetDevice(s->device));
    if (n > 0) {
        float* stage = nullptr;
        { int rc_ = stage_buffer(s, sizeof(float) * 16 * n, &stage); if (rc_) return rc_; }
        float* dpos = stage; float* dvel = stage + 3 * n; float* dC = stage + 6 * n; float* dm = stage + 15 * n;
        cudaMemcpyAsync(dpos, pos, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, s->stream);
        if (vel) cudaMemcpyAsync(dvel, vel, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, s->stream);
        if (C) cudaMemcpyAsync(dC, C, sizeof(float) * 9 * n, cudaMemcpyHostToDevice, s->stream);
        if (mass) cudaMemcpyAsync(dm, mass, sizeof(float) * n, cudaMemcpyHostToDevice, s->stream);
        launch_packed_to_soa(dpos, vel ? dvel : nullptr, C ? dC : nullptr, mass ? dm : nullptr, s->view(), 0, n, s->stream);
        launch_iota(s->orig_id, 0, n, s->stream);
        s->launches += 2;
        cudaError_t e = cudaStreamSynchronize(s->stream);
        if (e != cudaSuccess) return fail(s, MPM_ERR_CUDA, cudaGetErrorString(e));
    }
    s->n = n;
    particles_changed(s);
    return validate_positions(s, 0, n);
}

extern "C" int32_t mpm_download_particles(MpmSolver* s, MpmParticle80* ps, int64_t cap)
{
    if (!s || !ps) return MPM_ERR_INVALID;
    CK(cudaSetDevice(s->device));
    { int rc = comm_partition(s); if (rc) return rc; }
    if (cap < s->n) return fail(s, MPM_ERR_INVALID, "destination too small");
    if (s->n == 0) return MPM_OK;
    ensure_planes(s);
    float* stage = nullptr;
    { int rc_ = stage_buffer(s, sizeof(MpmParticle80) * s->n, &stage); if (rc_) return rc_; }
    // multi-GPU ranks return their local particles in slot order (global indices: mpm_download_ids)
    launch_soa_to_aos80(s->view(), s->comm ? nullptr : s->orig_id, stage, s->n, s->stream);
    s->launches += 1;
    cudaMemcpyAsync(ps, stage, sizeof(MpmParticle80) * s->n, cudaMemcpyDeviceToHost, s->stream);
    cudaError_t e = cudaStreamSynchronize(s->stream);
    if (e != cudaSuccess) return fail(s, MPM_ERR_CUDA, cudaGetErrorString(e));
    return MPM_OK;
}

extern "C" int32_t mpm_download_particles_soa(MpmSolver* s, float* pos, float* vel, float* C, float* mass, int64_t cap)
{
    if (!s) return MPM_ERR_INVALID;
    CK(cudaSetDevice(s->device));
    { int rc = comm_partition(s); if (rc) return rc; }
    if (cap < s->n) return fail(s, MPM_ERR_INVALID, "destination too small");
    const int64_t n = s->n;
    if (n == 0) return MPM_OK;
    ensure_planes(s);
    float* stage = nullptr;
    { int rc_ = stage_buffer(s, sizeof(float) * 16 * n, &stage); if (rc_) return rc_; }
    float* dpos = stage; float* dvel = stage + 3 * n; float* dC = stage + 6 * n; float* dm = stage + 15 * n;
    // multi-GPU ranks return their local particles in slot order (global indices: mpm_download_ids)
    launch_soa_to_packed(s->view(), s->comm ? nullptr : s->orig_id, dpos, dvel, dC, dm, n, s->stream);
    s->launches += 1;
    if (pos) cudaMemcpyAsync(pos, dpos, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, s->stream);
    if (vel) cudaMemcpyAsync(vel, dvel, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, s->stream);
    if (C) cudaMemcpyAsync(C, dC, sizeof(float) * 9 * n, cudaMemcpyDeviceToHost, s->stream);
    if (mass) cudaMemcpyAsync(mass, dm, sizeof(float) * n, cudaMemcpyDeviceToHost, s->stream);
    cudaError_t e = cudaStreamSynchronize(s->stream);
    if (e != cudaSuccess) return fail(s, MPM_ERR_CUDA, cudaGetErrorString(e));
    return MPM_OK;
}

extern "C" int32_t mpm_download_grid(MpmSolver* s, MpmCell16* cells, int64_t cap)
{
    if (!s || !cells) return MPM_ERR_INVALID;
    CK(cudaSetDevice(s->device));
    { int rc = comm_partition(s); if (rc) return rc; }
    if (cap < s->ncells) return fail(s, MPM_ERR_INVALID, "destination too small");
    if (s->grid_raw) { launch_update_grid(s->dp, s->grid, s->ncells, s->stream); s->launches += 1; s->grid_raw = false; }
    CK(cudaMemcpyAsync(cells, s->grid, 16 * s->ncells, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return MPM_OK;
}

// ---------------------------------------------------------------- checkpoint / resume
// File (little-endian): 64-byte header | parameter block | n particle records of 80 bytes (H:8-22) | [n original indices].
// Version 2 carries the parameters the state was produced with (MpmParams + collider list): loading into a solver whose
// physics differs is refused instead of silently resuming something else.  Under a communicator every rank writes its own
// particles with their global indices to "<path>.rank<r>of<w>"; mpm_load_state(path) finds those files, puts every
// particle back at its original index and uploads the whole set (on every rank when a communicator is attached: the slab
// partition then keeps each rank's share, for any world size).
struct StateHeader {
    char magic[8];
    uint32_t version;
    int32_t dim;
    int32_t grid[3];
    int32_t has_ids;     // v2: 1 = a table of original indices follows the records (a rank's share of a multi-GPU state)
    int64_t n;
    int64_t steps;
    int32_t rank, world; // v2 (0 / 1 for a single-GPU state)
    int64_t n_global;    // v2: particles of the whole scene
};
static_assert(sizeof(StateHeader) == 64, "checkpoint header is 64 bytes");
struct StateParams {
    MpmParams p;
    int32_t n_extra;
    float extra[MPM_MAX_EXTRA_SPHERES][4];
};
namespace mpm { int64_t comm_global_count(const MpmSolver* s); int comm_rank_world(const MpmSolver* s, int* rank, int* world); }

static bool same_physics(const MpmParams& a, const MpmParams& b, std::string& why)
{
#define SAME(f) if (memcmp(&a.f, &b.f, sizeof(a.f)) != 0) { why = #f; return false; }
    SAME(dim) SAME(grid_size) SAME(dt) SAME(gravity) SAME(rest_density) SAME(dynamic_viscosity) SAME(eos_stiffness) SAME(eos_power)
    SAME(grid_mode) SAME(fixed_point_mult) SAME(stress_form) SAME(eq16_order) SAME(bc_mode) SAME(bc_hi_off) SAME(bc_friction)
    SAME(clamp_min) SAME(clamp_max_off) SAME(wall_min) SAME(wall_max_off) SAME(wall_gain) SAME(interaction) SAME(sphere_radius)
    SAME(mouse_radius) SAME(math_mode)
#undef SAME
    return true;  // (sphere / mouse position are per-frame inputs, the kernel path and sort interval do not change results in strict mode)
}

extern "C" int32_t mpm_save_state(MpmSolver* s, const char* path)
{
    if (!s || !path) return MPM_ERR_INVALID;
    CK(cudaSetDevice(s->device));
    { int rc = comm_partition(s); if (rc) return rc; }
    std::vector<MpmParticle80> buf((size_t)std::max<int64_t>(s->n, 1));
    std::vector<uint32_t> ids;
    StateHeader h{};
    std::string file = path;
    if (s->comm) {  // a rank's local particles in slot order + their global indices
        int rank = 0, world = 1;
        comm_rank_world(s, &rank, &world);
        ensure_planes(s);
        float* stage = nullptr;
        if (s->n > 0) {
            { int rc_ = stage_buffer(s, sizeof(MpmParticle80) * s->n, &stage); if (rc_) return rc_; }
            launch_soa_to_aos80(s->view(), nullptr, stage, s->n, s->stream);
            CK(cudaMemcpyAsync(buf.data(), stage, sizeof(MpmParticle80) * s->n, cudaMemcpyDeviceToHost, s->stream));
            ids.resize((size_t)s->n);
            CK(cudaMemcpyAsync(ids.data(), s->orig_id, sizeof(uint32_t) * s->n, cudaMemcpyDeviceToHost, s->stream));
            CK(cudaStreamSynchronize(s->stream));
        }
        h.has_ids = 1; h.rank = rank; h.world = world; h.n_global = comm_global_count(s);
        file += ".rank" + std::to_string(rank) + "of" + std::to_string(world);
    } else {
        int rc = s->n > 0 ? mpm_download_particles(s, buf.data(), s->n) : MPM_OK;
        if (rc) return rc;
        h.world = 1; h.n_global = s->n;
    }
    memcpy(h.magic, "MPMB200", 8);
    h.version = 2; h.dim = s->hp.dim;
    for (int a = 0; a < 3; ++a) h.grid[a] = s->hp.grid_size[a];
    h.n = s->n; h.steps = s->steps;
    StateParams sp{};
    sp.p = s->hp; sp.n_extra = s->dp.n_extra;
    memcpy(sp.extra, s->dp.extra, sizeof(sp.extra));
    FILE* f = fopen(file.c_str(), "wb");
    if (!f) return fail(s, MPM_ERR_INVALID, std::string("cannot open for writing: ") + file);
    bool ok = fwrite(&h, sizeof(h), 1, f) == 1 && fwrite(&sp, sizeof(sp), 1, f) == 1 &&
              (s->n == 0 || fwrite(buf.data(), sizeof(MpmParticle80), (size_t)s->n, f) == (size_t)s->n);
    if (ok && h.has_ids && s->n > 0) ok = fwrite(ids.data(), sizeof(uint32_t), (size_t)s->n, f) == (size_t)s->n;
    if (fclose(f) != 0 || !ok) return fail(s, MPM_ERR_INVALID, std::string("short write: ") + file);
    return MPM_OK;
}

// one checkpoint file -> its records land in `all` (at their original index if the file carries one, else in file order)
static int read_state_file(MpmSolver* s, const std::string& file, std::vector<MpmParticle80>& all, StateHeader& h, int64_t* filled)
{
    FILE* f = fopen(file.c_str(), "rb");
    if (!f) return fail(s, MPM_ERR_INVALID, std::string("cannot open: ") + file);
    auto bad = [&](const std::string& why) { fclose(f); return fail(s, MPM_ERR_INVALID, file + ": " + why); };
    if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "MPMB200", 8) != 0 || (h.version != 1 && h.version != 2))
        return bad("not an mpm_b200 checkpoint (bad header)");
    if (h.dim != s->hp.dim || h.grid[0] != s->hp.grid_size[0] || h.grid[1] != s->hp.grid_size[1] || h.grid[2] != s->hp.grid_size[2])
        return bad("checkpoint was written for a different dim / grid size");
    if (h.version == 1) { h.has_ids = 0; h.rank = 0; h.world = 1; h.n_global = h.n; }
    else {
        StateParams sp{};
        if (fread(&sp, sizeof(sp), 1, f) != 1 || sp.p.struct_size != (int32_t)sizeof(MpmParams)) return bad("parameter block missing or of another ABI version");
        std::string why;
        if (!same_physics(sp.p, s->hp, why)) return bad("written with a different `" + why + "`: set the solver's parameters to the checkpoint's first");
        if (sp.n_extra != s->dp.n_extra || memcmp(sp.extra, s->dp.extra, sizeof(float) * 4 * (size_t)std::max(sp.n_extra, 0)) != 0)
            return bad("written with a different collider list (mpm_set_colliders)");
    }
    if (h.n < 0 || h.n_global < h.n || h.n_global > s->cap) return bad("holds more particles than max_particles");
    if ((int64_t)all.size() < h.n_global) all.resize((size_t)h.n_global);
    std::vector<MpmParticle80> buf((size_t)std::max<int64_t>(h.n, 1));
    if (h.n > 0 && fread(buf.data(), sizeof(MpmParticle80), (size_t)h.n, f) != (size_t)h.n) return bad("truncated");
    if (h.has_ids) {
        std::vector<uint32_t> ids((size_t)std::max<int64_t>(h.n, 1));
        if (h.n > 0 && fread(ids.data(), sizeof(uint32_t), (size_t)h.n, f) != (size_t)h.n) return bad("truncated (index table)");
        for (int64_t i = 0; i < h.n; ++i) {
            if ((int64_t)ids[(size_t)i] >= h.n_global) return bad("original index out of range");
            all[ids[(size_t)i]] = buf[(size_t)i];
        }
    } else {
        std::copy(buf.begin(), buf.begin() + h.n, all.begin());
    }
    fclose(f);
    *filled += h.n;
    return MPM_OK;
}

extern "C" int32_t mpm_load_state(MpmSolver* s, const char* path)
{
    if (!s || !path) return MPM_ERR_INVALID;
    std::vector<MpmParticle80> all;
    StateHeader h{};
    int64_t filled = 0;
    FILE* probe = fopen(path, "rb");
    if (probe) {
        fclose(probe);
        int rc = read_state_file(s, path, all, h, &filled);
        if (rc) return rc;
    } else {  // the per-rank files of a multi-GPU state
        int world = 0;
        for (int w = 1; w <= 64 && !world; ++w) {
            FILE* f = fopen((std::string(path) + ".rank0of" + std::to_string(w)).c_str(), "rb");
            if (f) { fclose(f); world = w; }
        }
        if (!world) return fail(s, MPM_ERR_INVALID, std::string("cannot open: ") + path + " (nor " + path + ".rank0of<world>)");
        for (int r = 0; r < world; ++r) {
            int rc = read_state_file(s, std::string(path) + ".rank" + std::to_string(r) + "of" + std::to_string(world), all, h, &filled);
            if (rc) return rc;
        }
    }
    if (filled != h.n_global) return fail(s, MPM_ERR_INVALID, "checkpoint files hold " + std::to_string(filled) + " of " + std::to_string(h.n_global) + " particles");
    int rc = mpm_upload_particles(s, all.data(), h.n_global);  // (validates the positions)
    if (rc) return rc;
    s->steps = h.steps;
    return MPM_OK;
}

// ---------------------------------------------------------------- step driver

struct PhaseTimer {
    MpmSolver* s;
    int phase;
    cudaEvent_t a = nullptr, b = nullptr;
    PhaseTimer(MpmSolver* s_, int phase_, size_t& cursor) : s(s_), phase(phase_)
    {
        if (s->timing != 1) return;
        while (s->ev.size() < cursor + 2) { cudaEvent_t e; cudaEventCreate(&e); s->ev.push_back(e); }
        a = s->ev[cursor]; b = s->ev[cursor + 1];
        cursor += 2;
        cudaEventRecord(a, s->stream);
    }
    ~PhaseTimer() { if (b) cudaEventRecord(b, s->stream); }
};

// clear / update restricted to the bounding box the binning found (cell path, one GPU, binning valid for this step)
static bool box_sweeps(const MpmSolver* s)
{
    static const bool off = getenv("MPM_DENSE_SWEEPS") != nullptr;
    return !off && s->path == MPM_PATH_CELL && s->sorted_valid && s->bin != nullptr;
}

static int run_phase(MpmSolver* s, int phase, size_t& cursor)
{
    const DevParams& P = s->dp;
    PhaseTimer t(s, phase, cursor);
    switch (phase) {
        case PH_SORT:
            if (s->path == MPM_PATH_TILED) { int rc = sort_particles(s); if (rc) return rc; }
            else if (s->path == MPM_PATH_CELL) { int rc = bin_particles(s); if (rc) return rc; }
            break;
        case PH_CLEAR:
            // Cell path: the binning of this step has run, so the blocks that can receive anything are known: clear and
            // update only their bounding box (29 % of the C4 grid at the start of the dam-break).  Multi-GPU slabs add the two
            // overlap planes at either end, whole: that is where the neighbours' halo contributions land.
            if (box_sweeps(s)) { launch_clear_box(P, s->grid, s->bin->box + 6, comm_halo_planes(s, 0), comm_halo_planes(s, 1), s->stream); s->launches += 1; }
            else CK(cudaMemsetAsync(s->grid, 0, 16 * s->ncells, s->stream));
            if (s->bin) s->bin->box_cleared = true;
            s->grid_raw = false;
            break;
        case PH_P2G1:
            if (s->path == MPM_PATH_TILED) { int rc = tiled_p2g1(s); if (rc) return rc; }
            else if (s->path == MPM_PATH_CELL) { int rc = cell_p2g1(s); if (rc) return rc; }
            else { launch_p2g1_ref(P, s->view(), s->n, s->grid, s->stream); s->launches += (s->n > 0); }
            break;
        case PH_P2G2:
            if (s->path == MPM_PATH_TILED) { int rc = tiled_p2g2(s); if (rc) return rc; }
            else if (s->path == MPM_PATH_CELL) { int rc = cell_p2g2(s); if (rc) return rc; }
            else { launch_p2g2_ref(P, s->view(), s->n, s->grid, s->stream); s->launches += (s->n > 0); }
            break;
        case PH_UPDATE:
            if (box_sweeps(s)) launch_update_box(P, s->grid, s->bin->box, comm_halo_planes(s, 0), comm_halo_planes(s, 1), s->stream);
            else launch_update_grid(P, s->grid, s->ncells, s->stream);
            s->launches += 1;
            s->grid_raw = false;
            break;
        case PH_G2P:
            if (s->path == MPM_PATH_TILED) { int rc = tiled_g2p(s); if (rc) return rc; }
            else if (s->path == MPM_PATH_CELL) { int rc = cell_g2p(s); if (rc) return rc; }
            else { launch_g2p_ref(P, s->view(), s->n, s->grid, s->orig_id, s->positions, s->stream); s->launches += (s->n > 0); }
            s->positions_valid = (s->path != MPM_PATH_CELL);  // the cell path produces the hand-off on demand
            break;
        default:
            return fail(s, MPM_ERR_INVALID, "unknown phase");
    }
    return MPM_OK;
}

static int collect_timing(MpmSolver* s, size_t used, const std::vector<int>& phases)
{
    if (!s->timing) return MPM_OK;
    CK(cudaStreamSynchronize(s->stream));
    if (s->timing != 1 || used == 0) return MPM_OK;
    for (size_t k = 0; k < phases.size(); ++k) {
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, s->ev[2 * k], s->ev[2 * k + 1]);
        s->ms_acc[phases[k]] += ms;
        if (phases[k] >= PH_EX_MASS && phases[k] <= PH_EX_MIG) s->ms_acc[PH_EXCHANGE] += ms;
    }
    return MPM_OK;
}

extern "C" int32_t mpm_step(MpmSolver* s, int32_t iterations)
{
    if (!s || iterations < 0) return MPM_ERR_INVALID;
    CK(cudaSetDevice(s->device));
    { int rc = comm_partition_for_step(s); if (rc) return rc; }
    if (s->timing) { for (double& v : s->ms_acc) v = 0; s->ms_step_acc = 0; s->timed_steps = 0; }
    size_t cursor = 0;
    std::vector<int> phases;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (s->timing) {
        size_t need = (size_t)iterations * 2 * 9 + 2;
        while (s->ev.size() < need) { cudaEvent_t e; cudaEventCreate(&e); s->ev.push_back(e); }
        e0 = s->ev[need - 2]; e1 = s->ev[need - 1];
        cudaEventRecord(e0, s->stream);
    }
    for (int it = 0; it < iterations; ++it) {
        int rc;
        const bool binned_path = s->path == MPM_PATH_TILED || s->path == MPM_PATH_CELL;
        if (binned_path && (!s->sorted_valid || s->steps_since_sort >= s->sort_interval)) {
            if ((rc = run_phase(s, PH_SORT, cursor))) return rc;
            phases.push_back(PH_SORT);
        }
        if ((rc = run_phase(s, PH_CLEAR, cursor))) return rc; phases.push_back(PH_CLEAR);
        if ((rc = run_phase(s, PH_P2G1, cursor))) return rc; phases.push_back(PH_P2G1);
        if (s->comm) { PhaseTimer t(s, PH_EX_MASS, cursor); phases.push_back(PH_EX_MASS); if ((rc = comm_exchange_halo(s, 0))) return rc; }
        if ((rc = run_phase(s, PH_P2G2, cursor))) return rc; phases.push_back(PH_P2G2);
        if (s->comm) { PhaseTimer t(s, PH_EX_MOM, cursor); phases.push_back(PH_EX_MOM); if ((rc = comm_exchange_halo(s, 1))) return rc; }
        // cell path: UpdateGrid is pointwise, so G2P can apply it while staging its tiles (MPM_FUSED_UPDATE=1).  Measured on
        // C4: 3.72 vs 3.75 ms/step at the start, 5.27 vs 5.22 ms after 100 steps -- no gain (the tile aprons redo 1.95x of
        // the update), so the separate kernel stays the default.
        static const bool fuse_update = getenv("MPM_FUSED_UPDATE") != nullptr;
        if (s->path == MPM_PATH_CELL && fuse_update) s->grid_raw = true;
        else { if ((rc = run_phase(s, PH_UPDATE, cursor))) return rc; phases.push_back(PH_UPDATE); }
        if ((rc = run_phase(s, PH_G2P, cursor))) return rc; phases.push_back(PH_G2P);
        if (s->comm) { PhaseTimer t(s, PH_EX_MIG, cursor); phases.push_back(PH_EX_MIG); if ((rc = comm_migrate(s))) return rc; }
        s->steps += 1;
        s->steps_since_sort += 1;
    }
    if (s->timing) {
        cudaEventRecord(e1, s->stream);
        int rc = collect_timing(s, cursor, phases);
        if (rc) return rc;
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, e0, e1);
        s->ms_step_acc = ms;
        s->timed_steps = iterations;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(s, MPM_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e));
    return MPM_OK;
}

extern "C" int32_t mpm_run_phase(MpmSolver* s, int32_t phase)
{
    if (!s) return MPM_ERR_INVALID;
    CK(cudaSetDevice(s->device));
    if (phase < 0 || phase > PH_SORT) return fail(s, MPM_ERR_INVALID, "phase must be 0..5");
    if (s->comm) return fail(s, MPM_ERR_STATE, "multi-GPU: phases cannot run one by one (halo exchanges sit between them); use mpm_step");
    if ((s->path == MPM_PATH_TILED || s->path == MPM_PATH_CELL) && phase != PH_SORT && phase != PH_CLEAR && phase != PH_UPDATE && !s->sorted_valid) {
        size_t c0 = 0; const int tm = s->timing; s->timing = 0;
        int rc = run_phase(s, PH_SORT, c0);
        s->timing = tm;
        if (rc) return rc;
    }
    const int tm = s->timing; s->timing = 0;
    size_t cursor = 0;
    int rc = run_phase(s, phase, cursor);
    s->timing = tm;
    if (rc) return rc;
    if (phase == PH_G2P) { s->steps += 1; s->steps_since_sort += 1; }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(s, MPM_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e));
    return MPM_OK;
}

extern "C" int32_t mpm_sync(MpmSolver* s)
{
    if (!s) return MPM_ERR_INVALID;
    CK(cudaSetDevice(s->device));
    CK(cudaStreamSynchronize(s->stream));
    int32_t f[2] = {0, 0};
    CK(cudaMemcpy(f, s->overflow_flag, sizeof(f), cudaMemcpyDeviceToHost));
    if (f[1]) {  // reported once, then cleared: the skipped particles are still there (unchanged) and the caller may go on
        CK(cudaMemset(s->overflow_flag + 1, 0, sizeof(int32_t)));
        return fail(s, MPM_ERR_DOMAIN, std::to_string(f[1]) + " particle position(s) non-finite or outside [1, R-1) were skipped since the last "
                                       "mpm_sync (the reference throws IndexOutOfRangeException here)");
    }
    if (s->hp.overflow_check && f[0]) return fail(s, MPM_ERR_OVERFLOW, "fixed-point grid accumulator overflowed int32 (lower fixed_point_mult)");
    return MPM_OK;
}

extern "C" int32_t mpm_get_positions(MpmSolver* s, float* dst4, int64_t cap, void** device_ptr, uint32_t* tex_width)
{
    if (!s) return MPM_ERR_INVALID;
    CK(cudaSetDevice(s->device));
    { int rc = comm_partition(s); if (rc) return rc; }
    // multi-GPU: the rank's local particles in slot order (G2P filled the array by global index instead)
    if ((!s->positions_valid || s->comm) && s->n > 0) {
        if (s->in_rec) launch_positions_rec(s->rview(), s->comm ? nullptr : s->orig_id, s->positions, s->n, s->stream);
        else launch_positions(s->view(), s->comm ? nullptr : s->orig_id, s->positions, s->n, s->stream);
        s->launches += 1;
        s->positions_valid = true;
    }
    if (device_ptr) *device_ptr = s->positions;
    if (tex_width) *tex_width = (uint32_t)sqrtf((float)s->n) + 1;  // MLSMPM3DFluidMultithreadGPU.cs:196
    if (dst4) {
        if (cap < s->n) return fail(s, MPM_ERR_INVALID, "destination too small");
        CK(cudaMemcpyAsync(dst4, s->positions, sizeof(float4) * s->n, cudaMemcpyDeviceToHost, s->stream));
        CK(cudaStreamSynchronize(s->stream));
    }
    return MPM_OK;
}

extern "C" int32_t mpm_export_positions(MpmSolver* s, int32_t* fd, uint64_t* bytes, uint32_t* tex_width)
{
    if (!s || !fd) return MPM_ERR_INVALID;
    CK(cudaSetDevice(s->device));
    if (!s->positions_mem.ptr) return fail(s, MPM_ERR_STATE, "the position array is not in an exportable allocation (driver without cuMemCreate / POSIX-FD handles)");
    int out = -1;
    if (!vmm_export_fd(s->positions_mem, &out)) return fail(s, MPM_ERR_CUDA, "cuMemExportToShareableHandle failed");
    *fd = out;
    if (bytes) *bytes = s->positions_mem.bytes;
    if (tex_width) *tex_width = (uint32_t)sqrtf((float)s->n) + 1;  // MLSMPM3DFluidMultithreadGPU.cs:196
    return MPM_OK;
}

static int32_t positions_async(MpmSolver* s, void* dst4, int64_t cap, bool q16)
{
    if (!s || !dst4) return MPM_ERR_INVALID;
    CK(cudaSetDevice(s->device));
    { int rc = comm_partition(s); if (rc) return rc; }
    if (cap < s->n) return fail(s, MPM_ERR_INVALID, "destination too small");
    if (!s->copy_stream) {  // first use: second device array, copy stream, events
        CK(cudaMalloc(&s->positions_b, sizeof(float4) * s->pitch));
        CK(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
        for (int k = 0; k < 2; ++k) {
            CK(cudaEventCreateWithFlags(&s->pos_ready[k], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&s->pos_copied[k], cudaEventDisableTiming));
            CK(cudaEventRecord(s->pos_copied[k], s->copy_stream));
        }
    }
    if (s->n == 0) return MPM_OK;
    const int b = s->pos_buf;
    float4* dev = b ? s->positions_b : s->positions;
    CK(cudaStreamWaitEvent(s->stream, s->pos_copied[b], 0));  // the copy that last read this device array is done
    if (q16) {
        const float scale[3] = {65535.0f / (float)s->hp.grid_size[0], 65535.0f / (float)s->hp.grid_size[1],
                                65535.0f / (float)std::max(s->hp.grid_size[2], 1)};
        if (s->in_rec) launch_positions_q16_rec(s->rview(), s->comm ? nullptr : s->orig_id, dev, s->n, scale, s->stream);
        else launch_positions_q16(s->view(), s->comm ? nullptr : s->orig_id, dev, s->n, scale, s->stream);
    } else if (s->in_rec) launch_positions_rec(s->rview(), s->comm ? nullptr : s->orig_id, dev, s->n, s->stream);
    else launch_positions(s->view(), s->comm ? nullptr : s->orig_id, dev, s->n, s->stream);
    s->launches += 1;
    CK(cudaEventRecord(s->pos_ready[b], s->stream));
    CK(cudaStreamWaitEvent(s->copy_stream, s->pos_ready[b], 0));
    CK(cudaMemcpyAsync(dst4, dev, (q16 ? sizeof(ushort4) : sizeof(float4)) * s->n, cudaMemcpyDeviceToHost, s->copy_stream));
    CK(cudaEventRecord(s->pos_copied[b], s->copy_stream));
    s->pos_buf ^= 1;
    if (b == 0) s->positions_valid = false;  // (the synchronous getter's array was just rewritten for this snapshot)
    return MPM_OK;
}

extern "C" int32_t mpm_get_positions_async(MpmSolver* s, float* dst4, int64_t cap) { return positions_async(s, dst4, cap, false); }
extern "C" int32_t mpm_get_positions_q16_async(MpmSolver* s, uint16_t* dst4, int64_t cap) { return positions_async(s, dst4, cap, true); }

extern "C" int32_t mpm_wait_positions(MpmSolver* s)
{
    if (!s) return MPM_ERR_INVALID;
    CK(cudaSetDevice(s->device));
    if (s->copy_stream) CK(cudaStreamSynchronize(s->copy_stream));
    return MPM_OK;
}

extern "C" int32_t mpm_num_particles(const MpmSolver* s, int64_t* n)
{
    if (!s || !n) return MPM_ERR_INVALID;
    *n = s->n;
    return MPM_OK;
}

extern "C" int32_t mpm_set_timing(MpmSolver* s, int32_t enabled)
{
    if (!s) return MPM_ERR_INVALID;
    s->timing = enabled == 2 ? 2 : (enabled != 0 ? 1 : 0);
    return MPM_OK;
}

namespace mpm { void comm_fill_stats(const MpmSolver* s, MpmStats* st); }

extern "C" int32_t mpm_get_stats(MpmSolver* s, MpmStats* st)
{
    if (!s || !st) return MPM_ERR_INVALID;
    if (s->comm) { cudaSetDevice(s->device); int rc = comm_partition(s); if (rc) return rc; }
    memset(st, 0, sizeof(*st));
    st->num_particles = s->n;
    st->local_particles = s->n;
    st->num_cells = s->ncells;
    st->steps = s->steps;
    st->kernel_launches = s->launches;
    st->kernel_path = s->path;
    st->world = 1;
    if (s->timed_steps > 0) {
        const double k = 1.0 / (double)s->timed_steps;
        st->ms_sort = (float)(s->ms_acc[PH_SORT] * k); st->ms_clear = (float)(s->ms_acc[PH_CLEAR] * k);
        st->ms_p2g1 = (float)(s->ms_acc[PH_P2G1] * k); st->ms_p2g2 = (float)(s->ms_acc[PH_P2G2] * k);
        st->ms_update = (float)(s->ms_acc[PH_UPDATE] * k); st->ms_g2p = (float)(s->ms_acc[PH_G2P] * k);
        st->ms_exchange = (float)(s->ms_acc[PH_EXCHANGE] * k); st->ms_step = (float)(s->ms_step_acc * k);
        st->ms_halo_mass = (float)(s->ms_acc[PH_EX_MASS] * k); st->ms_halo_momentum = (float)(s->ms_acc[PH_EX_MOM] * k);
        st->ms_migration = (float)(s->ms_acc[PH_EX_MIG] * k);
    }
    if (s->hp.overflow_check && s->overflow_flag) {
        cudaSetDevice(s->device);
        cudaStreamSynchronize(s->stream);
        cudaMemcpy(&st->overflow, s->overflow_flag, sizeof(int32_t), cudaMemcpyDeviceToHost);
    }
    if (s->bin && s->bin->far_n) {
        uint32_t f[4] = {0, 0, 0, 0};
        cudaSetDevice(s->device);
        cudaStreamSynchronize(s->stream);
        cudaMemcpy(f, s->bin->far_n, sizeof(f), cudaMemcpyDeviceToHost);
        st->unordered_binnings = f[1] + (f[3] ? 1 : 0);
        st->far_movers = f[2];
    }
    comm_fill_stats(s, st);
    return MPM_OK;
}

extern "C" int32_t mpm_debug_last_sort(MpmSolver* s, uint32_t* keys_before, uint32_t* perm, int64_t cap)
{
    if (!s) return MPM_ERR_INVALID;
    if (s->path != MPM_PATH_TILED && s->path != MPM_PATH_CELL) return fail(s, MPM_ERR_STATE, "the reference-shaped path does not bin particles");
    CK(cudaSetDevice(s->device));
    if (s->path == MPM_PATH_CELL) return bin_debug_last(s, keys_before, perm, cap);
    return sort_debug_last(s, keys_before, perm, cap);
}

extern "C" int32_t mpm_get_stream(MpmSolver* s, void** stream)
{
    if (!s || !stream) return MPM_ERR_INVALID;
    *stream = (void*)s->stream;
    return MPM_OK;
}

// pinned host memory for callers that want fast mpm_get_positions / downloads (C#: IntPtr)
extern "C" MPM_API int32_t mpm_host_alloc(int64_t bytes, void** out)
{
    if (!out || bytes <= 0) return MPM_ERR_INVALID;
    return cudaHostAlloc(out, (size_t)bytes, cudaHostAllocDefault) == cudaSuccess ? MPM_OK : MPM_ERR_CUDA;
}
extern "C" MPM_API int32_t mpm_host_free(void* p)
{
    return cudaFreeHost(p) == cudaSuccess ? MPM_OK : MPM_ERR_CUDA;
}
