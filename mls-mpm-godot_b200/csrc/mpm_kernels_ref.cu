// mpm_kernels_ref.cu -- the "reference-shaped" kernel path (MPM_PATH_REFERENCE): one thread per particle /
// per cell with global atomics, i.e. the launch shape of the reference's five compute shaders
// (mls-mpm/3d/fluid_multithread_gpu/compute_shaders/{clear_grid,p2g_1,p2g_2,update_grid,g2p}.glsl), but on
// SoA particle planes and with strict arithmetic.  It is the correctness anchor every tiled kernel is
// diffed against, and the only path for particle sets too small to be worth binning.
#include "mpm_kernels.h"
#include "mpm_particle_math.cuh"

namespace mpm {

template <bool FIXED>
__device__ __forceinline__ void cell_add(void* grid, int64_t ci, int ch, float val, const DevParams& P)
{
    if (FIXED) int_add_checked(reinterpret_cast<int*>(grid) + 4 * ci + ch, encode_fixed_checked(val, P), P);
    else atomicAdd(reinterpret_cast<float*>(grid) + 4 * ci + ch, val);
}

// ---------------------------------------------------------------- P2G_1  (p2g_1.glsl:40-94)
template <int DIM, bool FIXED>
__global__ void __launch_bounds__(256) k_p2g1_ref(DevParams P, ParticleView pv, int64_t n, void* grid)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ParticleIn p;
    p.px = pv.at(PX, i); p.py = pv.at(PY, i); p.pz = pv.at(PZ, i);
    p.vx = pv.at(VX, i); p.vy = pv.at(VY, i); p.vz = pv.at(VZ, i);
    p.m = pv.at(PM, i);
#pragma unroll
    for (int k = 0; k < 9; ++k) p.c[k] = pv.at(C0 + k, i);
    float wx[3], wy[3], wz[3] = {1.0f, 1.0f, 1.0f};
    const int cx = axis_weights(p.px, wx), cy = axis_weights(p.py, wy);
    const int cz = (DIM == 3) ? axis_weights(p.pz, wz) : 1;
    if (!stencil_in_grid(P, cx, cy, cz)) { flag_bad_particle(P); return; }
#pragma unroll
    for (int gx = 0; gx < 3; ++gx)
#pragma unroll
        for (int gy = 0; gy < 3; ++gy)
#pragma unroll
            for (int gz = 0; gz < (DIM == 3 ? 3 : 1); ++gz) {
                float weight = smul(wx[gx], wy[gy]);
                if (DIM == 3) weight = smul(weight, wz[gz]);
                const int nx = cx + gx - 1, ny = cy + gy - 1, nz = (DIM == 3) ? cz + gz - 1 : 0;
                const float dx = node_dist(nx, p.px), dy = node_dist(ny, p.py);
                const float dz = (DIM == 3) ? node_dist(nz, p.pz) : 0.0f;
                float mc, ox, oy, oz;
                p2g1_node<DIM>(p, weight, dx, dy, dz, mc, ox, oy, oz);
                const int64_t ci = cell_index(P, nx, ny, nz);
                cell_add<FIXED>(grid, ci, 3, mc, P);
                cell_add<FIXED>(grid, ci, 0, ox, P);
                cell_add<FIXED>(grid, ci, 1, oy, P);
                if (DIM == 3) cell_add<FIXED>(grid, ci, 2, oz, P);
            }
}

// ---------------------------------------------------------------- P2G_2  (p2g_2.glsl:52-154)
template <int DIM, bool FIXED>
__global__ void __launch_bounds__(256) k_p2g2_ref(DevParams P, ParticleView pv, int64_t n, void* grid)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float px = pv.at(PX, i), py = pv.at(PY, i), pz = pv.at(PZ, i), m = pv.at(PM, i);
    float c[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) c[k] = pv.at(C0 + k, i);
    float wx[3], wy[3], wz[3] = {1.0f, 1.0f, 1.0f};
    const int cx = axis_weights(px, wx), cy = axis_weights(py, wy);
    const int cz = (DIM == 3) ? axis_weights(pz, wz) : 1;
    if (!stencil_in_grid(P, cx, cy, cz)) return;  // (counted by P2G_1)

    float density = 0.0f;
#pragma unroll
    for (int gx = 0; gx < 3; ++gx)
#pragma unroll
        for (int gy = 0; gy < 3; ++gy)
#pragma unroll
            for (int gz = 0; gz < (DIM == 3 ? 3 : 1); ++gz) {
                float weight = smul(wx[gx], wy[gy]);
                if (DIM == 3) weight = smul(weight, wz[gz]);
                const int64_t ci = cell_index(P, cx + gx - 1, cy + gy - 1, (DIM == 3) ? cz + gz - 1 : 0);
                float gm;
                if (FIXED) gm = decode_fixed(reinterpret_cast<const int*>(grid)[4 * ci + 3], P.fmult);
                else gm = reinterpret_cast<const float*>(grid)[4 * ci + 3];
                density = sadd(density, smul(gm, weight));
            }
    float e[9];
    p2g2_stress<DIM>(P, c, m, density, e);
#pragma unroll
    for (int gx = 0; gx < 3; ++gx)
#pragma unroll
        for (int gy = 0; gy < 3; ++gy)
#pragma unroll
            for (int gz = 0; gz < (DIM == 3 ? 3 : 1); ++gz) {
                float weight = smul(wx[gx], wy[gy]);
                if (DIM == 3) weight = smul(weight, wz[gz]);
                const int nx = cx + gx - 1, ny = cy + gy - 1, nz = (DIM == 3) ? cz + gz - 1 : 0;
                const float dx = node_dist(nx, px), dy = node_dist(ny, py);
                const float dz = (DIM == 3) ? node_dist(nz, pz) : 0.0f;
                float ox, oy, oz;
                p2g2_node<DIM>(e, weight, dx, dy, dz, ox, oy, oz);
                const int64_t ci = cell_index(P, nx, ny, nz);
                cell_add<FIXED>(grid, ci, 0, ox, P);
                cell_add<FIXED>(grid, ci, 1, oy, P);
                if (DIM == 3) cell_add<FIXED>(grid, ci, 2, oz, P);
            }
}

// ---------------------------------------------------------------- UpdateGrid  (update_grid.glsl:36-74)
// fixed-point cell: decode, v = p / m, gravity on y, wall-normal component zeroed (ox / oy / oz), re-encode
__device__ __forceinline__ void update_cell_fixed(const DevParams& P, int4& c, bool ox, bool oy, bool oz)
{
    const float mm = decode_fixed(c.w, P.fmult);
    const float vx = sdiv(decode_fixed(c.x, P.fmult), mm);
    const float vy = sdiv(decode_fixed(c.y, P.fmult), mm);
    const float vz = sdiv(decode_fixed(c.z, P.fmult), mm);
    c.x = ox ? 0 : encode_fixed(vx, P.fmult);
    c.y = oy ? 0 : encode_fixed(sadd(vy, smul(P.dt, P.gravity)), P.fmult);
    c.z = oz ? 0 : encode_fixed(vz, P.fmult);
}

// One thread per local cell; a whole 16-B cell per thread keeps the access a coalesced 128-bit stream.
template <int DIM, bool FIXED>
__global__ void __launch_bounds__(256) k_update_grid(DevParams P, void* grid, int64_t ncells)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncells) return;
    int x, y, z;
    if (DIM == 3) { z = (int)(i % P.Rz); y = (int)(i / P.Rz % P.Ry); x = (int)(i / P.Rz / P.Ry) + P.gx0; }
    else { y = (int)(i % P.Ry); x = (int)(i / P.Ry) + P.gx0; z = 2; }
    const int hi = P.bc_hi_off;
    const bool ox = (x < 2 || x > P.Rx - hi), oy = (y < 2 || y > P.Ry - hi);
    const bool oz = (DIM == 3) && (z < 2 || z > P.Rz - hi);
    if (FIXED) {
        int4 c = reinterpret_cast<int4*>(grid)[i];
        if (c.w > 0) {
            update_cell_fixed(P, c, ox, oy, oz);
            reinterpret_cast<int4*>(grid)[i] = c;
        }
    } else {
        float4 c = reinterpret_cast<float4*>(grid)[i];
        if (c.w > 0.0f) {
            c.x = sadd(sdiv(c.x, c.w), smul(P.dt, 0.0f));
            c.y = sadd(sdiv(c.y, c.w), smul(P.dt, P.gravity));
            c.z = sadd(sdiv(c.z, c.w), smul(P.dt, 0.0f));
            if (P.bc_mode == 0) {
                if (ox) c.x = 0.0f;
                if (oy) c.y = 0.0f;
                if (oz) c.z = 0.0f;
            } else {
                const float f = P.bc_friction;
                if (ox) { c.y = smul(f, c.y); c.z = smul(f, c.z); c.x = 0.0f; }
                if (oy) { c.x = smul(f, c.x); c.z = smul(f, c.z); c.y = 0.0f; }
                if (oz) { c.x = smul(f, c.x); c.y = smul(f, c.y); c.z = 0.0f; }
            }
            reinterpret_cast<float4*>(grid)[i] = c;
        }
    }
}

// ---- the same two grid sweeps restricted to a box of cells (cell path, 3D fixed-point grid): the binning knows which grid
// blocks hold particles, and nothing outside their bounding box (+ the one-node apron P2G writes) is ever non-zero.
// box = {x0, x1, y0, y1, z0, z1}, half-open, global cell coordinates, already clamped to the local grid; it lives on the
// device (the scan of the binning writes it), so the launch covers the whole grid in strips of BOX_ROWS (x, y) rows and
// strips outside return at once.
constexpr int BOX_ROWS = 8;  // (x, y) rows per CTA: one row per CTA made the launch itself (65536 tiny CTAs on C4) cost 35 us

// (multi-GPU slabs: the first halo_lo and the last halo_hi stored planes are the overlap planes the neighbours add their
// contributions to, anywhere in y and z: they are swept whole, the planes between them inside the rank's own box)
__global__ void __launch_bounds__(256) k_clear_box(DevParams P, int4* __restrict__ grid, const int* __restrict__ box, int halo_lo, int halo_hi)
{
    pdl_prologue();
    const int x = P.gx0 + blockIdx.y;
    const bool whole = (int)blockIdx.y < halo_lo || (int)blockIdx.y >= P.nxl - halo_hi;
    if (!whole && (x < box[0] || x >= box[1])) return;
    const int by0 = whole ? 0 : box[2], by1 = whole ? P.Ry : box[3];
    const int y0 = max((int)blockIdx.x * BOX_ROWS, by0), y1 = min(min((int)blockIdx.x * BOX_ROWS + BOX_ROWS, by1), P.Ry);
    const int z0 = whole ? 0 : box[4], z1 = whole ? P.Rz : box[5];
    for (int y = y0; y < y1; ++y) {
        int4* row = grid + ((int64_t)blockIdx.y * P.Ry + y) * P.Rz;
        for (int z = z0 + threadIdx.x; z < z1; z += blockDim.x) row[z] = make_int4(0, 0, 0, 0);
    }
}

__global__ void __launch_bounds__(256) k_update_box(DevParams P, int4* __restrict__ grid, const int* __restrict__ box, int halo_lo, int halo_hi)
{
    pdl_prologue();
    const int x = P.gx0 + blockIdx.y;
    const bool whole = (int)blockIdx.y < halo_lo || (int)blockIdx.y >= P.nxl - halo_hi;
    if (!whole && (x < box[0] || x >= box[1])) return;
    const int by0 = whole ? 0 : box[2], by1 = whole ? P.Ry : box[3];
    const int y0 = max((int)blockIdx.x * BOX_ROWS, by0), y1 = min(min((int)blockIdx.x * BOX_ROWS + BOX_ROWS, by1), P.Ry);
    const int z0 = whole ? 0 : box[4], z1 = whole ? P.Rz : box[5];
    const int hi = P.bc_hi_off;
    const bool ox = (x < 2 || x > P.Rx - hi);
    for (int y = y0; y < y1; ++y) {
        int4* row = grid + ((int64_t)blockIdx.y * P.Ry + y) * P.Rz;
        const bool oy = (y < 2 || y > P.Ry - hi);
        for (int z = z0 + threadIdx.x; z < z1; z += blockDim.x) {
            int4 c = row[z];
            if (c.w > 0) {
                update_cell_fixed(P, c, ox, oy, z < 2 || z > P.Rz - hi);
                row[z] = c;
            }
        }
    }
}

// ---------------------------------------------------------------- G2P  (g2p.glsl:52-152)
template <int DIM, bool FIXED>
__global__ void __launch_bounds__(256) k_g2p_ref(DevParams P, ParticleView pv, int64_t n, const void* grid,
                                                 const uint32_t* orig_id, float4* positions)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float old[3] = {pv.at(PX, i), pv.at(PY, i), pv.at(PZ, i)};
    float wx[3], wy[3], wz[3] = {1.0f, 1.0f, 1.0f};
    const int cx = axis_weights(old[0], wx), cy = axis_weights(old[1], wy);
    const int cz = (DIM == 3) ? axis_weights(old[2], wz) : 1;
    if (!stencil_in_grid(P, cx, cy, cz)) return;  // (counted by P2G_1; the particle stays as it is)
    float B[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, v[3] = {0, 0, 0};
#pragma unroll
    for (int gx = 0; gx < 3; ++gx)
#pragma unroll
        for (int gy = 0; gy < 3; ++gy)
#pragma unroll
            for (int gz = 0; gz < (DIM == 3 ? 3 : 1); ++gz) {
                float weight = smul(wx[gx], wy[gy]);
                if (DIM == 3) weight = smul(weight, wz[gz]);
                const int nx = cx + gx - 1, ny = cy + gy - 1, nz = (DIM == 3) ? cz + gz - 1 : 0;
                const float dx = node_dist(nx, old[0]), dy = node_dist(ny, old[1]);
                const float dz = (DIM == 3) ? node_dist(nz, old[2]) : 0.0f;
                const int64_t ci = cell_index(P, nx, ny, nz);
                float gvx, gvy, gvz;
                if (FIXED) {
                    const int4 c = reinterpret_cast<const int4*>(grid)[ci];
                    gvx = decode_fixed(c.x, P.fmult); gvy = decode_fixed(c.y, P.fmult); gvz = decode_fixed(c.z, P.fmult);
                } else {
                    const float4 c = reinterpret_cast<const float4*>(grid)[ci];
                    gvx = c.x; gvy = c.y; gvz = c.z;
                }
                g2p_node<DIM>(gvx, gvy, gvz, weight, dx, dy, dz, B, v);
            }
    float np[3], c[9];
    g2p_finish<DIM>(P, old, B, v, np, c);
    pv.at(PX, i) = np[0]; pv.at(PY, i) = np[1]; pv.at(PZ, i) = np[2];
    pv.at(VX, i) = v[0]; pv.at(VY, i) = v[1]; pv.at(VZ, i) = v[2];
#pragma unroll
    for (int k = 0; k < 9; ++k) pv.at(C0 + k, i) = c[k];
    if (positions) {  // particle_pos texture (g2p.glsl:149-150), original index order
        const float len = __fsqrt_rn(sadd(sadd(smul(v[0], v[0]), smul(v[1], v[1])), smul(v[2], v[2])));
        positions[orig_id ? orig_id[i] : (uint32_t)i] = make_float4(np[0], np[1], np[2], len);
    }
}

// ---------------------------------------------------------------- launchers
static inline unsigned blocks_for(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

#define DISPATCH_DIM_FIXED(KERNEL, ...)                                              \
    do {                                                                             \
        if (P.dim == 3) {                                                            \
            if (P.grid_mode) KERNEL<3, true> __VA_ARGS__; else KERNEL<3, false> __VA_ARGS__; \
        } else {                                                                     \
            if (P.grid_mode) KERNEL<2, true> __VA_ARGS__; else KERNEL<2, false> __VA_ARGS__; \
        }                                                                            \
    } while (0)

void launch_p2g1_ref(const DevParams& P, ParticleView pv, int64_t n, void* grid, cudaStream_t st)
{
    if (n <= 0) return;
    DISPATCH_DIM_FIXED(k_p2g1_ref, <<<blocks_for(n, 256), 256, 0, st>>>(P, pv, n, grid));
}
void launch_p2g2_ref(const DevParams& P, ParticleView pv, int64_t n, void* grid, cudaStream_t st)
{
    if (n <= 0) return;
    DISPATCH_DIM_FIXED(k_p2g2_ref, <<<blocks_for(n, 256), 256, 0, st>>>(P, pv, n, grid));
}
void launch_update_grid(const DevParams& P, void* grid, int64_t ncells, cudaStream_t st)
{
    DISPATCH_DIM_FIXED(k_update_grid, <<<blocks_for(ncells, 256), 256, 0, st>>>(P, grid, ncells));
}
void launch_clear_box(const DevParams& P, void* grid, const int* box, int halo_lo, int halo_hi, cudaStream_t st)
{
    launch_pdl<PDL_SWEEP>(k_clear_box, dim3((unsigned)((P.Ry + BOX_ROWS - 1) / BOX_ROWS), (unsigned)P.nxl), dim3(256), 0, st, P, reinterpret_cast<int4*>(grid), box, halo_lo, halo_hi);
}
void launch_update_box(const DevParams& P, void* grid, const int* box, int halo_lo, int halo_hi, cudaStream_t st)
{
    launch_pdl<PDL_SWEEP>(k_update_box, dim3((unsigned)((P.Ry + BOX_ROWS - 1) / BOX_ROWS), (unsigned)P.nxl), dim3(256), 0, st, P, reinterpret_cast<int4*>(grid), box, halo_lo, halo_hi);
}
void launch_g2p_ref(const DevParams& P, ParticleView pv, int64_t n, const void* grid, const uint32_t* orig_id,
                    float4* positions, cudaStream_t st)
{
    if (n <= 0) return;
    DISPATCH_DIM_FIXED(k_g2p_ref, <<<blocks_for(n, 256), 256, 0, st>>>(P, pv, n, grid, orig_id, positions));
}

}  // namespace mpm
