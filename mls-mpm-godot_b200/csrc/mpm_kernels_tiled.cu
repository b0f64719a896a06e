// mpm_kernels_tiled.cu -- the B200 kernel path (MPM_PATH_TILED) for the 3D int32 fixed-point grid.
//
// Particles are binned by BxBxB grid block (mpm_sort.cu); one CTA owns one block's contiguous particle
// range and stages the block's (B+2)^3-node grid tile (block + 1-node apron = everything a quadratic
// B-spline stencil of a particle in the block can touch) in shared memory:
//   P2G_1 : smem tile of int32 accumulators (mass, momentum), native ATOMS.ADD, then one global RED per
//           touched node and channel -> global atomics drop from 108/particle to ~1-2/particle.
//   P2G_2 : mass tile loaded once and decoded once per node; momentum tile accumulated as above.
//   G2P   : velocity tile loaded and decoded once per node, 27-node gather from smem, fused advection,
//           clamp, interaction, predictive wall and the (x,y,z,|v|) hand-off write.
// A particle whose base cell has drifted out of its CTA's block since the last bin phase (sort_interval>1)
// takes a slow path straight to global memory, so the result never depends on how stale the binning is.
// Arithmetic is the strict path of mpm_particle_math.cuh: bit-identical to the reference-shaped kernels.
#include "mpm_kernels.h"
#include "mpm_particle_math.cuh"
#include "mpm_solver.h"
#include "mpm_tile.cuh"

#include <type_traits>

namespace mpm {

const uint32_t* sort_block_start(const MpmSolver* s);
void sort_geometry(const MpmSolver* s, int& B, int& nbx, int& nby, int& nbz, int64_t& nblocks);

// ---------------------------------------------------------------- P2G_1
template <int B>
__global__ void __launch_bounds__(TILED_THREADS) k_p2g1_tiled(DevParams P, TileGeom g, ParticleView pv,
                                                              const uint32_t* __restrict__ block_start, int* __restrict__ grid)
{
    using TL = Tile<B>;
    __shared__ int tile[4][TL::WORDS];
    const int b = blockIdx.x;
    const uint32_t s0 = block_start[b], s1 = block_start[b + 1];
    if (s0 == s1) return;
    TL tl; tl.init(g, b);
    for (int k = threadIdx.x; k < 4 * TL::WORDS; k += TILED_THREADS) (&tile[0][0])[k] = 0;
    __syncthreads();
    for (uint32_t i = s0 + threadIdx.x; i < s1; i += TILED_THREADS) {
        ParticleIn p;
        p.px = pv.at(PX, i); p.py = pv.at(PY, i); p.pz = pv.at(PZ, i);
        p.vx = pv.at(VX, i); p.vy = pv.at(VY, i); p.vz = pv.at(VZ, i);
        p.m = pv.at(PM, i);
#pragma unroll
        for (int k = 0; k < 9; ++k) p.c[k] = pv.at(C0 + k, i);
        float wx[3], wy[3], wz[3];
        const int cx = axis_weights(p.px, wx), cy = axis_weights(p.py, wy), cz = axis_weights(p.pz, wz);
        int base;
        const bool in_block = tl.stencil_base(cx, cy, cz, base);
        auto scatter = [&](auto in_tile) {
        constexpr bool inside = decltype(in_tile)::value;
#pragma unroll
        for (int gx = 0; gx < 3; ++gx)
#pragma unroll
            for (int gy = 0; gy < 3; ++gy)
#pragma unroll
                for (int gz = 0; gz < 3; ++gz) {
                    const float weight = smul(smul(wx[gx], wy[gy]), wz[gz]);
                    const int nx = cx + gx - 1, ny = cy + gy - 1, nz = cz + gz - 1;
                    float mc, ox, oy, oz;
                    p2g1_node<3>(p, weight, node_dist(nx, p.px), node_dist(ny, p.py), node_dist(nz, p.pz), mc, ox, oy, oz);
                    const int em = encode_fixed_checked(mc, P), ex = encode_fixed_checked(ox, P), ey = encode_fixed_checked(oy, P),
                              ez = encode_fixed_checked(oz, P);
                    if constexpr (inside) {
                        const int idx = base + gx * TL::PX + gy * TL::PY + gz;
                        int_add_checked(&tile[3][idx], em, P); int_add_checked(&tile[0][idx], ex, P);
                        int_add_checked(&tile[1][idx], ey, P); int_add_checked(&tile[2][idx], ez, P);
                    } else {
                        int* c = grid + 4 * cell_index(P, nx, ny, nz);
                        int_add_checked(c + 3, em, P); int_add_checked(c + 0, ex, P); int_add_checked(c + 1, ey, P); int_add_checked(c + 2, ez, P);
                    }
                }
        };
        if (in_block) scatter(std::true_type{});
        else if (stencil_in_grid(P, cx, cy, cz)) scatter(std::false_type{});
        else flag_bad_particle(P);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < TL::NODES; k += TILED_THREADS) {
        int idx; int64_t ci;
        const bool ok = tl.node(P, k, idx, ci);
        const int vx = tile[0][idx], vy = tile[1][idx], vz = tile[2][idx], m = tile[3][idx];
        if (!ok || (vx | vy | vz | m) == 0) continue;
        int* c = grid + 4 * ci;
        if (vx) int_add_checked(c + 0, vx, P);
        if (vy) int_add_checked(c + 1, vy, P);
        if (vz) int_add_checked(c + 2, vz, P);
        if (m) int_add_checked(c + 3, m, P);
    }
}

// ---------------------------------------------------------------- P2G_2
template <int B>
__global__ void __launch_bounds__(TILED_THREADS) k_p2g2_tiled(DevParams P, TileGeom g, ParticleView pv,
                                                              const uint32_t* __restrict__ block_start, int* __restrict__ grid)
{
    using TL = Tile<B>;
    __shared__ int tile[3][TL::WORDS];
    __shared__ float tmass[TL::WORDS];
    const int b = blockIdx.x;
    const uint32_t s0 = block_start[b], s1 = block_start[b + 1];
    if (s0 == s1) return;
    TL tl; tl.init(g, b);
    for (int k = threadIdx.x; k < 3 * TL::WORDS; k += TILED_THREADS) (&tile[0][0])[k] = 0;
    for (int k = threadIdx.x; k < TL::NODES; k += TILED_THREADS) {
        int idx; int64_t ci;
        const bool ok = tl.node(P, k, idx, ci);
        tmass[idx] = ok ? decode_fixed(grid[4 * ci + 3], P.fmult) : 0.0f;
    }
    __syncthreads();
    for (uint32_t i = s0 + threadIdx.x; i < s1; i += TILED_THREADS) {
        const float px = pv.at(PX, i), py = pv.at(PY, i), pz = pv.at(PZ, i), m = pv.at(PM, i);
        float c[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) c[k] = pv.at(C0 + k, i);
        float wx[3], wy[3], wz[3];
        const int cx = axis_weights(px, wx), cy = axis_weights(py, wy), cz = axis_weights(pz, wz);
        int base;
        const bool in_block = tl.stencil_base(cx, cy, cz, base);
        float density = 0.0f;
        auto gather = [&](auto in_tile) {
        constexpr bool inside = decltype(in_tile)::value;
#pragma unroll
        for (int gx = 0; gx < 3; ++gx)
#pragma unroll
            for (int gy = 0; gy < 3; ++gy)
#pragma unroll
                for (int gz = 0; gz < 3; ++gz) {
                    const float weight = smul(smul(wx[gx], wy[gy]), wz[gz]);
                    float gm;
                    if constexpr (inside) gm = tmass[base + gx * TL::PX + gy * TL::PY + gz];
                    else gm = decode_fixed(grid[4 * cell_index(P, cx + gx - 1, cy + gy - 1, cz + gz - 1) + 3], P.fmult);
                    density = sadd(density, smul(gm, weight));
                }
        };
        if (in_block) gather(std::true_type{});
        else if (stencil_in_grid(P, cx, cy, cz)) gather(std::false_type{});
        else continue;  // (a position outside the grid: counted by P2G_1, the particle stays as it is)
        float e[9];
        p2g2_stress<3>(P, c, m, density, e);
        auto scatter = [&](auto in_tile) {
        constexpr bool inside = decltype(in_tile)::value;
#pragma unroll
        for (int gx = 0; gx < 3; ++gx)
#pragma unroll
            for (int gy = 0; gy < 3; ++gy)
#pragma unroll
                for (int gz = 0; gz < 3; ++gz) {
                    const float weight = smul(smul(wx[gx], wy[gy]), wz[gz]);
                    const int nx = cx + gx - 1, ny = cy + gy - 1, nz = cz + gz - 1;
                    float ox, oy, oz;
                    p2g2_node<3>(e, weight, node_dist(nx, px), node_dist(ny, py), node_dist(nz, pz), ox, oy, oz);
                    const int ex = encode_fixed_checked(ox, P), ey = encode_fixed_checked(oy, P), ez = encode_fixed_checked(oz, P);
                    if constexpr (inside) {
                        const int idx = base + gx * TL::PX + gy * TL::PY + gz;
                        int_add_checked(&tile[0][idx], ex, P); int_add_checked(&tile[1][idx], ey, P); int_add_checked(&tile[2][idx], ez, P);
                    } else {
                        int* cc = grid + 4 * cell_index(P, nx, ny, nz);
                        int_add_checked(cc + 0, ex, P); int_add_checked(cc + 1, ey, P); int_add_checked(cc + 2, ez, P);
                    }
                }
        };
        if (in_block) scatter(std::true_type{});
        else if (stencil_in_grid(P, cx, cy, cz)) scatter(std::false_type{});
        else flag_bad_particle(P);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < TL::NODES; k += TILED_THREADS) {
        int idx; int64_t ci;
        const bool ok = tl.node(P, k, idx, ci);
        const int vx = tile[0][idx], vy = tile[1][idx], vz = tile[2][idx];
        if (!ok || (vx | vy | vz) == 0) continue;
        int* c = grid + 4 * ci;
        if (vx) int_add_checked(c + 0, vx, P);
        if (vy) int_add_checked(c + 1, vy, P);
        if (vz) int_add_checked(c + 2, vz, P);
    }
}

// ---------------------------------------------------------------- G2P
template <int B>
__global__ void __launch_bounds__(TILED_THREADS) k_g2p_tiled(DevParams P, TileGeom g, ParticleView pv,
                                                             const uint32_t* __restrict__ block_start, const int4* __restrict__ grid,
                                                             const uint32_t* __restrict__ orig_id, float4* __restrict__ positions)
{
    using TL = Tile<B>;
    __shared__ float tv[3][TL::WORDS];
    const int b = blockIdx.x;
    const uint32_t s0 = block_start[b], s1 = block_start[b + 1];
    if (s0 == s1) return;
    TL tl; tl.init(g, b);
    for (int k = threadIdx.x; k < TL::NODES; k += TILED_THREADS) {
        int idx; int64_t ci;
        float vx = 0.0f, vy = 0.0f, vz = 0.0f;
        if (tl.node(P, k, idx, ci)) {
            const int4 c = grid[ci];
            vx = decode_fixed(c.x, P.fmult); vy = decode_fixed(c.y, P.fmult); vz = decode_fixed(c.z, P.fmult);
        }
        tv[0][idx] = vx; tv[1][idx] = vy; tv[2][idx] = vz;
    }
    __syncthreads();
    for (uint32_t i = s0 + threadIdx.x; i < s1; i += TILED_THREADS) {
        const float old[3] = {pv.at(PX, i), pv.at(PY, i), pv.at(PZ, i)};
        float wx[3], wy[3], wz[3];
        const int cx = axis_weights(old[0], wx), cy = axis_weights(old[1], wy), cz = axis_weights(old[2], wz);
        int base;
        const bool in_block = tl.stencil_base(cx, cy, cz, base);
        float Bm[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, v[3] = {0, 0, 0};
        auto gather = [&](auto in_tile) {
        constexpr bool inside = decltype(in_tile)::value;
#pragma unroll
        for (int gx = 0; gx < 3; ++gx)
#pragma unroll
            for (int gy = 0; gy < 3; ++gy)
#pragma unroll
                for (int gz = 0; gz < 3; ++gz) {
                    const float weight = smul(smul(wx[gx], wy[gy]), wz[gz]);
                    const int nx = cx + gx - 1, ny = cy + gy - 1, nz = cz + gz - 1;
                    float gvx, gvy, gvz;
                    if constexpr (inside) {
                        const int idx = base + gx * TL::PX + gy * TL::PY + gz;
                        gvx = tv[0][idx]; gvy = tv[1][idx]; gvz = tv[2][idx];
                    } else {
                        const int4 c = grid[cell_index(P, nx, ny, nz)];
                        gvx = decode_fixed(c.x, P.fmult); gvy = decode_fixed(c.y, P.fmult); gvz = decode_fixed(c.z, P.fmult);
                    }
                    g2p_node<3>(gvx, gvy, gvz, weight, node_dist(nx, old[0]), node_dist(ny, old[1]), node_dist(nz, old[2]), Bm, v);
                }
        };
        if (in_block) gather(std::true_type{});
        else if (stencil_in_grid(P, cx, cy, cz)) gather(std::false_type{});
        else continue;  // (a position outside the grid: counted by P2G_1, the particle stays as it is)
        float np[3], c[9];
        g2p_finish<3>(P, old, Bm, v, np, c);
        pv.at(PX, i) = np[0]; pv.at(PY, i) = np[1]; pv.at(PZ, i) = np[2];
        pv.at(VX, i) = v[0]; pv.at(VY, i) = v[1]; pv.at(VZ, i) = v[2];
#pragma unroll
        for (int k = 0; k < 9; ++k) pv.at(C0 + k, i) = c[k];
        const float len = __fsqrt_rn(sadd(sadd(smul(v[0], v[0]), smul(v[1], v[1])), smul(v[2], v[2])));
        positions[orig_id[i]] = make_float4(np[0], np[1], np[2], len);
    }
}

// ---------------------------------------------------------------- host side
static int check_supported(MpmSolver* s)
{
    if (s->dp.dim != 3 || s->dp.grid_mode != MPM_GRID_FIXED) {
        s->err = "the tiled path implements the 3D fixed-point grid; use MPM_PATH_REFERENCE (or AUTO) for 2D / float grids";
        return MPM_ERR_INVALID;
    }
    return MPM_OK;
}

#define LAUNCH_TILED(KERNEL, ...)                                                                         \
    do {                                                                                                  \
        int B, nbx, nby, nbz; int64_t nblocks;                                                            \
        sort_geometry(s, B, nbx, nby, nbz, nblocks);                                                      \
        TileGeom g{nby, nbz, s->dp.gx0 + (s->comm ? 1 : 0)};                                              \
        if (B == 8) KERNEL<8><<<(unsigned)nblocks, TILED_THREADS, 0, s->stream>>>(s->dp, g, __VA_ARGS__); \
        else KERNEL<4><<<(unsigned)nblocks, TILED_THREADS, 0, s->stream>>>(s->dp, g, __VA_ARGS__);        \
        s->launches += 1;                                                                                 \
    } while (0)

// fast-math variants (mpm_kernels_fast.cu)
void fast_p2g1(MpmSolver* s);
void fast_p2g2(MpmSolver* s);
void fast_g2p(MpmSolver* s);

int tiled_p2g1(MpmSolver* s)
{
    int rc = check_supported(s);
    if (rc) return rc;
    if (s->n == 0) return MPM_OK;
    if (s->hp.math_mode == MPM_MATH_FAST) { fast_p2g1(s); return MPM_OK; }
    LAUNCH_TILED(k_p2g1_tiled, s->view(), sort_block_start(s), reinterpret_cast<int*>(s->grid));
    return MPM_OK;
}
int tiled_p2g2(MpmSolver* s)
{
    int rc = check_supported(s);
    if (rc) return rc;
    if (s->n == 0) return MPM_OK;
    if (s->hp.math_mode == MPM_MATH_FAST) { fast_p2g2(s); return MPM_OK; }
    LAUNCH_TILED(k_p2g2_tiled, s->view(), sort_block_start(s), reinterpret_cast<int*>(s->grid));
    return MPM_OK;
}
int tiled_g2p(MpmSolver* s)
{
    int rc = check_supported(s);
    if (rc) return rc;
    if (s->n == 0) return MPM_OK;
    if (s->hp.math_mode == MPM_MATH_FAST) { fast_g2p(s); return MPM_OK; }
    LAUNCH_TILED(k_g2p_tiled, s->view(), sort_block_start(s), reinterpret_cast<const int4*>(s->grid), s->orig_id, s->positions);
    return MPM_OK;
}

}  // namespace mpm
