// mpm_kernels_fast.cu -- MPM_MATH_FAST variants of the tiled kernels (3D, int32 fixed-point grid).
//
// Same algorithm, same tiles and the same slow path as mpm_kernels_tiled.cu, but the per-particle arithmetic is
// re-associated for instruction count (ncu on B200, profiles/r1: the strict kernels issue 33 / 52 / 38
// warp-instructions per particle and are issue-bound, not HBM-bound):
//   * FMA contraction, stencil distances and C*d / E*d partial sums hoisted per axis,
//   * the fixed-point scale folded into the particle weight (one multiply instead of four per node),
//   * G2P by sum factorisation: z-sums, then y, then x (279 FMA per particle instead of ~780 mul/add),
//   * reciprocal-multiply instead of IEEE division when decoding the grid tile, fp32 EOS power.
// Results differ from the strict path by rounding only; tests/test_parity_gpu.py states the tolerance.
//
// The tile is padded so that 32 consecutive cells of a block (z fastest, then y, then x) fall into 32 distinct
// shared-memory banks: row pitch PY = 24 words (8-cell rows advance the bank by -8), plane pitch PX = 256.
#include "mpm_kernels.h"
#include "mpm_particle_math.cuh"
#include "mpm_solver.h"
#include "mpm_tile.cuh"

#include <type_traits>

namespace mpm {

const uint32_t* sort_block_start(const MpmSolver* s);
void sort_geometry(const MpmSolver* s, int& B, int& nbx, int& nby, int& nbz, int64_t& nblocks);

// weights and node distances of one axis: d[g] = (c + g - 1 - p) + 0.5 = (g - 1) - cd
__device__ __forceinline__ int fast_axis(float p, float w[3], float d[3])
{
    const int c = __float2int_rz(p);
    const float cd = (p - (float)c) - 0.5f;
    const float a = 0.5f - cd, b = 0.5f + cd;
    w[0] = 0.5f * a * a;
    w[1] = 0.75f - cd * cd;
    w[2] = 0.5f * b * b;
    d[0] = -1.0f - cd; d[1] = -cd; d[2] = 1.0f - cd;
    return c;
}

// ---------------------------------------------------------------- P2G_1
template <int B>
__global__ void __launch_bounds__(TILED_THREADS) k_p2g1_fast(DevParams P, TileGeom g, ParticleView pv,
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
        const float px = pv.at(PX, i), py = pv.at(PY, i), pz = pv.at(PZ, i);
        const float vx = pv.at(VX, i), vy = pv.at(VY, i), vz = pv.at(VZ, i);
        const float ms = pv.at(PM, i) * P.fmult;  // mass in fixed-point units
        float c[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) c[k] = pv.at(C0 + k, i);
        float wx[3], wy[3], wz[3], dx[3], dy[3], dz[3];
        const int cx = fast_axis(px, wx, dx), cy = fast_axis(py, wy, dy), cz = fast_axis(pz, wz, dz);
        int base;
        const bool in_block = tl.stencil_base(cx, cy, cz, base);
        float wzs[3] = {wz[0] * ms, wz[1] * ms, wz[2] * ms};
        // the whole stencil is either in the tile (shared-memory atomics) or not (global atomics): decide once
        auto scatter = [&](auto in_tile) {
        constexpr bool inside = decltype(in_tile)::value;
#pragma unroll
        for (int gx = 0; gx < 3; ++gx) {
            const float ax = fmaf(c[0], dx[gx], vx), ay = fmaf(c[1], dx[gx], vy), az = fmaf(c[2], dx[gx], vz);
#pragma unroll
            for (int gy = 0; gy < 3; ++gy) {
                const float bx = fmaf(c[3], dy[gy], ax), by = fmaf(c[4], dy[gy], ay), bz = fmaf(c[5], dy[gy], az);
                const float wxy = wx[gx] * wy[gy];
#pragma unroll
                for (int gz = 0; gz < 3; ++gz) {
                    const float mc = wxy * wzs[gz];
                    const int em = f2i_checked(mc, P);
                    const int ex = f2i_checked(mc * fmaf(c[6], dz[gz], bx), P);
                    const int ey = f2i_checked(mc * fmaf(c[7], dz[gz], by), P);
                    const int ez = f2i_checked(mc * fmaf(c[8], dz[gz], bz), P);
                    if constexpr (inside) {
                        const int idx = base + gx * TL::PX + gy * TL::PY + gz;
                        int_add_checked(&tile[3][idx], em, P); int_add_checked(&tile[0][idx], ex, P);
                        int_add_checked(&tile[1][idx], ey, P); int_add_checked(&tile[2][idx], ez, P);
                    } else {
                        int* cc = grid + 4 * cell_index(P, cx + gx - 1, cy + gy - 1, cz + gz - 1);
                        int_add_checked(cc + 3, em, P); int_add_checked(cc + 0, ex, P); int_add_checked(cc + 1, ey, P); int_add_checked(cc + 2, ez, P);
                    }
                }
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
        const int ax = tile[0][idx], ay = tile[1][idx], az = tile[2][idx], m = tile[3][idx];
        if (!ok || (ax | ay | az | m) == 0) continue;
        int* cc = grid + 4 * ci;
        if (ax) int_add_checked(cc + 0, ax, P);
        if (ay) int_add_checked(cc + 1, ay, P);
        if (az) int_add_checked(cc + 2, az, P);
        if (m) int_add_checked(cc + 3, m, P);
    }
}

// ---------------------------------------------------------------- P2G_2
__device__ __forceinline__ float fast_eos_pow(float x, const DevParams& P)
{
    if (P.eos_pi > 0) {
        float r = x;
        for (int k = 1; k < P.eos_pi; ++k) r *= x;
        return r;
    }
    return __powf(x, P.eos_p);
}

template <int B>
__global__ void __launch_bounds__(TILED_THREADS) k_p2g2_fast(DevParams P, TileGeom g, ParticleView pv,
                                                             const uint32_t* __restrict__ block_start, int* __restrict__ grid)
{
    using TL = Tile<B>;
    __shared__ int tile[3][TL::WORDS];
    __shared__ float tmass[TL::WORDS];
    const int b = blockIdx.x;
    const uint32_t s0 = block_start[b], s1 = block_start[b + 1];
    if (s0 == s1) return;
    TL tl; tl.init(g, b);
    const float inv_mult = 1.0f / P.fmult;
    for (int k = threadIdx.x; k < 3 * TL::WORDS; k += TILED_THREADS) (&tile[0][0])[k] = 0;
    for (int k = threadIdx.x; k < TL::NODES; k += TILED_THREADS) {
        int idx; int64_t ci;
        const bool ok = tl.node(P, k, idx, ci);
        tmass[idx] = ok ? (float)grid[4 * ci + 3] * inv_mult : 0.0f;
    }
    __syncthreads();
    const float inv_rest = 1.0f / P.rest_density;
    for (uint32_t i = s0 + threadIdx.x; i < s1; i += TILED_THREADS) {
        const float px = pv.at(PX, i), py = pv.at(PY, i), pz = pv.at(PZ, i), m = pv.at(PM, i);
        float c[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) c[k] = pv.at(C0 + k, i);
        float wx[3], wy[3], wz[3], dx[3], dy[3], dz[3];
        const int cx = fast_axis(px, wx, dx), cy = fast_axis(py, wy, dy), cz = fast_axis(pz, wz, dz);
        int base;
        const bool in_block = tl.stencil_base(cx, cy, cz, base);
        float density = 0.0f;
        auto gather = [&](auto in_tile) {
        constexpr bool inside = decltype(in_tile)::value;
#pragma unroll
        for (int gx = 0; gx < 3; ++gx)
#pragma unroll
            for (int gy = 0; gy < 3; ++gy) {
                const float wxy = wx[gx] * wy[gy];
                float row = 0.0f;
#pragma unroll
                for (int gz = 0; gz < 3; ++gz) {
                    float gm;
                    if constexpr (inside) gm = tmass[base + gx * TL::PX + gy * TL::PY + gz];
                    else gm = (float)grid[4 * cell_index(P, cx + gx - 1, cy + gy - 1, cz + gz - 1) + 3] * inv_mult;
                    row = fmaf(gm, wz[gz], row);
                }
                density = fmaf(row, wxy, density);
            }
        };
        if (in_block) gather(std::true_type{});
        else if (stencil_in_grid(P, cx, cy, cz)) gather(std::false_type{});
        else continue;  // (a position outside the grid: counted by P2G_1, the particle stays as it is)
        // eq_16_term_0 = -volume * 4 * stress * dt, symmetric; pre-scaled to fixed-point units
        const float volume = __fdividef(m, density);
        const float pr = P.eos_k * (fast_eos_pow(density * inv_rest, P) - 1.0f);
        const float pressure = fmaxf(-0.1f, pr);
        const float s = -volume * 4.0f * P.dt * P.fmult;
        const float mu = P.visc;
        const float e00 = s * fmaf(2.0f * mu, c[0], -pressure), e11 = s * fmaf(2.0f * mu, c[4], -pressure),
                    e22 = s * fmaf(2.0f * mu, c[8], -pressure);
        const float e01 = s * mu * (c[1] + c[3]), e02 = s * mu * (c[2] + c[6]), e12 = s * mu * (c[5] + c[7]);
        auto scatter = [&](auto in_tile) {
        constexpr bool inside = decltype(in_tile)::value;
#pragma unroll
        for (int gx = 0; gx < 3; ++gx) {
            const float ax = e00 * dx[gx], ay = e01 * dx[gx], az = e02 * dx[gx];
#pragma unroll
            for (int gy = 0; gy < 3; ++gy) {
                const float bx = fmaf(e01, dy[gy], ax), by = fmaf(e11, dy[gy], ay), bz = fmaf(e12, dy[gy], az);
                const float wxy = wx[gx] * wy[gy];
#pragma unroll
                for (int gz = 0; gz < 3; ++gz) {
                    const float w = wxy * wz[gz];
                    const int ex = f2i_checked(w * fmaf(e02, dz[gz], bx), P);
                    const int ey = f2i_checked(w * fmaf(e12, dz[gz], by), P);
                    const int ez = f2i_checked(w * fmaf(e22, dz[gz], bz), P);
                    if constexpr (inside) {
                        const int idx = base + gx * TL::PX + gy * TL::PY + gz;
                        int_add_checked(&tile[0][idx], ex, P); int_add_checked(&tile[1][idx], ey, P); int_add_checked(&tile[2][idx], ez, P);
                    } else {
                        int* cc = grid + 4 * cell_index(P, cx + gx - 1, cy + gy - 1, cz + gz - 1);
                        int_add_checked(cc + 0, ex, P); int_add_checked(cc + 1, ey, P); int_add_checked(cc + 2, ez, P);
                    }
                }
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
        const int ax = tile[0][idx], ay = tile[1][idx], az = tile[2][idx];
        if (!ok || (ax | ay | az) == 0) continue;
        int* cc = grid + 4 * ci;
        if (ax) int_add_checked(cc + 0, ax, P);
        if (ay) int_add_checked(cc + 1, ay, P);
        if (az) int_add_checked(cc + 2, az, P);
    }
}

// ---------------------------------------------------------------- G2P
template <int B>
__global__ void __launch_bounds__(TILED_THREADS) k_g2p_fast(DevParams P, TileGeom g, ParticleView pv,
                                                            const uint32_t* __restrict__ block_start, const int4* __restrict__ grid,
                                                            const uint32_t* __restrict__ orig_id, float4* __restrict__ positions)
{
    using TL = Tile<B>;
    __shared__ float tv[3][TL::WORDS];
    const int b = blockIdx.x;
    const uint32_t s0 = block_start[b], s1 = block_start[b + 1];
    if (s0 == s1) return;
    TL tl; tl.init(g, b);
    const float inv_mult = 1.0f / P.fmult;
    for (int k = threadIdx.x; k < TL::NODES; k += TILED_THREADS) {
        int idx; int64_t ci;
        float vx = 0.0f, vy = 0.0f, vz = 0.0f;
        if (tl.node(P, k, idx, ci)) {
            const int4 c = grid[ci];
            vx = (float)c.x * inv_mult; vy = (float)c.y * inv_mult; vz = (float)c.z * inv_mult;
        }
        tv[0][idx] = vx; tv[1][idx] = vy; tv[2][idx] = vz;
    }
    __syncthreads();
    for (uint32_t i = s0 + threadIdx.x; i < s1; i += TILED_THREADS) {
        const float old[3] = {pv.at(PX, i), pv.at(PY, i), pv.at(PZ, i)};
        float wx[3], wy[3], wz[3], dx[3], dy[3], dz[3];
        const int cx = fast_axis(old[0], wx, dx), cy = fast_axis(old[1], wy, dy), cz = fast_axis(old[2], wz, dz);
        int base;
        const bool in_block = tl.stencil_base(cx, cy, cz, base);
        const float wdz[3] = {wz[0] * dz[0], wz[1] * dz[1], wz[2] * dz[2]};
        float v[3] = {0, 0, 0};
        float Bx[3] = {0, 0, 0}, By[3] = {0, 0, 0}, Bz[3] = {0, 0, 0};  // columns of B = sum w * v (x) d
        auto gather = [&](auto in_tile) {
        constexpr bool inside = decltype(in_tile)::value;
#pragma unroll
        for (int gx = 0; gx < 3; ++gx) {
            float S[3] = {0, 0, 0}, Ty[3] = {0, 0, 0}, Tz[3] = {0, 0, 0};  // sums over (gy, gz) for this gx
#pragma unroll
            for (int gy = 0; gy < 3; ++gy) {
                float s[3] = {0, 0, 0}, t[3] = {0, 0, 0};  // sums over gz
#pragma unroll
                for (int gz = 0; gz < 3; ++gz) {
                    float gv[3];
                    if constexpr (inside) {
                        const int idx = base + gx * TL::PX + gy * TL::PY + gz;
                        gv[0] = tv[0][idx]; gv[1] = tv[1][idx]; gv[2] = tv[2][idx];
                    } else {
                        const int4 c = grid[cell_index(P, cx + gx - 1, cy + gy - 1, cz + gz - 1)];
                        gv[0] = (float)c.x * inv_mult; gv[1] = (float)c.y * inv_mult; gv[2] = (float)c.z * inv_mult;
                    }
#pragma unroll
                    for (int k = 0; k < 3; ++k) { s[k] = fmaf(wz[gz], gv[k], s[k]); t[k] = fmaf(wdz[gz], gv[k], t[k]); }
                }
                const float wyd = wy[gy] * dy[gy];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    S[k] = fmaf(wy[gy], s[k], S[k]);
                    Ty[k] = fmaf(wyd, s[k], Ty[k]);
                    Tz[k] = fmaf(wy[gy], t[k], Tz[k]);
                }
            }
            const float wxd = wx[gx] * dx[gx];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                v[k] = fmaf(wx[gx], S[k], v[k]);
                Bx[k] = fmaf(wxd, S[k], Bx[k]);
                By[k] = fmaf(wx[gx], Ty[k], By[k]);
                Bz[k] = fmaf(wx[gx], Tz[k], Bz[k]);
            }
        }
        };
        if (in_block) gather(std::true_type{});
        else if (stencil_in_grid(P, cx, cy, cz)) gather(std::false_type{});
        else continue;  // (a position outside the grid: counted by P2G_1, the particle stays as it is)
        const float Bm[9] = {Bx[0], Bx[1], Bx[2], By[0], By[1], By[2], Bz[0], Bz[1], Bz[2]};
        float np[3], c[9];
        g2p_finish<3>(P, old, Bm, v, np, c);
        pv.at(PX, i) = np[0]; pv.at(PY, i) = np[1]; pv.at(PZ, i) = np[2];
        pv.at(VX, i) = v[0]; pv.at(VY, i) = v[1]; pv.at(VZ, i) = v[2];
#pragma unroll
        for (int k = 0; k < 9; ++k) pv.at(C0 + k, i) = c[k];
        const float len = sqrtf(fmaf(v[0], v[0], fmaf(v[1], v[1], v[2] * v[2])));
        positions[orig_id[i]] = make_float4(np[0], np[1], np[2], len);
    }
}

// ---------------------------------------------------------------- host side
#define LAUNCH_FAST(KERNEL, ...)                                                                          \
    do {                                                                                                  \
        int B, nbx, nby, nbz; int64_t nblocks;                                                            \
        sort_geometry(s, B, nbx, nby, nbz, nblocks);                                                      \
        TileGeom g{nby, nbz, s->dp.gx0 + (s->comm ? 1 : 0)};                                              \
        if (B == 8) KERNEL<8><<<(unsigned)nblocks, TILED_THREADS, 0, s->stream>>>(s->dp, g, __VA_ARGS__); \
        else KERNEL<4><<<(unsigned)nblocks, TILED_THREADS, 0, s->stream>>>(s->dp, g, __VA_ARGS__);        \
        s->launches += 1;                                                                                 \
    } while (0)

void fast_p2g1(MpmSolver* s) { LAUNCH_FAST(k_p2g1_fast, s->view(), sort_block_start(s), reinterpret_cast<int*>(s->grid)); }
void fast_p2g2(MpmSolver* s) { LAUNCH_FAST(k_p2g2_fast, s->view(), sort_block_start(s), reinterpret_cast<int*>(s->grid)); }
void fast_g2p(MpmSolver* s)
{
    LAUNCH_FAST(k_g2p_fast, s->view(), sort_block_start(s), reinterpret_cast<const int4*>(s->grid), s->orig_id, s->positions);
}

}  // namespace mpm
