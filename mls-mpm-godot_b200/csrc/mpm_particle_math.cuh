// mpm_particle_math.cuh -- per-particle arithmetic of the MLS-MPM fluid step in STRICT mode: one IEEE
// binary32 operation per operator of the C# statement it implements, in that statement's order, so that
// with the int32 fixed-point grid every grid word and particle float is reproducible bit for bit.
// Citations: F = mls-mpm/3d/fluid_multithread/MLSMPM3DFluidMultithread.cs,
// X = .../fluid_multithread_fixed_point/MLSMPM3DFluidMultithreadNew.cs, D = mls-mpm/2d/fluid/MLSMPM2DFluid.cs,
// M = mls-mpm/2d/fluid_multithread/MLSMPM2DFluidMultithread.cs, g2p.glsl = the GPU variant's shader.
#pragma once
#include "mpm_common.cuh"

namespace mpm {

struct ParticleIn {
    float px, py, pz, vx, vy, vz, m;
    float c[9];  // column-major: c[3*col+row]
};

// Distance of stencil node n to the particle along one axis: (cell_x - p.pos) + 0.5   (F:276)
__device__ __forceinline__ float node_dist(int n, float p) { return sadd(ssub(__int2float_rn(n), p), 0.5f); }

// P2G_1 contribution of one particle to one node (F:273-287, D:216-229).
// out = (mass_contrib * (v + C*dist)).xyz, mass_contrib
template <int DIM>
__device__ __forceinline__ void p2g1_node(const ParticleIn& p, float weight, float dx, float dy, float dz,
                                          float& mc, float& ox, float& oy, float& oz)
{
    float qx, qy, qz;
    if (DIM == 3) {  // Basis * Vector3 = row dots, ((a+b)+c)
        qx = sadd(sadd(smul(p.c[0], dx), smul(p.c[3], dy)), smul(p.c[6], dz));
        qy = sadd(sadd(smul(p.c[1], dx), smul(p.c[4], dy)), smul(p.c[7], dz));
        qz = sadd(sadd(smul(p.c[2], dx), smul(p.c[5], dy)), smul(p.c[8], dz));
    } else {         // Transform2D * Vector2 + zero origin
        qx = sadd(sadd(smul(p.c[0], dx), smul(p.c[3], dy)), 0.0f);
        qy = sadd(sadd(smul(p.c[1], dx), smul(p.c[4], dy)), 0.0f);
        qz = 0.0f;
    }
    mc = smul(weight, p.m);
    ox = smul(mc, sadd(p.vx, qx));
    oy = smul(mc, sadd(p.vy, qy));
    oz = (DIM == 3) ? smul(mc, sadd(p.vz, qz)) : 0.0f;
}

// Columns of eq_16_term_0 (F:326-347, D:263-285): e[3*col+row].
template <int DIM>
__device__ __forceinline__ void p2g2_stress(const DevParams& P, const float c[9], float m, float density,
                                            float e[9])
{
    const float volume = sdiv(m, density);
    const float pw = eos_pow(sdiv(density, P.rest_density), P);
    const float pr = smul(P.eos_k, ssub(pw, 1.0f));
    const float pressure = (-0.1f > pr) ? -0.1f : pr;
    const float np = -pressure;
    const float visc = P.visc, dt = P.dt;
    float t[9];
    if (DIM == 2) {
        const float trace = sadd(c[3], c[1]);
        t[0] = sadd(np, smul(visc, c[0]));
        t[1] = sadd(0.0f, smul(visc, trace));
        t[3] = sadd(0.0f, smul(visc, trace));
        t[4] = sadd(np, smul(visc, c[4]));
        t[2] = t[5] = t[6] = t[7] = t[8] = 0.0f;
    } else {
        t[0] = sadd(np, smul(sadd(c[0], c[0]), visc));
        t[1] = sadd(0.0f, smul(sadd(c[1], c[3]), visc));
        t[2] = sadd(0.0f, smul(sadd(c[2], c[6]), visc));
        t[3] = sadd(0.0f, smul(sadd(c[3], c[1]), visc));
        t[4] = sadd(np, smul(sadd(c[4], c[4]), visc));
        t[5] = sadd(0.0f, smul(sadd(c[5], c[7]), visc));
        t[6] = sadd(0.0f, smul(sadd(c[6], c[2]), visc));
        t[7] = sadd(0.0f, smul(sadd(c[7], c[5]), visc));
        t[8] = sadd(np, smul(sadd(c[8], c[8]), visc));
    }
    if (P.eq16_order == 1) {  // D:285  ((-dt*volume)*stress)*4
        const float s = smul(-dt, volume);
#pragma unroll
        for (int k = 0; k < 9; ++k) e[k] = smul(smul(s, t[k]), 4.0f);
    } else {                  // F:347  ((-volume*4)*stress)*dt
        const float s = smul(-volume, 4.0f);
#pragma unroll
        for (int k = 0; k < 9; ++k) e[k] = smul(smul(s, t[k]), dt);
    }
}

// P2G_2 momentum of one node (F:364, D:300): Basis(e.X*w, e.Y*w, e.Z*w) * dist
template <int DIM>
__device__ __forceinline__ void p2g2_node(const float e[9], float weight, float dx, float dy, float dz,
                                          float& ox, float& oy, float& oz)
{
    if (DIM == 3) {
        ox = sadd(sadd(smul(smul(e[0], weight), dx), smul(smul(e[3], weight), dy)), smul(smul(e[6], weight), dz));
        oy = sadd(sadd(smul(smul(e[1], weight), dx), smul(smul(e[4], weight), dy)), smul(smul(e[7], weight), dz));
        oz = sadd(sadd(smul(smul(e[2], weight), dx), smul(smul(e[5], weight), dy)), smul(smul(e[8], weight), dz));
    } else {
        ox = sadd(sadd(smul(smul(e[0], weight), dx), smul(smul(e[3], weight), dy)), 0.0f);
        oy = sadd(sadd(smul(smul(e[1], weight), dx), smul(smul(e[4], weight), dy)), 0.0f);
        oz = 0.0f;
    }
}

// G2P accumulation of one node (F:454-464): B += (w*v_i) (x) dist ; vel += w*v_i
template <int DIM>
__device__ __forceinline__ void g2p_node(float gvx, float gvy, float gvz, float weight, float dx, float dy,
                                         float dz, float B[9], float v[3])
{
    const float wx = smul(gvx, weight), wy = smul(gvy, weight);
    B[0] = sadd(B[0], smul(wx, dx));
    B[1] = sadd(B[1], smul(wy, dx));
    B[3] = sadd(B[3], smul(wx, dy));
    B[4] = sadd(B[4], smul(wy, dy));
    v[0] = sadd(v[0], wx);
    v[1] = sadd(v[1], wy);
    if (DIM == 3) {
        const float wz = smul(gvz, weight);
        B[2] = sadd(B[2], smul(wz, dx));
        B[5] = sadd(B[5], smul(wz, dy));
        B[6] = sadd(B[6], smul(wx, dz));
        B[7] = sadd(B[7], smul(wy, dz));
        B[8] = sadd(B[8], smul(wz, dz));
        v[2] = sadd(v[2], wz);
    }
}

// Tail of G2P (F:468-514, X:548-587, D:372-416, g2p.glsl:108-150): C = 4B, advect, clamp, interaction,
// predictive wall.  old = pre-advection position; writes new position/velocity/C into np/v/c.
// EXTRA = false compiles the sphere list out (the cell kernels instantiate both: at 128 registers the loop alone cost
// the list-less case 2 %).
template <int DIM, bool EXTRA = true>
__device__ __forceinline__ void g2p_finish(const DevParams& P, const float old[3], const float B[9], float v[3],
                                           float np[3], float c[9])
{
#pragma unroll
    for (int k = 0; k < 9; ++k) c[k] = smul(B[k], 4.0f);
    const float R[3] = {(float)P.Rx, (float)P.Ry, (float)P.Rz};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        if (a < DIM) np[a] = clampf(sadd(old[a], smul(v[a], P.dt)), P.clamp_min, ssub(R[a], P.clamp_max_off));
        else np[a] = old[a];
    }
    if (P.interaction == 1 || P.interaction == 2) {
        const bool post = P.interaction == 1;  // (scalars, not a pointer into np / old: those arrays must stay in registers)
        const float qx = post ? np[0] : old[0], qy = post ? np[1] : old[1], qz = post ? np[2] : old[2];
        const float dx = ssub(qx, P.sphere[0]), dy = ssub(qy, P.sphere[1]), dz = ssub(qz, P.sphere[2]);
        const float d2 = sadd(sadd(smul(dx, dx), smul(dy, dy)), smul(dz, dz));
        if (d2 < smul(P.sphere_r, P.sphere_r)) {
            float fx = 0.0f, fy = 0.0f, fz = 0.0f;
            if (d2 != 0.0f) {
                const float len = __fsqrt_rn(d2);
                fx = sdiv(dx, len); fy = sdiv(dy, len); fz = sdiv(dz, len);
            }
            v[0] = sadd(v[0], smul(fx, 1.0f));
            v[1] = sadd(v[1], smul(fy, 1.0f));
            v[2] = sadd(v[2], smul(fz, 1.0f));
        }
        if constexpr (EXTRA)
        for (int k = 0; k < P.n_extra; ++k) {  // sphere list (mpm_set_colliders): the same rule for every further sphere
            const float ex = ssub(qx, P.extra[k][0]), ey = ssub(qy, P.extra[k][1]), ez = ssub(qz, P.extra[k][2]);
            const float e2 = sadd(sadd(smul(ex, ex), smul(ey, ey)), smul(ez, ez));
            if (e2 < smul(P.extra[k][3], P.extra[k][3])) {
                float gx = 0.0f, gy = 0.0f, gz = 0.0f;
                if (e2 != 0.0f) {
                    const float len = __fsqrt_rn(e2);
                    gx = sdiv(ex, len); gy = sdiv(ey, len); gz = sdiv(ez, len);
                }
                v[0] = sadd(v[0], smul(gx, 1.0f));
                v[1] = sadd(v[1], smul(gy, 1.0f));
                v[2] = sadd(v[2], smul(gz, 1.0f));
            }
        }
    } else if (P.interaction == 3) {
        const float dx = ssub(np[0], P.mouse[0]), dy = ssub(np[1], P.mouse[1]);
        const float d2 = sadd(smul(dx, dx), smul(dy, dy));
        if (d2 < smul(P.mouse_r, P.mouse_r)) {
            const float len = __fsqrt_rn(d2);
            const float nf = sdiv(1.0f, sdiv(len, P.mouse_r));
            float nx = 0.0f, ny = 0.0f;
            if (d2 != 0.0f) { nx = sdiv(dx, len); ny = sdiv(dy, len); }
            const float fx = smul(smul(nx, nf), 0.1f), fy = smul(smul(ny, nf), 0.1f);
            if (!(isnan(fx) || isnan(fy))) { v[0] = sadd(v[0], fx); v[1] = sadd(v[1], fy); }
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        if (a < DIM) {
            const float xn = sadd(np[a], v[a]);
            const float wmin = P.wall_min, wmax = ssub(R[a], P.wall_max_off);
            float va = v[a];
            if (xn < wmin) va = sadd(va, smul(P.wall_gain, ssub(wmin, xn)));
            if (xn > wmax) va = sadd(va, smul(P.wall_gain, ssub(wmax, xn)));
            v[a] = va;
        }
    }
}

}  // namespace mpm
