// mpm_io.cu -- layout conversion between the reference's buffers and the solver's SoA planes, the
// position hand-off array, and the lattice scene generator.  None of this is on the per-step hot path.
#include <cuda_fp16.h>
#include "mpm_kernels.h"

namespace mpm {

// Reference particle record (MLSMPM3DFluidMultithreadGPU.cs:8-22): 20 floats =
// pos.xyz pad | vel.xyz mass | C_x.xyz pad | C_y.xyz pad | C_z.xyz pad
__global__ void __launch_bounds__(256) k_aos80_to_soa(const float4* __restrict__ aos, ParticleView pv,
                                                      int64_t dst_off, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 a = aos[5 * i], b = aos[5 * i + 1], c0 = aos[5 * i + 2], c1 = aos[5 * i + 3], c2 = aos[5 * i + 4];
    const int64_t d = dst_off + i;
    pv.at(PX, d) = a.x; pv.at(PY, d) = a.y; pv.at(PZ, d) = a.z;
    pv.at(VX, d) = b.x; pv.at(VY, d) = b.y; pv.at(VZ, d) = b.z; pv.at(PM, d) = b.w;
    pv.at(C0, d) = c0.x; pv.at(C1, d) = c0.y; pv.at(C2, d) = c0.z;
    pv.at(C3, d) = c1.x; pv.at(C4, d) = c1.y; pv.at(C5, d) = c1.z;
    pv.at(C6, d) = c2.x; pv.at(C7, d) = c2.y; pv.at(C8, d) = c2.z;
}

__global__ void __launch_bounds__(256) k_soa_to_aos80(ParticleView pv, const uint32_t* __restrict__ orig_id,
                                                      float4* __restrict__ aos, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t o = orig_id ? (int64_t)orig_id[i] : i;
    aos[5 * o] = make_float4(pv.at(PX, i), pv.at(PY, i), pv.at(PZ, i), 0.0f);
    aos[5 * o + 1] = make_float4(pv.at(VX, i), pv.at(VY, i), pv.at(VZ, i), pv.at(PM, i));
    aos[5 * o + 2] = make_float4(pv.at(C0, i), pv.at(C1, i), pv.at(C2, i), 0.0f);
    aos[5 * o + 3] = make_float4(pv.at(C3, i), pv.at(C4, i), pv.at(C5, i), 0.0f);
    aos[5 * o + 4] = make_float4(pv.at(C6, i), pv.at(C7, i), pv.at(C8, i), 0.0f);
}

__global__ void __launch_bounds__(256) k_packed_to_soa(const float* __restrict__ pos, const float* __restrict__ vel,
                                                       const float* __restrict__ C, const float* __restrict__ mass,
                                                       ParticleView pv, int64_t dst_off, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t d = dst_off + i;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        pv.at(PX + a, d) = pos[3 * i + a];
        pv.at(VX + a, d) = vel ? vel[3 * i + a] : 0.0f;
    }
    pv.at(PM, d) = mass ? mass[i] : 1.0f;
#pragma unroll
    for (int k = 0; k < 9; ++k) pv.at(C0 + k, d) = C ? C[9 * i + k] : 0.0f;
}

__global__ void __launch_bounds__(256) k_soa_to_packed(ParticleView pv, const uint32_t* __restrict__ orig_id,
                                                       float* pos, float* vel, float* C, float* mass, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t o = orig_id ? (int64_t)orig_id[i] : i;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        if (pos) pos[3 * o + a] = pv.at(PX + a, i);
        if (vel) vel[3 * o + a] = pv.at(VX + a, i);
    }
    if (mass) mass[o] = pv.at(PM, i);
    if (C) {
#pragma unroll
        for (int k = 0; k < 9; ++k) C[9 * o + k] = pv.at(C0 + k, i);
    }
}

__global__ void __launch_bounds__(256) k_iota(uint32_t* p, uint32_t start, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = start + (uint32_t)i;
}

// (x, y, z, |v|) of g2p.glsl:149-150, original index order
template <class View>
__global__ void __launch_bounds__(256) k_positions(View pv, const uint32_t* __restrict__ orig_id,
                                                   float4* __restrict__ positions, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float vx = pv.at(VX, i), vy = pv.at(VY, i), vz = pv.at(VZ, i);
    const float len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)), __fmul_rn(vz, vz)));
    positions[orig_id ? orig_id[i] : (uint32_t)i] = make_float4(pv.at(PX, i), pv.at(PY, i), pv.at(PZ, i), len);
}

// the same hand-off at half the bytes, for hosts that must copy it over PCIe: x, y, z as unsigned 16-bit fractions of the
// domain (code q = rint(p / R * 65535): a step of R / 65535 cells, 0.004 cell at R = 256), |v| as an IEEE half
template <class View>
__global__ void __launch_bounds__(256) k_positions_q16(View pv, const uint32_t* __restrict__ orig_id, ushort4* __restrict__ out,
                                                       int64_t n, float sx, float sy, float sz)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float vx = pv.at(VX, i), vy = pv.at(VY, i), vz = pv.at(VZ, i);
    const float len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)), __fmul_rn(vz, vz)));
    auto q = [](float p, float s) { return (unsigned short)__float2uint_rn(fminf(fmaxf(p * s, 0.0f), 65535.0f)); };
    out[orig_id ? orig_id[i] : (uint32_t)i] = make_ushort4(q(pv.at(PX, i), sx), q(pv.at(PY, i), sy), q(pv.at(PZ, i), sz),
                                                           __half_as_ushort(__float2half_rn(len)));
}

__global__ void __launch_bounds__(256) k_lattice(const float* __restrict__ xs, int nx, const float* __restrict__ ys,
                                                 int ny, const float* __restrict__ zs, int nz, ParticleView pv,
                                                 int64_t dst_off)
{
    const int64_t n = (int64_t)nx * ny * nz;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int iz = (int)(i % nz), iy = (int)(i / nz % ny), ix = (int)(i / nz / ny);
    const int64_t d = dst_off + i;
    pv.at(PX, d) = xs[ix]; pv.at(PY, d) = ys[iy]; pv.at(PZ, d) = zs[iz];
    pv.at(VX, d) = 0.0f; pv.at(VY, d) = 0.0f; pv.at(VZ, d) = 0.0f; pv.at(PM, d) = 1.0f;
#pragma unroll
    for (int k = 0; k < 9; ++k) pv.at(C0 + k, d) = 0.0f;
}

static inline unsigned nb(int64_t n) { return (unsigned)((n + 255) / 256); }

void launch_aos80_to_soa(const float* aos, ParticleView pv, int64_t dst_off, int64_t n, cudaStream_t st)
{
    if (n > 0) k_aos80_to_soa<<<nb(n), 256, 0, st>>>(reinterpret_cast<const float4*>(aos), pv, dst_off, n);
}
void launch_soa_to_aos80(ParticleView pv, const uint32_t* orig_id, float* aos, int64_t n, cudaStream_t st)
{
    if (n > 0) k_soa_to_aos80<<<nb(n), 256, 0, st>>>(pv, orig_id, reinterpret_cast<float4*>(aos), n);
}
void launch_packed_to_soa(const float* pos, const float* vel, const float* C, const float* mass, ParticleView pv,
                          int64_t dst_off, int64_t n, cudaStream_t st)
{
    if (n > 0) k_packed_to_soa<<<nb(n), 256, 0, st>>>(pos, vel, C, mass, pv, dst_off, n);
}
void launch_soa_to_packed(ParticleView pv, const uint32_t* orig_id, float* pos, float* vel, float* C, float* mass,
                          int64_t n, cudaStream_t st)
{
    if (n > 0) k_soa_to_packed<<<nb(n), 256, 0, st>>>(pv, orig_id, pos, vel, C, mass, n);
}
void launch_iota(uint32_t* p, uint32_t start, int64_t n, cudaStream_t st)
{
    if (n > 0) k_iota<<<nb(n), 256, 0, st>>>(p, start, n);
}
void launch_positions(ParticleView pv, const uint32_t* orig_id, float4* positions, int64_t n, cudaStream_t st)
{
    if (n > 0) k_positions<ParticleView><<<nb(n), 256, 0, st>>>(pv, orig_id, positions, n);
}
void launch_positions_rec(RecView rv, const uint32_t* orig_id, float4* positions, int64_t n, cudaStream_t st)
{
    if (n > 0) k_positions<RecView><<<nb(n), 256, 0, st>>>(rv, orig_id, positions, n);
}

void launch_positions_q16(ParticleView pv, const uint32_t* orig_id, void* out, int64_t n, const float scale[3], cudaStream_t st)
{
    if (n > 0) k_positions_q16<ParticleView><<<nb(n), 256, 0, st>>>(pv, orig_id, static_cast<ushort4*>(out), n, scale[0], scale[1], scale[2]);
}
void launch_positions_q16_rec(RecView rv, const uint32_t* orig_id, void* out, int64_t n, const float scale[3], cudaStream_t st)
{
    if (n > 0) k_positions_q16<RecView><<<nb(n), 256, 0, st>>>(rv, orig_id, static_cast<ushort4*>(out), n, scale[0], scale[1], scale[2]);
}

// 64-byte records (slot order) back into the grouped planes
__global__ void __launch_bounds__(256) k_rec_to_planes(const float4* __restrict__ rec, ParticleView pv, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float* q = pv.rec(i);
    const float4 a = rec[4 * i], b = rec[4 * i + 1], c = rec[4 * i + 2], d = rec[4 * i + 3];  // (record layout: mpm_common.cuh)
    q[PX * GROUP] = a.x; q[PY * GROUP] = a.y; q[PZ * GROUP] = a.z; q[PM * GROUP] = a.w;
    q[VX * GROUP] = b.x; q[VY * GROUP] = b.y; q[VZ * GROUP] = b.z; q[C2 * GROUP] = b.w;
    q[C0 * GROUP] = c.x; q[C1 * GROUP] = c.y; q[C3 * GROUP] = c.z; q[C4 * GROUP] = c.w;
    q[C6 * GROUP] = d.x; q[C7 * GROUP] = d.y; q[C5 * GROUP] = d.z; q[C8 * GROUP] = d.w;
}
void launch_rec_to_planes(RecView rv, ParticleView pv, int64_t n, cudaStream_t st)
{
    if (n > 0) k_rec_to_planes<<<nb(n), 256, 0, st>>>(reinterpret_cast<const float4*>(rv.base), pv, n);
}
void launch_lattice(const float* xs, int nx, const float* ys, int ny, const float* zs, int nz, ParticleView pv,
                    int64_t dst_off, cudaStream_t st)
{
    const int64_t n = (int64_t)nx * ny * nz;
    if (n > 0) k_lattice<<<nb(n), 256, 0, st>>>(xs, nx, ys, ny, zs, nz, pv, dst_off);
}

}  // namespace mpm
