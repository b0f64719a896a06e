// mpm_kernels.h -- host-callable launchers of every kernel in libmpm_b200.so.
#pragma once
#include "mpm_common.cuh"

namespace mpm {

// ---- reference-shaped path (mpm_kernels_ref.cu)
void launch_p2g1_ref(const DevParams& P, ParticleView pv, int64_t n, void* grid, cudaStream_t st);
void launch_p2g2_ref(const DevParams& P, ParticleView pv, int64_t n, void* grid, cudaStream_t st);
void launch_update_grid(const DevParams& P, void* grid, int64_t ncells, cudaStream_t st);
// clear / update restricted to a device-resident box of cells {x0, x1, y0, y1, z0, z1} (3D fixed-point grid)
// (halo_lo / halo_hi: stored planes at the low / high end of a multi-GPU slab that are swept whole, see k_clear_box)
void launch_clear_box(const DevParams& P, void* grid, const int* box, int halo_lo, int halo_hi, cudaStream_t st);
void launch_update_box(const DevParams& P, void* grid, const int* box, int halo_lo, int halo_hi, cudaStream_t st);
void launch_g2p_ref(const DevParams& P, ParticleView pv, int64_t n, const void* grid, const uint32_t* orig_id,
                    float4* positions, cudaStream_t st);

// ---- layout conversion / scene generation (mpm_io.cu)
// 80-byte AoS records (reference GPU buffer layout) <-> SoA planes
void launch_aos80_to_soa(const float* aos, ParticleView pv, int64_t dst_off, int64_t n, cudaStream_t st);
// writes record orig_id[i] of `aos` from slot i (un-permutes to original index order)
void launch_soa_to_aos80(ParticleView pv, const uint32_t* orig_id, float* aos, int64_t n, cudaStream_t st);
// packed SoA host layout (pos[3n], vel[3n], C[9n], mass[n]) staged on the device <-> planes
void launch_packed_to_soa(const float* pos, const float* vel, const float* C, const float* mass, ParticleView pv,
                          int64_t dst_off, int64_t n, cudaStream_t st);
void launch_soa_to_packed(ParticleView pv, const uint32_t* orig_id, float* pos, float* vel, float* C, float* mass,
                          int64_t n, cudaStream_t st);
void launch_iota(uint32_t* p, uint32_t start, int64_t n, cudaStream_t st);
void launch_positions(ParticleView pv, const uint32_t* orig_id, float4* positions, int64_t n, cudaStream_t st);
void launch_positions_rec(RecView rv, const uint32_t* orig_id, float4* positions, int64_t n, cudaStream_t st);
void launch_positions_q16(ParticleView pv, const uint32_t* orig_id, void* out, int64_t n, const float scale[3], cudaStream_t st);
void launch_positions_q16_rec(RecView rv, const uint32_t* orig_id, void* out, int64_t n, const float scale[3], cudaStream_t st);
void launch_rec_to_planes(RecView rv, ParticleView pv, int64_t n, cudaStream_t st);
// lattice block with per-axis coordinate tables (the fp32 accumulating loops run on the host: they are
// O(R) work), vel = 0, C = 0, mass = 1
void launch_lattice(const float* xs, int nx, const float* ys, int ny, const float* zs, int nz, ParticleView pv,
                    int64_t dst_off, cudaStream_t st);

}  // namespace mpm
