// mpm_bin.cu -- particle binning for the cell kernels (MPM_PATH_CELL): a one-pass counting sort by cell key.
//
// The reference never reorders particles (particle i keeps index i); binning is new.  The cell kernels
// (mpm_kernels_cell.cu) give every grid cell to one thread, which keeps the cell's 27-node stencil in registers
// across all particles of the cell.  That needs the particles of a cell to be found without searching, and the
// loads of a warp (32 consecutive cells = one "chunk") to be coalesced.  Layout of the particle planes:
//
//     sorted by (block, chunk, rank r inside the cell, cell inside the chunk)
//
// i.e. inside a chunk, first the rank-0 particle of every non-empty cell (in cell order), then the rank-1
// particles, ...  A warp at rank r reads slots chunk_start + S(r) + (number of lower lanes that still have a
// particle at rank r): consecutive addresses.  S(r) = sum over the chunk's cells of min(count, r) is carried
// as a running sum of ballots, so the only metadata are the per-cell counts and their exclusive scan.
//
// Per step:  counts of the NEW cells are accumulated by G2P itself (fire-and-forget RED on cnt[next], key stored
// per particle), so binning = clear cursor -> 3-kernel exclusive scan of the counts (also lists the non-empty
// blocks) -> k_place (rank by atomic cursor, destination slot from the chunk's counts) -> k_gather (16 planes
// + id, coalesced writes).  The rank comes from an atomic, so the order of the particles INSIDE a cell is not
// reproducible run to run; the fixed-point grid sums do not depend on it (int adds commute) and the
// MPM_MATH_FAST float accumulation is covered by its stated tolerance.
#include "mpm_bin.h"

#include <algorithm>

#include "mpm_kernels.h"
#include "mpm_tile.cuh"

namespace mpm {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // 4096 count entries per CTA

#define CKB(call)                                                          \
    do {                                                                   \
        cudaError_t e_ = (call);                                           \
        if (e_ != cudaSuccess) {                                           \
            s->err = std::string(#call) + ": " + cudaGetErrorString(e_);   \
            return MPM_ERR_CUDA;                                           \
        }                                                                  \
    } while (0)

// ---- keys + counts from positions (first step, after uploads, and every step in multi-GPU mode)
__global__ void __launch_bounds__(256) k_bin_keys(KeyGeom g, ParticleView pv, int64_t n, uint32_t nslots, uint32_t* __restrict__ keys,
                                                  uint32_t* __restrict__ cnt)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int cx = __float2int_rz(pv.at(PX, i)), cy = __float2int_rz(pv.at(PY, i)), cz = __float2int_rz(pv.at(PZ, i));
    uint32_t k = cell_key(g, cx, cy, cz);
    k = k < nslots ? k : nslots - 1;  // a NaN / out-of-slab position must not index outside the count array
    keys[i] = k;
    atomicAdd(&cnt[k], 1u);
}

// ---- exclusive scan of the counts, 3 kernels
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const uint32_t* __restrict__ cnt, int64_t nslots, uint32_t* __restrict__ tile_sums)
{
    __shared__ uint32_t wsum[SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t sum = 0;
    if (base + SCAN_ITEMS <= nslots) {
        const uint4* p = reinterpret_cast<const uint4*>(cnt + base);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS / 4; ++k) { const uint4 v = p[k]; sum += v.x + v.y + v.z + v.w; }
    } else {
        for (int k = 0; k < SCAN_ITEMS; ++k) if (base + k < nslots) sum += cnt[base + k];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int k = 0; k < SCAN_THREADS / 32; ++k) t += wsum[k];
        tile_sums[blockIdx.x] = t;
    }
}

// one CTA: exclusive scan of the tile sums in place; also resets the per-step counters
__global__ void __launch_bounds__(1024) k_scan_tiles(uint32_t* __restrict__ tile_sums, int ntiles, uint32_t* __restrict__ misc)
{
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    if (threadIdx.x < BIN_MISC_WORDS) misc[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int base = 0; base < ntiles; base += 1024) {
        const int i = base + threadIdx.x;
        const uint32_t v = (i < ntiles) ? tile_sums[i] : 0;
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) wsum[w] = x;
        __syncthreads();
        if (w == 0) {
            uint32_t s = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
            wsum[lane] = s;  // inclusive over warps
        }
        __syncthreads();
        const uint32_t carry = carry_s;
        const uint32_t woff = w ? wsum[w - 1] : 0;
        if (i < ntiles) tile_sums[i] = carry + woff + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + woff + x;
        __syncthreads();
    }
}

// per tile: exclusive scan with the tile's offset -> cell_start; append the non-empty blocks to the active list
template <int CELL_BITS>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const uint32_t* __restrict__ cnt, int64_t nslots, const uint32_t* __restrict__ tile_sums,
                                                             uint32_t* __restrict__ cell_start, uint32_t* __restrict__ active,
                                                             uint32_t* __restrict__ misc)
{
    __shared__ uint32_t wsum[SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    if (base + SCAN_ITEMS <= nslots) {
        const uint4* p = reinterpret_cast<const uint4*>(cnt + base);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS / 4; ++k) { const uint4 q = p[k]; v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w; }
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) v[k] = (base + k < nslots) ? cnt[base + k] : 0;
    }
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) sum += v[k];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t x = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) wsum[w] = x;
    __syncthreads();
    uint32_t woff = 0;
    for (int k = 0; k < w; ++k) woff += wsum[k];
    uint32_t run = tile_sums[blockIdx.x] + woff + x - sum;
    // grid blocks: 2^CELL_BITS consecutive entries.  CELL_BITS = 9 -> 32 threads per block; 6 -> 4 threads per block
    constexpr int THREADS_PER_BLOCK = (1 << CELL_BITS) / SCAN_ITEMS;
    static_assert(THREADS_PER_BLOCK >= 1 && THREADS_PER_BLOCK <= 32, "block must span 1..32 threads");
    uint32_t bsum = sum;
#pragma unroll
    for (int o = THREADS_PER_BLOCK / 2; o > 0; o >>= 1) bsum += __shfl_xor_sync(0xffffffffu, bsum, o);
    if ((threadIdx.x % THREADS_PER_BLOCK) == 0 && bsum > 0 && base < nslots)
        active[atomicAdd(&misc[BIN_N_ACTIVE], 1u)] = (uint32_t)(base >> CELL_BITS);
    if (base + SCAN_ITEMS <= nslots) {
        uint4* o4 = reinterpret_cast<uint4*>(cell_start + base);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS / 4; ++k) {
            uint4 q;
            q.x = run; run += v[4 * k]; q.y = run; run += v[4 * k + 1]; q.z = run; run += v[4 * k + 2]; q.w = run; run += v[4 * k + 3];
            o4[k] = q;
        }
    } else {
        for (int k = 0; k < SCAN_ITEMS; ++k) if (base + k < nslots) { cell_start[base + k] = run; run += v[k]; }
    }
    if (base <= nslots - 1 && nslots - 1 < base + SCAN_ITEMS) cell_start[nslots] = run;  // the thread holding the last entry: grand total
}

// rank by atomic cursor; destination slot inside the chunk from the chunk's 32 counts
__global__ void __launch_bounds__(256) k_place(const uint32_t* __restrict__ keys, int64_t n, const uint32_t* __restrict__ cnt,
                                               const uint32_t* __restrict__ cell_start, uint32_t* __restrict__ fill, uint32_t* __restrict__ src_of)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t key = keys[i];
    const uint32_t r = atomicAdd(&fill[key], 1u);
    const uint32_t chunk = key & ~31u, lane = key & 31u;
    const uint4* c4 = reinterpret_cast<const uint4*>(cnt + chunk);
    uint32_t below = 0;   // sum over the chunk's cells of min(count, r)
    uint32_t before = 0;  // lower lanes that still have a particle at rank r
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint4 q = c4[k];
        const uint32_t c[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            below += min(c[j], r);
            before += ((uint32_t)(4 * k + j) < lane && c[j] > r) ? 1u : 0u;
        }
    }
    src_of[cell_start[chunk] + below + before] = (uint32_t)i;
}

__global__ void __launch_bounds__(256) k_gather(ParticleView src, ParticleView dst, const uint32_t* __restrict__ src_of,
                                                const uint32_t* __restrict__ id_src, uint32_t* __restrict__ id_dst, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t j = src_of[i];
    float v[NPLANES];
#pragma unroll
    for (int k = 0; k < NPLANES; ++k) v[k] = src.at(k, j);
    const uint32_t id = id_src[j];
#pragma unroll
    for (int k = 0; k < NPLANES; ++k) dst.at(k, i) = v[k];
    id_dst[i] = id;
}

// block_start for the strict tiled kernels (they only need each block's contiguous particle range)
__global__ void __launch_bounds__(256) k_block_start(const uint32_t* __restrict__ cell_start, int cell_bits, int64_t nblocks, uint32_t* __restrict__ block_start)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b <= nblocks) block_start[b] = cell_start[b << cell_bits];
}

// ---------------------------------------------------------------- host side
static int ilog2_ceil64(int64_t v)
{
    int b = 0;
    while (((int64_t)1 << b) < v) ++b;
    return b;
}

int bin_create(MpmSolver* s)
{
    BinState* st = new BinState();
    s->bin = st;
    const int64_t cells = (int64_t)s->dp.nxl * s->dp.Ry * s->dp.Rz;
    st->B = (cells >= (int64_t)96 * 96 * 96) ? 8 : 4;
    st->logB = (st->B == 8) ? 3 : 2;
    st->cell_bits = 3 * st->logB;
    st->nbx = (s->dp.nxl + st->B - 1) / st->B;
    st->nby = (s->dp.Ry + st->B - 1) / st->B;
    st->nbz = (s->dp.Rz + st->B - 1) / st->B;
    st->nblocks = (int64_t)st->nbx * st->nby * st->nbz;
    if (ilog2_ceil64(st->nblocks) + st->cell_bits > 31) { s->err = "grid too large for 32-bit cell keys"; return MPM_ERR_INVALID; }
    st->nslots = st->nblocks << st->cell_bits;
    st->ntiles = (st->nslots + SCAN_TILE - 1) / SCAN_TILE;
    for (int k = 0; k < 2; ++k) {
        CKB(cudaMalloc(&st->cnt[k], sizeof(uint32_t) * (st->nslots + 32)));
        CKB(cudaMemsetAsync(st->cnt[k], 0, sizeof(uint32_t) * (st->nslots + 32), s->stream));
    }
    CKB(cudaMalloc(&st->cell_start, sizeof(uint32_t) * (st->nslots + 32)));
    CKB(cudaMemsetAsync(st->cell_start, 0, sizeof(uint32_t) * (st->nslots + 32), s->stream));
    CKB(cudaMalloc(&st->fill, sizeof(uint32_t) * st->nslots));
    CKB(cudaMalloc(&st->keys, sizeof(uint32_t) * s->pitch));
    CKB(cudaMalloc(&st->src_of, sizeof(uint32_t) * s->pitch));
    CKB(cudaMalloc(&st->tile_sums, sizeof(uint32_t) * (st->ntiles + 1)));
    CKB(cudaMalloc(&st->active, sizeof(uint32_t) * st->nblocks));
    CKB(cudaMalloc(&st->misc, sizeof(uint32_t) * BIN_MISC_WORDS));
    CKB(cudaMemsetAsync(st->misc, 0, sizeof(uint32_t) * BIN_MISC_WORDS, s->stream));
    CKB(cudaMalloc(&st->block_start, sizeof(uint32_t) * (st->nblocks + 1)));
    st->cur = 0;
    st->next_valid = false;
    return MPM_OK;
}

void bin_destroy(MpmSolver* s)
{
    BinState* st = s->bin;
    if (!st) return;
    cudaFree(st->cnt[0]); cudaFree(st->cnt[1]); cudaFree(st->cell_start); cudaFree(st->fill); cudaFree(st->keys);
    cudaFree(st->src_of); cudaFree(st->tile_sums); cudaFree(st->active); cudaFree(st->misc); cudaFree(st->block_start);
    delete st;
    s->bin = nullptr;
}

KeyGeom bin_key_geom(const MpmSolver* s)
{
    const BinState* st = s->bin;
    return KeyGeom{s->dp.dim, st->logB, st->nby, st->nbz, s->dp.gx0 + (s->comm ? 1 : 0)};
}

int bin_particles(MpmSolver* s)
{
    BinState* st = s->bin;
    const int64_t n = s->n;
    const int nxt = st->cur ^ 1;
    const unsigned nb = (unsigned)((n + 255) / 256);
    if (!st->next_valid) {  // no G2P has produced keys/counts for this particle set: compute them from the positions
        CKB(cudaMemsetAsync(st->cnt[nxt], 0, sizeof(uint32_t) * st->nslots, s->stream));
        if (n > 0) {
            k_bin_keys<<<nb, 256, 0, s->stream>>>(bin_key_geom(s), s->view(), n, (uint32_t)st->nslots, st->keys, st->cnt[nxt]);
            s->launches += 1;
        }
    }
    CKB(cudaMemsetAsync(st->fill, 0, sizeof(uint32_t) * st->nslots, s->stream));
    k_scan_reduce<<<(unsigned)st->ntiles, SCAN_THREADS, 0, s->stream>>>(st->cnt[nxt], st->nslots, st->tile_sums);
    k_scan_tiles<<<1, 1024, 0, s->stream>>>(st->tile_sums, (int)st->ntiles, st->misc);
    if (st->cell_bits == 9)
        k_scan_apply<9><<<(unsigned)st->ntiles, SCAN_THREADS, 0, s->stream>>>(st->cnt[nxt], st->nslots, st->tile_sums, st->cell_start, st->active, st->misc);
    else
        k_scan_apply<6><<<(unsigned)st->ntiles, SCAN_THREADS, 0, s->stream>>>(st->cnt[nxt], st->nslots, st->tile_sums, st->cell_start, st->active, st->misc);
    s->launches += 3;
    if (n > 0) {
        k_place<<<nb, 256, 0, s->stream>>>(st->keys, n, st->cnt[nxt], st->cell_start, st->fill, st->src_of);
        k_gather<<<nb, 256, 0, s->stream>>>(s->view(), s->view_alt(), st->src_of, s->orig_id, s->orig_id_alt, n);
        s->launches += 2;
        std::swap(s->part, s->part_alt);
        std::swap(s->orig_id, s->orig_id_alt);
    }
    st->cur = nxt;
    st->next_valid = false;
    // the other count buffer receives the next step's counts from G2P: clear it now
    CKB(cudaMemsetAsync(st->cnt[st->cur ^ 1], 0, sizeof(uint32_t) * st->nslots, s->stream));
    s->sorted_valid = true;
    s->steps_since_sort = 0;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { s->err = std::string("bin launch: ") + cudaGetErrorString(e); return MPM_ERR_CUDA; }
    return MPM_OK;
}

}  // namespace mpm
