// mpm_sort.cu -- particle binning: an integer LSD radix sort by cell key, hand-written for sm_100a.
//
// The reference never reorders particles (particle i keeps index i forever: SURVEY a13); binning is new.
// Key = id of the BxBxB grid block that holds the particle's base cell: the cell index formula of
// MLSMPM3DFluidMultithread.cs:282 applied to block coordinates, (bx*NBy + by)*NBz + bz, so a thread block of
// the tiled kernels owns one contiguous particle range and its grid tile fits in shared memory.  (Ordering
// by cell inside a block would buy nothing: the tile is in shared memory either way, and it costs a radix
// pass.)  The sort is stable (ties keep their previous relative order), so the permutation equals
// std::stable_sort on the same keys bit for bit (tests/test_parity_gpu.py).
//
// Bank-aware slot layout: measured on B200 (profiles/microbench/smem_atomics.cu, profiles/r1) a conflict-free
// ATOMS.ADD costs 0.78 ns/warp-instr/SM, 4.17 ns when 8 lanes hit one word, and the tiled P2G kernels were bound
// by shared-memory wavefronts (4 per ATOMS) rather than by HBM or issue slots.  So after the stable sort a
// per-block pass (k_block_layout) orders each block's particles by (rank inside the cell, cell): 32 consecutive
// slots then hold particles of 32 consecutive cells, which the padded tile of mpm_kernels_fast.cu maps to 32
// distinct banks.  The low key bits carry the cell-in-block id for that pass; the radix passes skip them.
// The rank inside a cell comes from a shared-memory atomic counter, so the slot order inside a block is not
// reproducible run to run -- nothing observable depends on it (int adds commute; downloads go through orig_id).
//
// Per pass (<= 8 key bits): k_tile_hist (per-tile digit counts) -> k_scan_rows (one CTA per digit scans
// its counts across tiles) -> k_scan_bins -> k_scatter (stable in-tile ranking with __match_any_sync,
// no atomics on the ranking path).  Then k_reorder gathers the 16 particle planes through the permutation.
#include <algorithm>
#include <string>

#include "mpm_kernels.h"
#include "mpm_solver.h"
#include "mpm_tile.cuh"

namespace mpm {

constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;  // 4096 keys per CTA
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int MAX_BINS = 256;

struct SortState {
    uint32_t* keys[2] = {nullptr, nullptr};
    uint32_t* vals[2] = {nullptr, nullptr};
    uint32_t* keys_before = nullptr;  // keys in pre-sort slot order (debug / parity)
    uint32_t* tile_hist = nullptr;    // [bins][ntiles]
    uint32_t* bin_total = nullptr;    // [bins]
    uint32_t* bin_base = nullptr;     // [bins]
    uint32_t* block_start = nullptr;  // [nblocks + 1]
    int64_t max_tiles = 0;
    int B = 8, logB = 3;              // grid block edge (cells)
    int nbx = 0, nby = 0, nbz = 0;
    int64_t nblocks = 0;
    int key_bits = 0, passes = 0, bits_per_pass = 0;
    int final_buf = 0;                // which ping-pong buffer holds the sorted result
    int64_t last_n = 0;
};

__global__ void __launch_bounds__(256) k_make_keys(KeyGeom g, ParticleView pv, int64_t n, uint32_t nkeys, uint32_t* keys, uint32_t* vals,
                                                   uint32_t* keys_before)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int cx = __float2int_rz(pv.at(PX, i)), cy = __float2int_rz(pv.at(PY, i));
    const int cz = (g.dim == 3) ? __float2int_rz(pv.at(PZ, i)) : 0;
    uint32_t k = cell_key(g, cx, cy, cz);
    k = k < nkeys ? k : nkeys - 1;  // a NaN / out-of-grid position must not produce a block id past block_start[] (the kernels skip such particles)
    keys[i] = k;
    vals[i] = (uint32_t)i;
    keys_before[i] = k >> (g.dim * g.logB);  // the sort key proper: the block id
}

__global__ void __launch_bounds__(SORT_THREADS) k_tile_hist(const uint32_t* __restrict__ keys, int64_t n, int shift, int bins,
                                                            int64_t ntiles, uint32_t* __restrict__ tile_hist)
{
    __shared__ uint32_t hist[MAX_BINS];
    for (int b = threadIdx.x; b < bins; b += SORT_THREADS) hist[b] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * SORT_TILE;
    const uint32_t mask = (uint32_t)bins - 1;
#pragma unroll 4
    for (int r = 0; r < SORT_ITEMS; ++r) {
        const int64_t i = base + (int64_t)r * SORT_THREADS + threadIdx.x;
        const bool valid = i < n;
        const uint32_t d = valid ? ((keys[i] >> shift) & mask) : 0xffffffffu;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        if (valid && (__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&hist[d], (uint32_t)__popc(peers));
    }
    __syncthreads();
    for (int b = threadIdx.x; b < bins; b += SORT_THREADS) tile_hist[(int64_t)b * ntiles + blockIdx.x] = hist[b];
}

// one CTA per digit: exclusive scan of its per-tile counts (in place) + total
__global__ void __launch_bounds__(256) k_scan_rows(uint32_t* tile_hist, int64_t ntiles, uint32_t* bin_total)
{
    __shared__ uint32_t warp_sum[8];
    __shared__ uint32_t carry_s;
    uint32_t* row = tile_hist + (int64_t)blockIdx.x * ntiles;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int64_t base = 0; base < ntiles; base += 256) {
        const int64_t i = base + threadIdx.x;
        const uint32_t v = (i < ntiles) ? row[i] : 0;
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sum[w] = x;
        __syncthreads();
        uint32_t woff = 0;
        for (int k = 0; k < w; ++k) woff += warp_sum[k];
        const uint32_t carry = carry_s;
        if (i < ntiles) row[i] = carry + woff + x - v;
        __syncthreads();
        if (threadIdx.x == 255) carry_s = carry + woff + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) bin_total[blockIdx.x] = carry_s;
}

__global__ void k_scan_bins(const uint32_t* bin_total, int bins, uint32_t* bin_base)
{
    // bins <= 256: one warp, 8 bins per lane
    const int lane = threadIdx.x;
    const int per = (bins + 31) / 32;
    uint32_t local[8];
    uint32_t sum = 0;
    for (int k = 0; k < per; ++k) {
        const int b = lane * per + k;
        local[k] = (b < bins) ? bin_total[b] : 0;
        sum += local[k];
    }
    uint32_t x = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    uint32_t run = x - sum;
    for (int k = 0; k < per; ++k) {
        const int b = lane * per + k;
        if (b < bins) bin_base[b] = run;
        run += local[k];
    }
}

__global__ void __launch_bounds__(SORT_THREADS) k_scatter(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                          uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int64_t n,
                                                          int shift, int bins, int64_t ntiles, const uint32_t* __restrict__ tile_hist,
                                                          const uint32_t* __restrict__ bin_base)
{
    __shared__ uint32_t wcnt[SORT_WARPS][MAX_BINS];  // per-warp digit counts, then per-warp offsets
    __shared__ uint32_t gbase[MAX_BINS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int b = threadIdx.x; b < SORT_WARPS * MAX_BINS; b += SORT_THREADS) (&wcnt[0][0])[b] = 0;
    __syncthreads();
    // warp w owns items [w*512, (w+1)*512) of the tile, 16 rounds of 32 consecutive keys: order = (warp, round, lane)
    const int64_t wbase = (int64_t)blockIdx.x * SORT_TILE + (int64_t)w * (32 * SORT_ITEMS);
    const uint32_t mask = (uint32_t)bins - 1;
    uint32_t key[SORT_ITEMS], val[SORT_ITEMS], rank[SORT_ITEMS];
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; ++r) {
        const int64_t i = wbase + r * 32 + lane;
        const bool valid = i < n;
        key[r] = valid ? keys_in[i] : 0xffffffffu;
        val[r] = valid ? vals_in[i] : 0;
    }
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; ++r) {
        const int64_t i = wbase + r * 32 + lane;
        const bool valid = i < n;
        const uint32_t d = valid ? ((key[r] >> shift) & mask) : 0xffffffffu;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t prev = 0;
        if (valid && lane == leader) {
            prev = wcnt[w][d];
            wcnt[w][d] = prev + (uint32_t)__popc(peers);
        }
        prev = __shfl_sync(0xffffffffu, prev, leader);
        rank[r] = prev + (uint32_t)__popc(peers & ((1u << lane) - 1u));
        __syncwarp();
    }
    __syncthreads();
    // per digit: exclusive scan over warps + global base of (digit, tile)
    for (int d = threadIdx.x; d < bins; d += SORT_THREADS) {
        uint32_t run = 0;
#pragma unroll
        for (int k = 0; k < SORT_WARPS; ++k) {
            const uint32_t c = wcnt[k][d];
            wcnt[k][d] = run;
            run += c;
        }
        gbase[d] = bin_base[d] + tile_hist[(int64_t)d * ntiles + blockIdx.x];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; ++r) {
        const int64_t i = wbase + r * 32 + lane;
        if (i < n) {
            const uint32_t d = (key[r] >> shift) & mask;
            const uint32_t dst = gbase[d] + wcnt[w][d] + rank[r];
            keys_out[dst] = key[r];
            vals_out[dst] = val[r];
        }
    }
}

// Per-block slot layout (see the header): slot order inside a block = (rank inside the cell, cell id).
// gather_src[slot] = pre-sort slot of the particle that moves there.
template <int CELL_BITS>
__global__ void __launch_bounds__(512) k_block_layout(const uint32_t* __restrict__ keys_sorted, const uint32_t* __restrict__ vals_sorted,
                                                      const uint32_t* __restrict__ block_start, uint32_t* __restrict__ rank_tmp,
                                                      uint32_t* __restrict__ gather_src)
{
    constexpr int NC = 1 << CELL_BITS;  // cells per block
    constexpr int RMAX = 16;            // rank levels laid out cell-interleaved; deeper ranks go to a cell-major tail
    __shared__ uint32_t count[NC];
    __shared__ uint16_t tab[RMAX][NC];
    __shared__ uint32_t tail_off[NC];
    __shared__ uint32_t lvl_base[RMAX + 1];
    __shared__ uint32_t wsum[16];
    const uint32_t b = blockIdx.x;
    const uint32_t s0 = block_start[b], s1 = block_start[b + 1];
    if (s0 == s1) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int L = tid; L < NC; L += 512) count[L] = 0;
    __syncthreads();
    for (uint32_t q = s0 + tid; q < s1; q += 512) rank_tmp[q] = atomicAdd(&count[keys_sorted[q] & (NC - 1)], 1u);
    __syncthreads();
    const uint32_t c = (tid < NC) ? count[tid] : 0u;
    uint32_t running = 0;
    for (int r = 0; r <= RMAX; ++r) {
        // r < RMAX: one slot per cell that has more than r particles; r == RMAX: all the remaining ones, cell-major
        const uint32_t v = (r < RMAX) ? (c > (uint32_t)r ? 1u : 0u) : (c > (uint32_t)RMAX ? c - RMAX : 0u);
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) wsum[warp] = x;
        __syncthreads();
        uint32_t woff = 0, total = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) { const uint32_t t = wsum[k]; if (k < warp) woff += t; total += t; }
        if (tid < NC) {
            if (r < RMAX) tab[r][tid] = (uint16_t)(woff + x - v);
            else tail_off[tid] = woff + x - v;
        }
        if (tid == 0) lvl_base[r] = running;
        running += total;
        __syncthreads();
    }
    for (uint32_t q = s0 + tid; q < s1; q += 512) {
        const uint32_t L = keys_sorted[q] & (NC - 1), r = rank_tmp[q];
        const uint32_t dst = (r < RMAX) ? lvl_base[r] + tab[r][L] : lvl_base[RMAX] + tail_off[L] + (r - RMAX);
        gather_src[s0 + dst] = vals_sorted[q];
    }
}

// gather the particle planes: slot i takes the particle from pre-sort slot gather_src[i]
__global__ void __launch_bounds__(256) k_reorder(ParticleView src, ParticleView dst, const uint32_t* __restrict__ gather_src,
                                                 const uint32_t* __restrict__ id_src, uint32_t* __restrict__ id_dst, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t j = gather_src[i];
    float v[NPLANES];
#pragma unroll
    for (int k = 0; k < NPLANES; ++k) v[k] = src.at(k, j);
    const uint32_t id = id_src[j];
#pragma unroll
    for (int k = 0; k < NPLANES; ++k) dst.at(k, i) = v[k];
    id_dst[i] = id;
}

// block_start[b] = first sorted rank whose block id is >= b  (b in [0, nblocks])
__global__ void __launch_bounds__(256) k_block_bounds(const uint32_t* __restrict__ keys, int64_t n, int cell_bits, int64_t nblocks,
                                                      uint32_t* __restrict__ block_start)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n == 0) {
        if (i <= nblocks) block_start[i] = 0;
        return;
    }
    if (i >= n) return;
    const int64_t b = keys[i] >> cell_bits;
    const int64_t bp = (i > 0) ? (int64_t)(keys[i - 1] >> cell_bits) : -1;
    for (int64_t bb = bp + 1; bb <= b; ++bb) block_start[bb] = (uint32_t)i;
    if (i == n - 1)
        for (int64_t bb = b + 1; bb <= nblocks; ++bb) block_start[bb] = (uint32_t)n;
}

// ---------------------------------------------------------------- host side

static int ilog2_ceil(int64_t v)
{
    int b = 0;
    while (((int64_t)1 << b) < v) ++b;
    return b;
}

#define CKS(call)                                                                 \
    do {                                                                          \
        cudaError_t e_ = (call);                                                  \
        if (e_ != cudaSuccess) {                                                  \
            s->err = std::string(#call) + ": " + cudaGetErrorString(e_);          \
            return MPM_ERR_CUDA;                                                  \
        }                                                                         \
    } while (0)

int sort_choose_block(const MpmSolver* s)
{
    // 8^3 blocks (10^3-node tile) once there are enough of them to fill the 148 SMs; 4^3 below that
    const int64_t cells = (int64_t)s->dp.nxl * s->dp.Ry * s->dp.Rz;
    return (cells >= (int64_t)96 * 96 * 96) ? 8 : 4;
}

int sort_create(MpmSolver* s)
{
    SortState* st = new SortState();
    s->sort = st;
    st->B = (s->dp.dim == 3) ? sort_choose_block(s) : 8;
    st->logB = (st->B == 8) ? 3 : 2;
    st->nbx = (s->dp.nxl + st->B - 1) / st->B;
    st->nby = (s->dp.Ry + st->B - 1) / st->B;
    st->nbz = (s->dp.dim == 3) ? (s->dp.Rz + st->B - 1) / st->B : 1;
    st->nblocks = (int64_t)st->nbx * st->nby * st->nbz;
    st->key_bits = std::max(1, ilog2_ceil(st->nblocks));  // sorted bits: the block id (above the cell-in-block bits)
    if (st->key_bits + s->dp.dim * st->logB > 32) { s->err = "grid too large for 32-bit cell keys"; return MPM_ERR_INVALID; }
    st->passes = (st->key_bits + 7) / 8;
    st->bits_per_pass = (st->key_bits + st->passes - 1) / st->passes;
    st->max_tiles = (s->pitch + SORT_TILE - 1) / SORT_TILE;
    for (int k = 0; k < 2; ++k) {
        CKS(cudaMalloc(&st->keys[k], sizeof(uint32_t) * s->pitch));
        CKS(cudaMalloc(&st->vals[k], sizeof(uint32_t) * s->pitch));
    }
    CKS(cudaMalloc(&st->keys_before, sizeof(uint32_t) * s->pitch));
    CKS(cudaMalloc(&st->tile_hist, sizeof(uint32_t) * MAX_BINS * st->max_tiles));
    CKS(cudaMalloc(&st->bin_total, sizeof(uint32_t) * MAX_BINS));
    CKS(cudaMalloc(&st->bin_base, sizeof(uint32_t) * MAX_BINS));
    CKS(cudaMalloc(&st->block_start, sizeof(uint32_t) * (st->nblocks + 1)));
    CKS(cudaMemsetAsync(st->block_start, 0, sizeof(uint32_t) * (st->nblocks + 1), s->stream));
    return MPM_OK;
}

void sort_destroy(MpmSolver* s)
{
    SortState* st = s->sort;
    if (!st) return;
    for (int k = 0; k < 2; ++k) { cudaFree(st->keys[k]); cudaFree(st->vals[k]); }
    cudaFree(st->keys_before); cudaFree(st->tile_hist); cudaFree(st->bin_total); cudaFree(st->bin_base);
    cudaFree(st->block_start);
    delete st;
    s->sort = nullptr;
}

int sort_particles(MpmSolver* s)
{
    SortState* st = s->sort;
    const int64_t n = s->n;
    st->last_n = n;
    const int cell_bits = s->dp.dim * st->logB;
    if (n == 0) {
        k_block_bounds<<<(unsigned)((st->nblocks + 256) / 256), 256, 0, s->stream>>>(nullptr, 0, cell_bits, st->nblocks, st->block_start);
        s->launches += 1;
        s->sorted_valid = true;
        s->steps_since_sort = 0;
        return MPM_OK;
    }
    KeyGeom g{s->dp.dim, st->logB, st->nby, st->nbz, s->dp.gx0 + (s->comm ? 1 : 0)};
    const unsigned nb = (unsigned)((n + 255) / 256);
    k_make_keys<<<nb, 256, 0, s->stream>>>(g, s->view(), n, (uint32_t)(st->nblocks << cell_bits), st->keys[0], st->vals[0], st->keys_before);
    s->launches += 1;
    const int64_t ntiles = (n + SORT_TILE - 1) / SORT_TILE;
    int cur = 0;
    for (int p = 0; p < st->passes; ++p) {
        const int shift = cell_bits + p * st->bits_per_pass;
        const int bits = std::min(st->bits_per_pass, st->key_bits - p * st->bits_per_pass);
        if (bits <= 0) break;
        const int bins = 1 << bits;
        k_tile_hist<<<(unsigned)ntiles, SORT_THREADS, 0, s->stream>>>(st->keys[cur], n, shift, bins, ntiles, st->tile_hist);
        k_scan_rows<<<bins, 256, 0, s->stream>>>(st->tile_hist, ntiles, st->bin_total);
        k_scan_bins<<<1, 32, 0, s->stream>>>(st->bin_total, bins, st->bin_base);
        k_scatter<<<(unsigned)ntiles, SORT_THREADS, 0, s->stream>>>(st->keys[cur], st->vals[cur], st->keys[cur ^ 1], st->vals[cur ^ 1], n,
                                                                    shift, bins, ntiles, st->tile_hist, st->bin_base);
        s->launches += 4;
        cur ^= 1;
    }
    st->final_buf = cur;
    k_block_bounds<<<nb, 256, 0, s->stream>>>(st->keys[cur], n, cell_bits, st->nblocks, st->block_start);
    // the spare ping-pong buffers hold the in-cell ranks and the gather list
    uint32_t* rank_tmp = st->keys[cur ^ 1];
    uint32_t* gather_src = st->vals[cur ^ 1];
    if (s->dp.dim == 3 && st->logB == 3)
        k_block_layout<9><<<(unsigned)st->nblocks, 512, 0, s->stream>>>(st->keys[cur], st->vals[cur], st->block_start, rank_tmp, gather_src);
    else if (s->dp.dim == 3)
        k_block_layout<6><<<(unsigned)st->nblocks, 512, 0, s->stream>>>(st->keys[cur], st->vals[cur], st->block_start, rank_tmp, gather_src);
    else
        k_block_layout<6><<<(unsigned)st->nblocks, 512, 0, s->stream>>>(st->keys[cur], st->vals[cur], st->block_start, rank_tmp, gather_src);
    k_reorder<<<nb, 256, 0, s->stream>>>(s->view(), s->view_alt(), gather_src, s->orig_id, s->orig_id_alt, n);
    s->fresh_particles = false;
    s->launches += 3;
    std::swap(s->part, s->part_alt);
    std::swap(s->orig_id, s->orig_id_alt);
    s->sorted_valid = true;
    s->steps_since_sort = 0;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { s->err = std::string("sort launch: ") + cudaGetErrorString(e); return MPM_ERR_CUDA; }
    return MPM_OK;
}

int sort_debug_last(MpmSolver* s, uint32_t* keys_before, uint32_t* perm, int64_t cap)
{
    SortState* st = s->sort;
    if (!st || !s->sorted_valid) { s->err = "no bin phase has run yet"; return MPM_ERR_STATE; }
    if (cap < st->last_n) { s->err = "destination too small"; return MPM_ERR_INVALID; }
    if (st->last_n == 0) return MPM_OK;
    if (keys_before) CKS(cudaMemcpyAsync(keys_before, st->keys_before, sizeof(uint32_t) * st->last_n, cudaMemcpyDeviceToHost, s->stream));
    if (perm) CKS(cudaMemcpyAsync(perm, st->vals[st->final_buf], sizeof(uint32_t) * st->last_n, cudaMemcpyDeviceToHost, s->stream));
    CKS(cudaStreamSynchronize(s->stream));
    return MPM_OK;
}

// ---------------------------------------------------------------- the radix sort on its own
// Stable LSD radix sort of (key, value) pairs on key bits [first_bit, first_bit + key_bits), in passes of at most 8
// bits (the kernels above).  keys[2] / vals[2] are ping-pong buffers of n entries, input in [0]; returns the index of
// the buffer that holds the result.  Used by the cell path's cold binning (mpm_bin.cu), where a particle set arrives
// in arbitrary order and the first layout has to be the stable sort by cell key.
int radix_sort_pairs(uint32_t* keys[2], uint32_t* vals[2], int64_t n, int first_bit, int key_bits, cudaStream_t stream, int64_t* launches, std::string* err)
{
    if (n <= 0 || key_bits <= 0) return 0;
    const int64_t ntiles = (n + SORT_TILE - 1) / SORT_TILE;
    uint32_t *tile_hist = nullptr, *bin_total = nullptr, *bin_base = nullptr;
    if (cudaMalloc(&tile_hist, sizeof(uint32_t) * MAX_BINS * ntiles) != cudaSuccess || cudaMalloc(&bin_total, sizeof(uint32_t) * MAX_BINS) != cudaSuccess ||
        cudaMalloc(&bin_base, sizeof(uint32_t) * MAX_BINS) != cudaSuccess) {
        cudaFree(tile_hist); cudaFree(bin_total); cudaFree(bin_base);
        if (err) *err = "radix sort: out of device memory";
        return -1;
    }
    const int passes = (key_bits + 7) / 8, per = (key_bits + passes - 1) / passes;
    int cur = 0;
    for (int p = 0; p < passes; ++p) {
        const int bits = std::min(per, key_bits - p * per);
        if (bits <= 0) break;
        const int shift = first_bit + p * per, bins = 1 << bits;
        k_tile_hist<<<(unsigned)ntiles, SORT_THREADS, 0, stream>>>(keys[cur], n, shift, bins, ntiles, tile_hist);
        k_scan_rows<<<bins, 256, 0, stream>>>(tile_hist, ntiles, bin_total);
        k_scan_bins<<<1, 32, 0, stream>>>(bin_total, bins, bin_base);
        k_scatter<<<(unsigned)ntiles, SORT_THREADS, 0, stream>>>(keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1], n, shift, bins, ntiles, tile_hist, bin_base);
        if (launches) *launches += 4;
        cur ^= 1;
    }
    cudaStreamSynchronize(stream);  // (cold path only) the work arrays are freed here
    cudaFree(tile_hist); cudaFree(bin_total); cudaFree(bin_base);
    return cur;
}

// accessors for the tiled kernels
const uint32_t* sort_block_start(const MpmSolver* s) { return s->sort->block_start; }
void sort_geometry(const MpmSolver* s, int& B, int& nbx, int& nby, int& nbz, int64_t& nblocks)
{
    B = s->sort->B; nbx = s->sort->nbx; nby = s->sort->nby; nbz = s->sort->nbz; nblocks = s->sort->nblocks;
}

}  // namespace mpm
