// mpm_bin.cu -- particle binning for the cell kernels (MPM_PATH_CELL): a one-pass counting sort by cell key.
//
// The reference never reorders particles (particle i keeps index i); binning is new.  The cell kernels
// (mpm_kernels_cell.cu) give every grid cell to one thread, which keeps the cell's 27-node stencil in registers
// across all particles of the cell.  That needs the particles of a cell to be found without searching, and the
// rows a warp works on (32 consecutive cells = one "chunk") to be contiguous.  Slot order of a step:
//
//     (block, chunk, rank r inside the cell, cell inside the chunk)
//
// where the cells of a block are first ORDERED BY PARTICLE COUNT, descending, and cut into chunks of 32: the 32
// lanes of a warp then carry nearly equal work -- and cells with more than 32 particles are first split into "virtual
// cells" of 32 so that no lane walks a pile-up alone (in cell order the warp ran to the maximum count of its 32 cells:
// 77 % lane efficiency measured on a settled lattice, ~50 % for Poisson-like counts).  Inside a chunk come first
// the rank-0 particles of its non-empty cells, then the rank-1 particles, ...  A warp at rank r reads slots
// chunk_start + S(r) + (number of lower lanes that still have a particle at rank r): consecutive addresses.
// S(r) = sum over the chunk's cells of min(count, r) is carried as a running sum of ballots.  Metadata per block:
// ord[pos] = cell, inv[cell] = pos, cnts[pos] = count, pstart[chunk] = first slot.
//
// Per step:  counts of the NEW cells are accumulated by G2P itself (fire-and-forget RED on cnt[next], key stored
// per particle), so binning = clear cursor -> block totals -> scan of the block totals (also the ordered list of
// non-empty blocks) -> per-block cell ordering + chunk starts -> k_place (rank by atomic cursor, destination
// slot from the chunk's counts; stores src_of[slot] = where the particle's record is, and moves its original index).
// No particle data moves here: the 64-byte records stay where G2P wrote them and the P2G kernels read them through
// src_of (mpm_kernels_cell.cu, RowStage).  The rank comes from an atomic, so the order of the particles INSIDE a cell
// is not reproducible run to run; the fixed-point grid sums do not depend on it (int adds commute) and the
// MPM_MATH_FAST float accumulation is covered by its stated tolerance.
#include "mpm_bin.h"

#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "mpm_kernels.h"
#include "mpm_tile.cuh"

namespace mpm {

// A binning with more "far movers" (particles that left their grid block's one-cell apron, see the stable ranking below)
// than this is ranked with the atomic cursor altogether and booked as unordered: bulk motion of more than a cell per step
// breaks the premise of the tile regions, and the exact fix-up is sized for outliers.
constexpr uint32_t FAR_LIMIT_DEFAULT = 1u << 17;  // (far_n[6] holds the limit in force: MPM_FAR_LIMIT overrides it)

#define CKB(call)                                                          \
    do {                                                                   \
        cudaError_t e_ = (call);                                           \
        if (e_ != cudaSuccess) {                                           \
            s->err = std::string(#call) + ": " + cudaGetErrorString(e_);   \
            return MPM_ERR_CUDA;                                           \
        }                                                                  \
    } while (0)

// ---- keys + counts from positions (first step, after uploads, and every step in multi-GPU mode)
template <class View>
__global__ void __launch_bounds__(256) k_bin_keys(KeyGeom g, View pv, int64_t first, int64_t n, uint32_t nslots, uint32_t* __restrict__ keys,
                                                  uint32_t* __restrict__ cnt)
{
    const int64_t i = first + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= first + n) return;
    const int cx = __float2int_rz(pv.at(PX, i)), cy = __float2int_rz(pv.at(PY, i)), cz = __float2int_rz(pv.at(PZ, i));
    uint32_t k = cell_key(g, cx, cy, cz);
    k = k < nslots ? k : nslots - 1;  // a NaN / out-of-slab position must not index outside the count array
    keys[i] = k;
    atomicAdd(&cnt[k], 1u);
}

// ---- per-block totals
template <int CELL_BITS>
__global__ void __launch_bounds__(256) k_block_sums(const uint32_t* __restrict__ cnt, int64_t nblocks, uint32_t* __restrict__ bsum)
{
    pdl_prologue();
    // one warp per grid block: 2^CELL_BITS counts (512 -> 4 uint4 per lane, 64 -> lanes 0..15 one uint4)
    const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (b >= nblocks) return;
    const uint4* p = reinterpret_cast<const uint4*>(cnt + (b << CELL_BITS));
    uint32_t sum = 0;
    constexpr int V = (1 << CELL_BITS) / 4;  // uint4 per block
#pragma unroll
    for (int k = lane; k < V; k += 32) { const uint4 v = p[k]; sum += v.x + v.y + v.z + v.w; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) bsum[b] = sum;
}

// one CTA: exclusive scan of the block totals -> bbase[nblocks + 1]; ordered list of the non-empty blocks;
// resets the per-step work counters.  Every thread takes a contiguous run of blocks (a multiple of 4, read and written
// as uint4: one SM's load/store unit handles every request of this kernel), so the CTA synchronises twice whatever the
// grid size; the round-per-1024-blocks version spent 36 us on C4's 32768 blocks, nearly all barrier and load latency.
// bsum and bbase are padded (SCAN_PAD entries, bsum's padding zero) so that the last runs may pass nblocks.
// Also the bounding box of the non-empty blocks, in cells and with the one-node apron their P2G tiles write: box[0..5] =
// {x0, x1, y0, y1, z0, z1} (half-open, global coordinates, clamped to the local grid) for this step's grid update, and
// box[6..11] = its union with the boxes of every binning since the last clear (`cleared` = a clear has run since the
// previous binning) -- whatever earlier steps left behind lies in there.
constexpr int64_t SCAN_PAD = 8192;
struct BoxGeom { int nby, nbz, B, x_owned0, gx0, nxl, Ry, Rz; };

__global__ void __launch_bounds__(1024) k_scan_blocks(const uint32_t* __restrict__ bsum, int64_t nblocks, uint32_t* __restrict__ bbase,
                                                      uint32_t* __restrict__ active, uint32_t* __restrict__ misc, BoxGeom bg,
                                                      int* __restrict__ box, int cleared, uint32_t* __restrict__ nact_out, uint32_t* __restrict__ far_n)
{
    pdl_prologue();
    if (threadIdx.x == 0) {  // stable ranking: the list of cells with far arrivals starts empty; last binning's verdict is booked
        if (far_n[3]) { far_n[1] += 1; far_n[3] = 0; }
        // A binning that had to give up (more than FAR_LIMIT far movers: bulk motion of more than a cell per step) marks the
        // scene as violent: the next 15 binnings do not even try -- far_n[2] starts above the limit, k_rank_count returns at
        // once and the placement is atomic -- and the 16th probes again.
        const uint32_t FAR_LIMIT = far_n[6];
        const bool violent = far_n[2] > FAR_LIMIT;
        far_n[4] = violent ? far_n[4] + 1 : 0;
        far_n[0] = 0; far_n[5] = 0;
        far_n[2] = (violent && (far_n[4] & 15u)) ? FAR_LIMIT + 1u : 0u;
    }
    __shared__ uint32_t wsum[32], wact[32];
    __shared__ int wbox[32][6];
    int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {-1, -1, -1};  // block coordinates of the non-empty blocks of this thread
    auto note = [&](uint32_t blk) {
        const int bz = (int)(blk % (uint32_t)bg.nbz), by = (int)(blk / (uint32_t)bg.nbz % (uint32_t)bg.nby), bx = (int)(blk / (uint32_t)(bg.nbz * bg.nby));
        lo[0] = min(lo[0], bx); hi[0] = max(hi[0], bx);
        lo[1] = min(lo[1], by); hi[1] = max(hi[1], by);
        lo[2] = min(lo[2], bz); hi[2] = max(hi[2], bz);
    };
    if (threadIdx.x < BIN_MISC_WORDS) misc[threadIdx.x] = 0;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t per4 = (((nblocks + 1023) / 1024 + 3) / 4);  // uint4 per thread
    const uint4* src = reinterpret_cast<const uint4*>(bsum) + (int64_t)threadIdx.x * per4;
    uint32_t s = 0, a = 0;
    for (int64_t k = 0; k < per4; ++k) {
        const uint4 v = src[k];
        s += v.x + v.y + v.z + v.w;
        a += (v.x > 0) + (v.y > 0) + (v.z > 0) + (v.w > 0);
    }
    uint32_t x = s, y = a;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t x2 = __shfl_up_sync(0xffffffffu, x, o), y2 = __shfl_up_sync(0xffffffffu, y, o);
        if (lane >= o) { x += x2; y += y2; }
    }
    if (lane == 31) { wsum[w] = x; wact[w] = y; }
    __syncthreads();
    if (w == 0) {
        uint32_t sx = wsum[lane], sy = wact[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t x2 = __shfl_up_sync(0xffffffffu, sx, o), y2 = __shfl_up_sync(0xffffffffu, sy, o);
            if (lane >= o) { sx += x2; sy += y2; }
        }
        wsum[lane] = sx; wact[lane] = sy;  // inclusive over warps
    }
    __syncthreads();
    uint32_t ox = (w ? wsum[w - 1] : 0) + x - s;  // exclusive prefix of the totals at this thread's first block
    uint32_t oy = (w ? wact[w - 1] : 0) + y - a;  // exclusive prefix of the non-empty flags
    uint4* dst = reinterpret_cast<uint4*>(bbase) + (int64_t)threadIdx.x * per4;
    uint32_t i = threadIdx.x * (uint32_t)(4 * per4);
    for (int64_t k = 0; k < per4; ++k, i += 4) {  // (blocks past nblocks have total 0: they get the grand total as base)
        const uint4 v = src[k];
        uint4 o;
        o.x = ox; if (v.x) { active[oy++] = i;     note(i); }     ox += v.x;
        o.y = ox; if (v.y) { active[oy++] = i + 1; note(i + 1); } ox += v.y;
        o.z = ox; if (v.z) { active[oy++] = i + 2; note(i + 2); } ox += v.z;
        o.w = ox; if (v.w) { active[oy++] = i + 3; note(i + 3); } ox += v.w;
        dst[k] = o;
    }
    if (threadIdx.x == 1023) { bbase[nblocks] = ox; misc[BIN_N_ACTIVE] = oy; *nact_out = oy; }  // (nothing but padding follows its run)
    // bounding box: warp reductions, then thread 0 over the 32 warps
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        lo[d] = __reduce_min_sync(0xffffffffu, lo[d]);
        hi[d] = __reduce_max_sync(0xffffffffu, hi[d]);
    }
    if (lane == 0) { for (int d = 0; d < 3; ++d) { wbox[w][2 * d] = lo[d]; wbox[w][2 * d + 1] = hi[d]; } }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 32; ++k)
            for (int d = 0; d < 3; ++d) { lo[d] = min(lo[d], wbox[k][2 * d]); hi[d] = max(hi[d], wbox[k][2 * d + 1]); }
        int nb[6] = {0, 0, 0, 0, 0, 0};  // empty unless there is a non-empty block
        if (hi[0] >= 0) {
            nb[0] = max(bg.x_owned0 + lo[0] * bg.B - 1, bg.gx0); nb[1] = min(bg.x_owned0 + (hi[0] + 1) * bg.B + 1, bg.gx0 + bg.nxl);
            nb[2] = max(lo[1] * bg.B - 1, 0);                    nb[3] = min((hi[1] + 1) * bg.B + 1, bg.Ry);
            nb[4] = max(lo[2] * bg.B - 1, 0);                    nb[5] = min((hi[2] + 1) * bg.B + 1, bg.Rz);
        }
        // union of: the new box, the previous binning's box, and (unless a clear ran since) the pending clear box
        int u[6] = {nb[0], nb[1], nb[2], nb[3], nb[4], nb[5]};
        auto join = [&](const int* q) {
            if (q[0] >= q[1]) return;  // empty
            if (u[0] >= u[1]) { for (int k = 0; k < 6; ++k) u[k] = q[k]; return; }
            for (int d = 0; d < 3; ++d) { u[2 * d] = min(u[2 * d], q[2 * d]); u[2 * d + 1] = max(u[2 * d + 1], q[2 * d + 1]); }
        };
        join(box);
        if (!cleared) join(box + 6);
        for (int k = 0; k < 6; ++k) { box[6 + k] = u[k]; box[k] = nb[k]; }
    }
}

// per non-empty block: split cells with more than VROWS particles into virtual cells of VROWS (within the block's budget
// of NC extra positions), order the virtual cells by particle count, descending (counting sort on min(count, 63)), so
// that the 32 virtual cells of a chunk (= the 32 lanes of a warp in the cell kernels) carry nearly equal work and no
// lane ever walks more than VROWS particles; then the start slot of every chunk.
//   ord[v0 + pos] = cell, cnts[v0 + pos] = count          (v0 = block * NV, NV = 2 * NC positions per block)
//   cellmeta[cell] = {position of the cell's first full virtual cell | position of the one holding the remainder << 16,
//                     number of full ones}
constexpr uint32_t VROWS = 32;
constexpr int STAB_ROW = 32;  // uint16 per chunk in stab[]: two 32-byte sectors (layout: k_block_order)

template <int CELL_BITS>
__global__ void __launch_bounds__(1 << CELL_BITS) k_block_order(const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ bbase,
                                                               const uint32_t* __restrict__ active, const uint32_t* __restrict__ misc,
                                                               uint16_t* __restrict__ ord, uint32_t* __restrict__ cnts,
                                                               uint2* __restrict__ cellmeta, uint32_t* __restrict__ pstart,
                                                               uint16_t* __restrict__ stab, uint32_t* __restrict__ fill, uint32_t* __restrict__ farcnt)
{
    pdl_prologue();
    constexpr int NC = 1 << CELL_BITS, NV = 2 * NC, NBIN = 64, NW = NC / 32, EXTRA = NV - NC;
    __shared__ uint32_t hist[NBIN], base[NBIN];
    __shared__ uint32_t wbin[NW][NBIN];  // per warp: cells of each bin, then the count in the lower warps
    __shared__ uint32_t s_cnt[NV];
    __shared__ uint16_t s_ord[NV];
    __shared__ uint32_t wsum[NW];
    __shared__ uint32_t carry_s;
    if (blockIdx.x >= misc[BIN_N_ACTIVE]) return;
    const uint32_t b = active[blockIdx.x];
    const uint32_t blk0 = b << CELL_BITS, v0 = b * NV;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t < NBIN) hist[t] = 0;
    for (int k = t; k < NW * NBIN; k += NC) (&wbin[0][0])[k] = 0;
    s_cnt[t] = 0; s_cnt[t + NC] = 0; s_ord[t] = 0; s_ord[t + NC] = 0;
    if (t == 0) carry_s = 0;
    __syncthreads();
    const uint32_t c = cnt[blk0 + t];
    fill[blk0 + t] = 0;  // the placement cursor and the far-arrival counter: only the cells of non-empty blocks are ever used,
    farcnt[blk0 + t] = 0;  // so they are reset here
    // extra virtual cells this cell wants, granted in cell order while the block's budget lasts
    const uint32_t want = c > VROWS ? (c + VROWS - 1) / VROWS - 1 : 0;
    uint32_t x = want;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) wsum[w] = x;
    __syncthreads();
    uint32_t pre = x - want, want_total = 0;
    for (int k = 0; k < NW; ++k) { if (k < w) pre += wsum[k]; want_total += wsum[k]; }
    const uint32_t g = pre >= (uint32_t)EXTRA ? 0u : min(want, (uint32_t)EXTRA - pre);
    const uint32_t rem = c - g * VROWS;
    const int binr = (int)min(rem, (uint32_t)(NBIN - 1));
    if (g) atomicAdd(&hist[VROWS], g);
    atomicAdd(&hist[binr], 1u);
    __syncthreads();
    if (t < NBIN) {
        uint32_t above = 0;
        for (int k = t + 1; k < NBIN; ++k) above += hist[k];
        base[t] = above;
    }
    __syncthreads();
    // position inside a bin: cells in cell order (deterministic -- an atomic cursor here made the lane a cell lands on,
    // hence the slot order, hence the NEXT binning's stable order, vary from run to run): rank of the cell among the cells
    // of its warp with the same bin (match.any), plus the bin's count in the lower warps.  Bin VROWS holds the full
    // virtual cells first (grants were handed out in cell order: the ones before this cell are min(pre, EXTRA)).
    const unsigned peers = __match_any_sync(0xffffffffu, binr);
    const uint32_t rank_w = __popc(peers & ((1u << lane) - 1u));
    if (lane == __ffs(peers) - 1) wbin[w][binr] = __popc(peers);
    __syncthreads();
    if (t < NBIN) {
        uint32_t run = 0;
        for (int k = 0; k < NW; ++k) { const uint32_t q = wbin[k][t]; wbin[k][t] = run; run += q; }
    }
    __syncthreads();
    const uint32_t granted_total = min(want_total, (uint32_t)EXTRA);
    const uint32_t pos_full = g ? base[VROWS] + min(pre, (uint32_t)EXTRA) : 0u;
    const uint32_t pos_rem = base[binr] + (binr == (int)VROWS ? granted_total : 0u) + wbin[w][binr] + rank_w;
    for (uint32_t k = 0; k < g; ++k) { s_cnt[pos_full + k] = VROWS; s_ord[pos_full + k] = (uint16_t)t; }
    s_cnt[pos_rem] = rem; s_ord[pos_rem] = (uint16_t)t;
    cellmeta[blk0 + t] = make_uint2(pos_full | (pos_rem << 16), g);  // one 8-byte load per particle in k_place
    __syncthreads();
    // counts, chunk starts and S tables: NV positions, NC threads -> two halves
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int p = t + half * NC;
        if (half == 1 && want_total == 0) {  // no cell was split (the usual case): the upper positions are simply empty
            cnts[v0 + p] = 0;
            continue;
        }
        const uint32_t v = s_cnt[p];
        cnts[v0 + p] = v;
        ord[v0 + p] = s_ord[p];
        uint32_t xs = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, xs, o); if (lane >= o) xs += y; }
        __syncthreads();  // (wsum / carry_s of the previous use are consumed)
        if (lane == 31) wsum[w] = xs;
        __syncthreads();
        const int chunk = p >> 5;
        // Chunk row for k_place.  First 32-byte sector: [0] flags chunks whose counts are not strictly ordered (a count >= 63
        // shares the last sort bin: those take the general path); [1..13] S(r) = sum over the chunk's virtual cells of
        // min(count, r), the first slot of the rank-r row; [14..15] the chunk's start slot (also in pstart[] for the walk).
        // Second sector: [16..31] = S(14..29), written only for chunks that have such rows -- the cells of a pile-up; without
        // them every particle of rank >= 14 took the general path (32 counts read and compared per particle), which in the
        // evolved dam-break is every warp of the placement.
        const uint32_t maxc = __reduce_max_sync(0xffffffffu, v);
        uint16_t* st = stab + ((size_t)(v0 >> 5) + chunk) * STAB_ROW;
        if (lane == 0) {
            uint32_t off = carry_s;
            for (int k = 0; k < w; ++k) off += wsum[k];
            const uint32_t ps = bbase[b] + off;
            pstart[(v0 >> 5) + chunk] = ps;
            st[0] = (maxc >= (uint32_t)(NBIN - 1)) ? 1 : 0;
            st[14] = (uint16_t)(ps & 0xffffu); st[15] = (uint16_t)(ps >> 16);
        }
#pragma unroll
        for (int r = 1; r < 14; ++r) {
            const uint32_t sr = __reduce_add_sync(0xffffffffu, min(v, (uint32_t)r));
            if (lane == 0) st[r] = (uint16_t)sr;
        }
        if (maxc > 14u) {  // (uniform over the warp)
#pragma unroll
            for (int r = 14; r < 30; ++r) {
                const uint32_t sr = __reduce_add_sync(0xffffffffu, min(v, (uint32_t)r));
                if (lane == 0) st[r + 2] = (uint16_t)sr;
            }
        }
        __syncthreads();
        if (t == NC - 1) {
            uint32_t tot = carry_s;
            for (int k = 0; k < NW; ++k) tot += wsum[k];
            carry_s = tot;
        }
    }
}

// general case of place_slot (a chunk whose counts are not strictly ordered, or a row beyond the chunk's S table)
__device__ __noinline__ uint32_t place_slot_general(const uint32_t* __restrict__ cnts_chunk, uint32_t chunk_start, uint32_t lane, uint32_t r)
{
    const uint4* c4 = reinterpret_cast<const uint4*>(cnts_chunk);
    uint32_t below = 0;   // sum over the chunk's virtual cells of min(count, r)
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
    return chunk_start + below + before;
}

// rank inside the (real) cell -> (virtual cell, row) -> destination slot from the chunk's counts
template <int CELL_BITS>
__device__ __forceinline__ uint32_t place_slot(uint32_t key, uint32_t rc, const uint2* __restrict__ cellmeta, const uint32_t* __restrict__ cnts,
                                               const uint32_t* __restrict__ pstart, const uint16_t* __restrict__ stab)
{
    constexpr uint32_t NV = 2u << CELL_BITS;
    const uint2 cm = cellmeta[key];  // x: position of the first full virtual cell | position of the remainder << 16; y: full ones
    const uint32_t g = cm.y;
    uint32_t pos, r;
    if (rc < g * VROWS) { pos = (cm.x & 0xffffu) + rc / VROWS; r = rc % VROWS; }
    else { pos = cm.x >> 16; r = rc - g * VROWS; }
    const uint32_t v0 = (key >> CELL_BITS) * NV;
    const uint32_t chunk = pos >> 5, lane = pos & 31u;
    const uint32_t gchunk = (v0 >> 5) + chunk;
    const uint16_t* row = stab + (size_t)gchunk * STAB_ROW;
    if (r < 30u && row[0] == 0) {
        // counts strictly ordered, descending: every lower lane still has a particle at rank r (r < own count <= theirs)
        const uint32_t below = r ? row[r < 14u ? r : r + 2u] : 0u;  // (rows 14..29: the chunk's second sector)
        const uint32_t ps = *reinterpret_cast<const uint32_t*>(row + 14);
        return ps + below + lane;
    }
    return place_slot_general(cnts + v0 + chunk * 32u, pstart[gchunk], lane, r);
}

// Atomic ranking (multi-GPU slabs, where migration leaves the slots in no particular order; MPM_ATOMIC_BINNING=1): the rank
// inside the cell comes from an atomic cursor, so the order of a cell's particles is not reproducible run to run.
template <int CELL_BITS>
__global__ void __launch_bounds__(256) k_place(const uint32_t* __restrict__ keys, int64_t n, const uint2* __restrict__ cellmeta,
                                               const uint32_t* __restrict__ cnts, const uint32_t* __restrict__ pstart,
                                               const uint16_t* __restrict__ stab, uint32_t* __restrict__ fill, uint32_t* __restrict__ src_of,
                                               const uint32_t* __restrict__ id_src, uint32_t* __restrict__ id_dst, const uint32_t* __restrict__ n_dev)
{
    pdl_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (n_dev ? (int64_t)*n_dev : n)) return;  // (multi-GPU: the count of a migration the host has not read yet)
    const uint32_t key = keys[i];
    const uint32_t id = id_src[i];  // the particle's original index moves to its new slot here (coalesced read, one more
                                    // scattered 4-byte store) rather than as a scattered read in P2G_1
    const uint32_t rc = atomicAdd(&fill[key], 1u);  // rank inside the (real) cell
    const uint32_t dest = place_slot<CELL_BITS>(key, rc, cellmeta, cnts, pstart, stab);
    src_of[dest] = (uint32_t)i;
    id_dst[dest] = id;
}

// ---------------------------------------------------------------- stable rank inside a cell (the default on one GPU)
// The binning permutation is the STABLE sort of the particles by cell key: inside a cell they keep the order of their old
// slots, so the layout -- and with it the fp32 accumulation order of the cell kernels -- is a pure function of the
// particle state: two runs give the same bits, and the permutation equals std::stable_sort on the same keys.
//
// A full radix sort every step would cost more than the atomic placement it replaces; the ranking below uses what the
// previous binning left: the old slots are grouped by old grid block ("tile" T = one contiguous run, tiles in block
// order), and a particle moves less than a cell per step, so the particles of tile T land in T's (B+2)^3 region of cells
// -- the same region as the P2G tile.  For a cell c and an old slot i in tile T
//     rank(i) = sum over the tiles T' < T whose region holds c of count(T', c)      (at most 7 tiles: c's own block and
//                                                                                    its face / edge / corner neighbours)
//             + number of slots j < i of T with the same cell.
// k_rank_count writes count(T, .) per tile (shared-memory counters, one row of (B+2)^3 words per tile); k_rank_place
// recounts per warp (warp w of the CTA owns the w-th quarter of the tile's rows), turns the counts into starting ranks
// (neighbour offset + prefix over the warps) and walks its rows in order: rank = counter + lower lanes of the row with
// the same cell (match.any), leader lane bumps the counter.  No global atomics, no dependence on scheduling.
//
// Particles that leave their tile's region (more than a cell per step past the block's edge: the fast part of a violent
// scene -- 2 % of the particles of the C4 dam-break after 100 steps) are "far movers".  They are placed BEHIND the regular
// arrivals of their cell (rank = count - far arrivals + an atomic ticket), and k_fix_far then restores the exact order
// cell by cell: the regular part of a cell is already in slot order, its few far arrivals are sorted and merged in.
// A cell with more than FIX_FAR_MAX far arrivals, or more than FIX_CAP such cells in one binning, stays as placed (a valid
// cell sort in atomic order) and the binning is counted in MpmStats.unordered_binnings.
struct RankGeom { int nbx, nby, nbz; };

// tiles with more rows than this are ranked by a whole CTA (k_rank_place_heavy), the others by one warp each; k_rank_count
// lists them (far_n[5] = how many; beyond HEAVY_CAP the heavy kernel falls back to scanning every tile)
#ifndef MPM_RANK_CTAS
#define MPM_RANK_CTAS 8  // CTAs of k_rank_place per SM (64 registers; 10 and 12 were measured: see profiles/r2/README.md)
#endif
constexpr uint32_t HEAVY_ROWS = 256;
constexpr int HEAVY_CAP = 4096;

template <int CELL_BITS>
struct RankCfg {
    static constexpr int LOGB = CELL_BITS / 3, B = 1 << LOGB, T = B + 2, RC = T * T * T;
    static constexpr int W = 4, THREADS = 32 * W;
};

// index of the cell `key` in the region of tile (tbx, tby, tbz) = block id `tile`, or -1 (a far mover).
// Nearly every particle stays in its block (first branch); one that crossed into a neighbouring block is decoded from the
// DIFFERENCE of the block ids -- d = dx nby nbz + dy nbz + dz with |d*| <= 1 splits by two roundings when nby, nbz >= 3 --
// and accepted only if that neighbour exists (at the grid's edge the same difference can also mean a block further away:
// those, and grids of fewer than 3 blocks across, take the divisions).
struct TileCtx {
    uint32_t tile;
    int tbx, tby, tbz;
};
__device__ __forceinline__ TileCtx tile_ctx(uint32_t tile, const RankGeom& g)
{
    TileCtx c;
    c.tile = tile;
    c.tbz = (int)(tile % (uint32_t)g.nbz); c.tby = (int)(tile / (uint32_t)g.nbz % (uint32_t)g.nby); c.tbx = (int)(tile / (uint32_t)(g.nbz * g.nby));
    return c;
}
template <int CELL_BITS>
__device__ __noinline__ int region_index_moved(uint32_t key, uint32_t tile, int tbx, int tby, int tbz, int nbx, int nby, int nbz)
{
    using C = RankCfg<CELL_BITS>;
    const uint32_t blk = key >> CELL_BITS;
    const int lx = (key >> (2 * C::LOGB)) & (C::B - 1), ly = (key >> C::LOGB) & (C::B - 1), lz = key & (C::B - 1);
    const int s_yz = nby * nbz;
    int dx, dy, dz;
    const int delta = (int)blk - (int)tile;
    bool decoded = false;
    if (nby >= 3 && nbz >= 3 && delta > -2 * s_yz && delta < 2 * s_yz) {
        dx = __float2int_rn((float)delta / (float)s_yz);
        const int rem = delta - dx * s_yz;
        dy = __float2int_rn((float)rem / (float)nbz);
        dz = rem - dy * nbz;
        decoded = dx >= -1 && dx <= 1 && dy >= -1 && dy <= 1 && dz >= -1 && dz <= 1 && (unsigned)(tbx + dx) < (unsigned)nbx &&
                  (unsigned)(tby + dy) < (unsigned)nby && (unsigned)(tbz + dz) < (unsigned)nbz;
    }
    if (!decoded) {
        const int bz = (int)(blk % (uint32_t)nbz), by = (int)(blk / (uint32_t)nbz % (uint32_t)nby), bx = (int)(blk / (uint32_t)(nbz * nby));
        dx = bx - tbx; dy = by - tby; dz = bz - tbz;
        if (dx < -1 || dx > 1 || dy < -1 || dy > 1 || dz < -1 || dz > 1) return -1;
    }
    const int rx = dx * C::B + lx + 1, ry = dy * C::B + ly + 1, rz = dz * C::B + lz + 1;
    if ((unsigned)rx >= (unsigned)C::T || (unsigned)ry >= (unsigned)C::T || (unsigned)rz >= (unsigned)C::T) return -1;
    return (rx * C::T + ry) * C::T + rz;
}
template <int CELL_BITS>
__device__ __forceinline__ int region_index(uint32_t key, const TileCtx& c, const RankGeom& g)
{
    using C = RankCfg<CELL_BITS>;
    if ((key >> CELL_BITS) == c.tile) {
        const int lx = (key >> (2 * C::LOGB)) & (C::B - 1), ly = (key >> C::LOGB) & (C::B - 1), lz = key & (C::B - 1);
        return ((lx + 1) * C::T + (ly + 1)) * C::T + (lz + 1);
    }
    return region_index_moved<CELL_BITS>(key, c.tile, c.tbx, c.tby, c.tbz, g.nbx, g.nby, g.nbz);  // (out of line: rare, and large)
}

template <int CELL_BITS>
__global__ void __launch_bounds__(RankCfg<CELL_BITS>::THREADS, 12) k_rank_count(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ bbase_prev,
                                                                                const uint32_t* __restrict__ active_prev, const uint32_t* __restrict__ nact_prev,
                                                                                RankGeom g, uint32_t* __restrict__ tcount, uint32_t* __restrict__ farcnt,
                                                                                uint32_t* __restrict__ fixlist, uint32_t* __restrict__ far_n,
                                                                                uint32_t* __restrict__ heavy)
{
    pdl_wait();
    using C = RankCfg<CELL_BITS>;
    constexpr int KB = 8;  // keys per thread and batch: their loads are issued together
    __shared__ uint32_t cnt[C::RC];
    const uint32_t FAR_LIMIT = far_n[6];
    if (far_n[2] > FAR_LIMIT) return;  // (a violent scene, not probing this time: k_scan_blocks)
    const uint32_t na = *nact_prev;
    for (uint32_t t = blockIdx.x; t < na; t += gridDim.x) {
        const uint32_t tile = active_prev[t];
        const TileCtx tc = tile_ctx(tile, g);
        const int tbx = tc.tbx, tby = tc.tby, tbz = tc.tbz;
        for (int k = threadIdx.x; k < C::RC; k += C::THREADS) cnt[k] = 0;
        __syncthreads();
        const uint32_t s0 = bbase_prev[tile], s1 = bbase_prev[tile + 1];
        if (threadIdx.x == 0 && ((s1 - s0 + 31u) >> 5) > HEAVY_ROWS) {  // the pile-up tiles get a whole CTA in k_rank_place_heavy
            const uint32_t h = atomicAdd(far_n + 5, 1u);
            if (h < (uint32_t)HEAVY_CAP) heavy[h] = tile;
        }
        const bool gave_up = *reinterpret_cast<volatile uint32_t*>(far_n + 2) > FAR_LIMIT;  // (then MpmStats.far_movers is a lower bound)
        for (uint32_t base = s0; base < s1; base += C::THREADS * KB) {
            uint32_t key[KB];
#pragma unroll
            for (int j = 0; j < KB; ++j) {
                const uint32_t i = base + j * C::THREADS + threadIdx.x;
                key[j] = i < s1 ? keys[i] : 0xffffffffu;
            }
#pragma unroll
            for (int j = 0; j < KB; ++j) {
                const uint32_t i = base + j * C::THREADS + threadIdx.x;
                const int r = i < s1 ? region_index<CELL_BITS>(key[j], tc, g) : 0;
                const bool far = i < s1 && r < 0;
                if (i < s1 && r >= 0) atomicAdd(&cnt[r], 1u);
                // far movers: counted per target cell (farcnt[] was cleared with the block's layout); the first one of a cell
                // lists the cell.  The two global counters are bumped once per warp, not per particle (a violent scene has
                // hundreds of thousands of far movers: per-particle atomics on one address took over a millisecond), and not at
                // all once the binning is past the point where it gives up on the stable order anyway.
                const unsigned mf = __ballot_sync(0xffffffffu, far && !gave_up);
                if (mf) {
                    const int lane = threadIdx.x & 31;
                    const bool first = ((mf >> lane) & 1u) && atomicAdd(&farcnt[key[j]], 1u) == 0u;
                    const unsigned ml = __ballot_sync(0xffffffffu, first);
                    uint32_t at = 0;
                    if (lane == __ffs(mf) - 1) {
                        atomicAdd(far_n + 2, (uint32_t)__popc(mf));
                        if (ml) at = atomicAdd(far_n, (uint32_t)__popc(ml));
                    }
                    at = __shfl_sync(0xffffffffu, at, __ffs(mf) - 1);
                    if (first) {
                        const uint32_t f = at + (uint32_t)__popc(ml & ((1u << lane) - 1u));
                        if (f < (uint32_t)FIX_CAP) fixlist[f] = key[j];
                    }
                }
            }
        }
        __syncthreads();
        for (int k = threadIdx.x; k < C::RC; k += C::THREADS) tcount[(size_t)tile * C::RC + k] = cnt[k];
        __syncthreads();
    }
}

struct RankArgs {
    const uint32_t* keys;
    const uint32_t* bsum_prev;
    const uint32_t* bbase_prev;
    const uint32_t* active_prev;
    const uint32_t* nact_prev;
    RankGeom g;
    const uint32_t* tcount;
    const uint32_t* cnt;     // particles per new cell (what G2P counted)
    const uint2* cellmeta;
    const uint32_t* cnts;
    const uint32_t* pstart;
    const uint16_t* stab;
    uint32_t* fill;      // atomic placement cursor (only when a binning gives up on the stable order)
    uint32_t* farcnt;    // far arrivals per cell: low half = count (k_rank_count), high half = tickets handed out
    uint32_t* far_n;
    const uint32_t* heavy;  // tiles for k_rank_place_heavy (k_rank_count)
    uint32_t n_total;    // particles
    uint32_t* src_of;
    const uint32_t* id_src;
    uint32_t* id_dst;
};




// One tile, W warps.  W = 1: a warp on its own (no block-wide barrier anywhere: the usual tile of ~100 rows is latency-bound,
// and thousands of independent warps hide that); W > 1: a CTA of W warps for the pile-up tiles of an evolved scene, each
// warp owning the w-th part of the rows, with a counting pass first so that every warp knows where its ranks start.
template <int CELL_BITS, int W>
__device__ __forceinline__ void rank_tile(const RankArgs& A, uint32_t tile, uint32_t s0, uint32_t s1, uint32_t* wcnt, uint32_t* off, uint32_t* nb_tile)
{
    using C = RankCfg<CELL_BITS>;
    constexpr int GS = 32 * W;  // threads working on the tile
    constexpr int RB = 4;       // rows per batch: their loads are issued together
    const RankGeom& g = A.g;
    const int lane = threadIdx.x & 31, w = (W == 1) ? 0 : (int)(threadIdx.x >> 5), gt = (W == 1) ? lane : (int)threadIdx.x;
    const unsigned lt = (1u << lane) - 1u;
    auto sync = [&]() { if (W == 1) __syncwarp(); else __syncthreads(); };
    const TileCtx tc = tile_ctx(tile, g);
    const int tbx = tc.tbx, tby = tc.tby, tbz = tc.tbz;
    for (int k = gt; k < W * C::RC; k += GS) wcnt[k] = 0;
    if (W > 1) for (int k = gt; k < C::RC; k += GS) off[k] = 0;
    // the 26 neighbouring tiles: those that come before this one and held particles contribute to the cells their
    // region shares with ours (a box of 2 or T cells per axis)
    if (gt < 27) {
        const int dz = gt % 3 - 1, dy = gt / 3 % 3 - 1, dx = gt / 9 - 1;
        const int bx = tbx + dx, by = tby + dy, bz = tbz + dz;
        uint32_t t2 = 0xffffffffu;
        if ((dx | dy | dz) != 0 && bx >= 0 && by >= 0 && bz >= 0 && bx < g.nbx && by < g.nby && bz < g.nbz) {
            t2 = (uint32_t)((bx * g.nby + by) * g.nbz + bz);
            if (t2 >= tile || A.bsum_prev[t2] == 0) t2 = 0xffffffffu;
        }
        nb_tile[gt] = t2;
    }
    sync();
    const uint32_t nrows = (s1 - s0 + 31u) >> 5, rpw = (nrows + W - 1) / W;
    const uint32_t r0 = min((uint32_t)w * rpw, nrows), r1 = min(r0 + rpw, nrows);
    if (W > 1) {
        // 1. counts of this warp's rows
        for (uint32_t row = r0; row < r1; row += 8) {
            uint32_t key[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t i = s0 + (row + j) * 32u + lane;
                key[j] = (row + j < r1 && i < s1) ? A.keys[i] : 0xffffffffu;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t i = s0 + (row + j) * 32u + lane;
                if (row + j < r1 && i < s1) {
                    const int r = region_index<CELL_BITS>(key[j], tc, g);
                    if (r >= 0) atomicAdd(&wcnt[w * C::RC + r], 1u);
                }
            }
        }
    }
    sync();
    // 2a. arrivals from lower tiles: for every such neighbour, the box of cells its region shares with ours (2 or T cells
    // per axis); the threads stride over the box, a neighbour's counts are one contiguous row of tcount
    uint32_t* acc = (W == 1) ? wcnt : off;
#pragma unroll 1
    for (int nb = 0; nb < 27; ++nb) {
        const uint32_t t2 = nb_tile[nb];
        if (t2 == 0xffffffffu) continue;
        const int dz = nb % 3 - 1, dy = nb / 3 % 3 - 1, dx = nb / 9 - 1;
        const int ey = dy ? 2 : C::T, ez = dz ? 2 : C::T;
        const int ncell = (dx ? 2 : C::T) * ey * ez;
        const uint32_t* row = A.tcount + (size_t)t2 * C::RC;
        // our region coordinate r = (d > 0 ? B : 0) + c and the neighbour's r' = r - d * B, per axis
        const int ox = dx > 0 ? C::B : 0, oy = dy > 0 ? C::B : 0, oz = dz > 0 ? C::B : 0;
        for (int c = gt; c < ncell; c += GS) {
            int cz, cy, cx, q = c;
            if (dz) { cz = q & 1; q >>= 1; } else { cz = q % C::T; q /= C::T; }
            if (dy) { cy = q & 1; q >>= 1; } else { cy = q % C::T; q /= C::T; }
            cx = q;
            const int rx = ox + cx, ry = oy + cy, rz = oz + cz;
            const uint32_t v = row[((rx - dx * C::B) * C::T + (ry - dy * C::B)) * C::T + (rz - dz * C::B)];
            if (v) atomicAdd(&acc[(rx * C::T + ry) * C::T + rz], v);
        }
    }
    sync();
    if (W > 1) {
        // 2b. starting rank of every (warp, region cell): arrivals from lower tiles, then the warps in order
        for (int k = gt; k < C::RC; k += GS) {
            uint32_t o = off[k];
#pragma unroll
            for (int q = 0; q < W; ++q) { const uint32_t c = wcnt[q * C::RC + k]; wcnt[q * C::RC + k] = o; o += c; }
        }
        sync();
    }
    // 3. the warp's rows in order: rank = counter + lower lanes of the row with the same cell
    uint32_t* ctr = wcnt + w * C::RC;
    for (uint32_t row = r0; row < r1; row += RB) {
        uint32_t key[RB], id[RB], rc[RB];
        int rg[RB];
        bool valid[RB];
#pragma unroll
        for (int j = 0; j < RB; ++j) {
            const uint32_t i = s0 + (row + j) * 32u + lane;
            valid[j] = row + j < r1 && i < s1;
            key[j] = valid[j] ? A.keys[i] : 0u;
            id[j] = valid[j] ? A.id_src[i] : 0u;
        }
        // Per row: the lanes with the same cell form a group (match.any); its lowest lane bumps the cell's counter by the
        // group's size and hands the old value to the others.  The four atomics of a batch are issued back to back --
        // shared-memory accesses of one warp are performed in order, so row j's sees row j - 1's -- and the returned
        // values are collected afterwards: one shared-memory round trip per batch instead of a load -> store chain per row.
        unsigned peers[RB];
        uint32_t old[RB];
#pragma unroll
        for (int j = 0; j < RB; ++j) {
            rg[j] = valid[j] ? region_index<CELL_BITS>(key[j], tc, g) : -1;
            const uint32_t tag = rg[j] >= 0 ? (uint32_t)rg[j] : (0x80000000u | (uint32_t)lane);
            peers[j] = __match_any_sync(0xffffffffu, tag);
            old[j] = 0;
            if (rg[j] >= 0 && lane == __ffs(peers[j]) - 1) old[j] = atomicAdd(&ctr[rg[j]], (uint32_t)__popc(peers[j]));
            __syncwarp();
        }
#pragma unroll
        for (int j = 0; j < RB; ++j) rc[j] = __shfl_sync(0xffffffffu, old[j], __ffs(peers[j]) - 1) + (uint32_t)__popc(peers[j] & lt);
        // far movers go behind their cell's regular arrivals: ticket from the high half of farcnt[] (the low half keeps the
        // cell's far count for k_fix_far, which puts them where the stable order wants them)
#pragma unroll
        for (int j = 0; j < RB; ++j) {
            if (valid[j] && rg[j] < 0) {
                const uint32_t oldf = atomicAdd(&A.farcnt[key[j]], 0x10000u);
                rc[j] = A.cnt[key[j]] - (oldf & 0xffffu) + (oldf >> 16);
            }
        }
#pragma unroll
        for (int j = 0; j < RB; ++j) {
            if (!valid[j]) continue;
            const uint32_t i = s0 + (row + j) * 32u + lane;
            const uint32_t dest = place_slot<CELL_BITS>(key[j], rc[j], A.cellmeta, A.cnts, A.pstart, A.stab);
            A.src_of[dest] = i;
            A.id_dst[dest] = id[j];
        }
    }
    sync();
}

// light tiles: one warp per tile, four tiles per CTA
template <int CELL_BITS>
__global__ void __launch_bounds__(128, MPM_RANK_CTAS) k_rank_place(const __grid_constant__ RankArgs A)
{
    pdl_wait();
    using C = RankCfg<CELL_BITS>;
    __shared__ uint32_t wcnt[4][C::RC];
    __shared__ uint32_t nb_tile[4][28];
    const int w = threadIdx.x >> 5;
    if (A.far_n[2] > A.far_n[6]) {
        // too violent a step for the stable order (uniform over the launch): every particle takes its rank from the atomic
        // cursor, all threads of the grid striding over the particles, four loads in flight each
        if (blockIdx.x == 0 && threadIdx.x == 0) A.far_n[3] = 1;
        const uint32_t n = A.n_total, stride = gridDim.x * blockDim.x;
        for (uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += 4 * stride) {
            uint32_t key[4], id[4], rk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { const uint32_t i = i0 + j * stride; key[j] = i < n ? A.keys[i] : 0u; id[j] = i < n ? A.id_src[i] : 0u; }
#pragma unroll
            for (int j = 0; j < 4; ++j) rk[j] = (i0 + j * stride < n) ? atomicAdd(&A.fill[key[j]], 1u) : 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t i = i0 + j * stride;
                if (i >= n) continue;
                const uint32_t dest = place_slot<CELL_BITS>(key[j], rk[j], A.cellmeta, A.cnts, A.pstart, A.stab);
                A.src_of[dest] = i;
                A.id_dst[dest] = id[j];
            }
        }
        return;
    }
    const uint32_t na = *A.nact_prev;
    for (uint32_t t = blockIdx.x * 4 + w; t < na; t += gridDim.x * 4) {
        const uint32_t tile = A.active_prev[t];
        const uint32_t s0 = A.bbase_prev[tile], s1 = A.bbase_prev[tile + 1];
        if (((s1 - s0 + 31u) >> 5) > HEAVY_ROWS) continue;  // (k_rank_place_heavy)
        rank_tile<CELL_BITS, 1>(A, tile, s0, s1, wcnt[w], nullptr, nb_tile[w]);
    }
}

// heavy tiles: a CTA of 8 warps per tile
template <int CELL_BITS>
__global__ void __launch_bounds__(256) k_rank_place_heavy(const __grid_constant__ RankArgs A)
{
    pdl_wait();
    using C = RankCfg<CELL_BITS>;
    constexpr int W = 8;
    __shared__ uint32_t wcnt[W * C::RC];
    __shared__ uint32_t off[C::RC];
    __shared__ uint32_t nb_tile[28];
    if (A.far_n[2] > A.far_n[6]) return;  // (k_rank_place ranks everything atomically)
    const uint32_t nheavy = A.far_n[5];
    const bool listed = nheavy <= (uint32_t)HEAVY_CAP;
    const uint32_t na = listed ? nheavy : *A.nact_prev;
    for (uint32_t t = blockIdx.x; t < na; t += gridDim.x) {
        const uint32_t tile = listed ? A.heavy[t] : A.active_prev[t];
        const uint32_t s0 = A.bbase_prev[tile], s1 = A.bbase_prev[tile + 1];
        if (((s1 - s0 + 31u) >> 5) <= HEAVY_ROWS) continue;  // (uniform over the CTA)
        rank_tile<CELL_BITS, W>(A, tile, s0, s1, wcnt, off, nb_tile);
    }
}

// One thread per cell that received far movers: its regular arrivals sit at ranks [0, c - f) in slot order, the f far
// arrivals behind them in ticket order.  Sort the far ones by old slot, then merge from the back (in place: a write never
// passes the regular entry that is read next).  far_n[3] flags cells that are left as they are.
constexpr int FIX_FAR_MAX = 32;
template <int CELL_BITS>
__global__ void __launch_bounds__(128) k_fix_far(const uint32_t* __restrict__ fixlist, uint32_t* __restrict__ far_n, const uint32_t* __restrict__ cnt,
                                                 const uint32_t* __restrict__ fill, const uint2* __restrict__ cellmeta, const uint32_t* __restrict__ cnts,
                                                 const uint32_t* __restrict__ pstart, const uint16_t* __restrict__ stab, uint32_t* __restrict__ src_of,
                                                 uint32_t* __restrict__ ids)
{
    pdl_wait();
    const uint32_t listed = far_n[0];
    if (listed == 0 || far_n[2] > far_n[6]) return;
    if (listed > (uint32_t)FIX_CAP && blockIdx.x == 0 && threadIdx.x == 0) far_n[3] = 1;
    const uint32_t ncell = min(listed, (uint32_t)FIX_CAP);
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < ncell; q += gridDim.x * blockDim.x) {
        const uint32_t key = fixlist[q];
        const uint32_t c = cnt[key], f = fill[key] & 0xffffu;
        if (f == 0 || f > c) continue;
        if (f > (uint32_t)FIX_FAR_MAX) { far_n[3] = 1; continue; }
        auto slot = [&](uint32_t r) { return place_slot<CELL_BITS>(key, r, cellmeta, cnts, pstart, stab); };
        uint32_t fv[FIX_FAR_MAX], fi[FIX_FAR_MAX];
        for (uint32_t k = 0; k < f; ++k) {  // insertion sort of the far arrivals by old slot
            const uint32_t sl = slot(c - f + k);
            const uint32_t v = src_of[sl], id = ids[sl];
            uint32_t j = k;
            while (j > 0 && fv[j - 1] > v) { fv[j] = fv[j - 1]; fi[j] = fi[j - 1]; --j; }
            fv[j] = v; fi[j] = id;
        }
        // merge from the back
        int a = (int)(c - f) - 1, b = (int)f - 1;  // last regular, last far
        uint32_t rv = 0, rid = 0;
        bool have = false;
        for (int w = (int)c - 1; b >= 0; --w) {
            if (a >= 0 && !have) { const uint32_t sl = slot((uint32_t)a); rv = src_of[sl]; rid = ids[sl]; have = true; }
            const uint32_t sw = slot((uint32_t)w);
            if (a >= 0 && rv > fv[b]) { src_of[sw] = rv; ids[sw] = rid; --a; have = false; }
            else { src_of[sw] = fv[b]; ids[sw] = fi[b]; --b; }
        }
    }
}

// check of a layout against ranks given from outside (mpm_debug_last_sort: the host derives them from the definition)
template <int CELL_BITS>
__global__ void __launch_bounds__(256) k_verify_layout(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ rank, int64_t n,
                                                       const uint2* __restrict__ cellmeta, const uint32_t* __restrict__ cnts,
                                                       const uint32_t* __restrict__ pstart, const uint16_t* __restrict__ stab,
                                                       const uint32_t* __restrict__ src_of, const uint32_t* __restrict__ id_src,
                                                       const uint32_t* __restrict__ id_dst, uint32_t* __restrict__ bad)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t dest = place_slot<CELL_BITS>(keys[i], rank[i], cellmeta, cnts, pstart, stab);
    if (src_of[dest] != (uint32_t)i || id_dst[dest] != id_src[i]) atomicAdd(bad, 1u);
}

// grouped planes -> 64-byte records (a freshly uploaded or edited particle set enters the cell path)
__global__ void __launch_bounds__(256) k_planes_to_rec(ParticleView pv, float4* __restrict__ rec, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* q = pv.rec(i);
    rec[4 * i + 0] = make_float4(q[PX * GROUP], q[PY * GROUP], q[PZ * GROUP], q[PM * GROUP]);
    rec[4 * i + 1] = make_float4(q[VX * GROUP], q[VY * GROUP], q[VZ * GROUP], q[C2 * GROUP]);
    rec[4 * i + 2] = make_float4(q[C0 * GROUP], q[C1 * GROUP], q[C3 * GROUP], q[C4 * GROUP]);
    rec[4 * i + 3] = make_float4(q[C6 * GROUP], q[C7 * GROUP], q[C5 * GROUP], q[C8 * GROUP]);
}

// What P2G_1 leaves for G2P -- position and mass planes in slot order -- for a G2P phase that is run without a P2G_1
// since the last binning (mpm_run_phase).
__global__ void __launch_bounds__(256) k_gather_g2p_inputs(const float4* __restrict__ rec, ParticleView dst, const uint32_t* __restrict__ src_of, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t j = src_of[i];
    reinterpret_cast<float4*>(dst.base)[i] = rec[4 * (size_t)j];  // (px, py, pz, m): one 16-byte element per slot, as P2G_1 leaves them
}

// ---- cold binning: a particle set in arbitrary order (upload, scene edit) is first brought into cell-key order by a
// stable radix sort, so that the stable ranking above finds its tiles
__global__ void __launch_bounds__(256) k_cold_keys(KeyGeom g, ParticleView pv, int64_t n, uint32_t nslots, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t k = cell_key(g, __float2int_rz(pv.at(PX, i)), __float2int_rz(pv.at(PY, i)), __float2int_rz(pv.at(PZ, i)));
    keys[i] = k < nslots ? k : nslots - 1;  // (as k_bin_keys)
    vals[i] = (uint32_t)i;
}

// planes (upload order) -> records in sorted order; keys and counts of that order; original indices follow
__global__ void __launch_bounds__(256) k_cold_gather(ParticleView pv, const uint32_t* __restrict__ skeys, const uint32_t* __restrict__ svals, int64_t n,
                                                     float4* __restrict__ rec, uint32_t* __restrict__ keys, uint32_t* __restrict__ cnt,
                                                     const uint32_t* __restrict__ id_src, uint32_t* __restrict__ id_dst)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t i = svals[p], k = skeys[p];
    const float* q = pv.rec(i);
    rec[4 * p + 0] = make_float4(q[PX * GROUP], q[PY * GROUP], q[PZ * GROUP], q[PM * GROUP]);
    rec[4 * p + 1] = make_float4(q[VX * GROUP], q[VY * GROUP], q[VZ * GROUP], q[C2 * GROUP]);
    rec[4 * p + 2] = make_float4(q[C0 * GROUP], q[C1 * GROUP], q[C3 * GROUP], q[C4 * GROUP]);
    rec[4 * p + 3] = make_float4(q[C6 * GROUP], q[C7 * GROUP], q[C5 * GROUP], q[C8 * GROUP]);
    keys[p] = k;
    id_dst[p] = id_src[i];
    atomicAdd(&cnt[k], 1u);
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
    if (const char* e = getenv("MPM_BLOCK_EDGE")) { if (atoi(e) == 4 || atoi(e) == 8) st->B = atoi(e); }  // (A/B: profiles/r2/README.md)
    st->logB = (st->B == 8) ? 3 : 2;
    st->cell_bits = 3 * st->logB;
    st->nbx = (s->dp.nxl + st->B - 1) / st->B;
    st->nby = (s->dp.Ry + st->B - 1) / st->B;
    st->nbz = (s->dp.Rz + st->B - 1) / st->B;
    st->nblocks = (int64_t)st->nbx * st->nby * st->nbz;
    if (ilog2_ceil64(st->nblocks) + st->cell_bits > 31) { s->err = "grid too large for 32-bit cell keys"; return MPM_ERR_INVALID; }
    // (the P2G kernels index 16-byte pieces of the records with 32 bits: record * 4 + piece)
    if (s->pitch >= ((int64_t)1 << 30)) { s->err = "MPM_PATH_CELL holds at most 2^30 particles per GPU"; return MPM_ERR_INVALID; }
    st->nslots = st->nblocks << st->cell_bits;
    for (int k = 0; k < 2; ++k) {
        CKB(cudaMalloc(&st->cnt[k], sizeof(uint32_t) * (st->nslots + 32)));
        CKB(cudaMemsetAsync(st->cnt[k], 0, sizeof(uint32_t) * (st->nslots + 32), s->stream));
    }
    const int64_t nvpos = 2 * st->nslots;  // virtual-cell positions: 2 per cell slot
    CKB(cudaMalloc(&st->cnts, sizeof(uint32_t) * (nvpos + 32)));
    CKB(cudaMemsetAsync(st->cnts, 0, sizeof(uint32_t) * (nvpos + 32), s->stream));
    CKB(cudaMalloc(&st->ord, sizeof(uint16_t) * nvpos));
    CKB(cudaMalloc(&st->cellmeta, sizeof(uint2) * st->nslots));
    CKB(cudaMalloc(&st->pstart, sizeof(uint32_t) * (nvpos >> 5)));
    CKB(cudaMalloc(&st->stab, sizeof(uint16_t) * STAB_ROW * (nvpos >> 5)));
    for (int k = 0; k < 2; ++k) {
        CKB(cudaMalloc(&st->bsum2[k], sizeof(uint32_t) * (st->nblocks + SCAN_PAD)));
        CKB(cudaMemsetAsync(st->bsum2[k], 0, sizeof(uint32_t) * (st->nblocks + SCAN_PAD), s->stream));
        CKB(cudaMalloc(&st->bbase2[k], sizeof(uint32_t) * (st->nblocks + SCAN_PAD)));
        CKB(cudaMemsetAsync(st->bbase2[k], 0, sizeof(uint32_t) * (st->nblocks + SCAN_PAD), s->stream));
        CKB(cudaMalloc(&st->active2[k], sizeof(uint32_t) * st->nblocks));
    }
    st->bsum = st->bsum2[0]; st->bbase = st->bbase2[0]; st->active = st->active2[0];
    CKB(cudaMalloc(&st->nact, sizeof(uint32_t) * 2));
    CKB(cudaMemsetAsync(st->nact, 0, sizeof(uint32_t) * 2, s->stream));
    // stable ranking is the default on one GPU; multi-GPU slabs (migration reshuffles the slots) and MPM_ATOMIC_BINNING=1
    // rank with an atomic cursor
    st->stable = getenv("MPM_ATOMIC_BINNING") == nullptr;
    const int64_t T = st->B + 2;
    CKB(cudaMalloc(&st->tcount, sizeof(uint32_t) * st->nblocks * T * T * T));
    CKB(cudaMalloc(&st->fixlist, sizeof(uint32_t) * FIX_CAP));
    CKB(cudaMalloc(&st->heavy, sizeof(uint32_t) * HEAVY_CAP));
    CKB(cudaMalloc(&st->far_n, sizeof(uint32_t) * 8));
    CKB(cudaMemsetAsync(st->far_n, 0, sizeof(uint32_t) * 8, s->stream));
    {
        const char* e = getenv("MPM_FAR_LIMIT");
        const uint32_t lim = e ? (uint32_t)strtoul(e, nullptr, 10) : FAR_LIMIT_DEFAULT;
        CKB(cudaMemcpyAsync(st->far_n + 6, &lim, sizeof(lim), cudaMemcpyHostToDevice, s->stream));
    }
    CKB(cudaMalloc(&st->fill, sizeof(uint32_t) * st->nslots));
    CKB(cudaMemsetAsync(st->fill, 0, sizeof(uint32_t) * st->nslots, s->stream));
    CKB(cudaMalloc(&st->farcnt, sizeof(uint32_t) * st->nslots));
    CKB(cudaMemsetAsync(st->farcnt, 0, sizeof(uint32_t) * st->nslots, s->stream));
    CKB(cudaMalloc(&st->keys, sizeof(uint32_t) * s->pitch));
    CKB(cudaMalloc(&st->src_of, sizeof(uint32_t) * (s->pitch + 512)));  // (the cell kernels read up to two units past the last slot)
    CKB(cudaMemsetAsync(st->src_of, 0, sizeof(uint32_t) * (s->pitch + 512), s->stream));

    CKB(cudaMalloc(&st->box, sizeof(int) * 12));
    CKB(cudaMemsetAsync(st->box, 0, sizeof(int) * 12, s->stream));  // empty boxes
    CKB(cudaMalloc(&st->misc, sizeof(uint32_t) * BIN_MISC_WORDS));
    CKB(cudaMemsetAsync(st->misc, 0, sizeof(uint32_t) * BIN_MISC_WORDS, s->stream));
    st->cur = 0;
    st->next_valid = false;
    return MPM_OK;
}

void bin_destroy(MpmSolver* s)
{
    BinState* st = s->bin;
    if (!st) return;
    cudaFree(st->cnt[0]); cudaFree(st->cnt[1]); cudaFree(st->cnts); cudaFree(st->ord); cudaFree(st->cellmeta); cudaFree(st->pstart); cudaFree(st->stab);
    for (int k = 0; k < 2; ++k) { cudaFree(st->bsum2[k]); cudaFree(st->bbase2[k]); cudaFree(st->active2[k]); }
    cudaFree(st->nact); cudaFree(st->tcount); cudaFree(st->fixlist); cudaFree(st->heavy); cudaFree(st->far_n);
    cudaFree(st->fill); cudaFree(st->farcnt); cudaFree(st->keys); cudaFree(st->src_of);
    cudaFree(st->misc); cudaFree(st->box);
    delete st;
    s->bin = nullptr;
}

KeyGeom bin_key_geom(const MpmSolver* s)
{
    const BinState* st = s->bin;
    return KeyGeom{s->dp.dim, st->logB, st->nby, st->nbz, s->dp.gx0 + (s->comm ? 1 : 0)};
}

static bool use_stable(const MpmSolver* s) { return s->bin->stable && !s->comm; }

// persistent grid of the ranking kernels (one CTA per tile, striding over the list of the previous layout's non-empty blocks)
static unsigned rank_grid(const BinState* st, int per_sm)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return (unsigned)std::min<int64_t>(st->nblocks, (int64_t)sms * per_sm);
}

// Cold binning: the particles sit in the planes in arbitrary (upload) order.  Stable radix sort of (cell key, index) on
// all key bits, then the records are written in that order (with keys, counts and original indices): afterwards the
// "previous layout" the stable ranking needs is simply the key order itself.
static int cold_sort(MpmSolver* s)
{
    BinState* st = s->bin;
    const int64_t n = s->n;
    const int nxt = st->cur ^ 1;
    const unsigned nb = (unsigned)((n + 255) / 256);
    if (s->in_rec) { int rc = ensure_planes(s); if (rc) return rc; }  // (a second binning without a G2P in between: back to the planes, in record order)
    CKB(cudaMemsetAsync(st->cnt[nxt], 0, sizeof(uint32_t) * st->nslots, s->stream));
    if (n > 0) {
        uint32_t* kb[2] = {nullptr, nullptr};
        uint32_t* vb[2] = {nullptr, nullptr};
        for (int k = 0; k < 2; ++k) { CKB(cudaMalloc(&kb[k], sizeof(uint32_t) * n)); CKB(cudaMalloc(&vb[k], sizeof(uint32_t) * n)); }
        k_cold_keys<<<nb, 256, 0, s->stream>>>(bin_key_geom(s), s->view(), n, (uint32_t)st->nslots, kb[0], vb[0]);
        const int key_bits = st->cell_bits + std::max(1, ilog2_ceil64(st->nblocks));
        const int fin = radix_sort_pairs(kb, vb, n, 0, key_bits, s->stream, &s->launches, &s->err);
        if (fin < 0) { for (int k = 0; k < 2; ++k) { cudaFree(kb[k]); cudaFree(vb[k]); } return MPM_ERR_CUDA; }
        k_cold_gather<<<nb, 256, 0, s->stream>>>(s->view(), kb[fin], vb[fin], n, reinterpret_cast<float4*>(s->rec), st->keys, st->cnt[nxt], s->orig_id, s->orig_id_alt);
        s->launches += 2;
        CKB(cudaStreamSynchronize(s->stream));
        for (int k = 0; k < 2; ++k) { cudaFree(kb[k]); cudaFree(vb[k]); }
        std::swap(s->orig_id, s->orig_id_alt);  // ids in record order
    }
    s->in_rec = true;
    st->lay_valid = false;  // the layout this binning computes doubles as the "previous" one
    return MPM_OK;
}

int bin_particles(MpmSolver* s)
{
    BinState* st = s->bin;
    const int64_t n = s->n + s->n_launch_extra;  // (launch size; see MpmSolver::n_launch_extra)
    const int nxt = st->cur ^ 1;
    const unsigned nb = (unsigned)((n + 255) / 256);
    const bool stable = use_stable(s);
    if (!st->next_valid && stable) {
        int rc = cold_sort(s);
        if (rc) return rc;
    } else {
        // On this path the particle state lives in the 64-byte records: G2P writes them, the P2G kernels read them through
        // src_of, and the binning never moves a particle.  A set that was just uploaded / edited is in the planes: convert once.
        if (!s->in_rec) {
            if (n > 0) { k_planes_to_rec<<<nb, 256, 0, s->stream>>>(s->view(), reinterpret_cast<float4*>(s->rec), n); s->launches += 1; }
            s->in_rec = true;
        }
        if (!st->next_valid) {  // no G2P has produced keys/counts for this particle set: compute them from the positions
            CKB(cudaMemsetAsync(st->cnt[nxt], 0, sizeof(uint32_t) * st->nslots, s->stream));
            if (n > 0) {
                k_bin_keys<RecView><<<nb, 256, 0, s->stream>>>(bin_key_geom(s), s->rview(), 0, n, (uint32_t)st->nslots, st->keys, st->cnt[nxt]);
                s->launches += 1;
            }
        }
    }
    const unsigned nbw = (unsigned)((st->nblocks * 32 + 255) / 256);
    const BoxGeom bg{st->nby, st->nbz, st->B, s->dp.gx0 + (s->comm ? 1 : 0), s->dp.gx0, s->dp.nxl, s->dp.Ry, s->dp.Rz};
    const int nl = st->lay ^ 1;            // buffers of the layout being built
    const int pl = st->lay_valid ? st->lay : nl;  // ... and of the one the records are in
    uint32_t* bsum = st->bsum2[nl];
    uint32_t* bbase = st->bbase2[nl];
    uint32_t* active = st->active2[nl];
    if (st->cell_bits == 9) {
        launch_pdl<PDL_BIN>(k_block_sums<9>, dim3(nbw), dim3(256), 0, s->stream, st->cnt[nxt], st->nblocks, bsum);
        launch_pdl<PDL_BIN>(k_scan_blocks, dim3(1), dim3(1024), 0, s->stream, bsum, st->nblocks, bbase, active, st->misc, bg, st->box, st->box_cleared ? 1 : 0, st->nact + nl, st->far_n);
        launch_pdl<PDL_BIN>(k_block_order<9>, dim3((unsigned)st->nblocks), dim3(512), 0, s->stream, st->cnt[nxt], bbase, active, st->misc, st->ord, st->cnts, st->cellmeta, st->pstart, st->stab, st->fill, st->farcnt);
    } else {
        launch_pdl<PDL_BIN>(k_block_sums<6>, dim3(nbw), dim3(256), 0, s->stream, st->cnt[nxt], st->nblocks, bsum);
        launch_pdl<PDL_BIN>(k_scan_blocks, dim3(1), dim3(1024), 0, s->stream, bsum, st->nblocks, bbase, active, st->misc, bg, st->box, st->box_cleared ? 1 : 0, st->nact + nl, st->far_n);
        launch_pdl<PDL_BIN>(k_block_order<6>, dim3((unsigned)st->nblocks), dim3(64), 0, s->stream, st->cnt[nxt], bbase, active, st->misc, st->ord, st->cnts, st->cellmeta, st->pstart, st->stab, st->fill, st->farcnt);
    }
    s->launches += 3;
    if (n > 0 && stable) {
        const RankGeom rg{st->nbx, st->nby, st->nbz};
        const unsigned grid_c = rank_grid(st, 12), grid_l = rank_grid(st, MPM_RANK_CTAS), grid_h = rank_grid(st, 4);  // (CTAs per SM)
        const RankArgs ra{st->keys, st->bsum2[pl], st->bbase2[pl], st->active2[pl], st->nact + pl, rg, st->tcount, st->cnt[nxt], st->cellmeta,
                          st->cnts, st->pstart, st->stab, st->fill, st->farcnt, st->far_n, st->heavy, (uint32_t)n, st->src_of, s->orig_id, s->orig_id_alt};
        const unsigned grid_f = rank_grid(st, 4);
        if (st->cell_bits == 9) {
            launch_pdl<PDL_RANK>(k_rank_count<9>, dim3(grid_c), dim3(RankCfg<9>::THREADS), 0, s->stream, st->keys, st->bbase2[pl], st->active2[pl], st->nact + pl, rg, st->tcount, st->farcnt, st->fixlist, st->far_n, st->heavy);
            launch_pdl<PDL_RANK>(k_rank_place<9>, dim3(grid_l), dim3(128), 0, s->stream, ra);
            launch_pdl<PDL_RANK>(k_rank_place_heavy<9>, dim3(grid_h), dim3(256), 0, s->stream, ra);
            launch_pdl<PDL_RANK>(k_fix_far<9>, dim3(grid_f), dim3(128), 0, s->stream, st->fixlist, st->far_n, st->cnt[nxt], st->farcnt, st->cellmeta, st->cnts, st->pstart, st->stab, st->src_of, s->orig_id_alt);
        } else {
            launch_pdl<PDL_RANK>(k_rank_count<6>, dim3(grid_c), dim3(RankCfg<6>::THREADS), 0, s->stream, st->keys, st->bbase2[pl], st->active2[pl], st->nact + pl, rg, st->tcount, st->farcnt, st->fixlist, st->far_n, st->heavy);
            launch_pdl<PDL_RANK>(k_rank_place<6>, dim3(grid_l), dim3(128), 0, s->stream, ra);
            launch_pdl<PDL_RANK>(k_rank_place_heavy<6>, dim3(grid_h), dim3(256), 0, s->stream, ra);
            launch_pdl<PDL_RANK>(k_fix_far<6>, dim3(grid_f), dim3(128), 0, s->stream, st->fixlist, st->far_n, st->cnt[nxt], st->farcnt, st->cellmeta, st->cnts, st->pstart, st->stab, st->src_of, s->orig_id_alt);
        }
        s->launches += 4;
    } else if (n > 0) {
        if (st->cell_bits == 9) launch_pdl<PDL_RANK>(k_place<9>, dim3(nb), dim3(256), 0, s->stream, st->keys, n, st->cellmeta, st->cnts, st->pstart, st->stab, st->fill, st->src_of, s->orig_id, s->orig_id_alt, s->n_dev);
        else launch_pdl<PDL_RANK>(k_place<6>, dim3(nb), dim3(256), 0, s->stream, st->keys, n, st->cellmeta, st->cnts, st->pstart, st->stab, st->fill, st->src_of, s->orig_id, s->orig_id_alt, s->n_dev);
        s->launches += 1;
    }
    st->prev_lay = pl;
    st->lay = nl;
    st->lay_valid = true;
    st->bsum = bsum; st->bbase = bbase; st->active = active;
    s->g2p_inputs = false;  // position / mass planes of this layout: written by P2G_1 (orig_id_alt holds the slot-order
                            // ids already; the two id arrays are swapped when G2P has rewritten the records in slot order)
    st->box_cleared = false;
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

// Binning introspection (mpm_debug_last_sort on the cell path).  Valid between a bin phase and the next G2P (keys[] still
// holds the keys that binning sorted by, in record order).  Re-runs the ranking kernels in verify mode: they write the
// rank inside the cell of every record and compare the layout in place (src_of, ids) with the one they derive; the
// permutation "cell-major position -> record index" is then assembled on the host from (key, rank).
int bin_debug_last(MpmSolver* s, uint32_t* keys_before, uint32_t* perm, int64_t cap)
{
    BinState* st = s->bin;
    if (!st || !s->sorted_valid || !st->lay_valid) { s->err = "no bin phase has run since the particles last moved"; return MPM_ERR_STATE; }
    if (!use_stable(s)) { s->err = "binning introspection needs the stable ranking (one GPU, MPM_ATOMIC_BINNING unset)"; return MPM_ERR_STATE; }
    const int64_t n = s->n;
    if (cap < n) { s->err = "destination too small"; return MPM_ERR_INVALID; }
    if (n == 0) return MPM_OK;
    // the definition: rank inside the cell = number of earlier records (lower old slot) with the same key
    std::vector<uint32_t> keys((size_t)n), rank((size_t)n);
    CKB(cudaMemcpyAsync(keys.data(), st->keys, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, s->stream));
    CKB(cudaStreamSynchronize(s->stream));
    std::vector<uint32_t> start((size_t)st->nslots + 1, 0u);
    for (int64_t i = 0; i < n; ++i) rank[(size_t)i] = start[keys[(size_t)i] + 1]++;
    // the layout in place must put record i at (cell keys[i], rank[i]), with its original index
    uint32_t *d_rank = nullptr, *d_bad = nullptr;
    uint32_t bad = 0;
    CKB(cudaMalloc(&d_rank, sizeof(uint32_t) * n));
    CKB(cudaMalloc(&d_bad, sizeof(uint32_t)));
    CKB(cudaMemsetAsync(d_bad, 0, sizeof(uint32_t), s->stream));
    CKB(cudaMemcpyAsync(d_rank, rank.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice, s->stream));
    const unsigned nb = (unsigned)((n + 255) / 256);
    if (st->cell_bits == 9) k_verify_layout<9><<<nb, 256, 0, s->stream>>>(st->keys, d_rank, n, st->cellmeta, st->cnts, st->pstart, st->stab, st->src_of, s->orig_id, s->orig_id_alt, d_bad);
    else k_verify_layout<6><<<nb, 256, 0, s->stream>>>(st->keys, d_rank, n, st->cellmeta, st->cnts, st->pstart, st->stab, st->src_of, s->orig_id, s->orig_id_alt, d_bad);
    CKB(cudaMemcpyAsync(&bad, d_bad, sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
    CKB(cudaStreamSynchronize(s->stream));
    cudaFree(d_rank); cudaFree(d_bad);
    if (bad) {
        s->err = "binning introspection: " + std::to_string(bad) + " records are not where the stable sort by cell key puts them";
        return MPM_ERR_STATE;
    }
    if (keys_before) std::copy(keys.begin(), keys.end(), keys_before);
    if (perm) {  // cell-major position -> record index, as the (verified) layout has it
        for (int64_t k = 0; k < st->nslots; ++k) start[(size_t)k + 1] += start[(size_t)k];
        for (int64_t i = 0; i < n; ++i) perm[(size_t)start[keys[(size_t)i]] + rank[(size_t)i]] = (uint32_t)i;
    }
    return MPM_OK;
}

// multi-GPU: keys of the particles that arrived by migration (slots [first, first + count)) join the keys and counts the
// last G2P produced for the ones that stayed
uint32_t* bin_next_keys(MpmSolver* s) { return (s->bin && s->bin->next_valid) ? s->bin->keys : nullptr; }
uint32_t* bin_next_counts(MpmSolver* s) { return s->bin->cnt[s->bin->cur ^ 1]; }
uint32_t bin_nslots(const MpmSolver* s) { return (uint32_t)s->bin->nslots; }

int bin_keys_range(MpmSolver* s, int64_t first, int64_t count)
{
    BinState* st = s->bin;
    if (!st || !st->next_valid || count <= 0) return MPM_OK;
    const unsigned nb = (unsigned)((count + 255) / 256);
    if (s->in_rec) k_bin_keys<RecView><<<nb, 256, 0, s->stream>>>(bin_key_geom(s), s->rview(), first, count, (uint32_t)st->nslots, st->keys, st->cnt[st->cur ^ 1]);
    else k_bin_keys<ParticleView><<<nb, 256, 0, s->stream>>>(bin_key_geom(s), s->view(), first, count, (uint32_t)st->nslots, st->keys, st->cnt[st->cur ^ 1]);
    s->launches += 1;
    return MPM_OK;
}

// G2P without a P2G_1 since the last binning (phases run one by one)
int bin_g2p_inputs(MpmSolver* s)
{
    if (s->g2p_inputs) return MPM_OK;
    if (s->n > 0) {
        k_gather_g2p_inputs<<<(unsigned)((s->n + 255) / 256), 256, 0, s->stream>>>(reinterpret_cast<const float4*>(s->rec), s->view(), s->bin->src_of, s->n);
        s->launches += 1;
    }
    s->g2p_inputs = true;
    return MPM_OK;
}

int ensure_planes(MpmSolver* s)
{
    if (!s->in_rec) return MPM_OK;
    launch_rec_to_planes(s->rview(), s->view(), s->n, s->stream);
    s->launches += (s->n > 0);
    s->in_rec = false;
    s->sorted_valid = false;
    return MPM_OK;
}

}  // namespace mpm
