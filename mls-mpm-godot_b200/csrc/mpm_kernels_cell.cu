// mpm_kernels_cell.cu -- the cell kernels (MPM_PATH_CELL, MPM_MATH_FAST): 3D, int32 fixed-point grid.
//
// ncu on B200 (profiles/r1/v3_*) showed the one-thread-per-particle tiled kernels bound by per-particle
// instruction streams that no amount of bandwidth shrinks: P2G_1 issues 108 shared-memory atomics and 108
// float->int conversions per particle (the conversion runs on the quarter-rate XU pipe, 61-69 % busy), G2P 81
// shared-memory loads per particle.  Here one thread owns one grid CELL and walks that cell's particles
// (cell-binned slots, mpm_bin.cu), keeping the cell's 27-node stencil in registers:
//   P2G_1 / P2G_2 : 108 / 81 fp32 accumulators per thread; fixed-point conversion and the shared-memory ATOMS
//                   happen once per cell instead of once per particle.
//   G2P           : the 27 node velocities (81 floats) are loaded from the shared-memory tile once per cell; results go
//                   out as 64-byte records (two full sectors), which the next step's P2G kernels read in place.
// A warp = 32 cells of a block (a "chunk"); at rank r it works on the rank-r particles of its cells, which sit in
// consecutive slots (a "row").  P2G fetches a row's records through the binning's index (RowStage below), G2P reads
// position + mass planes that P2G_1 wrote in slot order.  CTAs are persistent and
// pull non-empty grid blocks from a list with an atomic counter.  G2P also emits each particle's next cell key
// and counts it (RED), so the next step's binning needs no separate key pass.
//
// Arithmetic: MPM_MATH_FAST (FMA, per-axis hoisting, sum factorisation).  Contributions are accumulated in fp32
// and truncated to the int32 grid once per (cell, node) instead of once per (particle, node); the difference is
// below one fixed-point unit per particle and inside the FAST tolerance stated in tests/test_parity_gpu.py.
#include <cuda.h>  // CUtensorMap (the encoder itself is resolved through cudaGetDriverEntryPoint: no link to libcuda)

#include "mpm_bin.h"
#include "mpm_kernels.h"
#include "mpm_particle_math.cuh"
#include "mpm_tile.cuh"

// resident CTAs per SM the P2G kernels are compiled for (3 -> 168 registers, 4 -> 128 registers with spills)
#ifndef MPM_P2G_CTAS
#define MPM_P2G_CTAS 3
#endif
#ifndef MPM_G2P_CTAS
#define MPM_G2P_CTAS 3
#endif
// rows (of up to 32 particles) per pipeline unit: what one request brings in and one loop iteration of the walk consumes
#ifndef MPM_UNIT_ROWS
#define MPM_UNIT_ROWS 2
#endif

namespace mpm {

KeyGeom bin_key_geom(const MpmSolver* s);

// Tile of (B+2)^3 nodes for the cell kernels: DENSE, node (tx, ty, tz) at (tx * T + ty) * T + tz -- the order a bulk tensor
// copy of the box delivers and the order of a flat loop over the nodes.  (The tiled kernels pad their tile so that 32
// consecutive cells hit 32 banks; here a warp's lanes are the cells of a chunk in order of particle count, i.e. in no
// spatial order, so the padding bought nothing and cost 2.5x the shared memory: 40 KB against 16 KB for the four channels.)
template <int B>
struct CTile {
    static constexpr int T = B + 2;
    static constexpr int PY = T;
    static constexpr int PX = T * T;
    static constexpr int NODES = T * T * T;
    static constexpr int WORDS = (NODES + 3) / 4 * 4;  // per channel
    int ox, oy, oz;
    __device__ __forceinline__ void init(const TileGeom& g, int b)
    {
        const int bz = b % g.nbz, by = (b / g.nbz) % g.nby, bx = b / (g.nbz * g.nby);
        ox = g.x_owned0 + bx * B - 1; oy = by * B - 1; oz = bz * B - 1;
    }
    // k-th node of the tile -> its index (k itself) and global cell index (false if outside the local grid)
    __device__ __forceinline__ bool node(const DevParams& P, int k, int& idx, int64_t& ci) const
    {
        const int tz = k % T, ty = (k / T) % T, tx = k / (T * T);
        idx = k;
        const int nx = ox + tx, ny = oy + ty, nz = oz + tz;
        if (nx < P.gx0 || nx >= P.gx0 + P.nxl || ny < 0 || ny >= P.Ry || nz < 0 || nz >= P.Rz) return false;
        ci = cell_index(P, nx, ny, nz);
        return true;
    }
};

template <int B>
struct CellCfg {
    static constexpr int LOGB = (B == 8) ? 3 : 2;
    static constexpr int NC = B * B * B;
    static constexpr int NV = 2 * NC;        // virtual-cell positions per block (a cell with > 32 particles is split)
    static constexpr int NCHUNK = NV / 32;   // chunks of 32 virtual cells
    static constexpr int THREADS = (B == 8) ? 128 : 64;
    static constexpr int NWARP = THREADS / 32;
};

struct CellArgs {
    const uint32_t* cnts;    // particle count at each virtual-cell position (descending inside a block)
    const uint16_t* ord;     // virtual-cell position -> cell id inside the block
    const uint32_t* pstart;  // first slot of every chunk
    const uint32_t* active;  // non-empty blocks
    uint32_t* misc;          // list length, work counters
};

struct BlockWork {  // shared-memory hand-off of fetch_block
    int b;          // grid block, or -1
    int next_chunk; // chunks are handed to warps dynamically (heaviest first): next unclaimed chunk
};

// next non-empty grid block for this CTA, or -1
template <int NWARP>
__device__ __forceinline__ int fetch_block(const CellArgs& a, int which, BlockWork* w)
{
    __syncthreads();  // everyone is done with the previous block's shared memory
    if (threadIdx.x == 0) {
        const uint32_t bi = atomicAdd(&a.misc[which], 1u);
        w->b = (bi < a.misc[BIN_N_ACTIVE]) ? (int)a.active[bi] : -1;
        w->next_chunk = NWARP;  // chunk `warp` is every warp's first one
    }
    __syncthreads();
    return w->b;
}

// weights and node distances of one axis for a particle known to sit in cell `fc` (as float)
__device__ __forceinline__ void cell_axis(float p, float fc, float w[3], float d[3])
{
    const float cd = (p - fc) - 0.5f;
    const float a = 0.5f - cd, b = 0.5f + cd;
    w[0] = 0.5f * a * a;
    w[1] = 0.75f - cd * cd;
    w[2] = 0.5f * b * b;
    d[0] = -1.0f - cd; d[1] = -cd; d[2] = 1.0f - cd;
}

// The same for two axes at once (x and y as an aligned pair): every operation is one packed instruction.
__device__ __forceinline__ void cell_axis2(float2 p, float2 fc, float2 w[3], float2 d[3])
{
    const float2 h = make_float2(0.5f, 0.5f), one = make_float2(1.0f, 1.0f);
    const float2 cd = __fadd2_rn(__fadd2_rn(p, make_float2(-fc.x, -fc.y)), make_float2(-0.5f, -0.5f));
    const float2 ncd = make_float2(-cd.x, -cd.y);
    const float2 a = __fadd2_rn(h, ncd), b = __fadd2_rn(h, cd);
    w[0] = __fmul2_rn(__fmul2_rn(h, a), a);
    w[1] = __ffma2_rn(ncd, cd, make_float2(0.75f, 0.75f));
    w[2] = __fmul2_rn(__fmul2_rn(h, b), b);
    d[0] = __fadd2_rn(ncd, make_float2(-1.0f, -1.0f)); d[1] = ncd; d[2] = __fadd2_rn(ncd, one);
}
// a scalar as both halves of a pair: ptxas folds this into the packed instruction's scalar-operand form (`R.F32`), no move
__device__ __forceinline__ float2 bc(float s) { return make_float2(s, s); }

// ---------------------------------------------------------------- TMA: a block's grid tile as one bulk tensor copy
// The (B+2)^3 nodes a block's stencils touch are a dense box of the grid array [x][y][z][4 x int32].  One elected thread
// asks the TMA unit for the box (cp.async.bulk.tensor.4d, SASS UTMALDG): the 16 KB land in shared memory as [x][y][z][4]
// and the copy signals an mbarrier with its byte count.  Nodes outside the (local) grid come back as zeros, which is what
// the hand-written tile loaders wrote for them.  G2P issues the request for its NEXT block before walking the current one,
// so the tile load -- every thread of the CTA used to stall on it behind one barrier, 13 % of the kernel's stall samples
// -- overlaps a whole block of work.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// box of the 4-D view (channel, z, y, x) of the grid at (0, cz, cy, cx) -> dst; completion on bar
__device__ __forceinline__ void tma_load_tile(void* dst, const CUtensorMap* map, int cz, int cy, int cx, unsigned long long* bar)
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // (the buffer was last read through the generic proxy)
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(0), "r"(cz), "r"(cy), "r"(cx), "r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------- walking a block's chunks, software-pipelined
// ncu (first cell-kernel capture, summarised in DESIGN.md section 4): with 168 registers per thread only 12 warps fit on
// an SM, and a loop that loads a particle and then computes on it spends half its time in long-scoreboard stalls.  So the
// walk is pipelined: while the particle of rank r is being processed, the one of rank r+1 (or rank 0 of the warp's next
// chunk) is already in flight -- in registers where that is cheap (G2P: position + mass), as a cp.async into a per-warp
// shared-memory staging buffer otherwise (P2G: the whole 64-byte record; prefetch.global.L1 was tried first and left the
// L1 hit rate at 5 %).
//
// Round 2: the unit of the pipeline is TWO rows (up to 64 consecutive slots).  With one row per request the kernels were
// latency-bound, not issue-bound -- 12 warps per SM x <= 2 rows x 2 KB in flight is ~5 MB over the chip, which at the ~1.2 us
// a loaded DRAM access takes caps the read rate near 4 TB/s (measured: 3.26 GB in 0.82 ms) -- and cutting the instruction
// count by a fifth bought only 9 %.  A unit doubles the bytes in flight with the same two buffers and the same
// bookkeeping, and halves the per-row share of the walk's control instructions.
//
// Body interface (all calls warp-uniform except compute):
//                  begin_chunk(cell)      per-cell set-up (stencil registers / accumulators)
//                  fetch(base, n, k)      start bringing a unit in: slots base .. base + n - 1 (n <= 64).  k says how it was
//                                         announced: ROW_NEXT = it follows the unit fetched last, ROW_HINTED = it starts at the
//                                         slot given to hint_chunk, ROW_COLD = neither
//                  hint_chunk(slot)       the warp's next chunk will start at `slot`
//                  take(pending)          the oldest unit not yet taken becomes the current one; `pending` (0 or 1) units were
//                                         requested after it and may stay in flight
//                  compute(i, t)          process this lane's particle t of the current unit (slot i = unit base + t)
//                  end_chunk(has)         flush per-cell results
//                  finish()               after the warp's last chunk of the block
enum { ROW_COLD = 0, ROW_NEXT = 1, ROW_HINTED = 2 };

// asynchronous 16-byte global -> shared copy (LDGSTS: no register is held while the load is in flight), L2 only (the
// records are read once per kernel)
__device__ __forceinline__ void cp_async16(float* smem, const float* gmem)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem));
}
// the same, predicated (no branch around it: a row is copied by straight-line code)
__device__ __forceinline__ void cp_async16_if(unsigned sa, const void* gmem, bool on)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16;\n\t}" ::"r"(sa), "l"(gmem), "r"((int)on));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// Pile-ups (cells with thousands of particles against a wall or in a corner of the evolved dam-break) would leave one
// lane walking for ages while 31 idle; the binning therefore splits any cell with more than 32 particles into several
// VIRTUAL cells of at most 32 (mpm_bin.cu), each taken by its own lane.  The walk below only ever sees virtual cells:
// a block has up to 2 * NC of them (NCHUNK chunks), ordered by count, descending.
constexpr int RU = MPM_UNIT_ROWS;

template <int B, class Body>
__device__ __forceinline__ void walk_chunks(const CellArgs& a, int b, int lane, int warp, BlockWork* bw, Body& body)
{
    using CF = CellCfg<B>;
    const unsigned lt = (1u << lane) - 1u;
    const uint32_t v0 = (uint32_t)b * CF::NV;  // first virtual-cell position of the block
    bool have = false;  // the first unit of the coming chunk has already been requested
    int chunk_nx = warp;
    uint32_t c_nx = a.cnts[v0 + warp * 32 + lane];
    uint32_t L_nx = a.ord[v0 + warp * 32 + lane];
    uint32_t st_nx = a.pstart[(v0 >> 5) + warp];
    // slots of the first unit (rows 0 .. RU - 1) of a chunk whose lanes hold `cnt` particles
    auto first_unit = [&](uint32_t cnt) {
        uint32_t n = 0;
#pragma unroll
        for (int u = 0; u < RU; ++u) n += __popc(__ballot_sync(0xffffffffu, cnt > (uint32_t)u));
        return n;
    };
#pragma unroll 1
    for (;;) {
        const int chunk = chunk_nx;
        if (chunk >= CF::NCHUNK) break;
        const uint32_t c = c_nx, L = L_nx;
        uint32_t slot0 = st_nx;
        unsigned m[RU];  // lanes that have a particle in row r + u
#pragma unroll
        for (int u = 0; u < RU; ++u) m[u] = __ballot_sync(0xffffffffu, c > (uint32_t)u);
        if (!m[0]) break;  // virtual cells are ordered by count, descending: every later chunk of the block is empty too
        // claim the chunk after this one now, so that its metadata (and first unit) arrive while this one runs
        int g = 0;
        if (lane == 0) g = atomicAdd(&bw->next_chunk, 1);
        chunk_nx = __shfl_sync(0xffffffffu, g, 0);
        c_nx = 0; L_nx = 0; st_nx = 0;
        if (chunk_nx < CF::NCHUNK) {
            c_nx = a.cnts[v0 + chunk_nx * 32 + lane];
            L_nx = a.ord[v0 + chunk_nx * 32 + lane];
            st_nx = a.pstart[(v0 >> 5) + chunk_nx];
        }
        if (!have) {
            uint32_t n = 0;
#pragma unroll
            for (int u = 0; u < RU; ++u) n += __popc(m[u]);
            body.fetch(slot0, n, ROW_COLD);
        }
        have = false;
        body.begin_chunk((int)L);
#pragma unroll 1
        for (uint32_t r = 0;; r += RU) {
            // rows r .. r + RU - 1 are the current unit: its slots are slot0 .. slot0 + n_unit - 1, row after row
            uint32_t t[RU], n_unit = 0;
#pragma unroll
            for (int u = 0; u < RU; ++u) { t[u] = n_unit + __popc(m[u] & lt); n_unit += __popc(m[u]); }
            unsigned mn[RU];
            uint32_t n_next = 0;
#pragma unroll
            for (int u = 0; u < RU; ++u) { mn[u] = __ballot_sync(0xffffffffu, r + RU + u < c); n_next += __popc(mn[u]); }
            // the next unit is requested BEFORE the current one is waited for: two units in flight while the warp waits
            bool fetched = true;
            if (mn[0]) {
                body.fetch(slot0 + n_unit, n_next, ROW_NEXT);
            } else {
                const uint32_t n0 = first_unit(c_nx);
                if (n0) body.fetch(st_nx, n0, r > 0 ? ROW_HINTED : ROW_COLD);  // (single-unit chunk: no hint was given yet)
                else fetched = false;
                have = true;
            }
            body.take(fetched ? 1 : 0);
#pragma unroll 1
            for (uint32_t u = 0; u < (uint32_t)RU; ++u) {  // (one copy of the body: inlined copies get interleaved and spill)
                uint32_t tu = t[0];
#pragma unroll
                for (int k = 1; k < RU; ++k) tu = (u == (uint32_t)k) ? t[k] : tu;
                if (r + u < c) body.compute(slot0 + tu, tu);
            }
            if (r == 0) body.hint_chunk(st_nx);  // (here, not where the chunk is claimed: st_nx has arrived by now)
            slot0 += n_unit;
#pragma unroll
            for (int u = 0; u < RU; ++u) m[u] = mn[u];
            if (!m[0]) break;
        }
        body.end_chunk(c > 0);
    }
    body.finish();
}

template <int B>
struct CellPos {  // the cell (id L inside the block) this lane owns in the current chunk
    int base;
    float2 fcxy;  // (x, y) of the cell as an aligned pair: the x and y weights are computed packed
    float fcz;
    __device__ __forceinline__ void set(const CTile<B>& tl, int L)
    {
        constexpr int LOGB = CellCfg<B>::LOGB;
        const int lx = L >> (2 * LOGB), ly = (L >> LOGB) & (B - 1), lz = L & (B - 1);
        base = lx * CTile<B>::PX + ly * CTile<B>::PY + lz;
        fcxy = make_float2((float)(tl.ox + 1 + lx), (float)(tl.oy + 1 + ly)); fcz = (float)(tl.oz + 1 + lz);
    }
};

// ---------------------------------------------------------------- rows of records for the P2G kernels
// The particle state between steps is the array of 64-byte records G2P wrote (slot order of the previous step); the
// binning only produces src_of[slot] = record index.  The P2G kernels read the records THROUGH that index -- two full
// sectors per particle whatever the permutation -- which saves the gather pass of the binning (136 bytes per particle,
// 0.66 ms of 3.58 ms on C4).  A row (<= 32 particles in consecutive slots) is brought in by the whole warp: lane 4q + j
// copies 16-byte piece j of the row's records q, q + 8, q + 16, q + 24 (one cp.async.cg each: 4 per lane and row, where
// the plane layout needed 16), so the four lanes of a quad cover one record = one 64-byte request.  Piece j of record t
// lands at chunk (j ^ (t >> 1)) & 3 of the record's 64 bytes: the owner's four LDS.128 are then conflict-free (8
// consecutive records cover all 8 bank quads).  The record indices of a row come from one coalesced load that is issued
// a row ahead (ix_row) or, for the first row of the warp's next chunk, when that chunk is announced (ix_chunk), and are
// passed around with SHFL; only the first row after a block change pays the load's latency.
struct RowStage {
    static constexpr int UNIT = 32 * RU;     // records per buffer
    static constexpr int WORDS = UNIT * 16;  // one buffer, in 4-byte words
    const float4* rec4;
    const uint32_t* src_of;
    float* buf;   // the warp's two buffers
    unsigned sa;  // shared-memory address of the 16 bytes this lane fills for record q = lane / 4 of buffer 0
    int lane;
    uint32_t ix_row[RU], ix_chunk[RU];  // record indices of the UNIT slots that follow the last request / start the next chunk
    int wr = 0, rd = 0;
    __device__ __forceinline__ RowStage(const float* rec_, const uint32_t* src_of_, float* buf_, int lane_)
        : rec4(reinterpret_cast<const float4*>(rec_)), src_of(src_of_), buf(buf_), lane(lane_)
    {
        const int q = lane >> 2, j = lane & 3;
        sa = (unsigned)__cvta_generic_to_shared(buf + q * 16 + ((j ^ (q >> 1)) & 3) * 4);
#pragma unroll
        for (int u = 0; u < RU; ++u) { ix_row[u] = 0; ix_chunk[u] = 0; }
    }
    __device__ __forceinline__ void hint_chunk(uint32_t slot)
    {
#pragma unroll
        for (int u = 0; u < RU; ++u) ix_chunk[u] = src_of[slot + 32 * u + lane];
    }
    __device__ __forceinline__ void fetch(uint32_t base, uint32_t cnt, int kind)
    {
        uint32_t ix[RU];
#pragma unroll
        for (int u = 0; u < RU; ++u) ix[u] = (kind == ROW_NEXT) ? ix_row[u] : (kind == ROW_HINTED) ? ix_chunk[u] : src_of[base + 32 * u + lane];
        const unsigned dst = sa + wr * (WORDS * 4);
        const uint32_t q = lane >> 2, j = lane & 3;
#pragma unroll
        for (int it = 0; it < 4 * RU; ++it) {  // record t = 8 it + q; its swizzle (t >> 1) & 3 does not depend on it
            const uint32_t src = __shfl_sync(0xffffffffu, ix[it / 4], (8 * it + q) & 31);
            cp_async16_if(dst + it * 512, rec4 + (src * 4u + j), 8 * it + q < cnt);
        }
        cp_async_commit();
        wr ^= 1;
#pragma unroll
        for (int u = 0; u < RU; ++u) ix_row[u] = src_of[base + cnt + 32 * u + lane];  // (src_of is padded: reading past the last particle is harmless)
    }
    __device__ __forceinline__ void take(int pending)
    {
        if (pending) cp_async_wait_but_one(); else cp_async_wait_all();
        __syncwarp();  // the pieces of a record were copied by four different lanes
        rd = pending ? wr : wr ^ 1;  // (wr = where the next request goes = the older of two outstanding units)
    }
    // record t of the current unit: (px, py, pz, m) (vx, vy, vz, c2) (c0, c1, c3, c4) (c6, c7, c5, c8)
    __device__ __forceinline__ void load(uint32_t t, float4& a, float4& b, float4& c, float4& d) const
    {
        const float4* r = reinterpret_cast<const float4*>(buf + rd * WORDS) + 4 * t;
        const uint32_t s = (t >> 1) & 3u;
        a = r[s]; b = r[1u ^ s]; c = r[2u ^ s]; d = r[3u ^ s];
    }
    // the same without the velocity quad (P2G_2): returns c2, the one field of that quad it needs
    __device__ __forceinline__ float load_no_vel(uint32_t t, float4& a, float4& c, float4& d) const
    {
        const float4* r = reinterpret_cast<const float4*>(buf + rd * WORDS) + 4 * t;
        const uint32_t s = (t >> 1) & 3u;
        a = r[s]; c = r[2u ^ s]; d = r[3u ^ s];
        return reinterpret_cast<const float*>(r + (1u ^ s))[3];
    }
};

// units of (px, py, pz, m) quads for G2P: P2G_1 wrote them in slot order, so a unit is one contiguous run of 16-byte
// elements; lane l copies elements l, l + 32, ...
struct QuadStage {
    static constexpr int UNIT = 32 * RU;
    const float4* pm;
    float4* buf;  // the warp's two buffers of UNIT quads
    unsigned sa;
    int lane;
    int wr = 0, rd = 0;
    __device__ __forceinline__ QuadStage(const float4* pm_, float4* buf_, int lane_) : pm(pm_), buf(buf_), lane(lane_)
    {
        sa = (unsigned)__cvta_generic_to_shared(buf + lane);
    }
    __device__ __forceinline__ void fetch(uint32_t base, uint32_t cnt, int)
    {
        const unsigned dst = sa + wr * (UNIT * 16);
#pragma unroll
        for (int u = 0; u < RU; ++u) cp_async16_if(dst + 512 * u, pm + base + 32 * u + lane, (uint32_t)(lane + 32 * u) < cnt);
        cp_async_commit();
        wr ^= 1;
    }
    __device__ __forceinline__ void take(int pending)
    {
        if (pending) cp_async_wait_but_one(); else cp_async_wait_all();
        __syncwarp();
        rd = pending ? wr : wr ^ 1;
    }
    __device__ __forceinline__ float4 load(uint32_t t) const { return buf[rd * UNIT + t]; }
};

// ---------------------------------------------------------------- P2G_1
template <int B>
struct P2G1Body {
    using TL = CTile<B>;
    const DevParams& P;
    const ParticleView& pv;  // out: position and mass planes in slot order (what G2P reads)
    const TL& tl;
    int (*tile)[TL::WORDS];
    RowStage st;
    CellPos<B> cp;
    // 108 accumulators, packed for FFMA2 without register moves: (momentum x, momentum y) per node; momentum z and mass
    // per (gx, gy) as a pair over the nodes gz = 0, 1 plus a scalar for gz = 2.  [first version: (z, mass) per node, which
    // needs the pair (q_z, 1.0) built with a MOV for every node, and the record's old field order split (c0, c1), (c6, c7)
    // and (vx, vy) across odd register offsets: 92 of 479 warp-instructions per row were register moves]
    float2 axy[27];
    float2 az01[9], am01[9];
    float az2[9], am2[9];
    __device__ __forceinline__ P2G1Body(const DevParams& P_, const ParticleView& pv_, const TL& tl_, int (*tile_)[TL::WORDS], const RowStage& st_)
        : P(P_), pv(pv_), tl(tl_), tile(tile_), st(st_) {}
    __device__ __forceinline__ void begin_chunk(int L)
    {
        cp.set(tl, L);
#pragma unroll
        for (int n = 0; n < 27; ++n) axy[n] = make_float2(0.0f, 0.0f);
#pragma unroll
        for (int g = 0; g < 9; ++g) { az01[g] = make_float2(0.0f, 0.0f); am01[g] = make_float2(0.0f, 0.0f); az2[g] = 0.0f; am2[g] = 0.0f; }
    }
    __device__ __forceinline__ void fetch(uint32_t base, uint32_t cnt, int kind) { st.fetch(base, cnt, kind); }
    __device__ __forceinline__ void hint_chunk(uint32_t slot) { st.hint_chunk(slot); }
    __device__ __forceinline__ void take(int pending) { st.take(pending); }
    __device__ __forceinline__ void compute(uint32_t i, uint32_t t)
    {
        float4 ra, rb, rc, rd4;
        st.load(t, ra, rb, rc, rd4);  // (px, py, pz, m) (vx, vy, vz, c2) (c0, c1, c3, c4) (c6, c7, c5, c8)
        // G2P needs the position and the mass of slot i: one 16-byte element, in slot order
        reinterpret_cast<float4*>(pv.base)[i] = ra;
        const float ms = ra.w * P.fmult;  // mass in fixed-point units
        float2 wxy[3], dxy[3];            // (x, y) weights and node distances, computed packed
        float wz[3], dz[3];
        cell_axis2(make_float2(ra.x, ra.y), cp.fcxy, wxy, dxy);
        cell_axis(ra.z, cp.fcz, wz, dz);
        const float2 vxy = make_float2(rb.x, rb.y), c01 = make_float2(rc.x, rc.y), c34 = make_float2(rc.z, rc.w), c67 = make_float2(rd4.x, rd4.y);
        const float vz = rb.z, c2 = rb.w, c5 = rd4.z, c8 = rd4.w;
        const float2 wz01 = make_float2(wz[0], wz[1]), dz01 = make_float2(dz[0], dz[1]);
        // node value = mc * (v + C d) with mc = w * m.  (x, y) run as one pair per node; z and the mass as one pair per
        // two nodes (gz = 0, 1) and a scalar (gz = 2).  Scalars enter the packed instructions through their scalar-operand
        // form.
#pragma unroll
        for (int gx = 0; gx < 3; ++gx) {
            const float wxm = wxy[gx].x * ms;
            const float2 q0 = __ffma2_rn(c01, bc(dxy[gx].x), vxy);
            const float qz0 = fmaf(c2, dxy[gx].x, vz);
#pragma unroll
            for (int gy = 0; gy < 3; ++gy) {
                const int g = gx * 3 + gy;
                const float2 q1 = __ffma2_rn(c34, bc(dxy[gy].y), q0);
                const float qz1 = fmaf(c5, dxy[gy].y, qz0);
                const float wgt = wxm * wxy[gy].y;
                const float2 mc01 = __fmul2_rn(bc(wgt), wz01);
                const float mc2 = wgt * wz[2];
                az01[g] = __ffma2_rn(mc01, __ffma2_rn(bc(c8), dz01, bc(qz1)), az01[g]);
                az2[g] = fmaf(mc2, fmaf(c8, dz[2], qz1), az2[g]);
                am01[g] = __fadd2_rn(am01[g], mc01);
                am2[g] += mc2;
                axy[3 * g + 0] = __ffma2_rn(bc(mc01.x), __ffma2_rn(c67, bc(dz[0]), q1), axy[3 * g + 0]);
                axy[3 * g + 1] = __ffma2_rn(bc(mc01.y), __ffma2_rn(c67, bc(dz[1]), q1), axy[3 * g + 1]);
                axy[3 * g + 2] = __ffma2_rn(bc(mc2), __ffma2_rn(c67, bc(dz[2]), q1), axy[3 * g + 2]);
            }
        }
    }
    __device__ __forceinline__ void end_chunk(bool has)
    {
        if (!has) return;
#pragma unroll
        for (int gx = 0; gx < 3; ++gx)
#pragma unroll
            for (int gy = 0; gy < 3; ++gy)
#pragma unroll
                for (int gz = 0; gz < 3; ++gz) {
                    const int g = gx * 3 + gy, n = g * 3 + gz;
                    const int idx = cp.base + gx * TL::PX + gy * TL::PY + gz;
                    const float mz = gz == 0 ? az01[g].x : gz == 1 ? az01[g].y : az2[g];
                    const float mm = gz == 0 ? am01[g].x : gz == 1 ? am01[g].y : am2[g];
                    atomicAdd(&tile[3][idx], __float2int_rz(mm));
                    atomicAdd(&tile[0][idx], __float2int_rz(axy[n].x));
                    atomicAdd(&tile[1][idx], __float2int_rz(axy[n].y));
                    atomicAdd(&tile[2][idx], __float2int_rz(mz));
                }
    }
    __device__ __forceinline__ void finish() {}
};

template <int B>
__global__ void __launch_bounds__(CellCfg<B>::THREADS, (B == 8) ? MPM_P2G_CTAS : 6) k_p2g1_cell(DevParams P, TileGeom g, ParticleView pv, CellArgs a,
                                                                                     int* __restrict__ grid, const float* __restrict__ rec,
                                                                                     const uint32_t* __restrict__ src_of)
{
    pdl_prologue();
    using TL = CTile<B>;
    using CF = CellCfg<B>;
    extern __shared__ __align__(16) unsigned char dsm[];  // tile | staging (above the 48 KB static limit together)
    int (*tile)[TL::WORDS] = reinterpret_cast<int (*)[TL::WORDS]>(dsm);
    float* stg = reinterpret_cast<float*>(dsm + sizeof(int) * 4 * TL::WORDS);
    __shared__ BlockWork s_bw;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (;;) {
        const int b = fetch_block<CF::NWARP>(a, BIN_WORK_P2G1, &s_bw);
        if (b < 0) break;
        TL tl; tl.init(g, b);
        for (int k = threadIdx.x; k < 4 * TL::WORDS; k += CF::THREADS) reinterpret_cast<int*>(dsm)[k] = 0;
        __syncthreads();
        P2G1Body<B> body(P, pv, tl, tile, RowStage(rec, src_of, stg + warp * (2 * RowStage::WORDS), lane));
        walk_chunks<B>(a, b, lane, warp, &s_bw, body);
        __syncthreads();
        for (int k = threadIdx.x; k < TL::NODES; k += CF::THREADS) {
            int idx; int64_t ci;
            const bool ok = tl.node(P, k, idx, ci);
            const int vx = tile[0][idx], vy = tile[1][idx], vz = tile[2][idx], m = tile[3][idx];
            if (!ok || (vx | vy | vz | m) == 0) continue;
            int* cc = grid + 4 * ci;
            if (vx) atomicAdd(cc + 0, vx);
            if (vy) atomicAdd(cc + 1, vy);
            if (vz) atomicAdd(cc + 2, vz);
            if (m) atomicAdd(cc + 3, m);
        }
    }
}

// ---------------------------------------------------------------- P2G_2
__device__ __forceinline__ float cell_eos_pow(float x, const DevParams& P)
{
    if (P.eos_pi == 7) {  // the GPU variant's exponent (H:84): 4 multiplications, no loop
        const float x2 = x * x, x3 = x2 * x;
        return (x3 * x3) * x;
    }
    if (P.eos_pi == 4) {  // the CPU variants' exponent (F:40)
        const float x2 = x * x;
        return x2 * x2;
    }
    if (P.eos_pi > 0) {
        float r = x;
        for (int k = 1; k < P.eos_pi; ++k) r *= x;
        return r;
    }
    return __powf(x, P.eos_p);
}

template <int B>
struct P2G2Body {
    using TL = CTile<B>;
    const DevParams& P;
    const TL& tl;
    int (*tile)[TL::WORDS];
    const float* tmass;
    float inv_rest;
    RowStage st;
    CellPos<B> cp;
    // node masses of the cell's stencil: per (gx, gz) the pair over gy = 0, 1 and a scalar for gy = 2 (the density sum
    // then runs packed); 81 accumulators: (x, y) per node, z per (gx, gy) as a pair over gz = 0, 1 and a scalar for gz = 2
    float2 gm01[9];
    float gm2[9];
    float2 axy[27];
    float2 az01[9];
    float az2[9];
    __device__ __forceinline__ P2G2Body(const DevParams& P_, const TL& tl_, int (*tile_)[TL::WORDS], const float* tmass_, const RowStage& st_)
        : P(P_), tl(tl_), tile(tile_), tmass(tmass_), inv_rest(1.0f / P_.rest_density), st(st_) {}
    __device__ __forceinline__ void begin_chunk(int L)
    {
        cp.set(tl, L);
#pragma unroll
        for (int gx = 0; gx < 3; ++gx)
#pragma unroll
            for (int gz = 0; gz < 3; ++gz) {
                const float* t = tmass + cp.base + gx * TL::PX + gz;
                gm01[gx * 3 + gz] = make_float2(t[0], t[TL::PY]);
                gm2[gx * 3 + gz] = t[2 * TL::PY];
            }
#pragma unroll
        for (int n = 0; n < 27; ++n) axy[n] = make_float2(0.0f, 0.0f);
#pragma unroll
        for (int g = 0; g < 9; ++g) { az01[g] = make_float2(0.0f, 0.0f); az2[g] = 0.0f; }
    }
    __device__ __forceinline__ void fetch(uint32_t base, uint32_t cnt, int kind) { st.fetch(base, cnt, kind); }
    __device__ __forceinline__ void hint_chunk(uint32_t slot) { st.hint_chunk(slot); }
    __device__ __forceinline__ void take(int pending) { st.take(pending); }
    __device__ __forceinline__ void compute(uint32_t, uint32_t t)
    {
        float4 ra, rc, rd4;
        const float c2 = st.load_no_vel(t, ra, rc, rd4);  // (px, py, pz, m) . (c0, c1, c3, c4) (c6, c7, c5, c8)
        const float mass = ra.w;
        float2 wxy[3], dxy[3];
        float wz[3], dz[3];
        cell_axis2(make_float2(ra.x, ra.y), cp.fcxy, wxy, dxy);
        cell_axis(ra.z, cp.fcz, wz, dz);
        // density = sum over the stencil of mass * weight, by sum factorisation (z, then y, then x); rows gy = 0, 1 packed
        float density = 0.0f;
#pragma unroll
        for (int gx = 0; gx < 3; ++gx) {
            const float2 r01 = __ffma2_rn(gm01[gx * 3 + 2], bc(wz[2]), __ffma2_rn(gm01[gx * 3 + 1], bc(wz[1]), __fmul2_rn(gm01[gx * 3], bc(wz[0]))));
            const float r2 = fmaf(gm2[gx * 3 + 2], wz[2], fmaf(gm2[gx * 3 + 1], wz[1], gm2[gx * 3] * wz[0]));
            const float sx = fmaf(r2, wxy[2].y, fmaf(r01.y, wxy[1].y, r01.x * wxy[0].y));
            density = fmaf(sx, wxy[gx].x, density);
        }
        // eq_16_term_0 = -volume * 4 * stress * dt (symmetric), pre-scaled to fixed-point units
        const float volume = __fdividef(mass, density);
        const float pr = P.eos_k * (cell_eos_pow(density * inv_rest, P) - 1.0f);
        const float pressure = fmaxf(-0.1f, pr);
        const float s = -volume * 4.0f * P.dt * P.fmult;
        const float smu = s * P.visc, sp = s * pressure;
        const float e00 = fmaf(2.0f * smu, rc.x, -sp), e11 = fmaf(2.0f * smu, rc.w, -sp), e22 = fmaf(2.0f * smu, rd4.w, -sp);
        const float e01 = smu * (rc.y + rc.z), e02 = smu * (c2 + rd4.x), e12 = smu * (rd4.z + rd4.y);
        // node value = w * (E d), E symmetric; (x, y) as one pair per node, z as a pair over gz = 0, 1 and a scalar
        const float2 ex = make_float2(e00, e01), ey = make_float2(e01, e11), ez = make_float2(e02, e12);
        const float2 wz01 = make_float2(wz[0], wz[1]), dz01 = make_float2(dz[0], dz[1]);
#pragma unroll
        for (int gx = 0; gx < 3; ++gx) {
            const float2 f0 = __fmul2_rn(ex, bc(dxy[gx].x));
            const float fz0 = e02 * dxy[gx].x;
#pragma unroll
            for (int gy = 0; gy < 3; ++gy) {
                const int g = gx * 3 + gy;
                const float2 f1 = __ffma2_rn(ey, bc(dxy[gy].y), f0);
                const float fz1 = fmaf(e12, dxy[gy].y, fz0);
                const float wgt = wxy[gx].x * wxy[gy].y;
                const float2 w01 = __fmul2_rn(bc(wgt), wz01);
                const float w2 = wgt * wz[2];
                az01[g] = __ffma2_rn(w01, __ffma2_rn(bc(e22), dz01, bc(fz1)), az01[g]);
                az2[g] = fmaf(w2, fmaf(e22, dz[2], fz1), az2[g]);
                axy[3 * g + 0] = __ffma2_rn(bc(w01.x), __ffma2_rn(ez, bc(dz[0]), f1), axy[3 * g + 0]);
                axy[3 * g + 1] = __ffma2_rn(bc(w01.y), __ffma2_rn(ez, bc(dz[1]), f1), axy[3 * g + 1]);
                axy[3 * g + 2] = __ffma2_rn(bc(w2), __ffma2_rn(ez, bc(dz[2]), f1), axy[3 * g + 2]);
            }
        }
    }
    __device__ __forceinline__ void end_chunk(bool has)
    {
        if (!has) return;
#pragma unroll
        for (int gx = 0; gx < 3; ++gx)
#pragma unroll
            for (int gy = 0; gy < 3; ++gy)
#pragma unroll
                for (int gz = 0; gz < 3; ++gz) {
                    const int g = gx * 3 + gy, n = g * 3 + gz;
                    const int idx = cp.base + gx * TL::PX + gy * TL::PY + gz;
                    const float mz = gz == 0 ? az01[g].x : gz == 1 ? az01[g].y : az2[g];
                    atomicAdd(&tile[0][idx], __float2int_rz(axy[n].x));
                    atomicAdd(&tile[1][idx], __float2int_rz(axy[n].y));
                    atomicAdd(&tile[2][idx], __float2int_rz(mz));
                }
    }
    __device__ __forceinline__ void finish() {}
};

template <int B>
struct P2G2Smem {  // byte sizes of the dynamic shared-memory regions of k_p2g2_cell
    static constexpr size_t RAW = (sizeof(int4) * CTile<B>::NODES + 127) / 128 * 128;  // the tile as the TMA unit delivers it
    static constexpr size_t ACC = sizeof(int) * 3 * CTile<B>::WORDS;                     // momentum accumulators
    static constexpr size_t MASS = sizeof(float) * CTile<B>::WORDS;                      // node masses of the stencils
    static constexpr size_t STAGE = sizeof(float) * CellCfg<B>::NWARP * 2 * RowStage::WORDS;
    static constexpr size_t TOTAL = RAW + ACC + MASS + STAGE;
};

// The node masses P2G_2's density sums need are the .w words of the block's grid tile: like G2P, the kernel asks the TMA
// unit for the NEXT block's tile before it walks the current one (the whole 16-byte cells: the mass words alone would be
// 4-byte pieces at a 16-byte stride, and the sectors are the same).
template <int B>
__global__ void __launch_bounds__(CellCfg<B>::THREADS, (B == 8) ? MPM_P2G_CTAS : 6) k_p2g2_cell(DevParams P, TileGeom g, ParticleView pv, CellArgs a,
                                                                                     int* __restrict__ grid, const float* __restrict__ rec,
                                                                                     const uint32_t* __restrict__ src_of,
                                                                                     const __grid_constant__ CUtensorMap grid_map)
{
    pdl_prologue();
    using TL = CTile<B>;
    using CF = CellCfg<B>;
    using SM = P2G2Smem<B>;
    extern __shared__ __align__(128) unsigned char dsm128[];  // raw tile | accumulators | mass tile | staging
    int4* raw = reinterpret_cast<int4*>(dsm128);
    int (*tile)[TL::WORDS] = reinterpret_cast<int (*)[TL::WORDS]>(dsm128 + SM::RAW);
    float* tmass = reinterpret_cast<float*>(dsm128 + SM::RAW + SM::ACC);
    float* stg = reinterpret_cast<float*>(dsm128 + SM::RAW + SM::ACC + SM::MASS);
    __shared__ __align__(8) unsigned long long tile_bar;
    __shared__ BlockWork s_bw;
    __shared__ int s_next;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float inv_mult = 1.0f / P.fmult;
    constexpr unsigned TILE_BYTES = sizeof(int4) * TL::NODES;
    auto request_tile = [&](int blk) {  // (one thread)
        TL t2; t2.init(g, blk);
        mbar_expect_tx(&tile_bar, TILE_BYTES);
        tma_load_tile(raw, &grid_map, t2.oz, t2.oy, t2.ox - P.gx0, &tile_bar);
    };
    auto claim = [&]() -> int {  // (one thread) next non-empty block of the list, or -1
        const uint32_t bi = atomicAdd(&a.misc[BIN_WORK_P2G2], 1u);
        return (bi < a.misc[BIN_N_ACTIVE]) ? (int)a.active[bi] : -1;
    };
    if (threadIdx.x == 0) {
        mbar_init(&tile_bar, 1);
        s_next = claim();
        if (s_next >= 0) request_tile(s_next);
    }
    __syncthreads();
    int b = s_next;
    unsigned phase = 0;
    while (b >= 0) {
        TL tl; tl.init(g, b);
        for (int k = threadIdx.x; k < 3 * TL::WORDS; k += CF::THREADS) (&tile[0][0])[k] = 0;
        mbar_wait(&tile_bar, phase);
        phase ^= 1u;
        for (int k = threadIdx.x; k < TL::NODES; k += CF::THREADS) tmass[k] = (float)raw[k].w * inv_mult;  // (nodes outside the grid: 0)
        __syncthreads();  // mass tile and cleared accumulators are in place, raw is free again
        if (threadIdx.x == 0) {
            s_bw.next_chunk = CF::NWARP;  // chunk `warp` is every warp's first one
            s_next = claim();
            if (s_next >= 0) request_tile(s_next);  // in flight while this block is walked
        }
        __syncthreads();
        const int b_next = s_next;
        P2G2Body<B> body(P, tl, tile, tmass, RowStage(rec, src_of, stg + warp * (2 * RowStage::WORDS), lane));
        walk_chunks<B>(a, b, lane, warp, &s_bw, body);
        __syncthreads();
        for (int k = threadIdx.x; k < TL::NODES; k += CF::THREADS) {
            int idx; int64_t ci;
            const bool ok = tl.node(P, k, idx, ci);
            const int vx = tile[0][idx], vy = tile[1][idx], vz = tile[2][idx];
            if (!ok || (vx | vy | vz) == 0) continue;
            int* cc = grid + 4 * ci;
            if (vx) atomicAdd(cc + 0, vx);
            if (vy) atomicAdd(cc + 1, vy);
            if (vz) atomicAdd(cc + 2, vz);
        }
        __syncthreads();  // the accumulators are cleared again at the top
        b = b_next;
    }
}

// ---------------------------------------------------------------- G2P
template <int B, bool COMM, bool EXTRA>  // COMM: multi-GPU (lists the particles that leave the slab); EXTRA: sphere list
struct G2PBody {
    using TL = CTile<B>;
    const DevParams& P;
    const ParticleView& pv;
    const TL& tl;
    const float (*tv)[TL::WORDS];
    const KeyGeom& kg;
    uint32_t nslots;
    uint32_t* keys;
    uint32_t* cnt_next;
    const MigClassify& mg;
    int lane;
    CellPos<B> cp;
    float2 gxy[27];        // node velocities of the cell's stencil: (x, y) packed for FFMA2, z separate
    float gvz[27];
    QuadStage st;          // (px, py, pz, m) of the current / next unit, staged in shared memory
    float4* rec;           // output records
    __device__ __forceinline__ G2PBody(const DevParams& P_, const ParticleView& pv_, const TL& tl_, const float (*tv_)[TL::WORDS],
                                       const KeyGeom& kg_, uint32_t nslots_, uint32_t* keys_, uint32_t* cnt_next_, const MigClassify& mg_, int lane_, float4* rec_,
                                       const QuadStage& st_)
        : P(P_), pv(pv_), tl(tl_), tv(tv_), kg(kg_), nslots(nslots_), keys(keys_), cnt_next(cnt_next_), mg(mg_), lane(lane_), st(st_), rec(rec_) {}
    __device__ __forceinline__ void begin_chunk(int L)
    {
        cp.set(tl, L);
#pragma unroll
        for (int gx = 0; gx < 3; ++gx)
#pragma unroll
            for (int gy = 0; gy < 3; ++gy)
#pragma unroll
                for (int gz = 0; gz < 3; ++gz) {
                    const int n = (gx * 3 + gy) * 3 + gz, idx = cp.base + gx * TL::PX + gy * TL::PY + gz;
                    gxy[n] = make_float2(tv[0][idx], tv[1][idx]); gvz[n] = tv[2][idx];
                }
    }
    __device__ __forceinline__ void fetch(uint32_t base, uint32_t cnt, int kind) { st.fetch(base, cnt, kind); }
    __device__ __forceinline__ void hint_chunk(uint32_t) {}
    __device__ __forceinline__ void take(int pending) { st.take(pending); }
    __device__ __forceinline__ void compute(uint32_t i, uint32_t t)
    {
        const float4 pm = st.load(t);
        const float old[3] = {pm.x, pm.y, pm.z};
        float wx[3], wy[3], wz[3], dx[3], dy[3], dz[3];
        // The particle sits in the cell its thread owns (that is what the binning key says), so the cell coordinate is
        // trunc(p): taking it from the position frees three registers here (G2P: 20 -> 4 bytes of spills, -1 %; in the
        // P2G kernels the same change was slower, they keep the per-thread floats).
        float2 wxy[3], dxy[3];
        cell_axis2(make_float2(old[0], old[1]), make_float2(truncf(old[0]), truncf(old[1])), wxy, dxy);
        cell_axis(old[2], truncf(old[2]), wz, dz);
#pragma unroll
        for (int k = 0; k < 3; ++k) { wx[k] = wxy[k].x; wy[k] = wxy[k].y; dx[k] = dxy[k].x; dy[k] = dxy[k].y; }
        // Sum factorisation (z, then y, then x) with the x and y components packed: fma.rn.f32x2 (FFMA2, new on sm_100)
        // does two FMAs per issue slot, and this loop is issue-bound (ncu: 59 % issue-active at 3 warps per scheduler).
        // The z component rides as pairs too: (s_z, t_z) = sum over gz of (w_z, w_z d_z) * v_z, then (S_z, T_zz), then
        // (v_z, B_zz) -- natural pairs whose other operand is a scalar [first version: scalar FFMAs, 39 more per particle].
        const float wdz[3] = {wz[0] * dz[0], wz[1] * dz[1], wz[2] * dz[2]};
        const float2 wwd[3] = {make_float2(wz[0], wdz[0]), make_float2(wz[1], wdz[1]), make_float2(wz[2], wdz[2])};
        float2 vxy = make_float2(0.f, 0.f), Bxxy = vxy, Byxy = vxy, Bzxy = vxy;  // (x, y) components of v and of B's columns
        float2 vzBzz = vxy;                                                      // (v_z, B_zz)
        float Bxz = 0.f, Byz = 0.f;
#pragma unroll
        for (int gx = 0; gx < 3; ++gx) {
            float2 Sxy = make_float2(0.f, 0.f), Tyxy = Sxy, Tzxy = Sxy, SzTzz = Sxy;
            float Tyz = 0.f;
#pragma unroll
            for (int gy = 0; gy < 3; ++gy) {
                const int n = (gx * 3 + gy) * 3;
                const float2 sxy = __ffma2_rn(bc(wz[2]), gxy[n + 2], __ffma2_rn(bc(wz[1]), gxy[n + 1], __fmul2_rn(bc(wz[0]), gxy[n])));
                const float2 txy = __ffma2_rn(bc(wdz[2]), gxy[n + 2], __ffma2_rn(bc(wdz[1]), gxy[n + 1], __fmul2_rn(bc(wdz[0]), gxy[n])));
                const float2 stz = __ffma2_rn(wwd[2], bc(gvz[n + 2]), __ffma2_rn(wwd[1], bc(gvz[n + 1]), __fmul2_rn(wwd[0], bc(gvz[n]))));
                const float wyd = wy[gy] * dy[gy];
                Sxy = __ffma2_rn(bc(wy[gy]), sxy, Sxy); Tyxy = __ffma2_rn(bc(wyd), sxy, Tyxy); Tzxy = __ffma2_rn(bc(wy[gy]), txy, Tzxy);
                SzTzz = __ffma2_rn(bc(wy[gy]), stz, SzTzz); Tyz = fmaf(wyd, stz.x, Tyz);
            }
            const float wxd = wx[gx] * dx[gx];
            vxy = __ffma2_rn(bc(wx[gx]), Sxy, vxy); Bxxy = __ffma2_rn(bc(wxd), Sxy, Bxxy);
            Byxy = __ffma2_rn(bc(wx[gx]), Tyxy, Byxy); Bzxy = __ffma2_rn(bc(wx[gx]), Tzxy, Bzxy);
            vzBzz = __ffma2_rn(bc(wx[gx]), SzTzz, vzBzz); Bxz = fmaf(wxd, SzTzz.x, Bxz); Byz = fmaf(wx[gx], Tyz, Byz);
        }
        const float vz = vzBzz.x, Bzz = vzBzz.y;
        float v[3] = {vxy.x, vxy.y, vz};
        const float Bx[3] = {Bxxy.x, Bxxy.y, Bxz}, By[3] = {Byxy.x, Byxy.y, Byz}, Bz[3] = {Bzxy.x, Bzxy.y, Bzz};
        const float Bm[9] = {Bx[0], Bx[1], Bx[2], By[0], By[1], By[2], Bz[0], Bz[1], Bz[2]};
        float np[3], cm[9];
        g2p_finish<3, EXTRA>(P, old, Bm, v, np, cm);
        // multi-GPU: particles whose new base cell left this rank's slab are listed for the migration (few per warp).  One
        // that would land beyond the neighbouring slab (|v| dt larger than that slab is wide: numerical outliers of a
        // violent scene) is held back in the neighbour's far plane for this step and counted (MpmStats.slab_jump_clamps).
        bool stays = true;
        int side = 0;
        if constexpr (COMM) {
            const int cx = __float2int_rz(np[0]);
            if (cx < mg.x0 || cx >= mg.x1) {
                stays = false;
                side = cx >= mg.x1;
                if (cx < mg.xl0) { np[0] = (float)mg.xl0 + 0.5f; atomicAdd(mg.cnt + 8, 1u); }
                else if (cx >= mg.xr1) { np[0] = (float)mg.xr1 - 0.5f; atomicAdd(mg.cnt + 8, 1u); }
            }
        }
        // one 64-byte record per particle (field order of the planes): the next step's P2G kernels read it from here
        float4* q = rec + 4 * (size_t)i;
        q[0] = make_float4(np[0], np[1], np[2], pm.w);
        q[1] = make_float4(v[0], v[1], v[2], cm[2]);
        q[2] = make_float4(cm[0], cm[1], cm[3], cm[4]);
        q[3] = make_float4(cm[6], cm[7], cm[5], cm[8]);
        if constexpr (COMM) {
            if (!stays) {
                const uint32_t slot = atomicAdd(mg.cnt + side, 1u);
                if (slot < mg.rec_cap) (side ? mg.leaveR : mg.leaveL)[slot] = i;
            }
        }
        if (cnt_next && stays) {  // bin key of the NEW position for the next step (leavers get theirs where they arrive)
            uint32_t k = cell_key(kg, __float2int_rz(np[0]), __float2int_rz(np[1]), __float2int_rz(np[2]));
            k = k < nslots ? k : nslots - 1;
            keys[i] = k;
            atomicAdd(&cnt_next[k], 1u);  // fire-and-forget RED (taking the rank from the return value here was measured
                                          // slower: +0.11 ms in G2P against -0.04 ms in k_place on C4)
        }
    }
    __device__ __forceinline__ void end_chunk(bool) {}
    __device__ __forceinline__ void finish() {}
};

// 4 CTAs per SM (128 registers, ~30 bytes of spills): measured 0.778 vs 0.823 ms on C4 against 3 CTAs at 158 registers.
// The multi-GPU classification is a separate instantiation: as a run-time branch it cost the single-GPU kernel 0.045 ms.
template <int B>
struct G2PSmem {  // byte sizes of the three dynamic shared-memory regions of k_g2p_cell
    static constexpr size_t RAW = (sizeof(int4) * CTile<B>::NODES + 127) / 128 * 128;
    static constexpr size_t TV = sizeof(float) * 3 * CTile<B>::WORDS;
    static constexpr size_t QUADS = sizeof(float4) * CellCfg<B>::NWARP * 2 * QuadStage::UNIT;
    static constexpr size_t TOTAL = RAW + TV + QUADS;
};

template <int B, bool COMM, bool EXTRA>
__global__ void __launch_bounds__(CellCfg<B>::THREADS, (B == 8) ? MPM_G2P_CTAS : 8) k_g2p_cell(DevParams P, TileGeom g, ParticleView pv, CellArgs a,
                                                                                    const __grid_constant__ CUtensorMap grid_map, int raw_grid, KeyGeom kg, uint32_t nslots,
                                                                                    uint32_t* __restrict__ keys, uint32_t* __restrict__ cnt_next,
                                                                                    MigClassify mg, float4* __restrict__ rec)
{
    pdl_prologue();
    using TL = CTile<B>;
    using CF = CellCfg<B>;
    // dynamic shared memory (above the 48 KB static limit together): the tile as the TMA unit delivers it -- [x][y][z]
    // cells of 4 x int32 -- | the velocity tile the stencils read | the per-warp staging of (px, py, pz, m) units
    // (its own symbol: the bulk tensor copy needs a 128-byte aligned destination, and the alignment of a dynamic array is
    // the one of its first declaration in the translation unit)
    extern __shared__ __align__(128) unsigned char dsm128[];
    int4* raw = reinterpret_cast<int4*>(dsm128);
    float (*tv)[TL::WORDS] = reinterpret_cast<float (*)[TL::WORDS]>(dsm128 + G2PSmem<B>::RAW);
    float4 (*s_quads)[2 * QuadStage::UNIT] = reinterpret_cast<float4 (*)[2 * QuadStage::UNIT]>(dsm128 + G2PSmem<B>::RAW + G2PSmem<B>::TV);
    __shared__ __align__(8) unsigned long long tile_bar;
    __shared__ BlockWork s_bw;
    __shared__ int s_next;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float inv_mult = 1.0f / P.fmult;
    constexpr unsigned TILE_BYTES = sizeof(int4) * TL::NODES;
    auto request_tile = [&](int blk) {  // (one thread)
        TL t2; t2.init(g, blk);
        mbar_expect_tx(&tile_bar, TILE_BYTES);
        tma_load_tile(raw, &grid_map, t2.oz, t2.oy, t2.ox - P.gx0, &tile_bar);
    };
    auto claim = [&]() -> int {  // (one thread) next non-empty block of the list, or -1
        const uint32_t bi = atomicAdd(&a.misc[BIN_WORK_G2P], 1u);
        return (bi < a.misc[BIN_N_ACTIVE]) ? (int)a.active[bi] : -1;
    };
    if (threadIdx.x == 0) {
        mbar_init(&tile_bar, 1);
        s_next = claim();
        if (s_next >= 0) request_tile(s_next);
    }
    __syncthreads();
    int b = s_next;
    unsigned phase = 0;
    while (b >= 0) {
        TL tl; tl.init(g, b);
        mbar_wait(&tile_bar, phase);
        phase ^= 1u;
        for (int k = threadIdx.x; k < TL::NODES; k += CF::THREADS) {
            const int tz = k % TL::T, ty = (k / TL::T) % TL::T, tx = k / (TL::T * TL::T);
            const int idx = k;  // (dense tile: the order the tensor copy delivered)
            const int4 c = raw[k];
            float vx = 0.0f, vy = 0.0f, vz = 0.0f;
            if (!raw_grid) {  // the grid update already ran: cells hold velocities
                vx = (float)c.x * inv_mult; vy = (float)c.y * inv_mult; vz = (float)c.z * inv_mult;
            } else if (c.w > 0) {
                // UpdateGrid fused into the tile load (update_grid.glsl:44-66): v = p / m, gravity on y, zero the
                // wall-normal component for idx < 2 || idx > R - bc_hi_off.  (m and p carry the same fixed-point scale.)
                const float im = __frcp_rn((float)c.w);
                const int nx = tl.ox + tx, ny = tl.oy + ty, nz = tl.oz + tz;
                const int hi = P.bc_hi_off;
                vx = (nx < 2 || nx > P.Rx - hi) ? 0.0f : (float)c.x * im;
                vy = (ny < 2 || ny > P.Ry - hi) ? 0.0f : fmaf((float)c.y, im, P.dt * P.gravity);
                vz = (nz < 2 || nz > P.Rz - hi) ? 0.0f : (float)c.z * im;
            }
            tv[0][idx] = vx; tv[1][idx] = vy; tv[2][idx] = vz;
        }
        __syncthreads();  // tv is complete, raw is free again
        if (threadIdx.x == 0) {
            s_bw.next_chunk = CF::NWARP;  // chunk `warp` is every warp's first one
            s_next = claim();
            if (s_next >= 0) request_tile(s_next);  // in flight while this block is walked
        }
        __syncthreads();
        const int b_next = s_next;
        G2PBody<B, COMM, EXTRA> body(P, pv, tl, const_cast<const float (*)[TL::WORDS]>(tv), kg, nslots, keys, cnt_next, mg, lane, rec,
                                     QuadStage(reinterpret_cast<const float4*>(pv.base), s_quads[warp], lane));
        walk_chunks<B>(a, b, lane, warp, &s_bw, body);
        __syncthreads();  // everyone is done with tv and the work counters
        b = b_next;
    }
}

// ---------------------------------------------------------------- host side
static int check_cell_supported(MpmSolver* s)
{
    if (s->dp.dim != 3 || s->dp.grid_mode != MPM_GRID_FIXED || s->hp.math_mode != MPM_MATH_FAST) {
        s->err = "MPM_PATH_CELL implements dim = 3, MPM_GRID_FIXED, MPM_MATH_FAST";
        return MPM_ERR_INVALID;
    }
    if (!s->sorted_valid) { s->err = "cell kernels need freshly binned particles"; return MPM_ERR_STATE; }
    return MPM_OK;
}

// One process may drive several GPUs (k solvers on the in-process transport): launch geometry and the opt-in for more
// than 48 KB of dynamic shared memory are looked up / set once per (kernel, device).
constexpr int MAX_DEVICES = 64;

template <typename K>
static unsigned persistent_grid(K kernel, int threads, size_t smem)
{
    int per_sm = 1, dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (smem > 0) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    return (unsigned)(per_sm * sms);
}

// SMEM8 / SMEM4: dynamic shared memory of the B = 8 / B = 4 instantiation
#define LAUNCH_CELL(KERNEL, SMEM8, SMEM4, ...)                                                                            \
    do {                                                                                                                  \
        BinState* st = s->bin;                                                                                            \
        TileGeom g{st->nby, st->nbz, s->dp.gx0 + (s->comm ? 1 : 0)};                                                      \
        CellArgs a{st->cnts, st->ord, st->pstart, st->active, st->misc};                                       \
        if (st->B == 8) {                                                                                                 \
            static unsigned grid8_dev[MAX_DEVICES] = {};  /* per device: the shared-memory opt-in is a per-device attribute */ \
            unsigned& grid8 = grid8_dev[s->device & (MAX_DEVICES - 1)];                                                   \
            if (!grid8) grid8 = persistent_grid(KERNEL<8>, CellCfg<8>::THREADS, (SMEM8));                                 \
            launch_pdl<PDL_CELL>(KERNEL<8>, dim3((unsigned)std::min<int64_t>(grid8, st->nblocks)), dim3(CellCfg<8>::THREADS), (SMEM8), s->stream, s->dp, g, s->view(), a, __VA_ARGS__); \
        } else {                                                                                                          \
            static unsigned grid4_dev[MAX_DEVICES] = {};                                                                  \
            unsigned& grid4 = grid4_dev[s->device & (MAX_DEVICES - 1)];                                                   \
            if (!grid4) grid4 = persistent_grid(KERNEL<4>, CellCfg<4>::THREADS, (SMEM4));                                 \
            launch_pdl<PDL_CELL>(KERNEL<4>, dim3((unsigned)std::min<int64_t>(grid4, st->nblocks)), dim3(CellCfg<4>::THREADS), (SMEM4), s->stream, s->dp, g, s->view(), a, __VA_ARGS__); \
        }                                                                                                                 \
        s->launches += 1;                                                                                                 \
    } while (0)

template <int B>
constexpr size_t p2g1_smem() { return sizeof(int) * 4 * CTile<B>::WORDS + sizeof(float) * CellCfg<B>::NWARP * 2 * RowStage::WORDS; }

// 4-D tensor map (channel, z, y, x) of the local grid, box = one block's tile; encoded once per solver
static int grid_tensor_map(MpmSolver* s)
{
    BinState* bs = s->bin;
    if (bs->grid_map_valid) return MPM_OK;
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn || q != cudaDriverEntryPointSuccess) {
        s->err = "cuTensorMapEncodeTiled is not available from this driver";
        return MPM_ERR_CUDA;
    }
    const cuuint32_t T = (cuuint32_t)(bs->B + 2);
    const cuuint64_t dims[4] = {4, (cuuint64_t)s->dp.Rz, (cuuint64_t)s->dp.Ry, (cuuint64_t)s->dp.nxl};
    const cuuint64_t strides[3] = {16, 16ull * s->dp.Rz, 16ull * s->dp.Rz * s->dp.Ry};  // bytes, dimensions 1..3
    const cuuint32_t box[4] = {4, T, T, T}, estr[4] = {1, 1, 1, 1};
    const CUresult r = reinterpret_cast<EncodeFn>(fn)(&bs->grid_map, CU_TENSOR_MAP_DATA_TYPE_INT32, 4, s->grid, dims, strides, box, estr,
                                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { s->err = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")"; return MPM_ERR_CUDA; }
    bs->grid_map_valid = true;
    return MPM_OK;
}

int cell_p2g1(MpmSolver* s)
{
    int rc = check_cell_supported(s);
    if (rc) return rc;
    if (s->n + s->n_launch_extra == 0) return MPM_OK;
    LAUNCH_CELL(k_p2g1_cell, p2g1_smem<8>(), p2g1_smem<4>(), reinterpret_cast<int*>(s->grid), s->rec, s->bin->src_of);
    s->g2p_inputs = true;
    return MPM_OK;
}

int cell_p2g2(MpmSolver* s)
{
    int rc = check_cell_supported(s);
    if (rc) return rc;
    if (s->n + s->n_launch_extra == 0) return MPM_OK;
    if ((rc = grid_tensor_map(s))) return rc;
    LAUNCH_CELL(k_p2g2_cell, P2G2Smem<8>::TOTAL, P2G2Smem<4>::TOTAL, reinterpret_cast<int*>(s->grid), s->rec, s->bin->src_of, s->bin->grid_map);
    return MPM_OK;
}

// the two G2P instantiations as names the launch macro can take
template <int B>
constexpr auto k_g2p_cell_single = k_g2p_cell<B, false, false>;
template <int B>
constexpr auto k_g2p_cell_comm = k_g2p_cell<B, true, false>;
template <int B>
constexpr auto k_g2p_cell_single_x = k_g2p_cell<B, false, true>;
template <int B>
constexpr auto k_g2p_cell_comm_x = k_g2p_cell<B, true, true>;

int cell_g2p(MpmSolver* s)
{
    int rc = check_cell_supported(s);
    if (rc) return rc;
    if (s->n + s->n_launch_extra == 0) return MPM_OK;
    BinState* bs = s->bin;
    if ((rc = bin_g2p_inputs(s))) return rc;
    // multi-GPU: particles that leave the slab are not counted here; the migration moves the keys of the particles it
    // relocates and adds keys + counts for the arrivals (bin_keys_range)
    const bool fuse = true;
    uint32_t* cnt_next = bs->cnt[bs->cur ^ 1];
    MigClassify mg;
    rc = comm_begin_classify(s, &mg);
    if (rc) return rc;
    // The (x, y, z, |v|) hand-off in original index order is a 16-B scatter per particle (0.30 ms of 1.17 ms on C4 when
    // fused here): on this path it is produced on demand by mpm_get_positions instead of every step.
    if ((rc = grid_tensor_map(s))) return rc;
#define G2P_ARGS bs->grid_map, s->grid_raw ? 1 : 0, bin_key_geom(s), (uint32_t)bs->nslots, bs->keys, cnt_next, mg, \
                 reinterpret_cast<float4*>(s->rec)
    const bool extra = s->dp.n_extra > 0;
    if (mg.cnt && extra) LAUNCH_CELL(k_g2p_cell_comm_x, G2PSmem<8>::TOTAL, G2PSmem<4>::TOTAL, G2P_ARGS);
    else if (mg.cnt) LAUNCH_CELL(k_g2p_cell_comm, G2PSmem<8>::TOTAL, G2PSmem<4>::TOTAL, G2P_ARGS);
    else if (extra) LAUNCH_CELL(k_g2p_cell_single_x, G2PSmem<8>::TOTAL, G2PSmem<4>::TOTAL, G2P_ARGS);
    else LAUNCH_CELL(k_g2p_cell_single, G2PSmem<8>::TOTAL, G2PSmem<4>::TOTAL, G2P_ARGS);
#undef G2P_ARGS
    std::swap(s->orig_id, s->orig_id_alt);  // the records are in this step's slot order now, and so are the ids P2G_1 wrote
    s->g2p_inputs = false;
    s->in_rec = true;  // the particle state lives in the records (ensure_planes converts for the plane readers)
    bs->next_valid = fuse;
    s->sorted_valid = false;  // positions moved: the layout is exact for one step only
    return MPM_OK;
}

}  // namespace mpm
