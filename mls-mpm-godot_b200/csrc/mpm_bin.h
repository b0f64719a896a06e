// mpm_bin.h -- state of the counting-sort binning used by the cell kernels (mpm_bin.cu, mpm_kernels_cell.cu).
#pragma once
#include <cuda.h>

#include "mpm_solver.h"

namespace mpm {

// words of BinState::misc (device): reset by the scan every step
enum { BIN_N_ACTIVE = 0, BIN_WORK_P2G1, BIN_WORK_P2G2, BIN_WORK_G2P, BIN_MISC_WORDS = 8 };

struct BinState {
    int B = 8, logB = 3, cell_bits = 9;  // grid block edge (cells), bits of the cell-in-block id
    int nbx = 0, nby = 0, nbz = 0;
    int64_t nblocks = 0;
    int64_t nslots = 0;        // nblocks << cell_bits: one count entry per (block, cell)
    uint32_t* cnt[2] = {nullptr, nullptr};  // particle count per (block, cell); cnt[cur] describes the current layout,
    int cur = 0;                            // cnt[cur ^ 1] is being accumulated by G2P for the next binning
    bool next_valid = false;                // keys[] and cnt[cur ^ 1] were produced by the last G2P
    // layout metadata of the current binning: per block 2 * NC "virtual cell" positions (a cell with more than 32 particles
    // is split), ordered by count, descending, in chunks of 32
    uint32_t* cnts = nullptr;        // [2 * nslots] count at each position
    uint16_t* ord = nullptr;         // [2 * nslots] position -> cell id inside the block
    uint2* cellmeta = nullptr;       // [nslots] cell -> {first full virtual cell | remainder's position << 16, number of full ones}
    uint32_t* pstart = nullptr;      // [2 * nslots / 32] first slot of every chunk
    uint16_t* stab = nullptr;        // [2 * nslots / 32][32] per chunk: [0] irregular flag, [1..13] and [16..31] first slot of the rank-r row (r = 1..13, 14..29)
                                     // relative to the chunk start, [14..15] the chunk start (one 32-B sector for k_place)
    // per-block totals, their exclusive scan and the list of non-empty blocks, double-buffered: [lay] describes the current
    // layout, [lay ^ 1] the previous one -- which is the order the records (and keys[]) are in when the next binning runs,
    // and what its stable ranking walks (a block's particles = one contiguous run of old slots = one "tile")
    uint32_t* bsum2[2] = {nullptr, nullptr};    // [nblocks + pad] particles per block
    uint32_t* bbase2[2] = {nullptr, nullptr};   // [nblocks + 1 + pad] exclusive scan
    uint32_t* active2[2] = {nullptr, nullptr};  // [nblocks] non-empty blocks, ascending
    uint32_t* nact = nullptr;                   // [2] length of active2[k]
    int lay = 0;
    int prev_lay = 0;                // buffers the last binning used as the "previous layout" (= lay for a cold binning)
    bool lay_valid = false;          // [lay] describes the order of the records (false after an upload / migration)
    uint32_t* bsum = nullptr;        // = bsum2[lay]
    uint32_t* bbase = nullptr;       // = bbase2[lay]
    uint32_t* fill = nullptr;        // [nslots] placement cursor (atomic ranking only)
    uint32_t* farcnt = nullptr;      // [nslots] far arrivals per cell (stable ranking): count | tickets << 16
    // stable ranking (k_rank_count / k_rank_place): particles of old block T that go to cell r of T's (B+2)^3 region
    uint32_t* tcount = nullptr;      // [nblocks][(B+2)^3]; row T is valid while bsum2[previous][T] > 0
    uint32_t* fixlist = nullptr;     // [FIX_CAP] cells that received "far movers" (particles that left their block's region) in this binning
    uint32_t* heavy = nullptr;       // [HEAVY_CAP] tiles with more than HEAVY_ROWS rows (ranked by a whole CTA each)
    uint32_t* far_n = nullptr;       // [8]: [0] cells listed, [1] binnings with a cell left in atomic order, [2] far movers of this binning,
                                     // [3] flag: this binning left a cell unordered (booked into [1] by the next scan), [4] consecutive
                                     // violent binnings, [5] heavy tiles listed, [6] the far-mover limit in force (MPM_FAR_LIMIT)
    CUtensorMap grid_map;            // TMA descriptor of the local grid (box = one block's tile), for G2P's tile prefetch
    bool grid_map_valid = false;     // (re-encoded when the grid is re-created: multi-GPU re-cuts)
    bool stable = true;              // rank inside a cell = order of the old slots (std::stable_sort); false: atomic cursor
    uint32_t* keys = nullptr;        // [pitch] cell key of each particle for the NEXT binning (slot order)
    uint32_t* src_of = nullptr;      // [pitch + 64] slot -> index of the particle's record (the records stay where G2P wrote them)
    uint32_t* active = nullptr;      // = active2[lay]
    uint32_t* misc = nullptr;        // BIN_MISC_WORDS counters
    int* box = nullptr;              // [12] cells: bounding box of the non-empty blocks (+ apron) of the current binning, and its
                                     // union with those since the last clear (what a sparse clear has to cover); see k_scan_blocks
    bool box_cleared = true;         // the grid has been cleared since the last binning
};

int bin_create(MpmSolver* s);
void bin_destroy(MpmSolver* s);
int bin_particles(MpmSolver* s);
int bin_g2p_inputs(MpmSolver* s);  // position / mass planes in slot order, when no P2G_1 ran since the binning
// cell keys of the last binning's input (record order) and its permutation (cell-major rank -> input index), re-derived
// by a verification run of the ranking kernels that also checks the layout in place against it
int bin_debug_last(MpmSolver* s, uint32_t* keys_before, uint32_t* perm, int64_t cap);
constexpr int FIX_CAP = 1 << 21;  // cells with far arrivals a binning can put in order (8 MB)

// cell kernels (mpm_kernels_cell.cu)
int cell_p2g1(MpmSolver* s);
int cell_p2g2(MpmSolver* s);
int cell_g2p(MpmSolver* s);

}  // namespace mpm
