// mpm_tile.cuh -- geometry of the shared-memory grid tile used by the tiled kernels (strict and fast).
#pragma once
#include "mpm_common.cuh"

namespace mpm {

constexpr int TILED_THREADS = 256;

struct TileGeom {
    int nby, nbz;
    int x_owned0;  // global x of the first plane covered by blocks (gx0, or gx0+1 with a ghost plane)
};

// Tile of (B+2)^3 nodes, padded so that 32 consecutive cells of a block (z fastest, then y, then x) fall into
// 32 distinct shared-memory banks for any fixed stencil offset: B = 8 -> row pitch 24 words (each 8-cell row
// moves the bank octet by -1 mod 4), plane pitch 256; B = 4 -> row pitch 12, plane pitch 80 (row r -> bank 12r).
template <int B>
struct Tile {
    static constexpr int T = B + 2;
    static constexpr int PY = (B == 8) ? 24 : 12;
    static constexpr int PX = (B == 8) ? 256 : 80;
    static constexpr int WORDS = T * PX;  // per channel, padding included
    static constexpr int NODES = T * T * T;
    int ox, oy, oz;
    __device__ __forceinline__ void init(const TileGeom& g, int b)
    {
        const int bz = b % g.nbz, by = (b / g.nbz) % g.nby, bx = b / (g.nbz * g.nby);
        ox = g.x_owned0 + bx * B - 1; oy = by * B - 1; oz = bz * B - 1;
    }
    __device__ __forceinline__ bool stencil_base(int cx, int cy, int cz, int& idx) const
    {
        const int tx = cx - 1 - ox, ty = cy - 1 - oy, tz = cz - 1 - oz;
        idx = tx * PX + ty * PY + tz;
        return (unsigned)tx <= (unsigned)(T - 3) && (unsigned)ty <= (unsigned)(T - 3) && (unsigned)tz <= (unsigned)(T - 3);
    }
    // k-th real node of the tile -> padded index and global cell index (false if outside the local grid)
    __device__ __forceinline__ bool node(const DevParams& P, int k, int& idx, int64_t& ci) const
    {
        const int tz = k % T, ty = (k / T) % T, tx = k / (T * T);
        idx = tx * PX + ty * PY + tz;
        const int nx = ox + tx, ny = oy + ty, nz = oz + tz;
        if (nx < P.gx0 || nx >= P.gx0 + P.nxl || ny < 0 || ny >= P.Ry || nz < 0 || nz >= P.Rz) return false;
        ci = cell_index(P, nx, ny, nz);
        return true;
    }
};

// ---- bin key: block-major, cell-minor.  2D uses BxB blocks with bz = lz = 0.  The block id is the reference's
// cell index formula (MLSMPM3DFluidMultithread.cs:282) applied to block coordinates.
struct KeyGeom {
    int dim, logB, nby, nbz, gx0;  // gx0 = first OWNED x plane (block origin of the slab)
};

__device__ __forceinline__ uint32_t cell_key(const KeyGeom& g, int cx, int cy, int cz)
{
    const int m = (1 << g.logB) - 1;
    const int lx = cx - g.gx0;
    const int bx = lx >> g.logB, by = cy >> g.logB, bz = cz >> g.logB;
    const uint32_t blk = (uint32_t)((bx * g.nby + by) * g.nbz + bz);
    if (g.dim == 3) return (blk << (3 * g.logB)) | (uint32_t)(((((lx & m) << g.logB) | (cy & m)) << g.logB) | (cz & m));
    return (blk << (2 * g.logB)) | (uint32_t)(((lx & m) << g.logB) | (cy & m));
}

}  // namespace mpm
