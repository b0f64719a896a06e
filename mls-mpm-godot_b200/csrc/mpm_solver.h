// mpm_solver.h -- the opaque solver object behind the C ABI (include/mpm_b200.h).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/mpm_b200.h"
#include "mpm_common.cuh"

namespace mpm {
// device memory that can be exported to another API / process as a POSIX file descriptor (mpm_vmm.cu)
struct ExportableAlloc {
    void* ptr = nullptr;
    size_t bytes = 0;
    unsigned long long handle = 0;
};
bool vmm_alloc(int device, size_t bytes, ExportableAlloc* out);
void vmm_free(ExportableAlloc* a);
bool vmm_export_fd(const ExportableAlloc& a, int* fd);
struct SortState;  // mpm_sort.cu
struct BinState;   // mpm_bin.cu
struct CommState;  // mpm_comm.cu
}  // namespace mpm

// PH_EX_MASS / PH_EX_MOM / PH_EX_MIG: the three exchanges of a multi-GPU step (their sum is reported as ms_exchange too)
enum { PH_CLEAR = 0, PH_P2G1, PH_P2G2, PH_UPDATE, PH_G2P, PH_SORT, PH_EXCHANGE, PH_EX_MASS, PH_EX_MOM, PH_EX_MIG, PH_COUNT };

struct MpmSolver {
    MpmParams hp{};
    mpm::DevParams dp{};
    int device = 0;
    cudaStream_t stream = nullptr;

    int64_t cap = 0;    // max particles
    int64_t pitch = 0;  // plane stride (floats)
    int64_t n = 0;      // particles currently held (local)
    // multi-GPU, migration by peer stores: between the moment a migration is enqueued and the moment the host reads its counts,
    // n is the count BEFORE the exchange; kernels launched in between are sized for n + n_launch_extra and take the true count
    // from *n_dev (both are 0 / null otherwise)
    int64_t n_launch_extra = 0;
    const uint32_t* n_dev = nullptr;
    float* part = nullptr;      // NPLANES * pitch floats
    float* part_alt = nullptr;  // reorder target (tiled path)
    // Cell path: the particle state lives in 64-byte records (the 16 fields of a slot, contiguous).  G2P writes them in
    // the slot order of its step; the next binning only computes src_of[new slot] = old slot, and the P2G kernels read the
    // records through it: a record is two full 32-B sectors wherever it sits, while the 16 fields of a grouped slot are
    // 16 different sectors (measured on the evolved C4 dam-break: 25.8 GB of L2 traffic for a 4.4 GB gather from planes).
    float* rec = nullptr;       // [pitch] x 16 floats
    bool grid_raw = false;      // cell path: the grid holds mass + momentum (P2G done, UpdateGrid not yet applied): G2P applies
                                // the update while it loads its tiles, and a grid download applies it first
    bool in_rec = false;        // the particle state of slots [0, n) currently lives in `rec`, not in `part`
    bool g2p_inputs = false;    // cell path: P2G_1 has written the position / mass planes of the current layout (what G2P reads)
    uint32_t* orig_id = nullptr;      // original (global) index of the particle in each slot
    uint32_t* orig_id_alt = nullptr;  // binned paths: ids in the NEW slot order (cell path: swapped in when G2P has rewritten the records)
    void* grid = nullptr;  // ncells_local * 16 B
    int64_t ncells = 0;    // local cells (nxl * Ry * Rz)
    float4* positions = nullptr;  // (x, y, z, |v|) in original index order
    mpm::ExportableAlloc positions_mem;  // ... inside an exportable allocation when the driver offers one (mpm_export_positions)
    bool positions_valid = false;
    // pipelined hand-off (mpm_get_positions_async): second device array, copy stream, events
    float4* positions_b = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t pos_ready[2] = {nullptr, nullptr}, pos_copied[2] = {nullptr, nullptr};
    int pos_buf = 0;
    int32_t* overflow_flag = nullptr;  // device: [0] overflow detector, [1] particles skipped for their position (DevParams::flags)
    void* stage = nullptr;             // grow-only staging buffer of the upload / download calls
    size_t stage_bytes = 0;

    int path = MPM_PATH_REFERENCE;  // resolved kernel path
    int sort_interval = 1;
    int64_t steps = 0, launches = 0;
    int64_t steps_since_sort = 0;
    bool sorted_valid = false;
    bool fresh_particles = true;  // particle set changed since the last bin phase (-> lane interleave once)
    mpm::SortState* sort = nullptr;
    mpm::BinState* bin = nullptr;
    mpm::CommState* comm = nullptr;

    // per-phase timing
    int timing = 0;  // MPM_TIMING_*: 0 off, 1 events around every phase, 2 events around the whole mpm_step() call only
    std::vector<cudaEvent_t> ev;  // 2 events per phase per step, recycled
    double ms_acc[PH_COUNT] = {0};
    double ms_step_acc = 0;
    int64_t timed_steps = 0;

    std::string err;

    mpm::ParticleView view() const { return mpm::ParticleView{part, pitch}; }
    mpm::ParticleView view_alt() const { return mpm::ParticleView{part_alt, pitch}; }
    mpm::RecView rview() const { return mpm::RecView{rec}; }
};

namespace mpm {
// binning (mpm_sort.cu): computes keys, stable radix sort, reorders particle planes into part_alt and swaps.
int sort_create(MpmSolver* s);
void sort_destroy(MpmSolver* s);
int sort_particles(MpmSolver* s);
int sort_debug_last(MpmSolver* s, uint32_t* keys_before, uint32_t* perm, int64_t cap);
int radix_sort_pairs(uint32_t* keys[2], uint32_t* vals[2], int64_t n, int first_bit, int key_bits, cudaStream_t stream, int64_t* launches, std::string* err);
// tiled kernels (mpm_kernels_tiled.cu)
int tiled_p2g1(MpmSolver* s);
int tiled_p2g2(MpmSolver* s);
int tiled_g2p(MpmSolver* s);
// particle state back into the grouped planes if the last G2P left it in `rec` (mpm_bin.cu)
int ensure_planes(MpmSolver* s);
// multi-GPU (mpm_comm.cu)
struct MigClassify {  // handed to a G2P kernel that classifies leaving particles itself; cnt == nullptr: nothing to do
    int x0, x1, xl0, xr1;
    uint32_t* cnt;      // [0] leaving left, [1] leaving right, [8] crossed more than one slab
    uint32_t* leaveL;
    uint32_t* leaveR;
    uint32_t rec_cap;
};
int comm_begin_classify(MpmSolver* s, MigClassify* out);
void comm_destroy(MpmSolver* s);
int comm_exchange_halo(MpmSolver* s, int pass);  // pass 0: after P2G_1 (4 words), 1: after P2G_2 (3 words)
int comm_migrate(MpmSolver* s);
}  // namespace mpm
