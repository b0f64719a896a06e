// mpm_comm.cu -- multi-GPU x-slab decomposition: slab cuts, halo exchange-add, particle migration.
//
// The reference is single-device; this is new.  A rank owns grid planes [x0, x1) (contiguous in the reference's
// cell order x*Ry*Rz + y*Rz + z, MLSMPM3DFluidMultithread.cs:282) and the particles whose base cell x lies
// there; it stores planes [x0-1, x1+1).  The quadratic B-spline stencil reaches +-1 cell around the base cell
// (MLSMPM3DFluidMultithread.cs:267-275), so a rank scatters into [x0-1, x1] and only nearest neighbours talk.
//
// Per step and neighbour pair
//   after P2G_1 : both sides send their copy of the two overlap planes (the ghost plane and the boundary owned
//                 plane are adjacent in memory: one contiguous block of 2*Ry*Rz cells) and add what they receive
//                 -> both hold complete sums (P2G_2's density gather needs complete mass on the ghost plane).
//   after P2G_2 : the same with the increment since the first exchange (block - snapshot).
//   grid update : pointwise, run redundantly on the ghost planes (no exchange).
//   after G2P   : particles whose base cell left [x0, x1) move to the neighbour (counts first, then records).
// All grid traffic is int32 adds, which commute: k ranks give the bits 1 rank gives.
//
// Transports: NCCL send/recv (one process per GPU) and LOCAL (k solvers in one process: peer copies ordered by
// CUDA events, rendezvous through a mutex/condvar mailbox per directed edge).  Both are stream-ordered.
#include <dlfcn.h>
#include <nccl.h>  // types only: every NCCL function is resolved with dlsym

#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <vector>

#include "mpm_kernels.h"
#include "mpm_solver.h"
#include "mpm_tile.cuh"

namespace mpm {

int sort_create(MpmSolver* s);
void sort_destroy(MpmSolver* s);
int bin_create(MpmSolver* s);
void bin_destroy(MpmSolver* s);
uint32_t* bin_next_keys(MpmSolver* s);
int bin_keys_range(MpmSolver* s, int64_t first, int64_t count);

// ================================================================ transports
struct Transport {
    int rank = 0, world = 1;
    virtual ~Transport() {}
    // Stream-ordered neighbour exchange.  L = rank-1, R = rank+1.  A size of 0 skips that message; both ends of an
    // edge must agree on its size.  Send buffers may be rewritten by work enqueued on `st` after the call returns.
    virtual int exchange(const void* sendL, size_t nsl, void* recvL, size_t nrl, const void* sendR, size_t nsr,
                         void* recvR, size_t nrr, cudaStream_t st, std::string& err) = 0;
    // true when every rank is its own process on its own GPU: halo planes can then be written straight into the
    // neighbour's memory over NVLink (CUDA IPC) with kernels that wait on each other's flags
    virtual bool separate_gpus() const { return false; }
    // ranks of one process (LOCAL): hand every rank the device pointers of its neighbours' receive allocation, so that
    // the same peer-store halo kernels run as with CUDA IPC between processes.  `mine` may be null (set-up failed here);
    // returns true only if EVERY rank of the chain offered a pointer (all ranks then take the peer-store path).
    virtual bool share_pointers(void* mine, int device, void** left, void** right, int* dev_left, int* dev_right) { return false; }
    // ... and order the neighbours' streams by events instead of letting kernels spin on each other's flags: ranks that share
    // ONE GPU cannot rely on their kernels being co-resident (a kernel spinning for a flag can keep the kernel that would
    // set it from ever being scheduled).  notify: my push of message `seq` (pass 0 / 1) towards side L / R is enqueued on st;
    // await: make st wait for the push of message `seq` from that side.
    virtual int notify(int side, int pass, uint32_t seq, cudaStream_t st, std::string& err) { return MPM_OK; }
    virtual int await(int side, int pass, uint32_t seq, cudaStream_t st, std::string& err) { return MPM_OK; }
};

// ---------------------------------------------------------------- NCCL
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi* nccl_api(std::string& err)
{
    static NcclApi api;
    static std::mutex m;
    std::lock_guard<std::mutex> lk(m);
    if (api.handle) return &api;
    // prefer the copy this process already has (a Python host has torch's), else the system library
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { err = std::string("cannot load libnccl.so.2: ") + dlerror(); return nullptr; }
#define SYM(field, name)                                                     \
    do {                                                                     \
        *(void**)(&api.field) = dlsym(h, name);                              \
        if (!api.field) { err = std::string("libnccl lacks ") + name; return nullptr; } \
    } while (0)
    SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank"); SYM(CommDestroy, "ncclCommDestroy");
    SYM(Send, "ncclSend"); SYM(Recv, "ncclRecv"); SYM(GroupStart, "ncclGroupStart"); SYM(GroupEnd, "ncclGroupEnd");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    api.handle = h;
    return &api;
}

struct NcclTransport : Transport {
    NcclApi* api = nullptr;
    ncclComm_t comm = nullptr;
    ~NcclTransport() override { if (comm) api->CommDestroy(comm); }
    bool separate_gpus() const override { return true; }
    int exchange(const void* sendL, size_t nsl, void* recvL, size_t nrl, const void* sendR, size_t nsr, void* recvR,
                 size_t nrr, cudaStream_t st, std::string& err) override
    {
        ncclResult_t r = api->GroupStart();
        if (rank > 0) {
            if (r == ncclSuccess && nsl) r = api->Send(sendL, nsl, ncclInt8, rank - 1, comm, st);
            if (r == ncclSuccess && nrl) r = api->Recv(recvL, nrl, ncclInt8, rank - 1, comm, st);
        }
        if (rank < world - 1) {
            if (r == ncclSuccess && nsr) r = api->Send(sendR, nsr, ncclInt8, rank + 1, comm, st);
            if (r == ncclSuccess && nrr) r = api->Recv(recvR, nrr, ncclInt8, rank + 1, comm, st);
        }
        ncclResult_t e = api->GroupEnd();
        if (r == ncclSuccess) r = e;
        if (r != ncclSuccess) { err = std::string("NCCL: ") + api->GetErrorString(r); return MPM_ERR_COMM; }
        return MPM_OK;
    }
};

// ---------------------------------------------------------------- LOCAL (k solvers in one process)
struct Mailbox {  // one directed edge src -> dst
    std::mutex m;
    std::condition_variable cv;
    uint64_t posted = 0, consumed = 0;
    const void* ptr = nullptr;
    size_t bytes = 0;
    int src_device = 0;
    cudaEvent_t ready = nullptr;  // sender: data final on its stream
    cudaEvent_t done = nullptr;   // receiver: its copy has been enqueued up to here
};

}  // namespace mpm

struct MpmLocalHub {
    int world = 0;
    std::vector<mpm::Mailbox> to_right;  // [r]: r -> r+1
    std::vector<mpm::Mailbox> to_left;   // [r]: r -> r-1
    int timeout_s = 60;
    // one-time rendezvous of the peer-store halo set-up: every rank publishes the device pointer of its receive allocation
    std::mutex pm;
    std::condition_variable pcv;
    std::vector<void*> p2p_ptr;
    std::vector<int> p2p_dev;
    int p2p_arrived = 0;
    // halo signals of the peer-store path between ranks of one process: [rank][side][pass]
    struct Signal {
        std::mutex m;
        std::condition_variable cv;
        uint32_t posted = 0, consumed = 0;
        cudaEvent_t ev = nullptr;
    };
    std::vector<Signal> sig;  // world * 4
};

namespace mpm {

struct LocalTransport : Transport {
    MpmLocalHub* hub = nullptr;
    int device = 0;

    int post(Mailbox& mb, const void* ptr, size_t bytes, cudaStream_t st)
    {
        std::lock_guard<std::mutex> lk(mb.m);
        if (!mb.ready) cudaEventCreateWithFlags(&mb.ready, cudaEventDisableTiming);
        cudaEventRecord(mb.ready, st);
        mb.ptr = ptr; mb.bytes = bytes; mb.src_device = device;
        mb.posted += 1;
        mb.cv.notify_all();
        return MPM_OK;
    }
    int pull(Mailbox& mb, void* dst, size_t bytes, cudaStream_t st, std::string& err)
    {
        std::unique_lock<std::mutex> lk(mb.m);
        if (!mb.cv.wait_for(lk, std::chrono::seconds(hub->timeout_s), [&] { return mb.posted > mb.consumed; })) {
            err = "local transport: the neighbouring rank did not arrive (every rank must be stepping concurrently)";
            return MPM_ERR_COMM;
        }
        if (mb.bytes != bytes) { err = "local transport: message size mismatch between neighbours"; return MPM_ERR_COMM; }
        cudaStreamWaitEvent(st, mb.ready, 0);
        cudaError_t e = (mb.src_device == device) ? cudaMemcpyAsync(dst, mb.ptr, bytes, cudaMemcpyDeviceToDevice, st)
                                                  : cudaMemcpyPeerAsync(dst, device, mb.ptr, mb.src_device, bytes, st);
        if (e != cudaSuccess) { err = std::string("local transport copy: ") + cudaGetErrorString(e); return MPM_ERR_CUDA; }
        if (!mb.done) cudaEventCreateWithFlags(&mb.done, cudaEventDisableTiming);
        cudaEventRecord(mb.done, st);
        mb.consumed += 1;
        mb.cv.notify_all();
        return MPM_OK;
    }
    // my later work on `st` (which may rewrite the send buffer) must follow the receiver's copy
    int settle(Mailbox& mb, cudaStream_t st, std::string& err)
    {
        std::unique_lock<std::mutex> lk(mb.m);
        if (!mb.cv.wait_for(lk, std::chrono::seconds(hub->timeout_s), [&] { return mb.consumed == mb.posted; })) {
            err = "local transport: the neighbouring rank did not take its message";
            return MPM_ERR_COMM;
        }
        cudaStreamWaitEvent(st, mb.done, 0);
        return MPM_OK;
    }
    bool share_pointers(void* mine, int dev, void** left, void** right, int* dev_left, int* dev_right) override
    {
        std::unique_lock<std::mutex> lk(hub->pm);
        if ((int)hub->p2p_ptr.size() != world) { hub->p2p_ptr.assign(world, nullptr); hub->p2p_dev.assign(world, 0); }
        hub->p2p_ptr[rank] = mine; hub->p2p_dev[rank] = dev;
        hub->p2p_arrived += 1;
        hub->pcv.notify_all();
        if (!hub->pcv.wait_for(lk, std::chrono::seconds(hub->timeout_s), [&] { return hub->p2p_arrived >= world; })) return false;
        bool all = true;
        for (int r = 0; r < world; ++r) all = all && hub->p2p_ptr[r] != nullptr;
        *left = rank > 0 ? hub->p2p_ptr[rank - 1] : nullptr;
        *right = rank < world - 1 ? hub->p2p_ptr[rank + 1] : nullptr;
        *dev_left = rank > 0 ? hub->p2p_dev[rank - 1] : dev;
        *dev_right = rank < world - 1 ? hub->p2p_dev[rank + 1] : dev;
        return all;
    }
    MpmLocalHub::Signal& signal_of(int from_rank, int side, int pass) { return hub->sig[(size_t)(from_rank * 2 + side) * 3 + pass]; }  // pass 2 = migration
    int notify(int side, int pass, uint32_t seq, cudaStream_t st, std::string& err) override
    {
        MpmLocalHub::Signal& sg = signal_of(rank, side, pass);
        std::unique_lock<std::mutex> lk(sg.m);
        // the event is recorded anew for every message: the receiver must have taken the previous one
        if (!sg.cv.wait_for(lk, std::chrono::seconds(hub->timeout_s), [&] { return sg.consumed + 1 >= seq; })) {
            err = "local transport: the neighbouring rank did not take the previous halo message";
            return MPM_ERR_COMM;
        }
        if (!sg.ev) cudaEventCreateWithFlags(&sg.ev, cudaEventDisableTiming);
        cudaEventRecord(sg.ev, st);
        sg.posted = seq;
        sg.cv.notify_all();
        return MPM_OK;
    }
    int await(int side, int pass, uint32_t seq, cudaStream_t st, std::string& err) override
    {
        // the message from my LEFT neighbour is the one it sent to ITS right side (1), and vice versa
        MpmLocalHub::Signal& sg = signal_of(side == 0 ? rank - 1 : rank + 1, side == 0 ? 1 : 0, pass);
        std::unique_lock<std::mutex> lk(sg.m);
        if (!sg.cv.wait_for(lk, std::chrono::seconds(hub->timeout_s), [&] { return sg.posted >= seq; })) {
            err = "local transport: the neighbouring rank did not arrive (every rank must be stepping concurrently)";
            return MPM_ERR_COMM;
        }
        cudaStreamWaitEvent(st, sg.ev, 0);
        sg.consumed = seq;
        sg.cv.notify_all();
        return MPM_OK;
    }
    int exchange(const void* sendL, size_t nsl, void* recvL, size_t nrl, const void* sendR, size_t nsr, void* recvR,
                 size_t nrr, cudaStream_t st, std::string& err) override
    {
        const bool hasL = rank > 0, hasR = rank < world - 1;
        int rc;
        if (hasL && nsl) post(hub->to_left[rank], sendL, nsl, st);
        if (hasR && nsr) post(hub->to_right[rank], sendR, nsr, st);
        if (hasL && nrl && (rc = pull(hub->to_right[rank - 1], recvL, nrl, st, err))) return rc;
        if (hasR && nrr && (rc = pull(hub->to_left[rank + 1], recvR, nrr, st, err))) return rc;
        if (hasL && nsl && (rc = settle(hub->to_left[rank], st, err))) return rc;
        if (hasR && nsr && (rc = settle(hub->to_right[rank], st, err))) return rc;
        return MPM_OK;
    }
};

// ================================================================ state
constexpr int REC_WORDS = NPLANES + 1;  // 16 particle planes + original index

struct CommState {
    Transport* tr = nullptr;
    int rank = 0, world = 1;
    std::vector<int> cuts;  // world + 1 global x planes
    int x0 = 0, x1 = 0;     // owned planes
    bool slab_set = false;    // the planes hold this rank's local particles and the grid is the slab
    bool pending = false;     // the planes hold the GLOBAL set: partition before the next step / download
    // halo: [0] = left block (local planes 0,1), [1] = right block (local planes nxl-2, nxl-1)
    int4* halo_recv[2] = {nullptr, nullptr};
    int4* halo_send[2] = {nullptr, nullptr};
    int4* halo_snap[2] = {nullptr, nullptr};
    int64_t halo_cells = 0;  // 2 * Ry * Rz
    // direct peer stores for the halo planes (CUDA IPC over NVLink): own = [side][pass] receive regions + flags
    bool p2p_tried = false, p2p_ready = false, p2p_inproc = false;
    uint8_t* p2p_own = nullptr;
    uint8_t* p2p_peer[2] = {nullptr, nullptr};  // the left / right neighbour's allocation, mapped here
    uint32_t p2p_seq[2] = {0, 0};               // messages sent so far, per pass
    uint32_t* p2p_done = nullptr;               // block-completion counters of the push kernels + error flag
    // migration
    uint32_t* d_cnt = nullptr;  // 0 nL, 1 nR (leaving), 2 mL, 3 mR (arriving), 4 holes, 5 fillers, 8 held-back outliers
    uint32_t* h_cnt = nullptr;  // pinned mirror
    uint32_t* send_rec[2] = {nullptr, nullptr};
    uint32_t* recv_rec[2] = {nullptr, nullptr};
    uint32_t* holes = nullptr;
    uint32_t* fillers = nullptr;
    uint32_t* leave[2] = {nullptr, nullptr};  // slots of the particles leaving to the left / right (device lists)
    bool classified = false;                  // this step's G2P already filled leave[] and the counts
    int64_t rec_cap = 0;
    int64_t migrated_out = 0, migrated_in = 0, overflow_rounds = 0;
    int64_t recuts = 0;            // times mpm_comm_rebalance moved a cut
    int64_t slab_jump_clamps = 0;  // particles held back because they would have crossed more than one slab in a step
    int64_t n_global = 0;          // particles of the whole scene (the set the last upload / init handed to every rank)
    uint32_t sent_prev[2] = {0, 0}, recv_prev[2] = {0, 0};  // particles that crossed each edge in the previous step
    // migration by peer stores (cell path, peer-store halos available): the leavers are written straight into the neighbour's
    // receive region, the counts stay on the device, and the host reads them one phase later (comm_finish_migration)
    uint32_t mig_seq = 0;         // migration messages sent so far
    uint32_t* n_dev = nullptr;    // [1] local particle count as the device knows it (written by the unpack kernel)
    cudaEvent_t mig_done = nullptr;
    bool mig_pending = false;     // an exchange is enqueued whose counts the host has not read yet
    int64_t mig_n_before = 0;     // s->n when that exchange was enqueued
};

#define CKM(call)                                                          \
    do {                                                                   \
        cudaError_t e_ = (call);                                           \
        if (e_ != cudaSuccess) {                                           \
            s->err = std::string(#call) + ": " + cudaGetErrorString(e_);   \
            return MPM_ERR_CUDA;                                           \
        }                                                                  \
    } while (0)

static void free_slab_buffers(CommState* c)
{
    for (int k = 0; k < 2; ++k) {
        cudaFree(c->halo_recv[k]); cudaFree(c->halo_send[k]); cudaFree(c->halo_snap[k]);
        c->halo_recv[k] = c->halo_send[k] = c->halo_snap[k] = nullptr;
    }
}

void comm_destroy(MpmSolver* s)
{
    CommState* c = s->comm;
    if (!c) return;
    free_slab_buffers(c);
    if (!c->p2p_inproc) for (int k = 0; k < 2; ++k) if (c->p2p_peer[k]) cudaIpcCloseMemHandle(c->p2p_peer[k]);
    cudaFree(c->p2p_own); cudaFree(c->p2p_done); cudaFree(c->n_dev);
    if (c->mig_done) cudaEventDestroy(c->mig_done);
    for (int k = 0; k < 2; ++k) { cudaFree(c->send_rec[k]); cudaFree(c->recv_rec[k]); }
    cudaFree(c->d_cnt); cudaFree(c->holes); cudaFree(c->fillers); cudaFree(c->leave[0]); cudaFree(c->leave[1]);
    if (c->h_cnt) cudaFreeHost(c->h_cnt);
    delete c->tr;
    delete c;
    s->comm = nullptr;
}

static int comm_attach(MpmSolver* s, Transport* tr, int rank, int world)
{
    if (s->comm) { delete tr; s->err = "a communicator is already attached"; return MPM_ERR_STATE; }
    if (s->hp.dim != 3 || s->hp.grid_mode != MPM_GRID_FIXED) {
        delete tr;
        s->err = "multi-GPU slabs need dim = 3 and MPM_GRID_FIXED (integer halo sums)";
        return MPM_ERR_INVALID;
    }
    CommState* c = new CommState();
    c->tr = tr; c->rank = rank; c->world = world;
    tr->rank = rank; tr->world = world;
    s->comm = c;
    c->rec_cap = std::max<int64_t>(s->cap / 8, 1 << 16);
    if (c->rec_cap > s->cap) c->rec_cap = s->cap;
    CKM(cudaMalloc(&c->d_cnt, 16 * sizeof(uint32_t)));
    CKM(cudaHostAlloc(&c->h_cnt, 16 * sizeof(uint32_t), cudaHostAllocDefault));
    for (int k = 0; k < 2; ++k) {
        CKM(cudaMalloc(&c->send_rec[k], sizeof(uint32_t) * (64 + REC_WORDS * c->rec_cap)));
        CKM(cudaMalloc(&c->recv_rec[k], sizeof(uint32_t) * (64 + REC_WORDS * c->rec_cap)));
    }
    for (int k = 0; k < 2; ++k) CKM(cudaMalloc(&c->leave[k], sizeof(uint32_t) * c->rec_cap));
    CKM(cudaMalloc(&c->holes, sizeof(uint32_t) * 2 * c->rec_cap));
    CKM(cudaMalloc(&c->fillers, sizeof(uint32_t) * 2 * c->rec_cap));
    if (!s->part_alt) {
        CKM(cudaMalloc(&s->part_alt, sizeof(float) * NPLANES * s->pitch));
        CKM(cudaMalloc(&s->orig_id_alt, sizeof(uint32_t) * s->pitch));
    }
    s->sort_interval = 1;  // arrivals are appended unbinned: re-bin every step
    return MPM_OK;
}

int comm_halo_planes(const MpmSolver* s, int side)
{
    const CommState* c = s->comm;
    if (!c || c->world < 2) return 0;
    return (side == 0 ? c->rank > 0 : c->rank < c->world - 1) ? 2 : 0;
}
int64_t comm_global_count(const MpmSolver* s) { return s->comm ? s->comm->n_global : s->n; }
int comm_rank_world(const MpmSolver* s, int* rank, int* world)
{
    *rank = s->comm ? s->comm->rank : 0;
    *world = s->comm ? s->comm->world : 1;
    return MPM_OK;
}

void comm_fill_stats(const MpmSolver* s, MpmStats* st)
{
    if (!s->comm) return;
    st->rank = s->comm->rank;
    st->world = s->comm->world;
    st->slab_jump_clamps = s->comm->slab_jump_clamps;
    st->migrated = s->comm->migrated_out;
    st->halo_peer_exchanges = s->comm->p2p_seq[0] + s->comm->p2p_seq[1];
}

// ================================================================ slab set-up at upload time
template <class View>
__global__ void __launch_bounds__(256) k_xhist(View pv, int64_t n, int rx, unsigned long long* __restrict__ hist)
{
    extern __shared__ uint32_t sh[];
    for (int k = threadIdx.x; k < rx; k += blockDim.x) sh[k] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int cx = __float2int_rz(pv.at(PX, i));
        cx = cx < 0 ? 0 : (cx >= rx ? rx - 1 : cx);
        atomicAdd(&sh[cx], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < rx; k += blockDim.x)
        if (sh[k]) atomicAdd(&hist[k], (unsigned long long)sh[k]);
}

// keep the particles whose base cell x is in [x0, x1): unordered compaction into the alternate planes
__global__ void __launch_bounds__(256) k_filter_slab(ParticleView src, ParticleView dst, const uint32_t* __restrict__ id_src,
                                                     uint32_t* __restrict__ id_dst, int64_t n, int x0, int x1, uint32_t* counter)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool keep = false;
    if (i < n) {
        const int cx = __float2int_rz(src.at(PX, i));
        keep = cx >= x0 && cx < x1;
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (!m) return;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (!keep) return;
    const uint32_t d = base + (uint32_t)__popc(m & ((1u << lane) - 1u));
#pragma unroll
    for (int k = 0; k < NPLANES; ++k) dst.at(k, d) = src.at(k, i);
    id_dst[d] = id_src[i];
}

static int slab_cuts_host(const int64_t* hist, int rx, int world, int min_width, int* cuts)
{
    if (!hist || !cuts || world < 1 || min_width < 1 || (int64_t)world * min_width > rx) return MPM_ERR_INVALID;
    int64_t total = 0;
    for (int x = 0; x < rx; ++x) total += hist[x];
    cuts[0] = 0;
    int x = 0;
    int64_t cum = 0;  // particles in planes < x
    for (int k = 1; k < world; ++k) {
        const int64_t target = (total * k + world - 1) / world;
        while (x < rx && cum < target) cum += hist[x++];
        int c = x;
        c = std::max(c, cuts[k - 1] + min_width);
        c = std::min(c, rx - (world - k) * min_width);
        cuts[k] = c;
        while (x < c) cum += hist[x++];  // keep (x, cum) consistent if the clamp moved the cut right
    }
    cuts[world] = rx;
    return MPM_OK;
}

constexpr int MIN_SLAB_WIDTH = 4;

bool comm_partitioned(const MpmSolver* s) { return s->comm && s->comm->slab_set; }
void comm_mark_global(MpmSolver* s) { s->comm->pending = true; s->comm->slab_set = false; }

// The planes hold the GLOBAL particle set (s->n particles, ids in orig_id): cut the slabs, keep this rank's.
int comm_finish_migration(MpmSolver* s);

// what mpm_step calls: only the slab partition of a freshly uploaded set (a migration whose counts the host has not read
// yet stays pending: the step's first kernels take the count from the device)
int comm_partition_for_step(MpmSolver* s);

int comm_partition(MpmSolver* s)
{
    CommState* c = s->comm;
    if (!c) return MPM_OK;
    if (c->mig_pending) { int rc = comm_finish_migration(s); if (rc) return rc; }  // every caller but mpm_step needs s->n
    return comm_partition_for_step(s);
}

int comm_partition_for_step(MpmSolver* s)
{
    CommState* c = s->comm;
    if (!c || !c->pending) return MPM_OK;
    if (c->mig_pending) { int rc = comm_finish_migration(s); if (rc) return rc; }
    const int64_t n_global = s->n;
    c->n_global = n_global;
    const int rx = s->dp.Rx;
    // 1. x-plane histogram -> equal-count cuts (identical on every rank: same data, same arithmetic)
    unsigned long long* d_hist = nullptr;
    CKM(cudaMalloc(&d_hist, sizeof(unsigned long long) * rx));
    CKM(cudaMemsetAsync(d_hist, 0, sizeof(unsigned long long) * rx, s->stream));
    if (n_global > 0) {
        const int blocks = (int)std::min<int64_t>((n_global + 255) / 256, 148 * 8);
        k_xhist<ParticleView><<<blocks, 256, sizeof(uint32_t) * rx, s->stream>>>(s->view(), n_global, rx, d_hist);
        s->launches += 1;
    }
    std::vector<int64_t> hist(rx);
    static_assert(sizeof(unsigned long long) == sizeof(int64_t), "");
    cudaError_t e = cudaMemcpyAsync(hist.data(), d_hist, sizeof(int64_t) * rx, cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    cudaFree(d_hist);
    if (e != cudaSuccess) { s->err = cudaGetErrorString(e); return MPM_ERR_CUDA; }
    c->cuts.assign(c->world + 1, 0);
    if (slab_cuts_host(hist.data(), rx, c->world, MIN_SLAB_WIDTH, c->cuts.data())) {
        s->err = "grid too narrow in x for this many ranks (each slab needs >= 4 planes)";
        return MPM_ERR_INVALID;
    }
    c->x0 = c->cuts[c->rank]; c->x1 = c->cuts[c->rank + 1];
    // 2. local grid: planes [x0-1, x1+1)
    s->dp.gx0 = c->x0 - 1; s->dp.nxl = c->x1 - c->x0 + 2;
    cudaFree(s->grid); s->grid = nullptr;
    s->ncells = (int64_t)s->dp.nxl * s->dp.Ry * s->dp.Rz;
    CKM(cudaMalloc(&s->grid, 16 * s->ncells));
    CKM(cudaMemsetAsync(s->grid, 0, 16 * s->ncells, s->stream));
    free_slab_buffers(c);
    c->halo_cells = 2 * (int64_t)s->dp.Ry * s->dp.Rz;
    for (int k = 0; k < 2; ++k) {
        CKM(cudaMalloc(&c->halo_recv[k], 16 * c->halo_cells));
        CKM(cudaMalloc(&c->halo_send[k], 16 * c->halo_cells));
        CKM(cudaMalloc(&c->halo_snap[k], 16 * c->halo_cells));
    }
    if (s->path == MPM_PATH_TILED) {  // block grid follows the slab
        sort_destroy(s);
        int rc = sort_create(s);
        if (rc) return rc;
    } else if (s->path == MPM_PATH_CELL) {
        bin_destroy(s);
        int rc = bin_create(s);
        if (rc) return rc;
    }
    // 3. keep own particles
    CKM(cudaMemsetAsync(c->d_cnt, 0, 9 * sizeof(uint32_t), s->stream));  // (word 9 is the sticky halo-timeout flag)
    if (n_global > 0) {
        k_filter_slab<<<(unsigned)((n_global + 255) / 256), 256, 0, s->stream>>>(s->view(), s->view_alt(), s->orig_id, s->orig_id_alt,
                                                                                 n_global, c->x0, c->x1, c->d_cnt);
        s->launches += 1;
    }
    CKM(cudaMemcpyAsync(c->h_cnt, c->d_cnt, sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
    CKM(cudaStreamSynchronize(s->stream));
    std::swap(s->part, s->part_alt);
    std::swap(s->orig_id, s->orig_id_alt);
    s->n = c->h_cnt[0];
    c->slab_set = true;
    c->pending = false;
    s->sorted_valid = false;
    s->positions_valid = false;
    return MPM_OK;
}

// ================================================================ halo exchange
// block += received; snap (optional) keeps the completed block for the increment of the second exchange
__global__ void __launch_bounds__(256) k_halo_add(int4* __restrict__ blockL, const int4* __restrict__ recvL, int4* __restrict__ snapL,
                                                  int4* __restrict__ blockR, const int4* __restrict__ recvR, int4* __restrict__ snapR,
                                                  int64_t cells)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cells) return;
    if (blockL) {
        int4 a = blockL[i]; const int4 b = recvL[i];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        blockL[i] = a;
        if (snapL) snapL[i] = a;
    }
    if (blockR) {
        int4 a = blockR[i]; const int4 b = recvR[i];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        blockR[i] = a;
        if (snapR) snapR[i] = a;
    }
}

__global__ void __launch_bounds__(256) k_halo_diff(const int4* __restrict__ blockL, const int4* __restrict__ snapL, int4* __restrict__ outL,
                                                   const int4* __restrict__ blockR, const int4* __restrict__ snapR, int4* __restrict__ outR,
                                                   int64_t cells)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cells) return;
    if (blockL) { const int4 a = blockL[i], b = snapL[i]; outL[i] = make_int4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
    if (blockR) { const int4 a = blockR[i], b = snapR[i]; outR[i] = make_int4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
}

// ---------------------------------------------------------------- halo planes by direct peer stores (one process per GPU)
// Each rank owns one allocation with four receive regions ([side][pass], 2 planes each) and four flags; the neighbours
// map it through CUDA IPC.  A halo exchange is then two kernels and no library call:
//   k_halo_push      writes this rank's two overlap planes (pass 1: their increment since pass 0) straight into the
//                    neighbours' receive regions over NVLink; the last block to finish publishes the message number in
//                    the neighbours' flags (system-scope release).
//   k_halo_wait_add  waits for the neighbours' flags to reach the message number, then adds the received planes.
// No acknowledgement is needed: a region is rewritten one step later, and between two pushes of the same pass this rank
// has waited for a message that the neighbour only sent after consuming the previous one (push0, wait0, push1, wait1
// alternate on both sides).  A wait gives up after ~10 s and raises an error flag instead of hanging the GPU.
constexpr int P2P_FLAG_STRIDE = 32;  // uint32 words between flags (one 128-B line each)

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256) k_halo_push(const int4* __restrict__ blockL, const int4* __restrict__ snapL, int4* __restrict__ dstL,
                                                   uint32_t* flagL, const int4* __restrict__ blockR, const int4* __restrict__ snapR,
                                                   int4* __restrict__ dstR, uint32_t* flagR, int64_t cells, uint32_t seq, uint32_t* done)
{
    pdl_prologue();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cells) {
        if (dstL) {
            int4 a = blockL[i];
            if (snapL) { const int4 b = snapL[i]; a.x -= b.x; a.y -= b.y; a.z -= b.z; a.w -= b.w; }
            dstL[i] = a;
        }
        if (dstR) {
            int4 a = blockR[i];
            if (snapR) { const int4 b = snapR[i]; a.x -= b.x; a.y -= b.y; a.z -= b.z; a.w -= b.w; }
            dstR[i] = a;
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t prev = atomicAdd(done, 1u);
        if (prev == gridDim.x - 1) {  // every block's stores are fenced: publish
            *done = 0;
            __threadfence_system();
            if (flagL) st_release_sys(flagL, seq);
            if (flagR) st_release_sys(flagR, seq);
        }
    }
}

__global__ void __launch_bounds__(256) k_halo_wait_add(int4* __restrict__ blockL, const int4* recvL, int4* __restrict__ snapL, const uint32_t* flagL,
                                                       int4* __restrict__ blockR, const int4* recvR, int4* __restrict__ snapR, const uint32_t* flagR,
                                                       int64_t cells, uint32_t seq, uint32_t* err)
{
    pdl_prologue();
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        for (int side = 0; side < 2; ++side) {
            const uint32_t* f = side ? flagR : flagL;
            if (!f) continue;
            while ((int32_t)(ld_acquire_sys(f) - seq) < 0) {
                if (clock64() - t0 > 20000000000ll) { atomicExch(err, 1u); break; }  // ~10 s
                __nanosleep(200);
            }
        }
    }
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cells) return;
    if (blockL) {
        int4 a = blockL[i]; const int4 b = __ldcv(recvL + i);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        blockL[i] = a;
        if (snapL) snapL[i] = a;
    }
    if (blockR) {
        int4 a = blockR[i]; const int4 b = __ldcv(recvR + i);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        blockR[i] = a;
        if (snapR) snapR[i] = a;
    }
}

// Layout of a rank's receive allocation: 4 halo regions [side][pass] | 6 flags (4 halo, 2 migration) | 2 migration regions
// [side]: MIG_HDR header words ([0] = particles in the message) + 17 planes (16 fields + original index) of rec_cap words.
constexpr int MIG_HDR = 32;
static size_t p2p_flags_offset(const CommState* c) { return 4 * 16 * (size_t)c->halo_cells; }
static size_t p2p_mig_offset(const CommState* c) { return (p2p_flags_offset(c) + 6 * P2P_FLAG_STRIDE * sizeof(uint32_t) + 255) / 256 * 256; }
static size_t p2p_mig_bytes(const CommState* c) { return (sizeof(uint32_t) * ((size_t)MIG_HDR + (size_t)REC_WORDS * c->rec_cap) + 255) / 256 * 256; }
static size_t p2p_total_bytes(const CommState* c) { return p2p_mig_offset(c) + 2 * p2p_mig_bytes(c); }
static uint32_t* p2p_mig_region(const CommState* c, uint8_t* base, int side) { return reinterpret_cast<uint32_t*>(base + p2p_mig_offset(c) + side * p2p_mig_bytes(c)); }
static uint32_t* p2p_mig_flag(const CommState* c, uint8_t* base, int side) { return reinterpret_cast<uint32_t*>(base + p2p_flags_offset(c)) + (4 + side) * P2P_FLAG_STRIDE; }

// one-time set-up: allocate, exchange IPC handles with both neighbours (through the transport), map
static void p2p_setup(MpmSolver* s)
{
    CommState* c = s->comm;
    c->p2p_tried = true;
    if (getenv("MPM_NO_P2P")) return;
    const bool hasL = c->rank > 0, hasR = c->rank < c->world - 1;
    const size_t region = 16 * (size_t)c->halo_cells, total = p2p_total_bytes(c);
    if (!c->tr->separate_gpus()) {
        // ranks of one process: plain device pointers instead of IPC handles (peer access enabled when the devices differ)
        bool ok = cudaMalloc(&c->p2p_own, total) == cudaSuccess && cudaMemset(c->p2p_own, 0, p2p_mig_offset(c)) == cudaSuccess &&
                  cudaMalloc(&c->p2p_done, 8 * sizeof(uint32_t)) == cudaSuccess && cudaMemset(c->p2p_done, 0, 8 * sizeof(uint32_t)) == cudaSuccess;
        void *pl = nullptr, *pr = nullptr;
        int dl = s->device, dr = s->device;
        ok = c->tr->share_pointers(ok ? c->p2p_own : nullptr, s->device, &pl, &pr, &dl, &dr) && ok;
        for (int d : {dl, dr})
            if (ok && d != s->device) {
                const cudaError_t e = cudaDeviceEnablePeerAccess(d, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = false;
                cudaGetLastError();
            }
        // (a rank whose peer access failed cannot tell the others any more: the halo kernels' 10 s time-out reports it)
        c->p2p_peer[0] = (uint8_t*)pl; c->p2p_peer[1] = (uint8_t*)pr;
        c->p2p_inproc = true;
        c->p2p_ready = ok;
        return;
    }
    uint8_t *hs = nullptr, *hr = nullptr;  // device staging: [0..63] handle for/from the left, [64..127] right
    bool ok = cudaMalloc(&c->p2p_own, total) == cudaSuccess && cudaMemset(c->p2p_own, 0, p2p_mig_offset(c)) == cudaSuccess &&
              cudaMalloc(&c->p2p_done, 8 * sizeof(uint32_t)) == cudaSuccess && cudaMemset(c->p2p_done, 0, 8 * sizeof(uint32_t)) == cudaSuccess &&
              cudaMalloc(&hs, 128) == cudaSuccess && cudaMalloc(&hr, 128) == cudaSuccess;
    cudaIpcMemHandle_t mine, theirs[2];
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    ok = ok && cudaIpcGetMemHandle(&mine, c->p2p_own) == cudaSuccess;
    // every rank takes part in the exchange even if its own set-up failed (an all-zero handle says so)
    if (!ok) memset(&mine, 0, sizeof(mine));
    if (hs && hr) {
        cudaMemcpyAsync(hs, &mine, 64, cudaMemcpyHostToDevice, s->stream);
        cudaMemcpyAsync(hs + 64, &mine, 64, cudaMemcpyHostToDevice, s->stream);
        cudaMemsetAsync(hr, 0, 128, s->stream);
        std::string err;
        if (c->tr->exchange(hs, hasL ? 64 : 0, hr, hasL ? 64 : 0, hs + 64, hasR ? 64 : 0, hr + 64, hasR ? 64 : 0, s->stream, err)) ok = false;
        cudaMemcpyAsync(theirs, hr, 128, cudaMemcpyDeviceToHost, s->stream);
        cudaStreamSynchronize(s->stream);
    } else {
        ok = false;
    }
    cudaFree(hs); cudaFree(hr);
    const cudaIpcMemHandle_t zero = {};
    for (int side = 0; side < 2 && ok; ++side) {
        if (!(side ? hasR : hasL)) continue;
        if (memcmp(&theirs[side], &zero, 64) == 0) { ok = false; break; }
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, theirs[side], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
        c->p2p_peer[side] = (uint8_t*)p;
    }
    // every rank must take the same path: global AND of `ok` over the chain (world - 1 rounds of neighbour exchange)
    {
        uint32_t* d = nullptr;  // [0] mine, [1] from the left, [2] from the right
        uint32_t h[3] = {ok ? 1u : 0u, 1u, 1u};
        if (cudaMalloc(&d, 3 * sizeof(uint32_t)) == cudaSuccess) {
            for (int round = 1; round < c->world; ++round) {
                cudaMemcpyAsync(d, h, sizeof(h), cudaMemcpyHostToDevice, s->stream);
                std::string err;
                if (c->tr->exchange(d, hasL ? 4 : 0, d + 1, hasL ? 4 : 0, d, hasR ? 4 : 0, d + 2, hasR ? 4 : 0, s->stream, err)) h[0] = 0;
                uint32_t g[3] = {0, 1, 1};
                cudaMemcpyAsync(g, d, sizeof(g), cudaMemcpyDeviceToHost, s->stream);
                cudaStreamSynchronize(s->stream);
                h[0] = (h[0] && (!hasL || g[1]) && (!hasR || g[2])) ? 1u : 0u;
            }
            cudaFree(d);
        } else {
            h[0] = 0;  // (cannot even take part: the neighbours will time out in exchange() -- out of memory is fatal anyway)
        }
        ok = h[0] != 0;
    }
    if (!ok) {
        for (int k = 0; k < 2; ++k) if (c->p2p_peer[k]) { cudaIpcCloseMemHandle(c->p2p_peer[k]); c->p2p_peer[k] = nullptr; }
        fprintf(stderr, "[mpm_b200 rank %d] direct peer halo stores unavailable (CUDA IPC); set MPM_NO_P2P=1 on every rank to use NCCL for the halos\n", c->rank);
    }
    c->p2p_ready = ok;
}

static int exchange_halo_p2p(MpmSolver* s, int pass, int4* blockL, int4* blockR)
{
    CommState* c = s->comm;
    const size_t region = 16 * (size_t)c->halo_cells;
    const unsigned nb = (unsigned)((c->halo_cells + 255) / 256);
    const uint32_t seq = ++c->p2p_seq[pass];
    auto region_of = [&](uint8_t* base, int side) { return reinterpret_cast<int4*>(base + (size_t)(side * 2 + pass) * region); };
    auto flag_of = [&](uint8_t* base, int side) { return reinterpret_cast<uint32_t*>(base + 4 * region) + (side * 2 + pass) * P2P_FLAG_STRIDE; };
    // my left neighbour receives from its RIGHT side (1), my right neighbour from its LEFT side (0)
    int4* dstL = blockL ? region_of(c->p2p_peer[0], 1) : nullptr;
    int4* dstR = blockR ? region_of(c->p2p_peer[1], 0) : nullptr;
    uint32_t* fL = blockL ? flag_of(c->p2p_peer[0], 1) : nullptr;
    uint32_t* fR = blockR ? flag_of(c->p2p_peer[1], 0) : nullptr;
    launch_pdl<PDL_HALO>(k_halo_push, dim3(nb), dim3(256), 0, s->stream, blockL, pass == 1 ? c->halo_snap[0] : nullptr, dstL, fL, blockR, pass == 1 ? c->halo_snap[1] : nullptr,
                                           dstR, fR, c->halo_cells, seq, c->p2p_done + pass);
    if (c->p2p_inproc) {  // ranks of one process: the streams are ordered by events, the flags are already set when the wait kernel runs
        int rc;
        if (blockL && (rc = c->tr->notify(0, pass, seq, s->stream, s->err))) return rc;
        if (blockR && (rc = c->tr->notify(1, pass, seq, s->stream, s->err))) return rc;
        if (blockL && (rc = c->tr->await(0, pass, seq, s->stream, s->err))) return rc;
        if (blockR && (rc = c->tr->await(1, pass, seq, s->stream, s->err))) return rc;
    }
    launch_pdl<PDL_HALO>(k_halo_wait_add, dim3(nb), dim3(256), 0, s->stream, blockL, region_of(c->p2p_own, 0), pass == 0 ? c->halo_snap[0] : nullptr,
                                               blockL ? flag_of(c->p2p_own, 0) : nullptr, blockR, region_of(c->p2p_own, 1),
                                               pass == 0 ? c->halo_snap[1] : nullptr, blockR ? flag_of(c->p2p_own, 1) : nullptr, c->halo_cells,
                                               seq, c->d_cnt + 9);
    s->launches += 2;
    return MPM_OK;
}

int comm_exchange_halo(MpmSolver* s, int pass)
{
    CommState* c = s->comm;
    if (!c->slab_set) { s->err = "multi-GPU: upload the particle set after mpm_comm_init*"; return MPM_ERR_STATE; }
    { int rc = comm_finish_migration(s); if (rc) return rc; }  // the binning and P2G_1 of this step are enqueued: now the host catches up
    const bool hasL = c->rank > 0, hasR = c->rank < c->world - 1;
    if (!hasL && !hasR) return MPM_OK;
    int4* grid = reinterpret_cast<int4*>(s->grid);
    const int64_t plane = (int64_t)s->dp.Ry * s->dp.Rz;
    int4* blockL = hasL ? grid : nullptr;
    int4* blockR = hasR ? grid + (int64_t)(s->dp.nxl - 2) * plane : nullptr;
    if (!c->p2p_tried) p2p_setup(s);
    if (c->p2p_ready) return exchange_halo_p2p(s, pass, blockL, blockR);
    const size_t bytes = 16 * (size_t)c->halo_cells;
    const unsigned nb = (unsigned)((c->halo_cells + 255) / 256);
    const int4 *sendL = blockL, *sendR = blockR;
    if (pass == 1) {
        k_halo_diff<<<nb, 256, 0, s->stream>>>(blockL, c->halo_snap[0], c->halo_send[0], blockR, c->halo_snap[1], c->halo_send[1], c->halo_cells);
        s->launches += 1;
        sendL = c->halo_send[0]; sendR = c->halo_send[1];
    }
    int rc = c->tr->exchange(sendL, hasL ? bytes : 0, c->halo_recv[0], hasL ? bytes : 0, sendR, hasR ? bytes : 0, c->halo_recv[1],
                             hasR ? bytes : 0, s->stream, s->err);
    if (rc) return rc;
    k_halo_add<<<nb, 256, 0, s->stream>>>(blockL, c->halo_recv[0], pass == 0 ? c->halo_snap[0] : nullptr, blockR, c->halo_recv[1],
                                          pass == 0 ? c->halo_snap[1] : nullptr, c->halo_cells);
    s->launches += 1;
    return MPM_OK;
}

// ================================================================ particle migration
struct MigGeom {
    int x0, x1;      // owned planes
    int xl0, xr1;    // the left neighbour's first plane, the right neighbour's end plane (jump guard)
};

// d_cnt words: 0 nL, 1 nR (particles leaving left / right), 4 holes, 5 fillers, 8 "crossed more than one slab" flag
__device__ __forceinline__ int mig_side(const MigGeom& g, float px)
{
    const int cx = __float2int_rz(px);
    return cx < g.x0 ? 0 : (cx >= g.x1 ? 1 : -1);
}

// generic classification pass (kernel paths whose G2P does not classify): lists of the leaving slots + counts
template <class View>
__global__ void __launch_bounds__(256) k_mig_scan(MigGeom g, View pv, int64_t n, uint32_t rec_cap, uint32_t* __restrict__ leaveL,
                                                  uint32_t* __restrict__ leaveR, uint32_t* __restrict__ cnt)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float px = pv.at(PX, i);
    const int side = mig_side(g, px);
    if (side < 0) return;
    // a particle that would land beyond the neighbouring slab is held back in that slab's far plane for this step
    const int cx = __float2int_rz(px);
    if (cx < g.xl0) { pv.at(PX, i) = (float)g.xl0 + 0.5f; atomicAdd(cnt + 8, 1u); }
    else if (cx >= g.xr1) { pv.at(PX, i) = (float)g.xr1 - 0.5f; atomicAdd(cnt + 8, 1u); }
    const uint32_t slot = atomicAdd(cnt + side, 1u);
    if (slot < rec_cap) (side ? leaveR : leaveL)[slot] = (uint32_t)i;
}

// Migration message: MIG_HDR header words ([0] = number of particles leaving over this edge), then the first `cap`
// records as SoA with stride cap (word MIG_HDR + k * cap + slot), then any overflow records as 17-word AoS.  `cap` is
// agreed by both ends without talking: it is a function of the count that crossed this edge in the previous step, which
// sender and receiver both know.  So a step needs ONE exchange and ONE host sync; only when the count grows by more than
// the head-room from one step to the next does a second (overflow) exchange follow.

static inline uint32_t mig_capacity(uint32_t prev)  // the same on both ends of an edge: depends on nothing rank-local
{
    const uint64_t want = (uint64_t)prev + prev / 4 + 4096ull;
    return (uint32_t)std::min<uint64_t>((want + 4095ull) & ~4095ull, 1u << 30);
}

// leavers (from the lists) -> messages; leavers below n_stay leave holes.  Everything is sized by device-side counts.
template <class View>
__global__ void __launch_bounds__(256) k_mig_pack(View pv, const uint32_t* __restrict__ ids, int64_t n, uint32_t capL, uint32_t capR,
                                                  uint32_t rec_cap, const uint32_t* __restrict__ leaveL, const uint32_t* __restrict__ leaveR,
                                                  uint32_t* __restrict__ sendL, uint32_t* __restrict__ sendR, uint32_t* __restrict__ holes,
                                                  uint32_t* __restrict__ fillers, MigGeom g, uint32_t* __restrict__ cnt)
{
    const uint32_t nL = min(cnt[0], rec_cap), nR = min(cnt[1], rec_cap);
    const int64_t n_stay = n - cnt[0] - cnt[1];
    // stayers in the tail [n_stay, n) are the fillers of the holes the leavers below n_stay leave
    {
        const uint32_t nt = cnt[0] + cnt[1];
        for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += gridDim.x * blockDim.x) {
            const int64_t i = n_stay + t;
            if (mig_side(g, pv.at(PX, i)) < 0) fillers[atomicAdd(cnt + 5, 1u)] = (uint32_t)i;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { sendL[0] = cnt[0]; sendR[0] = cnt[1]; }
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < nL + nR; j += gridDim.x * blockDim.x) {
        const int side = j >= nL;
        const uint32_t slot = side ? j - nL : j;
        const uint32_t i = (side ? leaveR : leaveL)[slot];
        if ((int64_t)i < n_stay) holes[atomicAdd(cnt + 4, 1u)] = i;
        uint32_t* out = side ? sendR : sendL;
        const uint32_t cap = side ? capR : capL;
        if (slot < cap) {
#pragma unroll
            for (int k = 0; k < NPLANES; ++k) out[MIG_HDR + (size_t)k * cap + slot] = __float_as_uint(pv.at(k, i));
            out[MIG_HDR + (size_t)NPLANES * cap + slot] = ids[i];
        } else {
            uint32_t* o = out + MIG_HDR + (size_t)REC_WORDS * cap + (size_t)REC_WORDS * (slot - cap);
#pragma unroll
            for (int k = 0; k < NPLANES; ++k) o[k] = __float_as_uint(pv.at(k, i));
            o[NPLANES] = ids[i];
        }
    }
}

template <class View>
__global__ void __launch_bounds__(256) k_mig_fill(View pv, uint32_t* __restrict__ ids, uint32_t* __restrict__ keys,
                                                  const uint32_t* __restrict__ holes, const uint32_t* __restrict__ fillers,
                                                  const uint32_t* __restrict__ cnt)
{
    pdl_prologue();
    const uint32_t nh = cnt[4];
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < nh; j += gridDim.x * blockDim.x) {
        const uint32_t dst = holes[j], src = fillers[j];
#pragma unroll
        for (int k = 0; k < NPLANES; ++k) pv.at(k, dst) = pv.at(k, src);
        ids[dst] = ids[src];
        if (keys) keys[dst] = keys[src];  // next-step bin keys written by G2P travel with the particle
    }
}

// arrivals of one message appended behind the stayers: left arrivals first, then right arrivals.
// overflow = false: the SoA part (first min(count, cap) records); true: the AoS overflow records (count - cap).
template <class View>
__global__ void __launch_bounds__(256) k_mig_unpack(View pv, uint32_t* __restrict__ ids, int64_t n, uint32_t* __restrict__ cnt,
                                                    const uint32_t* __restrict__ msgL, const uint32_t* __restrict__ msgR, uint32_t capL,
                                                    uint32_t capR, int only_side, bool overflow)
{
    // blockIdx.y = side (0: from the left neighbour, 1: from the right one); a missing neighbour has a null message
    const int side = only_side >= 0 ? only_side : (int)blockIdx.y;
    const uint32_t* msg = side ? msgR : msgL;
    if (!msg) return;
    const uint32_t cap = side ? capR : capL;
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t total = msg[0];
    if (j == 0 && !overflow) cnt[2 + side] = total;  // arrivals from this side, for the host
    const int64_t base = n - cnt[0] - cnt[1] + (side == 1 && msgL ? (int64_t)msgL[0] : 0);
    if (!overflow) {
        if (j >= min(total, cap)) return;
#pragma unroll
        for (int k = 0; k < NPLANES; ++k) pv.at(k, base + j) = __uint_as_float(msg[MIG_HDR + (size_t)k * cap + j]);
        ids[base + j] = msg[MIG_HDR + (size_t)NPLANES * cap + j];
    } else {
        if (total <= cap || j >= total - cap) return;
        const uint32_t* o = msg + MIG_HDR + (size_t)REC_WORDS * cap + (size_t)REC_WORDS * j;
#pragma unroll
        for (int k = 0; k < NPLANES; ++k) pv.at(k, base + cap + j) = __uint_as_float(o[k]);
        ids[base + cap + j] = o[NPLANES];
    }
}

// ---------------------------------------------------------------- migration by peer stores (cell path)
// The leavers go straight into the neighbour's receive region -- 17 planes of rec_cap words, so a warp's stores are full
// lines over NVLink and no message size has to be agreed on (the NCCL path sizes its messages from the previous step's
// count and needs a second round when a burst exceeds the head-room) -- and the last block publishes count + flag.
__global__ void __launch_bounds__(256) k_mig_push(RecView pv, const uint32_t* __restrict__ ids, const uint32_t* __restrict__ n_dev, uint32_t rec_cap,
                                                  const uint32_t* __restrict__ leaveL, const uint32_t* __restrict__ leaveR, uint32_t* dstL, uint32_t* dstR,
                                                  uint32_t* flagL, uint32_t* flagR, uint32_t* __restrict__ holes, uint32_t* __restrict__ fillers, MigGeom g,
                                                  uint32_t* __restrict__ cnt, uint32_t seq, uint32_t* done)
{
    pdl_prologue();
    const int64_t n = *n_dev;
    const uint32_t nL = min(cnt[0], rec_cap), nR = min(cnt[1], rec_cap);
    const int64_t n_stay = n - cnt[0] - cnt[1];
    {   // stayers in the tail [n_stay, n) are the fillers of the holes the leavers below n_stay leave
        const uint32_t nt = cnt[0] + cnt[1];
        for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += gridDim.x * blockDim.x) {
            const int64_t i = n_stay + t;
            if (mig_side(g, pv.at(PX, i)) < 0) fillers[atomicAdd(cnt + 5, 1u)] = (uint32_t)i;
        }
    }
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < nL + nR; j += gridDim.x * blockDim.x) {
        const int side = j >= nL;
        const uint32_t slot = side ? j - nL : j;
        const uint32_t i = (side ? leaveR : leaveL)[slot];
        if ((int64_t)i < n_stay) holes[atomicAdd(cnt + 4, 1u)] = i;
        uint32_t* out = side ? dstR : dstL;
        if (!out) continue;  // (no neighbour on that side: nothing can leave through a wall; the slot is a hole all the same)
#pragma unroll
        for (int k = 0; k < NPLANES; ++k) out[MIG_HDR + (size_t)k * rec_cap + slot] = __float_as_uint(pv.at(k, i));
        out[MIG_HDR + (size_t)NPLANES * rec_cap + slot] = ids[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t prev = atomicAdd(done, 1u);
        if (prev == gridDim.x - 1) {  // every block's stores are fenced: publish count, then flag
            *done = 0;
            if (dstL) dstL[0] = cnt[0];
            if (dstR) dstR[0] = cnt[1];
            __threadfence_system();
            if (flagL) st_release_sys(flagL, seq);
            if (flagR) st_release_sys(flagR, seq);
        }
    }
}

// arrivals appended behind the stayers (left arrivals first), their bin keys and counts for the next binning, and the new
// particle count -- all from device-side numbers; the host reads cnt[] later
__global__ void __launch_bounds__(256) k_mig_pull(RecView pv, uint32_t* __restrict__ ids, uint32_t* n_dev, uint32_t rec_cap, const uint32_t* msgL,
                                                  const uint32_t* msgR, const uint32_t* flagL, const uint32_t* flagR, uint32_t seq, uint32_t* cnt,
                                                  KeyGeom kg, uint32_t nslots, uint32_t* __restrict__ keys, uint32_t* __restrict__ cnt_next, uint32_t* done)
{
    pdl_prologue();
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        for (int side = 0; side < 2; ++side) {
            const uint32_t* f = side ? flagR : flagL;
            if (!f) continue;
            while ((int32_t)(ld_acquire_sys(f) - seq) < 0) {
                if (clock64() - t0 > 20000000000ll) { atomicExch(cnt + 9, 1u); break; }  // ~10 s
                __nanosleep(200);
            }
        }
    }
    __syncthreads();
    const int64_t n = *n_dev;
    const uint32_t mL = msgL ? min(__ldcv(msgL), rec_cap) : 0u, mR = msgR ? min(__ldcv(msgR), rec_cap) : 0u;
    const int64_t n_stay = n - cnt[0] - cnt[1];
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < mL + mR; j += gridDim.x * blockDim.x) {
        const int side = j >= mL;
        const uint32_t* msg = side ? msgR : msgL;
        const uint32_t q = side ? j - mL : j;
        const int64_t dst = n_stay + j;
#pragma unroll
        for (int k = 0; k < NPLANES; ++k) pv.at(k, dst) = __uint_as_float(__ldcv(msg + MIG_HDR + (size_t)k * rec_cap + q));
        ids[dst] = __ldcv(msg + MIG_HDR + (size_t)NPLANES * rec_cap + q);
        if (keys) {  // (as bin_keys_range does for the NCCL path)
            uint32_t key = cell_key(kg, __float2int_rz(pv.at(PX, dst)), __float2int_rz(pv.at(PY, dst)), __float2int_rz(pv.at(PZ, dst)));
            key = key < nslots ? key : nslots - 1;
            keys[dst] = key;
            atomicAdd(&cnt_next[key], 1u);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t prev = atomicAdd(done, 1u);
        if (prev == gridDim.x - 1) {  // every block has read the old count: publish the new one
            *done = 0;
            cnt[2] = msgL ? __ldcv(msgL) : 0u; cnt[3] = msgR ? __ldcv(msgR) : 0u;  // (unclipped: the host checks them against rec_cap)
            *n_dev = (uint32_t)(n_stay + mL + mR);
        }
    }
}

KeyGeom bin_key_geom(const MpmSolver* s);
uint32_t* bin_next_counts(MpmSolver* s);
uint32_t bin_nslots(const MpmSolver* s);

static bool mig_by_peer_stores(const MpmSolver* s)
{
    static const bool off = getenv("MPM_NCCL_MIGRATION") != nullptr;
    const CommState* c = s->comm;
    return !off && c->p2p_ready && s->path == MPM_PATH_CELL && s->in_rec && bin_next_keys(const_cast<MpmSolver*>(s)) != nullptr;
}

// enqueue one migration by peer stores; the host learns the counts in comm_finish_migration
static int migrate_p2p(MpmSolver* s)
{
    CommState* c = s->comm;
    const bool hasL = c->rank > 0, hasR = c->rank < c->world - 1;
    MigGeom g{c->x0, c->x1, hasL ? c->cuts[c->rank - 1] : c->x0, hasR ? c->cuts[c->rank + 2] : c->x1};
    if (!c->n_dev) {
        CKM(cudaMalloc(&c->n_dev, sizeof(uint32_t)));
        CKM(cudaEventCreateWithFlags(&c->mig_done, cudaEventDisableTiming));
    }
    if (!c->mig_pending) {  // (the device count is current unless the host has changed the set since: refresh it)
        const uint32_t n32 = (uint32_t)s->n;
        CKM(cudaMemcpyAsync(c->n_dev, &n32, sizeof(uint32_t), cudaMemcpyHostToDevice, s->stream));
    }
    const uint32_t seq = ++c->mig_seq;
    uint32_t* dstL = hasL ? p2p_mig_region(c, c->p2p_peer[0], 1) : nullptr;  // my left neighbour receives on its RIGHT side
    uint32_t* dstR = hasR ? p2p_mig_region(c, c->p2p_peer[1], 0) : nullptr;
    uint32_t* fL = hasL ? p2p_mig_flag(c, c->p2p_peer[0], 1) : nullptr;
    uint32_t* fR = hasR ? p2p_mig_flag(c, c->p2p_peer[1], 0) : nullptr;
    launch_pdl<PDL_MIG>(k_mig_push, dim3(296), dim3(256), 0, s->stream, s->rview(), s->orig_id, c->n_dev, (uint32_t)c->rec_cap, c->leave[0], c->leave[1], dstL, dstR, fL, fR, c->holes,
                                           c->fillers, g, c->d_cnt, seq, c->p2p_done + 2);
    launch_pdl<PDL_MIG>(k_mig_fill<RecView>, dim3(296), dim3(256), 0, s->stream, s->rview(), s->orig_id, bin_next_keys(s), c->holes, c->fillers, c->d_cnt);
    if (c->p2p_inproc) {  // ranks of one process: order the streams by events (see exchange_halo_p2p)
        int rc;
        if (hasL && (rc = c->tr->notify(0, 2, seq, s->stream, s->err))) return rc;
        if (hasR && (rc = c->tr->notify(1, 2, seq, s->stream, s->err))) return rc;
        if (hasL && (rc = c->tr->await(0, 2, seq, s->stream, s->err))) return rc;
        if (hasR && (rc = c->tr->await(1, 2, seq, s->stream, s->err))) return rc;
    }
    launch_pdl<PDL_MIG>(k_mig_pull, dim3(296), dim3(256), 0, s->stream, s->rview(), s->orig_id, c->n_dev, (uint32_t)c->rec_cap, hasL ? p2p_mig_region(c, c->p2p_own, 0) : nullptr,
                                           hasR ? p2p_mig_region(c, c->p2p_own, 1) : nullptr, hasL ? p2p_mig_flag(c, c->p2p_own, 0) : nullptr,
                                           hasR ? p2p_mig_flag(c, c->p2p_own, 1) : nullptr, seq, c->d_cnt, bin_key_geom(s), bin_nslots(s), bin_next_keys(s),
                                           bin_next_counts(s), c->p2p_done + 3);
    s->launches += 3;
    CKM(cudaMemcpyAsync(c->h_cnt, c->d_cnt, 16 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
    CKM(cudaEventRecord(c->mig_done, s->stream));
    c->classified = false;
    c->mig_pending = true;
    c->mig_n_before = s->n;
    // until the host has read the counts, launches are sized for the most this rank can hold after the exchange
    s->n_launch_extra = std::min<int64_t>(2 * c->rec_cap, s->cap - s->n);
    s->n_dev = c->n_dev;
    s->sorted_valid = false;
    s->positions_valid = false;
    return MPM_OK;
}

// the host side of a migration enqueued by migrate_p2p: counts, checks, bookkeeping.  Called when the next step has its
// binning and P2G_1 enqueued (the GPU is busy, the copy of the counts arrived long ago), or by any call that needs s->n.
int comm_finish_migration(MpmSolver* s)
{
    CommState* c = s->comm;
    if (!c || !c->mig_pending) return MPM_OK;
    c->mig_pending = false;
    s->n_launch_extra = 0;
    s->n_dev = nullptr;
    CKM(cudaEventSynchronize(c->mig_done));
    const bool hasL = c->rank > 0, hasR = c->rank < c->world - 1;
    const uint32_t nL = c->h_cnt[0], nR = c->h_cnt[1], mL = hasL ? c->h_cnt[2] : 0, mR = hasR ? c->h_cnt[3] : 0;
    c->slab_jump_clamps += c->h_cnt[8];
    if (c->h_cnt[9]) { s->err = "multi-GPU: a neighbour's halo planes / migrants did not arrive within 10 s (peer-store path)"; return MPM_ERR_COMM; }
    if ((int64_t)std::max(std::max(nL, nR), std::max(mL, mR)) > c->rec_cap) {
        s->err = "multi-GPU: more particles crossed a slab boundary in one step than the migration buffer holds (raise max_particles)";
        return MPM_ERR_COMM;
    }
    const int64_t n_stay = c->mig_n_before - nL - nR;
    if (n_stay + mL + mR > s->cap) { s->err = "multi-GPU: arriving particles exceed max_particles of this rank"; return MPM_ERR_COMM; }
    s->n = n_stay + mL + mR;
    c->sent_prev[0] = nL; c->sent_prev[1] = nR; c->recv_prev[0] = mL; c->recv_prev[1] = mR;
    c->migrated_out += nL + nR;
    c->migrated_in += mL + mR;
    static const bool trace = getenv("MPM_COMM_TRACE") != nullptr;
    if (trace) fprintf(stderr, "[mig r%d p2p] n=%lld nL=%u nR=%u mL=%u mR=%u\n", c->rank, (long long)s->n, nL, nR, mL, mR);
    return MPM_OK;
}

template <class View>
static int migrate_impl(MpmSolver* s, View pv, bool recut = false)
{
    CommState* c = s->comm;
    const bool hasL = c->rank > 0, hasR = c->rank < c->world - 1;
    const int64_t n = s->n;
    MigGeom g{c->x0, c->x1, hasL ? c->cuts[c->rank - 1] : c->x0, hasR ? c->cuts[c->rank + 2] : c->x1};
    if (recut) { g.xl0 = INT_MIN; g.xr1 = INT_MAX; }  // particles change owner because the cuts moved, not because they did
    const unsigned nb = (unsigned)((n + 255) / 256);
    const uint32_t capS[2] = {mig_capacity(c->sent_prev[0]), mig_capacity(c->sent_prev[1])};
    const uint32_t capR[2] = {mig_capacity(c->recv_prev[0]), mig_capacity(c->recv_prev[1])};
    if ((int64_t)std::max(std::max(capS[0], capS[1]), std::max(capR[0], capR[1])) > c->rec_cap) {
        s->err = "multi-GPU: migration buffer too small for the traffic over a slab boundary (raise max_particles)";
        return MPM_ERR_COMM;
    }
    if (!c->classified) {  // this path's G2P did not classify: one pass over the positions
        CKM(cudaMemsetAsync(c->d_cnt, 0, 9 * sizeof(uint32_t), s->stream));  // (word 9 is the sticky halo-timeout flag)
        if (n > 0) { k_mig_scan<View><<<nb, 256, 0, s->stream>>>(g, pv, n, (uint32_t)c->rec_cap, c->leave[0], c->leave[1], c->d_cnt); s->launches += 1; }
    }
    c->classified = false;
    k_mig_pack<View><<<296, 256, 0, s->stream>>>(pv, s->orig_id, n, capS[0], capS[1], (uint32_t)c->rec_cap, c->leave[0], c->leave[1], c->send_rec[0],
                                                 c->send_rec[1], c->holes, c->fillers, g, c->d_cnt);
    k_mig_fill<View><<<296, 256, 0, s->stream>>>(pv, s->orig_id, bin_next_keys(s), c->holes, c->fillers, c->d_cnt);
    s->launches += 2;
    auto msg_bytes = [](uint32_t cap) { return sizeof(uint32_t) * ((size_t)MIG_HDR + (size_t)REC_WORDS * cap); };
    int rc = c->tr->exchange(c->send_rec[0], hasL ? msg_bytes(capS[0]) : 0, c->recv_rec[0], hasL ? msg_bytes(capR[0]) : 0, c->send_rec[1],
                             hasR ? msg_bytes(capS[1]) : 0, c->recv_rec[1], hasR ? msg_bytes(capR[1]) : 0, s->stream, s->err);
    if (rc) return rc;
    const uint32_t* msgL = hasL ? c->recv_rec[0] : nullptr;
    const uint32_t* msgR = hasR ? c->recv_rec[1] : nullptr;
    {
        const dim3 grid((std::max(capR[0], capR[1]) + 255) / 256, 2);
        k_mig_unpack<View><<<grid, 256, 0, s->stream>>>(pv, s->orig_id, n, c->d_cnt, msgL, msgR, capR[0], capR[1], -1, false);
        s->launches += 1;
    }
    // counts to the host in one copy: [0] nL, [1] nR, [2] mL, [3] mR (written by the unpack), [8] bad
    CKM(cudaMemcpyAsync(c->h_cnt, c->d_cnt, 16 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
    CKM(cudaStreamSynchronize(s->stream));
    const uint32_t nL = c->h_cnt[0], nR = c->h_cnt[1], mL = hasL ? c->h_cnt[2] : 0, mR = hasR ? c->h_cnt[3] : 0;
    c->slab_jump_clamps += c->h_cnt[8];
    if (c->h_cnt[9]) { s->err = "multi-GPU: a neighbour's halo planes did not arrive within 10 s (peer-store path)"; return MPM_ERR_COMM; }
    if ((int64_t)std::max(std::max(nL, nR), std::max(mL, mR)) > c->rec_cap) { s->err = "multi-GPU: migration buffer too small"; return MPM_ERR_COMM; }
    const int64_t n_stay = n - nL - nR;
    if (n_stay + mL + mR > s->cap) { s->err = "multi-GPU: arriving particles exceed max_particles of this rank"; return MPM_ERR_COMM; }
    // overflow round (only when an edge's count more than doubled since the previous step): sizes are known to both ends now
    const size_t rb = sizeof(uint32_t) * REC_WORDS;
    const size_t oSL = nL > capS[0] ? rb * (nL - capS[0]) : 0, oSR = nR > capS[1] ? rb * (nR - capS[1]) : 0;
    const size_t oRL = mL > capR[0] ? rb * (mL - capR[0]) : 0, oRR = mR > capR[1] ? rb * (mR - capR[1]) : 0;
    // every rank must take the same decision: an edge overflows on both of its ends or on neither, but a rank whose edges
    // are both quiet still has nothing to do here, and its neighbours' other edges do not involve it
    if (oSL || oSR || oRL || oRR) {
        rc = c->tr->exchange(c->send_rec[0] + MIG_HDR + (size_t)REC_WORDS * capS[0], oSL, c->recv_rec[0] + MIG_HDR + (size_t)REC_WORDS * capR[0], oRL,
                             c->send_rec[1] + MIG_HDR + (size_t)REC_WORDS * capS[1], oSR, c->recv_rec[1] + MIG_HDR + (size_t)REC_WORDS * capR[1], oRR,
                             s->stream, s->err);
        if (rc) return rc;
        if (oRL) { k_mig_unpack<View><<<(mL - capR[0] + 255) / 256, 256, 0, s->stream>>>(pv, s->orig_id, n, c->d_cnt, msgL, msgR, capR[0], capR[1], 0, true); s->launches += 1; }
        if (oRR) { k_mig_unpack<View><<<(mR - capR[1] + 255) / 256, 256, 0, s->stream>>>(pv, s->orig_id, n, c->d_cnt, msgL, msgR, capR[0], capR[1], 1, true); s->launches += 1; }
        c->overflow_rounds += 1;
    }
    static const bool trace = getenv("MPM_COMM_TRACE") != nullptr;
    if (trace)
        fprintf(stderr, "[mig r%d] n=%lld nL=%u nR=%u mL=%u mR=%u capS=%u,%u capR=%u,%u overflow=%d\n", c->rank, (long long)n, nL, nR, mL, mR,
                capS[0], capS[1], capR[0], capR[1], (int)(oSL || oSR || oRL || oRR));
    c->sent_prev[0] = nL; c->sent_prev[1] = nR; c->recv_prev[0] = mL; c->recv_prev[1] = mR;
    s->n = n_stay + mL + mR;
    rc = bin_keys_range(s, n_stay, (int64_t)mL + mR);  // arrivals: bin keys + counts for the next step
    if (rc) return rc;
    c->migrated_out += nL + nR;
    c->migrated_in += mL + mR;
    if (nL + nR + mL + mR) { s->sorted_valid = false; s->positions_valid = false; }
    return MPM_OK;
}

// The cell-path G2P classifies its particles itself (it has the new position in registers): zero the counters before it
// runs and hand it the geometry and lists.
int comm_begin_classify(MpmSolver* s, MigClassify* out)
{
    CommState* c = s->comm;
    out->cnt = nullptr;
    if (!c || c->world < 2) return MPM_OK;
    const bool hasL = c->rank > 0, hasR = c->rank < c->world - 1;
    CKM(cudaMemsetAsync(c->d_cnt, 0, 9 * sizeof(uint32_t), s->stream));  // (word 9 is the sticky halo-timeout flag)
    out->x0 = c->x0; out->x1 = c->x1;
    out->xl0 = hasL ? c->cuts[c->rank - 1] : c->x0;
    out->xr1 = hasR ? c->cuts[c->rank + 2] : c->x1;
    out->cnt = c->d_cnt; out->leaveL = c->leave[0]; out->leaveR = c->leave[1]; out->rec_cap = (uint32_t)c->rec_cap;
    c->classified = true;
    return MPM_OK;
}

int comm_migrate(MpmSolver* s)
{
    CommState* c = s->comm;
    if (c->world < 2) return MPM_OK;
    { int rc = comm_finish_migration(s); if (rc) return rc; }  // (normally done already, after the step's P2G_1 was enqueued)
    if (mig_by_peer_stores(s)) return migrate_p2p(s);
    // after a cell-path G2P the particle state lives in the 64-byte records: migrate those
    return s->in_rec ? migrate_impl<RecView>(s, s->rview()) : migrate_impl<ParticleView>(s, s->view());
}

// ================================================================ re-cutting the slabs
__global__ void __launch_bounds__(256) k_hist_round(const unsigned long long* __restrict__ own, const unsigned long long* __restrict__ accL,
                                                    const unsigned long long* __restrict__ accR, unsigned long long* __restrict__ toR,
                                                    unsigned long long* __restrict__ toL, int rx)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= rx) return;
    toR[x] = accL[x] + own[x];  // everything at or left of this rank
    toL[x] = accR[x] + own[x];  // everything at or right of this rank
}

__global__ void __launch_bounds__(256) k_hist_scale(unsigned long long* __restrict__ own, int rx, unsigned long long w)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x < rx) own[x] *= w;
}

// Global x-plane histogram (world - 1 rounds of neighbour exchange: each round passes "all ranks on my left, plus me" to
// the right and the mirror image to the left), new cuts moved at most max_shift planes from the current ones, one
// migration round, local grid and binning re-created.  Collective over all ranks; call between steps.
// wq = 1 on every rank: equal particle counts.  Otherwise every rank's planes count wq times (its cost per particle in
// 1/4096 of the caller's unit): the cuts equalise the summed COST -- what a host does that has measured that its ranks
// need different times for the same number of particles (denser cells, more pile-ups, a wider empty grid to scan).
static int rebalance(MpmSolver* s, int max_shift, unsigned long long wq)
{
    CommState* c = s->comm;
    { int rc = comm_finish_migration(s); if (rc) return rc; }
    const int rx = s->dp.Rx, world = c->world;
    const bool hasL = c->rank > 0, hasR = c->rank < world - 1;
    max_shift = std::max(1, std::min(max_shift, MIN_SLAB_WIDTH - 1));  // a particle never has to travel further than the next rank
    unsigned long long* d = nullptr;  // own | accL | accR | toR | toL
    CKM(cudaMalloc(&d, sizeof(unsigned long long) * 5 * rx));
    CKM(cudaMemsetAsync(d, 0, sizeof(unsigned long long) * 5 * rx, s->stream));
    unsigned long long *own = d, *accL = d + rx, *accR = d + 2 * rx, *toR = d + 3 * rx, *toL = d + 4 * rx;
    if (s->n > 0) {
        const int blocks = (int)std::min<int64_t>((s->n + 255) / 256, 148 * 8);
        if (s->in_rec) k_xhist<RecView><<<blocks, 256, sizeof(uint32_t) * rx, s->stream>>>(s->rview(), s->n, rx, own);
        else k_xhist<ParticleView><<<blocks, 256, sizeof(uint32_t) * rx, s->stream>>>(s->view(), s->n, rx, own);
        s->launches += 1;
    }
    if (wq != 1ull) { k_hist_scale<<<(rx + 255) / 256, 256, 0, s->stream>>>(own, rx, wq); s->launches += 1; }
    const size_t hb = sizeof(unsigned long long) * rx;
    int rc = MPM_OK;
    for (int round = 1; round < world && rc == MPM_OK; ++round) {
        k_hist_round<<<(rx + 255) / 256, 256, 0, s->stream>>>(own, accL, accR, toR, toL, rx);
        s->launches += 1;
        rc = c->tr->exchange(toL, hasL ? hb : 0, accL, hasL ? hb : 0, toR, hasR ? hb : 0, accR, hasR ? hb : 0, s->stream, s->err);
    }
    std::vector<int64_t> hist(3 * rx);
    cudaError_t e = cudaMemcpyAsync(hist.data(), d, 3 * hb, cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    cudaFree(d);
    if (rc) return rc;
    if (e != cudaSuccess) { s->err = cudaGetErrorString(e); return MPM_ERR_CUDA; }
    for (int x = 0; x < rx; ++x) hist[x] += hist[rx + x] + hist[2 * rx + x];
    std::vector<int> target(world + 1), cuts(c->cuts);
    if (slab_cuts_host(hist.data(), rx, world, MIN_SLAB_WIDTH, target.data())) return MPM_OK;
    bool moved = false, valid = true;
    for (int k = 1; k < world; ++k) {
        cuts[k] = std::max(c->cuts[k] - max_shift, std::min(c->cuts[k] + max_shift, target[k]));
        moved |= cuts[k] != c->cuts[k];
    }
    for (int k = 0; k < world; ++k) valid &= cuts[k + 1] - cuts[k] >= MIN_SLAB_WIDTH;
    if (!valid) moved = false;  // two cuts closing in on each other: leave everything as it is this time
    if (!moved) return MPM_OK;  // (the same decision on every rank: same histogram, same arithmetic)
    c->cuts = cuts;
    c->x0 = cuts[c->rank]; c->x1 = cuts[c->rank + 1];
    c->classified = false;
    rc = s->in_rec ? migrate_impl<RecView>(s, s->rview(), true) : migrate_impl<ParticleView>(s, s->view(), true);
    if (rc) return rc;
    // local grid and binning follow the slab
    s->dp.gx0 = c->x0 - 1; s->dp.nxl = c->x1 - c->x0 + 2;
    CKM(cudaStreamSynchronize(s->stream));
    cudaFree(s->grid); s->grid = nullptr;
    s->ncells = (int64_t)s->dp.nxl * s->dp.Ry * s->dp.Rz;
    CKM(cudaMalloc(&s->grid, 16 * s->ncells));
    CKM(cudaMemsetAsync(s->grid, 0, 16 * s->ncells, s->stream));
    s->grid_raw = false;
    if (s->path == MPM_PATH_TILED) { sort_destroy(s); rc = sort_create(s); }
    else if (s->path == MPM_PATH_CELL) { bin_destroy(s); rc = bin_create(s); }
    if (rc) return rc;
    s->sorted_valid = false;
    s->positions_valid = false;
    c->recuts += 1;
    return MPM_OK;
}

}  // namespace mpm

// ================================================================ C ABI
using namespace mpm;

extern "C" int32_t mpm_slab_cuts(const int64_t* hist, int32_t rx, int32_t world, int32_t min_width, int32_t* cuts)
{
    return slab_cuts_host(hist, rx, world, min_width, cuts);
}

extern "C" int32_t mpm_comm_unique_id(uint8_t id[MPM_COMM_ID_BYTES])
{
    static_assert(sizeof(ncclUniqueId) == MPM_COMM_ID_BYTES, "ncclUniqueId size");
    if (!id) return MPM_ERR_INVALID;
    std::string err;
    NcclApi* api = nccl_api(err);
    if (!api) return MPM_ERR_COMM;
    ncclUniqueId u;
    if (api->GetUniqueId(&u) != ncclSuccess) return MPM_ERR_COMM;
    memcpy(id, &u, sizeof(u));
    return MPM_OK;
}

extern "C" int32_t mpm_comm_init(MpmSolver* s, const uint8_t id[MPM_COMM_ID_BYTES], int32_t rank, int32_t world)
{
    if (!s || !id || world < 1 || rank < 0 || rank >= world) return MPM_ERR_INVALID;
    if (cudaSetDevice(s->device) != cudaSuccess) { s->err = "cudaSetDevice failed"; return MPM_ERR_CUDA; }
    NcclApi* api = nccl_api(s->err);
    if (!api) return MPM_ERR_COMM;
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    NcclTransport* tr = new NcclTransport();
    tr->api = api;
    ncclResult_t r = api->CommInitRank(&tr->comm, world, u, rank);
    if (r != ncclSuccess) {
        s->err = std::string("ncclCommInitRank: ") + api->GetErrorString(r);
        tr->comm = nullptr;
        delete tr;
        return MPM_ERR_COMM;
    }
    return comm_attach(s, tr, rank, world);
}

extern "C" int32_t mpm_local_hub_create(int32_t world, MpmLocalHub** hub)
{
    if (!hub || world < 1 || world > 64) return MPM_ERR_INVALID;
    MpmLocalHub* h = new MpmLocalHub();
    h->world = world;
    h->to_right = std::vector<Mailbox>(world);
    h->to_left = std::vector<Mailbox>(world);
    h->sig = std::vector<MpmLocalHub::Signal>((size_t)world * 6);
    *hub = h;
    return MPM_OK;
}

extern "C" int32_t mpm_local_hub_destroy(MpmLocalHub* hub)
{
    if (!hub) return MPM_OK;
    for (auto* v : {&hub->to_right, &hub->to_left})
        for (Mailbox& mb : *v) {
            if (mb.ready) cudaEventDestroy(mb.ready);
            if (mb.done) cudaEventDestroy(mb.done);
        }
    for (auto& sg : hub->sig) if (sg.ev) cudaEventDestroy(sg.ev);
    delete hub;
    return MPM_OK;
}

extern "C" int32_t mpm_comm_init_local(MpmSolver* s, MpmLocalHub* hub, int32_t rank, int32_t world)
{
    if (!s || !hub || world != hub->world || rank < 0 || rank >= world) return MPM_ERR_INVALID;
    if (cudaSetDevice(s->device) != cudaSuccess) { s->err = "cudaSetDevice failed"; return MPM_ERR_CUDA; }
    LocalTransport* tr = new LocalTransport();
    tr->hub = hub;
    tr->device = s->device;
    return comm_attach(s, tr, rank, world);
}

extern "C" int32_t mpm_comm_rebalance(MpmSolver* s, int32_t max_shift)
{
    if (!s) return MPM_ERR_INVALID;
    if (!s->comm || s->comm->world < 2) return MPM_OK;
    if (cudaSetDevice(s->device) != cudaSuccess) { s->err = "cudaSetDevice failed"; return MPM_ERR_CUDA; }
    int rc = comm_partition(s);
    if (rc) return rc;
    return rebalance(s, max_shift, 1ull);
}

extern "C" int32_t mpm_comm_rebalance_weighted(MpmSolver* s, int32_t max_shift, float cost_per_particle)
{
    if (!s) return MPM_ERR_INVALID;
    if (!s->comm || s->comm->world < 2) return MPM_OK;
    if (cudaSetDevice(s->device) != cudaSuccess) { s->err = "cudaSetDevice failed"; return MPM_ERR_CUDA; }
    int rc = comm_partition(s);
    if (rc) return rc;
    // quantised to 1/4096 of the caller's unit, at least one step, at most 2^24 steps (a plane of 10^6 particles stays far
    // below 2^63); not-a-number or non-positive costs count as 1.0
    const float c = (cost_per_particle > 0.0f && cost_per_particle < 4096.0f) ? cost_per_particle : 1.0f;
    const unsigned long long wq = std::max<unsigned long long>(1ull, (unsigned long long)llroundf(c * 4096.0f));
    return rebalance(s, max_shift, wq);
}

extern "C" int32_t mpm_comm_slab(const MpmSolver* s, int32_t* x0, int32_t* x1, int32_t* gx0, int32_t* nxl)
{
    if (!s) return MPM_ERR_INVALID;
    if (x0) *x0 = s->comm ? s->comm->x0 : 0;
    if (x1) *x1 = s->comm ? s->comm->x1 : s->dp.Rx;
    if (gx0) *gx0 = s->dp.gx0;
    if (nxl) *nxl = s->dp.nxl;
    return MPM_OK;
}

extern "C" int32_t mpm_download_ids(MpmSolver* s, uint32_t* ids, int64_t cap)
{
    if (!s || !ids) return MPM_ERR_INVALID;
    if (cap < s->n) { s->err = "destination too small"; return MPM_ERR_INVALID; }
    if (cudaSetDevice(s->device) != cudaSuccess) { s->err = "cudaSetDevice failed"; return MPM_ERR_CUDA; }
    if (s->n == 0) return MPM_OK;
    CKM(cudaMemcpyAsync(ids, s->orig_id, sizeof(uint32_t) * s->n, cudaMemcpyDeviceToHost, s->stream));
    CKM(cudaStreamSynchronize(s->stream));
    return MPM_OK;
}
