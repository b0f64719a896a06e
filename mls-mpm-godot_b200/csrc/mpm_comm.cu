// mpm_comm.cu -- multi-GPU slab decomposition (one process per GPU): halo exchange and particle migration.
// PLACEHOLDER in this revision: the entry points exist so the ABI is stable, and report MPM_ERR_COMM.
#include "mpm_kernels.h"
#include "mpm_solver.h"

namespace mpm {
void comm_destroy(MpmSolver*) {}
int comm_exchange_halo(MpmSolver*, int) { return MPM_OK; }
int comm_migrate(MpmSolver*) { return MPM_OK; }
int comm_filter_upload(MpmSolver*, int64_t) { return MPM_OK; }
void comm_fill_stats(const MpmSolver*, MpmStats*) {}
}  // namespace mpm

extern "C" int32_t mpm_comm_unique_id(uint8_t id[MPM_COMM_ID_BYTES])
{
    (void)id;
    return MPM_ERR_COMM;
}
extern "C" int32_t mpm_comm_init(MpmSolver* s, const uint8_t id[MPM_COMM_ID_BYTES], int32_t rank, int32_t world)
{
    (void)id; (void)rank; (void)world;
    if (s) s->err = "multi-GPU slabs are not implemented in this build";
    return MPM_ERR_COMM;
}
