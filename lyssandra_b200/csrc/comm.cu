// comm.cu — peer-mapped exchange buffers for the multi-rank K-SVD sweep (one process per GPU).
// Placeholder: creation succeeds for world == 1 only until the P2P path lands.
#include "common.cuh"

extern "C" int lys_comm_create(int rank, int world, void** comm)
{
    LYS_CHECK_ARG(comm, "lys_comm_create: null out pointer");
    LYS_CHECK_ARG(world >= 1 && rank >= 0 && rank < world, "lys_comm_create: bad rank/world");
    *comm = nullptr;
    if (world == 1) return LYS_OK;
    lys::set_error("lys_comm_create: multi-rank exchange not available in this build");
    return LYS_EUNSUPPORTED;
}
extern "C" int lys_comm_export(void*, unsigned char*) { lys::set_error("lys_comm_export: not available"); return LYS_EUNSUPPORTED; }
extern "C" int lys_comm_connect(void*, const unsigned char*) { lys::set_error("lys_comm_connect: not available"); return LYS_EUNSUPPORTED; }
extern "C" int lys_comm_destroy(void*) { return LYS_OK; }
