// comm.cu — peer-mapped mailboxes for the multi-rank K-SVD sweep (one process per GPU).
// Each rank cudaMalloc's one buffer (2.2 MB), exports its CUDA IPC handle (64 bytes); the host
// (torch.distributed all_gather) hands every rank all handles; peers are opened with
// cudaIpcOpenMemHandle, which maps them over NVLink/NVSwitch.  The sweep kernel then writes its
// per-atom sums straight into every peer's mailbox as tagged 8-byte words (comm.cuh).
#include "comm.cuh"
#include <string.h>

using namespace lys;

static_assert(sizeof(cudaIpcMemHandle_t) == LYS_COMM_HANDLE_BYTES, "IPC handle size");

extern "C" int lys_comm_create(int rank, int world, void** comm)
{
    LYS_CHECK_ARG(comm, "lys_comm_create: null out pointer");
    LYS_CHECK_ARG(world >= 1 && world <= COMM_MAX_RANKS && rank >= 0 && rank < world,
                  "lys_comm_create: bad rank/world (%d/%d, max %d ranks)", rank, world, COMM_MAX_RANKS);
    *comm = nullptr;
    if (world == 1) return LYS_OK;             // NULL comm == single rank
    CommHost* h = new CommHost();
    h->rank = rank; h->world = world;
    LYS_CUDA(cudaGetDevice(&h->device));
    LYS_CUDA(cudaMalloc(&h->local, COMM_BUFFER_BYTES));
    LYS_CUDA(cudaMemset(h->local, 0, COMM_BUFFER_BYTES));
    LYS_CUDA(cudaDeviceSynchronize());
    h->peer[rank] = h->local;
    *comm = h;
    return LYS_OK;
}

extern "C" int lys_comm_export(void* comm, unsigned char handle[LYS_COMM_HANDLE_BYTES])
{
    LYS_CHECK_ARG(comm && handle, "lys_comm_export: null argument");
    CommHost* h = reinterpret_cast<CommHost*>(comm);
    cudaIpcMemHandle_t ipc;
    LYS_CUDA(cudaIpcGetMemHandle(&ipc, h->local));
    memcpy(handle, &ipc, LYS_COMM_HANDLE_BYTES);
    return LYS_OK;
}

extern "C" int lys_comm_connect(void* comm, const unsigned char* all_handles)
{
    LYS_CHECK_ARG(comm && all_handles, "lys_comm_connect: null argument");
    CommHost* h = reinterpret_cast<CommHost*>(comm);
    for (int r = 0; r < h->world; ++r) {
        if (r == h->rank) continue;
        cudaIpcMemHandle_t ipc;
        memcpy(&ipc, all_handles + (size_t)r * LYS_COMM_HANDLE_BYTES, LYS_COMM_HANDLE_BYTES);
        LYS_CUDA(cudaIpcOpenMemHandle(&h->peer[r], ipc, cudaIpcMemLazyEnablePeerAccess));
    }
    h->dev.rank = h->rank; h->dev.world = h->world;
    for (int r = 0; r < COMM_MAX_RANKS; ++r)
        h->dev.box[r] = reinterpret_cast<unsigned long long*>(r < h->world ? h->peer[r] : nullptr);
    h->connected = true;
    return LYS_OK;
}

extern "C" int lys_comm_destroy(void* comm)
{
    if (!comm) return LYS_OK;
    CommHost* h = reinterpret_cast<CommHost*>(comm);
    for (int r = 0; r < h->world; ++r)
        if (r != h->rank && h->peer[r]) cudaIpcCloseMemHandle(h->peer[r]);
    if (h->local) cudaFree(h->local);
    delete h;
    return LYS_OK;
}
