// comm.cuh — peer-mapped mailboxes shared by comm.cu (host side) and ksvd_sweep.cu (device side).
//
// Every rank owns ONE buffer of 64-bit words, mapped by all its peers over NVLink/NVSwitch (CUDA IPC).  A word is
// self-validating: the low 16 bits carry a generation tag, the high 48 bits the payload, and it is written with a
// single 8-byte store — so the exchange needs no flags and no fences: a reader spins on the word itself until the
// tag it expects shows up.
//
// Layout of a rank's buffer: mailbox[slot][source rank][COMM_LD] words, slot = atom % COMM_WINDOW for the per-atom
// sums of the K-SVD sweep, slot COMM_WINDOW for the one-off exchange of the fixed-point scale bound.
#pragma once
#include "common.cuh"

namespace lys {

constexpr int COMM_MAX_RANKS = 8;
constexpr int COMM_LD = 520;                 // words per (slot, rank): 2n + 3 <= 515 for n <= 256, padded
constexpr int COMM_WINDOW = 64;              // atoms in flight per mailbox ring (a rank can run at most a few atoms ahead)
constexpr size_t COMM_BUFFER_BYTES = (size_t)(COMM_WINDOW + 1) * COMM_MAX_RANKS * COMM_LD * sizeof(unsigned long long);

// passed BY VALUE to the sweep kernels; box[r] points at rank r's buffer
struct PeerComm {
    int rank, world;
    unsigned long long* box[COMM_MAX_RANKS];
};

__host__ __device__ inline size_t comm_word(int slot, int src_rank, int t)
{
    return ((size_t)slot * COMM_MAX_RANKS + src_rank) * COMM_LD + t;
}
// tag of generation g: never 0 (a zeroed buffer matches nothing)
__host__ __device__ inline unsigned comm_tag(unsigned g) { return g % 65535u + 1u; }

struct CommHost {
    int rank = 0, world = 1, device = 0;
    void* local = nullptr;
    void* peer[COMM_MAX_RANKS] = {nullptr};
    bool connected = false;
    unsigned generation = 1;     // next unused mailbox generation (identical on every rank: every rank makes the same calls)
    PeerComm dev{};
};

int transpose(const float* src, int64_t lds, float* dst, int64_t ldd, int rows, int cols, cudaStream_t st);

}  // namespace lys
