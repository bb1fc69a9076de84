// comm.cuh — peer-mapped exchange buffers shared by comm.cu (host side) and ksvd_sweep.cu (device side).
#pragma once
#include "common.cuh"

namespace lys {

constexpr int COMM_MAX_RANKS = 8;
constexpr int COMM_LD = 264;                 // floats per (parity, rank) slot: n + 2 <= 258, padded
constexpr size_t COMM_SLOT_FLOATS = 2 * COMM_MAX_RANKS * COMM_LD;
constexpr size_t COMM_FLAG_OFFSET_BYTES = COMM_SLOT_FLOATS * sizeof(float);        // flags follow the slots
constexpr size_t COMM_BUFFER_BYTES = COMM_FLAG_OFFSET_BYTES + 2 * COMM_MAX_RANKS * sizeof(unsigned) + 256;

// passed BY VALUE to the sweep kernel; slots[r]/flags[r] point into rank r's buffer
struct PeerComm {
    int rank, world;
    float* slots[COMM_MAX_RANKS];
    unsigned* flags[COMM_MAX_RANKS];
};

struct CommHost {
    int rank = 0, world = 1, device = 0;
    void* local = nullptr;
    void* peer[COMM_MAX_RANKS] = {nullptr};
    bool connected = false;
    unsigned epoch = 1;          // sequence base of the next sweep launch (identical on every rank)
    PeerComm dev{};
};

int transpose(const float* src, int64_t lds, float* dst, int64_t ldd, int rows, int cols, cudaStream_t st);

}  // namespace lys
