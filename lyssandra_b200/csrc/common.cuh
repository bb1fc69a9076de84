// common.cuh — shared helpers for the lyssa_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/lyssa_b200.h"

namespace lys {

// 2^-52: the reference's `np.finfo(float).eps` (lyssa/utils/math.py:61,65;
// lyssa/dict_learning/online_dict_learn.py:56).  Used where the reference ADDS eps to a
// denominator (normalize / norm_cols / ODL): representable in float32, it keeps
// "zero vector -> zeros" without perturbing any non-degenerate value.
constexpr float kRefEps = 2.220446049250313e-16f;
// Cholesky pivot floor (lyssa/sparse_coding.py:316,:335,:345 compare 1 - w.w with the
// machine epsilon of the arithmetic in use): float32 epsilon here.
constexpr float kPivotEps = 1.1920928955078125e-07f;

void set_error(const char* fmt, ...);

#define LYS_CHECK_ARG(cond, ...)                                   \
    do {                                                           \
        if (!(cond)) {                                             \
            ::lys::set_error(__VA_ARGS__);                         \
            return LYS_EINVAL;                                     \
        }                                                          \
    } while (0)

#define LYS_CUDA(call)                                                                      \
    do {                                                                                    \
        cudaError_t _e = (call);                                                            \
        if (_e != cudaSuccess) {                                                            \
            ::lys::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e),        \
                             __FILE__, __LINE__);                                           \
            return LYS_ECUDA;                                                               \
        }                                                                                   \
    } while (0)

#define LYS_LAUNCH_CHECK(name)                                                              \
    do {                                                                                    \
        cudaError_t _e = cudaGetLastError();                                                \
        if (_e != cudaSuccess) {                                                            \
            ::lys::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));      \
            return LYS_ECUDA;                                                               \
        }                                                                                   \
    } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int sm_count();   // SMs of the current device (cached per device)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// generic fp32 GEMM used for the small dense contractions (Gram, D*A) and by the
// generic-shape encode path: C(i,j) = sum_p A(i,p) * B(p,j), all operands strided.
int sgemm_strided(const float* A, int64_t sai, int64_t sap,
                  const float* B, int64_t sbp, int64_t sbj,
                  float* C, int64_t sci, int64_t scj,
                  int64_t M, int64_t Nc, int Kd, cudaStream_t stream);

}  // namespace lys
