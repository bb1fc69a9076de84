// odl.cu — Mairal online dictionary learning as the reference implements it
// (lyssa/dict_learning/online_dict_learn.py:79-98), on sparse codes.
//
//   K12 odl_accumulate   A = beta*A + Z_b Z_b^T ; B = beta*B + X_b Z_b^T        :84-85
//       The reference multiplies DENSE K x b matrices (2 K^2 b flop, >99 % zeros); with
//       (idx,val)[b][k] it is b*k^2 + b*k*n scattered adds.
//   K13 odl_update_dict  D <- norm_cols(clamp(D + (B - D A) diag(1/(A_kk + eps))))  :91-98
//       Jacobi: D A is formed ONCE from the old D (one GEMM), not refreshed per atom.
#include "common.cuh"
#include <algorithm>

namespace lys {
namespace {

__global__ void scale_kernel(float* __restrict__ p, int64_t count, float beta)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < count; i += stride) p[i] *= beta;
}

// one warp per signal: A[idx_a][idx_b] += z_a z_b ; B[:, idx_a] += x z_a
__global__ void __launch_bounds__(256)
odl_accumulate_kernel(const float* __restrict__ X, int64_t xfs, int64_t xss,
                      const int32_t* __restrict__ idx, const float* __restrict__ val,
                      int n, int K, int64_t b, int k, float* __restrict__ A, float* __restrict__ B)
{
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = w; i < b; i += nw) {
        const int32_t* ii = idx + i * k;
        const float* vv = val + i * k;
        // A: k*k pairs spread over the lanes
        for (int p = lane; p < k * k; p += 32) {
            int a = p / k, c = p % k;
            int ia = ii[a], ic = ii[c];
            if (ia >= 0 && ic >= 0) atomicAdd(&A[(int64_t)ia * K + ic], vv[a] * vv[c]);
        }
        // B: n features over the lanes, k atoms
        for (int f = lane; f < n; f += 32) {
            const float x = X[(int64_t)f * xfs + i * xss];
            for (int a = 0; a < k; ++a) {
                int ia = ii[a];
                if (ia >= 0) atomicAdd(&B[(int64_t)f * K + ia], x * vv[a]);
            }
        }
    }
}

// one warp per atom: u = D[:,c] + (B[:,c] - DA[:,c]) / (A[c,c] + eps); clamp; normalise
__global__ void __launch_bounds__(256)
odl_update_kernel(float* __restrict__ D, int64_t ldd, const float* __restrict__ A,
                  const float* __restrict__ B, const float* __restrict__ DA, int n, int K, int non_neg)
{
    const int lane = threadIdx.x & 31;
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= K) return;
    const float inv = 1.f / (A[(int64_t)c * K + c] + kRefEps);                 // :94
    float u[LYS_MAX_FEATURES / 32];
    float sq = 0.f;
#pragma unroll
    for (int q = 0; q < LYS_MAX_FEATURES / 32; ++q) {
        int f = lane + 32 * q;
        float v = 0.f;
        if (f < n) {
            v = inv * (B[(int64_t)f * K + c] - DA[(int64_t)f * K + c]) + D[(int64_t)f * ldd + c];
            if (non_neg && v < 0.f) v = 0.f;                                  // :96-97
        }
        u[q] = v;
        sq = fmaf(v, v, sq);
    }
    sq = warp_sum(sq);
    const float s = 1.f / (sqrtf(sq) + kRefEps);                               // :98 norm_cols
#pragma unroll
    for (int q = 0; q < LYS_MAX_FEATURES / 32; ++q) {
        int f = lane + 32 * q;
        if (f < n) D[(int64_t)f * ldd + c] = u[q] * s;
    }
}

}  // namespace
}  // namespace lys

using namespace lys;

extern "C" int lys_odl_accumulate(const float* Xb, int64_t xfs, int64_t xss, const int32_t* idx, const float* val,
                                  int n, int K, int64_t b, int k, float beta, float* A, float* B, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    LYS_CHECK_ARG(A && B, "lys_odl_accumulate: null statistics");
    LYS_CHECK_ARG(n >= 1 && n <= LYS_MAX_FEATURES && K >= 1 && K <= LYS_MAX_ATOMS && b >= 0 && k >= 1 && k <= LYS_MAX_NONZERO,
                  "lys_odl_accumulate: bad shape");
    const int blocks = sm_count() * 4;
    if (beta == 0.f) {       // beta_0 = 0 wipes the statistics exactly (0 * x), also for inf/nan-free inputs
        LYS_CUDA(cudaMemsetAsync(A, 0, sizeof(float) * (size_t)K * K, stream));
        LYS_CUDA(cudaMemsetAsync(B, 0, sizeof(float) * (size_t)n * K, stream));
    } else if (beta != 1.f) {
        scale_kernel<<<blocks, 256, 0, stream>>>(A, (int64_t)K * K, beta);
        scale_kernel<<<blocks, 256, 0, stream>>>(B, (int64_t)n * K, beta);
        LYS_LAUNCH_CHECK("scale_kernel");
    }
    if (b == 0) return LYS_OK;
    LYS_CHECK_ARG(Xb && idx && val, "lys_odl_accumulate: null pointer");
    int64_t want = (b + 7) / 8;
    odl_accumulate_kernel<<<(unsigned)std::min<int64_t>(want, blocks * 4), 256, 0, stream>>>(Xb, xfs, xss, idx, val, n, K, b, k, A, B);
    LYS_LAUNCH_CHECK("odl_accumulate_kernel");
    return LYS_OK;
}

extern "C" size_t lys_odl_update_workspace_bytes(int n, int K) { return align_up((size_t)n * K * 4, 256); }

extern "C" int lys_odl_update_dict(float* D, int64_t ldd, const float* A, const float* B, int n, int K,
                                   int non_neg, void* workspace, size_t workspace_bytes, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    LYS_CHECK_ARG(D && A && B && workspace, "lys_odl_update_dict: null pointer");
    LYS_CHECK_ARG(n >= 1 && n <= LYS_MAX_FEATURES && K >= 1 && K <= LYS_MAX_ATOMS && ldd >= K, "lys_odl_update_dict: bad shape");
    if (workspace_bytes < lys_odl_update_workspace_bytes(n, K)) { set_error("lys_odl_update_dict: workspace too small"); return LYS_EWORKSPACE; }
    float* DA = reinterpret_cast<float*>(workspace);
    int rc = sgemm_strided(D, ldd, 1, A, K, 1, DA, K, 1, n, K, K, stream);      // DA = D A   (:91)
    if (rc) return rc;
    odl_update_kernel<<<(K * 32 + 255) / 256, 256, 0, stream>>>(D, ldd, A, B, DA, n, K, non_neg);
    LYS_LAUNCH_CHECK("odl_update_kernel");
    return LYS_OK;
}
