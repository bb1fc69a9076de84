// spm.cu — ScSPM spatial-pyramid pooling straight from sparse codes ("next" row 1).
//
// Replaces the per-image loop of sc_spm_extractor.encode
// (lyssa/feature_extract/spatial_pyramid.py:45-97) and the pooling operators of
// lyssa/feature_extract/pooling.py:4-26 for a whole batch of images at once: the dense code
// matrix (K x patches) the reference pools over is never formed.
//   cell of patch p at level `lev` (:80-83):  floor(cy / (H/lev)) * lev + floor(cx / (W/lev)),
//       cy = py + psize/2 - 0.5, cx = px + psize/2 - 0.5 (:62-63), evaluated in float64 like the reference
//   feature (cell, atom) = pool over the patches of the cell of Z[atom, patch]:
//       sc_max_pooling  max |z|   (pooling.py:4-7)     -> atomicMax on the bit pattern of |z| (>= 0 orders like uint)
//       sum_pooling     sum z     (pooling.py:16-19)   -> atomicAdd
//       average_pooling sum z / #patches of the cell (pooling.py:22-26)
//   optional per-cell l2 normalisation x / (||x|| + eps) (feature_extract/preproc.py:8-15, utils/math.py:61-62)
//   empty cells stay zero (:74,:88); output order level-major, cell-major, atom-minor (:94-96).
// Bound: HBM/atomics — 8k bytes of codes + 12 bytes of position per patch in, n_levels * k atomics per patch.
#include "common.cuh"

namespace lys {
namespace {

constexpr int kMaxLevels = 8;

struct Levels { int lev[kMaxLevels]; int off[kMaxLevels]; int n; int total; };

__device__ __forceinline__ int cell_of(double cy, double cx, int H, int W, int lev)
{
    const double hunit = (double)H / lev, wunit = (double)W / lev;           // :78-79
    const double b = floor(cy / hunit) * lev + floor(cx / wunit);            // :83
    if (!(b >= 0.0) || b >= (double)(lev * lev)) return -1;                   // never equals a cell index j (:84-88)
    return (int)b;
}

__global__ void spm_count_kernel(const int32_t* __restrict__ patch_img, const float* __restrict__ pos, float psize,
                                 const int32_t* __restrict__ img_hw, int64_t N, Levels L, int32_t* __restrict__ count)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const int img = patch_img[p];
    const int H = img_hw[2 * img], W = img_hw[2 * img + 1];
    const double cy = (double)pos[2 * p] + (double)psize / 2 - 0.5, cx = (double)pos[2 * p + 1] + (double)psize / 2 - 0.5;
    for (int l = 0; l < L.n; ++l) {
        const int c = cell_of(cy, cx, H, W, L.lev[l]);
        if (c >= 0) atomicAdd(count + (int64_t)img * L.total + L.off[l] + c, 1);
    }
}

__global__ void spm_pool_kernel(const int32_t* __restrict__ idx, const float* __restrict__ val, int64_t N, int k, int K,
                                const int32_t* __restrict__ patch_img, const float* __restrict__ pos, float psize,
                                const int32_t* __restrict__ img_hw, Levels L, int pooling, float* __restrict__ F)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;       // one thread per (patch, code slot)
    if (t >= N * k) return;
    const int64_t p = t / k;
    const int a = idx[t];
    const float v = val[t];
    if (a < 0 || v == 0.f) return;                                            // not in the support: contributes |0|, +0
    const int img = patch_img[p];
    const int H = img_hw[2 * img], W = img_hw[2 * img + 1];
    const double cy = (double)pos[2 * p] + (double)psize / 2 - 0.5, cx = (double)pos[2 * p + 1] + (double)psize / 2 - 0.5;
    for (int l = 0; l < L.n; ++l) {
        const int c = cell_of(cy, cx, H, W, L.lev[l]);
        if (c < 0) continue;
        float* dst = F + ((int64_t)img * L.total + L.off[l] + c) * K + a;
        if (pooling == 0) atomicMax(reinterpret_cast<unsigned int*>(dst), __float_as_uint(fabsf(v)));
        else atomicAdd(dst, v);
    }
}

// one warp per (image, cell): average (divide by the patches of the cell) and / or l2-normalise
__global__ void spm_finalize_kernel(float* __restrict__ F, const int32_t* __restrict__ count, int64_t n_cells, int K,
                                    int average, int l2)
{
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= n_cells) return;
    float* row = F + w * K;
    const int cnt = count[w];
    if (cnt == 0) return;                                                     // empty cell: zeros, never normalised (:88)
    const float scale = average ? 1.f / (float)cnt : 1.f;
    double ss = 0.0;
    for (int c = lane; c < K; c += 32) { const float x = row[c] * scale; row[c] = x; ss += (double)x * x; }
    if (!l2) return;
    ss = warp_sum(ss);
    const float inv = 1.f / ((float)sqrt(ss) + kRefEps);
    for (int c = lane; c < K; c += 32) row[c] *= inv;
}

}  // namespace
}  // namespace lys

using namespace lys;

extern "C" int lys_spm_total_cells(const int32_t* levels, int n_levels)
{
    if (!levels || n_levels < 1 || n_levels > kMaxLevels) return -1;
    int t = 0;
    for (int l = 0; l < n_levels; ++l) { if (levels[l] < 1 || levels[l] > 64) return -1; t += levels[l] * levels[l]; }
    return t;
}

extern "C" int lys_spm_pool(const int32_t* idx, const float* val, int64_t N, int k, int K,
                            const int32_t* patch_img, const float* patch_pos, float patch_size,
                            const int32_t* img_hw, int n_imgs, const int32_t* levels_host, int n_levels,
                            int pooling, int l2_normalize, float* F, int32_t* cell_count, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    LYS_CHECK_ARG(n_levels >= 1 && n_levels <= kMaxLevels && levels_host, "lys_spm_pool: 1..%d pyramid levels", kMaxLevels);
    LYS_CHECK_ARG(pooling >= 0 && pooling <= 2, "lys_spm_pool: pooling must be 0 (max |z|), 1 (sum) or 2 (average)");
    LYS_CHECK_ARG(N >= 0 && k >= 1 && k <= LYS_MAX_NONZERO && K >= 1 && n_imgs >= 1, "lys_spm_pool: bad shape");
    LYS_CHECK_ARG(F && cell_count && (N == 0 || (idx && val && patch_img && patch_pos && img_hw)), "lys_spm_pool: null pointer");
    Levels L; L.n = n_levels; L.total = 0;
    for (int l = 0; l < n_levels; ++l) {
        LYS_CHECK_ARG(levels_host[l] >= 1 && levels_host[l] <= 64, "lys_spm_pool: level %d out of range", levels_host[l]);
        L.lev[l] = levels_host[l]; L.off[l] = L.total; L.total += levels_host[l] * levels_host[l];
    }
    const int64_t n_cells = (int64_t)n_imgs * L.total;
    LYS_CUDA(cudaMemsetAsync(F, 0, (size_t)n_cells * K * sizeof(float), stream));
    LYS_CUDA(cudaMemsetAsync(cell_count, 0, (size_t)n_cells * sizeof(int32_t), stream));
    if (N == 0) return LYS_OK;
    spm_count_kernel<<<(unsigned)((N + 255) / 256), 256, 0, stream>>>(patch_img, patch_pos, patch_size, img_hw, N, L, cell_count);
    LYS_LAUNCH_CHECK("spm_count_kernel");
    spm_pool_kernel<<<(unsigned)((N * k + 255) / 256), 256, 0, stream>>>(idx, val, N, k, K, patch_img, patch_pos, patch_size,
                                                                         img_hw, L, pooling, F);
    LYS_LAUNCH_CHECK("spm_pool_kernel");
    if (pooling == 2 || l2_normalize) {
        spm_finalize_kernel<<<(unsigned)((n_cells * 32 + 255) / 256), 256, 0, stream>>>(F, cell_count, n_cells, K, pooling == 2, l2_normalize);
        LYS_LAUNCH_CHECK("spm_finalize_kernel");
    }
    return LYS_OK;
}
