// dsift.cu — dense SIFT descriptors ("next" row 2): the producer of the 128-d
// inputs of the ScSPM pipeline, lyssa/feature_extract/dsift.py:75-162 (Lazebnik-style dense SIFT).
//   1. orientation maps (:127-135): I_h, I_w = image (*) 5x5 Gaussian-derivative kernels (true
//      convolution, zero padding, 'same'); mag = |grad|; map_i = mag * max(cos(theta - angle_i)^9, 0),
//      8 angles.  cos(theta - a) = (I_w cos a + I_h sin a) / mag, so no atan2 is needed.
//   2. descriptors (:136-141): for every grid patch and angle, the 16 spatial bins are the bilinear
//      weight matrix (16 x ps^2, separable: w[bi][row] * w[bj][col], :57-73) times the patch of the map.
//   3. Lowe normalisation (:146-162): f /= max(|f|, nrml_thres); f = min(f, sift_thres); renormalise the
//      high-contrast ones.
// Patch order and positions as in process_image (:106-118): p = a * n_h + b at (h_b, w_a).
// Bound: L2 — each patch stages 8 x ps x ps map values (overlapping patches re-read them from L2).
#include "common.cuh"
#include <algorithm>

namespace lys {
namespace {

constexpr int kAngles = 8, kBins = 4, kSamples = 16, kDesc = 128, kMaxPs = 32;

struct DsiftParams {
    float gh[25], gw[25];            // Gaussian-derivative kernels, row-major 5x5 (:23-35)
    float ca[kAngles], sa[kAngles];  // cos / sin of the 8 angles (:21)
    float wt[kBins * kMaxPs];        // bilinear bin weights w[bin][pixel] (:57-73, separable factor)
};

constexpr int kMaxBatch = 128;       // images of one size per launch (their pointers travel as a kernel parameter)
struct DsiftBatch {
    const float* img[kMaxBatch];
};

// blockIdx.z = image of the batch; orientation maps of image z at orient + z * 8 * H * W
__global__ void dsift_orient_kernel(DsiftBatch B, int64_t row_stride, int H, int W, DsiftParams P,
                                    float* __restrict__ orient_all)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const float* __restrict__ img = B.img[blockIdx.z];
    float* __restrict__ orient = orient_all + (size_t)blockIdx.z * kAngles * H * W;
    float ih = 0.f, iw = 0.f;
#pragma unroll
    for (int a = 0; a < 5; ++a) {
        const int yy = y - (a - 2);
        if (yy < 0 || yy >= H) continue;
#pragma unroll
        for (int b = 0; b < 5; ++b) {
            const int xx = x - (b - 2);
            if (xx < 0 || xx >= W) continue;
            const float v = __ldg(img + (int64_t)yy * row_stride + xx);
            ih = fmaf(P.gh[a * 5 + b], v, ih);
            iw = fmaf(P.gw[a * 5 + b], v, iw);
        }
    }
    const float mag = sqrtf(ih * ih + iw * iw);                               // :131
    const float inv = mag > 0.f ? 1.f / mag : 0.f;
#pragma unroll
    for (int i = 0; i < kAngles; ++i) {
        const float c = (iw * P.ca[i] + ih * P.sa[i]) * inv;                  // cos(theta - angle_i)
        const float c2 = c * c, c4 = c2 * c2;
        orient[((int64_t)i * H + y) * W + x] = mag * fmaxf(c * c4 * c4, 0.f); // alpha = 9 (:19,:135)
    }
}

// one block of 128 threads per patch.  The patch of the 8 orientation maps (8 x ps x ps floats) is staged in
// shared memory once; the separable bilinear weighting runs as a horizontal pass (angle, row, column-bin) and a
// vertical pass (angle, row-bin, column-bin) — same summation order as a direct double loop, 3x fewer
// multiply-adds and no repeated global reads.  Row pitches are padded by one float against bank conflicts.
// blockIdx.y = image of the batch.  PS > 0 fixes the patch size at compile time (16: the reference's default use).
template <int PS>
__global__ void __launch_bounds__(kDesc)
dsift_desc_kernel(const float* __restrict__ orient_all, int H, int W, int ps_rt, int gs, int off_h, int off_w,
                  int n_h, int n_w, DsiftParams P, float nrml_thres, float sift_thres,
                  float* __restrict__ desc_all, float* __restrict__ pos_all)
{
    extern __shared__ float dsm[];
    const int ps = PS > 0 ? PS : ps_rt;
    const float* __restrict__ orient = orient_all + (size_t)blockIdx.y * kAngles * H * W;
    float* __restrict__ desc = desc_all + (size_t)blockIdx.y * n_h * n_w * kDesc;
    float* __restrict__ pos = pos_all + (size_t)blockIdx.y * n_h * n_w * 2;
    const int pitch = ps + 1;
    float* smap = dsm;                                   // [angle][row][pitch]
    float* srow = smap + kAngles * ps * pitch;           // [angle][row][column bin]
    float* swt = srow + kAngles * ps * kBins;            // [bin][kMaxPs + 1]
    const int p = blockIdx.x;
    const int a = p / n_h, b = p % n_h;                                       // :107-109 meshgrid order
    const int h0 = off_h + b * gs, w0 = off_w + a * gs;
    const int t = threadIdx.x, ang = t / kSamples, bin = t % kSamples, bi = bin / kBins, bj = bin % kBins;
    for (int e = t; e < kAngles * ps * ps; e += kDesc) {
        const int j = e % ps, i = (e / ps) % ps, g = e / (ps * ps);
        smap[(g * ps + i) * pitch + j] = __ldg(orient + ((int64_t)g * H + (h0 + i)) * W + w0 + j);
    }
    if (t < kBins * kMaxPs) swt[(t / kMaxPs) * (kMaxPs + 1) + (t % kMaxPs)] = P.wt[t];
    __syncthreads();
    for (int o = t; o < kAngles * ps * kBins; o += kDesc) {
        const int cb = o % kBins, gi = o / kBins;                             // gi = angle * ps + row
        const float* r = smap + gi * pitch;
        const float* w = swt + cb * (kMaxPs + 1);
        float sacc = 0.f;
#pragma unroll
        for (int j = 0; j < ps; ++j) sacc = fmaf(w[j], r[j], sacc);
        srow[o] = sacc;
    }
    __syncthreads();
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < ps; ++i) acc = fmaf(swt[bi * (kMaxPs + 1) + i], srow[(ang * ps + i) * kBins + bj], acc);
    // ---- :146-162
    __shared__ float red[4];
    __shared__ float bc;
    float ss = warp_sum(acc * acc);
    if ((t & 31) == 0) red[t >> 5] = ss;
    __syncthreads();
    if (t == 0) bc = sqrtf(red[0] + red[1] + red[2] + red[3]);
    __syncthreads();
    const float siftlen = bc;
    const bool hcontrast = siftlen >= nrml_thres;
    float f = acc / fmaxf(siftlen, nrml_thres);
    f = fminf(f, sift_thres);
    __syncthreads();
    ss = warp_sum(f * f);
    if ((t & 31) == 0) red[t >> 5] = ss;
    __syncthreads();
    if (hcontrast) f /= sqrtf(red[0] + red[1] + red[2] + red[3]);
    desc[(int64_t)p * kDesc + t] = f;
    if (t == 0) { pos[2 * p] = (float)h0; pos[2 * p + 1] = (float)w0; }
}

}  // namespace
}  // namespace lys

using namespace lys;

extern "C" size_t lys_dsift_workspace_bytes(int H, int W)
{
    if (H < 1 || W < 1) return 0;
    return align_up((size_t)kAngles * H * W * sizeof(float), 256);
}

extern "C" int lys_dsift_grid(int H, int W, int grid_spacing, int patch_size, int* n_h, int* n_w, int* off_h, int* off_w)
{
    LYS_CHECK_ARG(H >= patch_size && W >= patch_size && grid_spacing >= 1 && patch_size >= 4 && patch_size <= kMaxPs,
                  "lys_dsift_grid: image %dx%d, patch %d (4..%d), spacing %d", H, W, patch_size, kMaxPs, grid_spacing);
    const int rem_h = (H - patch_size) % grid_spacing, rem_w = (W - patch_size) % grid_spacing;       // :101-102
    const int oh = rem_h / 2, ow = rem_w / 2;                                                           // :106-107
    if (off_h) *off_h = oh;
    if (off_w) *off_w = ow;
    if (n_h) *n_h = (H - patch_size - oh) / grid_spacing + 1;                                          // range(off, h-ps+1, gs)
    if (n_w) *n_w = (W - patch_size - ow) / grid_spacing + 1;
    return LYS_OK;
}

namespace lys {
namespace {
int dsift_run(const float* const* imgs, int n_imgs, int64_t row_stride, int H, int W, int grid_spacing, int patch_size,
              float nrml_thres, float sift_thres, const float* gh25, const float* gw25, const float* bin_weights,
              float* desc, float* pos, void* workspace, size_t workspace_bytes, cudaStream_t stream, const char* who)
{
    LYS_CHECK_ARG(imgs && n_imgs >= 1 && gh25 && gw25 && bin_weights && desc && pos && workspace, "%s: null pointer", who);
    int n_h, n_w, off_h, off_w;
    int rc = lys_dsift_grid(H, W, grid_spacing, patch_size, &n_h, &n_w, &off_h, &off_w);
    if (rc) return rc;
    const size_t per_img = lys_dsift_workspace_bytes(H, W);
    if (workspace_bytes < per_img) { set_error("%s: workspace too small", who); return LYS_EWORKSPACE; }
    DsiftParams P;
    for (int i = 0; i < 25; ++i) { P.gh[i] = gh25[i]; P.gw[i] = gw25[i]; }
    for (int i = 0; i < kAngles; ++i) { const double ang = i * 2.0 * 3.14159265358979323846 / kAngles; P.ca[i] = (float)cos(ang); P.sa[i] = (float)sin(ang); }
    for (int b = 0; b < kBins; ++b)
        for (int i = 0; i < kMaxPs; ++i) P.wt[b * kMaxPs + i] = (i < patch_size) ? bin_weights[b * patch_size + i] : 0.f;
    float* orient = reinterpret_cast<float*>(workspace);
    const int per_launch = (int)std::min<size_t>(kMaxBatch, workspace_bytes / per_img);
    const size_t desc_smem = ((size_t)kAngles * patch_size * (patch_size + 1) + (size_t)kAngles * patch_size * kBins +
                              (size_t)kBins * (kMaxPs + 1)) * sizeof(float);
    const size_t n_patch = (size_t)n_h * n_w;
    for (int i0 = 0; i0 < n_imgs; i0 += per_launch) {
        const int nb = std::min(per_launch, n_imgs - i0);
        DsiftBatch B;
        for (int i = 0; i < nb; ++i) {
            LYS_CHECK_ARG(imgs[i0 + i] != nullptr, "%s: null image", who);
            B.img[i] = imgs[i0 + i];
        }
        dim3 blk(32, 8), grd((W + 31) / 32, (H + 7) / 8, nb);
        dsift_orient_kernel<<<grd, blk, 0, stream>>>(B, row_stride, H, W, P, orient);
        LYS_LAUNCH_CHECK("dsift_orient_kernel");
        dim3 dgrd((unsigned)n_patch, nb);
        float* d0 = desc + (size_t)i0 * n_patch * kDesc;
        float* p0 = pos + (size_t)i0 * n_patch * 2;
        if (patch_size == 16)
            dsift_desc_kernel<16><<<dgrd, kDesc, desc_smem, stream>>>(orient, H, W, patch_size, grid_spacing, off_h, off_w, n_h, n_w, P,
                                                                      nrml_thres, sift_thres, d0, p0);
        else
            dsift_desc_kernel<0><<<dgrd, kDesc, desc_smem, stream>>>(orient, H, W, patch_size, grid_spacing, off_h, off_w, n_h, n_w, P,
                                                                     nrml_thres, sift_thres, d0, p0);
        LYS_LAUNCH_CHECK("dsift_desc_kernel");
    }
    return LYS_OK;
}
}  // namespace
}  // namespace lys

extern "C" int lys_dsift(const float* img, int64_t row_stride, int H, int W, int grid_spacing, int patch_size,
                         float nrml_thres, float sift_thres, const float* gh25, const float* gw25, const float* bin_weights,
                         float* desc, float* pos, void* workspace, size_t workspace_bytes, void* stream_)
{
    const float* one[1] = {img};
    return dsift_run(one, 1, row_stride, H, W, grid_spacing, patch_size, nrml_thres, sift_thres, gh25, gw25, bin_weights,
                     desc, pos, workspace, workspace_bytes, (cudaStream_t)stream_, "lys_dsift");
}

extern "C" int lys_dsift_batch(const float* const* imgs, int n_imgs, int64_t row_stride, int H, int W, int grid_spacing, int patch_size,
                               float nrml_thres, float sift_thres, const float* gh25, const float* gw25, const float* bin_weights,
                               float* desc, float* pos, void* workspace, size_t workspace_bytes, void* stream_)
{
    return dsift_run(imgs, n_imgs, row_stride, H, W, grid_spacing, patch_size, nrml_thres, sift_thres, gh25, gw25, bin_weights,
                     desc, pos, workspace, workspace_bytes, (cudaStream_t)stream_, "lys_dsift_batch");
}
