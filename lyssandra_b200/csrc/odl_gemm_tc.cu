// odl_gemm_tc.cu — the one dense contraction of the ODL dictionary update, DA = D A
// (lyssa/dict_learning/online_dict_learn.py:91; n x K times K x K: 1.07 GFLOP at cfg4), on the 5th-generation
// tensor cores: tcgen05.mma kind::f16 with TMEM accumulators, hand-written for sm_100a.
//
//     out[f][c] = sum_j D[f][j] A[j][c]            f < n <= 128 on the TMEM lanes (MMA M = 128, zero padded),
//                                                   c: 128 columns per CTA (MMA N), j: the contraction
// A is symmetric (A = sum beta^t Z Z^T), so column c of A is read as the contiguous row c: both operands are K-major
// without a transpose.  The contraction is split over the CTAs of a column tile (grid = K/128 x SPLIT ~ one CTA per SM);
// every CTA writes its partial tile and the column-update kernel (odl.cu) adds the SPLIT partials in a fixed order:
// deterministic, no atomics.
// Precision: fp32-faithful by the same exact two-plane fp16 split as the fused encode (hi = rn16(s v), lo = rn16(s v - hi),
// three products lo.hi + hi.lo + hi.hi accumulated in fp32 in TMEM, dropped term 2^-22).  The statistics A span many
// orders of magnitude (a large diagonal, small decayed off-diagonals), so every ROW of either operand gets its own
// power-of-two scale s (row maximum over the CTA's contraction range -> [16, 32)), undone exactly in the epilogue.
//
// Roles (160 threads): warps 0-3: thread t produces row t of both operand tiles (fp32 -> planes in the canonical K-major
// no-swizzle layout: 16-byte chunk kc of row r at kc*(128*16) + r*16), then reads TMEM lane t in the epilogue;
// warp 4: TMEM allocation + the single MMA-issuing thread.  Two operand stages of 64 contraction steps, mbarrier
// full/empty pairs, tcgen05.commit frees a stage.
#include "common.cuh"
#include "tc_ptx.cuh"
#include <cuda_fp16.h>
#include <algorithm>

namespace lys {
namespace {

using namespace tc;

constexpr int GM = 128;                  // rows of either operand tile (MMA M and N)
constexpr int GK = 64;                   // contraction steps per stage
constexpr int G_PLANE = GM * GK * 2;     // 16 KB: one fp16 plane of one operand tile
constexpr int G_OPER = 2 * G_PLANE;      // hi + lo
constexpr int G_STAGE = 2 * G_OPER;      // A operand (D rows) + B operand (rows of A)
constexpr int G_STAGES = 2;
constexpr int G_THREADS = 160;
constexpr int G_SMEM = G_STAGES * G_STAGE + 1024;
// kind::f16: D = F32 (bit 4), A = B = F16, both K-major, N = 128, M = 128
constexpr uint32_t kGIdesc = (1u << 4) | ((uint32_t)(GM >> 3) << 17) | ((uint32_t)(GM >> 4) << 24);

__device__ __forceinline__ float row_scale(float amax)
{
    int es = 258 - (int)(__float_as_uint(amax) >> 23);           // max |s v| in [16, 32)
    es = min(max(es, 1), 254);
    return __uint_as_float((uint32_t)es << 23);
}

// 64 consecutive floats of one row (zeros outside [0, limit)) -> hi/lo planes of row `row` of an operand tile
__device__ __forceinline__ void produce_row(unsigned char* oper, int row, const float* __restrict__ src, int j0, int limit, float s, bool live)
{
#pragma unroll
    for (int kc = 0; kc < GK / 8; ++kc) {
        float v[8];
        const int j = j0 + kc * 8;
        if (live && j + 8 <= limit && ((reinterpret_cast<uintptr_t>(src + j) & 15) == 0)) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(src + j)), b = __ldg(reinterpret_cast<const float4*>(src + j + 4));
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = (live && j + e < limit) ? __ldg(src + j + e) : 0.f;
        }
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float a = v[2 * e] * s, b = v[2 * e + 1] * s;
            const __half2 h = __floats2half2_rn(a, b);
            const float2 hf = __half22float2(h);
            const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
            hi[e] = *reinterpret_cast<const uint32_t*>(&h);
            lo[e] = *reinterpret_cast<const uint32_t*>(&l);
        }
        *reinterpret_cast<uint4*>(oper + kc * (GM * 16) + row * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(oper + G_PLANE + kc * (GM * 16) + row * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}

__global__ void __launch_bounds__(G_THREADS, 1)
da_gemm_tc_kernel(const float* __restrict__ D, int64_t ldd, const float* __restrict__ A, int n, int K, int jl /* contraction per CTA, multiple of 64 */,
                  float* __restrict__ partial /* [split][n][K] */)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G_STAGES * G_STAGE);      // full[2], empty[2], done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    float* s_cscale = reinterpret_cast<float*>(bars + 16);                         // [128] 1 / scale of the CTA's columns
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int c0 = blockIdx.x * GM, split = blockIdx.y, j_lo = split * jl, j_hi = min(K, j_lo + jl);
    const int n_chunks = (j_hi - j_lo + GK - 1) / GK;

    if (tid == 0) {
        for (int s = 0; s < G_STAGES; ++s) { mbar_init(smem_u32(&bars[s]), 4); mbar_init(smem_u32(&bars[2 + s]), 1); }
        mbar_init(smem_u32(&bars[4]), 1);
        mbar_init_fence();
    }
    if (warp == 4) tmem_alloc<1>(smem_u32(tmem_slot), 128);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t bar0 = smem_u32(&bars[0]);

    if (warp < 4) {
        // ------------------------------------------------------------ producers: thread t = row t of both tiles
        const int f = tid, c = c0 + tid;
        const bool f_live = f < n, c_live = c < K;
        const float* drow = D + (int64_t)f * ldd;
        const float* arow = A + (int64_t)c * K;                       // row c of A == column c (symmetric)
        // row maxima over this CTA's contraction range: 16 independent 128-bit loads in flight per operand (a scalar loop
        // waits one L2 round trip per element: 85 us for the whole product, measured)
        auto row_absmax = [&](const float* __restrict__ row, bool live) {
            float m = 0.f;
            if (!live) return m;
            int j = j_lo;
            if ((reinterpret_cast<uintptr_t>(row + j_lo) & 15) == 0) {
                for (; j + 64 <= j_hi; j += 64) {
                    float4 v[16];
#pragma unroll
                    for (int q = 0; q < 16; ++q) v[q] = __ldg(reinterpret_cast<const float4*>(row + j) + q);
#pragma unroll
                    for (int q = 0; q < 16; ++q)
                        m = fmaxf(fmaxf(m, fmaxf(fabsf(v[q].x), fabsf(v[q].y))), fmaxf(fabsf(v[q].z), fabsf(v[q].w)));
                }
            }
            for (; j < j_hi; ++j) m = fmaxf(m, fabsf(__ldg(row + j)));
            return m;
        };
        const float dmax = row_absmax(drow, f_live), amax = row_absmax(arow, c_live);
        const float sd = row_scale(dmax), sa = row_scale(amax);
        s_cscale[tid] = 1.f / sa;                                     // exact: powers of two
        for (int ch = 0; ch < n_chunks; ++ch) {
            const int st = ch % G_STAGES;
            if (ch >= G_STAGES) mbar_wait(bar0 + 8 * (2 + st), ((ch / G_STAGES) - 1) & 1);      // stage consumed by the MMAs
            unsigned char* stage = smem + st * G_STAGE;
            produce_row(stage, tid, drow, j_lo + ch * GK, j_hi, sd, f_live);
            produce_row(stage + G_OPER, tid, arow, j_lo + ch * GK, j_hi, sa, c_live);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa(bar0 + 8 * st, 0));
        }
        // ------------------------------------------------------------ epilogue: TMEM lane t = row f of the partial tile
        mbar_wait(bar0 + 8 * 4, 0);
        fence_after();
        const uint32_t tq = tmem_base + ((uint32_t)(warp * 32) << 16);
        const float inv_sd = 1.f / sd;
        float* out = partial + ((int64_t)split * n + f) * K + c0;
#pragma unroll 1
        for (int p = 0; p < GM / 32; ++p) {
            uint32_t r[32];
            LYS_TMEM_LD_X32(tq + p * 32, r);
            LYS_TMEM_WAIT_X32(r);
            if (f_live) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int cc = p * 32 + i;
                    if (c0 + cc < K) out[cc] = __uint_as_float(r[i]) * inv_sd * s_cscale[cc];
                }
            }
        }
        fence_before();
    } else {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            constexpr uint32_t LBO = GM * 16, SBO = 128;
            for (int ch = 0; ch < n_chunks; ++ch) {
                const int st = ch % G_STAGES;
                mbar_wait(bar0 + 8 * st, (ch / G_STAGES) & 1);
                fence_after();
                const uint32_t a_hi = smem_u32(smem + st * G_STAGE), a_lo = a_hi + G_PLANE;
                const uint32_t b_hi = a_hi + G_OPER, b_lo = b_hi + G_PLANE;
#pragma unroll
                for (int ks = 0; ks < GK / 16; ++ks)
                    mma_f16<1>(tmem_base, make_desc(a_lo + ks * 2 * LBO, LBO, SBO), make_desc(b_hi + ks * 2 * LBO, LBO, SBO), kGIdesc, (ch > 0 || ks > 0));
#pragma unroll
                for (int ks = 0; ks < GK / 16; ++ks)
                    mma_f16<1>(tmem_base, make_desc(a_hi + ks * 2 * LBO, LBO, SBO), make_desc(b_lo + ks * 2 * LBO, LBO, SBO), kGIdesc, 1);
#pragma unroll
                for (int ks = 0; ks < GK / 16; ++ks)
                    mma_f16<1>(tmem_base, make_desc(a_hi + ks * 2 * LBO, LBO, SBO), make_desc(b_hi + ks * 2 * LBO, LBO, SBO), kGIdesc, 1);
                commit<1>(bar0 + 8 * (2 + st));                       // the stage may be refilled
            }
            commit<1>(bar0 + 8 * 4);                                  // accumulator complete
        }
        __syncwarp();
    }
    __syncthreads();
    if (warp == 4) {
        fence_after();
        tmem_dealloc<1>(tmem_base, 128);
    }
}

}  // namespace

// number of contraction splits for K: about one CTA per SM, at most 16, every split a whole number (>= 1) of 64-step chunks
int da_gemm_tc_splits(int K)
{
    const int tiles = (K + GM - 1) / GM;
    const int chunks = std::max(1, (K + GK - 1) / GK);
    const int want = std::max(1, std::min(chunks, std::min(16, sm_count() / std::max(tiles, 1))));
    const int per = (chunks + want - 1) / want;
    return (chunks + per - 1) / per;
}
bool da_gemm_tc_supported(int n, int K) { return n >= 1 && n <= GM && K >= 1; }

// partial: [da_gemm_tc_splits(K)][n][K] floats; the caller adds the partials in split order
int da_gemm_tc(const float* D, int64_t ldd, const float* A, int n, int K, float* partial, cudaStream_t stream)
{
    const int tiles = (K + GM - 1) / GM;
    const int chunks = std::max(1, (K + GK - 1) / GK);
    const int splits = da_gemm_tc_splits(K);
    const int jl = ((chunks + splits - 1) / splits) * GK;
    LYS_CUDA(cudaFuncSetAttribute(da_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM));
    da_gemm_tc_kernel<<<dim3((unsigned)tiles, (unsigned)splits), G_THREADS, G_SMEM, stream>>>(D, ldd, A, n, K, jl, partial);
    LYS_LAUNCH_CHECK("da_gemm_tc_kernel");
    return LYS_OK;
}

}  // namespace lys
