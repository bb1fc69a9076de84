// gemm.cu — strided fp32 SIMT GEMM for the SMALL dense contractions of the path:
//   K1  Gram = D^T D            (lyssa/sparse_coding.py:630)      134 MFLOP, once per encode
//   K13 DA   = D A              (lyssa/dict_learning/online_dict_learn.py:91)
// and the correlation Alpha = X^T D of the generic-shape encode path (sparse_coding.py:631).
// fp32 FFMA with a fixed p = 0..Kd-1 accumulation order per output element (deterministic,
// fp32-faithful: the argmax decisions of Batch-OMP are sensitive to ~1e-6 relative error in
// Alpha, SURVEY.md §7 hard part 1, so no single-pass TF32/BF16 here).
#include "common.cuh"

namespace lys {

namespace {

constexpr int BM = 128, BN = 128, BK = 8, TM = 8, TN = 8, NT = 256;

__global__ void __launch_bounds__(NT)
sgemm_kernel(const float* __restrict__ A, int64_t sai, int64_t sap,
             const float* __restrict__ B, int64_t sbp, int64_t sbj,
             float* __restrict__ C, int64_t sci, int64_t scj,
             int64_t M, int64_t Nc, int Kd)
{
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int t = threadIdx.x;
    const int64_t i0 = (int64_t)blockIdx.y * BM;
    const int64_t j0 = (int64_t)blockIdx.x * BN;
    const int ty = t / 16, tx = t % 16;      // 16 x 16 threads, each 8 x 8 outputs

    float acc[TM][TN];
#pragma unroll
    for (int a = 0; a < TM; ++a)
#pragma unroll
        for (int b = 0; b < TN; ++b) acc[a][b] = 0.f;

    const bool a_p_contig = (sap == 1);
    const bool b_j_contig = (sbj == 1);

    for (int p0 = 0; p0 < Kd; p0 += BK) {
#pragma unroll
        for (int e = 0; e < (BM * BK) / NT; ++e) {
            int id = t + e * NT;
            int i, p;
            if (a_p_contig) { i = id / BK; p = id % BK; } else { i = id % BM; p = id / BM; }
            int64_t gi = i0 + i; int gp = p0 + p;
            As[p][i] = (gi < M && gp < Kd) ? A[gi * sai + gp * sap] : 0.f;
        }
#pragma unroll
        for (int e = 0; e < (BN * BK) / NT; ++e) {
            int id = t + e * NT;
            int j, p;
            if (b_j_contig) { j = id % BN; p = id / BN; } else { p = id % BK; j = id / BK; }
            int64_t gj = j0 + j; int gp = p0 + p;
            Bs[p][j] = (gj < Nc && gp < Kd) ? B[gp * sbp + gj * sbj] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int p = 0; p < BK; ++p) {
            float av[TM], bv[TN];
#pragma unroll
            for (int a = 0; a < TM; a += 4) {
                float4 v = *reinterpret_cast<const float4*>(&As[p][ty * 4 + a * 16]);
                av[a] = v.x; av[a + 1] = v.y; av[a + 2] = v.z; av[a + 3] = v.w;
            }
#pragma unroll
            for (int b = 0; b < TN; b += 4) {
                float4 v = *reinterpret_cast<const float4*>(&Bs[p][tx * 4 + b * 16]);
                bv[b] = v.x; bv[b + 1] = v.y; bv[b + 2] = v.z; bv[b + 3] = v.w;
            }
#pragma unroll
            for (int a = 0; a < TM; ++a)
#pragma unroll
                for (int b = 0; b < TN; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
        }
        __syncthreads();
    }
    // rows: ty*4 + {0..3} and 64 + ty*4 + {0..3}; cols likewise with tx
#pragma unroll
    for (int a = 0; a < TM; ++a) {
        int64_t gi = i0 + ty * 4 + (a / 4) * 64 + (a % 4);
        if (gi >= M) continue;
#pragma unroll
        for (int b = 0; b < TN; ++b) {
            int64_t gj = j0 + tx * 4 + (b / 4) * 64 + (b % 4);
            if (gj < Nc) C[gi * sci + gj * scj] = acc[a][b];
        }
    }
}

// small-tile variant for skinny problems (few output rows, e.g. D*A with n = 128 rows): 32x32
// tiles keep all SMs busy where the 128x128 kernel would launch a handful of CTAs
constexpr int SB = 32, SBK = 32;

__global__ void __launch_bounds__(NT)
sgemm_small_kernel(const float* __restrict__ A, int64_t sai, int64_t sap,
                   const float* __restrict__ B, int64_t sbp, int64_t sbj,
                   float* __restrict__ C, int64_t sci, int64_t scj,
                   int64_t M, int64_t Nc, int Kd)
{
    __shared__ float As[SBK][SB + 1];
    __shared__ float Bs[SBK][SB + 1];
    const int t = threadIdx.x;
    const int64_t i0 = (int64_t)blockIdx.y * SB, j0 = (int64_t)blockIdx.x * SB;
    const int ty = t / 16, tx = t % 16;            // each thread: rows {ty, ty+16}, cols {tx, tx+16}
    float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    const bool a_p_contig = (sap == 1), b_j_contig = (sbj == 1);
    for (int p0 = 0; p0 < Kd; p0 += SBK) {
#pragma unroll
        for (int e = 0; e < (SB * SBK) / NT; ++e) {
            const int id = t + e * NT;
            int i, p;
            if (a_p_contig) { i = id / SBK; p = id % SBK; } else { i = id % SB; p = id / SB; }
            const int64_t gi = i0 + i; const int gp = p0 + p;
            As[p][i] = (gi < M && gp < Kd) ? A[gi * sai + gp * sap] : 0.f;
            int j, q;
            if (b_j_contig) { j = id % SB; q = id / SB; } else { q = id % SBK; j = id / SBK; }
            const int64_t gj = j0 + j; const int gq = p0 + q;
            Bs[q][j] = (gj < Nc && gq < Kd) ? B[gq * sbp + gj * sbj] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int p = 0; p < SBK; ++p) {
            const float a0 = As[p][ty], a1 = As[p][ty + 16], b0 = Bs[p][tx], b1 = Bs[p][tx + 16];
            acc[0][0] = fmaf(a0, b0, acc[0][0]); acc[0][1] = fmaf(a0, b1, acc[0][1]);
            acc[1][0] = fmaf(a1, b0, acc[1][0]); acc[1][1] = fmaf(a1, b1, acc[1][1]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const int64_t gi = i0 + ty + 16 * a;
        if (gi >= M) continue;
#pragma unroll
        for (int bq = 0; bq < 2; ++bq) {
            const int64_t gj = j0 + tx + 16 * bq;
            if (gj < Nc) C[gi * sci + gj * scj] = acc[a][bq];
        }
    }
}

}  // namespace

int sgemm_strided(const float* A, int64_t sai, int64_t sap,
                  const float* B, int64_t sbp, int64_t sbj,
                  float* C, int64_t sci, int64_t scj,
                  int64_t M, int64_t Nc, int Kd, cudaStream_t stream)
{
    if (M <= 0 || Nc <= 0) return LYS_OK;
    const int64_t big_ctas = ((Nc + BN - 1) / BN) * ((M + BM - 1) / BM);
    if (big_ctas < sm_count() && (Nc + SB - 1) / SB <= 65535 && (M + SB - 1) / SB <= 65535) {
        dim3 g((unsigned)((Nc + SB - 1) / SB), (unsigned)((M + SB - 1) / SB));
        sgemm_small_kernel<<<g, NT, 0, stream>>>(A, sai, sap, B, sbp, sbj, C, sci, scj, M, Nc, Kd);
        LYS_LAUNCH_CHECK("sgemm_small_kernel");
        return LYS_OK;
    }
    dim3 grid((unsigned)((Nc + BN - 1) / BN), (unsigned)((M + BM - 1) / BM));
    if (grid.y > 65535u) { set_error("sgemm: M too large for one launch"); return LYS_EINVAL; }
    sgemm_kernel<<<grid, NT, 0, stream>>>(A, sai, sap, B, sbp, sbj, C, sci, scj, M, Nc, Kd);
    LYS_LAUNCH_CHECK("sgemm_kernel");
    return LYS_OK;
}

}  // namespace lys

namespace lys {
bool corr_gemm_tc_supported(int n, int K);
size_t corr_gemm_tc_planes_bytes(int n, int K);
int corr_gemm_tc_prepare(const float* D, int64_t ldd, int n, int K, void* planes, cudaStream_t stream);
int corr_gemm_tc(const float* X, int64_t xfs, int64_t xss, const void* planes,
                 int n, int K, int64_t C, float* alpha, cudaStream_t stream);
}

extern "C" size_t lys_gram_workspace_bytes(int n, int K)
{
    return lys::corr_gemm_tc_supported(n, K) ? lys::align_up(lys::corr_gemm_tc_planes_bytes(n, K), 256) : 0;
}

// Gram with a scratch buffer: on the tcgen05 correlation GEMM where its shapes allow (n = 64 or 128, K a multiple of
// 256: G = D^T D is the correlation of the atoms with the dictionary), else the fp32 SIMT GEMM of lys_gram
extern "C" int lys_gram_ws(const float* D, int64_t ldd, int n, int K, float* G, void* workspace, size_t workspace_bytes, void* stream)
{
    LYS_CHECK_ARG(D && G, "lys_gram_ws: null pointer");
    LYS_CHECK_ARG(n >= 1 && n <= LYS_MAX_FEATURES && K >= 1 && K <= LYS_MAX_ATOMS && ldd >= K,
                  "lys_gram_ws: bad shape n=%d K=%d ldd=%lld", n, K, (long long)ldd);
    if (lys::corr_gemm_tc_supported(n, K) && workspace && workspace_bytes >= lys_gram_workspace_bytes(n, K)) {
        int rc = lys::corr_gemm_tc_prepare(D, ldd, n, K, workspace, (cudaStream_t)stream);
        if (rc) return rc;
        return lys::corr_gemm_tc(D, ldd, 1, workspace, n, K, K, G, (cudaStream_t)stream);      // "signals" = the atoms themselves
    }
    return lys::sgemm_strided(D, 1, ldd, D, ldd, 1, G, K, 1, K, K, n, (cudaStream_t)stream);
}

extern "C" int lys_gram(const float* D, int64_t ldd, int n, int K, float* G, void* stream)
{
    LYS_CHECK_ARG(D && G, "lys_gram: null pointer");
    LYS_CHECK_ARG(n >= 1 && n <= LYS_MAX_FEATURES && K >= 1 && K <= LYS_MAX_ATOMS && ldd >= K,
                  "lys_gram: bad shape n=%d K=%d ldd=%lld", n, K, (long long)ldd);
    // G(i,j) = sum_f D[f][i] D[f][j]
    return lys::sgemm_strided(D, 1, ldd, D, ldd, 1, G, K, 1, K, K, n, (cudaStream_t)stream);
}
