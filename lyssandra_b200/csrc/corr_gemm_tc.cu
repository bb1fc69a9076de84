// corr_gemm_tc.cu — the correlation GEMM Alpha = X^T D (lyssa/sparse_coding.py:631) on the
// 5th-generation tensor cores: tcgen05.mma with TMEM accumulators, hand-written for sm_100a.
//
// Precision: Batch-OMP's argmax decisions need fp32-faithful correlations (SURVEY.md §7 hard
// part 1: a single-pass TF32/BF16 GEMM flips ~1 % of the supports), so every fp32 operand is
// split EXACTLY into three bf16 planes, x = x1 + x2 + x3 (8+8+8 mantissa bits), and the six
// products of order <= 2^-16 are accumulated in fp32 in TMEM:
//     x.d ~= x3 d1 + x2 d2 + x1 d3 + x2 d1 + x1 d2 + x1 d1        (dropped terms <= 2^-24 |x||d|)
// i.e. 6 bf16 MMAs per k-step instead of 1 — still ~20x cheaper than the SIMT fp32 GEMM.
//
// Decomposition: C (signals x atoms).  One CTA owns ONE 256-atom slice of D for the whole
// kernel (its three bf16 planes, 96 KB, stay resident in shared memory) and walks over
// 128-signal tiles; M = 128 signals on the TMEM lanes, N = 256 atoms on the columns, K-dim = 64
// features = 4 k-steps of 16.  Roles (warp-specialised, 288 threads):
//     warps 0-3  epilogue: tcgen05.ld (32 lanes x 32 columns) -> 128-byte row segments of Alpha
//     warps 4-7  producers: load the fp32 X tile, split to bf16 planes, store them in the
//                canonical K-major (no-swizzle) core-matrix layout, fence.proxy.async, arrive
//     warp  8    TMEM allocation + the single MMA-issuing thread
// Pipelines: two A stages and two 256-column accumulator stages, mbarrier full/empty pairs;
// tcgen05.commit releases the A stage and publishes the accumulator in one go.
//
// Shared-memory operand layout (both operands K-major, SWIZZLE_NONE): a plane of R rows x 64
// bf16 is 8 K-chunks of 16 bytes; chunk kc of row r lives at  kc*(R*16) + r*16  bytes, so a
// core matrix (8 rows x 16 B) is 128 contiguous bytes: SBO = 128 B (next 8 rows),
// LBO = R*16 B (next K-chunk); k-step ks starts at +ks*2*LBO.
#include "common.cuh"
#include <cuda_bf16.h>

namespace lys {
namespace {

constexpr int TM = 128;          // signals per tile (MMA M, TMEM lanes)
constexpr int TN = 256;          // atoms per CTA slice (MMA N, TMEM columns per stage)
constexpr int NF = 64;           // features (MMA K extent)
constexpr int A_PLANE = TM * NF * 2;       // 16 KB
constexpr int B_PLANE = TN * NF * 2;       // 32 KB
constexpr int A_STAGE = 3 * A_PLANE;       // 48 KB
constexpr int SMEM_B = 3 * B_PLANE;        // 96 KB
constexpr int SMEM_A = 2 * A_STAGE;        // 96 KB
constexpr int SMEM_BAR = 128;
constexpr int SMEM_TOTAL = SMEM_B + SMEM_A + SMEM_BAR;
constexpr int THREADS = 288;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tc_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (sm_100: version field = 1)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

// kind::f16 instruction descriptor: D = F32, A = B = BF16, both K-major, M = 128, N = 256
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

__device__ __forceinline__ void split3(float x, __nv_bfloat16& h1, __nv_bfloat16& h2, __nv_bfloat16& h3)
{
    h1 = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(h1);          // exact
    h2 = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(h2);         // exact
    h3 = __float2bfloat16_rn(r2);
}
__device__ __forceinline__ uint32_t pack2(__nv_bfloat16 lo, __nv_bfloat16 hi)
{
    return (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
}

// write 8 consecutive K elements (one 16-byte chunk) of one row into the three planes
__device__ __forceinline__ void store_chunk(unsigned char* plane0, int plane_bytes, int chunk_off, const float* v)
{
    uint32_t w[3][4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        __nv_bfloat16 a1, a2, a3, b1, b2, b3;
        split3(v[2 * e], a1, a2, a3);
        split3(v[2 * e + 1], b1, b2, b3);
        w[0][e] = pack2(a1, b1); w[1][e] = pack2(a2, b2); w[2][e] = pack2(a3, b3);
    }
#pragma unroll
    for (int p = 0; p < 3; ++p)
        *reinterpret_cast<uint4*>(plane0 + p * plane_bytes + chunk_off) = make_uint4(w[p][0], w[p][1], w[p][2], w[p][3]);
}

#define TMEM_LD_32x32b_x32(taddr, r)                                                                          \
    asm volatile(                                                                                             \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                             \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                             \
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"             \
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),     \
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), \
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), \
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) \
        : "r"(taddr))

__global__ void __launch_bounds__(THREADS, 1)
corr_gemm_tc_kernel(const float* __restrict__ X, int64_t xfs, int64_t xss,
                    const uint4* __restrict__ planes /* per 256-atom slice: 3 planes in smem layout */,
                    int K, int64_t C, float* __restrict__ alpha, int accumulate)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sB = smem;
    unsigned char* sA = smem + SMEM_B;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_B + SMEM_A);
    // bars[0..1] a_full, [2..3] a_empty, [4..5] acc_full, [6..7] acc_empty, then the TMEM base
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_slices = K / TN;
    const int slice = blockIdx.x % n_slices;
    const int group = blockIdx.x / n_slices;
    const int n_groups = gridDim.x / n_slices;
    const int64_t n_tiles = (C + TM - 1) / TM;
    const int a0 = slice * TN;

    if (tid == 0) {
        mbar_init(smem_u32(&bars[0]), 128); mbar_init(smem_u32(&bars[1]), 128);
        mbar_init(smem_u32(&bars[2]), 1);   mbar_init(smem_u32(&bars[3]), 1);
        mbar_init(smem_u32(&bars[4]), 1);   mbar_init(smem_u32(&bars[5]), 1);
        mbar_init(smem_u32(&bars[6]), 128); mbar_init(smem_u32(&bars[7]), 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    // B planes: this CTA's 256 atoms x 64 features (pre-split by split_dict_planes_kernel),
    // 96 KB copied once and resident for the whole kernel
    {
        const uint4* src = planes + (size_t)slice * (SMEM_B / 16);
        uint4* dst = reinterpret_cast<uint4*>(sB);
        for (int item = tid; item < SMEM_B / 16; item += THREADS) dst[item] = __ldg(src + item);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= 4 && warp < 8) {
        // ------------------------------------------------------------- producers (one row each)
        const int row = tid - 128;
        int it = 0;
        for (int64_t t = group; t < n_tiles; t += n_groups, ++it) {
            const int st = it & 1;
            mbar_wait(smem_u32(&bars[2 + st]), ((it >> 1) & 1) ^ 1);
            const int64_t sig = t * TM + row;
            unsigned char* dst = sA + st * A_STAGE;
            const bool ok = sig < C;
            const float* xp = X + sig * xss;
            float xin[NF];
            if (xfs == 1) {
#pragma unroll
                for (int q = 0; q < NF / 4; ++q) {
                    const float4 v4 = ok ? __ldg(reinterpret_cast<const float4*>(xp) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                    xin[4 * q] = v4.x; xin[4 * q + 1] = v4.y; xin[4 * q + 2] = v4.z; xin[4 * q + 3] = v4.w;
                }
            } else {
#pragma unroll
                for (int f = 0; f < NF; ++f) xin[f] = ok ? __ldg(xp + (int64_t)f * xfs) : 0.f;
            }
#pragma unroll
            for (int kc = 0; kc < NF / 8; ++kc)
                store_chunk(dst, A_PLANE, kc * (TM * 16) + row * 16, xin + kc * 8);
            fence_async_smem();
            mbar_arrive(smem_u32(&bars[0 + st]));
        }
    } else if (warp == 8) {
        // ------------------------------------------------------------------- MMA issuer
        if (lane == 0) {
            const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
            constexpr uint32_t LBO_A = TM * 16, LBO_B = TN * 16, SBO = 128;
            // small products first: (3,1) (2,2) (1,3) (2,1) (1,2) (1,1)   [1-based plane indices]
            const int pa[6] = {2, 1, 0, 1, 0, 0};
            const int pb[6] = {0, 1, 2, 0, 1, 0};
            int it = 0;
            for (int64_t t = group; t < n_tiles; t += n_groups, ++it) {
                const int st = it & 1;
                const uint32_t ph = (it >> 1) & 1;
                mbar_wait(smem_u32(&bars[0 + st]), ph);          // A planes landed
                mbar_wait(smem_u32(&bars[6 + st]), ph ^ 1);      // accumulator stage drained
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(st * TN);
                uint32_t acc = 0;
#pragma unroll
                for (int p = 0; p < 6; ++p) {
#pragma unroll
                    for (int ks = 0; ks < NF / 16; ++ks) {
                        const uint64_t ad = make_desc(a_base + st * A_STAGE + pa[p] * A_PLANE + ks * 2 * LBO_A, LBO_A, SBO);
                        const uint64_t bd = make_desc(b_base + pb[p] * B_PLANE + ks * 2 * LBO_B, LBO_B, SBO);
                        tc_mma_bf16(d_tmem, ad, bd, kIdesc, acc);
                        acc = 1;
                    }
                }
                tc_commit(smem_u32(&bars[2 + st]));              // A stage free once these MMAs retire
                tc_commit(smem_u32(&bars[4 + st]));              // accumulator ready
            }
        }
        __syncwarp();
    } else {
        // -------------------------------------------------------------------- epilogue
        int it = 0;
        for (int64_t t = group; t < n_tiles; t += n_groups, ++it) {
            const int st = it & 1;
            mbar_wait(smem_u32(&bars[4 + st]), (it >> 1) & 1);
            tc_fence_after();
            // every thread owns one signal row and writes whole 128-byte lines of it (8 x 16 B);
            // measured faster than a shared-memory transpose to 512-byte warp-contiguous stores
            const int64_t sig = t * TM + warp * 32 + lane;
            float* out = alpha + sig * (int64_t)K + a0;
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(st * TN);
#pragma unroll 1
            for (int c = 0; c < TN / 32; ++c) {
                uint32_t r[32];
                TMEM_LD_32x32b_x32(taddr + c * 32, r);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (sig < C) {
                    if (accumulate) {                         // second 64-feature half of a 128-feature product
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float4 o = *reinterpret_cast<const float4*>(out + c * 32 + q * 4);
                            r[4 * q] = __float_as_uint(o.x + __uint_as_float(r[4 * q]));
                            r[4 * q + 1] = __float_as_uint(o.y + __uint_as_float(r[4 * q + 1]));
                            r[4 * q + 2] = __float_as_uint(o.z + __uint_as_float(r[4 * q + 2]));
                            r[4 * q + 3] = __float_as_uint(o.w + __uint_as_float(r[4 * q + 3]));
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        *reinterpret_cast<uint4*>(out + c * 32 + q * 4) = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
                }
            }
            tc_fence_before();
            mbar_arrive(smem_u32(&bars[6 + st]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// D (n=64, K) fp32 -> for every 256-atom slice, the three bf16 planes in the kernel's smem layout
__global__ void split_dict_planes_kernel(const float* __restrict__ D, int64_t ldd, int K, unsigned char* __restrict__ planes)
{
    const int item = blockIdx.x * blockDim.x + threadIdx.x;        // one 16-byte chunk of one atom
    if (item >= K * (NF / 8)) return;
    const int atom = item % K, kc = item / K;
    const int slice = atom / TN, a_local = atom % TN;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = __ldg(D + (int64_t)(kc * 8 + e) * ldd + atom);
    store_chunk(planes + (size_t)slice * SMEM_B, B_PLANE, kc * (TN * 16) + a_local * 16, v);
}

}  // namespace

// n = 64, or n = 128 as two 64-feature halves: the second launch adds its product to the first one's Alpha
bool corr_gemm_tc_supported(int n, int K)
{
    return (n == NF || n == 2 * NF) && K >= TN && (K % TN) == 0 && K <= LYS_MAX_ATOMS;
}

static size_t half_planes_bytes(int K) { return (size_t)(K / TN) * SMEM_B; }

size_t corr_gemm_tc_planes_bytes(int n, int K) { return corr_gemm_tc_supported(n, K) ? (size_t)(n / NF) * half_planes_bytes(K) : 0; }

int corr_gemm_tc_prepare(const float* D, int64_t ldd, int n, int K, void* planes, cudaStream_t stream)
{
    if (!corr_gemm_tc_supported(n, K)) return LYS_EUNSUPPORTED;
    const int items = K * (NF / 8);
    for (int h = 0; h < n / NF; ++h) {
        split_dict_planes_kernel<<<(items + 255) / 256, 256, 0, stream>>>(D + (int64_t)h * NF * ldd, ldd, K,
                                                                          reinterpret_cast<unsigned char*>(planes) + h * half_planes_bytes(K));
        LYS_LAUNCH_CHECK("split_dict_planes_kernel");
    }
    return LYS_OK;
}

// Alpha (C, K) row-major fp32 = X^T D for a chunk of C signals; `planes` from corr_gemm_tc_prepare
int corr_gemm_tc(const float* X, int64_t xfs, int64_t xss, const void* planes,
                 int n, int K, int64_t C, float* alpha, cudaStream_t stream)
{
    if (!corr_gemm_tc_supported(n, K)) return LYS_EUNSUPPORTED;
    if (C <= 0) return LYS_OK;
    // function attributes are per device and one process may drive several GPUs: set it on every launch
    LYS_CUDA(cudaFuncSetAttribute(corr_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    const int n_slices = K / TN;
    const int64_t n_tiles = (C + TM - 1) / TM;
    int groups = sm_count() / n_slices;
    if (groups < 1) groups = 1;
    if ((int64_t)groups > n_tiles) groups = (int)n_tiles;
    for (int h = 0; h < n / NF; ++h) {
        const uint4* pl = reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(planes) + h * half_planes_bytes(K));
        corr_gemm_tc_kernel<<<groups * n_slices, THREADS, SMEM_TOTAL, stream>>>(X + (int64_t)h * NF * xfs, xfs, xss, pl, K, C, alpha, h);
        LYS_LAUNCH_CHECK("corr_gemm_tc_kernel");
    }
    return LYS_OK;
}

}  // namespace lys
