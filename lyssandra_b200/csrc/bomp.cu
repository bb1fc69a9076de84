// bomp.cu — C-ABI entry points of the Batch-OMP encode path (lyssa/sparse_coding.py:629-635,
// :708-726 -> batch_omp :302-367) and the host-buffer pipeline behind
// sparse_encoder.encode(X_numpy, D_numpy).
#include "common.cuh"
#include <mutex>
#include <atomic>
#include <thread>
#include <chrono>
#include <emmintrin.h>
#include <vector>
#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace lys {

int bomp_greedy_generic(const float* alpha, const float* G, int K, int64_t C, int k,
                        int32_t* idx, float* val, int32_t* nsel,
                        float* Z, int64_t zas, int64_t zss, cudaStream_t stream);

int omp_greedy_generic(const float* alpha, const float* G, int K, int64_t C, int k, const float* xnorm2, float tol, int strict,
                       int32_t* idx, float* val, int32_t* nsel, float* Z, int64_t zas, int64_t zss, int32_t* truncated,
                       cudaStream_t stream);
int col_norm2(const float* X, int64_t xfs, int64_t xss, int n, int64_t C, float* out, cudaStream_t stream);

bool bomp_fast_supported(int K, int k, int64_t zas, bool has_Z, const float* Z, int64_t zss);
int bomp_greedy_fast(const float* alpha, const float* G, int K, int64_t C, int k,
                     int32_t* idx, float* val, int32_t* nsel, float* Z, int64_t zss, cudaStream_t stream);

bool corr_gemm_tc_supported(int n, int K);
size_t corr_gemm_tc_planes_bytes(int n, int K);
int corr_gemm_tc_prepare(const float* D, int64_t ldd, int n, int K, void* planes, cudaStream_t stream);
int corr_gemm_tc(const float* X, int64_t xfs, int64_t xss, const void* planes,
                 int n, int K, int64_t C, float* alpha, cudaStream_t stream);

// fused tcgen05 path (bomp_fused.cu); returns LYS_EUNSUPPORTED for shapes it is not built for
int bomp_encode_fused(const float* X, int64_t xfs, int64_t xss, const float* D, int64_t ldd,
                      const float* G, int n, int K, int64_t N, int k,
                      int32_t* idx, float* val, int32_t* nsel,
                      float* Z, int64_t zas, int64_t zss,
                      void* workspace, size_t workspace_bytes, int screen, cudaStream_t stream);
size_t bomp_fused_workspace_bytes(int n, int K, int64_t N, int k);
int bomp_fused_launch_count(int n, int K, int64_t N, int k);

// ---- dominant-kernel profiling (see lys_profile_enable in the header)
struct ProfState {
    bool on = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> used, pool;
    const char* name = "none";
};
static ProfState g_prof;
static std::mutex g_prof_mu;

bool profile_begin(cudaStream_t st, const char* name, cudaEvent_t* stop_out)
{
    if (!g_prof.on) return false;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    std::pair<cudaEvent_t, cudaEvent_t> ev;
    if (!g_prof.pool.empty()) { ev = g_prof.pool.back(); g_prof.pool.pop_back(); }
    else {
        if (cudaEventCreate(&ev.first) != cudaSuccess || cudaEventCreate(&ev.second) != cudaSuccess) return false;
    }
    cudaEventRecord(ev.first, st);
    g_prof.used.push_back(ev);
    g_prof.name = name;
    *stop_out = ev.second;
    return true;
}

namespace {

constexpr int NF_MAX_HOOK = 128;

// signals per correlation chunk: the (chunk x K) fp32 Alpha tile is written by the GEMM kernel
// and read once by the greedy kernel.  Measured on B200 (profiles/README.md): per-launch fixed
// costs (D-plane staging, tail waves) outweigh L2 residency of the tile — 16/64/128/512 MB
// chunks give 10.8/6.9/6.2/5.5 ms per 1M-patch encode — so the chunk is 512 MB.
constexpr int64_t kChunkBytesTarget = 512ll << 20;

int64_t generic_chunk(int K, int64_t N)
{
    const int64_t target = kChunkBytesTarget;
    int64_t c = target / ((int64_t)K * 4);
    c = std::max<int64_t>(1024, c / 1024 * 1024);
    return std::min<int64_t>(c, std::max<int64_t>(N, 1));
}

__global__ void dense_fill_kernel(const int32_t* __restrict__ idx, const float* __restrict__ val,
                                  int64_t N, int k, int K, float* __restrict__ Z,
                                  int64_t zas, int64_t zss)
{
    // one warp per signal: zero the dense row/column, then scatter the k coefficients
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = w; i < N; i += nw) {
        float* z = Z + i * zss;
        if (zas == 1 && (K % 4) == 0 && ((reinterpret_cast<uintptr_t>(z) & 15) == 0)) {
            float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int c = lane * 4; c < K; c += 128) *reinterpret_cast<float4*>(z + c) = zero;
        } else {
            for (int c = lane; c < K; c += 32) z[(int64_t)c * zas] = 0.f;
        }
        __syncwarp();
        for (int j = lane; j < k; j += 32) {
            int a = idx[i * k + j];
            if (a >= 0) z[(int64_t)a * zas] = val[i * k + j];
        }
    }
}

}  // namespace
}  // namespace lys

using namespace lys;

extern "C" size_t lys_bomp_workspace_bytes(int n, int K, int64_t N, int k)
{
    if (n < 1 || K < 1 || N < 0 || k < 1) return 0;
    size_t generic = align_up((size_t)generic_chunk(K, N) * (size_t)K * sizeof(float), 256) + align_up(corr_gemm_tc_planes_bytes(n, K), 256) + 256;
    size_t fused = bomp_fused_workspace_bytes(n, K, N, k);
    return std::max(generic, fused);
}

extern "C" int lys_bomp_encode(const float* X, int64_t xfs, int64_t xss,
                               const float* D, int64_t ldd, const float* G,
                               int n, int K, int64_t N, int k,
                               int32_t* idx, float* val, int32_t* nsel,
                               float* Z, int64_t zas, int64_t zss,
                               void* workspace, size_t workspace_bytes, void* stream_)
{
    return lys_bomp_encode_ex(X, xfs, xss, D, ldd, G, n, K, N, k, idx, val, nsel, Z, zas, zss, workspace, workspace_bytes, 0, stream_);
}

extern "C" int lys_bomp_encode_ex(const float* X, int64_t xfs, int64_t xss,
                                  const float* D, int64_t ldd, const float* G,
                                  int n, int K, int64_t N, int k,
                                  int32_t* idx, float* val, int32_t* nsel,
                                  float* Z, int64_t zas, int64_t zss,
                                  void* workspace, size_t workspace_bytes, int flags, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    LYS_CHECK_ARG((flags & ~LYS_BOMP_SCREEN) == 0, "lys_bomp_encode_ex: unknown flags 0x%x", flags);
    LYS_CHECK_ARG(n >= 1 && n <= LYS_MAX_FEATURES, "lys_bomp_encode: n=%d out of range [1,%d]", n, LYS_MAX_FEATURES);
    LYS_CHECK_ARG(K >= 1 && K <= LYS_MAX_ATOMS, "lys_bomp_encode: K=%d out of range [1,%d]", K, LYS_MAX_ATOMS);
    LYS_CHECK_ARG(k >= 1 && k <= LYS_MAX_NONZERO && k <= K,
                  "lys_bomp_encode: n_nonzero_coefs=%d must be in [1, min(K=%d, %d)]", k, K, LYS_MAX_NONZERO);
    LYS_CHECK_ARG(N >= 0, "lys_bomp_encode: N < 0");
    if (N == 0) return LYS_OK;
    LYS_CHECK_ARG(X && D && G && idx && val, "lys_bomp_encode: null pointer");
    LYS_CHECK_ARG(ldd >= K, "lys_bomp_encode: ldd < K");
    LYS_CHECK_ARG(!Z || (zas >= 1 && zss >= 1), "lys_bomp_encode: bad Z strides");
    LYS_CHECK_ARG(workspace != nullptr, "lys_bomp_encode: workspace is null");
    if (workspace_bytes < lys_bomp_workspace_bytes(n, K, N, k)) {
        set_error("lys_bomp_encode: workspace %zu B < required %zu B", workspace_bytes,
                  lys_bomp_workspace_bytes(n, K, N, k));
        return LYS_EWORKSPACE;
    }

    int rc = bomp_encode_fused(X, xfs, xss, D, ldd, G, n, K, N, k, idx, val, nsel, Z, zas, zss,
                               workspace, workspace_bytes, (flags & LYS_BOMP_SCREEN) != 0, stream);
    if (rc != LYS_EUNSUPPORTED) return rc;

    // generic path: per chunk, Alpha = X_chunk^T D (fp32 GEMM) then one warp per signal
    float* alpha = reinterpret_cast<float*>(workspace);
    const int64_t chunk = generic_chunk(K, N);
    void* planes = reinterpret_cast<unsigned char*>(workspace) + align_up((size_t)chunk * (size_t)K * sizeof(float), 256);
    const bool use_tc = corr_gemm_tc_supported(n, K);
    const bool fast = bomp_fast_supported(K, k, zas, Z != nullptr, Z, zss);
    if (use_tc) {
        rc = corr_gemm_tc_prepare(D, ldd, n, K, planes, stream);
        if (rc) return rc;
    }
    for (int64_t s0 = 0; s0 < N; s0 += chunk) {
        const int64_t C = std::min(chunk, N - s0);
        if (use_tc) rc = corr_gemm_tc(X + s0 * xss, xfs, xss, planes, n, K, C, alpha, stream);
        else rc = sgemm_strided(X + s0 * xss, xss, xfs, D, ldd, 1, alpha, K, 1, C, K, n, stream);
        if (rc) return rc;
        cudaEvent_t stop_ev;
        const bool prof = profile_begin(stream, fast ? "bomp_fast_kernel" : "bomp_warp_kernel", &stop_ev);
        if (fast)
            rc = bomp_greedy_fast(alpha, G, K, C, k, idx + s0 * k, val + s0 * k, nsel ? nsel + s0 : nullptr,
                                  Z ? Z + s0 * zss : nullptr, zss, stream);
        else
            rc = bomp_greedy_generic(alpha, G, K, C, k, idx + s0 * k, val + s0 * k,
                                     nsel ? nsel + s0 : nullptr,
                                     Z ? Z + s0 * zss : nullptr, zas, zss, stream);
        if (prof) cudaEventRecord(stop_ev, stream);
        if (rc) return rc;
    }
    return LYS_OK;
}

extern "C" size_t lys_omp_workspace_bytes(int n, int K, int64_t N, int k_max)
{
    if (n < 1 || K < 1 || N < 0 || k_max < 1) return 0;
    const int64_t chunk = generic_chunk(K, N);
    return align_up((size_t)chunk * (size_t)K * sizeof(float), 256) + align_up(corr_gemm_tc_planes_bytes(n, K), 256) +
           align_up((size_t)chunk * sizeof(float), 256) + 256;
}

extern "C" int lys_omp_encode(const float* X, int64_t xfs, int64_t xss, const float* D, int64_t ldd, const float* G,
                              int n, int K, int64_t N, int k, float tol, int strict,
                              int32_t* idx, float* val, int32_t* nsel,
                              float* Z, int64_t zas, int64_t zss, int32_t* truncated,
                              void* workspace, size_t workspace_bytes, void* stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    LYS_CHECK_ARG(n >= 1 && n <= LYS_MAX_FEATURES, "lys_omp_encode: n=%d out of range [1,%d]", n, LYS_MAX_FEATURES);
    LYS_CHECK_ARG(K >= 1 && K <= LYS_MAX_ATOMS, "lys_omp_encode: K=%d out of range [1,%d]", K, LYS_MAX_ATOMS);
    LYS_CHECK_ARG(k >= 1 && k <= LYS_OMP_MAX_NONZERO && k <= K,
                  "lys_omp_encode: k_max=%d must be in [1, min(K=%d, %d)]", k, K, LYS_OMP_MAX_NONZERO);
    LYS_CHECK_ARG(tol >= 0.f, "lys_omp_encode: tol < 0");
    LYS_CHECK_ARG(N >= 0, "lys_omp_encode: N < 0");
    if (N == 0) return LYS_OK;
    LYS_CHECK_ARG(X && D && G && idx && val && workspace, "lys_omp_encode: null pointer");
    LYS_CHECK_ARG(ldd >= K, "lys_omp_encode: ldd < K");
    LYS_CHECK_ARG(!Z || (zas >= 1 && zss >= 1), "lys_omp_encode: bad Z strides");
    if (workspace_bytes < lys_omp_workspace_bytes(n, K, N, k)) {
        set_error("lys_omp_encode: workspace %zu B < required %zu B", workspace_bytes, lys_omp_workspace_bytes(n, K, N, k));
        return LYS_EWORKSPACE;
    }
    const int64_t chunk = generic_chunk(K, N);
    unsigned char* p = reinterpret_cast<unsigned char*>(workspace);
    float* alpha = reinterpret_cast<float*>(p); p += align_up((size_t)chunk * (size_t)K * sizeof(float), 256);
    void* planes = p; p += align_up(corr_gemm_tc_planes_bytes(n, K), 256);
    float* xn2 = reinterpret_cast<float*>(p);
    const bool use_tc = corr_gemm_tc_supported(n, K);
    int rc = LYS_OK;
    if (use_tc && (rc = corr_gemm_tc_prepare(D, ldd, n, K, planes, stream))) return rc;
    for (int64_t s0 = 0; s0 < N; s0 += chunk) {
        const int64_t C = std::min(chunk, N - s0);
        if (use_tc) rc = corr_gemm_tc(X + s0 * xss, xfs, xss, planes, n, K, C, alpha, stream);
        else rc = sgemm_strided(X + s0 * xss, xss, xfs, D, ldd, 1, alpha, K, 1, C, K, n, stream);
        if (rc) return rc;
        if ((rc = col_norm2(X + s0 * xss, xfs, xss, n, C, xn2, stream))) return rc;
        rc = omp_greedy_generic(alpha, G, K, C, k, xn2, tol, strict, idx + s0 * k, val + s0 * k, nsel ? nsel + s0 : nullptr,
                                Z ? Z + s0 * zss : nullptr, zas, zss, truncated, stream);
        if (rc) return rc;
    }
    return LYS_OK;
}

extern "C" int lys_bomp_launch_count(int n, int K, int64_t N, int k)
{
    if (n < 1 || K < 1 || N < 1 || k < 1) return 0;
    int fused = bomp_fused_launch_count(n, K, N, k);
    if (fused > 0) return fused;
    const int64_t chunk = generic_chunk(K, N);
    const int halves = corr_gemm_tc_supported(n, K) ? (n > 64 ? 2 : 1) : 0;      // tensor-core GEMM launches per chunk
    return (int)(((halves ? halves : 1) + 1) * ((N + chunk - 1) / chunk)) + halves;
}

extern "C" int lys_profile_enable(int on)
{
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.on = on != 0;
    return LYS_OK;
}

extern "C" int lys_profile_fetch(double* kernel_ms, int64_t* launches, const char** kernel_name, int reset)
{
    std::lock_guard<std::mutex> lk(g_prof_mu);
    double total = 0.0;
    for (auto& ev : g_prof.used) {
        LYS_CUDA(cudaEventSynchronize(ev.second));
        float ms = 0.f;
        LYS_CUDA(cudaEventElapsedTime(&ms, ev.first, ev.second));
        total += ms;
    }
    if (kernel_ms) *kernel_ms = total;
    if (launches) *launches = (int64_t)g_prof.used.size();
    if (kernel_name) *kernel_name = g_prof.name;
    if (reset) {
        for (auto& ev : g_prof.used) g_prof.pool.push_back(ev);
        g_prof.used.clear();
    }
    return LYS_OK;
}

extern "C" int lys_corr_gemm(const float* X, int64_t xfs, int64_t xss, const float* D, int64_t ldd,
                             int n, int K, int64_t C, float* alpha, int impl, void* stream)
{
    LYS_CHECK_ARG(X && D && alpha && n >= 1 && K >= 1 && C >= 0 && ldd >= K, "lys_corr_gemm: bad argument");
    if (impl == 0) impl = corr_gemm_tc_supported(n, K) ? 2 : 1;
    if (impl == 1) return sgemm_strided(X, xss, xfs, D, ldd, 1, alpha, K, 1, C, K, n, (cudaStream_t)stream);
    if (impl == 2) {
        if (!corr_gemm_tc_supported(n, K)) { set_error("lys_corr_gemm: tcgen05 path needs n=64 or 128, K multiple of 256"); return LYS_EUNSUPPORTED; }
        static void* hook_planes[64] = {nullptr};          // bring-up hook only: cached per device, never freed
        int dev = 0;
        LYS_CUDA(cudaGetDevice(&dev));
        if (!hook_planes[dev & 63]) LYS_CUDA(cudaMalloc(&hook_planes[dev & 63], corr_gemm_tc_planes_bytes(NF_MAX_HOOK, LYS_MAX_ATOMS)));
        int rc2 = corr_gemm_tc_prepare(D, ldd, n, K, hook_planes[dev & 63], (cudaStream_t)stream);
        if (rc2) return rc2;
        return corr_gemm_tc(X, xfs, xss, hook_planes[dev & 63], n, K, C, alpha, (cudaStream_t)stream);
    }
    set_error("lys_corr_gemm: unknown impl %d", impl);
    return LYS_EINVAL;
}

extern "C" int lys_codes_to_dense(const int32_t* idx, const float* val, int64_t N, int k, int K,
                                  float* Z, int64_t zas, int64_t zss, void* stream)
{
    LYS_CHECK_ARG(idx && val && Z, "lys_codes_to_dense: null pointer");
    LYS_CHECK_ARG(k >= 1 && k <= K && K >= 1 && N >= 0, "lys_codes_to_dense: bad shape");
    if (N == 0) return LYS_OK;
    int64_t blocks = std::min<int64_t>((N + 7) / 8, (int64_t)sm_count() * 8);
    dense_fill_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(idx, val, N, k, K, Z, zas, zss);
    LYS_LAUNCH_CHECK("dense_fill_kernel");
    return LYS_OK;
}

// --------------------------------------------------------------------------- host pipeline
namespace {

// One pipeline per device: staging buffers, two copy/compute streams and events are created once and reused.  Each
// pipeline has its own lock, so sparse_encoder(n_jobs=G).encode(numpy) — one host thread per GPU
// (lyssa/utils/__init__.py:92-146 forks one worker per job) — drives G devices concurrently; two callers that target the
// SAME device take turns.
struct HostPipe {
    std::mutex mu;
    bool ready = false;
    size_t bytes = 0;
    unsigned char* base = nullptr;
    cudaStream_t streams[2] = {nullptr, nullptr};
    cudaEvent_t dict_ready = nullptr;
    size_t stage_bytes = 0;                  // pinned host staging for the sparse codes of one call
    unsigned char* stage = nullptr;
    std::vector<cudaEvent_t> chunk_done;     // one event per chunk, grow-only
};
constexpr int kMaxDevices = 64;
HostPipe g_pipes[kMaxDevices];

// caller holds p.mu and has made `device` current
int pipe_reserve(HostPipe& p, size_t bytes, size_t stage_bytes, size_t n_events)
{
    if (!p.ready) {
        for (int s = 0; s < 2; ++s) LYS_CUDA(cudaStreamCreateWithFlags(&p.streams[s], cudaStreamNonBlocking));
        LYS_CUDA(cudaEventCreateWithFlags(&p.dict_ready, cudaEventDisableTiming));
        p.ready = true;
    }
    if (p.bytes < bytes) {
        if (p.base) { LYS_CUDA(cudaFree(p.base)); p.base = nullptr; p.bytes = 0; }
        LYS_CUDA(cudaMalloc(&p.base, bytes));
        p.bytes = bytes;
    }
    if (p.stage_bytes < stage_bytes) {
        if (p.stage) { LYS_CUDA(cudaFreeHost(p.stage)); p.stage = nullptr; p.stage_bytes = 0; }
        LYS_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&p.stage), stage_bytes, cudaHostAllocDefault));
        p.stage_bytes = stage_bytes;
    }
    while (p.chunk_done.size() < n_events) {
        cudaEvent_t ev;
        LYS_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        p.chunk_done.push_back(ev);
    }
    return LYS_OK;
}

// on any failure inside the chunk loop: no async copy may still be writing into the caller's buffers when we return
int pipe_fail(HostPipe& p, int rc)
{
    cudaStreamSynchronize(p.streams[0]);
    cudaStreamSynchronize(p.streams[1]);
    return rc;
}

#define LYS_PIPE_CUDA(pipe, call)                                                            \
    do {                                                                                     \
        cudaError_t _e = (call);                                                             \
        if (_e != cudaSuccess) {                                                             \
            ::lys::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return pipe_fail(pipe, LYS_ECUDA);                                               \
        }                                                                                    \
    } while (0)

// Dense rows on the HOST.  The reference's contract returns the dense (K x N) code matrix, 99.5 % zeros at cfg2: sending
// it over PCIe costs 4.3 GB per million patches.  Instead the device returns the sparse codes (46 MB) and host threads of
// this library write the rows of Z — the zero-fill + scatter `Z[Dx, i] = z` of sparse_coding.py:308,:365 and
// utils/__init__.py:69,89, nothing numerical — chunk by chunk while the device encodes the next chunks: a row is composed
// in a cache-resident scratch (zeros + its k coefficients) and leaves with non-temporal 16-byte stores, so every byte of Z
// is written once and never read.
struct DensifyJob {
    const int32_t* idx; const float* val;      // (N, k) pinned staging
    float* Z; int64_t zss; int K, k;
    int64_t chunk, N;
    std::atomic<int64_t> ready{0};             // chunks whose codes have landed
    std::atomic<bool> abort{false};
};

void densify_worker(DensifyJob* job, int w, int T)
{
    const int K = job->K, k = job->k;
    std::vector<float> scratch_v((size_t)K + 16, 0.f);
    float* scratch = scratch_v.data();
    while ((reinterpret_cast<uintptr_t>(scratch) & 15) != 0) ++scratch;
    const int64_t n_chunks = (job->N + job->chunk - 1) / job->chunk;
    for (int64_t c = 0; c < n_chunks; ++c) {
        while (job->ready.load(std::memory_order_acquire) <= c) {
            if (job->abort.load(std::memory_order_relaxed)) return;
            std::this_thread::sleep_for(std::chrono::microseconds(20));
        }
        const int64_t c0 = c * job->chunk, C = std::min(job->chunk, job->N - c0);
        const int64_t r0 = c0 + C * w / T, r1 = c0 + C * (w + 1) / T;
        for (int64_t i = r0; i < r1; ++i) {
            const int32_t* ii = job->idx + i * k;
            const float* vv = job->val + i * k;
            for (int j = 0; j < k; ++j) if (ii[j] >= 0 && ii[j] < K) scratch[ii[j]] = vv[j];
            float* row = job->Z + i * job->zss;
            if (((reinterpret_cast<uintptr_t>(row) & 15) == 0) && (K % 4) == 0) {
                for (int c4 = 0; c4 < K; c4 += 4) _mm_stream_ps(row + c4, _mm_load_ps(scratch + c4));
            } else {
                memcpy(row, scratch, sizeof(float) * (size_t)K);
            }
            for (int j = 0; j < k; ++j) if (ii[j] >= 0 && ii[j] < K) scratch[ii[j]] = 0.f;
        }
    }
    _mm_sfence();
}

int host_threads()
{
    const unsigned hw = std::thread::hardware_concurrency();
    return (int)std::max(1u, std::min(16u, hw > 1 ? hw - 1 : 1u));
}

}  // namespace

extern "C" int lys_bomp_encode_host(const float* X, int64_t xfs, int64_t xss,
                                    const float* D, int64_t ldd,
                                    int n, int K, int64_t N, int k,
                                    int32_t* idx, float* val, int32_t* nsel,
                                    float* Z, int64_t zas, int64_t zss, int device)
{
    LYS_CHECK_ARG(X && D, "lys_bomp_encode_host: null input");
    LYS_CHECK_ARG(n >= 1 && n <= LYS_MAX_FEATURES && K >= 1 && K <= LYS_MAX_ATOMS && ldd >= K,
                  "lys_bomp_encode_host: bad dictionary shape");
    LYS_CHECK_ARG(k >= 1 && k <= LYS_MAX_NONZERO && k <= K, "lys_bomp_encode_host: bad n_nonzero_coefs=%d", k);
    LYS_CHECK_ARG((xfs == 1 && xss >= n) || (xss == 1 && xfs >= N),
                  "lys_bomp_encode_host: X must be signal-major (feat stride 1) or feature-major (signal stride 1)");
    LYS_CHECK_ARG(!Z || (zas == 1 && zss >= K) || (zss == 1 && zas >= N),
                  "lys_bomp_encode_host: Z must be signal-major (atom stride 1) or atom-major (signal stride 1)");
    if (N == 0) return LYS_OK;
    if (device >= 0) LYS_CUDA(cudaSetDevice(device));
    else LYS_CUDA(cudaGetDevice(&device));
    LYS_CHECK_ARG(device < kMaxDevices, "lys_bomp_encode_host: device %d out of range", device);
    HostPipe& pipe = g_pipes[device];
    std::lock_guard<std::mutex> lock(pipe.mu);

    const bool x_sig_major = (xfs == 1);
    const bool z_sig_major = (zas == 1);
    const bool host_dense = Z && z_sig_major;         // rows of Z written by host threads from the sparse codes
    const bool dev_dense = Z && !z_sig_major;         // atom-major Z: dense block from the device (strided host scatter would thrash)
    const int64_t chunk = std::min<int64_t>(N, 32768);
    const int64_t n_chunks = (N + chunk - 1) / chunk;
    const size_t ws_bytes = align_up(lys_bomp_workspace_bytes(n, K, chunk, k), 256);
    const size_t d_bytes = align_up((size_t)n * K * 4, 256), g_bytes = align_up((size_t)K * K * 4, 256);
    const size_t x_bytes = align_up((size_t)chunk * n * 4, 256);
    const size_t i_bytes = align_up((size_t)chunk * k * 4, 256);
    const size_t s_bytes = align_up((size_t)chunk * 4, 256);
    const size_t z_bytes = dev_dense ? align_up((size_t)chunk * K * 4, 256) : 0;
    const size_t slot_bytes = x_bytes + 2 * i_bytes + s_bytes + z_bytes + ws_bytes;
    const size_t codes_bytes = align_up((size_t)N * k * 4, 256);
    int rc = pipe_reserve(pipe, d_bytes + g_bytes + 2 * slot_bytes, host_dense ? 2 * codes_bytes : 0, (size_t)n_chunks);
    if (rc) return rc;

    unsigned char* p = pipe.base;
    float* dD = reinterpret_cast<float*>(p); p += d_bytes;
    float* dG = reinterpret_cast<float*>(p); p += g_bytes;
    struct Slot { float* x; int32_t* idx; float* val; int32_t* nsel; float* z; void* ws; } slot[2];
    for (int s = 0; s < 2; ++s) {
        slot[s].x = reinterpret_cast<float*>(p); p += x_bytes;
        slot[s].idx = reinterpret_cast<int32_t*>(p); p += i_bytes;
        slot[s].val = reinterpret_cast<float*>(p); p += i_bytes;
        slot[s].nsel = reinterpret_cast<int32_t*>(p); p += s_bytes;
        slot[s].z = dev_dense ? reinterpret_cast<float*>(p) : nullptr; p += z_bytes;
        slot[s].ws = p; p += ws_bytes;
    }
    int32_t* st_idx = host_dense ? reinterpret_cast<int32_t*>(pipe.stage) : nullptr;
    float* st_val = host_dense ? reinterpret_cast<float*>(pipe.stage + codes_bytes) : nullptr;

    cudaStream_t s0 = pipe.streams[0], s1 = pipe.streams[1];
    LYS_PIPE_CUDA(pipe, cudaMemcpy2DAsync(dD, (size_t)K * 4, D, (size_t)ldd * 4, (size_t)K * 4, n, cudaMemcpyHostToDevice, s0));
    rc = lys_gram(dD, K, n, K, dG, s0);
    if (rc) return pipe_fail(pipe, rc);
    LYS_PIPE_CUDA(pipe, cudaEventRecord(pipe.dict_ready, s0));
    LYS_PIPE_CUDA(pipe, cudaStreamWaitEvent(s1, pipe.dict_ready, 0));

    DensifyJob job;
    std::vector<std::thread> workers;
    if (host_dense) {
        job.idx = st_idx; job.val = st_val; job.Z = Z; job.zss = zss; job.K = K; job.k = k; job.chunk = chunk; job.N = N;
        const int T = host_threads();
        for (int w = 0; w < T; ++w) workers.emplace_back(densify_worker, &job, w, T);
    }
    auto stop_workers = [&](bool abort) {
        if (abort) job.abort.store(true);
        for (auto& t : workers) t.join();
        workers.clear();
    };

    int which = 0;
    int64_t issued = 0;
    for (int64_t c0 = 0; c0 < N; c0 += chunk, which ^= 1, ++issued) {
        const int64_t C = std::min(chunk, N - c0);
        cudaStream_t st = pipe.streams[which];
        Slot& sl = slot[which];
        int64_t dxfs, dxss;
        cudaError_t e;
        if (x_sig_major) {       // host rows of n floats, row stride xss -> device (C, n)
            e = cudaMemcpy2DAsync(sl.x, (size_t)n * 4, X + c0 * xss, (size_t)xss * 4, (size_t)n * 4, C, cudaMemcpyHostToDevice, st);
            dxfs = 1; dxss = n;
        } else {                 // host (n, N) feature-major -> device (n, C)
            e = cudaMemcpy2DAsync(sl.x, (size_t)C * 4, X + c0, (size_t)xfs * 4, (size_t)C * 4, n, cudaMemcpyHostToDevice, st);
            dxfs = C; dxss = 1;
        }
        rc = (e == cudaSuccess) ? LYS_OK : LYS_ECUDA;
        if (!rc) rc = lys_bomp_encode(sl.x, dxfs, dxss, dD, K, dG, n, K, C, k, sl.idx, sl.val, sl.nsel,
                                      sl.z, C, 1, sl.ws, ws_bytes, st);
        if (!rc && host_dense) {
            if (cudaMemcpyAsync(st_idx + c0 * k, sl.idx, (size_t)C * k * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                cudaMemcpyAsync(st_val + c0 * k, sl.val, (size_t)C * k * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = LYS_ECUDA;
        }
        if (!rc && !host_dense) {
            if (idx && cudaMemcpyAsync(idx + c0 * k, sl.idx, (size_t)C * k * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = LYS_ECUDA;
            if (val && cudaMemcpyAsync(val + c0 * k, sl.val, (size_t)C * k * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = LYS_ECUDA;
        }
        if (!rc && nsel && cudaMemcpyAsync(nsel + c0, sl.nsel, (size_t)C * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = LYS_ECUDA;
        if (!rc && dev_dense &&
            cudaMemcpy2DAsync(Z + c0, (size_t)zas * 4, sl.z, (size_t)C * 4, (size_t)C * 4, K, cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = LYS_ECUDA;
        if (!rc && cudaEventRecord(pipe.chunk_done[issued], st) != cudaSuccess) rc = LYS_ECUDA;
        if (rc) {
            if (rc == LYS_ECUDA) set_error("lys_bomp_encode_host: a copy or launch of chunk %lld failed: %s", (long long)issued, cudaGetErrorString(cudaGetLastError()));
            stop_workers(true);
            return pipe_fail(pipe, rc);
        }
        // hand finished chunks to the host threads while later chunks are still being issued (two chunks stay in flight)
        if (host_dense && issued >= 2) {
            if (cudaEventSynchronize(pipe.chunk_done[issued - 2]) != cudaSuccess) { stop_workers(true); set_error("lys_bomp_encode_host: chunk failed"); return pipe_fail(pipe, LYS_ECUDA); }
            job.ready.store(issued - 1, std::memory_order_release);
        }
    }
    if (host_dense) {
        for (int64_t c = std::max<int64_t>(0, issued - 2); c < issued; ++c) {
            if (cudaEventSynchronize(pipe.chunk_done[c]) != cudaSuccess) { stop_workers(true); set_error("lys_bomp_encode_host: chunk failed"); return pipe_fail(pipe, LYS_ECUDA); }
            job.ready.store(c + 1, std::memory_order_release);
        }
        stop_workers(false);
        if (idx) memcpy(idx, st_idx, (size_t)N * k * 4);
        if (val) memcpy(val, st_val, (size_t)N * k * 4);
    }
    LYS_PIPE_CUDA(pipe, cudaStreamSynchronize(s0));
    LYS_PIPE_CUDA(pipe, cudaStreamSynchronize(s1));
    return LYS_OK;
}
