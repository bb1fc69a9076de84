// sweep_common.cuh — helpers shared by the approximate (ksvd_sweep.cu) and exact (ksvd_exact.cu) K-SVD sweeps.
#pragma once
#include "common.cuh"

namespace lys {

using u64 = unsigned long long;

__device__ __forceinline__ u64 ld_relaxed_gpu(const u64* p)
{
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ u64 ld_relaxed_sys(const u64* p)
{
    u64 v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(u64* p, u64 v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void red_add(u64* p, u64 v)
{
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// entry id -> signal index: ent / k by multiply-shift (k <= 32, ent < 2^31; M = ceil(2^(32+s)/k) is exact
// for every 32-bit numerator because M*k - 2^(32+s) < k <= 2^s)
struct FastDiv { unsigned long long M; int s; };
__device__ __forceinline__ int fdiv(int ent, FastDiv d)
{
    return (int)(((unsigned long long)(unsigned)ent * d.M) >> (32 + d.s));
}

inline FastDiv make_fastdiv(int k)
{
    FastDiv d;
    d.s = 0;
    while ((1 << d.s) < k) ++d.s;
    d.M = ((1ull << (32 + d.s)) + (unsigned long long)k - 1) / (unsigned long long)k;
    return d;
}

// one CTA per SM, at most 255 (the contribution counter of an accumulator word has 8 bits)
int sweep_grid_size();
// bounds[c][b] = first CSR position of atom c whose signal is >= b*S, S = ceil(N / grid): CTA b owns signals [bS, (b+1)S)
int sweep_bounds(const int32_t* rowptr, const int32_t* entries, int K, int k, int grid, int64_t N, int32_t* bounds, cudaStream_t stream);

}  // namespace lys
