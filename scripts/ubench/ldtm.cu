// Micro-benchmark: tcgen05.ld (TMEM -> registers) throughput on sm_100a, 32x32b shape.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldtm ldtm.cu && ./ldtm
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define LD_X32(taddr, r)                                                                                      \
    asm volatile(                                                                                             \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                             \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                             \
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"             \
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),     \
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), \
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), \
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) \
        : "r"(taddr))
#define LD_X16(taddr, r)                                                                                      \
    asm volatile(                                                                                             \
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                             \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"                      \
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),     \
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) \
        : "r"(taddr))

template <int X, int DEPTH>
__global__ void k_ldtm(uint32_t* out, long long* cyc, int iters)
{
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint32_t r[DEPTH][32];
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            const uint32_t col = (uint32_t)(((it * DEPTH + d) * X) & 511);
            if (X == 32) { LD_X32(base + col, r[d]); } else { LD_X16(base + col, r[d]); }
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) acc ^= r[d][0] ^ r[d][X - 1];
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512));
}

template <int X, int DEPTH> void run(int warps)
{
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 2048;
    k_ldtm<X, DEPTH><<<148, warps * 32>>>(out, cyc, iters);
    k_ldtm<X, DEPTH><<<148, warps * 32>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    const double bytes_per_warp = (double)iters * DEPTH * X * 32 * 4;
    printf("x%-3d depth %d warps %2d: %9.0f cycles  -> %6.1f B/clk per warp, %7.1f B/clk per SM  (%s)\n", X, DEPTH, warps, avg,
           bytes_per_warp / avg, bytes_per_warp * warps / avg, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    for (int w : {1, 4, 8, 16}) { run<32, 1>(w); run<32, 2>(w); run<16, 2>(w); run<16, 4>(w); }
    return 0;
}
