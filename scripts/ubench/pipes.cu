// Micro-benchmark of sm_100a issue rates for the instructions the TMEM scan is built from.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define REP8(x) x x x x x x x x
#define BODY(NAME, ASM)                                                                      \
    __global__ void NAME(float* out, long long* cyc, float seed)                             \
    {                                                                                        \
        float a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4,   \
              a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7, b = seed * 0.5f, c = seed * 0.25f;       \
        long long t0 = clock64();                                                            \
        _Pragma("unroll 1") for (int it = 0; it < 64; ++it) {                                                  \
            REP8(ASM)                                                                        \
        }                                                                                    \
        long long t1 = clock64();                                                            \
        out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;  \
        if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                                     \
    }

#define OPS8(OP) \
    asm volatile(OP(%0) OP(%1) OP(%2) OP(%3) OP(%4) OP(%5) OP(%6) OP(%7) \
        : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7) : "f"(b), "f"(c));

#define FMNMX3_OP(r) "max.f32 " #r ", " #r ", %8, %9;\n"       /* 3-input max (ptxas fuses? no: explicit below) */
#define FMAX3_OP(r) "max.f32 " #r ", " #r ", %8, %9;\n"
#define FMAX2_OP(r) "max.f32 " #r ", " #r ", %8;\n"
#define FADD_OP(r) "add.f32 " #r ", " #r ", %8;\n"
#define FMUL_OP(r) "mul.f32 " #r ", " #r ", %8;\n"
#define FFMA_OP(r) "fma.rn.f32 " #r ", " #r ", %8, %9;\n"
#define FFMAI_OP(r) "fma.rn.f32 " #r ", " #r ", %8, 0f40400000;\n"
#define LOP_OP(r) "{.reg .b32 t; mov.b32 t, " #r "; xor.b32 t, t, 0x12345; mov.b32 " #r ", t;}\n"
#define IADD_OP(r) "{.reg .b32 t; mov.b32 t, " #r "; add.s32 t, t, 77; mov.b32 " #r ", t;}\n"
#define IMAD_OP(r) "{.reg .b32 t; mov.b32 t, " #r "; mad.lo.s32 t, t, 3, 7; mov.b32 " #r ", t;}\n"

BODY(k_max3, OPS8(FMAX3_OP))
BODY(k_max2, OPS8(FMAX2_OP))
BODY(k_fadd, OPS8(FADD_OP))
BODY(k_fmul, OPS8(FMUL_OP))
BODY(k_ffma, OPS8(FFMA_OP))
BODY(k_ffmai, OPS8(FFMAI_OP))
BODY(k_lop, OPS8(LOP_OP))
BODY(k_iadd, OPS8(IADD_OP))
BODY(k_imad, OPS8(IMAD_OP))
// mixes: 8 ALU + 8 FMA-pipe ops interleaved
#define MIX_OP(r) "max.f32 " #r ", " #r ", %8, %9;\n" "mul.f32 " #r ", " #r ", %8;\n"
BODY(k_mix_max3_fmul, OPS8(MIX_OP))

template <typename K> void run(const char* name, K kern, int warps_per_sm, int ops_per_iter)
{
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    kern<<<148, warps_per_sm * 32>>>(out, cyc, 1.5f);
    kern<<<148, warps_per_sm * 32>>>(out, cyc, 1.5f);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    const double n_instr = 64.0 * 8 * ops_per_iter;             // per warp
    printf("%-18s warps/SM %2d: %8.0f cycles, %.2f cycles per warp-instr per SMSP (%.2f warps/SMSP)\n", name, warps_per_sm, avg,
           avg / (n_instr * (warps_per_sm / 4.0 > 1 ? warps_per_sm / 4.0 : 1)), warps_per_sm / 4.0);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    for (int w : {4, 8, 16}) {
        run("max3", k_max3, w, 8); run("max2", k_max2, w, 8); run("fadd", k_fadd, w, 8); run("fmul", k_fmul, w, 8);
        run("ffma", k_ffma, w, 8); run("ffma imm", k_ffmai, w, 8); run("lop3", k_lop, w, 8); run("iadd", k_iadd, w, 8);
        run("imad", k_imad, w, 8); run("max3+fmul", k_mix_max3_fmul, w, 16);
    }
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
