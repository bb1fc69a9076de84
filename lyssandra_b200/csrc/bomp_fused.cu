// bomp_fused.cu — fused correlation + greedy Batch-OMP kernel for the benchmark shapes.
// (placeholder until the tcgen05 path lands: every shape is reported unsupported so that
// lys_bomp_encode takes the generic path.)
#include "common.cuh"

namespace lys {

size_t bomp_fused_workspace_bytes(int, int, int64_t, int) { return 0; }
int bomp_fused_launch_count(int, int, int64_t, int) { return 0; }

int bomp_encode_fused(const float*, int64_t, int64_t, const float*, int64_t, const float*,
                      int, int, int64_t, int, int32_t*, float*, int32_t*, float*, int64_t, int64_t,
                      void*, size_t, cudaStream_t)
{
    return LYS_EUNSUPPORTED;
}

}  // namespace lys
