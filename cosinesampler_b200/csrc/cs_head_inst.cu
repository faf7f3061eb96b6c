// cs_head_inst.cu -- instantiations of the fused PDE-residual head: the tensor-core kernel
// (cs_head_mma.cuh) for C in {8, 16, 32}, the SIMT kernel (cs_head.cuh) for C = 4 and, for
// comparison, for every C when COSINE_SAMPLER_HEAD=simt is set in the environment.
#include <cstdlib>
#include <cstring>

#include "cs_head_mma.cuh"

namespace cs {
cudaError_t launch_head_any(int dim, int C, const HeadParams& p, cudaStream_t s) {
    static const bool simt = [] { const char* e = getenv("COSINE_SAMPLER_HEAD"); return e && !strcmp(e, "simt"); }();
    if (!simt) {
#define CS_HEAD_MMA_CASE(D, CC) if (dim == D && C == CC) return launch_head_mma<D, CC>(p, s);
        CS_HEAD_MMA_CASE(2, 8) CS_HEAD_MMA_CASE(2, 16) CS_HEAD_MMA_CASE(2, 32)
        CS_HEAD_MMA_CASE(3, 8) CS_HEAD_MMA_CASE(3, 16) CS_HEAD_MMA_CASE(3, 32)
#undef CS_HEAD_MMA_CASE
    }
#define CS_HEAD_CASE(D, CC) if (dim == D && C == CC) return launch_head<D, CC>(p, s);
    CS_HEAD_CASE(2, 4) CS_HEAD_CASE(2, 8) CS_HEAD_CASE(2, 16) CS_HEAD_CASE(2, 32)
    CS_HEAD_CASE(3, 4) CS_HEAD_CASE(3, 8) CS_HEAD_CASE(3, 16) CS_HEAD_CASE(3, 32)
#undef CS_HEAD_CASE
    return cudaErrorInvalidValue;
}
}  // namespace cs
