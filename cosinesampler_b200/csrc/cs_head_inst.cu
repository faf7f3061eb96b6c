// cs_head_inst.cu -- instantiations of the fused PDE-residual head (cs_head.cuh).
#include "cs_head.cuh"

namespace cs {
cudaError_t launch_head_any(int dim, int C, const HeadParams& p, cudaStream_t s) {
#define CS_HEAD_CASE(D, CC) if (dim == D && C == CC) return launch_head<D, CC>(p, s);
    CS_HEAD_CASE(2, 4) CS_HEAD_CASE(2, 8) CS_HEAD_CASE(2, 16) CS_HEAD_CASE(2, 32)
    CS_HEAD_CASE(3, 4) CS_HEAD_CASE(3, 8) CS_HEAD_CASE(3, 16) CS_HEAD_CASE(3, 32)
#undef CS_HEAD_CASE
    return cudaErrorInvalidValue;
}
}  // namespace cs
