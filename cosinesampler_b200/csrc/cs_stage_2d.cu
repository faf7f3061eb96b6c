// 2D instantiations of the stage engine.
#include "cs_launch.cuh"
namespace cs {
cudaError_t launch_stage_2d(int vec, int stage, bool has_u, bool has_x2, const StageParams& p, cudaStream_t s) {
    return launch_stage_dim<2>(vec, stage, has_u, has_x2, p, s);
}
}  // namespace cs
