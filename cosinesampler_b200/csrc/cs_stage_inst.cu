// cs_stage_inst.cu -- one translation unit per (CS_DIM, CS_VEC, CS_LSHIFT) variant of the
// stage engine; compiled several times with different -D flags so the variants build in
// parallel (see _build.py).  Exports cs::launch_d<DIM>_v<VEC>_l<LSHIFT>.
#include "cs_launch.cuh"

#ifndef CS_DIM
#error "compile with -DCS_DIM=2|3 -DCS_VEC=4|1 -DCS_LSHIFT=0..3"
#endif
#define CS_CAT_(a, b, c, d, e, f) a##b##c##d##e##f
#define CS_CAT(a, b, c, d, e, f) CS_CAT_(a, b, c, d, e, f)
#define CS_FN CS_CAT(launch_d, CS_DIM, _v, CS_VEC, _l, CS_LSHIFT)

namespace cs {
cudaError_t CS_FN(int stage, bool has_u, bool has_x2, const StageParams& p, cudaStream_t s) {
    return launch_variant<CS_DIM, CS_VEC, CS_LSHIFT>(stage, has_u, has_x2, p, s);
}
}  // namespace cs
