// cs_jet_inst.cu -- one translation unit per (CS_DIM, CS_LSHIFT) variant of the jet kernels
// (cs_jet.cuh); compiled several times with different -D flags so the variants build in
// parallel (see _build.py).  Exports cs::launch_jet_d<DIM>_l<LSHIFT>.
#include "cs_jet.cuh"

#ifndef CS_DIM
#error "compile with -DCS_DIM=2|3 -DCS_LSHIFT=0..3"
#endif
#define CS_CAT_(a, b, c, d) a##b##c##d
#define CS_CAT(a, b, c, d) CS_CAT_(a, b, c, d)
#define CS_FN CS_CAT(launch_jet_d, CS_DIM, _l, CS_LSHIFT)

namespace cs {
cudaError_t CS_FN(int order, bool backward, JetParams& p, cudaStream_t s) {
    return launch_jet_variant<CS_DIM, CS_LSHIFT>(order, backward, p, s);
}
}  // namespace cs
