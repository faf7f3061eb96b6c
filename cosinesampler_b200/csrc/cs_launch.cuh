// cs_launch.cuh -- kernel selection and launch geometry for one dimensionality.
// Included by cs_stage_2d.cu / cs_stage_3d.cu (one translation unit per DIM so
// the two compile in parallel).
#pragma once
#include "cs_engine.cuh"

namespace cs {

struct LaunchInfo {
    int sm_count;
};

inline int device_sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (cached[dev] == 0) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        cached[dev] = n > 0 ? n : 148;
    }
    return cached[dev];
}

template <int DIM, int VEC, int STAGE, bool HAS_U, bool HAS_X2>
cudaError_t launch_one(const StageParams& p, cudaStream_t stream) {
    using RL = RecLayout<DIM, STAGE, HAS_X2>;
    auto kern = cs_stage_kernel<DIM, VEC, STAGE, HAS_U, HAS_X2>;
    constexpr int threads = 256;
    constexpr int wpb = threads / 32;
    const int pts = 128 >> p.lshift;
    const size_t smem = (size_t)wpb * RL::FIELDS * pts * sizeof(float);
    // per (device, kernel) one-time setup: opt in to >48 KiB and query occupancy
    struct PerDevice { size_t smem; int occ[4]; };
    static PerDevice cfg[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    PerDevice& c = cfg[dev];
    if (smem > c.smem) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        c.smem = smem;
    }
    int& occ = c.occ[p.lshift & 3];
    if (occ == 0) {
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem);
        if (e != cudaSuccess) return e;
        if (occ < 1) occ = 1;
    }
    const long long total_tiles = p.num_ptiles * p.N;
    long long blocks = (total_tiles + wpb - 1) / wpb;
    const long long cap = (long long)device_sm_count() * occ;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) return cudaSuccess;
    kern<<<(unsigned)blocks, threads, smem, stream>>>(p);
    return cudaGetLastError();
}

template <int DIM, int VEC>
cudaError_t launch_stage_vec(int stage, bool has_u, bool has_x2, const StageParams& p, cudaStream_t s) {
    switch (stage) {
        case ST_F: return launch_one<DIM, VEC, ST_F, false, false>(p, s);
        case ST_B: return launch_one<DIM, VEC, ST_B, false, false>(p, s);
        case ST_BB:
            return has_u ? launch_one<DIM, VEC, ST_BB, true, false>(p, s)
                         : launch_one<DIM, VEC, ST_BB, false, false>(p, s);
        case ST_BBB:
            return has_x2 ? launch_one<DIM, VEC, ST_BBB, false, true>(p, s)
                          : launch_one<DIM, VEC, ST_BBB, false, false>(p, s);
    }
    return cudaErrorInvalidValue;
}

template <int DIM>
cudaError_t launch_stage_dim(int vec, int stage, bool has_u, bool has_x2, const StageParams& p,
                             cudaStream_t s) {
    return vec == 4 ? launch_stage_vec<DIM, 4>(stage, has_u, has_x2, p, s)
                    : launch_stage_vec<DIM, 1>(stage, has_u, has_x2, p, s);
}

}  // namespace cs
