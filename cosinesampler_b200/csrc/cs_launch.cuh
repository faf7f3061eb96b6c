// cs_launch.cuh -- kernel selection and launch geometry for one (DIM, VEC, LSHIFT) variant.
#pragma once
#include "cs_engine.cuh"

#ifndef CS_SMEM_TARGET
#define CS_SMEM_TARGET (72 * 1024)
#endif

namespace cs {

inline int device_index() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev < 0 || dev >= 64) ? 0 : dev;
}

inline int device_sm_count(int dev) {
    static int cached[64] = {0};
    if (cached[dev] == 0) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        cached[dev] = n > 0 ? n : 148;
    }
    return cached[dev];
}

template <int DIM, int VEC, int LSHIFT, int STAGE, bool HAS_U, bool HAS_X2>
cudaError_t launch_one(const StageParams& p, cudaStream_t stream) {
    using WS = WarpSmem<DIM, VEC, LSHIFT, STAGE, HAS_U, HAS_X2>;
    auto kern = cs_stage_kernel<DIM, VEC, LSHIFT, STAGE, HAS_U, HAS_X2>;
    constexpr size_t smem_per_warp = (size_t)WS::TOTAL * sizeof(float4);
    // block size: as many warps as fit CS_SMEM_TARGET bytes of shared memory, at most
    // CS_THREADS threads; several blocks then share an SM
    int warps = CS_THREADS / 32;
    while (warps > 1 && warps * smem_per_warp > CS_SMEM_TARGET) warps >>= 1;
    const int threads = warps * 32;
    const size_t smem = warps * smem_per_warp;
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    // per (device, kernel) one-time setup: opt in to >48 KiB and query occupancy
    static int occ_cache[64] = {0};
    const int dev = device_index();
    int& occ = occ_cache[dev];
    if (occ == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem);
        if (e != cudaSuccess) return e;
        if (occ < 1) occ = 1;
    }
    const long long total_tiles = p.num_ptiles * p.N;
    long long blocks = (total_tiles + warps - 1) / warps;
    const long long cap = (long long)device_sm_count(dev) * occ;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) return cudaSuccess;
    kern<<<(unsigned)blocks, threads, smem, stream>>>(p);
    return cudaGetLastError();
}

template <int DIM, int VEC, int LSHIFT>
cudaError_t launch_variant(int stage, bool has_u, bool has_x2, const StageParams& p, cudaStream_t s) {
    switch (stage) {
        case ST_F: return launch_one<DIM, VEC, LSHIFT, ST_F, false, false>(p, s);
        case ST_B: return launch_one<DIM, VEC, LSHIFT, ST_B, false, false>(p, s);
        case ST_BB:
            return has_u ? launch_one<DIM, VEC, LSHIFT, ST_BB, true, false>(p, s)
                         : launch_one<DIM, VEC, LSHIFT, ST_BB, false, false>(p, s);
        case ST_BBB:
            return has_x2 ? launch_one<DIM, VEC, LSHIFT, ST_BBB, false, true>(p, s)
                          : launch_one<DIM, VEC, LSHIFT, ST_BBB, false, false>(p, s);
    }
    return cudaErrorInvalidValue;
}

}  // namespace cs
