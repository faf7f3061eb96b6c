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

// Shared-memory small-cell kernel (cs_small_kernel): taken when the fields of one cell that this
// call touches, plus the per-warp records, fit comfortably in shared memory.
template <int DIM, int VEC, int LSHIFT, int STAGE, bool HAS_U, bool HAS_X2>
bool try_launch_small(const StageParams& p, cudaStream_t stream, cudaError_t& err) {
    if (VEC != 4 || p.small_cell == 1) return false;
    using RL = RecLayout<DIM, STAGE, HAS_U, HAS_X2>;
    constexpr int PTS = 128 >> LSHIFT;
    const bool want_y = p.y != nullptr;
    const bool want_g = p.ggrid != nullptr && (STAGE == ST_B || STAGE == ST_BB);
    const bool want_s = p.acc != nullptr && STAGE != ST_F;
    const int nfields = ((want_y || want_g) ? 1 : 0) + (HAS_U ? 1 : 0) + (want_s ? 1 : 0);
    const size_t cell_bytes = (size_t)p.cell_stride * sizeof(float);
    const size_t rec_per_warp = (size_t)RL::FIELDS4 * PTS * sizeof(float4);
    // auto mode: gather-only calls on cells of at most 32 KiB per field (e.g. 16 channels x 22^2;
    // at 64 KiB per field too few warps fit next to the cell and the global kernel wins).
    // Scattering calls stay on the global kernel: measured on the reference's own shapes
    // ([96,4,16,16], 100 000 points) contended shared-memory atomics are 2x slower than
    // red.global.add.v4.f32 into the L2-resident accumulator (profiles/README.md).
    if (p.small_cell == 0 && (cell_bytes > 32 * 1024 || want_s)) return false;
    int warps = 8;
    while (warps > 1 && nfields * cell_bytes + warps * rec_per_warp > 100 * 1024) warps >>= 1;
    const size_t smem = nfields * cell_bytes + warps * rec_per_warp;
    if (smem > 200 * 1024 || p.N > 65535) return false;
    auto kern = cs_small_kernel<DIM, VEC, LSHIFT, STAGE, HAS_U, HAS_X2>;
    static size_t configured[64] = {0};
    const int dev = device_index();
    if (smem > configured[dev]) {
        err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return true;
        configured[dev] = smem;
    }
    // blocks per cell: about two waves over the GPU, but at least 4 tiles per warp so that the
    // stage-in / flush of the cell is amortised
    const long long per_sm = (long long)(227 * 1024) / (long long)(smem + 1024);
    long long want_blocks = 2ll * device_sm_count(dev) * (per_sm < 1 ? 1 : per_sm);
    long long splits = (want_blocks + p.N - 1) / p.N;
    const long long max_splits = (p.num_ptiles + 4ll * warps - 1) / (4ll * warps);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    kern<<<dim3((unsigned)splits, (unsigned)p.N), warps * 32, smem, stream>>>(p);
    err = cudaGetLastError();
    return true;
}

template <int DIM, int VEC, int LSHIFT, int STAGE, bool HAS_U, bool HAS_X2, bool SCAT>
cudaError_t launch_global(const StageParams& p, cudaStream_t stream) {
    using WS = WarpSmem<DIM, VEC, LSHIFT, STAGE, HAS_U, HAS_X2>;
    auto kern = cs_stage_kernel<DIM, VEC, LSHIFT, STAGE, HAS_U, HAS_X2, SCAT>;
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

template <int DIM, int VEC, int LSHIFT, int STAGE, bool HAS_U, bool HAS_X2>
cudaError_t launch_one(const StageParams& p, cudaStream_t stream) {
    {
        cudaError_t err = cudaSuccess;
        if (try_launch_small<DIM, VEC, LSHIFT, STAGE, HAS_U, HAS_X2>(p, stream, err)) return err;
    }
    constexpr bool can_scatter = (STAGE != ST_F);
    if (can_scatter && p.acc != nullptr)
        return launch_global<DIM, VEC, LSHIFT, STAGE, HAS_U, HAS_X2, can_scatter>(p, stream);
    // the fused b_input stream only feeds the scatter: a gather-only BBB never has it
    return launch_global<DIM, VEC, LSHIFT, STAGE, HAS_U, false, false>(p, stream);
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
