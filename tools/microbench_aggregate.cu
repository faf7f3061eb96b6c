// microbench_aggregate.cu -- three ways to scatter 2^24 contributions of 64 bytes (4 lanes x float4, one "walker" per
// contribution, 8 walkers per warp) into a 16 MiB table on B200, for the three point orders the sampler meets:
// unsorted texels, binned texels with runs of 4 and of 32 identical texels.
//   direct    : one red.global.add.v4.f32 per lane and contribution (the drop-in stage kernels)
//   matchany  : __match_any_sync over the warp's 8 texel ids, shuffle reduction over the matching walkers, the lowest
//               matching walker issues the red (the north star's "warp-level aggregation")
//   runs      : every walker takes 4 consecutive contributions and sums those with identical texel ids in registers
//               before the red (what the one-pass kernel does with binned points)
// Prints one JSON object per line.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s at %s:%d\"}\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__device__ __forceinline__ void red4(float* p, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// texel of contribution g: runs of `run` consecutive contributions share a texel (run = 1: unsorted)
__device__ __forceinline__ uint32_t texel_of(long long g, int run, uint32_t nseg) { return hash32((uint32_t)(g / run) * 2654435761u + 777u) % nseg; }
__device__ __forceinline__ float4 value_of(long long g, int j) { const float f = (float)((g * 4 + j) & 1023) * 1e-3f; return make_float4(f, f + 1.f, f + 2.f, f + 3.f); }

template <int MODE>   // 0 direct, 1 matchany, 2 runs of 4 per walker
__global__ void __launch_bounds__(256) scatter_kernel(float* __restrict__ table, uint32_t nseg, long long n, int run) {
    const int lane = threadIdx.x & 31, j = lane & 3, q = lane >> 2;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    if (MODE == 2) {
        // a warp tile = 32 consecutive contributions, walker q takes 4q .. 4q+3
        for (long long t = warp; t * 32 < n; t += nwarps) {
            uint32_t cur_seg = 0; float4 cur = make_float4(0, 0, 0, 0);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const long long g = t * 32 + 4 * q + i;
                const uint32_t seg = texel_of(g, run, nseg);
                const float4 v = value_of(g, j);
                if (i > 0 && seg == cur_seg) { cur.x += v.x; cur.y += v.y; cur.z += v.z; cur.w += v.w; }
                else { if (i > 0) red4(table + ((size_t)cur_seg * 4 + j) * 4, cur); cur = v; cur_seg = seg; }
            }
            red4(table + ((size_t)cur_seg * 4 + j) * 4, cur);
        }
    } else {
        // a warp round = 8 consecutive contributions, one per walker
        for (long long t = warp; t * 8 < n; t += nwarps) {
            const long long g = t * 8 + q;
            const uint32_t seg = texel_of(g, run, nseg);
            float4 v = value_of(g, j);
            if (MODE == 1) {
                const unsigned peers = __match_any_sync(0xffffffffu, seg * 4u + (uint32_t)j);
                float4 sum = v;
#pragma unroll
                for (int w = 1; w < 8; ++w) {
                    const int src = (lane + 4 * w) & 31;
                    const float4 o = make_float4(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src),
                                                 __shfl_sync(0xffffffffu, v.z, src), __shfl_sync(0xffffffffu, v.w, src));
                    if ((peers >> src) & 1u) { sum.x += o.x; sum.y += o.y; sum.z += o.z; sum.w += o.w; }
                }
                if ((int)(__ffs(peers) - 1) == lane) red4(table + ((size_t)seg * 4 + j) * 4, sum);
            } else {
                red4(table + ((size_t)seg * 4 + j) * 4, v);
            }
        }
    }
}

template <int MODE>
static void run(const char* name, float* table, uint32_t nseg, long long n, int runlen, int sm) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 2; ++i) scatter_kernel<MODE><<<sm * 8, 256>>>(table, nseg, n, runlen);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 10; ++i) scatter_kernel<MODE><<<sm * 8, 256>>>(table, nseg, n, runlen);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 10;
    printf("{\"bench\": \"scatter64B_%s\", \"run_of_identical_texels\": %d, \"contributions\": %lld, \"ms\": %.4f, \"Gcontrib_per_s\": %.2f}\n",
           name, runlen, n, ms, n / ms / 1e6);
}

int main() {
    int sm = 148; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    const size_t bytes = 16u << 20;
    float* table; CK(cudaMalloc(&table, bytes)); CK(cudaMemset(table, 0, bytes));
    const uint32_t nseg = bytes / 64;
    const long long n = 1ll << 24;
    for (int runlen : {1, 4, 32}) {
        run<0>("direct", table, nseg, n, runlen, sm);
        run<1>("matchany_shuffle", table, nseg, n, runlen, sm);
        run<2>("register_runs_of_4", table, nseg, n, runlen, sm);
    }
    return 0;
}
