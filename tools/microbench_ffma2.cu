// microbench_ffma2.cu -- issue / pipe throughput of scalar FFMA against packed FFMA2 (fma.rn.f32x2, sm_100+) on
// B200: 8 independent accumulator chains per lane, 12 warps per SM sub-partition worth of blocks.
// Prints one JSON object per line.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s at %s:%d\"}\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float ffma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

constexpr int CH = 8;
template <int MODE>   // 0: scalar FFMA (3 registers), 1: FFMA2 with packed operands, 2: FFMA2 with a broadcast scalar operand
__global__ void __launch_bounds__(128) fma_kernel(float* out, int iters, float s0, float s1) {
    float a[CH], b[CH];
    u64 pa[CH];
    const float x = s0 + threadIdx.x * 1e-9f, y = s1;
#pragma unroll
    for (int c = 0; c < CH; ++c) { a[c] = c * 0.25f; b[c] = c * 0.5f; pa[c] = pk(a[c], b[c]); }
    const u64 px = pk(x, x * 1.0001f), py = pk(y, y);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            if (MODE == 0) { a[c] = ffma1(a[c], x, y); b[c] = ffma1(b[c], x, y); }
            else if (MODE == 1) pa[c] = ffma2(pa[c], px, py);
            else pa[c] = ffma2(pa[c], pk(x, x), py);
        }
    }
    float r = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        if (MODE == 0) r += a[c] + b[c];
        else { float u, v; upk(pa[c], u, v); r += u + v; }
    }
    if (r == 123.456f) out[0] = r;
}

template <int MODE>
static void run(const char* name, int blocks_per_sm) {
    float* out; CK(cudaMalloc(&out, 4));
    const int iters = 4096;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int blocks = 148 * blocks_per_sm;
    fma_kernel<MODE><<<blocks, 128>>>(out, iters, 0.999f, 0.001f);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    fma_kernel<MODE><<<blocks, 128>>>(out, iters, 0.999f, 0.001f);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double fmas = (double)blocks * 128 * iters * CH * 2;            // scalar FMAs
    const double winstr = (double)blocks * 4 * iters * CH * (MODE == 0 ? 2 : 1);
    printf("{\"bench\": \"%s\", \"warps_per_SM\": %d, \"ms\": %.4f, \"TFMA_per_s\": %.2f, \"fma_lanes_per_clk_per_SM\": %.1f, \"warp_instr_per_clk_per_SMSP\": %.3f}\n",
           name, blocks_per_sm * 4, ms, fmas / ms / 1e9, fmas / (ms * 1e-3) / 148 / 1.965e9, winstr / (ms * 1e-3) / 148 / 4 / 1.965e9);
    CK(cudaFree(out));
}

int main() {
    for (int bps : {1, 3, 6}) {
        run<0>("ffma_scalar", bps);
        run<1>("ffma2_packed", bps);
        run<2>("ffma2_broadcast_scalar", bps);
    }
    return 0;
}
