// microbench.cu -- design-deciding micro-measurements for the sampler kernels on B200
// (SURVEY Appendix B item 4): L2->SM gather bandwidth at 64 B / 128 B granularity on a 16 MiB and
// a 64 MiB table, red.global.add.v4.f32 vs scalar red throughput on random texels, and a
// streaming copy for scale.  Prints one JSON object per line.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s at %s:%d\"}\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

// each group of L lanes reads SEG = 16*L contiguous bytes at a random SEG-aligned position
template <int L, int UNROLL>
__global__ void gather_kernel(const float4* __restrict__ table, uint32_t nseg, float4* __restrict__ out, long long groups) {
    const long long gid = (blockIdx.x * (long long)blockDim.x + threadIdx.x);
    const int j = threadIdx.x % L;
    long long g = gid / L;
    const long long gstride = (long long)gridDim.x * blockDim.x / L;
    float4 acc = make_float4(0, 0, 0, 0);
    for (; g < groups; g += gstride * UNROLL) {
        float4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            long long gg = g + u * gstride;
            uint32_t seg = hash32((uint32_t)gg * 2654435761u + 12345u) % nseg;
            v[u] = (gg < groups) ? __ldg(table + (size_t)seg * L + j) : make_float4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    if (acc.x == 123.456f) out[gid] = acc;   // keep the loads alive
}

template <int L, bool VEC>
__global__ void red_kernel(float* __restrict__ table, uint32_t nseg, long long groups) {
    const long long gid = (blockIdx.x * (long long)blockDim.x + threadIdx.x);
    const int j = threadIdx.x % L;
    long long g = gid / L;
    const long long gstride = (long long)gridDim.x * blockDim.x / L;
    for (; g < groups; g += gstride) {
        uint32_t seg = hash32((uint32_t)g * 2654435761u + 777u) % nseg;
        float* p = table + ((size_t)seg * L + j) * 4;
        if (VEC) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p), "f"(1.f), "f"(2.f), "f"(3.f), "f"(4.f) : "memory");
        } else {
            atomicAdd(p, 1.f); atomicAdd(p + 1, 2.f); atomicAdd(p + 2, 3.f); atomicAdd(p + 3, 4.f);
        }
    }
}

__global__ void copy_kernel(const float4* __restrict__ a, float4* __restrict__ b, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long s = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += s) b[i] = a[i];
}

template <typename F>
float time_ms(F f, int iters) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < 3; ++i) f();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < iters; ++i) f();
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    return ms / iters;
}

int main() {
    int sm = 148; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    const size_t big = (size_t)1 << 30;
    float4 *A, *B; CK(cudaMalloc(&A, big)); CK(cudaMalloc(&B, big));
    CK(cudaMemset(A, 0, big)); CK(cudaMemset(B, 0, big));
    {
        long long n = big / 16;
        float ms = time_ms([&] { copy_kernel<<<sm * 16, 512>>>(A, B, n); }, 10);
        printf("{\"bench\": \"copy_1GiB\", \"ms\": %.4f, \"GBps_rw\": %.1f}\n", ms, 2.0 * big / ms / 1e6);
    }
    const long long groups = 1ll << 24;   // 2^22 pairs x 4 corners
    for (size_t table_mib : {16, 64}) {
        const size_t bytes = table_mib << 20;
        {
            uint32_t nseg = bytes / 64;
            float ms = time_ms([&] { gather_kernel<4, 4><<<sm * 8, 256>>>(A, nseg, B, groups); }, 10);
            printf("{\"bench\": \"gather64B_L4\", \"table_MiB\": %zu, \"ms\": %.4f, \"Gseg_per_s\": %.2f, \"GBps\": %.1f}\n",
                   table_mib, ms, groups / ms / 1e6, groups * 64.0 / ms / 1e6);
            ms = time_ms([&] { gather_kernel<4, 8><<<sm * 8, 256>>>(A, nseg, B, groups); }, 10);
            printf("{\"bench\": \"gather64B_L4_u8\", \"table_MiB\": %zu, \"ms\": %.4f, \"Gseg_per_s\": %.2f, \"GBps\": %.1f}\n",
                   table_mib, ms, groups / ms / 1e6, groups * 64.0 / ms / 1e6);
        }
        {
            uint32_t nseg = bytes / 128;
            long long g2 = groups / 2;
            float ms = time_ms([&] { gather_kernel<8, 4><<<sm * 8, 256>>>(A, nseg, B, g2); }, 10);
            printf("{\"bench\": \"gather128B_L8\", \"table_MiB\": %zu, \"ms\": %.4f, \"Gseg_per_s\": %.2f, \"GBps\": %.1f}\n",
                   table_mib, ms, g2 / ms / 1e6, g2 * 128.0 / ms / 1e6);
        }
        {
            uint32_t nseg = bytes / 16;
            long long g4 = groups * 4;
            float ms = time_ms([&] { gather_kernel<1, 8><<<sm * 8, 256>>>(A, nseg, B, g4); }, 5);
            printf("{\"bench\": \"gather16B_L1\", \"table_MiB\": %zu, \"ms\": %.4f, \"Gseg_per_s\": %.2f, \"GBps\": %.1f}\n",
                   table_mib, ms, g4 / ms / 1e6, g4 * 16.0 / ms / 1e6);
        }
        {
            uint32_t nseg = bytes / 64;
            float ms = time_ms([&] { red_kernel<4, true><<<sm * 8, 256>>>((float*)A, nseg, groups); }, 10);
            printf("{\"bench\": \"red_v4_64B_L4\", \"table_MiB\": %zu, \"ms\": %.4f, \"Gseg_per_s\": %.2f, \"GBps\": %.1f}\n",
                   table_mib, ms, groups / ms / 1e6, groups * 64.0 / ms / 1e6);
            {
                uint32_t nseg8 = bytes / 128;
                long long g8 = groups / 2;
                float ms8 = time_ms([&] { red_kernel<8, true><<<sm * 8, 256>>>((float*)A, nseg8, g8); }, 10);
                printf("{\"bench\": \"red_v4_128B_L8\", \"table_MiB\": %zu, \"ms\": %.4f, \"Gseg_per_s\": %.2f, \"GBps\": %.1f}\n",
                       table_mib, ms8, g8 / ms8 / 1e6, g8 * 128.0 / ms8 / 1e6);
                uint32_t nseg2 = bytes / 32;
                long long g2 = groups * 2;
                float ms2 = time_ms([&] { red_kernel<2, true><<<sm * 8, 256>>>((float*)A, nseg2, g2); }, 10);
                printf("{\"bench\": \"red_v4_32B_L2\", \"table_MiB\": %zu, \"ms\": %.4f, \"Gseg_per_s\": %.2f, \"GBps\": %.1f}\n",
                       table_mib, ms2, g2 / ms2 / 1e6, g2 * 32.0 / ms2 / 1e6);
            }
            ms = time_ms([&] { red_kernel<4, false><<<sm * 8, 256>>>((float*)A, nseg, groups); }, 5);
            printf("{\"bench\": \"red_scalar_64B_L4\", \"table_MiB\": %zu, \"ms\": %.4f, \"Gseg_per_s\": %.2f, \"GBps\": %.1f}\n",
                   table_mib, ms, groups / ms / 1e6, groups * 64.0 / ms / 1e6);
        }
    }
    // small table that fits L1/smem-sized working sets: 256 KiB (PIXEL-sized cells)
    {
        uint32_t nseg = (256 << 10) / 64;
        float ms = time_ms([&] { gather_kernel<4, 4><<<sm * 8, 256>>>(A, nseg, B, groups); }, 10);
        printf("{\"bench\": \"gather64B_L4\", \"table_MiB\": 0.25, \"ms\": %.4f, \"Gseg_per_s\": %.2f, \"GBps\": %.1f}\n",
               ms, groups / ms / 1e6, groups * 64.0 / ms / 1e6);
        ms = time_ms([&] { red_kernel<4, true><<<sm * 8, 256>>>((float*)A, nseg, groups); }, 10);
        printf("{\"bench\": \"red_v4_64B_L4\", \"table_MiB\": 0.25, \"ms\": %.4f, \"Gseg_per_s\": %.2f, \"GBps\": %.1f}\n",
               ms, groups / ms / 1e6, groups * 64.0 / ms / 1e6);
    }
    return 0;
}
