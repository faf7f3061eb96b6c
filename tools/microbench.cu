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


// ---- TMA (non-tensor bulk copy, SASS UBLKCP) as the gather engine: every lane fetches one random 64-byte
// segment into shared memory with cp.async.bulk and the warp waits on one mbarrier; two buffers in flight
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity) {
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n"
        :: "r"(mbar), "r"(parity) : "memory");
}
__global__ void __launch_bounds__(256) bulk_gather_kernel(const float4* __restrict__ table, uint32_t nseg, float4* __restrict__ out,
                                                          long long groups) {
    __shared__ __align__(128) float4 buf[8][2][32][4];     // [warp][stage][lane][64 bytes]
    __shared__ __align__(8) unsigned long long bars[8][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(&bars[warp][0]);
    const unsigned bar1 = (unsigned)__cvta_generic_to_shared(&bars[warp][1]);
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar0));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const long long wid = blockIdx.x * 8ll + warp, nw = gridDim.x * 8ll;
    const long long iters = (groups / 32 + nw - 1) / nw;
    float4 acc = make_float4(0, 0, 0, 0);
    auto issue = [&](long long it, int st) {
        const long long g = (wid + it * nw) * 32 + lane;
        const uint32_t seg = hash32((uint32_t)g * 2654435761u + 12345u) % nseg;
        const unsigned bar = st ? bar1 : bar0;
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(32 * 64) : "memory");
        __syncwarp();
        const unsigned dst = (unsigned)__cvta_generic_to_shared(&buf[warp][st][lane][0]);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 64, [%2];"
                     :: "r"(dst), "l"(table + (size_t)seg * 4), "r"(bar) : "memory");
    };
    if (iters > 0) issue(0, 0);
    for (long long it = 0; it < iters; ++it) {
        const int st = it & 1;
        if (it + 1 < iters) issue(it + 1, st ^ 1);
        mbar_wait(st ? bar1 : bar0, (unsigned)((it >> 1) & 1));
        const float4 v = buf[warp][st][lane][lane & 3];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        __syncwarp();
    }
    if (acc.x == 123.456f) out[blockIdx.x * 256 + threadIdx.x] = acc;
}

// ---- cp.reduce.async.bulk as the scatter engine: every lane stages one 64-byte segment in shared memory and
// one bulk reduction adds it to a random 64-byte segment of the table (instead of 4 lanes x red.global.add.v4.f32)
__global__ void __launch_bounds__(256) bulk_red_kernel(float* __restrict__ table, uint32_t nseg, long long groups) {
    __shared__ __align__(128) float4 buf[256][4];
    const long long gid = blockIdx.x * 256ll + threadIdx.x;
    const long long stride = gridDim.x * 256ll;
    const unsigned src = (unsigned)__cvta_generic_to_shared(&buf[threadIdx.x][0]);
    for (long long g = gid; g < groups; g += stride) {
        const uint32_t seg = hash32((uint32_t)g * 2654435761u + 777u) % nseg;
#pragma unroll
        for (int i = 0; i < 4; ++i) buf[threadIdx.x][i] = make_float4(1.f, 2.f, 3.f, 4.f);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 64;"
                     :: "l"(table + (size_t)seg * 16), "r"(src) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

// ---- a [N,C,P] stream read through a permutation (what tile-binned points would do to the drop-in operator's
// streams): random 4-byte reads from a 256 MiB array
__global__ void gather4_kernel(const float* __restrict__ a, uint32_t n, float* __restrict__ out, long long reads) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long s = (long long)gridDim.x * blockDim.x;
    float acc = 0.f;
    for (; i < reads; i += s) acc += __ldg(a + hash32((uint32_t)i * 2654435761u + 99u) % n);
    if (acc == 123.456f) out[0] = acc;
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
    // ---- TMA bulk copy / bulk reduction against the LSU path ("what is not used")
    {
        const size_t bytes = (size_t)16 << 20;
        uint32_t nseg = bytes / 64;
        float ms = time_ms([&] { bulk_gather_kernel<<<sm * 4, 256>>>(A, nseg, B, groups); }, 10);
        printf("{\"bench\": \"bulkcopy_gather64B\", \"table_MiB\": 16, \"ms\": %.4f, \"Gseg_per_s\": %.2f, \"GBps\": %.1f, \"note\": \"cp.async.bulk (UBLKCP) 64 B per lane, mbarrier per warp, 2 stages\"}\n",
               ms, groups / ms / 1e6, groups * 64.0 / ms / 1e6);
        ms = time_ms([&] { bulk_red_kernel<<<sm * 8, 256>>>((float*)A, nseg, groups); }, 10);
        printf("{\"bench\": \"bulkreduce_red64B\", \"table_MiB\": 16, \"ms\": %.4f, \"Gseg_per_s\": %.2f, \"GBps\": %.1f, \"note\": \"cp.reduce.async.bulk.add.f32 64 B per lane from shared memory\"}\n",
               ms, groups / ms / 1e6, groups * 64.0 / ms / 1e6);
        const long long reads = 1ll << 26;
        ms = time_ms([&] { gather4_kernel<<<sm * 16, 256>>>((const float*)A, (uint32_t)((256u << 20) / 4), (float*)B, reads); }, 5);
        printf("{\"bench\": \"gather4B_permuted_stream\", \"table_MiB\": 256, \"ms\": %.4f, \"Greads_per_s\": %.2f, \"useful_GBps\": %.1f, \"note\": \"a 2^26-element stream read through a random permutation\"}\n",
               ms, reads / ms / 1e6, reads * 4.0 / ms / 1e6);
    }
    return 0;
}
