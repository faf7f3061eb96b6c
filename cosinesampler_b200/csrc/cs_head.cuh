// cs_head.cuh -- fused PDE-residual head on jets (PIXEL caller glue).
//
// The reference's scripts put a small MLP head, Linear(C,16)-Tanh-Linear(16,1), on the sampled
// features and obtain u_a, u_aa by nested autograd through it (test_2d.py:42-127); under torch
// that is ~200 elementwise / GEMM kernels per step and dominates the step time once the sampler
// itself is fast (profiles/README.md).  With the jet operator (cs_jet.cuh) the head only needs
// second-order Taylor mode, which is closed-form per point:
//
//     h  = W1 z + b1        hd_a = W1 z_a          hdd_a = W1 z_aa
//     t  = tanh h           s1 = 1 - t^2           s2 = -2 t s1         s3 = -2 (s1^2 + t s2)
//     u  = w2.t + b2        u_a = w2.(s1 hd_a)     u_aa = w2.(s2 hd_a^2 + s1 hdd_a)
//     f  = c_u u + c_u3 u^3 + sum_a (c1_a u_a + c2_a u_aa)          loss = scale * sum_p f^2
//
// This kernel evaluates loss AND its gradient w.r.t. the jets and the head parameters in one
// pass: a block stages a [J*C rows] x [TP points] tile of the jets in shared memory (coalesced),
// the 16 lanes of a half-warp are the 16 hidden units and walk groups of 4 points (weights and
// weight-gradient accumulators live in registers; jets are read as 128-bit broadcasts), the
// gradient w.r.t. the jets overwrites the tile and leaves coalesced.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cs {

struct HeadParams {
    long long P;
    const float* jets;      // [J, C, P]
    float* gjets;           // [J, C, P] (may alias jets)
    const float* W1;        // [16, C]
    const float* b1;        // [16]
    const float* w2;        // [16]
    const float* b2;        // [1]
    float* gW1;             // [16, C]   accumulated (+=)
    float* gb1;             // [16]
    float* gw2;             // [16]
    float* gb2;             // [1]
    float* loss_sum;        // [1]       += sum_p f^2 (unscaled)
    float* f_out;           // [P] or nullptr
    float c_u, c_u3, c1[3], c2[3];
    float scale;            // gradients are those of scale * sum_p f^2
    int vec;                // rows may be accessed as 16-byte vectors
};

constexpr int HEAD_K = 16;          // hidden width (test_2d.py:44)
constexpr int HEAD_THREADS = 128;
constexpr int HEAD_TP = 64;         // points per tile

template <int DIM, int C> struct HeadSmem {
    static constexpr int J = 1 + 2 * DIM;
    static constexpr int ROWS = J * C;
    static constexpr int TILE_F4 = ROWS * (HEAD_TP / 4);
    static constexpr int XCH_F4 = (HEAD_THREADS / 16) * J * HEAD_K;       // one [J][K] float4 array per half-warp
    static constexpr int RED_F = (HEAD_THREADS / 32) * (HEAD_K * C + 2 * HEAD_K + 2);
    static constexpr size_t BYTES = (size_t)(TILE_F4 + XCH_F4) * 16;      // the reduction scratch reuses the tile
};

__device__ __forceinline__ float half_sum16(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// tanh x = -em / (2 + em), em = expm1(-2|x|): accurate for small and large |x| alike, and without
// the data-dependent branches of tanhf (which, unrolled over the 4 points of a group, made the
// compiler spill the weight and accumulator registers)
__device__ __forceinline__ float tanh_branchfree(float x) {
    const float em = expm1f(-2.f * fabsf(x));
    return copysignf(-em / (2.f + em), x);
}

template <int DIM, int C>
__global__ void __launch_bounds__(HEAD_THREADS, (C <= 16 ? 4 : 2))
cs_pde_head_kernel(const HeadParams p) {
    using HS = HeadSmem<DIM, C>;
    constexpr int J = HS::J;
    constexpr int K = HEAD_K;
    constexpr int ROWS = HS::ROWS;
    constexpr int TP = HEAD_TP;
    constexpr int TP4 = TP / 4;
    constexpr int CPL = (C + 15) / 16;          // channels per lane in the W1^T pass

    extern __shared__ float4 smem4[];
    float4* tile4 = smem4;                       // [ROWS][TP4]
    float4* xch_all = smem4 + HS::TILE_F4;       // [8][J][K]
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int k = lane & 15;
    const int hw = tid >> 4;                     // half-warp index in the block
    float4* xch = xch_all + hw * J * K;

    // ---- weights in registers
    float w1row[C];                              // W1[k][:]
#pragma unroll
    for (int c = 0; c < C; ++c) w1row[c] = __ldg(p.W1 + k * C + c);
    float w1col[CPL][K];                         // W1[:][c], c = k + 16 i
#pragma unroll
    for (int i = 0; i < CPL; ++i)
#pragma unroll
        for (int kk = 0; kk < K; ++kk) {
            const int c = k + 16 * i;
            w1col[i][kk] = (c < C) ? __ldg(p.W1 + kk * C + c) : 0.f;
        }
    const float b1k = __ldg(p.b1 + k);
    const float w2k = __ldg(p.w2 + k);
    const float b2 = __ldg(p.b2);

    float gW1acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) gW1acc[c] = 0.f;
    float gb1acc = 0.f, gw2acc = 0.f, gb2acc = 0.f, lossacc = 0.f;

    const long long ntiles = (p.P + TP - 1) / TP;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long p0 = tile * TP;
        // ---- stage the tile: rows of the [J*C, P] array, TP consecutive points each
        for (int idx = tid; idx < ROWS * TP4; idx += HEAD_THREADS) {
            const int r = idx / TP4;
            const int v = idx - r * TP4;
            const long long pp = p0 + 4 * v;
            const float* src = p.jets + (long long)r * p.P + pp;
            float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.vec) {
                if (pp < p.P) val = __ldcs(reinterpret_cast<const float4*>(src));
            } else {
                if (pp < p.P) val.x = __ldcs(src);
                if (pp + 1 < p.P) val.y = __ldcs(src + 1);
                if (pp + 2 < p.P) val.z = __ldcs(src + 2);
                if (pp + 3 < p.P) val.w = __ldcs(src + 3);
            }
            tile4[idx] = val;
        }
        __syncthreads();

        for (int grp = hw; grp < TP4; grp += HEAD_THREADS / 16) {
            const long long gp0 = p0 + 4 * grp;
            // ---- hidden pre-activations of 4 points for hidden unit k, all jets
            float h[J][4];
#pragma unroll
            for (int j = 0; j < J; ++j) {
                h[j][0] = h[j][1] = h[j][2] = h[j][3] = 0.f;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const float4 z = tile4[(j * C + c) * TP4 + grp];
                    h[j][0] = fmaf(w1row[c], z.x, h[j][0]);
                    h[j][1] = fmaf(w1row[c], z.y, h[j][1]);
                    h[j][2] = fmaf(w1row[c], z.z, h[j][2]);
                    h[j][3] = fmaf(w1row[c], z.w, h[j][3]);
                }
            }
            // ---- per point: activation derivatives, residual, and the gradient w.r.t. h
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float t = tanh_branchfree(h[0][i] + b1k);
                const float s1 = 1.f - t * t;
                const float s2 = -2.f * t * s1;
                const float s3 = -2.f * (s1 * s1 + t * s2);
                float u = half_sum16(w2k * t) + b2;
                float f = p.c_u * u + p.c_u3 * u * u * u;
                float ua[DIM], uaa[DIM];
#pragma unroll
                for (int a = 0; a < DIM; ++a) {
                    const float hd = h[1 + a][i], hdd = h[1 + DIM + a][i];
                    ua[a] = half_sum16(w2k * s1 * hd);
                    uaa[a] = half_sum16(w2k * (s2 * hd * hd + s1 * hdd));
                    f += p.c1[a] * ua[a] + p.c2[a] * uaa[a];
                }
                const bool valid = gp0 + i < p.P;
                const float g = valid ? 2.f * p.scale * f : 0.f;
                if (k == 0 && valid) {
                    lossacc += f * f;
                    if (p.f_out) p.f_out[gp0 + i] = f;
                }
                const float gu = g * (p.c_u + 3.f * p.c_u3 * u * u);
                float gw2 = gu * t;
                float gh = gu * s1;
#pragma unroll
                for (int a = 0; a < DIM; ++a) {
                    const float hd = h[1 + a][i], hdd = h[1 + DIM + a][i];
                    const float g1 = g * p.c1[a], g2 = g * p.c2[a];
                    gw2 += g1 * s1 * hd + g2 * (s2 * hd * hd + s1 * hdd);
                    gh += g1 * s2 * hd + g2 * (s3 * hd * hd + s2 * hdd);
                    h[1 + a][i] = w2k * (g1 * s1 + g2 * 2.f * s2 * hd);
                    h[1 + DIM + a][i] = w2k * g2 * s1;
                }
                gh *= w2k;
                h[0][i] = gh;
                gb1acc += gh;
                gw2acc += gw2;
                if (k == 0) gb2acc += gu;
            }
            // ---- weight gradient: gW1[k][c] += sum_j sum_i gH_j[k][i] * z_j[c][i]
            // (compiler barrier: re-read z from shared memory instead of keeping 4*J*C values live)
            asm volatile("" ::: "memory");
#pragma unroll
            for (int j = 0; j < J; ++j)
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const float4 z = tile4[(j * C + c) * TP4 + grp];
                    gW1acc[c] = fmaf(h[j][0], z.x, gW1acc[c]);
                    gW1acc[c] = fmaf(h[j][1], z.y, gW1acc[c]);
                    gW1acc[c] = fmaf(h[j][2], z.z, gW1acc[c]);
                    gW1acc[c] = fmaf(h[j][3], z.w, gW1acc[c]);
                }
            // ---- exchange gH over the hidden units, then gZ_j[c] = sum_k W1[k][c] gH_j[k]
#pragma unroll
            for (int j = 0; j < J; ++j) xch[j * K + k] = make_float4(h[j][0], h[j][1], h[j][2], h[j][3]);
            __syncwarp();                        // also: every lane has finished reading z of this group
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
                const int c = k + 16 * i;
                if (c < C) {
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int kk = 0; kk < K; ++kk) {
                            const float4 gq = xch[j * K + kk];
                            acc.x = fmaf(w1col[i][kk], gq.x, acc.x);
                            acc.y = fmaf(w1col[i][kk], gq.y, acc.y);
                            acc.z = fmaf(w1col[i][kk], gq.z, acc.z);
                            acc.w = fmaf(w1col[i][kk], gq.w, acc.w);
                        }
                        tile4[(j * C + c) * TP4 + grp] = acc;
                    }
                }
            }
            __syncwarp();                        // xch is reused by the next group
        }
        __syncthreads();

        // ---- the tile now holds d loss / d jets: store it
        for (int idx = tid; idx < ROWS * TP4; idx += HEAD_THREADS) {
            const int r = idx / TP4;
            const int v = idx - r * TP4;
            const long long pp = p0 + 4 * v;
            float* dst = p.gjets + (long long)r * p.P + pp;
            const float4 val = tile4[idx];
            if (p.vec) {
                if (pp < p.P) __stcs(reinterpret_cast<float4*>(dst), val);
            } else {
                if (pp < p.P) __stcs(dst, val.x);
                if (pp + 1 < p.P) __stcs(dst + 1, val.y);
                if (pp + 2 < p.P) __stcs(dst + 2, val.z);
                if (pp + 3 < p.P) __stcs(dst + 3, val.w);
            }
        }
        __syncthreads();
    }

    // ---- parameter gradients and loss: halves -> warps (shared memory) -> one atomic per block
    float* red = reinterpret_cast<float*>(smem4);            // [4 warps][K*C + 2K + 2], reuses the tile
    constexpr int RW = K * C + 2 * K + 2;
#pragma unroll
    for (int c = 0; c < C; ++c) gW1acc[c] += __shfl_xor_sync(0xffffffffu, gW1acc[c], 16);
    gb1acc += __shfl_xor_sync(0xffffffffu, gb1acc, 16);
    gw2acc += __shfl_xor_sync(0xffffffffu, gw2acc, 16);
    gb2acc += __shfl_xor_sync(0xffffffffu, gb2acc, 16);
    lossacc += __shfl_xor_sync(0xffffffffu, lossacc, 16);
    if (lane < 16) {
        float* rw = red + warp * RW;
#pragma unroll
        for (int c = 0; c < C; ++c) rw[k * C + c] = gW1acc[c];
        rw[K * C + k] = gb1acc;
        rw[K * C + K + k] = gw2acc;
        if (k == 0) { rw[K * C + 2 * K] = gb2acc; rw[K * C + 2 * K + 1] = lossacc; }
    }
    __syncthreads();
    for (int e = tid; e < RW; e += HEAD_THREADS) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < HEAD_THREADS / 32; ++w) s += red[w * RW + e];
        float* dst = (e < K * C) ? p.gW1 + e : (e < K * C + K) ? p.gb1 + (e - K * C)
                   : (e < K * C + 2 * K) ? p.gw2 + (e - K * C - K)
                   : (e == K * C + 2 * K) ? p.gb2 : p.loss_sum;
        atomicAdd(dst, s);
    }
}

template <int DIM, int C>
cudaError_t launch_head(const HeadParams& p, cudaStream_t stream) {
    using HS = HeadSmem<DIM, C>;
    auto kern = cs_pde_head_kernel<DIM, C>;
    size_t smem = HS::BYTES;
    const size_t red_bytes = (size_t)HS::RED_F * 4;
    if (smem < red_bytes) smem = red_bytes;
    static int occ_cache[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    int& occ = occ_cache[dev];
    if (occ == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, HEAD_THREADS, smem);
        if (e != cudaSuccess) return e;
        if (occ < 1) occ = 1;
    }
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    long long blocks = (p.P + HEAD_TP - 1) / HEAD_TP;
    const long long cap = (long long)sms * occ;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) return cudaSuccess;
    kern<<<(unsigned)blocks, HEAD_THREADS, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace cs
