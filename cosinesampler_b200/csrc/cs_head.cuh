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
// pass: a block stages a [J*C rows] x [TP points] tile of the jets in shared memory (coalesced);
// 8 lanes, two hidden units each, share a group of 4 consecutive points (weights and
// weight-gradient accumulators live in registers; the jets are read as 128-bit broadcasts, each
// feeding 8 FFMA); sums over the hidden units are 3-step shuffles; the gradient w.r.t. the jets
// goes through a small shared-memory exchange, overwrites the tile and leaves coalesced.
// Measured alternatives (profiles/README.md): one hidden unit per lane (4 FFMA per broadcast
// LDS.128) is bound by the shared-memory pipe, 0.53 ms per 2^20 points; one thread per point with
// the weights as constant-bank operands and the weight gradient re-partitioned through shared
// memory is latency-bound at 12 warps per SM, 0.74 ms.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "cs_engine.cuh"

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
#ifndef CS_HEAD_KPL
#define CS_HEAD_KPL 2
#endif
#ifndef CS_HEAD_BLOCKS
#define CS_HEAD_BLOCKS 3
#endif
constexpr int HEAD_KPL = CS_HEAD_KPL;   // hidden units per lane
constexpr int HEAD_LPG = HEAD_K / HEAD_KPL;             // lanes per point group (8)
constexpr int HEAD_GPB = HEAD_THREADS / HEAD_LPG;       // point groups per block pass (16)

template <int DIM, int C> struct HeadSmem {
    static constexpr int J = 1 + 2 * DIM;
    static constexpr int ROWS = J * C;
    static constexpr int TILE_F4 = ROWS * (HEAD_TP / 4);
    static constexpr int XCH_F4 = HEAD_GPB * J * HEAD_K;                 // one [J][K] float4 array per point group
    static constexpr int W1T_F4 = C * HEAD_K / 4;                        // W1^T [C][K], read in the W1^T pass
    static constexpr int RED_F = (HEAD_THREADS / 32) * (HEAD_K * C + 2 * HEAD_K + 2);
    static constexpr size_t BYTES = (size_t)(2 * TILE_F4 + XCH_F4 + W1T_F4) * 16;   // two tiles: the next one is prefetched
};

// sum over the HEAD_LPG lanes that share a point group
__device__ __forceinline__ float group_sum8(float v) {
#pragma unroll
    for (int o = HEAD_LPG / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// tanh x = -em / (2 + em), em = expm1(-2|x|): accurate for small and large |x| alike, and without
// the data-dependent branches of tanhf (which, unrolled over the points of a group, made the
// compiler spill the weight and accumulator registers)
__device__ __forceinline__ float tanh_branchfree(float x) {
    const float em = expm1f(-2.f * fabsf(x));
    return copysignf(-em / (2.f + em), x);
}

template <int DIM, int C>
__global__ void __launch_bounds__(HEAD_THREADS, (C <= 16 ? CS_HEAD_BLOCKS : 2))
cs_pde_head_kernel(const HeadParams p) {
    using HS = HeadSmem<DIM, C>;
    constexpr int J = HS::J;
    constexpr int K = HEAD_K;
    constexpr int KPL = HEAD_KPL;
    constexpr int ROWS = HS::ROWS;
    constexpr int TP = HEAD_TP;
    constexpr int TP4 = TP / 4;
    constexpr int CPL = (C + HEAD_LPG - 1) / HEAD_LPG;    // channels per lane in the W1^T pass

    extern __shared__ float4 smem4[];
    float4* tiles = smem4;                       // [2][ROWS][TP4]
    float4* xch_all = smem4 + 2 * HS::TILE_F4;   // [GPB][J][K]
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int kq = lane & (HEAD_LPG - 1);        // this lane's hidden units: KPL*kq .. KPL*kq + KPL-1
    const int gslot = tid / HEAD_LPG;            // point-group slot in the block
    float4* xch = xch_all + gslot * J * K;

    // ---- weights in registers
    float w1row[KPL][C];                         // W1[KPL*kq + i][:]
#pragma unroll
    for (int i = 0; i < KPL; ++i)
#pragma unroll
        for (int c = 0; c < C; ++c) w1row[i][c] = __ldg(p.W1 + (KPL * kq + i) * C + c);
    // W1^T in shared memory for the W1^T pass (the columns would cost another C*K/8 registers)
    float* w1t = reinterpret_cast<float*>(smem4 + 2 * HS::TILE_F4 + HS::XCH_F4);       // [C][K]
    for (int e = tid; e < C * K; e += HEAD_THREADS) w1t[e] = __ldg(p.W1 + (e % K) * C + (e / K));
    const float4* w1t4 = reinterpret_cast<const float4*>(w1t);
    __syncthreads();
    float b1k[KPL], w2k[KPL];
#pragma unroll
    for (int i = 0; i < KPL; ++i) { b1k[i] = __ldg(p.b1 + KPL * kq + i); w2k[i] = __ldg(p.w2 + KPL * kq + i); }
    const float b2 = __ldg(p.b2);

    float gW1acc[KPL][C];
    float gb1acc[KPL], gw2acc[KPL];
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
        gb1acc[i] = 0.f; gw2acc[i] = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) gW1acc[i][c] = 0.f;
    }
    float gb2acc = 0.f, lossacc = 0.f;

    // Stage rows of the [J*C, P] array, TP consecutive points each, into a tile buffer: cp.async
    // (16 bytes, zero-filled past P) when the rows can be accessed as vectors, plain loads otherwise.
    auto prefetch = [&](long long tile, float4* dst4) {
        const long long p0 = tile * TP;
        const unsigned sbase = (unsigned)__cvta_generic_to_shared(dst4);
        for (int idx = tid; idx < ROWS * TP4; idx += HEAD_THREADS) {
            const int r = idx / TP4;
            const int v = idx - r * TP4;
            const long long pp = p0 + 4 * v;
            const float* src = p.jets + (long long)r * p.P + pp;
            if (p.vec) {
                const bool ok = pp < p.P;
                cp16z(sbase + idx * 16, ok ? src : p.jets, ok);
            } else {
                float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
                if (pp < p.P) val.x = __ldcs(src);
                if (pp + 1 < p.P) val.y = __ldcs(src + 1);
                if (pp + 2 < p.P) val.z = __ldcs(src + 2);
                if (pp + 3 < p.P) val.w = __ldcs(src + 3);
                dst4[idx] = val;
            }
        }
        cp_async_commit();
    };

    const long long ntiles = (p.P + TP - 1) / TP;
    int buf = 0;
    if ((long long)blockIdx.x < ntiles) prefetch(blockIdx.x, tiles);
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long p0 = tile * TP;
        float4* tile4 = tiles + buf * HS::TILE_F4;
        // the next tile streams in while this one is computed
        const long long next = tile + gridDim.x;
        if (next < ntiles) { prefetch(next, tiles + (buf ^ 1) * HS::TILE_F4); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();

        for (int grp = gslot; grp < TP4; grp += HEAD_GPB) {
            const long long gp0 = p0 + 4 * grp;
            // ---- A: hidden pre-activations of 4 points for this lane's hidden units, all jets
            float h[KPL][J][4];
#pragma unroll
            for (int j = 0; j < J; ++j) {
#pragma unroll
                for (int i = 0; i < KPL; ++i) h[i][j][0] = h[i][j][1] = h[i][j][2] = h[i][j][3] = 0.f;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const float4 z = tile4[(j * C + c) * TP4 + grp];
#pragma unroll
                    for (int i = 0; i < KPL; ++i) {
                        h[i][j][0] = fmaf(w1row[i][c], z.x, h[i][j][0]);
                        h[i][j][1] = fmaf(w1row[i][c], z.y, h[i][j][1]);
                        h[i][j][2] = fmaf(w1row[i][c], z.z, h[i][j][2]);
                        h[i][j][3] = fmaf(w1row[i][c], z.w, h[i][j][3]);
                    }
                }
            }
            // ---- B: per point: activation derivatives, residual, and the gradient w.r.t. h
#pragma unroll
            for (int pt = 0; pt < 4; ++pt) {
                float t[KPL], s1[KPL], s2[KPL];
                float pu = 0.f, pua[DIM], puaa[DIM];
#pragma unroll
                for (int a = 0; a < DIM; ++a) { pua[a] = 0.f; puaa[a] = 0.f; }
#pragma unroll
                for (int i = 0; i < KPL; ++i) {
                    t[i] = tanh_branchfree(h[i][0][pt] + b1k[i]);
                    s1[i] = 1.f - t[i] * t[i];
                    s2[i] = -2.f * t[i] * s1[i];
                    pu = fmaf(w2k[i], t[i], pu);
#pragma unroll
                    for (int a = 0; a < DIM; ++a) {
                        const float hd = h[i][1 + a][pt], hdd = h[i][1 + DIM + a][pt];
                        pua[a] = fmaf(w2k[i], s1[i] * hd, pua[a]);
                        puaa[a] = fmaf(w2k[i], s2[i] * hd * hd + s1[i] * hdd, puaa[a]);
                    }
                }
                const float u = group_sum8(pu) + b2;
                float f = p.c_u * u + p.c_u3 * u * u * u;
#pragma unroll
                for (int a = 0; a < DIM; ++a) f += p.c1[a] * group_sum8(pua[a]) + p.c2[a] * group_sum8(puaa[a]);
                const bool valid = gp0 + pt < p.P;
                const float g = valid ? 2.f * p.scale * f : 0.f;
                if (kq == 0 && valid) {
                    lossacc += f * f;
                    if (p.f_out) p.f_out[gp0 + pt] = f;
                }
                const float gu = g * (p.c_u + 3.f * p.c_u3 * u * u);
                if (kq == 0) gb2acc += gu;
#pragma unroll
                for (int i = 0; i < KPL; ++i) {
                    const float s3 = -2.f * (s1[i] * s1[i] + t[i] * s2[i]);
                    float gw2 = gu * t[i];
                    float gh = gu * s1[i];
#pragma unroll
                    for (int a = 0; a < DIM; ++a) {
                        const float hd = h[i][1 + a][pt], hdd = h[i][1 + DIM + a][pt];
                        const float g1 = g * p.c1[a], g2 = g * p.c2[a];
                        gw2 += g1 * s1[i] * hd + g2 * (s2[i] * hd * hd + s1[i] * hdd);
                        gh += g1 * s2[i] * hd + g2 * (s3 * hd * hd + s2[i] * hdd);
                        h[i][1 + a][pt] = w2k[i] * (g1 * s1[i] + g2 * 2.f * s2[i] * hd);
                        h[i][1 + DIM + a][pt] = w2k[i] * g2 * s1[i];
                    }
                    gh *= w2k[i];
                    h[i][0][pt] = gh;
                    gb1acc[i] += gh;
                    gw2acc[i] += gw2;
                }
            }
            // ---- D: weight gradient gW1[k][c] += sum_j sum_pt gH_j[k][pt] * z_j[c][pt]
            // (compiler barrier: re-read z from shared memory instead of keeping 4*J*C values live)
            asm volatile("" ::: "memory");
#pragma unroll
            for (int j = 0; j < J; ++j)
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const float4 z = tile4[(j * C + c) * TP4 + grp];
#pragma unroll
                    for (int i = 0; i < KPL; ++i) {
                        gW1acc[i][c] = fmaf(h[i][j][0], z.x, gW1acc[i][c]);
                        gW1acc[i][c] = fmaf(h[i][j][1], z.y, gW1acc[i][c]);
                        gW1acc[i][c] = fmaf(h[i][j][2], z.z, gW1acc[i][c]);
                        gW1acc[i][c] = fmaf(h[i][j][3], z.w, gW1acc[i][c]);
                    }
                }
            // ---- C: exchange gH over the hidden units, then gZ_j[c] = sum_k W1[k][c] gH_j[k]
#pragma unroll
            for (int j = 0; j < J; ++j)
#pragma unroll
                for (int i = 0; i < KPL; ++i)
                    xch[j * K + KPL * kq + i] = make_float4(h[i][j][0], h[i][j][1], h[i][j][2], h[i][j][3]);
            __syncwarp();                        // also: every lane has finished reading z of this group
#pragma unroll
            for (int ci = 0; ci < CPL; ++ci) {
                const int c = kq + HEAD_LPG * ci;
                if (c < C) {
#pragma unroll 1
                    for (int j = 0; j < J; ++j) {
                        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int k4 = 0; k4 < K / 4; ++k4) {
                            const float4 wc = w1t4[c * (K / 4) + k4];           // W1[4 k4 .. 4 k4 + 3][c]
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                const float4 gq = xch[j * K + 4 * k4 + kk];
                                const float w = kk == 0 ? wc.x : kk == 1 ? wc.y : kk == 2 ? wc.z : wc.w;
                                acc.x = fmaf(w, gq.x, acc.x);
                                acc.y = fmaf(w, gq.y, acc.y);
                                acc.z = fmaf(w, gq.z, acc.z);
                                acc.w = fmaf(w, gq.w, acc.w);
                            }
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
        __syncthreads();                         // the buffer is free for the prefetch after next
        buf ^= 1;
    }

    // ---- parameter gradients and loss: point groups of a warp (shuffles) -> warps (shared memory)
    // -> one atomic per block and element
    float* red = reinterpret_cast<float*>(smem4);            // [4 warps][K*C + 2K + 2], reuses the tile
    constexpr int RW = K * C + 2 * K + 2;
#pragma unroll
    for (int o = HEAD_LPG; o < 32; o <<= 1) {
#pragma unroll
        for (int i = 0; i < KPL; ++i) {
#pragma unroll
            for (int c = 0; c < C; ++c) gW1acc[i][c] += __shfl_xor_sync(0xffffffffu, gW1acc[i][c], o);
            gb1acc[i] += __shfl_xor_sync(0xffffffffu, gb1acc[i], o);
            gw2acc[i] += __shfl_xor_sync(0xffffffffu, gw2acc[i], o);
        }
        gb2acc += __shfl_xor_sync(0xffffffffu, gb2acc, o);
        lossacc += __shfl_xor_sync(0xffffffffu, lossacc, o);
    }
    if (lane < HEAD_LPG) {
        float* rw = red + warp * RW;
#pragma unroll
        for (int i = 0; i < KPL; ++i) {
#pragma unroll
            for (int c = 0; c < C; ++c) rw[(KPL * kq + i) * C + c] = gW1acc[i][c];
            rw[K * C + KPL * kq + i] = gb1acc[i];
            rw[K * C + K + KPL * kq + i] = gw2acc[i];
        }
        if (kq == 0) { rw[K * C + 2 * K] = gb2acc; rw[K * C + 2 * K + 1] = lossacc; }
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
