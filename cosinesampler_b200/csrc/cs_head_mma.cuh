// cs_head_mma.cuh -- the fused PDE-residual head (cs_head.cuh) with its three small contractions
// on the tensor cores.
//
// Per tile of points the head is three GEMM-shaped products with a 16-wide hidden layer,
//     A   H_j  [16 k x p]  = W1   [16 k x C]  . Z_j  [C x p]          (j = value, d/da, d2/da2)
//     D   gW1  [16 k x C] += gH_j [16 k x p]  . Z_j^T [p x C]
//     C   gZ_j [C x p]     = W1^T [C x 16 k]  . gH_j [16 k x p]
// and an elementwise stage between A and D/C (tanh and its derivatives, the residual, the loss).
// The SIMT version (cs_head.cuh) is bound by the shared-memory pipe: every FFMA needs an operand
// that another lane also needs, and a broadcast LDS.128 costs four wavefronts whatever it delivers
// (0.53-0.62 ms per 2^20 points, ncu: shared-memory wavefronts 80 % busy; profiles/README.md).
// Here the operands are mma.sync.m16n8k8 fragments: each shared-memory wavefront delivers 32
// distinct words, the accumulators of A land in exactly the (k, p) layout the elementwise stage
// wants, gH is the A operand of D without moving (the contraction index p may be permuted), and
// only C needs an exchange through shared memory.
//
// Precision: fp32 parity is kept with the error-compensated 3xTF32 scheme -- every operand x is
// split into big = tf32(x) and small = x - big and a product is accumulated as
// small*big + big*small + big*big in fp32 -- which carries 21-22 mantissa bits per operand.
#pragma once
#include "cs_head.cuh"

namespace cs {

constexpr int HM_THREADS = 128;
constexpr int HM_TP = 64;               // points per tile
constexpr int HM_TPS = HM_TP + 8;       // row stride of the tile in floats: 72 = 8 mod 32 -> conflict-free fragments
constexpr int HM_K = 16;

// big = x rounded to the 10 explicit mantissa bits of TF32 (round half away from zero on the magnitude
// bits, the same rounding as cvt.rna.tf32.f32, which ptxas expands into a 4-instruction sequence with a
// NaN/Inf guard that finite data does not need); small = x - big exactly, handed to the tensor core as
// it is (the hardware ignores its low 13 mantissa bits, an error below 2^-21 |x|).
struct Tf32x2 { uint32_t big, small; };
__device__ __forceinline__ Tf32x2 split_tf32(float x) {
    Tf32x2 r;
    r.big = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
    r.small = __float_as_uint(x - __uint_as_float(r.big));
    return r;
}

// tanh from one ex2.approx and one rcp.approx: absolute error a few 1e-7 (|tanh| <= 1 is the scale
// that matters downstream: u = w2 . tanh, s1 = 1 - tanh^2).  The SIMT kernel keeps the expm1f form.
__device__ __forceinline__ float tanh_fast(float x) {
    const float e = __expf(-2.f * fabsf(x));
    return copysignf(__fdividef(1.f - e, 1.f + e), x);
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int DIM, int C> struct HeadMmaSmem {
    static constexpr int J = 1 + 2 * DIM;
    static constexpr int ROWS = J * C;
    static constexpr int TILE_F = ROWS * HM_TPS;                          // one tile buffer (floats)
    static constexpr int XCH_F = (HM_THREADS / 32) * J * HM_K * 8;        // per warp: gH [J*16][8 points]
    static constexpr int RED_F = (HM_THREADS / 32) * (HM_K * C + 2 * HM_K + 2);
    static constexpr size_t BYTES = (size_t)(2 * TILE_F + XCH_F) * 4;     // two tiles: the next one is prefetched
};

template <int DIM, int C>
__global__ void __launch_bounds__(HM_THREADS, (C <= 16 ? 4 : 2))
cs_pde_head_mma_kernel(const HeadParams p) {
    using HS = HeadMmaSmem<DIM, C>;
    constexpr int J = HS::J;
    constexpr int K = HM_K;
    constexpr int ROWS = HS::ROWS;
    constexpr int TP = HM_TP;
    constexpr int TPS = HM_TPS;
    constexpr int TP4 = TP / 4;
    constexpr int TPS4 = TPS / 4;
    constexpr int KSA = C / 8;                  // k-steps of product A (contraction over the C channels)
    constexpr int NBD = C / 8;                  // n-blocks of product D (C output columns)
    constexpr int MBC = (C + 15) / 16;          // m-blocks of product C (C output rows)
    static_assert(C % 8 == 0, "the tensor-core head needs C % 8 == 0");

    extern __shared__ float4 smem4[];
    float* tiles = reinterpret_cast<float*>(smem4);                   // [2][ROWS][TPS]
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int g = lane >> 2;                    // fragment group id
    const int t = lane & 3;                     // thread id in group
    float* xch = tiles + 2 * HS::TILE_F + warp * (J * K * 8);         // [J*16][8]

    // ---- constant fragments of W1 (product A: rows k, columns c) and W1^T (product C: rows c, columns k)
    uint32_t wab[KSA][4], was[KSA][4];
#pragma unroll
    for (int s = 0; s < KSA; ++s) {
        const float v[4] = {__ldg(p.W1 + g * C + 8 * s + t), __ldg(p.W1 + (g + 8) * C + 8 * s + t),
                            __ldg(p.W1 + g * C + 8 * s + t + 4), __ldg(p.W1 + (g + 8) * C + 8 * s + t + 4)};
#pragma unroll
        for (int e = 0; e < 4; ++e) { const Tf32x2 sp = split_tf32(v[e]); wab[s][e] = sp.big; was[s][e] = sp.small; }
    }
    uint32_t wcb[MBC][2][4], wcs[MBC][2][4];
#pragma unroll
    for (int mb = 0; mb < MBC; ++mb)
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const int c0 = 16 * mb + g, c1 = 16 * mb + g + 8;
            const int k0 = 8 * ks + t, k1 = 8 * ks + t + 4;
            const float v[4] = {c0 < C ? __ldg(p.W1 + k0 * C + c0) : 0.f, c1 < C ? __ldg(p.W1 + k0 * C + c1) : 0.f,
                                c0 < C ? __ldg(p.W1 + k1 * C + c0) : 0.f, c1 < C ? __ldg(p.W1 + k1 * C + c1) : 0.f};
#pragma unroll
            for (int e = 0; e < 4; ++e) { const Tf32x2 sp = split_tf32(v[e]); wcb[mb][ks][e] = sp.big; wcs[mb][ks][e] = sp.small; }
        }
    const float b1g[2] = {__ldg(p.b1 + g), __ldg(p.b1 + g + 8)};
    const float w2g[2] = {__ldg(p.w2 + g), __ldg(p.w2 + g + 8)};
    const float b2 = __ldg(p.b2);

    float accW[NBD][4];                         // gW1 fragments: rows k = g, g+8; columns c = 8 nb + 2t, 2t+1
    float accWs[NBD][4];                        // the small (correction) terms: a second accumulator chain
#pragma unroll
    for (int nb = 0; nb < NBD; ++nb) {
        accW[nb][0] = accW[nb][1] = accW[nb][2] = accW[nb][3] = 0.f;
        accWs[nb][0] = accWs[nb][1] = accWs[nb][2] = accWs[nb][3] = 0.f;
    }
    float gb1acc[2] = {0.f, 0.f}, gw2acc[2] = {0.f, 0.f};
    float gb2acc = 0.f, lossacc = 0.f;

    auto prefetch = [&](long long tile, float* dst) {
        const long long p0 = tile * TP;
        const unsigned sbase = (unsigned)__cvta_generic_to_shared(dst);
        for (int idx = tid; idx < ROWS * TP4; idx += HM_THREADS) {
            const int r = idx / TP4;
            const int v = idx - r * TP4;
            const long long pp = p0 + 4 * v;
            const float* src = p.jets + (long long)r * p.P + pp;
            if (p.vec) {
                const bool ok = pp < p.P;
                cp16z(sbase + (r * TPS4 + v) * 16, ok ? src : p.jets, ok);
            } else {
                float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
                if (pp < p.P) val.x = __ldcs(src);
                if (pp + 1 < p.P) val.y = __ldcs(src + 1);
                if (pp + 2 < p.P) val.z = __ldcs(src + 2);
                if (pp + 3 < p.P) val.w = __ldcs(src + 3);
                reinterpret_cast<float4*>(dst)[r * TPS4 + v] = val;
            }
        }
        cp_async_commit();
    };

    const long long ntiles = (p.P + TP - 1) / TP;
    int buf = 0;
    if ((long long)blockIdx.x < ntiles) prefetch(blockIdx.x, tiles);
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long p0 = tile * TP;
        float* tl = tiles + buf * HS::TILE_F;
        const long long next = tile + gridDim.x;
        if (next < ntiles) { prefetch(next, tiles + (buf ^ 1) * HS::TILE_F); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();

        for (int nb8 = warp; nb8 < TP / 8; nb8 += HM_THREADS / 32) {
            const int pb = 8 * nb8;                             // first point of this block in the tile
            // ---- A: H_j = W1 Z_j.  B fragment: Z_j[c = 8s + t (+4)][p = g]
            float hf[J][4];
#pragma unroll
            for (int j = 0; j < J; ++j) hf[j][0] = hf[j][1] = hf[j][2] = hf[j][3] = 0.f;
            // the J accumulator chains are independent: issue the three 3xTF32 terms jet by jet so that
            // consecutive MMAs never wait for each other
#pragma unroll
            for (int s = 0; s < KSA; ++s) {
                Tf32x2 b0[J], b1[J];
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    b0[j] = split_tf32(tl[(j * C + 8 * s + t) * TPS + pb + g]);
                    b1[j] = split_tf32(tl[(j * C + 8 * s + t + 4) * TPS + pb + g]);
                }
#pragma unroll
                for (int j = 0; j < J; ++j) mma_tf32(hf[j], was[s], b0[j].big, b1[j].big);
#pragma unroll
                for (int j = 0; j < J; ++j) mma_tf32(hf[j], wab[s], b0[j].small, b1[j].small);
#pragma unroll
                for (int j = 0; j < J; ++j) mma_tf32(hf[j], wab[s], b0[j].big, b1[j].big);
            }
            // hf[j][2*half + e] = H_j[k = g + 8 half][p = 2t + e]
            // ---- B: per point (e = 0, 1): sums over the 16 hidden units = this lane's two + the 8 groups
            float gsc[2], g1c[2][DIM], g2c[2][DIM];             // per point: gu and g*c1[a], g*c2[a]
            float tt[2][2], s1[2][2], s2[2][2];                 // [half][e]
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                float pu = 0.f, pua[DIM], puaa[DIM];
#pragma unroll
                for (int a = 0; a < DIM; ++a) { pua[a] = 0.f; puaa[a] = 0.f; }
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int q = 2 * half + e;
                    const float th = tanh_fast(hf[0][q] + b1g[half]);
                    tt[half][e] = th;
                    s1[half][e] = 1.f - th * th;
                    s2[half][e] = -2.f * th * s1[half][e];
                    pu = fmaf(w2g[half], th, pu);
#pragma unroll
                    for (int a = 0; a < DIM; ++a) {
                        const float hd = hf[1 + a][q], hdd = hf[1 + DIM + a][q];
                        pua[a] = fmaf(w2g[half], s1[half][e] * hd, pua[a]);
                        puaa[a] = fmaf(w2g[half], s2[half][e] * hd * hd + s1[half][e] * hdd, puaa[a]);
                    }
                }
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) {
                    pu += __shfl_xor_sync(0xffffffffu, pu, o);
#pragma unroll
                    for (int a = 0; a < DIM; ++a) {
                        pua[a] += __shfl_xor_sync(0xffffffffu, pua[a], o);
                        puaa[a] += __shfl_xor_sync(0xffffffffu, puaa[a], o);
                    }
                }
                const float u = pu + b2;
                float f = p.c_u * u + p.c_u3 * u * u * u;
#pragma unroll
                for (int a = 0; a < DIM; ++a) f += p.c1[a] * pua[a] + p.c2[a] * puaa[a];
                const long long pi = p0 + pb + 2 * t + e;
                const bool valid = pi < p.P;
                const float gg = valid ? 2.f * p.scale * f : 0.f;
                if (g == 0 && valid) {
                    lossacc += f * f;
                    if (p.f_out) p.f_out[pi] = f;
                }
                gsc[e] = gg * (p.c_u + 3.f * p.c_u3 * u * u);
                if (g == 0) gb2acc += gsc[e];
#pragma unroll
                for (int a = 0; a < DIM; ++a) { g1c[e][a] = gg * p.c1[a]; g2c[e][a] = gg * p.c2[a]; }
            }
            // gradient w.r.t. h / hd / hdd, in place
#pragma unroll
            for (int half = 0; half < 2; ++half)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int q = 2 * half + e;
                    const float th = tt[half][e], d1 = s1[half][e], d2 = s2[half][e];
                    const float d3 = -2.f * (d1 * d1 + th * d2);
                    float gw2 = gsc[e] * th;
                    float gh = gsc[e] * d1;
#pragma unroll
                    for (int a = 0; a < DIM; ++a) {
                        const float hd = hf[1 + a][q], hdd = hf[1 + DIM + a][q];
                        gw2 += g1c[e][a] * d1 * hd + g2c[e][a] * (d2 * hd * hd + d1 * hdd);
                        gh += g1c[e][a] * d2 * hd + g2c[e][a] * (d3 * hd * hd + d2 * hdd);
                        hf[1 + a][q] = w2g[half] * (g1c[e][a] * d1 + g2c[e][a] * 2.f * d2 * hd);
                        hf[1 + DIM + a][q] = w2g[half] * g2c[e][a] * d1;
                    }
                    gh *= w2g[half];
                    hf[0][q] = gh;
                    gb1acc[half] += gh;
                    gw2acc[half] += gw2;
                }
            // ---- D: gW1 += gH_j Z_j^T.  The contraction runs over the 8 points in the order
            // (0,2,4,6,1,3,5,7), so that the accumulator fragment of A is the A fragment here:
            // a0 = (g, 2t) a1 = (g+8, 2t) a2 = (g, 2t+1) a3 = (g+8, 2t+1); B: Z_j[c = 8nb + g][p = 2t, 2t+1]
#pragma unroll
            for (int j = 0; j < J; ++j) {
                uint32_t ab[4], as[4];
                const float av[4] = {hf[j][0], hf[j][2], hf[j][1], hf[j][3]};
#pragma unroll
                for (int e = 0; e < 4; ++e) { const Tf32x2 sp = split_tf32(av[e]); ab[e] = sp.big; as[e] = sp.small; }
                Tf32x2 z0[NBD], z1[NBD];
#pragma unroll
                for (int nb = 0; nb < NBD; ++nb) {
                    const float2 z = *reinterpret_cast<const float2*>(tl + (j * C + 8 * nb + g) * TPS + pb + 2 * t);
                    z0[nb] = split_tf32(z.x);
                    z1[nb] = split_tf32(z.y);
                }
#pragma unroll
                for (int nb = 0; nb < NBD; ++nb) mma_tf32(accWs[nb], as, z0[nb].big, z1[nb].big);
#pragma unroll
                for (int nb = 0; nb < NBD; ++nb) mma_tf32(accW[nb], ab, z0[nb].big, z1[nb].big);
#pragma unroll
                for (int nb = 0; nb < NBD; ++nb) mma_tf32(accWs[nb], ab, z0[nb].small, z1[nb].small);
            }
            // ---- C: gZ_j = W1^T gH_j.  gH goes through shared memory: [j*16 + k][8 points]
#pragma unroll
            for (int j = 0; j < J; ++j) {
                *reinterpret_cast<float2*>(xch + (j * K + g) * 8 + 2 * t) = make_float2(hf[j][0], hf[j][1]);
                *reinterpret_cast<float2*>(xch + (j * K + g + 8) * 8 + 2 * t) = make_float2(hf[j][2], hf[j][3]);
            }
            __syncwarp();                        // also: every lane has finished reading z of this point block
#pragma unroll
            for (int mb = 0; mb < MBC; ++mb) {
                float d[J][4];
#pragma unroll
                for (int j = 0; j < J; ++j) d[j][0] = d[j][1] = d[j][2] = d[j][3] = 0.f;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    Tf32x2 b0[J], b1[J];
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        b0[j] = split_tf32(xch[(j * K + 8 * ks + t) * 8 + g]);
                        b1[j] = split_tf32(xch[(j * K + 8 * ks + t + 4) * 8 + g]);
                    }
#pragma unroll
                    for (int j = 0; j < J; ++j) mma_tf32(d[j], wcs[mb][ks], b0[j].big, b1[j].big);
#pragma unroll
                    for (int j = 0; j < J; ++j) mma_tf32(d[j], wcb[mb][ks], b0[j].small, b1[j].small);
#pragma unroll
                    for (int j = 0; j < J; ++j) mma_tf32(d[j], wcb[mb][ks], b0[j].big, b1[j].big);
                }
                // d[j][2*half + e] = gZ_j[c = 16mb + g + 8 half][p = 2t + e]: overwrites z in the tile
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    if (16 * mb + g < C)
                        *reinterpret_cast<float2*>(tl + (j * C + 16 * mb + g) * TPS + pb + 2 * t) = make_float2(d[j][0], d[j][1]);
                    if (16 * mb + g + 8 < C)
                        *reinterpret_cast<float2*>(tl + (j * C + 16 * mb + g + 8) * TPS + pb + 2 * t) = make_float2(d[j][2], d[j][3]);
                }
            }
            __syncwarp();                        // xch is reused by the next point block
        }
        __syncthreads();

        // ---- the tile now holds d loss / d jets: store it
        for (int idx = tid; idx < ROWS * TP4; idx += HM_THREADS) {
            const int r = idx / TP4;
            const int v = idx - r * TP4;
            const long long pp = p0 + 4 * v;
            float* dst = p.gjets + (long long)r * p.P + pp;
            const float4 val = reinterpret_cast<const float4*>(tl)[r * TPS4 + v];
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

    // ---- parameter gradients and loss: lanes -> warps (shared memory) -> one atomic per block and element
    float* red = tiles;                          // [4 warps][K*C + 2K + 2], reuses the tile buffers
    constexpr int RW = K * C + 2 * K + 2;
    // gb1 / gw2: the four lanes of a group hold partial sums over different points
#pragma unroll
    for (int half = 0; half < 2; ++half) {
#pragma unroll
        for (int o = 1; o < 4; o <<= 1) {
            gb1acc[half] += __shfl_xor_sync(0xffffffffu, gb1acc[half], o);
            gw2acc[half] += __shfl_xor_sync(0xffffffffu, gw2acc[half], o);
        }
    }
    // loss / gb2 live in the lanes of group 0
#pragma unroll
    for (int o = 1; o < 4; o <<= 1) {
        gb2acc += __shfl_xor_sync(0xffffffffu, gb2acc, o);
        lossacc += __shfl_xor_sync(0xffffffffu, lossacc, o);
    }
    {
        float* rw = red + warp * RW;
#pragma unroll
        for (int nb = 0; nb < NBD; ++nb) {
            rw[g * C + 8 * nb + 2 * t] = accW[nb][0] + accWs[nb][0];
            rw[g * C + 8 * nb + 2 * t + 1] = accW[nb][1] + accWs[nb][1];
            rw[(g + 8) * C + 8 * nb + 2 * t] = accW[nb][2] + accWs[nb][2];
            rw[(g + 8) * C + 8 * nb + 2 * t + 1] = accW[nb][3] + accWs[nb][3];
        }
        if (t == 0) {
            rw[K * C + g] = gb1acc[0]; rw[K * C + g + 8] = gb1acc[1];
            rw[K * C + K + g] = gw2acc[0]; rw[K * C + K + g + 8] = gw2acc[1];
        }
        if (lane == 0) { rw[K * C + 2 * K] = gb2acc; rw[K * C + 2 * K + 1] = lossacc; }
    }
    __syncthreads();
    for (int e = tid; e < RW; e += HM_THREADS) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < HM_THREADS / 32; ++w) s += red[w * RW + e];
        float* dst = (e < K * C) ? p.gW1 + e : (e < K * C + K) ? p.gb1 + (e - K * C)
                   : (e < K * C + 2 * K) ? p.gw2 + (e - K * C - K)
                   : (e == K * C + 2 * K) ? p.gb2 : p.loss_sum;
        atomicAdd(dst, s);
    }
}

template <int DIM, int C>
cudaError_t launch_head_mma(const HeadParams& p, cudaStream_t stream) {
    using HS = HeadMmaSmem<DIM, C>;
    auto kern = cs_pde_head_mma_kernel<DIM, C>;
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
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, HM_THREADS, smem);
        if (e != cudaSuccess) return e;
        if (occ < 1) occ = 1;
    }
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    long long blocks = (p.P + HM_TP - 1) / HM_TP;
    const long long cap = (long long)sms * occ;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) return cudaSuccess;
    kern<<<(unsigned)blocks, HM_THREADS, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace cs
