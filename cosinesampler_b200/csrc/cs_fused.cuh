// cs_fused.cuh -- the whole PIXEL training step for a Linear(C,K)-Tanh-Linear(K,1) head in ONE pass over
// the collocation points: gather -> head / residual / gradients -> scatter (SURVEY section 8f ranks 1 + 2,
// VERDICT r1 "next" items 1 and 2).
//
// What the reference does per step (test_2d.py:36-127, modules_2d.py:22-111): 14-20 operator launches that
// re-gather the same corners, the caller's replicate / sum-over-cells / MLP head under three levels of
// autograd, and one scalar red.global per (point, corner, channel) for every scatter (cu2d:340-354,464-473).
//
// Three observations restructure it:
//
//  1. The sampler is linear in the cells and so is the head's first layer, so the two commute:
//         W1 . (sum_n sum_q V[n,q] coef_j,q)  =  sum_n sum_q (W1 . V[n,q]) coef_j,q .
//     Mixing the CELLS with W1 once per step (cs_head_premix: Vh = W1 . V, a grid-sized pass, 1/128 of the
//     per-point work at 2^25 points) makes the gather deliver the hidden pre-activations H_j directly; the
//     adjoint scatters d loss / d H_j into gVh and one grid-sized pass (cs_head_postmix) returns
//     gInput = W1^T . gVh and gW1 = sum_texels gVh (x) V.  The three per-point matrix products of the head
//     (cs_head_mma.cuh: 3 x J x K x C MACs per point, 3xTF32 on the tensor cores, 0.23 ms per 2^20 points)
//     disappear; what is left per point is the elementwise tanh / residual stage, which runs in the very
//     registers the gather accumulates into.  No jets, no d loss / d jets in HBM: the per-point HBM traffic of
//     the step is the coordinates (4*dim bytes).
//  2. Loss and gradients do not depend on the order of the points, so the step may bin them by texel first
//     (cs_bin_points: counting sort on a tile-major texel key).  Consecutive points then hit the same corners.
//  3. With binned points the scatter pre-reduces before it touches L2: the points are binned on a key that
//     also holds the sub-texel quadrant (for N cells offset by n/N of a texel, every cell's corners are the same
//     for all points of one sub-bin), each walker (the L = K/4 lanes that share a point) takes consecutive
//     binned points, and contributions of consecutive points with identical corners are summed in registers
//     and leave as ONE red.global.add.v4.f32 per corner and run instead of one per corner and point.
//     Measured alternatives (profiles/README.md): walker-private 3x3-texel windows in shared memory (plain
//     ld/st.shared read-modify-write, flushed when the walker moves on) cut the reds 7-28x but cost 22 KB of
//     shared memory per warp (6 warps per SM) and a serial LDS->FADD->STS chain per corner: 1.29 ms per 2^20
//     points against 0.53 ms without; ATOMS-based privatisation costs 2 cycles per lane; __match_any_sync
//     needs the points of a warp to agree in corner AND sub-texel shift plus a segmented shuffle per value.
#pragma once
#include <limits.h>
#include <stdlib.h>
#include <string.h>

#include "cs_jet.cuh"

namespace cs {

// cells per unrolled round of the one-pass kernel's gather (0: all NC cells; experiments)
#ifndef CS_FUSED_GATHER_UNROLL
#define CS_FUSED_GATHER_UNROLL 0
#endif
constexpr int FUSED_GATHER_UNROLL = CS_FUSED_GATHER_UNROLL;

// ---------------------------------------------------------------------------------------------------------
// Point binning: counting sort of the coordinates on a tile-major texel key (+ sub-texel quadrant)
// ---------------------------------------------------------------------------------------------------------
struct BinParams {
    int dim;
    int size[3];
    int shift;                 // texel coordinates are coarsened by >> shift so that nbins stays bounded
    int sub;                   // log2 of the sub-bins per texel and axis (0, 1, 2)
    int ntx, nty, ntz;         // tiles per axis (2D: 8x8 texels, 3D: 4x4x4 texels per tile)
    unsigned nbins;
    long long P;
    const float* coords;       // [P, dim]
    const float* offset;       // [N] (device); cell 0's offset enters the key.  nullable = 0
    int align, multicell, index_mode;
};
constexpr int BIN_SCAN_CHUNK = 2048;   // bins per block of the two-level scan

// low-corner texel of cell 0 along one axis, clamped into the cell, and the sub-bin of the fractional position:
// a locality key, never a correctness matter (the fused kernel compares the real corner indices of every point)
__device__ __forceinline__ int bin_axis(float g, int size, float off0, const BinParams& p, int& subbin) {
    float i;
    if (p.align) {
        const float sf = (float)(size - 1 - (p.multicell ? 1 : 0));
        const float h = __fmul_rn(__fadd_rn(g, 1.f), 0.5f);
        i = (p.index_mode == 0) ? __fadd_rn(__fmul_rn(h, sf), off0) : __fmaf_rn(h, sf, off0);
    } else {
        const float sf = (float)size;
        i = __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(g, 1.f), sf), -1.f), 0.5f), off0);
    }
    if (!(fabsf(i) < 1.0e9f)) i = 0.f;
    const float lf = floorf(i);
    int l = (int)lf;
    subbin = min((1 << p.sub) - 1, max(0, (int)((i - lf) * (float)(1 << p.sub))));
    if (l < 0) { l = 0; subbin = 0; }
    if (l > size - 1) { l = size - 1; subbin = 0; }
    return l >> p.shift;
}

__device__ __forceinline__ unsigned bin_key(const float* gp, const BinParams& p) {
    const float off0 = p.offset ? __ldg(p.offset) : 0.f;
    int sx, sy, sz = 0;
    const int lx = bin_axis(__ldg(gp), p.size[0], off0, p, sx);
    const int ly = bin_axis(__ldg(gp + 1), p.size[1], off0, p, sy);
    unsigned key;
    if (p.dim == 2) {
        const unsigned tile = (unsigned)((ly >> 3) * p.ntx + (lx >> 3));
        key = (tile << 6) | (unsigned)(((ly & 7) << 3) | (lx & 7));
        return (key << (2 * p.sub)) | (unsigned)((sy << p.sub) | sx);
    }
    const int lz = bin_axis(__ldg(gp + 2), p.size[2], off0, p, sz);
    const unsigned tile = (unsigned)(((lz >> 2) * p.nty + (ly >> 2)) * p.ntx + (lx >> 2));
    key = (tile << 6) | (unsigned)(((lz & 3) << 4) | ((ly & 3) << 2) | (lx & 3));
    return (key << (3 * p.sub)) | (unsigned)((((sz << p.sub) | sy) << p.sub) | sx);
}

// Both sweeps over the points keep BIN_ILP independent points in flight per thread: with one load per thread
// the resident threads hold ~1 MB in flight and the sweeps ran at 1 TB/s (latency-bound).
constexpr int BIN_ILP = 4;

static __global__ void __launch_bounds__(256) cs_bin_count_kernel(const BinParams p, unsigned* __restrict__ hist,
                                                                  unsigned* __restrict__ keys,
                                                                  unsigned* __restrict__ rank) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < p.P; i0 += BIN_ILP * stride) {
        unsigned key[BIN_ILP];
#pragma unroll
        for (int u = 0; u < BIN_ILP; ++u) {
            const long long i = i0 + u * stride;
            key[u] = (i < p.P) ? bin_key(p.coords + i * p.dim, p) : 0u;
        }
        unsigned r[BIN_ILP];
#pragma unroll
        for (int u = 0; u < BIN_ILP; ++u) r[u] = (i0 + u * stride < p.P) ? atomicAdd(hist + key[u], 1u) : 0u;
#pragma unroll
        for (int u = 0; u < BIN_ILP; ++u)
            if (i0 + u * stride < p.P) { keys[i0 + u * stride] = key[u]; rank[i0 + u * stride] = r[u]; }
    }
}

// two-level exclusive scan: every block scans BIN_SCAN_CHUNK bins in place and writes its total; one block
// scans the totals; the scatter adds both
static __global__ void __launch_bounds__(1024) cs_bin_scan_chunk_kernel(unsigned* __restrict__ hist, unsigned nbins,
                                                                        unsigned* __restrict__ totals) {
    __shared__ unsigned part[1024];
    const unsigned b0 = blockIdx.x * BIN_SCAN_CHUNK + 2 * threadIdx.x;
    const unsigned c0 = b0 < nbins ? hist[b0] : 0u;
    const unsigned c1 = b0 + 1 < nbins ? hist[b0 + 1] : 0u;
    const unsigned s = c0 + c1;
    part[threadIdx.x] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const unsigned v = (threadIdx.x >= (unsigned)o) ? part[threadIdx.x - o] : 0u;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    const unsigned excl = part[threadIdx.x] - s;
    if (b0 < nbins) hist[b0] = excl;
    if (b0 + 1 < nbins) hist[b0 + 1] = excl + c0;
    if (threadIdx.x == 1023) totals[blockIdx.x] = part[1023];
}

static __global__ void __launch_bounds__(1024) cs_bin_scan_totals_kernel(unsigned* __restrict__ totals, unsigned n) {
    __shared__ unsigned part[1024];
    const unsigned per = (n + 1023u) / 1024u;
    const unsigned b0 = threadIdx.x * per;
    const unsigned b1 = min(n, b0 + per);
    unsigned s = 0;
    for (unsigned b = b0; b < b1; ++b) s += totals[b];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const unsigned v = (threadIdx.x >= (unsigned)o) ? part[threadIdx.x - o] : 0u;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    unsigned run = part[threadIdx.x] - s;
    for (unsigned b = b0; b < b1; ++b) {
        const unsigned c = totals[b];
        totals[b] = run;
        run += c;
    }
}

// The destinations are random: a single sweep writes 4*dim bytes into random 32-byte sectors of a buffer larger
// than L2 and every sector is read-modified-written in DRAM (1.19 ms for 2^25 points).  The sweep is therefore
// repeated per destination window [lo, hi) small enough to stay in L2, where the partial sectors merge before
// they are written back.  The first sweep turns (key, rank) into the final position and stores it over the rank,
// so that the later sweeps read 4 bytes per point and the coordinates only of the points they place.
template <int DIM>
static __global__ void __launch_bounds__(256) cs_bin_scatter_kernel(const float* __restrict__ coords, long long P, int vec2,
                                                                    const unsigned* __restrict__ offs,
                                                                    const unsigned* __restrict__ totals,
                                                                    const unsigned* __restrict__ keys,
                                                                    unsigned* __restrict__ rank, int first,
                                                                    unsigned lo, unsigned hi,
                                                                    float* __restrict__ sorted, int* __restrict__ perm) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < P; i0 += BIN_ILP * stride) {
        unsigned pos[BIN_ILP];
        if (first) {
            unsigned key[BIN_ILP], r[BIN_ILP];
#pragma unroll
            for (int u = 0; u < BIN_ILP; ++u) {
                const long long i = i0 + u * stride;
                key[u] = (i < P) ? __ldcs(keys + i) : 0u;
                r[u] = (i < P) ? rank[i] : 0u;
            }
#pragma unroll
            for (int u = 0; u < BIN_ILP; ++u) {
                pos[u] = __ldg(offs + key[u]) + __ldg(totals + key[u] / BIN_SCAN_CHUNK) + r[u];
                if (i0 + u * stride < P) rank[i0 + u * stride] = pos[u];
                else pos[u] = 0xffffffffu;
            }
        } else {
#pragma unroll
            for (int u = 0; u < BIN_ILP; ++u) pos[u] = (i0 + u * stride < P) ? __ldcs(rank + i0 + u * stride) : 0xffffffffu;
        }
#pragma unroll
        for (int u = 0; u < BIN_ILP; ++u) {
            if (pos[u] >= lo && pos[u] < hi) {
                const unsigned i = (unsigned)(i0 + u * stride);
                if (DIM == 2 && vec2) {
                    reinterpret_cast<float2*>(sorted)[pos[u]] = __ldg(reinterpret_cast<const float2*>(coords) + i);
                } else {
                    const float* gp = coords + (size_t)i * DIM;
                    float* dst = sorted + (size_t)pos[u] * DIM;
#pragma unroll
                    for (int a = 0; a < DIM; ++a) dst[a] = __ldg(gp + a);
                }
                if (perm) perm[pos[u]] = (int)i;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Grid-sized mixes with the head's first layer
// ---------------------------------------------------------------------------------------------------------
constexpr int MIX_MAXK = 64;

// Vh[n, t, k] = sum_c W1[k, c] * V[n, c, t]   (channel-first in, channel-last out: the staging transpose of
// cs_to_channel_last and the first Linear layer in one pass); Vh[N*T, :] = 0 (the texel out-of-bounds corners read)
template <int K>
__global__ void __launch_bounds__(256) cs_head_premix_kernel(const float* __restrict__ V, const float* __restrict__ W1,
                                                             float* __restrict__ Vh, int C, long long T, long long NT) {
    extern __shared__ float w1s[];           // [C][K]: w1s[c*K + k] = W1[k*C + c]
    for (int e = threadIdx.x; e < K * C; e += blockDim.x) w1s[(e % C) * K + (e / C)] = __ldg(W1 + e);
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x < K) Vh[NT * K + threadIdx.x] = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < NT;
         i += (long long)gridDim.x * blockDim.x) {
        const long long n = i / T, t = i - n * T;
        const float* src = V + n * C * T + t;
        float out[K];
#pragma unroll
        for (int k = 0; k < K; ++k) out[k] = 0.f;
        for (int c = 0; c < C; ++c) {
            const float v = __ldg(src + (long long)c * T);
            const float4* wr = reinterpret_cast<const float4*>(w1s + c * K);
#pragma unroll
            for (int k4 = 0; k4 < K / 4; ++k4) {
                const float4 w = wr[k4];
                out[4 * k4] = fmaf(w.x, v, out[4 * k4]);
                out[4 * k4 + 1] = fmaf(w.y, v, out[4 * k4 + 1]);
                out[4 * k4 + 2] = fmaf(w.z, v, out[4 * k4 + 2]);
                out[4 * k4 + 3] = fmaf(w.w, v, out[4 * k4 + 3]);
            }
        }
        float4* dst = reinterpret_cast<float4*>(Vh + i * K);
#pragma unroll
        for (int k4 = 0; k4 < K / 4; ++k4)
            dst[k4] = make_float4(out[4 * k4], out[4 * k4 + 1], out[4 * k4 + 2], out[4 * k4 + 3]);
    }
}

// gInput[n, c, t] (+)= sum_k W1[k, c] * gVh[n, t, k]      and      gW1[k, c] += sum_{n,t} gVh[n, t, k] * V[n, c, t]
// gVh is channel-last [N,T,K] (KFIRST = false) or channel-first [N,K,T] (the output of the transposing peer reduce).
// A block owns tiles of TT texels: one texel per thread for gInput (the tile of gVh and V is parked in shared
// memory on the way), then one (k, c) pair per thread for the tile's share of gW1, four texels per 128-bit
// shared-memory load; the pairs' sums stay in registers over all tiles of the block and leave as one atomic each.
constexpr int POSTMIX_TT = 128;              // texels per block tile
constexpr int POSTMIX_TS = POSTMIX_TT + 4;   // row stride in shared memory (floats): rows 4 banks apart
template <int K, bool KFIRST>
__global__ void __launch_bounds__(POSTMIX_TT) cs_head_postmix_kernel(const float* __restrict__ gVh,
                                                                     const float* __restrict__ V,
                                                                     const float* __restrict__ W1,
                                                                     float* __restrict__ gInput, int accumulate,
                                                                     float* __restrict__ gW1, int C, long long T,
                                                                     long long ntiles_per_cell, long long ntiles) {
    extern __shared__ float4 sm4[];
    float* sm = reinterpret_cast<float*>(sm4);
    float* gs = sm;                                   // [K][TS]
    float* vs = gs + K * POSTMIX_TS;                  // [C][TS]
    float* w1s = vs + C * POSTMIX_TS;                 // [C][K]
    const int tid = threadIdx.x;
    for (int e = tid; e < K * C; e += POSTMIX_TT) w1s[(e % C) * K + (e / C)] = __ldg(W1 + e);
    const int npairs = K * C;
    constexpr int MAXPP = (MIX_MAXK * 64 + POSTMIX_TT - 1) / POSTMIX_TT;     // pairs per thread, C <= 64
    float wacc[MAXPP];
#pragma unroll
    for (int i = 0; i < MAXPP; ++i) wacc[i] = 0.f;
    __syncthreads();
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long n = tile / ntiles_per_cell;
        const long long t0 = (tile - n * ntiles_per_cell) * POSTMIX_TT;
        const long long t = t0 + tid;
        const bool ok = t < T;
        float g[K];
        if (ok) {
            if (KFIRST) {
#pragma unroll
                for (int k = 0; k < K; ++k) g[k] = __ldg(gVh + (n * K + k) * T + t);
            } else {
                const float4* src = reinterpret_cast<const float4*>(gVh + (n * T + t) * K);
#pragma unroll
                for (int k4 = 0; k4 < K / 4; ++k4) {
                    const float4 v = __ldg(src + k4);
                    g[4 * k4] = v.x; g[4 * k4 + 1] = v.y; g[4 * k4 + 2] = v.z; g[4 * k4 + 3] = v.w;
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < K; ++k) g[k] = 0.f;
        }
#pragma unroll
        for (int k = 0; k < K; ++k) gs[k * POSTMIX_TS + tid] = g[k];
        const float* vp = V + n * C * T + t;
        float* op = gInput ? gInput + n * C * T + t : nullptr;
#pragma unroll 4
        for (int c = 0; c < C; ++c) {
            const float v = ok ? __ldg(vp + (long long)c * T) : 0.f;
            vs[c * POSTMIX_TS + tid] = v;
            if (ok && op) {
                float gi = 0.f;
                const float* wr = w1s + c * K;
#pragma unroll
                for (int k = 0; k < K; ++k) gi = fmaf(wr[k], g[k], gi);
                if (accumulate) op[(long long)c * T] += gi; else op[(long long)c * T] = gi;
            }
        }
        __syncthreads();
        if (gW1) {
#pragma unroll
            for (int i = 0; i < MAXPP; ++i) {
                const int pr = tid + i * POSTMIX_TT;
                if (pr < npairs) {
                    const float4* gr = reinterpret_cast<const float4*>(gs + (pr % K) * POSTMIX_TS);
                    const float4* vr = reinterpret_cast<const float4*>(vs + (pr / K) * POSTMIX_TS);
                    float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
                    for (int t4 = 0; t4 < POSTMIX_TT / 4; ++t4) {
                        const float4 a = gr[t4], b = vr[t4];
                        s0 = fmaf(a.x, b.x, s0); s1 = fmaf(a.y, b.y, s1);
                        s0 = fmaf(a.z, b.z, s0); s1 = fmaf(a.w, b.w, s1);
                    }
                    wacc[i] += s0 + s1;
                }
            }
        }
        __syncthreads();
    }
    if (gW1) {
#pragma unroll
        for (int i = 0; i < MAXPP; ++i) {
            const int pr = tid + i * POSTMIX_TT;
            if (pr < npairs) atomicAdd(gW1 + (pr % K) * C + (pr / K), wacc[i]);
        }
    }
}

// The same pass for channel-last gVh, software-pipelined: the tile after this one streams into the other half of
// shared memory with cp.async while this one is consumed (the kernel above waits for its loads tile by tile:
// 31 % occupancy, 11 long-scoreboard stalls per issue, 0.19 ms for the 64 MiB fields of config 4).
template <int K>
__global__ void __launch_bounds__(POSTMIX_TT) cs_head_postmix_pipe_kernel(const float* __restrict__ gVh,
                                                                          const float* __restrict__ V,
                                                                          const float* __restrict__ W1,
                                                                          float* __restrict__ gInput, int accumulate,
                                                                          float* __restrict__ gW1, int C, long long T,
                                                                          long long ntiles_per_cell, long long ntiles) {
    constexpr int TT = POSTMIX_TT;
    extern __shared__ float4 sm4[];
    float* sm = reinterpret_cast<float*>(sm4);
    constexpr int VS = TT + 4;                        // row stride of v: rows 4 banks apart
    const int tile_f = TT * K + C * VS;               // floats per stage: g [TT][K] then v [C][VS]
    float* w1s = sm + 2 * tile_f;                     // [C][K]
    const int tid = threadIdx.x;
    for (int e = tid; e < K * C; e += TT) w1s[(e % C) * K + (e / C)] = __ldg(W1 + e);
    const int npairs4 = (K / 4) * C;                  // a work item = 4 hidden units x 1 channel x half of the texels
    constexpr int MAXW = (MIX_MAXK / 4 * 64 * 2 + TT - 1) / TT;
    float4 wacc[MAXW];
#pragma unroll
    for (int i = 0; i < MAXW; ++i) wacc[i] = make_float4(0.f, 0.f, 0.f, 0.f);

    auto issue = [&](long long tile, int st) {
        const long long n = tile / ntiles_per_cell;
        const long long t0 = (tile - n * ntiles_per_cell) * TT;
        float* gs = sm + st * tile_f;
        float* vs = gs + TT * K;
        const unsigned gsb = (unsigned)__cvta_generic_to_shared(gs);
        const unsigned vsb = (unsigned)__cvta_generic_to_shared(vs);
        // g: TT texels x K floats, contiguous in global memory (16-byte pieces; zero-filled past T)
        const float* gsrc = gVh + (n * T + t0) * K;
        for (int e = tid; e < TT * K / 4; e += TT) {
            const bool ok = t0 + (4 * e) / K < T;
            cp16z(gsb + e * 16, ok ? gsrc + 4 * e : gVh, ok);
        }
        // v: C rows of TT texels
        const bool vec = (T % 4 == 0);
        for (int e = tid; e < C * TT / 4; e += TT) {
            const int c = e / (TT / 4), q4 = e - c * (TT / 4);
            const long long t = t0 + 4 * q4;
            const float* src = V + (n * C + c) * T + t;
            const unsigned dst = vsb + (c * VS + 4 * q4) * 4;
            if (vec) {
                const bool ok = t < T;
                cp16z(dst, ok ? src : V, ok);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) cp4z(dst + 4 * i, (t + i < T) ? src + i : V, t + i < T);
            }
        }
        cp_async_commit();
    };

    __syncthreads();
    long long tile = blockIdx.x;
    int st = 0;
    if (tile < ntiles) issue(tile, 0);
    for (; tile < ntiles; tile += gridDim.x, st ^= 1) {
        const long long next = tile + gridDim.x;
        if (next < ntiles) { issue(next, st ^ 1); cp_async_wait<1>(); } else cp_async_wait<0>();
        __syncthreads();
        const float* gs = sm + st * tile_f;
        const float* vs = gs + TT * K;
        const long long n = tile / ntiles_per_cell;
        const long long t = (tile - n * ntiles_per_cell) * TT + tid;
        if (gInput && t < T) {
            float g[K];
#pragma unroll
            for (int k4 = 0; k4 < K / 4; ++k4) {
                const float4 v = reinterpret_cast<const float4*>(gs + tid * K)[k4];
                g[4 * k4] = v.x; g[4 * k4 + 1] = v.y; g[4 * k4 + 2] = v.z; g[4 * k4 + 3] = v.w;
            }
            float* op = gInput + n * C * T + t;
#pragma unroll 4
            for (int c = 0; c < C; ++c) {
                float gi = 0.f;
                const float* wr = w1s + c * K;
#pragma unroll
                for (int k = 0; k < K; ++k) gi = fmaf(wr[k], g[k], gi);
                if (accumulate) op[(long long)c * T] += gi; else op[(long long)c * T] = gi;
            }
        }
        if (gW1) {
#pragma unroll
            for (int i = 0; i < MAXW; ++i) {
                const int w = tid + i * TT;
                if (w < 2 * npairs4) {
                    const int half = w / npairs4, pr = w - half * npairs4;
                    const int kq = pr % (K / 4), c = pr / (K / 4);
                    const float* gp = gs + (half * (TT / 2)) * K + 4 * kq;
                    const float* vp = vs + c * VS + half * (TT / 2);
                    float4 a = wacc[i];
#pragma unroll 8
                    for (int tt = 0; tt < TT / 2; ++tt) {
                        const float4 gq = *reinterpret_cast<const float4*>(gp + tt * K);
                        const float v = vp[tt];
                        a.x = fmaf(gq.x, v, a.x); a.y = fmaf(gq.y, v, a.y); a.z = fmaf(gq.z, v, a.z); a.w = fmaf(gq.w, v, a.w);
                    }
                    wacc[i] = a;
                }
            }
        }
        __syncthreads();                                // the stage is free for the tile after next
    }
    if (gW1) {
#pragma unroll
        for (int i = 0; i < MAXW; ++i) {
            const int w = tid + i * TT;
            if (w < 2 * npairs4) {
                const int pr = w % npairs4;
                const int kq = pr % (K / 4), c = pr / (K / 4);
                atomicAdd(gW1 + (4 * kq + 0) * C + c, wacc[i].x);
                atomicAdd(gW1 + (4 * kq + 1) * C + c, wacc[i].y);
                atomicAdd(gW1 + (4 * kq + 2) * C + c, wacc[i].z);
                atomicAdd(gW1 + (4 * kq + 3) * C + c, wacc[i].w);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// The fused step
// ---------------------------------------------------------------------------------------------------------
struct FusedParams {
    int N;                    // cells
    int size[3];
    int tstride[3];
    long long T;              // texels per cell
    long long P;
    const float* Vh;          // [N*T + 1, K]  W1-mixed cells (cs_head_premix); the extra texel is zero
    float* gVh;               // [N*T + 1, K]  accumulated; the extra texel absorbs out-of-bounds corners
    const float* coords;      // [P, DIM]   (binned: cs_bin_points)
    const float* offset;      // [N]
    const float* b1;          // [K]
    const float* w2;          // [K]
    const float* b2;          // [1]
    float* gb1;               // [K]   +=
    float* gw2;               // [K]   +=
    float* gb2;               // [1]   +=
    float* loss_sum;          // [1]   += sum_p f^2 (unscaled)
    float c_u, c_u3, c1[3], c2[3];
    float scale;
    int cvec2;
    int pad, align, kernel, multicell, index_mode;
    int aggregate;            // pre-reduce the scatter over runs of points that share their corners
    long long num_ptiles;     // warp tiles of PTS consecutive points
    long long tiles_per_warp; // tiles per warp (the order in which the warps take them: see the kernel)
    int group_blocks;         // blocks per SM when the grid is one full wave (else 1): they interleave their tiles
};

// Phase 1 for one (cell, point).  Record fields (float4 each, [field][point]):
//   0 .. CQ-1 : texel index of each of the 2^DIM corners relative to this cell's first texel (int bits, corner c
//               = bit a set -> high corner along axis a); a corner outside the cell points at the extra texel
//               behind the last cell (zero in Vh, a dump in gVh): gathers and reds need no predicate
//   CQ + a    : (w1, m k', -m^2 k'', w0) of axis a: weight of the high corner, d weight / d coordinate of the
//               high corner, d2 weight / d coordinate^2 of the high corner (opposite signs for the low corner)
// Record slot of point i of a tile: one spare float4 per 8 points.  The L lanes of a walker read the same record and
// the walkers of a warp read points PPQ apart: with slot = i their 16-byte records fall on 2 of the 8 bank groups
// (4 wavefronts per LDS.128 at K = 16; ncu: 40 % of the kernel's shared-memory wavefronts were bank conflicts), with
// the spare slots they fall on 8 different ones (1 wavefront), for every hidden width; writes stay conflict-free.
__device__ __forceinline__ int rec_slot(int i) { return i + (i >> 3); }

template <int DIM, int PTS>
__device__ __forceinline__ void build_fused_record(float4* rec4, int i, const float (&g)[DIM], bool in_range,
                                                   float off, int pad_index, const FusedParams& p) {
    constexpr int NCORN = 1 << DIM;
    constexpr int CQ = NCORN / 4;
    int idx[NCORN];
    float4 ax[DIM];
#pragma unroll
    for (int a = 0; a < DIM; ++a) ax[a] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int c = 0; c < NCORN; ++c) idx[c] = pad_index;
    if (in_range) {
        bool ok = true;
        bool lo_ok[DIM], hi_ok[DIM];
        int base = 0;
#pragma unroll
        for (int a = 0; a < DIM; ++a) {
            const AxisRec ar = axis_setup(g[a], p.size[a], off, p, p.align != 0, 2);
            ok = ok && ar.ok;
            base += ar.l * p.tstride[a];
            lo_ok[a] = (ar.l >= 0) && (ar.l < p.size[a]);
            hi_ok[a] = (ar.l + 1 >= 0) && (ar.l + 1 < p.size[a]);
            ax[a] = make_float4(ar.w1, ar.d, -ar.e, ar.w0);
        }
        if (ok) {
#pragma unroll
            for (int c = 0; c < NCORN; ++c) {
                bool valid = true;
                int o = base;
#pragma unroll
                for (int a = 0; a < DIM; ++a) {
                    const bool hi = (c >> a) & 1;
                    valid = valid && (hi ? hi_ok[a] : lo_ok[a]);
                    o += hi ? p.tstride[a] : 0;
                }
                if (valid) idx[c] = o;
            }
        }
    }
    const int sl = rec_slot(i);                  // PTS = padded stride of a field
#pragma unroll
    for (int h = 0; h < CQ; ++h)
        rec4[h * PTS + sl] = make_float4(__int_as_float(idx[4 * h]), __int_as_float(idx[4 * h + 1]),
                                         __int_as_float(idx[4 * h + 2]), __int_as_float(idx[4 * h + 3]));
#pragma unroll
    for (int a = 0; a < DIM; ++a) rec4[(CQ + a) * PTS + sl] = ax[a];
}

__device__ __forceinline__ float tanh_ex2(float x) {
    // one ex2.approx + one rcp.approx: absolute error a few 1e-7 (|tanh| <= 1 is the scale that matters:
    // u = w2 . tanh, s1 = 1 - tanh^2); saturates cleanly to +-1 for large |x|
    // (explicit .ftz forms: __expf / __fdividef without -use_fast_math wrap the MUFU in denormal and range
    // handling, 37 instructions per tanh instead of 7 -- 16 % of the kernel's instructions, ncu source page)
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-2.885390081777927f * fabsf(x)));       // exp(-2|x|)
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
    return copysignf((1.f - e) * r, x);
}

__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// Packed fp32 pairs (fma.rn.f32x2 -> FFMA2, sm_100+): two IEEE fp32 operations per issued instruction.  The FMA
// pipe delivers the same 128 lanes per clock and SM either way (tools/microbench_ffma2.cu: 120 vs 122), but this
// kernel is bound by instruction ISSUE (issue slots 66-71 % busy, FMA pipe 43 %), and the 4 hidden units a lane owns
// are two natural pairs: every per-unit operation below is issued once per pair.  A scalar operand (a weight of the
// point) is the broadcast form of the instruction (pk(s, s) folds into Rx.F32), a negated operand folds into the
// operand modifier: neither costs an instruction.  Results are bit-identical to the scalar formulation.
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float a, float b) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ f2 bc(float s) { return pk(s, s); }
__device__ __forceinline__ float lo(f2 v) { return __uint_as_float((unsigned)(v & 0xffffffffull)); }
__device__ __forceinline__ float hi(f2 v) { return __uint_as_float((unsigned)(v >> 32)); }
__device__ __forceinline__ f2 neg2(f2 v) { return pk(-lo(v), -hi(v)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { f2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 f4lo(const float4& v) { return pk(v.x, v.y); }
__device__ __forceinline__ f2 f4hi(const float4& v) { return pk(v.z, v.w); }

// Two channels of one (x, y) slab: bilinear-type blend with the high-corner weights w1x, w1y (w0 = 1 - w1 for
// all three kernels).  A = value, T = d/dx / (m k'x) = d2/dx2 / (-m^2 k''x), DA = d/dy / (m k'y).
struct SlabOut { f2 A, T, DA; };
__device__ __forceinline__ SlabOut slab_blend(f2 v00, f2 v10, f2 v01, f2 v11, f2 w1x, f2 w1y) {
    const f2 d0 = sub2(v10, v00), d1 = sub2(v11, v01);
    const f2 a0 = fma2(d0, w1x, v00), a1 = fma2(d1, w1x, v01);
    SlabOut o;
    o.DA = sub2(a1, a0);
    o.A = fma2(o.DA, w1y, a0);
    o.T = fma2(sub2(d1, d0), w1y, d0);
    return o;
}

// Per-lane head parameters (hidden units 4j .. 4j+3) and the gradient / loss accumulators of this lane
struct HeadLane {
    float b1k[4];
    f2 w2p[2];
    float b2;
    f2 gb1acc[2], gw2acc[2];
    float gb2acc, lossacc;
};

__device__ __forceinline__ void init_head_lane(HeadLane& hl, const FusedParams& p, int j) {
#pragma unroll
    for (int k = 0; k < 4; ++k) hl.b1k[k] = __ldg(p.b1 + 4 * j + k);
    hl.w2p[0] = pk(__ldg(p.w2 + 4 * j), __ldg(p.w2 + 4 * j + 1));
    hl.w2p[1] = pk(__ldg(p.w2 + 4 * j + 2), __ldg(p.w2 + 4 * j + 3));
    hl.b2 = __ldg(p.b2);
    hl.gb1acc[0] = hl.gb1acc[1] = hl.gw2acc[0] = hl.gw2acc[1] = pk(0.f, 0.f);
    hl.gb2acc = 0.f; hl.lossacc = 0.f;
}

template <int DIM, int LSHIFT>
struct FusedGeom {
    static constexpr int NCORN = 1 << DIM;
    static constexpr int CQ = NCORN / 4;
    static constexpr int J = 1 + 2 * DIM;
    static constexpr int L = 1 << LSHIFT;
    static constexpr int K = 4 * L;                         // hidden width: 4 units per lane, L lanes per point
    static constexpr int NW = 32 >> LSHIFT;                 // walkers (point slots) per warp
    static constexpr int PPQ = (DIM == 2) ? 4 : 2;          // consecutive points per walker and tile
    static constexpr int PTS = PPQ * NW;                    // points per warp tile
    static constexpr int PPL = (PTS + 31) / 32;
    static constexpr int PTSP = PTS + (PTS + 7) / 8;        // padded points per record field (rec_slot)
    static constexpr int REC1 = (CQ + DIM) * PTSP;          // float4 per record buffer (one cell)
};

// Phase A for one point (record slot ri): h[jt][hh] = H_jt, hidden units 4j + 2hh, 4j + 2hh + 1, summed over the cells
template <int DIM, int LSHIFT, int NC>
__device__ __forceinline__ void fused_gather_point(const FusedParams& p, const float4* recw, int ri, int ncells, int j,
                                                   f2 (&h)[1 + 2 * DIM][2]) {
    using G = FusedGeom<DIM, LSHIFT>;
    constexpr int NCORN = G::NCORN, CQ = G::CQ, J = G::J, K = G::K, PTSP = G::PTSP, REC1 = G::REC1;
    const f2 zero2 = pk(0.f, 0.f);
#pragma unroll
    for (int jt = 0; jt < J; ++jt) { h[jt][0] = zero2; h[jt][1] = zero2; }
#pragma unroll (NC > 0 ? (FUSED_GATHER_UNROLL > 0 ? FUSED_GATHER_UNROLL : NC) : 2)
    for (int n = 0; n < ncells; ++n) {
        const float4* rec = recw + n * REC1;
        const float* vsrc = p.Vh + (long long)n * p.T * K + 4 * j;
        float4 v[NCORN];
#pragma unroll
        for (int hq = 0; hq < CQ; ++hq) {
            const float4 ix = rec[hq * PTSP + ri];
            v[4 * hq + 0] = ldg_f4(vsrc + (long long)__float_as_int(ix.x) * K);
            v[4 * hq + 1] = ldg_f4(vsrc + (long long)__float_as_int(ix.y) * K);
            v[4 * hq + 2] = ldg_f4(vsrc + (long long)__float_as_int(ix.z) * K);
            v[4 * hq + 3] = ldg_f4(vsrc + (long long)__float_as_int(ix.w) * K);
        }
        const float4 ax = rec[CQ * PTSP + ri];            // (w1, d, -e, w0)
        const float4 ay = rec[(CQ + 1) * PTSP + ri];
        const f2 w1x = bc(ax.x), w1y = bc(ay.x);
        if (DIM == 2) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const SlabOut o = hh == 0 ? slab_blend(f4lo(v[0]), f4lo(v[1]), f4lo(v[2]), f4lo(v[3]), w1x, w1y)
                                          : slab_blend(f4hi(v[0]), f4hi(v[1]), f4hi(v[2]), f4hi(v[3]), w1x, w1y);
                h[0][hh] = add2(h[0][hh], o.A);
                h[1][hh] = fma2(bc(ax.y), o.T, h[1][hh]);
                h[2][hh] = fma2(bc(ay.y), o.DA, h[2][hh]);
                h[3][hh] = fma2(bc(ax.z), o.T, h[3][hh]);
                h[4][hh] = fma2(bc(ay.z), o.DA, h[4][hh]);
            }
        } else {
            const float4 az = rec[(CQ + (DIM == 3 ? 2 : 1)) * PTSP + ri];
            const f2 w1z = bc(az.x);
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const SlabOut lo_ = hh == 0 ? slab_blend(f4lo(v[0]), f4lo(v[1]), f4lo(v[2]), f4lo(v[3]), w1x, w1y)
                                            : slab_blend(f4hi(v[0]), f4hi(v[1]), f4hi(v[2]), f4hi(v[3]), w1x, w1y);
                const SlabOut hi_ = hh == 0 ? slab_blend(f4lo(v[4 % NCORN]), f4lo(v[5 % NCORN]), f4lo(v[6 % NCORN]), f4lo(v[7 % NCORN]), w1x, w1y)
                                            : slab_blend(f4hi(v[4 % NCORN]), f4hi(v[5 % NCORN]), f4hi(v[6 % NCORN]), f4hi(v[7 % NCORN]), w1x, w1y);
                const f2 dA = sub2(hi_.A, lo_.A);
                const f2 tz = fma2(sub2(hi_.T, lo_.T), w1z, lo_.T);
                const f2 daz = fma2(sub2(hi_.DA, lo_.DA), w1z, lo_.DA);
                h[0][hh] = add2(h[0][hh], fma2(dA, w1z, lo_.A));
                h[1][hh] = fma2(bc(ax.y), tz, h[1][hh]);
                h[2][hh] = fma2(bc(ay.y), daz, h[2][hh]);
                h[DIM][hh] = fma2(bc(az.y), dA, h[DIM][hh]);
                h[1 + DIM][hh] = fma2(bc(ax.z), tz, h[1 + DIM][hh]);
                h[(2 + DIM) % J][hh] = fma2(bc(ay.z), daz, h[(2 + DIM) % J][hh]);
                h[2 * DIM][hh] = fma2(bc(az.z), dA, h[2 * DIM][hh]);
            }
        }
    }

}

// Phase B for one point: head + residual + loss, and d loss / d H_jt in place of h
template <int DIM, int LSHIFT>
__device__ __forceinline__ void fused_head_point(const FusedParams& p, HeadLane& hl, bool valid, int j,
                                                 f2 (&h)[1 + 2 * DIM][2]) {
    constexpr int L = 1 << LSHIFT;
    // B: head + residual + loss, and d loss / d H_jt in place (test_2d.py:42-127 in closed form):
    //   t = tanh h, s1 = 1 - t^2, s2 = -2 t s1, s3 = -2 (s1^2 + t s2)
    //   u = w2.t + b2, u_a = w2.(s1 hd_a), u_aa = w2.(s2 hd_a^2 + s1 hdd_a)
    f2 th[2], s1[2], ws1[2], ws2[2];
    const f2 zero2 = pk(0.f, 0.f);
    f2 ppu = zero2, ppua[DIM], ppuaa[DIM];
#pragma unroll
    for (int a = 0; a < DIM; ++a) { ppua[a] = zero2; ppuaa[a] = zero2; }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
        th[hh] = pk(tanh_ex2(lo(h[0][hh]) + hl.b1k[2 * hh]), tanh_ex2(hi(h[0][hh]) + hl.b1k[2 * hh + 1]));
        s1[hh] = fma2(neg2(th[hh]), th[hh], bc(1.f));
        ws1[hh] = mul2(hl.w2p[hh], s1[hh]);                                 // w2 s1
        ws2[hh] = mul2(mul2(bc(-2.f), th[hh]), ws1[hh]);                 // w2 s2
        ppu = fma2(hl.w2p[hh], th[hh], ppu);
#pragma unroll
        for (int a = 0; a < DIM; ++a) {
            const f2 hd = h[1 + a][hh];
            ppua[a] = fma2(ws1[hh], hd, ppua[a]);
            ppuaa[a] = fma2(ws2[hh], mul2(hd, hd), fma2(ws1[hh], h[1 + DIM + a][hh], ppuaa[a]));
        }
    }
    float pu = lo(ppu) + hi(ppu), pua[DIM], puaa[DIM];
#pragma unroll
    for (int a = 0; a < DIM; ++a) { pua[a] = lo(ppua[a]) + hi(ppua[a]); puaa[a] = lo(ppuaa[a]) + hi(ppuaa[a]); }
#pragma unroll
    for (int o = 1; o < L; o <<= 1) {
        pu += __shfl_xor_sync(0xffffffffu, pu, o);
#pragma unroll
        for (int a = 0; a < DIM; ++a) {
            pua[a] += __shfl_xor_sync(0xffffffffu, pua[a], o);
            puaa[a] += __shfl_xor_sync(0xffffffffu, puaa[a], o);
        }
    }
    const float u = pu + hl.b2;
    float f = p.c_u * u + p.c_u3 * u * u * u;
#pragma unroll
    for (int a = 0; a < DIM; ++a) f += p.c1[a] * pua[a] + p.c2[a] * puaa[a];
    const float gg = valid ? 2.f * p.scale * f : 0.f;
    const float gsc = gg * (p.c_u + 3.f * p.c_u3 * u * u);
    if (j == 0) {
        if (valid) hl.lossacc = fmaf(f, f, hl.lossacc);
        hl.gb2acc += gsc;
    }
    float g1c[DIM], g2c[DIM];
#pragma unroll
    for (int a = 0; a < DIM; ++a) { g1c[a] = gg * p.c1[a]; g2c[a] = gg * p.c2[a]; }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
        // with G1 = sum_a g1c_a hd_a + g2c_a hdd_a and G2 = sum_a g2c_a hd_a^2:
        //   d loss / d w2_k = gsc t + s1 G1 + s2 G2        d loss / d h_k = w2 (gsc s1 + s2 G1 + s3 G2)
        const f2 ws3 = mul2(bc(-2.f), fma2(ws1[hh], s1[hh], mul2(th[hh], ws2[hh])));      // w2 s3
        f2 G1 = zero2, G2 = zero2;
#pragma unroll
        for (int a = 0; a < DIM; ++a) {
            const f2 hd = h[1 + a][hh];
            const f2 gh2 = mul2(bc(g2c[a]), hd);
            G1 = fma2(bc(g1c[a]), hd, fma2(bc(g2c[a]), h[1 + DIM + a][hh], G1));
            G2 = fma2(gh2, hd, G2);
            h[1 + a][hh] = fma2(ws1[hh], bc(g1c[a]), mul2(mul2(bc(2.f), ws2[hh]), gh2));
            h[1 + DIM + a][hh] = mul2(ws1[hh], bc(g2c[a]));
        }
        const f2 gh = fma2(ws3, G2, fma2(ws2[hh], G1, mul2(ws1[hh], bc(gsc))));
        const f2 gw2 = fma2(mul2(mul2(bc(-2.f), th[hh]), s1[hh]), G2, fma2(s1[hh], G1, mul2(bc(gsc), th[hh])));
        h[0][hh] = gh;
        hl.gb1acc[hh] = add2(hl.gb1acc[hh], gh);
        hl.gw2acc[hh] = add2(hl.gw2acc[hh], gw2);
    }
}

// Head-parameter gradients and loss: walkers of a warp (shuffles) -> warps (shared memory) -> one atomic per block
// and element.  smem4 is reused: every warp must be done with its records (the barrier inside).
template <int LSHIFT>
__device__ __forceinline__ void fused_reduce_head(const FusedParams& p, HeadLane& hl, float4* smem4, int lane, int warp,
                                                  int wpb) {
    constexpr int L = 1 << LSHIFT;
    constexpr int K = 4 * L;
    const int q = lane >> LSHIFT;
    const int j = lane & (L - 1);
    // ---- head-parameter gradients and loss: walkers of a warp (shuffles) -> warps (shared memory) -> one
    // atomic per block and element
    float gb1s[4] = {lo(hl.gb1acc[0]), hi(hl.gb1acc[0]), lo(hl.gb1acc[1]), hi(hl.gb1acc[1])};
    float gw2s[4] = {lo(hl.gw2acc[0]), hi(hl.gw2acc[0]), lo(hl.gw2acc[1]), hi(hl.gw2acc[1])};
#pragma unroll
    for (int o = L; o < 32; o <<= 1) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            gb1s[k] += __shfl_xor_sync(0xffffffffu, gb1s[k], o);
            gw2s[k] += __shfl_xor_sync(0xffffffffu, gw2s[k], o);
        }
        hl.gb2acc += __shfl_xor_sync(0xffffffffu, hl.gb2acc, o);
        hl.lossacc += __shfl_xor_sync(0xffffffffu, hl.lossacc, o);
    }
    __syncthreads();                                 // every warp is done with its records
    float* red = reinterpret_cast<float*>(smem4);    // [wpb][2K + 2]
    constexpr int RW = 2 * K + 2;
    if (q == 0) {
        float* rw = red + warp * RW;
#pragma unroll
        for (int k = 0; k < 4; ++k) { rw[4 * j + k] = gb1s[k]; rw[K + 4 * j + k] = gw2s[k]; }
        if (j == 0) { rw[2 * K] = hl.gb2acc; rw[2 * K + 1] = hl.lossacc; }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < RW; e += blockDim.x) {
        float s = 0.f;
        for (int w = 0; w < wpb; ++w) s += red[w * RW + e];
        float* dst = (e < K) ? p.gb1 + e : (e < 2 * K) ? p.gw2 + (e - K) : (e == 2 * K) ? p.gb2 : p.loss_sum;
        atomicAdd(dst, s);
    }
}

#ifndef CS_FUSED_BLOCKS
#define CS_FUSED_BLOCKS 1
#endif
// Unrolling over the cells pays for the gather only (NC = 4: 6.68 -> 6.14 ms per 2^25 points, its loads are in
// flight together).  Unrolling phase 1 and the scatter by 2 / 4 cells grows the code and LOSES: 6.34 / 7.30 ms
// (profiles/README.md): the kernel lives at the edge of the instruction cache.
#ifndef CS_FUSED_INTERLEAVE
#define CS_FUSED_INTERLEAVE -1        // -1: the default order (see the kernel)
#endif
#ifndef CS_FUSED_UNROLL_P1
#define CS_FUSED_UNROLL_P1 1
#endif
#ifndef CS_FUSED_UNROLL_C
#define CS_FUSED_UNROLL_C 1
#endif
// ONE block of 12 warps per SM (168 registers): the warps of a block are on one SM by construction, which is what the
// interleaved tile order needs; 3 blocks of 4 warps relied on the hardware's block placement and were 4 % slower
#ifndef CS_FUSED_THREADS
#define CS_FUSED_THREADS 384
#endif
constexpr int FUSED_THREADS = CS_FUSED_THREADS;
constexpr int FUSED_UNROLL_P1 = CS_FUSED_UNROLL_P1;
constexpr int FUSED_UNROLL_C = CS_FUSED_UNROLL_C;
constexpr int FUSED_MAX_CELLS = 32;      // the records of all cells of a tile live in shared memory

// Code size matters here: the first version (scalar arithmetic, three inlined phases unrolled over the points of a
// walker) was 8000 SASS instructions (128 KB) and stalled on instruction fetch.  Phase 1 has one call site, corners
// need no validity predicates (dump texel), the arithmetic is packed (two hidden units per instruction): 4096
// instructions with gather + head unrolled over the PPQ points, `no_instruction` stalls 0.15 per issue.
// NC: number of cells when it is known at compile time (4: the PIXEL configurations; the gather loop over the
// cells is then fully unrolled and all its loads are in flight together), 0 = run-time p.N.
template <int DIM, int LSHIFT, int NC>
__global__ void __launch_bounds__(FUSED_THREADS, CS_FUSED_BLOCKS)
cs_pde_fused_kernel(const FusedParams p) {
    constexpr int NCORN = 1 << DIM;
    constexpr int CQ = NCORN / 4;
    constexpr int J = 1 + 2 * DIM;
    constexpr int L = 1 << LSHIFT;
    constexpr int K = 4 * L;                         // hidden width: 4 units per lane, L lanes per point
    constexpr int NW = 32 >> LSHIFT;                 // walkers (point slots) per warp
    constexpr int PPQ = (DIM == 2) ? 4 : 2;          // consecutive points per walker and tile
    constexpr int PTS = PPQ * NW;                    // points per warp tile
    constexpr int PPL = (PTS + 31) / 32;
    constexpr int PTSP = PTS + (PTS + 7) / 8;        // padded points per record field (rec_slot)
    constexpr int REC1 = (CQ + DIM) * PTSP;          // float4 per record buffer (one cell)

    extern __shared__ float4 smem4[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int q = lane >> LSHIFT;
    const int j = lane & (L - 1);
    const int ncells = NC > 0 ? NC : p.N;
    float4* recw = smem4 + (size_t)warp * ncells * REC1;

    HeadLane hl;
    init_head_lane(hl, p, j);
    const f2 zero2 = pk(0.f, 0.f);

    // Tile order.  Binned points: neighbouring tiles touch the same texels, so the warps that share an L1 should work
    // on neighbouring tiles at the same time.  IL = 1 (default): a block owns a contiguous range and its warps take
    // the tiles round-robin.  IL = 2: the `group_blocks` blocks that the hardware places on one SM (blocks b,
    // b + groups, ... of a grid launched in one wave) share one range and interleave all their warps.  IL = 0: every
    // warp walks its own contiguous range (round 2's first version).  Measured per 2^25 (2D) / 2^22 (3D) points with
    // 3 blocks of 4 warps per SM: 5.72 / 1.57 ms (0), 5.66 / 1.53 (1), 5.54 / 1.57 (2); with one block of 12 warps and
    // IL = 1: 5.34 / 1.52 ms.
    constexpr int IL = CS_FUSED_INTERLEAVE >= 0 ? CS_FUSED_INTERLEAVE : 1;
    long long tile_step, tile_begin, tile_end;
    if (IL == 2) {
        const int groups = (int)gridDim.x / p.group_blocks;
        const int grp = (int)blockIdx.x % groups, member = (int)blockIdx.x / groups;
        tile_step = (long long)wpb * p.group_blocks;
        tile_begin = (long long)grp * tile_step * p.tiles_per_warp + member * wpb + warp;
        tile_end = (long long)(grp + 1) * tile_step * p.tiles_per_warp;
    } else if (IL == 1) {
        tile_step = wpb;
        tile_begin = (long long)blockIdx.x * wpb * p.tiles_per_warp + warp;
        tile_end = (long long)(blockIdx.x + 1) * wpb * p.tiles_per_warp;
    } else {
        tile_step = 1;
        tile_begin = ((long long)blockIdx.x * wpb + warp) * p.tiles_per_warp;
        tile_end = tile_begin + p.tiles_per_warp;
    }
    if (tile_end > p.num_ptiles) tile_end = p.num_ptiles;

    // phase 1 runs one (cell, point) per lane: lane -> point (u * 32 + lane) % PTS; when a tile has fewer than 32
    // points (3D: 16) the lanes beyond PTS take the same points for the NEXT cell (CPL cells per round)
    constexpr int CPL = (PTS < 32) ? 32 / PTS : 1;
    auto load_coords = [&](float (&g)[PPL][DIM], bool (&inr)[PPL], long long tile) {
#pragma unroll
        for (int u = 0; u < PPL; ++u) {
            const int i = (u * 32 + lane) % PTS;
            const long long pi = tile * PTS + i;
            inr[u] = (u * 32 + lane < PTS * CPL) && (pi < p.P);
#pragma unroll
            for (int a = 0; a < DIM; ++a) g[u][a] = 0.f;
            if (inr[u]) {
                const float* gp = p.coords + pi * DIM;
                if (DIM == 2 && p.cvec2) {
                    const float2 t = __ldg(reinterpret_cast<const float2*>(gp));
                    g[u][0] = t.x; g[u][1] = t.y;
                } else {
#pragma unroll
                    for (int a = 0; a < DIM; ++a) g[u][a] = __ldg(gp + a);
                }
            }
        }
    };

    float gcur[PPL][DIM], gnext[PPL][DIM];
    bool icur[PPL], inext[PPL];
    if (tile_begin < tile_end) load_coords(gcur, icur, tile_begin);

#pragma unroll 1
    for (long long tile = tile_begin; tile < tile_end; tile += tile_step) {
        const bool have_next = tile + tile_step < tile_end;
        if (have_next) load_coords(gnext, inext, tile + tile_step);
        const long long qp0 = tile * PTS + (long long)q * PPQ;      // first point of this walker

        // ---- phase 1: records of every cell for the PTS points of this tile, one point per lane
        __syncwarp();                                   // everyone is done reading the previous tile's records
#pragma unroll (NC > 0 ? FUSED_UNROLL_P1 : 1)
        for (int n0 = 0; n0 < ncells; n0 += CPL) {
            const int n = n0 + (CPL > 1 ? lane / PTS : 0);
            if (n < ncells) {
                const float off = __ldg(p.offset + n);
                const int pad_index = (int)((long long)(ncells - n) * p.T);
#pragma unroll
                for (int u = 0; u < PPL; ++u) {
                    const int i = (u * 32 + lane) % PTS;
                    if (u * 32 + lane < PTS * CPL)
                        build_fused_record<DIM, PTSP>(recw + n * REC1, i, gcur[u], icur[u], off, pad_index, p);
                }
            }
        }
        __syncwarp();                                   // records are visible

        // acc[jt][t][h] = d loss / d H_jt of point t of this walker, hidden units 4j + 2h, 4j + 2h + 1
        f2 acc[J][PPQ][2];

        // ---- A + B for the PPQ points of this walker.  The loop is unrolled: d loss / d H of point t lands in its
        // slot of the register tile directly.  (Round 2's first versions kept it a real loop, with the finished point
        // rotated into the tile by 80 moves, because the unrolled scalar code -- 8000 instructions -- stalled on
        // instruction fetch; with packed pairs the unrolled kernel is 4096 instructions and faster than the loop:
        // 5.84 -> 5.74 ms per 2^25 points, 3D 1.62 -> 1.58 ms per 2^22; unrolling the scatter over the cells or
        // phase 1 as well still loses, 5.92 / 6.05 ms, profiles/README.md.)
#pragma unroll
        for (int t = 0; t < PPQ; ++t) {
            const int ri = rec_slot(PPQ * q + t);
            f2 h[J][2];
            fused_gather_point<DIM, LSHIFT, NC>(p, recw, ri, ncells, j, h);
            fused_head_point<DIM, LSHIFT>(p, hl, qp0 + t < p.P, j, h);
#pragma unroll
            for (int jt = 0; jt < J; ++jt) { acc[jt][t][0] = h[jt][0]; acc[jt][t][1] = h[jt][1]; }
        }

        // ---- C: scatter d loss / d H_jt into gVh (separable adjoint: per corner  Wy u_x +- Wx beta), pre-reduced in
        // registers over runs of consecutive points of this walker with identical corners
#pragma unroll (NC > 0 ? FUSED_UNROLL_C : 1)
        for (int n = 0; n < ncells; ++n) {
            const float4* rec = recw + n * REC1;
            float* gcell = p.gVh + (long long)n * p.T * K + 4 * j;
            f2 cur[NCORN][2];
            int cix[NCORN];
#pragma unroll
            for (int t = 0; t < PPQ; ++t) {
                const int ri = rec_slot(PPQ * q + t);
                int ix[NCORN];
#pragma unroll
                for (int hq = 0; hq < CQ; ++hq) {
                    const float4 f = rec[hq * PTSP + ri];
                    ix[4 * hq] = __float_as_int(f.x); ix[4 * hq + 1] = __float_as_int(f.y);
                    ix[4 * hq + 2] = __float_as_int(f.z); ix[4 * hq + 3] = __float_as_int(f.w);
                }
                const float4 ax = rec[CQ * PTSP + ri];            // (w1, d, -e, w0)
                const float4 ay = rec[(CQ + 1) * PTSP + ri];
                f2 cv[NCORN][2];
                if (DIM == 2) {
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const f2 alpha = fma2(acc[3][t][hh], bc(ax.z), mul2(acc[1][t][hh], bc(ax.y)));
                        const f2 beta = fma2(acc[4][t][hh], bc(ay.z), mul2(acc[2][t][hh], bc(ay.y)));
                        const f2 u0 = fma2(acc[0][t][hh], bc(ax.w), neg2(alpha));
                        const f2 u1 = fma2(acc[0][t][hh], bc(ax.x), alpha);
                        const f2 bx0 = mul2(bc(ax.w), beta), bx1 = mul2(bc(ax.x), beta);
                        cv[0][hh] = fma2(bc(ay.w), u0, neg2(bx0));
                        cv[1][hh] = fma2(bc(ay.w), u1, neg2(bx1));
                        cv[2][hh] = fma2(bc(ay.x), u0, bx0);
                        cv[3][hh] = fma2(bc(ay.x), u1, bx1);
                    }
                } else {
                    const float4 az = rec[(CQ + (DIM == 3 ? 2 : 1)) * PTSP + ri];
                    const float p00 = ax.w * ay.w, p10 = ax.x * ay.w, p01 = ax.w * ay.x, p11 = ax.x * ay.x;
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const f2 alpha = fma2(acc[1 + DIM][t][hh], bc(ax.z), mul2(acc[1][t][hh], bc(ax.y)));
                        const f2 beta = fma2(acc[(2 + DIM) % J][t][hh], bc(ay.z), mul2(acc[2][t][hh], bc(ay.y)));
                        const f2 gamma = fma2(acc[2 * DIM][t][hh], bc(az.z), mul2(acc[DIM][t][hh], bc(az.y)));
                        const f2 u0 = fma2(acc[0][t][hh], bc(ax.w), neg2(alpha));
                        const f2 u1 = fma2(acc[0][t][hh], bc(ax.x), alpha);
                        const f2 bx0 = mul2(bc(ax.w), beta), bx1 = mul2(bc(ax.x), beta);
                        const f2 c00 = fma2(bc(ay.w), u0, neg2(bx0)), c10 = fma2(bc(ay.w), u1, neg2(bx1));
                        const f2 c01 = fma2(bc(ay.x), u0, bx0), c11 = fma2(bc(ay.x), u1, bx1);
                        const f2 g00 = mul2(bc(p00), gamma), g10 = mul2(bc(p10), gamma);
                        const f2 g01 = mul2(bc(p01), gamma), g11 = mul2(bc(p11), gamma);
                        cv[0][hh] = fma2(bc(az.w), c00, neg2(g00));
                        cv[1][hh] = fma2(bc(az.w), c10, neg2(g10));
                        cv[2][hh] = fma2(bc(az.w), c01, neg2(g01));
                        cv[3][hh] = fma2(bc(az.w), c11, neg2(g11));
                        cv[4 % NCORN][hh] = fma2(bc(az.x), c00, g00);
                        cv[5 % NCORN][hh] = fma2(bc(az.x), c10, g10);
                        cv[6 % NCORN][hh] = fma2(bc(az.x), c01, g01);
                        cv[7 % NCORN][hh] = fma2(bc(az.x), c11, g11);
                    }
                }
                if (t > 0) {
                    bool same = p.aggregate != 0;
#pragma unroll
                    for (int c = 0; c < NCORN; ++c) same = same && (ix[c] == cix[c]);
                    if (same) {
#pragma unroll
                        for (int c = 0; c < NCORN; ++c) {
                            cv[c][0] = add2(cv[c][0], cur[c][0]);
                            cv[c][1] = add2(cv[c][1], cur[c][1]);
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < NCORN; ++c)
                            red_add_v4(gcell + (long long)cix[c] * K, lo(cur[c][0]), hi(cur[c][0]), lo(cur[c][1]), hi(cur[c][1]));
                    }
                }
#pragma unroll
                for (int c = 0; c < NCORN; ++c) {
                    cix[c] = ix[c];
                    cur[c][0] = cv[c][0];
                    cur[c][1] = cv[c][1];
                }
            }
#pragma unroll
            for (int c = 0; c < NCORN; ++c)
                red_add_v4(gcell + (long long)cix[c] * K, lo(cur[c][0]), hi(cur[c][0]), lo(cur[c][1]), hi(cur[c][1]));
        }

        if (have_next) {
#pragma unroll
            for (int u = 0; u < PPL; ++u) {
                icur[u] = inext[u];
#pragma unroll
                for (int a = 0; a < DIM; ++a) gcur[u][a] = gnext[u][a];
            }
        }
    }

    fused_reduce_head<LSHIFT>(p, hl, smem4, lane, warp, wpb);
}

// ---------------------------------------------------------------------------------------------------------
// launch
// ---------------------------------------------------------------------------------------------------------
template <int DIM, int LSHIFT>
cudaError_t launch_fused_one(FusedParams& p, cudaStream_t stream) {
    constexpr int NW = 32 >> LSHIFT;
    constexpr int K = 4 << LSHIFT;
    constexpr int PPQ = (DIM == 2) ? 4 : 2;
    constexpr int PTS = PPQ * NW;
    constexpr int REC1 = ((1 << DIM) / 4 + DIM) * (PTS + (PTS + 7) / 8);
    if (p.N > FUSED_MAX_CELLS) return cudaErrorInvalidConfiguration;
    auto kern = (p.N == 4) ? cs_pde_fused_kernel<DIM, LSHIFT, 4> : cs_pde_fused_kernel<DIM, LSHIFT, 0>;
    const size_t per_warp = (size_t)p.N * REC1 * sizeof(float4);
    int wpb = FUSED_THREADS / 32;
    while (wpb > 1 && wpb * per_warp > (216 / CS_FUSED_BLOCKS) * 1024) wpb >>= 1;   // the records of a block must fit
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    p.num_ptiles = (p.P + PTS - 1) / PTS;
    // few points: smaller blocks, so that the tiles still spread over all SMs
    if (p.num_ptiles < (long long)sms * wpb) wpb = (int)((p.num_ptiles + sms - 1) / sms);
    if (wpb < 1) wpb = 1;
    const int threads = wpb * 32;
    size_t smem = wpb * per_warp;
    const size_t red_bytes = (size_t)wpb * (2 * K + 2) * sizeof(float);
    if (smem < red_bytes) smem = red_bytes;
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) occ = 1;
    long long blocks = (long long)sms * occ;
    const long long need = (p.num_ptiles + wpb - 1) / wpb;
    if (blocks > need) blocks = need;
    if (blocks < 1) return cudaSuccess;
    p.tiles_per_warp = (p.num_ptiles + blocks * wpb - 1) / (blocks * wpb);
    p.group_blocks = (blocks == (long long)sms * occ) ? occ : 1;
    kern<<<(unsigned)blocks, threads, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace cs
