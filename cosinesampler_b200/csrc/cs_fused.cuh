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
//  3. With binned points the scatter pre-reduces before it touches L2: every *walker* (the L = K/4 lanes that
//     share a point) owns a contiguous range of the binned points and keeps, per cell, a private 3x3-texel
//     window of the accumulator in shared memory (the N cells of a point differ by the sub-texel multicell
//     offset, so their corners lie within one texel of each other: 3x3 covers every cell's 2x2 corners).
//     Contributions are added with plain ld/st.shared (the window is private: no atomics), and a window is
//     flushed with one red.global.add.v4.f32 per touched texel when the walker moves to the next texel.  At
//     16-65 points per texel that is 7-28x fewer reds than one per (point, corner).
//     (Warp-level __match_any_sync aggregation was the alternative: it needs the points of a warp to agree in
//     corner AND sub-texel shift, and a 3-step segmented shuffle per contribution; ATOMS-based privatisation
//     costs 2 cycles per lane.  Private windows need neither.)
#pragma once
#include <limits.h>

#include "cs_jet.cuh"

namespace cs {

// ---------------------------------------------------------------------------------------------------------
// Point binning: counting sort of the coordinates on a tile-major texel key
// ---------------------------------------------------------------------------------------------------------
struct BinParams {
    int dim;
    int size[3];
    int shift;                 // texel coordinates are coarsened by >> shift so that nbins stays bounded
    int ntx, nty, ntz;         // tiles per axis (2D: 8x8 texels, 3D: 4x4x4 texels per tile)
    unsigned nbins;
    long long P;
    const float* coords;       // [P, dim]
    const float* offset;       // [N] (device); cell 0's offset enters the key.  nullable = 0
    int align, multicell, index_mode;
};

// low-corner texel of cell 0 along one axis, clamped into the cell: a locality key, never a correctness
// matter (the fused kernel compares the real corner indices of every point)
__device__ __forceinline__ int bin_axis(float g, int size, float off0, const BinParams& p) {
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
    int l = (int)floorf(i);
    l = max(0, min(size - 1, l));
    return l >> p.shift;
}

__device__ __forceinline__ unsigned bin_key(const float* gp, const BinParams& p) {
    const float off0 = p.offset ? __ldg(p.offset) : 0.f;
    const int lx = bin_axis(__ldg(gp), p.size[0], off0, p);
    const int ly = bin_axis(__ldg(gp + 1), p.size[1], off0, p);
    if (p.dim == 2) {
        const unsigned tile = (unsigned)((ly >> 3) * p.ntx + (lx >> 3));
        return (tile << 6) | (unsigned)(((ly & 7) << 3) | (lx & 7));
    }
    const int lz = bin_axis(__ldg(gp + 2), p.size[2], off0, p);
    const unsigned tile = (unsigned)(((lz >> 2) * p.nty + (ly >> 2)) * p.ntx + (lx >> 2));
    return (tile << 6) | (unsigned)(((lz & 3) << 4) | ((ly & 3) << 2) | (lx & 3));
}

static __global__ void __launch_bounds__(256) cs_bin_count_kernel(const BinParams p, unsigned* __restrict__ hist,
                                                           unsigned* __restrict__ rank) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.P;
         i += (long long)gridDim.x * blockDim.x) {
        const unsigned key = bin_key(p.coords + i * p.dim, p);
        rank[i] = atomicAdd(hist + key, 1u);
    }
}

// exclusive scan of hist[0..nbins) in place, one block
static __global__ void __launch_bounds__(1024) cs_bin_scan_kernel(unsigned* __restrict__ hist, unsigned nbins) {
    __shared__ unsigned part[1024];
    const unsigned per = (nbins + 1023u) / 1024u;
    const unsigned b0 = threadIdx.x * per;
    const unsigned b1 = min(nbins, b0 + per);
    unsigned s = 0;
    for (unsigned b = b0; b < b1; ++b) s += hist[b];
    part[threadIdx.x] = s;
    __syncthreads();
    // Hillis-Steele inclusive scan over the 1024 partial sums
    for (int o = 1; o < 1024; o <<= 1) {
        unsigned v = (threadIdx.x >= (unsigned)o) ? part[threadIdx.x - o] : 0u;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    unsigned run = part[threadIdx.x] - s;
    for (unsigned b = b0; b < b1; ++b) {
        const unsigned c = hist[b];
        hist[b] = run;
        run += c;
    }
}

static __global__ void __launch_bounds__(256) cs_bin_scatter_kernel(const BinParams p, const unsigned* __restrict__ offs,
                                                             const unsigned* __restrict__ rank,
                                                             float* __restrict__ sorted, int* __restrict__ perm) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.P;
         i += (long long)gridDim.x * blockDim.x) {
        const float* gp = p.coords + i * p.dim;
        const unsigned key = bin_key(gp, p);
        const long long pos = (long long)__ldg(offs + key) + __ldg(rank + i);
        float* dst = sorted + pos * p.dim;
        for (int a = 0; a < p.dim; ++a) dst[a] = __ldg(gp + a);
        if (perm) perm[pos] = (int)i;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Grid-sized mixes with the head's first layer
// ---------------------------------------------------------------------------------------------------------
constexpr int MIX_MAXK = 32;

// Vh[n, t, k] = sum_c W1[k, c] * V[n, c, t]   (channel-first in, channel-last out: the staging transpose of
// cs_to_channel_last and the first Linear layer in one pass)
template <int K>
__global__ void __launch_bounds__(256) cs_head_premix_kernel(const float* __restrict__ V, const float* __restrict__ W1,
                                                             float* __restrict__ Vh, int C, long long T, long long NT) {
    extern __shared__ float w1s[];           // [C][K]: w1s[c*K + k] = W1[k*C + c]
    for (int e = threadIdx.x; e < K * C; e += blockDim.x) w1s[(e % C) * K + (e / C)] = __ldg(W1 + e);
    __syncthreads();
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
// gVh is channel-last [N,T,K] (KFIRST = false) or channel-first [N,K,T] (the output of the peer reduce).
constexpr int POSTMIX_TT = 128;              // texels per block tile
template <int K, bool KFIRST>
__global__ void __launch_bounds__(POSTMIX_TT) cs_head_postmix_kernel(const float* __restrict__ gVh,
                                                                     const float* __restrict__ V,
                                                                     const float* __restrict__ W1,
                                                                     float* __restrict__ gInput, int accumulate,
                                                                     float* __restrict__ gW1, int C, long long T,
                                                                     long long ntiles_per_cell, long long ntiles) {
    extern __shared__ float sm[];
    float* w1s = sm;                                  // [C][K]
    float* gs = w1s + C * K;                          // [TT][K+1]
    float* vs = gs + POSTMIX_TT * (K + 1);            // [C][TT]
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
        for (int k = 0; k < K; ++k) gs[tid * (K + 1) + k] = g[k];
        for (int c = 0; c < C; ++c) {
            float v = 0.f;
            if (ok) v = __ldg(V + (n * C + c) * T + t);
            if (ok && gInput) {
                float gi = 0.f;
                const float* wr = w1s + c * K;
#pragma unroll
                for (int k = 0; k < K; ++k) gi = fmaf(wr[k], g[k], gi);
                float* o = gInput + (n * C + c) * T + t;
                if (accumulate) *o += gi; else *o = gi;
            }
            vs[c * POSTMIX_TT + tid] = v;
        }
        __syncthreads();
        if (gW1) {
#pragma unroll
            for (int i = 0; i < MAXPP; ++i) {
                const int pr = tid + i * POSTMIX_TT;
                if (pr < npairs) {
                    const int k = pr % K, c = pr / K;
                    float s = 0.f;
#pragma unroll 8
                    for (int tt = 0; tt < POSTMIX_TT; ++tt) s = fmaf(gs[tt * (K + 1) + k], vs[c * POSTMIX_TT + tt], s);
                    wacc[i] += s;
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

// ---------------------------------------------------------------------------------------------------------
// The fused step
// ---------------------------------------------------------------------------------------------------------
struct FusedParams {
    int N, K;                 // cells, hidden width = channels of Vh / gVh
    int size[3];
    int tstride[3];
    long long P;
    long long cell_stride;    // T * K
    const float* Vh;          // [N, T, K]  W1-mixed cells (cs_head_premix)
    float* gVh;               // [N, T, K]  accumulated
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
    long long pts_per_walker; // AGG: contiguous points per walker (a multiple of PPQ)
    long long num_ptiles;     // !AGG: warp tiles of PTS consecutive points
    int win_stride;           // AGG: float4 per walker window (N*9*L padded)
};

// Phase 1 for one (cell, point): field 0 = (base texel, corner-valid mask, low corner x, low corner y) as int
// bits, field 1 + a = (w0, w1, m k', m^2 k'') of axis a.
template <int DIM, int PTS>
__device__ __forceinline__ void build_fused_record(float4* rec4, int i, const float (&g)[DIM], bool in_range,
                                                   float off, const FusedParams& p) {
    int base = 0, mask = 0, l0 = 0, l1 = 0;
    float4 ax[DIM];
#pragma unroll
    for (int a = 0; a < DIM; ++a) ax[a] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (in_range) {
        bool ok = true;
        bool lo_ok[DIM], hi_ok[DIM];
#pragma unroll
        for (int a = 0; a < DIM; ++a) {
            const AxisRec ar = axis_setup(g[a], p.size[a], off, p, p.align != 0, 2);
            ok = ok && ar.ok;
            base += ar.l * p.tstride[a];
            if (a == 0) l0 = ar.l;
            if (a == 1) l1 = ar.l;
            lo_ok[a] = (ar.l >= 0) && (ar.l < p.size[a]);
            hi_ok[a] = (ar.l + 1 >= 0) && (ar.l + 1 < p.size[a]);
            ax[a] = make_float4(ar.w0, ar.w1, ar.d, ar.e);
        }
        if (ok) {
#pragma unroll
            for (int c = 0; c < (1 << DIM); ++c) {
                bool valid = true;
#pragma unroll
                for (int a = 0; a < DIM; ++a) valid = valid && (((c >> a) & 1) ? hi_ok[a] : lo_ok[a]);
                if (valid) mask |= 1 << c;
            }
        }
    }
    rec4[i] = make_float4(__int_as_float(base), __int_as_float(mask), __int_as_float(l0), __int_as_float(l1));
#pragma unroll
    for (int a = 0; a < DIM; ++a) rec4[(1 + a) * PTS + i] = ax[a];
}

__device__ __forceinline__ float tanh_ex2(float x) {
    // one ex2.approx + one rcp.approx: absolute error a few 1e-7 (|tanh| <= 1 is the scale that matters:
    // u = w2 . tanh, s1 = 1 - tanh^2); saturates cleanly to +-1 for large |x|
    const float e = __expf(-2.f * fabsf(x));
    return copysignf(__fdividef(1.f - e, 1.f + e), x);
}

__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// flush one (walker, cell) window: one red.global.add.v4.f32 per touched, in-bounds texel; the window is
// zeroed again.  Called by the walkers that move on while the others idle (warp-divergent by design).
static __device__ __noinline__ void flush_window(float4* w, int L, int touched, int ox, int oy, float* gcell, int K,
                                          int W, int H) {
#pragma unroll 1
    for (int s = 0; s < 9; ++s) {
        if ((touched >> s) & 1) {
            const float4 v = w[s * L];
            w[s * L] = make_float4(0.f, 0.f, 0.f, 0.f);
            const int x = ox + (s % 3), y = oy + (s / 3);
            if ((unsigned)x < (unsigned)W && (unsigned)y < (unsigned)H)
                red_add_v4(gcell + ((long long)y * W + x) * K, v.x, v.y, v.z, v.w);
        }
    }
}

// register budget: 9 warps per SM (3 blocks of 96 threads, the shared-memory limit of the aggregating variant
// at N = 4 cells) leave 227 registers per thread
#ifndef CS_FUSED_MAXREG_AGG
#define CS_FUSED_MAXREG_AGG 224
#endif
#ifndef CS_FUSED_MAXREG
#define CS_FUSED_MAXREG 224
#endif
constexpr int FUSED_THREADS_AGG = 96;
constexpr int FUSED_THREADS = 96;

template <int DIM, int LSHIFT, bool AGG>
__global__ void __maxnreg__(AGG ? CS_FUSED_MAXREG_AGG : CS_FUSED_MAXREG)
cs_pde_fused_kernel(const FusedParams p) {
    constexpr int NCORN = 1 << DIM;
    constexpr int J = 1 + 2 * DIM;
    constexpr int L = 1 << LSHIFT;
    constexpr int NW = 32 >> LSHIFT;                 // walkers per warp
    constexpr int PPQ = (DIM == 2) ? 4 : 2;          // points per walker and iteration
    constexpr int PTS = PPQ * NW;                    // points per warp and iteration
    constexpr int PPL = (PTS + 31) / 32;
    constexpr int REC1 = (1 + DIM) * PTS;            // float4 per record buffer
    constexpr int FULL = (1 << NCORN) - 1;
    static_assert(!AGG || DIM == 2, "aggregation windows are two-dimensional");

    extern __shared__ float4 smem4[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int q = lane >> LSHIFT;
    const int j = lane & (L - 1);
    const int ncells = p.N;
    const int K = p.K;
    const int per_warp = REC1 + (AGG ? NW * p.win_stride + NW * ncells : 0);
    float4* rec = smem4 + (size_t)warp * per_warp;
    float4* win = rec + REC1 + (AGG ? q * p.win_stride + j : 0);               // this lane's column of its walker's windows
    int4* hdr = reinterpret_cast<int4*>(rec + REC1 + (AGG ? NW * p.win_stride : 0)) + (AGG ? q * ncells : 0);

    if (AGG) {
        for (int s = 0; s < ncells * 9; ++s) win[s * L] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j == 0)
            for (int n = 0; n < ncells; ++n) hdr[n] = make_int4(INT_MIN, INT_MIN, 0, 0);
        __syncwarp();
    }

    int coff[NCORN];
#pragma unroll
    for (int c = 0; c < NCORN; ++c) {
        coff[c] = 0;
#pragma unroll
        for (int a = 0; a < DIM; ++a) coff[c] += ((c >> a) & 1) * p.tstride[a];
    }
    float b1k[4], w2k[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { b1k[k] = __ldg(p.b1 + 4 * j + k); w2k[k] = __ldg(p.w2 + 4 * j + k); }
    const float b2 = __ldg(p.b2);
    float gb1acc[4] = {0.f, 0.f, 0.f, 0.f}, gw2acc[4] = {0.f, 0.f, 0.f, 0.f};
    float gb2acc = 0.f, lossacc = 0.f;

    const long long gw = (long long)blockIdx.x * wpb + warp;
    const long long tw = (long long)gridDim.x * wpb;
    const long long R = p.pts_per_walker;
    // first point of (walker qq, iteration it)
    auto first_point = [&](int qq, long long it) -> long long {
        return AGG ? (gw * NW + qq) * R + it * PPQ : (gw + it * tw) * PTS + (long long)qq * PPQ;
    };
    auto iteration_live = [&](long long it) -> bool {
        return AGG ? (it * PPQ < R && first_point(0, it) < p.P) : (gw + it * tw < p.num_ptiles);
    };
    auto load_coords = [&](float (&g)[PPL][DIM], bool (&inr)[PPL], long long it) {
#pragma unroll
        for (int u = 0; u < PPL; ++u) {
            const int i = u * 32 + lane;
            const long long pi = first_point(i / PPQ, it) + (i % PPQ);
            inr[u] = (i < PTS) && (pi < p.P);
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
    // records of cell n for the PTS points of this iteration; returns "every corner of every point is valid"
    auto phase1 = [&](const float (&g)[PPL][DIM], const bool (&inr)[PPL], int n) -> bool {
        const float off = __ldg(p.offset + n);
        __syncwarp();                               // everyone is done reading the record buffer
        bool allv = true;
#pragma unroll
        for (int u = 0; u < PPL; ++u) {
            const int i = u * 32 + lane;
            if (i < PTS) {
                build_fused_record<DIM, PTS>(rec, i, g[u], inr[u], off, p);
                allv = allv && (__float_as_int(rec[i].y) == FULL);
            }
        }
        return __all_sync(0xffffffffu, allv);       // also a warp barrier: records are visible
    };

    float gcur[PPL][DIM], gnext[PPL][DIM];
    bool icur[PPL], inext[PPL];
    if (iteration_live(0)) {
    load_coords(gcur, icur, 0);
    for (long long it = 0;; ++it) {
        const bool have_next = iteration_live(it + 1);
        if (have_next) load_coords(gnext, inext, it + 1);

        // ---- A: gather.  acc[jt][t][k] = H_jt of point t, hidden unit 4j + k, summed over the cells
        float acc[J][PPQ][4];
#pragma unroll
        for (int jt = 0; jt < J; ++jt)
#pragma unroll
            for (int t = 0; t < PPQ; ++t)
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[jt][t][k] = 0.f;
        for (int n = 0; n < ncells; ++n) {
            const bool allv = phase1(gcur, icur, n);
            const float* vsrc = p.Vh + (long long)n * p.cell_stride + 4 * j;
            float4 v[PPQ][NCORN];
#pragma unroll
            for (int t = 0; t < PPQ; ++t) {
                const float4 hd = rec[PPQ * q + t];
                const int base = __float_as_int(hd.x);
                const int mask = __float_as_int(hd.y);
#pragma unroll
                for (int c = 0; c < NCORN; ++c) {
                    if (allv) v[t][c] = ldg_f4(vsrc + (long long)(base + coff[c]) * K);
                    else v[t][c] = ((mask >> c) & 1) ? ldg_f4(vsrc + (long long)(base + coff[c]) * K)
                                                     : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int t = 0; t < PPQ; ++t) {
                const int ri = PPQ * q + t;
                const float4 ax = rec[1 * PTS + ri];
                const float4 ay = rec[2 * PTS + ri];
                if (DIM == 2) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const SlabJet sj = slab_contract(f4get(v[t][0], k), f4get(v[t][1], k), f4get(v[t][2], k),
                                                         f4get(v[t][3], k), ax, ay);
                        acc[0][t][k] += sj.A;
                        acc[1][t][k] += sj.X;
                        acc[2][t][k] += sj.Y;
                        acc[3][t][k] += sj.XX;
                        acc[4][t][k] += sj.YY;
                    }
                } else {
                    const float4 az = rec[(DIM == 3 ? 3 : 2) * PTS + ri];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const SlabJet lo = slab_contract(f4get(v[t][0], k), f4get(v[t][1], k), f4get(v[t][2], k),
                                                         f4get(v[t][3], k), ax, ay);
                        const SlabJet hi = slab_contract(f4get(v[t][4 % NCORN], k), f4get(v[t][5 % NCORN], k),
                                                         f4get(v[t][6 % NCORN], k), f4get(v[t][7 % NCORN], k), ax, ay);
                        acc[0][t][k] += fmaf(hi.A, az.y, lo.A * az.x);
                        acc[1][t][k] += fmaf(hi.X, az.y, lo.X * az.x);
                        acc[2][t][k] += fmaf(hi.Y, az.y, lo.Y * az.x);
                        acc[DIM][t][k] += (hi.A - lo.A) * az.z;
                        acc[1 + DIM][t][k] += fmaf(hi.XX, az.y, lo.XX * az.x);
                        acc[(2 + DIM) % J][t][k] += fmaf(hi.YY, az.y, lo.YY * az.x);
                        acc[2 * DIM][t][k] += (lo.A - hi.A) * az.w;
                    }
                }
            }
        }

        // ---- B: head + residual + loss, and d loss / d H_jt in place (test_2d.py:42-127 in closed form)
#pragma unroll
        for (int t = 0; t < PPQ; ++t) {
            float th[4], s1[4], s2[4];
            float pu = 0.f, pua[DIM], puaa[DIM];
#pragma unroll
            for (int a = 0; a < DIM; ++a) { pua[a] = 0.f; puaa[a] = 0.f; }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                th[k] = tanh_ex2(acc[0][t][k] + b1k[k]);
                s1[k] = 1.f - th[k] * th[k];
                s2[k] = -2.f * th[k] * s1[k];
                pu = fmaf(w2k[k], th[k], pu);
#pragma unroll
                for (int a = 0; a < DIM; ++a) {
                    const float hd = acc[1 + a][t][k], hdd = acc[1 + DIM + a][t][k];
                    pua[a] = fmaf(w2k[k], s1[k] * hd, pua[a]);
                    puaa[a] = fmaf(w2k[k], s2[k] * hd * hd + s1[k] * hdd, puaa[a]);
                }
            }
#pragma unroll
            for (int o = 1; o < L; o <<= 1) {
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
            const bool valid = first_point(q, it) + t < p.P;
            const float gg = valid ? 2.f * p.scale * f : 0.f;
            const float gsc = gg * (p.c_u + 3.f * p.c_u3 * u * u);
            if (j == 0) {
                if (valid) lossacc = fmaf(f, f, lossacc);
                gb2acc += gsc;
            }
            float g1c[DIM], g2c[DIM];
#pragma unroll
            for (int a = 0; a < DIM; ++a) { g1c[a] = gg * p.c1[a]; g2c[a] = gg * p.c2[a]; }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float d1 = s1[k], d2 = s2[k];
                const float d3 = -2.f * (d1 * d1 + th[k] * d2);
                float gw2 = gsc * th[k];
                float gh = gsc * d1;
#pragma unroll
                for (int a = 0; a < DIM; ++a) {
                    const float hd = acc[1 + a][t][k], hdd = acc[1 + DIM + a][t][k];
                    gw2 += g1c[a] * d1 * hd + g2c[a] * (d2 * hd * hd + d1 * hdd);
                    gh += g1c[a] * d2 * hd + g2c[a] * (d3 * hd * hd + d2 * hdd);
                    acc[1 + a][t][k] = w2k[k] * (g1c[a] * d1 + g2c[a] * 2.f * d2 * hd);
                    acc[1 + DIM + a][t][k] = w2k[k] * g2c[a] * d1;
                }
                gh *= w2k[k];
                acc[0][t][k] = gh;
                gb1acc[k] += gh;
                gw2acc[k] += gw2;
            }
        }

        // ---- C: scatter d loss / d H_jt into gVh: per corner  Wy u_x + Wx (+-beta)  (separable adjoint)
        int l0x[PPQ], l0y[PPQ];                      // low corner of each point in cell 0 (window anchors)
#pragma unroll
        for (int t = 0; t < PPQ; ++t) { l0x[t] = 0; l0y[t] = 0; }
        for (int n = 0; n < ncells; ++n) {
            phase1(gcur, icur, n);
            float* gcell = p.gVh + (long long)n * p.cell_stride + 4 * j;
            float4* wn = win + (AGG ? n * 9 * L : 0);
            int ox = 0, oy = 0, touched = 0;
            if (AGG) { const int4 h = hdr[n]; ox = h.x; oy = h.y; touched = h.z; }
#pragma unroll
            for (int t = 0; t < PPQ; ++t) {
                const int ri = PPQ * q + t;
                const float4 hd = rec[ri];
                const int base = __float_as_int(hd.x);
                const int mask = __float_as_int(hd.y);
                const float4 ax = rec[1 * PTS + ri];
                const float4 ay = rec[2 * PTS + ri];
                float cv[NCORN][4];
                if (DIM == 2) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float alpha = acc[1][t][k] * ax.z - acc[3][t][k] * ax.w;
                        const float beta = acc[2][t][k] * ay.z - acc[4][t][k] * ay.w;
                        const float u0 = fmaf(acc[0][t][k], ax.x, -alpha);
                        const float u1 = fmaf(acc[0][t][k], ax.y, alpha);
                        const float bx0 = ax.x * beta, bx1 = ax.y * beta;
                        cv[0][k] = fmaf(ay.x, u0, -bx0);
                        cv[1][k] = fmaf(ay.x, u1, -bx1);
                        cv[2][k] = fmaf(ay.y, u0, bx0);
                        cv[3][k] = fmaf(ay.y, u1, bx1);
                    }
                } else {
                    const float4 az = rec[(DIM == 3 ? 3 : 2) * PTS + ri];
                    const float p00 = ax.x * ay.x, p10 = ax.y * ay.x, p01 = ax.x * ay.y, p11 = ax.y * ay.y;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float alpha = acc[1][t][k] * ax.z - acc[1 + DIM][t][k] * ax.w;
                        const float beta = acc[2][t][k] * ay.z - acc[(2 + DIM) % J][t][k] * ay.w;
                        const float gamma = acc[DIM][t][k] * az.z - acc[2 * DIM][t][k] * az.w;
                        const float u0 = fmaf(acc[0][t][k], ax.x, -alpha);
                        const float u1 = fmaf(acc[0][t][k], ax.y, alpha);
                        const float bx0 = ax.x * beta, bx1 = ax.y * beta;
                        const float c00 = fmaf(ay.x, u0, -bx0), c10 = fmaf(ay.x, u1, -bx1);
                        const float c01 = fmaf(ay.y, u0, bx0), c11 = fmaf(ay.y, u1, bx1);
                        cv[0][k] = fmaf(az.x, c00, -p00 * gamma);
                        cv[1][k] = fmaf(az.x, c10, -p10 * gamma);
                        cv[2][k] = fmaf(az.x, c01, -p01 * gamma);
                        cv[3][k] = fmaf(az.x, c11, -p11 * gamma);
                        cv[4 % NCORN][k] = fmaf(az.y, c00, p00 * gamma);
                        cv[5 % NCORN][k] = fmaf(az.y, c10, p10 * gamma);
                        cv[6 % NCORN][k] = fmaf(az.y, c01, p01 * gamma);
                        cv[7 % NCORN][k] = fmaf(az.y, c11, p11 * gamma);
                    }
                }
                bool direct = true;
                if (AGG) {
                    const int lx = __float_as_int(hd.z), ly = __float_as_int(hd.w);
                    if (n == 0) { l0x[t] = lx; l0y[t] = ly; }
                    if (mask == FULL) {
                        // cell n's low corner is that of cell 0 plus 0 or 1 per axis (the sub-texel multicell
                        // offset), alternating from point to point inside one texel of cell 0: a window anchored
                        // at cell 0's low corner covers both.  Re-anchor only when the point does not fit.
                        int dx = lx - ox, dy = ly - oy;
                        if ((unsigned)dx > 1u || (unsigned)dy > 1u) {
                            if (touched) flush_window(wn, L, touched, ox, oy, gcell, K, p.size[0], p.size[1]);
                            touched = 0;
                            ox = l0x[t]; oy = l0y[t];
                            dx = lx - ox; dy = ly - oy;
                        }
                        if ((unsigned)dx <= 1u && (unsigned)dy <= 1u) {
                            const int s0 = dy * 3 + dx;
#pragma unroll
                            for (int c = 0; c < NCORN; ++c) {
                                float4* slot = wn + (s0 + (c & 1) + 3 * (c >> 1)) * L;
                                float4 o = *slot;
                                o.x += cv[c][0]; o.y += cv[c][1]; o.z += cv[c][2]; o.w += cv[c][3];
                                *slot = o;
                            }
                            touched |= 0x1B << s0;
                            direct = false;
                        }
                    }
                }
                if (direct && mask) {
#pragma unroll
                    for (int c = 0; c < NCORN; ++c)
                        if ((mask >> c) & 1)
                            red_add_v4(gcell + (long long)(base + coff[c]) * K, cv[c][0], cv[c][1], cv[c][2], cv[c][3]);
                }
            }
            if (AGG) {
                __syncwarp();
                if (j == 0) hdr[n] = make_int4(ox, oy, touched, 0);
            }
        }

        if (!have_next) break;
#pragma unroll
        for (int u = 0; u < PPL; ++u) {
            icur[u] = inext[u];
#pragma unroll
            for (int a = 0; a < DIM; ++a) gcur[u][a] = gnext[u][a];
        }
    }

    if (AGG) {
        __syncwarp();
        for (int n = 0; n < ncells; ++n) {
            const int4 h = hdr[n];
            if (h.z) flush_window(win + n * 9 * L, L, h.z, h.x, h.y, p.gVh + (long long)n * p.cell_stride + 4 * j, K,
                                  p.size[0], p.size[1]);
        }
    }
    }

    // ---- head-parameter gradients and loss: walkers of a warp (shuffles) -> warps (shared memory) -> one
    // atomic per block and element
#pragma unroll
    for (int o = L; o < 32; o <<= 1) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            gb1acc[k] += __shfl_xor_sync(0xffffffffu, gb1acc[k], o);
            gw2acc[k] += __shfl_xor_sync(0xffffffffu, gw2acc[k], o);
        }
        gb2acc += __shfl_xor_sync(0xffffffffu, gb2acc, o);
        lossacc += __shfl_xor_sync(0xffffffffu, lossacc, o);
    }
    __syncthreads();                                 // every warp is done with its records / windows
    float* red = reinterpret_cast<float*>(smem4);    // [wpb][2K + 2]
    const int RW = 2 * K + 2;
    if (q == 0) {
        float* rw = red + warp * RW;
#pragma unroll
        for (int k = 0; k < 4; ++k) { rw[4 * j + k] = gb1acc[k]; rw[K + 4 * j + k] = gw2acc[k]; }
        if (j == 0) { rw[2 * K] = gb2acc; rw[2 * K + 1] = lossacc; }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < RW; e += blockDim.x) {
        float s = 0.f;
        for (int w = 0; w < wpb; ++w) s += red[w * RW + e];
        float* dst = (e < K) ? p.gb1 + e : (e < 2 * K) ? p.gw2 + (e - K) : (e == 2 * K) ? p.gb2 : p.loss_sum;
        atomicAdd(dst, s);
    }
}

// ---------------------------------------------------------------------------------------------------------
// launch
// ---------------------------------------------------------------------------------------------------------
inline int fused_win_stride(int N, int L) {
    int s = N * 9 * L;
    s += (4 - (s % 8) + 8) % 8;          // walker stride = 4 mod 8 float4: neighbouring walkers use opposite bank halves
    return s;
}

template <int DIM, int LSHIFT, bool AGG>
cudaError_t launch_fused_one(FusedParams& p, cudaStream_t stream) {
    constexpr int L = 1 << LSHIFT;
    constexpr int NW = 32 >> LSHIFT;
    constexpr int PPQ = (DIM == 2) ? 4 : 2;
    constexpr int PTS = PPQ * NW;
    constexpr int REC1 = (1 + DIM) * PTS;
    auto kern = cs_pde_fused_kernel<DIM, LSHIFT, AGG>;
    const int threads = AGG ? FUSED_THREADS_AGG : FUSED_THREADS;
    const int wpb = threads / 32;
    p.win_stride = AGG ? fused_win_stride(p.N, L) : 0;
    const size_t per_warp = (size_t)(REC1 + (AGG ? NW * p.win_stride + NW * p.N : 0)) * sizeof(float4);
    size_t smem = wpb * per_warp;
    const size_t red_bytes = (size_t)wpb * (2 * p.K + 2) * sizeof(float);
    if (smem < red_bytes) smem = red_bytes;
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) occ = 1;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    long long blocks = (long long)sms * occ;
    p.num_ptiles = (p.P + PTS - 1) / PTS;
    if (AGG) {
        // every walker owns pts_per_walker contiguous points; short inputs use fewer blocks so that a walker
        // still sees a few texels' worth of points
        long long walkers = blocks * wpb * NW;
        long long R = (p.P + walkers - 1) / walkers;
        const long long minR = 8 * PPQ;
        if (R < minR) {
            R = minR;
            walkers = (p.P + R - 1) / R;
            blocks = (walkers + wpb * NW - 1) / (wpb * NW);
        }
        R = (R + PPQ - 1) / PPQ * PPQ;
        p.pts_per_walker = R;
    } else {
        const long long need = (p.num_ptiles + wpb - 1) / wpb;
        if (blocks > need) blocks = need;
    }
    if (blocks < 1) return cudaSuccess;
    kern<<<(unsigned)blocks, threads, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace cs
