// cs_scalar.cuh -- double- and half-precision paths of the four stages.
//
// The reference dispatches float / double / half (AT_DISPATCH_FLOATING_TYPES_AND_HALF, cu2d:905) but its
// double and half instantiations cannot run: the per-cell offset is always a float tensor (modules_2d.py:25)
// and is read through TensorInfo<scalar_t> (cu2d:914).  This file provides what that dispatch promises for
// fp64: the same four entry points on double tensors (the offset stays fp32, as the reference builds it),
// semantics and quirks identical to the fp32 engine (cs_engine.cuh), computed entirely in
// double.  It is a correctness path -- one thread per (cell, point), serial channel loop over the reference's
// channel-first layout, scalar atomicAdd(double) -- for callers who validate a PDE residual in fp64; it makes no
// throughput claim.
//
// fp16 (ScalarParamsT<__half>): tensors are stored as half, every load is widened to float, all arithmetic is
// float (the same formulas), outputs are rounded to half once; gInput is accumulated in a float workspace with
// atomicAdd(float) and rounded to half by cs_f32_to_f16_kernel afterwards -- accumulating thousands of half
// atomics would lose the sum.  Coordinates are half too, as the reference's dispatch would have them (grid is read
// through scalar_t, cu2d:912): 11 significant bits, i.e. 1/8 of a texel at 256 texels; that is the caller's choice.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "cs_engine.cuh"

namespace cs {

template <typename T> struct ScalarTypes;
template <> struct ScalarTypes<double> {
    typedef double CT;           // arithmetic type
    typedef double AT;           // accumulator of gInput
    static __device__ __forceinline__ double ld(const double* p) { return *p; }
    static __device__ __forceinline__ void st(double* p, double v) { *p = v; }
};
template <> struct ScalarTypes<__half> {
    typedef float CT;
    typedef float AT;
    static __device__ __forceinline__ float ld(const __half* p) { return __half2float(*p); }
    static __device__ __forceinline__ void st(__half* p, float v) { *p = __float2half_rn(v); }
};
__device__ __forceinline__ void cs_sincospi(double x, double* s, double* c) { sincospi(x, s, c); }
__device__ __forceinline__ void cs_sincospi(float x, float* s, float* c) { sincospif(x, s, c); }

template <typename S>
struct ScalarParamsT {
    typedef typename ScalarTypes<S>::AT AT;
    int N, C;
    int size[3];
    int tstride[3];
    long long T;                 // texels per cell
    long long P;
    const S* V;                  // input        [N, C, T]
    const S* U;                  // gOutInput    [N, C, T] or nullptr            (BB)
    AT* acc;                     // gInput       [N, C, T] or nullptr, zero-initialised (fp16: the float workspace)
    const S* x1; long long x1_sn, x1_sc;          // gOut       strided [N, C, P]
    const S* x2; long long x2_sn, x2_sc;          // gOutggOut  (BBB, fused b_input pass) or nullptr
    S* y;                        // out / ggOut  [N, C, P] or nullptr
    const S* grid; long long grid_sn;             // [N, P, dim]
    const S* gog;                // gOutGrid     [N, P, dim]
    const S* gogg;               // gOutgGrid    [N, P, dim]
    S* ggrid;                    // gGrid        [N, P, dim] or nullptr
    const float* offset;         // [N]  fp32, as modules_2d.py:24-27 builds it
    int pad, align, kernel, multicell;
};
typedef ScalarParamsT<double> ScalarParams;

template <typename CT>
struct AxisT {
    int l;
    CT w[2], dw[2], ew[2];       // weight, d weight / d coordinate, d2 weight / d coordinate^2 for (low, high)
    bool ok, inb[2];
};

template <typename CT>
__device__ __forceinline__ CT clip_grad_t(CT& i, int size) {
    const CT hi = (CT)(size - 1);
    if (i <= CT(0)) { i = CT(0); return CT(0); }
    if (i >= hi) { i = hi; return CT(0); }
    return CT(1);
}
template <typename CT>
__device__ __forceinline__ CT reflect_grad_t(CT& i, int twice_low, int twice_high) {
    if (twice_low == twice_high) { i = CT(0); return CT(0); }
    const CT lo = (CT)twice_low * CT(0.5);
    const CT span = (CT)(twice_high - twice_low) * CT(0.5);
    CT x = i - lo, sign = CT(1);
    if (x < CT(0)) { sign = CT(-1); x = -x; }
    const CT extra = fmod(x, span);
    const long long flips = (long long)floor(x / span);
    if ((flips & 1) == 0) { i = extra + lo; return sign; }
    i = span - extra + lo;
    return -sign;
}

// cu2d:53-261 in CT (index map, padding, kernel functions and their derivatives)
template <typename CT, typename P>
__device__ __forceinline__ AxisT<CT> axis_setup_t(CT g, int size, CT off, const P& p, bool align) {
    AxisT<CT> a;
    CT i, m;
    if (align) {
        const CT s = (CT)(size - 1 - (p.multicell ? 1 : 0));
        m = s * CT(0.5);
        i = ((g + CT(1)) * CT(0.5)) * s + off;
    } else {
        const CT s = (CT)size;
        m = s * CT(0.5);
        i = ((g + CT(1)) * s - CT(1)) * CT(0.5) + off;
    }
    const CT huge = sizeof(CT) == 8 ? CT(1.0e300) : CT(3.0e38);
    a.ok = (p.pad == 1) ? (fabs(i) <= huge) : (fabs(i) < CT(1.0e9));
    if (!a.ok) i = CT(0);
    if (p.pad == 1) {
        m *= clip_grad_t<CT>(i, size);
    } else if (p.pad == 2) {
        const CT gr = align ? reflect_grad_t<CT>(i, 0, 2 * (size - 2)) : reflect_grad_t<CT>(i, -1, 2 * size - 1);
        m *= gr * clip_grad_t<CT>(i, size);
    }
    const CT lf = floor(i);
    a.l = (int)lf;
    const CT r = (lf + CT(1)) - i;
    CT k0, k1, k2, whi;
    if (p.kernel == 0) {
        CT sn, cn;
        cs_sincospi(r, &sn, &cn);
        k0 = CT(0.5) * (CT(1) - cn);
        k1 = CT(0.5 * 3.14159265358979323846) * sn;
        k2 = CT(0.5 * 9.86960440108935861883) * cn;
        whi = CT(1) - k0;
    } else if (p.kernel == 2) {
        k0 = r * r * (CT(3) - CT(2) * r);
        k1 = CT(6) * r * (CT(1) - r);
        k2 = CT(6) - CT(12) * r;
        whi = CT(1) - k0;
    } else {
        k0 = r; k1 = CT(1); k2 = CT(0);
        whi = i - lf;
    }
    a.w[0] = k0; a.w[1] = whi;
    a.dw[0] = -m * k1; a.dw[1] = m * k1;
    a.ew[0] = m * m * k2; a.ew[1] = -m * m * k2;
    a.inb[0] = (a.l >= 0) && (a.l < size);
    a.inb[1] = (a.l + 1 >= 0) && (a.l + 1 < size);
    return a;
}

// One thread per (cell, point).  STAGE as in cs_engine.cuh (ST_F, ST_B, ST_BB, ST_BBB).
template <int DIM, int STAGE, typename T>
__global__ void __launch_bounds__(256) cs_scalar_stage_kernel(const ScalarParamsT<T> p) {
    typedef ScalarTypes<T> TT;
    typedef typename TT::CT CT;
    constexpr int NCORN = 1 << DIM;
    const long long total = (long long)p.N * p.P;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(idx / p.P);
        const long long pi = idx - (long long)n * p.P;
        const CT off = (CT)__ldg(p.offset + n);
        const T* gp = p.grid + (long long)n * p.grid_sn + pi * DIM;
        // the 2D forward ignores align_corners and uses 1 (cu2d:307-308)
        const bool align = (STAGE == ST_F && DIM == 2) ? true : (p.align != 0);
        AxisT<CT> ax[DIM];
        bool ok = true;
        int base = 0;
#pragma unroll
        for (int a = 0; a < DIM; ++a) {
            ax[a] = axis_setup_t<CT>(TT::ld(gp + a), p.size[a], off, p, align);
            ok = ok && ax[a].ok;
            base += ax[a].l * p.tstride[a];
        }
        CT gog[DIM], gogg[DIM];
#pragma unroll
        for (int a = 0; a < DIM; ++a) {
            gog[a] = (STAGE >= ST_BB) ? TT::ld(p.gog + ((long long)n * p.P + pi) * DIM + a) : CT(0);
            gogg[a] = (STAGE == ST_BBB) ? TT::ld(p.gogg + ((long long)n * p.P + pi) * DIM + a) : CT(0);
        }
        // per-corner coefficients
        CT w[NCORN], A[NCORN], E[NCORN], D1[NCORN][DIM], H[NCORN][DIM];
        bool valid[NCORN];
        int texel[NCORN];
#pragma unroll
        for (int q = 0; q < NCORN; ++q) {
            int b[DIM];
            bool v = ok;
            int t = base;
#pragma unroll
            for (int a = 0; a < DIM; ++a) {
                b[a] = (q >> a) & 1;
                v = v && ax[a].inb[b[a]];
                t += b[a] * p.tstride[a];
            }
            valid[q] = v;
            texel[q] = t;
            CT wall = CT(1);
#pragma unroll
            for (int a = 0; a < DIM; ++a) wall *= ax[a].w[b[a]];
            w[q] = wall;
            A[q] = CT(0); E[q] = CT(0);
#pragma unroll
            for (int a = 0; a < DIM; ++a) {
                CT wo = CT(1);                       // prod_{c != a} W_c
#pragma unroll
                for (int c = 0; c < DIM; ++c) if (c != a) wo *= ax[c].w[b[c]];
                const CT d1 = ax[a].dw[b[a]] * wo;
                const CT d2 = ax[a].ew[b[a]] * wo;
                D1[q][a] = d1;
                A[q] += d1 * gog[a];
                E[q] += d2 * gogg[a] * gog[a];
                // BB gGrid coefficient of axis a: 2D pure second derivative only (cu2d:675-678,705-706); 3D full
                // Hessian row (cu3d:836-856)
                CT h = d2 * gog[a];
                if (DIM == 3) {
#pragma unroll
                    for (int c = 0; c < DIM; ++c) {
                        if (c == a) continue;
                        CT wr = CT(1);               // the remaining axis
#pragma unroll
                        for (int e = 0; e < DIM; ++e) if (e != a && e != c) wr *= ax[e].w[b[e]];
                        h += ax[a].dw[b[a]] * ax[c].dw[b[c]] * wr * gog[c];
                    }
                }
                H[q][a] = h;
            }
        }
        CT gg[DIM];
#pragma unroll
        for (int a = 0; a < DIM; ++a) gg[a] = CT(0);
        const T* Vn = p.V ? p.V + (long long)n * p.C * p.T : nullptr;
        const T* Un = p.U ? p.U + (long long)n * p.C * p.T : nullptr;
        typename TT::AT* An = p.acc ? p.acc + (long long)n * p.C * p.T : nullptr;
        for (int c = 0; c < p.C; ++c) {
            const CT go = (STAGE != ST_F) ? TT::ld(p.x1 + n * p.x1_sn + (long long)c * p.x1_sc + pi) : CT(0);
            const CT go2 = (STAGE == ST_BBB && p.x2) ? TT::ld(p.x2 + n * p.x2_sn + (long long)c * p.x2_sc + pi) : CT(0);
            CT out = CT(0);
#pragma unroll
            for (int q = 0; q < NCORN; ++q) {
                if (!valid[q]) continue;
                const long long o = (long long)c * p.T + texel[q];
                const CT v = Vn ? TT::ld(Vn + o) : CT(0);
                if (STAGE == ST_F) {
                    out += v * w[q];
                } else if (STAGE == ST_B) {
                    if (An) atomicAdd(An + o, w[q] * go);
                    if (p.ggrid) {
#pragma unroll
                        for (int a = 0; a < DIM; ++a) gg[a] += go * v * D1[q][a];
                    }
                } else if (STAGE == ST_BB) {
                    out += v * A[q];
                    const CT u = Un ? TT::ld(Un + o) : CT(0);
                    if (Un) out += u * w[q];
                    if (An) atomicAdd(An + o, go * A[q]);
                    if (p.ggrid) {
#pragma unroll
                        for (int a = 0; a < DIM; ++a) {
                            gg[a] += go * v * H[q][a];
                            if (DIM == 3 && Un) gg[a] += go * u * D1[q][a];      // cu3d:848-856
                        }
                    }
                } else {
                    out += v * E[q];
                    if (An) atomicAdd(An + o, go * E[q] + go2 * A[q]);
                }
            }
            if (p.y && STAGE != ST_B) TT::st(p.y + ((long long)n * p.C + c) * p.P + pi, out);
        }
        if (p.ggrid && (STAGE == ST_B || STAGE == ST_BB)) {
#pragma unroll
            for (int a = 0; a < DIM; ++a) TT::st(p.ggrid + ((long long)n * p.P + pi) * DIM + a, gg[a]);
        }
    }
}

template <int DIM, typename T>
cudaError_t launch_scalar(int stage, const ScalarParamsT<T>& p, cudaStream_t s) {
    const long long total = (long long)p.N * p.P;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) return cudaSuccess;
    switch (stage) {
        case ST_F: cs_scalar_stage_kernel<DIM, ST_F, T><<<(unsigned)blocks, 256, 0, s>>>(p); break;
        case ST_B: cs_scalar_stage_kernel<DIM, ST_B, T><<<(unsigned)blocks, 256, 0, s>>>(p); break;
        case ST_BB: cs_scalar_stage_kernel<DIM, ST_BB, T><<<(unsigned)blocks, 256, 0, s>>>(p); break;
        default: cs_scalar_stage_kernel<DIM, ST_BBB, T><<<(unsigned)blocks, 256, 0, s>>>(p); break;
    }
    return cudaGetLastError();
}

// gInput of the fp16 path: the float workspace rounded to half
static __global__ void __launch_bounds__(256) cs_f32_to_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst,
                                                                   long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = __float2half_rn(src[i]);
}

}  // namespace cs
