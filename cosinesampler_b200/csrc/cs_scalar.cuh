// cs_scalar.cuh -- double-precision path of the four stages.
//
// The reference dispatches float / double / half (AT_DISPATCH_FLOATING_TYPES_AND_HALF, cu2d:905) but its
// double and half instantiations cannot run: the per-cell offset is always a float tensor (modules_2d.py:25)
// and is read through TensorInfo<scalar_t> (cu2d:914).  This file provides what that dispatch promises for
// fp64: the same four entry points on double tensors (the offset stays fp32, as the reference builds it),
// semantics and quirks identical to the fp32 engine (cs_engine.cuh), computed entirely in
// double.  It is a correctness path -- one thread per (cell, point), serial channel loop over the reference's
// channel-first layout, scalar atomicAdd(double) -- for callers who validate a PDE residual in fp64; it makes no
// throughput claim.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "cs_engine.cuh"

namespace cs {

struct ScalarParams {
    int N, C;
    int size[3];
    int tstride[3];
    long long T;                 // texels per cell
    long long P;
    const double* V;             // input        [N, C, T]
    const double* U;             // gOutInput    [N, C, T] or nullptr            (BB)
    double* acc;                 // gInput       [N, C, T] or nullptr, zero-initialised
    const double* x1; long long x1_sn, x1_sc;     // gOut       strided [N, C, P]
    const double* x2; long long x2_sn, x2_sc;     // gOutggOut  (BBB, fused b_input pass) or nullptr
    double* y;                   // out / ggOut  [N, C, P] or nullptr
    const double* grid; long long grid_sn;        // [N, P, dim]
    const double* gog;           // gOutGrid     [N, P, dim]
    const double* gogg;          // gOutgGrid    [N, P, dim]
    double* ggrid;               // gGrid        [N, P, dim] or nullptr
    const float* offset;         // [N]  fp32, as modules_2d.py:24-27 builds it
    int pad, align, kernel, multicell;
};

struct AxisD {
    int l;
    double w[2], dw[2], ew[2];   // weight, d weight / d coordinate, d2 weight / d coordinate^2 for (low, high)
    bool ok, inb[2];
};

__device__ __forceinline__ double clip_grad_d(double& i, int size) {
    const double hi = (double)(size - 1);
    if (i <= 0.0) { i = 0.0; return 0.0; }
    if (i >= hi) { i = hi; return 0.0; }
    return 1.0;
}
__device__ __forceinline__ double reflect_grad_d(double& i, int twice_low, int twice_high) {
    if (twice_low == twice_high) { i = 0.0; return 0.0; }
    const double lo = (double)twice_low * 0.5;
    const double span = (double)(twice_high - twice_low) * 0.5;
    double x = i - lo, sign = 1.0;
    if (x < 0.0) { sign = -1.0; x = -x; }
    const double extra = fmod(x, span);
    const long long flips = (long long)floor(x / span);
    if ((flips & 1) == 0) { i = extra + lo; return sign; }
    i = span - extra + lo;
    return -sign;
}

// cu2d:53-261 in double (index map, padding, kernel functions and their derivatives)
__device__ __forceinline__ AxisD axis_setup_d(double g, int size, double off, const ScalarParams& p, bool align) {
    AxisD a;
    double i, m;
    if (align) {
        const double s = (double)(size - 1 - (p.multicell ? 1 : 0));
        m = s * 0.5;
        i = ((g + 1.0) * 0.5) * s + off;
    } else {
        const double s = (double)size;
        m = s * 0.5;
        i = ((g + 1.0) * s - 1.0) * 0.5 + off;
    }
    a.ok = (p.pad == 1) ? (fabs(i) <= 1.0e300) : (fabs(i) < 1.0e9);
    if (!a.ok) i = 0.0;
    if (p.pad == 1) {
        m *= clip_grad_d(i, size);
    } else if (p.pad == 2) {
        const double gr = align ? reflect_grad_d(i, 0, 2 * (size - 2)) : reflect_grad_d(i, -1, 2 * size - 1);
        m *= gr * clip_grad_d(i, size);
    }
    const double lf = floor(i);
    a.l = (int)lf;
    const double r = (lf + 1.0) - i;
    double k0, k1, k2, whi;
    if (p.kernel == 0) {
        double sn, cn;
        sincospi(r, &sn, &cn);
        k0 = 0.5 * (1.0 - cn);
        k1 = 0.5 * 3.14159265358979323846 * sn;
        k2 = 0.5 * 9.86960440108935861883 * cn;
        whi = 1.0 - k0;
    } else if (p.kernel == 2) {
        k0 = r * r * (3.0 - 2.0 * r);
        k1 = 6.0 * r * (1.0 - r);
        k2 = 6.0 - 12.0 * r;
        whi = 1.0 - k0;
    } else {
        k0 = r; k1 = 1.0; k2 = 0.0;
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
template <int DIM, int STAGE>
__global__ void __launch_bounds__(256) cs_scalar_stage_kernel(const ScalarParams p) {
    constexpr int NCORN = 1 << DIM;
    const long long total = (long long)p.N * p.P;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(idx / p.P);
        const long long pi = idx - (long long)n * p.P;
        const double off = (double)__ldg(p.offset + n);
        const double* gp = p.grid + (long long)n * p.grid_sn + pi * DIM;
        // the 2D forward ignores align_corners and uses 1 (cu2d:307-308)
        const bool align = (STAGE == ST_F && DIM == 2) ? true : (p.align != 0);
        AxisD ax[DIM];
        bool ok = true;
        int base = 0;
#pragma unroll
        for (int a = 0; a < DIM; ++a) {
            ax[a] = axis_setup_d(gp[a], p.size[a], off, p, align);
            ok = ok && ax[a].ok;
            base += ax[a].l * p.tstride[a];
        }
        double gog[DIM], gogg[DIM];
#pragma unroll
        for (int a = 0; a < DIM; ++a) {
            gog[a] = (STAGE >= ST_BB) ? p.gog[((long long)n * p.P + pi) * DIM + a] : 0.0;
            gogg[a] = (STAGE == ST_BBB) ? p.gogg[((long long)n * p.P + pi) * DIM + a] : 0.0;
        }
        // per-corner coefficients
        double w[NCORN], A[NCORN], E[NCORN], D1[NCORN][DIM], H[NCORN][DIM];
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
            double wall = 1.0;
#pragma unroll
            for (int a = 0; a < DIM; ++a) wall *= ax[a].w[b[a]];
            w[q] = wall;
            A[q] = 0.0; E[q] = 0.0;
#pragma unroll
            for (int a = 0; a < DIM; ++a) {
                double wo = 1.0;                     // prod_{c != a} W_c
#pragma unroll
                for (int c = 0; c < DIM; ++c) if (c != a) wo *= ax[c].w[b[c]];
                const double d1 = ax[a].dw[b[a]] * wo;
                const double d2 = ax[a].ew[b[a]] * wo;
                D1[q][a] = d1;
                A[q] += d1 * gog[a];
                E[q] += d2 * gogg[a] * gog[a];
                // BB gGrid coefficient of axis a: 2D pure second derivative only (cu2d:675-678,705-706); 3D full
                // Hessian row (cu3d:836-856)
                double h = d2 * gog[a];
                if (DIM == 3) {
#pragma unroll
                    for (int c = 0; c < DIM; ++c) {
                        if (c == a) continue;
                        double wr = 1.0;             // the remaining axis
#pragma unroll
                        for (int e = 0; e < DIM; ++e) if (e != a && e != c) wr *= ax[e].w[b[e]];
                        h += ax[a].dw[b[a]] * ax[c].dw[b[c]] * wr * gog[c];
                    }
                }
                H[q][a] = h;
            }
        }
        double gg[DIM];
#pragma unroll
        for (int a = 0; a < DIM; ++a) gg[a] = 0.0;
        const double* Vn = p.V ? p.V + (long long)n * p.C * p.T : nullptr;
        const double* Un = p.U ? p.U + (long long)n * p.C * p.T : nullptr;
        double* An = p.acc ? p.acc + (long long)n * p.C * p.T : nullptr;
        for (int c = 0; c < p.C; ++c) {
            const double go = (STAGE != ST_F) ? p.x1[n * p.x1_sn + (long long)c * p.x1_sc + pi] : 0.0;
            const double go2 = (STAGE == ST_BBB && p.x2) ? p.x2[n * p.x2_sn + (long long)c * p.x2_sc + pi] : 0.0;
            double out = 0.0;
#pragma unroll
            for (int q = 0; q < NCORN; ++q) {
                if (!valid[q]) continue;
                const long long o = (long long)c * p.T + texel[q];
                const double v = Vn ? Vn[o] : 0.0;
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
                    const double u = Un ? Un[o] : 0.0;
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
            if (p.y && STAGE != ST_B) p.y[((long long)n * p.C + c) * p.P + pi] = out;
        }
        if (p.ggrid && (STAGE == ST_B || STAGE == ST_BB)) {
#pragma unroll
            for (int a = 0; a < DIM; ++a) p.ggrid[((long long)n * p.P + pi) * DIM + a] = gg[a];
        }
    }
}

template <int DIM>
cudaError_t launch_scalar(int stage, const ScalarParams& p, cudaStream_t s) {
    const long long total = (long long)p.N * p.P;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) return cudaSuccess;
    switch (stage) {
        case ST_F: cs_scalar_stage_kernel<DIM, ST_F><<<(unsigned)blocks, 256, 0, s>>>(p); break;
        case ST_B: cs_scalar_stage_kernel<DIM, ST_B><<<(unsigned)blocks, 256, 0, s>>>(p); break;
        case ST_BB: cs_scalar_stage_kernel<DIM, ST_BB><<<(unsigned)blocks, 256, 0, s>>>(p); break;
        default: cs_scalar_stage_kernel<DIM, ST_BBB><<<(unsigned)blocks, 256, 0, s>>>(p); break;
    }
    return cudaGetLastError();
}

}  // namespace cs
