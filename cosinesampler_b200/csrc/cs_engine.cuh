// cs_engine.cuh -- the gather / contract / scatter engine behind every stage.
//
// One kernel template serves forward (F), backward (B), double backward (BB)
// and triple backward (BBB) in 2D and 3D.  What the stages share (reference:
// cosine_sampler_2d_kernel.cu:265-891, cosine_sampler_3d_kernel.cu:250-1071):
//
//   * a (cell n, point p) pair touches 2^dim corner texels of cell n;
//   * per pair and corner there is a handful of scalar coefficients that depend
//     only on the point's fractional position (products of per-axis kernel
//     values and derivatives);
//   * channel-wise work is: gather V (and optionally U = gOutInput) at the
//     corners, read one or two [N,C,P] point streams, write one point stream,
//     reduce over channels into a per-point gradient, and scatter-add into a
//     grid-shaped accumulator.
//
// B200 mapping (nothing here is translated from the reference, which runs one
// thread per pair with a serial channel loop over NCHW):
//
//   * a warp owns a *tile* of 4*(32/L) consecutive points of one cell; L lanes
//     (1,2,4,8) cooperate on a quad of 4 consecutive points, each lane owning
//     VEC channels at a time.  With the channel-last field layout and VEC=4 a
//     corner gather is one 16-byte load per lane and the L lanes of a quad read
//     one contiguous 16*L-byte segment; a scatter is one red.global.add.v4.f32
//     per lane.
//   * because a lane holds 4 consecutive points of a channel, every access to a
//     [N,C,P] stream is a 16-byte access and the 32/L quads of a warp cover
//     contiguous 128*(4/L)-byte runs per channel: streams stay fully coalesced
//     in their reference layout.
//   * phase 1 of a tile computes the per-point record (index map, padding,
//     kernel values) once per point, one point per lane, into shared memory;
//     phase 2 reads it back as broadcasts.  The transcendental work is thus
//     never replicated across the lanes of a quad.
//   * persistent grid: blocks loop over tiles, cell index fastest so that the N
//     cells of one point (which share coordinates and an expanded gOut in
//     PIXEL) are in flight together.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cs {

enum Stage { ST_F = 0, ST_B = 1, ST_BB = 2, ST_BBB = 3 };

struct StageParams {
    // geometry
    int N, C;
    int size[3];         // extent per axis a (0:W 1:H 2:D)
    int tstride[3];      // texel stride per axis, in texels
    long long P;
    long long num_ptiles;  // tiles per cell
    int lshift;          // log2(lanes per quad)
    // grid-shaped fields (same layout for V, U, acc)
    long long cell_stride;  // elements between cells
    int texel_stride;       // elements between texels (C channel-last, 1 channel-first)
    int chan_stride;        // elements between channels (1 channel-last, T channel-first)
    const float* V;         // input
    const float* U;         // gOutInput (BB) or nullptr
    float* acc;             // gInput accumulator or nullptr
    // point streams
    const float* x1; long long x1_sn, x1_sc;   // gOut
    const float* x2; long long x2_sn, x2_sc;   // gOutggOut (BBB fused b_input) or nullptr
    float* y;                                  // out / ggOut (contiguous [N,C,P]) or nullptr
    int svec4;                                 // streams may be accessed as float4
    // per-point arrays
    const float* grid; long long grid_sn;      // [N,P,dim]
    int grid_vec2;                             // 2D coordinates may be loaded as float2
    const float* gog;                          // gOutGrid  [N,P,dim]
    const float* gogg;                         // gOutgGrid [N,P,dim]
    float* ggrid;                              // gGrid [N,P,dim] or nullptr
    const float* offset;                       // [N]
    // modes
    int pad, align, kernel, multicell, index_mode;
};

// ---------------------------------------------------------------------------
// Coordinate map, padding and kernel functions.  Semantics follow
// cu2d:53-261 (grid_sampler_unnormalize[_set_grad], clip/reflect_coordinates
// [_set_grad], cosine/smoothstep and their derivatives); the arithmetic is
// spelled out with explicit roundings so that cell decisions are reproducible.
// ---------------------------------------------------------------------------
struct AxisRec {
    int l;        // low corner index
    float w0;     // weight of the low corner  k(r)
    float w1;     // weight of the high corner 1-k(r)  (linear: i - l)
    float d;      // m * k'(r)      (d weight / d normalised coord: -d low, +d high)
    float e;      // m^2 * k''(r)   (+e low, -e high)
    bool ok;      // coordinate is finite and in a sane range
};

__device__ __forceinline__ float clip_grad(float& i, int size) {
    const float hi = (float)(size - 1);
    if (i <= 0.f) { i = 0.f; return 0.f; }
    if (i >= hi) { i = hi; return 0.f; }
    return 1.f;
}

__device__ __forceinline__ float reflect_grad(float& i, int twice_low, int twice_high) {
    if (twice_low == twice_high) { i = 0.f; return 0.f; }
    const float lo = (float)twice_low * 0.5f;
    const float span = (float)(twice_high - twice_low) * 0.5f;
    float x = i - lo;
    float sign = 1.f;
    if (x < 0.f) { sign = -1.f; x = -x; }
    const float extra = fmodf(x, span);
    const int flips = (int)floorf(x / span);
    if ((flips & 1) == 0) { i = extra + lo; return sign; }
    i = span - extra + lo;
    return -sign;
}

__device__ __forceinline__ AxisRec axis_setup(float g, int size, float off, const StageParams& p,
                                              bool align, int order) {
    AxisRec a;
    float i, m;
    if (align) {
        const int s = size - 1 - (p.multicell ? 1 : 0);           // cu2d:56-61
        const float sf = (float)s;
        m = sf * 0.5f;                                            // cu2d:80
        const float h = __fmul_rn(__fadd_rn(g, 1.f), 0.5f);
        i = (p.index_mode == 0) ? __fadd_rn(__fmul_rn(h, sf), off) : __fmaf_rn(h, sf, off);
    } else {
        const float sf = (float)size;
        m = sf * 0.5f;                                            // cu2d:84
        if (p.index_mode == 0) {
            i = __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(g, 1.f), sf), -1.f), 0.5f), off);
        } else {
            i = __fmaf_rn(__fmaf_rn(__fadd_rn(g, 1.f), sf, -1.f), 0.5f, off);
        }
    }
    a.ok = fabsf(i) < 1.0e9f;                                     // rejects NaN / inf / absurd
    if (!a.ok) i = 0.f;
    if (p.pad == 1) {                                             // border, cu2d:220-223
        m *= clip_grad(i, size);
    } else if (p.pad == 2) {                                      // reflection, cu2d:224-233
        const float gr = align ? reflect_grad(i, 0, 2 * (size - 2))
                               : reflect_grad(i, -1, 2 * size - 1);
        const float gc = clip_grad(i, size);
        m *= gr * gc;
    }
    const float lf = floorf(i);
    a.l = (int)lf;
    const float r = __fsub_rn(lf + 1.f, i);
    if (p.kernel == 0) {                                          // cosine, cu2d:251-261
        float sn, cn;
        sincospif(r, &sn, &cn);
        a.w0 = 0.5f * (1.f - cn);
        a.w1 = 1.f - a.w0;
        if (order >= 1) a.d = (0.5f * 3.14159265358979323846f) * sn * m;
        if (order >= 2) a.e = (0.5f * 9.86960440108935861883f) * cn * (m * m);
    } else if (p.kernel == 2) {                                   // smoothstep, cu2d:239-249
        a.w0 = r * r * (3.f - 2.f * r);
        a.w1 = 1.f - a.w0;
        if (order >= 1) a.d = 6.f * r * (1.f - r) * m;
        if (order >= 2) a.e = (6.f - 12.f * r) * (m * m);
    } else {                                                      // linear
        a.w0 = r;
        a.w1 = __fsub_rn(i, lf);      // ATen's (ix - ix_nw): keeps F bit-equal to grid_sample
        if (order >= 1) a.d = m;
        if (order >= 2) a.e = 0.f;
    }
    return a;
}

// ---------------------------------------------------------------------------
// Small memory helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_f32(float* addr, float a) {
    asm volatile("red.global.add.f32 [%0], %1;" :: "l"(addr), "f"(a) : "memory");
}

// 4 consecutive points of one channel of a [N,C,P] stream
__device__ __forceinline__ void stream_load4(float (&v)[4], const float* base, long long p0,
                                             long long P, bool vec) {
    if (vec) {
        if (p0 < P) {
            const float4 t = __ldcs(reinterpret_cast<const float4*>(base + p0));
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else { v[0] = v[1] = v[2] = v[3] = 0.f; }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = (p0 + i < P) ? __ldcs(base + p0 + i) : 0.f;
    }
}
__device__ __forceinline__ void stream_store4(const float (&v)[4], float* base, long long p0,
                                              long long P, bool vec) {
    if (vec) {
        if (p0 < P) __stcs(reinterpret_cast<float4*>(base + p0), make_float4(v[0], v[1], v[2], v[3]));
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) if (p0 + i < P) __stcs(base + p0 + i, v[i]);
    }
}

template <int VEC> struct FieldVec;
template <> struct FieldVec<4> {
    float v[4];
    __device__ __forceinline__ void load(const float* p, int /*chan_stride*/) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    __device__ __forceinline__ static void red(float* p, const float (&a)[4]) {
        red_add_v4(p, a[0], a[1], a[2], a[3]);
    }
};
template <> struct FieldVec<1> {
    float v[1];
    __device__ __forceinline__ void load(const float* p, int) { v[0] = __ldg(p); }
    __device__ __forceinline__ static void red(float* p, const float (&a)[1]) { red_add_f32(p, a[0]); }
};

// record layout in shared memory: SoA, rec[field * pts + point]
//   field 0: base texel (int bits)   field 1: corner-valid mask (int bits)
//   then per axis a: w0, w1 [, d [, e]] and stage extras
template <int DIM, int STAGE, bool HAS_X2> struct RecLayout {
    // per-axis fields
    static constexpr int W0 = 0, W1 = 1, D1 = 2, E2 = 3, G1 = 4, G2 = 5;
    // number of per-axis fields by stage
    //  F: w0 w1                     B: w0 w1 d
    //  BB: w0 w1 d e gog            BBB: w0 w1 e*gog*gogg [d*gog]
    static constexpr int PER_AXIS = (STAGE == ST_F) ? 2 : (STAGE == ST_B) ? 3 :
                                    (STAGE == ST_BB) ? 5 : (HAS_X2 ? 4 : 3);
    static constexpr int FIELDS = 2 + DIM * PER_AXIS;
};

// ---------------------------------------------------------------------------
// The stage kernel
// ---------------------------------------------------------------------------
template <int DIM, int VEC, int STAGE, bool HAS_U, bool HAS_X2>
__global__ void __launch_bounds__(256)
cs_stage_kernel(const StageParams p) {
    using RL = RecLayout<DIM, STAGE, HAS_X2>;
    constexpr int NCORN = 1 << DIM;
    constexpr int PA = RL::PER_AXIS;
    constexpr bool HAS_X1 = (STAGE != ST_F);

    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int L = 1 << p.lshift;
    const int pts = 128 >> p.lshift;               // points per tile
    const int q = lane >> p.lshift;                // quad within the tile
    const int j = lane & (L - 1);                  // lane within the quad
    float* rec = smem + (size_t)warp * RL::FIELDS * pts;

    const bool want_y = (p.y != nullptr);
    const bool want_g = (p.ggrid != nullptr);
    const bool want_s = (p.acc != nullptr);
    const bool need_v = want_y || want_g;
    const bool svec = p.svec4 != 0;
    const int V = p.C / VEC;                       // channel vectors per texel
    // 2D forward always maps with align_corners = 1 (cu2d:307-308)
    const bool align = (STAGE == ST_F && DIM == 2) ? true : (p.align != 0);
    constexpr int ORDER = (STAGE == ST_F) ? 0 : (STAGE == ST_B) ? 1 : 2;

    const long long total = p.num_ptiles * p.N;
    for (long long tile = (long long)blockIdx.x * wpb + warp; tile < total;
         tile += (long long)gridDim.x * wpb) {
        const int n = (int)(tile % p.N);
        const long long pt0 = (tile / p.N) * pts;

        // ---------------- phase 1: one point per lane -> record -------------
        const float off = __ldg(p.offset + n);
        for (int i = lane; i < pts; i += 32) {
            const long long pi = pt0 + i;
            int base = 0, mask = 0;
            if (pi < p.P) {
                const float* gp = p.grid + (long long)n * p.grid_sn + pi * DIM;
                float g[DIM];
                if (DIM == 2 && p.grid_vec2) {
                    const float2 t = __ldg(reinterpret_cast<const float2*>(gp));
                    g[0] = t.x; g[1] = t.y;
                } else {
#pragma unroll
                    for (int a = 0; a < DIM; ++a) g[a] = __ldg(gp + a);
                }
                float gog[DIM], gogg[DIM];
                if (STAGE >= ST_BB) {
                    const float* s = p.gog + ((long long)n * p.P + pi) * DIM;
#pragma unroll
                    for (int a = 0; a < DIM; ++a) gog[a] = __ldg(s + a);
                }
                if (STAGE == ST_BBB) {
                    const float* s = p.gogg + ((long long)n * p.P + pi) * DIM;
#pragma unroll
                    for (int a = 0; a < DIM; ++a) gogg[a] = __ldg(s + a);
                }
                bool ok = true;
                int lo_ok[DIM], hi_ok[DIM];
#pragma unroll
                for (int a = 0; a < DIM; ++a) {
                    const AxisRec ar = axis_setup(g[a], p.size[a], off, p, align, ORDER);
                    ok = ok && ar.ok;
                    base += ar.l * p.tstride[a];
                    lo_ok[a] = (ar.l >= 0) && (ar.l < p.size[a]);
                    hi_ok[a] = (ar.l + 1 >= 0) && (ar.l + 1 < p.size[a]);
                    float* ra = rec + (2 + a * PA) * pts + i;
                    ra[RL::W0 * pts] = ar.w0;
                    ra[RL::W1 * pts] = ar.w1;
                    if (STAGE == ST_B) ra[2 * pts] = ar.d;
                    if (STAGE == ST_BB) {
                        ra[2 * pts] = ar.d;
                        ra[3 * pts] = ar.e;
                        ra[4 * pts] = gog[a];
                    }
                    if (STAGE == ST_BBB) {
                        ra[2 * pts] = ar.e * gogg[a] * gog[a];
                        if (HAS_X2) ra[3 * pts] = ar.d * gog[a];
                    }
                }
                if (ok) {
#pragma unroll
                    for (int c = 0; c < NCORN; ++c) {
                        bool v = true;
#pragma unroll
                        for (int a = 0; a < DIM; ++a) v = v && (((c >> a) & 1) ? hi_ok[a] : lo_ok[a]);
                        mask |= (v ? 1 : 0) << c;
                    }
                }
            }
            rec[0 * pts + i] = __int_as_float(base);
            rec[1 * pts + i] = __int_as_float(mask);
        }
        __syncwarp();

        // ---------------- phase 2: L lanes per quad, VEC channels per lane ---
        const long long qp0 = pt0 + 4 * q;          // first point of this lane's quad
        const float* Vn = p.V + (long long)n * p.cell_stride;
        const float* Un = HAS_U ? p.U + (long long)n * p.cell_stride : nullptr;
        float* An = want_s ? p.acc + (long long)n * p.cell_stride : nullptr;

        float gg[4][DIM];
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int a = 0; a < DIM; ++a) gg[t][a] = 0.f;

        for (int jj = j; jj < V; jj += L) {
            const int chan0 = jj * VEC;
            float x1[VEC][4], x2[VEC][4], y[VEC][4];
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                if (HAS_X1) stream_load4(x1[k], p.x1 + n * p.x1_sn + (long long)(chan0 + k) * p.x1_sc, qp0, p.P, svec);
                if (HAS_X2) stream_load4(x2[k], p.x2 + n * p.x2_sn + (long long)(chan0 + k) * p.x2_sc, qp0, p.P, svec);
#pragma unroll
                for (int t = 0; t < 4; ++t) y[k][t] = 0.f;
            }
            const int foff = chan0 * p.chan_stride;

#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int ri = 4 * q + t;
                const int base = __float_as_int(rec[0 * pts + ri]);
                const int mask = __float_as_int(rec[1 * pts + ri]);
                if (mask == 0) continue;
                float w[DIM][2], dw[DIM][2], ew[DIM][2], gog[DIM], fx[DIM], hx[DIM];
#pragma unroll
                for (int a = 0; a < DIM; ++a) {
                    const float* ra = rec + (2 + a * PA) * pts + ri;
                    w[a][0] = ra[0];
                    w[a][1] = ra[pts];
                    if (STAGE == ST_B || STAGE == ST_BB) { dw[a][1] = ra[2 * pts]; dw[a][0] = -dw[a][1]; }
                    if (STAGE == ST_BB) { ew[a][0] = ra[3 * pts]; ew[a][1] = -ew[a][0]; gog[a] = ra[4 * pts]; }
                    if (STAGE == ST_BBB) {
                        fx[a] = ra[2 * pts];                 // e * gogg * gog (low corner sign)
                        if (HAS_X2) hx[a] = ra[3 * pts];     // d * gog        (high corner sign)
                    }
                }
#pragma unroll
                for (int c = 0; c < NCORN; ++c) {
                    if (!((mask >> c) & 1)) continue;
                    int b[DIM];
                    int texel = base;
#pragma unroll
                    for (int a = 0; a < DIM; ++a) { b[a] = (c >> a) & 1; texel += b[a] * p.tstride[a]; }
                    // products of the other axes' weights
                    float wo[DIM];          // prod_{c != a} W_c
                    float wall;             // prod_a W_a, in x,y,z order
                    if (DIM == 2) {
                        wo[0] = w[1][b[1]]; wo[1] = w[0][b[0]];
                        wall = w[0][b[0]] * w[1][b[1]];
                    } else {
                        wo[0] = w[1][b[1]] * w[2][b[2]];
                        wo[1] = w[0][b[0]] * w[2][b[2]];
                        wo[2] = w[0][b[0]] * w[1][b[1]];
                        wall = (w[0][b[0]] * w[1][b[1]]) * w[2][b[2]];
                    }
                    // stage coefficients for this (point, corner)
                    float cy = 0.f;        // y      += V * cy
                    float cu = 0.f;        // y      += U * cu
                    float cs1 = 0.f;       // acc    += x1 * cs1
                    float cs2 = 0.f;       // acc    += x2 * cs2
                    float cg[DIM];         // gg[a]  += dot(V,x1) * cg[a]
                    float cgu[DIM];        // gg[a]  += dot(U,x1) * cgu[a]
                    if (STAGE == ST_F) {
                        cy = wall;
                    } else if (STAGE == ST_B) {
                        cs1 = wall;
#pragma unroll
                        for (int a = 0; a < DIM; ++a) cg[a] = dw[a][b[a]] * wo[a];
                    } else if (STAGE == ST_BB) {
                        float A = 0.f;
#pragma unroll
                        for (int a = 0; a < DIM; ++a) A += dw[a][b[a]] * wo[a] * gog[a];
                        cy = A; cs1 = A; cu = wall;
                        if (DIM == 2) {
                            // pure second derivatives only (cu2d:705-706)
#pragma unroll
                            for (int a = 0; a < DIM; ++a) cg[a] = ew[a][b[a]] * wo[a] * gog[a];
                        } else {
                            // full Hessian row + gOutInput term (cu3d:836-856)
#pragma unroll
                            for (int a = 0; a < DIM; ++a) {
                                float s = ew[a][b[a]] * wo[a] * gog[a];
#pragma unroll
                                for (int bb = 0; bb < DIM; ++bb) {
                                    if (bb == a) continue;
                                    const int cc = 3 - a - bb;
                                    s += dw[a][b[a]] * dw[bb][b[bb]] * w[cc][b[cc]] * gog[bb];
                                }
                                cg[a] = s;
                                cgu[a] = dw[a][b[a]] * wo[a];
                            }
                        }
                    } else {  // BBB: pure second derivatives (cu2d:876-885, cu3d:1054-1065)
                        float E = 0.f, A = 0.f;
#pragma unroll
                        for (int a = 0; a < DIM; ++a) {
                            E += (b[a] ? -fx[a] : fx[a]) * wo[a];
                            if (HAS_X2) A += (b[a] ? hx[a] : -hx[a]) * wo[a];
                        }
                        cy = E; cs1 = E; cs2 = A;
                    }

                    const long long fo = (long long)texel * p.texel_stride + foff;
                    FieldVec<VEC> vv, uu;
                    if (need_v) {
                        vv.load(Vn + fo, p.chan_stride);
                        if (want_y) {
#pragma unroll
                            for (int k = 0; k < VEC; ++k) y[k][t] = fmaf(vv.v[k], cy, y[k][t]);
                        }
                    }
                    if (HAS_U) {
                        uu.load(Un + fo, p.chan_stride);
                        if (want_y) {
#pragma unroll
                            for (int k = 0; k < VEC; ++k) y[k][t] = fmaf(uu.v[k], cu, y[k][t]);
                        }
                    }
                    if (HAS_X1 && STAGE != ST_BBB && want_g) {
                        float dot = 0.f;
#pragma unroll
                        for (int k = 0; k < VEC; ++k) dot = fmaf(vv.v[k], x1[k][t], dot);
#pragma unroll
                        for (int a = 0; a < DIM; ++a) gg[t][a] = fmaf(dot, cg[a], gg[t][a]);
                        if (HAS_U && DIM == 3) {
                            float du = 0.f;
#pragma unroll
                            for (int k = 0; k < VEC; ++k) du = fmaf(uu.v[k], x1[k][t], du);
#pragma unroll
                            for (int a = 0; a < DIM; ++a) gg[t][a] = fmaf(du, cgu[a], gg[t][a]);
                        }
                    }
                    if (HAS_X1 && want_s) {
                        float s[VEC];
#pragma unroll
                        for (int k = 0; k < VEC; ++k) {
                            s[k] = x1[k][t] * cs1;
                            if (HAS_X2) s[k] = fmaf(x2[k][t], cs2, s[k]);
                        }
                        FieldVec<VEC>::red(An + fo, s);
                    }
                }
            }
            if (want_y) {
#pragma unroll
                for (int k = 0; k < VEC; ++k)
                    stream_store4(y[k], p.y + ((long long)n * p.C + chan0 + k) * p.P, qp0, p.P, svec);
            }
        }

        if ((STAGE == ST_B || STAGE == ST_BB) && want_g) {
            // sum the channel-partial gradients over the L lanes of the quad
            for (int o = L >> 1; o > 0; o >>= 1) {
#pragma unroll
                for (int t = 0; t < 4; ++t)
#pragma unroll
                    for (int a = 0; a < DIM; ++a) gg[t][a] += __shfl_xor_sync(0xffffffffu, gg[t][a], o);
            }
            if (j == 0) {
                float* out = p.ggrid + ((long long)n * p.P + qp0) * DIM;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    if (qp0 + t < p.P) {
#pragma unroll
                        for (int a = 0; a < DIM; ++a) out[t * DIM + a] = gg[t][a];
                    }
                }
            }
        }
        __syncwarp();
    }
}

}  // namespace cs
