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
//   * phase 1 of a tile computes, once per point (one point per lane), the index
//     map, padding, kernel values AND the finished per-corner coefficients of
//     the stage into a shared-memory record; phase 2 reads them back as
//     128-bit broadcasts (one float4 = the 4 corners of a coefficient).  Neither
//     the transcendental work nor the coefficient products are replicated over
//     the L lanes of a quad, and phase 2 is nothing but address + gather + fma.
//   * the random corner gathers run through a per-warp cp.async ring (one stage
//     ahead, data in flight in shared memory), the streams are prefetched one
//     item ahead with plain loads; warps never synchronise with each other.
//   * persistent grid: blocks loop over tiles, cell index fastest so that the N
//     cells of one point (which share coordinates and an expanded gOut in
//     PIXEL) are in flight together -- or cell by cell when the fields touched
//     at random would not fit in L2 together (StageParams::cell_major).
//   * cs_small_kernel (end of this file) is the variant for cells that fit in
//     shared memory.
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
    int small_cell;      // 0 auto, 1 never, 2 whenever it fits: shared-memory small-cell kernel
    int cell_major;      // tile order: 1 = all point tiles of cell 0, then cell 1, ... (bounds the L2 working set
                         // to one cell's fields); 0 = cell index fastest (cells of a point run together)
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
    int gvec4;                                 // gGrid quads may be stored as float4
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

template <class ParamsT>
__device__ __forceinline__ AxisRec axis_setup(float g, int size, float off, const ParamsT& p,
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
    // NaN / inf are rejected (the reference's floor -> int cast is undefined there, cu2d:310-311).
    // Huge finite indices are out of range for zeros padding (contribute nothing, like the
    // reference), clip to the border for border padding, and are rejected for reflection (whose
    // flip count overflows in the reference as well).
    a.ok = (p.pad == 1) ? (fabsf(i) <= 3.0e38f) : (fabsf(i) < 1.0e9f);
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

__device__ __forceinline__ float f4get(const float4& v, int i) {
    return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w;
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
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
            const float x = (p0 + i < P) ? __ldcs(base + p0 + i) : 0.f;
            if (i == 0) t.x = x; else if (i == 1) t.y = x; else if (i == 2) t.z = x; else t.w = x;
        }
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
}
__device__ __forceinline__ void stream_store4(const float (&v)[4], float* base, long long p0,
                                              long long P, bool vec) {
    if (vec) {
        if (p0 < P) __stcs(reinterpret_cast<float4*>(base + p0), make_float4(v[0], v[1], v[2], v[3]));
    } else {
        // ragged / unaligned streams: a real loop, so that the vector path above is not
        // burdened with four predicated-off scalar stores
        const float4 t = make_float4(v[0], v[1], v[2], v[3]);
#pragma unroll 1
        for (int i = 0; i < 4; ++i) if (p0 + i < P) __stcs(base + p0 + i, f4get(t, i));
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

// Record layout in shared memory: float4 fields, rec4[field * PTS + point].
//   field 0            : (base texel, corner-valid mask, -, -) as int bits
//   field 1 + k*CQ + h : coefficient k for corners 4h..4h+3   (CQ = corner quads = NCORN/4)
// Coefficient index k by stage:
//   F   : 0 = w_q
//   B   : 0 = w_q (scatter)            1+a = D_a,q (gGrid)
//   BB  : 0 = A_q (ggOut and scatter)  1+a = cg_a,q (gGrid)   [U: 1+DIM = w_q ; 3D: 2+DIM+a = D_a,q]
//   BBB : 0 = E_q (ggOut and scatter)  [X2: 1 = A_q]
template <int DIM, int STAGE, bool HAS_U, bool HAS_X2> struct RecLayout {
    static constexpr int NCORN = 1 << DIM;
    static constexpr int CQ = NCORN / 4;
    static constexpr int K = (STAGE == ST_F) ? 1 : (STAGE == ST_B) ? 1 + DIM :
                             (STAGE == ST_BB) ? (1 + DIM + (HAS_U ? 1 + (DIM == 3 ? DIM : 0) : 0))
                                              : (HAS_X2 ? 2 : 1);
    static constexpr int FIELDS4 = 1 + K * CQ;           // float4 fields per point
};

// Per-point inputs of phase 1, prefetched one tile ahead so that their latency hides
// behind phase 2 of the current tile.
template <int DIM, int STAGE> struct PointIn {
    float g[DIM];
    float gog[(STAGE >= ST_BB) ? DIM : 1];
    float gogg[(STAGE == ST_BBB) ? DIM : 1];
};

template <int DIM, int STAGE>
__device__ __forceinline__ void load_point(PointIn<DIM, STAGE>& in, const StageParams& p, int n,
                                           long long pi) {
    if (pi < p.P) {
        const float* gp = p.grid + (long long)n * p.grid_sn + pi * DIM;
        if (DIM == 2 && p.grid_vec2) {
            const float2 t = __ldg(reinterpret_cast<const float2*>(gp));
            in.g[0] = t.x; in.g[1] = t.y;
        } else {
#pragma unroll
            for (int a = 0; a < DIM; ++a) in.g[a] = __ldg(gp + a);
        }
        if (STAGE >= ST_BB) {
            const float* s = p.gog + ((long long)n * p.P + pi) * DIM;
#pragma unroll
            for (int a = 0; a < DIM; ++a) in.gog[a] = __ldg(s + a);
        }
        if (STAGE == ST_BBB) {
            const float* s = p.gogg + ((long long)n * p.P + pi) * DIM;
#pragma unroll
            for (int a = 0; a < DIM; ++a) in.gogg[a] = __ldg(s + a);
        }
    }
}

// Phase 1 for one point: index map -> per-corner coefficients -> shared-memory record.
template <int DIM, int STAGE, bool HAS_U, bool HAS_X2, int PTS>
__device__ __forceinline__ void build_record(float4* rec4, int i, const PointIn<DIM, STAGE>& in,
                                             bool in_range, float off, const StageParams& p, bool align) {
    using RL = RecLayout<DIM, STAGE, HAS_U, HAS_X2>;
    constexpr int NCORN = RL::NCORN;
    constexpr int CQ = RL::CQ;
    constexpr int K = RL::K;
    constexpr int ORDER = (STAGE == ST_F) ? 0 : (STAGE == ST_B) ? 1 : 2;
    int base = 0, mask = 0;
    float coef[K][NCORN];
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int c = 0; c < NCORN; ++c) coef[k][c] = 0.f;
    if (in_range) {
        bool ok = true;
        bool lo_ok[DIM], hi_ok[DIM];
        float w[DIM][2], dw[DIM][2], ew[DIM][2];
#pragma unroll
        for (int a = 0; a < DIM; ++a) {
            const AxisRec ar = axis_setup(in.g[a], p.size[a], off, p, align, ORDER);
            ok = ok && ar.ok;
            base += ar.l * p.tstride[a];
            lo_ok[a] = (ar.l >= 0) && (ar.l < p.size[a]);
            hi_ok[a] = (ar.l + 1 >= 0) && (ar.l + 1 < p.size[a]);
            w[a][0] = ar.w0; w[a][1] = ar.w1;
            if (ORDER >= 1) { dw[a][0] = -ar.d; dw[a][1] = ar.d; }
            if (ORDER >= 2) { ew[a][0] = ar.e; ew[a][1] = -ar.e; }
        }
#pragma unroll
        for (int c = 0; c < NCORN; ++c) {
            int b[DIM];
            bool valid = ok;
#pragma unroll
            for (int a = 0; a < DIM; ++a) { b[a] = (c >> a) & 1; valid = valid && (b[a] ? hi_ok[a] : lo_ok[a]); }
            if (!valid) continue;
            mask |= 1 << c;
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
            if (STAGE == ST_F) {
                coef[0][c] = wall;
            } else if (STAGE == ST_B) {
                coef[0][c] = wall;
#pragma unroll
                for (int a = 0; a < DIM; ++a) coef[1 + a][c] = dw[a][b[a]] * wo[a];
            } else if (STAGE == ST_BB) {
                float A = 0.f;
#pragma unroll
                for (int a = 0; a < DIM; ++a) A += dw[a][b[a]] * wo[a] * in.gog[a];
                coef[0][c] = A;
                if (DIM == 2) {
                    // pure second derivatives only (cu2d:705-706)
#pragma unroll
                    for (int a = 0; a < DIM; ++a) coef[1 + a][c] = ew[a][b[a]] * wo[a] * in.gog[a];
                } else {
                    // full Hessian row (cu3d:848-856)
#pragma unroll
                    for (int a = 0; a < DIM; ++a) {
                        float sacc = ew[a][b[a]] * wo[a] * in.gog[a];
#pragma unroll
                        for (int bb = 0; bb < DIM; ++bb) {
                            if (bb == a) continue;
                            const int cc = 3 - a - bb;
                            sacc += dw[a][b[a]] * dw[bb][b[bb]] * w[cc][b[cc]] * in.gog[bb];
                        }
                        coef[1 + a][c] = sacc;
                    }
                }
                if (HAS_U) {
                    coef[1 + DIM][c] = wall;                             // ggOut += gOutInput * w_q
                    if (DIM == 3) {                                      // gGrid += <gOutInput, gOut> D_a,q (cu3d:836-840)
#pragma unroll
                        for (int a = 0; a < DIM; ++a) coef[2 + DIM + a][c] = dw[a][b[a]] * wo[a];
                    }
                }
            } else {  // BBB: pure second derivatives (cu2d:876-885, cu3d:1054-1065)
                float E = 0.f, A = 0.f;
#pragma unroll
                for (int a = 0; a < DIM; ++a) {
                    E += ew[a][b[a]] * wo[a] * (in.gogg[a] * in.gog[a]);
                    if (HAS_X2) A += dw[a][b[a]] * wo[a] * in.gog[a];
                }
                coef[0][c] = E;
                if (HAS_X2) coef[1][c] = A;
            }
        }
    }
    rec4[i] = make_float4(__int_as_float(base), __int_as_float(mask), 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int h = 0; h < CQ; ++h)
            rec4[(1 + k * CQ + h) * PTS + i] =
                make_float4(coef[k][4 * h], coef[k][4 * h + 1], coef[k][4 * h + 2], coef[k][4 * h + 3]);
}


// ---------------------------------------------------------------------------
// The stage kernel: a per-warp software pipeline
//
//   work item  = (tile, channel-vector iteration jj); a tile is PTS points of one cell
//   stage      = PG of the 4 points a lane owns in the item (NST = 4/PG stages per item)
//
// The random part of a stage's input -- the 2^dim corner vectors of PG points (and of
// gOutInput) -- is fetched with cp.async (LDGSTS, 16 B per lane, zero-fill for corners that
// are out of bounds) into lane-private shared-memory slots, one stage ahead of its use.
// The gathers in flight live in shared memory, not in registers, so a warp keeps
// PG*2^dim*16 B per lane in flight while it computes the previous stage; phase 1 of the next
// tile (index map, sincospif, coefficients) and the plain coalesced prefetch of the next
// item's slice of the gOut / gOutggOut streams (registers) overlap with them.
// ---------------------------------------------------------------------------
#ifndef CS_THREADS
#define CS_THREADS 128
#endif
#ifndef CS_MIN_BLOCKS
#define CS_MIN_BLOCKS 3
#endif
#ifndef CS_MIN_BLOCKS_GATHER
#define CS_MIN_BLOCKS_GATHER 4
#endif
#ifndef CS_MIN_BLOCKS_3D
#define CS_MIN_BLOCKS_3D 3
#endif
#ifndef CS_PG2
#define CS_PG2 2
#endif
#ifndef CS_PG3
#define CS_PG3 1
#endif

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory");
}

// Shared-memory plan of one warp (in float4 units).
template <int DIM, int VEC, int LSHIFT, int STAGE, bool HAS_U, bool HAS_X2> struct WarpSmem {
    using RL = RecLayout<DIM, STAGE, HAS_U, HAS_X2>;
    static constexpr int NCORN = 1 << DIM;
    static constexpr int PTS = 128 >> LSHIFT;
    static constexpr int PG = (DIM == 2) ? CS_PG2 : CS_PG3;  // points per stage
    static constexpr int NST = 4 / PG;
    static constexpr int GSLOTS = PG * NCORN * (HAS_U ? 2 : 1);   // gather slots per stage (V then U)
    static constexpr int REC = 2 * RL::FIELDS4 * PTS;        // records, double-buffered by tile
    static constexpr int GBUF = 2 * GSLOTS * 32;             // gathers, double-buffered by stage
    static constexpr int TOTAL = REC + GBUF;                 // float4 per warp
};

// cp.async with a precomputed 32-bit shared address
__device__ __forceinline__ void cp16(unsigned sdst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sdst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp16z(unsigned sdst, const void* gsrc, bool valid) {
    const int n = valid ? 16 : 0;                       // src-size 0 -> 16 bytes of zeros
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(sdst), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp4z(unsigned sdst, const void* gsrc, bool valid) {
    const int n = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" :: "r"(sdst), "l"(gsrc), "r"(n) : "memory");
}

// Per-item constants a lane needs to issue and consume the stages of one work item.
struct ItemCtx {
    const char* vsrc;      // input field of cell n at this lane's channel vector (bytes)
    const char* usrc;      // gOutInput field, same
    char* adst;            // gInput accumulator, same
    int tsb;               // texel stride in bytes
};

// Issue the corner gathers of stage `st` (PG points) of an item into gbuf[st & 1].
template <int DIM, int VEC, int PG, int F4, int PTS, bool HAS_U, bool ALLV>
__device__ __forceinline__ void issue_stage(const float4* rec, int q, int st, const ItemCtx& ic,
                                            const int (&coff)[1 << DIM], unsigned gdst, bool need_v) {
    constexpr int NCORN = 1 << DIM;
    constexpr int GS = PG * NCORN * (HAS_U ? 2 : 1);
    const unsigned d0 = gdst + (st & 1) * GS * 512;
#pragma unroll
    for (int s = 0; s < PG; ++s) {
        const float4 hd = rec[4 * q + st * PG + s];
        const int base = __float_as_int(hd.x);
        const int mask = ALLV ? ((1 << NCORN) - 1) : __float_as_int(hd.y);
#pragma unroll
        for (int c = 0; c < NCORN; ++c) {
            const bool valid = ALLV || ((mask >> c) & 1);
            const long long fo = (long long)(valid ? base + coff[c] : 0) * ic.tsb;
            if (VEC == 4) {
                if (need_v) { if (ALLV) cp16(d0 + (s * NCORN + c) * 512, ic.vsrc + fo);
                              else cp16z(d0 + (s * NCORN + c) * 512, ic.vsrc + fo, valid); }
                if (HAS_U) { if (ALLV) cp16(d0 + ((PG + s) * NCORN + c) * 512, ic.usrc + fo);
                             else cp16z(d0 + ((PG + s) * NCORN + c) * 512, ic.usrc + fo, valid); }
            } else {
                if (need_v) cp4z(d0 + (s * NCORN + c) * 512, ic.vsrc + fo, valid);
                if (HAS_U) cp4z(d0 + ((PG + s) * NCORN + c) * 512, ic.usrc + fo, valid);
            }
        }
    }
}

// Consume stage `st`: contract the gathered corner vectors of PG points.
// SMEMF: the fields (V, U, accumulator) of the cell live in shared memory (small-cell kernel):
// corner vectors are read straight from there and the scatter uses shared-memory atomics.
// Corners are handled four at a time (one float4 of each coefficient), which keeps the live
// state of a 3D point (8 corners) at the size of a 2D one.
template <int DIM, int VEC, int STAGE, int PG, int PTS, bool HAS_U, bool HAS_X2, bool ALLV, bool SMEMF = false>
__device__ __forceinline__ void consume_stage(const float4* rec, const float4* gb, int q, int lane, int st,
                                              const ItemCtx& ic, const int (&coff)[1 << DIM],
                                              bool want_y, bool want_g, bool want_s, bool need_v,
                                              const float (&x1)[4][VEC], const float (&x2)[4][VEC],
                                              float (&y)[4][VEC], float (&gg)[4][DIM]) {
    using RL = RecLayout<DIM, STAGE, HAS_U, HAS_X2>;
    constexpr int NCORN = 1 << DIM;
    constexpr int CQ = RL::CQ;
#pragma unroll
    for (int s = 0; s < PG; ++s) {
        const int t = st * PG + s;
        const int ri = 4 * q + t;
        int base = 0, mask = (1 << NCORN) - 1;
        if (!ALLV || want_s || SMEMF) {
            const float4 hd = rec[ri];
            base = __float_as_int(hd.x);
            if (!ALLV) mask = __float_as_int(hd.y);
        }
        if (!ALLV && mask == 0) continue;
#pragma unroll
        for (int h = 0; h < CQ; ++h) {
            const float4 k0 = rec[(1 + h) * PTS + ri];          // coefficient 0 of corners 4h..4h+3
            float vv[4][VEC], uu[HAS_U ? 4 : 1][VEC];
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                const int c = 4 * h + cc;
                if (SMEMF) {
                    const bool valid = ALLV || ((mask >> c) & 1);
                    const long long fo = (long long)(valid ? base + coff[c] : 0) * ic.tsb;
#pragma unroll
                    for (int k = 0; k < VEC; ++k) { vv[cc][k] = 0.f; if (HAS_U) uu[HAS_U ? cc : 0][k] = 0.f; }
                    if (valid) {
                        if (VEC == 4) {
                            if (need_v) {
                                const float4 a4 = *reinterpret_cast<const float4*>(ic.vsrc + fo);
                                vv[cc][0] = a4.x; vv[cc][1 % VEC] = a4.y; vv[cc][2 % VEC] = a4.z; vv[cc][3 % VEC] = a4.w;
                            }
                            if (HAS_U) {
                                const float4 b4 = *reinterpret_cast<const float4*>(ic.usrc + fo);
                                uu[HAS_U ? cc : 0][0] = b4.x; uu[HAS_U ? cc : 0][1 % VEC] = b4.y;
                                uu[HAS_U ? cc : 0][2 % VEC] = b4.z; uu[HAS_U ? cc : 0][3 % VEC] = b4.w;
                            }
                        } else {
                            if (need_v) vv[cc][0] = *reinterpret_cast<const float*>(ic.vsrc + fo);
                            if (HAS_U) uu[HAS_U ? cc : 0][0] = *reinterpret_cast<const float*>(ic.usrc + fo);
                        }
                    }
                } else if (VEC == 4) {
                    if (need_v) {
                        const float4 a4 = gb[(s * NCORN + c) * 32 + lane];
                        vv[cc][0] = a4.x; vv[cc][1 % VEC] = a4.y; vv[cc][2 % VEC] = a4.z; vv[cc][3 % VEC] = a4.w;
                    }
                    if (HAS_U) {
                        const float4 b4 = gb[((PG + s) * NCORN + c) * 32 + lane];
                        uu[HAS_U ? cc : 0][0] = b4.x; uu[HAS_U ? cc : 0][1 % VEC] = b4.y;
                        uu[HAS_U ? cc : 0][2 % VEC] = b4.z; uu[HAS_U ? cc : 0][3 % VEC] = b4.w;
                    }
                } else {
                    if (need_v) vv[cc][0] = gb[(s * NCORN + c) * 32 + lane].x;
                    if (HAS_U) uu[HAS_U ? cc : 0][0] = gb[((PG + s) * NCORN + c) * 32 + lane].x;
                }
            }
            if (want_y) {
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    const float cy = f4get(k0, cc);
#pragma unroll
                    for (int k = 0; k < VEC; ++k) y[t][k] = fmaf(vv[cc][k], cy, y[t][k]);
                }
                if (HAS_U) {
                    const float4 ku = rec[(1 + (1 + DIM) * CQ + h) * PTS + ri];
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc)
#pragma unroll
                        for (int k = 0; k < VEC; ++k)
                            y[t][k] = fmaf(uu[HAS_U ? cc : 0][k], f4get(ku, cc), y[t][k]);
                }
            }
            if (want_g) {
                float dot[4];
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    dot[cc] = 0.f;
#pragma unroll
                    for (int k = 0; k < VEC; ++k) dot[cc] = fmaf(vv[cc][k], x1[t][k], dot[cc]);
                }
#pragma unroll
                for (int a = 0; a < DIM; ++a) {
                    const float4 kg = rec[(1 + (1 + a) * CQ + h) * PTS + ri];
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) gg[t][a] = fmaf(dot[cc], f4get(kg, cc), gg[t][a]);
                }
                if (HAS_U && DIM == 3) {
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        dot[cc] = 0.f;
#pragma unroll
                        for (int k = 0; k < VEC; ++k) dot[cc] = fmaf(uu[HAS_U ? cc : 0][k], x1[t][k], dot[cc]);
                    }
#pragma unroll
                    for (int a = 0; a < DIM; ++a) {
                        const float4 kg = rec[(1 + (2 + DIM + a) * CQ + h) * PTS + ri];
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) gg[t][a] = fmaf(dot[cc], f4get(kg, cc), gg[t][a]);
                    }
                }
            }
            if (want_s) {
                float4 k1 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (HAS_X2) k1 = rec[(1 + CQ + h) * PTS + ri];
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    const int c = 4 * h + cc;
                    if (ALLV || ((mask >> c) & 1)) {
                        const float cs1 = f4get(k0, cc);
                        float sv[VEC];
#pragma unroll
                        for (int k = 0; k < VEC; ++k) {
                            sv[k] = x1[t][k] * cs1;
                            if (HAS_X2) sv[k] = fmaf(x2[t][k], f4get(k1, cc), sv[k]);
                        }
                        float* dst = reinterpret_cast<float*>(ic.adst + (long long)(base + coff[c]) * ic.tsb);
                        if (SMEMF) {
#pragma unroll
                            for (int k = 0; k < VEC; ++k) atomicAdd(dst + k, sv[k]);    // red.shared.add.f32
                        } else {
                            FieldVec<VEC>::red(dst, sv);
                        }
                    }
                }
            }
        }
    }
}

// SCAT: this instantiation can scatter into the accumulator.  Gather-only calls (u_x, u_xx: the
// majority of a PDE step) run the SCAT = false instantiation, which carries no reduction code,
// needs fewer registers and is compiled for one more resident block per SM.
template <int DIM, int VEC, int LSHIFT, int STAGE, bool HAS_U, bool HAS_X2, bool SCAT>
__global__ void __launch_bounds__(CS_THREADS, (DIM == 2 ? (SCAT ? CS_MIN_BLOCKS : CS_MIN_BLOCKS_GATHER)
                                                        : CS_MIN_BLOCKS_3D))
cs_stage_kernel(const StageParams p) {
    using RL = RecLayout<DIM, STAGE, HAS_U, HAS_X2>;
    using WS = WarpSmem<DIM, VEC, LSHIFT, STAGE, HAS_U, HAS_X2>;
    constexpr int NCORN = 1 << DIM;
    constexpr bool HAS_X1 = (STAGE != ST_F);
    constexpr int L = 1 << LSHIFT;                 // lanes per point quad
    constexpr int PTS = WS::PTS;                   // points per warp tile
    constexpr int PPL = (PTS + 31) / 32;           // points per lane in phase 1
    constexpr int PG = WS::PG;
    constexpr int NST = WS::NST;
    constexpr int F4 = RL::FIELDS4;
    constexpr int FULL = (1 << NCORN) - 1;

    extern __shared__ float4 smem4[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int q = lane >> LSHIFT;                  // quad within the tile
    const int j = lane & (L - 1);                  // lane within the quad
    float4* wbase = smem4 + (size_t)warp * WS::TOTAL;
    float4* recbuf = wbase;                        // [2][F4][PTS]
    float4* gbuf = wbase + WS::REC;                // [2][GSLOTS][32]
    const unsigned gdst = (unsigned)__cvta_generic_to_shared(gbuf + lane);

    const bool want_y = (p.y != nullptr);
    const bool want_g = (p.ggrid != nullptr) && (STAGE == ST_B || STAGE == ST_BB);
    const bool want_s = SCAT && (p.acc != nullptr) && HAS_X1;
    const bool need_v = want_y || want_g;
    const bool svec = p.svec4 != 0;
    const int V = p.C / VEC;                       // channel vectors per texel; L divides V (API)
    const int items_per_tile = V >> LSHIFT;
    const int tsb = p.texel_stride * 4;
    const int csb = p.chan_stride * 4 * VEC;       // bytes between channel vectors
    const long long cellb = p.cell_stride * 4;
    // 2D forward always maps with align_corners = 1 (cu2d:307-308)
    const bool align = (STAGE == ST_F && DIM == 2) ? true : (p.align != 0);

    int coff[NCORN];                               // texel offset of each corner
#pragma unroll
    for (int c = 0; c < NCORN; ++c) {
        coff[c] = 0;
#pragma unroll
        for (int a = 0; a < DIM; ++a) coff[c] += ((c >> a) & 1) * p.tstride[a];
    }

    const long long total = p.num_ptiles * p.N;
    const long long tstep = (long long)gridDim.x * wpb;
    long long tile = (long long)blockIdx.x * wpb + warp;
    if (tile >= total) return;
    // (cell, point-tile) of a tile index, advanced incrementally: one 64-bit division per
    // thread instead of several per tile
    struct TileId { int n; int pt; };                // point-tile index fits 32 bits (API check)
    const bool cell_major = p.cell_major != 0;
    const int nptiles = (int)p.num_ptiles;
    const int step_n = cell_major ? (int)(tstep / nptiles) : (int)(tstep % p.N);
    const int step_p = cell_major ? (int)(tstep % nptiles) : (int)(tstep / p.N);
    auto advance = [&](TileId t) -> TileId {
        t.n += step_n; t.pt += step_p;
        if (cell_major) { if (t.pt >= nptiles) { t.pt -= nptiles; t.n += 1; } }
        else { if (t.n >= p.N) { t.n -= p.N; t.pt += 1; } }
        return t;
    };
    TileId tcur;
    tcur.n = cell_major ? (int)(tile / nptiles) : (int)(tile % p.N);
    tcur.pt = cell_major ? (int)(tile % nptiles) : (int)(tile / p.N);
    TileId tnext = advance(tcur);
    TileId tnext2 = advance(tnext);

    PointIn<DIM, STAGE> pin[PPL];

    auto load_inputs = [&](TileId t) {
        const long long pt0 = (long long)t.pt * PTS;
#pragma unroll
        for (int u = 0; u < PPL; ++u)
            if (u * 32 + lane < PTS) load_point<DIM, STAGE>(pin[u], p, t.n, pt0 + u * 32 + lane);
    };
    // phase 1 of a tile into record buffer `par`; returns "every corner of every point valid"
    auto phase1 = [&](TileId t, int par) -> bool {
        const long long pt0 = (long long)t.pt * PTS;
        const float off = __ldg(p.offset + t.n);
        __syncwarp();                               // everyone is done reading this record buffer
        bool allv = true;
#pragma unroll
        for (int u = 0; u < PPL; ++u) {
            const int i = u * 32 + lane;
            if (i < PTS) {
                float4* rb = recbuf + par * F4 * PTS;
                build_record<DIM, STAGE, HAS_U, HAS_X2, PTS>(rb, i, pin[u], pt0 + i < p.P, off, p, align);
                allv = allv && (__float_as_int(rb[i].y) == FULL);
            }
        }
        return __all_sync(0xffffffffu, allv);       // also a warp barrier: records are visible
    };
    auto make_ctx = [&](TileId t, int jj) -> ItemCtx {
        ItemCtx ic;
        const long long o = (long long)t.n * cellb + (long long)jj * csb;
        ic.vsrc = reinterpret_cast<const char*>(p.V) + o;
        ic.usrc = reinterpret_cast<const char*>(p.U) + o;
        ic.adst = reinterpret_cast<char*>(p.acc) + o;
        ic.tsb = tsb;
        return ic;
    };
    // Stream slices of the *next* item are prefetched straight into registers (plain coalesced
    // 16-byte loads, one item ahead of their use); only the random corner gathers go through
    // the shared-memory cp.async ring.
    float4 xn1[VEC], xn2[HAS_X2 ? VEC : 1];
    auto issue_streams = [&](TileId t, int it) {
        if (!HAS_X1) return;
        const long long qp0 = (long long)t.pt * PTS + 4 * q;
        const int chan0 = (j + it * L) * VEC;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            float tmp[4];
            stream_load4(tmp, p.x1 + t.n * p.x1_sn + (long long)(chan0 + k) * p.x1_sc, qp0, p.P, svec);
            xn1[k] = make_float4(tmp[0], tmp[1], tmp[2], tmp[3]);
            if (HAS_X2) {
                stream_load4(tmp, p.x2 + t.n * p.x2_sn + (long long)(chan0 + k) * p.x2_sc, qp0, p.P, svec);
                xn2[HAS_X2 ? k : 0] = make_float4(tmp[0], tmp[1], tmp[2], tmp[3]);
            }
        }
    };
    auto issue = [&](const float4* rec, int st, const ItemCtx& ic, bool allv) {
        if (!(need_v || HAS_U)) return;
        if (allv) issue_stage<DIM, VEC, PG, F4, PTS, HAS_U, true>(rec, q, st, ic, coff, gdst, need_v);
        else issue_stage<DIM, VEC, PG, F4, PTS, HAS_U, false>(rec, q, st, ic, coff, gdst, need_v);
    };

    // ---- prologue: first item's records, streams and stage-0 gathers ------------------
    load_inputs(tcur);
    bool allv_cur = phase1(tcur, 0);
    if (tile + tstep < total) load_inputs(tnext);
    ItemCtx ic_cur = make_ctx(tcur, j);
    issue_streams(tcur, 0);
    issue(recbuf, 0, ic_cur, allv_cur);
    cp_async_commit();

    int tpar = 0;                                   // record buffer of the current tile
    for (; tile < total; tile += tstep) {
        const int n = tcur.n;
        const long long qp0 = (long long)tcur.pt * PTS + 4 * q;
        const float4* rec = recbuf + tpar * F4 * PTS;
        const bool next_tile = tile + tstep < total;
        bool allv_next = true;

        float gg[4][DIM];
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int a = 0; a < DIM; ++a) gg[t][a] = 0.f;

        for (int it = 0; it < items_per_tile; ++it) {
            const int jj = j + it * L;
            const bool last_item = (it + 1 == items_per_tile);
            ItemCtx ic_next = ic_cur;
            float x1[4][VEC], x2[4][VEC], y[4][VEC];
#pragma unroll
            for (int t = 0; t < 4; ++t)
#pragma unroll
                for (int k = 0; k < VEC; ++k) { x1[t][k] = 0.f; x2[t][k] = 0.f; y[t][k] = 0.f; }
            if (HAS_X1) {
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    x1[0][k] = xn1[k].x; x1[1][k] = xn1[k].y; x1[2][k] = xn1[k].z; x1[3][k] = xn1[k].w;
                    if (HAS_X2) {
                        x2[0][k] = xn2[HAS_X2 ? k : 0].x; x2[1][k] = xn2[HAS_X2 ? k : 0].y;
                        x2[2][k] = xn2[HAS_X2 ? k : 0].z; x2[3][k] = xn2[HAS_X2 ? k : 0].w;
                    }
                }
            }

#pragma unroll
            for (int st = 0; st < NST; ++st) {
                // ---- produce: next stage of this item, or stage 0 of the next item
                if (st + 1 < NST) {
                    issue(rec, st + 1, ic_cur, allv_cur);
                } else if (!last_item) {
                    ic_next = make_ctx(tcur, jj + L);
                    issue_streams(tcur, it + 1);
                    issue(rec, 0, ic_next, allv_cur);
                } else if (next_tile) {
                    allv_next = phase1(tnext, tpar ^ 1);
                    if (tile + 2 * tstep < total) load_inputs(tnext2);
                    ic_next = make_ctx(tnext, j);
                    issue_streams(tnext, 0);
                    issue(recbuf + (tpar ^ 1) * F4 * PTS, 0, ic_next, allv_next);
                }
                cp_async_commit();
                cp_async_wait<1>();                 // everything but the group just committed has landed

                // ---- consume stage st
                const float4* gb = gbuf + (st & 1) * WS::GSLOTS * 32;
                if (allv_cur)
                    consume_stage<DIM, VEC, STAGE, PG, PTS, HAS_U, HAS_X2, true>(
                        rec, gb, q, lane, st, ic_cur, coff, want_y, want_g, want_s, need_v, x1, x2, y, gg);
                else
                    consume_stage<DIM, VEC, STAGE, PG, PTS, HAS_U, HAS_X2, false>(
                        rec, gb, q, lane, st, ic_cur, coff, want_y, want_g, want_s, need_v, x1, x2, y, gg);
            }
            if (want_y) {
                const int chan0 = jj * VEC;
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    const float tmp[4] = {y[0][k], y[1][k], y[2][k], y[3][k]};
                    stream_store4(tmp, p.y + ((long long)n * p.C + chan0 + k) * p.P, qp0, p.P, svec);
                }
            }
            ic_cur = ic_next;
        }

        if (want_g) {
            // sum the channel-partial gradients over the L lanes of the quad
#pragma unroll
            for (int o = L >> 1; o > 0; o >>= 1) {
#pragma unroll
                for (int t = 0; t < 4; ++t)
#pragma unroll
                    for (int a = 0; a < DIM; ++a) gg[t][a] += __shfl_xor_sync(0xffffffffu, gg[t][a], o);
            }
            if (j == 0) {
                float* out = p.ggrid + ((long long)n * p.P + qp0) * DIM;
                if (p.gvec4 && qp0 < p.P) {
                    // 4 points x DIM floats, 16-byte aligned when P % 4 == 0
                    const float* f = &gg[0][0];
#pragma unroll
                    for (int v4 = 0; v4 < DIM; ++v4)
                        reinterpret_cast<float4*>(out)[v4] = make_float4(f[4 * v4], f[4 * v4 + 1], f[4 * v4 + 2], f[4 * v4 + 3]);
                } else {
                    float flat[4 * DIM];
#pragma unroll
                    for (int t = 0; t < 4; ++t)
#pragma unroll
                        for (int a = 0; a < DIM; ++a) flat[t * DIM + a] = gg[t][a];
                    const int nvalid = (int)((p.P - qp0 < 4 ? (p.P - qp0 > 0 ? p.P - qp0 : 0) : 4)) * DIM;
#pragma unroll
                    for (int e = 0; e < 4 * DIM; ++e) if (e < nvalid) out[e] = flat[e];
                }
            }
        }
        tpar ^= 1;
        allv_cur = allv_next;
        tcur = tnext; tnext = tnext2; tnext2 = advance(tnext2);
    }
    cp_async_wait<0>();
}

// ---------------------------------------------------------------------------
// Small-cell kernel: when one cell's fields fit in shared memory (the shapes of
// the reference's own scripts, e.g. [96,4,16,16] = 4 KiB per cell, test_2d.py:26)
// a block stages the cell once, gathers from shared memory and accumulates the
// scatter in a shared-memory private copy that is flushed with one vector red
// per 16 bytes at the end -- the "shared-memory-privatised accumulation when the
// grid fits" of the north star.  Many points fall on few texels here, so the
// global-atomic version serialises in L2; the private copy does not.
//
// grid = (splits, N): block (s, n) handles point tiles s, s+splits, ... of cell n,
// its warps taking them round-robin.
// ---------------------------------------------------------------------------
template <int DIM, int VEC, int LSHIFT, int STAGE, bool HAS_U, bool HAS_X2>
__global__ void __launch_bounds__(256)
cs_small_kernel(const StageParams p) {
    using RL = RecLayout<DIM, STAGE, HAS_U, HAS_X2>;
    constexpr int NCORN = 1 << DIM;
    constexpr bool HAS_X1 = (STAGE != ST_F);
    constexpr int L = 1 << LSHIFT;
    constexpr int PTS = 128 >> LSHIFT;
    constexpr int PPL = (PTS + 31) / 32;
    constexpr int F4 = RL::FIELDS4;
    constexpr int FULL = (1 << NCORN) - 1;

    extern __shared__ float4 smem4[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int q = lane >> LSHIFT;
    const int j = lane & (L - 1);
    const int n = blockIdx.y;

    const bool want_y = (p.y != nullptr);
    const bool want_g = (p.ggrid != nullptr) && (STAGE == ST_B || STAGE == ST_BB);
    const bool want_s = (p.acc != nullptr) && HAS_X1;
    const bool need_v = want_y || want_g;
    const bool svec = p.svec4 != 0;
    const int V = p.C / VEC;
    const int items_per_tile = V >> LSHIFT;
    const int tsb = p.texel_stride * 4;
    const int csb = p.chan_stride * 4 * VEC;
    const bool align = (STAGE == ST_F && DIM == 2) ? true : (p.align != 0);
    const int cell_f4 = (int)(p.cell_stride / 4);          // cell size in float4 (cell elements % 4 == 0, API)

    // shared memory: [V][U][acc] fields of the cell, then one record buffer per warp
    float4* sV = smem4;
    float4* sU = sV + (need_v ? cell_f4 : 0);
    float4* sA = sU + (HAS_U ? cell_f4 : 0);
    float4* recs = sA + (want_s ? cell_f4 : 0);
    float4* rec4 = recs + (size_t)warp * F4 * PTS;

    {
        const float4* gV = reinterpret_cast<const float4*>(p.V + (long long)n * p.cell_stride);
        const float4* gU = reinterpret_cast<const float4*>(p.U + (long long)n * p.cell_stride);
        for (int i = threadIdx.x; i < cell_f4; i += blockDim.x) {
            if (need_v) sV[i] = __ldg(gV + i);
            if (HAS_U) sU[i] = __ldg(gU + i);
            if (want_s) sA[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    __syncthreads();

    int coff[NCORN];
#pragma unroll
    for (int c = 0; c < NCORN; ++c) {
        coff[c] = 0;
#pragma unroll
        for (int a = 0; a < DIM; ++a) coff[c] += ((c >> a) & 1) * p.tstride[a];
    }
    const float off = __ldg(p.offset + n);

    for (long long pt = (long long)blockIdx.x * wpb + warp; pt < p.num_ptiles; pt += (long long)gridDim.x * wpb) {
        const long long pt0 = pt * PTS;
        const long long qp0 = pt0 + 4 * q;
        PointIn<DIM, STAGE> pin[PPL];
#pragma unroll
        for (int u = 0; u < PPL; ++u)
            if (u * 32 + lane < PTS) load_point<DIM, STAGE>(pin[u], p, n, pt0 + u * 32 + lane);

        // stream slices of the first item: issued before phase 1 so their latency overlaps it
        float x1[4][VEC], x2[4][VEC], y[4][VEC];
        auto load_streams = [&](int it) {
            const int chan0 = (j + it * L) * VEC;
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                float tmp[4] = {0.f, 0.f, 0.f, 0.f};
                if (HAS_X1) stream_load4(tmp, p.x1 + n * p.x1_sn + (long long)(chan0 + k) * p.x1_sc, qp0, p.P, svec);
#pragma unroll
                for (int t = 0; t < 4; ++t) x1[t][k] = tmp[t];
                if (HAS_X2) stream_load4(tmp, p.x2 + n * p.x2_sn + (long long)(chan0 + k) * p.x2_sc, qp0, p.P, svec);
#pragma unroll
                for (int t = 0; t < 4; ++t) x2[t][k] = HAS_X2 ? tmp[t] : 0.f;
            }
        };
        load_streams(0);

        __syncwarp();                                   // previous tile's records are no longer read
        bool allv = true;
#pragma unroll
        for (int u = 0; u < PPL; ++u) {
            const int i = u * 32 + lane;
            if (i < PTS) {
                build_record<DIM, STAGE, HAS_U, HAS_X2, PTS>(rec4, i, pin[u], pt0 + i < p.P, off, p, align);
                allv = allv && (__float_as_int(rec4[i].y) == FULL);
            }
        }
        allv = __all_sync(0xffffffffu, allv);

        float gg[4][DIM];
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int a = 0; a < DIM; ++a) gg[t][a] = 0.f;

        for (int it = 0; it < items_per_tile; ++it) {
            const int jj = j + it * L;
            if (it > 0) load_streams(it);
#pragma unroll
            for (int t = 0; t < 4; ++t)
#pragma unroll
                for (int k = 0; k < VEC; ++k) y[t][k] = 0.f;
            ItemCtx ic;
            ic.vsrc = reinterpret_cast<const char*>(sV) + (long long)jj * csb;
            ic.usrc = reinterpret_cast<const char*>(sU) + (long long)jj * csb;
            ic.adst = reinterpret_cast<char*>(sA) + (long long)jj * csb;
            ic.tsb = tsb;
            if (allv)
                consume_stage<DIM, VEC, STAGE, 4, PTS, HAS_U, HAS_X2, true, true>(
                    rec4, nullptr, q, lane, 0, ic, coff, want_y, want_g, want_s, need_v, x1, x2, y, gg);
            else
                consume_stage<DIM, VEC, STAGE, 4, PTS, HAS_U, HAS_X2, false, true>(
                    rec4, nullptr, q, lane, 0, ic, coff, want_y, want_g, want_s, need_v, x1, x2, y, gg);
            if (want_y) {
                const int chan0 = jj * VEC;
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    const float tmp[4] = {y[0][k], y[1][k], y[2][k], y[3][k]};
                    stream_store4(tmp, p.y + ((long long)n * p.C + chan0 + k) * p.P, qp0, p.P, svec);
                }
            }
        }

        if (want_g) {
#pragma unroll
            for (int o = L >> 1; o > 0; o >>= 1) {
#pragma unroll
                for (int t = 0; t < 4; ++t)
#pragma unroll
                    for (int a = 0; a < DIM; ++a) gg[t][a] += __shfl_xor_sync(0xffffffffu, gg[t][a], o);
            }
            if (j == 0) {
                float* out = p.ggrid + ((long long)n * p.P + qp0) * DIM;
                if (p.gvec4 && qp0 < p.P) {
                    const float* f = &gg[0][0];
#pragma unroll
                    for (int v4 = 0; v4 < DIM; ++v4)
                        reinterpret_cast<float4*>(out)[v4] = make_float4(f[4 * v4], f[4 * v4 + 1], f[4 * v4 + 2], f[4 * v4 + 3]);
                } else {
                    float flat[4 * DIM];
#pragma unroll
                    for (int t = 0; t < 4; ++t)
#pragma unroll
                        for (int a = 0; a < DIM; ++a) flat[t * DIM + a] = gg[t][a];
                    const int nvalid = (int)((p.P - qp0 < 4 ? (p.P - qp0 > 0 ? p.P - qp0 : 0) : 4)) * DIM;
#pragma unroll
                    for (int e = 0; e < 4 * DIM; ++e) if (e < nvalid) out[e] = flat[e];
                }
            }
        }
    }

    if (want_s) {
        // flush the private accumulator: one 16-byte vector red per float4, coalesced
        __syncthreads();
        float* gA = p.acc + (long long)n * p.cell_stride;
        for (int i = threadIdx.x; i < cell_f4; i += blockDim.x) {
            const float4 v = sA[i];
            if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) red_add_v4(gA + 4 * i, v.x, v.y, v.z, v.w);
        }
    }
}

}  // namespace cs
