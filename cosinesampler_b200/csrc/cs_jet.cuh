// cs_jet.cuh -- fused multi-cell "jet" sampler (SURVEY section 8f ranks 1 + 2).
//
// The reference evaluates a PDE residual by calling the operator 14-20 times per step
// (forward, a first backward per u_a, a double backward per u_aa, triple backwards to the
// cells; modules_2d.py:38-111), each call streaming [N,C,P] tensors through HBM, and the
// caller replicates the coordinates over the N cells and sums the result over the cells
// (test_2d.py:38,51).  Every one of those calls evaluates, at the same corners with the
// same per-axis kernel values, one of
//
//     z      = sum_n sum_q V_q w_q                    (value)
//     z_a    = sum_n sum_q V_q D_a,q                  (d / d coordinate a)
//     z_aa   = sum_n sum_q V_q D_aa,q                 (pure second derivative)
//
// (formulas: SURVEY section 7.0; cu2d:315-353, 476-503, 694-706).  The jet operator
// produces all of them in ONE gather pass -- jets[J][C][P], J = 1 + ORDER*DIM, summed over
// the cells inside the kernel -- and its adjoint scatters d loss / d jets back into the
// cells in ONE scatter pass.  The caller applies the chain rule of its head to the jets
// (cosinesampler_b200/jet.py), so a training step needs first-order autograd only.
//
// Work decomposition follows cs_engine.cuh: a warp owns a tile of PPQ*(32/L) consecutive
// points, L = C/4 lanes share a quad of PPQ points, each lane holds 4 channels (one 16-byte
// gather per corner, one red.global.add.v4.f32 per corner); the warp walks the N cells of
// its tile and keeps the sums over the cells in registers.  Phase 1 (one point per lane)
// writes a shared-memory record per (cell, point): the per-axis kernel values in the forward
// pass (the jets are separable in the axes, so the corners are contracted axis by axis), the
// finished per-corner coefficients of all J jets in the backward pass (one red per corner);
// the corner gathers run through a per-warp cp.async ring two stages ahead of their use.
#pragma once
#include "cs_engine.cuh"

namespace cs {

struct JetParams {
    int N, C;
    int size[3];            // extent per axis a (0:W 1:H 2:D)
    int tstride[3];         // texel stride per axis, in texels
    long long P;
    long long num_ptiles;
    long long cell_stride;  // elements between cells (T*C)
    const float* V;         // input, channel-last [N,T,C]                   (forward)
    float* acc;             // gInput accumulator, channel-last [N,T,C]      (backward)
    const float* coords;    // [P,DIM]
    const float* offset;    // [N]
    float* jets;            // [J,C,P]                                        (forward)
    const float* gjets;     // [J,C,P]                                        (backward)
    int svec;               // point rows may be accessed as 16-byte vectors
    int cvec2;              // 2D coordinates may be loaded as float2
    int pad, align, kernel, multicell, index_mode;
};

// ORDER 1: value + first derivatives; 2: + pure second derivatives; 3: + the mixed second derivatives
// ((x,y) in 2D; (x,y), (x,z), (y,z) in 3D) -- what the reference's 3D double backward contracts (cu3d:836-856)
template <int DIM, int ORDER> struct JetLayout {
    static constexpr int NCORN = 1 << DIM;
    static constexpr int CQ = NCORN / 4;
    static constexpr int NMIX = (ORDER >= 3) ? DIM * (DIM - 1) / 2 : 0;
    static constexpr int J = 1 + (ORDER >= 2 ? 2 : 1) * DIM + NMIX;
    static constexpr int FIELDS4 = 1 + J * CQ;     // float4 fields per point: header + J coefficient sets
};

// Phase 1 for one (cell, point): record field 0 = (base texel, corner-valid mask);
// field 1 + jt*CQ + h = coefficient of jet jt for corners 4h..4h+3.
//   jt = 0: w_q     jt = 1+a: D_a,q     jt = 1+DIM+a: D_aa,q
template <int DIM, int ORDER, int PTS>
__device__ __forceinline__ void build_jet_record(float4* rec4, int i, const float (&g)[DIM], bool in_range,
                                                 float off, const JetParams& p) {
    using JL = JetLayout<DIM, ORDER>;
    constexpr int NCORN = JL::NCORN;
    constexpr int CQ = JL::CQ;
    constexpr int J = JL::J;
    int base = 0, mask = 0;
    float coef[J][NCORN];
#pragma unroll
    for (int k = 0; k < J; ++k)
#pragma unroll
        for (int c = 0; c < NCORN; ++c) coef[k][c] = 0.f;
    if (in_range) {
        bool ok = true;
        bool lo_ok[DIM], hi_ok[DIM];
        float w[DIM][2], dw[DIM][2], ew[DIM][2];
#pragma unroll
        for (int a = 0; a < DIM; ++a) {
            const AxisRec ar = axis_setup(g[a], p.size[a], off, p, p.align != 0, ORDER);
            ok = ok && ar.ok;
            base += ar.l * p.tstride[a];
            lo_ok[a] = (ar.l >= 0) && (ar.l < p.size[a]);
            hi_ok[a] = (ar.l + 1 >= 0) && (ar.l + 1 < p.size[a]);
            w[a][0] = ar.w0; w[a][1] = ar.w1;
            dw[a][0] = -ar.d; dw[a][1] = ar.d;
            if (ORDER >= 2) { ew[a][0] = ar.e; ew[a][1] = -ar.e; } else { ew[a][0] = ew[a][1] = 0.f; }
        }
#pragma unroll
        for (int c = 0; c < NCORN; ++c) {
            int b[DIM];
            bool valid = ok;
#pragma unroll
            for (int a = 0; a < DIM; ++a) { b[a] = (c >> a) & 1; valid = valid && (b[a] ? hi_ok[a] : lo_ok[a]); }
            if (!valid) continue;
            mask |= 1 << c;
            float wo[DIM], wall;
            if (DIM == 2) {
                wo[0] = w[1][b[1]]; wo[1] = w[0][b[0]];
                wall = w[0][b[0]] * w[1][b[1]];
            } else {
                wo[0] = w[1][b[1]] * w[2][b[2]];
                wo[1] = w[0][b[0]] * w[2][b[2]];
                wo[2] = w[0][b[0]] * w[1][b[1]];
                wall = (w[0][b[0]] * w[1][b[1]]) * w[2][b[2]];
            }
            coef[0][c] = wall;
#pragma unroll
            for (int a = 0; a < DIM; ++a) {
                coef[1 + a][c] = dw[a][b[a]] * wo[a];
                if (ORDER >= 2) coef[1 + DIM + a][c] = ew[a][b[a]] * wo[a];
            }
            if (ORDER >= 3) {
                if (DIM == 2) {
                    coef[(1 + 2 * DIM) % J][c] = dw[0][b[0]] * dw[1][b[1]];
                } else {
                    coef[(1 + 2 * DIM) % J][c] = dw[0][b[0]] * dw[1][b[1]] * w[DIM - 1][b[DIM - 1]];
                    coef[(2 + 2 * DIM) % J][c] = dw[0][b[0]] * dw[DIM - 1][b[DIM - 1]] * w[1][b[1]];
                    coef[(3 + 2 * DIM) % J][c] = dw[1][b[1]] * dw[DIM - 1][b[DIM - 1]] * w[0][b[0]];
                }
            }
        }
    }
    rec4[i] = make_float4(__int_as_float(base), __int_as_float(mask), 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < J; ++k)
#pragma unroll
        for (int h = 0; h < CQ; ++h)
            rec4[(1 + k * CQ + h) * PTS + i] =
                make_float4(coef[k][4 * h], coef[k][4 * h + 1], coef[k][4 * h + 2], coef[k][4 * h + 3]);
}

// Forward-pass record: the jets are separable in the axes, so the gather pass only needs the
// per-axis kernel values (w0, w1, m k', m^2 k'') -- 1 + DIM float4 per (cell, point) instead of
// 1 + J * 2^DIM / 4 -- and contracts the corners axis by axis (slab_contract below):
//   field 0 = (base texel, corner-valid mask), field 1 + a = (w0, w1, d, e) of axis a.
// Corners that are out of bounds are zero-filled by the gather, which is what skipping them means.
template <int DIM, int ORDER, int PTS>
__device__ __forceinline__ void build_jet_axis_record(float4* rec4, int i, const float (&g)[DIM], bool in_range,
                                                      float off, const JetParams& p) {
    int base = 0, mask = 0;
    float4 ax[DIM];
#pragma unroll
    for (int a = 0; a < DIM; ++a) ax[a] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (in_range) {
        bool ok = true;
        bool lo_ok[DIM], hi_ok[DIM];
#pragma unroll
        for (int a = 0; a < DIM; ++a) {
            const AxisRec ar = axis_setup(g[a], p.size[a], off, p, p.align != 0, ORDER);
            ok = ok && ar.ok;
            base += ar.l * p.tstride[a];
            lo_ok[a] = (ar.l >= 0) && (ar.l < p.size[a]);
            hi_ok[a] = (ar.l + 1 >= 0) && (ar.l + 1 < p.size[a]);
            ax[a] = make_float4(ar.w0, ar.w1, ar.d, (ORDER >= 2) ? ar.e : 0.f);
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
    rec4[i] = make_float4(__int_as_float(base), __int_as_float(mask), 0.f, 0.f);
#pragma unroll
    for (int a = 0; a < DIM; ++a) rec4[(1 + a) * PTS + i] = ax[a];
}

// One channel of one (x, y) slab of corners: v00 = (x low, y low), v10 = (x high, y low), v01, v11;
// ax, ay = (w0, w1, d, e) of the two axes.  dW = (-d, +d), d2W = (+e, -e).
struct SlabJet { float A, X, XX, Y, YY, XY; };
__device__ __forceinline__ SlabJet slab_contract(float v00, float v10, float v01, float v11, const float4& ax,
                                                 const float4& ay) {
    const float a0 = fmaf(v10, ax.y, v00 * ax.x);
    const float a1 = fmaf(v11, ax.y, v01 * ax.x);
    const float d0 = v10 - v00, d1 = v11 - v01;
    SlabJet s;
    s.A = fmaf(a1, ay.y, a0 * ay.x);
    s.X = ax.z * fmaf(d1, ay.y, d0 * ay.x);
    s.XX = -ax.w * fmaf(d1, ay.y, d0 * ay.x);
    const float da = a1 - a0;
    s.Y = da * ay.z;
    s.YY = -da * ay.w;
    s.XY = ax.z * ay.z * (d1 - d0);
    return s;
}

// PPQ consecutive points of one row of a [rows, P] array
template <int PPQ>
__device__ __forceinline__ void row_store(const float (&v)[PPQ], float* row, long long p0, long long P, bool vec) {
    if (vec) {
        if (p0 < P) {
            if (PPQ == 4) __stcs(reinterpret_cast<float4*>(row + p0), make_float4(v[0], v[1 % PPQ], v[2 % PPQ], v[3 % PPQ]));
            else __stcs(reinterpret_cast<float2*>(row + p0), make_float2(v[0], v[1 % PPQ]));
        }
    } else {
#pragma unroll
        for (int i = 0; i < PPQ; ++i) if (p0 + i < P) __stcs(row + p0 + i, v[i]);
    }
}
template <int PPQ>
__device__ __forceinline__ void row_load(float (&v)[PPQ], const float* row, long long p0, long long P, bool vec) {
    if (vec) {
        if (p0 < P) {
            if (PPQ == 4) {
                const float4 t = __ldcs(reinterpret_cast<const float4*>(row + p0));
                v[0] = t.x; v[1 % PPQ] = t.y; v[2 % PPQ] = t.z; v[3 % PPQ] = t.w;
            } else {
                const float2 t = __ldcs(reinterpret_cast<const float2*>(row + p0));
                v[0] = t.x; v[1 % PPQ] = t.y;
            }
        } else {
#pragma unroll
            for (int i = 0; i < PPQ; ++i) v[i] = 0.f;
        }
    } else {
#pragma unroll
        for (int i = 0; i < PPQ; ++i) v[i] = (p0 + i < P) ? __ldcs(row + p0 + i) : 0.f;
    }
}

#ifndef CS_JET_RING3D
#define CS_JET_RING3D 3
#endif
template <int DIM, int LSHIFT, int ORDER, int PPQ> struct JetSmem {
    using JL = JetLayout<DIM, ORDER>;
    static constexpr int NCORN = 1 << DIM;
    static constexpr int PTS = PPQ * (32 >> LSHIFT);            // points per warp tile
    static constexpr int PG = (PPQ >= 2) ? PPQ / 2 : 1;           // points per pipeline stage (two stages per item)
    static constexpr int NST = PPQ / PG;
    static constexpr int GSLOTS = PG * NCORN;
    static constexpr int REC1 = JL::FIELDS4 * PTS;               // one record buffer of the backward pass (float4)
    static constexpr int REC1F = (1 + DIM) * PTS;                // one record buffer of the forward pass (per-axis values)
    // ring depth: two stages in flight behind the one consumed (2D 0.217 -> 0.207 ms per 2^20 points; in 3D it
    // only pays with the small per-axis records, which leave room for it at full occupancy)
    static constexpr int RING = (DIM == 2) ? 3 : CS_JET_RING3D;
    static constexpr int GBUF = RING * GSLOTS * 32;              // gather ring
    static constexpr int TOTAL_FWD = 2 * REC1F + GBUF;
    static constexpr int TOTAL_BWD = REC1;
};

#ifndef CS_JET_BLOCKS_2D
#define CS_JET_BLOCKS_2D 3
#endif
#ifndef CS_JET_BLOCKS_3D
#define CS_JET_BLOCKS_3D 3
#endif

// ---------------------------------------------------------------------------
// forward: jets[jt][c][p] = sum_n sum_q V[n][corner q][c] * coef_jt,q(n, p)
// ---------------------------------------------------------------------------
template <int DIM, int LSHIFT, int ORDER, int PPQ>
__global__ void __launch_bounds__(128, (DIM == 2 ? CS_JET_BLOCKS_2D : CS_JET_BLOCKS_3D))
cs_jet_fwd_kernel(const JetParams p) {
    using JL = JetLayout<DIM, ORDER>;
    using WS = JetSmem<DIM, LSHIFT, ORDER, PPQ>;
    constexpr int NCORN = JL::NCORN;
    constexpr int J = JL::J;
    constexpr int L = 1 << LSHIFT;
    constexpr int PTS = WS::PTS;
    constexpr int PPL = (PTS + 31) / 32;
    constexpr int PG = WS::PG;
    constexpr int NST = WS::NST;
    constexpr int GS = WS::GSLOTS;
    constexpr int FULL = (1 << NCORN) - 1;

    extern __shared__ float4 smem4[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int q = lane >> LSHIFT;
    const int j = lane & (L - 1);
    float4* recbuf = smem4 + (size_t)warp * WS::TOTAL_FWD;   // [2][F4][PTS]
    float4* gbuf = recbuf + 2 * WS::REC1F;                    // [RING][GS][32]
    const unsigned gdst = (unsigned)__cvta_generic_to_shared(gbuf + lane);

    const bool svec = p.svec != 0;
    const int tsb = p.C * 4;
    const long long cellb = p.cell_stride * 4;
    const char* vlane = reinterpret_cast<const char*>(p.V) + j * 16;
    const int ncells = p.N;
    const int nptiles = (int)p.num_ptiles;
    const int tstep = (int)gridDim.x * wpb;

    int coff[NCORN];
#pragma unroll
    for (int c = 0; c < NCORN; ++c) {
        coff[c] = 0;
#pragma unroll
        for (int a = 0; a < DIM; ++a) coff[c] += ((c >> a) & 1) * p.tstride[a];
    }

    int pt = (int)blockIdx.x * wpb + warp;
    if (pt >= nptiles) return;

    float gcur[PPL][DIM], gnext[PPL][DIM];
    auto load_coords = [&](float (&g)[PPL][DIM], int tile) {
        const long long pt0 = (long long)tile * PTS;
#pragma unroll
        for (int u = 0; u < PPL; ++u) {
            const int i = u * 32 + lane;
            const long long pi = pt0 + i;
#pragma unroll
            for (int a = 0; a < DIM; ++a) g[u][a] = 0.f;
            if (i < PTS && pi < p.P) {
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
    auto phase1 = [&](const float (&g)[PPL][DIM], int tile, int n, int par) -> bool {
        const long long pt0 = (long long)tile * PTS;
        const float off = __ldg(p.offset + n);
        float4* rb = recbuf + par * WS::REC1F;
        __syncwarp();                               // everyone is done reading this record buffer
        bool allv = true;
#pragma unroll
        for (int u = 0; u < PPL; ++u) {
            const int i = u * 32 + lane;
            if (i < PTS) {
                build_jet_axis_record<DIM, ORDER, PTS>(rb, i, g[u], pt0 + i < p.P, off, p);
                allv = allv && (__float_as_int(rb[i].y) == FULL);
            }
        }
        return __all_sync(0xffffffffu, allv);       // also a warp barrier: records are visible
    };
    auto issue = [&](const float4* rec, int st, int slot, int n, bool allv) {
        const char* vsrc = vlane + (long long)n * cellb;
        const unsigned d0 = gdst + slot * GS * 512;
#pragma unroll
        for (int s = 0; s < PG; ++s) {
            const float4 hd = rec[PPQ * q + st * PG + s];
            const int base = __float_as_int(hd.x);
            const int mask = __float_as_int(hd.y);
#pragma unroll
            for (int c = 0; c < NCORN; ++c) {
                if (allv) {
                    cp16(d0 + (s * NCORN + c) * 512, vsrc + (long long)(base + coff[c]) * tsb);
                } else {
                    const bool valid = (mask >> c) & 1;
                    cp16z(d0 + (s * NCORN + c) * 512, vsrc + (long long)(valid ? base + coff[c] : 0) * tsb, valid);
                }
            }
        }
    };
    static_assert(NST == 2, "the ring schedule below assumes two stages per work item");

    // ---- prologue: records of the first item, both of its stages in flight
    load_coords(gcur, pt);
    bool allv_cur = phase1(gcur, pt, 0, 0);
    int ptn = pt + tstep;
    if (ptn < nptiles) load_coords(gnext, ptn);
    issue(recbuf, 0, 0, 0, allv_cur);
    cp_async_commit();
    if (WS::RING == 3) {
        issue(recbuf, 1, 1, 0, allv_cur);
        cp_async_commit();
    }
    int par = 0;
    int n = 0;
    int slot0 = 0;                              // ring slot of stage 0 of the current item

    // One loop over the work items (tile, cell): the item after (pt, n) is (pt, n+1) or, after the
    // last cell, (next tile, 0).  While stage st of an item is consumed, stage st of the NEXT item is
    // issued, so two stages (2 x PG points x 2^DIM corners x 16 B per lane) are always in flight:
    // the kernel is bound by bytes in flight over L2 latency (one stage ahead: 5.0 TB/s of gathers).
    float acc[J][PPQ][4];
    while (pt < nptiles) {
        if (n == 0) {
#pragma unroll
            for (int jt = 0; jt < J; ++jt)
#pragma unroll
                for (int t = 0; t < PPQ; ++t)
#pragma unroll
                    for (int k = 0; k < 4; ++k) acc[jt][t][k] = 0.f;
        }
        const bool last_cell = (n + 1 == ncells);
        const int pt_nx = last_cell ? ptn : pt;
        const int n_nx = last_cell ? 0 : n + 1;
        const bool have_next = pt_nx < nptiles;
        const float4* rec = recbuf + par * WS::REC1F;
        const float4* rec_nx = recbuf + (par ^ 1) * WS::REC1F;
        bool allv_next = true;
        if (have_next) {
            float gsel[PPL][DIM];
#pragma unroll
            for (int u = 0; u < PPL; ++u)
#pragma unroll
                for (int a = 0; a < DIM; ++a) gsel[u][a] = last_cell ? gnext[u][a] : gcur[u][a];
            allv_next = phase1(gsel, pt_nx, n_nx, par ^ 1);
        }
#pragma unroll
        for (int st = 0; st < NST; ++st) {
            int slot_c;
            if (WS::RING == 3) {
                // ---- produce: the same stage of the next item, two ring slots ahead
                slot_c = slot0 + st;
                if (slot_c >= 3) slot_c -= 3;
                int slot_p = slot_c + 2;
                if (slot_p >= 3) slot_p -= 3;
                if (have_next) issue(rec_nx, st, slot_p, n_nx, allv_next);
                cp_async_commit();
                cp_async_wait<2>();             // everything but the two groups just committed has landed
            } else {
                // ---- produce: the next stage of this item, or stage 0 of the next item
                slot_c = st;
                if (st + 1 < NST) issue(rec, st + 1, st + 1, n, allv_cur);
                else if (have_next) issue(rec_nx, 0, 0, n_nx, allv_next);
                cp_async_commit();
                cp_async_wait<1>();             // everything but the group just committed has landed
            }

            // ---- consume stage st: corners contracted axis by axis with the per-axis kernel values
            const float4* gb = gbuf + slot_c * GS * 32;
#pragma unroll
            for (int s = 0; s < PG; ++s) {
                const int t = st * PG + s;
                const int ri = PPQ * q + t;
                const float4 ax = rec[1 * PTS + ri];
                const float4 ay = rec[2 * PTS + ri];
                if (DIM == 2) {
                    float4 v[4];
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) v[cc] = gb[(s * NCORN + cc) * 32 + lane];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const SlabJet sj = slab_contract(f4get(v[0], k), f4get(v[1], k), f4get(v[2], k), f4get(v[3], k), ax, ay);
                        acc[0][t][k] += sj.A;
                        acc[1][t][k] += sj.X;
                        acc[2][t][k] += sj.Y;
                        if (ORDER >= 2) { acc[3 % J][t][k] += sj.XX; acc[4 % J][t][k] += sj.YY; }
                        if (ORDER >= 3) acc[5 % J][t][k] += sj.XY;
                    }
                } else {
                    const float4 az = rec[(DIM == 3 ? 3 : 2) * PTS + ri];
                    float4 v0[4], v1[4];
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        v0[cc] = gb[(s * NCORN + cc) * 32 + lane];
                        v1[cc] = gb[(s * NCORN + 4 + cc) * 32 + lane];
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const SlabJet lo = slab_contract(f4get(v0[0], k), f4get(v0[1], k), f4get(v0[2], k), f4get(v0[3], k), ax, ay);
                        const SlabJet hi = slab_contract(f4get(v1[0], k), f4get(v1[1], k), f4get(v1[2], k), f4get(v1[3], k), ax, ay);
                        acc[0][t][k] += fmaf(hi.A, az.y, lo.A * az.x);
                        acc[1][t][k] += fmaf(hi.X, az.y, lo.X * az.x);
                        acc[2][t][k] += fmaf(hi.Y, az.y, lo.Y * az.x);
                        acc[DIM][t][k] += (hi.A - lo.A) * az.z;
                        if (ORDER >= 2) {
                            acc[1 + DIM][t][k] += fmaf(hi.XX, az.y, lo.XX * az.x);
                            acc[2 + DIM][t][k] += fmaf(hi.YY, az.y, lo.YY * az.x);
                            acc[2 * DIM][t][k] += (lo.A - hi.A) * az.w;
                        }
                        if (ORDER >= 3) {
                            acc[(1 + 2 * DIM) % J][t][k] += fmaf(hi.XY, az.y, lo.XY * az.x);
                            acc[(2 + 2 * DIM) % J][t][k] += (hi.X - lo.X) * az.z;
                            acc[(3 + 2 * DIM) % J][t][k] += (hi.Y - lo.Y) * az.z;
                        }
                    }
                }
            }
        }
        slot0 += NST;
        if (slot0 >= WS::RING) slot0 -= WS::RING;
        par ^= 1;
        allv_cur = allv_next;
        if (last_cell) {
            // ---- the sums over the cells leave the registers once per tile
            const long long qp0 = (long long)pt * PTS + PPQ * q;
#pragma unroll
            for (int jt = 0; jt < J; ++jt)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float tmp[PPQ];
#pragma unroll
                    for (int t = 0; t < PPQ; ++t) tmp[t] = acc[jt][t][k];
                    row_store<PPQ>(tmp, p.jets + ((long long)jt * p.C + 4 * j + k) * p.P, qp0, p.P, svec);
                }
#pragma unroll
            for (int u = 0; u < PPL; ++u)
#pragma unroll
                for (int a = 0; a < DIM; ++a) gcur[u][a] = gnext[u][a];
            pt = ptn;
            ptn += tstep;
            if (ptn < nptiles) load_coords(gnext, ptn);
            n = 0;
        } else {
            ++n;
        }
    }
    cp_async_wait<0>();
}

// ---------------------------------------------------------------------------
// backward (adjoint): acc[n][corner q][c] += sum_jt gjets[jt][c][p] * coef_jt,q(n, p)
// ---------------------------------------------------------------------------
template <int DIM, int LSHIFT, int ORDER, int PPQ>
__global__ void __launch_bounds__(128, (DIM == 2 ? CS_JET_BLOCKS_2D : CS_JET_BLOCKS_3D))
cs_jet_bwd_kernel(const JetParams p) {
    using JL = JetLayout<DIM, ORDER>;
    using WS = JetSmem<DIM, LSHIFT, ORDER, PPQ>;
    constexpr int NCORN = JL::NCORN;
    constexpr int CQ = JL::CQ;
    constexpr int J = JL::J;
    constexpr int L = 1 << LSHIFT;
    constexpr int PTS = WS::PTS;
    constexpr int PPL = (PTS + 31) / 32;
    constexpr int FULL = (1 << NCORN) - 1;

    extern __shared__ float4 smem4[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int q = lane >> LSHIFT;
    const int j = lane & (L - 1);
    float4* rec = smem4 + (size_t)warp * WS::TOTAL_BWD;

    const bool svec = p.svec != 0;
    const int ncells = p.N;
    const int nptiles = (int)p.num_ptiles;
    const int tstep = (int)gridDim.x * wpb;

    int coff[NCORN];
#pragma unroll
    for (int c = 0; c < NCORN; ++c) {
        coff[c] = 0;
#pragma unroll
        for (int a = 0; a < DIM; ++a) coff[c] += ((c >> a) & 1) * p.tstride[a];
    }

    for (int pt = (int)blockIdx.x * wpb + warp; pt < nptiles; pt += tstep) {
        const long long pt0 = (long long)pt * PTS;
        const long long qp0 = pt0 + PPQ * q;
        // the J x 4 rows of d loss / d jets of this lane's quad: read once, used for all N cells
        float x[J][PPQ][4];
#pragma unroll
        for (int jt = 0; jt < J; ++jt)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float tmp[PPQ];
                row_load<PPQ>(tmp, p.gjets + ((long long)jt * p.C + 4 * j + k) * p.P, qp0, p.P, svec);
#pragma unroll
                for (int t = 0; t < PPQ; ++t) x[jt][t][k] = tmp[t];
            }
        float g[PPL][DIM];
#pragma unroll
        for (int u = 0; u < PPL; ++u) {
            const int i = u * 32 + lane;
            const long long pi = pt0 + i;
#pragma unroll
            for (int a = 0; a < DIM; ++a) g[u][a] = 0.f;
            if (i < PTS && pi < p.P) {
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

        for (int n = 0; n < ncells; ++n) {
            const float off = __ldg(p.offset + n);
            __syncwarp();                           // the previous cell's records are no longer read
            bool allv = true;
#pragma unroll
            for (int u = 0; u < PPL; ++u) {
                const int i = u * 32 + lane;
                if (i < PTS) {
                    build_jet_record<DIM, ORDER, PTS>(rec, i, g[u], pt0 + i < p.P, off, p);
                    allv = allv && (__float_as_int(rec[i].y) == FULL);
                }
            }
            allv = __all_sync(0xffffffffu, allv);
            float* adst = p.acc + (long long)n * p.cell_stride + 4 * j;
#pragma unroll
            for (int t = 0; t < PPQ; ++t) {
                const int ri = PPQ * q + t;
                const float4 hd = rec[ri];
                const int base = __float_as_int(hd.x);
                const int mask = allv ? FULL : __float_as_int(hd.y);
                if (mask == 0) continue;
#pragma unroll
                for (int h = 0; h < CQ; ++h) {
                    float sv[4][4];
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc)
#pragma unroll
                        for (int k = 0; k < 4; ++k) sv[cc][k] = 0.f;
#pragma unroll
                    for (int jt = 0; jt < J; ++jt) {
                        const float4 k4 = rec[(1 + jt * CQ + h) * PTS + ri];
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) {
                            const float cf = f4get(k4, cc);
#pragma unroll
                            for (int k = 0; k < 4; ++k) sv[cc][k] = fmaf(x[jt][t][k], cf, sv[cc][k]);
                        }
                    }
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        const int c = 4 * h + cc;
                        if ((mask >> c) & 1)
                            red_add_v4(adst + (long long)(base + coff[c]) * p.C, sv[cc][0], sv[cc][1], sv[cc][2], sv[cc][3]);
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// launch
// ---------------------------------------------------------------------------
template <int DIM, int LSHIFT, int ORDER, int PPQ, bool BWD>
cudaError_t launch_jet_one(const JetParams& p, cudaStream_t stream) {
    using WS = JetSmem<DIM, LSHIFT, ORDER, PPQ>;
    auto kern = BWD ? cs_jet_bwd_kernel<DIM, LSHIFT, ORDER, PPQ> : cs_jet_fwd_kernel<DIM, LSHIFT, ORDER, PPQ>;
    constexpr size_t smem_per_warp = (size_t)(BWD ? WS::TOTAL_BWD : WS::TOTAL_FWD) * sizeof(float4);
    int warps = 4;
    while (warps > 1 && warps * smem_per_warp > 72 * 1024) warps >>= 1;
    const int threads = warps * 32;
    const size_t smem = warps * smem_per_warp;
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    static int occ_cache[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    int& occ = occ_cache[dev];
    if (occ == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem);
        if (e != cudaSuccess) return e;
        if (occ < 1) occ = 1;
    }
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    long long blocks = (p.num_ptiles + warps - 1) / warps;
    const long long cap = (long long)sms * occ;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) return cudaSuccess;
    kern<<<(unsigned)blocks, threads, smem, stream>>>(p);
    return cudaGetLastError();
}

#ifndef CS_JET_PPQ2D
#define CS_JET_PPQ2D 4
#endif
template <int DIM> struct JetPPQ { static constexpr int value = (DIM == 2) ? CS_JET_PPQ2D : 2; };

template <int DIM, int LSHIFT>
cudaError_t launch_jet_variant(int order, bool backward, JetParams& p, cudaStream_t s) {
    constexpr int PPQ = JetPPQ<DIM>::value;
    constexpr int PTS = PPQ * (32 >> LSHIFT);
    p.num_ptiles = (p.P + PTS - 1) / PTS;
    if (order == 3)
        return backward ? launch_jet_one<DIM, LSHIFT, 3, PPQ, true>(p, s) : launch_jet_one<DIM, LSHIFT, 3, PPQ, false>(p, s);
    if (order == 2)
        return backward ? launch_jet_one<DIM, LSHIFT, 2, PPQ, true>(p, s) : launch_jet_one<DIM, LSHIFT, 2, PPQ, false>(p, s);
    return backward ? launch_jet_one<DIM, LSHIFT, 1, PPQ, true>(p, s) : launch_jet_one<DIM, LSHIFT, 1, PPQ, false>(p, s);
}

}  // namespace cs
