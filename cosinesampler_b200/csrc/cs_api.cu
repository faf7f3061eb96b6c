// cs_api.cu -- extern "C" boundary of libcosine_sampler_b200.so
// (see include/cosine_sampler_b200.h for the contract and the reference
// interfaces each entry point replaces) plus the layout-staging kernels.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>

#include "../../include/cosine_sampler_b200.h"
#include "cs_engine.cuh"
#include "cs_jet.cuh"
#include "cs_head.cuh"
#include "cs_fused.cuh"
#include "cs_scalar.cuh"

namespace cs {
// one function per (dim, field vector width, log2 lanes) variant, each in its own object file
#define CS_DECL(name) cudaError_t name(int stage, bool has_u, bool has_x2, const StageParams& p, cudaStream_t s);
CS_DECL(launch_d2_v4_l0) CS_DECL(launch_d2_v4_l1) CS_DECL(launch_d2_v4_l2) CS_DECL(launch_d2_v4_l3)
CS_DECL(launch_d3_v4_l0) CS_DECL(launch_d3_v4_l1) CS_DECL(launch_d3_v4_l2) CS_DECL(launch_d3_v4_l3)
CS_DECL(launch_d2_v1_l0) CS_DECL(launch_d3_v1_l0)
#undef CS_DECL
#define CS_DECLJ(name) cudaError_t name(int order, bool backward, JetParams& p, cudaStream_t s);
CS_DECLJ(launch_jet_d2_l0) CS_DECLJ(launch_jet_d2_l1) CS_DECLJ(launch_jet_d2_l2) CS_DECLJ(launch_jet_d2_l3)
CS_DECLJ(launch_jet_d3_l0) CS_DECLJ(launch_jet_d3_l1) CS_DECLJ(launch_jet_d3_l2) CS_DECLJ(launch_jet_d3_l3)
#undef CS_DECLJ
cudaError_t launch_head_any(int dim, int C, const HeadParams& p, cudaStream_t s);
#define CS_DECLF(name) cudaError_t name(FusedParams& p, cudaStream_t s);
CS_DECLF(launch_fused_d2_l0) CS_DECLF(launch_fused_d2_l1) CS_DECLF(launch_fused_d2_l2) CS_DECLF(launch_fused_d2_l3)
CS_DECLF(launch_fused_d2_l4) CS_DECLF(launch_fused_d3_l4)
CS_DECLF(launch_fused_d3_l0) CS_DECLF(launch_fused_d3_l1) CS_DECLF(launch_fused_d3_l2) CS_DECLF(launch_fused_d3_l3)
#undef CS_DECLF
}  // namespace cs

namespace {

thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char* what) {
    snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return (int)e;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int gcd8(int v) {  // largest power of two <= 8 dividing v
    int l = 1;
    while (l < 8 && v % (l * 2) == 0) l *= 2;
    return l;
}

constexpr int CS_NOTHING_TO_DO = 1 << 30;   // internal: valid but empty problem

// Validate the problem and fill the geometry part of StageParams.
int setup(const cs_problem* pb, cs::StageParams& p, const float* grid, const float* offset) {
    if (!pb) return fail(CS_EINVAL, "cs_problem is NULL");
    if (pb->dim != 2 && pb->dim != 3) return fail(CS_EINVAL, "dim must be 2 or 3, got %d", pb->dim);
    if (pb->N < 0 || pb->C < 0 || pb->P < 0) return fail(CS_EINVAL, "negative size");
    if (pb->H < 1 || pb->W < 1 || pb->D < 1) return fail(CS_EINVAL, "cell extent must be >= 1");
    if (pb->dim == 2 && pb->D != 1) return fail(CS_EINVAL, "D must be 1 when dim == 2");
    if (pb->padding_mode < 0 || pb->padding_mode > 2) return fail(CS_EINVAL, "bad padding_mode %d", pb->padding_mode);
    if (pb->kernel < 0 || pb->kernel > 2) return fail(CS_EINVAL, "bad kernel %d", pb->kernel);
    if (pb->index_mode < 0 || pb->index_mode > 1) return fail(CS_EINVAL, "bad index_mode %d", pb->index_mode);
    if (pb->field_layout < 0 || pb->field_layout > 1) return fail(CS_EINVAL, "bad field_layout %d", pb->field_layout);
    if (pb->grad_order < 0 || pb->grad_order > 1) return fail(CS_EINVAL, "bad grad_order %d", pb->grad_order);
    if (pb->small_cell < 0 || pb->small_cell > 2) return fail(CS_EINVAL, "bad small_cell %d", pb->small_cell);
    if (pb->lanes != 0 && pb->lanes != 1 && pb->lanes != 2 && pb->lanes != 4 && pb->lanes != 8)
        return fail(CS_EINVAL, "lanes must be 0,1,2,4 or 8");
    if (pb->N == 0 || pb->C == 0 || pb->P == 0) return CS_NOTHING_TO_DO;   // empty problem (cu2d:904)
    if (!grid || !offset) return fail(CS_EINVAL, "grid/offset pointer is NULL");
    const long long T = (long long)pb->D * pb->H * pb->W;
    if (T * (long long)pb->C >= (1ll << 31))
        return fail(CS_EUNSUPPORTED, "a cell has %lld elements; the per-cell index is 32-bit", T * pb->C);
    memset(&p, 0, sizeof(p));
    p.N = pb->N; p.C = pb->C; p.P = pb->P;
    p.size[0] = pb->W; p.size[1] = pb->H; p.size[2] = pb->D;
    p.tstride[0] = 1; p.tstride[1] = pb->W; p.tstride[2] = pb->W * pb->H;
    p.cell_stride = T * pb->C;
    if (pb->field_layout == CS_LAYOUT_CHANNEL_LAST) { p.texel_stride = pb->C; p.chan_stride = 1; }
    else { p.texel_stride = 1; p.chan_stride = (int)T; }
    p.grid = grid; p.grid_sn = pb->grid_stride_n; p.offset = offset;
    p.grid_vec2 = ((reinterpret_cast<uintptr_t>(grid) & 7u) == 0) && (pb->grid_stride_n % 2 == 0);
    p.pad = pb->padding_mode; p.align = pb->align_corners; p.kernel = pb->kernel;
    p.multicell = pb->multicell; p.index_mode = pb->index_mode;
    p.small_cell = pb->small_cell;
    return 0;
}

int run(const cs_problem* pb, cs::StageParams& p, int stage, bool has_u, bool has_x2, void* stream) {
    if (p.N == 0 || p.C == 0 || p.P == 0) return 0;   // empty problem: nothing to launch (cu2d:904)
    // field vector width
    bool vec4 = (pb->field_layout == CS_LAYOUT_CHANNEL_LAST) && (p.C % 4 == 0) &&
                aligned16(p.V) && aligned16(p.U) && aligned16(p.acc);
    const int vec = vec4 ? 4 : 1;
    // scalar fields: one lane walks all channels of its point quad (channel-first strides
    // give the lanes of a quad nothing to share)
    int lanes = vec4 ? gcd8(p.C / 4) : 1;
    if (vec4 && pb->lanes && (p.C / 4) % pb->lanes == 0) lanes = pb->lanes;   // must divide C/4
    p.lshift = (lanes == 1) ? 0 : (lanes == 2) ? 1 : (lanes == 4) ? 2 : 3;
    const int pts = 128 >> p.lshift;
    p.num_ptiles = (p.P + pts - 1) / pts;
    if (p.num_ptiles >= (1ll << 31)) return fail(CS_EUNSUPPORTED, "too many points per cell (%lld)", (long long)p.P);
    // stream vector width
    p.svec4 = (p.P % 4 == 0) && aligned16(p.x1) && aligned16(p.x2) && aligned16(p.y) &&
              p.x1_sn % 4 == 0 && p.x1_sc % 4 == 0 && p.x2_sn % 4 == 0 && p.x2_sc % 4 == 0;
    p.gvec4 = (p.P % 4 == 0) && aligned16(p.ggrid);
    {   // L2 working set of the grid-shaped fields this call touches at random: if all cells
        // together exceed about half of the 126 MB L2, walk the cells one after the other
        const bool reads_v = (p.V != nullptr) && (p.y != nullptr || p.ggrid != nullptr);
        const int nfields = (reads_v ? 1 : 0) + (p.U ? 1 : 0) + (p.acc ? 1 : 0);
        const long long footprint = (long long)nfields * p.N * p.cell_stride * 4;
        static const long long limit_mb = [] { const char* e = getenv("CS_CELL_MAJOR_MB"); return e ? atoll(e) : 80ll; }();
        p.cell_major = footprint > (limit_mb << 20);
    }
    using Fn = cudaError_t (*)(int, bool, bool, const cs::StageParams&, cudaStream_t);
    static const Fn table[2][5] = {
        {cs::launch_d2_v4_l0, cs::launch_d2_v4_l1, cs::launch_d2_v4_l2, cs::launch_d2_v4_l3, cs::launch_d2_v1_l0},
        {cs::launch_d3_v4_l0, cs::launch_d3_v4_l1, cs::launch_d3_v4_l2, cs::launch_d3_v4_l3, cs::launch_d3_v1_l0}};
    const Fn fn = table[pb->dim - 2][vec4 ? p.lshift : 4];
    cudaError_t e = fn(stage, has_u, has_x2, p, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "stage kernel launch");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

// ---------------------------------------------------------------------------
// Layout staging: [N, C, T] <-> [N, T, C]
// A block moves a 32-texel x 32-channel tile through shared memory so that both
// the T-contiguous side and the C-contiguous side are accessed in runs.
// ---------------------------------------------------------------------------
template <bool TO_CL, bool ACCUM>
__global__ void __launch_bounds__(256) cs_layout_kernel(const float* __restrict__ src,
                                                        float* __restrict__ dst, int C, long long T) {
    __shared__ float tile[32][33];          // [channel][texel]
    const int n = blockIdx.z;
    const long long t0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    const int ct = min(32, C - c0);
    const float* s = src + (long long)n * C * T;
    float* d = dst + (long long)n * C * T;
    if (TO_CL) {
        // read [c][t] rows (t contiguous)
        for (int c = ty; c < ct; c += 8)
            if (t0 + tx < T) tile[c][tx] = s[(long long)(c0 + c) * T + t0 + tx];
        __syncthreads();
        // write [t][c]: linear over the (texel, channel) pairs of the tile
        for (int idx = threadIdx.x; idx < 32 * ct; idx += 256) {
            const int t = idx / ct, c = idx - t * ct;
            if (t0 + t < T) d[(t0 + t) * C + c0 + c] = tile[c][t];
        }
    } else {
        for (int idx = threadIdx.x; idx < 32 * ct; idx += 256) {
            const int t = idx / ct, c = idx - t * ct;
            if (t0 + t < T) tile[c][t] = s[(t0 + t) * C + c0 + c];
        }
        __syncthreads();
        for (int c = ty; c < ct; c += 8) {
            if (t0 + tx < T) {
                float* o = d + (long long)(c0 + c) * T + t0 + tx;
                if (ACCUM) *o += tile[c][tx]; else *o = tile[c][tx];
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Reference-order gGrid (cs_problem.grad_order == 1): one thread per (cell, point), channels
// in order, corners in order (x bit fastest), the products and fused multiply-adds written
// exactly as the reference / ATen write them:
//     gix -= nw_val * (iy_se - iy) * gOut      ->   t = nw_val * wy ;  gix = fma(-t, gOut, gix)
// (cu2d:476-495, cu3d:525-572; nvcc contracts a - b*c into an fma in both code bases), and
// gGrid = mult * gix [* k'] at the end (cu2d:502-503).  Not a fast path: it exists so that the
// linear / multicell=False backward can be compared bit for bit with torch grid_sample.
// ---------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(256) cs_backward_reforder_kernel(const cs::StageParams p) {
    constexpr int NCORN = 1 << DIM;
    const long long total = (long long)p.N * p.P;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(idx / p.P);
        const long long pi = idx - (long long)n * p.P;
        const float off = __ldg(p.offset + n);
        const float* gp = p.grid + (long long)n * p.grid_sn + pi * DIM;
        cs::AxisRec ar[DIM];
        bool ok = true;
        int base = 0;
#pragma unroll
        for (int a = 0; a < DIM; ++a) {
            ar[a] = cs::axis_setup(__ldg(gp + a), p.size[a], off, p, p.align != 0, 1);
            ok = ok && ar[a].ok;
            base += ar[a].l * p.tstride[a];
        }
        float gg[DIM];
#pragma unroll
        for (int a = 0; a < DIM; ++a) gg[a] = 0.f;
        const float* Vn = p.V + (long long)n * p.cell_stride;
        for (int c = 0; c < p.C; ++c) {
            const float g = __ldg(p.x1 + n * p.x1_sn + (long long)c * p.x1_sc + pi);
#pragma unroll
            for (int q = 0; q < NCORN; ++q) {
                int b[DIM];
                bool valid = ok;
                int texel = base;
#pragma unroll
                for (int a = 0; a < DIM; ++a) {
                    b[a] = (q >> a) & 1;
                    const int la = ar[a].l + b[a];
                    valid = valid && (la >= 0) && (la < p.size[a]);
                    texel += b[a] * p.tstride[a];
                }
                if (!valid) continue;
                const float v = __ldg(Vn + (long long)texel * p.texel_stride + (long long)c * p.chan_stride);
#pragma unroll
                for (int a = 0; a < DIM; ++a) {
                    float t = v;
#pragma unroll
                    for (int o = 0; o < DIM; ++o)
                        if (o != a) t = __fmul_rn(t, b[o] ? ar[o].w1 : ar[o].w0);
                    gg[a] = __fmaf_rn(b[a] ? t : -t, g, gg[a]);
                }
            }
        }
        float* out = p.ggrid + ((long long)n * p.P + pi) * DIM;
#pragma unroll
        for (int a = 0; a < DIM; ++a) out[a] = __fmul_rn(ar[a].d, gg[a]);   // d = mult * k'  (k' = 1 when linear)
    }
}

int launch_reforder(const cs_problem* pb, const cs::StageParams& p, void* stream) {
    const long long total = (long long)p.N * p.P;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    if (blocks < 1) return 0;
    if (pb->dim == 2) cs_backward_reforder_kernel<2><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
    else cs_backward_reforder_kernel<3><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "reference-order backward launch");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

// ---------------------------------------------------------------------------
// Gradient all-reduce over peer memory, fused with the layout change.
// Every rank holds a channel-last accumulator [N, T, C] in symmetric memory (mapped into all peers
// over NVLink / NVSwitch).  The (32 texel x 32 channel) tiles are dealt round-robin to the ranks;
// the owner of a tile loads it from every peer's accumulator, sums in rank order, transposes through
// shared memory and stores the channel-first result into every peer's output [N, C, T]: a
// reduce-scatter and an all-gather in one kernel, each byte crossing NVLink once per direction, and
// the cs_from_channel_last pass for free.  One extra block sums a small vector (head gradients, loss)
// from all peers for this rank.  The caller brackets the launch with symmetric-memory barriers.
// ---------------------------------------------------------------------------
struct PeerParams {
    const float* acc[CS_MAX_PEERS];
    float* out[CS_MAX_PEERS];
    const float* small_in[CS_MAX_PEERS];
    float* small_out;
    int world, rank, N, C, small_n, tiles_y;
    int vec4;                   // every accumulator is 16-byte aligned
    long long T, tiles_x, total_tiles, my_tiles;
};

__device__ __forceinline__ float ld_peer(const float* p) {
    float v;
    asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ float4 ld_peer_v4(const float* p) {
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

__global__ void __launch_bounds__(256) cs_peer_reduce_kernel(const PeerParams p) {
    __shared__ float tile[32][33];          // [channel][texel]
    const long long b = blockIdx.x;
    if (b < p.my_tiles) {
        const long long id = b * p.world + p.rank;
        const long long bx = id % p.tiles_x;
        const long long rem = id / p.tiles_x;
        const int by = (int)(rem % p.tiles_y);
        const int n = (int)(rem / p.tiles_y);
        const long long t0 = bx * 32;
        const int c0 = by * 32;
        const int ct = min(32, p.C - c0);
        if ((ct & 3) == 0 && (p.C & 3) == 0 && p.vec4) {
            // 16-byte peer loads: a texel's channels are contiguous
            const int cq = ct >> 2;
            for (int idx = threadIdx.x; idx < 32 * cq; idx += 256) {
                const int t = idx / cq, c = 4 * (idx - t * cq);
                if (t0 + t < p.T) {
                    const long long off = ((long long)n * p.T + t0 + t) * p.C + c0 + c;
                    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int r = 0; r < p.world; ++r) {
                        const float4 v = ld_peer_v4(p.acc[r] + off);
                        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
                    }
                    tile[c][t] = s.x; tile[c + 1][t] = s.y; tile[c + 2][t] = s.z; tile[c + 3][t] = s.w;
                }
            }
        } else {
            for (int idx = threadIdx.x; idx < 32 * ct; idx += 256) {
                const int t = idx / ct, c = idx - t * ct;
                if (t0 + t < p.T) {
                    const long long off = ((long long)n * p.T + t0 + t) * p.C + c0 + c;
                    float s = 0.f;
                    for (int r = 0; r < p.world; ++r) s += ld_peer(p.acc[r] + off);
                    tile[c][t] = s;
                }
            }
        }
        __syncthreads();
        const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
        for (int c = ty; c < ct; c += 8) {
            if (t0 + tx < p.T) {
                const float v = tile[c][tx];
                const long long off = ((long long)n * p.C + c0 + c) * p.T + t0 + tx;
                for (int r = 0; r < p.world; ++r) p.out[r][off] = v;
            }
        }
    } else {
        for (int i = threadIdx.x; i < p.small_n; i += 256) {
            float s = 0.f;
            for (int r = 0; r < p.world; ++r) s += ld_peer(p.small_in[r] + i);
            p.small_out[i] = s;
        }
    }
}


// ---------------------------------------------------------------------------
// Layout-preserving all-reduce over peer memory (the one-pass step: its accumulator is consumed channel-last by
// cs_head_postmix, so nothing has to be transposed).  The buffer is cut into 16-byte words dealt to the ranks
// in contiguous slices; the owner of a slice sums it over all ranks and stores the sum into every rank's
// output.  Two data paths:
//   * multimem (NVSwitch multicast objects, sm_90+): ONE multimem.ld_reduce.add.v4.f32 returns the sum of all
//     ranks' copies, reduced inside the switch, and ONE multimem.st.v4.f32 broadcasts it: every byte crosses
//     this GPU's links once in each direction whatever the world size;
//   * peer pointers: world 16-byte ld.volatile loads in rank order and world 16-byte stores.
// ---------------------------------------------------------------------------
struct PeerFlatParams {
    const float* acc[CS_MAX_PEERS];
    float* out[CS_MAX_PEERS];
    const float* small_in[CS_MAX_PEERS];
    const float* acc_mc;        // multicast address of the accumulators (nullptr: peer-pointer path)
    float* out_mc;              // multicast address of the outputs
    float* small_out;
    int world, rank, small_n;
    long long n4;               // 16-byte words in the buffer
    long long begin4, end4;     // this rank's slice
};

__device__ __forceinline__ float4 multimem_ld_reduce_v4(const float* p) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void multimem_st_v4(float* p, const float4& v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <bool MULTIMEM>
__global__ void __launch_bounds__(256) cs_peer_flat_reduce_kernel(const PeerFlatParams p) {
    if (blockIdx.x + 1 == gridDim.x && p.small_n > 0) {
        // the last block sums the small vector (head gradients, loss) of all ranks for this rank
        for (int i = threadIdx.x; i < p.small_n; i += 256) {
            float s = 0.f;
            for (int r = 0; r < p.world; ++r) s += ld_peer(p.small_in[r] + i);
            p.small_out[i] = s;
        }
        return;
    }
    const long long stride = (long long)(gridDim.x - (p.small_n > 0 ? 1 : 0)) * 256;
    for (long long w = p.begin4 + (long long)blockIdx.x * 256 + threadIdx.x; w < p.end4; w += stride) {
        if (MULTIMEM) {
            const float4 v = multimem_ld_reduce_v4(p.acc_mc + 4 * w);
            multimem_st_v4(p.out_mc + 4 * w, v);
        } else {
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int r = 0; r < p.world; ++r) {
                const float4 v = ld_peer_v4(p.acc[r] + 4 * w);
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
            for (int r = 0; r < p.world; ++r) *reinterpret_cast<float4*>(p.out[r] + 4 * w) = s;
        }
    }
}

int layout_launch(const float* src, float* dst, int N, int C, long long T, int mode, void* stream) {
    if (N < 0 || C < 0 || T < 0) return fail(CS_EINVAL, "negative size");
    if (N == 0 || C == 0 || T == 0) return 0;
    if (!src || !dst) return fail(CS_EINVAL, "NULL buffer");
    const long long bx = (T + 31) / 32;
    if (bx > 0x7fffffffll || N > 65535 || (C + 31) / 32 > 65535)
        return fail(CS_EUNSUPPORTED, "layout kernel grid too large");
    dim3 grid((unsigned)bx, (unsigned)((C + 31) / 32), (unsigned)N);
    cudaStream_t s = (cudaStream_t)stream;
    if (mode == 0) cs_layout_kernel<true, false><<<grid, 256, 0, s>>>(src, dst, C, T);
    else if (mode == 1) cs_layout_kernel<false, false><<<grid, 256, 0, s>>>(src, dst, C, T);
    else cs_layout_kernel<false, true><<<grid, 256, 0, s>>>(src, dst, C, T);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "layout kernel launch");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

// Jet operator (cs_jet.cuh): validation shared by cs_jet_forward / cs_jet_backward.
int jet_run(const cs_problem* pb, int order, bool backward, const float* field_in, float* field_out,
            const float* coords, const float* offset, float* jets_out, const float* jets_in, void* stream) {
    if (!pb) return fail(CS_EINVAL, "cs_problem is NULL");
    if (pb->dim != 2 && pb->dim != 3) return fail(CS_EINVAL, "dim must be 2 or 3, got %d", pb->dim);
    if (order < 1 || order > 3) return fail(CS_EINVAL, "jet order must be 1, 2 or 3, got %d", order);
    if (pb->N < 0 || pb->C < 0 || pb->P < 0) return fail(CS_EINVAL, "negative size");
    if (pb->H < 1 || pb->W < 1 || pb->D < 1) return fail(CS_EINVAL, "cell extent must be >= 1");
    if (pb->dim == 2 && pb->D != 1) return fail(CS_EINVAL, "D must be 1 when dim == 2");
    if (pb->padding_mode < 0 || pb->padding_mode > 2) return fail(CS_EINVAL, "bad padding_mode %d", pb->padding_mode);
    if (pb->kernel < 0 || pb->kernel > 2) return fail(CS_EINVAL, "bad kernel %d", pb->kernel);
    if (pb->index_mode < 0 || pb->index_mode > 1) return fail(CS_EINVAL, "bad index_mode %d", pb->index_mode);
    if (pb->N == 0 || pb->C == 0 || pb->P == 0) return 0;
    if (pb->field_layout != CS_LAYOUT_CHANNEL_LAST)
        return fail(CS_EUNSUPPORTED, "the jet operator needs channel-last fields (cs_to_channel_last)");
    const int v = pb->C / 4;
    if (pb->C % 4 != 0 || !(v == 1 || v == 2 || v == 4 || v == 8))
        return fail(CS_EUNSUPPORTED, "the jet operator needs C in {4, 8, 16, 32}, got %d", pb->C);
    if (!coords || !offset) return fail(CS_EINVAL, "coords/offset pointer is NULL");
    const float* field = backward ? field_out : field_in;
    const float* rows = backward ? jets_in : jets_out;
    if (!field || !rows) return fail(CS_EINVAL, "field/jets pointer is NULL");
    if (!aligned16(field)) return fail(CS_EINVAL, "field must be 16-byte aligned");
    const long long T = (long long)pb->D * pb->H * pb->W;
    if (T * (long long)pb->C >= (1ll << 31))
        return fail(CS_EUNSUPPORTED, "a cell has %lld elements; the per-cell index is 32-bit", T * pb->C);
    if (pb->P >= (1ll << 33)) return fail(CS_EUNSUPPORTED, "too many points (%lld)", (long long)pb->P);
    cs::JetParams p;
    memset(&p, 0, sizeof(p));
    p.N = pb->N; p.C = pb->C; p.P = pb->P;
    p.size[0] = pb->W; p.size[1] = pb->H; p.size[2] = pb->D;
    p.tstride[0] = 1; p.tstride[1] = pb->W; p.tstride[2] = pb->W * pb->H;
    p.cell_stride = T * pb->C;
    p.V = field_in; p.acc = field_out; p.coords = coords; p.offset = offset;
    p.jets = jets_out; p.gjets = jets_in;
    p.svec = (p.P % 4 == 0) && aligned16(rows);
    p.cvec2 = (reinterpret_cast<uintptr_t>(coords) & 7u) == 0;
    p.pad = pb->padding_mode; p.align = pb->align_corners; p.kernel = pb->kernel;
    p.multicell = pb->multicell; p.index_mode = pb->index_mode;
    using Fn = cudaError_t (*)(int, bool, cs::JetParams&, cudaStream_t);
    static const Fn table[2][4] = {
        {cs::launch_jet_d2_l0, cs::launch_jet_d2_l1, cs::launch_jet_d2_l2, cs::launch_jet_d2_l3},
        {cs::launch_jet_d3_l0, cs::launch_jet_d3_l1, cs::launch_jet_d3_l2, cs::launch_jet_d3_l3}};
    const int lshift = (v == 1) ? 0 : (v == 2) ? 1 : (v == 4) ? 2 : 3;
    cudaError_t e = table[pb->dim - 2][lshift](order, backward, p, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "jet kernel launch");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}


// ---------------------------------------------------------------------------
// Fused step (cs_fused.cuh): point binning, W1 mixes, the one-pass kernel
// ---------------------------------------------------------------------------
int bin_setup(const cs_problem* pb, cs::BinParams& b) {
    if (!pb) return fail(CS_EINVAL, "cs_problem is NULL");
    if (pb->dim != 2 && pb->dim != 3) return fail(CS_EINVAL, "dim must be 2 or 3, got %d", pb->dim);
    if (pb->P < 0) return fail(CS_EINVAL, "negative size");
    if (pb->H < 1 || pb->W < 1 || pb->D < 1) return fail(CS_EINVAL, "cell extent must be >= 1");
    if (pb->dim == 2 && pb->D != 1) return fail(CS_EINVAL, "D must be 1 when dim == 2");
    if (pb->index_mode < 0 || pb->index_mode > 1) return fail(CS_EINVAL, "bad index_mode %d", pb->index_mode);
    if (pb->P >= (1ll << 31)) return fail(CS_EUNSUPPORTED, "cs_bin_points: at most 2^31 - 1 points per call (%lld)", (long long)pb->P);
    memset(&b, 0, sizeof(b));
    b.dim = pb->dim;
    b.size[0] = pb->W; b.size[1] = pb->H; b.size[2] = pb->D;
    b.P = pb->P;
    b.align = pb->align_corners; b.multicell = pb->multicell; b.index_mode = pb->index_mode;
    for (int shift = 0;; ++shift) {
        const int tl = (pb->dim == 2) ? 3 : 2;       // log2 tile edge
        const long long nx = ((((long long)pb->W - 1) >> shift) >> tl) + 1;
        const long long ny = ((((long long)pb->H - 1) >> shift) >> tl) + 1;
        const long long nz = (pb->dim == 3) ? ((((long long)pb->D - 1) >> shift) >> tl) + 1 : 1;
        const long long nb = nx * ny * nz * 64;
        if (nb <= (1ll << 21)) {
            b.shift = shift; b.ntx = (int)nx; b.nty = (int)ny; b.ntz = (int)nz; b.nbins = (unsigned)nb;
            break;
        }
    }
    // sub-texel bins (the cells of a multicell stack are offset by n/N of a texel: inside one sub-bin every
    // cell sees the same corners): as fine as the point density allows, at least 2 points per sub-bin (config 4,
    // 16 points per texel: 8 sub-bins per texel instead of none, one-pass kernel 1.72 -> 1.62 ms; 4, 1 and 0.5 measured)
    if (b.shift == 0) {
        const double density = (double)pb->P / ((double)pb->W * pb->H * pb->D);
        double minpts = 2.0;
        if (const char* e = getenv("COSINE_SAMPLER_BIN_MINPTS")) { const double v = atof(e); if (v > 0.0) minpts = v; }
        while (b.sub < 2 && density >= minpts * (double)(1ll << (pb->dim * (b.sub + 1))) &&
               ((long long)b.nbins << pb->dim) <= (1ll << 22)) {
            b.sub += 1;
            b.nbins <<= pb->dim;
        }
    }
    return 0;
}

inline long long round256(long long v) { return (v + 255) / 256 * 256; }

template <int K>
int premix_launch(const float* V, const float* W1, float* Vh, int C, long long T, long long NT, cudaStream_t s) {
    const size_t smem = (size_t)C * K * sizeof(float);
    long long blocks = (NT + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    cs::cs_head_premix_kernel<K><<<(unsigned)blocks, 256, smem, s>>>(V, W1, Vh, C, T, NT);
    return 0;
}

template <int K, bool KFIRST>
cudaError_t postmix_launch(const float* gVh, const float* V, const float* W1, float* gInput, int accumulate,
                           float* gW1, int N, int C, long long T, cudaStream_t s) {
    const long long tpc = (T + cs::POSTMIX_TT - 1) / cs::POSTMIX_TT;
    const long long ntiles = tpc * N;
    static const long long mult = [] { const char* e = getenv("CS_POSTMIX_BLOCKS"); return e ? atoll(e) : 4ll; }();
    long long blocks = ntiles < 148 * mult ? ntiles : 148 * mult;
    if (!KFIRST) {
        // channel-last gVh: the software-pipelined kernel (two stages of [TT][K] + [C][TT] in shared memory)
        auto kern = cs::cs_head_postmix_pipe_kernel<K>;
        const size_t smem = (size_t)(2 * (cs::POSTMIX_TT * K + C * (cs::POSTMIX_TT + 4)) + C * K) * sizeof(float);
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<(unsigned)blocks, cs::POSTMIX_TT, smem, s>>>(gVh, V, W1, gInput, accumulate, gW1, C, T, tpc, ntiles);
        return cudaGetLastError();
    }
    auto kern = cs::cs_head_postmix_kernel<K, KFIRST>;
    const size_t smem = (size_t)((K + C) * cs::POSTMIX_TS + C * K) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<(unsigned)blocks, cs::POSTMIX_TT, smem, s>>>(gVh, V, W1, gInput, accumulate, gW1, C, T, tpc, ntiles);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Double-precision path (cs_scalar.cuh)
// ---------------------------------------------------------------------------
template <typename S>
int scalar_setup(const cs_problem* pb, cs::ScalarParamsT<S>& p, const S* grid, const float* offset) {
    if (!pb) return fail(CS_EINVAL, "cs_problem is NULL");
    if (pb->dim != 2 && pb->dim != 3) return fail(CS_EINVAL, "dim must be 2 or 3, got %d", pb->dim);
    if (pb->N < 0 || pb->C < 0 || pb->P < 0) return fail(CS_EINVAL, "negative size");
    if (pb->H < 1 || pb->W < 1 || pb->D < 1) return fail(CS_EINVAL, "cell extent must be >= 1");
    if (pb->dim == 2 && pb->D != 1) return fail(CS_EINVAL, "D must be 1 when dim == 2");
    if (pb->padding_mode < 0 || pb->padding_mode > 2) return fail(CS_EINVAL, "bad padding_mode %d", pb->padding_mode);
    if (pb->kernel < 0 || pb->kernel > 2) return fail(CS_EINVAL, "bad kernel %d", pb->kernel);
    if (pb->field_layout != CS_LAYOUT_CHANNEL_FIRST)
        return fail(CS_EUNSUPPORTED, "the double / half precision paths work on the reference (channel-first) layout");
    if (pb->N == 0 || pb->C == 0 || pb->P == 0) return CS_NOTHING_TO_DO;
    if (!grid || !offset) return fail(CS_EINVAL, "grid/offset pointer is NULL");
    const long long T = (long long)pb->D * pb->H * pb->W;
    if (T >= (1ll << 31)) return fail(CS_EUNSUPPORTED, "a cell has %lld texels; the texel index is 32-bit", T);
    memset(&p, 0, sizeof(p));
    p.N = pb->N; p.C = pb->C; p.P = pb->P; p.T = T;
    p.size[0] = pb->W; p.size[1] = pb->H; p.size[2] = pb->D;
    p.tstride[0] = 1; p.tstride[1] = pb->W; p.tstride[2] = pb->W * pb->H;
    p.grid = grid; p.grid_sn = pb->grid_stride_n; p.offset = offset;
    p.pad = pb->padding_mode; p.align = pb->align_corners; p.kernel = pb->kernel; p.multicell = pb->multicell;
    return 0;
}

template <typename T>
int scalar_run(const cs_problem* pb, const cs::ScalarParamsT<T>& p, int stage, void* stream) {
    cudaError_t e = (pb->dim == 2) ? cs::launch_scalar<2, T>(stage, p, (cudaStream_t)stream)
                                   : cs::launch_scalar<3, T>(stage, p, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "double / half precision stage kernel launch");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}


// fp16: gInput accumulates in a float workspace (zeroed here) and is rounded to half after the stage kernel
int half_begin(const cs_problem* pb, const void* gInput, float* workspace, const char* who, void* stream) {
    if (!gInput) return 0;
    if (!workspace) return fail(CS_EINVAL, "%s: gInput needs the float workspace (N*C*D*H*W floats)", who);
    const size_t n = (size_t)pb->N * pb->C * pb->D * pb->H * pb->W;
    cudaError_t e = cudaMemsetAsync(workspace, 0, n * sizeof(float), (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "fp16 workspace memset");
    return 0;
}
int half_finish(const cs_problem* pb, cs_half* gInput, const float* workspace, void* stream) {
    if (!gInput) return 0;
    const long long n = (long long)pb->N * pb->C * pb->D * pb->H * pb->W;
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) return 0;
    cs::cs_f32_to_f16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(workspace, reinterpret_cast<__half*>(gInput), n);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "fp16 conversion kernel launch");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}
inline const __half* hp(const cs_half* p) { return reinterpret_cast<const __half*>(p); }
inline __half* hp(cs_half* p) { return reinterpret_cast<__half*>(p); }

}  // namespace

extern "C" {

int cs_version(void) { return CS_VERSION; }
const char* cs_last_error(void) { return g_err; }
uint64_t cs_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int cs_forward(const cs_problem* pb, const float* input, const float* grid, const float* offset,
               float* out, void* stream) {
    cs::StageParams p;
    if (int rc = setup(pb, p, grid, offset)) return rc == CS_NOTHING_TO_DO ? 0 : rc;
    if (!input || !out) return fail(CS_EINVAL, "cs_forward: input/out is NULL");
    p.V = input; p.y = out;
    return run(pb, p, cs::ST_F, false, false, stream);
}

int cs_backward(const cs_problem* pb, cs_stream gOut, const float* input, const float* grid,
                const float* offset, float* gInput, float* gGrid, void* stream) {
    cs::StageParams p;
    if (int rc = setup(pb, p, grid, offset)) return rc == CS_NOTHING_TO_DO ? 0 : rc;
    if (!gOut.ptr) return fail(CS_EINVAL, "cs_backward: gOut is NULL");
    if (gGrid && !input) return fail(CS_EINVAL, "cs_backward: gGrid needs input");
    if (!gInput && !gGrid) return 0;
    p.V = input; p.acc = gInput; p.ggrid = gGrid;
    p.x1 = gOut.ptr; p.x1_sn = gOut.stride_n; p.x1_sc = gOut.stride_c;
    if (pb->grad_order == 1 && gGrid) {
        // gGrid in reference order by the per-pair kernel, gInput (if wanted) by the regular one
        if (int rc = launch_reforder(pb, p, stream)) return rc;
        if (!gInput) return 0;
        p.ggrid = nullptr;
    }
    return run(pb, p, cs::ST_B, false, false, stream);
}

int cs_backward_backward(const cs_problem* pb, const float* gOutInput, const float* gOutGrid,
                         const float* input, const float* grid, cs_stream gOut, const float* offset,
                         float* gInput, float* gGrid, float* ggOut, void* stream) {
    cs::StageParams p;
    if (int rc = setup(pb, p, grid, offset)) return rc == CS_NOTHING_TO_DO ? 0 : rc;
    if (!gOut.ptr || !gOutGrid) return fail(CS_EINVAL, "cs_backward_backward: gOut/gOutGrid is NULL");
    if ((gGrid || ggOut) && !input) return fail(CS_EINVAL, "cs_backward_backward: gGrid/ggOut need input");
    if (!gInput && !gGrid && !ggOut) return 0;
    // gOutInput only feeds ggOut (2D and 3D) and gGrid (3D): skip its gather when neither is wanted
    const bool has_u = gOutInput && (ggOut || (gGrid && pb->dim == 3));
    p.V = input; p.U = has_u ? gOutInput : nullptr; p.acc = gInput; p.ggrid = gGrid; p.y = ggOut;
    p.gog = gOutGrid;
    p.x1 = gOut.ptr; p.x1_sn = gOut.stride_n; p.x1_sc = gOut.stride_c;
    return run(pb, p, cs::ST_BB, has_u, false, stream);
}

int cs_backward_backward_backward(const cs_problem* pb, const float* input, const float* grid,
                                  cs_stream gOut, const float* gOutGrid, const float* gOutgGrid,
                                  cs_stream gOutggOut, const float* offset, float* gInput,
                                  float* ggOut, void* stream) {
    cs::StageParams p;
    if (int rc = setup(pb, p, grid, offset)) return rc == CS_NOTHING_TO_DO ? 0 : rc;
    if (!gOut.ptr || !gOutGrid || !gOutgGrid)
        return fail(CS_EINVAL, "cs_backward_backward_backward: gOut/gOutGrid/gOutgGrid is NULL");
    if (ggOut && !input) return fail(CS_EINVAL, "cs_backward_backward_backward: ggOut needs input");
    if (!gInput && !ggOut) return 0;
    const bool has_x2 = gOutggOut.ptr && gInput;
    p.V = input; p.acc = gInput; p.y = ggOut;
    p.gog = gOutGrid; p.gogg = gOutgGrid;
    p.x1 = gOut.ptr; p.x1_sn = gOut.stride_n; p.x1_sc = gOut.stride_c;
    if (has_x2) { p.x2 = gOutggOut.ptr; p.x2_sn = gOutggOut.stride_n; p.x2_sc = gOutggOut.stride_c; }
    return run(pb, p, cs::ST_BBB, false, has_x2, stream);
}

int cs_forward_f64(const cs_problem* pb, const double* input, const double* grid, const float* offset, double* out,
                   void* stream) {
    cs::ScalarParams p;
    if (int rc = scalar_setup(pb, p, grid, offset)) return rc == CS_NOTHING_TO_DO ? 0 : rc;
    if (!input || !out) return fail(CS_EINVAL, "cs_forward_f64: input/out is NULL");
    p.V = input; p.y = out;
    return scalar_run(pb, p, cs::ST_F, stream);
}

int cs_backward_f64(const cs_problem* pb, cs_stream_f64 gOut, const double* input, const double* grid,
                    const float* offset, double* gInput, double* gGrid, void* stream) {
    cs::ScalarParams p;
    if (int rc = scalar_setup(pb, p, grid, offset)) return rc == CS_NOTHING_TO_DO ? 0 : rc;
    if (!gOut.ptr) return fail(CS_EINVAL, "cs_backward_f64: gOut is NULL");
    if (gGrid && !input) return fail(CS_EINVAL, "cs_backward_f64: gGrid needs input");
    if (!gInput && !gGrid) return 0;
    p.V = input; p.acc = gInput; p.ggrid = gGrid;
    p.x1 = gOut.ptr; p.x1_sn = gOut.stride_n; p.x1_sc = gOut.stride_c;
    return scalar_run(pb, p, cs::ST_B, stream);
}

int cs_backward_backward_f64(const cs_problem* pb, const double* gOutInput, const double* gOutGrid,
                             const double* input, const double* grid, cs_stream_f64 gOut, const float* offset,
                             double* gInput, double* gGrid, double* ggOut, void* stream) {
    cs::ScalarParams p;
    if (int rc = scalar_setup(pb, p, grid, offset)) return rc == CS_NOTHING_TO_DO ? 0 : rc;
    if (!gOut.ptr || !gOutGrid) return fail(CS_EINVAL, "cs_backward_backward_f64: gOut/gOutGrid is NULL");
    if ((gGrid || ggOut) && !input) return fail(CS_EINVAL, "cs_backward_backward_f64: gGrid/ggOut need input");
    if (!gInput && !gGrid && !ggOut) return 0;
    p.V = input; p.U = gOutInput; p.acc = gInput; p.ggrid = gGrid; p.y = ggOut; p.gog = gOutGrid;
    p.x1 = gOut.ptr; p.x1_sn = gOut.stride_n; p.x1_sc = gOut.stride_c;
    return scalar_run(pb, p, cs::ST_BB, stream);
}

int cs_backward_backward_backward_f64(const cs_problem* pb, const double* input, const double* grid,
                                      cs_stream_f64 gOut, const double* gOutGrid, const double* gOutgGrid,
                                      cs_stream_f64 gOutggOut, const float* offset, double* gInput, double* ggOut,
                                      void* stream) {
    cs::ScalarParams p;
    if (int rc = scalar_setup(pb, p, grid, offset)) return rc == CS_NOTHING_TO_DO ? 0 : rc;
    if (!gOut.ptr || !gOutGrid || !gOutgGrid)
        return fail(CS_EINVAL, "cs_backward_backward_backward_f64: gOut/gOutGrid/gOutgGrid is NULL");
    if (ggOut && !input) return fail(CS_EINVAL, "cs_backward_backward_backward_f64: ggOut needs input");
    if (!gInput && !ggOut) return 0;
    p.V = input; p.acc = gInput; p.y = ggOut; p.gog = gOutGrid; p.gogg = gOutgGrid;
    p.x1 = gOut.ptr; p.x1_sn = gOut.stride_n; p.x1_sc = gOut.stride_c;
    if (gOutggOut.ptr && gInput) { p.x2 = gOutggOut.ptr; p.x2_sn = gOutggOut.stride_n; p.x2_sc = gOutggOut.stride_c; }
    return scalar_run(pb, p, cs::ST_BBB, stream);
}

int cs_forward_f16(const cs_problem* pb, const cs_half* input, const cs_half* grid, const float* offset, cs_half* out,
                   void* stream) {
    cs::ScalarParamsT<__half> p;
    if (int rc = scalar_setup<__half>(pb, p, hp(grid), offset)) return rc == CS_NOTHING_TO_DO ? 0 : rc;
    if (!input || !out) return fail(CS_EINVAL, "cs_forward_f16: input/out is NULL");
    p.V = hp(input); p.y = hp(out);
    return scalar_run(pb, p, cs::ST_F, stream);
}

int cs_backward_f16(const cs_problem* pb, cs_stream_f16 gOut, const cs_half* input, const cs_half* grid,
                    const float* offset, cs_half* gInput, cs_half* gGrid, float* workspace, void* stream) {
    cs::ScalarParamsT<__half> p;
    if (int rc = scalar_setup<__half>(pb, p, hp(grid), offset)) return rc == CS_NOTHING_TO_DO ? 0 : rc;
    if (!gOut.ptr) return fail(CS_EINVAL, "cs_backward_f16: gOut is NULL");
    if (gGrid && !input) return fail(CS_EINVAL, "cs_backward_f16: gGrid needs input");
    if (!gInput && !gGrid) return 0;
    if (int rc = half_begin(pb, gInput, workspace, "cs_backward_f16", stream)) return rc;
    p.V = hp(input); p.acc = gInput ? workspace : nullptr; p.ggrid = hp(gGrid);
    p.x1 = hp(gOut.ptr); p.x1_sn = gOut.stride_n; p.x1_sc = gOut.stride_c;
    if (int rc = scalar_run(pb, p, cs::ST_B, stream)) return rc;
    return half_finish(pb, gInput, workspace, stream);
}

int cs_backward_backward_f16(const cs_problem* pb, const cs_half* gOutInput, const cs_half* gOutGrid,
                             const cs_half* input, const cs_half* grid, cs_stream_f16 gOut, const float* offset,
                             cs_half* gInput, cs_half* gGrid, cs_half* ggOut, float* workspace, void* stream) {
    cs::ScalarParamsT<__half> p;
    if (int rc = scalar_setup<__half>(pb, p, hp(grid), offset)) return rc == CS_NOTHING_TO_DO ? 0 : rc;
    if (!gOut.ptr || !gOutGrid) return fail(CS_EINVAL, "cs_backward_backward_f16: gOut/gOutGrid is NULL");
    if ((gGrid || ggOut) && !input) return fail(CS_EINVAL, "cs_backward_backward_f16: gGrid/ggOut need input");
    if (!gInput && !gGrid && !ggOut) return 0;
    if (int rc = half_begin(pb, gInput, workspace, "cs_backward_backward_f16", stream)) return rc;
    p.V = hp(input); p.U = hp(gOutInput); p.acc = gInput ? workspace : nullptr; p.ggrid = hp(gGrid); p.y = hp(ggOut);
    p.gog = hp(gOutGrid);
    p.x1 = hp(gOut.ptr); p.x1_sn = gOut.stride_n; p.x1_sc = gOut.stride_c;
    if (int rc = scalar_run(pb, p, cs::ST_BB, stream)) return rc;
    return half_finish(pb, gInput, workspace, stream);
}

int cs_backward_backward_backward_f16(const cs_problem* pb, const cs_half* input, const cs_half* grid,
                                      cs_stream_f16 gOut, const cs_half* gOutGrid, const cs_half* gOutgGrid,
                                      cs_stream_f16 gOutggOut, const float* offset, cs_half* gInput, cs_half* ggOut,
                                      float* workspace, void* stream) {
    cs::ScalarParamsT<__half> p;
    if (int rc = scalar_setup<__half>(pb, p, hp(grid), offset)) return rc == CS_NOTHING_TO_DO ? 0 : rc;
    if (!gOut.ptr || !gOutGrid || !gOutgGrid)
        return fail(CS_EINVAL, "cs_backward_backward_backward_f16: gOut/gOutGrid/gOutgGrid is NULL");
    if (ggOut && !input) return fail(CS_EINVAL, "cs_backward_backward_backward_f16: ggOut needs input");
    if (!gInput && !ggOut) return 0;
    if (int rc = half_begin(pb, gInput, workspace, "cs_backward_backward_backward_f16", stream)) return rc;
    p.V = hp(input); p.acc = gInput ? workspace : nullptr; p.y = hp(ggOut); p.gog = hp(gOutGrid); p.gogg = hp(gOutgGrid);
    p.x1 = hp(gOut.ptr); p.x1_sn = gOut.stride_n; p.x1_sc = gOut.stride_c;
    if (gOutggOut.ptr && gInput) { p.x2 = hp(gOutggOut.ptr); p.x2_sn = gOutggOut.stride_n; p.x2_sc = gOutggOut.stride_c; }
    if (int rc = scalar_run(pb, p, cs::ST_BBB, stream)) return rc;
    return half_finish(pb, gInput, workspace, stream);
}

int cs_jet_forward(const cs_problem* pb, int32_t order, const float* input, const float* coords,
                   const float* offset, float* jets, void* stream) {
    return jet_run(pb, order, false, input, nullptr, coords, offset, jets, nullptr, stream);
}

int cs_jet_backward(const cs_problem* pb, int32_t order, const float* gJets, const float* coords,
                    const float* offset, float* gInput, void* stream) {
    return jet_run(pb, order, true, nullptr, gInput, coords, offset, nullptr, gJets, stream);
}

int cs_pde_head_step(int32_t dim, int32_t C, int64_t P, const float* jets, const float* W1, const float* b1,
                     const float* w2, const float* b2, const cs_pde_residual* res, float scale,
                     float* gJets, float* gW1, float* gb1, float* gw2, float* gb2, float* loss_sum,
                     float* f_out, void* stream) {
    if (dim != 2 && dim != 3) return fail(CS_EINVAL, "dim must be 2 or 3, got %d", dim);
    if (!(C == 4 || C == 8 || C == 16 || C == 32))
        return fail(CS_EUNSUPPORTED, "cs_pde_head_step needs C in {4, 8, 16, 32}, got %d", C);
    if (P < 0) return fail(CS_EINVAL, "negative size");
    if (P == 0) return 0;
    if (!jets || !W1 || !b1 || !w2 || !b2 || !res || !gJets || !gW1 || !gb1 || !gw2 || !gb2 || !loss_sum)
        return fail(CS_EINVAL, "cs_pde_head_step: NULL pointer");
    cs::HeadParams p;
    memset(&p, 0, sizeof(p));
    p.P = P; p.jets = jets; p.gjets = gJets;
    p.W1 = W1; p.b1 = b1; p.w2 = w2; p.b2 = b2;
    p.gW1 = gW1; p.gb1 = gb1; p.gw2 = gw2; p.gb2 = gb2; p.loss_sum = loss_sum; p.f_out = f_out;
    p.c_u = res->c_u; p.c_u3 = res->c_u3;
    for (int a = 0; a < 3; ++a) { p.c1[a] = res->c1[a]; p.c2[a] = res->c2[a]; }
    p.scale = scale;
    p.vec = (P % 4 == 0) && aligned16(jets) && aligned16(gJets);
    cudaError_t e = cs::launch_head_any(dim, C, p, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "pde head kernel launch");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}


int cs_bin_workspace_bytes(const cs_problem* pb, int64_t* bytes) {
    cs::BinParams b;
    if (int rc = bin_setup(pb, b)) return rc;
    if (!bytes) return fail(CS_EINVAL, "cs_bin_workspace_bytes: bytes is NULL");
    *bytes = round256((long long)b.nbins * 4) + round256(((long long)b.nbins / cs::BIN_SCAN_CHUNK + 1) * 4) +
             2 * round256((long long)pb->P * 4);
    return 0;
}

int cs_bin_points(const cs_problem* pb, const float* coords, const float* offset, float* sorted, int32_t* perm,
                  void* workspace, int64_t workspace_bytes, void* stream) {
    cs::BinParams b;
    if (int rc = bin_setup(pb, b)) return rc;
    if (pb->P == 0) return 0;
    if (!coords || !sorted || !workspace) return fail(CS_EINVAL, "cs_bin_points: NULL pointer");
    if (coords == sorted) return fail(CS_EINVAL, "cs_bin_points: sorted must not alias coords");
    const long long hist_bytes = round256((long long)b.nbins * 4);
    const unsigned nchunks = (b.nbins + cs::BIN_SCAN_CHUNK - 1) / cs::BIN_SCAN_CHUNK;
    const long long tot_bytes = round256(((long long)b.nbins / cs::BIN_SCAN_CHUNK + 1) * 4);
    const long long pt_bytes = round256((long long)pb->P * 4);
    if (workspace_bytes < hist_bytes + tot_bytes + 2 * pt_bytes)
        return fail(CS_EINVAL, "cs_bin_points: workspace too small (see cs_bin_workspace_bytes)");
    b.coords = coords; b.offset = offset;
    unsigned* hist = reinterpret_cast<unsigned*>(workspace);
    unsigned* totals = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(workspace) + hist_bytes);
    unsigned* rank = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(workspace) + hist_bytes + tot_bytes);
    unsigned* keys = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(workspace) + hist_bytes + tot_bytes + pt_bytes);
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(hist, 0, (size_t)b.nbins * 4, s);
    if (e != cudaSuccess) return cuda_fail(e, "cs_bin_points memset");
    long long blocks = (pb->P + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    cs::cs_bin_count_kernel<<<(unsigned)blocks, 256, 0, s>>>(b, hist, keys, rank);
    cs::cs_bin_scan_chunk_kernel<<<nchunks, 1024, 0, s>>>(hist, b.nbins, totals);
    cs::cs_bin_scan_totals_kernel<<<1, 1024, 0, s>>>(totals, nchunks);
    // destination windows of 2^23 points (64-96 MB: they stay in L2 while they fill up)
    const long long window = 1ll << 23;
    const int vec2 = ((reinterpret_cast<uintptr_t>(coords) | reinterpret_cast<uintptr_t>(sorted)) & 7u) == 0;
    int sweeps = 0;
    for (long long lo = 0; lo < pb->P; lo += window, ++sweeps) {
        const unsigned hi = (unsigned)((lo + window < pb->P) ? lo + window : pb->P);
        if (pb->dim == 2)
            cs::cs_bin_scatter_kernel<2><<<(unsigned)blocks, 256, 0, s>>>(coords, pb->P, vec2, hist, totals, keys, rank,
                                                                         sweeps == 0 ? 1 : 0, (unsigned)lo, hi, sorted, perm);
        else
            cs::cs_bin_scatter_kernel<3><<<(unsigned)blocks, 256, 0, s>>>(coords, pb->P, 0, hist, totals, keys, rank,
                                                                         sweeps == 0 ? 1 : 0, (unsigned)lo, hi, sorted, perm);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "cs_bin_points launch");
    g_launches.fetch_add(3 + sweeps, std::memory_order_relaxed);
    return 0;
}

static int mix_check(int32_t N, int32_t C, int64_t T, int32_t K) {
    if (N < 0 || C < 0 || T < 0) return fail(CS_EINVAL, "negative size");
    if (!(K == 4 || K == 8 || K == 16 || K == 32 || K == 64))
        return fail(CS_EUNSUPPORTED, "hidden width must be 4, 8, 16, 32 or 64, got %d", K);
    if (C > 64) return fail(CS_EUNSUPPORTED, "the W1 mixes support at most 64 input channels, got %d", C);
    return 0;
}

int cs_head_premix(int32_t N, int32_t C, int64_t T, int32_t K, const float* input, const float* W1, float* Vh,
                   void* stream) {
    if (int rc = mix_check(N, C, T, K)) return rc;
    if (N == 0 || C == 0 || T == 0) return 0;
    if (!input || !W1 || !Vh) return fail(CS_EINVAL, "cs_head_premix: NULL pointer");
    if (!aligned16(Vh)) return fail(CS_EINVAL, "cs_head_premix: Vh must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    const long long NT = (long long)N * T;
    switch (K) {
        case 4: premix_launch<4>(input, W1, Vh, C, T, NT, s); break;
        case 8: premix_launch<8>(input, W1, Vh, C, T, NT, s); break;
        case 16: premix_launch<16>(input, W1, Vh, C, T, NT, s); break;
        case 32: premix_launch<32>(input, W1, Vh, C, T, NT, s); break;
        default: premix_launch<64>(input, W1, Vh, C, T, NT, s); break;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "cs_head_premix launch");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

int cs_head_postmix(int32_t N, int32_t C, int64_t T, int32_t K, const float* gVh, int32_t hidden_first,
                    const float* input, const float* W1, float* gInput, int32_t accumulate, float* gW1,
                    void* stream) {
    if (int rc = mix_check(N, C, T, K)) return rc;
    if (N == 0 || C == 0 || T == 0) return 0;
    if (!gVh || !input || !W1) return fail(CS_EINVAL, "cs_head_postmix: NULL pointer");
    if (!gInput && !gW1) return 0;
    if (!hidden_first && !aligned16(gVh)) return fail(CS_EINVAL, "cs_head_postmix: gVh must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e;
#define CS_POSTMIX(KK) e = hidden_first ? postmix_launch<KK, true>(gVh, input, W1, gInput, accumulate, gW1, N, C, T, s) \
                                        : postmix_launch<KK, false>(gVh, input, W1, gInput, accumulate, gW1, N, C, T, s)
    switch (K) {
        case 4: CS_POSTMIX(4); break;
        case 8: CS_POSTMIX(8); break;
        case 16: CS_POSTMIX(16); break;
        case 32: CS_POSTMIX(32); break;
        default: CS_POSTMIX(64); break;
    }
#undef CS_POSTMIX
    if (e != cudaSuccess) return cuda_fail(e, "cs_head_postmix launch");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

int cs_pde_fused_step(const cs_problem* pb, const float* Vh, const float* coords, const float* offset,
                      const float* b1, const float* w2, const float* b2, const cs_pde_residual* res, float scale,
                      float* gVh, float* gb1, float* gw2, float* gb2, float* loss_sum, int32_t aggregate,
                      void* stream) {
    if (!pb) return fail(CS_EINVAL, "cs_problem is NULL");
    if (pb->dim != 2 && pb->dim != 3) return fail(CS_EINVAL, "dim must be 2 or 3, got %d", pb->dim);
    if (pb->N < 0 || pb->C < 0 || pb->P < 0) return fail(CS_EINVAL, "negative size");
    if (pb->H < 1 || pb->W < 1 || pb->D < 1) return fail(CS_EINVAL, "cell extent must be >= 1");
    if (pb->dim == 2 && pb->D != 1) return fail(CS_EINVAL, "D must be 1 when dim == 2");
    if (pb->padding_mode < 0 || pb->padding_mode > 2) return fail(CS_EINVAL, "bad padding_mode %d", pb->padding_mode);
    if (pb->kernel < 0 || pb->kernel > 2) return fail(CS_EINVAL, "bad kernel %d", pb->kernel);
    if (pb->index_mode < 0 || pb->index_mode > 1) return fail(CS_EINVAL, "bad index_mode %d", pb->index_mode);
    if (aggregate < 0 || aggregate > 1) return fail(CS_EINVAL, "aggregate must be 0 or 1");
    if (pb->N == 0 || pb->C == 0 || pb->P == 0) return 0;
    if (pb->field_layout != CS_LAYOUT_CHANNEL_LAST)
        return fail(CS_EUNSUPPORTED, "cs_pde_fused_step needs the channel-last mixed cells of cs_head_premix");
    const int K = pb->C;
    if (!(K == 4 || K == 8 || K == 16 || K == 32 || K == 64))
        return fail(CS_EUNSUPPORTED, "hidden width must be 4, 8, 16, 32 or 64, got %d", K);
    if (!Vh || !coords || !offset || !b1 || !w2 || !b2 || !res || !gVh || !gb1 || !gw2 || !gb2 || !loss_sum)
        return fail(CS_EINVAL, "cs_pde_fused_step: NULL pointer");
    if (!aligned16(Vh) || !aligned16(gVh)) return fail(CS_EINVAL, "cs_pde_fused_step: Vh / gVh must be 16-byte aligned");
    const long long T = (long long)pb->D * pb->H * pb->W;
    if (((long long)pb->N * T + 1) * (long long)K >= (1ll << 31))
        return fail(CS_EUNSUPPORTED, "the mixed cells have %lld elements; the texel index is 32-bit", ((long long)pb->N * T + 1) * K);
    if (pb->P >= (1ll << 33)) return fail(CS_EUNSUPPORTED, "too many points (%lld)", (long long)pb->P);
    if (pb->N > cs::FUSED_MAX_CELLS)
        return fail(CS_EUNSUPPORTED, "cs_pde_fused_step holds the records of all cells in shared memory: at most %d cells, got %d",
                    cs::FUSED_MAX_CELLS, pb->N);
    cs::FusedParams p;
    memset(&p, 0, sizeof(p));
    p.N = pb->N; p.P = pb->P; p.T = T;
    p.size[0] = pb->W; p.size[1] = pb->H; p.size[2] = pb->D;
    p.tstride[0] = 1; p.tstride[1] = pb->W; p.tstride[2] = pb->W * pb->H;
    p.Vh = Vh; p.gVh = gVh; p.coords = coords; p.offset = offset;
    p.b1 = b1; p.w2 = w2; p.b2 = b2;
    p.gb1 = gb1; p.gw2 = gw2; p.gb2 = gb2; p.loss_sum = loss_sum;
    p.c_u = res->c_u; p.c_u3 = res->c_u3;
    for (int a = 0; a < 3; ++a) { p.c1[a] = res->c1[a]; p.c2[a] = res->c2[a]; }
    p.scale = scale;
    p.cvec2 = (reinterpret_cast<uintptr_t>(coords) & 7u) == 0;
    p.pad = pb->padding_mode; p.align = pb->align_corners; p.kernel = pb->kernel;
    p.multicell = pb->multicell; p.index_mode = pb->index_mode;
    const int v = K / 4;
    const int lshift = (v == 1) ? 0 : (v == 2) ? 1 : (v == 4) ? 2 : (v == 8) ? 3 : 4;
    p.aggregate = aggregate ? 1 : 0;
    using Fn = cudaError_t (*)(cs::FusedParams&, cudaStream_t);
    static const Fn table[2][5] = {
        {cs::launch_fused_d2_l0, cs::launch_fused_d2_l1, cs::launch_fused_d2_l2, cs::launch_fused_d2_l3, cs::launch_fused_d2_l4},
        {cs::launch_fused_d3_l0, cs::launch_fused_d3_l1, cs::launch_fused_d3_l2, cs::launch_fused_d3_l3, cs::launch_fused_d3_l4}};
    cudaError_t e = table[pb->dim - 2][lshift](p, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "fused step kernel launch");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

int cs_peer_allreduce_from_channel_last(int32_t world, int32_t rank, const float* const* acc_ptrs,
                                        float* const* out_ptrs, int32_t N, int32_t C, int64_t T,
                                        const float* const* small_ptrs, float* small_out, int32_t small_n,
                                        void* stream) {
    if (world < 1 || world > CS_MAX_PEERS) return fail(CS_EUNSUPPORTED, "world must be 1..%d, got %d", CS_MAX_PEERS, world);
    if (rank < 0 || rank >= world) return fail(CS_EINVAL, "bad rank %d of %d", rank, world);
    if (N < 0 || C < 0 || T < 0 || small_n < 0) return fail(CS_EINVAL, "negative size");
    if (!acc_ptrs || !out_ptrs) return fail(CS_EINVAL, "NULL pointer table");
    if (small_n > 0 && (!small_ptrs || !small_out)) return fail(CS_EINVAL, "small vector given without pointers");
    PeerParams p;
    memset(&p, 0, sizeof(p));
    for (int r = 0; r < world; ++r) {
        if (!acc_ptrs[r] || !out_ptrs[r]) return fail(CS_EINVAL, "NULL peer buffer for rank %d", r);
        p.acc[r] = acc_ptrs[r]; p.out[r] = out_ptrs[r];
        if (small_n > 0) {
            if (!small_ptrs[r]) return fail(CS_EINVAL, "NULL small buffer for rank %d", r);
            p.small_in[r] = small_ptrs[r];
        }
    }
    p.small_out = small_out; p.small_n = small_n;
    p.vec4 = 1;
    for (int r = 0; r < world; ++r) if (!aligned16(acc_ptrs[r])) p.vec4 = 0;
    p.world = world; p.rank = rank; p.N = N; p.C = C; p.T = T;
    p.tiles_x = (T + 31) / 32; p.tiles_y = (C + 31) / 32;
    p.total_tiles = (long long)N * p.tiles_y * p.tiles_x;
    p.my_tiles = p.total_tiles > rank ? (p.total_tiles - rank + world - 1) / world : 0;
    const long long blocks = p.my_tiles + (small_n > 0 ? 1 : 0);
    if (blocks < 1) return 0;
    if (blocks > 0x7fffffffll) return fail(CS_EUNSUPPORTED, "peer reduce grid too large");
    cs_peer_reduce_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "peer reduce launch");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}


int cs_peer_allreduce(int32_t world, int32_t rank, const float* const* acc_ptrs, float* const* out_ptrs, int64_t n,
                      const float* acc_multicast, float* out_multicast, const float* const* small_ptrs,
                      float* small_out, int32_t small_n, void* stream) {
    if (world < 1 || world > CS_MAX_PEERS) return fail(CS_EUNSUPPORTED, "world must be 1..%d, got %d", CS_MAX_PEERS, world);
    if (rank < 0 || rank >= world) return fail(CS_EINVAL, "bad rank %d of %d", rank, world);
    if (n < 0 || small_n < 0) return fail(CS_EINVAL, "negative size");
    if (n % 4 != 0) return fail(CS_EINVAL, "cs_peer_allreduce: n must be a multiple of 4 floats, got %lld", (long long)n);
    if (!acc_ptrs || !out_ptrs) return fail(CS_EINVAL, "NULL pointer table");
    if (small_n > 0 && (!small_ptrs || !small_out)) return fail(CS_EINVAL, "small vector given without pointers");
    if ((acc_multicast == nullptr) != (out_multicast == nullptr))
        return fail(CS_EINVAL, "cs_peer_allreduce: give both multicast addresses or neither");
    PeerFlatParams p;
    memset(&p, 0, sizeof(p));
    for (int r = 0; r < world; ++r) {
        if (!acc_ptrs[r] || !out_ptrs[r]) return fail(CS_EINVAL, "NULL peer buffer for rank %d", r);
        if (!aligned16(acc_ptrs[r]) || !aligned16(out_ptrs[r])) return fail(CS_EINVAL, "peer buffers must be 16-byte aligned");
        p.acc[r] = acc_ptrs[r]; p.out[r] = out_ptrs[r];
        if (small_n > 0) {
            if (!small_ptrs[r]) return fail(CS_EINVAL, "NULL small buffer for rank %d", r);
            p.small_in[r] = small_ptrs[r];
        }
    }
    if (acc_multicast && (!aligned16(acc_multicast) || !aligned16(out_multicast)))
        return fail(CS_EINVAL, "multicast addresses must be 16-byte aligned");
    p.acc_mc = acc_multicast; p.out_mc = out_multicast;
    p.small_out = small_out; p.small_n = small_n;
    p.world = world; p.rank = rank; p.n4 = n / 4;
    const long long per = (p.n4 + world - 1) / world;
    p.begin4 = per * rank < p.n4 ? per * rank : p.n4;
    p.end4 = p.begin4 + per < p.n4 ? p.begin4 + per : p.n4;
    long long blocks = (p.end4 - p.begin4 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    blocks += (small_n > 0 ? 1 : 0);
    if (blocks < 1) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    if (acc_multicast) cs_peer_flat_reduce_kernel<true><<<(unsigned)blocks, 256, 0, s>>>(p);
    else cs_peer_flat_reduce_kernel<false><<<(unsigned)blocks, 256, 0, s>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "peer all-reduce launch");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

int cs_to_channel_last(const float* src, float* dst, int32_t N, int32_t C, int64_t T, void* stream) {
    return layout_launch(src, dst, N, C, T, 0, stream);
}

int cs_from_channel_last(const float* src, float* dst, int32_t N, int32_t C, int64_t T,
                         int32_t accumulate, void* stream) {
    return layout_launch(src, dst, N, C, T, accumulate ? 2 : 1, stream);
}

}  // extern "C"
