/*
 * cosine_sampler_b200.h -- C ABI of libcosine_sampler_b200.so
 *
 * Drop-in boundary for the hot path of NamGyuKang/CosineSampler: the four
 * entry points the reference exports from each of its two pybind modules
 * (`_cosine_2d`, `_cosine_3d`):
 *
 *   forward                     cosine_sampler_2d.cpp:47   cosine_sampler_3d.cpp:50
 *   backward                    cosine_sampler_2d.cpp:64   cosine_sampler_3d.cpp:67
 *   backward_backward           cosine_sampler_2d.cpp:87   cosine_sampler_3d.cpp:90
 *   backward_backward_backward  cosine_sampler_2d.cpp:108  cosine_sampler_3d.cpp:112
 *   (pybind tables: cosine_sampler_2d.cpp:130-135, cosine_sampler_3d.cpp:133-138)
 *
 * re-expressed with plain pointers and sizes: no torch types, no hidden
 * allocation.  The caller owns every buffer (the reference's extension
 * allocated its outputs with torch::empty / zeros_like, cpp2d:57,75,80,99-101,
 * 119-120; here the host layer does that and passes raw device pointers).
 *
 * All device buffers are fp32.  All functions launch on `stream` (a
 * cudaStream_t passed as void*) of the *current* device and return without
 * synchronising.  Return value: 0 on success, a negative CS_E* code for bad
 * arguments, a positive cudaError_t when a launch failed.  cs_last_error()
 * gives a thread-local description of the most recent failure.
 *
 * Shapes (reference layouts, contiguous unless a stride is given):
 *   input / gOutInput / gInput   [N, C, (D,) H, W]      "grid-shaped fields"
 *   grid / gOutGrid / gOutgGrid / gGrid   [N, P, dim]   P = points per cell
 *   out / gOut / ggOut / gOutggOut        [N, C, P]     "point streams"
 *   offset                                [N]
 * grid[...,0] runs along W, grid[...,1] along H, grid[...,2] along D
 * (cosine_sampler_2d_kernel.cu:304-308, cosine_sampler_3d_kernel.cu:295-301).
 */
#ifndef COSINE_SAMPLER_B200_H
#define COSINE_SAMPLER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CS_VERSION 200

/* padding_mode_enum, modules_2d.py:4-10 */
#define CS_PAD_ZEROS 0
#define CS_PAD_BORDER 1
#define CS_PAD_REFLECTION 2
/* kernel_enum, modules_2d.py:12-18 */
#define CS_KERNEL_COSINE 0
#define CS_KERNEL_LINEAR 1
#define CS_KERNEL_SMOOTHSTEP 2

/* Layout of the grid-shaped fields (input, gOutInput, gInput) of one call. */
#define CS_LAYOUT_CHANNEL_FIRST 0 /* [N, C, (D,) H, W]: the reference boundary layout   */
#define CS_LAYOUT_CHANNEL_LAST 1  /* [N, (D,) H, W, C]: staged copy, vector gathers/reds */

/* index_mode: how i = ((g+1)/2)*s + offset is rounded. */
#define CS_INDEX_SEPARATE 0 /* multiply, then add: what test/grid_sampler.py:37-38 does */
#define CS_INDEX_FUSED 1    /* one fma: what the reference's --use_fast_math build emits */

#define CS_EINVAL (-1)
#define CS_EUNSUPPORTED (-2)

typedef struct cs_problem {
    int32_t dim;           /* 2 or 3 */
    int32_t N;             /* cells (batch of input) */
    int32_t C;             /* channels */
    int32_t D, H, W;       /* extent of a cell; D = 1 when dim == 2 */
    int64_t P;             /* points per cell = prod(grid.shape[1:-1]) */
    int32_t padding_mode;  /* CS_PAD_* */
    int32_t align_corners; /* 0/1.  2D forward ignores it and uses 1, as cu2d:307-308 */
    int32_t kernel;        /* CS_KERNEL_* */
    int32_t multicell;     /* 0/1: shrink the index range by one cell (cu2d:57-59) */
    int32_t index_mode;    /* CS_INDEX_* */
    int32_t field_layout;  /* CS_LAYOUT_* of input, gOutInput and gInput in this call */
    int64_t grid_stride_n; /* elements between cells of `grid`; P*dim if contiguous, 0 if expanded */
    int32_t lanes;         /* 0 = auto; else 1,2,4,8 lanes cooperating on one point quad */
    int32_t small_cell;    /* 0 = auto: cells whose fields fit in shared memory take the shared-memory
                              kernel; 1 = never; 2 = whenever it fits */
    int32_t grad_order;    /* cs_backward only.  0 = fast (channels split over lanes, shuffle reduction);
                              1 = reference order: gGrid accumulated channel by channel, corner by corner
                              with the reference's operation order (cu2d:464-503, ATen GridSampler.cu),
                              one thread per (cell, point) -- bit-comparable with torch grid_sample */
    int32_t reserved;
} cs_problem;

/* Strided view of a [N, C, P] point stream whose P axis is contiguous.
 * PIXEL's `val.sum(0)` hands the op a gOut that is expanded along N
 * (stride_n == 0); the reference copies it (modules_2d.py:42), we read it in place. */
typedef struct cs_stream {
    const float *ptr;
    int64_t stride_n; /* elements */
    int64_t stride_c; /* elements */
} cs_stream;

int cs_version(void);
const char *cs_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
uint64_t cs_launch_count(void);

/* out[n,c,p] = sum_q input[n,c,corner q] * w_q            (cu2d:265-356, cu3d:250-371) */
int cs_forward(const cs_problem *pb, const float *input, const float *grid, const float *offset,
               float *out, void *stream);

/* gInput[n,c,corner q] += w_q gOut[n,c,p]      (nullable: the reference's input_requires_grad)
 * gGrid[n,p,a] = sum_c gOut sum_q V_q dw_q/dg_a (nullable)   (cu2d:359-507, cu3d:373-584)
 * gInput must be zero-initialised by the caller (cpp2d:75). */
int cs_backward(const cs_problem *pb, cs_stream gOut, const float *input, const float *grid,
                const float *offset, float *gInput, float *gGrid, void *stream);

/* (gInput, gGrid, ggOut) of cu2d:509-717 / cu3d:587-870.  gOutInput nullable (the
 * reference's input_requires_grad == false); each output nullable = not wanted.
 * gInput must be zero-initialised by the caller (cpp2d:99). */
int cs_backward_backward(const cs_problem *pb, const float *gOutInput, const float *gOutGrid,
                         const float *input, const float *grid, cs_stream gOut,
                         const float *offset, float *gInput, float *gGrid, float *ggOut,
                         void *stream);

/* (gInput, ggOut) of cu2d:722-891 / cu3d:875-1071.  Either output nullable.
 * gOutggOut (nullable) fuses the second call of modules_2d.py:109 into the same pass:
 * gInput += scatter(gOutggOut * sum_a dw_q/dg_a gOutGrid_a), i.e. `gInput + b_input`
 * of modules_2d.py:111 in one kernel.  gInput must be zero-initialised by the caller. */
int cs_backward_backward_backward(const cs_problem *pb, const float *input, const float *grid,
                                  cs_stream gOut, const float *gOutGrid, const float *gOutgGrid,
                                  cs_stream gOutggOut, const float *offset, float *gInput,
                                  float *ggOut, void *stream);

/* ---- Double precision -----------------------------------------------------------
 * The reference dispatches double through AT_DISPATCH_FLOATING_TYPES_AND_HALF (cu2d:905,948,1009,1076) but feeds
 * its float offset tensor (modules_2d.py:25) through TensorInfo<scalar_t> (cu2d:914), so that instantiation cannot
 * run.  These four entry points are what the dispatch promises: the stages above on double tensors in the
 * reference (channel-first) layout, offset still fp32, every nullable / zero-initialisation rule unchanged,
 * all arithmetic in double.  A correctness path (one thread per (cell, point)); no throughput claim. */
typedef struct cs_stream_f64 {
    const double *ptr;
    int64_t stride_n; /* elements */
    int64_t stride_c; /* elements */
} cs_stream_f64;
int cs_forward_f64(const cs_problem *pb, const double *input, const double *grid, const float *offset,
                   double *out, void *stream);
int cs_backward_f64(const cs_problem *pb, cs_stream_f64 gOut, const double *input, const double *grid,
                    const float *offset, double *gInput, double *gGrid, void *stream);
int cs_backward_backward_f64(const cs_problem *pb, const double *gOutInput, const double *gOutGrid,
                             const double *input, const double *grid, cs_stream_f64 gOut, const float *offset,
                             double *gInput, double *gGrid, double *ggOut, void *stream);
int cs_backward_backward_backward_f64(const cs_problem *pb, const double *input, const double *grid,
                                      cs_stream_f64 gOut, const double *gOutGrid, const double *gOutgGrid,
                                      cs_stream_f64 gOutggOut, const float *offset, double *gInput,
                                      double *ggOut, void *stream);

/* ---- Half precision --------------------------------------------------------------
 * The third type of the reference's AT_DISPATCH_FLOATING_TYPES_AND_HALF (cu2d:905,948,1009,1076); like the double
 * instantiation it cannot run in the reference (float offset tensor through TensorInfo<at::Half>, cu2d:914).
 * Tensors are IEEE binary16 (cs_half = the 16 raw bits), coordinates included, offset stays fp32; every value is
 * widened to float, the arithmetic is the fp32 formulas, results are rounded to half once.  gInput is accumulated
 * in `workspace` (N*C*D*H*W floats, device; zeroed by the call; nullable when gInput is NULL) and then rounded
 * into gInput, which therefore needs no zero-initialisation.  A correctness path; no throughput claim. */
typedef uint16_t cs_half;
typedef struct cs_stream_f16 {
    const cs_half *ptr;
    int64_t stride_n; /* elements */
    int64_t stride_c; /* elements */
} cs_stream_f16;
int cs_forward_f16(const cs_problem *pb, const cs_half *input, const cs_half *grid, const float *offset,
                   cs_half *out, void *stream);
int cs_backward_f16(const cs_problem *pb, cs_stream_f16 gOut, const cs_half *input, const cs_half *grid,
                    const float *offset, cs_half *gInput, cs_half *gGrid, float *workspace, void *stream);
int cs_backward_backward_f16(const cs_problem *pb, const cs_half *gOutInput, const cs_half *gOutGrid,
                             const cs_half *input, const cs_half *grid, cs_stream_f16 gOut, const float *offset,
                             cs_half *gInput, cs_half *gGrid, cs_half *ggOut, float *workspace, void *stream);
int cs_backward_backward_backward_f16(const cs_problem *pb, const cs_half *input, const cs_half *grid,
                                      cs_stream_f16 gOut, const cs_half *gOutGrid, const cs_half *gOutgGrid,
                                      cs_stream_f16 gOutggOut, const float *offset, cs_half *gInput,
                                      cs_half *ggOut, float *workspace, void *stream);

/* ---- Fused multi-cell jet operator (not in the reference; SURVEY section 8f ranks 1 + 2) ------------
 * One gather pass replaces the forward, first-backward (gGrid) and double-backward (gGrid) calls of
 * modules_2d.py:22-74 for a point set shared by all N cells (test_2d.py:36-38) and the caller's sum
 * over the cells (test_2d.py:51):
 *     jets[0]        [C,P] = sum_n sum_q input[n, corner q] * w_q             (cu2d:315-353)
 *     jets[1+a]      [C,P] = sum_n sum_q input[n, corner q] * dw_q/dg_a       (cu2d:476-503)
 *     jets[1+dim+a]  [C,P] = sum_n sum_q input[n, corner q] * d2w_q/dg_a^2    (cu2d:694-706; order 2)
 *     jets[1+2*dim+m][C,P] = sum_n sum_q input[n, corner q] * d2w_q/dg_a dg_b    (cu3d:836-856; order 3: the
 *                            mixed second derivatives, m over (x,y) in 2D and (x,y), (x,z), (y,z) in 3D)
 * a = 0..dim-1 in the order of grid[..., a].  jets is [J, C, P] contiguous with J = 1 + dim (order 1),
 * 1 + 2*dim (order 2) or 1 + 2*dim + dim*(dim-1)/2 (order 3); coords [P, dim].
 * pb->field_layout must be CS_LAYOUT_CHANNEL_LAST (input is [N, T, C]) and C in {4, 8, 16, 32};
 * pb->grid_stride_n, lanes, small_cell, grad_order are ignored; align_corners is honoured in 2D too. */
int cs_jet_forward(const cs_problem *pb, int32_t order, const float *input, const float *coords,
                   const float *offset, float *jets, void *stream);
/* Adjoint of cs_jet_forward: gInput[n, corner q] += sum_j gJets[j] * coef_j,q  -- the triple-backward
 * scatters of modules_2d.py:98-111 in one pass.  gInput is channel-last [N, T, C], zero-initialised
 * by the caller. */
int cs_jet_backward(const cs_problem *pb, int32_t order, const float *gJets, const float *coords,
                    const float *offset, float *gInput, void *stream);

/* ---- Fused PDE-residual head on jets (PIXEL caller glue; not in the reference) ---------------------
 * The head of the reference's scripts, Linear(C,16)-Tanh-Linear(16,1) (test_2d.py:42-47), applied to the
 * jets of cs_jet_forward in second-order Taylor mode, a residual
 *     f = c_u u + c_u3 u^3 + sum_a (c1[a] u_a + c2[a] u_aa)
 * (test_2d.py:221: c1[1] = 2, c_u3 = 5, c_u = -5, c2[0] = -1e-4; test_3d.py:270: c_u = 1, c2[a] = 1;
 * Helmholtz: c_u = k^2, c2[a] = 1), and in the same pass the gradients of scale * sum_p f^2 with respect
 * to the jets (gJets [1+2*dim, C, P], may alias jets) and the head parameters (gW1 [16,C], gb1 [16],
 * gw2 [16], gb2 [1]: accumulated with +=, as is loss_sum [1] += sum_p f^2, unscaled).  f_out [P] nullable. */
typedef struct cs_pde_residual {
    float c_u, c_u3;
    float c1[3];
    float c2[3];
} cs_pde_residual;
int cs_pde_head_step(int32_t dim, int32_t C, int64_t P, const float *jets, const float *W1, const float *b1,
                     const float *w2, const float *b2, const cs_pde_residual *res, float scale,
                     float *gJets, float *gW1, float *gb1, float *gw2, float *gb2, float *loss_sum,
                     float *f_out, void *stream);

/* ---- One-pass fused training step (not in the reference; SURVEY section 8f ranks 1 + 2) ------------------
 * For a head Linear(C,K)-Tanh-Linear(K,1) the whole step of test_2d.py:36-127 -- replicate, sample, sum over
 * cells, head, nested autograd for u_a / u_aa, residual, loss, and the triple-backward scatters of
 * modules_2d.py:98-111 -- is ONE pass over the points.  The sampler and the head's first layer are both
 * linear, so they commute: the cells are mixed with W1 once per step (cs_head_premix), the gather then
 * delivers the hidden pre-activations, the adjoint scatters d loss / d hidden into gVh, and cs_head_postmix
 * returns gInput = W1^T gVh and gW1 = sum_texels gVh (x) input.  Points may be given in any order; binned
 * by texel (cs_bin_points) the kernel gathers from cache and pre-reduces its scatter in registers. */

/* Counting sort of coords [P, dim] on a tile-major texel key (8x8 texels in 2D, 4x4x4 in 3D) of cell 0,
 * refined by the sub-texel quadrant when the point density allows (>= 2 points per sub-bin).
 * sorted [P, dim]; perm [P] (nullable): perm[i] = index in coords of sorted point i.  offset [N] device
 * (nullable).  workspace: cs_bin_workspace_bytes() bytes of device scratch.  Order inside a bin is not
 * deterministic.  Uses pb->dim, D/H/W, P, align_corners, multicell, index_mode. */
int cs_bin_workspace_bytes(const cs_problem *pb, int64_t *bytes);
int cs_bin_points(const cs_problem *pb, const float *coords, const float *offset, float *sorted,
                  int32_t *perm, void *workspace, int64_t workspace_bytes, void *stream);

/* Vh [N*T + 1, K]: Vh[n*T + t, :] = W1 [K, C] applied to input [n, :, t] (channel-first in, channel-last out:
 * the staging transpose and the first Linear layer, test_2d.py:44, in one pass); the extra texel Vh[N*T, :] is
 * set to zero -- out-of-bounds corners of cs_pde_fused_step read it.  K in {4, 8, 16, 32, 64}, C <= 64. */
int cs_head_premix(int32_t N, int32_t C, int64_t T, int32_t K, const float *input, const float *W1,
                   float *Vh, void *stream);
/* gInput [N, C, T] (= or +=, nullable) = W1^T gVh;  gW1 [K, C] (+=, nullable) = sum_{n,t} gVh[n,t,:] (x) input[n,:,t].
 * gVh is [N, T, K] (hidden_first = 0) or [N, K, T] (hidden_first = 1: the output of
 * cs_peer_allreduce_from_channel_last). */
int cs_head_postmix(int32_t N, int32_t C, int64_t T, int32_t K, const float *gVh, int32_t hidden_first,
                    const float *input, const float *W1, float *gInput, int32_t accumulate, float *gW1,
                    void *stream);

/* The pass itself (at most 32 cells).  pb describes the MIXED cells: pb->C = K (hidden width, in {4, 8, 16, 32, 64}),
 * pb->field_layout = CS_LAYOUT_CHANNEL_LAST, P = number of points; coords [P, dim] shared by all cells.
 * Residual f = c_u u + c_u3 u^3 + sum_a (c1[a] u_a + c2[a] u_aa) with u = w2 . tanh(H + b1) + b2 and H the
 * sampled mixed cells summed over the N cells; loss_sum [1] += sum_p f^2 (unscaled); gradients of
 * scale * sum_p f^2: gVh [N*T + 1, K] += (zero-initialised by the caller; like Vh it carries one extra texel,
 * which absorbs the contributions of out-of-bounds corners and is to be ignored), gb1 [K] +=, gw2 [K] +=,
 * gb2 [1] +=.  Vh is the [N*T + 1, K] output of cs_head_premix.
 * aggregate: 0 = one red.global.add.v4.f32 per (cell, point, corner); 1 = contributions of consecutive
 * points with identical corners are summed in registers first (pays with binned points). */
int cs_pde_fused_step(const cs_problem *pb, const float *Vh, const float *coords, const float *offset,
                      const float *b1, const float *w2, const float *b2, const cs_pde_residual *res,
                      float scale, float *gVh, float *gb1, float *gw2, float *gb2, float *loss_sum,
                      int32_t aggregate, void *stream);

/* ---- Gradient all-reduce over peer memory, fused with the layout change ------------
 * acc_ptrs[r] / out_ptrs[r] (r < world <= CS_MAX_PEERS): device addresses, valid on THIS device, of rank r's
 * channel-last accumulator [N, T, C] and channel-first output [N, C, T] (symmetric memory: NVLink-mapped
 * peer allocations).  Tiles are dealt round-robin to the ranks; the owner sums a tile over all peers in rank
 * order, transposes it and stores it into every peer's output, so after a barrier every rank holds the same
 * bits: sum_r acc_r in the reference layout.  small_ptrs[r] (nullable when small_n == 0): a vector of small_n
 * floats per rank, summed over the ranks into this rank's small_out.  The caller must order the launch
 * between two cross-rank barriers (all scatters done before; nobody reuses the buffers until after). */
#define CS_MAX_PEERS 8
int cs_peer_allreduce_from_channel_last(int32_t world, int32_t rank, const float *const *acc_ptrs,
                                        float *const *out_ptrs, int32_t N, int32_t C, int64_t T,
                                        const float *const *small_ptrs, float *small_out, int32_t small_n,
                                        void *stream);

/* Layout-preserving variant for the one-pass step (its accumulator is consumed as it is by cs_head_postmix):
 * out_r[i] = sum_r acc_r[i] for i < n (n a multiple of 4), every rank owning a contiguous slice.  With
 * acc_multicast / out_multicast (the NVSwitch multicast addresses of the same symmetric allocations; both or
 * neither) the sum is taken inside the switch: one multimem.ld_reduce.add.v4.f32 and one multimem.st.v4.f32 per
 * 16 bytes; without them the owner loads the slice from every peer (16-byte loads, rank order) and stores it to
 * every peer.  Same barrier contract and small vector as above. */
int cs_peer_allreduce(int32_t world, int32_t rank, const float *const *acc_ptrs, float *const *out_ptrs, int64_t n,
                      const float *acc_multicast, float *out_multicast, const float *const *small_ptrs,
                      float *small_out, int32_t small_n, void *stream);

/* Staging between the reference layout and the channel-last layout.
 * src [N, C, T] -> dst [N, T, C]   (T = D*H*W) */
int cs_to_channel_last(const float *src, float *dst, int32_t N, int32_t C, int64_t T, void *stream);
/* src [N, T, C] -> dst [N, C, T]; accumulate != 0 adds into dst instead of overwriting */
int cs_from_channel_last(const float *src, float *dst, int32_t N, int32_t C, int64_t T,
                         int32_t accumulate, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* COSINE_SAMPLER_B200_H */
