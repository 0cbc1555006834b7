/* stc_b200.h -- C ABI of libstc_b200.so: the B200 (sm_100a) implementation of STC-GNN's
 * per-timestep spatio-temporal-categorical graph-convolution GRU cell.
 *
 * What it replaces (reference = underdoc-wang/STC-GNN, framework/STC_GNN.py):
 *   STC_Cell.forward                STC_GNN.py:65-79   -> stc_cell_fwd
 *   BDG_Dif.forward (x2 per cell)   STC_GNN.py:31-47   -> inside stc_cell_fwd
 *   BDG_Dif.cheby_poly              STC_GNN.py:24-29   -> inside (feature-side recurrence for Gs,
 *                                                         matrix-space for the small Gc)
 *   autograd of the above           Model_Trainer.py:81 (loss.backward()) -> stc_cell_bwd
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer unless it says "host".
 *   - all tensors are fp32, row-major, feature axis fastest:  X[b][n][c][l].
 *   - the library never allocates or frees device memory.  The forward pass writes what backward
 *     needs into `saved` (size from stc_cell_saved_bytes; keep it untouched until stc_cell_bwd has
 *     run -- it is the autograd node's "saved tensors"); backward additionally takes `scratch`
 *     (size from stc_cell_bwd_scratch_bytes) which may be reused by any later call.
 *   - launches go to `stream` (a cudaStream_t passed as void*); no host synchronisation inside.
 *   - return value: 0 on success, a negative StcStatus otherwise; stc_last_error() gives the text
 *     (thread-local).  Nothing throws across the boundary.  There is no CPU fallback.
 */
#ifndef STC_B200_H_
#define STC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STC_ABI_VERSION 1

typedef enum {
  STC_OK = 0,
  STC_ERR_BAD_ARG = -1,      /* null pointer, non-positive size, bad enum */
  STC_ERR_UNSUPPORTED = -2,  /* shape outside what the kernels tile (message says which) */
  STC_ERR_WORKSPACE = -3,    /* saved/scratch buffer smaller than the *_bytes() query says */
  STC_ERR_CUDA = -4,         /* a CUDA runtime call or launch failed */
  STC_ERR_ARCH = -5          /* device is not compute capability 10.x */
} StcStatus;

enum { STC_ACT_NONE = 0, STC_ACT_RELU = 1 };          /* BDG_Dif(activation=...), STC_GNN.py:13,46 */
enum { STC_SUPPORT_DENSE = 0, STC_SUPPORT_CSR = 1 };

/* Static shape of one cell call.  Mirrors STC_Cell.__init__ (STC_GNN.py:52-58) plus the batch. */
typedef struct {
  int32_t B;        /* batch                                  */
  int32_t N;        /* num_nodes  (regions)                   */
  int32_t C;        /* num_categories (incident types)        */
  int32_t Din;      /* input_dim of Xt                        */
  int32_t h;        /* hidden_dim                             */
  int32_t Ks;       /* number of spatial Chebyshev terms  (orders 0..Ks-1) */
  int32_t Kc;       /* number of categorical Chebyshev terms               */
  int32_t act;      /* STC_ACT_*                              */
  int32_t has_bias; /* use_bias                               */
} StcDims;

/* Spatial support Gs [N,N].  The forward pass contracts over Gs's FIRST index
 * ('bncl,nm->bmcl', STC_GNN.py:37), i.e. applies Gs^T; backward applies Gs.
 *   dense: vals = Gs row-major [N][N]  (vals[n*N+m] = Gs[n,m]); the other fields are ignored.
 *   CSR  : (rowptr, col, vals) is Gs in CSR (row n lists the m with Gs[n,m] != 0, used by backward),
 *          (t_rowptr, t_col, t_vals) is Gs^T in CSR (row m lists the n, used by forward).
 *          Column indices are int32; nnz < 2^31. */
typedef struct {
  int32_t kind;
  int64_t nnz;
  const float* vals;
  const int32_t* rowptr;
  const int32_t* col;
  const float* t_vals;
  const int32_t* t_rowptr;
  const int32_t* t_col;
} StcSupport;

int stc_abi_version(void);
const char* stc_last_error(void);

/* Bytes of the `saved` buffer of one cell call (forward intermediates; also the forward's scratch). */
size_t stc_cell_saved_bytes(const StcDims* d);
/* Bytes of the `scratch` buffer stc_cell_bwd needs. */
size_t stc_cell_bwd_scratch_bytes(const StcDims* d);

/* H' = STC_Cell(Gs, Gc, Xt, H)                                             STC_GNN.py:65-79
 *   gc        [C][C]
 *   xt        [B][N][C][Din], inner [N][C][Din] slab contiguous, batch stride xt_batch_stride ELEMENTS
 *             (the encoder hands in a [:,t] view: STC_GNN.py:111)
 *   h_prev    [B][N][C][h] contiguous
 *   Wg        [(Din+h)*Ks*Kc][2h]  rows ordered (n, c, l)  (STC_GNN.py:17,35-41);  bg [2h] or NULL
 *   Wc        [(Din+h)*Ks*Kc][h];                                                   bc [h]  or NULL
 *   h_out     [B][N][C][h] contiguous, must not alias an input                                   */
int stc_cell_fwd(const StcDims* d, const StcSupport* gs, const float* gc,
                 const float* xt, int64_t xt_batch_stride, const float* h_prev,
                 const float* Wg, const float* bg, const float* Wc, const float* bc,
                 float* h_out, void* saved, size_t saved_bytes, void* stream);

/* ---- the cell split at its spatial-support stages (row-partitioned graphs, stc_gnn_b200/halo.py) ----
 * A rank that owns a block of nodes cannot run the spatial hops inside stc_cell_fwd: each hop needs the
 * boundary-node features of its neighbours (an NCCL halo exchange).  The caller therefore
 *   1. fills the spatial terms of Xt and H (regions STC_SAVED_YX, STC_SAVED_YH of `saved`, Ks-1 terms each,
 *      [B][N][C][Din or h] contiguous) with stc_support_apply + its exchange,
 *   2. calls stage STC_STAGE_GATES  (node-local: Gc terms, gates conv, writes u, r and r*H = term 0 of STC_SAVED_YR),
 *   3. fills terms 1..Ks-1 of STC_SAVED_YR the same way,
 *   4. calls stage STC_STAGE_CANDI  (node-local: candidate conv, tanh, GRU blend -> h_out).
 * d->N is the number of LOCAL nodes.  stc_cell_saved_layout writes the offset (in floats) of every region of
 * `saved` into offsets[STC_SAVED_REGIONS]. */
enum { STC_STAGE_GATES = 0, STC_STAGE_CANDI = 1 };
enum { STC_SAVED_U = 0, STC_SAVED_R, STC_SAVED_C, STC_SAVED_YR, STC_SAVED_YX, STC_SAVED_YH, STC_SAVED_Q,
       STC_SAVED_PG, STC_SAVED_PC, STC_SAVED_REGIONS };
int stc_cell_saved_layout(const StcDims* d, int64_t* offsets_floats, int32_t n_offsets);
int stc_cell_fwd_stage(const StcDims* d, int32_t stage, const float* gc,
                       const float* xt, int64_t xt_batch_stride, const float* h_prev,
                       const float* Wg, const float* bg, const float* Wc, const float* bc,
                       float* h_out, void* saved, size_t saved_bytes, void* stream);

/* Gradients of stc_cell_fwd given d_h_out.  `saved` is the buffer the matching forward call filled.
 *   d_xt      [B][N][C][Din] contiguous, or NULL when Xt needs no gradient
 *   d_h_prev  [B][N][C][h]
 *   dWg,dbg,dWc,dbc  same shapes as the parameters (dbg/dbc NULL when has_bias == 0)
 *   dGs       [N][N] dense or NULL (must be NULL for a CSR support: constants by construction)
 *   dGc       [C][C] or NULL
 *   accumulate_params != 0: parameter/support gradients are ADDED to the buffers' contents
 *   (lets a time loop accumulate dW over steps); == 0: they are overwritten.
 *   d_xt / d_h_prev are always overwritten.                                                     */
int stc_cell_bwd(const StcDims* d, const StcSupport* gs, const float* gc,
                 const float* xt, int64_t xt_batch_stride, const float* h_prev,
                 const float* Wg, const float* Wc, const float* d_h_out,
                 float* d_xt, float* d_h_prev,
                 float* dWg, float* dbg, float* dWc, float* dbc, float* dGs, float* dGc,
                 int32_t accumulate_params, const void* saved, size_t saved_bytes,
                 void* scratch, size_t scratch_bytes, void* stream);

/* ---- the Xt-side spatial terms hoisted out of the time loop ----
 * In the encoder's first layer Xt does not depend on the recurrence (STC_GNN.py:107-118), so its spatial terms for ALL
 * timesteps can be produced by one batched stc_support_apply and handed in:
 *   yx_terms      [Ks-1][B][N][C][Din] contiguous = Y_1 .. Y_{Ks-1} of this step's Xt (NULL: computed inside, as
 *                 stc_cell_fwd / stc_cell_bwd do); the forward then launches no Xt-side hop.
 *   dyx_terms_out [Ks-1][B][N][C][Din]: stc_cell_bwd_x writes the adjoints of those terms there and does NOT fold them:
 *                 no Xt-side adjoint hop and no Xt-side dGs contribution is launched; d_xt (if given) holds the term-0
 *                 part only.  The caller folds all timesteps at once (dGs += sum_t X_t (x) dY_1,t via
 *                 stc_support_outer, dXt += Gs dY_1 via stc_support_apply).  Both pointers go together. */
int stc_cell_fwd_x(const StcDims* d, const StcSupport* gs, const float* gc,
                   const float* xt, int64_t xt_batch_stride, const float* h_prev,
                   const float* Wg, const float* bg, const float* Wc, const float* bc,
                   float* h_out, void* saved, size_t saved_bytes, const float* yx_terms, void* stream);
int stc_cell_bwd_x(const StcDims* d, const StcSupport* gs, const float* gc,
                   const float* xt, int64_t xt_batch_stride, const float* h_prev,
                   const float* Wg, const float* Wc, const float* d_h_out,
                   float* d_xt, float* d_h_prev,
                   float* dWg, float* dbg, float* dWc, float* dbc, float* dGs, float* dGc,
                   int32_t accumulate_params, const void* saved, size_t saved_bytes,
                   void* scratch, size_t scratch_bytes, const float* yx_terms, float* dyx_terms_out, void* stream);
/* dGs[n][m] += coef * sum_{b,j} A[b][n][j] * Bm[b][m][j]  -- the gradient of the mode product Y = Gs^T X w.r.t. a dense
 * Gs [N][N] with A = X ([B][N][width], batch stride in elements) and Bm = dL/dY ([B][N][width] contiguous). */
int stc_support_outer(int32_t N, int32_t B, int32_t width, const float* a, int64_t a_batch_stride, const float* b,
                      float coef, float* dGs, void* stream);

/* ---- the backward split the same way (gradients through a row-partitioned graph) ----
 * Order: stage STC_STAGE_CANDI first, then STC_STAGE_GATES (the reverse of the forward).
 *   1. STC_STAGE_CANDI: clears the parameter and Gc gradients unless accumulate_params, runs the candidate-conv adjoint.  Leaves
 *      d(r*H spatial terms) in region STC_SCRATCH_DYR of `scratch` (Ks terms, [B][N][C][h] each), the x-part adjoint
 *      of term 0 in d_xt and of terms 1..Ks-1 in region STC_SCRATCH_DYX.
 *   2. the caller folds terms Ks-1..1 of STC_SCRATCH_DYR into term 0 with the adjoint hops
 *      (ybar[k-1] += (k >= 2 ? 2 : 1) Gs ybar[k], ybar[k-2] -= ybar[k]; stc_support_apply with transpose = 0 + its exchange),
 *   3. STC_STAGE_GATES: GRU adjoint + gates-conv adjoint (reads term 0 of STC_SCRATCH_DYR as d(r*H)); adds its x-part
 *      adjoints to d_xt / STC_SCRATCH_DYX, writes the h-part adjoints to d_h_prev / STC_SCRATCH_DYH (Ks-1 terms),
 *      finishes dGc,
 *   4. the caller folds STC_SCRATCH_DYX into d_xt and STC_SCRATCH_DYH into d_h_prev the same way.
 * Rows whose d_h_out and STC_SCRATCH_DYR term-0 entries are zero contribute nothing to any parameter gradient (that
 * is how the halo rows of an extended node set are kept out of dW).  A constant (CSR) support has no dGs.      */
enum { STC_SCRATCH_DYR = 0, STC_SCRATCH_DYX, STC_SCRATCH_DYH, STC_SCRATCH_REGIONS };
int stc_cell_bwd_scratch_layout(const StcDims* d, int64_t* offsets_floats, int32_t n_offsets);
int stc_cell_bwd_stage(const StcDims* d, int32_t stage, const float* gc,
                       const float* xt, int64_t xt_batch_stride, const float* h_prev,
                       const float* Wg, const float* Wc, const float* d_h_out,
                       float* d_xt, float* d_h_prev,
                       float* dWg, float* dbg, float* dWc, float* dbc, float* dGc,
                       int32_t accumulate_params, const void* saved, size_t saved_bytes,
                       void* scratch, size_t scratch_bytes, void* stream);

/* Y[b,m,:] = alpha * sum_n A(m,n) X[b,n,:] + beta * Z[b,m,:]  with A = Gs^T (transpose != 0, the
 * forward mode product of STC_GNN.py:37) or A = Gs.  X,Z,Y: [B][N][width]; X and Z may carry a batch
 * stride (elements), Y is contiguous; Z may be NULL when beta == 0; Y may alias Z.  Exposed because it
 * is the support kernel the roofline is quoted on, and the building block of the halo-partitioned path. */
int stc_support_apply(const StcSupport* gs, int32_t N, int32_t B, int32_t width, int32_t transpose,
                      const float* x, int64_t x_batch_stride, const float* z, int64_t z_batch_stride,
                      float* y, float alpha, float beta, void* stream);

/* The same product restricted to the output nodes listed in rows[n_rows] (int32, device): the other rows of Y are
 * not touched.  CSR supports only.  The row-partitioned path (stc_gnn_b200/halo.py) computes the rows whose
 * neighbours are all local while the halo exchange of the hop is in flight, then the boundary rows. */
int stc_support_apply_rows(const StcSupport* gs, int32_t N, int32_t B, int32_t width, int32_t transpose,
                           const float* x, int64_t x_batch_stride, const float* z, int64_t z_batch_stride,
                           float* y, float alpha, float beta, const int32_t* rows, int32_t n_rows, void* stream);

/* Halo rows of a row-partitioned hop.  x_ext is an extended tensor [B][nloc + nhalo][width] (batch stride in
 * elements).  pack: send[j][b][:] = x_ext[b][idx[j]][:] for the n_rows local boundary nodes idx (int32, device;
 * grouped by destination rank, so `send` is the input of one all-to-all with row counts as split sizes);
 * unpack: x_ext[b][row0 + j][:] = recv[j][b][:] (row0 = nloc: the halo rows, ordered by owner rank). */
int stc_halo_pack(const float* x_ext, int64_t x_batch_stride, int32_t width, int32_t B, const int32_t* idx,
                  int32_t n_rows, float* send, void* stream);
int stc_halo_unpack(const float* recv, int32_t width, int32_t B, int32_t row0, int32_t n_rows, float* x_ext,
                    int64_t x_batch_stride, void* stream);

/* Building block self-test / microbenchmark: D[M][N] = A[M][K] * B[K][N], row-major fp32, evaluated as a
 * 3xTF32 tcgen05 product with TMEM accumulation (the same code path as the gate contraction).  N <= 256. */
int stc_tf32x3_gemm(const float* a, const float* b, float* d, int32_t M, int32_t N, int32_t K, void* stream);

/* Independent kernels of one stc_cell_fwd / stc_cell_bwd call (dW beside the adjoint hops, the dGs outer products, the
 * Xt-side hops beside the H-side hops) may be forked onto internal side streams and are joined back with events
 * before the call returns; by default only for problems too small to fill the device.  mode: -1 = by problem size
 * (default; the STC_CONCURRENCY environment variable sets the initial value), 0 = always one stream, 1 = always fork. */
int stc_concurrency_set(int32_t mode);

/* Number of kernel launches the last stc_cell_fwd / stc_cell_bwd on this thread issued
 * (bench.py reports gpu_launches from these). */
int stc_last_launch_count(void);

/* Optional per-kernel instrumentation for bench.py's roofline line (off by default; when on, every kernel
 * launch is bracketed by CUDA events on its own stream).  stc_timing_collect synchronises the recorded
 * events, ADDS per-kind device milliseconds / launch counts / algorithmic bytes (the compulsory HBM traffic
 * of each launch, formulas next to each launcher) into the caller's arrays of length n_kinds (host
 * pointers), clears the record and returns the number of kinds the library knows. */
int stc_timing_enable(int32_t on);
int stc_timing_collect(double* ms_by_kind, int64_t* launches_by_kind, double* alg_bytes_by_kind, int32_t n_kinds);
const char* stc_kernel_kind_name(int32_t kind);

/* Diagnostic: while dev_buf (int64 [n_slots], device memory, caller-owned) is registered, CTA 0 of the tcgen05
 * gate-convolution kernels stamps clock64() at its phase boundaries for its first n_slots/32 tiles
 * (16 stamps per tile; the gates convolution uses the first half of the buffer, the candidate convolution the
 * second; tools/trace_conv.py prints the deltas).  Pass NULL to switch it off (the default). */
int stc_debug_trace_set(void* dev_buf, int64_t n_slots);

#ifdef __cplusplus
}
#endif
#endif /* STC_B200_H_ */
