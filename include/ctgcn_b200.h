/* ctgcn_b200 — C-ABI of the B200-native CTGCN forward hot path.
 *
 * The reference (jhljx/CTGCN) is pure Python: its "operator API" for this path is the
 * torch.nn.Module surface of layers.py / models.py, whose arithmetic is delegated to torch
 * library calls.  Each entry point below replaces one group of those call sites (cited as
 * reference file:line, relative to the reference repo root).  Python modules with the
 * reference's exact signatures (ctgcn_b200/layers.py, ctgcn_b200/models.py) bind these
 * through ctypes; see INTEGRATION.md for the stub a reference maintainer would add.
 *
 * Conventions
 *  - every function returns 0 on success, a negative CTGCN_E* code on failure; the message is
 *    available (thread-local) from ctgcn_last_error().  Nothing throws across the ABI.
 *  - all tensor pointers are DEVICE pointers to fp32 row-major data unless the name says host;
 *    `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *  - compute entry points never allocate, never free caller memory and never synchronise the
 *    stream.  Scratch comes from a caller-owned workspace whose size is queried first.
 *  - the library owns only opaque plan handles (ctgcn_plan) and their device arrays.
 */
#ifndef CTGCN_B200_H
#define CTGCN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTGCN_OK 0
#define CTGCN_EINVAL (-1)   /* bad argument / unsupported shape           */
#define CTGCN_ECUDA (-2)    /* a CUDA runtime call or kernel launch failed */
#define CTGCN_ENOMEM (-3)   /* workspace too small / allocation failed     */
#define CTGCN_ENODEV (-4)   /* no usable sm_100 device                     */

#define CTGCN_MAX_CORES 64
#define CTGCN_MAX_NEG 64    /* largest neg_num of the negative-sampling loss entry points */

/* activation codes for ctgcn_linear_fwd / ctgcn_spmm_linear_fwd (layers.py:98-99,104-105) */
#define CTGCN_ACT_NONE 0
#define CTGCN_ACT_SELU 1

/* output modes of ctgcn_gru_seq_fwd */
#define CTGCN_GRU_SUM_LN 0  /* y[n,:]   = LayerNorm(sum_s h_s)   layers.py:59-62  */
#define CTGCN_GRU_EACH_LN 1 /* y[n,s,:] = LayerNorm(h_s)         models.py:249-250 */

/* implementation selector (ctgcn_set_gru_impl): all are CUDA.  AUTO picks the fastest tensor-core kernel the shapes allow
 * (the CTA-pair kernel of csrc/gru_tc2.cu, else the fp32 sequence kernel); TCGEN05 = AUTO but an error instead of the
 * fp32 kernel when no tensor-core kernel takes the shape.  The last three select one specific tensor-core build (A/B measurements
 * and tests): the one-CTA kernel of round 1 (csrc/gru_tc.cu), the round-2 kernel built WITHOUT pairing (cta_group::1), and the
 * step-per-launch kernel for wide hidden states (csrc/gru_wide_tc.cu: what AUTO runs for H = 256 / 384 / 512) on every shape it
 * takes, H = 128 included.  Workspace queries depend on the selection: set it before asking. */
#define CTGCN_IMPL_AUTO 0
#define CTGCN_IMPL_SIMT 1
#define CTGCN_IMPL_TCGEN05 2
#define CTGCN_IMPL_TC_ONE_CTA_R1 3
#define CTGCN_IMPL_TC_UNPAIRED 4
#define CTGCN_IMPL_TC_WIDE 5

/* recurrent cell of the sequence kernels: the reference's rnn_type (layers.py:26-30, models.py:232-237) */
#define CTGCN_CELL_GRU 0
#define CTGCN_CELL_LSTM 1

typedef struct ctgcn_plan ctgcn_plan;

int ctgcn_version(void);
const char* ctgcn_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
int64_t ctgcn_launch_count(void);
int ctgcn_device_check(void); /* 0 iff the current device is compute capability 10.x */

/* Per-kernel-class device timing for bench.py's roofline: while enabled, every launch is bracketed by CUDA
 * events on its stream.  ctgcn_prof_collect synchronises the device and returns, per class
 * (0 cumulative SpMM, 1 GRU(+LayerNorm), 2 dense linear, 3 weight packing/transposes, 4 sparse-input linear),
 * the summed milliseconds and launch counts since the last reset.  Arrays of CTGCN_PROF_NCLASS entries. */
#define CTGCN_PROF_NCLASS 5
int ctgcn_prof_enable(int on);
int ctgcn_prof_collect(double* ms, int64_t* counts, int reset);

/* ---------------------------------------------------------------- graph plan
 * Replaces the per-call work torch.sparse.mm does on the reference's adj_list
 * (layers.py:41-45: K uncoalesced COO matrices, helper.py:51-82 / utils.py:89-95).
 * The K matrices of ONE snapshot are merged once into a "union CSR": every distinct (row,col)
 * is stored once with the first core index it belongs to ("level"); entries that are not a
 * suffix of the core list (e.g. the +I diagonal that helper.py:72 adds to the first matrix only,
 * or any non-nested / per-core-weighted entry) are stored as "one-shot" entries of their level.
 * Duplicate COO entries inside one matrix are summed (torch.sparse.mm semantics).
 *
 * rows/cols: K pointers to int64 index arrays, vals: K pointers to fp32, nnz: K counts (host array).
 * on_device != 0: index/value arrays are device pointers, else host pointers.
 * n_rows x n_cols is the (common) shape; the adjacency case has n_rows == n_cols.
 * Synchronises the stream (one-off build). */
int ctgcn_plan_create_coo(int64_t n_rows, int64_t n_cols, int k, const int64_t* const* rows,
                          const int64_t* const* cols, const float* const* vals, const int64_t* nnz,
                          int on_device, void* stream, ctgcn_plan** out);

/* Same plan from an already merged CSR: rowptr[n_rows+1] int32, col/val/level per entry; level byte =
 * first core index (bits 0..6) | 0x80 for one-shot entries; entries of a row sorted by level.
 * nnz_raw_sum is recorded for statistics only (sum of the K matrices' stored non-zeros). */
int ctgcn_plan_create_csr(int64_t n_rows, int64_t n_cols, int k, const int32_t* rowptr, const int32_t* col,
                          const float* val, const uint8_t* level, int64_t nnz_raw_sum, int on_device,
                          void* stream, ctgcn_plan** out);
int ctgcn_plan_destroy(ctgcn_plan* plan);
/* stats[0..7] = n_rows, n_cols, k, entries in the union CSR, sum of raw nnz, sum of coalesced nnz,
 *               one-shot entries, device bytes held */
int ctgcn_plan_stats(const ctgcn_plan* plan, int64_t stats[8]);
/* copy the plan arrays into caller-owned DEVICE buffers (any may be NULL to skip): rowptr[n_rows+1],
 * col/val/level[entries].  For tests / inspection; asynchronous on `stream`. */
int ctgcn_plan_arrays(const ctgcn_plan* plan, int32_t* rowptr, int32_t* col, float* val, uint8_t* level, void* stream);

/* ---------------------------------------------------------------- cumulative k-core SpMM
 * layers.py:41-48:  S_i = S_{i-1} + A_i x ; U_i = relu(S_i)   for i = 0..K-1, in ONE pass over the
 * union CSR.  x: [n_cols, d] with row stride ldx (elements); u: [n_rows, K, d] contiguous. */
int ctgcn_cumspmm_fwd(const ctgcn_plan* plan, const float* x, int64_t ldx, int d, float* u, void* stream);

/* Same pass with the relu switchable: relu = 0 returns the cumulative sums S_i themselves (with a K = 1 plan: the plain
 * SpMM A·x, used for the weight gradient of a sparse-input Linear, layers.py:97 with a COO x). */
int ctgcn_cumspmm_fwd_ex(const ctgcn_plan* plan, const float* x, int64_t ldx, int d, int relu, float* u, void* stream);

/* Backward of the cumulative SpMM w.r.t. x (the autograd of layers.py:41-47).
 * plan_t: plan of the TRANSPOSED list [A_0^T … A_{K-1}^T] (shape n x m; the k-core adjacencies are symmetric, but the module
 *         API does not promise it).  g: [m, K, d] contiguous = dL/dS_i (the caller has applied the relu mask).
 * dx[n, d] (row stride lddx) = sum_j A_j^T (sum_{i>=j} g_i).  workspace: ctgcn_cumspmm_bwd_workspace_bytes(plan_t, d). */
size_t ctgcn_cumspmm_bwd_workspace_bytes(const ctgcn_plan* plan_t, int d);
int ctgcn_cumspmm_bwd(const ctgcn_plan* plan_t, const float* g, int d, float* dx, int64_t lddx, void* workspace,
                      size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- GRU over a short sequence + LayerNorm
 * layers.py:59-62 (sequence = core axis, mode SUM_LN) and models.py:249-250 (sequence = snapshot axis,
 * mode EACH_LN).  nn.GRU(num_layers=1, batch_first=True), h0 = 0, PyTorch packing [r;z;n]:
 *   w_ih [3H, d_in], w_hh [3H, H], b_ih/b_hh [3H] or NULL (bias=False).
 * seq element (n, s, j) is read at seq[n*seq_row_stride + s*seq_step_stride + j].
 * SUM_LN : y[n*y_row_stride + j];  EACH_LN: y[n*y_row_stride + s*y_step_stride + j].
 * workspace: ctgcn_gru_workspace_bytes(d_in, h) bytes (packed weights). */
size_t ctgcn_gru_workspace_bytes(int d_in, int h);
int ctgcn_gru_seq_fwd(const float* seq, int64_t seq_row_stride, int64_t seq_step_stride, int64_t n, int steps,
                      int d_in, int h, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                      const float* ln_w, const float* ln_b, float eps, int mode, float* y, int64_t y_row_stride,
                      int64_t y_step_stride, void* workspace, size_t workspace_bytes, void* stream);
int ctgcn_set_gru_impl(int impl);
/* CoreDiffusion as ONE launch (default on): for 128-wide GRU layers ctgcn_core_diffusion_fwd* run the cumulative SpMM inside the
 * tensor-core GRU kernel (gather warps fetch the next tile's feature rows with bulk copies while the tensor cores work on the
 * current tile).  0 = always the two-kernel path (SpMM → U → GRU); used by tests and A/B measurements. */
int ctgcn_set_fusion(int on);
/* debug: when non-NULL, block 0 of every following tcgen05 GRU launch writes clock64() stamps of its pipeline events
 * into device_buf[32 events][64 steps] (int64); NULL switches it off. */
int ctgcn_debug_gru_trace(int64_t* device_buf);
/* test hook: one half-step of a GRU cell's pre-activations for d_in = 64 through the tcgen05 weight packer, chunk images,
 * descriptors, split-bf16 MMAs and TMEM loads of the GRU kernel:
 *   out[128,256] = [ x W_in^T | x W_ir^T + h W_hr^T | x W_iz^T + h W_hz^T | h W_hn^T ]  for hidden features 0..63,
 * x [128,64], h [128,128], w_ih [384,64], w_hh [384,128] (no bias, no pre-scaling).  workspace >= 512 KB device memory. */
int ctgcn_selftest_umma(const float* x, const float* h, const float* w_ih, const float* w_hh, float* out, void* workspace,
                        size_t workspace_bytes, void* stream);

/* rnn_type-generic form of the two functions above (cell = CTGCN_CELL_GRU: identical to ctgcn_gru_seq_fwd).
 * CTGCN_CELL_LSTM: nn.LSTM(num_layers=1, batch_first=True), h0 = c0 = 0, PyTorch packing [i;f;g;o]:
 *   w_ih [4H, d_in], w_hh [4H, H], b_ih/b_hh [4H] or NULL; only the output sequence h_s is used (layers.py:59, models.py:249). */
size_t ctgcn_rnn_workspace_bytes(int cell, int d_in, int h);
int ctgcn_rnn_seq_fwd(int cell, const float* seq, int64_t seq_row_stride, int64_t seq_step_stride, int64_t n, int steps,
                      int d_in, int h, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                      const float* ln_w, const float* ln_b, float eps, int mode, float* y, int64_t y_row_stride,
                      int64_t y_step_stride, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- CoreDiffusion.forward (layers.py:38-63)
 * y[n_rows, h] (row stride ldy) = LayerNorm(sum_i GRU(relu(cumsum_i A_i x))).
 * workspace: ctgcn_core_diffusion_workspace_bytes(plan, d_in, h). */
size_t ctgcn_core_diffusion_workspace_bytes(const ctgcn_plan* plan, int d_in, int h);
int ctgcn_core_diffusion_fwd(const ctgcn_plan* plan, const float* x, int64_t ldx, int d_in, int h,
                             const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                             const float* ln_w, const float* ln_b, float eps, float* y, int64_t ldy,
                             void* workspace, size_t workspace_bytes, void* stream);

/* The per-core sums of a CoreDiffusion call ([rows, K, d_in] fp32) live in the workspace between its two kernels.  When all
 * rows would need more than `bytes` (default 8 GiB; 0 restores it) the layer is evaluated in row chunks — whole waves of the
 * persistent sequence kernel — and the workspace queries return the bounded size (BASELINE.json configs[4]: 5 M nodes x K 20 x
 * 256-d would need 102 GB otherwise).  Results do not depend on the chunking.  Process-wide setting. */
int ctgcn_set_workspace_cap(size_t bytes);

/* Same layer with the snapshot exchange of CTGCN.forward (models.py:248) fused into the epilogue: output row r is stored
 * into the buffer of the node slice that owns it (n_slices balanced contiguous slices of the n_rows nodes, the first
 * n_rows % n_slices of them one row longer) at
 *     slice_ptrs[g][(r - start_g) * slice_row_stride + slice_col_offset + j],   j < h.
 * slice_ptrs is a DEVICE array of n_slices pointers; with peer-mapped (NVLink) pointers every GPU writes its snapshot's rows
 * straight into the [rows, T, D] sequence buffer of the rank that runs the temporal GRU on them.  The caller provides the
 * cross-GPU barrier before those buffers are read. */
int ctgcn_core_diffusion_fwd_scatter(const ctgcn_plan* plan, const float* x, int64_t ldx, int d_in, int h,
                                     const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                                     const float* ln_w, const float* ln_b, float eps, float* const* slice_ptrs, int n_slices,
                                     int64_t slice_row_stride, int64_t slice_col_offset, void* workspace,
                                     size_t workspace_bytes, void* stream);

/* rnn_type-generic CoreDiffusion.forward: `cell` selects nn.GRU / nn.LSTM over the core axis (layers.py:26-30).
 * y != NULL: plain output (row stride ldy); y == NULL: the scatter form above (slice_ptrs …). */
size_t ctgcn_core_diffusion_rnn_workspace_bytes(const ctgcn_plan* plan, int cell, int d_in, int h);
int ctgcn_core_diffusion_rnn_fwd(const ctgcn_plan* plan, int cell, const float* x, int64_t ldx, int d_in, int h,
                                 const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                                 const float* ln_w, const float* ln_b, float eps, float* y, int64_t ldy,
                                 float* const* slice_ptrs, int n_slices, int64_t slice_row_stride, int64_t slice_col_offset,
                                 void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- MLP layers (layers.py:95-106)
 * dense:  y[n, d_out] = act(x[n, d_in] w^T + b),  w [d_out, d_in] (nn.Linear layout), b may be NULL.
 * workspace: ctgcn_linear_workspace_bytes(d_in, d_out). */
size_t ctgcn_linear_workspace_bytes(int64_t d_in, int64_t d_out);
int ctgcn_linear_fwd(const float* x, int64_t ldx, int64_t n, int64_t d_in, const float* w, const float* b,
                     int64_t d_out, int act, float* y, int64_t ldy, void* workspace, size_t workspace_bytes,
                     void* stream);
/* sparse COO input (the one-hot identity of helper.py:169-172, the 'combine'/'adj' degree features of
 * helper.py:136-155): x given as a K=1 plan of shape [n, d_in].  Same workspace size query. */
int ctgcn_spmm_linear_fwd(const ctgcn_plan* x_plan, const float* w, const float* b, int64_t d_out, int act,
                          float* y, int64_t ldy, void* workspace, size_t workspace_bytes, void* stream);

/* test hook: one GRU half-step through tcgen05.mma.cta_group::2 on a CTA pair (csrc/umma2_selftest.cu) — the mechanisms of
 * csrc/gru_tc2.cu in isolation (cluster launch, 2-CTA TMEM allocation, half weight chunks + remote-arrive relay, multicast commit).  out[256,256] = [x W_in^T | x W_ir^T + h W_hr^T | x W_iz^T + h W_hz^T | h W_hn^T] for hidden
 * features 0..63 from x[256,64], h[256,128], w_ih[384,64], w_hh[384,128]; workspace >= 512 KB of device memory. */
int ctgcn_selftest_umma_pair(const float* x, const float* h, const float* w_ih, const float* w_hh, float* out, void* workspace,
                             size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- negative-sampling loss (metrics.py:18-93)
 * The unsupervised loss the trainer evaluates on every batch (embedding.py:347).  One snapshot per call.
 *
 * ctgcn_neg_sample replaces NegativeSamplingLoss.__get_node_indices (metrics.py:68-93, a Python loop with random.sample per
 * batch node).  pair_ptr int64[n_nodes+1] / pair_idx int32[]: CSR of the walk co-occurrence lists (helper.py:85-94);
 * freq int32[freq_len]: the frequency-expanded negative list (helper.py:97-106); batch int64[n_batch] node ids.
 * Outputs: pos int32[n_batch, neg_num] — the kept neighbours of batch node b in pos[b, 0..count[b]), -1 after; all of them in
 * stored order when the node has at most neg_num, else neg_num distinct ones drawn uniformly; count int32[n_batch];
 * neg int32[neg_num] — node ids at neg_num distinct positions of freq.  Batch ids outside [0, n_nodes) keep nothing.
 * Counter-based generator: the same (seed, inputs) gives the same draw on every device. */
int ctgcn_neg_sample(const int64_t* pair_ptr, const int32_t* pair_idx, int64_t n_nodes, const int32_t* freq, int64_t freq_len,
                     const int64_t* batch, int64_t n_batch, int neg_num, uint64_t seed, int32_t* pos, int32_t* count,
                     int32_t* neg, void* stream);
/* loss[0] = mean_s softplus(-<e_node, e_pos>) + q * mean_s softplus(<e_node, sum_j e_neg_j>) over the S = sum_b count[b]
 * samples (metrics.py:55-61: BCEWithLogits, mean reduction; pos_score by mul+sum, neg_score by matmul+sum); 0 when S = 0.
 * emb [n_nodes, d] fp32 with row stride ld.  The backward ACCUMULATES grad_loss[0] * dloss/demb into grad_emb (row stride
 * ldg; the caller zeroes it) and must get the workspace the forward filled for the same arguments. */
size_t ctgcn_neg_loss_workspace_bytes(int64_t n_batch, int d);
int ctgcn_neg_loss_fwd(const float* emb, int64_t ld, int64_t n_nodes, int d, const int64_t* batch, int64_t n_batch,
                       const int32_t* pos, const int32_t* count, const int32_t* neg, int neg_num, float q, float* loss,
                       void* workspace, size_t workspace_bytes, void* stream);
int ctgcn_neg_loss_bwd(const float* emb, int64_t ld, int64_t n_nodes, int d, const int64_t* batch, int64_t n_batch,
                       const int32_t* pos, const int32_t* count, const int32_t* neg, int neg_num, float q,
                       const float* grad_loss, float* grad_emb, int64_t ldg, void* workspace, size_t workspace_bytes,
                       void* stream);

/* ---------------------------------------------------------------- host helper (no GPU needed)
 * Exact k-core numbers by bucket peeling (replaces networkx.core_number used at
 * preprocessing/structure_generation.py:35 for the synthetic generators).  CSR of an undirected simple
 * graph: rowptr int64[n+1], col int32[rowptr[n]]; core_out int32[n]. */
int ctgcn_kcore_numbers(int64_t n, const int64_t* rowptr, const int32_t* col, int32_t* core_out);

#ifdef __cplusplus
}
#endif
#endif /* CTGCN_B200_H */
