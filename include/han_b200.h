/*
 * han_b200.h — C-ABI of libhan_sm100.so: the B200 (sm_100a) implementation of the HAN
 * node-level + semantic-level attention hot path.
 *
 * The reference (CG-Labs/HAN) has no FFI: its "operator API" is four Python callables that build
 * TF1 graph nodes.  Each entry point below names the reference lines it replaces; the Python host
 * (han_b200/) binds these through ctypes and exposes the reference's own names.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes.  All pointers are DEVICE pointers unless named host_*.
 *  - The caller owns every buffer; the library never allocates, frees or synchronises.
 *  - Every call is asynchronous on `stream` (a cudaStream_t passed as void*).
 *  - Return value: 0 = ok, <0 = invalid argument, >0 = cudaError_t.  han_last_error() returns a
 *    thread-local message for the last non-zero return.
 *  - fp32 row-major everywhere; CSR = int64 indptr[n+1] + int32 indices[nnz], columns ascending.
 *  - K = heads, H = hidden units per head, D = K*H.  Supported (K,H): see han_attn_shape_supported.
 *  - Node table T: [n][TS] fp32, TS = han_table_stride(K,H) = D:
 *        T[j][0:D]   = S_j  = X_j W           (utils/layers.py:20)
 *    f2_j = S_j a2 + b2 (utils/layers.py:24) is NOT stored: the kernels that need it (K-B, the by-source pass of
 *    K-D, han_attn_coefs) recompute it from the row they fetch anyway (8 FMAs per head), which keeps a gathered
 *    record at D floats -- 256 B for K = H = 8, a whole number of 64-byte DRAM fetches -- and takes a2 / b2.
 *  - Row record R: [n][RS] fp32, RS = han_record_stride(K,H) = roundup(D + 3K, 4):
 *        [ dV (D) | f1 (K) | lse (K) | delta (K) ]
 *        f1 = S a1 + b1 (utils/layers.py:23); lse = log-sum-exp of the row's logits, so that
 *        alpha_ij = exp(leaky_relu(f1_i + f2_j) - lse_i) (:27); dV, delta are filled by the backward.
 */
#ifndef HAN_B200_H
#define HAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* han_stream_t; /* cudaStream_t */

enum { HAN_ACT_IDENTITY = 0, HAN_ACT_ELU = 1 };          /* models/gat.py:10,28 */
enum { HAN_SEM_REFERENCE = 0, HAN_SEM_PAPER = 1 };        /* utils/layers.py:156 vs han.pdf Eq.7-9 */
enum { HAN_DENSE_ADJ = 0, HAN_DENSE_BIAS = 1, HAN_DENSE_POSITIVE = 2 }; /* meaning of a dense N x N input */
enum { HAN_F32 = 0, HAN_F64 = 1 };

int han_version(void);
const char* han_last_error(void);
int han_attn_shape_supported(int K, int H);
int han_table_stride(int K, int H);
int han_record_stride(int K, int H);

/* ---- K-0: graph builder. Replaces utils/process.py:14-25 (adj_to_bias) --------------------- */

/* Row counts of the mask of a dense n x n matrix (leading dimension ld, dtype HAN_F32/HAN_F64).
 *   HAN_DENSE_ADJ : entry (i,j) is an edge iff adj[i][j] + (i==j) > 0   (process.py:20,23 with nhood=1)
 *   HAN_DENSE_BIAS: entry is an edge iff bias[i][j] == 0 (what -1e9*(1-mt) leaves, process.py:25);
 *                   *bad_count receives the number of entries that are neither 0 nor <= -1e8.
 *   HAN_DENSE_POSITIVE: entry is an edge iff value > 0 (an already-formed mt, process.py:21-24; nhood > 1)
 * row_counts[n] int32.  bad_count may be NULL for HAN_DENSE_ADJ. */
int han_dense_row_counts(const void* dense, int dtype, int kind, int64_t n, int64_t ld,
                         int32_t* row_counts, int32_t* bad_count, han_stream_t stream);

/* Exclusive scan int32 counts[n] -> int64 indptr[n+1].  ws: han_scan_workspace_bytes(n). */
size_t han_scan_workspace_bytes(int64_t n);
int han_scan_counts(const int32_t* counts, int64_t n, int64_t* indptr, void* ws, size_t ws_bytes,
                    han_stream_t stream);

/* Column indices in ascending order per row == np.nonzero(mask) order. */
int han_dense_fill_indices(const void* dense, int dtype, int kind, int64_t n, int64_t ld,
                           const int64_t* indptr, int32_t* indices, han_stream_t stream);

/* Transposed structure (CSC of the same pattern) + edge permutation: for transposed edge t,
 * t_indices[t] = destination row i, perm[t] = position of edge (i,j) in the CSR.  Rows ascending
 * within each column, so the result is deterministic.  perm may be NULL (only edge weights need it).
 * ws: han_transpose_workspace_bytes. */
size_t han_transpose_workspace_bytes(int64_t n_rows, int64_t n_cols, int64_t nnz);
int han_csr_transpose(int64_t n_rows, int64_t n_cols, int64_t nnz, const int64_t* indptr,
                      const int32_t* indices, int64_t* t_indptr, int32_t* t_indices, int32_t* perm,
                      void* ws, size_t ws_bytes, han_stream_t stream);

/* Sort helper for COO-built graphs: sorts each row's columns ascending in place (payload, nullable,
 * follows its key) and writes to *dup_count the number of adjacent duplicates (callers dedupe on
 * their side).  scratch: int32[n_rows + 64]. */
int han_csr_sort_rows(int64_t n_rows, const int64_t* indptr, int32_t* indices, int32_t* payload,
                      int32_t* scratch, int32_t* dup_count, han_stream_t stream);

/* ---- K-A: projection. Replaces utils/layers.py:20,23,24 for G groups of K heads at once ----- */
/* X [n][ldx] (F columns used), W [F][G*D] (meta-path g, head k in columns g*D + k*H ...),
 * a1 [G][K][H], b1 [G][K].  Writes T [G][n][TS] and f1 = S a1 + b1 into R[g][:, D:D+K] (R [G][n][RS]).
 * mode must be 0: exact-FP32 CUDA-core FFMA, any supported (K,H), any alignment; the tensor-core
 * variants are han_project_fwd_tc below. */
int han_project_fwd(const float* X, int64_t n, int64_t F, int64_t ldx, const float* W, int G, int K,
                    int H, const float* a1, const float* b1, float* T, float* R, int mode, han_stream_t stream);

/* The same projection on the tcgen05 tensor cores (TMA-staged operands, TMEM accumulators, f1
 * fused into the epilogue).  K = H = 8, 1 <= G <= 4 (<= 256 accumulator columns), X 16-byte aligned
 * with ldx % 4 == 0.  mode 1 = 3xTF32 (X and W split hi+lo: FP32-grade), 2 = 2xTF32 (X exactly
 * representable in tf32, e.g. 0/1 features; only W split), 3 = plain TF32.
 * ws: han_project_tc_workspace_bytes (transposed hi/lo copies of W). */
size_t han_project_tc_workspace_bytes(int64_t F, int G, int K, int H);
int han_project_fwd_tc(const float* X, int64_t n, int64_t F, int64_t ldx, const float* W, int G, int K,
                       int H, const float* a1, const float* b1, float* T, float* R, float* T_mc, int64_t t_rows, int64_t t_row0, int64_t r_rows,
                       int mode, void* ws, size_t ws_bytes, han_stream_t stream);
/* Destination addressing: meta-path g, local row i goes to T + ((g*t_rows + t_row0 + i)*TS) and
 * R + ((g*r_rows + i)*RS); t_rows / r_rows = 0 mean n (tables are exactly [G][n][.]); larger values
 * let the kernel write this rank's slice of full-size tables shared with other ranks.
 * T_mc (nullable): NVLS MULTICAST address of a symmetric node table [G][t_rows][TS].  When given, the
 * epilogue writes its rows with multimem.st instead, so the GEMM and the all-gather of its output are
 * one kernel: NVSwitch replicates every store into all ranks' tables.  T is then unused. */

/* Copies local floats to a multicast address (for producers without a fused multicast epilogue). */
int han_multicast_copy(const float* src, float* dst_mc, int64_t n_floats, han_stream_t stream);

/* dW [F][G*D] = X^T dS  (split over rows + deterministic reduce).  dS [G][n][D].
 * ws: han_project_bwd_workspace_bytes. */
size_t han_project_bwd_workspace_bytes(int64_t n, int64_t F, int G, int D);
int han_project_bwd(const float* X, int64_t n, int64_t F, int64_t ldx, const float* dS, int G, int D,
                    float* dW, void* ws, size_t ws_bytes, int mode, han_stream_t stream);

/* The same dW on the tcgen05 tensor cores (K = H = 8, 1 <= G <= 4): split-K over the node index,
 * producer warps transpose + hi/lo-split the operands into the swizzled K-major layout, accumulators
 * in TMEM, deterministic second-stage reduce.  mode as in han_project_fwd_tc. */
size_t han_project_bwd_tc_workspace_bytes(int64_t n, int64_t F, int G);
int han_project_bwd_tc(const float* X, int64_t n, int64_t F, int64_t ldx, const float* dS, int G,
                       float* dW, void* ws, size_t ws_bytes, int mode, han_stream_t stream);

/* ---- K-B: fused CSR edge-softmax-aggregate. Replaces utils/layers.py:26-35,46 for K heads ---- */
/* For destination rows [0,n_dst): alpha_ij = softmax_j(leaky_relu_0.2(f1_i + f2_j)), V_i = sum_j
 * alpha_ij S_j, out_i = act(V_i + bias).  T is indexed by the CSR's column ids; f2_j = T_j a2 + b2 (a2 [K][H],
 * b2 [K] of this meta-path) is computed per gathered row; f1 is read from
 * R[:, D:D+K]; lse is written to R[:, D+K:D+2K]; V to vsave [n_dst][D]; out to
 * out + i*out_stride (so K-B writes straight into Z[n][P][D], models/gat.py:46,58,60).
 * colmean (nullable) [D]: value used for rows with no edge at all (dense-path uniform 1/N row).
 * edge_w (nullable) [nnz], CSR order: sp_attn_head's stored adjacency values (utils/layers.py:95-96),
 * l_ij = w_ij (f1_i + f2_j); NULL = the 0/1 adjacency of attn_head.  Entry points: han_attn_fwd_chunked* below. */

/* Per-edge coefficients alpha [nnz][K] (utils/layers.py:43-44 return_coef), from the saved lse. */
int han_attn_coefs(const int64_t* indptr, const int32_t* indices, int64_t n_dst, const float* T, const float* a2,
                   const float* b2, const float* R, int K, int H, const float* edge_w, float* alpha,
                   han_stream_t stream);

/* Chunked edge-stream kernels of K-B / the by-source pass of K-D: a warp owns a
 * contiguous chunk of whole rows (~han_csr_chunk_edges(nnz) edges, boundaries precomputed once per graph) and pulls the
 * gathered rows through a shared-memory cp.async ring, so bytes in flight do not depend on registers
 * and work is balanced by edges, not rows.  chunk_rows: int32[han_csr_num_chunks(nnz) + 1]. */
int64_t han_csr_chunk_edges(int64_t nnz);   /* ~nnz/(148*32) clamped to [128, 2048] */
int64_t han_csr_num_chunks(int64_t nnz);
int han_csr_chunk_rows(const int64_t* indptr, int64_t n_rows, int64_t nnz, int32_t* chunk_rows,
                       han_stream_t stream);
/* indptr may point at row r0 of a larger CSR (then n_rows / nnz are those of the sub-range and chunk_rows holds
 * row numbers relative to r0). */
int han_attn_fwd_chunked(const int64_t* indptr, const int32_t* indices, const int32_t* chunk_rows,
                         int64_t n_chunks, int64_t n_dst, const float* T, const float* a2, const float* b2,
                         float* R, const float* bias,
                         int K, int H, int act, float* out, int64_t out_stride, float* vsave,
                         const float* colmean, const float* edge_w, const float* resid, int64_t resid_stride,
                         float* const* out2_tab, int64_t out2_block_rows, int64_t out2_stride, float* vsave2,
                         float* csave, const uint32_t* seed_ptr, float coef_keep, float in_keep, int metapath,
                         int64_t row0, han_stream_t stream);
/* vsave2 [n_dst][D] / csave [n_dst][K] (both or neither; NULL for inference): the second aggregate kept for the
 * backward.  With k_ij = leaky_relu'(l_ij) (times w_ij with edge weights),
 *     V'_i = sum_j alpha~_ij k_ij S_j          c_i = sum_j alpha_ij k_ij
 * make df1_i = sum_j dl_ij = <dV_i, V'_i> - delta_i c_i a ROW-LOCAL quantity (han_attn_bwd_prep): the backward needs no
 * per-edge dl array, no by-destination pass and, sharded, no reduce-scatter of df1. */
/* out2_tab (nullable): DEVICE array of base pointers, one per block of out2_block_rows destination rows: row i is also
 * stored to out2_tab[i / out2_block_rows] + (i % out2_block_rows) * out2_stride.  In tile-sharded multi-GPU runs the
 * pointers are peer-mapped addresses inside the other GPUs' semantic-layer inputs (symmetric memory): the re-sharding
 * all-to-all of Z is fused into K-B's epilogue as plain stores over NVLink (models/gat.py:58-60 across GPUs). */
/* resid (nullable) [n_dst][resid_stride]: the residual term of utils/layers.py:38-40, added before the
 * activation: out_i = act(V_i + bias + resid_i).  Its gradient is dV (R[:, 0:D] after han_attn_bwd_prep). */
int han_attn_bwd_src_chunked(const int64_t* t_indptr, const int32_t* t_indices,
                             const int32_t* chunk_rows, int64_t n_chunks, int64_t n_src,
                             const float* Tsrc, const float* a2, const float* b2, const float* R, int K, int H,
                             float* dS_agg, float* df2, const float* edge_w_t, const uint32_t* seed_ptr,
                             float coef_keep, float in_keep, int metapath, int64_t row0, han_stream_t stream);
/* edge_w_t (nullable) [nnz]: the edge weights in TRANSPOSED-edge order (edge_w[perm[t]]); dl then carries the
 * factor w_ij (d l_ij / d f1_i = d l_ij / d f2_j = w_ij). */
/* Training-mode dropout of the attention coefficients (utils/layers.py:29-30: coefs scaled 1/keep where
 * kept, zeroed elsewhere, NOT re-normalised): coef_keep = 1 - coef_drop in (0,1]; 1 disables it.  The
 * mask bit of edge (dst i, src j), head k of meta-path `metapath` is a pure function of (*seed_ptr, i, j,
 * k, metapath) (han_rng.cuh), recomputed identically by the forward, the backward and every rank; row0
 * is the global id of local row 0.  seed_ptr is a DEVICE word so a captured CUDA graph can advance it.
 * Training-mode dropout of the projected features (utils/layers.py:31-32): in_keep = 1 - ffd_drop in (0,1]; 1 disables
 * it.  The table keeps the UN-dropped S (f1 / f2 come from it, :23-24); the gather kernels mask the rows they fetch --
 * bit of (node j, column d) a pure function of (*seed_ptr, j, d, metapath), j a GLOBAL node id (the CSR's column ids
 * forward, row0 + local source row backward) -- so the aggregate (:33) and its gradient see the dropped S. */

/* ---- K-D: backward of K-B ----------------------------------------------------------------- */
/* prep (row-local): dV = dout * act'(.), delta = <dV, V> per head -> R[:, 0:D], R[:, D+2K:D+3K];
 * dbias_partial [han_reduce_blocks()][D] per-block column sums of dV. */
int han_reduce_blocks(void);
int han_attn_bwd_prep(const float* dout, int64_t dout_stride, const float* out, int64_t out_stride,
                      const float* vsave, float* R, int64_t n_dst, int K, int H, int act,
                      float* dbias_partial, float* R_mc, int64_t r_row0, const float* vsave2, const float* csave,
                      float* df1, han_stream_t stream);
/* df1 (nullable) [n_dst][K] = <dV, vsave2> - delta * csave per head: the gradient w.r.t. f1 (see han_attn_fwd_chunked). */
/* R_mc (nullable): multicast address of a symmetric record table [rows][RS]; when given, the complete
 * record (dV | f1 | lse | delta) of local row i is written to row r_row0 + i of every rank's copy
 * (prep fused with the all-gather of the records); f1 and lse are read from the local R. */

/* by-source pass over the transposed structure (han_attn_bwd_src_chunked above): for source rows [0,n_src):
 *   dS_agg_j = sum_i alpha_ij dV_i ; df2_j = sum_i dl_ij
 * where dl_ij = alpha_ij (dV_i.S_j - delta_i) * leaky'(f1_i + f2_j).  Tsrc rows = local sources. */

/* Heavy rows (power-law meta-paths).  The same two passes over a VIRTUAL-row CSR: indptr_v [n_v+1] is the
 * CSR's offsets with cut points inserted so that no row exceeds a fixed number of edges (column array and
 * perm unchanged; chunk_rows built over indptr_v).  vmap [n_v][2] = (real row, partial slot or -1 for a row
 * that was not cut); part receives the per-segment partial state: [n_slots][K][2H+3] forward (running max, normaliser,
 * c, the two un-normalised aggregates), [n_slots][K][H+2] backward (df2 and dS sums); heavy_rows [n_heavy] / heavy_ptr
 * [n_heavy+1] list the cut rows and their slot ranges for the merge kernel (one warp per cut row), which is
 * launched right after the stream kernel.  Results are identical to the un-split entry points up to the
 * order of floating-point additions inside a cut row. */
int han_attn_fwd_chunked_split(const int64_t* indptr_v, const int32_t* indices, const int32_t* chunk_rows,
                               int64_t n_chunks, int64_t n_dst, const float* T, const float* a2, const float* b2,
                               float* R, const float* bias,
                               int K, int H, int act, float* out, int64_t out_stride, float* vsave,
                               const float* colmean, const float* edge_w, const float* resid, int64_t resid_stride,
                               float* const* out2_tab, int64_t out2_block_rows, int64_t out2_stride, float* vsave2,
                               float* csave, const uint32_t* seed_ptr, float coef_keep, float in_keep, int metapath,
                               int64_t row0, const int32_t* vmap,
                               float* part, const int32_t* heavy_rows, const int32_t* heavy_ptr, int n_heavy,
                               han_stream_t stream);
int han_attn_bwd_src_chunked_split(const int64_t* t_indptr_v, const int32_t* t_indices,
                                   const int32_t* chunk_rows, int64_t n_chunks, int64_t n_src,
                                   const float* Tsrc, const float* a2, const float* b2, const float* R, int K, int H,
                                   float* dS_agg, float* df2, const float* edge_w_t, const uint32_t* seed_ptr,
                                   float coef_keep, float in_keep, int metapath, int64_t row0, const int32_t* vmap,
                                   float* part,
                                   const int32_t* heavy_rows, const int32_t* heavy_ptr, int n_heavy,
                                   han_stream_t stream);

/* finish (row-local): dS_tot = dS_agg + df1 a1^T + df2 a2^T (in place into dS_agg);
 * partial sums for da1,da2 [K][H], db1,db2 [K]: part [han_reduce_blocks()][2*D + 2*K]. */
int han_attn_bwd_finish(const float* T, int64_t n, int K, int H, const float* a1, const float* a2,
                        const float* df1, const float* df2, float* dS, float* part,
                        const uint32_t* seed_ptr, float in_keep, int metapath, int64_t row0,
                        han_stream_t stream);
/* in_keep < 1 (training-mode dropout of the projected features, utils/layers.py:31-32): dS_agg is the
 * gradient w.r.t. the dropped S and is passed through the same mask; T holds the un-dropped S. */

/* ---- training-mode projection with feed-forward dropout (utils/layers.py:18-19,31-32) ------------- */
/* Every head of every meta-path draws its own mask over the input features (one tf.nn.dropout per
 * attn_head call), so S_k = (X * m_k / keep) W_k.  T receives this S (f1 / f2 come from it); the second dropout
 * of :31-32 is applied by the gather kernels to the rows they fetch (in_keep of han_attn_fwd_chunked).  FP32 FFMA
 * with the mask bits generated while X is staged; K <= 8.  in_keep = 1 - ffd_drop in (0,1). */
int han_project_fwd_drop(const float* X, int64_t n, int64_t F, int64_t ldx, const float* W, int64_t ldw,
                         int G, int K, int H, const float* a1, const float* b1, float* T, float* R,
                         const uint32_t* seed_ptr, float in_keep, int metapath0, int64_t row0, han_stream_t stream);
size_t han_project_bwd_drop_workspace_bytes(int64_t n, int64_t F, int D);
int han_project_bwd_drop(const float* X, int64_t n, int64_t F, int64_t ldx, const float* dS, int G, int K,
                         int H, float* dW, int64_t ldw, void* ws, size_t ws_bytes, const uint32_t* seed_ptr,
                         float in_keep, int metapath0, int64_t row0, han_stream_t stream);

/* Gradient of the projection w.r.t. its input, for stacked attention layers (models/gat.py:48-57 feed the
 * concatenated heads of one layer to the next; utils/layers.py:18-20 for the per-head input dropout):
 *   dX[n][f] (+)= sum_k m_k(n,f)/keep * sum_h dS[n][k*H+h] * W[f][k*H+h]          (one meta-path per call)
 * dS [n][K*H]; W points at the meta-path's first column of the (F x ldw) weight matrix; dX [n][ldx];
 * accumulate != 0 adds to dX (meta-paths that share an input).  seed_ptr (device uint32, nullable) and
 * in_keep < 1 regenerate the forward's masks of han_project_fwd_drop; otherwise plain dS W^T. */
int han_project_dx(const float* dS, int64_t n, int K, int H, const float* W, int64_t ldw, int64_t F,
                   float* dX, int64_t ldx, int accumulate, const void* seed_ptr, float in_keep,
                   int metapath, int64_t row0, han_stream_t stream);

/* Deterministic column sums of partial buffers: outv[c] = sum_b part[b][c]. */
int han_reduce_partials(const float* part, int nblocks, int64_t cols, float* outv, han_stream_t stream);

/* ---- Training update. Replaces models/base_gattn.py:12-24 (L2 on every variable + tf.train.AdamOptimizer) */
/* p, g, m, v: flat FP32 buffers of n elements (n % 4 == 0, 16-byte aligned) holding every trainable variable
 * at padded offsets.  step_ptr: device int32, the 1-based step count t of THIS update (the caller bumps it
 * on the stream before the call, so a captured CUDA graph advances it too).
 *   g' = g + l2_coef*p;  m = b1 m + (1-b1) g';  v = b2 v + (1-b2) g'^2;
 *   p -= lr*sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v) + eps)        (epsilon outside the bias correction) */
int han_adam_l2_step(float* p, const float* g, float* m, float* v, int64_t n, const int* step_ptr,
                     float lr, float beta1, float beta2, float eps, float l2_coef, han_stream_t stream);

/* ---- Classifier and training loss. Replace models/gat.py:66-72 (tf.layers.dense) and models/base_gattn.py:41-48 ------ */
/* Exact-FP32 kernels (han_b200/csrc/dense_ce.cu) for D <= 64, C <= 384.
 * han_dense_fwd: Y [n][C] = X [n][ldx] (D columns used) * W [D][C] + b [C].
 * han_dense_bwd: dX [n][D] (nullable) = dY W^T; per-CTA partials of dW [D][C] | db [C] into
 *   part [han_dense_blocks()][D*C + C] (reduce with han_reduce_partials); scale (nullable): device scalar multiplied
 *   into dY (the upstream gradient of a scalar loss).
 * han_masked_ce: loss = sum_i mask_i * (-sum_c labels_ic log_softmax(logits_i)_c) / *mask_total as per-CTA partials
 *   loss_part [han_dense_blocks()]; dlogits (nullable) [n][C] = (softmax_i * sum_c labels_ic - labels_i) * mask_i /
 *   *mask_total.  mask_total is a DEVICE scalar (sum of the mask over ALL ranks in sharded runs). */
int han_dense_blocks(void);
int han_dense_fwd(const float* X, int64_t n, int D, int64_t ldx, const float* W, int C, const float* b, float* Y,
                  han_stream_t stream);
int han_dense_bwd(const float* X, int64_t n, int D, int64_t ldx, const float* W, int C, const float* dY,
                  const float* scale, float* dX, float* part, han_stream_t stream);
int han_masked_ce(const float* logits, const float* labels, const float* mask, const float* mask_total, int64_t n,
                  int C, float* loss_part, float* dlogits, han_stream_t stream);

/* ---- K-C / K-F: semantic attention. Replaces utils/layers.py:152-159 ------------------------ */
/* Z [n][P][D], w [D][A], b [A], u [A].  Supported (D,A): han_semantic_shape_supported.
 * HAN_SEM_REFERENCE: writes out [n][D] and beta [n][P] (per-node softmax over meta-paths, :156).
 * HAN_SEM_PAPER    : writes only scores [n][P]; the caller averages them over nodes (all-reduce when
 *                    sharded), takes one softmax, then calls han_semantic_combine.
 * vsave [n*P][A] (nullable) keeps tanh(Zw+b) for the backward; scores (nullable in reference mode). */
int han_semantic_shape_supported(int D, int A);
int han_semantic_fwd(const float* Z, int64_t n, int P, int D, int A, const float* w, const float* b,
                     const float* u, int mode, float* out, float* beta, float* vsave, float* scores,
                     han_stream_t stream);
/* Default for D = 64, A = 128 (HAN_SEM_TC=0 selects han_semantic_fwd): the same forward on
 * tcgen05 tensor cores -- persistent CTAs, w^T resident in shared memory, TMA ring for Z, 3xTF32 accumulation in
 * double-buffered TMEM, epilogue of tile i under the MMAs of tile i+1 (han_b200/csrc/semantic_tc.cu).
 * ws: han_semantic_tc_workspace_bytes() bytes (the transposed hi/lo split of w).  epilogue_groups: 1, 2 or 4
 * (4 / 8 / 16 epilogue warps; groups share each row's columns). */
size_t han_semantic_tc_workspace_bytes(void);
int han_semantic_fwd_tc(const float* Z, int64_t n, int P, int D, int A, const float* w, const float* b,
                        const float* u, int mode, float* out, float* beta, float* vsave, float* scores, void* ws,
                        size_t ws_bytes, int epilogue_groups, han_stream_t stream);
/* K-F on tcgen05 tensor cores (D = 64, A = 128; han_b200/csrc/semantic_tc.cu): v = tanh(Z w + b) is RECOMPUTED from Z
 * (no vsave: the forward need not store it), dv w^T and Z^T dv run as 3xTF32 tcgen05.mma from one shared-memory copy
 * of each operand (MN-major descriptors for the transposed uses), dw accumulates in TMEM with a drain every 256 rows.
 * Same outputs and dz_tab routing as han_semantic_bwd.  ws: han_semantic_bwd_tc_workspace_bytes(). */
size_t han_semantic_bwd_tc_workspace_bytes(void);
int han_semantic_bwd_tc(const float* dout, const float* Z, const float* beta, int64_t n, int P, int D, int A,
                        const float* w, const float* b, const float* u, int mode, const float* dsbar, float* dZ,
                        float* dw, float* db, float* du, void* ws, size_t ws_bytes, float* const* dz_tab,
                        int64_t dz_stride, han_stream_t stream);
/* out[n] = sum_p beta_vec[p] Z[n,p]; beta (nullable) [n][P] receives the broadcast (han.pdf Eq. 9). */
int han_semantic_combine(const float* Z, int64_t n, int P, int D, const float* beta_vec, float* out,
                         float* beta, han_stream_t stream);

size_t han_semantic_bwd_workspace_bytes(int P, int D, int A);
/* dout [n][D] -> dZ [n][P][D], dw [D][A], db [A], du [A] (deterministic two-stage reduce).
 * dsbar [P] (paper mode only): d(loss)/d(s_bar_p) / N, the per-row score gradient. */
int han_semantic_bwd(const float* dout, const float* Z, const float* beta, const float* vsave,
                     int64_t n, int P, int D, int A, const float* w, const float* u, int mode,
                     const float* dsbar, float* dZ, float* dw, float* db, float* du, void* ws,
                     size_t ws_bytes, float* const* dz_tab, int64_t dz_stride, han_stream_t stream);
/* dz_tab (nullable): DEVICE array of P base pointers; when given, the gradient row of (node, meta-path p) goes to
 * dz_tab[p] + node * dz_stride instead of dZ[node][p][:] (dZ may then be NULL).  Tile-sharded multi-GPU runs pass
 * peer-mapped addresses inside the GPUs that own each meta-path, fusing the all-to-all of dZ into this kernel. */

#ifdef __cplusplus
}
#endif
#endif /* HAN_B200_H */
