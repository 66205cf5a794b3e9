/*
 * tdrb200.h — C ABI of libtdrb200.so, the B200 (sm_100a) engine for TorchDR's
 * neighbor-embedding hot path.
 *
 * Every entry point is what a binding for the reference's plugin seams would
 * call; the reference interface each one replaces is cited (paths relative to
 * the TorchDR tree).  Conventions:
 *   - all data pointers are DEVICE pointers into caller-owned buffers (PyTorch
 *     tensors in the Python host), row-major, fp32 unless stated;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*);
 *   - return value: 0 = ok, < 0 = error (TDR_E_*); the message is available
 *     from tdr_last_error() (thread-local);
 *   - the library allocates nothing: scratch comes from a caller-provided
 *     workspace whose size the *_workspace_bytes twin reports;
 *   - no torch types, no NCCL dependency (collectives stay in the host language,
 *     torch.distributed), and no global state: every option is a per-call argument;
 *     the only process-wide datum is the thread-local error string.
 */
#ifndef TDRB200_H
#define TDRB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TDR_OK 0
#define TDR_E_INVALID (-1)   /* bad argument (shape, k, metric ...) */
#define TDR_E_WORKSPACE (-2) /* workspace too small / misaligned */
#define TDR_E_CUDA (-3)      /* CUDA runtime error, see tdr_last_error() */
#define TDR_E_UNSUPPORTED (-4)

#define TDR_METRIC_SQEUCLIDEAN 0
#define TDR_METRIC_EUCLIDEAN 1

#define TDR_MAX_K 160 /* top-k lists live in shared memory */

typedef void* tdr_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define TDR_API __attribute__((visibility("default")))
#else
#define TDR_API
#endif

/* ---- library ----------------------------------------------------------- */
TDR_API int tdr_abi_version(void);
TDR_API const char* tdr_last_error(void);
/* sm count, compute capability of the current device; TDR_E_UNSUPPORTED unless sm_100 */
TDR_API int tdr_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- (i) distances / exact kNN ------------------------------------------
 * Replaces pairwise_distances(..., k=k, backend=None) -> pairwise_distances_torch
 * (torchdr/distance/base.py:22-249, distance/torch.py:21-125) + kmin
 * (utils/utils.py:173-216): expanded form ||x||^2+||y||^2-2xy in fp32, self
 * excluded by global row id (torch.py:111-116 adds 1e12 on the diagonal),
 * k smallest per row in ascending order, int32 indices.  Ties: lower index
 * first.  Xq are rows [q_row0, q_row0+nq) of the database when exclude_self
 * (the distributed chunk of distance/base.py:184-186). */
TDR_API size_t tdr_knn_workspace_bytes(int64_t nq, int64_t ndb, int d, int k);

/* Per-call options of the two kNN entry points (no process-wide switches):
 * path   kernel selection: AUTO = tcgen05 tensor-core kernel (fp16 hi/lo split) whenever the tile shapes fit
 *        (d <= 256, k <= 96 and query tile + lists + a two-stage ring <= 227 KB of shared memory: k <= 96 at
 *        d <= 128, k <= 33 at d <= 256), else the fp32 SIMT kernel; SIMT / TC force one (TC returns TDR_E_UNSUPPORTED if the shape does not fit).
 * prune  tile-pruned sweep of the tensor-core kernel (queries that are rows of the database, >= 64 database
 *        tiles): database tiles whose bounding box is farther from the query tile's box than a bound on the
 *        tile's k-th neighbour distances are not swept.  Results are bit-identical to the full sweep.
 *        OFF = sweep every tile; ON (= DEFAULT); CERTIFIED = thresholds that ignore outlier rows, then a
 *        certification pass and a second sweep of the uncertified query tiles — for inputs that were brought
 *        into a locality-creating order first (torchdr_b200/reorder.py), also bit-identical.
 * sweep_stats  optional device pointer to two uint64 counters, [0] += tiles swept, [1] += tiles a full sweep
 *        would visit (summed over query tiles); NULL disables them.
 * db_labels    optional int32[ndb]: the id reported in out_idx for database row c is db_labels[c], and candidates
 *        that tie in distance are ranked by that id — the k neighbours are the k smallest by (distance, label)
 *        whatever the order of the database rows.  Used when the database was re-ordered (torchdr_b200/reorder.py)
 *        and labels = the rows' ids in the caller's order: the result then does not depend on the re-ordering.
 *        NULL: the row index is the label.  (The fp32 SIMT kernel re-labels after the search: ties by row index.) */
#define TDR_KNN_PATH_AUTO 0
#define TDR_KNN_PATH_SIMT 1
#define TDR_KNN_PATH_TC 2
#define TDR_KNN_PRUNE_DEFAULT (-1)
#define TDR_KNN_PRUNE_OFF 0
#define TDR_KNN_PRUNE_ON 1
#define TDR_KNN_PRUNE_CERTIFIED 2
TDR_API int tdr_knn_f32(const float* Xq, int64_t nq, int64_t q_row0,
                const float* Xdb, int64_t ndb, int d, int k,
                int exclude_self, int metric,
                float* out_dist /*[nq,k]*/, int32_t* out_idx /*[nq,k]*/,
                int path, int prune, uint64_t* sweep_stats, const int32_t* db_labels,
                void* ws, size_t ws_bytes, tdr_stream_t stream);

/* Full matrix C[n,m] (k=None path, distance/torch.py:81-116).  Y may equal X.  path as above: AUTO = the
 * tcgen05 kernel with a dense epilogue for d <= 256, else the fp32 SIMT tile kernel.
 * Workspace: tdr_knn_workspace_bytes(n, m, d, 1). */
TDR_API int tdr_pairwise_full_f32(const float* X, int64_t n, const float* Y, int64_t m, int d,
                          int metric, int exclude_diag, float* C /*[n,m]*/, int path,
                          void* ws, size_t ws_bytes, tdr_stream_t stream);

/* pairwise_distances_indexed, per-query key lists (distance/base.py:252-405, 2-D key_indices):
 * out[i,s] = dist(X[query_idx[i]], Y[key_idx[i,s]]) in the exact-difference form of base.py:384-385.
 * query_idx may be NULL (identity); key_idx is int32 or int64 [nq,k]; negative keys wrap as in torch. */
TDR_API int tdr_indexed_dist_f32(const float* X, const int64_t* query_idx, int64_t nq,
                                 const float* Y, int64_t ny, int d,
                                 const void* key_idx, int key_is_int64, int k, int metric,
                                 float* out /*[nq,k]*/, tdr_stream_t stream);

/* Row ordering for inputs without index locality (torchdr_b200/reorder.py; no counterpart in the reference, which
 * hands unordered rows to FAISS): nearest-centre assignment of one level of the Voronoi tree.  Entry i is row rows[i]
 * of X[., d], a member of tree node node[i]; centres is [n_nodes, B, d], cnorm [n_nodes, B] holds |c|^2 (+inf = not a
 * centre); child_out[i] = argmin_c |c|^2 - 2 x.c, ties to the lower c.  d <= 512, B <= 64. */
TDR_API int tdr_tree_assign_f32(const float* X, int d, const int64_t* rows, const int64_t* node, int64_t m,
                                const float* centres, const float* cnorm, int B, int64_t* child_out,
                                tdr_stream_t stream);
/* Lloyd step of the same tree: sums[(node, child), d] and cnt[(node, child)] of the member rows (entries grouped by
 * node as above; child[i] in [0, B), B <= 16, d <= 512).  Both outputs are zeroed by the call and must hold
 * n_nodes + TDR_TREE_SPARE_NODES nodes (a tile's shared-memory accumulators are flushed for up to that many nodes past
 * the tile's first). */
#define TDR_TREE_SPARE_NODES 2
TDR_API int tdr_tree_accumulate_f32(const float* X, int d, const int64_t* rows, const int64_t* node,
                                    const int64_t* child, int64_t m, int64_t n_nodes, int B,
                                    float* sums, float* cnt, tdr_stream_t stream);

/* ---- (ii) per-row bandwidth search --------------------------------------
 * UMAPAffinity rows: rho = row min, sigma by bracket+bisection
 * (torchdr/affinity/knn_normalized.py:445-468, utils/root_search.py:17-198),
 * P = exp(-(C-rho)/sigma).  C rows are the kNN distances. */
TDR_API int tdr_umap_affinity_f32(const float* C /*[n,k]*/, int64_t n, int k, int max_iter,
                          float* P /*[n,k]*/, float* rho /*[n]*/, float* sigma /*[n]*/,
                          tdr_stream_t stream);

/* EntropicAffinity rows (torchdr/affinity/entropic.py:230-312): eps by
 * bisection on the row entropy, log_P = -C/eps - logsumexp - log(n_total).
 * use_bounds selects the Vladymyrov bracket (entropic.py:51-115); its four
 * data-independent scalars (host-computed, fp32) are
 *   b_num = tN*log(tN/perp), b_den = tN-1, b_lr = log(tN/perp),
 *   b_logp1 = log((tN-1)*p1/(1-p1)). */
TDR_API int tdr_entropic_affinity_f32(const float* C /*[n,k]*/, int64_t n, int k,
                              float target_entropy, float log_n_total,
                              int use_bounds, float b_num, float b_den, float b_lr, float b_logp1,
                              int max_iter,
                              float* logP /*[n,k]*/, float* eps /*[n]*/, float* log_norm /*[n]*/,
                              tdr_stream_t stream);

/* Dense variant (EntropicAffinity(sparsity=False), entropic.py:266-268): C is the full
 * [n_rows, m] distance matrix (tdr_pairwise_full_f32, diagonal carrying the 1e12 penalty).  One
 * CTA per row, one streaming pass over the row per bisection step.  logP may be NULL (only eps /
 * log_norm wanted) or alias C (in place). */
TDR_API int tdr_entropic_dense_f32(const float* C /*[n_rows,m]*/, int64_t n_rows, int64_t m,
                                   float target_entropy, float log_n_total,
                                   int use_bounds, float b_num, float b_den, float b_lr, float b_logp1,
                                   int max_iter,
                                   float* logP /*[n_rows,m] or NULL*/, float* eps, float* log_norm,
                                   tdr_stream_t stream);

/* Fused (i)+(ii): kNN mainloop with the UMAP sigma/rho search in the epilogue
 * (the "affinity kernel" of BASELINE.json).  Same workspace as tdr_knn_f32. */
TDR_API int tdr_knn_umap_fused_f32(const float* Xq, int64_t nq, int64_t q_row0,
                           const float* Xdb, int64_t ndb, int d, int k,
                           int exclude_self, int max_iter,
                           float* out_dist /*[nq,k] or NULL*/, int32_t* out_idx /*[nq,k]*/,
                           float* P /*[nq,k]*/, float* rho /*[nq]*/, float* sigma /*[nq]*/,
                           int path, int prune, uint64_t* sweep_stats, const int32_t* db_labels,
                           void* ws, size_t ws_bytes, tdr_stream_t stream);

/* ---- graph stage ---------------------------------------------------------
 * symmetrize_sparse(mode="sum_minus_prod") (torchdr/utils/sparse.py:170-206):
 * Q = P + P^T - P o P^T on the kNN graph, output as CSR of the local rows with
 * ascending columns (the reference's padded ELL is CSR + padding, sparse.py:118-140).
 * ext_* (may be NULL, n_ext=0) are transposed edges received from other ranks
 * (distributed_symmetrize_sparse, sparse.py:209-342): ext_row is the GLOBAL row
 * (owned locally), ext_col the global column, ext_val = P[col,row].
 * Capacity of col/val must be >= 2*n_local*k + n_ext.  *nnz_out is a device int64.
 * mode: TDR_SYM_SUM_MINUS_PROD (UMAP, sparse.py:163-164) or TDR_SYM_SUM = P + P^T (sparse.py:159-160; the union graph
 * the row-local LargeVis step pulls from). */
#define TDR_SYM_SUM_MINUS_PROD 0
#define TDR_SYM_SUM 1
TDR_API size_t tdr_symmetrize_workspace_bytes(int64_t n_local, int k, int64_t n_ext);
TDR_API int tdr_symmetrize_csr_f32(const float* P /*[n_local,k]*/, const int32_t* idx /*[n_local,k]*/,
                           int64_t n_local, int k, int64_t row0, int64_t n_total,
                           const int64_t* ext_row, const int32_t* ext_col, const float* ext_val,
                           int64_t n_ext, int transpose_local, int mode,
                           int64_t* rowptr /*[n_local+1]*/, int32_t* col, float* val,
                           int64_t* nnz_out, void* ws, size_t ws_bytes, tdr_stream_t stream);

/* Edges j of local rows whose transposed entry belongs to another rank: counts per
 * destination and packed (row=j global, col=i global, val) triples ordered by rank.
 * Used by the host to build the all_to_all of sparse.py:259-309 with int indices. */
TDR_API int tdr_symmetrize_export_f32(const float* P, const int32_t* idx, int64_t n_local, int k,
                              int64_t row0, int64_t n_total, int world, int rank,
                              int64_t* send_counts /*[world] device*/,
                              int64_t* out_row, int32_t* out_col, float* out_val /*[n_local*k]*/,
                              tdr_stream_t stream);

/* CSR -> (values[n,W], indices[n,W] int64, -1 padded) for the SparseAffinity seam
 * (affinity/base.py:407-431). */
TDR_API int tdr_csr_to_ell_f32(const int64_t* rowptr, const int32_t* col, const float* val,
                       int64_t n_local, int64_t width, float pad_val,
                       float* ell_val, int64_t* ell_idx, tdr_stream_t stream);

/* max over val[0..nnz) -> *out (device float); UMAP.on_affinity_computation_end, umap.py:218 */
TDR_API int tdr_max_f32(const float* val, int64_t nnz, float* out, tdr_stream_t stream);

/* UMAP edge schedule (umap.py:215-234): epochs_per_sample = A_max/(val+1e-3), inf when
 * val <= A_max/max_iter; epoch_of_next_sample = copy.  a_max is a host float. */
TDR_API int tdr_umap_schedule_f32(const float* val, int64_t nnz, float a_max, int max_iter,
                          float* epochs_per_sample, float* epoch_of_next_sample,
                          tdr_stream_t stream);

/* Drop never-sampled edges (epochs_per_sample == inf): compact CSR for the step kernel.
 * out_rowptr[n_local+1], out_col/out_eps/out_eons capacity nnz. */
TDR_API size_t tdr_compact_workspace_bytes(int64_t n_local, int64_t nnz);
TDR_API int tdr_umap_compact_f32(const int64_t* rowptr, const int32_t* col, const float* eps,
                         int64_t n_local, int64_t nnz,
                         int64_t* out_rowptr, int32_t* out_col, float* out_eps, float* out_eons,
                         int64_t* nnz_out, void* ws, size_t ws_bytes, tdr_stream_t stream);

/* ---- (iii) optimisation steps -------------------------------------------
 * One UMAP iteration for local rows [row0, row0+n_local): attraction
 * (umap.py:236-264) with the epoch_of_next_sample update, repulsion
 * (umap.py:266-292) on the first 5*active negatives, clamp +-4 each, then the
 * SGD update z -= lr*g (affinity_matcher.py:427; plain SGD, umap.py:139).
 * Jacobi: reads Z_in (all N rows), writes Z_out[row0..] (may not alias Z_in).
 * neg: int64 [n_local, n_neg] adjusted negatives (NE base.py:629-636) or NULL to
 * draw them in-kernel with Philox4x32-7 keyed by (seed, n_iter, global row, slot / 4).
 * a, b are the python doubles of umap.py:19-36 (the kernel derives the fp32 constants
 * (float)a, (float)b, (float)(b-1), (float)(2ab), (float)(-2b) exactly as the torch ops do).
 * precise = 1 evaluates pow in fp64, one warp per row (parity mode); 0 = throughput kernel.
 * grad_out (nullable) [n_local,2] receives the gradient; gnorm_sq (nullable,
 * device double) accumulates ||g||^2; nan_flag (nullable, device int) is set
 * if a NaN is written (check_NaNs, affinity_matcher.py:315); stats (nullable, device
 * uint64[2]) accumulates the number of sampled edges and of negatives used. */
TDR_API int tdr_umap_step_f32(const float* Z_in, float* Z_out, int64_t n_total, int64_t row0, int64_t n_local,
                      const int64_t* rowptr, const int32_t* col,
                      const float* epochs_per_sample, float* epoch_of_next_sample,
                      const int64_t* neg, int n_neg, int negative_sample_rate,
                      uint64_t seed, int64_t n_iter,
                      double a, double b, float lam, float repulsion, float lr,
                      int precise, float* grad_out, double* gnorm_sq, int* nan_flag,
                      uint64_t* stats, tdr_stream_t stream);

/* n_steps single-GPU iterations ping-ponging Z_a <-> Z_b (in-kernel negatives) in ONE persistent
 * cooperative launch per 128 iterations: the iterations are separated by a grid barrier inside the
 * kernel.  lrs is a HOST array [n_steps].  The result is in Z_a if n_steps is even, else Z_b.
 * sync_words: caller-owned device buffer of TDR_RUN_SYNC_WORDS uint32, zero-initialised once and
 * reused by every call on the same embedding (barrier counters, work counters, status word
 * sync_words[TDR_RUN_STATUS_WORD]: 0 = ok); epoch0 = total number of iterations already run on
 * this sync buffer.  precise != 0 runs the parity kernel with one launch per iteration. */
#define TDR_RUN_SYNC_WORDS 8
#define TDR_RUN_STATUS_WORD 4
TDR_API int tdr_umap_run_f32(float* Z_a, float* Z_b, int64_t n_total,
                     const int64_t* rowptr, const int32_t* col,
                     const float* epochs_per_sample, float* epoch_of_next_sample,
                     int n_neg, int negative_sample_rate, uint64_t seed, int64_t n_iter0,
                     int n_steps, const float* lrs_host,
                     double a, double b, float lam, float repulsion,
                     int precise, double* gnorm_sq, int* nan_flag, uint64_t* stats,
                     uint32_t* sync_words, uint32_t epoch0, tdr_stream_t stream);

/* Fused step + exchange for the row-sharded multi-GPU loop, one launch per iteration: same iteration
 * as tdr_umap_step_f32 (throughput kernel, in-kernel negatives), but every updated row is additionally
 * stored into the Z_out buffer of each peer GPU through NVLink peer mappings (peer_out_ptrs: HOST
 * array of n_peers <= 8 device addresses, e.g. torch symmetric-memory buffer_ptrs without this
 * rank's own).  The caller provides the cross-GPU barrier before the next iteration reads the buffers. */
TDR_API int tdr_umap_step_p2p_f32(const float* Z_in, float* Z_out, int64_t n_total, int64_t row0, int64_t n_local,
                                  const int64_t* rowptr, const int32_t* col,
                                  const float* epochs_per_sample, float* epoch_of_next_sample,
                                  int n_neg, int negative_sample_rate, uint64_t seed, int64_t n_iter,
                                  double a, double b, float lam, float repulsion, float lr,
                                  double* gnorm_sq, int* nan_flag,
                                  const uint64_t* peer_out_ptrs, int n_peers, tdr_stream_t stream);

/* n_steps row-sharded iterations in ONE persistent cooperative launch per 128 iterations — step, exchange
 * and barrier in the same kernel (replaces the per-iteration collective of affinity_matcher.py:395-413):
 * every updated row is stored into the destination buffer of this rank AND of each NVLink peer; at the end
 * of an iteration the last CTA of the rank announces it in the peers' flag words (st.release.sys) and waits
 * for theirs (ld.acquire.sys) before it releases the local grid barrier.  Z_a/Z_b are this rank's two
 * embedding buffers, peers_a/peers_b the peers' addresses of the same two buffers (HOST arrays,
 * n_peers = world - 1 entries, ranks ascending without this rank); my_flags is this rank's zero-initialised
 * uint32[world] flag buffer and peer_flags the peers' addresses of theirs; sync_words / epoch0 as in
 * tdr_umap_run_f32 (epoch0 must be the same on every rank).  A peer that does not arrive within timeout_s
 * seconds (<= 0: 60 s) aborts the run on this rank: sync_words[TDR_RUN_STATUS_WORD] becomes 1 (2 = a local
 * CTA was missing) and the kernel exits instead of spinning forever; the buffers are then undefined.
 * Result in Z_a if n_steps is even. */
TDR_API int tdr_umap_run_p2p_f32(float* Z_a, float* Z_b, int64_t n_total, int64_t row0, int64_t n_local,
                                 const int64_t* rowptr, const int32_t* col,
                                 const float* epochs_per_sample, float* epoch_of_next_sample,
                                 int n_neg, int negative_sample_rate, uint64_t seed, int64_t n_iter0,
                                 int n_steps, const float* lrs_host,
                                 double a, double b, float lam, float repulsion,
                                 double* gnorm_sq, int* nan_flag, uint64_t* stats, uint32_t* sync_words,
                                 const uint64_t* peers_a, const uint64_t* peers_b,
                                 uint32_t* my_flags, const uint64_t* peer_flags, int n_peers,
                                 int rank, int world, uint32_t epoch0, double timeout_s, tdr_stream_t stream);

/* LargeVis gradient (largevis.py:181-201 differentiated): accumulates into
 * grad[n_total,2] (zeroed by the caller) with atomics — the autograd scatter of
 * affinity_matcher.py:418-425 — for local rows; P/idx are the directed kNN rows. */
TDR_API int tdr_largevis_grad_f32(const float* Z, int64_t n_total, int64_t row0, int64_t n_local,
                          const float* P /*[n_local,k]*/, const int32_t* idx /*[n_local,k]*/, int k,
                          const int64_t* neg /*[n_local,n_neg] or NULL*/, int n_neg,
                          uint64_t seed, int64_t n_iter, float lam, float repulsion,
                          float* grad /*[n_total,2]*/, tdr_stream_t stream);

/* One LargeVis iteration for local rows [row0, row0+n_local) in ROW-LOCAL form — gradient, momentum SGD and row
 * exchange without the N x 2 all-reduce of affinity_matcher.py:418-425 (in-kernel negatives only):
 *   rowptr/col/val = CSR of S = P + P^T for the local rows (tdr_symmetrize_csr_f32, mode TDR_SYM_SUM): the attraction
 *   of row i, 2 lam S_ij Q_ij (z_i - z_j), then collects its own edges and the edges pointing at it in one gather sum;
 *   the push of a negative pair (i', j) onto j is computed by the owner of j, which re-generates the counter-based
 *   negative stream of all N rows and keeps the pairs that hit its rows (fp32 atomics into grad_scratch[n_local,2]);
 *   then buf = mu*buf + g (first: buf = g), z -= lr*buf on the local rows (torch.optim.SGD, NE base.py:331-343),
 *   written to Z_out (Jacobi: Z_in is read for all N rows) and to the Z_out of n_peers NVLink peers
 *   (peer_out_ptrs: HOST array of device addresses, may be NULL when n_peers = 0).
 * mom: [n_local,2] momentum buffer; gnorm_sq / nan_flag as in tdr_umap_step_f32. */
TDR_API int tdr_largevis_step_f32(const float* Z_in, float* Z_out, int64_t n_total, int64_t row0, int64_t n_local,
                                  const int64_t* rowptr, const int32_t* col, const float* val,
                                  int n_neg, uint64_t seed, int64_t n_iter, float lam, float repulsion,
                                  float* grad_scratch /*[n_local,2]*/, float* mom /*[n_local,2]*/,
                                  float lr, float momentum, int first,
                                  double* gnorm_sq, int* nan_flag,
                                  const uint64_t* peer_out_ptrs, int n_peers, tdr_stream_t stream);

/* t-SNE gradient (tsne.py:162-180 differentiated): sparse attraction on the kNN rows
 * plus the dense N x N repulsion of the logsumexp normaliser (expanded-form distances,
 * diagonal included).  Rows [row0,row0+n_local) are this rank's share of the double sum.
 *   phase 0: zero ws; U_i = sum_j w_ij^2 (z_i - z_j) and the partial normaliser S (a device
 *            double at ws[0]) for local rows; attraction scattered into grad (caller-zeroed).
 *   (host: all-reduce S when distributed)
 *   phase 1: grad[local rows] += -4 repulsion U_i / S   (repulsion = repulsion_strength, NE base.py:237-241).
 * ws: tdr_tsne_workspace_bytes(n_local), 16-byte aligned. */
TDR_API size_t tdr_tsne_workspace_bytes(int64_t n_local);
TDR_API int tdr_tsne_grad_f32(const float* Z, int64_t n_total, int64_t row0, int64_t n_local,
                      const float* P, const int32_t* idx, int k, float lam, float repulsion, int phase,
                      float* grad, void* ws, size_t ws_bytes, tdr_stream_t stream);

/* InfoTSNE gradient (infotsne.py:179-197 differentiated): t-SNE attraction on the kNN rows plus the
 * noise-contrastive repulsion (1/N) sum_i log sum_{s in Neg(i)} 1/(1+D_is); both scatter into row i
 * and the sampled rows like autograd's index_put(accumulate).  Arguments as tdr_largevis_grad_f32
 * (neg = NULL draws the n_neg negatives per row in-kernel, NE base.py:629-636); grad is caller-zeroed. */
TDR_API int tdr_infotsne_grad_f32(const float* Z, int64_t n_total, int64_t row0, int64_t n_local,
                          const float* P, const int32_t* idx, int k,
                          const int64_t* neg, int n_neg, uint64_t seed, int64_t n_iter,
                          float lam, float repulsion, float* grad /*[n_total,2]*/, tdr_stream_t stream);

/* SNE gradient (sne.py:162-179 differentiated): attraction lam * sum P_ij D_ij on the kNN rows plus the
 * dense repulsion (1/N) sum_i log sum_j exp(-C_ij) (expanded-form C on the embedding, diagonal included).
 *   phase 0: row_sums[row0 .. row0+n_local) = sum_j exp(-C_ij); attraction scattered into grad (caller-zeroed).
 *   (host: all-gather row_sums when distributed)
 *   phase 1: grad[local rows] += -(2 repulsion / N) sum_j e_ij (1/S_i + 1/S_j)(z_i - z_j).
 * row_sums: float [n_total]. */
TDR_API int tdr_sne_grad_f32(const float* Z, int64_t n_total, int64_t row0, int64_t n_local,
                     const float* P, const int32_t* idx, int k, float lam, float repulsion, int phase,
                     float* grad, float* row_sums, tdr_stream_t stream);

/* SGD with momentum (torch.optim.SGD semantics: buf = mu*buf + g; z -= lr*buf; first
 * step buf = g), NE base.py:331-343.  first != 0 initialises the buffer. */
TDR_API int tdr_sgd_momentum_f32(float* Z, float* buf, const float* grad, int64_t n_elems,
                         float lr, float momentum, int first,
                         double* gnorm_sq, int* nan_flag, tdr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TDRB200_H */
