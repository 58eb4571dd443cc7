/* ia_b200.h -- C ABI of the B200-native two-tower vector-similarity path.
 *
 * Drop-in boundary for the hot path of sunzeyeah/item-alignment (file:line below are into the
 * reference tree).  The reference is pure Python/PyTorch, so "the reference's FFI for this path"
 * is a ctypes binding: item_alignment_b200/_lib.py is that binding and INTEGRATION.md shows the
 * stub a reference maintainer would add.  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *   - every entry point returns 0 (IA_OK) or a negative ia_status; ia_last_error() gives the
 *     thread-local message.  Nothing here falls back to the CPU.
 *   - device entry points are asynchronous on `stream` (a cudaStream_t passed as void*), never
 *     allocate user-visible memory and never synchronise.  *_host entry points take HOST buffers,
 *     own their staging/streams and return after the result is on the host.
 *   - x, y are row-major [n, d] with leading dimensions ldx, ldy in ELEMENTS; sim/probs/loss are fp32.
 *   - labels are int64 {0,1} exactly as the reference's collate functions produce them
 *     (src/data/data.py:237); t = 2*label-1 is formed inside the kernels (src/models/text.py:1471,1475).
 */
#ifndef IA_B200_H
#define IA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* ia_stream_t; /* cudaStream_t */

typedef enum {
  IA_OK = 0,
  IA_ERR_INVALID = -1,     /* bad argument (reference raises ValueError, src/models/base.py:64,86) */
  IA_ERR_UNSUPPORTED = -2, /* dtype / shape not supported by the CUDA path */
  IA_ERR_CUDA = -3,        /* CUDA runtime / driver error */
  IA_ERR_WORKSPACE = -4    /* workspace too small */
} ia_status;

/* config.similarity_measure, src/models/base.py:54-62 */
typedef enum { IA_INNER = 0, IA_COSINE = 1, IA_L1 = 2, IA_L2 = 3 } ia_measure;
/* config.loss_type on a VecSim head, src/models/text.py:1400-1409 ("ce" belongs to the softmax head) */
typedef enum { IA_LOSS_BCE = 0, IA_LOSS_HINGE = 1, IA_LOSS_EUCLIDEAN = 2, IA_LOSS_COSINE = 3 } ia_loss;
typedef enum { IA_F32 = 0, IA_BF16 = 1, IA_F16 = 2, IA_F64 = 3 /* scores of ia_best_f1_threshold only */ } ia_dtype;
/* _Loss.reduction, src/models/loss.py:58-68,122-134 */
typedef enum { IA_RED_NONE = 0, IA_RED_MEAN = 1, IA_RED_SUM = 2 } ia_reduction;

const char* ia_version(void);
const char* ia_last_error(void);
/* Bytes of zero-initialised device scratch the *_fwd_bwd entry points need (deterministic two-stage
 * loss reduction).  One workspace per concurrently used stream; kernels leave it zeroed. */
size_t ia_workspace_bytes(void);
/* Number of kernels this library has launched since load (bench.py's gpu_launches evidence). */
int64_t ia_launch_count(void);

/* ---- pair scoring, forward only ------------------------------------------------------------
 * Replaces: InnerProduct.forward (src/models/base.py:29-34), nn.CosineSimilarity / nn.PairwiseDistance
 * as constructed at base.py:54-62, the probability map of VecSimClassificationHead.forward
 * (base.py:79-86) and the threshold labelling `probs >= threshold` (finetune_text.py:576-580;
 * the comparison is done in double like numpy does with an np.arange threshold).
 * probs and labels_out may be NULL. */
int ia_pair_score_fwd(int measure, int dtype, const void* x, const void* y, int64_t n, int64_t d,
                      int64_t ldx, int64_t ldy, float* sim, float* probs, double threshold,
                      uint8_t* labels_out, ia_stream_t stream);

/* ---- fused score + loss, forward AND backward in one HBM pass --------------------------------
 * Replaces, for one batch: the similarity + probs above, the loss ladder (src/models/text.py:1468-1477:
 * bce -> nn.BCEWithLogitsLoss on sim; hinge -> HingeLoss, src/models/loss.py:126-134; euclidean ->
 * EuclideanDistanceLoss, loss.py:61-68; cosine -> nn.CosineEmbeddingLoss on the embeddings) and the
 * autograd backward of all of it (finetune_text.py:479-482).
 *   loss_out  : 1 float (mean / sum) or n floats (IA_RED_NONE)
 *   dx, dy    : [n, d] gradients in grad_dtype (= dtype, or IA_F32), leading dims lddx / lddy;
 *               both NULL -> forward + loss only
 *   grad_scale: upstream d(total)/d(loss) known on the HOST, folded into dx, dy (1.0 for a plain loss.backward())
 *   upstream_dev: NULL, or a DEVICE float with the upstream scalar autograd hands to backward() (GradScaler's 65536
 *               under --fp16, 1/accumulation_steps, ...: finetune_text.py:479-482 scales the loss before backward()).
 *               It multiplies grad_scale BEFORE the single rounding to grad_dtype, as the reference's fp32 backward does,
 *               so 16-bit gradients neither underflow nor round twice.  With skip_if_one != 0 the launch is a device-side
 *               no-op when *upstream_dev == 1 (the gradients a previous forward wrote are already exact) -- no host sync.
 *               With upstream_dev, loss_out may be NULL (mean / sum reductions: gradients only).
 *   sim, probs may be NULL. */
int ia_pair_score_loss_fwd_bwd(int measure, int loss, float margin, int reduction, int dtype,
                               int grad_dtype, const void* x, const void* y, int64_t ldx, int64_t ldy,
                               const int64_t* labels, int64_t n, int64_t d, float* sim, float* probs,
                               float* loss_out, void* dx, void* dy, int64_t lddx, int64_t lddy,
                               float grad_scale, const float* upstream_dev, int skip_if_one, void* workspace,
                               size_t workspace_bytes, ia_stream_t stream);

/* The same launch with the backward of the head's activation folded into the gradient store (training path of the head
 * projection, base.py:67-75): x, y are the head's OUTPUTS drop(tanh(dense(.))) and, with act_bwd_scale = 1/(1-p) (1.0 = no
 * dropout; 0 = plain gradients, identical to ia_pair_score_loss_fwd_bwd), dx / dy receive
 *     d_pre = dL/d(out) * keep/(1-p) * (1 - tanh^2)      -- the gradient w.r.t. the dense layer's output,
 * ready for ia_project_dgrad / ia_project_wgrad: no separate pass over the [n, h] gradients.  With dropout active a dropped
 * element is recognised by its exact zero. */
int ia_pair_score_loss_act_bwd(int measure, int loss, float margin, int reduction, int dtype,
                               int grad_dtype, const void* x, const void* y, int64_t ldx, int64_t ldy,
                               const int64_t* labels, int64_t n, int64_t d, float* sim, float* probs,
                               float* loss_out, void* dx, void* dy, int64_t lddx, int64_t lddy,
                               float grad_scale, const float* upstream_dev, int skip_if_one, float act_bwd_scale,
                               void* workspace, size_t workspace_bytes, ia_stream_t stream);

/* ---- backward of the score alone (upstream gradient is an arbitrary per-pair vector) ----------
 * Replaces autograd of InnerProduct / CosineSimilarity / PairwiseDistance when the caller keeps the
 * reference's unfused head -> loss module sequence.  gsim: [n] fp32 = dL/dsim. */
int ia_pair_score_bwd(int measure, int dtype, int grad_dtype, const void* x, const void* y,
                      int64_t ldx, int64_t ldy, const float* gsim, int64_t n, int64_t d, void* dx,
                      void* dy, int64_t lddx, int64_t lddy, ia_stream_t stream);

/* ---- gather-and-score: pair r = (row xi[r] of ex, row yi[r] of ey) --------------------------------
 * Replaces the per-pair Python loop of GCNTwoTower.forward (src/models/graph.py:87-117: one head call per pair,
 * torch.cat per iteration) and scores id pairs (item_train_pair.jsonl) straight against an embedding matrix such
 * as pred_text.py:158-192's feature_matrix.  ex and ey may be the same matrix.  Gradients come back dense per pair
 * ([n, d]); the caller scatters them into the embedding matrix gradient (index_add).
 * rows_x / rows_y are the row counts of ex / ey: an index outside [0, rows) never reads out of bounds -- that pair's
 * sim / probs / loss contribution / gradients become NaN (torch indexing in the reference loop would raise). */
int ia_pair_score_gather_fwd(int measure, int dtype, const void* ex, const void* ey, int64_t rows_x,
                             int64_t rows_y, int64_t ldx, int64_t ldy, const int64_t* xi, const int64_t* yi,
                             int64_t n, int64_t d, float* sim, float* probs, double threshold,
                             uint8_t* labels_out, ia_stream_t stream);
int ia_pair_score_gather_loss_fwd_bwd(int measure, int loss, float margin, int reduction, int dtype,
                                      int grad_dtype, const void* ex, const void* ey, int64_t rows_x,
                                      int64_t rows_y, int64_t ldx, int64_t ldy,
                                      const int64_t* xi, const int64_t* yi, const int64_t* labels, int64_t n,
                                      int64_t d, float* sim, float* probs, float* loss_out, void* dx, void* dy,
                                      int64_t lddx, int64_t lddy, float grad_scale, void* workspace,
                                      size_t workspace_bytes, ia_stream_t stream);

/* ---- threshold sweep: confusion counts of `probs >= thresholds[k]` vs labels, k < nthr <= 32 ----------
 * Replaces the numpy/sklearn loop of finetune_text.py:576-580 (np.arange(0.1, 1.0, 0.1); fp32 probs compared in
 * double).  counts: [nthr][4] u64 = tp, fp, fn, tn; precision/recall/F1 follow on the host from the integers. */
int ia_threshold_sweep(const float* probs, const int64_t* labels, int64_t n, const double* thresholds,
                       int nthr, uint64_t* counts, ia_stream_t stream);

/* ---- best-F1 threshold search: reference finetune_bert.py:72-106 --------------------------------------
 * Replaces the Python `sorted(zip(scores, labels), reverse=...)` + loop: a stable LSD radix sort of (score key, label) on
 * the device (equal scores keep their input order, like Python's sort), a prefix count of the positives (label == 1), the F1
 * of every cut point i in [0, n-2] in float64 with the loop's own operations, and the EARLIEST maximum (the loop's strict
 * '>').  scores: fp32 (IA_F32) or fp64 (IA_F64); labels int64 {0,1}; n < 2^31.
 * out5 (DEVICE, 5 doubles) = best accuracy, F1, precision, recall, threshold = mean of the scores on both sides of the cut;
 * all zero when n < 2 or no cut has F1 > 0, as the reference returns.  workspace: ia_best_f1_workspace_bytes(n, score_dtype)
 * bytes of device scratch, 256-byte aligned, no initialisation. */
size_t ia_best_f1_workspace_bytes(int64_t n, int score_dtype);
int ia_best_f1_threshold(int score_dtype, const void* scores, const int64_t* labels, int64_t n,
                         int high_score_more_similar, double* out5, void* workspace, size_t workspace_bytes,
                         ia_stream_t stream);

/* ---- elementwise losses on a score vector (the reference's loss modules on their own) ---------
 * HingeLoss.forward (loss.py:126-134), EuclideanDistanceLoss.forward (loss.py:61-68), BCEWithLogits
 * (text.py:1403).  target_pm1: int64 in {-1,+1} for hinge / euclidean, {0,1} for bce.
 * loss_out: 1 float or n floats; gsim (may be NULL): d loss / d sim * grad_scale. */
int ia_score_loss_fwd_bwd(int loss, float margin, int reduction, const float* sim,
                          const int64_t* target, int64_t n, float* loss_out, float* gsim,
                          float grad_scale, void* workspace, size_t workspace_bytes,
                          ia_stream_t stream);

/* ---- "softmax" measure: TwoTowerClassificationHead (base.py:103-117) + CrossEntropyLoss --------
 * logits = [x ; y] . W^T + b with W [2, 2h] fp32 row-major, probs = softmax(logits) (numpy twin:
 * submit/similarity.py:19-24).  labels NULL -> forward only.  With labels: loss (mean CE,
 * text.py:1408-1409,1473) and, when non-NULL, dx, dy (grad_dtype), dW [2,2h], db [2] (fp32).
 * upstream_dev / skip_if_one: as in ia_pair_score_loss_fwd_bwd (device-resident upstream scalar folded into every
 * gradient before rounding; optional device-side no-op when it is 1; loss_out may then be NULL).
 * Labels must be 0 or 1 (any non-zero value counts as class 1; nn.CrossEntropyLoss's ignore_index is not offered).
 * workspace: ia_softmax_head_workspace_bytes(h). */
size_t ia_softmax_head_workspace_bytes(int64_t h);
int ia_softmax_head_fwd_bwd(int dtype, int grad_dtype, const void* x, const void* y, int64_t ldx,
                            int64_t ldy, const float* w, const float* b, const int64_t* labels,
                            int64_t n, int64_t h, float* logits, float* probs, float* loss_out,
                            void* dx, void* dy, int64_t lddx, int64_t lddy, float* dw, float* db,
                            float grad_scale, const float* upstream_dev, int skip_if_one, void* workspace,
                            size_t workspace_bytes, ia_stream_t stream);

/* Diagnostics: cycle counters of the last tensor-core softmax-head launch made with IA_HEAD_DEBUG=16 in the environment (summed
 * over CTAs and warps): [0] forward warps waiting for rows, [1] at their barrier, [2] loader warps waiting for a released stage,
 * [3] forward warps total, [4] backward warps waiting for deltas, [5] backward warps total, [6] warp 0's softmax section. */
int ia_softmax_head_last_stats(uint64_t* out8);

/* ---- in-place scale of gradients by a device scalar (autograd upstream != 1, e.g. GradScaler) --
 * No-op on the device when *g == 1.0f. */
int ia_scale_inplace(int dtype, void* a, void* b, int64_t count, const float* g, ia_stream_t stream);

/* ---- head projection in front of the score ----------------------------------
 * Replaces, with dropout inactive (eval mode or p = 0): `x = tanh(dense(f1))`, `y = tanh(dense(f2))` of
 * VecSimClassificationHead.forward (src/models/base.py:67-75; dense = nn.Linear(H*len(cls_layers) or H, H),
 * base.py:47-49) -- one tcgen05 GEMM for both sides with bias + tanh + rounding in its epilogue.
 * f1, f2: [n, k_in]; w: [h, k_in] (nn.Linear layout); bias: [h] fp32 or NULL; x, y: [n, h] in `dtype`.
 * fast_tanh = 0: tanh to ~1e-6 relative in fp32, rounded once (the default everywhere in the Python mirror);
 * fast_tanh = 1: the hardware tanh.approx (2^-11 relative: still within one bf16 ulp of the exact value) -- opt-in.
 * dtype is IA_BF16 or IA_F16 (fp32 -> IA_ERR_UNSUPPORTED: keep torch's GEMM); k_in, h and all leading
 * dimensions are multiples of 8 elements, pointers 16-byte aligned. */
int ia_project_tanh_fwd(int dtype, const void* f1, const void* f2, int64_t ldf1, int64_t ldf2, int64_t n,
                        int64_t k_in, const void* w, int64_t ldw, const float* bias, int64_t h, void* x,
                        void* y, int64_t ldx, int64_t ldy, int fast_tanh, ia_stream_t stream);
/* The same GEMM with the pair score of base.py:77-86 computed in its epilogue from the rounded
 * embeddings (results equal ia_project_tanh_fwd followed by ia_pair_score_fwd up to fp32 summation
 * order); x and y may be NULL, in which case the embeddings never reach HBM (inference scoring).
 * workspace: ia_project_score_workspace_bytes(n, k_in, h) bytes of device scratch, no initialisation. */
size_t ia_project_score_workspace_bytes(int64_t n, int64_t k_in, int64_t h);
/* Diagnostics: cycle counters of the last projection launch made with IA_PROJ_DEBUG=2 in the environment
 * (summed over CTAs): [0] MMA thread waiting for the epilogue, [1] for TMA, [2] MMA thread total,
 * [3] epilogue warps waiting for the MMAs, [4] epilogue warps total. */
int ia_project_last_stats(uint64_t* out8);
int ia_project_score_fwd(int measure, int dtype, const void* f1, const void* f2, int64_t ldf1, int64_t ldf2,
                         int64_t n, int64_t k_in, const void* w, int64_t ldw, const float* bias, int64_t h,
                         void* x, void* y, int64_t ldx, int64_t ldy, float* sim, float* probs,
                         double threshold, uint8_t* labels_out, int fast_tanh, void* workspace,
                         size_t workspace_bytes, ia_stream_t stream);

/* ---- training side of the head projection: x = drop(tanh(dense(drop(f)))) and its backward ---------------------------
 * Replaces VecSimClassificationHead.forward in train() mode (src/models/base.py:50-53,67-75: nn.Dropout on the features and on
 * the tanh outputs) and the autograd backward of dense / tanh / dropout, for bf16 / fp16 tensors.  Dropout masks are
 * counter-based (Philox4x32-10): element (row, col) of a [rows, cols] tensor is kept iff the 16-bit lane (row & 7) of
 * philox(key = seed, counter = ((row >> 3) * cols + col, stream, step)) is >= round(p * 2^16); kept values are scaled by
 * 1/(1-p).  streams: 0 = f1, 1 = f2, 2 = x, 3 = y; `step` is the caller's call counter.  (torch's own RNG stream cannot be
 * reproduced inside a GEMM epilogue; the masks are replayed on the host by the tests.)
 *   ia_dropout_fwd                out = x * mask / (1 - p)                (input dropout of one side; cols % 8 == 0)
 *   ia_project_tanh_dropout_fwd   one tcgen05 GEMM for both sides, bias + tanh + OUTPUT dropout in its epilogue; f1, f2 are the
 *                                 already dropped features.  p_drop = 0 is ia_project_tanh_fwd.
 *   ia_tanh_dropout_bwd           d_pre = g * [out != 0 or no dropout] * s * (1 - (out / s)^2), s = keep_scale = 1/(1-p)
 *                                 (0 or 1: no dropout) -- for an arbitrary upstream gradient g; the fused pair-loss launch
 *                                 ia_pair_score_loss_act_bwd writes d_pre itself
 *   ia_transpose16                dst[c][r] = src[r][c] for 16-bit elements (W -> W^T once per step for the data gradient)
 *   ia_project_dgrad              df1, df2 = (d_pre . W) * input mask / (1 - p): the forward GEMM kernel on (d_pre, W^T [k_in, h])
 *                                 with the input-dropout mask (streams 0, 1) in its epilogue
 *   ia_project_wgrad              dW [h, k_in] fp32 = d_pre1^T f1' + d_pre2^T f2' (f' = the dropped features), db [h] fp32 =
 *                                 column sums of d_pre1, d_pre2 (NULL: skipped): a tcgen05 GEMM contracting over the ROWS
 *                                 (MN-major operands), split-K with a fixed-order reduce.  workspace:
 *                                 ia_project_wgrad_workspace_bytes(n, h, k_in) bytes, 256-byte aligned. */
int ia_dropout_fwd(int dtype, const void* x, int64_t ldx, int64_t rows, int64_t cols, float p_drop, uint64_t seed,
                   uint32_t step, uint32_t stream_id, void* out, int64_t ldo, ia_stream_t stream);
int ia_project_tanh_dropout_fwd(int dtype, const void* f1, const void* f2, int64_t ldf1, int64_t ldf2, int64_t n,
                                int64_t k_in, const void* w, int64_t ldw, const float* bias, int64_t h, void* x,
                                void* y, int64_t ldx, int64_t ldy, float p_drop, uint64_t seed, uint32_t step,
                                ia_stream_t stream);
int ia_tanh_dropout_bwd(int dtype, const void* g, int64_t ldg, const void* out, int64_t ldo, int64_t rows,
                        int64_t cols, float keep_scale, void* dpre, int64_t ldd, ia_stream_t stream);
int ia_transpose16(const void* src, int64_t rows, int64_t cols, int64_t lds, void* dst, int64_t ldd,
                   ia_stream_t stream);
int ia_project_dgrad(int dtype, const void* d1, const void* d2, int64_t ldd1, int64_t ldd2, int64_t n, int64_t h,
                     const void* wt, int64_t ldwt, int64_t k_in, void* df1, void* df2, int64_t lddf1,
                     int64_t lddf2, float p_drop, uint64_t seed, uint32_t step, ia_stream_t stream);
size_t ia_project_wgrad_workspace_bytes(int64_t n, int64_t h, int64_t k_in);
int ia_project_wgrad(int dtype, const void* d1, const void* d2, int64_t ldd, const void* f1, const void* f2,
                     int64_t ldf, int64_t n, int64_t h, int64_t k_in, float* dw, float* db, void* workspace,
                     size_t workspace_bytes, ia_stream_t stream);

/* ---- row inverse norms: 1 / max(||row||, eps) (cosine retrieval pre-pass, base.py:58 eps) ------ */
int ia_row_inv_norm(int dtype, const void* x, int64_t n, int64_t d, int64_t ldx, float eps,
                    float* out, ia_stream_t stream);

/* ---- catalog retrieval: all-pairs score + per-query top-k --------------------------------------
 * No reference implementation exists (README.md:12,16 motivate it); semantics = the pairwise
 * similarity above for every (query, catalog row), top-k by a stable sort (ties -> lower index;
 * nearest precedent torchkge/torchkge/inference.py:243-246).  Results are 64-bit keys
 *   key = (orderable(score) << 32) | (0xFFFFFFFF - global_row)      (distances: score word complemented)
 * so one unsigned max-compare means "better score, then lower index" in the kernel epilogue, the
 * shard merge and after the NCCL all-gather alike.
 *
 * A catalog handle borrows the device catalog [c, d] (bf16 for the tensor-core measures; fp32/bf16
 * for l1/l2) and owns inverse norms, TMA descriptors and scratch. */
typedef struct ia_catalog ia_catalog;
int ia_catalog_create(ia_catalog** out, int dtype, const void* catalog, int64_t c, int64_t d,
                      int64_t ld, int64_t row_base, ia_stream_t stream);
void ia_catalog_destroy(ia_catalog* cat);
/* keys_out: [q, k] u64 sorted best-first.  k <= IA_MAX_K. */
#define IA_MAX_K 128
int ia_catalog_topk(ia_catalog* cat, int measure, const void* queries, int64_t q, int64_t ldq,
                    int k, uint64_t* keys_out, ia_stream_t stream);
/* Same, with per-query lower bounds known to the caller: tau_init[q] (int64 holding the upper 32 bits of a key,
 * 0 = none) promises that the final k-th best key of query q -- over ALL shards the caller will merge -- is
 * >= tau_init[q] << 32, so candidates below it are never collected.  This is how row shards share what they
 * learnt in a cheap probe pass (ShardedCatalogIndex: min over ranks of each rank's ceil(k/G)-th best). */
int ia_catalog_topk_seeded(ia_catalog* cat, int measure, const void* queries, int64_t q, int64_t ldq,
                           int k, const int64_t* tau_init, uint64_t* keys_out, ia_stream_t stream);
/* Probe pass of a row shard (ShardedCatalogIndex): `groups` disjoint row groups of rows_per_group rows (a multiple of 256) at the
 * head of the catalog are scanned with a top-kp each (kp <= 16: the kernel's register top-k path, no threshold sharing between
 * groups); bound_out[q] (int64, DEVICE) = the smallest of the groups' kp-th best key words, 0 = no bound.  The ranks take the
 * MIN over all ranks (one small all-reduce): (ranks * groups * kp) >= k rows meet that key, so it is a valid tau_init for
 * ia_catalog_topk_seeded on every shard.  ia_catalog_topk runs the same probe on its own for k > 16 (IA_RETR_PROBE=0: off). */
int ia_catalog_probe_bound(ia_catalog* cat, int measure, const void* queries, int64_t q, int64_t ldq, int kp,
                           int groups, int64_t rows_per_group, int64_t* bound_out, ia_stream_t stream);
/* Telemetry of the last ia_catalog_topk on this handle (synchronises the device): out8[0] keys appended to the
 * per-query buffers, [1] buffer->list merges, [2] 32-column groups that left the fast path, [3] rare-path
 * iterations, and epilogue-warp cycle sums: [4] waiting for accumulators, [5] in merges, [6] total, [7] rare path. */
int ia_catalog_last_stats(ia_catalog* cat, uint64_t* out8);
/* The work decomposition ia_catalog_topk chose last: catalog splits and 256-row tiles per split. */
int ia_catalog_last_plan(ia_catalog* cat, int* splits, int* tiles_per_split);
/* merge `parts` sorted key lists [parts, q, k] -> [q, k] (after the all-gather of shard results) */
int ia_topk_merge(const uint64_t* keys_in, int parts, int64_t q, int k, uint64_t* keys_out,
                  ia_stream_t stream);
/* keys -> fp32 scores + int64 global rows */
int ia_unpack_keys(const uint64_t* keys, int64_t count, int descending, float* scores,
                   int64_t* rows, ia_stream_t stream);

/* ---- host-buffer entry points (what a CPU-side caller of the reference binds) -------------------
 * Same maths with HOST inputs/outputs: pinned or pageable host pointers, chunked H2D overlapped
 * with the kernels on internal streams, result copied back; returns when the outputs are valid.
 * Replaces the vectorised CPU scoring of finetune_text.py:566-580 and per-pair compute() loops
 * (submit/similarity.py:27, pred_bert.py:47-52) for a whole batch. */
int ia_pair_score_host(int measure, int dtype, const void* x, const void* y, int64_t n, int64_t d,
                       float* sim, float* probs, double threshold, uint8_t* labels_out, int device);
/* Page-locked host buffers for these entry points (full PCIe speed, no driver staging).  write_combined = 1 for buffers the CPU
 * only writes and the device reads (x, y, labels): DMA reads skip the CPU-cache snoop; never read such a buffer on the CPU. */
int ia_host_alloc(void** ptr, size_t bytes, int write_combined);
int ia_host_free(void* ptr);
int ia_pair_score_loss_host(int measure, int loss, float margin, int reduction, int dtype,
                            const void* x, const void* y, const int64_t* labels, int64_t n, int64_t d,
                            float* loss_out, void* dx, void* dy, int device);

/* ---- dissimilarity ranking with explicit eps / squaring -----------------------
 * The l1 / l2 all-pairs kernel with the two knobs that separate nn.PairwiseDistance (eps = 1e-6 added to
 * every difference, base.py:59-62) from torchkge's plain norms (torchkge/torchkge/utils/dissimilarities.py:11-25:
 * l1 = ||a-b||_1, l2 = ||a-b||_2^2; eps = 0, squared = 1), so that candidate ranking for link prediction
 * (torchkge/torchkge/inference.py:216-246: scores = -dissimilarity(h + r, candidates), sort descending,
 * top_k) runs through the catalog: keys hold the k SMALLEST distances, ties -> lower row.
 * p_norm is 1 or 2; `squared` only matters for p_norm = 2. */
int ia_catalog_topk_dissimilarity(ia_catalog* cat, int p_norm, float eps, int squared, const void* queries,
                                  int64_t q, int64_t ldq, int k, uint64_t* keys_out, ia_stream_t stream);

/* ---- binary catalog files ---------------------------------------------------
 * Replaces the embedding JSONL of finetune_text.py:784-792 (one pair per line, embeddings as
 * stringified float lists, read back with eval at model_ensemble.py:112) as the input of retrieval:
 * a 64-byte header, the row-major [rows, dim] matrix at offset 4096 and an optional id table
 * (uint64 offsets[rows+1] + UTF-8 bytes).  Host-side functions; only _upload touches the device. */
typedef struct ia_catalog_file ia_catalog_file;
/* data: rows x dim elements already in `dtype`; ids: rows C strings or NULL. */
int ia_catalog_file_write(const char* path, int dtype, const void* data, int64_t rows, int64_t dim,
                          const char* const* ids);
int ia_catalog_file_open(const char* path, ia_catalog_file** out);        /* mmap, validates the layout */
void ia_catalog_file_close(ia_catalog_file* f);
int ia_catalog_file_info(const ia_catalog_file* f, int* dtype, int64_t* rows, int64_t* dim, int* has_ids);
const void* ia_catalog_file_data(const ia_catalog_file* f);               /* the mapped matrix */
const char* ia_catalog_file_id(const ia_catalog_file* f, int64_t row, int64_t* len);   /* not NUL-terminated */
/* rows [row_begin, row_end) -> device_dst (row-major, ld = dim): a rank uploads its shard_bounds() slice.
 * Double-buffered pinned staging; synchronises `stream` before returning. */
int ia_catalog_file_upload(const ia_catalog_file* f, int64_t row_begin, int64_t row_end, void* device_dst,
                           ia_stream_t stream);
/* Native converter: reference JSONL -> catalog file.  side 0 = src_item_*, 1 = tgt_item_*, 2 = both;
 * an item id is stored once (first occurrence).  Values are float32(float64(decimal)) -- what eval + numpy give
 * a consumer of the reference's file, bit for bit -- then rounded to `dtype` to nearest even.  `threads` workers
 * (0 = all hardware threads) parse runs of whole lines; the output does not depend on the thread count. */
int ia_embedding_jsonl_to_catalog(const char* jsonl_path, const char* out_path, int dtype, int side,
                                  int threads, int64_t* rows_out, int64_t* dim_out);

#ifdef __cplusplus
}
#endif
#endif /* IA_B200_H */
