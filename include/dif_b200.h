/* dif_b200.h - C ABI of libdif_b200.so: the embedding-space distance path of
 * sandyz1000/deep-insight-face, rebuilt for NVIDIA B200 (sm_100a).
 *
 * The reference has no FFI for this path: its boundary is plain Python (a Keras Loss subclass,
 * numpy evaluators, verify()).  Each entry point below names the reference function whose
 * arithmetic it replaces (paths relative to the reference repo root).  The Python mirror in
 * deep_insight_face_b200/ binds these with ctypes; INTEGRATION.md shows the stub a maintainer of
 * the reference would add.
 *
 * Conventions
 *   - every function returns 0 (DIF_OK) or a negative DIF_ERR_* code and never throws;
 *     dif_last_error() returns a thread-local description of the last failure;
 *   - pointers are caller-owned DEVICE pointers unless the function name ends in _host;
 *   - functions are stream-ordered on `stream` (a cudaStream_t, NULL = default stream) and
 *     return without synchronising, except the *_host variants, which synchronise before returning;
 *   - a handle must not be used from two threads at once;
 *   - there is no CPU fallback: without a usable sm_100 device every compute call fails.
 */
#ifndef DIF_B200_H_
#define DIF_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DIF_OK 0
#define DIF_ERR_INVALID (-1)  /* bad argument */
#define DIF_ERR_CUDA (-2)     /* CUDA runtime / driver error */
#define DIF_ERR_NO_DEVICE (-3)
#define DIF_ERR_CAPACITY (-4) /* gallery full / workspace too small */
#define DIF_ERR_STATE (-5)

/* metric codes follow deep_insight_face/evaluation/utility.py:52-66 (`distance_metric`) */
#define DIF_METRIC_SQL2 0   /* sum((a-b)^2)                      utility.py:55-56 */
#define DIF_METRIC_COSINE 1 /* arccos(cos_sim)/pi for pairs; cosine similarity for search */

/* Precision of the tensor-core pass.  Gallery search re-ranks every surviving candidate with the
 * canonical fp32 reduction and proves the candidate lists complete from the pass's error bound, so
 * its RESULTS are identical in all three modes; the losses report values of the chosen precision. */
#define DIF_PREC_TF32X3 0 /* fp32-exact: hi/lo TF32 planes, 3 tensor-core products per term (default) */
#define DIF_PREC_BF16 1   /* bf16 operands, fp32 accumulate (reported separately) */
#define DIF_PREC_TF32X1 2 /* fp32 storage read as TF32 by the tensor cores, one product per term */
#define DIF_PREC_BF16X3 3 /* x = b0 + b1 in bf16; b1*b0 + b0*b1 + b0*b0, fp32 accumulate: error <= 1.1e-5 of sum |a||b|,
                           * the 3xTF32 scheme on the bf16 pipe at half its cost (gallery filter only) */

#define DIF_MAX_TOPK 24

/* ---- library ------------------------------------------------------------------------------ */
/* Selects the device, checks it is sm_100, warms the context.  One process drives ONE GPU (the multi-GPU model
 * is one process per GPU over torch.distributed): the library's workspaces and kernel attributes belong to the
 * device of the first successful call; a later call naming another device fails with DIF_ERR_STATE. */
int dif_init(int device);
const char* dif_last_error(void);
int dif_sync(void* stream);
const char* dif_version(void);
int64_t dif_launch_count(void);       /* kernels launched by this library since load (bench evidence) */
/* The loss entry points keep per-thread workspaces sized for the largest batch seen; a block that is outgrown is parked
 * rather than freed, because a CUDA graph captured at the smaller size may still replay kernels that point at it.  A
 * sweep over batch sizes therefore holds every size's blocks until this call frees the parked ones: call it when no
 * captured graph that contains library kernels is alive.  Returns the number of blocks freed. */
int64_t dif_release_retired(void);

/* ---- 1:N gallery search -------------------------------------------------------------------
 * New API (the reference only has 1:1 verify: predictions.py:104-150 `TripletPrediction.verify`,
 * api.py:94-104 `face_distance`).  Rows are L2-normalised on add for DIF_METRIC_COSINE
 * (networks/inceptionv3.py:305, networks/triplet.py:138 contract) and stored raw for DIF_METRIC_SQL2.
 * Search returns, per query, the k best rows ordered by (score best-first, row index ascending);
 * scores are cosine similarity (descending) or squared L2 distance (ascending), computed with the
 * canonical fp32 reduction of oracle/dif_oracle.c so results are bit-reproducible on the CPU. */
typedef struct dif_gallery dif_gallery_t;

dif_gallery_t* dif_gallery_create(int device, int64_t capacity_rows, int dim, int metric, int precision);
void dif_gallery_destroy(dif_gallery_t* g);
/* ids NULL: id = id_base + row index.  The first add WITH ids makes the default ids of the rows enrolled so far
 * explicit; from then on every add must carry ids (DIF_ERR_STATE otherwise). */
int dif_gallery_add(dif_gallery_t* g, const float* rows, const int64_t* ids, int64_t n, void* stream);
int dif_gallery_add_host(dif_gallery_t* g, const float* rows_host, const int64_t* ids_host, int64_t n);
/* rows [row0, row0+n) of the synthetic gallery of oracle/dif_oracle.c:dif_synth_value (seed, dim) */
int dif_gallery_fill_synth(dif_gallery_t* g, uint64_t seed, int64_t row0, int64_t n, void* stream);
int dif_gallery_set_id_base(dif_gallery_t* g, int64_t id_base); /* default ids = id_base + local row */
/* Incremental delete (the persistent identity index replacing the python dict of
 * deep_insight_face/predictions.py:112).  rows_host: n strictly ascending row numbers; the remaining rows keep
 * their order (so "ties -> lower row" still means "earlier enrolled") and their ids - default ids (id_base + row)
 * are frozen first; rows added later without ids continue the default sequence at id_base + (rows ever enrolled), so
 * an id is never reused.  Synchronises the stream. */
int dif_gallery_remove(dif_gallery_t* g, const int64_t* rows_host, int64_t n, void* stream);
/* ids of rows [row0, row0+n) into a DEVICE buffer */
int dif_gallery_get_ids(dif_gallery_t* g, int64_t row0, int64_t n, int64_t* out, void* stream);
int64_t dif_gallery_size(const dif_gallery_t* g);
/* tuning / test knobs: "gemm_ctas" = 1 | 2 (CTA pair, default); "force_fallback" = 0 | 1 (send every
 * query through the exact brute-force path as well); "resident_queries" = -1 auto | 0 | 1 (keep the
 * query block resident in shared memory while gallery tiles stream); "splits" = 0 auto | n;
 * "l2_prefetch" = 0 | 1 (producer prefetches gallery tiles into L2 two tiles ahead; off by default) */
int dif_gallery_set_option(dif_gallery_t* g, const char* name, int value);
int dif_gallery_reset(dif_gallery_t* g);
/* scores [Q*k] fp32, ids [Q*k] int64, rows [Q*k] int32 local row index (may be NULL).
 * Slots beyond the gallery size hold id -1 / row -1. */
int dif_gallery_search(dif_gallery_t* g, const float* queries, int n_queries, int k, float* scores,
                       int64_t* ids, int32_t* rows, void* stream);
/* same call with HOST buffers: H2D of queries (straight from the caller's buffer when it is page-locked, through
 * the handle's pinned staging otherwise), D2H of results, synchronises */
int dif_gallery_search_host(dif_gallery_t* g, const float* queries_host, int n_queries, int k,
                            float* scores_host, int64_t* ids_host, int32_t* rows_host);
/* counters of the last search: [0] queries that took the exact fallback, [1] kernels launched,
 * [2] candidate splits, [3] candidates per split, [4] 1 if the resident-query schedule ran,
 * [5] reserved (0) */
int dif_gallery_last_stats(const dif_gallery_t* g, int64_t out[6]);
/* duration in ms of the last search's tensor-core kernel (CUDA events on the launch stream);
 * only valid after the stream has been synchronised */
int dif_gallery_last_kernel_ms(dif_gallery_t* g, float* ms);
/* durations in ms of the last search's phases (CUDA events on the launch stream, valid after the stream has been
 * synchronised): [0] query prep, [1] tensor-core filter, [2] window + canonical re-rank, [3] exact path for flagged
 * queries, [4] candidate exchange + merge (sharded searches, else 0) */
int dif_gallery_last_phase_ms(dif_gallery_t* g, float out[5]);
/* copy canonical stored rows (normalised for cosine) back out: out [n*dim] fp32 */
int dif_gallery_get_rows(dif_gallery_t* g, int64_t row0, int64_t n, float* out, void* stream);

/* merge `world` per-shard results (each [Q*k], shard-major) into the global top-k, ordering
 * (score best-first, global row ascending).  grows = global row index of each candidate. */
int dif_topk_merge(const float* scores, const int64_t* grows, const int64_t* ids, int world, int n_queries,
                   int k, int metric, float* out_scores, int64_t* out_grows, int64_t* out_ids, void* stream);

/* ---- row-sharded search over several GPUs (one process per GPU) ----------------------
 * Rank r's gallery holds the contiguous global rows [shard_row0, shard_row0 + size).  One search =
 *   local search (the four passes above) -> every rank's k candidates per query are exchanged -> every rank
 *   merges world*k candidates per query with the key (score best-first, GLOBAL row ascending),
 * so the result equals the single-gallery search of the concatenated rows bit for bit.  All ranks must call
 * with the same n_queries / k / queries and galleries of the same metric that either all carry explicit ids
 * or none does (default id = id_base + local row).
 *
 * The exchanged unit is a packed chunk per rank: scores f32 [Q*k] | local rows i32 [Q*k] | ids i64 [Q*k] (explicit
 * ids only), 8 B per candidate without ids (5.24 MB per rank at 65 536 queries x top-10).
 *
 * transport (chosen at attach time): DIF_TRANSPORT_NCCL = one ncclAllGather of the chunks, then the merge
 * kernel; DIF_TRANSPORT_PEER = ONE kernel that stores this rank's chunk straight into every peer's exchange buffer
 * over NVLink (buffers mapped with CUDA IPC at attach time), publishes a per-peer epoch flag, waits for the
 * peers' flags and merges - the candidate exchange fused with its consumer, no collective launch. */
#define DIF_TRANSPORT_NCCL 0
#define DIF_TRANSPORT_PEER 1
#define DIF_NCCL_ID_BYTES 128
/* NCCL is bound at run time (dlopen of the libnccl.so.2 already loaded in the process, else the system one), so
 * the library still loads on a box without NCCL; these fail with DIF_ERR_STATE there. */
int dif_nccl_unique_id(void* id_out_host /* DIF_NCCL_ID_BYTES */);
/* ncclCommInitRank on the device of dif_init; *comm_out is an ncclComm_t owned by the caller (dif_nccl_comm_destroy) */
int dif_nccl_comm_create(int world, int rank, const void* id_host, void** comm_out);
int dif_nccl_comm_destroy(void* nccl_comm);
/* bytes of one packed chunk */
int64_t dif_shard_chunk_bytes(int n_queries, int k, int with_ids);
/* local search writing the packed chunk (DEVICE pointer, dif_shard_chunk_bytes(...) bytes, 16-byte aligned) */
int dif_gallery_search_packed(dif_gallery_t* g, const float* queries, int n_queries, int k, int with_ids, void* chunk,
                              void* stream);
/* merge of `world` packed chunks laid out back to back; shard_info DEVICE [world][2] int64 = {shard_row0, id_base}.
 * scores [Q*k] f32, ids [Q*k] i64, grows [Q*k] i64 global rows (may be NULL); empty slots id -1 / row -1. */
int dif_shard_merge(const void* chunks, const int64_t* shard_info, int world, int n_queries, int k, int metric,
                    int with_ids, float* scores, int64_t* ids, int64_t* grows, void* stream);
/* Binds the gallery to its place in the job: exchanges {shard_row0, id_base, ids?} of every rank once (and, for
 * DIF_TRANSPORT_PEER, the CUDA IPC handles of the exchange buffers sized for max_queries x max_k), synchronises.
 * nccl_comm: an ncclComm_t of `world` ranks (torch's ProcessGroupNCCL._comm_ptr(), or dif_nccl_comm_create). */
int dif_gallery_shard_attach(dif_gallery_t* g, void* nccl_comm, int rank, int world, int64_t shard_row0,
                             int max_queries, int max_k, int transport);
/* The sharded search itself, stream-ordered, no host synchronisation.  queries: DEVICE [Q*D], identical on all ranks. */
int dif_gallery_search_sharded(dif_gallery_t* g, void* nccl_comm, int rank, int world, const float* queries,
                               int n_queries, int k, float* scores, int64_t* ids, int64_t* grows, void* stream);
/* HOST buffers.  How the (identical) query batch reaches every GPU:
 *   bcast_root = DIF_UPLOAD_EACH   (-1)  every rank uploads its own full copy over its PCIe link;
 *   bcast_root = DIF_UPLOAD_SLICED (-2)  every rank uploads 1/world of the rows of its copy and one ncclAllGather
 *                                        over NVLink assembles the batch on every GPU (PCIe bytes per rank / world);
 *   bcast_root = r >= 0                  only rank r's queries_host is read, uploaded once and ncclBroadcast (other
 *                                        ranks may pass NULL).
 * Uploads go straight from the caller's buffer when it is page-locked.  scores_host / ids_host / grows_host may be NULL
 * on ranks that do not need the result (they skip the D2H).  Synchronises. */
#define DIF_UPLOAD_EACH (-1)
#define DIF_UPLOAD_SLICED (-2)
int dif_gallery_search_sharded_host(dif_gallery_t* g, void* nccl_comm, int rank, int world, const float* queries_host,
                                    int bcast_root, int n_queries, int k, float* scores_host, int64_t* ids_host,
                                    int64_t* grows_host);

/* raw synthetic rows shared with the oracle (oracle/dif_oracle.c:dif_or_synth_value):
 * out[r*D + d] = synth(seed, row, d, D) with row = rows_idx ? rows_idx[r] : row0 + r (device pointers) */
int dif_synth_fill(float* out, uint64_t seed, int64_t row0, const int64_t* rows_idx, int64_t n, int D, void* stream);

/* diagnostic: C[M,N] (row-major, ldc = N) = A[M,K] * B[N,K]^T through the same tcgen05/TMA skeleton the
 * search and loss kernels use.  precision as above; ctas = 1 | 2 (+16: resident-A schedule).
 * Used by tests/test_gemm_gpu.py. */
int dif_debug_nt_gemm(const float* A, const float* B, int M, int N, int K, float* C, int precision, int ctas,
                      int n_splits, void* stream);

/* diagnostic for the skeleton's operand-major and tile-width options: C [M][N] = A B^T with A given as [M][K]
 * (a_mn = 0) or [K][M] (a_mn = 1: read through an MN-major shared-memory descriptor, no transposed copy) and B as
 * [N][K] or [K][N]; precision DIF_PREC_TF32X3 | DIF_PREC_BF16X3; CTA pairs; tile width bn <= 256, a multiple of 32
 * (of 64 for TF32 / 128 for bf16 planes when b_mn = 1).  Used by tests/test_gemm_gpu.py. */
int dif_debug_gemm_layout(const float* A, const float* B, int M, int N, int K, float* C, int precision, int a_mn, int b_mn,
                          int bn, int n_splits, void* stream);

/* diagnostic: average ms of the NT-GEMM main loop alone (checksum epilogue) on synthetic operands;
 * ctas = 1 | 2 (+16: resident-A schedule; + (bn << 8): tile width bn instead of 256) */
int dif_debug_gemm_time(int M, int N, int K, int precision, int ctas, int n_splits, int iters, float* ms_out);

/* ---- batch-hard triplet losses ------------------------------------------------------------
 * deep_insight_face/common/losses.py:33-51 (BatchHardTripletLoss, cosine),
 * :54-85 (BatchHardTripletLossEuclidean), :88-128 (...AutoAlpha: pass the current auto_alpha as
 * `alpha`; the caller's next value is stats[0] * alpha_scale, losses.py:113).
 * labels are int32 class ids (argmax of the one-hot, losses.py:35; see dif_labels_from_onehot).  Outputs:
 *   loss [B]; pos_idx/neg_idx [B] mined column (first index on ties, -1 if a filler won; may be NULL);
 *   stats [4] = mean(dists), mean(hardest_pos), mean(hardest_neg), max(dists) (losses.py:72-80,70);
 *   demb [B*D] gradient of sum_i dloss[i]*loss[i] (dloss NULL -> 1/B each, Keras AUTO mean); NULL skips backward.
 * Arithmetic is the canonical fp32 reduction of oracle/dif_oracle.c, so mined indices are reproducible
 * bit for bit on the CPU; `precision` must be DIF_PREC_TF32X3 (fp32-exact). */
#define DIF_LOSS_BH_COSINE 0
#define DIF_LOSS_BH_EUCLIDEAN 1
/* OR into `variant`: soft margin log(1 + exp(hn - hp | hp - hn)) of arXiv 1703.07737 eq. 4 instead of the
 * reference's hard margin max(. + alpha, 0); alpha is ignored.  Not in the reference. */
#define DIF_LOSS_SOFT_MARGIN 4
int dif_batch_hard(const float* emb, const int32_t* labels, int B, int D, int variant, float alpha, float* loss,
                   int32_t* pos_idx, int32_t* neg_idx, float* stats, const float* dloss, float* demb,
                   int precision, void* stream);
int dif_batch_hard_host(const float* emb_host, const int32_t* labels_host, int B, int D, int variant, float alpha,
                        float* loss_host, int32_t* pos_idx_host, int32_t* neg_idx_host, float* stats_host,
                        const float* dloss_host, float* demb_host, int precision);
/* The calling thread's page-locked staging block for a B x D step, as eight pointers the caller may work in directly:
 * bufs[0..7] = emb [B*D] float, labels [B] int32, dloss [B] float (inputs); loss [B] float, pos_idx [B] int32, neg_idx [B]
 * int32, stats [4] float, demb [B*D] float (outputs).  dif_batch_hard_host called with these very pointers skips its
 * copy-in / copy-out: at B <= 128 a step is then one kernel launch and one wait, nothing else (the kernel reads and writes
 * the block across PCIe).  The pointers stay valid until the thread asks for a larger block. */
int dif_batch_hard_host_buffers(int B, int D, void** bufs);
/* test knob: 0 = automatic (B <= 128: the whole step in one thread-block-cluster launch; B >= 512: tensor-core filter +
 * canonical re-rank; CUDA-core miner + fused merge / gradient in between), 1 = always the CUDA-core miner (two
 * launches), 2 = always the tensor-core miner, 3 = the cluster step where it fits.  Mined indices, losses and
 * gradients are identical; the printed statistics' float sums are folded in a path-specific fixed order. */
int dif_batch_hard_set_path(int path);
/* deep_insight_face/common/losses.py:131-148 (BatchAllTripletLoss, cosine): loss [B] = pos_loss + neg_loss;
 * demb [B*D] gradient of sum_i dloss[i]*loss[i] (dloss NULL -> 1/B each; demb NULL skips the backward pass). */
int dif_batch_all(const float* emb, const int32_t* labels, int B, int D, float alpha, float* loss, const float* dloss,
                  float* demb, void* stream);
/* tensorflow_addons TripletHardLoss / TripletSemiHardLoss with their defaults, the third-party losses of
 * deep_insight_face/networks/triplet.py:196,209,211 (arithmetic of tensorflow_addons/losses/triplet.py and
 * metric_learning.py, not in the reference tree).  labels [B] sparse int32; kind = DIF_TFA_HARD | DIF_TFA_SEMIHARD,
 * optionally | DIF_TFA_SOFT (hard only: log1p(exp(hp - hn))) | DIF_TFA_SQUARED (distance_metric "squared-L2",
 * default "L2").  loss [1] scalar (mean over anchors | sum over positive pairs / their number, NaN when there
 * is none, as in tfa); pos_idx / neg_idx [B] optional, hard only: mined columns, first index on ties, -1 if the
 * anchor has no positive / negative; demb [B*D] optional = dloss * d loss / d emb.  B <= 8192, D <= 512. */
#define DIF_TFA_HARD 0
#define DIF_TFA_SEMIHARD 1
#define DIF_TFA_SOFT 4
#define DIF_TFA_SQUARED 8
int dif_tfa_triplet(const float* emb, const int32_t* labels, int B, int D, int kind, float margin, float* loss,
                    int32_t* pos_idx, int32_t* neg_idx, float dloss, float* demb, void* stream);
/* tf.argmax(labels, axis=1) of a one-hot [B, C] fp32 matrix (first maximum), losses.py:35 */
int dif_labels_from_onehot(const float* onehot, int B, int C, int32_t* labels, void* stream);

/* Embedding head: tf.nn.l2_normalize(x, axis=1) of deep_insight_face/networks/inceptionv3.py:305
 * and networks/triplet.py:138, y = x * rsqrt(max(sum x^2, 1e-12)) in the canonical fp32 arithmetic; inv_norm [n]
 * optional.  Backward: dx = inv * (g - y (y . g)) (a plain scaling for rows whose squared norm was clamped). */
int dif_l2_normalize(const float* x, int64_t n, int D, float* y, float* inv_norm, void* stream);
int dif_l2_normalize_bwd(const float* g, const float* y, const float* inv_norm, int64_t n, int D, float* dx, void* stream);

/* explicit-triplet loss on [B, 3D] rows (anchor|positive|negative):
 * deep_insight_face/networks/triplet.py:16-46 `triplet_loss`; loss [B]; dy [B*3D] optional (dloss NULL -> 1) */
int dif_triplet_apn(const float* y_pred, int B, int D, float alpha, float* loss, const float* dloss, float* dy,
                    void* stream);
/* deep_insight_face/networks/siamese.py:22-24 `euclidean_distance`: out [B] = sqrt(max(sum((x-y)^2), eps)) */
int dif_euclidean_distance(const float* x, const float* y, int B, int D, float eps, float* out, void* stream);
/* deep_insight_face/networks/siamese.py:32-39 `contrastive_loss` (margin 1): out [1] mean loss; dd [B] optional */
int dif_contrastive_loss(const float* y_true, const float* dist, int B, float margin, float* out, float* dd,
                         void* stream);

/* ---- ArcFace additive-angular-margin logits + softmax cross-entropy ------------------------
 * Absent from the reference; spec in DESIGN.md / oracle/losses_oracle.py:arcface (arXiv 1801.07698):
 * logits = s*cos(theta + m*onehot) over L2-normalised X [B,D] and W [C,D] (easy-margin fallback past
 * pi - m); loss [B] = per-sample cross-entropy; dX [B*D], dW [C*D] = gradients of sum_b dloss[b]*loss[b]
 * (dloss NULL -> 1/B each); pass dX = dW = NULL to skip the backward pass.  precision: DIF_PREC_TF32X3
 * (TF32 hi/lo operand planes, D % 4 == 0, D >= 32) or DIF_PREC_BF16X3 (two bf16 planes per operand, D % 8 == 0,
 * D >= 64; half the plane bytes and tensor time); both are within 1e-4 of the fp64 oracle. */
int dif_arcface(const float* X, const float* W, const int32_t* y, int B, int C, int D, float s, float m,
                float* loss, const float* dloss, float* dX, float* dW, int precision, void* stream);
int dif_arcface_host(const float* X_host, const float* W_host, const int32_t* y_host, int B, int C, int D, float s,
                     float m, float* loss_host, const float* dloss_host, float* dX_host, float* dW_host,
                     int precision);

/* ---- pair verification --------------------------------------------------------------------
 * deep_insight_face/evaluation/utility.py:52-66 `distance`: out [N]; metric 0 squared L2,
 * metric 1 arccos(cosine)/pi; any other metric fails with "Undefined distance metric %d" (utility.py:64).
 * mean [D] (may be NULL) is subtracted from both rows first (utility.py:98-102,144-148 `subtract_mean`). */
int dif_pair_distance(const float* e1, const float* e2, int64_t N, int D, int metric, const float* mean, float* out,
                      void* stream);
int dif_pair_distance_host(const float* e1_host, const float* e2_host, int64_t N, int D, int metric,
                           const float* mean_host, float* out_host);
/* train-fold means of utility.py:98-102: folds are the contiguous KFold(shuffle=False) ranges
 * [fold_begin[f], fold_begin[f+1]); mean [n_folds*D] row f = mean of both embedding sets over the rows
 * NOT in fold f.  workspace: n_folds*D doubles. */
int dif_fold_mean(const float* e1, const float* e2, const int64_t* fold_begin, int n_folds, int D, double* workspace,
                  float* mean, void* stream);
/* deep_insight_face/evaluation/utility.py:36-49 `calculate_accuracy` / :69-77 `calculate_val_far` for T
 * thresholds and every fold in one pass: counts [n_folds*T*4] int64 = tp, fp, tn, fn over the pairs with
 * fold[i] == f (fold NULL: one fold holding every pair), predict = (double)dist < thresholds[t]
 * (strict, np.less against the float64 thresholds of np.arange).  ascending != 0 promises sorted
 * thresholds (histogram + scan path; workspace n_folds*2*(T+1) u32); 0 takes the direct path. */
int dif_threshold_sweep(const float* dist, const uint8_t* issame, const int32_t* fold, int64_t N, int n_folds,
                        const double* thresholds, int T, int ascending, unsigned int* workspace, int64_t* counts,
                        void* stream);
int dif_threshold_sweep_host(const float* dist_host, const uint8_t* issame_host, const int32_t* fold_host, int64_t N,
                             int n_folds, const double* thresholds_host, int T, int64_t* counts_host);

#ifdef __cplusplus
}
#endif
#endif /* DIF_B200_H_ */
