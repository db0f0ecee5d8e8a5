/*
 * ured_chamfer.h -- C ABI of the B200-native Chamfer / density-aware Chamfer (DCD) hot path.
 *
 * This is the drop-in boundary for ONE path of the U-RED reference: the native Chamfer op
 * (pybind module `chamfer_3D`) and the torch-op body of `calc_dcd` / `calc_cd` built on it.
 * Paths below are relative to the reference tree; DCD/ = Density_aware_Chamfer_Distance/.
 *
 *   reference interface                                              replaced by
 *   ---------------------------------------------------------------  ---------------------------
 *   chamfer_3D.forward(xyz1,xyz2,dist1,dist2,idx1,idx2)              ured_chamfer_forward
 *       DCD/utils_v2/metrics/CD/chamfer3D/chamfer_cuda.cpp:17-19,31
 *       -> chamfer_cuda_forward            .../chamfer3D.cu:136-154
 *       -> NmDistanceKernel x2             .../chamfer3D.cu:12-134
 *   chamfer_3D.backward(xyz1,xyz2,gradxyz1,gradxyz2,graddist1,       ured_chamfer_backward
 *                       graddist2,idx1,idx2)
 *       .../chamfer_cuda.cpp:22-26,32 -> chamfer3D.cu:155-195
 *   torch-op body of calc_cd / calc_dcd                              ured_dcd_forward,
 *       DCD/utils_v2/model_utils.py:13-51, 53-70                     ured_dcd_backward
 *   torch.topk(cd_m, k, largest=False) ranking                       ured_topk_smallest, ured_merge_topk,
 *       dataset/dataset_utils.py:1043-1051                           ured_topk_exchange (sharded over GPUs)
 *   fscore(dist1, dist2, threshold)                                  ured_dcd_forward_ex (fscore output)
 *       DCD/utils_v2/metrics/CD/fscore.py:3-16
 *   pytorch3d knn_points(K=1) of residual_retrieval_loss              ured_nn_packed(URED_FLAG_ONE_DIRECTION),
 *       loss/basic_loss.py:249-265                                   ured_nn_backward_one_direction
 *   emd.forward / emd.backward (auction EMD)                         ured_emd_forward, ured_emd_backward
 *       DCD/utils_v2/metrics/EMD/emd.cpp:13-30 -> emd_cuda.cu:226-316
 *
 * Conventions (all entry points)
 *   - plain pointers and sizes only; every pointer is DEVICE memory owned by the caller
 *     (the library never allocates or frees device memory -- the one exception is the peer-exchange buffer of
 *     ured_xchg_alloc, which must be an exportable cudaMalloc block; its only global state is a launch counter);
 *   - clouds are row-major contiguous float32 [count, n, 3]; distances float32; indices int32
 *     (same dtypes as the reference: dist_chamfer_3D.py:33-37);
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued asynchronously on it
 *     and no entry point synchronises the device;
 *   - return value: 0 on success, a positive cudaError_t value on a CUDA failure, a negative
 *     URED_E_* value on an argument error.  ured_last_error_string() describes the last
 *     failure seen by the calling thread.  (The reference printf()s and returns 0/1, and its
 *     Python wrapper ignores the result: chamfer3D.cu:145-151, dist_chamfer_3D.py:45.)
 *   - re-entrant: concurrent calls from several host threads / streams are allowed as long as
 *     they do not share output or workspace buffers.
 *
 * Pair addressing.  A call evaluates B ordered cloud pairs.  Pair b uses cloud-1 entry
 * (b / rep1) and cloud-2 entry (b % mod2); count1 = ceil(B / rep1) and count2 = min(B, mod2)
 * clouds must be present.  rep1 = 1, mod2 = B is the reference's plain batched call;
 * rep1 = K scores one target against its K candidates; rep1 = S, mod2 = S scores every target
 * against every one of S library shapes.
 */
#ifndef URED_CHAMFER_H_
#define URED_CHAMFER_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define URED_ABI_VERSION 1

/* argument errors (negative so that they never collide with cudaError_t) */
#define URED_E_NULL      (-1) /* a required pointer is NULL */
#define URED_E_SHAPE     (-2) /* negative size, or rep1/mod2 out of range */
#define URED_E_WORKSPACE (-3) /* workspace too small or misaligned (needs 256-byte alignment) */
#define URED_E_RANGE     (-4) /* k, n_lambda, ... out of the supported range */

/* flags for the forward entry points */
#define URED_FLAG_EXACT_ONLY 1u /* skip the screening pass; run the difference-form kernel on every pair (FP32 pipes) */
#define URED_FLAG_ONE_DIRECTION 4u /* ured_nn_packed: only cloud-1 points search cloud 2 (dist2/idx2 untouched, may be NULL) */
#define URED_FLAG_NON_REG    2u /* ured_dcd_forward with lengths: clamp the DCD fractions at 1 (calc_dcd non_reg=True) */
#define URED_FLAG_FP32_SCREEN 8u /* ured_nn_packed / ured_chamfer_forward: screen on the FP32 pipes (nn_kernel) instead of the tensor cores */
#define URED_NN_VARIANT_TENSOR 100 /* ured_nn_launch_shape: the tensor-core screening kernel (nn_tc_kernel) */

/* Ragged batches.  Entry points that take `len1` / `len2` (device int32 arrays, one entry per cloud-1 / cloud-2
 * entry, or NULL) treat cloud c as having only its first len[c] points (clamped to [0, n]); n is then the row
 * stride.  Outputs past a cloud's valid length are written as 0; a pair with an empty side yields zeros, like the
 * reference op which never touches its zero-filled outputs then.  This is what lets U-RED's per-sample / per-part
 * loop (loss/chamfer_loss.py:13-30, sizes mask.sum(1)*1024 and ragged target parts) run as two batched calls with
 * the lengths staying on the device. */

int ured_abi_version(void);
const char *ured_last_error_string(void);
/* number of kernels this library has launched in this process so far (statistics only) */
unsigned long long ured_kernel_launches(void);

/* ---- packed clouds ---------------------------------------------------------------------
 * The nearest-neighbour kernels read both clouds from a packed, padded structure-of-arrays image (the tensor-core
 * kernel builds its bf16 operand rows from it and copies X|Y|Z into shared memory with TMA bulk copies; the
 * FP32-pipe kernel streams all four arrays through shared memory the same way): per cloud one block
 *     X[np] | Y[np] | Z[np] | W[np] | tail[32]      (floats; np = n rounded up to 32)
 * with W = x^2+y^2+z^2, padding = copies of the last point, tail[0] = max W of the cloud.
 * Blocks are independent: a pointer to block i is a valid packed image of clouds i, i+1, ...
 * (block stride = ured_packed_bytes(1, n) rounded down to (4*np + 32) * 4 bytes), so a library that
 * is scored against many targets is packed once, kept resident, and addressed by slices. */
size_t ured_packed_bytes(int count, int n);
int ured_pack_clouds(const float *xyz, int count, int n, const int *len, void *packed, void *stream);

/* ---- nearest neighbours on packed clouds (both directions, one launch) --------------------
 * dist1/idx1: [B, n1]  for every point of cloud 1, squared distance to / index of its nearest
 *                      point in cloud 2 (lowest index on ties);  dist2/idx2: [B, n2] vice versa.
 * Queries are read from the packed images too, so xyz1 / xyz2 may be NULL when the corresponding image is given
 * (xyz1 is needed only with URED_FLAG_ONE_DIRECTION and packed1 == NULL).
 * Results are bit-identical to NmDistanceKernel (chamfer3D.cu:12-134) for finite inputs, whichever kernel runs:
 *   default                  nn_tc_kernel: the screening scores |c|^2 - 2 q.c come from the tensor cores (tcgen05.mma on
 *                            exact bf16x3 operand splits, fp32 accumulators in TMEM), the winning 32-candidate chunk of every
 *                            query is then evaluated in the reference's exact difference form; needs no scratch;
 *   URED_FLAG_FP32_SCREEN    nn_kernel: the same screen with packed FP32 FMAs;
 *   URED_FLAG_EXACT_ONLY     nn_kernel: the difference form on every pair.
 * Non-finite coordinates:
 * memory-safe (indices stay in range) but unspecified -- the reference admits a NaN distance only as the first
 * candidate of each of its 512-candidate tiles (chamfer3D.cu:36,126), an artefact this library does not reproduce;
 * the Python layer offers an opt-in check that raises instead (URED_CHECK_FINITE=1).
 *
 * FP32-pipe kernels only: for shapes whose grid would be too small (few pairs), very large clouds, and launches whose
 * last wave of CTAs would be mostly empty, the candidate range of some or all work items is split over several CTAs and
 * merged afterwards; that needs `scratch`: ured_nn_scratch_bytes(B, n1, n2) bytes (0 for most shapes, in which case
 * scratch may be NULL), 256-byte aligned.  (The tensor-core kernel scans clouds of more than 2048 candidates range by
 * range inside one CTA and ignores scratch.) */
size_t ured_nn_scratch_bytes(int B, int n1, int n2);
/* The launch plan ured_nn_packed will use for this problem and these flags (reporting / tests).  FP32-pipe kernels:
 * kernel variant id, queries per CTA, threads per CTA, the number of work items (pair, direction, query tile), how many
 * of them -- always the LAST ones of the launch -- are cut into `nsplit` candidate ranges, and nsplit itself (1 when
 * nothing is split).  Tensor-core kernel: variant = URED_NN_VARIANT_TENSOR, queries per work item (a group of 128-query
 * tiles), threads per CTA, work items (pair, direction, query group; taken in turn by one persistent CTA per SM),
 * nsplit = candidate ranges of 2048 scanned one after the other, split_items = 0.  Any output pointer may be NULL. */
int ured_nn_launch_shape(int B, int n1, int n2, unsigned flags, int *variant, int *queries_per_cta, int *threads, int *nsplit,
                         int *items, int *split_items);
int ured_nn_packed(const float *xyz1, const void *packed1, int n1,
                   const float *xyz2, const void *packed2, int n2,
                   int B, int rep1, int mod2, const int *len1, const int *len2,
                   float *dist1, float *dist2, int *idx1, int *idx2,
                   void *scratch, size_t scratch_bytes,
                   unsigned flags, void *stream);

/* ---- drop-in for chamfer_3D.forward ---------------------------------------------------------
 * workspace: ured_chamfer_workspace_bytes(B, n1, n2) bytes, 256-byte aligned. */
size_t ured_chamfer_workspace_bytes(int B, int n1, int n2);
int ured_chamfer_forward(const float *xyz1, const float *xyz2, int B, int n1, int n2,
                         const int *len1, const int *len2,
                         float *dist1, float *dist2, int *idx1, int *idx2,
                         void *workspace, size_t workspace_bytes,
                         unsigned flags, void *stream);

/* ---- drop-in for chamfer_3D.backward ---------------------------------------------------------
 * gradxyz1 [count1, n1, 3] and gradxyz2 [count2, n2, 3] are OVERWRITTEN (the reference
 * accumulates into caller-zeroed buffers; here the zero-fill is part of the call).
 * graddist1 / graddist2 may be NULL (treated as zeros). */
int ured_chamfer_backward(const float *xyz1, const float *xyz2, int B, int n1, int n2,
                          int rep1, int mod2, const int *len1, const int *len2,
                          const float *graddist1, const float *graddist2,
                          const int *idx1, const int *idx2,
                          float *gradxyz1, float *gradxyz2, void *stream);

/* Backward of the one-direction search (ured_nn_packed with URED_FLAG_ONE_DIRECTION; K=1 knn_points of
 * loss/basic_loss.py:249-265): only cloud-1 points carry a distance, so cloud-2 points get the scatter terms alone. */
int ured_nn_backward_one_direction(const float *xyz1, const float *xyz2, int B, int n1, int n2, const int *len2,
                                   const float *graddist1, const int *idx1,
                                   float *gradxyz1, float *gradxyz2, void *stream);

/* ---- calc_cd / calc_dcd epilogue ------------------------------------------------------------
 * From (dist1, idx1) [B, n1] and (dist2, idx2) [B, n2] of chamfer(gt, x) -- note the
 * reference's argument swap, model_utils.py:56: cloud 1 = gt, cloud 2 = x -- computes per pair
 *   cd_p = (mean sqrt(dist1) + mean sqrt(dist2)) / 2                     model_utils.py:57
 *   cd_t =  mean dist1 + mean dist2                                      model_utils.py:58
 *   loss = (mean_i(1 - exp(-alpha d1_i) w1_i) + mean_j(1 - exp(-alpha d2_j) w2_j)) / 2
 *          w1_i = frac_21 / (count1[idx1_i]^n_lambda + 1e-6), count1 = histogram of idx1
 *          over the n2 points of cloud 2; symmetric for w2 with frac_12     model_utils.py:31-45
 * ew1 [B, n1] / ew2 [B, n2] (optional, may be NULL) receive exp(-alpha d) * w per point, the
 * only per-point state the backward pass needs.
 *
 * Reduction order.  The per-pair means are accumulated in float32 in the order of torch's own reduction kernel for a
 * contiguous [B >= 16, n] tensor (one warp per row, 4-wide vector slots, shuffle-down tree; ATen Reduce.cuh), for
 * 128 < n < 8192, so that loss / cd_p / cd_t carry the same bits as the reference's `.mean(1)` calls on the same
 * dist/idx and a ranking cannot differ by a last-bit swap (tests/test_gpu_ranking.py holds this against torch).
 * Outside that range, and for ragged pairs, the order is still fixed (results are deterministic) but torch picks a
 * different block shape, so agreement is to float32 round-off only.
 *
 * ured_dcd_forward_ex additionally fills fscore [3, B] (f-score, precision_1, precision_2 at `f_threshold`,
 * metrics/CD/fscore.py:3-16; NULL to skip); loss/cd_p/cd_t/ew may each be NULL. */
int ured_dcd_forward_ex(const float *dist1, const float *dist2, const int *idx1, const int *idx2,
                        int B, int n1, int n2, int rep1, int mod2, const int *len1, const int *len2,
                        float alpha, float n_lambda, float frac_12, float frac_21, unsigned flags,
                        float *loss, float *cd_p, float *cd_t, float *ew1, float *ew2,
                        float *fscore, float f_threshold, void *stream);
/* the same without the F-score outputs */
int ured_dcd_forward(const float *dist1, const float *dist2, const int *idx1, const int *idx2,
                     int B, int n1, int n2, int rep1, int mod2, const int *len1, const int *len2,
                     float alpha, float n_lambda, float frac_12, float frac_21, unsigned flags,
                     float *loss, float *cd_p, float *cd_t,
                     float *ew1, float *ew2, void *stream);

/* Fused backward of [loss, cd_p, cd_t, dist1, dist2] w.r.t. both clouds: builds the per-point
 * d(out)/d(dist) coefficient and applies the Chamfer backward (chamfer3D.cu:155-174) in the
 * same pass.  Any of g_loss/g_cd_p/g_cd_t [B] and g_dist1 [B,n1] / g_dist2 [B,n2] may be NULL.
 * ew1/ew2 are required when g_loss is given.  gradxyz1/gradxyz2 are overwritten. */
int ured_dcd_backward(const float *xyz1, const float *xyz2, int B, int n1, int n2,
                      int rep1, int mod2, const int *len1, const int *len2,
                      const float *dist1, const float *dist2, const int *idx1, const int *idx2,
                      const float *ew1, const float *ew2, float alpha,
                      const float *g_loss, const float *g_cd_p, const float *g_cd_t,
                      const float *g_dist1, const float *g_dist2,
                      float *gradxyz1, float *gradxyz2, void *stream);

/* ---- ranking ----------------------------------------------------------------------------------
 * For each of `rows` score rows of length `cols`: the k smallest entries in ascending
 * (score, index) order -- torch.topk(..., largest=False) with its unspecified tie order pinned
 * to "lowest index first"; NaN sorts last.  out_idx gets index + idx_offset (global shape ids
 * of a library shard).  k <= cols is required; k <= 1024. */
int ured_topk_smallest(const float *scores, int rows, int cols, int k, int idx_offset,
                       float *out_scores, int *out_idx, void *stream);

/* Merge step of sharded retrieval: rows of `cols` candidates (score, id) gathered from all shards -> the k
 * smallest in ascending (score, id) order.  ids must be unique per row; ids < 0 mark padding; if a row holds
 * fewer than k real entries the tail is (+inf, -1). */
int ured_merge_topk(const float *scores, const int *ids, int rows, int cols, int k,
                    float *out_scores, int *out_ids, void *stream);

/* ---- EMD by the auction algorithm (re-rank of the Chamfer top-k) ---------------------------------------------------
 * Replaces the pybind module `emd` of the reference: emd.forward / emd.backward
 *     DCD/utils_v2/metrics/EMD/emd.cpp:13-30 -> emd_cuda.cu:226-316 (seven kernels per iteration, iterated from the host);
 * Python surface emdFunction / emdModule (EMD/emd_module.py:39-91) and calc_emd (utils_v2/model_utils.py:72-77).
 * xyz1 (prediction) and xyz2 (ground truth) are [B, n, 3] clouds of EQUAL size, coordinates expected in [0, 1] as in the
 * reference (its value "3.0 - distance - price" assumes it).  dist [B, n] receives the squared distance of every xyz1 point
 * to its assigned xyz2 point, assignment [B, n] that point's index (not necessarily a bijection, exactly as in the
 * reference: the last iteration assigns every remaining bidder to the object it bid on).  One launch runs the whole
 * auction: a cluster of 8 CTAs per pair, phases separated by cluster barriers.  Results equal the reference's whenever the
 * reference's own outcome is defined; when two bidders for one object are within its 1e-6 tolerance the reference lets the
 * last store win, this library the lowest point index.  No restriction on n (the reference needs n % 1024 == 0) or B (<= 512 there).
 * workspace: ured_emd_workspace_bytes(B, n) bytes, 256-byte aligned.  ured_emd_backward fills gradxyz1 [B, n, 3]
 * (the reference gives xyz2 no gradient: emd_module.py:81-84). */
size_t ured_emd_workspace_bytes(int B, int n);
int ured_emd_forward(const float *xyz1, const float *xyz2, int B, int n, float eps, int iters,
                     float *dist, int *assignment, void *workspace, size_t workspace_bytes, void *stream);
int ured_emd_backward(const float *xyz1, const float *xyz2, int B, int n,
                      const float *graddist, const int *assignment, float *gradxyz1, void *stream);

/* ---- measurement aid ------------------------------------------------------------------------------------------------
 * Launches a pure FFMA stream (8 independent chains per thread, blocks x 256 threads, 64 FFMA per iteration) and
 * reports its FLOP count; bench.py times it with CUDA events to obtain the FP32 FMA peak of the device it runs on. */
int ured_probe_ffma(float *sink, int blocks, int iters, double *flop, void *stream);

/* ---- sharded retrieval: fused local top-k + peer exchange + merge ---------------------------------------------------
 * The library is split over `world` ranks (one process per GPU); every rank scores its shard and needs the global k
 * best (score, shape id) per query.  The reference has no counterpart (it is single-GPU: engine/generate_pair.py:69-122
 * ranks with torch.topk, dataset/dataset_utils.py:1043-1051); SURVEY.md 8(e) defines the exchange.  Instead of
 * "top-k kernel, NCCL all_gather, merge kernel" ONE kernel selects the local top-k, stores it into every peer's
 * exchange buffer over NVLink, publishes a flag, waits for the peers' flags in its own buffer and merges.
 *
 * Set-up (once): every rank allocates a buffer of ured_xchg_bytes(world, rows, k) with ured_xchg_alloc, exports a
 * 64-byte handle, the handles travel over any side channel (torch.distributed all_gather in this package), every
 * rank imports the others' handles and passes the table bufs[world] (own buffer at [rank]) to ured_topk_exchange.
 * Any other peer mapping of buffers of that size (zero-initialised) works as well, e.g. torch symmetric memory.
 * This buffer is the only device memory the library ever allocates (it must be a whole cudaMalloc block to be
 * exportable).  Limits: world <= 16, rows <= 512 per call (all CTAs of a call wait for their peers while resident),
 * k <= 64.  All ranks must issue the same sequence of ured_topk_exchange calls with the same rows / k.
 * A peer that does not arrive within timeout_ms (0 = 5000) makes the call write ids = -2 and set the buffer's status
 * word (ured_xchg_status) instead of hanging the GPU. */
#define URED_XCHG_HANDLE_BYTES 64
size_t ured_xchg_bytes(int world, int rows, int k);
int ured_xchg_alloc(size_t bytes, void **dev_ptr);
int ured_xchg_free(void *dev_ptr);
int ured_xchg_export(void *dev_ptr, void *handle64);
int ured_xchg_import(const void *handle64, void **peer_ptr);
int ured_xchg_close(void *peer_ptr);
/* synchronises `stream`; status 0 = every exchange so far completed, 1 = some exchange timed out */
int ured_xchg_status(const void *own_buf, int *status, unsigned *epoch, void *stream);
/* scores [rows, cols] of this rank's shard (column c = shape id c + idx_offset) -> out_scores / out_ids [rows, k]:
 * the k smallest over ALL ranks in ascending (score, id) order, identical on every rank; (+inf, -1) pads a library
 * of fewer than k shapes.  cols may be 0 (an empty shard still takes part in the exchange). */
int ured_topk_exchange(const float *scores, int rows, int cols, int k, int idx_offset,
                       void *const *bufs, int world, int rank, int buf_rows,
                       float *out_scores, int *out_ids, unsigned timeout_ms, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* URED_CHAMFER_H_ */
