/*
 * dspmb.h -- C ABI of the B200-native multibox hot path (libdspmb.so).
 *
 * Drop-in boundary for the three MXNet contrib operators the reference patches into MXNet
 * (operator/multibox_{prior,target,detection}{-inl.h,.cc,.cu}) and for the Cython NMS helpers
 * (cython/{cpu_nms.pyx,gpu_nms.pyx,nms_kernel.cu,gpu_nms.hpp}).  Every entry point names the reference
 * interface it replaces (paths relative to the reference tree).
 *
 * Conventions
 *   - plain C types only; `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - unless marked HOST, pointers are device pointers on the current CUDA device, caller-owned,
 *     contiguous fp32 in exactly the reference's layouts; outputs are fully overwritten;
 *   - calls are asynchronous on `stream` (except the *_host entry) and allocate nothing: scratch comes
 *     from the caller via the *_workspace_bytes queries (the analogue of ResourceRequest::kTempSpace,
 *     operator/multibox_target-inl.h:258-261, multibox_detection-inl.h:183-186);
 *   - return value: 0 on success, a negative DSPMB_ERR_* code otherwise (the reference aborts through
 *     CHECK_* -> dmlc::Error; here the failure is a code plus dspmb_last_error()).  Data-dependent CHECKs
 *     (label padding, mining candidates) are evaluated on the device and latched in the workspace; read
 *     them with dspmb_status().
 *   - there is no CPU fallback: without a CUDA device every compute entry returns DSPMB_ERR_CUDA.
 */
#ifndef DSPMB_H_
#define DSPMB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSPMB_VERSION 100

#define DSPMB_OK 0
#define DSPMB_ERR_BAD_ARG -1           /* shape / parameter CHECKs of the Param ctors and InferShape       */
#define DSPMB_ERR_LABEL_PADDING -2     /* operator/multibox_target.cc:98-101                                */
#define DSPMB_ERR_MINING_CANDIDATES -3 /* operator/multibox_target.cc:236 CHECK_GE(temp.size(), num_neg)    */
#define DSPMB_ERR_MINING_THRESH -4     /* operator/multibox_target.cc:184 CHECK_GT(negative_mining_thresh)  */
#define DSPMB_ERR_WORKSPACE -5         /* workspace pointer NULL / too small / misaligned                   */
#define DSPMB_ERR_CUDA -6              /* CUDA runtime error (message in dspmb_last_error)                  */
#define DSPMB_ERR_INTERNAL -7          /* an internal invariant failed (never expected; please report)      */

int dspmb_version(void);

/* Thread-local description of the last failure in this thread ("" if none). */
const char *dspmb_last_error(void);

/*
 * The reference's CPU operators call the host libm (std::exp / std::log on float,
 * operator/multibox_target.cc:53-54,228,230 and multibox_detection.cc:117-118).  The kernels evaluate
 * glibc's expf/logf algorithm bit-for-bit in fp64; glibc selects an FMA or a non-FMA build of it at load
 * time from the CPU flags.  mode: 0 = non-FMA (sse2) variant, 1 = FMA variant, -1 = detect from this host.
 * Default is -1.  Returns the mode now in effect.
 */
int dspmb_set_libm_mode(int mode);

/* ---------------------------------------------------------------------------------------------------
 * MultiBoxPrior -- replaces MultiBoxPriorOp::Forward (operator/multibox_prior-inl.h:97-129) and
 * MultiBoxPriorForward (operator/multibox_prior.cc:29-71, .cu:61-101).
 * out: (in_height*in_width*(num_sizes+num_ratios-1), 4).  sizes/ratios: HOST arrays (op parameters).
 * steps <= 0 select the automatic 1/H, 1/W steps.  One launch for all anchor kinds.
 * ------------------------------------------------------------------------------------------------- */
int dspmb_prior_f32(float *out, int in_height, int in_width, const float *sizes, int num_sizes,
                    const float *ratios, int num_ratios, float step_y, float step_x, float offset_y,
                    float offset_x, int clip, void *stream);

/* All feature maps of a head in ONE launch, written back to back into out (sum_i H_i*W_i*K_i, 4) -- the
 * MultiBoxPrior-per-scale + Flatten + Concat of symbol/common.py:415-432.  All arrays HOST; sizes/ratios are
 * the per-map lists concatenated, with num_sizes[i]/num_ratios[i] entries for map i; steps holds (y, x) pairs
 * and offsets (y, x) pairs per map.  At most DSPMB_MAX_MAPS maps, DSPMB_MAX_KINDS sizes+ratios in total. */
#define DSPMB_MAX_MAPS 16
#define DSPMB_MAX_KINDS 128
int dspmb_prior_multi_f32(float *out, int num_maps, const int *heights, const int *widths, const float *sizes,
                          const int *num_sizes, const float *ratios, const int *num_ratios, const float *steps,
                          const float *offsets, int clip, void *stream);

/* ---------------------------------------------------------------------------------------------------
 * MultiBoxTarget -- replaces MultiBoxTargetOp::Forward (operator/multibox_target-inl.h:89-171) and
 * MultiBoxTargetForward (operator/multibox_target.cc:72-284; CPU semantics, not the divergent .cu ones).
 * anchors (A,4) | labels (B,L,label_width>=6) [cls,xmin,ymin,xmax,ymax,dist,...] | cls_preds (B,C,A)
 * -> loc_target (B,A*5), loc_mask (B,A*5), cls_target (B,A).
 * variances: HOST float[4].  Optional device outputs (NULL to skip):
 *   match_out (B,A) int32 : ground-truth index each positive anchor was matched to, -1 otherwise;
 *   stats_out (B,4) int32 : num_valid_gt, num_positive, num_negative, num_bipartite_matches.
 * ------------------------------------------------------------------------------------------------- */
size_t dspmb_target_workspace_bytes(int B, int A, int L, int C);
int dspmb_target_f32(const float *anchors, const float *labels, const float *cls_preds, float *loc_target,
                     float *loc_mask, float *cls_target, int B, int A, int L, int label_width, int C,
                     float overlap_threshold, float ignore_label, float negative_mining_ratio,
                     float negative_mining_thresh, int minimum_negative_samples, const float *variances,
                     int32_t *match_out, int32_t *stats_out, void *workspace, size_t workspace_bytes,
                     void *stream);

/* ---------------------------------------------------------------------------------------------------
 * MultiBoxDetection -- replaces MultiBoxDetectionOp::Forward (operator/multibox_detection-inl.h:81-107)
 * and MultiBoxDetectionForward (operator/multibox_detection.cc:53-169; CPU semantics).
 * cls_prob (B,C,A) | loc_pred (B,A*5) | anchors (A,4) -> out (B,A,7) [id,score,xmin,ymin,xmax,ymax,dist].
 * Optional device output valid_count_out (B) int32 = rows emitted by pass 1 per image.
 * ------------------------------------------------------------------------------------------------- */
size_t dspmb_detection_workspace_bytes(int B, int A, int C);
int dspmb_detection_f32(const float *cls_prob, const float *loc_pred, const float *anchors, float *out, int B,
                        int A, int C, float threshold, int clip, const float *variances, float nms_threshold,
                        int force_suppress, int nms_topk, int32_t *valid_count_out, void *workspace,
                        size_t workspace_bytes, void *stream);

/* MultiBoxDetection fed by the per-scale prediction heads -- replaces, in ONE operator,
 * the chain of symbol/common.py:399-412,424-432 and symbol/symbol_builder.py:161-165:
 *   per scale  transpose(0,2,3,1) -> Flatten, then Concat -> Reshape(0,-1,C) -> transpose(0,2,1)
 *   -> SoftmaxActivation(mode='channel') -> MultiBoxDetection.
 * cls_heads[k] (B, na_k*C, H_k, W_k), loc_heads[k] (B, na_k*5, H_k, W_k): device pointers to the conv outputs (NCHW,
 * channel = anchor_in_cell * C + class); head_hw = {H_0, W_0, H_1, W_1, ...}; head_na[k] = anchors per cell.  The class
 * tensor (B,C,A) is never written: the stream kernel reads the logits, evaluates the channel softmax (bit-exact glibc
 * expf for every anchor that can reach the threshold) and continues as dspmb_detection_f32.  C = 21 or 9; at most 8
 * scales; anchors as in dspmb_detection_f32 (scale-major, cell-major, then anchor-in-cell: the order
 * dspmb_prior_multi_f32 produces). */
size_t dspmb_detection_heads_workspace_bytes(int B, int A, int C, const int *head_hw, const int *head_na, int nscales);
int dspmb_detection_heads_f32(const float *const *cls_heads, const float *const *loc_heads, const int *head_hw,
                              const int *head_na, int nscales, const float *anchors, float *out, int B, int A, int C,
                              float threshold, int clip, const float *variances, float nms_threshold,
                              int force_suppress, int nms_topk, int32_t *valid_count_out, void *workspace,
                              size_t workspace_bytes, void *stream);

/* Forward of the training graph behind MultiBoxTarget and the MultiBoxMetric statistics
 * -- replaces symbol/symbol_builder.py:82-88 (SoftmaxOutput(ignore_label=-1, use_ignore, multi_output,
 * normalization='valid') forward = channel softmax; MakeLoss(smooth_l1(loc_mask * (loc_preds - loc_target), scalar=1)))
 * and the reductions of train/metric.py:27-46, in one pass over the tensors:
 *   cls_prob (B,C,A) and loc_loss (B,A*5): optional outputs (NULL: not written);
 *   stats (B,4) double, device: [#(cls_target >= 0), sum -log(prob[label] + eps), sum(loc_loss), #(loc_loss > 0)]
 *   (the last one is the count MakeLoss(normalization='valid') divides its gradient by).
 * Softmax as in multibox_target.cc:220-231 (MXNet's kernel is not part of the reference tree), bit-exact expf/logf. */
size_t dspmb_multibox_loss_workspace_bytes(int B, int A);
int dspmb_multibox_loss_f32(const float *cls_preds, const float *loc_preds, const float *loc_target,
                            const float *loc_mask, const float *cls_target, float *cls_prob, float *loc_loss,
                            double *stats, int B, int A, int C, float eps, void *workspace, size_t workspace_bytes,
                            void *stream);

/* Ordered compaction of the surviving detections: for every image the rows of `out` (B,A,7) with id >= 0, in
 * row order, at most K of them, into dst (B,K,7) (padded with -1) and their number into counts (B) -- the
 * `det[:,0] >= 0` filter of detect/multitask_detector.py:268-271 / multi_solver.py:419-432, on the device.
 * valid_count (B) optional: rows at and beyond it are known to be empty and are not scanned. */
int dspmb_detection_compact_f32(const float *out, const int32_t *valid_count, int B, int A, int K, float *dst,
                                int32_t *counts, void *stream);

/* ---------------------------------------------------------------------------------------------------
 * Fused compaction + all-gather over NVLink peer memory (the multi-GPU exchange step: surviving detections and
 * per-image target statistics).
 * Every rank allocates one gather buffer with dspmb_p2p_alloc(dspmb_gather_buffer_bytes(B, K, stats_width, world)),
 * publishes the 64-byte CUDA IPC handle to its peers (any host channel) and maps theirs with dspmb_p2p_open.
 * dspmb_detection_gather_f32 compacts the surviving rows (id >= 0, row order, at most K -- a multiple of 4, 0 to send
 * no rows --, padded with -1) of this rank's B images, appends `stats_width` int32 per image from `stats`
 * (B, stats_width; NULL with stats_width 0) and stores both into section `rank` of EVERY rank's buffer (HOST array
 * peer_bases[world] of device pointers, own buffer at index `rank`), slot 0/1; the last CTA publishes the sequence
 * number `seq` (use step + 1; it must grow by 2 between two uses of a slot) in every peer's flag word.
 * Flow control: a slot may only be rewritten after every reader has acknowledged the previous generation -- the
 * gather kernel itself waits (bounded) for ack >= seq - 2 from every rank, so each rank MUST call
 * dspmb_detection_gather_ack(slot, seq) once it has finished with generation seq (after gather_wait / gather_read),
 * whether it read the data or not.  dspmb_detection_gather_wait enqueues a kernel that returns once every rank's
 * flag for `slot` has reached `seq`.  A wait that runs out latches an error word: dspmb_gather_error (synchronous).
 * Buffer layout per slot: rows (world,B,K,7) float | counts (world,B) int32 | stats (world,B,stats_width) int32 |
 * flags (world) u64, each part padded to 256 B; then acks (2,world) u64, CTA counters, error word.
 * ------------------------------------------------------------------------------------------------- */
size_t dspmb_gather_buffer_bytes(int B, int K, int stats_width, int world);
int dspmb_p2p_alloc(size_t bytes, void **dev_ptr, unsigned char *ipc_handle_out /* 64 bytes */);
int dspmb_p2p_open(const unsigned char *ipc_handle, void **dev_ptr);
int dspmb_p2p_close(void *dev_ptr);
int dspmb_p2p_free(void *dev_ptr);
int dspmb_detection_gather_f32(const float *out, const int32_t *valid_count, const int32_t *stats, int B, int A, int K,
                               int stats_width, int rank, int world, void *const *peer_bases, int slot, long long seq,
                               void *stream);
int dspmb_detection_gather_wait(void *local_base, int B, int K, int stats_width, int world, int slot, long long seq,
                                void *stream);
int dspmb_detection_gather_ack(int B, int K, int stats_width, int rank, int world, void *const *peer_bases, int slot,
                               long long seq, void *stream);
/* The per-step exchange of a rank in one host call (what P2PDetectionGatherer.submit issues).  The context owns a side
 * stream and the events that order it against the caller's compute stream.  dspmb_gather_submit enqueues: compute
 * stream waits for the gather kernel that read `out` two steps ago (prev_seq > 0: the slot has been used before);
 * release_prev: bounded wait + acknowledgement, on the side stream, of generation prev_seq that this rank never read;
 * side stream waits for everything enqueued on compute_stream so far; the gather kernel of generation seq. */
void *dspmb_gather_ctx_create(void);
void *dspmb_gather_ctx_side_stream(void *ctx);
int dspmb_gather_ctx_destroy(void *ctx);
int dspmb_gather_submit(void *ctx, const float *out, const int32_t *valid_count, const int32_t *stats, int B, int A, int K,
                        int stats_width, int rank, int world, void *const *peer_bases, int slot, long long seq,
                        long long prev_seq, int release_prev, void *compute_stream);
/* Copies slot `slot` of this rank's buffer into rows_out (world*B, K, 7), counts_out (world*B) and stats_out
 * (world*B, stats_width) (device, async; NULL pointers are skipped). */
int dspmb_detection_gather_read(const void *local_base, int B, int K, int stats_width, int world, int slot,
                                float *rows_out, int32_t *counts_out, int32_t *stats_out, void *stream);
/* 0, or 1 if a bounded wait of the exchange ran out (synchronous read of the buffer's error word). */
int dspmb_gather_error(const void *local_base, int B, int K, int stats_width, int world);

/* Synchronises `stream` and returns the data-dependent status latched by the last target/detection call
 * that used `workspace` (0 or a DSPMB_ERR_* code). */
int dspmb_status(const void *workspace, void *stream);

/* ---------------------------------------------------------------------------------------------------
 * Standalone NMS -- replaces cpu_nms (cython/cpu_nms.pyx:17-68), gpu_nms (cython/gpu_nms.pyx:16-31) and
 * _nms (cython/nms_kernel.cu:91-144).  Pixel "+1" IoU convention.
 * mode 0: suppress iff (double)iou >= thresh  (cpu_nms);  mode 1: suppress iff iou > (float)thresh (gpu_nms,
 * detect/nms.py::nms).  dets (N,dim>=5) rows [x1,y1,x2,y2,score,(class)...].
 * class_col >= 5 restricts suppression to boxes with equal dets[:,class_col] (force_suppress off);
 * class_col < 0 is the reference's single-class behaviour.
 * presorted != 0: rows are already in descending score order (the _nms contract); otherwise the library sorts
 * on the device (score descending; ties -> higher index first, i.e. a stable argsort()[::-1]).
 * keep (N) int32 receives kept ORIGINAL row indices in score order, num_keep (1) int32 their count.
 * ------------------------------------------------------------------------------------------------- */
size_t dspmb_nms_workspace_bytes(int N);
int dspmb_nms_f32(const float *dets, int N, int dim, double thresh, int mode, int class_col, int presorted,
                  int32_t *keep, int32_t *num_keep, void *workspace, size_t workspace_bytes, void *stream);

/* Binary-compatible with `void _nms(int*, int*, const float*, int, int, float, int)` of cython/gpu_nms.hpp:1-2
 * apart from the int return: HOST pointers, boxes sorted by descending score, synchronous, selects
 * device_id, strict-greater rule.  keep_out receives sorted positions. */
int dspmb_nms_host(int *keep_out, int *num_out, const float *boxes_host, int boxes_num, int boxes_dim,
                   float nms_overlap_thresh, int device_id);

/* ---------------------------------------------------------------------------------------------------
 * bbox_overlaps_cython -- cython/bbox.pyx:15-55.
 * boxes (N,4), query_boxes (K,4) float64 [x1,y1,x2,y2] -> overlaps (N,K) float64, row-major; "+1" pixel convention,
 * 0 unless the intersection has positive width and height.  Device pointers, asynchronous on `stream`.
 * ------------------------------------------------------------------------------------------------- */
int dspmb_bbox_overlaps_f64(const double *boxes, int N, const double *query_boxes, int K, double *overlaps,
                            void *stream);

/* ---------------------------------------------------------------------------------------------------
 * Evaluation consumers of the detection output.
 *
 * dspmb_detection_postfilter_f32 -- multi_solver.py:419-432: per image the rows of `out` (B,A,7) with id >= 0 and
 *   score > score_thresh (0.25 there), in row order, at most K (200 there; the reference raises beyond, here the rest
 *   is cut), into rows (B,K,7) padded with -1; counts (B) int32.  valid_count (B) may bound the scan or be NULL.
 * dspmb_map_match_f32 -- the per-image TP/FP matching of MApMetric.update (evaluate/eval_metric.py:113-160):
 *   labels (B,L,label_width>=5) [cls,xmin,ymin,xmax,ymax,(difficult)], preds (B,M,pred_width>=6)
 *   [id,score,xmin,ymin,xmax,ymax,...] -> flags (B,M) int32: 0 not recorded (id < 0, or the best gt is difficult and
 *   use_difficult == 0), 1 true positive, 2 false positive.  Record accumulation and AP stay on the host
 *   (dspnet_b200.evalmap).  Device pointers, asynchronous on `stream`.
 * ------------------------------------------------------------------------------------------------- */
int dspmb_detection_postfilter_f32(const float *out, const int32_t *valid_count, int B, int A, int K,
                                   float score_thresh, float *rows, int32_t *counts, void *stream);
int dspmb_map_match_f32(const float *labels, int B, int L, int label_width, const float *preds, int M, int pred_width,
                        float ovp_thresh, int use_difficult, int32_t *flags, void *stream);

/* ---------------------------------------------------------------------------------------------------
 * Per-kernel timing (used by bench.py for the roofline line).  While enabled, every kernel the library launches
 * is bracketed by a cudaEvent pair recorded on the launching stream.  dspmb_profile_read synchronises those
 * events, adds the elapsed milliseconds and launch counts per kernel slot into ms[] / launches[] (up to
 * max_slots entries), clears the record list and returns the number of slots.  Slot names come from
 * dspmb_profile_kernel_name.  Not for use inside CUDA graph capture.
 * ------------------------------------------------------------------------------------------------- */
int dspmb_profile_enable(int on);
/* Kernels (and memset nodes) the last dspmb_detection_f32 / dspmb_target_f32 call of this thread put on the GPU,
 * whether launched directly or replayed from the graph cache (bench.py's gpu_launches). */
int dspmb_last_launch_count(void);
int dspmb_profile_read(float *ms, int *launches, int max_slots);
const char *dspmb_profile_kernel_name(int slot);

/* Path-selection knobs.  The detection pipeline picks between equivalent code paths by problem size (class
 * bucketing in the sort kernel vs. in the NMS kernel, bit-mask vs. chunk-sweep NMS, shared vs. global staging);
 * the tests use these to force every path on small inputs.  Returns the previous value, or -1 for a bad knob. */
#define DSPMB_TUNE_DET_STREAM_VARIANT 0 /* detection stream kernel: 0 generic, 1 persistent TMA ring, 2 (default) one
                                           TMA-fed tile per CTA, >2 register-resident (4: + 4-anchor target variant) */
#define DSPMB_TUNE_NMS_MASK_ROWS 1      /* largest segment for the shared-memory bit-mask NMS (default/max 320) */
#define DSPMB_TUNE_NMS_SMEM_ROWS 2      /* largest segment staged in shared memory by the sweep NMS (max 1024) */
#define DSPMB_TUNE_SORT_SMEM_KEYS 3     /* 64-bit sort keys kept in shared memory (default/max 8192)           */
#define DSPMB_TUNE_PHASES 4             /* bit mask of the launches a detection/target call performs (default 31:
                                           1 stream, 2 sort (+rank) / match, 4 nms or pair tests + resolve / select) --
                                           bench.py times one at a time                                             */
#define DSPMB_TUNE_GRAPH_CACHE 5        /* 1 (default): a detection/target call repeated with identical arguments is
                                           captured into a CUDA graph on its second sighting and replayed from then
                                           on (one cudaGraphLaunch instead of 3-4 kernel launches); 0: always launch */
#define DSPMB_TUNE_DET_PIPELINE 6       /* 1 (default): detection runs stream -> {sort || class pair tests + resolve}
                                           (the pair kernel is a programmatic dependent of the sort kernel) where its
                                           preconditions hold; 0: stream -> sort+rank -> nms in final row order      */
#define DSPMB_TUNE_TARGET_PIPELINE 7    /* 1: thread-block-cluster target matcher (8 CTAs per image, DSMEM histograms);
                                           0 (default): one 1024-thread CTA per image -- measured at SSD-512 B=64:
                                           84.6 us/step with one CTA, 88.8 us with the cluster (the matcher is a chain
                                           of ~8 dependent global-memory round trips and barriers, not a throughput
                                           problem, so eight times the threads buy nothing)                          */
#define DSPMB_TUNE_NMS_PIPELINE 8       /* 1 (default): tiled standalone NMS; 0: full-mask kernels                   */
#define DSPMB_TUNE_DET_PREFETCH 9        /* detection stream kernel: once a CTA's own tile has landed, one thread issues
                                           an L2 prefetch for the class tile of the CTA this many launches ahead, so
                                           that DRAM keeps working while the resident CTAs compute (the CTAs of a wave
                                           load together and compute together).  Default 600; measured stream kernel
                                           24.6 us without, 24.3 / 22.9 / 22.6 / 22.6 / 23.8 / 24.7 us at 150 / 300 /
                                           500 / 700 / 900 / 1332 CTAs.  (Issued at CTA start, as in round 2's first
                                           attempt, the prefetch cost +2.5 us.)  0 = off                              */
#define DSPMB_TUNE_TARGET_PREFETCH 10    /* the same late prefetch (logits tile of the CTA this many launches ahead) in
                                           the target stream kernel; default 0 = off: measured 47.4 us without, 49.3 /
                                           49.0 / 49.4 us at 300 / 740 / 1500 CTAs -- that kernel is bound by latency
                                           and instruction issue, its DRAM queues are never empty                    */
#define DSPMB_TUNE_DET_SPLIT 11           /* detection fork/join pipeline: image groups whose post-processing overlaps the
                                           stream kernel of the next group inside the graph (default 1 = no split, max 4;
                                           measured at SSD-512 B=32: 54.0 / 56.2 / 61.7 / 69.5 us for 1 / 2 / 3 / 4 groups:
                                           the 1024-thread sort CTAs of a group only find room once the next group's
                                           stream CTAs have drained, and the stream kernel itself slows down)          */
#define DSPMB_TUNE_DET_LEAN 12            /* 1: the TMA-fed detection stream kernel stages only the class rows; survivors
                                           fetch their loc_pred / anchor values from global memory (9 instead of 7
                                           CTAs per SM); 2 (default): the same with the class tile as ONE
                                           cp.async.bulk.tensor 2-D copy (tensor map, UTMALDG) instead of NFG 1-D bulk
                                           copies (measured 51.8 vs 53.3 us per step); 0: loc_pred and anchors are
                                           staged by bulk copies as well                                             */
#define DSPMB_TUNE_TARGET_SMALL 13        /* target stream kernel with ONE anchor per thread: 0 never, 1 (default) when the
                                           two-anchor grid is at most about two waves of CTAs (latency-bound batches:
                                           55.5 -> 53.0 us at 8 images of SSD-512, 57.5 -> 55.4 us at 16), 2 always       */
#define DSPMB_TUNE_TARGET_SHORTLIST 14    /* 1 (default): the target matcher mines the hard negatives on a shortlist -- a
                                           2048-key sample bounds the pivot, one pass compacts the keys below the bound
                                           into shared memory, range / histogram / pivot / final pass run on that list
                                           (verified by its count; falls back to the full walk); 0: always four passes
                                           over all A keys                                                            */
#define DSPMB_TUNE_TARGET_PDL 15          /* 1 (default): the target matcher is a programmatic dependent of the stream
                                           kernel (resident and through its label-only prologue before the stream
                                           kernel has drained); 0: plain stream order                                */
#define DSPMB_TUNE_DET_SORT_PDL 16        /* 1 (default): the detection sort kernel is a programmatic dependent of the stream
                                           kernel (resident while that grid drains; it releases the pair-test kernel
                                           only after its own griddepcontrol.wait); 0: plain stream order               */
#define DSPMB_TUNE_NMS_PDL 17             /* 1 (default): the cull / resolve kernels of the standalone NMS chunk loop are
                                           launched as programmatic dependents of their predecessors (resident before
                                           it ends, griddepcontrol.wait first); 0: plain launches                       */
#define DSPMB_NUM_TUNING 18
int dspmb_set_tuning(int knob, int value);

/* Debug timeline of the detection kernels: device_buffer (16 x 2 uint64, caller-initialised to UINT64_MAX / 0 pairs)
 * receives per kernel [earliest CTA start, latest CTA end] in %globaltimer ns (slots: 0 stream, 1 sort, 2 pair, 3 tail,
 * 4 resolve); NULL switches it off.  Not part of the operator boundary. */
int dspmb_debug_trace(unsigned long long *device_buffer);

/* Device self-test hooks used by the parity tests: y[i] = expf(x[i]) / logf(x[i]) through the same
 * glibc-compatible routines the kernels use.  x, y device pointers. */
int dspmb_test_expf(const float *x, float *y, long n, void *stream);
int dspmb_test_logf(const float *x, float *y, long n, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DSPMB_H_ */
