/*
 * einx.h -- C ABI of the B200-native (sm_100a) extraction-and-matching path of EI-Nexus.
 *
 * The reference (ZhonghuaYi/EI-Nexus_official) is pure Python/PyTorch and has no FFI; the
 * boundary it exposes for this path is four groups of Python callables.  Each entry point
 * below replaces the ATen op sequence behind one of them (paths relative to the reference
 * root); the Python host layer in ei-nexus_official_b200/ keeps the reference signatures and
 * binds these symbols with ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - every data pointer is a DEVICE pointer owned by the caller (outputs included);
 *   - calls are asynchronous on `stream` (a cudaStream_t passed as void*) and never
 *     synchronise with the host;
 *   - return value: 0 (EINX_OK) or a negative EINX_ERR_*; text via einx_last_error();
 *   - no C++ exception crosses the ABI; one context is not thread-safe, distinct contexts are;
 *   - sm_100a only: einx_create() fails with EINX_ERR_ARCH on any other device.  There is no
 *     CPU or generic-GPU fallback anywhere behind this header.
 */
#ifndef EINX_H_
#define EINX_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EINX_OK 0
#define EINX_ERR_INVALID (-1)     /* bad argument                                   */
#define EINX_ERR_CUDA (-2)        /* a CUDA runtime call failed                     */
#define EINX_ERR_ARCH (-3)        /* device is not compute capability 10.0 (B200)   */
#define EINX_ERR_NOMEM (-4)       /* workspace allocation failed                    */
#define EINX_ERR_UNSUPPORTED (-5) /* shape outside what the kernels were built for  */

#define EINX_SAMPLE_GATHER 0   /* sparsify_full_resolution_descriptors */
#define EINX_SAMPLE_BILINEAR 1 /* sparsify_low_resolution_descriptors  */
#define EINX_SAMPLE_GATHER_NHWC 2 /* sparsify_full_resolution_descriptors on a channels-last map: raw is
                                   * (B, Hd, Wd, C), the memory of a torch.channels_last (B, C, Hd, Wd) tensor --
                                   * a keypoint's C channels are one contiguous, coalesced read              */

#define EINX_MNN_FP32 0  /* FFMA tiles, fp32 accumulate: index-exact reference path */
#define EINX_MNN_TF32X3 1 /* tcgen05 kind::tf32, 3-term split (hi*hi + hi*lo + lo*hi) */
#define EINX_MNN_BF16 2  /* tcgen05 kind::f16 on bf16-rounded descriptors             */
#define EINX_MNN_FP16X3 3 /* tcgen05 kind::f16, 3-term fp16 split of 2^10 * d (same 22 significant bits as
                           * TF32X3 at twice the tensor rate); requires |d| < 63, which L2-normalised
                           * descriptors times any scale_factor < 63 satisfy -- larger values saturate */

typedef struct einx_ctx einx_ctx;
typedef void* einx_stream; /* cudaStream_t */

/* ABI version of this header (major*100 + minor). */
int einx_version(void);

/* Context: owns the per-device workspace (NMS survivor lists, radix-select histograms,
 * voxel statistics, MNN row/column best keys).  `device` is a CUDA ordinal. */
int einx_create(int device, einx_ctx** out);
void einx_destroy(einx_ctx* ctx);
/* Last error text of `ctx`; pass NULL for the error of a failed einx_create(). */
const char* einx_last_error(const einx_ctx* ctx);

/*
 * Event voxelisation.  Replaces datasets/representations.py:8-22 (time_normalization) and
 * :66-124 (events_to_voxel_grid) for a ragged batch of B event windows.
 *   x, y, p   : (N_total) fp32 -- the reference casts to float32 at :73-75
 *   t         : (N_total) fp64 -- stays double across the ABI: the offset subtraction of :19-20
 *               happens in fp64 before the fp32 cast of :76 (epoch-scale timestamps)
 *   ev_offsets: (B+1) int64, window b owns events [ev_offsets[b], ev_offsets[b+1]), time-sorted
 *   out       : (B, bins, H, W) fp32, fully overwritten
 *   normalize : non-zero -> mean / unbiased-std normalisation over cells != 0 (:114-122)
 * A window with a single event yields an all-zero grid (the reference's 0/0 time maps to an
 * out-of-range bin); empty windows yield zeros (the reference raises IndexError -- the host
 * layer keeps that behaviour).
 */
int einx_voxelize(einx_ctx* ctx, const float* x, const float* y, const double* t, const float* p,
                  const int64_t* ev_offsets, int B, int bins, int H, int W, int normalize,
                  float* out, einx_stream stream);

/*
 * Detection post-processing.  Replaces core/modules/utils/detector_util.py:80-135
 * (prob_map_to_points_map: remove_border_points :138-164, fast_nms :243-337, the quantile
 * top-k threshold :108-133) and :451-484 (prob_map_to_positions_with_prob, ordering 'yx').
 *   score     : (B, Hp, Wp) fp32 >= 0, border frame zeroed IN PLACE like the reference
 *   mask      : optional (B, Hp, Wp) uint8; where 0 the score is zeroed first (in place), i.e.
 *               `score[~mask] = 0` of EventExtractors.py:374-375 fused in; NULL to skip
 *   top_k     : <= 0 means None; prob_thresh as in the reference (1.0 in every shipped config)
 *   nms_map   : optional (B, Hp, Wp) fp32 dense result (the `nms` tensor); NULL to skip
 *   kpts      : (B, kcap, 3) fp32 rows (y+0.5, x+0.5, prob) in raster order; rows >= count
 *               are left untouched
 *   counts    : (B) int32 number of keypoints found (may exceed kcap: only kcap rows written)
 * Bit-exact with the reference for non-negative, NaN-free maps.
 */
int einx_detect(einx_ctx* ctx, float* score, const uint8_t* mask, int B, int Hp, int Wp,
                int nms_radius, int border, float prob_thresh, int top_k, float* nms_map,
                float* kpts, int kcap, int32_t* counts, einx_stream stream);

/*
 * The same for the two sides of a batch of pairs in ONE launch (EIM.forward runs the extractor on the event side
 * and on the image side, core/modules/EIM.py:89-93: two independent batches of B maps of one size).  Every argument
 * is as in einx_detect, once per side; B is the batch of one side.  Each image is owned by one CTA (or cluster), so
 * 2B images fill the machine where two launches of B would run back to back.
 */
int einx_detect_pair(einx_ctx* ctx, float* score0, float* score1, const uint8_t* mask0,
                     const uint8_t* mask1, int B, int Hp, int Wp, int nms_radius, int border,
                     float prob_thresh, int top_k, float* nms_map0, float* nms_map1, float* kpts0,
                     float* kpts1, int kcap, int32_t* counts0, int32_t* counts1, einx_stream stream);

/*
 * Descriptor sampling + L2 normalisation.  Replaces core/modules/utils/descriptor_util.py:21-28
 * (normalize_descriptors), :50-71 (gather, SiLK type) and :74-128 (bilinear grid_sample with
 * align_corners=False on the padded-image grid, SuperPoint type).
 *   raw   : (B, C, Hd, Wd) fp32 descriptor map (Hd, Wd = padded image for gather; coarse map
 *           for bilinear); (B, Hd, Wd, C) for EINX_SAMPLE_GATHER_NHWC
 *   Hp,Wp : padded image size the positions refer to (bilinear mode only)
 *   kpts  : (B, kcap, 3) as written by einx_detect; counts (B) int32 (clamped to kcap)
 *   desc  : (B, kcap, C) fp32; rows >= count are zero-filled
 */
int einx_sample(einx_ctx* ctx, const float* raw, int B, int C, int Hd, int Wd, int mode, int Hp,
                int Wp, const float* kpts, const int32_t* counts, int kcap, float scale,
                int normalize, float* desc, einx_stream stream);

/*
 * einx_sample with a second output for the matcher: `split` (NULL to skip) receives the fp16 operand pair of
 * the FP16X3 tensor-core mode, [hi (B, kcap, C) | lo (B, kcap, C)] with hi = fp16(2^10 d), lo = fp16(2^10 d - hi),
 * written while the descriptor is in registers.  einx_mnn_split() consumes it, so the similarity kernel's tile
 * pipeline is TMA -> MMA with nothing to convert.  Needs C % 8 == 0 and a 16-byte aligned buffer of
 * 2 * B * kcap * C halves.
 */
int einx_sample_split(einx_ctx* ctx, const float* raw, int B, int C, int Hd, int Wd, int mode,
                      int Hp, int Wp, const float* kpts, const int32_t* counts, int kcap, float scale,
                      int normalize, float* desc, uint16_t* split, einx_stream stream);

/*
 * Mutual-nearest-neighbour matching.  Replaces core/modules/matchers/MNN.py:11-22 (find_nn),
 * :25-32 (mutual_check) and :88-129 of NearestNeighborMatcher.forward.  The similarity matrix
 * is never written to memory: row/column argmax are fused into the tile epilogue.
 *   d0 (B, ncap, D), d1 (B, mcap, D) fp32; n0, n1 (B) int32 valid rows (NULL = all)
 *   ratio_thresh / distance_thresh <= 0 disable the test (None/False in the reference)
 *   m0 (B, ncap), m1 (B, mcap) int64, -1 = unmatched; s0, s1 fp32 (m > -1)
 *   kpts0/kpts1 (B, cap, 3) + mk0/mk1 (B, ncap, 3) + nmatch (B): optional compaction of the
 *   matched keypoint rows in ascending i (:103-129); pass NULL kpts0 to skip
 * Ties resolve to the lowest index like topk(1).
 */
int einx_mnn(einx_ctx* ctx, const float* d0, const float* d1, const int32_t* n0,
             const int32_t* n1, int B, int ncap, int mcap, int D, float ratio_thresh,
             float distance_thresh, int mutual, int precision, int64_t* m0, int64_t* m1,
             float* s0, float* s1, const float* kpts0, const float* kpts1, float* mk0,
             float* mk1, int32_t* nmatch, einx_stream stream);

/*
 * einx_mnn with the FP16X3 operands already split (einx_sample_split): split0 / split1 are the [hi | lo] fp16
 * buffers of d0 / d1 (both or neither; NULL = derive them from d0 / d1 in one pre-pass, which is what einx_mnn
 * does).  Ignored by the other precisions.  Results are identical either way.
 */
int einx_mnn_split(einx_ctx* ctx, const float* d0, const float* d1, const uint16_t* split0,
                   const uint16_t* split1, const int32_t* n0, const int32_t* n1, int B, int ncap,
                   int mcap, int D, float ratio_thresh, float distance_thresh, int mutual,
                   int precision, int64_t* m0, int64_t* m1, float* s0, float* s1, const float* kpts0,
                   const float* kpts1, float* mk0, float* mk1, int32_t* nmatch, einx_stream stream);

/*
 * Opt-in dense by-products of MNN.py:88,96-98 for callers that really want them
 * (`similarity` (B,N,M) and `log_assignment` (B,N+1,M+1)); fp32 FFMA path.
 */
int einx_mnn_dense(einx_ctx* ctx, const float* d0, const float* d1, int B, int N, int M, int D,
                   float* similarity, float* log_assignment, einx_stream stream);

/* ---- rows adjacent to the path (SURVEY.md section 8 f) ---------------------------------------- */

/*
 * Event accumulation image.  Replaces datasets/visualize.py:23-49 (draw_events_accumulation_image,
 * dict branch: one count per event at (int(y), int(x)), then (c - min) / (max - min) * 255 in fp64,
 * clipped to 255 and truncated to uint8) for a ragged batch -- the reference runs a per-event Python
 * loop here (datasets/MVSEC.py:850, datasets/EC.py:300).
 *   x, y      : (N_total) event coordinates, fp64 if coord_f64 != 0 (what the datasets hold:
 *               truncation then matches the reference bit for bit) else fp32 (the voxeliser's SoA;
 *               identical unless a coordinate lies within one fp32 ulp below an integer)
 *   image     : (B, H, W) uint8.  A window whose counts are all equal (max == min) yields zeros
 *               (the reference divides 0/0 and casts NaN).  Events outside [0,W)x[0,H) are skipped
 *               (the reference raises IndexError or wraps negative indices).
 */
int einx_events_image(einx_ctx* ctx, const void* x, const void* y, int coord_f64,
                      const int64_t* ev_offsets, int B, int H, int W, uint8_t* image,
                      einx_stream stream);

/*
 * The (N, 4) ndarray branch of the same function (datasets/visualize.py:41-44): every event adds 2 * p - 1 to its
 * pixel (a signed histogram), then the same min-max normalisation.  p has the dtype of x / y; 2 * p - 1 must be
 * integer-valued (polarities 0/1 or -1/1; the host layer checks), which keeps the fp64 sums of the reference exact.
 */
int einx_events_image_signed(einx_ctx* ctx, const void* x, const void* y, const void* p, int coord_f64,
                             const int64_t* ev_offsets, int B, int H, int W, uint8_t* image,
                             einx_stream stream);

/*
 * The reference's other scatter representations, selectable through `representation_type`
 * (datasets/MVSEC.py:706-718).  Same ragged SoA events as einx_voxelize (x, y, p fp32; t fp64, time-sorted);
 * time_normalization (datasets/representations.py:8-22) runs on the device in fp64, and an event belongs to
 * bin i iff  i*dt <= t <= i*dt + dt  with both ends inclusive -- exactly np.searchsorted(.., 'left') ..
 * searchsorted(.., 'right') of the reference, so boundary events count in two bins.  Events outside
 * [0,W)x[0,H) are skipped.
 *   einx_event_stack : datasets/representations.py:177-214 -- out (B, bins, H, W) fp32, the integer sum of
 *                      2*int(p) - 1 per cell.  Bit-exact.
 *   einx_time_surface: datasets/representations.py:25-63  -- out (B, bins, H, W) fp32 with bins // 2 time
 *                      bins; channel 2*i + int(p) holds the latest normalised time of that polarity (the
 *                      reference's last-write-wins fancy assignment on time-sorted events = the maximum).
 *                      Polarities other than 0/1 have no channel and are skipped.  Bit-exact.
 */
int einx_event_stack(einx_ctx* ctx, const float* x, const float* y, const double* t, const float* p,
                     const int64_t* ev_offsets, int B, int bins, int H, int W, float* out,
                     einx_stream stream);
int einx_time_surface(einx_ctx* ctx, const float* x, const float* y, const double* t, const float* p,
                      const int64_t* ev_offsets, int B, int bins, int H, int W, float* out,
                      einx_stream stream);

/*
 * Event distance map.  Replaces datasets/representations.py:215-248 (events_to_distance_map): per time bin
 * (i/bins <= t <= (i+1)/bins after time_normalization, both ends inclusive) the event pixels are marked and
 * cv.distanceTransform(1 - event_map, cv.DIST_L2, 3) gives every pixel its 3x3 chamfer distance (axial 0.955,
 * diagonal 1.3693) to the nearest event pixel.  Computed in closed form, fp64 rounded once: within 1e-6 relative
 * of OpenCV (whose own IPP / fixed-point paths differ from each other by that much), bit-exact against the
 * oracle.  A bin without events is FLT_MAX everywhere.  Polarity is not used.  Events outside [0,W)x[0,H) are
 * skipped (the reference wraps negative indices or raises).
 *   x, y fp32, t fp64 time-sorted, ev_offsets (B + 1) int64; out (B, bins, H, W) fp32; H, W <= 65534.
 */
int einx_distance_map(einx_ctx* ctx, const float* x, const float* y, const double* t,
                      const int64_t* ev_offsets, int B, int bins, int H, int W, float* out,
                      einx_stream stream);

/*
 * Compact host->device wire format for integer-pixel events (what the sensor and the EC dataset
 * deliver: datasets/rectify_ec.py:66-83 rounds to pixels, polarity 0/1): x, y as uint16 and p as int8
 * expand to the fp32 arrays einx_voxelize reads -- exactly the values the reference's
 * astype(float32) of representations.py:73-75 produces -- so an event crosses PCIe in 13 bytes
 * (with its fp64 timestamp) instead of 20.  The host layer only chooses this format when it is lossless.
 */
int einx_unpack_events(einx_ctx* ctx, const uint16_t* x, const uint16_t* y, const int8_t* p, int64_t n,
                       float* xo, float* yo, float* po, einx_stream stream);

/*
 * Event mask for the detector.  Replaces `events_image > 0` (train_extractor.py:225), the constant
 * padding of Padder.pad for bool tensors (core/modules/utils/util.py:17-32) and the 3x3 box
 * convolution + `> 0` of core/modules/event_extractors/EventExtractors.py:357-363, i.e. a 3x3 binary
 * dilation of the zero-extended mask.  The result is the `mask` argument of einx_detect, which applies
 * `score[~mask] = 0` (:374-375) on load.
 *   image (B, H, W) uint8 -> mask (B, Hp, Wp) uint8 in {0,1}; the image sits at (pad_top, pad_left).
 */
int einx_mask_dilate(einx_ctx* ctx, const uint8_t* image, int B, int H, int W, int pad_top,
                     int pad_left, int Hp, int Wp, uint8_t* mask, einx_stream stream);

/*
 * Detector head post-processing.  Replaces core/modules/utils/detector_util.py:18-39
 * (logits_to_prob: softmax over the channel dimension, or 1 / (1 + exp(-x)) for one channel) and
 * :42-77 (depth_to_space: drop the dustbin channel, pixel-shuffle by `cell`).
 *   logits : (B, C, Hc, Wc) fp32 with C == cell*cell + 1 (cell > 1) or C == 1 (cell == 1)
 *   mode   : EINX_HEAD_SCORE  -> out (B, 1, Hc*cell, Wc*cell)   depth_to_space(logits_to_prob(x))
 *            EINX_HEAD_PROB   -> out (B, C, Hc, Wc)             logits_to_prob(x)
 *            EINX_HEAD_SHUFFLE-> out (B, 1, Hc*cell, Wc*cell)   depth_to_space(x), x already probabilities
 * fp32 exp / sum: agrees with torch to a few ulp (1e-6 relative), not bit-exact.
 */
#define EINX_HEAD_SCORE 0
#define EINX_HEAD_PROB 1
#define EINX_HEAD_SHUFFLE 2
int einx_logits_to_score(einx_ctx* ctx, const float* logits, int B, int C, int Hc, int Wc, int cell,
                         int mode, float* out, einx_stream stream);

/*
 * LightGlue match filtering.  Replaces core/modules/matchers/lightglue.py:402-418 (filter_matches)
 * on a log-assignment matrix: row / column argmax of scores[:, :-1, :-1] (first index on ties),
 * mutual check, exp of the row maxima, threshold.  One pass over the matrix.
 *   scores (B, M+1, N+1) fp32; m0 (B, M), m1 (B, N) int64 (-1 = none); ms0 (B, M), ms1 (B, N) fp32
 */
int einx_filter_matches(einx_ctx* ctx, const float* scores, int B, int M, int N, float th,
                        int64_t* m0, int64_t* m1, float* ms0, float* ms1, einx_stream stream);

/*
 * LightGlue log-assignment matrix.  Replaces core/modules/matchers/lightglue.py:365-377
 * (sigmoid_log_double_softmax, called from MatchAssignment.forward :393):
 *   scores[b, i, j] = log_softmax(sim, 2)[b, i, j] + log_softmax(sim, 1)[b, i, j]
 *                     + logsigmoid(z0[b, i]) + logsigmoid(z1[b, j])          i < M, j < N
 *   scores[b, i, N] = logsigmoid(-z0[b, i]);  scores[b, M, j] = logsigmoid(-z1[b, j]);  scores[b, M, N] = 0
 *   sim (B, M, N), z0 (B, M), z1 (B, N), scores (B, M+1, N+1), all fp32 and contiguous.
 * Row and column (max, log-sum-exp) statistics in one pass over sim, the matrix written in a second
 * pass (12 bytes of HBM traffic per element).
 * fp32 exp / log / summation order: agrees with torch to ~1e-6 of the magnitude, not bit-exact.
 *   best_keys : optional (B*M + B*N) uint64, NULL to skip.  The write pass then also reduces the values it
 *               stores to (max, first index) per row and per column of scores[:, :-1, :-1]: what
 *               filter_matches (:402-418) would recompute from the matrix.  einx_filter_matches_keys()
 *               turns them into matches without another pass over the matrix, with results identical
 *               to einx_filter_matches() on `scores`.  Opaque layout: rows of all items, then columns.
 */
int einx_log_double_softmax(einx_ctx* ctx, const float* sim, const float* z0, const float* z1,
                            int B, int M, int N, float* scores, uint64_t* best_keys,
                            einx_stream stream);
int einx_filter_matches_keys(einx_ctx* ctx, const uint64_t* best_keys, int B, int M, int N, float th,
                             int64_t* m0, int64_t* m1, float* ms0, float* ms1, einx_stream stream);

/*
 * Metric-side N x M reductions (the distance matrix is reduced along both axes and never stored).
 *
 * einx_pairwise_min_dist -- core/metrics/keypoints_metrics.py:110-124 (Repeatability.update_one):
 *   norm = ||a[:, None, :2] - b[None, :, :2]||_2  (difference in fp32, squares / sum / sqrt in fp64, rounded to
 *   fp32: what torch.linalg.norm does on the CPU);  rowmin[i] = min_j norm[i, j]  (torch.min(norm, 1), :120),
 *   colmin[j] = min_i norm[i, j]  (torch.min(norm, 0), :117).  The caller counts `<= distance_thresh`.
 *   a (B, N, 2), b (B, M, 2) fp32; na, nb (B) int32 valid counts or NULL; rowmin (B, N), colmin (B, M) fp32
 *   (+inf beyond the counts and for an empty other side).
 *
 * einx_gt_assign -- core/geometry/gt_generation.py:96-126 (gt_matches_from_pose_depth, between `project` and the
 *   epipolar pass):  dist0 = |kp0_1[i] - kp1[j]|^2, dist1 = |kp0[i] - kp1_0[j]|^2 (fp32), dist = max(dist0, dist1)
 *   where visible0[i] & visible1[j] else +inf; min0 / min1 = first argmin along j / i; positive = mutual argmin
 *   & dist < pos_th^2; negative0 = min_j dist0 > neg_th^2 & valid0 (negative1 alike with dist1);
 *   m0 = -1 where negative0, else min0 where positive, else -2 (IGNORE_FEATURE); m1 alike.
 *   kp* (B, N|M, 2) fp32 in the order the reference indexes them (after its `ordering` flip); visible*, valid*
 *   (B, N|M) uint8; m0 (B, N), m1 (B, M) int64; min0, min1 optional int32 argmins (NULL to skip) from which the
 *   caller scatters the dense `assignment`.  N, M >= 1 (an empty side is the reference's early return, :63-71).
 */
int einx_pairwise_min_dist(einx_ctx* ctx, const float* a, const float* b, const int32_t* na,
                           const int32_t* nb, int B, int N, int M, float* rowmin, float* colmin,
                           einx_stream stream);
int einx_gt_assign(einx_ctx* ctx, const float* kp0, const float* kp1, const float* kp0_1,
                   const float* kp1_0, const uint8_t* visible0, const uint8_t* visible1,
                   const uint8_t* valid0, const uint8_t* valid1, int B, int N, int M, float pos_th,
                   float neg_th, int64_t* m0, int64_t* m1, int32_t* min0, int32_t* min1,
                   einx_stream stream);

/* Number of kernel launches issued through `ctx` so far (bench.py's gpu_launches). */
int64_t einx_launch_count(const einx_ctx* ctx);

/*
 * Measurement hook.  With profiling on, every entry point brackets its dominant kernel with CUDA
 * events recorded on the caller's stream (slot 0: voxel scatter, 1: detect, 2: sample, 3: MNN
 * similarity tiles).  einx_profile_read() waits for those events and returns the elapsed
 * milliseconds of the most recent launch per slot (-1 if none); it is the only call in this
 * header that synchronises with the host.
 */
#define EINX_PROFILE_SLOTS 4
int einx_profile_enable(einx_ctx* ctx, int on);
int einx_profile_read(einx_ctx* ctx, float* ms_out /* [EINX_PROFILE_SLOTS] */);

#ifdef __cplusplus
}
#endif
#endif /* EINX_H_ */
