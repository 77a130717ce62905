/*
 * vdet_b200.h -- C ABI of libvdet_b200.so: the B200 (sm_100a) implementation of vdetlib's
 * post-CNN hot path (per-frame NMS, IoU scoring / tubelet linking, temporal smoothing).
 *
 * vdetlib has no FFI layer: its only compiled boundary is the CPython extension
 * `utils.cython_nms` (reference setup.py:8-15, built from utils/nms.pyx) plus plain NumPy
 * functions.  Each entry point below names the reference interface it stands in for; the
 * Python adapters in vdetlib_b200/{utils,vdet}/ keep the reference's function names and
 * call these through ctypes (see INTEGRATION.md for the binding a vdetlib maintainer adds).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *   - plain pointers and sizes only, no torch / numpy types;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); entry points are
 *     asynchronous unless documented "synchronous";
 *   - no allocation inside: outputs and scratch (`ws`, sized by the matching
 *     *_workspace_bytes query) are supplied by the caller;
 *   - return value: VDET_OK or a negative VDET_ERR_* ; vdet_last_error() gives the
 *     thread-local message.  Data-dependent conditions that the reference reports as Python
 *     exceptions (ZeroDivisionError of nms.pyx:64) are written to a caller-supplied device
 *     `status` word (bit flags VDET_STATUS_*), read by the adapter after it synchronises.
 *   - "segments" are frames: rows [seg_offsets[s], seg_offsets[s+1]) of the packed arrays
 *     belong to frame s.  Boxes are (x1,y1,x2,y2), "+1" pixel convention throughout.
 */
#ifndef VDET_B200_H_
#define VDET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VDET_ABI_VERSION 2

#define VDET_OK                 0
#define VDET_ERR_INVALID       (-1)   /* bad argument (adapter raises ValueError)            */
#define VDET_ERR_CUDA          (-2)   /* CUDA runtime error (adapter raises RuntimeError)    */
#define VDET_ERR_WORKSPACE     (-3)   /* ws_bytes too small                                  */
#define VDET_ERR_UNSUPPORTED   (-4)   /* size outside what this build handles                */

#define VDET_STATUS_ZERO_DIVISION 1u  /* a same-frame pair had union == 0 (nms.pyx:64)       */
#define VDET_STATUS_ALL_MISSING   2u  /* a tubelet row had no valid score (tubelet_cls.py:295 IndexError) */

#define VDET_DTYPE_F32 0
#define VDET_DTYPE_F64 1

#define VDET_POOL_ARGMAX_SCORE 0      /* dets_spatial_max_pooling, tubelet_cls.py:334-340    */
#define VDET_POOL_ARGMAX_IOU   1      /* anchor_propagate, tubelet_cls.py:375-377            */
#define VDET_POOL_MAX_IOU      2      /* tubelets_overlap, utils/protocol.py:467-489: out_score = max IoU */

#define VDET_PAD_ZERO 0
#define VDET_PAD_EDGE 1

/* Output layouts of vdet_nms_frames_f32 (n_s = boxes of frame s, o_s = seg_offsets[s]):
 *   CLASS_MAJOR: keep_idx / keep_mask [C, n_rows]: (class c, frame s) block at c*n_rows + o_s;
 *                keep_cnt [C, S].  One plane per class = what apply_vid_nms returns per class.
 *   FRAME_MAJOR: (frame s, class c) block at o_s*C + c*n_s; keep_cnt [S, C].  A contiguous range
 *                of frames is a contiguous range of every output (chunked / pipelined D2H). */
#define VDET_LAYOUT_CLASS_MAJOR 0
#define VDET_LAYOUT_FRAME_MAJOR 1

int         vdet_abi_version(void);
const char* vdet_last_error(void);
/* Number of SMs of `device` (grid sizing for callers); <0 on error. */
int         vdet_sm_count(int device);
/* Persistent kernels (the NMS grid) normally occupy every SM.  A multi-GPU caller that overlaps a
 * collective with them reserves a few SMs so that the collective's kernel can be scheduled at
 * once instead of waiting for the persistent grid to drain (process-wide setting, default 0). */
int         vdet_set_reserved_sms(int n);
/* Host-side staging copy into a pinned upload buffer with non-temporal (streaming) stores.
 * A buffer filled with ordinary stores sits dirty in the CPU caches and the GPU's DMA reads of it
 * run at about half the PCIe rate on the measured host (profiles/r01_pcie.md); streamed lines are
 * in DRAM when the copy engine asks for them.  Plain host memory in and out, no CUDA call. */
int         vdet_host_copy_stream(void* dst, const void* src, size_t bytes);
/* The same over `n_threads` host threads, each streaming a range of whole 64-byte lines (<= 0: half the
 * hardware threads, at most 8, for copies of 4 MB and more, else 1).  One core streams ~10 GB/s; staging a 41 MB shard with one
 * thread takes longer than the GPU needs for the whole step. */
int         vdet_host_copy_stream_mt(void* dst, const void* src, size_t bytes, int n_threads);

/* ---------------------------------------------------------------------------------------
 * Per-frame greedy NMS on class-shared boxes.
 * Replaces: utils.cython_nms.nms (utils/nms.pyx:17-68) for S=1,C=1; the per-frame
 * suppression of vid_nms (nms.pyx:71-125); and the 30 per-class apply_vid_nms passes
 * (vdet/video_det.py:51-61) in one launch.
 *
 *   boxes        [n_rows, 4] float32, row stride `box_ld` floats (4 when packed; 6 with
 *                `boxes` pointing at column 1 of a [M,6] vid_nms array)
 *   scores       element (row r, class c) at scores[r*score_ldr + c*score_ldc]
 *   seg_offsets  [S+1] int32, ascending, seg_offsets[S] == number of packed rows
 *   row_ids      optional [n_rows] int32: original row of packed row p (NULL = identity).
 *                Boxes/scores are read at row_ids[p]; outputs report row_ids[p].
 *   thresh       IoU threshold as the reference's Python float (double); suppression test
 *                is (double)iou_f32 >= thresh  (nms.pyx:65)
 *   keep_idx     int32, C*n_rows entries: the (frame s, class c) block (see VDET_LAYOUT_*) holds
 *                the kept rows in descending score (ties: ascending row), then -1 padding
 *   keep_cnt     int32, C*S entries
 *   keep_mask    optional uint8, C*n_rows entries (1 = kept), same blocks, indexed by packed row
 *   status       [1] uint32, OR-ed with VDET_STATUS_* (never cleared by the library)
 * max_seg_len: an upper bound of the longest frame (chooses the kernel variant): <= 1024 keeps the
 * frame's bit matrix in shared memory, <= 2048 uses `ws` (vdet_nms_frames_workspace_bytes);
 * longer frames are handled by the drop-in entry points below (any length, one class).
 * Measurement hook (not part of the contract): the environment variable VDET_NMS_PER_SM caps the
 * resident CTAs per SM of the <= 1024-box variant (tools/nms_time.py).
 * ------------------------------------------------------------------------------------- */
size_t vdet_nms_frames_workspace_bytes(int max_seg_len, int n_classes, int device);
int vdet_nms_frames_f32(const float* boxes, int box_ld,
                        const float* scores, int64_t score_ldr, int64_t score_ldc,
                        const int32_t* seg_offsets, int n_segs, int max_seg_len,
                        const int32_t* row_ids, int n_classes, double thresh,
                        int32_t* keep_idx, int32_t* keep_cnt, uint8_t* keep_mask,
                        int64_t n_rows, int out_layout, uint32_t* status,
                        void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * The keep lists of vdet_nms_frames_f32 (FRAME_MAJOR layout) as ONE contiguous array: what
 * utils/nms.pyx:43-66 returns per problem -- the kept rows in descending score -- for every
 * (frame, class), in (frame, class) order, plus prefix offsets:
 *   keep_off [S*C + 1] int32 exclusive prefix of keep_cnt (device; keep_off_mirror: optional second
 *            destination, e.g. mapped pinned host memory)
 *   keep_out [sum K] entries of (frame s, class c) at keep_off[s*C + c] ...:
 *            VDET_KEEP_U16_LOCAL: uint16 index within the frame; VDET_KEEP_I32_ROW: int32 packed row
 *   keep_bits (optional) [S*C, ceil(max_seg_len/32)] uint32: bit i of block (s,c) = box i of frame s kept
 * keep_out / keep_off_mirror / keep_bits may be MAPPED PINNED HOST pointers: the kernels write them
 * across PCIe directly, which takes a data-dependent sum(K) home without a host synchronisation.
 * ------------------------------------------------------------------------------------- */
#define VDET_KEEP_U16_LOCAL 0
#define VDET_KEEP_I32_ROW   1
int vdet_compact_keep(const int32_t* keep_idx, const int32_t* keep_cnt, const int32_t* seg_offsets,
                      int n_segs, int max_seg_len, int n_classes, int out_dtype,
                      int32_t* keep_off, int32_t* keep_off_mirror, void* keep_out,
                      uint32_t* keep_bits, void* stream);

/* ---------------------------------------------------------------------------------------
 * Drop-in NMS entry points (SYNCHRONOUS: they return the kept count).
 * Replace utils.cython_nms.nms / vid_nms / track_det_nms (utils/nms.pyx:17,71,128).
 *   dets   [n, ncol] float32 row-major with row stride `ld` floats;
 *          nms: (x1,y1,x2,y2,score); vid_nms: (frame,x1,y1,x2,y2,score)
 *   keep   [n] int64: kept ORIGINAL row indices in GLOBAL descending score order
 *   returns kept count (>= 0) or VDET_ERR_*; ZeroDivision is reported via *status_host.
 * ------------------------------------------------------------------------------------- */
size_t vdet_nms_workspace_bytes(int64_t n, int device);
int64_t vdet_nms_f32(const float* dets, int64_t n, int ld, double thresh,
                     int64_t* keep, uint32_t* status_host,
                     void* ws, size_t ws_bytes, void* stream);
int64_t vdet_vid_nms_f32(const float* dets, int64_t n, int ld, double thresh,
                         int64_t* keep, uint32_t* status_host,
                         void* ws, size_t ws_bytes, void* stream);
/* tracks [q,5] = (frame,x1,y1,x2,y2), dets [k,6]; round 1 nms.pyx:163-183, round 2 :186-187 */
int64_t vdet_track_det_nms_f32(const float* tracks, int64_t q, int tracks_ld,
                               const float* dets, int64_t k, int dets_ld, double thresh,
                               int64_t* keep, uint32_t* status_host,
                               void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * One suppression step of the greedy tubelet proposal (vdet/track.py:172-183, :238-249).
 * Device-resident state between tracker calls:
 *   det_info   [m,6] float32 sorted by descending score (track.py:135-137 / :200-201)
 *   seg_offsets/row_ids : frames of det_info (row_ids ascending inside a frame), from
 *                         vdet_segment_by_frame
 *   keep       [m] uint8, updated in place
 * track_boxes [q,4] float32 with track_seg[q] = frame SEGMENT index (or -1: frame has no
 * dets).  The boxes of one call must lie on DISTINCT frames (one tracklet); successive calls are
 * sequenced by the stream, which preserves the reference's per-box order.
 * ------------------------------------------------------------------------------------- */
int vdet_track_nms_step_f32(const float* det_info, int64_t m,
                            const int32_t* seg_offsets, const int32_t* row_ids, int n_segs,
                            const float* track_boxes, const int32_t* track_seg, int q,
                            double thresh, uint8_t* keep, uint32_t* status, void* stream);

/* Group rows by the float32 frame column (bit-equal frames, -0.0 == +0.0), stable.
 * SYNCHRONOUS (returns counts through host pointers).
 *   frames       pointer to the frame value of row 0, row stride `ld` floats
 *   row_valid    optional [n] uint8: rows with 0 are dropped
 *   scores       optional (row stride scores_ld): rows inside a frame are then ordered by DESCENDING
 *                score, ties by ascending row (frame_top_detections, utils/protocol.py:341-351)
 *   row_ids_out  [n] int32 packed row -> original row (ascending inside a frame)
 *   seg_offsets_out [n+1] int32 (first *n_segs_host+1 entries valid)
 *   seg_frame_out   optional [n] float32: frame value of each segment                    */
size_t vdet_segment_workspace_bytes(int64_t n);
int vdet_segment_by_frame(const float* frames, int ld, int64_t n, const uint8_t* row_valid,
                          const void* scores, int scores_ld, int scores_dtype,
                          int32_t* row_ids_out, int32_t* seg_offsets_out, float* seg_frame_out,
                          int32_t* n_segs_host, int32_t* max_seg_len_host, int64_t* n_packed_host,
                          void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * Dense IoU matrix.  Replaces utils.common.iou (utils/common.py:451-468): the f64 entry
 * reproduces its float64 arithmetic bit for bit; the f32 entry uses the float32 pair
 * arithmetic of nms.pyx:57-64 (the HBM-roofline kernel of BASELINE.json).
 *   a [na,4], b [nb,4] packed; out [na, nb] row-major (out_ld = nb).  union==0 -> NaN.
 * ------------------------------------------------------------------------------------- */
int vdet_iou_matrix_f32(const float* a, int64_t na, const float* b, int64_t nb,
                        float* out, void* stream);
int vdet_iou_matrix_f64(const double* a, int64_t na, const double* b, int64_t nb,
                        double* out, void* stream);
/* Suppression bit matrix of ONE frame in original index space:
 * bit j of mask[i*words + j/32] = ((double)iou_f32(i,j) >= thresh); words = ceil(n/32). */
int vdet_iou_bitmask_f32(const float* boxes, int n, double thresh, uint32_t* mask,
                         uint32_t* status, void* stream);

/* ---------------------------------------------------------------------------------------
 * Frame-to-frame link (build-defined; IoU = nms.pyx pair arithmetic,
 * FIRST arg-max as np.argmax in tubelet_cls.py:375-376).
 * For every packed row p of frame s < S-1: succ[p] = packed row of the best box of frame
 * s+1 (or -1 if that frame is empty), best_iou[p] = its IoU.  Rows of the LAST frame are
 * linked against `halo_boxes` [n_halo,4] (the first frame of the next shard, from the
 * boundary allgather); succ is then halo_row_base + the index into the halo (-1 when n_halo == 0).
 * halo_row_base lets a caller split one video into two calls (frames [0,k) with frame k as the
 * halo and halo_row_base = its packed row offset) and still get packed-row successors; a shard of
 * a multi-GPU run passes its own row count, so that halo successors lie BEYOND the local rows
 * (succ >= n_rows = "continues on the next shard") instead of aliasing them.
 * n_halo_dev (optional, device): the halo's box count when only the device knows it (ragged
 * frames: it arrives with the boundary all-gather); n_halo is then the capacity of halo_boxes.
 * ws (optional, vdet_link_workspace_bytes): with a workspace and frames of 512..2048 boxes every frame
 * is sorted by x1 first and only the pairs that can overlap in x are evaluated -- about half of
 * them on BASELINE's synthetic frames, same results bit for bit; without one, every pair is.
 * ------------------------------------------------------------------------------------- */
size_t vdet_link_workspace_bytes(int64_t n_rows, int n_segs, int n_halo);
int vdet_link_frames_f32(const float* boxes, const int32_t* seg_offsets, int n_segs,
                         int max_seg_len, const float* halo_boxes, int n_halo,
                         const int32_t* n_halo_dev, int halo_row_base,
                         int32_t* succ, float* best_iou, int64_t n_rows,
                         void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * Spatial max-pooling of detections onto tubelet boxes (vdet/tubelet_cls.py:330-347,
 * :515-532; mode ARGMAX_IOU = anchor_propagate :375-377).  IoU in float64 exactly as
 * utils/common.py:451-468.
 *   tub_boxes [p,4], det_boxes [n_rows,4]: float32 or float64 (box_dtype)
 *   tub_seg   [p] int32 frame segment of each tubelet box (-1 = frame has no dets)
 *   det_scores: element r at det_scores[r*score_ld], float32 or float64 (score_dtype)
 *   out_arg   [p] int32 packed det row chosen, -1 when no det has IoU > thresh (strict)
 *   out_score [p] float64: chosen det's score, or -1e5
 * ------------------------------------------------------------------------------------- */
int vdet_spatial_maxpool(const void* tub_boxes, const int32_t* tub_seg, int64_t p,
                         const void* det_boxes, int box_dtype,
                         const void* det_scores, int64_t score_ld, int score_dtype,
                         const int32_t* det_seg_offsets, int n_segs,
                         double thresh, int mode,
                         int32_t* out_arg, double* out_score, void* stream);

/* ---------------------------------------------------------------------------------------
 * Temporal score smoothing on [n_rows, L] score rows (one row per tubelet x class), row
 * stride `ld` elements, optional per-row lengths (NULL = all L).  dtype F32 or F64; the
 * proto adapters use F64 (the reference works on Python floats).
 *   completion : do_score_completion, vdet/tubelet_cls.py:284-303 (in place).  Rows with no
 *                valid score set VDET_STATUS_ALL_MISSING.
 *   maxpool    : score_proto_temporal_maxpool, :386-414 (window odd, pad -1e5)
 *   conv1d     : depthwise temporal convolution, the build-defined stand-in for
 *                score_conv_cls :15-51 (taps [n_channels, w]; row r uses channel r % n_channels)
 * ------------------------------------------------------------------------------------- */
size_t vdet_score_completion_workspace_bytes(int64_t n_rows, int64_t L, int dtype);
int vdet_score_completion(void* scores, int dtype, int64_t n_rows, int64_t L, int64_t ld,
                          const int32_t* lengths, double miss_thr, uint32_t* status,
                          void* ws, size_t ws_bytes, void* stream);
/* The same for rows that are ONE FRAME RANGE of longer tubelets (frame-sharded completion): a run that
 * touches the shard's first / last column is completed from the nearest valid score of the neighbouring shards.
 *   bounds [n_rows, 4] (row dtype): left gap, left value, right gap, right value; gap = number of (missing)
 *   frames between that score and this shard's first / last column, < 0 = no valid score on that side.
 * A row with no valid score anywhere (both gaps < 0, nothing valid inside) sets VDET_STATUS_ALL_MISSING. */
int vdet_score_completion_bounded(void* scores, int dtype, int64_t n_rows, int64_t L, int64_t ld,
                                  const int32_t* lengths, double miss_thr, const void* bounds,
                                  uint32_t* status, void* ws, size_t ws_bytes, void* stream);
int vdet_temporal_maxpool(const void* scores, void* out, int dtype, int64_t n_rows, int64_t L,
                          int64_t ld, const int32_t* lengths, int window, double pad, void* stream);
int vdet_temporal_conv1d(const void* x, void* out, int dtype, int64_t n_rows, int64_t L,
                         int64_t ld, const int32_t* lengths, const void* taps, int n_channels,
                         int window, int pad_mode, void* stream);

/* ---------------------------------------------------------------------------------------
 * Post-CNN per-class score floor + cap (vdet/video_det.py:88-100), all frames x classes.
 *   scores [n_rows, n_classes] float32 (class 0 = background, skipped)
 *   idx_out [S, n_classes, k] int32 local row index within the frame, -1 padded: ascending
 *           row order when count <= k, else the top k in descending score
 *   cnt_out [S, n_classes] int32 = min(count, k)
 * ------------------------------------------------------------------------------------- */
int vdet_threshold_topk_f32(const float* scores, const int32_t* seg_offsets, int n_segs,
                            int max_seg_len, int n_classes, float thresh, int k,
                            int32_t* idx_out, int32_t* cnt_out, void* stream);

/* ---------------------------------------------------------------------------------------
 * Densify strided tubelets (score_proto_interpolation, vdet/tubelet_cls.py:430-490).
 * Every tubelet has >= 2 knots (frames ascending).  knot_x [n_knots] frame ids as float64, knot_y
 * [n_fields, n_knots] the fields to interpolate, knot_off [n_tubelets+1]; the dense frames of
 * tubelet k are dense_first[k] + 0,1,... occupying out columns [dense_off[k], dense_off[k+1]);
 * dense_tub [n_dense] = tubelet of each dense column.  out [n_fields, n_dense] float64.  Linear
 * interpolation as numpy.interp inside the knots, linear extrapolation (extrap1d) outside.
 * ------------------------------------------------------------------------------------- */
int vdet_tubelet_interpolate_f64(const double* knot_x, const double* knot_y, int64_t n_knots,
                                 const int32_t* knot_off, const int32_t* dense_off,
                                 const int32_t* dense_first, const int32_t* dense_tub,
                                 int n_tubelets, int n_fields, int64_t n_dense, double* out,
                                 void* stream);

/* ---------------------------------------------------------------------------------------
 * Links -> tubelet score rows (build-defined glue between vdet_link_frames_f32 and the temporal
 * kernels; the reference obtains tubelets from external trackers, vdet/track.py:18-106).
 *   follow_links: chain k starts at packed row start[k] and follows succ[]; it ends at succ == -1,
 *                 at succ >= n_rows (the link leaves this shard: halo successors, see
 *                 vdet_link_frames_f32) or when link_iou < min_iou.
 *                 chain_rows [n_frames, n_chains] int32 (-1 after the end).
 *   gather_chain_scores: out [n_chains, n_classes, n_frames] float32 = class scores along each chain,
 *                 `missing` (-1e5, utils/protocol.py:459) after its end -- the rows that
 *                 vdet_score_completion / vdet_temporal_maxpool / vdet_temporal_conv1d consume.
 * ------------------------------------------------------------------------------------- */
int vdet_follow_links(const int32_t* succ, const float* link_iou, int64_t n_rows, const int32_t* start,
                      int n_chains, int n_frames, float min_iou, int32_t* chain_rows, void* stream);
int vdet_gather_chain_scores_f32(const float* scores, int n_classes, const int32_t* chain_rows,
                                 int n_chains, int n_frames, float missing, float* out, void* stream);

/* ---------------------------------------------------------------------------------------
 * Stable sort of (score, id) pairs by DESCENDING score (equal scores keep their input order).
 * The merge step of a frame-sharded vid_nms: every rank all-gathers its kept (score, global row)
 * list and sorts the concatenation into the reference's global keep order (utils/nms.pyx:80,97);
 * also the ranking of top_detections (utils/protocol.py:330-339).  dtype F32 or F64 (Python floats).
 * ------------------------------------------------------------------------------------- */
size_t vdet_sort_workspace_bytes(int64_t n);
int vdet_sort_by_score_desc(const void* scores, int dtype, const int64_t* ids, int64_t n,
                            void* scores_out, int64_t* ids_out,
                            void* ws, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* VDET_B200_H_ */
