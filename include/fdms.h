/*
 * fdms.h -- C ABI of libfdms.so: the B200-native (sm_100a) multi-scale face-detection
 * pre/post-processing hot path.
 *
 * The reference (NguyendaiAI-pixta/face_detection_multi_scale) has no FFI of its own: its
 * boundary is a set of Python callables.  Each entry point below names the
 * reference callable (file:line, relative to the reference root) whose arithmetic it
 * replaces; the modules of face_detection_multi_scale_b200/ bind them with ctypes and re-exports the
 * reference's Python signatures on top (INTEGRATION.md shows the binding).
 *
 * Conventions
 *   - every function returns FDMS_OK (0) or a negative fdms_status; none throws;
 *   - all data pointers are DEVICE pointers unless the name says host;
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream);
 *   - no entry point synchronises the stream or allocates on the hot path: scratch lives in
 *     the ctx and only grows (outside steady state) when a call needs more than any before;
 *   - a ctx is bound to one device and must be used by one host thread / one stream at a
 *     time (its scratch is shared by consecutive calls); create one ctx per worker thread;
 *   - there is no CPU fallback: without a CUDA device fdms_create fails with FDMS_ERR_CUDA.
 */
#ifndef FDMS_H_
#define FDMS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FDMS_API __attribute__((visibility("default")))
#define FDMS_VERSION 100 /* 0.1.0 */

typedef struct fdms_ctx fdms_ctx;
typedef void *fdms_stream;

typedef enum fdms_status {
    FDMS_OK = 0,
    FDMS_ERR_INVALID = -1,     /* bad argument (message in fdms_last_error) */
    FDMS_ERR_CUDA = -2,        /* CUDA runtime error (no device, launch failure, ...) */
    FDMS_ERR_NOMEM = -3,       /* scratch allocation failed */
    FDMS_ERR_UNSUPPORTED = -4  /* shape outside compiled limits */
} fdms_status;

typedef enum fdms_dtype { FDMS_U8 = 0, FDMS_F16 = 1, FDMS_F32 = 2 } fdms_dtype;
typedef enum fdms_layout { FDMS_CHW = 0, FDMS_HWC = 1 } fdms_layout;
/* torchvision.ops.nms arithmetic flavours */
typedef enum fdms_iou_mode {
    FDMS_IOU_TV_CPU = 0,  /* torchvision CPU kernel: (area_i + area_j) - inter, (double)iou > thr */
    FDMS_IOU_TV_CUDA = 1, /* torchvision CUDA kernel: fma(w_j, h_j, area_i) - inter, iou > (float)thr */
    FDMS_IOU_NCNN = 2     /* cpp/yolov7-face-ncnn/src/yolov7face.cpp:43-74: +1 pixel extents, iou >= (float)thr */
} fdms_iou_mode;

#define FDMS_MAX_SCALES 8
#define FDMS_MAX_LEVELS 32
#define FDMS_MAX_ANCHORS 8   /* anchors per level of fdms_decode_dense_head */
#define FDMS_MAX_HEAD_NO 383 /* columns per row of fdms_decode_dense_head (5 + nc, or 6 + 3 * nkpt) */
#define FDMS_MAX_NMS 30000 /* utils/general.py:520 */

/* ---- context ---------------------------------------------------------------------------- */
FDMS_API int fdms_version(void);
FDMS_API const char *fdms_build_info(void);
FDMS_API int fdms_create(fdms_ctx **out, int device);
FDMS_API void fdms_destroy(fdms_ctx *ctx);
/* message of the last failing call made with this ctx (empty string if none) */
FDMS_API const char *fdms_last_error(const fdms_ctx *ctx);
/* Test hook (no reference counterpart): overwrite the shared memory of every SM with `byte_value`, so that a later
 * kernel that reads shared memory it never wrote fails reproducibly instead of depending on its predecessor. */
FDMS_API int fdms_debug_poison_smem(fdms_ctx *ctx, int byte_value, fdms_stream stream);
/* bytes of device scratch currently owned by the ctx */
FDMS_API int64_t fdms_scratch_bytes(const fdms_ctx *ctx);

/* ---- letterbox geometry (host arithmetic only) -------------------------------------------
 * utils/datasets.py:873-903 `letterbox`: ratio, new_unpad, dw/dh (mod stride when auto),
 * Python round-half-even on the halved paddings. */
typedef struct fdms_geom {
    int32_t dst_h, dst_w;     /* output image size */
    int32_t unpad_h, unpad_w; /* resized size (new_unpad) */
    int32_t top, bottom, left, right;
    int32_t resize;           /* 1 when the reference would call cv2.resize (:898) */
    int32_t pad_;
    double ratio_w, ratio_h;  /* returned `ratio` */
    double dw, dh;            /* returned `(dw, dh)` (halved paddings) */
} fdms_geom;

FDMS_API int fdms_letterbox_geometry(int src_h, int src_w, int new_h, int new_w, int auto_, int scale_fill,
                                     int scaleup, int stride, fdms_geom *out);

/* ---- fused pyramid: resize + pad + channel flip + /255 + pack ----------------------------
 * Replaces, for every image of a batch and every scale in ONE launch:
 *   utils/datasets.py:873-903 letterbox  (cv2.resize INTER_LINEAR u8 fixed point, constant border)
 *   multi_scale_face_detector.py:96-105  (BGR->RGB, HWC->CHW, .half()/.float(), /255)
 *   utils/preprocess_yolo_predict.py:273-290,345-378 (api mode: virtual zero-padded square)
 * src: n_img images, u8 HWC, image i at src + i*img_stride + y*row_pitch + x*3 (bytes).
 * virt_h/virt_w >= src_h/src_w: size of the virtual zero-padded source the resize runs over
 * (== src size for the letterbox mode; max(h,w) square for the api mode).
 * Per scale the caller gives the geometry (from fdms_letterbox_geometry on the virtual size)
 * and the output tensor: element (img, c, y, x) at out + img*out_img_stride + c*cs + y*ys + x*xs
 * with (cs, ys, xs) = (H*W, W, 1) for FDMS_CHW or (1, 3W, 3) for FDMS_HWC (element units).
 * out_dtype: FDMS_F16 / FDMS_F32 store value/255 (true division rounded once to the type),
 * FDMS_U8 stores the pixel.  flip_rb != 0 reverses the channel order (img[:, :, ::-1]). */
typedef struct fdms_scale_out {
    fdms_geom geom;
    void *out;
    int64_t out_img_stride; /* elements */
    int32_t out_dtype;      /* fdms_dtype */
    int32_t out_layout;     /* fdms_layout */
} fdms_scale_out;

FDMS_API int fdms_pyramid(fdms_ctx *ctx, const uint8_t *src, int n_img, int64_t img_stride, int64_t row_pitch,
                          int src_h, int src_w, int virt_h, int virt_w, const fdms_scale_out *scales, int n_scales,
                          int flip_rb, int border, fdms_stream stream);

/* ---- head decode -------------------------------------------------------------------------
 * models/yolo.py:267-308 (IKeypoint inference branch; same math as Detect :59-106 and IDetect
 * :152-207), nc=1, nkpt=5 => no=21, na=3.
 * A level is one raw conv output (bs, na*no, ny, nx), channel = a*no + k, dtype f16 or f32
 * (the math is fp32 on the up-cast values).  `scale` groups levels into pyramid scales;
 * row = row_offset + a*ny*nx + gy*nx + gx addresses the reference's pred[b, row, :]. */
typedef struct fdms_level {
    const void *data;
    int32_t dtype; /* FDMS_F16 | FDMS_F32 */
    int32_t ny, nx;
    int32_t scale;      /* which pyramid scale this level belongs to, 0..n_scales-1 */
    int32_t row_offset; /* first row of this level inside its scale's pred */
    float stride;       /* 8 / 16 / 32 / 64 */
    float anchors[6];   /* (w,h) x 3 anchors, PIXELS (IKeypoint.anchor_grid) */
} fdms_level;

/* Full decode ("compat level"): writes pred (bs, rows_total, 21) f32 for one scale's levels
 * (all levels must have scale == 0).  models/yolo.py:308 `cat(z, 1)`. */
FDMS_API int fdms_decode_dense(fdms_ctx *ctx, const fdms_level *levels, int n_levels, int bs, float *pred,
                               int64_t rows_total, fdms_stream stream);

/* The same full decode for ANY detection head the reference can run in eval mode -- Detect (models/yolo.py:59-106),
 * IDetect (:152-207), IKeypoint (:267-308) with na anchors per level and no columns per row:
 *   nkpt > 0  : the keypoint heads (inplace branch; the reference slices `[..., :6]`, i.e. nc = 1, so no must equal
 *               6 + 3 * nkpt): columns 0..5 as in fdms_decode_dense, then (x, y, conf) per keypoint;
 *   nkpt == 0 : every column through the sigmoid, xy / wh decoded, obj and the nc = no - 5 class columns as they are (the
 *               plain YOLO head: `inplace=False` branch :95-99 for nc > 1; nc = 1 also through the inplace branch).
 * levels[l].anchors is ignored; anchors_px is a HOST array (n_levels, na, 2) in pixels (`anchor_grid`).  All levels
 * must have scale == 0.  pred: (bs, rows_total, no) f32, rows level-major then anchor, y, x (:308 `cat(z, 1)`). */
FDMS_API int fdms_decode_dense_head(fdms_ctx *ctx, const fdms_level *levels, int n_levels, int bs, int na, int no, int nkpt,
                                    const float *anchors_px, float *pred, int64_t rows_total, fdms_stream stream);

/* Fused decode + confidence filter + xywh2xyxy + stream compaction ("fused level"):
 * models/yolo.py:267-308 then utils/general.py:515,531,547-563 and :299-306.
 * Segment g = b*n_scales + scale.  For every row with rn(cls*obj) > (float)conf_thres one
 * candidate [x1,y1,x2,y2,conf,0,15 kpt] (21 floats) is appended to cand[g] (unordered: the
 * NMS sort key restores the reference order) and its row index to cand_row[g].
 * count[g] receives the TOTAL number of passing rows, which may exceed cap (only the first
 * cap arrivals are stored); callers treat count > cap as overflow. */
FDMS_API int fdms_decode_filter(fdms_ctx *ctx, const fdms_level *levels, int n_levels, int n_scales, int bs,
                                float conf_thres, float *cand, int32_t *cand_row, int32_t *count, int cap,
                                fdms_stream stream);

/* fdms_decode_filter with options.  FDMS_DECODE_BOXES_ONLY: the 15 landmark columns of a candidate row are neither
 * read from the head nor written (columns 0..5 are identical to the full decode).  This is what the multi-scale
 * path needs -- multi_scale_face_detector.py:151-153 drops the landmarks before the merge -- and it cuts the DRAM
 * sectors touched per candidate from 21 to 6 (the head is plane-major). */
#define FDMS_DECODE_BOXES_ONLY 1
FDMS_API int fdms_decode_filter_ex(fdms_ctx *ctx, const fdms_level *levels, int n_levels, int n_scales, int bs,
                                   float conf_thres, float *cand, int32_t *cand_row, int32_t *count, int cap, int flags,
                                   fdms_stream stream);

/* Same filter + box conversion, reading an already decoded pred (bs, rows, no) f16/f32 with
 * no = 6 + 3*kpt_label columns (nc == 1): the front half of utils/general.py:506-600
 * non_max_suppression for kpt_label = 5 (no = 21, every cfg of the reference), kpt_label = 0
 * (no = 6, the `pred[..., :6]` fallback at multi_scale_face_detector.py:134) or any other keypoint
 * count.  Candidate rows have `no` floats.  Segment g = b. */
FDMS_API int fdms_filter_pred(fdms_ctx *ctx, const void *pred, int dtype, int bs, int64_t rows, int no,
                              float conf_thres, float *cand, int32_t *cand_row, int32_t *count, int cap,
                              fdms_stream stream);

/* ---- batched NMS -------------------------------------------------------------------------
 * torchvision.ops.nms as called at utils/general.py:584 and multi_scale_face_detector.py:234:
 * stable descending sort on score (ties -> ascending `tie`), greedy suppression with strict
 * IoU > thr, no +1, NaN never suppresses (iou_mode FDMS_IOU_TV_CPU / FDMS_IOU_TV_CUDA; FDMS_IOU_NCNN
 * is the ncnn demo's flavour: +1 pixel extents, IoU >= thr).  n_seg independent segments in one call.
 * Segment g reads n = min(count[g], cap) items: box (x1,y1,x2,y2) at
 * rows + (g*cap + i)*row_stride + box_off, score at ... + score_off, tie[g*cap + i]
 * (tie == NULL means tie = i; tie values must be unique inside a segment).
 * If n > max_nms only the max_nms best-ranked are considered
 * (utils/general.py:578-579).  keep[g*cap + k] = item index i of the k-th kept box in score
 * order; n_keep[g] = number kept. */
FDMS_API int fdms_nms(fdms_ctx *ctx, const float *rows, int row_stride, int box_off, int score_off,
                      const int32_t *tie, const int32_t *count, int n_seg, int cap, double iou_thres, int iou_mode,
                      int max_nms, int32_t *keep, int32_t *n_keep, fdms_stream stream);

/* out[g, k, :] = rows[g, keep[g,k], 0:n_cols] for k < n_keep[g]  (utils/general.py:595 x[i]) */
FDMS_API int fdms_gather_rows(fdms_ctx *ctx, const float *rows, int row_stride, int n_cols, const int32_t *keep,
                              const int32_t *n_keep, int n_seg, int cap, float *out, int out_cap,
                              fdms_stream stream);

/* ---- rescale -----------------------------------------------------------------------------
 * utils/general.py:374-404 scale_coords (+clip_coords) with ratio_pad=None, in place on
 * coords (n, cols) f32 with row stride `row_stride` floats.  kpt_step == 0: box path on
 * columns 0..3 (the reference clips ROWS 0..3 only, :389); kpt_step == 3: keypoint path over
 * `cols` columns, every row clipped. */
FDMS_API int fdms_scale_coords(fdms_ctx *ctx, float *coords, int64_t n, int cols, int64_t row_stride, int in_h,
                               int in_w, int orig_h, int orig_w, int kpt_step, fdms_stream stream);
/* The same with an explicit ratio_pad = ((gain, gain), (pad_x, pad_y)) (general.py:379-383; test.py:172 passes what
 * the dataset's letterbox returned): gain and pad are the caller's Python floats, the original size only clips. */
FDMS_API int fdms_scale_coords_ratio_pad(fdms_ctx *ctx, float *coords, int64_t n, int cols, int64_t row_stride,
                                         double gain, double pad_x, double pad_y, int orig_h, int orig_w,
                                         int kpt_step, fdms_stream stream);
/* utils/preprocess_yolo_predict.py:122-157 scale_coords_api_approach on (n,4) boxes;
 * do_round applies torch.round (half to even) as multi_scale_face_detector.py:144 does. */
FDMS_API int fdms_scale_coords_api(fdms_ctx *ctx, float *boxes, int64_t n, int64_t row_stride, int in_h,
                                   int orig_h, int orig_w, int do_round, fdms_stream stream);

/* ---- cross-scale rescale + merge -----------------------------------------------------------
 * multi_scale_face_detector.py:138-166 (rescale, round, drop landmarks, append scale_idx),
 * :168-201 calculate_scale_weights, :203-240 weighted_nms, :242-288 detect_multi_scale.
 * Input: per-scale NMS survivors of n_img images (segment g = img*n_scales + s) given as
 * cand/keep/n_keep from fdms_decode_filter + fdms_nms.  in_h[s] is the model-input height of
 * scale s; every image has original size (orig_h, orig_w).  The rescale is ALWAYS the api
 * transform, as in the reference (:144).
 * Output: dets (n_img, out_cap, 7) f32 [x1,y1,x2,y2,conf,cls,scale_idx] in weighted-score
 * order and n_out[img]; merged_in (n_img) optionally receives sum_s n_keep (may be NULL).  At most out_cap
 * per-scale survivors of an image enter the merge: callers treat merged_in > out_cap as overflow.
 * If kpts_out != NULL it receives (n_img, out_cap, 15) landmarks of the kept detections
 * rescaled with the same transform (extension: the reference drops them, :151-153). */
FDMS_API int fdms_merge_scales(fdms_ctx *ctx, const float *cand, const int32_t *keep, const int32_t *n_keep,
                               int n_img, int n_scales, int cap, const int32_t *in_h_host, int orig_h, int orig_w,
                               double iou_thres, float *dets, int32_t *n_out, int out_cap, int32_t *merged_in,
                               float *kpts_out, fdms_stream stream);
/* The same for a RAGGED batch: images of different original sizes whose letterboxed inputs share their shapes (the
 * reference handles arbitrary images per call, multi_scale_face_detector.py:242-288; autoShape batches mixed sizes,
 * models/common.py:618-636).  orig_hw_dev = (n_img, 2) int32 on the device, (H0, W0) per image. */
FDMS_API int fdms_merge_scales_ragged(fdms_ctx *ctx, const float *cand, const int32_t *keep, const int32_t *n_keep,
                                      int n_img, int n_scales, int cap, const int32_t *in_h_host,
                                      const int32_t *orig_hw_dev, double iou_thres, float *dets, int32_t *n_out,
                                      int out_cap, int32_t *merged_in, float *kpts_out, fdms_stream stream);

/* multi_scale_face_detector.py:168-201,224-229 for the standalone weighted_nms(detections) call:
 * dets (n,7) float64 [x1,y1,x2,y2,conf,cls,scale_idx] -> rows_out (n,8) f32
 * [x1,y1,x2,y2, (float)(conf*weight), conf, cls, scale_idx]; feed rows_out to fdms_nms
 * (row_stride 8, box_off 0, score_off 4, tie NULL, TV_CPU) to get the keep list of :234.
 * weights_out (n) f64, optional: the weights calculate_scale_weights returns. */
FDMS_API int fdms_weighted_rows(fdms_ctx *ctx, const double *dets, int64_t n, int n_scales, float *rows_out,
                                double *weights_out, fdms_stream stream);

/* ---- whole post-processing step in one host call ---------------------------------------------
 * decode_filter -> per-scale NMS -> merge, back to back on `stream` without host sync.
 * cand/cand_row/count/keep/n_keep are caller-owned scratch with the shapes given above
 * (n_seg = bs*n_scales).  Counts can be read back afterwards for overflow checks. */
/* flags of fdms_post_desc.  FDMS_POST_SMALL_ONLY: launch only the single-CTA kernels (segments / images of at most
 * 1024 boxes -- the common case).  A segment with more candidates gets n_keep = -2 and an image with more merged rows
 * n_out = 0: the caller MUST compare count[] and merged_in[] with FDMS_SMALL_LIMIT afterwards and redo the call
 * without the flag when either exceeds it (the counts are exact).  Saves the launches that a capacity above 1024 would
 * otherwise add to every call although they find nothing to do. */
#define FDMS_POST_SMALL_ONLY 1
#define FDMS_SMALL_LIMIT 1024
typedef struct fdms_post_desc {
    const fdms_level *levels; /* host array */
    int32_t n_levels, n_scales, bs, cap;
    float conf_thres;
    int32_t iou_mode; /* per-scale NMS flavour; the merge NMS is always TV_CPU (:234 runs on CPU) */
    double iou_thres;
    const int32_t *in_h_host; /* [n_scales] */
    int32_t orig_h, orig_w;
    float *cand; int32_t *cand_row; int32_t *count; int32_t *keep; int32_t *n_keep;
    float *dets; int32_t *n_out; int32_t out_cap;
    int32_t flags; /* FDMS_POST_* */
    int32_t *merged_in; /* may be NULL */
    float *kpts_out;    /* may be NULL; (bs, out_cap, 15) landmarks of the final detections.  cand[..., 6:21] is NOT filled by
                         * fdms_postprocess in either case: the landmarks are decoded from the head maps for the final
                         * detections only (the standalone fdms_decode_filter fills all 21 columns) */
    const int32_t *orig_hw_dev; /* NULL, or (bs, 2) int32 ON THE DEVICE: (H0, W0) per image -- a batch whose images have
                                 * different original sizes but the same letterboxed input shapes (orig_h / orig_w unused) */
} fdms_post_desc;

FDMS_API int fdms_postprocess(fdms_ctx *ctx, const fdms_post_desc *desc, fdms_stream stream);
/* Floating-point flavour of the stages whose reference arithmetic depends on the device the reference runs on:
 *   FDMS_NUMERICS_EXACT (default)  sigmoid with the correctly rounded exponential, `x / s` as a true division -- what the
 *       CPU oracle restates and within 1 ulp of ATen on either device;
 *   FDMS_NUMERICS_ATEN_CUDA        what ATen computes ON A CUDA DEVICE: sigmoid = 1 / (1 + expf(-x)) with
 *       libdevice's expf (models/yolo.py:285-306), and tensor / python-scalar = tensor * (1 / scalar)
 *       (utils/general.py:386 `coords[:, :4] /= gain`, multi_scale_face_detector.py:102 `img /= 255.0` in fp32,
 *       models/yolo.py:405 `p[..., :4] /= scale`), and fdms_scale_img with the FMA contraction of ATen's CUDA bilinear
 *       kernel (utils/torch_utils.py:254 F.interpolate): bit-identical to the reference running on the same GPU --
 *       in fp32 for f32 head maps, and for F16 head maps (a half model, the reference's default on a GPU, msfd:49) with
 *       the roundings of ATen's half tensors: sigmoid, `* 2.`, `- 0.5`, `** 2`, `* anchor_grid` round to half, `+ grid`
 *       and `* stride` are float32 (the grid is created float32, models/yolo.py:310-313), the keypoint coordinates are
 *       stored back into the half tensor; pred is float32 in the reference too (torch.cat promotes), and everything
 *       after the head is float32 arithmetic in both.
 * Also read from the environment variable FDMS_NUMERICS (exact | aten_cuda) when a context is created. */
#define FDMS_NUMERICS_EXACT 0
#define FDMS_NUMERICS_ATEN_CUDA 1
FDMS_API int fdms_set_numerics(fdms_ctx *ctx, int mode);
FDMS_API int fdms_get_numerics(const fdms_ctx *ctx);

/* The same contract as FDMS_POST_SMALL_ONLY for fdms_nms and fdms_merge_scales called one by one on this context
 * (until switched off again): only the single-CTA kernels are launched, the caller checks the counts. */
FDMS_API int fdms_set_small_only(fdms_ctx *ctx, int on);

/* ---- "next" rows around the hot path ---------------------------------------------
 * widerface_evaluate/box_overlaps.pyx:15-55 bbox_overlaps: (n,4) x (k,4) float64 boxes -> (n,k) float64
 * IoU with the +1 pixel convention (0 where the boxes do not overlap). */
FDMS_API int fdms_bbox_overlaps(fdms_ctx *ctx, const double *boxes, int64_t n, const double *query, int64_t k,
                                double *overlaps, fdms_stream stream);

/* Padded per-frame output tensors of the production driver: normalised boxes x/img_w, y/img_h in float64
 * (utils/preprocess_yolo_predict.py:26-49, yolov7_face_multi_scale_dataframe_predict.py:699-716) padded with
 * -1 to max_faces per frame (utils/preprocess_yolo_predict.py:478-520).  dets/n_out as produced by
 * fdms_merge_scales.  bboxes (n_img, max_faces, 4) f64, confidence (n_img, max_faces) f64,
 * class_indexes / scale_used (n_img, max_faces) i32. */
FDMS_API int fdms_pack_frames(fdms_ctx *ctx, const float *dets, const int32_t *n_out, int n_img, int out_cap,
                              int max_faces, int img_w, int img_h, double *bboxes, double *confidence,
                              int32_t *class_indexes, int32_t *scale_used, fdms_stream stream);

/* Test-time augmentation around the (unchanged) backbone, models/yolo.py:363-374 Model.forward_augment:
 * fdms_scale_img = utils/torch_utils.py:247-257 scale_img: F.interpolate(bilinear, align_corners=False) of an NCHW
 * f16/f32 batch (n, c, h, w) to (resized_h, resized_w) = (int(h*ratio), int(w*ratio)), read through an optional flip
 * of the source (0 none, 2 up-down, 3 left-right: `x.flip(fi)` at :368), and padded with pad_value (0.447) on the
 * right / bottom to (out_h, out_w), the next multiple of the grid size.  dst is (n, c, out_h, out_w), same dtype. */
FDMS_API int fdms_scale_img(fdms_ctx *ctx, const void *src, int dtype, int n, int c, int h, int w, int resized_h,
                            int resized_w, int out_h, int out_w, int flip, float pad_value, void *dst,
                            fdms_stream stream);

/* models/yolo.py:402-417 _descale_pred on pred (bs, rows, no) f32: columns 0..3 /= scale, then y = img_h - y
 * (flip 2) or x = img_w - x (flip 3); the other columns are copied (the reference leaves the landmarks as they
 * are).  The rows are written at row_offset of dst (bs, rows_total, no): the torch.cat(y, 1) of :374.  dst may be
 * pred itself when rows_total == rows and row_offset == 0. */
FDMS_API int fdms_descale_pred(fdms_ctx *ctx, const float *pred, int bs, int64_t rows, int no, float scale, int flip,
                               int img_h, int img_w, float *dst, int64_t rows_total, int64_t row_offset,
                               fdms_stream stream);

/* ---- device JPEG ingest: multi_scale_face_detector.py:84 cv2.imread -> libjpeg-turbo -------------
 * The entropy decode stays on the host (one image per thread), dequantisation + islow IDCT + fancy chroma upsampling +
 * YCbCr -> BGR run on the device, bit-identical to cv2.imread.  Baseline / extended sequential 8-bit JPEG, one interleaved
 * scan, grey or YCbCr with 4:4:4 / 4:2:2 / 4:2:0 sampling, no EXIF rotation; anything else is reported as unsupported and the
 * caller decodes that file with the host decoder. */
typedef struct fdms_jpeg_desc {
    int32_t width, height, ncomp, supported;
    int32_t h[3], v[3];                    /* sampling factors per component */
    int32_t blocks_w[3], blocks_h[3];      /* whole 8x8 blocks per component (MCU-padded) */
    int32_t down_w[3], down_h[3];          /* real samples per component */
    int64_t coef_off[3], coef_count;       /* layout of the int16 coefficient buffer: (blocks_h, blocks_w, 64) per component */
    uint16_t qt[3][64];                    /* quantisation table per component, natural (row-major) order */
    char reason[64];                       /* why `supported` is 0 */
} fdms_jpeg_desc;
FDMS_API int fdms_jpeg_probe(const uint8_t *data, size_t len, fdms_jpeg_desc *out);      /* host only */
FDMS_API int fdms_jpeg_decode_host(const uint8_t *data, size_t len, int16_t *coef);      /* host only: entropy decode */
/* Decodes n JPEG byte streams (host pointers) of one size h x w into dst = (n, h, w, 3) uint8 BGR ON THE DEVICE (row pitch /
 * image stride in bytes): Huffman decode on n_threads host threads (0 = all), one upload of the coefficients, two kernels on
 * `stream`.  status[i] (host): 0 = decoded on the device; 1 = not supported by the device path (progressive, 4:1:1, EXIF
 * rotation, ...): slot i of dst is left untouched and the caller decodes that file on the host; < 0 = corrupt stream or
 * wrong size.  Returns after the host work; the device work is asynchronous. */
FDMS_API int fdms_jpeg_decode_batch(fdms_ctx *ctx, const uint8_t *const *data, const size_t *len, int n, int n_threads,
                                    uint8_t *dst, int64_t img_stride, int64_t pitch, int h, int w, int32_t *status,
                                    fdms_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* FDMS_H_ */
