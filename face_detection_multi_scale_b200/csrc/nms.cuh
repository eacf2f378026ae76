// Internal (non-exported) entry points shared between translation units of libfdms.
#pragma once
#include "common.cuh"

int fdms_nms_impl(fdms_ctx *ctx, const float *rows, int row_stride, int box_off, int score_off, const int32_t *tie,
                  const int32_t *count, int n_seg, int cap, double iou_thres, int iou_mode, int max_nms, int32_t *keep,
                  int32_t *n_keep, cudaStream_t st);
void fdms_fold_threshold(double thr, int iou_mode, float *thr_f, int *inclusive);
// The fused single-CTA merge (merge rows -> NMS -> emit) for images of at most fdms_merge_small_limit() rows; larger
// images only get their rows written.  Callers then run fdms_nms_impl with ctx->nms_small_done set (the segments of
// at most that many rows are done) and merge_emit_kernel for the rest.
struct MergeParams;
struct EmitParams;
int fdms_merge_small_limit(const fdms_ctx *ctx);
int fdms_merge_small_launch(fdms_ctx *ctx, const MergeParams &P, const EmitParams &E, int n_img, float thr_f, int inclusive,
                            cudaStream_t st);
