// Internal (non-exported) entry points shared between translation units of libfdms.
#pragma once
#include "common.cuh"

int fdms_nms_impl(fdms_ctx *ctx, const float *rows, int row_stride, int box_off, int score_off, const int32_t *tie,
                  const int32_t *count, int n_seg, int cap, double iou_thres, int iou_mode, int max_nms, int32_t *keep,
                  int32_t *n_keep, cudaStream_t st);
void fdms_fold_threshold(double thr, int iou_mode, float *thr_f, int *inclusive);
