// Rescale to original-image coordinates and cross-scale merge.
//
//   scale_coords_kernel / scale_coords_api_kernel : in-place drop-ins for
//       utils/general.py:374-404 (ratio_pad None or given) and utils/preprocess_yolo_predict.py:122-157.
//   The merge of an image (merge_rows.cuh holds the per-image device code):
//     rows : multi_scale_face_detector.py:138-166 -- api rescale, clamp, round half-to-even, drop landmarks, append
//            scale_idx -- and :168-201,224-229 (size-aware weights, weighted score in float64 then cast to
//            float32), the scales of an image concatenated in scale order (np.vstack, :273);
//     NMS  : TV_CPU flavour (the reference runs it on CPU, :234);
//     emit : detections[keep] with the ORIGINAL confidence (:237-238).
//   Images of at most 1024 merged rows do all three in one CTA (merge_small_kernel, nms.cu); larger ones use
//   merge_prepare_kernel (here, when the fused kernel is off) / the rows written by the fused kernel, the
//   large-segment NMS kernels and merge_emit_kernel.
//   fdms_postprocess : decode_filter -> per-scale NMS -> merge as one host call, no host sync.
#include <math.h>

#include "common.cuh"
#include "merge_rows.cuh"
#include "nms.cuh"

namespace {

constexpr int NO = MERGE_NO;

// utils/general.py:374-404 (ratio_pad None or given).  gain/pad are computed in double on the host like
// the Python floats they are, then cast to the tensor dtype by the in-place ops.
// inv_gain != 0 (FDMS_NUMERICS_ATEN_CUDA): `coords /= gain` as ATen's CUDA kernel computes a division by a Python
// scalar, coords * (1 / gain) with the reciprocal rounded to fp32 first; else the true division of the CPU kernel.
__global__ void scale_coords_kernel(float *coords, long long n, int cols, long long row_stride, float padx, float pady,
                                    float gain, float inv_gain, float w0, float h0, int kpt_step) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float *r = coords + i * row_stride;
    auto f_div = [&](float a, float g) { return inv_gain != 0.f ? ::f_mul(a, inv_gain) : ::f_div(a, g); };
    if (kpt_step == 0) {
        float x1 = f_div(f_sub(r[0], padx), gain), y1 = f_div(f_sub(r[1], pady), gain);
        float x2 = f_div(f_sub(r[2], padx), gain), y2 = f_div(f_sub(r[3], pady), gain);
        if (i < 4) {  // clip_coords(coords[0:4], ...) slices ROWS (general.py:389)
            x1 = clampf(x1, 0.f, w0);
            x2 = clampf(x2, 0.f, w0);
            y1 = clampf(y1, 0.f, h0);
            y2 = clampf(y2, 0.f, h0);
        }
        r[0] = x1; r[1] = y1; r[2] = x2; r[3] = y2;
    } else {
        for (int c = 0; c < cols; c += kpt_step) {
            r[c] = clampf(f_div(f_sub(r[c], padx), gain), 0.f, w0);
            if (c + 1 < cols) r[c + 1] = clampf(f_div(f_sub(r[c + 1], pady), gain), 0.f, h0);
        }
    }
}

__global__ void scale_coords_api_kernel(float *boxes, long long n, long long row_stride, float k, float w0, float h0,
                                        int do_round) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float *r = boxes + i * row_stride;
    float4 b = api_box(make_float4(r[0], r[1], r[2], r[3]), k, w0, h0, do_round != 0);
    r[0] = b.x; r[1] = b.y; r[2] = b.z; r[3] = b.w;
}

__global__ void merge_prepare_kernel(const __grid_constant__ MergeParams P) {
    merge_prepare_image(P, blockIdx.x, threadIdx.x, blockDim.x);
}

// msfd:168-201 + 224-229 on an (n,7) float64 detection array (the weighted_nms drop-in):
// rows_out (n,8) f32 = [x1,y1,x2,y2 as float32, weighted score, conf, cls, scale_idx]
__global__ void weighted_rows_kernel(const double *__restrict__ dets, long long n, int n_scales,
                                     float *__restrict__ rows_out, double *__restrict__ weights_out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *d = dets + i * 7;
    double area = __dmul_rn(__dsub_rn(d[2], d[0]), __dsub_rn(d[3], d[1]));
    int s = (int)d[6];
    double wt = 1.0;
    if (area < 1024.0 && s >= n_scales - 2) wt *= 1.2;
    if (area >= 1024.0 && area <= 16384.0 && s == n_scales / 2) wt *= 1.1;
    if (area > 16384.0 && s <= 1) wt *= 1.2;
    float *o = rows_out + i * MROW;
    o[0] = (float)d[0]; o[1] = (float)d[1]; o[2] = (float)d[2]; o[3] = (float)d[3];
    o[4] = (float)__dmul_rn(d[4], wt);
    o[5] = (float)d[4]; o[6] = (float)d[5]; o[7] = (float)d[6];
    if (weights_out) weights_out[i] = wt;
}

// skip_le >= 0: images with at most that many merged rows were emitted by the fused kernel already
// grid (parts, images): a dense image's rows (thousands, each gathering its landmarks from scattered candidate rows)
// are spread over several CTAs -- one CTA per image left the kernel waiting on ~40 dependent loads per thread
// (ncu: 96 % long-scoreboard stalls, 111 us for 32 CTAs on the dense-crowd workload).
__global__ void merge_emit_kernel(const __grid_constant__ EmitParams P, const int32_t *__restrict__ mcount, int skip_le) {
    const int img = blockIdx.y;
    if (skip_le >= 0 && mcount[img] <= skip_le) return;
    merge_emit_image(P, img, P.n_keep2[img], threadIdx.x, blockDim.x, blockIdx.x, gridDim.x);
}

}  // namespace

extern "C" {

FDMS_API int fdms_scale_coords(fdms_ctx *ctx, float *coords, int64_t n, int cols, int64_t row_stride, int in_h,
                               int in_w, int orig_h, int orig_w, int kpt_step, fdms_stream stream) {
    if (in_h <= 0 || in_w <= 0 || orig_h <= 0 || orig_w <= 0) {
        FDMS_BIND(ctx);
        return n == 0 ? FDMS_OK : fdms_fail(ctx, FDMS_ERR_INVALID, "fdms_scale_coords: bad arguments");
    }
    double gain = fmin((double)in_h / orig_h, (double)in_w / orig_w);  // general.py:377
    double padx = ((double)in_w - orig_w * gain) / 2, pady = ((double)in_h - orig_h * gain) / 2;  // :378
    return fdms_scale_coords_ratio_pad(ctx, coords, n, cols, row_stride, gain, padx, pady, orig_h, orig_w, kpt_step,
                                       stream);
}

FDMS_API int fdms_scale_coords_ratio_pad(fdms_ctx *ctx, float *coords, int64_t n, int cols, int64_t row_stride,
                                         double gain, double pad_x, double pad_y, int orig_h, int orig_w,
                                         int kpt_step, fdms_stream stream) {
    FDMS_BIND(ctx);
    if (n == 0) return FDMS_OK;
    if (!coords || n < 0 || cols <= 0 || row_stride < cols || orig_h <= 0 || orig_w <= 0 ||
        (kpt_step != 0 && kpt_step < 2) || (kpt_step == 0 && cols < 4))
        return fdms_fail(ctx, FDMS_ERR_INVALID, "fdms_scale_coords: bad arguments");
    // the Python floats of general.py:377-381 meet an fp32 tensor: each in-place op rounds its scalar to fp32 first
    scale_coords_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(
        coords, n, cols, row_stride, (float)pad_x, (float)pad_y, (float)gain,
        ctx->numerics == FDMS_NUMERICS_ATEN_CUDA ? 1.0f / (float)gain : 0.f, (float)orig_w, (float)orig_h, kpt_step);
    FDMS_LAUNCH_CHECK(ctx, "scale_coords_kernel");
    return FDMS_OK;
}

FDMS_API int fdms_scale_coords_api(fdms_ctx *ctx, float *boxes, int64_t n, int64_t row_stride, int in_h, int orig_h,
                                   int orig_w, int do_round, fdms_stream stream) {
    FDMS_BIND(ctx);
    if (n == 0) return FDMS_OK;
    if (!boxes || n < 0 || row_stride < 4 || in_h <= 0 || orig_h <= 0 || orig_w <= 0)
        return fdms_fail(ctx, FDMS_ERR_INVALID, "fdms_scale_coords_api: bad arguments");
    int sq = orig_h > orig_w ? orig_h : orig_w;     // ppyp:141
    float k = (float)((double)sq / (double)in_h);   // ppyp:142-145 (Python float times fp32 tensor)
    scale_coords_api_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(
        boxes, n, row_stride, k, (float)orig_w, (float)orig_h, do_round);
    FDMS_LAUNCH_CHECK(ctx, "scale_coords_api_kernel");
    return FDMS_OK;
}

FDMS_API int fdms_weighted_rows(fdms_ctx *ctx, const double *dets, int64_t n, int n_scales, float *rows_out,
                                double *weights_out, fdms_stream stream) {
    FDMS_BIND(ctx);
    if (n == 0) return FDMS_OK;
    if (!dets || !rows_out || n < 0 || n_scales <= 0)
        return fdms_fail(ctx, FDMS_ERR_INVALID, "fdms_weighted_rows: bad arguments");
    weighted_rows_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(dets, n, n_scales, rows_out,
                                                                                           weights_out);
    FDMS_LAUNCH_CHECK(ctx, "weighted_rows_kernel");
    return FDMS_OK;
}

static int merge_scales_impl(fdms_ctx *ctx, const float *cand, const int32_t *keep, const int32_t *n_keep, int n_img,
                             int n_scales, int cap, const int32_t *in_h_host, int orig_h, int orig_w,
                             const int32_t *orig_hw_dev, double iou_thres, float *dets, int32_t *n_out, int out_cap,
                             int32_t *merged_in, float *kpts_out, fdms_stream stream, const fdms_level *lazy_levels = nullptr,
                             int n_lazy_levels = 0, const int32_t *cand_row = nullptr) {
    FDMS_BIND(ctx);
    if (orig_hw_dev) orig_h = orig_w = 1;  // per-image sizes on the device; the scalars are unused
    if (!cand || !keep || !n_keep || !in_h_host || !dets || !n_out || n_img <= 0 || n_scales <= 0 ||
        n_scales > FDMS_MAX_SCALES || cap <= 0 || out_cap <= 0 || orig_h <= 0 || orig_w <= 0)
        return fdms_fail(ctx, FDMS_ERR_INVALID, "fdms_merge_scales: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    // capacity of the merged (pre-NMS) list of an image: it can never need more than n_scales * cap, and the
    // caller sizes out_cap to hold it (merged_in reports the exact demand, so overflow is detectable)
    long long mcap_ll = (long long)n_scales * cap;
    if (mcap_ll > out_cap) mcap_ll = out_cap;
    const int mcap = (int)mcap_ll;
    if (int rc = fdms_reserve(ctx, ctx->merge_rows, (size_t)n_img * mcap * (MROW * 4 + 4))) return rc;
    if (int rc = fdms_reserve(ctx, ctx->merge_cnt, (size_t)n_img * 2 * 4)) return rc;
    if (int rc = fdms_reserve(ctx, ctx->merge_keep, (size_t)n_img * mcap * 4)) return rc;
    float *mrows = reinterpret_cast<float *>(ctx->merge_rows.p);
    int32_t *msrc = reinterpret_cast<int32_t *>(mrows + (size_t)n_img * mcap * MROW);
    int32_t *mcount = reinterpret_cast<int32_t *>(ctx->merge_cnt.p);
    int32_t *n_keep2 = mcount + n_img;
    int32_t *keep2 = reinterpret_cast<int32_t *>(ctx->merge_keep.p);

    MergeParams P;
    memset(&P, 0, sizeof(P));
    P.cand = cand;
    P.keep = keep;
    P.n_keep = n_keep;
    P.n_scales = n_scales;
    P.cap = cap;
    P.mcap = mcap;
    const int sq = orig_h > orig_w ? orig_h : orig_w;
    for (int s = 0; s < n_scales; s++) {
        if (in_h_host[s] <= 0) return fdms_fail(ctx, FDMS_ERR_INVALID, "fdms_merge_scales: bad in_h[%d]", s);
        P.k[s] = (float)((double)sq / (double)in_h_host[s]);
        P.in_h[s] = in_h_host[s];
    }
    P.orig_hw = orig_hw_dev;
    P.w0 = (float)orig_w;
    P.h0 = (float)orig_h;
    P.mrows = mrows;
    P.msrc = msrc;
    P.mcount = mcount;
    P.merged_in = merged_in;
    EmitParams E;
    memset(&E, 0, sizeof(E));
    E.mrows = mrows;
    E.msrc = msrc;
    E.keep2 = keep2;
    E.n_keep2 = n_keep2;
    E.cand = cand;
    E.n_scales = n_scales;
    E.cap = cap;
    E.mcap = mcap;
    E.out_cap = out_cap;
    for (int s = 0; s < n_scales; s++) {
        E.k[s] = P.k[s];
        E.in_h[s] = P.in_h[s];
    }
    E.orig_hw = orig_hw_dev;
    E.w0 = P.w0;
    E.h0 = P.h0;
    E.dets = dets;
    E.n_out = n_out;
    E.kpts = kpts_out;
    if (kpts_out && lazy_levels && cand_row && n_lazy_levels > 0 && n_lazy_levels <= FDMS_MAX_LEVELS) {
        E.n_levels = n_lazy_levels;
        E.aten = ctx->numerics == FDMS_NUMERICS_ATEN_CUDA;
        E.cand_row = cand_row;
        for (int l = 0; l < n_lazy_levels; l++) {
            E.lv[l].data = lazy_levels[l].data;
            E.lv[l].dtype = lazy_levels[l].dtype;
            E.lv[l].ny = lazy_levels[l].ny;
            E.lv[l].nx = lazy_levels[l].nx;
            E.lv[l].scale = lazy_levels[l].scale;
            E.lv[l].row_offset = lazy_levels[l].row_offset;
            E.lv[l].stride = lazy_levels[l].stride;
        }
    }
    // torchvision.ops.nms on CPU tensors in the reference (msfd:234): TV_CPU flavour, no max_nms cut
    const int small_limit = fdms_merge_small_limit(ctx);
    int skip_le = -1;
    if (small_limit > 0 && ctx->merge_fused) {
        // images of at most small_limit merged rows (the common case): rows -> NMS -> emit in ONE launch
        float thr_f;
        int inclusive;
        fdms_fold_threshold(iou_thres, FDMS_IOU_TV_CPU, &thr_f, &inclusive);
        if (int rc = fdms_merge_small_launch(ctx, P, E, n_img, thr_f, inclusive, st)) return rc;
        if (mcap <= small_limit || ctx->post_small_only || ctx->small_only_sticky) return FDMS_OK;  // no image can have more rows / the caller checks
        skip_le = small_limit;
        ctx->nms_small_done = true;
    } else {
        merge_prepare_kernel<<<n_img, 256, 0, st>>>(P);
        FDMS_LAUNCH_CHECK(ctx, "merge_prepare_kernel");
    }
    const int rc = fdms_nms_impl(ctx, mrows, MROW, 0, 4, nullptr, mcount, n_img, mcap, iou_thres, FDMS_IOU_TV_CPU, 0, keep2,
                                 n_keep2, st);
    ctx->nms_small_done = false;
    if (rc) return rc;
    {
        int parts = ceil_div(mcap < out_cap ? mcap : out_cap, 256);
        if (parts > 16) parts = 16;
        merge_emit_kernel<<<dim3(parts, n_img), 256, 0, st>>>(E, mcount, skip_le);
    }
    FDMS_LAUNCH_CHECK(ctx, "merge_emit_kernel");
    return FDMS_OK;
}

FDMS_API int fdms_merge_scales(fdms_ctx *ctx, const float *cand, const int32_t *keep, const int32_t *n_keep,
                               int n_img, int n_scales, int cap, const int32_t *in_h_host, int orig_h, int orig_w,
                               double iou_thres, float *dets, int32_t *n_out, int out_cap, int32_t *merged_in,
                               float *kpts_out, fdms_stream stream) {
    return merge_scales_impl(ctx, cand, keep, n_keep, n_img, n_scales, cap, in_h_host, orig_h, orig_w, nullptr, iou_thres, dets,
                             n_out, out_cap, merged_in, kpts_out, stream);
}

FDMS_API int fdms_merge_scales_ragged(fdms_ctx *ctx, const float *cand, const int32_t *keep, const int32_t *n_keep,
                                      int n_img, int n_scales, int cap, const int32_t *in_h_host,
                                      const int32_t *orig_hw_dev, double iou_thres, float *dets, int32_t *n_out,
                                      int out_cap, int32_t *merged_in, float *kpts_out, fdms_stream stream) {
    if (!orig_hw_dev) return fdms_fail(ctx, FDMS_ERR_INVALID, "fdms_merge_scales_ragged: null orig_hw");
    return merge_scales_impl(ctx, cand, keep, n_keep, n_img, n_scales, cap, in_h_host, 1, 1, orig_hw_dev, iou_thres, dets, n_out,
                             out_cap, merged_in, kpts_out, stream);
}

FDMS_API int fdms_postprocess(fdms_ctx *ctx, const fdms_post_desc *d, fdms_stream stream) {
    FDMS_BIND(ctx);
    if (!d) return fdms_fail(ctx, FDMS_ERR_INVALID, "fdms_postprocess: null descriptor");
    if ((d->flags & FDMS_POST_SMALL_ONLY) && (fdms_merge_small_limit(ctx) != FDMS_SMALL_LIMIT || !ctx->merge_fused))
        return fdms_fail(ctx, FDMS_ERR_UNSUPPORTED, "fdms_postprocess: FDMS_POST_SMALL_ONLY needs the fused single-CTA kernels");
    struct Guard {  // the flag is per call
        fdms_ctx *c;
        ~Guard() { c->post_small_only = false; }
    } guard{ctx};
    ctx->post_small_only = (d->flags & FDMS_POST_SMALL_ONLY) != 0;
    // landmarks: never for every candidate.  When the caller wants them in the merged output they are decoded by the emit
    // step from the raw head maps, for the final detections only (ctx->lazy_kpts, FDMS_LAZY_KPTS=0: decoded with the boxes)
    const bool lazy = d->kpts_out && ctx->lazy_kpts;
    if (int rc = fdms_decode_filter_ex(ctx, d->levels, d->n_levels, d->n_scales, d->bs, d->conf_thres, d->cand, d->cand_row,
                                       d->count, d->cap, (d->kpts_out && !lazy) ? 0 : FDMS_DECODE_BOXES_ONLY, stream))
        return rc;
    if (int rc = fdms_nms_impl(ctx, d->cand, NO, 0, 4, d->cand_row, d->count, d->bs * d->n_scales, d->cap, d->iou_thres,
                               d->iou_mode, FDMS_MAX_NMS, d->keep, d->n_keep, (cudaStream_t)stream))
        return rc;
    return merge_scales_impl(ctx, d->cand, d->keep, d->n_keep, d->bs, d->n_scales, d->cap, d->in_h_host, d->orig_h, d->orig_w,
                             d->orig_hw_dev, d->iou_thres, d->dets, d->n_out, d->out_cap, d->merged_in, d->kpts_out, stream,
                             lazy ? d->levels : nullptr, lazy ? d->n_levels : 0, lazy ? d->cand_row : nullptr);
}

}  // extern "C"
