// "Next" rows around the hot path: WIDER-FACE evaluation IoU matrix and the padded
// per-frame output tensors of the production driver.
#include "common.cuh"

namespace {

// widerface_evaluate/box_overlaps.pyx:15-55 -- float64, +1 pixel convention, 0 where the boxes do not
// overlap.  One thread per (n, k) pair; consecutive threads share the box row and walk the query boxes,
// so the (N, K) row-major output is written coalesced.
__global__ void bbox_overlaps_kernel(const double *__restrict__ boxes, long long N, const double *__restrict__ query,
                                     long long K, double *__restrict__ out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * K) return;
    const long long n = idx / K, k = idx - n * K;
    const double b0 = boxes[n * 4], b1 = boxes[n * 4 + 1], b2 = boxes[n * 4 + 2], b3 = boxes[n * 4 + 3];
    const double q0 = query[k * 4], q1 = query[k * 4 + 1], q2 = query[k * 4 + 2], q3 = query[k * 4 + 3];
    double v = 0.0;
    const double iw = __dadd_rn(__dsub_rn(fmin(b2, q2), fmax(b0, q0)), 1.0);  // :38-41
    if (iw > 0) {
        const double ih = __dadd_rn(__dsub_rn(fmin(b3, q3), fmax(b1, q1)), 1.0);  // :43-46
        if (ih > 0) {
            const double box_area = __dmul_rn(__dadd_rn(__dsub_rn(q2, q0), 1.0), __dadd_rn(__dsub_rn(q3, q1), 1.0));
            const double area_n = __dmul_rn(__dadd_rn(__dsub_rn(b2, b0), 1.0), __dadd_rn(__dsub_rn(b3, b1), 1.0));
            const double inter = __dmul_rn(iw, ih);
            const double ua = __dsub_rn(__dadd_rn(area_n, box_area), inter);  // :48-52
            v = __ddiv_rn(inter, ua);                                        // :53
        }
    }
    out[idx] = v;
}

// yolov7_face_multi_scale_dataframe_predict.py:699-716 (normalize_bbox = x / img_width, y / img_height in
// float64, utils/preprocess_yolo_predict.py:26-49) and utils/preprocess_yolo_predict.py:478-520 (pad every
// frame to max_faces with -1): dets (n_img, out_cap, 7) f32 + n_out -> bboxes (n_img, max_faces, 4) f64,
// confidence (n_img, max_faces) f64, class_indexes (n_img, max_faces) i32, scale_used (n_img, max_faces) i32.
__global__ void pack_frames_kernel(const float *__restrict__ dets, const int32_t *__restrict__ n_out, int out_cap,
                                   int max_faces, double img_w, double img_h, double *__restrict__ bboxes,
                                   double *__restrict__ conf, int32_t *__restrict__ cls, int32_t *__restrict__ scale) {
    const int img = blockIdx.y;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= max_faces) return;
    const int n = min(max(n_out[img], 0), min(out_cap, max_faces));
    const long long o = (long long)img * max_faces + f;
    if (f < n) {
        const float *d = dets + ((long long)img * out_cap + f) * 7;
        bboxes[o * 4 + 0] = __ddiv_rn((double)d[0], img_w);
        bboxes[o * 4 + 1] = __ddiv_rn((double)d[1], img_h);
        bboxes[o * 4 + 2] = __ddiv_rn((double)d[2], img_w);
        bboxes[o * 4 + 3] = __ddiv_rn((double)d[3], img_h);
        conf[o] = (double)d[4];
        cls[o] = (int32_t)d[5];
        scale[o] = (int32_t)d[6];
    } else {
        bboxes[o * 4 + 0] = bboxes[o * 4 + 1] = bboxes[o * 4 + 2] = bboxes[o * 4 + 3] = -1.0;
        conf[o] = -1.0;
        cls[o] = -1;
        scale[o] = -1;
    }
}

}  // namespace

extern "C" {

FDMS_API int fdms_bbox_overlaps(fdms_ctx *ctx, const double *boxes, int64_t n, const double *query, int64_t k,
                                double *overlaps, fdms_stream stream) {
    FDMS_BIND(ctx);
    if (n == 0 || k == 0) return FDMS_OK;
    if (!boxes || !query || !overlaps || n < 0 || k < 0 || n * k > 0x7fffffff00LL)
        return fdms_fail(ctx, FDMS_ERR_INVALID, "fdms_bbox_overlaps: bad arguments");
    bbox_overlaps_kernel<<<(unsigned)ceil_div64(n * k, 256), 256, 0, (cudaStream_t)stream>>>(boxes, n, query, k, overlaps);
    FDMS_LAUNCH_CHECK(ctx, "bbox_overlaps_kernel");
    return FDMS_OK;
}

FDMS_API int fdms_pack_frames(fdms_ctx *ctx, const float *dets, const int32_t *n_out, int n_img, int out_cap,
                              int max_faces, int img_w, int img_h, double *bboxes, double *confidence,
                              int32_t *class_indexes, int32_t *scale_used, fdms_stream stream) {
    FDMS_BIND(ctx);
    if (max_faces == 0 || n_img == 0) return FDMS_OK;
    if (!dets || !n_out || !bboxes || !confidence || !class_indexes || !scale_used || n_img < 0 || n_img > 65535 ||
        out_cap <= 0 || max_faces < 0 || img_w <= 0 || img_h <= 0)
        return fdms_fail(ctx, FDMS_ERR_INVALID, "fdms_pack_frames: bad arguments");
    dim3 grid(ceil_div(max_faces, 128), n_img);
    pack_frames_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(dets, n_out, out_cap, max_faces, (double)img_w,
                                                               (double)img_h, bboxes, confidence, class_indexes,
                                                               scale_used);
    FDMS_LAUNCH_CHECK(ctx, "pack_frames_kernel");
    return FDMS_OK;
}

}  // extern "C"
