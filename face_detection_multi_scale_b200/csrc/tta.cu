// "Next" row: the elementwise steps of Model.forward_augment (models/yolo.py:363-374) that sit
// on either side of the backbone -- scale_img (utils/torch_utils.py:247-257) before it, _descale_pred
// (models/yolo.py:402-417) after it.  The backbone forward itself stays the reference's PyTorch code.
#include "common.cuh"

namespace {

template <typename T> __device__ __forceinline__ float tta_ld(const T *p, long long i);
template <> __device__ __forceinline__ float tta_ld<float>(const float *p, long long i) { return __ldg(p + i); }
template <> __device__ __forceinline__ float tta_ld<__half>(const __half *p, long long i) { return __half2float(__ldg(p + i)); }
template <typename T> __device__ __forceinline__ void tta_st(T *p, long long i, float v);
template <> __device__ __forceinline__ void tta_st<float>(float *p, long long i, float v) { p[i] = v; }
template <> __device__ __forceinline__ void tta_st<__half>(__half *p, long long i, float v) { p[i] = __float2half_rn(v); }

// F.interpolate(mode='bilinear', align_corners=False) with an explicit output size: ATen's
// area_pixel_compute_source_index -- scale = in / out in fp32, src = scale * (dst + 0.5) - 0.5 clamped at 0,
// the upper neighbour clamps at the edge -- and the four-tap sum of upsample_bilinear2d.  The optional flip is
// x.flip(3) / x.flip(2) applied to the SOURCE (models/yolo.py:368); columns / rows beyond the resized image are the
// F.pad(value=0.447) of scale_img.  One thread per output element, x fastest.
template <typename T>
__global__ void scale_img_kernel(const T *__restrict__ src, int planes, int h, int w, int rh, int rw, int oh, int ow,
                                 int flip, float pad_value, float sy, float sx, int aten, T *__restrict__ dst) {
    const long long total = (long long)planes * oh * ow;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % ow);
        const long long t = idx / ow;
        const int y = (int)(t % oh);
        const long long p = t / oh;
        float v = pad_value;
        if (y < rh && x < rw) {
            const float fy = fmaxf(__fmaf_rn(sy, __fadd_rn((float)y, 0.5f), -0.5f), 0.f);  // ATen contracts this to an FMA
            const float fx = fmaxf(__fmaf_rn(sx, __fadd_rn((float)x, 0.5f), -0.5f), 0.f);
            const int y0 = (int)fy, x0 = (int)fx;
            const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
            const float ly = __fsub_rn(fy, (float)y0), lx = __fsub_rn(fx, (float)x0);
            const float hy = __fsub_rn(1.f, ly), hx = __fsub_rn(1.f, lx);
            const int ya = flip == 2 ? h - 1 - y0 : y0, yb = flip == 2 ? h - 1 - y1 : y1;
            const int xa = flip == 3 ? w - 1 - x0 : x0, xb = flip == 3 ? w - 1 - x1 : x1;
            const long long base = p * h * w;
            const float v00 = tta_ld<T>(src, base + (long long)ya * w + xa), v01 = tta_ld<T>(src, base + (long long)ya * w + xb);
            const float v10 = tta_ld<T>(src, base + (long long)yb * w + xa), v11 = tta_ld<T>(src, base + (long long)yb * w + xb);
            if (aten)  // upsample_bilinear2d_out_frame as nvcc contracted it in torch's sm_100 build (read from its SASS):
                       // h0 * (w0 * a + w1 * b) + h1 * (w0 * c + w1 * d) = fma(h0, fma(w0, a, w1 * b), h1 * fma(w0, c, w1 * d))
                v = __fmaf_rn(hy, __fmaf_rn(hx, v00, __fmul_rn(lx, v01)), __fmul_rn(ly, __fmaf_rn(hx, v10, __fmul_rn(lx, v11))));
            else  // every product and sum rounded (the CPU oracle's statement; <= 2 ulp from ATen on either device)
                v = __fadd_rn(__fmul_rn(hy, __fadd_rn(__fmul_rn(hx, v00), __fmul_rn(lx, v01))),
                              __fmul_rn(ly, __fadd_rn(__fmul_rn(hx, v10), __fmul_rn(lx, v11))));
        }
        tta_st<T>(dst, idx, v);
    }
}

// models/yolo.py:402-417: p[..., :4] /= scale; flips == 2: y = img_h - y; flips == 3: x = img_w - x.  Landmark
// columns are left as they are (the reference does not de-scale them).  Rows are written to `dst` at row offset
// `row0` of a (bs, rows_total, no) tensor: that is the torch.cat(y, 1) of forward_augment (:374).
__global__ void descale_pred_kernel(const float *__restrict__ pred, int bs, long long rows, int no, float scale,
                                    float inv_scale, int flip, float img_h, float img_w, float *__restrict__ dst, long long rows_total,
                                    long long row0) {
    const long long total = (long long)bs * rows * no;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % no);
        const long long t = idx / no;
        const long long r = t % rows, b = t / rows;
        float v = pred[idx];
        if (c < 4) v = inv_scale != 0.f ? __fmul_rn(v, inv_scale) : __fdiv_rn(v, scale);  // ATen CUDA: v * (1 / scale)
        if (flip == 2 && c == 1) v = __fsub_rn(img_h, v);
        if (flip == 3 && c == 0) v = __fsub_rn(img_w, v);
        dst[((long long)b * rows_total + row0 + r) * no + c] = v;
    }
}

}  // namespace

extern "C" {

FDMS_API int fdms_scale_img(fdms_ctx *ctx, const void *src, int dtype, int n, int c, int h, int w, int resized_h,
                            int resized_w, int out_h, int out_w, int flip, float pad_value, void *dst,
                            fdms_stream stream) {
    FDMS_BIND(ctx);
    if (!src || !dst || n <= 0 || c <= 0 || h <= 0 || w <= 0 || resized_h <= 0 || resized_w <= 0 || out_h < resized_h ||
        out_w < resized_w || (flip != 0 && flip != 2 && flip != 3) || (dtype != FDMS_F16 && dtype != FDMS_F32))
        return fdms_fail(ctx, FDMS_ERR_INVALID, "fdms_scale_img: bad arguments");
    const long long total = (long long)n * c * out_h * out_w;
    const int threads = 256;
    const int blocks = (int)min((long long)ctx->sm_count * 16, (long long)ceil_div64(total, threads));
    // ATen: area_pixel_compute_scale<float>(in, out, align_corners=false, scales=nullopt) = (float)in / out
    const float sy = (float)h / (float)resized_h, sx = (float)w / (float)resized_w;
    cudaStream_t st = (cudaStream_t)stream;
    const int aten = ctx->numerics == FDMS_NUMERICS_ATEN_CUDA;
    if (dtype == FDMS_F16)
        scale_img_kernel<__half><<<blocks, threads, 0, st>>>(reinterpret_cast<const __half *>(src), n * c, h, w, resized_h,
                                                            resized_w, out_h, out_w, flip, pad_value, sy, sx, aten,
                                                            reinterpret_cast<__half *>(dst));
    else
        scale_img_kernel<float><<<blocks, threads, 0, st>>>(reinterpret_cast<const float *>(src), n * c, h, w, resized_h,
                                                           resized_w, out_h, out_w, flip, pad_value, sy, sx, aten,
                                                           reinterpret_cast<float *>(dst));
    FDMS_LAUNCH_CHECK(ctx, "scale_img_kernel");
    return FDMS_OK;
}

FDMS_API int fdms_descale_pred(fdms_ctx *ctx, const float *pred, int bs, int64_t rows, int no, float scale, int flip,
                               int img_h, int img_w, float *dst, int64_t rows_total, int64_t row_offset,
                               fdms_stream stream) {
    FDMS_BIND(ctx);
    if (!pred || !dst || bs <= 0 || rows <= 0 || no < 4 || !(scale > 0.f) || (flip != 0 && flip != 2 && flip != 3) ||
        row_offset < 0 || row_offset + rows > rows_total)
        return fdms_fail(ctx, FDMS_ERR_INVALID, "fdms_descale_pred: bad arguments");
    const long long total = (long long)bs * rows * no;
    const int threads = 256;
    const int blocks = (int)min((long long)ctx->sm_count * 16, (long long)ceil_div64(total, threads));
    descale_pred_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(pred, bs, rows, no, scale, ctx->numerics == FDMS_NUMERICS_ATEN_CUDA ? 1.0f / scale : 0.f, flip, (float)img_h,
                                                                      (float)img_w, dst, rows_total, row_offset);
    FDMS_LAUNCH_CHECK(ctx, "descale_pred_kernel");
    return FDMS_OK;
}

}  // extern "C"
