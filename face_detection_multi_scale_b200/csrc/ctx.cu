// Context, error reporting and the host-side letterbox geometry of libfdms.
#include <math.h>
#include <stdlib.h>
#include <new>

#include "common.cuh"

namespace {
// Test hook: fills the whole shared memory of an SM with one byte value, so that a kernel reading shared memory it
// never wrote sees that value instead of whatever the previous kernel left there.
__global__ void poison_smem_kernel(unsigned word, int n_words) {
    extern __shared__ unsigned poison_buf[];
    for (int i = threadIdx.x; i < n_words; i += blockDim.x) poison_buf[i] = word;
    __syncthreads();
    if (poison_buf[(threadIdx.x * 97) % n_words] != word) __trap();  // keeps the stores alive
}
}  // namespace

extern "C" {

FDMS_API int fdms_version(void) { return FDMS_VERSION; }

FDMS_API int fdms_debug_poison_smem(fdms_ctx *ctx, int byte_value, fdms_stream stream) {
    FDMS_BIND(ctx);
    const int bytes = ctx->smem_optin > 0 ? ctx->smem_optin : 48 * 1024;
    FDMS_CUDA(ctx, cudaFuncSetAttribute(poison_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    const unsigned b = (unsigned)byte_value & 0xffu, word = b * 0x01010101u;
    // one CTA owns the whole shared memory of an SM; a few waves so that every SM gets at least one
    poison_smem_kernel<<<4 * ctx->sm_count, 256, bytes, (cudaStream_t)stream>>>(word, bytes / 4);
    FDMS_LAUNCH_CHECK(ctx, "poison_smem_kernel");
    return FDMS_OK;
}

FDMS_API const char *fdms_build_info(void) {
    return "libfdms " __DATE__ " " __TIME__ " sm_100a nvcc " FDMS_STR_(__CUDACC_VER_MAJOR__) "." FDMS_STR_(__CUDACC_VER_MINOR__)
           " fmad=off";
}

FDMS_API int fdms_create(fdms_ctx **out, int device) {
    if (!out) return FDMS_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return FDMS_ERR_CUDA;  // no CUDA device: there is no CPU fallback
    }
    if (device < 0 || device >= n) return FDMS_ERR_INVALID;
    fdms_ctx *c = new (std::nothrow) fdms_ctx();
    if (!c) return FDMS_ERR_NOMEM;
    c->device = device;
    int prev_dev = -1;
    cudaGetDevice(&prev_dev);
    if (cudaSetDevice(device) != cudaSuccess) {
        delete c;
        return FDMS_ERR_CUDA;
    }
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&c->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (c->sm_count <= 0) c->sm_count = FDMS_SM_COUNT_FALLBACK;
    if (const char *e = getenv("FDMS_NUMERICS")) c->numerics = (!strcmp(e, "aten_cuda") || !strcmp(e, "1")) ? FDMS_NUMERICS_ATEN_CUDA : FDMS_NUMERICS_EXACT;
    if (const char *e = getenv("FDMS_PYR_SPECIAL")) c->pyr_special = atoi(e);
    if (const char *e = getenv("FDMS_PYR_SMEM_KB")) c->pyr_smem_budget = atoi(e) * 1024;
    if (const char *e = getenv("FDMS_PYR_DIRECT")) c->pyr_direct = atoi(e);
    if (const char *e = getenv("FDMS_PYR_STRIP")) c->pyr_strip = atoi(e);
    if (const char *e = getenv("FDMS_PYR_STRIP_STAGE_KB")) c->pyr_strip_stage_bytes = atoi(e) * 1024;
    if (const char *e = getenv("FDMS_PYR_STRIP_REGS")) c->pyr_strip_regs = atoi(e);
    if (const char *e = getenv("FDMS_PYR_STRIP_WARPS")) c->pyr_strip_warps = atoi(e);
    if (const char *e = getenv("FDMS_PYR_STRIP_SMEM_KB")) c->pyr_strip_smem = atoi(e) * 1024;
    if (const char *e = getenv("FDMS_PYR_STRIP_CTAS")) c->pyr_strip_ctas = atoi(e);
    if (const char *e = getenv("FDMS_PYR_STRIP_RESERVE_KB")) c->pyr_strip_reserve = atoi(e) * 1024;
    if (const char *e = getenv("FDMS_PYR_TH_DIRECT")) c->pyr_th_direct = atoi(e) < 1 ? 1 : atoi(e);
    if (const char *e = getenv("FDMS_PYR_TH")) c->pyr_th = atoi(e) < 1 ? 1 : (atoi(e) > 31 ? 31 : atoi(e));  // rows + sentinel are written by one warp
    if (const char *e = getenv("FDMS_NMS_SMALL_LIMIT")) c->nms_small_limit = atoi(e);
    if (const char *e = getenv("FDMS_DECODE_FUSED")) c->decode_fused = atoi(e);
    if (const char *e = getenv("FDMS_MERGE_FUSED")) c->merge_fused = atoi(e);
    if (const char *e = getenv("FDMS_NMS_GREEDY_MIN_SEG")) c->nms_greedy_min_seg = atoi(e);
    if (const char *e = getenv("FDMS_PDL")) c->pdl = atoi(e);
    if (const char *e = getenv("FDMS_LAZY_KPTS")) c->lazy_kpts = atoi(e);
    if (const char *e = getenv("FDMS_NMS_BIN")) c->nms_bin = atoi(e);
    if (const char *e = getenv("FDMS_NMS_REG_SORT")) c->nms_reg_sort = atoi(e);
    if (const char *e = getenv("FDMS_MASK_BUDGET_MB")) c->mask_budget = (size_t)atoll(e) << 20;
    if (prev_dev >= 0 && prev_dev != device) cudaSetDevice(prev_dev);  // the caller's current device is left as it was
    *out = c;
    return FDMS_OK;
}

FDMS_API void fdms_destroy(fdms_ctx *ctx) {
    if (!ctx) return;
    int prev_dev = -1;
    cudaGetDevice(&prev_dev);
    cudaSetDevice(ctx->device);
    fdms_buf *bufs[] = {&ctx->sort_a, &ctx->idx_a, &ctx->idx_b, &ctx->boxes, &ctx->mask,
                        &ctx->seginfo, &ctx->merge_rows, &ctx->merge_cnt, &ctx->merge_keep, &ctx->pyr_tab, &ctx->dec_tab, &ctx->pyr_cnt, &ctx->jpg_coef, &ctx->jpg_planes};
    for (fdms_buf *b : bufs)
        if (b->p) cudaFree(b->p);
    if (ctx->jpg_host) cudaFreeHost(ctx->jpg_host);
    if (ctx->jpg_event) cudaEventDestroy(ctx->jpg_event);
    if (prev_dev >= 0 && prev_dev != ctx->device) cudaSetDevice(prev_dev);
    delete ctx;
}

FDMS_API int fdms_set_numerics(fdms_ctx *ctx, int mode) {
    if (!ctx || (mode != FDMS_NUMERICS_EXACT && mode != FDMS_NUMERICS_ATEN_CUDA)) return FDMS_ERR_INVALID;
    ctx->numerics = mode;
    return FDMS_OK;
}

FDMS_API int fdms_get_numerics(const fdms_ctx *ctx) { return ctx ? ctx->numerics : FDMS_ERR_INVALID; }

FDMS_API int fdms_set_small_only(fdms_ctx *ctx, int on) {
    if (!ctx) return FDMS_ERR_INVALID;
    ctx->small_only_sticky = on != 0;
    return FDMS_OK;
}

FDMS_API const char *fdms_last_error(const fdms_ctx *ctx) { return ctx ? ctx->err : "null ctx"; }

FDMS_API int64_t fdms_scratch_bytes(const fdms_ctx *ctx) {
    if (!ctx) return 0;
    const fdms_buf *bufs[] = {&ctx->sort_a, &ctx->idx_a, &ctx->idx_b, &ctx->boxes, &ctx->mask,
                              &ctx->seginfo, &ctx->merge_rows, &ctx->merge_cnt, &ctx->merge_keep, &ctx->pyr_tab, &ctx->dec_tab, &ctx->pyr_cnt, &ctx->jpg_coef, &ctx->jpg_planes};
    int64_t t = 0;
    for (const fdms_buf *b : bufs) t += (int64_t)b->bytes;
    return t;
}

// utils/datasets.py:873-903.  Python's round() on a float is round-half-to-even in double
// (nearbyint under the default rounding mode); np.mod takes the sign of the divisor.
FDMS_API int fdms_letterbox_geometry(int src_h, int src_w, int new_h, int new_w, int auto_, int scale_fill,
                                     int scaleup, int stride, fdms_geom *g) {
    if (!g || src_h <= 0 || src_w <= 0 || new_h <= 0 || new_w <= 0 || stride <= 0) return FDMS_ERR_INVALID;
    double r = fmin((double)new_h / src_h, (double)new_w / src_w);  // :880
    if (!scaleup) r = fmin(r, 1.0);                                  // :881-882
    int unpad_w = (int)nearbyint(src_w * r), unpad_h = (int)nearbyint(src_h * r);  // :886
    double dw = new_w - unpad_w, dh = new_h - unpad_h;                               // :887
    double rw = r, rh = r;
    if (auto_) {  // :888-889
        int idw = (int)dw % stride, idh = (int)dh % stride;
        if (idw < 0) idw += stride;
        if (idh < 0) idh += stride;
        dw = idw;
        dh = idh;
    } else if (scale_fill) {  // :890-893
        dw = 0.0;
        dh = 0.0;
        unpad_w = new_w;
        unpad_h = new_h;
        rw = (double)new_w / src_w;
        rh = (double)new_h / src_h;
    }
    dw /= 2;  // :895-896
    dh /= 2;
    memset(g, 0, sizeof(*g));
    g->resize = !(src_w == unpad_w && src_h == unpad_h);  // :898
    g->top = (int)nearbyint(dh - 0.1);                    // :900
    g->bottom = (int)nearbyint(dh + 0.1);
    g->left = (int)nearbyint(dw - 0.1);  // :901
    g->right = (int)nearbyint(dw + 0.1);
    g->unpad_h = unpad_h;
    g->unpad_w = unpad_w;
    g->dst_h = unpad_h + g->top + g->bottom;
    g->dst_w = unpad_w + g->left + g->right;
    g->ratio_w = rw;
    g->ratio_h = rh;
    g->dw = dw;
    g->dh = dh;
    if (unpad_h <= 0 || unpad_w <= 0) return FDMS_ERR_INVALID;
    return FDMS_OK;
}

}  // extern "C"
