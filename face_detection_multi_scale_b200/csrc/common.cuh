// Shared host/device helpers for libfdms (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>

#include "../../include/fdms.h"

#define FDMS_SM_COUNT_FALLBACK 148
#define FDMS_STR2_(x) #x
#define FDMS_STR_(x) FDMS_STR2_(x)

struct fdms_buf {
    void *p = nullptr;
    size_t bytes = 0;
};

struct fdms_ctx {
    int device = 0;
    int sm_count = FDMS_SM_COUNT_FALLBACK;
    int smem_optin = 0;
    char err[512] = {0};
    int pyr_attr = 0;
    int pyr_special = 1;  // FDMS_PYR_SPECIAL=0 disables the integer-scale fast paths (testing)
    int pyr_smem_budget = 32 * 1024, pyr_th = 12;   // FDMS_PYR_SMEM_KB / FDMS_PYR_TH (tuning knobs)
    int pyr_direct = 3, pyr_th_direct = 16;  // FDMS_PYR_DIRECT (bit 0: integer factors, bit 1: the 3:2 factor) / FDMS_PYR_TH_DIRECT: unstaged tiles
    // FDMS_PYR_STRIP=0 disables the strip kernel (one source pass for all exact-factor scales); stage / ring size knobs
    int pyr_strip = 1, pyr_strip_stage_bytes = 36 * 1024, pyr_strip_smem = 108 * 1024, pyr_strip_warps = 14, pyr_strip_regs = 64;
    int pyr_strip_ctas = 0, pyr_strip_reserve = 0;  // FDMS_PYR_STRIP_CTAS (0 = one per SM), FDMS_PYR_STRIP_RESERVE_KB: shared memory claimed per CTA beyond the stages (keeps other kernels' CTAs off its SM)
    bool pyr_strip_attr = false, pyr_cnt_ready = false;
    bool nms_attr = false, nms_small_attr = false;
    int nms_small_limit = 1024;   // FDMS_NMS_SMALL_LIMIT: largest segment of the fused kernel (0 disables it)
    int decode_fused = 1;         // FDMS_DECODE_FUSED=0: candidates are decoded by a second kernel (decode_candidates_kernel)
    int merge_fused = 1;          // FDMS_MERGE_FUSED=0: the three-kernel merge (prepare, NMS, emit) for every image
    int numerics = FDMS_NUMERICS_EXACT;  // fdms_set_numerics / FDMS_NUMERICS
    bool post_small_only = false; // set around fdms_postprocess(FDMS_POST_SMALL_ONLY): no large-segment launches
    bool small_only_sticky = false;  // fdms_set_small_only: the same for fdms_nms / fdms_merge_scales called one by one
    bool nms_small_done = false;  // set around fdms_nms_impl by the fused merge: segments <= the limit are done already
    int nms_greedy_min_seg = 8;   // FDMS_NMS_GREEDY_MIN_SEG: fewer segments than this -> bit-matrix path
    int nms_bin = 1;              // FDMS_NMS_BIN: 1 = nms_binned_kernel (kept boxes in a size-stratified grid), 0 = the same
                                  // kernel without the grid, 2 = nms_greedy_kernel (the list sweep it replaced; A/B)
    bool nms_bin_attr = false;
    int lazy_kpts = 1;            // FDMS_LAZY_KPTS: fdms_postprocess decodes landmarks at emit time, for the final detections only
    int pdl = 0;                  // FDMS_PDL: 1 = the per-scale NMS and the merge are launched as programmatic dependents of the kernel before them
    int nms_reg_sort = 1;         // FDMS_NMS_REG_SORT: 0 = every bitonic stage through shared memory (A/B)
    size_t mask_budget = (size_t)8 << 30;  // bytes; FDMS_MASK_BUDGET_MB overrides
    // scratch (grow-only)
    fdms_buf sort_a;           // NMS: 64-bit keys of segments too large for shared memory
    fdms_buf idx_a, idx_b;     // NMS: item indices (sort scratch / score order)
    fdms_buf boxes;            // NMS: boxes in score order (float4)
    fdms_buf mask;             // NMS: suppression bit matrix
    fdms_buf seginfo;          // NMS: per segment {n, nblk, mask_off}
    fdms_buf merge_rows;       // merge: (n_img, mcap, 8) rows
    fdms_buf merge_cnt;        // merge: counts
    fdms_buf merge_keep;       // merge: keep + n_keep
    fdms_buf pyr_cnt;          // strip kernel: work-item counter + finished-CTA counter (self-resetting)
    fdms_buf pyr_tab;          // pyramid tap tables of the last geometry
    unsigned long long pyr_tab_key = 0;
    fdms_buf jpg_coef, jpg_planes;  // JPEG ingest: coefficients + descriptors, component sample planes (device)
    void *jpg_host = nullptr;       // JPEG ingest: pinned staging of the coefficients (written by the host Huffman threads)
    size_t jpg_host_bytes = 0;
    cudaEvent_t jpg_event = nullptr;  // the upload of the staging buffer has completed
    bool jpg_pending = false;
    fdms_buf dec_tab;          // decode filter: per-CTA table of the last head geometry
    unsigned long long dec_tab_key = 0;
};

static inline int fdms_fail(fdms_ctx *ctx, int code, const char *fmt, ...) {
    if (ctx) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
        va_end(ap);
    }
    return code;
}

#define FDMS_CUDA(ctx, call)                                                                          \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fdms_fail((ctx), FDMS_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                             __FILE__, __LINE__);                                                     \
    } while (0)

#define FDMS_LAUNCH_CHECK(ctx, name)                                                                  \
    do {                                                                                              \
        cudaError_t e_ = cudaGetLastError();                                                          \
        if (e_ != cudaSuccess)                                                                        \
            return fdms_fail((ctx), FDMS_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e_)); \
    } while (0)

// Grow-only scratch.  Growth synchronises the device (cudaFree) -- it only happens when a call
// needs more than any call before it.
static inline int fdms_reserve(fdms_ctx *ctx, fdms_buf &b, size_t bytes) {
    if (bytes <= b.bytes) return FDMS_OK;
    size_t want = bytes + bytes / 4 + 256;
    if (b.p) {
        cudaDeviceSynchronize();
        cudaFree(b.p);
        b.p = nullptr;
        b.bytes = 0;
    }
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fdms_fail(ctx, FDMS_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    }
    b.bytes = want;
    return FDMS_OK;
}

// Makes the context's device current for the duration of an API call and restores the caller's device afterwards
// (torch's current device must not drift when a tensor lives on another GPU than the current one).
// Programmatic dependent launch (sm_90+): a kernel launched with the attribute may be scheduled while the kernel before it
// in the stream still runs; fdms_pdl_wait() blocks until that kernel has completed and its writes are visible (a no-op for
// a normal launch), fdms_pdl_trigger() lets the dependents of THIS kernel be scheduled from now on.
#ifdef __CUDACC__
__device__ __forceinline__ void fdms_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void fdms_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

struct fdms_device_guard {
    int prev = -1;
    bool switched = false;
    int bind(fdms_ctx *ctx) {
        if (!ctx) return FDMS_ERR_INVALID;
        cudaError_t e = cudaGetDevice(&prev);
        if (e != cudaSuccess) return fdms_fail(ctx, FDMS_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
        if (prev != ctx->device) {
            e = cudaSetDevice(ctx->device);
            if (e != cudaSuccess) return fdms_fail(ctx, FDMS_ERR_CUDA, "cudaSetDevice(%d): %s", ctx->device, cudaGetErrorString(e));
            switched = true;
        }
        return FDMS_OK;
    }
    ~fdms_device_guard() {
        if (switched) cudaSetDevice(prev);
    }
};
#define FDMS_BIND(ctx)                            \
    fdms_device_guard fdms_guard_;                \
    if (int rc_ = fdms_guard_.bind(ctx)) return rc_

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device helpers ------------------------------------------------------------------------
// Every fp32 operation that the reference rounds separately is written with an explicit
// round-to-nearest intrinsic so that neither nvcc (-fmad) nor a later edit can contract it.
__device__ __forceinline__ float f_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float f_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float f_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float f_div(float a, float b) { return __fdiv_rn(a, b); }

// 1 / (1 + exp(-x)).  The reference's sigmoid is ATen's: on the CPU a vectorised expf (Sleef, <= 1 ulp), on a CUDA
// device `one / (one + std::exp(-a))` with libdevice's expf (<= 2 ulp) -- two different roundings of the same value.
// Default (`aten` false): exp evaluated in fp64 and rounded once, i.e. the correctly rounded expf, which the CPU oracle
// restates bit for bit and which lies within 1 ulp of either.  `aten` true (FDMS_NUMERICS_ATEN_CUDA): the very expf
// ATen's CUDA kernel calls, so that the decode is bit-identical to the reference running in fp32 on the same GPU.
__device__ __forceinline__ float sigmoid_ref(float x, bool aten) {
    const float e = aten ? expf(-x) : (float)exp(-(double)x);
    return f_div(1.0f, f_add(1.0f, e));
}

// Round to half precision and back: one ATen operation on a half tensor (opmath float, the result stored as c10::Half).
__device__ __forceinline__ float h_rn(float x) { return __half2float(__float2half_rn(x)); }

// The head decode as ATen computes it ON A CUDA DEVICE FOR A HALF MODEL (the reference's default, msfd:49 `model.half()`):
// the conv outputs, the sigmoid, `y * 2.`, `- 0.5`, `** 2` and `* anchor_grid` are half tensors (every operation rounds to
// half), `self.grid[i]` is created float32 (models/yolo.py:310-313), so `+ grid` promotes to float32 and `* stride` stays
// there; the keypoint coordinates are written back into the half tensor x_kpt (rounded once more) and `torch.cat` promotes
// the row to float32 -- pred is a float32 tensor of half-valued columns 2..20 and float32 xy.  Verified op by op against the
// reference module on half tensors (tests/ref_live_check.py) and on the GPU (tests/ref_gpu_check.py, part C).
__device__ __forceinline__ float h_sigmoid(float x) { return h_rn(f_div(1.0f, f_add(1.0f, expf(-x)))); }
__device__ __forceinline__ float h_xy(float s, float g, float stride) {  // (s * 2. - 0.5 + grid) * stride, s a half value
    return f_mul(f_add(h_rn(f_sub(h_rn(f_mul(s, 2.f)), 0.5f)), g), stride);
}
__device__ __forceinline__ float h_wh(float s, float anchor) {  // (s * 2) ** 2 * anchor_grid, all half tensors
    const float t = h_rn(f_mul(s, 2.f));
    return h_rn(f_mul(h_rn(f_mul(t, t)), h_rn(anchor)));
}

__device__ __forceinline__ float ld_f(const void *p, int dtype, int64_t i) {
    return dtype == FDMS_F16 ? __half2float(__ldg(reinterpret_cast<const __half *>(p) + i))
                             : __ldg(reinterpret_cast<const float *>(p) + i);
}
