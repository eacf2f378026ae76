// Declarations shared by the two pyramid kernels (pyramid.cu: tiles per scale; pyramid_strip.cu: one pass over the
// source for all exact-factor scales).
#pragma once
#include "common.cuh"

// Exact integer scale factors make OpenCV's fixed-point weights constant: an odd factor k gives
// f == 0 on both axes (weights 2048/0: the output IS source pixel (k*d + (k-1)/2)), an even factor gives f == 0.5
// (weights 1024/1024: (((1024*(64*(a+b)))>>16) + ... + 2) >> 2 == (a + b + c + d + 2) >> 2).  Same results as the
// general path, fewer instructions.
enum { PM_GENERAL = 0, PM_POINT = 1, PM_HALF = 2, PM_RATIO32 = 3 };

struct PyrScale {
    void *out;
    long long img_stride, cs, ys, xs;  // element strides
    int dst_h, dst_w, unpad_h, unpad_w, top, left;
    int tile_begin, tiles_x;
    int tw, th;        // output tile
    int dense_rows;    // row slots cover a contiguous source-row range (scale_y <= 2)
    int slot_bytes;    // multiple of 16
    int smem_need;     // host only: dynamic shared memory of a staged tile of this scale
    int vec;           // paired stores allowed
    int mode;          // PM_GENERAL / PM_POINT / PM_HALF: strength-reduced arithmetic for exact integer scales
    int direct;        // exact integer scale, no column padding: taps are read straight from global memory (no staging)
    int kx, ky;        // the integer factors when mode != PM_GENERAL
    int real_rows;     // direct tiles: interior rows whose taps are real image rows; the rest read the virtual zeros
    double scale_x, scale_y;  // cv2: 1.0 / (dst / src) over the virtual extent
    const int2 *xtab;         // per output column: {source column sx, a0 | a1 << 16}; sx = -1 on the border
    const int4 *ytab;         // per output row: {r0, r1, b0 | b1 << 16, s}: clamped tap rows, weights, unclamped upper tap row; r0 = -1 on the border
};

struct PyrParams {
    const uint8_t *src;
    long long img_stride, pitch;
    int src_h, src_w, virt_h, virt_w;
    int n_scales, flip, border, vec_ok;
    int shifted;  // !vec_ok: rows are staged with a per-slot byte shift (see TileShared::slot_shift)
    PyrScale sc[FDMS_MAX_SCALES];
};

static __device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

static __device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
static __device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
static __device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE;\n"
        "bra LAB_WAIT;\n"
        "LAB_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA bulk copy global -> shared, 16 B aligned on both sides, size a multiple of 16.
static __device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}


// pyramid_strip.cu: launches the strip kernel for the scales of P it can handle (exact factors, fp16 CHW, rows-only
// padding, 16-byte aligned source) and returns their bit mask in *handled; the caller runs the rest through the tile kernel.
int fdms_pyramid_strip(fdms_ctx *ctx, const PyrParams &P, int n_img, unsigned *handled, cudaStream_t st);
