// Fused letterbox pyramid: bilinear resize (OpenCV INTER_LINEAR u8 fixed point, bit exact),
// constant border, channel flip, /255 and pack -- every image x every scale in ONE launch.
//
// Replaces utils/datasets.py:873-903 (letterbox), multi_scale_face_detector.py:96-105 (pack)
// and utils/preprocess_yolo_predict.py:273-290,345-378 (api mode: virtual zero-padded square).
//
// Work decomposition (HBM-bound byte kernel; no tensor cores):
//   grid = (tiles of all scales of one image, image).  All scales of an image are adjacent
//   in launch order, so the u8 source (6.2 MB at 1080p) is fetched from HBM once and the
//   other scales re-read it from the 126 MB L2.
//   A CTA owns a tw x th tile of one scale's output.  It stages the source row segments it
//   needs into shared memory with one bulk async copy (TMA, cp.async.bulk -> UBLKCP) per row,
//   completion tracked by an mbarrier; ragged/misaligned sources take a scalar staging path.
//   Each thread then produces 2 adjacent output pixels x th rows: horizontal pass with
//   16x8-bit dot products (IDP.2A) out of shared memory, vertical pass in registers, one
//   32-bit store per channel row (a warp writes one full 128 B line of the fp16 plane).
#include "common.cuh"

namespace {

constexpr int PYR_THREADS = 128;
constexpr int PYR_MAX_TH = 16;
constexpr int PYR_SMEM_BUDGET = 64 * 1024;  // per CTA, keeps >= 3 CTAs per SM

struct PyrScale {
    void *out;
    long long img_stride, cs, ys, xs;  // element strides
    int dst_h, dst_w, unpad_h, unpad_w, top, left;
    int dtype;
    int tile_begin, tiles_x;
    int tw, th;        // output tile
    int dense_rows;    // row slots cover a contiguous source-row range (scale_y <= 2)
    int n_slots;       // row slots per tile
    int slot_bytes;    // multiple of 16
    double scale_x, scale_y;  // cv2: 1.0 / (dst / src) over the virtual extent
};

struct PyrParams {
    const uint8_t *src;
    long long img_stride, pitch;
    int src_h, src_w, virt_h, virt_w;
    int n_scales, flip, border, vec_ok;
    PyrScale sc[FDMS_MAX_SCALES];
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
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
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// OpenCV resize.cpp coefficient set-up for INTER_LINEAR:
//   f = (float)((d + 0.5) * scale - 0.5); s = floor(f); f -= s
__device__ __forceinline__ void lin_coeff(int d, double scale, int &s, float &f) {
    double t = __dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5);
    float ff = __double2float_rn(t);
    int si = (int)floorf(ff);
    s = si;
    f = __fsub_rn(ff, (float)si);
}
__device__ __forceinline__ unsigned pack_w(float f) {  // (short)cvRound((1-f)*2048) | cvRound(f*2048) << 16
    int w0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
    int w1 = __float2int_rn(__fmul_rn(f, 2048.f));
    return (unsigned)w0 | ((unsigned)w1 << 16);
}
// x axis: taps past either edge get weight 0 and the base clamps (HResize set-up)
__device__ __forceinline__ void coeff_x(int d, double scale, int n_src, int &s, unsigned &a01) {
    float f;
    lin_coeff(d, scale, s, f);
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= n_src - 1) { f = 0.f; s = n_src - 1; }
    a01 = pack_w(f);
}
// y axis: weights are kept, the two source ROWS are clamped (resizeGeneric_Invoker)
__device__ __forceinline__ void coeff_y(int d, double scale, int n_src, int &r0, int &r1, unsigned &b01) {
    int s;
    float f;
    lin_coeff(d, scale, s, f);
    r0 = min(max(s, 0), n_src - 1);
    r1 = min(max(s + 1, 0), n_src - 1);
    b01 = pack_w(f);
}

// Horizontal pass for one output pixel out of a staged row: o = byte offset of source pixel
// sx inside the slot; the pixel and its right neighbour are the 6 bytes at o.  Returns
// (S[sx]*a0 + S[sx+1]*a1) >> 4 per channel (15 bits), source channel order.
__device__ __forceinline__ void hpass(const uint8_t *slot, int o, unsigned a01, int (&r)[3]) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(slot + (o & ~3));
    unsigned sh = (unsigned)(o & 3) * 8u;
    uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
    uint32_t lo = __funnelshift_r(w0, w1, sh);  // c0 c1 c2 n0
    uint32_t hi = __funnelshift_r(w1, w2, sh);  // n1 n2 .. ..
    uint32_t p01 = __byte_perm(lo, hi, 0x4130);  // c0 n0 c1 n1
    uint32_t p2 = __byte_perm(lo, hi, 0x0052);   // c2 n2
    r[0] = (int)(__dp2a_lo(a01, p01, 0u) >> 4);
    r[1] = (int)(__dp2a_hi(a01, p01, 0u) >> 4);
    r[2] = (int)(__dp2a_lo(a01, p2, 0u) >> 4);
}

__device__ __forceinline__ int vpass(int r0, int r1, int b0, int b1) {
    int v = (((b0 * r0) >> 16) + ((b1 * r1) >> 16) + 2) >> 2;
    return min(v, 255);
}

template <typename T> struct OutCvt;
template <> struct OutCvt<__half> {
    // rn16(rn32(v / 255)): ATen computes half division in float and rounds once (msfd:101-102)
    static __device__ __forceinline__ __half cvt(int v) { return __float2half_rn(__fdiv_rn((float)v, 255.0f)); }
};
template <> struct OutCvt<float> {
    static __device__ __forceinline__ float cvt(int v) { return __fdiv_rn((float)v, 255.0f); }  // true division
};
template <> struct OutCvt<uint8_t> {
    static __device__ __forceinline__ uint8_t cvt(int v) { return (uint8_t)v; }
};

template <typename T>
__device__ __forceinline__ void store_pair(T *p, long long xs, T a, T b, bool vec) {
    p[0] = a;
    p[xs] = b;
}
template <>
__device__ __forceinline__ void store_pair<__half>(__half *p, long long xs, __half a, __half b, bool vec) {
    if (vec) {
        *reinterpret_cast<__half2 *>(p) = __halves2half2(a, b);
    } else {
        p[0] = a;
        p[xs] = b;
    }
}
template <>
__device__ __forceinline__ void store_pair<float>(float *p, long long xs, float a, float b, bool vec) {
    if (vec) {
        *reinterpret_cast<float2 *>(p) = make_float2(a, b);
    } else {
        p[0] = a;
        p[xs] = b;
    }
}

template <typename T>
__device__ __forceinline__ void run_tile(const PyrParams &P, const PyrScale &S, int tile, int img, uint8_t *smem,
                                         const T *lut) {
    __shared__ int s_r0[PYR_MAX_TH], s_r1[PYR_MAX_TH], s_slot0[PYR_MAX_TH], s_slot1[PYR_MAX_TH];
    __shared__ unsigned s_b[PYR_MAX_TH];
    __shared__ int s_slot_row[2 * PYR_MAX_TH + 4];
    __shared__ int s_misc[4];  // n_used_slots, g0, row_bytes, n_bulk_rows
    __shared__ __align__(8) uint64_t s_bar;

    const int tid = threadIdx.x;
    const int ty = tile / S.tiles_x, tx = tile - ty * S.tiles_x;
    const int oy0 = ty * S.th, ox0 = tx * S.tw;
    const int y_end = min(oy0 + S.th, S.dst_h), x_end = min(ox0 + S.tw, S.dst_w);
    // interior (resized image) part of this tile, in resized-image coordinates
    const int dya = max(oy0, S.top) - S.top, dyb = min(y_end, S.top + S.unpad_h) - S.top;      // [dya, dyb)
    const int dxa = max(ox0, S.left) - S.left, dxb = min(x_end, S.left + S.unpad_w) - S.left;  // [dxa, dxb)
    const bool has_interior = dya < dyb && dxa < dxb;

    T *out = reinterpret_cast<T *>(S.out) + (long long)img * S.img_stride;
    const T border = lut[P.border];
    const bool vec = (S.xs == 1) && ((S.dst_w & 1) == 0) && ((S.cs & 1) == 0) && ((S.img_stride & 1) == 0) &&
                     sizeof(T) > 1;
    const int x = ox0 + 2 * tid;  // this thread's first pixel

    if (has_interior) {
        const uint8_t *src = P.src + (long long)img * P.img_stride;
        if (tid < S.th) {
            int dy = oy0 + tid - S.top;
            int r0 = 0, r1 = 0;
            unsigned b01 = 0;
            if (dy >= 0 && dy < S.unpad_h && oy0 + tid < S.dst_h) coeff_y(dy, S.scale_y, P.virt_h, r0, r1, b01);
            s_r0[tid] = r0;
            s_r1[tid] = r1;
            s_b[tid] = b01;
        }
        if (tid == 0) {
            mbar_init(&s_bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            // slot assignment + column range
            int ja = dya + S.top - oy0, jb = dyb + S.top - oy0;  // valid tile rows [ja, jb)
            int n_used = 0;
            if (S.dense_rows) {
                int rmin = s_r0[ja], rmax = s_r1[jb - 1];
                n_used = min(rmax - rmin + 1, S.n_slots);
                for (int k = 0; k < n_used; k++) s_slot_row[k] = rmin + k;
                for (int j = ja; j < jb; j++) {
                    s_slot0[j] = s_r0[j] - rmin;
                    s_slot1[j] = s_r1[j] - rmin;
                }
            } else {
                for (int j = ja; j < jb; j++) {
                    int b1 = (int)(s_b[j] >> 16);
                    s_slot0[j] = n_used;
                    s_slot_row[n_used++] = s_r0[j];
                    if (b1 != 0) {
                        s_slot1[j] = n_used;
                        s_slot_row[n_used++] = s_r1[j];
                    } else {
                        s_slot1[j] = s_slot0[j];
                    }
                }
            }
            int sxa, sxb;
            unsigned dummy;
            coeff_x(dxa, S.scale_x, P.virt_w, sxa, dummy);
            coeff_x(dxb - 1, S.scale_x, P.virt_w, sxb, dummy);
            sxb = min(sxb + 1, P.virt_w - 1);
            int g0 = (3 * sxa) & ~15;
            int g1 = (3 * sxb + 3 + 15) & ~15;
            s_misc[0] = n_used;
            s_misc[1] = g0;
            s_misc[2] = g1 - g0;
        }
        __syncthreads();
        const int n_used = s_misc[0], g0 = s_misc[1], row_bytes = s_misc[2];
        // ---- stage source rows ---------------------------------------------------------------
        const int real_end = 3 * P.src_w;  // bytes of real pixels in a row; beyond: virtual zeros
        int bulk_bytes = 0;
        if (P.vec_ok) bulk_bytes = max(0, min(g0 + row_bytes, real_end & ~15) - g0);
        if (tid < 32) {
            // one bulk copy per staged row, issued by warp 0
            int n_bulk = 0;
            for (int k = tid; k < n_used; k += 32) {
                int row = s_slot_row[k];
                if (row < P.src_h && bulk_bytes > 0) {
                    bulk_g2s(smem + (size_t)k * S.slot_bytes, src + (long long)row * P.pitch + g0, (uint32_t)bulk_bytes,
                             &s_bar);
                    n_bulk++;
                }
            }
            n_bulk = __reduce_add_sync(0xffffffffu, n_bulk);
            if (tid == 0) mbar_expect_tx(&s_bar, (uint32_t)(n_bulk * bulk_bytes));
        }
        // scalar part: bytes not covered by the bulk copies (ragged tail, misaligned source,
        // virtual zero rows/columns of the api mode)
        {
            const int rest = row_bytes - bulk_bytes;
            for (int k = 0; k < n_used; k++) {
                int row = s_slot_row[k];
                uint8_t *dst = smem + (size_t)k * S.slot_bytes;
                if (row >= P.src_h) {
                    for (int i = tid * 4; i < row_bytes; i += PYR_THREADS * 4) *reinterpret_cast<uint32_t *>(dst + i) = 0u;
                } else {
                    const uint8_t *srow = src + (long long)row * P.pitch;
                    for (int i = tid; i < rest; i += PYR_THREADS) {
                        int pos = g0 + bulk_bytes + i;
                        dst[bulk_bytes + i] = pos < real_end ? __ldg(srow + pos) : (uint8_t)0;
                    }
                }
            }
        }
        mbar_wait(&s_bar, 0);
        __syncthreads();

        // ---- per-thread x coefficients ----------------------------------------------------------
        int oA = 0, oB = 0;
        unsigned aA = 0, aB = 0;
        const int dxA = x - S.left, dxB = x + 1 - S.left;
        const bool inA = x < x_end && dxA >= 0 && dxA < S.unpad_w;
        const bool inB = x + 1 < x_end && dxB >= 0 && dxB < S.unpad_w;
        if (inA) {
            int sx;
            coeff_x(dxA, S.scale_x, P.virt_w, sx, aA);
            oA = 3 * sx - g0;
        }
        if (inB) {
            int sx;
            coeff_x(dxB, S.scale_x, P.virt_w, sx, aB);
            oB = 3 * sx - g0;
        }
        // ---- rows -------------------------------------------------------------------------------
        int c0 = -1, c1 = -1;  // cached slots
        int A0[3] = {0, 0, 0}, A1[3] = {0, 0, 0}, B0[3] = {0, 0, 0}, B1[3] = {0, 0, 0};
        if (x < x_end) {
            for (int j = 0; j < S.th; j++) {
                const int oy = oy0 + j;
                if (oy >= S.dst_h) break;
                T vA[3], vB[3];
                const int dy = oy - S.top;
                if (dy < 0 || dy >= S.unpad_h || (!inA && !inB)) {
                    vA[0] = vA[1] = vA[2] = vB[0] = vB[1] = vB[2] = border;
                } else {
                    const int k0 = s_slot0[j], k1 = s_slot1[j];
                    const unsigned b01 = s_b[j];
                    const int b0 = (int)(b01 & 0xffffu), b1 = (int)(b01 >> 16);
                    int nA0[3], nB0[3], nA1[3], nB1[3];
                    // row 0 (uniform branches: k0/k1/c0/c1 are the same for the whole CTA)
                    if (k0 == c0) {
#pragma unroll
                        for (int c = 0; c < 3; c++) { nA0[c] = A0[c]; nB0[c] = B0[c]; }
                    } else if (k0 == c1) {
#pragma unroll
                        for (int c = 0; c < 3; c++) { nA0[c] = A1[c]; nB0[c] = B1[c]; }
                    } else {
                        const uint8_t *slot = smem + (size_t)k0 * S.slot_bytes;
                        if (inA) hpass(slot, oA, aA, nA0);
                        if (inB) hpass(slot, oB, aB, nB0);
                    }
                    if (b1 == 0 || k1 == k0) {
#pragma unroll
                        for (int c = 0; c < 3; c++) { nA1[c] = nA0[c]; nB1[c] = nB0[c]; }
                    } else if (k1 == c1) {
#pragma unroll
                        for (int c = 0; c < 3; c++) { nA1[c] = A1[c]; nB1[c] = B1[c]; }
                    } else if (k1 == c0) {
#pragma unroll
                        for (int c = 0; c < 3; c++) { nA1[c] = A0[c]; nB1[c] = B0[c]; }
                    } else {
                        const uint8_t *slot = smem + (size_t)k1 * S.slot_bytes;
                        if (inA) hpass(slot, oA, aA, nA1);
                        if (inB) hpass(slot, oB, aB, nB1);
                    }
#pragma unroll
                    for (int c = 0; c < 3; c++) { A0[c] = nA0[c]; A1[c] = nA1[c]; B0[c] = nB0[c]; B1[c] = nB1[c]; }
                    c0 = k0;
                    c1 = (b1 == 0) ? k0 : k1;
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        vA[c] = inA ? lut[vpass(A0[c], A1[c], b0, b1)] : border;
                        vB[c] = inB ? lut[vpass(B0[c], B1[c], b0, b1)] : border;
                    }
                }
                T *row = out + (long long)oy * S.ys + (long long)x * S.xs;
                const bool two = x + 1 < x_end;
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const int sc = P.flip ? 2 - c : c;
                    T *p = row + (long long)c * S.cs;
                    if (two) store_pair<T>(p, S.xs, vA[sc], vB[sc], vec);
                    else p[0] = vA[sc];
                }
            }
        }
    } else {
        // pure border tile
        if (x < x_end) {
            const bool two = x + 1 < x_end;
            for (int oy = oy0; oy < y_end; oy++) {
                T *row = out + (long long)oy * S.ys + (long long)x * S.xs;
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    T *p = row + (long long)c * S.cs;
                    if (two) store_pair<T>(p, S.xs, border, border, vec);
                    else p[0] = border;
                }
            }
        }
    }
}

__global__ void __launch_bounds__(PYR_THREADS) pyramid_kernel(const __grid_constant__ PyrParams P) {
    extern __shared__ __align__(128) uint8_t dyn_smem[];
    __shared__ __align__(16) uint8_t s_lut_raw[256 * 4];
    int tile = blockIdx.x;
    int si = 0;
#pragma unroll 1
    for (int s = 1; s < P.n_scales; s++)
        if (tile >= P.sc[s].tile_begin) si = s;
    const PyrScale &S = P.sc[si];
    tile -= S.tile_begin;
    // value LUT: out = T(v / 255) for every byte value (exact division once per CTA)
    if (S.dtype == FDMS_F16) {
        __half *lut = reinterpret_cast<__half *>(s_lut_raw);
        for (int i = threadIdx.x; i < 256; i += PYR_THREADS) lut[i] = OutCvt<__half>::cvt(i);
        __syncthreads();
        run_tile<__half>(P, S, tile, blockIdx.y, dyn_smem, lut);
    } else if (S.dtype == FDMS_F32) {
        float *lut = reinterpret_cast<float *>(s_lut_raw);
        for (int i = threadIdx.x; i < 256; i += PYR_THREADS) lut[i] = OutCvt<float>::cvt(i);
        __syncthreads();
        run_tile<float>(P, S, tile, blockIdx.y, dyn_smem, lut);
    } else {
        uint8_t *lut = s_lut_raw;
        for (int i = threadIdx.x; i < 256; i += PYR_THREADS) lut[i] = (uint8_t)i;
        __syncthreads();
        run_tile<uint8_t>(P, S, tile, blockIdx.y, dyn_smem, lut);
    }
}

}  // namespace

extern "C" FDMS_API int fdms_pyramid(fdms_ctx *ctx, const uint8_t *src, int n_img, int64_t img_stride,
                                     int64_t row_pitch, int src_h, int src_w, int virt_h, int virt_w,
                                     const fdms_scale_out *scales, int n_scales, int flip_rb, int border,
                                     fdms_stream stream) {
    if (int rc = fdms_bind(ctx)) return rc;
    if (!src || !scales || n_img <= 0 || n_scales <= 0 || n_scales > FDMS_MAX_SCALES || src_h <= 0 || src_w <= 0)
        return fdms_fail(ctx, FDMS_ERR_INVALID, "fdms_pyramid: bad arguments");
    if (n_img > 65535) return fdms_fail(ctx, FDMS_ERR_UNSUPPORTED, "fdms_pyramid: n_img > 65535");
    if (virt_h < src_h || virt_w < src_w || row_pitch < (int64_t)src_w * 3 || border < 0 || border > 255)
        return fdms_fail(ctx, FDMS_ERR_INVALID, "fdms_pyramid: bad extents");
    PyrParams P;
    memset(&P, 0, sizeof(P));
    P.src = src;
    P.img_stride = img_stride;
    P.pitch = row_pitch;
    P.src_h = src_h;
    P.src_w = src_w;
    P.virt_h = virt_h;
    P.virt_w = virt_w;
    P.n_scales = n_scales;
    P.flip = flip_rb ? 1 : 0;
    P.border = border;
    P.vec_ok = ((uintptr_t)src % 16 == 0) && (img_stride % 16 == 0) && (row_pitch % 16 == 0);
    int tiles = 0, smem_max = 0;
    for (int s = 0; s < n_scales; s++) {
        const fdms_scale_out &in = scales[s];
        const fdms_geom &g = in.geom;
        if (!in.out || g.dst_h <= 0 || g.dst_w <= 0 || g.unpad_h <= 0 || g.unpad_w <= 0 || g.top < 0 || g.left < 0 ||
            g.top + g.unpad_h > g.dst_h || g.left + g.unpad_w > g.dst_w)
            return fdms_fail(ctx, FDMS_ERR_INVALID, "fdms_pyramid: bad geometry for scale %d", s);
        if (in.out_dtype != FDMS_U8 && in.out_dtype != FDMS_F16 && in.out_dtype != FDMS_F32)
            return fdms_fail(ctx, FDMS_ERR_INVALID, "fdms_pyramid: bad out_dtype for scale %d", s);
        PyrScale &S = P.sc[s];
        S.out = in.out;
        S.img_stride = in.out_img_stride;
        if (in.out_layout == FDMS_CHW) {
            S.cs = (long long)g.dst_h * g.dst_w;
            S.ys = g.dst_w;
            S.xs = 1;
        } else {
            S.cs = 1;
            S.ys = 3LL * g.dst_w;
            S.xs = 3;
        }
        S.dst_h = g.dst_h;
        S.dst_w = g.dst_w;
        S.unpad_h = g.unpad_h;
        S.unpad_w = g.unpad_w;
        S.top = g.top;
        S.left = g.left;
        S.dtype = in.out_dtype;
        // cv2.resize: scale = 1.0 / (dsize / ssize), computed in double (inv_scale first)
        S.scale_x = 1.0 / ((double)g.unpad_w / virt_w);
        S.scale_y = 1.0 / ((double)g.unpad_h / virt_h);
        // tile shape: widest tile whose staged rows fit the shared-memory budget
        int tw = 2 * PYR_THREADS, th = 8;
        for (;;) {
            S.dense_rows = S.scale_y <= 2.0;
            S.n_slots = S.dense_rows ? (int)((th - 1) * S.scale_y) + 4 : 2 * th;
            if (S.n_slots > 2 * PYR_MAX_TH + 4) S.n_slots = 2 * PYR_MAX_TH + 4;
            int span_px = (int)(tw * S.scale_x) + 4;
            S.slot_bytes = ((span_px * 3 + 15) & ~15) + 48;
            if ((long long)S.n_slots * S.slot_bytes <= PYR_SMEM_BUDGET) break;
            if (th > 1) th >>= 1;
            else if (tw > 32) tw >>= 1;
            else break;
        }
        if ((long long)S.n_slots * S.slot_bytes > 200 * 1024)
            return fdms_fail(ctx, FDMS_ERR_UNSUPPORTED, "fdms_pyramid: scale %d needs %lld B of shared memory", s,
                             (long long)S.n_slots * S.slot_bytes);
        S.tw = tw;
        S.th = th;
        S.tiles_x = ceil_div(g.dst_w, tw);
        S.tile_begin = tiles;
        tiles += S.tiles_x * ceil_div(g.dst_h, th);
        if (S.n_slots * S.slot_bytes > smem_max) smem_max = S.n_slots * S.slot_bytes;
    }
    if (!ctx->pyr_attr) {  // dynamic shared memory opt-in, once per ctx
        FDMS_CUDA(ctx, cudaFuncSetAttribute(pyramid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        ctx->pyr_attr = true;
    }
    dim3 grid(tiles, n_img);
    pyramid_kernel<<<grid, PYR_THREADS, smem_max, (cudaStream_t)stream>>>(P);
    FDMS_LAUNCH_CHECK(ctx, "pyramid_kernel");
    return FDMS_OK;
}
