// Fused letterbox pyramid: bilinear resize (OpenCV INTER_LINEAR u8 fixed point, bit exact),
// constant border, channel flip, /255 and pack -- every image x every scale in ONE launch.
//
// Replaces utils/datasets.py:873-903 (letterbox), multi_scale_face_detector.py:96-105 (pack)
// and utils/preprocess_yolo_predict.py:273-290,345-378 (api mode: virtual zero-padded square).
//
// Work decomposition (HBM-bound byte kernel; no tensor cores):
//   grid = (tiles of all scales of one image, image).  All scales of an image are adjacent
//   in launch order, so the u8 source (6.2 MB at 1080p) is fetched from HBM once and the
//   other scales re-read it from the 126 MB L2.  A CTA (160 threads, each owning 2 adjacent
//   output columns) owns a tw x th tile of one scale's output; the host picks one of two
//   tile kinds per scale:
//   * staged tiles (any scale factor).  Warp 0 derives the vertical taps of the tile's rows,
//     assigns the source rows to shared-memory slots and stages every slot with ONE bulk async
//     copy (TMA, cp.async.bulk -> UBLKCP) completing on an mbarrier; ragged / misaligned /
//     virtual-zero sources take a scalar staging path.  Threads stream over the slots in
//     source-row order: horizontal pass of the slot for their two pixels (16x8-bit dot
//     products, IDP.2A) into registers, then every output row whose lower tap is that slot is
//     emitted (vertical pass, /255, pack, one 32-bit store per channel: a warp writes 128
//     contiguous bytes of a plane).  Horizontal results are computed once per (source row,
//     column) even when output rows share source rows (scale < 2).
//   * direct tiles (exact factors k:1 and 3:2 with rows-only padding).  The taps are an affine /
//     periodic function of the output position, a thread's window of every source row is a few
//     aligned words next to its neighbours' windows, so plain loads coalesce: no shared memory,
//     no barrier, no tap prologue, several rows of loads in flight per thread.
#include "common.cuh"
#include "pyramid.cuh"

namespace {

constexpr int PYR_THREADS = 160;  // 2 columns per thread: 320-column tiles divide 640/960/1280 evenly
#ifndef PYR_UNROLL32
#define PYR_UNROLL32 2  // row pairs of the 3:2 direct tile whose loads are in flight together
#endif
#ifndef PYR_UNROLL_INT
#define PYR_UNROLL_INT 4  // rows of the integer-factor direct tiles in flight together
#endif
#ifndef PYR_MIN_CTAS
#define PYR_MIN_CTAS 8  // resident CTAs per SM the register allocation of the fp16 instantiation is held to (48 registers)
#endif
constexpr int K_UNROLL32 = PYR_UNROLL32, K_UNROLL_INT = PYR_UNROLL_INT;
constexpr int PYR_MAX_TH = 32;    // rows per tile are handled by the lanes of warp 0
constexpr int PYR_MAX_SLOTS = 64;
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
__device__ __forceinline__ void coeff_y(int d, double scale, int n_src, int &r0, int &r1, unsigned &b01, int &s) {
    float f;
    lin_coeff(d, scale, s, f);
    r0 = min(max(s, 0), n_src - 1);
    r1 = min(max(s + 1, 0), n_src - 1);
    b01 = pack_w(f);
}

// Horizontal pass of one output pixel out of a staged row.  `w` = shared-memory address of the
// aligned word holding the first byte of source pixel sx, sh = 8 * (byte offset in that word);
// the pixel and its right neighbour are the 6 bytes from there.  h[c] = (S[sx][c]*a0 +
// S[sx+1][c]*a1) >> 4 (15 bits, source channel order).
__device__ __forceinline__ void hpass(uint32_t w, unsigned sh, unsigned a01, int *h) {
    uint32_t w0, w1, w2;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(w));
    asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(w1) : "r"(w));
    asm volatile("ld.shared.u32 %0, [%1+8];" : "=r"(w2) : "r"(w));
    uint32_t lo = __funnelshift_r(w0, w1, sh);   // c0 c1 c2 n0
    uint32_t hi = __funnelshift_r(w1, w2, sh);   // n1 n2 .. ..
    uint32_t p01 = __byte_perm(lo, hi, 0x4130);  // c0 n0 c1 n1
    uint32_t p2 = __byte_perm(lo, hi, 0x0052);   // c2 n2
    h[0] = (int)(__dp2a_lo(a01, p01, 0u) >> 4);
    h[1] = (int)(__dp2a_hi(a01, p01, 0u) >> 4);
    h[2] = (int)(__dp2a_lo(a01, p2, 0u) >> 4);
}

// PM_POINT: the output pixel is source pixel sx; PRMT drops byte c straight into the 2^23 magic number, so the
// three values leave as float bits (2^23 + v).
__device__ __forceinline__ void hpass_point(uint32_t w, unsigned sh, unsigned *h) {
    uint32_t w0, w1;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(w));
    asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(w1) : "r"(w));
    const uint32_t lo = __funnelshift_r(w0, w1, sh);  // c0 c1 c2 ..
    h[0] = __byte_perm(lo, 0x4B000000u, 0x7650);
    h[1] = __byte_perm(lo, 0x4B000000u, 0x7651);
    h[2] = __byte_perm(lo, 0x4B000000u, 0x7652);
}
// PM_HALF: h[c] = S[sx][c] + S[sx+1][c]
__device__ __forceinline__ void hpass_half(uint32_t w, unsigned sh, int *h) {
    uint32_t w0, w1, w2;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(w));
    asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(w1) : "r"(w));
    asm volatile("ld.shared.u32 %0, [%1+8];" : "=r"(w2) : "r"(w));
    const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
    const uint32_t p01 = __byte_perm(lo, hi, 0x4130), p2 = __byte_perm(lo, hi, 0x0052);
    h[0] = (int)__dp2a_lo(0x00010001u, p01, 0u);
    h[1] = (int)__dp2a_hi(0x00010001u, p01, 0u);
    h[2] = (int)__dp2a_lo(0x00010001u, p2, 0u);
}

// Vertical pass.  OpenCV: (((b0*(R0>>4))>>16) + ((b1*(R1>>4))>>16) + 2) >> 2.  Returns the sum t before the
// final shift: two full-rate IMADs, one PRMT gathering the two high halves, one IDP.2A adding them (+2).
// t >> 2 cannot exceed 255: b0+b1 <= 2049 and R>>4 <= 32654, so the two floors sum to at most 1020.
__device__ __forceinline__ unsigned vsum(unsigned r0, unsigned r1, unsigned b0, unsigned b1) {
    const unsigned p0 = b0 * r0, p1 = b1 * r1;            // < 2^27
    const unsigned hh = __byte_perm(p0, p1, 0x7632);      // (p0 >> 16) | (p1 >> 16) << 16
    return __dp2a_lo(hh, 0x0101u, 2u);                    // (p0 >> 16) + (p1 >> 16) + 2
}

// ---- value conversion + paired store, per output type.  Inputs are the sums t (pixel = t >> 2). ----
template <typename T> struct Out;
template <> struct Out<__half> {
    // rn16(rn32(v/255)) (ATen divides in float and rounds once, msfd:101-102) == rn16(rn32(v * rn32(1/255)))
    // for all 256 byte values (checked exhaustively in tests/test_host_logic.py).  float(v) comes
    // from the 2^23 magic number: (t >> 2) + 0x4B000000 is one mad.hi, and
    // fma(as_float(0x4B000000 + v), c, -2^23*c) is exactly rn32(v*c).
    static __device__ __forceinline__ float f(unsigned t) {
        const float c = 0.003921568859368563f;  // rn32(1/255)
        const unsigned bits = (t >> 2) + 0x4B000000u;  // LEA.HI
        return __fmaf_rn(__uint_as_float(bits), c, -8388608.0f * c);
    }
    static __device__ __forceinline__ void store2(__half *p, long long xs, unsigned a, unsigned b, bool vec) {
        if (vec) {
            *reinterpret_cast<__half2 *>(p) = __floats2half2_rn(f(a), f(b));
        } else {
            p[0] = __float2half_rn(f(a));
            p[xs] = __float2half_rn(f(b));
        }
    }
    static __device__ __forceinline__ void store1(__half *p, unsigned a) { p[0] = __float2half_rn(f(a)); }
    // value already in magic-number form (float bits of 2^23 + v)
    static __device__ __forceinline__ void store2_bits(__half *p, unsigned a, unsigned b) {
        const float c = 0.003921568859368563f;
        *reinterpret_cast<__half2 *>(p) = __floats2half2_rn(__fmaf_rn(__uint_as_float(a), c, -8388608.0f * c),
                                                            __fmaf_rn(__uint_as_float(b), c, -8388608.0f * c));
    }
};
template <> struct Out<float> {
    static __device__ __forceinline__ void store2_bits(float *p, unsigned a, unsigned b) {
        *reinterpret_cast<float2 *>(p) = make_float2(f((a & 0xffu) << 2), f((b & 0xffu) << 2));
    }
    static __device__ __forceinline__ float f(unsigned t) { return __fdiv_rn((float)(t >> 2), 255.0f); }  // true division
    static __device__ __forceinline__ void store2(float *p, long long xs, unsigned a, unsigned b, bool vec) {
        if (vec) {
            *reinterpret_cast<float2 *>(p) = make_float2(f(a), f(b));
        } else {
            p[0] = f(a);
            p[xs] = f(b);
        }
    }
    static __device__ __forceinline__ void store1(float *p, unsigned a) { p[0] = f(a); }
};
// fp32 output as a CUDA device computes it (FDMS_NUMERICS_ATEN_CUDA): ATen's `img /= 255.0` on a CUDA tensor is
// img * (1 / 255) with the reciprocal rounded to fp32 (one bit away from the division for some bytes).  A distinct
// element type, so that the choice costs the other instantiations nothing.
struct f32r {
    float v;
};
template <> struct Out<f32r> {
    static __device__ __forceinline__ float f(unsigned t) { return __fmul_rn((float)(t >> 2), 0.003921568859368563f); }
    static __device__ __forceinline__ void store2_bits(f32r *p, unsigned a, unsigned b) {
        *reinterpret_cast<float2 *>(p) = make_float2(f((a & 0xffu) << 2), f((b & 0xffu) << 2));
    }
    static __device__ __forceinline__ void store2(f32r *p, long long xs, unsigned a, unsigned b, bool vec) {
        if (vec) {
            *reinterpret_cast<float2 *>(p) = make_float2(f(a), f(b));
        } else {
            p[0].v = f(a);
            p[xs].v = f(b);
        }
    }
    static __device__ __forceinline__ void store1(f32r *p, unsigned a) { p[0].v = f(a); }
};
template <> struct Out<uint8_t> {
    static __device__ __forceinline__ void store2_bits(uint8_t *p, unsigned a, unsigned b) {
        p[0] = (uint8_t)a;
        p[1] = (uint8_t)b;
    }
    static __device__ __forceinline__ void store2(uint8_t *p, long long xs, unsigned a, unsigned b, bool) {
        p[0] = (uint8_t)(a >> 2);
        p[xs] = (uint8_t)(b >> 2);
    }
    static __device__ __forceinline__ void store1(uint8_t *p, unsigned a) { p[0] = (uint8_t)(a >> 2); }
};

struct TileShared {
    // per row: x = slot of the row's lower tap -- the row is emitted once that slot's horizontal pass is done
    // (-2: border row, -4: virtual zero row), y / z = weights of the slot before / of that slot; one sentinel entry
    // past the last row.  A clamped row (both taps on the first / last source row) gets that row staged twice and a
    // row with a zero lower weight is emitted from its upper tap alone with weights (0, b0), so the two taps of a
    // row are ALWAYS the previous and the current slot.
    int4 rowmeta[PYR_MAX_TH + 1];
    int slot_row[PYR_MAX_SLOTS];
    // sources whose rows do not start on 16 bytes: a slot is staged from the 16-byte boundary below its first byte,
    // so its data starts slot_shift bytes into the slot; the bulk copy covers the first slot_cov bytes of the slot.
    // The two arrays live outside this struct (only the SHIFTED kernel allocates them: static shared memory of the
    // common kernel decides its L1 carve-out).
    int *slot_shift, *slot_cov;
    int n_used, g0, row_bytes, bulk_bytes, need_scalar, n_rows, tail_max;
    alignas(8) unsigned long long bar;
};

// One emitted output row for this thread's two pixels.  R0/R1: horizontal results of the upper /
// lower source row, [0..2] pixel A, [3..5] pixel B.  pc[c] + off addresses channel c of pixel A.
template <typename T, bool EDGE>
__device__ __forceinline__ void emit_row(T *const *pc, unsigned off, long long xs, unsigned b0s, unsigned b1s,  // plain weights
                                         const int *R0, const int *R1, int flip, unsigned tborder, bool inA, bool inB,
                                         bool two, bool vec) {
    unsigned tA[3], tB[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        tA[c] = vsum((unsigned)R0[c], (unsigned)R1[c], b0s, b1s);
        tB[c] = vsum((unsigned)R0[3 + c], (unsigned)R1[3 + c], b0s, b1s);
        if (EDGE) {
            if (!inA) tA[c] = tborder;
            if (!inB) tB[c] = tborder;
        }
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const unsigned a = flip ? tA[2 - c] : tA[c], b = flip ? tB[2 - c] : tB[c];
        T *p = pc[c] + off;
        if (!EDGE || two) Out<T>::store2(p, xs, a, b, vec);
        else Out<T>::store1(p, a);
    }
}

template <typename T>
__device__ __forceinline__ void border_row(T *const *pc, unsigned off, long long xs, unsigned tborder, bool two,
                                           bool vec) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
        T *p = pc[c] + off;
        if (two) Out<T>::store2(p, xs, tborder, tborder, vec);
        else Out<T>::store1(p, tborder);
    }
}

// The per-thread streaming loop of pyramid_kernel: horizontal pass of each slot in source-row order, then every
// output row whose lower tap is that slot.  FAST = interior tile with paired stores (no per-pixel predicates).
template <typename T, bool FAST, int MODE, bool SHIFT = false>
__device__ __forceinline__ void stream_tile(const TileShared &sh, T *const *pc, unsigned ys, long long xs_, uint32_t wA,
                                            unsigned shA, unsigned aA, uint32_t wB, unsigned shB, unsigned aB,
                                            int n_used, int n_rows, int slot_bytes, unsigned tborder, bool inA_,
                                            bool inB_, bool two_, bool vec_) {
    const long long xs = FAST ? 1 : xs_;
    const bool inA = FAST ? true : inA_, inB = FAST ? true : inB_, two = FAST ? true : two_, vec = FAST ? true : vec_;
    const bool edge = !(inA && inB);
    unsigned off = 0;
    int j = 0;
    const int4 *rmp = sh.rowmeta;
    int4 rm = rmp[0];

    if (MODE == PM_POINT) {  // one slot per interior row; border rows interleave
        for (int k = 0;; k++) {
            while (rm.x < 0) {  // border rows due before slot k
                border_row<T>(pc, off, xs, rm.x == -4 ? 0u : tborder, true, true);
                off += ys;
                ++j;
                rm = *++rmp;
            }
            if (k >= n_used) break;
            unsigned hA[3], hB[3];
            hpass_point(wA + (uint32_t)k * slot_bytes, shA, hA);
            hpass_point(wB + (uint32_t)k * slot_bytes, shB, hB);
#pragma unroll
            for (int c = 0; c < 3; c++) Out<T>::store2_bits(pc[c] + off, hA[c], hB[c]);
            off += ys;
            ++j;
            rm = *++rmp;
        }
        return;
    }
    int E[6] = {0, 0, 0, 0, 0, 0}, O[6] = {0, 0, 0, 0, 0, 0};

    // emit every row whose lower tap is slot k; `cur` = horizontal results of slot k, `other` = slot k-1
    auto flush = [&](int k, int *cur, int *other) {
        while (rm.x <= k) {
            if (rm.x < 0) {  // -2: letterbox border, -4: virtual zero rows of the api mode
                border_row<T>(pc, off, xs, rm.x == -4 ? 0u : tborder, two, vec);
            } else {
                const unsigned b0s = (unsigned)rm.y, b1s = (unsigned)rm.z;  // weights of slot k-1 / slot k
                if (MODE == PM_HALF) {
#pragma unroll
                    for (int c = 0; c < 3; c++)
                        Out<T>::store2(pc[c] + off, 1, (unsigned)(other[c] + cur[c] + 2), (unsigned)(other[3 + c] + cur[3 + c] + 2),
                                       true);
                } else if (!FAST && edge) {
                    emit_row<T, true>(pc, off, xs, b0s, b1s, other, cur, 0, tborder, inA, inB, two, vec);
                } else {
                    emit_row<T, false>(pc, off, xs, b0s, b1s, other, cur, 0, tborder, true, true, true, vec);
                }
            }
            off += ys;
            j++;
            rm = *++rmp;
        }
    };

    // horizontal pass of slot k for both pixels.  SHIFT: the slot's data starts slot_shift[k] bytes into the slot
    // (rows that do not start on 16 bytes), shA / shB are byte offsets and the word / bit split is redone per slot.
    auto hslot = [&](int k, uint32_t slot_off, int *h) {
        if (SHIFT) {
            const unsigned d = (unsigned)sh.slot_shift[k], oa = shA + d, ob = shB + d;
            hpass(wA + slot_off + (oa & ~3u), (oa & 3u) * 8u, aA, h);
            hpass(wB + slot_off + (ob & ~3u), (ob & 3u) * 8u, aB, h + 3);
        } else if (MODE == PM_HALF) {
            hpass_half(wA + slot_off, shA, h);
            hpass_half(wB + slot_off, shB, h + 3);
        } else {
            hpass(wA + slot_off, shA, aA, h);
            hpass(wB + slot_off, shB, aB, h + 3);
        }
    };
    uint32_t slot_off = 0;
    for (int k = 0; k < n_used; k += 2) {
        hslot(k, slot_off, E);
        flush(k, E, O);
        slot_off += slot_bytes;
        if (k + 1 < n_used) {
            hslot(k + 1, slot_off, O);
            flush(k + 1, O, E);
        }
        slot_off += slot_bytes;
    }
    // rows not tied to any slot (a tile without interior)
    while (j < n_rows) {
        border_row<T>(pc, off, xs, rm.x == -4 ? 0u : tborder, two, vec);
        off += ys;
        j++;
        rm = *++rmp;
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Direct tiles.  With an exact integer factor every tap position is an affine function of the output coordinate,
// a thread's two pixels read a fixed, word-aligned byte window per source row and neighbouring threads read
// neighbouring windows: the loads coalesce by themselves, so these tiles skip the staging (no shared memory, no
// barrier, no tap prologue) and run at much higher occupancy.
//   PM_HALF, kx == 2: pixels A and B are the 12 contiguous bytes [a b | c d] (4 source pixels) at 6 * dx; the sum of
//                     the 2 x 2 block of each channel is a chain of byte dot products (IDP.4A) seeded with the
//                     rounding constant, the second source row accumulating onto the first.
//   PM_POINT:         the pixel is source pixel (kx * dx + kx / 2, ky * dy + ky / 2).
// Only rows are padded here (host guarantees left == 0 and dst_w == unpad_w), so a row is either all border or all
// interior.
// ---------------------------------------------------------------------------------------------------------------
// Column pairs of a direct tile that lie in the left / right letterbox border: rare (portrait sources), kept out of line.
template <typename T>
__device__ __noinline__ void border_columns(T *const *pc, unsigned ys, unsigned tborder, int rows) {
    unsigned off = 0;
    for (int y = 0; y < rows; y++, off += ys) border_row<T>(pc, off, 1, tborder, true, true);
}

// COLS (CTA-uniform): the tile contains letterbox border columns.
template <typename T, int MODE, bool COLS>
__device__ __forceinline__ void direct_tile(const PyrParams &P, const PyrScale &S, int img, int oy0, int ox0, int tid) {
    const int x = ox0 + 2 * tid;
    if (x >= S.dst_w || 2 * tid >= S.tw) return;
    const int4 e = __ldg(reinterpret_cast<const int4 *>(S.xtab + x));
    const int sxA = e.x, sxB = e.z;
    const uint8_t *src = P.src + (long long)img * P.img_stride;
    T *pc[3];
    {
        T *p0 = reinterpret_cast<T *>(S.out) + (long long)img * S.img_stride + (long long)oy0 * S.ys + x;
        pc[0] = P.flip ? p0 + 2 * S.cs : p0;
        pc[1] = p0 + S.cs;
        pc[2] = P.flip ? p0 : p0 + 2 * S.cs;
    }
    const unsigned ys = (unsigned)S.ys;
    const unsigned tborder = (unsigned)P.border * 4u;
    const int y_end = min(oy0 + S.th, S.dst_h);
    if (COLS && sxA < 0) {  // a column pair of the left / right letterbox border (left, unpad_w even: never half a pair)
        border_columns<T>(pc, ys, tborder, y_end - oy0);
        return;
    }
    const int ya = min(max(oy0, S.top), y_end), yz = max(min(y_end, S.top + S.unpad_h), ya);  // interior rows [ya, yz)
    const int yb = max(min(yz, S.top + S.real_rows), ya);  // [ya, yb) have real taps, [yb, yz) are virtual zeros (api mode)
    unsigned off = 0;
    for (int y = oy0; y < ya; y++, off += ys) border_row<T>(pc, off, 1, tborder, true, true);
    const long long pitch = P.pitch;
    if (MODE == PM_HALF) {
        const int ky = S.ky;
        const uint8_t *row = src + 3 * sxA + (long long)(ky * (ya - S.top) + (ky >> 1) - 1) * pitch;
        const long long step = (long long)ky * pitch;
#pragma unroll(K_UNROLL_INT)
        for (int y = ya; y < yb; y++, off += ys, row += step) {
            const uint32_t *p0 = reinterpret_cast<const uint32_t *>(row);
            const uint32_t *p1 = reinterpret_cast<const uint32_t *>(row + pitch);
            const uint32_t w0 = __ldg(p0), w1 = __ldg(p0 + 1), w2 = __ldg(p0 + 2);
            const uint32_t v0 = __ldg(p1), v1 = __ldg(p1 + 1), v2 = __ldg(p1 + 2);
            // bytes: w0 = a0 a1 a2 b0, w1 = b1 b2 c0 c1, w2 = c2 d0 d1 d2
            unsigned tA[3], tB[3];
            tA[0] = __dp4a(v0, 0x01000001u, __dp4a(w0, 0x01000001u, 2u));
            tA[1] = __dp4a(v1, 0x00000001u, __dp4a(v0, 0x00000100u, __dp4a(w1, 0x00000001u, __dp4a(w0, 0x00000100u, 2u))));
            tA[2] = __dp4a(v1, 0x00000100u, __dp4a(v0, 0x00010000u, __dp4a(w1, 0x00000100u, __dp4a(w0, 0x00010000u, 2u))));
            tB[0] = __dp4a(v2, 0x00000100u, __dp4a(v1, 0x00010000u, __dp4a(w2, 0x00000100u, __dp4a(w1, 0x00010000u, 2u))));
            tB[1] = __dp4a(v2, 0x00010000u, __dp4a(v1, 0x01000000u, __dp4a(w2, 0x00010000u, __dp4a(w1, 0x01000000u, 2u))));
            tB[2] = __dp4a(v2, 0x01000001u, __dp4a(w2, 0x01000001u, 2u));
#pragma unroll
            for (int c = 0; c < 3; c++) Out<T>::store2(pc[c] + off, 1, tA[c], tB[c], true);
        }
    } else {
        const int ky = S.ky;
        const int oA = 3 * sxA, oB = 3 * sxB;
        const unsigned shA = (unsigned)(oA & 3) * 8u, shB = (unsigned)(oB & 3) * 8u;
        const bool a2 = (oA & 3) >= 2, b2 = (oB & 3) >= 2;  // the 3 bytes spill into the next word
        const uint8_t *row = src + (long long)(ky * (ya - S.top) + (ky >> 1)) * pitch;
        const long long step = (long long)ky * pitch;
        const int wa = oA & ~3, wb = oB & ~3;
#pragma unroll(K_UNROLL_INT)
        for (int y = ya; y < yb; y++, off += ys, row += step) {
            const uint32_t *pa = reinterpret_cast<const uint32_t *>(row + wa);
            const uint32_t *pb = reinterpret_cast<const uint32_t *>(row + wb);
            const uint32_t a0 = __ldg(pa), a1 = a2 ? __ldg(pa + 1) : 0u;
            const uint32_t b0 = __ldg(pb), b1 = b2 ? __ldg(pb + 1) : 0u;
            const uint32_t lA = __funnelshift_r(a0, a1, shA), lB = __funnelshift_r(b0, b1, shB);
#pragma unroll
            for (int c = 0; c < 3; c++)
                Out<T>::store2_bits(pc[c] + off, __byte_perm(lA, 0x4B000000u, 0x7650 + c), __byte_perm(lB, 0x4B000000u, 0x7650 + c));
        }
    }
    for (int y = yb; y < yz; y++, off += ys) border_row<T>(pc, off, 1, 0u, true, true);  // all taps zero -> pixel 0
    for (int y = yz; y < y_end; y++, off += ys) border_row<T>(pc, off, 1, tborder, true, true);
}


// PM_RATIO32: the exact 3:2 scale (1920x1080 -> 1280x720 and every other 1.5:1 letterbox).  The taps are periodic:
// output columns 2m, 2m+1 read source pixels 3m..3m+2 (nine contiguous bytes at 9m) and output rows 2n, 2n+1 read
// source rows 3n..3n+2, so like the integer factors a thread's window is a fixed affine function of its position --
// three aligned words per source row, no staging, no per-row tap look-up, and all loads of two row pairs are issued
// before the first is reduced.  Per source row the nine bytes are turned into the six horizontal sums with 3 SHF +
// 3 PRMT + 6 IDP.2A (the staged path needs two independent 6-byte windows: 26 instructions).  The WEIGHTS are not
// assumed: the two column phases and the two row phases are read from the same tap tables the general path uses.
template <typename T, bool COLS>
__device__ __forceinline__ void direct_tile_32(const PyrParams &P, const PyrScale &S, int img, int oy0, int ox0, int tid) {
    const int x = ox0 + 2 * tid;
    if (x >= S.dst_w || 2 * tid >= S.tw) return;
    const int4 e = __ldg(reinterpret_cast<const int4 *>(S.xtab + x));  // {3m, a(even), 3m+1, a(odd)}
    const unsigned aA = (unsigned)e.y, aB = (unsigned)e.w;
    const int o = 3 * e.x;  // byte offset of source pixel 3m
    const unsigned sh = (unsigned)(o & 3) * 8u;
    const uint8_t *src = P.src + (long long)img * P.img_stride + (o & ~3);
    T *pc[3];
    {
        T *p0 = reinterpret_cast<T *>(S.out) + (long long)img * S.img_stride + (long long)oy0 * S.ys + x;
        pc[0] = P.flip ? p0 + 2 * S.cs : p0;
        pc[1] = p0 + S.cs;
        pc[2] = P.flip ? p0 : p0 + 2 * S.cs;
    }
    const unsigned ys = (unsigned)S.ys;
    const unsigned tborder = (unsigned)P.border * 4u;
    const int y_end = min(oy0 + S.th, S.dst_h);
    if (COLS && e.x < 0) {  // a column pair of the left / right letterbox border
        border_columns<T>(pc, ys, tborder, y_end - oy0);
        return;
    }
    // interior rows [ya, yz): [ya, yb) have real taps, [yb, yz) read the virtual zeros of the api mode.  top, th and
    // the interior heights are even (host), so every range starts on an even row and holds whole row pairs.
    const int ya = min(max(oy0, S.top), y_end), yz = max(min(y_end, S.top + S.unpad_h), ya);
    const int yb = max(min(yz, S.top + S.real_rows), ya);
    unsigned off = 0;
    for (int y = oy0; y < ya; y++, off += ys) border_row<T>(pc, off, 1, tborder, true, true);
    if (ya < yb) {
        const int4 r0 = __ldg(S.ytab + ya), r1 = __ldg(S.ytab + ya + 1);  // the two row phases: {3n, ., b01, .}, {3n+1, ...}
        const unsigned b0e = (unsigned)r0.z & 0xffffu, b1e = (unsigned)r0.z >> 16;
        const unsigned b0o = (unsigned)r1.z & 0xffffu, b1o = (unsigned)r1.z >> 16;
        const long long pitch = P.pitch;
        const uint8_t *row = src + (long long)r0.x * pitch;
        auto hsum = [&](const uint8_t *p, int *h) {
            const uint32_t *q = reinterpret_cast<const uint32_t *>(p);
            const uint32_t w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2);
            const uint32_t lo = __funnelshift_r(w0, w1, sh);   // p0c0 p0c1 p0c2 p1c0
            const uint32_t mid = __funnelshift_r(w1, w2, sh);  // p1c1 p1c2 p2c0 p2c1
            const uint32_t hi = w2 >> sh;                      // p2c2 (byte 0)
            const uint32_t a01 = __byte_perm(lo, mid, 0x4130);   // (p0c0 p1c0) (p0c1 p1c1)
            const uint32_t a2b0 = __byte_perm(lo, mid, 0x6352);  // (p0c2 p1c2) (p1c0 p2c0)
            const uint32_t b12 = __byte_perm(mid, hi, 0x4130);   // (p1c1 p2c1) (p1c2 p2c2)
            h[0] = (int)(__dp2a_lo(aA, a01, 0u) >> 4);
            h[1] = (int)(__dp2a_hi(aA, a01, 0u) >> 4);
            h[2] = (int)(__dp2a_lo(aA, a2b0, 0u) >> 4);
            h[3] = (int)(__dp2a_hi(aB, a2b0, 0u) >> 4);
            h[4] = (int)(__dp2a_lo(aB, b12, 0u) >> 4);
            h[5] = (int)(__dp2a_hi(aB, b12, 0u) >> 4);
        };
#pragma unroll(K_UNROLL32)
        for (int y = ya; y < yb; y += 2, off += 2 * ys, row += 3 * pitch) {
            int h0[6], h1[6], h2[6];
            hsum(row, h0);
            hsum(row + pitch, h1);
            hsum(row + 2 * pitch, h2);
            emit_row<T, false>(pc, off, 1, b0e, b1e, h0, h1, 0, tborder, true, true, true, true);
            emit_row<T, false>(pc, off + ys, 1, b0o, b1o, h1, h2, 0, tborder, true, true, true, true);
        }
    }
    for (int y = yb; y < yz; y++, off += ys) border_row<T>(pc, off, 1, 0u, true, true);
    for (int y = yz; y < y_end; y++, off += ys) border_row<T>(pc, off, 1, tborder, true, true);
}

// Warp 0, all 32 lanes: vertical taps of one tile (lane = row; `e` = its ytab entry, r0 = -1 when the row is border
// or the lane has no row), slot assignment, bulk copies into `buf`, row meta into `sh`, then the full-barrier arrive.
template <bool SHIFTED>
__device__ __forceinline__ void tile_prologue(const PyrParams &P, const PyrScale &S, TileShared &sh, uint8_t *buf,
                                              const uint8_t *src, int lane, int n_rows, int4 e, int g0, int row_bytes,
                                              int bulk_bytes) {
    const int r0 = e.x, r1 = e.y;
    const unsigned b01 = (unsigned)e.z;
    const bool zero_row = r0 >= P.src_h;  // api mode: both taps in the virtual zero padding
    const bool valid = r0 >= 0 && !zero_row;
    const int b1 = (int)(b01 >> 16);
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    const int slot_bytes = S.slot_bytes;
    int slot0 = 0, slot1 = 0, n_used = 0, rmin = 0;
    if (vmask) {
        if (S.dense_rows) {  // slot k = source row clamp(rmin + k): slots follow the UNCLAMPED tap rows
            const int u0 = e.w, u1 = (b1 != 0) ? u0 + 1 : u0;
            rmin = __shfl_sync(0xffffffffu, u0, __ffs(vmask) - 1);
            const int rmax = __shfl_sync(0xffffffffu, u1, 31 - __clz(vmask));
            n_used = min(rmax - rmin + 1, PYR_MAX_SLOTS);
            slot0 = u0 - rmin;
            slot1 = u1 - rmin;
        } else {
            const int cnt = valid ? (b1 != 0 ? 2 : 1) : 0;
            int incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            n_used = __shfl_sync(0xffffffffu, incl, 31);
            slot0 = incl - cnt;
            slot1 = (b1 != 0) ? slot0 + 1 : slot0;
        }
    }
    int tx_bytes = 0, tail = 0;
    bool virt_row = false;
    uint64_t *bar = reinterpret_cast<uint64_t *>(&sh.bar);
    auto stage = [&](int k, int row) {
        sh.slot_row[k] = row;
        if (row >= P.src_h) {  // virtual zero row (api mode): zero-filled from the slot base by the scalar staging
            virt_row = true;
            if (SHIFTED) {  // stream_tile adds slot_shift[k] to every read of the slot: it must not be left-over shared memory
                sh.slot_shift[k] = 0;
                sh.slot_cov[k] = 0;
            }
            return;
        }
        if (!SHIFTED) {
            if (bulk_bytes > 0) {
                bulk_g2s(buf + (size_t)k * slot_bytes, src + (long long)row * P.pitch + g0, (uint32_t)bulk_bytes, bar);
                tx_bytes += bulk_bytes;
            }
            return;
        }
        // g0 = first needed byte of the row (unaligned), row_bytes = needed bytes.  Copy the whole 16-byte chunks
        // between the boundary below that byte and the end of the row's real bytes; the rest is the scalar tail.
        const uint8_t *srow = src + (long long)row * P.pitch;
        const uintptr_t g = (uintptr_t)(srow + g0), a = g & ~(uintptr_t)15;
        const int d = (int)(g - a);
        const uintptr_t row_end = (uintptr_t)srow + (uintptr_t)(3 * P.src_w);
        long long cov = (long long)(row_end & ~(uintptr_t)15) - (long long)a;
        cov = min(cov, (long long)((d + row_bytes + 15) & ~15));
        if (cov < 0 || a < (uintptr_t)P.src) cov = 0;
        sh.slot_shift[k] = d;
        sh.slot_cov[k] = (int)cov;
        // bytes after the bulk copy up to the end of the real data (rounded up to a word; what follows are the
        // virtual zero columns of the api mode, written as words)
        tail = max(tail, min(d + row_bytes, (d + 3 * P.src_w - g0 + 3) & ~3) - (int)cov);
        if (cov > 0) {
            bulk_g2s(buf + (size_t)k * slot_bytes, reinterpret_cast<const void *>(a), (uint32_t)cov, bar);
            tx_bytes += (int)cov;
        }
    };
    if (S.dense_rows) {
        for (int k = lane; k < n_used; k += 32) stage(k, min(max(rmin + k, 0), P.virt_h - 1));
    } else if (valid) {
        stage(slot0, r0);
        if (b1 != 0) stage(slot0 + 1, r1);
    }
    tx_bytes = __reduce_add_sync(0xffffffffu, tx_bytes);
    tail = __reduce_max_sync(0xffffffffu, tail);
    const bool any_virt = __any_sync(0xffffffffu, virt_row);
    if (lane <= n_rows) {
        const int b0 = (int)(b01 & 0xffffu);
        sh.rowmeta[lane] = lane == n_rows ? make_int4(0x7ffffffe, 0, 0, 0)
                                          : make_int4(valid ? slot1 : (zero_row ? -4 : -2), b1 != 0 ? b0 : 0,
                                                      b1 != 0 ? b1 : b0, 0);
    }
    if (lane == 0) {
        sh.n_used = n_used;
        sh.need_scalar = SHIFTED ? (n_used > 0 && (tail > 0 || any_virt)) : ((bulk_bytes < row_bytes && n_used > 0) || any_virt);
        sh.tail_max = tail;
        sh.n_rows = n_rows;
    }
    __syncwarp();  // every lane's meta is ordered before lane 0's releasing arrive
    if (lane == 0) mbar_expect_tx(bar, (uint32_t)tx_bytes);
}

// Tap tables of a pyramid geometry (OpenCV's coefficient set-up).  Runs once per geometry:
// the tables are cached in the ctx, so the main kernel carries no fp64 and no conversions.
__global__ void pyramid_tables_kernel(const __grid_constant__ PyrParams P) {
    const PyrScale &S = P.sc[blockIdx.x];
    int2 *xt = const_cast<int2 *>(S.xtab);
    int4 *yt = const_cast<int4 *>(S.ytab);
    for (int x = threadIdx.x; x < S.dst_w; x += blockDim.x) {
        const int dx = x - S.left;
        int sx = -1;
        unsigned a01 = 0;
        if (dx >= 0 && dx < S.unpad_w) coeff_x(dx, S.scale_x, P.virt_w, sx, a01);
        xt[x] = make_int2(sx, (int)a01);
    }
    for (int y = threadIdx.x; y < S.dst_h; y += blockDim.x) {
        const int dy = y - S.top;
        int r0 = -1, r1 = -1, sy = 0;
        unsigned b01 = 0;
        if (dy >= 0 && dy < S.unpad_h) coeff_y(dy, S.scale_y, P.virt_h, r0, r1, b01, sy);
        yt[y] = make_int4(r0, r1, (int)b01, sy);
    }
}

// SHIFTED (launch-uniform): the source rows do not start on 16 bytes, see TileShared::slot_shift.  A separate
// instantiation keeps that machinery out of the kernel every aligned source runs.
template <typename T, bool SHIFTED>
__global__ void __launch_bounds__(PYR_THREADS, sizeof(T) == 2 ? PYR_MIN_CTAS : 6) pyramid_kernel(const __grid_constant__ PyrParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ TileShared sh;
    __shared__ int s_slotx[SHIFTED ? 2 * PYR_MAX_SLOTS : 2];

    const int tid = threadIdx.x;
    int tile = blockIdx.x;
    int si = 0;
#pragma unroll 1
    for (int s = 1; s < P.n_scales; s++)
        if (tile >= P.sc[s].tile_begin) si = s;
    const PyrScale &S = P.sc[si];
    tile -= S.tile_begin;
    const int img = blockIdx.y;

    const int tw = S.tw, th = S.th;
    const int ty = tile / S.tiles_x, tx = tile - ty * S.tiles_x;
    const int oy0 = ty * th, ox0 = tx * tw;
    if (S.direct) {  // CTA-uniform
        const bool cols = ox0 < S.left || ox0 + tw > S.left + S.unpad_w;  // portrait sources: rare, own instantiation
        if (cols) {
            if (S.mode == PM_HALF) direct_tile<T, PM_HALF, true>(P, S, img, oy0, ox0, tid);
            else if (S.mode == PM_RATIO32) direct_tile_32<T, true>(P, S, img, oy0, ox0, tid);
            else direct_tile<T, PM_POINT, true>(P, S, img, oy0, ox0, tid);
        } else if (S.mode == PM_HALF) direct_tile<T, PM_HALF, false>(P, S, img, oy0, ox0, tid);
        else if (S.mode == PM_RATIO32) direct_tile_32<T, false>(P, S, img, oy0, ox0, tid);
        else direct_tile<T, PM_POINT, false>(P, S, img, oy0, ox0, tid);
        return;
    }
    const int dst_h = S.dst_h, dst_w = S.dst_w, top = S.top, left = S.left, unpad_h = S.unpad_h, unpad_w = S.unpad_w;
    const int y_end = min(oy0 + th, dst_h), x_end = min(ox0 + tw, dst_w);
    const int n_rows = y_end - oy0;
    // interior (resized image) columns of this tile, in resized-image coordinates: [dxa, dxb)
    const int dxa = max(ox0, left) - left, dxb = min(x_end, left + unpad_w) - left;
    const bool has_cols = dxa < dxb;
    const int slot_bytes = S.slot_bytes;
    const uint8_t *src = P.src + (long long)img * P.img_stride;

    // ---- per-thread columns: horizontal taps (independent of warp 0's work below) -------------------
    const int x = ox0 + 2 * tid;
    const bool active = x < x_end && tid * 2 < tw;
    const bool two = x + 1 < x_end;
    int sxA = -1, sxB = -1;
    unsigned aA = 0, aB = 0;
    if (active) {
        if (two) {  // x is even and the table is 16-byte aligned: both entries in one load
            const int4 e = __ldg(reinterpret_cast<const int4 *>(S.xtab + x));
            sxA = e.x; aA = (unsigned)e.y; sxB = e.z; aB = (unsigned)e.w;
        } else {
            const int2 e = __ldg(S.xtab + x);
            sxA = e.x; aA = (unsigned)e.y;
        }
    }
    const bool inA = sxA >= 0, inB = sxB >= 0;

    // ---- warp 0: vertical taps, slot assignment, bulk copies ------------------------------------
    if (tid < 32) {
        const int lane = tid;
        if (lane == 0) {
            mbar_init(reinterpret_cast<uint64_t *>(&sh.bar), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            sh.slot_shift = s_slotx;
            sh.slot_cov = s_slotx + (SHIFTED ? PYR_MAX_SLOTS : 1);
        }
        __syncwarp();
        int4 e = make_int4(-1, 0, 0, 0);
        if (has_cols && lane < n_rows) e = __ldg(S.ytab + oy0 + lane);
        // source byte range of the tile's columns
        int sx = 0;
        if (has_cols && lane < 2) sx = __ldg(S.xtab + left + (lane == 0 ? dxa : dxb - 1)).x;
        const int sxa = __shfl_sync(0xffffffffu, sx, 0);
        const int sxb = min(__shfl_sync(0xffffffffu, sx, 1) + 1, P.virt_w - 1);
        const int g0 = SHIFTED ? 3 * sxa : ((3 * sxa) & ~15);
        const int row_bytes = SHIFTED ? 3 * (sxb - sxa + 1) : ((3 * sxb + 3 + 15) & ~15) - g0;
        const int real_end = 3 * P.src_w;
        const int bulk_bytes = P.vec_ok ? max(0, min(g0 + row_bytes, real_end & ~15) - g0) : 0;
        if (sxa >= P.src_w && e.x >= 0) e.x = P.src_h;  // the whole tile reads virtual zero columns (api mode, portrait)
        if (lane == 0) {
            sh.g0 = g0;
            sh.row_bytes = row_bytes;
            sh.bulk_bytes = bulk_bytes;
        }
        __syncwarp();
        tile_prologue<SHIFTED>(P, S, sh, smem, src, lane, n_rows, e, g0, row_bytes, bulk_bytes);
    }
    __syncthreads();
    const int n_used = sh.n_used, g0 = sh.g0;

    // ---- per-thread byte offsets inside a slot --------------------------------------------------------
    // (shifted sources: wA / wB stay the slot base and shA / shB carry the BYTE offset; the per-slot shift is added
    // in stream_tile)
    uint32_t wA = smem_u32(smem), wB = wA;  // out-of-image pixels read in bounds; their value is replaced
    unsigned shA = 0, shB = 0;
    if (inA) {
        const int o = 3 * sxA - g0;
        if (SHIFTED) shA = (unsigned)o;
        else {
            wA += (uint32_t)(o & ~3);
            shA = (unsigned)(o & 3) * 8u;
        }
    }
    if (inB) {
        const int o = 3 * sxB - g0;
        if (SHIFTED) shB = (unsigned)o;
        else {
            wB += (uint32_t)(o & ~3);
            shB = (unsigned)(o & 3) * 8u;
        }
    }

    // ---- scalar staging: whatever the bulk copies do not cover ---------------------------------------
    if (SHIFTED && sh.need_scalar) {
        // tails of the shifted slots, flattened over (slot, byte) so that the usual <= 31 bytes per slot cost a few
        // iterations in total; virtual rows are zero-filled whole
        const int need = sh.row_bytes, tail_max = sh.tail_max;
        const int real_len = 3 * P.src_w - g0;  // real bytes of a row from g0 on (may exceed `need`, may be <= 0)
        for (int idx = tid; idx < n_used * tail_max; idx += PYR_THREADS) {
            const int k = idx / tail_max, j = idx - k * tail_max;
            const int row = sh.slot_row[k];
            if (row >= P.src_h) continue;
            const int d = sh.slot_shift[k], i = sh.slot_cov[k] + j;
            if (i >= min(d + need, (d + real_len + 3) & ~3)) continue;
            const uint8_t *srow = src + (long long)row * P.pitch;
            const int pos = g0 - d + i;  // byte of the row this slot byte mirrors
            smem[(size_t)k * slot_bytes + i] = (pos >= 0 && pos < 3 * P.src_w) ? __ldg(srow + pos) : (uint8_t)0;
        }
        if (real_len < need) {  // CTA-uniform: the tile reaches into the virtual zero columns -> word-wise zeros
            const int zwords = (need - max(real_len, 0)) / 4 + 2;
            for (int idx = tid; idx < n_used * zwords; idx += PYR_THREADS) {
                const int k = idx / zwords, j = idx - k * zwords;
                if (sh.slot_row[k] >= P.src_h) continue;
                const int d = sh.slot_shift[k];
                const int i = ((d + max(real_len, 0) + 3) & ~3) + 4 * j;
                if (i < d + need) *reinterpret_cast<uint32_t *>(smem + (size_t)k * slot_bytes + i) = 0u;
            }
        }
        for (int k = 0; k < n_used; k++)
            if (sh.slot_row[k] >= P.src_h)
                for (int i = tid * 4; i < need + 16; i += PYR_THREADS * 4)
                    *reinterpret_cast<uint32_t *>(smem + (size_t)k * slot_bytes + i) = 0u;
        __syncthreads();
    } else if (!SHIFTED && sh.need_scalar) {
        const int row_bytes = sh.row_bytes, bulk_bytes = sh.bulk_bytes, rest = row_bytes - bulk_bytes;
        const int real_end = 3 * P.src_w;
        for (int k = 0; k < n_used; k++) {
            const int row = sh.slot_row[k];
            uint8_t *dst = smem + (size_t)k * slot_bytes;
            if (row >= P.src_h) {  // virtual zero rows of the api mode
                for (int i = tid * 4; i < row_bytes; i += PYR_THREADS * 4) *reinterpret_cast<uint32_t *>(dst + i) = 0u;
            } else {
                const uint8_t *srow = src + (long long)row * P.pitch;
                const uint8_t *from = srow + g0 + bulk_bytes;  // g0 and bulk_bytes are multiples of 16
                const int valid = real_end - (g0 + bulk_bytes);  // bytes of the row that exist from here on
                if (((uintptr_t)from & 7) == 0 && (real_end & 7) == 0) {  // rows that start on 8 bytes (e.g. 1080 px wide)
                    for (int i = tid * 8; i < rest; i += PYR_THREADS * 8)
                        *reinterpret_cast<uint2 *>(dst + bulk_bytes + i) =
                            i < valid ? __ldg(reinterpret_cast<const uint2 *>(from + i)) : make_uint2(0u, 0u);
                } else if (((uintptr_t)from & 3) == 0 && (real_end & 3) == 0) {
                    for (int i = tid * 4; i < rest; i += PYR_THREADS * 4)
                        *reinterpret_cast<uint32_t *>(dst + bulk_bytes + i) =
                            i < valid ? __ldg(reinterpret_cast<const uint32_t *>(from + i)) : 0u;
                } else {
                    for (int i = tid; i < rest; i += PYR_THREADS) dst[bulk_bytes + i] = i < valid ? __ldg(from + i) : (uint8_t)0;
                }
            }
        }
        __syncthreads();
    }
    if (n_used > 0) mbar_wait(reinterpret_cast<uint64_t *>(&sh.bar), 0);

    if (!active) return;

    // ---- stream over the slots ------------------------------------------------------------------------
    T *pc[3];
    {
        T *p0 = reinterpret_cast<T *>(S.out) + (long long)img * S.img_stride + (long long)oy0 * S.ys +
                (long long)x * S.xs;
        // channel flip (img[:, :, ::-1]) = source channel c goes to output plane 2 - c: swap the plane pointers
        pc[0] = P.flip ? p0 + 2 * S.cs : p0;
        pc[1] = p0 + S.cs;
        pc[2] = P.flip ? p0 : p0 + 2 * S.cs;
    }
    const unsigned tborder = (unsigned)P.border * 4u;
    // fast tiles: every column interior, every thread owns a full pixel pair, paired stores allowed -- the
    // common case (letterbox pads rows only; 320 | 640, 960, 1280); CTA-uniform, so no divergence
    const bool fast = S.vec != 0 && ox0 >= left && x_end <= left + unpad_w && ((x_end - ox0) & 1) == 0;
    if (fast && S.mode == PM_POINT)
        stream_tile<T, true, PM_POINT>(sh, pc, (unsigned)S.ys, 1, wA, shA, aA, wB, shB, aB, n_used, n_rows, slot_bytes,
                                       tborder, true, true, true, true);
    else if (fast && S.mode == PM_HALF)
        stream_tile<T, true, PM_HALF>(sh, pc, (unsigned)S.ys, 1, wA, shA, aA, wB, shB, aB, n_used, n_rows, slot_bytes,
                                      tborder, true, true, true, true);
    else if (fast && SHIFTED)
        stream_tile<T, true, PM_GENERAL, true>(sh, pc, (unsigned)S.ys, 1, wA, shA, aA, wB, shB, aB, n_used, n_rows,
                                               slot_bytes, tborder, true, true, true, true);
    else if (fast)
        stream_tile<T, true, PM_GENERAL>(sh, pc, (unsigned)S.ys, 1, wA, shA, aA, wB, shB, aB, n_used, n_rows,
                                         slot_bytes, tborder, true, true, true, true);
    else if (SHIFTED)
        stream_tile<T, false, PM_GENERAL, true>(sh, pc, (unsigned)S.ys, S.xs, wA, shA, aA, wB, shB, aB, n_used, n_rows,
                                                slot_bytes, tborder, inA, inB, two, S.vec != 0);
    else
        stream_tile<T, false, PM_GENERAL>(sh, pc, (unsigned)S.ys, S.xs, wA, shA, aA, wB, shB, aB, n_used, n_rows,
                                          slot_bytes, tborder, inA, inB, two, S.vec != 0);
}


template <typename T> struct OutBit { static constexpr int value = 1 << (int)sizeof(T); };  // 2, 4, 16
template <> struct OutBit<f32r> { static constexpr int value = 8; };

template <typename T>
int launch_group(fdms_ctx *ctx, const PyrParams &P, int tiles, int n_img, int smem, cudaStream_t st) {
    const int bit = OutBit<T>::value;
    if (!(ctx->pyr_attr & bit)) {  // dynamic shared memory opt-in, once per ctx and instantiation
        FDMS_CUDA(ctx, cudaFuncSetAttribute(pyramid_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        FDMS_CUDA(ctx, cudaFuncSetAttribute(pyramid_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        ctx->pyr_attr |= bit;
    }
    dim3 grid(tiles, n_img);
    if (P.shifted) pyramid_kernel<T, true><<<grid, PYR_THREADS, smem, st>>>(P);
    else pyramid_kernel<T, false><<<grid, PYR_THREADS, smem, st>>>(P);
    FDMS_LAUNCH_CHECK(ctx, "pyramid_kernel");
    return FDMS_OK;
}

}  // namespace

extern "C" FDMS_API int fdms_pyramid(fdms_ctx *ctx, const uint8_t *src, int n_img, int64_t img_stride,
                                     int64_t row_pitch, int src_h, int src_w, int virt_h, int virt_w,
                                     const fdms_scale_out *scales, int n_scales, int flip_rb, int border,
                                     fdms_stream stream) {
    FDMS_BIND(ctx);
    if (!src || !scales || n_img <= 0 || n_scales <= 0 || n_scales > FDMS_MAX_SCALES || src_h <= 0 || src_w <= 0)
        return fdms_fail(ctx, FDMS_ERR_INVALID, "fdms_pyramid: bad arguments");
    if (n_img > 65535) return fdms_fail(ctx, FDMS_ERR_UNSUPPORTED, "fdms_pyramid: n_img > 65535");
    if (virt_h < src_h || virt_w < src_w || row_pitch < (int64_t)src_w * 3 || border < 0 || border > 255)
        return fdms_fail(ctx, FDMS_ERR_INVALID, "fdms_pyramid: bad extents");
    for (int s = 0; s < n_scales; s++) {
        const fdms_scale_out &in = scales[s];
        const fdms_geom &g = in.geom;
        if (!in.out || g.dst_h <= 0 || g.dst_w <= 0 || g.unpad_h <= 0 || g.unpad_w <= 0 || g.top < 0 || g.left < 0 ||
            g.top + g.unpad_h > g.dst_h || g.left + g.unpad_w > g.dst_w)
            return fdms_fail(ctx, FDMS_ERR_INVALID, "fdms_pyramid: bad geometry for scale %d", s);
        if (in.out_dtype != FDMS_U8 && in.out_dtype != FDMS_F16 && in.out_dtype != FDMS_F32)
            return fdms_fail(ctx, FDMS_ERR_INVALID, "fdms_pyramid: bad out_dtype for scale %d", s);
    }
    // one launch per output dtype present (normally one: every scale of a pyramid shares it)
    for (int dt = FDMS_U8; dt <= FDMS_F32; dt++) {
        PyrParams P;
        memset(&P, 0, sizeof(P));
        P.src = src;
        P.img_stride = img_stride;
        P.pitch = row_pitch;
        P.src_h = src_h;
        P.src_w = src_w;
        P.virt_h = virt_h;
        P.virt_w = virt_w;
        P.flip = flip_rb ? 1 : 0;
        P.border = border;
        P.vec_ok = ((uintptr_t)src % 16 == 0) && (img_stride % 16 == 0) && (row_pitch % 16 == 0);
        P.shifted = !P.vec_ok;
        // direct tiles read aligned 32-bit words: every source row must start on a word
        const bool words_ok = ((uintptr_t)src % 4 == 0) && (img_stride % 4 == 0) && (row_pitch % 4 == 0);
        int tiles = 0, smem_max = 0, n = 0;
        for (int s = 0; s < n_scales; s++) {
            const fdms_scale_out &in = scales[s];
            if (in.out_dtype != dt) continue;
            const fdms_geom &g = in.geom;
            PyrScale &S = P.sc[n++];
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
            const size_t esz = dt == FDMS_F32 ? 4 : (dt == FDMS_F16 ? 2 : 1);
            S.vec = (S.xs == 1) && (g.dst_w % 2 == 0) && (S.cs % 2 == 0) && (S.img_stride % 2 == 0) && esz > 1 &&
                    ((uintptr_t)in.out % (2 * esz) == 0);
            S.dst_h = g.dst_h;
            S.dst_w = g.dst_w;
            S.unpad_h = g.unpad_h;
            S.unpad_w = g.unpad_w;
            S.top = g.top;
            S.left = g.left;
            // cv2.resize: scale = 1.0 / (dsize / ssize), computed in double (inv_scale first)
            S.scale_x = 1.0 / ((double)g.unpad_w / virt_w);
            S.scale_y = 1.0 / ((double)g.unpad_h / virt_h);
            S.dense_rows = S.scale_y <= 2.0;
            // tile: 320 columns (fewer when the staged rows would not fit), as many rows as the slots allow
            // odd integer scale: every vertical tap has weight 0 on the lower row -> one slot per output row
            const bool y_point = S.scale_y == (double)(long long)S.scale_y && (((long long)S.scale_y) & 1);
            {
                // exact integer factors on both axes, no clamped taps (virt == k * unpad), no virtual zeros
                const long long kx = (long long)S.scale_x, ky = (long long)S.scale_y;
                // (factor 1 -- the image already has the model size -- is the odd factor whose point sample is the pixel)
                const bool int_x = S.scale_x == (double)kx && kx >= 1 && kx * g.unpad_w == virt_w;
                const bool int_y = S.scale_y == (double)ky && ky >= 1 && ky * g.unpad_h == virt_h;
                const bool real = virt_h == src_h && virt_w == src_w;
                // api mode on a landscape image: the virtual square only adds zero ROWS below the image; when the
                // image height is a multiple of the factor no output row mixes real and virtual taps
                const bool api_rows = int_y && virt_w == src_w && virt_h > src_h && src_h % ky == 0;
                S.mode = PM_GENERAL;
                if (int_x && int_y && (real || api_rows) && ctx->pyr_special) {
                    if ((kx & 1) && (ky & 1)) S.mode = PM_POINT;
                    else if (!(kx & 1) && !(ky & 1)) S.mode = PM_HALF;
                }
                S.kx = (int)kx;
                S.ky = (int)ky;
                // direct tiles: paired stores, word-aligned source rows that end on a word
                // (column padding is fine as long as no column PAIR straddles the border)
                const bool cols_ok = g.left % 2 == 0 && g.unpad_w % 2 == 0 && g.dst_w % 2 == 0;
                S.direct = S.mode != PM_GENERAL && ctx->pyr_direct && S.vec && words_ok && cols_ok && src_w % 4 == 0 &&
                           (S.mode == PM_POINT || kx == 2);
                // the staged specialisations assume no virtual rows and rows that start on 16 bytes
                if ((!real || P.shifted) && !S.direct) S.mode = PM_GENERAL;
                S.real_rows = (S.direct && !real) ? src_h / (int)ky : g.unpad_h;
                // the exact 3:2 scale as a direct tile too (direct only: there is no staged flavour of it)
                const bool r32 = 2LL * virt_w == 3LL * g.unpad_w && 2LL * virt_h == 3LL * g.unpad_h && S.scale_x == 1.5 &&
                                 S.scale_y == 1.5;
                const bool rows32 = real || (virt_w == src_w && virt_h > src_h && src_h % 3 == 0);
                if (S.mode == PM_GENERAL && r32 && rows32 && (ctx->pyr_direct & 2) && ctx->pyr_special && S.vec && words_ok &&
                    cols_ok && src_w % 4 == 0 && g.top % 2 == 0 && ctx->pyr_th_direct % 2 == 0) {
                    S.mode = PM_RATIO32;
                    S.direct = 1;
                    S.real_rows = real ? g.unpad_h : 2 * (src_h / 3);
                }
            }
            int tw = 2 * PYR_THREADS, th = ctx->pyr_th;
            if (S.direct) {  // no staging: no shared-memory constraint on the tile
                th = ctx->pyr_th_direct;
                S.slot_bytes = 0;
            } else
            for (;;) {
                int n_slots = S.dense_rows ? (int)((th - 1) * S.scale_y) + 4 : (y_point ? th : 2 * th);
                int span_px = (int)(tw * S.scale_x) + 4;
                // slack: tap pairs past the span and whole-word reads; shifted sources start up to 15 bytes into the slot
                S.slot_bytes = ((span_px * 3 + 15) & ~15) + (P.shifted ? 64 : 48);
                if (n_slots <= PYR_MAX_SLOTS && (long long)n_slots * S.slot_bytes <= ctx->pyr_smem_budget) {
                    S.smem_need = n_slots * S.slot_bytes;
                    if (S.smem_need > smem_max) smem_max = S.smem_need;
                    break;
                }
                if (th > 1) th >>= 1;
                else if (tw > 32) tw >>= 1;
                else return fdms_fail(ctx, FDMS_ERR_UNSUPPORTED, "fdms_pyramid: scale %d needs too much shared memory", s);
            }
            S.tw = tw;
            S.th = th;
            S.tiles_x = ceil_div(g.dst_w, tw);
            S.tile_begin = tiles;
            tiles += S.tiles_x * ceil_div(g.dst_h, th);
        }
        if (n == 0) continue;
        P.n_scales = n;
        // tap tables: cached per geometry (FNV-1a over everything they depend on)
        {
            size_t bytes = 0;
            unsigned long long key = 1469598103934665603ull;
            auto mix = [&](long long v) {
                for (int q = 0; q < 8; q++) { key ^= (unsigned long long)(v >> (8 * q)) & 0xff; key *= 1099511628211ull; }
            };
            mix(dt); mix(src_h); mix(src_w); mix(virt_h); mix(virt_w); mix(n);
            for (int q = 0; q < n; q++) {
                const PyrScale &S = P.sc[q];
                mix(S.dst_h); mix(S.dst_w); mix(S.unpad_h); mix(S.unpad_w); mix(S.top); mix(S.left);
                bytes += (((size_t)S.dst_w * 8 + 15) & ~(size_t)15) + (size_t)S.dst_h * 16;
            }
            if (int rc = fdms_reserve(ctx, ctx->pyr_tab, bytes)) return rc;
            uint8_t *base = reinterpret_cast<uint8_t *>(ctx->pyr_tab.p);
            for (int q = 0; q < n; q++) {
                PyrScale &S = P.sc[q];
                S.xtab = reinterpret_cast<const int2 *>(base);
                base += ((size_t)S.dst_w * 8 + 15) & ~(size_t)15;
                S.ytab = reinterpret_cast<const int4 *>(base);
                base += (size_t)S.dst_h * 16;
            }
            mix((long long)(uintptr_t)ctx->pyr_tab.p);
            if (key != ctx->pyr_tab_key) {
                pyramid_tables_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(P);
                FDMS_LAUNCH_CHECK(ctx, "pyramid_tables_kernel");
                ctx->pyr_tab_key = key;
            }
        }
        // exact-factor scales of an aligned fp16 CHW pyramid: one pass over the source (pyramid_strip.cu); the tile
        // kernel below runs whatever is left
        if (dt == FDMS_F16) {
            unsigned handled = 0u;
            if (int rc = fdms_pyramid_strip(ctx, P, n_img, &handled, (cudaStream_t)stream)) return rc;
            if (handled) {
                int m = 0;
                tiles = 0;
                smem_max = 0;
                for (int q = 0; q < n; q++) {
                    if ((handled >> q) & 1u) continue;
                    PyrScale S = P.sc[q];
                    S.tile_begin = tiles;
                    tiles += S.tiles_x * ceil_div(S.dst_h, S.th);
                    if (S.smem_need > smem_max) smem_max = S.smem_need;
                    P.sc[m++] = S;
                }
                if (m == 0) continue;
                P.n_scales = m;
            }
        }
        int rc;
        if (dt == FDMS_F16) rc = launch_group<__half>(ctx, P, tiles, n_img, smem_max, (cudaStream_t)stream);
        else if (dt == FDMS_F32 && ctx->numerics == FDMS_NUMERICS_ATEN_CUDA)
            rc = launch_group<f32r>(ctx, P, tiles, n_img, smem_max, (cudaStream_t)stream);
        else if (dt == FDMS_F32) rc = launch_group<float>(ctx, P, tiles, n_img, smem_max, (cudaStream_t)stream);
        else rc = launch_group<uint8_t>(ctx, P, tiles, n_img, smem_max, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return FDMS_OK;
}
