/*
 * fdms_oracle.c -- CPU restatement of the reference's multi-scale face-detection
 * pre/post-processing hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for the CUDA kernels in
 * face_detection_multi_scale_b200/csrc.  It is never linked into, imported by or
 * called from the product library; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Parity pinning: every function here is checked by tests/test_oracle_golden.py
 * against golden vectors produced by the reference's OWN functions run in the
 * build container (tests/golden/make_golden.py imports /root/reference through
 * baseline/refshim.py) and, where the arithmetic lives in a third-party library
 * that the reference merely calls (cv2.resize, torchvision.ops.nms, torch
 * elementwise ops), against that library directly.
 *
 * Reference citations are relative to /root/reference.
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -fno-fast-math (oracle/Makefile).
 * -ffp-contract=off matters: several formulas below depend on every fp32
 * operation being rounded separately.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define FO_EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* B1  letterbox geometry -- utils/datasets.py:873-903                        */
/* ------------------------------------------------------------------------- */

/* Python round() on a float: round-half-to-even in double. */
static double py_round(double v) { return nearbyint(v); }

typedef struct {
    int dst_h, dst_w;     /* output image size */
    int unpad_h, unpad_w; /* resized image size (new_unpad) */
    int top, bottom, left, right;
    int resize;           /* 1 if cv2.resize is called (shape[::-1] != new_unpad) */
    double ratio_w, ratio_h;
    double dw, dh;        /* halved paddings as returned by the reference */
} fo_geom;

FO_EXPORT int fo_letterbox_geometry(int src_h, int src_w, int new_h, int new_w, int auto_, int scale_fill,
                                    int scaleup, int stride, fo_geom *g) {
    /* utils/datasets.py:880-883 */
    double r = fmin((double)new_h / src_h, (double)new_w / src_w);
    if (!scaleup) r = fmin(r, 1.0);
    /* :886-887 */
    int unpad_w = (int)py_round(src_w * r), unpad_h = (int)py_round(src_h * r);
    double dw = new_w - unpad_w, dh = new_h - unpad_h;
    double rw = r, rh = r;
    if (auto_) { /* :888-889, np.mod on ints: result has the sign of the divisor */
        int idw = (int)dw % stride, idh = (int)dh % stride;
        if (idw < 0) idw += stride;
        if (idh < 0) idh += stride;
        dw = idw; dh = idh;
    } else if (scale_fill) { /* :890-893 */
        dw = 0.0; dh = 0.0;
        unpad_w = new_w; unpad_h = new_h;
        rw = (double)new_w / src_w; rh = (double)new_h / src_h;
    }
    dw /= 2; dh /= 2; /* :895-896 */
    g->resize = !(src_w == unpad_w && src_h == unpad_h); /* :898 */
    g->top = (int)py_round(dh - 0.1); g->bottom = (int)py_round(dh + 0.1); /* :900 */
    g->left = (int)py_round(dw - 0.1); g->right = (int)py_round(dw + 0.1); /* :901 */
    g->unpad_h = unpad_h; g->unpad_w = unpad_w;
    g->dst_h = unpad_h + g->top + g->bottom; g->dst_w = unpad_w + g->left + g->right;
    g->ratio_w = rw; g->ratio_h = rh; g->dw = dw; g->dh = dh;
    return 0;
}

/* ------------------------------------------------------------------------- */
/* B2  cv2.resize(INTER_LINEAR) on uint8 -- third party (OpenCV imgproc,      */
/*     resize.cpp: resize() coefficient set-up, HResizeLinear<uchar,int,short, */
/*     INTER_RESIZE_COEF_SCALE=2048>, VResizeLinear<uchar,int,short,           */
/*     FixedPtCast<int,uchar,22>>).  Call sites: utils/datasets.py:899,        */
/*     utils/preprocess_yolo_predict.py:336.                                   */
/* ------------------------------------------------------------------------- */

static int cv_round_to_short(float v) {
    /* saturate_cast<short>(float) == cvRound (round half to even) then saturate */
    long r = lrintf(v);
    if (r > 32767) r = 32767;
    if (r < -32768) r = -32768;
    return (int)r;
}

/* x axis: taps past either edge get weight 0 (resize.cpp coefficient loop). */
static void axis_x(int n_dst, int n_src, int *ofs, short *a0, short *a1) {
    double scale = 1.0 / ((double)n_dst / n_src);
    for (int d = 0; d < n_dst; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= s;
        if (s < 0) { f = 0; s = 0; }
        if (s >= n_src - 1) { f = 0; s = n_src - 1; }
        ofs[d] = s;
        a0[d] = (short)cv_round_to_short((1.f - f) * 2048.f);
        a1[d] = (short)cv_round_to_short(f * 2048.f);
    }
}

/* y axis: weights are kept, source ROWS are clamped (resizeGeneric_Invoker). */
static void axis_y(int n_dst, int n_src, int *r0, int *r1, short *b0, short *b1) {
    double scale = 1.0 / ((double)n_dst / n_src);
    for (int d = 0; d < n_dst; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= s;
        int s0 = s, s1 = s + 1;
        if (s0 < 0) s0 = 0;
        if (s0 >= n_src) s0 = n_src - 1;
        if (s1 < 0) s1 = 0;
        if (s1 >= n_src) s1 = n_src - 1;
        r0[d] = s0; r1[d] = s1;
        b0[d] = (short)cv_round_to_short((1.f - f) * 2048.f);
        b1[d] = (short)cv_round_to_short(f * 2048.f);
    }
}

/* Reads the source through a virtual zero-padded extent (virt_h, virt_w) >= (src_h, src_w):
 * pixels outside the real image are 0.  virt == src for the plain letterbox path; the
 * "api" path (utils/preprocess_yolo_predict.py:273-290) resizes a zero-padded square. */
static inline int src_px(const uint8_t *src, int src_h, int src_w, int64_t pitch, int cn, int y, int x, int c) {
    if (y >= src_h || x >= src_w) return 0;
    return src[(int64_t)y * pitch + (int64_t)x * cn + c];
}

FO_EXPORT int fo_resize_linear_u8(const uint8_t *src, int src_h, int src_w, int64_t src_pitch, int virt_h,
                                  int virt_w, uint8_t *dst, int dst_h, int dst_w, int64_t dst_pitch, int cn) {
    if (virt_h < src_h || virt_w < src_w) return -1;
    int *xo = (int *)malloc(sizeof(int) * dst_w), *y0 = (int *)malloc(sizeof(int) * dst_h),
        *y1 = (int *)malloc(sizeof(int) * dst_h);
    short *a0 = (short *)malloc(2 * dst_w), *a1 = (short *)malloc(2 * dst_w), *b0 = (short *)malloc(2 * dst_h),
          *b1 = (short *)malloc(2 * dst_h);
    axis_x(dst_w, virt_w, xo, a0, a1);
    axis_y(dst_h, virt_h, y0, y1, b0, b1);
    for (int dy = 0; dy < dst_h; dy++) {
        for (int dx = 0; dx < dst_w; dx++) {
            int sx = xo[dx], sx1 = sx + 1 < virt_w ? sx + 1 : sx; /* a1 == 0 whenever sx+1 is past the edge */
            for (int c = 0; c < cn; c++) {
                int r0 = src_px(src, src_h, src_w, src_pitch, cn, y0[dy], sx, c) * a0[dx] +
                         src_px(src, src_h, src_w, src_pitch, cn, y0[dy], sx1, c) * a1[dx];
                int r1 = src_px(src, src_h, src_w, src_pitch, cn, y1[dy], sx, c) * a0[dx] +
                         src_px(src, src_h, src_w, src_pitch, cn, y1[dy], sx1, c) * a1[dx];
                int v = (((b0[dy] * (r0 >> 4)) >> 16) + ((b1[dy] * (r1 >> 4)) >> 16) + 2) >> 2;
                if (v < 0) v = 0;
                if (v > 255) v = 255;
                dst[(int64_t)dy * dst_pitch + (int64_t)dx * cn + c] = (uint8_t)v;
            }
        }
    }
    free(xo); free(y0); free(y1); free(a0); free(a1); free(b0); free(b1);
    return 0;
}

/* letterbox = resize (if needed) + cv2.copyMakeBorder(constant) -- utils/datasets.py:898-902.
 * dst is HWC u8 of size g.dst_h x g.dst_w (dense).  virt_* as above. */
FO_EXPORT int fo_letterbox_u8(const uint8_t *src, int src_h, int src_w, int64_t src_pitch, int virt_h, int virt_w,
                              const fo_geom *g, int border, uint8_t *dst) {
    int cn = 3;
    int64_t dpitch = (int64_t)g->dst_w * cn;
    memset(dst, border, (size_t)dpitch * g->dst_h);
    uint8_t *inner = dst + (int64_t)g->top * dpitch + (int64_t)g->left * cn;
    if (g->resize) return fo_resize_linear_u8(src, src_h, src_w, src_pitch, virt_h, virt_w, inner, g->unpad_h,
                                              g->unpad_w, dpitch, cn);
    for (int y = 0; y < g->unpad_h; y++)
        for (int x = 0; x < g->unpad_w; x++)
            for (int c = 0; c < cn; c++) inner[y * dpitch + x * cn + c] = (uint8_t)src_px(src, src_h, src_w, src_pitch, cn, y, x, c);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* B3  pack -- multi_scale_face_detector.py:96-105                             */
/*     HWC u8 -> CHW float, optional channel flip (img[:, :, ::-1]), /255.     */
/*     fp32: true division.  fp16: rn16(rn32(v/255)) (ATen computes half ops   */
/*     in float and rounds once to half), returned as raw binary16 bits.       */
/* ------------------------------------------------------------------------- */

static uint16_t f32_to_f16_bits(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    int32_t exp = (int32_t)((x >> 23) & 0xff) - 127 + 15;
    uint32_t man = x & 0x7fffffu;
    if (((x >> 23) & 0xff) == 0xff) return (uint16_t)(sign | 0x7c00u | (man ? 0x200u : 0));
    if (exp >= 31) return (uint16_t)(sign | 0x7c00u);
    if (exp <= 0) {
        if (exp < -10) return (uint16_t)sign;
        man |= 0x800000u;
        int shift = 14 - exp;
        uint32_t half = man >> shift, rem = man & ((1u << shift) - 1), mid = 1u << (shift - 1);
        if (rem > mid || (rem == mid && (half & 1))) half++;
        return (uint16_t)(sign | half);
    }
    uint32_t half = ((uint32_t)exp << 10) | (man >> 13);
    uint32_t rem = man & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (half & 1))) half++;
    return (uint16_t)(sign | half);
}

FO_EXPORT void fo_pack_chw(const uint8_t *hwc, int h, int w, int flip, int out_f16, void *out) {
    int64_t plane = (int64_t)h * w;
    for (int c = 0; c < 3; c++) {
        int sc = flip ? 2 - c : c;
        for (int64_t i = 0; i < plane; i++) {
            float v = (float)hwc[i * 3 + sc] / 255.0f;
            if (out_f16) ((uint16_t *)out)[c * plane + i] = f32_to_f16_bits(v);
            else ((float *)out)[c * plane + i] = v;
        }
    }
}

/* ------------------------------------------------------------------------- */
/* B5  head decode -- models/yolo.py:267-308 (IKeypoint inference branch)      */
/*     x: (bs, na*no, ny, nx) fp32, channel = a*no + k; pred: (bs, rows_total, */
/*     no) where this level's rows start at row_offset.  nc=1, nkpt=5 => no=21. */
/* ------------------------------------------------------------------------- */

/* torch.sigmoid in fp32 = 1 / (1 + exp(-x)).  ATen's vectorised expf (Sleef, 1 ulp) cannot be restated bit for bit;
 * the exponential is taken correctly rounded here (fp64 exp rounded once) instead of libm's expf (0.502 ulp: a
 * randomised run found one confidence in ~40 000 that differs in the last bit between the two).  The reference
 * fixtures pin the decode within 1e-4 px either way (tests/test_oracle_golden.py). */
static float sigmoidf_(float v) { return 1.0f / (1.0f + (float)exp(-(double)v)); }

FO_EXPORT void fo_decode_level(const float *x, int bs, int na, int no, int ny, int nx, float stride,
                               const float *anchors_px /* na*2: (w,h) */, int64_t rows_total, int64_t row_offset,
                               float *pred) {
    int64_t P = (int64_t)ny * nx;
    for (int b = 0; b < bs; b++)
        for (int a = 0; a < na; a++)
            for (int gy = 0; gy < ny; gy++)
                for (int gx = 0; gx < nx; gx++) {
                    const float *in = x + ((int64_t)b * na * no + (int64_t)a * no) * P + (int64_t)gy * nx + gx;
                    float *o = pred + ((int64_t)b * rows_total + row_offset + a * P + (int64_t)gy * nx + gx) * no;
                    float s0 = sigmoidf_(in[0 * P]), s1 = sigmoidf_(in[1 * P]);
                    float s2 = sigmoidf_(in[2 * P]), s3 = sigmoidf_(in[3 * P]);
                    float t;
                    t = s0 * 2.f; t = t - 0.5f; t = t + (float)gx; o[0] = t * stride; /* yolo.py:290 */
                    t = s1 * 2.f; t = t - 0.5f; t = t + (float)gy; o[1] = t * stride;
                    t = s2 * 2.f; t = t * t; o[2] = t * anchors_px[a * 2 + 0];         /* yolo.py:291 */
                    t = s3 * 2.f; t = t * t; o[3] = t * anchors_px[a * 2 + 1];
                    o[4] = sigmoidf_(in[4 * P]);
                    for (int k = 5; k < no - 15; k++) o[k] = sigmoidf_(in[k * P]);
                    int k0 = no - 15;
                    for (int j = 0; j < 5; j++) { /* yolo.py:293-295: raw logits for x/y, sigmoid for conf */
                        t = in[(k0 + 3 * j) * P] * 2.f; t = t - 0.5f; t = t + (float)gx; o[k0 + 3 * j] = t * stride;
                        t = in[(k0 + 3 * j + 1) * P] * 2.f; t = t - 0.5f; t = t + (float)gy; o[k0 + 3 * j + 1] = t * stride;
                        o[k0 + 3 * j + 2] = sigmoidf_(in[(k0 + 3 * j + 2) * P]);
                    }
                }
}

/* The same inference branch for every head variant the reference can run (Detect models/yolo.py:59-106, IDetect
 * :152-207, IKeypoint :267-308): nkpt > 0 -> 6 detection columns (nc = 1) + (x, y, conf) per keypoint, inplace
 * branch; nkpt == 0 -> all columns through the sigmoid, xy / wh decoded (the plain head: :95-99 for nc > 1). */
FO_EXPORT void fo_decode_level_head(const float *x, int bs, int na, int no, int nkpt, int ny, int nx, float stride,
                                    const float *anchors_px /* na*2: (w,h) */, int64_t rows_total, int64_t row_offset,
                                    float *pred) {
    int64_t P = (int64_t)ny * nx;
    int n_sig = nkpt > 0 ? 6 : no;
    for (int b = 0; b < bs; b++)
        for (int a = 0; a < na; a++)
            for (int gy = 0; gy < ny; gy++)
                for (int gx = 0; gx < nx; gx++) {
                    const float *in = x + ((int64_t)b * na * no + (int64_t)a * no) * P + (int64_t)gy * nx + gx;
                    float *o = pred + ((int64_t)b * rows_total + row_offset + a * P + (int64_t)gy * nx + gx) * no;
                    for (int c = 0; c < no; c++) {
                        float v = in[c * P], t;
                        if (c < n_sig) {
                            float sg = sigmoidf_(v);
                            if (c < 2) { t = sg * 2.f; t = t - 0.5f; t = t + (float)(c == 0 ? gx : gy); o[c] = t * stride; }
                            else if (c < 4) { t = sg * 2.f; t = t * t; o[c] = t * anchors_px[a * 2 + (c - 2)]; }
                            else o[c] = sg;
                        } else if ((c - 6) % 3 == 2) o[c] = sigmoidf_(v);
                        else { t = v * 2.f; t = t - 0.5f; t = t + (float)((c - 6) % 3 == 0 ? gx : gy); o[c] = t * stride; }
                    }
                }
}

/* ------------------------------------------------------------------------- */
/* B6  confidence filter + xywh2xyxy -- utils/general.py:515,531,547-563,299-306 */
/*     nc == 1, kpt_label == 5.  Output rows: [x1,y1,x2,y2,conf,0,15 kpt] (21), */
/*     plus the source row index in src_row[].  Candidates stay in row order.   */
/* ------------------------------------------------------------------------- */

FO_EXPORT int64_t fo_filter_xyxy(const float *pred, int64_t rows, int no, float conf_thres, float *cand,
                                 int32_t *src_row) {
    int64_t n = 0;
    for (int64_t r = 0; r < rows; r++) {
        const float *p = pred + r * no;
        if (!(p[4] > conf_thres)) continue; /* general.py:515,531 */
        float conf = p[5] * p[4];           /* :547 */
        if (!(conf > conf_thres)) continue; /* :563 */
        float *o = cand + n * no;
        float hw = p[2] / 2, hh = p[3] / 2; /* :302-305 */
        o[0] = p[0] - hw; o[1] = p[1] - hh; o[2] = p[0] + hw; o[3] = p[1] + hh;
        o[4] = conf; o[5] = 0.f;
        for (int k = 6; k < no; k++) o[k] = p[k];
        src_row[n] = (int32_t)r;
        n++;
    }
    return n;
}

/* ------------------------------------------------------------------------- */
/* B7  torchvision.ops.nms -- third party; call sites utils/general.py:584,    */
/*     multi_scale_face_detector.py:234.                                       */
/*   mode 0 (TV_CPU):  torchvision/csrc/ops/cpu/nms_kernel.cpp                  */
/*        ovr = inter / ((area_i + area_j) - inter); suppress iff               */
/*        (double)ovr > iou_threshold (threshold stays double).                 */
/*   mode 1 (TV_CUDA): torchvision/csrc/ops/cuda/nms_kernel.cu as compiled for  */
/*        sm_100 in torchvision 0.26: the column box's area is */
/*        fused, den = fma(bw_j, bh_j, area_i) - inter; compare in fp32 against */
/*        (float)iou_threshold.                                                 */
/*   Both: stable descending sort on score, strict '>', NaN never suppresses.   */
/* ------------------------------------------------------------------------- */

typedef struct { float s; int64_t i; } fo_si;
static int cmp_desc_stable(const void *a, const void *b) {
    const fo_si *x = (const fo_si *)a, *y = (const fo_si *)b;
    if (x->s > y->s) return -1;
    if (x->s < y->s) return 1;
    return x->i < y->i ? -1 : (x->i > y->i ? 1 : 0);
}

FO_EXPORT int64_t fo_nms(const float *boxes /* n*4 xyxy */, const float *scores, int64_t n, double iou_thres,
                         int mode, int64_t *keep) {
    if (n <= 0) return 0;
    fo_si *ord = (fo_si *)malloc(sizeof(fo_si) * n);
    uint8_t *sup = (uint8_t *)calloc(n, 1);
    for (int64_t i = 0; i < n; i++) { ord[i].s = scores[i]; ord[i].i = i; }
    qsort(ord, n, sizeof(fo_si), cmp_desc_stable);
    float thr_f = (float)iou_thres;
    int64_t nk = 0;
    for (int64_t oi = 0; oi < n; oi++) {
        int64_t i = ord[oi].i;
        if (sup[i]) continue;
        keep[nk++] = i;
        const float *a = boxes + i * 4;
        float area_i = (a[2] - a[0]) * (a[3] - a[1]);
        for (int64_t oj = oi + 1; oj < n; oj++) {
            int64_t j = ord[oj].i;
            if (sup[j]) continue;
            const float *b = boxes + j * 4;
            float xx1 = a[0] > b[0] ? a[0] : b[0], yy1 = a[1] > b[1] ? a[1] : b[1];
            float xx2 = a[2] < b[2] ? a[2] : b[2], yy2 = a[3] < b[3] ? a[3] : b[3];
            if (mode == 2) { /* cpp/yolov7-face-ncnn/src/yolov7face.cpp:43-74: +1 extents, ovr >= thresh */
                float wn = xx2 - xx1 + 1, hn = yy2 - yy1 + 1;
                if (!(wn > 0.f)) wn = 0.f;
                if (!(hn > 0.f)) hn = 0.f;
                float in_ = wn * hn;
                float ai = (a[2] - a[0] + 1) * (a[3] - a[1] + 1), aj = (b[2] - b[0] + 1) * (b[3] - b[1] + 1);
                float ovr = in_ / (ai + aj - in_);
                if (ovr >= thr_f) sup[j] = 1;
                continue;
            }
            float w = xx2 - xx1, h = yy2 - yy1;
            if (!(w > 0.f)) w = 0.f;
            if (!(h > 0.f)) h = 0.f;
            float inter = w * h;
            float bw = b[2] - b[0], bh = b[3] - b[1];
            if (mode == 0) {
                float area_j = bw * bh;
                float ovr = inter / ((area_i + area_j) - inter);
                if ((double)ovr > iou_thres) sup[j] = 1;
            } else {
                float den = fmaf(bw, bh, area_i) - inter;
                float ovr = inter / den;
                if (ovr > thr_f) sup[j] = 1;
            }
        }
    }
    free(ord); free(sup);
    return nk;
}

/* ------------------------------------------------------------------------- */
/* B8  rescale                                                                 */
/* ------------------------------------------------------------------------- */

static float clampf_(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* utils/preprocess_yolo_predict.py:122-157 (+ .round() at multi_scale_face_detector.py:144 when do_round) */
FO_EXPORT void fo_scale_coords_api(float *boxes, int64_t n, int in_h, int orig_h, int orig_w, int do_round) {
    int sq = orig_h > orig_w ? orig_h : orig_w;
    float k = (float)((double)sq / in_h); /* Python double scalar, multiplied into an fp32 tensor */
    for (int64_t i = 0; i < n; i++) {
        float *b = boxes + i * 4;
        b[0] = clampf_(b[0] * k, 0.f, (float)orig_w); b[2] = clampf_(b[2] * k, 0.f, (float)orig_w);
        b[1] = clampf_(b[1] * k, 0.f, (float)orig_h); b[3] = clampf_(b[3] * k, 0.f, (float)orig_h);
        if (do_round) for (int c = 0; c < 4; c++) b[c] = nearbyintf(b[c]); /* torch.round: half to even */
    }
}

/* utils/general.py:374-404.  kpt_step == 0: box path (cols 0..3; clip touches ROWS 0..3 only -- the
 * reference slices coords[0:4], general.py:389).  kpt_step == 3: keypoint path over `cols` columns,
 * cols 0::3 and 1::3 shifted/scaled and clipped on every row, conf columns untouched. */
/* general.py:379-383: gain / pad handed in by the caller (ratio_pad); the Python floats are rounded to fp32 by each
 * in-place tensor op. */
FO_EXPORT void fo_scale_coords_rp(float *coords, int64_t n, int cols, double gain_d, double padx_d, double pady_d,
                                  int orig_h, int orig_w, int kpt_step) {
    float padx = (float)padx_d, pady = (float)pady_d, gain = (float)gain_d;
    for (int64_t i = 0; i < n; i++) {
        float *r = coords + i * cols;
        if (!kpt_step) {
            r[0] = (r[0] - padx) / gain; r[2] = (r[2] - padx) / gain;
            r[1] = (r[1] - pady) / gain; r[3] = (r[3] - pady) / gain;
            if (i < 4) {
                r[0] = clampf_(r[0], 0.f, (float)orig_w); r[2] = clampf_(r[2], 0.f, (float)orig_w);
                r[1] = clampf_(r[1], 0.f, (float)orig_h); r[3] = clampf_(r[3], 0.f, (float)orig_h);
            }
        } else {
            for (int c = 0; c < cols; c += kpt_step) {
                r[c] = clampf_((r[c] - padx) / gain, 0.f, (float)orig_w);
                if (c + 1 < cols) r[c + 1] = clampf_((r[c + 1] - pady) / gain, 0.f, (float)orig_h);
            }
        }
    }
}

FO_EXPORT void fo_scale_coords_std(float *coords, int64_t n, int cols, int in_h, int in_w, int orig_h, int orig_w,
                                   int kpt_step) {
    double gain = fmin((double)in_h / orig_h, (double)in_w / orig_w);  /* general.py:377-378 */
    fo_scale_coords_rp(coords, n, cols, gain, ((double)in_w - orig_w * gain) / 2, ((double)in_h - orig_h * gain) / 2,
                       orig_h, orig_w, kpt_step);
}

/* ------------------------------------------------------------------------- */
/* B9  cross-scale merge -- multi_scale_face_detector.py:168-240               */
/*     dets: (n,7) float64 [x1,y1,x2,y2,conf,cls,scale_idx]; returns keep      */
/*     indices into dets in weighted-score order.                              */
/* ------------------------------------------------------------------------- */

FO_EXPORT void fo_scale_weights(const double *dets, int64_t n, int n_scales, double *w) {
    for (int64_t i = 0; i < n; i++) {
        const double *d = dets + i * 7;
        double area = (d[2] - d[0]) * (d[3] - d[1]);
        int s = (int)d[6];
        double wt = 1.0;
        if (area < 1024 && s >= n_scales - 2) wt *= 1.2;                     /* :188-189 */
        if (area >= 1024 && area <= 16384 && s == n_scales / 2) wt *= 1.1;   /* :192-194 */
        if (area > 16384 && s <= 1) wt *= 1.2;                               /* :197-198 */
        w[i] = wt;
    }
}

FO_EXPORT int64_t fo_weighted_nms(const double *dets, int64_t n, int n_scales, double iou_thres, int64_t *keep) {
    if (n <= 0) return 0;
    double *w = (double *)malloc(sizeof(double) * n);
    float *boxes = (float *)malloc(sizeof(float) * 4 * n), *scores = (float *)malloc(sizeof(float) * n);
    fo_scale_weights(dets, n, n_scales, w);
    for (int64_t i = 0; i < n; i++) {
        for (int c = 0; c < 4; c++) boxes[i * 4 + c] = (float)dets[i * 7 + c]; /* :228 */
        scores[i] = (float)(dets[i * 7 + 4] * w[i]);                             /* :225,229 */
    }
    int64_t nk = fo_nms(boxes, scores, n, iou_thres, 0 /* reference runs this NMS on CPU, :234 */, keep);
    free(w); free(boxes); free(scores);
    return nk;
}

/* ------------------------------------------------------------------------- */
/* "next" row: WIDER-FACE evaluation IoU matrix             */
/*     widerface_evaluate/box_overlaps.pyx:15-55 (float64, +1 pixel convention; */
/*     entries stay 0 where iw <= 0 or ih <= 0).  overlaps is (N, K) row major.  */
/* ------------------------------------------------------------------------- */
FO_EXPORT void fo_bbox_overlaps(const double *boxes, int64_t N, const double *query, int64_t K, double *overlaps) {
    for (int64_t i = 0; i < N * K; i++) overlaps[i] = 0.0;
    for (int64_t k = 0; k < K; k++) {
        const double *q = query + k * 4;
        double box_area = (q[2] - q[0] + 1) * (q[3] - q[1] + 1); /* :33-36 */
        for (int64_t n = 0; n < N; n++) {
            const double *b = boxes + n * 4;
            double iw = fmin(b[2], q[2]) - fmax(b[0], q[0]) + 1; /* :38-41 */
            if (iw > 0) {
                double ih = fmin(b[3], q[3]) - fmax(b[1], q[1]) + 1; /* :43-46 */
                if (ih > 0) {
                    double ua = (b[2] - b[0] + 1) * (b[3] - b[1] + 1) + box_area - iw * ih; /* :48-52 */
                    overlaps[n * K + k] = iw * ih / ua;                                    /* :53 */
                }
            }
        }
    }
}
