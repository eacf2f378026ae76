// Cross-scale merge rows: the per-image pieces of multi_scale_face_detector.py:138-166 (rescale, round, drop
// landmarks, append scale_idx), :168-201,224-229 (size-aware weights) and :237-238 (detections[keep]) as device
// functions, shared by the standalone kernels of merge.cu and the fused single-CTA merge of nms.cu.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

struct MergeParams {
    const float *cand;
    const int32_t *keep, *n_keep;
    int n_scales, cap, mcap;
    float k[FDMS_MAX_SCALES];  // max(H0, W0) / in_h per scale
    float w0, h0;
    const int32_t *orig_hw;    // ragged batches: (n_img, 2) original (H0, W0) per image on the device; k / w0 / h0 above
    int in_h[FDMS_MAX_SCALES]; // are then derived per image from it and the scales' input heights
    float *mrows;      // (n_img, mcap, 8)
    int32_t *msrc;     // (n_img, mcap): s * cap + item index
    int32_t *mcount;   // (n_img)
    int32_t *merged_in;
};

struct EmitParams {
    const float *mrows;
    const int32_t *msrc, *keep2, *n_keep2;
    const float *cand;
    int n_scales, cap, mcap, out_cap;
    float k[FDMS_MAX_SCALES];
    float w0, h0;
    const int32_t *orig_hw;
    int in_h[FDMS_MAX_SCALES];
    float *dets;
    int32_t *n_out;
    float *kpts;
    // landmarks decoded HERE, for the final detections only, from the raw head maps (fdms_postprocess): the candidates
    // then carry boxes only -- 6 instead of 21 gathered sectors per candidate in the filter, and a dense crowd has ~15
    // candidates per final detection.  n_levels == 0: the candidate rows hold decoded landmarks (fdms_merge_scales).
    int n_levels, aten;
    const int32_t *cand_row;  // head row of every candidate (fdms_decode_filter)
    struct Lv {
        const void *data;
        int dtype, ny, nx, scale, row_offset;
        float stride;
    } lv[FDMS_MAX_LEVELS];
};

namespace {

constexpr int MERGE_NO = 21;
constexpr int MROW = 8;  // merge row: x1 y1 x2 y2 wscore conf cls scale

__device__ __forceinline__ float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

__device__ __forceinline__ float4 api_box(float4 b, float k, float w0, float h0, bool do_round) {
    b.x = clampf(f_mul(b.x, k), 0.f, w0);
    b.z = clampf(f_mul(b.z, k), 0.f, w0);
    b.y = clampf(f_mul(b.y, k), 0.f, h0);
    b.w = clampf(f_mul(b.w, k), 0.f, h0);
    if (do_round) {  // torch.round: half to even
        b.x = rintf(b.x); b.y = rintf(b.y); b.z = rintf(b.z); b.w = rintf(b.w);
    }
    return b;
}

// Rescale geometry of (image, scale): the batch-wide values, or -- ragged batches -- from the image's own size:
// k = (float)(max(H0, W0) / in_h) computed in double like the Python float it is (ppyp:141-142).
template <typename PT>
__device__ __forceinline__ void merge_geom(const PT &P, int img, int s, float &k, float &w0, float &h0) {
    if (P.orig_hw) {
        const int H0 = P.orig_hw[2 * img], W0 = P.orig_hw[2 * img + 1];
        k = (float)((double)max(H0, W0) / (double)P.in_h[s]);
        w0 = (float)W0;
        h0 = (float)H0;
    } else {
        k = P.k[s];
        w0 = P.w0;
        h0 = P.h0;
    }
}

// Rows of image `img` (all threads of the CTA take part); returns the number of merged rows (CTA-uniform).
__device__ __forceinline__ int merge_prepare_image(const MergeParams &P, int img, int tid, int n_threads) {
    int off[FDMS_MAX_SCALES + 1];
    off[0] = 0;
    for (int s = 0; s < P.n_scales; s++) {
        int nk = P.n_keep[img * P.n_scales + s];
        off[s + 1] = off[s] + max(nk, 0);
    }
    const int total = min(off[P.n_scales], P.mcap);
    if (tid == 0) {
        P.mcount[img] = total;
        if (P.merged_in) P.merged_in[img] = off[P.n_scales];
    }
    for (int pos = tid; pos < total; pos += n_threads) {
        int s = 0;
        while (pos >= off[s + 1]) s++;
        const int kidx = pos - off[s];
        const long long seg = (long long)img * P.n_scales + s;
        const int item = P.keep[seg * P.cap + kidx];
        const float *row = P.cand + (seg * P.cap + item) * MERGE_NO;
        float gk, gw, gh;
        merge_geom(P, img, s, gk, gw, gh);
        float4 b = api_box(make_float4(row[0], row[1], row[2], row[3]), gk, gw, gh, true);  // msfd:144
        const float conf = row[4], cls = row[5];
        // msfd:168-201 on the rounded boxes, float64
        double area = ((double)b.z - (double)b.x) * ((double)b.w - (double)b.y);
        double wt = 1.0;
        if (area < 1024.0 && s >= P.n_scales - 2) wt *= 1.2;
        if (area >= 1024.0 && area <= 16384.0 && s == P.n_scales / 2) wt *= 1.1;
        if (area > 16384.0 && s <= 1) wt *= 1.2;
        float wscore = (float)__dmul_rn((double)conf, wt);  // msfd:225,229
        float *o = P.mrows + ((long long)img * P.mcap + pos) * MROW;
        o[0] = b.x; o[1] = b.y; o[2] = b.z; o[3] = b.w;
        o[4] = wscore; o[5] = conf; o[6] = cls; o[7] = (float)s;
        P.msrc[(long long)img * P.mcap + pos] = s * P.cap + item;
    }
    return total;
}

// detections[keep] of image `img` with nk kept rows (all threads of the CTA take part).  The rows and the keep list
// may have been written by this very CTA (fused merge): plain loads, never the read-only path.
// `part` / `n_parts`: the rows of an image may be split over several CTAs (the standalone emit kernel).
__device__ __forceinline__ void merge_emit_image(const EmitParams &P, int img, int nk, int tid, int n_threads, int part = 0,
                                                 int n_parts = 1) {
    if (tid == 0 && part == 0) P.n_out[img] = nk;
    const int n = min(max(nk, 0), P.out_cap);
    for (int k = part * n_threads + tid; k < n; k += n_parts * n_threads) {
        const int pos = P.keep2[(long long)img * P.mcap + k];
        const float *m = P.mrows + ((long long)img * P.mcap + pos) * MROW;
        float *o = P.dets + ((long long)img * P.out_cap + k) * 7;
        o[0] = m[0]; o[1] = m[1]; o[2] = m[2]; o[3] = m[3];
        o[4] = m[5];  // original confidence (msfd:238 indexes the unweighted array)
        o[5] = m[6];
        o[6] = m[7];
        if (P.kpts) {
            const int src = P.msrc[(long long)img * P.mcap + pos];
            const int s = src / P.cap, item = src - s * P.cap;
            const long long slot = ((long long)img * P.n_scales + s) * P.cap + item;
            float *ko = P.kpts + ((long long)img * P.out_cap + k) * 15;
            float gk, gw, gh;
            merge_geom(P, img, s, gk, gw, gh);
            float kp[15];
            if (P.n_levels > 0) {
                // models/yolo.py:296-301 on the raw channels 6..20 of this detection's head cell (same operations, same
                // order as decode_cell in decode.cu)
                const int hrow = P.cand_row[slot];
                int li = -1;
                for (int l = 0; l < P.n_levels; l++)
                    if (P.lv[l].scale == s && hrow >= P.lv[l].row_offset) li = l;  // levels of a scale ascend in row_offset
                const EmitParams::Lv &L = P.lv[li];
                const int cells = L.ny * L.nx, rel = hrow - L.row_offset;
                const int a = rel / cells, cell = rel - a * cells;
                const int gy = cell / L.nx, gx = cell - gy * L.nx;
                const long long e0 = (((long long)img * 3 + a) * MERGE_NO + 6) * cells + cell;
                float v[15];
                if (L.dtype == FDMS_F16) {
                    const __half *p = reinterpret_cast<const __half *>(L.data) + e0;
#pragma unroll
                    for (int c = 0; c < 15; c++) v[c] = __half2float(__ldg(p + (long long)c * cells));
                } else {
                    const float *p = reinterpret_cast<const float *>(L.data) + e0;
#pragma unroll
                    for (int c = 0; c < 15; c++) v[c] = __ldg(p + (long long)c * cells);
                }
                const bool hf = P.aten != 0 && L.dtype == FDMS_F16;  // ATen's CUDA arithmetic for a half model (common.cuh)
#pragma unroll
                for (int j = 0; j < 5; j++) {
                    if (hf) {
                        kp[3 * j] = h_rn(h_xy(v[3 * j], (float)gx, L.stride));
                        kp[3 * j + 1] = h_rn(h_xy(v[3 * j + 1], (float)gy, L.stride));
                        kp[3 * j + 2] = h_sigmoid(v[3 * j + 2]);
                    } else {
                        kp[3 * j] = f_mul(f_add(f_sub(f_mul(v[3 * j], 2.f), 0.5f), (float)gx), L.stride);
                        kp[3 * j + 1] = f_mul(f_add(f_sub(f_mul(v[3 * j + 1], 2.f), 0.5f), (float)gy), L.stride);
                        kp[3 * j + 2] = sigmoid_ref(v[3 * j + 2], P.aten != 0);
                    }
                }
            } else {
                const float *row = P.cand + slot * MERGE_NO;
#pragma unroll
                for (int c = 0; c < 15; c++) kp[c] = row[6 + c];
            }
            for (int j = 0; j < 5; j++) {
                ko[3 * j] = clampf(f_mul(kp[3 * j], gk), 0.f, gw);
                ko[3 * j + 1] = clampf(f_mul(kp[3 * j + 1], gk), 0.f, gh);
                ko[3 * j + 2] = kp[3 * j + 2];
            }
        }
    }
}

}  // namespace
