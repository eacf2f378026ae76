// Head decode kernels.
//
//   decode_filter_kernel : IKeypoint decode (models/yolo.py:267-308) fused with the confidence
//                          filter, xywh2xyxy and stream compaction of non_max_suppression
//                          (utils/general.py:515,531,547-563,299-306).  Reads the raw conv
//                          outputs (bs, 63, ny, nx) in place: no view/permute/contiguous/clone/cat
//                          and no (rows x 21) tensor.  Only the obj/cls planes are streamed; the
//                          other 19 channel planes are touched only at surviving cells (by
//                          decode_candidates_kernel, one candidate per thread; 4 planes instead of 19
//                          when the caller does not need landmarks).
//   decode_dense_kernel  : the same decode writing every row (compat level, IKeypoint.forward's
//                          first return value).
//   filter_pred_kernel   : the same filter reading an already decoded pred (compat level).
//
// Arithmetic contract: every fp32 operation the reference rounds separately
// is rounded separately here (explicit _rn intrinsics, -fmad=false); sigmoid is
// 1/(1+exp(-x)) with exp evaluated in fp64 and rounded once to fp32 (correctly rounded up to
// double rounding), IEEE division.
#include <math.h>

#include "common.cuh"

namespace {

constexpr int DEC_THREADS = 128;
constexpr int DEC_VEC = 8;                            // cells per 128-bit (fp16) load
#ifndef FDMS_DEC_UNROLL
#define FDMS_DEC_UNROLL 4  // vectors per plane and thread, all loads issued before the first compare (1, 2 or 4)
#endif
constexpr int DEC_UNROLL = FDMS_DEC_UNROLL;
constexpr int DEC_GROUP = DEC_THREADS * DEC_VEC;      // cells of one vector round of a CTA
constexpr int DEC_CELLS = DEC_GROUP * DEC_UNROLL;     // cells per CTA
static_assert(DEC_UNROLL >= 1 && DEC_UNROLL * DEC_VEC <= 32, "the pending-cell mask is one 32-bit word");
constexpr int NO = 21, NA = 3;

struct DecLevel {
    const void *data;
    int dtype, ny, nx, scale, row_offset;
    int block_begin, chunks;  // first CTA of this level; CTAs per (b, a) plane
    float stride;
    float anchors[6];
};

struct DecParams {
    DecLevel lv[FDMS_MAX_LEVELS];
    int n_levels, n_scales, bs, cap;
    int scale_first[FDMS_MAX_SCALES + 1];  // levels of scale s are lv[scale_first[s] .. scale_first[s+1])
    float conf_thres;
    float prod_pass, prod_fail;  // (1+e^-cls)(1+e^-obj) below / above which the fast decision is trusted
    float logit_lo;  // conservative lower bound on logit(conf_thres): x < logit_lo => sigmoid(x) <= thres
    float *cand;
    int32_t *cand_row;
    int32_t *count;
    const struct DecCta *cta_tab;  // per CTA of decode_filter_kernel (cached per geometry in the ctx)
    int aten;  // FDMS_NUMERICS_ATEN_CUDA: sigmoid with libdevice's expf, like ATen's CUDA kernel (see sigmoid_ref)
};

// What a CTA of decode_filter_kernel needs to know about itself.  Deriving it in the kernel (level search, two
// integer divisions, 64-bit plane offsets) cost about half of the instructions of a CTA that finds no candidate,
// i.e. of nearly every CTA; the table is built once per head geometry by decode_tab_kernel.
struct DecCta {
    long long elem_off;  // element offset of channel 0 of this CTA's (image, anchor) plane group in the level's data
    int level, cells;    // level index; cells per plane (ny * nx)
    int cell0;           // first cell of this CTA inside the plane
    int row_base;        // row index of cell 0 of this plane: row_offset + a * cells
    int seg;             // candidate segment: b * n_scales + scale
    int pad_;
};

// models/yolo.py:290-295 for one cell; v[k] = raw channel k (already fp32), out = 21 floats
// [cx, cy, w, h, obj, cls, (kx, ky, kc) x 5].  hf: the half-model arithmetic of ATen's CUDA kernels (common.cuh).
__device__ __forceinline__ void decode_box(const float (&v)[NO], float gx, float gy, float stride, float aw, float ah,
                                           float (&o)[NO], bool aten, bool hf) {
    if (hf) {
        const float s0 = h_sigmoid(v[0]), s1 = h_sigmoid(v[1]), s2 = h_sigmoid(v[2]), s3 = h_sigmoid(v[3]);
        o[0] = h_xy(s0, gx, stride);
        o[1] = h_xy(s1, gy, stride);
        o[2] = h_wh(s2, aw);
        o[3] = h_wh(s3, ah);
        o[4] = h_sigmoid(v[4]);
        o[5] = h_sigmoid(v[5]);
        return;
    }
    float s0 = sigmoid_ref(v[0], aten), s1 = sigmoid_ref(v[1], aten), s2 = sigmoid_ref(v[2], aten), s3 = sigmoid_ref(v[3], aten);
    o[0] = f_mul(f_add(f_sub(f_mul(s0, 2.f), 0.5f), gx), stride);
    o[1] = f_mul(f_add(f_sub(f_mul(s1, 2.f), 0.5f), gy), stride);
    float t2 = f_mul(s2, 2.f), t3 = f_mul(s3, 2.f);
    o[2] = f_mul(f_mul(t2, t2), aw);
    o[3] = f_mul(f_mul(t3, t3), ah);
    o[4] = sigmoid_ref(v[4], aten);
    o[5] = sigmoid_ref(v[5], aten);
}

__device__ __forceinline__ void decode_cell(const float (&v)[NO], float gx, float gy, float stride, float aw, float ah,
                                            float (&o)[NO], bool aten, bool hf) {
    decode_box(v, gx, gy, stride, aw, ah, o, aten, hf);
#pragma unroll
    for (int j = 0; j < 5; j++) {
        if (hf) {  // the float32 product is stored into the half tensor x_kpt (models/yolo.py:293-295)
            o[6 + 3 * j] = h_rn(h_xy(v[6 + 3 * j], gx, stride));
            o[7 + 3 * j] = h_rn(h_xy(v[7 + 3 * j], gy, stride));
            o[8 + 3 * j] = h_sigmoid(v[8 + 3 * j]);
        } else {
            o[6 + 3 * j] = f_mul(f_add(f_sub(f_mul(v[6 + 3 * j], 2.f), 0.5f), gx), stride);  // raw logits, no sigmoid
            o[7 + 3 * j] = f_mul(f_add(f_sub(f_mul(v[7 + 3 * j], 2.f), 0.5f), gy), stride);
            o[8 + 3 * j] = sigmoid_ref(v[8 + 3 * j], aten);
        }
    }
}

// utils/general.py:547-563 + 299-306 on one decoded row: o -> candidate row (in place)
__device__ __forceinline__ void row_to_candidate(float (&o)[NO], float conf) {
    float hw = f_div(o[2], 2.f), hh = f_div(o[3], 2.f);
    float cx = o[0], cy = o[1];
    o[0] = f_sub(cx, hw);
    o[1] = f_sub(cy, hh);
    o[2] = f_add(cx, hw);
    o[3] = f_add(cy, hh);
    o[4] = conf;
    o[5] = 0.f;  // class index j (nc == 1)
}

__device__ __forceinline__ float to_f32(float x) { return x; }
__device__ __forceinline__ float to_f32(__half x) { return __half2float(x); }

template <typename T>
__device__ __forceinline__ void load_vec8(const T *p, long long i, int n_valid, bool aligned, float (&v)[DEC_VEC]);

template <>
__device__ __forceinline__ void load_vec8<__half>(const __half *p, long long i, int n_valid, bool aligned,
                                                  float (&v)[DEC_VEC]) {
    if (aligned && n_valid == DEC_VEC) {
        uint4 raw = __ldg(reinterpret_cast<const uint4 *>(p + i));
        const __half2 *h = reinterpret_cast<const __half2 *>(&raw);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            float2 f = __half22float2(h[k]);
            v[2 * k] = f.x;
            v[2 * k + 1] = f.y;
        }
    } else {
#pragma unroll
        for (int k = 0; k < DEC_VEC; k++) v[k] = k < n_valid ? __half2float(__ldg(p + i + k)) : -INFINITY;
    }
}
template <>
__device__ __forceinline__ void load_vec8<float>(const float *p, long long i, int n_valid, bool aligned,
                                                 float (&v)[DEC_VEC]) {
    if (aligned && n_valid == DEC_VEC) {
        float4 a = __ldg(reinterpret_cast<const float4 *>(p + i));
        float4 b = __ldg(reinterpret_cast<const float4 *>(p + i + 4));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int k = 0; k < DEC_VEC; k++) v[k] = k < n_valid ? __ldg(p + i + k) : -INFINITY;
    }
}

// One full, 16-byte aligned vector of 8 cells of a plane, kept as loaded.
template <typename T> struct RawVec;
template <> struct RawVec<__half> {
    uint4 r;
    __device__ __forceinline__ void load(const __half *p) { r = __ldg(reinterpret_cast<const uint4 *>(p)); }
    __device__ __forceinline__ unsigned ge_mask(float lo) const {  // bit k: cell k >= lo (false for NaN)
        const __half2 *h = reinterpret_cast<const __half2 *>(&r);
        unsigned m = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float2 f = __half22float2(h[k]);
            m |= (f.x >= lo ? 1u : 0u) << (2 * k);
            m |= (f.y >= lo ? 1u : 0u) << (2 * k + 1);
        }
        return m;
    }
};
template <> struct RawVec<float> {
    float4 a, b;
    __device__ __forceinline__ void load(const float *p) {
        a = __ldg(reinterpret_cast<const float4 *>(p));
        b = __ldg(reinterpret_cast<const float4 *>(p + 4));
    }
    __device__ __forceinline__ unsigned ge_mask(float lo) const {
        return (a.x >= lo ? 1u : 0u) | (a.y >= lo ? 2u : 0u) | (a.z >= lo ? 4u : 0u) | (a.w >= lo ? 8u : 0u) |
               (b.x >= lo ? 16u : 0u) | (b.y >= lo ? 32u : 0u) | (b.z >= lo ? 64u : 0u) | (b.w >= lo ? 128u : 0u);
    }
};

// FUSE: 0 = only record the survivors' rows (decode_candidates_kernel decodes them afterwards); 1 / 2 = the CTA decodes
// the survivors it has just recorded itself, boxes only / with landmarks.  They are dense now (one per thread), belong
// to this CTA's (image, anchor) plane group -- no level search, no row arithmetic beyond one division -- and their
// obj / cls sectors are still in L1; the second kernel, its launch and its tail (9 us on a latency-bound chain) go away.
template <typename T, int FUSE>
__device__ __forceinline__ void decode_filter_body(const DecParams &P, const DecLevel &L, const DecCta &C, int *s_rows,
                                                   int &s_n, int &s_base) {
    const int cells = C.cells;
    const int cell0 = C.cell0 + (int)threadIdx.x * DEC_VEC;  // first cell of this thread's vector u is cell0 + u * DEC_GROUP
    const T *base = reinterpret_cast<const T *>(L.data) + C.elem_off;
    const bool aligned = (cells % DEC_VEC == 0) && ((uintptr_t)L.data % 16 == 0);

    // cheap necessary condition on the raw logits: sigmoid(obj) > thres and sigmoid(cls) > thres.  The kernel is
    // bound by the latency of these loads (13.5 MB in all): every thread issues its 2 * DEC_UNROLL vector loads
    // before it looks at any of them.
    unsigned pre = 0;  // bit u * DEC_VEC + k: cell cell0 + u * DEC_GROUP + k is pending
    // fp32 maps take twice the registers per vector: they go in two passes of half as many vectors.  Full aligned
    // vectors stay RAW (4 registers per fp16 vector) until their turn to be compared; ragged ones (a plane whose
    // size is not a multiple of 8, the last vector of a plane) are rare and compared as they are read.
    constexpr int UT = sizeof(T) == 2 ? DEC_UNROLL : (DEC_UNROLL > 1 ? DEC_UNROLL / 2 : 1);
    const float lo = P.logit_lo;
    const T *pobj = base + 4LL * cells, *pcls = base + 5LL * cells;
#pragma unroll
    for (int u0 = 0; u0 < DEC_UNROLL; u0 += UT) {
        RawVec<T> ro[UT], rc[UT];
        int nv[UT];
#pragma unroll
        for (int u = 0; u < UT; u++) {
            const int c0 = cell0 + (u0 + u) * DEC_GROUP;
            nv[u] = max(0, min(DEC_VEC, cells - c0));
            if (aligned && nv[u] == DEC_VEC) {
                ro[u].load(pobj + c0);
                rc[u].load(pcls + c0);
            }
        }
#pragma unroll
        for (int u = 0; u < UT; u++) {
            const int c0 = cell0 + (u0 + u) * DEC_GROUP;
            unsigned m = 0;
            if (aligned && nv[u] == DEC_VEC) {
                m = ro[u].ge_mask(lo) & rc[u].ge_mask(lo);
            } else if (nv[u] > 0) {
                float xo[DEC_VEC], xc[DEC_VEC];
                load_vec8<T>(pobj, c0, nv[u], false, xo);
                load_vec8<T>(pcls, c0, nv[u], false, xc);
#pragma unroll
                for (int k = 0; k < DEC_VEC; k++)
                    if (k < nv[u] && xo[k] >= lo && xc[k] >= lo) m |= 1u << k;
            }
            pre |= m << ((u0 + u) * DEC_VEC);
        }
    }
    // The few cells that get here (~1 % of a head) are decided DENSELY: every lane with a pending cell takes its
    // lowest one, so the warp runs the exponentials once per round (1-2 rounds) instead of once per cell index
    // whenever any lane passes.  conf = sigmoid(cls) * sigmoid(obj) > thres  <=>  (1 + e^-cls)(1 + e^-obj) < 1/thres:
    // two expf and no division; only a +-4e-6 relative band around the threshold needs the exact arithmetic
    // (general.py:547,563; the exact confidence itself is recomputed by decode_candidates_kernel, and
    // obj > thres (general.py:515) is implied by cls*obj > thres).
    // Survivors are only RECORDED (row index; the order inside a segment is irrelevant, the NMS sort key restores
    // the reference order): first in shared memory, then with ONE global atomic per CTA -- all CTAs of an image and
    // scale reserve slots in the same counter, and same-address atomics serialise in L2.
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const T *po = base + 4LL * cells + cell0, *pc = base + 5LL * cells + cell0;
    while (__any_sync(0xffffffffu, pre != 0)) {
        const bool have = pre != 0;
        const int bit = have ? __ffs(pre) - 1 : 0;
        const int k = (bit / DEC_VEC) * DEC_GROUP + (bit % DEC_VEC);  // cell offset from cell0
        pre &= pre - 1u;
        bool ok = false;
        if (have) {
            const float lo = to_f32(__ldg(po + k)), lc = to_f32(__ldg(pc + k));  // L1 hits: just read as vectors
            const float prod = f_mul(f_add(1.0f, expf(-lc)), f_add(1.0f, expf(-lo)));
            ok = prod < P.prod_pass;
            if (!ok && !(prod > P.prod_fail)) {
                const bool hf = P.aten != 0 && sizeof(T) == 2;
                const float sc = hf ? h_sigmoid(lc) : sigmoid_ref(lc, P.aten != 0), so = hf ? h_sigmoid(lo) : sigmoid_ref(lo, P.aten != 0);
                ok = f_mul(sc, so) > P.conf_thres;
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (m) {
            int basepos = 0;
            if (lane == 0) basepos = atomicAdd(&s_n, __popc(m));
            basepos = __shfl_sync(0xffffffffu, basepos, 0);
            if (ok) s_rows[basepos + __popc(m & ((1u << lane) - 1u))] = C.row_base + cell0 + k;
        }
    }
    __syncthreads();
    const int n_cta = s_n;
    if (n_cta == 0) return;  // CTA-uniform
    const int seg = C.seg;
    if (threadIdx.x == 0) s_base = atomicAdd(P.count + seg, n_cta);
    __syncthreads();
    const int a = (C.row_base - L.row_offset) / cells;  // the anchor of this CTA's plane group
    for (int i = threadIdx.x; i < n_cta; i += DEC_THREADS) {
        const int pos = s_base + i;
        if (pos >= P.cap) continue;
        const long long slot = (long long)seg * P.cap + pos;
        const int row = s_rows[i];
        P.cand_row[slot] = row;
        if (FUSE) {
            constexpr int NV = FUSE == 2 ? NO : 6;  // planes read / columns written
            const int cell = row - C.row_base;
            const int gy = cell / L.nx, gx = cell - gy * L.nx;
            const T *p = base + cell;
            float v[NO], o[NO];
#pragma unroll
            for (int c = NV; c < NO; c++) v[c] = 0.f;
            T raw[NV];
#pragma unroll
            for (int c = 0; c < NV; c++) raw[c] = __ldg(p + (long long)c * cells);  // all loads before the first conversion
#pragma unroll
            for (int c = 0; c < NV; c++) v[c] = to_f32(raw[c]);
            const bool hf = P.aten != 0 && sizeof(T) == 2;
            if (FUSE == 2) decode_cell(v, (float)gx, (float)gy, L.stride, L.anchors[2 * a], L.anchors[2 * a + 1], o, P.aten != 0, hf);
            else decode_box(v, (float)gx, (float)gy, L.stride, L.anchors[2 * a], L.anchors[2 * a + 1], o, P.aten != 0, hf);
            row_to_candidate(o, f_mul(o[5], o[4]));  // general.py:547: cls *= obj
            float *dst = P.cand + slot * NO;
#pragma unroll
            for (int c = 0; c < NV; c++) dst[c] = o[c];
        }
    }
}

__global__ void decode_tab_kernel(const __grid_constant__ DecParams P, DecCta *tab, int n_blocks) {
    const int blk = blockIdx.x * blockDim.x + threadIdx.x;
    if (blk >= n_blocks) return;
    int li = 0;
    for (int l = 1; l < P.n_levels; l++)
        if (blk >= P.lv[l].block_begin) li = l;
    const DecLevel &L = P.lv[li];
    const int local = blk - L.block_begin, planes = P.bs * NA;
    const int chunk = local / planes;  // chunk-major: neighbouring CTAs read different planes
    const int plane = local - chunk * planes;
    const int b = plane / NA, a = plane - b * NA;
    const long long cells = (long long)L.ny * L.nx;
    DecCta c;
    c.elem_off = ((long long)b * NA * NO + (long long)a * NO) * cells;
    c.level = li;
    c.cells = (int)cells;
    c.cell0 = chunk * DEC_CELLS;
    c.row_base = L.row_offset + a * (int)cells;
    c.seg = b * P.n_scales + L.scale;
    c.pad_ = 0;
    tab[blk] = c;
}

template <int FUSE>
__global__ void __launch_bounds__(DEC_THREADS) decode_filter_kernel(const __grid_constant__ DecParams P) {
    __shared__ int s_rows[DEC_CELLS];  // one copy for both instantiations of the body
    __shared__ int s_n, s_base;
    fdms_pdl_trigger();  // the NMS launched behind this kernel may be scheduled; it waits for this grid to complete
    const DecCta C = P.cta_tab[blockIdx.x];
    const DecLevel &L = P.lv[C.level];
    if (L.dtype == FDMS_F16) decode_filter_body<__half, FUSE>(P, L, C, s_rows, s_n, s_base);
    else decode_filter_body<float, FUSE>(P, L, C, s_rows, s_n, s_base);
}

// Second half of the fused decode: one thread per recorded candidate (dense warps after the compaction):
// 21 strided loads, exact sigmoids, models/yolo.py:290-295, then utils/general.py:547 (conf = cls*obj) and
// :299-306 (xywh2xyxy), one 84-byte row per thread.
constexpr int CAND_THREADS = 128;

template <bool KPT>
__global__ void __launch_bounds__(CAND_THREADS) decode_candidates_kernel(const __grid_constant__ DecParams P) {
    constexpr int NV = KPT ? NO : 6;  // planes read / columns written
    const int seg = blockIdx.y;
    const int n = min(P.count[seg], P.cap);
    const int b = seg / P.n_scales, scale = seg - b * P.n_scales;
    for (int i = blockIdx.x * CAND_THREADS + threadIdx.x; i < n; i += gridDim.x * CAND_THREADS) {
        const long long slot = (long long)seg * P.cap + i;
        const int row = P.cand_row[slot];
        int li = -1;  // level of this row inside its scale (rows are level-major)
#pragma unroll 1
        for (int l = P.scale_first[scale]; l < P.scale_first[scale + 1]; l++)
            if (row >= P.lv[l].row_offset) li = l;
        if (li < 0) continue;
        const DecLevel &L = P.lv[li];
        const int cells = L.ny * L.nx;
        const int r = row - L.row_offset;
        const int a = r / cells, cell = r - a * cells;
        const int gy = cell / L.nx, gx = cell - gy * L.nx;
        const long long base = ((long long)b * NA * NO + (long long)a * NO) * cells + cell;
        float v[NO], o[NO];
        // Every plane of a candidate is a different 32-byte DRAM sector (the head is plane-major), so this kernel
        // is bound by sector traffic: ~21 sectors per candidate with landmarks, 6 without.  All loads are issued
        // before the first conversion.
#pragma unroll
        for (int c = NV; c < NO; c++) v[c] = 0.f;
        if (L.dtype == FDMS_F16) {
            const __half *p = reinterpret_cast<const __half *>(L.data) + base;
            unsigned short raw[NV];
#pragma unroll
            for (int c = 0; c < NV; c++) raw[c] = __ldg(reinterpret_cast<const unsigned short *>(p + (long long)c * cells));
#pragma unroll
            for (int c = 0; c < NV; c++) v[c] = __half2float(__ushort_as_half(raw[c]));
        } else {
            const float *p = reinterpret_cast<const float *>(L.data) + base;
#pragma unroll
            for (int c = 0; c < NV; c++) v[c] = __ldg(p + (long long)c * cells);
        }
        const bool hf = P.aten != 0 && L.dtype == FDMS_F16;
        if (KPT) decode_cell(v, (float)gx, (float)gy, L.stride, L.anchors[2 * a], L.anchors[2 * a + 1], o, P.aten != 0, hf);
        else decode_box(v, (float)gx, (float)gy, L.stride, L.anchors[2 * a], L.anchors[2 * a + 1], o, P.aten != 0, hf);
        row_to_candidate(o, f_mul(o[5], o[4]));  // general.py:547: cls *= obj
        float *dst = P.cand + slot * NO;
#pragma unroll
        for (int c = 0; c < NV; c++) dst[c] = o[c];
    }
}

// ---- dense decode ---------------------------------------------------------------------------
struct DenseParams {
    DecLevel lv[FDMS_MAX_LEVELS];
    int n_levels, bs;
    long long rows_total;
    float *pred;
    int aten;  // see DecParams::aten
};

constexpr int DENSE_THREADS = 128;

__global__ void __launch_bounds__(DENSE_THREADS) decode_dense_kernel(const __grid_constant__ DenseParams P) {
    __shared__ float s_out[DENSE_THREADS * NO];
    int li = 0;
#pragma unroll 1
    for (int l = 1; l < P.n_levels; l++)
        if ((int)blockIdx.x >= P.lv[l].block_begin) li = l;
    const DecLevel &L = P.lv[li];
    const int local_block = blockIdx.x - L.block_begin;
    const int planes = P.bs * NA;
    const int chunk = local_block / planes, plane = local_block - chunk * planes;
    const int b = plane / NA, a = plane - b * NA;
    const long long cells = (long long)L.ny * L.nx;
    const long long cell = (long long)chunk * DENSE_THREADS + threadIdx.x;
    const long long cbase = ((long long)b * NA * NO + (long long)a * NO) * cells;
    if (cell < cells) {
        float v[NO], o[NO];
#pragma unroll
        for (int c = 0; c < NO; c++) v[c] = ld_f(L.data, L.dtype, cbase + (long long)c * cells + cell);
        int gy = (int)(cell / L.nx), gx = (int)(cell - (long long)gy * L.nx);
        decode_cell(v, (float)gx, (float)gy, L.stride, L.anchors[2 * a], L.anchors[2 * a + 1], o, P.aten != 0,
                    P.aten != 0 && L.dtype == FDMS_F16);
#pragma unroll
        for (int c = 0; c < NO; c++) s_out[threadIdx.x * NO + c] = o[c];  // stride 21 words: conflict free
    }
    __syncthreads();
    // coalesced write of this CTA's contiguous (<=128 rows x 21) block of pred
    const long long cell_first = (long long)chunk * DENSE_THREADS;
    const int n_rows = (int)min((long long)DENSE_THREADS, cells - cell_first);
    float *dst = P.pred + ((long long)b * P.rows_total + L.row_offset + (long long)a * cells + cell_first) * NO;
    for (int i = threadIdx.x; i < n_rows * NO; i += DENSE_THREADS) dst[i] = s_out[i];
}

// ---- dense decode of any detection head of models/yolo.py ------------------------------------
// Detect (:59-106), IDetect (:152-207) and IKeypoint (:267-308) share one inference branch; what differs between
// the head variants the reference can actually run is the column semantics:
//   nkpt > 0 (inplace branch, nc == 1, no == 6 + 3 nkpt): columns 0..5 sigmoid -> xy / wh / obj / cls, then per
//            keypoint (raw*2 - 0.5 + gx) * stride, (raw*2 - 0.5 + gy) * stride, sigmoid(conf);
//   nkpt == 0: every column through the sigmoid, xy / wh decoded, the rest (obj, nc classes) as they are -- the plain
//            YOLO head (the non-inplace branch :95-99 for nc > 1; the inplace branch for nc == 1, where x_kpt is empty).
// A CTA takes `blockDim.x` cells of one (image, anchor) plane group: channel planes are read coalesced (thread = cell),
// the (cells x no) block of pred is staged in shared memory with an odd pitch and written out contiguously.
struct HeadLevel {
    const void *data;
    int dtype, ny, nx, row_offset;
    int block_begin, chunks;
    float stride;
    float anchors[2 * FDMS_MAX_ANCHORS];
};

struct HeadParams {
    HeadLevel lv[FDMS_MAX_LEVELS];
    int n_levels, bs, na, no, nkpt, pitch;
    long long rows_total;
    float *pred;
    int aten;
};

__global__ void __launch_bounds__(128) decode_head_kernel(const __grid_constant__ HeadParams P) {
    extern __shared__ float s_rows[];  // [blockDim.x][pitch]
    int li = 0;
#pragma unroll 1
    for (int l = 1; l < P.n_levels; l++)
        if ((int)blockIdx.x >= P.lv[l].block_begin) li = l;
    const HeadLevel &L = P.lv[li];
    const int T = blockDim.x, tid = threadIdx.x, no = P.no, na = P.na;
    const int local_block = blockIdx.x - L.block_begin;
    const int planes = P.bs * na;
    const int chunk = local_block / planes, plane = local_block - chunk * planes;
    const int b = plane / na, a = plane - b * na;
    const long long cells = (long long)L.ny * L.nx;
    const long long cell_first = (long long)chunk * T, cell = cell_first + tid;
    const long long cbase = ((long long)b * na + a) * no * cells;
    const bool aten = P.aten != 0, hf = aten && L.dtype == FDMS_F16;  // hf: ATen's CUDA arithmetic for a half model (common.cuh)
    if (cell < cells) {
        const int gyi = (int)(cell / L.nx), gxi = (int)(cell - (long long)gyi * L.nx);
        const float gx = (float)gxi, gy = (float)gyi, stride = L.stride;
        const int n_sig = P.nkpt > 0 ? 6 : no;  // leading columns that go through the sigmoid
        float *row = s_rows + tid * P.pitch;
#pragma unroll 1
        for (int c = 0; c < no; c++) {
            const float v = ld_f(L.data, L.dtype, cbase + (long long)c * cells + cell);
            float o;
            if (c < n_sig) {
                const float sg = hf ? h_sigmoid(v) : sigmoid_ref(v, aten);
                if (c < 2) {  // :73 / :166 / :290
                    o = hf ? h_xy(sg, c == 0 ? gx : gy, stride) : f_mul(f_add(f_sub(f_mul(sg, 2.f), 0.5f), c == 0 ? gx : gy), stride);
                } else if (c < 4) {  // :74 / :167 / :291
                    const float t = f_mul(sg, 2.f);
                    o = hf ? h_wh(sg, L.anchors[2 * a + (c - 2)]) : f_mul(f_mul(t, t), L.anchors[2 * a + (c - 2)]);
                } else o = sg;
            } else {
                const int k = (c - 6) % 3;
                if (k == 2) o = hf ? h_sigmoid(v) : sigmoid_ref(v, aten);  // :81 / :183 / :295
                else if (hf) o = h_rn(h_xy(v, k == 0 ? gx : gy, stride));  // stored back into the half tensor x_kpt
                else o = f_mul(f_add(f_sub(f_mul(v, 2.f), 0.5f), k == 0 ? gx : gy), stride);  // raw logits, :76-77
            }
            row[c] = o;
        }
    }
    __syncthreads();
    const int n_rows = (int)min((long long)T, cells - cell_first);
    float *dst = P.pred + ((long long)b * P.rows_total + L.row_offset + (long long)a * cells + cell_first) * no;
    for (int i = tid; i < n_rows * no; i += T) {
        const int r = i / no, c = i - r * no;
        dst[i] = s_rows[r * P.pitch + c];
    }
}

// ---- filter on decoded pred -------------------------------------------------------------------
constexpr int FLT_THREADS = 256;

__global__ void __launch_bounds__(FLT_THREADS)
    filter_pred_kernel(const void *__restrict__ pred, int dtype, int bs, long long rows, int no, float thres,
                       float *__restrict__ cand, int32_t *__restrict__ cand_row, int32_t *__restrict__ count, int cap) {
    const int b = blockIdx.y;
    const long long r = (long long)blockIdx.x * FLT_THREADS + threadIdx.x;
    bool pass = false;
    float conf = 0.f;
    const long long base = ((long long)b * rows + r) * no;
    if (r < rows) {
        float obj = ld_f(pred, dtype, base + 4);
        if (obj > thres) {                                       // general.py:515
            conf = f_mul(ld_f(pred, dtype, base + 5), obj);      // :547
            pass = conf > thres;                                 // :563
        }
    }
    unsigned m = __ballot_sync(0xffffffffu, pass);
    if (m == 0) return;
    const int lane = threadIdx.x & 31;
    int basepos = 0;
    if (lane == 0) basepos = atomicAdd(count + b, __popc(m));
    basepos = __shfl_sync(0xffffffffu, basepos, 0);
    if (pass) {
        int pos = basepos + __popc(m & ((1u << lane) - 1u));
        if (pos < cap) {
            const float cx = ld_f(pred, dtype, base), cy = ld_f(pred, dtype, base + 1);
            const float hw = f_div(ld_f(pred, dtype, base + 2), 2.f), hh = f_div(ld_f(pred, dtype, base + 3), 2.f);
            float *dst = cand + ((long long)b * cap + pos) * no;
            dst[0] = f_sub(cx, hw);  // xywh2xyxy, general.py:299-306
            dst[1] = f_sub(cy, hh);
            dst[2] = f_add(cx, hw);
            dst[3] = f_add(cy, hh);
            dst[4] = conf;
            dst[5] = 0.f;  // class index (nc == 1)
            for (int c = 6; c < no; c++) dst[c] = ld_f(pred, dtype, base + c);  // keypoints pass through (:561)
            cand_row[(long long)b * cap + pos] = (int)r;
        }
    }
}

// Largest float L such that every x < L certainly has sigmoid(x) <= thres (margin 1e-3 in
// logit space dwarfs every rounding in the fp32 sigmoid).
float logit_lower_bound(float thres) {
    if (!(thres > 0.f)) return -INFINITY;
    if (thres >= 1.f) return INFINITY;
    double l = log((double)thres / (1.0 - (double)thres));
    return (float)(l - 1e-3 - 1e-6 * fabs(l));
}

int fill_levels(fdms_ctx *ctx, const fdms_level *levels, int n_levels, int n_scales, int bs, int cells_per_block,
                DecLevel *out, int *n_blocks, const char *who) {
    if (!levels || n_levels <= 0 || n_levels > FDMS_MAX_LEVELS || bs <= 0)
        return fdms_fail(ctx, FDMS_ERR_INVALID, "%s: bad level table", who);
    long long blocks = 0;
    for (int l = 0; l < n_levels; l++) {
        const fdms_level &in = levels[l];
        if (!in.data || in.ny <= 0 || in.nx <= 0 || (in.dtype != FDMS_F16 && in.dtype != FDMS_F32) || in.scale < 0 ||
            in.scale >= n_scales || in.row_offset < 0)
            return fdms_fail(ctx, FDMS_ERR_INVALID, "%s: bad level %d", who, l);
        if ((long long)in.ny * in.nx * NA + in.row_offset > 0x7fffffffLL)
            return fdms_fail(ctx, FDMS_ERR_UNSUPPORTED, "%s: level %d too large", who, l);
        DecLevel &L = out[l];
        L.data = in.data;
        L.dtype = in.dtype;
        L.ny = in.ny;
        L.nx = in.nx;
        L.scale = in.scale;
        L.row_offset = in.row_offset;
        L.stride = in.stride;
        for (int k = 0; k < 6; k++) L.anchors[k] = in.anchors[k];
        L.chunks = (int)ceil_div64((long long)in.ny * in.nx, cells_per_block);
        L.block_begin = (int)blocks;
        blocks += (long long)L.chunks * bs * NA;
        if (blocks > 0x7fffffffLL) return fdms_fail(ctx, FDMS_ERR_UNSUPPORTED, "%s: grid too large", who);
    }
    *n_blocks = (int)blocks;
    return FDMS_OK;
}

}  // namespace

extern "C" {

FDMS_API int fdms_decode_filter_ex(fdms_ctx *ctx, const fdms_level *levels, int n_levels, int n_scales, int bs,
                                   float conf_thres, float *cand, int32_t *cand_row, int32_t *count, int cap, int flags,
                                   fdms_stream stream) {
    FDMS_BIND(ctx);
    if (!cand || !cand_row || !count || cap <= 0 || n_scales <= 0 || n_scales > FDMS_MAX_SCALES)
        return fdms_fail(ctx, FDMS_ERR_INVALID, "fdms_decode_filter: bad arguments");
    DecParams P;
    memset(&P, 0, sizeof(P));
    P.aten = ctx->numerics == FDMS_NUMERICS_ATEN_CUDA;
    int n_blocks = 0;
    if (int rc = fill_levels(ctx, levels, n_levels, n_scales, bs, DEC_CELLS, P.lv, &n_blocks, "fdms_decode_filter"))
        return rc;
    P.n_levels = n_levels;
    P.n_scales = n_scales;
    P.bs = bs;
    P.cap = cap;
    P.conf_thres = conf_thres;
    // levels must be grouped by scale, ascending row_offset inside a scale (how every caller builds them)
    for (int sidx = 0, l = 0; sidx <= n_scales; sidx++) {
        while (l < n_levels && levels[l].scale < sidx) l++;
        P.scale_first[sidx] = l;
    }
    for (int l = 1; l < n_levels; l++)
        if (levels[l].scale < levels[l - 1].scale ||
            (levels[l].scale == levels[l - 1].scale && levels[l].row_offset <= levels[l - 1].row_offset))
            return fdms_fail(ctx, FDMS_ERR_INVALID, "fdms_decode_filter: levels must be sorted by (scale, row_offset)");
    // half-model arithmetic (aten_cuda numerics on fp16 maps): both sigmoids are rounded to half (2^-11 relative each)
    // before the product, so the fast decisions keep a 2e-3 band instead of 4e-6 and the logit bound moves accordingly
    bool hf = false;
    for (int l = 0; l < n_levels; l++) hf = hf || (P.aten && levels[l].dtype == FDMS_F16);
    const double band = hf ? 2e-3 : 4e-6;
    P.logit_lo = logit_lower_bound(hf && conf_thres > 0.f ? conf_thres * (1.0f - 2e-3f) : conf_thres);
    {   // conf > thres * (1 + band) certainly passes, conf < thres * (1 - band) certainly fails (the fast product is
        // good to ~3e-7 relative); non-positive thresholds have no fast "fail" side
        const double hi = (double)conf_thres + band * fabs((double)conf_thres) + 1e-37;
        const double lo = (double)conf_thres - band * fabs((double)conf_thres) - 1e-37;
        P.prod_pass = hi > 0.0 ? (float)(1.0 / hi) : INFINITY;
        P.prod_fail = lo > 0.0 ? (float)(1.0 / lo) : INFINITY;
    }
    P.cand = cand;
    P.cand_row = cand_row;
    P.count = count;
    cudaStream_t st = (cudaStream_t)stream;
    {   // per-CTA table, cached per head geometry (FNV-1a over everything it depends on)
        unsigned long long key = 1469598103934665603ull;
        auto mix = [&](long long v) {
            for (int q = 0; q < 8; q++) { key ^= (unsigned long long)(v >> (8 * q)) & 0xff; key *= 1099511628211ull; }
        };
        mix(n_levels); mix(n_scales); mix(bs); mix(n_blocks);
        for (int l = 0; l < n_levels; l++) { mix(levels[l].ny); mix(levels[l].nx); mix(levels[l].scale); mix(levels[l].row_offset); }
        if (int rc = fdms_reserve(ctx, ctx->dec_tab, (size_t)n_blocks * sizeof(DecCta))) return rc;
        mix((long long)(uintptr_t)ctx->dec_tab.p);
        if (key != ctx->dec_tab_key) {
            decode_tab_kernel<<<ceil_div(n_blocks, 256), 256, 0, st>>>(P, reinterpret_cast<DecCta *>(ctx->dec_tab.p), n_blocks);
            FDMS_LAUNCH_CHECK(ctx, "decode_tab_kernel");
            ctx->dec_tab_key = key;
        }
        P.cta_tab = reinterpret_cast<const DecCta *>(ctx->dec_tab.p);
    }
    FDMS_CUDA(ctx, cudaMemsetAsync(count, 0, sizeof(int32_t) * (size_t)bs * n_scales, st));
    const int n_seg = bs * n_scales;
    if (ctx->decode_fused) {  // filter + candidate decode in one launch
        if (flags & FDMS_DECODE_BOXES_ONLY) decode_filter_kernel<1><<<n_blocks, DEC_THREADS, 0, st>>>(P);
        else decode_filter_kernel<2><<<n_blocks, DEC_THREADS, 0, st>>>(P);
        FDMS_LAUNCH_CHECK(ctx, "decode_filter_kernel");
        return FDMS_OK;
    }
    decode_filter_kernel<0><<<n_blocks, DEC_THREADS, 0, st>>>(P);
    FDMS_LAUNCH_CHECK(ctx, "decode_filter_kernel");
    if (n_seg > 65535) return fdms_fail(ctx, FDMS_ERR_UNSUPPORTED, "fdms_decode_filter: more than 65535 segments");
    int gx = ceil_div(cap, CAND_THREADS);
    if (gx > 8) gx = 8;
    if (flags & FDMS_DECODE_BOXES_ONLY) decode_candidates_kernel<false><<<dim3(gx, n_seg), CAND_THREADS, 0, st>>>(P);
    else decode_candidates_kernel<true><<<dim3(gx, n_seg), CAND_THREADS, 0, st>>>(P);
    FDMS_LAUNCH_CHECK(ctx, "decode_candidates_kernel");
    return FDMS_OK;
}

FDMS_API int fdms_decode_filter(fdms_ctx *ctx, const fdms_level *levels, int n_levels, int n_scales, int bs,
                                float conf_thres, float *cand, int32_t *cand_row, int32_t *count, int cap,
                                fdms_stream stream) {
    return fdms_decode_filter_ex(ctx, levels, n_levels, n_scales, bs, conf_thres, cand, cand_row, count, cap, 0, stream);
}

FDMS_API int fdms_decode_dense(fdms_ctx *ctx, const fdms_level *levels, int n_levels, int bs, float *pred,
                               int64_t rows_total, fdms_stream stream) {
    FDMS_BIND(ctx);
    if (!pred || rows_total <= 0) return fdms_fail(ctx, FDMS_ERR_INVALID, "fdms_decode_dense: bad arguments");
    DenseParams P;
    memset(&P, 0, sizeof(P));
    P.aten = ctx->numerics == FDMS_NUMERICS_ATEN_CUDA;
    int n_blocks = 0;
    if (int rc = fill_levels(ctx, levels, n_levels, 1, bs, DENSE_THREADS, P.lv, &n_blocks, "fdms_decode_dense"))
        return rc;
    for (int l = 0; l < n_levels; l++)
        if ((int64_t)levels[l].row_offset + (int64_t)NA * levels[l].ny * levels[l].nx > rows_total)
            return fdms_fail(ctx, FDMS_ERR_INVALID, "fdms_decode_dense: level %d exceeds rows_total", l);
    P.n_levels = n_levels;
    P.bs = bs;
    P.rows_total = rows_total;
    P.pred = pred;
    decode_dense_kernel<<<n_blocks, DENSE_THREADS, 0, (cudaStream_t)stream>>>(P);
    FDMS_LAUNCH_CHECK(ctx, "decode_dense_kernel");
    return FDMS_OK;
}

FDMS_API int fdms_decode_dense_head(fdms_ctx *ctx, const fdms_level *levels, int n_levels, int bs, int na, int no, int nkpt,
                                    const float *anchors_px, float *pred, int64_t rows_total, fdms_stream stream) {
    FDMS_BIND(ctx);
    if (!pred || rows_total <= 0 || !anchors_px || !levels || n_levels <= 0 || n_levels > FDMS_MAX_LEVELS || bs <= 0)
        return fdms_fail(ctx, FDMS_ERR_INVALID, "fdms_decode_dense_head: bad arguments");
    if (na <= 0 || na > FDMS_MAX_ANCHORS || no < 6 || no > FDMS_MAX_HEAD_NO)
        return fdms_fail(ctx, FDMS_ERR_UNSUPPORTED, "fdms_decode_dense_head: na must be 1..%d and no 6..%d (got na=%d no=%d)",
                         FDMS_MAX_ANCHORS, FDMS_MAX_HEAD_NO, na, no);
    if (nkpt < 0 || (nkpt > 0 && no != 6 + 3 * nkpt))
        return fdms_fail(ctx, FDMS_ERR_INVALID,
                         "fdms_decode_dense_head: a keypoint head has no = 6 + 3 * nkpt columns (nc = 1); got no=%d nkpt=%d", no,
                         nkpt);
    HeadParams P;
    memset(&P, 0, sizeof(P));
    P.aten = ctx->numerics == FDMS_NUMERICS_ATEN_CUDA;
    P.pitch = no | 1;  // odd: the staging rows never share a bank pattern
    int T = 128;       // cells per CTA: as many as 48 KB of staging allow
    while (T > 32 && (size_t)T * P.pitch * sizeof(float) > 48 * 1024) T >>= 1;
    long long blocks = 0;
    for (int l = 0; l < n_levels; l++) {
        const fdms_level &in = levels[l];
        if (!in.data || in.ny <= 0 || in.nx <= 0 || (in.dtype != FDMS_F16 && in.dtype != FDMS_F32) || in.row_offset < 0)
            return fdms_fail(ctx, FDMS_ERR_INVALID, "fdms_decode_dense_head: bad level %d", l);
        if ((int64_t)in.row_offset + (int64_t)na * in.ny * in.nx > rows_total || rows_total > 0x7fffffffLL)
            return fdms_fail(ctx, FDMS_ERR_INVALID, "fdms_decode_dense_head: level %d exceeds rows_total", l);
        HeadLevel &L = P.lv[l];
        L.data = in.data;
        L.dtype = in.dtype;
        L.ny = in.ny;
        L.nx = in.nx;
        L.row_offset = in.row_offset;
        L.stride = in.stride;
        for (int k = 0; k < 2 * na; k++) L.anchors[k] = anchors_px[(size_t)l * 2 * na + k];
        L.chunks = (int)ceil_div64((long long)in.ny * in.nx, T);
        L.block_begin = (int)blocks;
        blocks += (long long)L.chunks * bs * na;
        if (blocks > 0x7fffffffLL) return fdms_fail(ctx, FDMS_ERR_UNSUPPORTED, "fdms_decode_dense_head: grid too large");
    }
    P.n_levels = n_levels;
    P.bs = bs;
    P.na = na;
    P.no = no;
    P.nkpt = nkpt;
    P.rows_total = rows_total;
    P.pred = pred;
    decode_head_kernel<<<(unsigned)blocks, T, (size_t)T * P.pitch * sizeof(float), (cudaStream_t)stream>>>(P);
    FDMS_LAUNCH_CHECK(ctx, "decode_head_kernel");
    return FDMS_OK;
}

FDMS_API int fdms_filter_pred(fdms_ctx *ctx, const void *pred, int dtype, int bs, int64_t rows, int no,
                              float conf_thres, float *cand, int32_t *cand_row, int32_t *count, int cap,
                              fdms_stream stream) {
    FDMS_BIND(ctx);
    if (!pred || !cand || !cand_row || !count || bs <= 0 || bs > 65535 || rows <= 0 || cap <= 0 || no < 6 ||
        (no - 6) % 3 != 0 || (dtype != FDMS_F16 && dtype != FDMS_F32) || rows > 0x7fffffffLL)
        return fdms_fail(ctx, FDMS_ERR_INVALID, "fdms_filter_pred: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    FDMS_CUDA(ctx, cudaMemsetAsync(count, 0, sizeof(int32_t) * (size_t)bs, st));
    dim3 grid((unsigned)ceil_div64(rows, FLT_THREADS), bs);
    filter_pred_kernel<<<grid, FLT_THREADS, 0, st>>>(pred, dtype, bs, rows, no, conf_thres, cand, cand_row, count, cap);
    FDMS_LAUNCH_CHECK(ctx, "filter_pred_kernel");
    return FDMS_OK;
}

}  // extern "C"
