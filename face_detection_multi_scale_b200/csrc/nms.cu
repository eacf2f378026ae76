// Batched NMS with torchvision.ops.nms semantics (call sites utils/general.py:584 and
// multi_scale_face_detector.py:234), all segments (image x scale) of a batch per launch and
// no host synchronisation between the stages:
//
//   nms_small_kernel : segments of <= 1024 boxes (the common case), one CTA per segment, everything in shared
//                      memory: register-sorted 32-key runs + binary-search ranks, then a blocked greedy
//                      sweep (kept-list test, balanced two-pass 256x256 pair masks, ballot resolver).
//   merge_small_kernel: the cross-scale merge of an image of <= 1024 rows in one CTA -- merge rows
//                      (merge_rows.cuh), the same NMS body, emit -- instead of three launches.
//   Larger segments:
//   nms_sort_kernel  : one CTA per segment.  64-bit key = (score descending, tie ascending),
//                      bitonic network in shared memory (<= 8192 items, else in L2-resident
//                      global scratch); emits item indices and boxes in score order and, in
//                      the last CTA to finish, the plan (mask offsets, tile prefix) for stage 2.
//   nms_mask_kernel  : persistent grid over (segment, 64-row block, 256-column group) tiles of
//                      the upper triangle; column boxes staged in shared memory; one 64-bit
//                      suppression word per (row, column block).
//   nms_sweep_kernel : one CTA per segment.  Warp 0 resolves a 64-box block with register
//                      shuffles, then the CTA ORs the mask rows of the boxes it kept into the
//                      `removed` bitmap in shared memory (coalesced row reads).
//
//   nms_greedy_kernel: >= 8 large segments: one CTA per segment, the blocked greedy sweep of the small kernel
//                      with the kept list streamed through shared-memory tiles (no K^2 memory).
//
// IoU arithmetic, selected by iou_mode:
//   TV_CPU : inter / ((area_i + area_j) - inter), suppressed iff (double)iou > thr (double)
//   TV_CUDA: inter / (fma(w_j, h_j, area_i) - inter), suppressed iff iou > (float)thr
//   NCNN   : +1 pixel on every extent, suppressed iff iou >= (float)thr (cpp/yolov7-face-ncnn/src/yolov7face.cpp:43)
// The double comparison is folded on the host into a float threshold plus an inclusive flag.
#include <math.h>

#include "common.cuh"
#include "merge_rows.cuh"
#include "nms.cuh"

namespace {

constexpr int SORT_THREADS = 1024;
constexpr int SORT_SMEM_ITEMS = 8192;  // 8 B key + 4 B index each (+ 1/16 skew of the register-blocked sort) = 102 KB
constexpr int SORT_SMEM_SLOTS = SORT_SMEM_ITEMS + SORT_SMEM_ITEMS / 16;
constexpr int MASK_THREADS = 256;      // 64 rows x 4 column blocks
constexpr int SWEEP_THREADS = 512;
constexpr int SWEEP_MAX_NBLK = 4096;   // 262144 boxes per segment
constexpr int SMALL_THREADS = 1024;
// Register cap of the single-CTA kernels.  A resident 1024-thread CTA leaves the rest of the SM's register file to the
// CTAs of whatever runs beside it -- in the pipelined path the HBM-bound pyramid kernel of the next batch, 7680
// registers per CTA: at 46-48 registers (what __launch_bounds__(1024) gives) two of them fit next to an NMS CTA, at
// 40 three, at 32 four.  Measured on cfg2 with both streams running: 263.1 / 266.8 / 270.1 k images/s for 48 / 40 / 32
// (the pyramid launches inside the timed region 0.1224 -> 0.1198 ms); 24 registers buy nothing more (the NMS itself
// slows from 26.6 to 28.6 us on 76 spilled bytes).  At 32 the kernels spill 4-12 bytes and take 26.6 instead of 26.3 us.
#ifndef FDMS_NMS_SMALL_REGS
#define FDMS_NMS_SMALL_REGS 32
#endif
#if FDMS_NMS_SMALL_REGS > 0
#define FDMS_SMALL_BOUNDS __maxnreg__(FDMS_NMS_SMALL_REGS)
#else
#define FDMS_SMALL_BOUNDS __launch_bounds__(1024)
#endif
constexpr int SMALL_LIMIT = 1024;      // largest segment the fused single-CTA kernel takes (one box per thread)
constexpr int GREEDY_THREADS = 1024;
constexpr int GREEDY_BS = 256;         // candidates per greedy step
constexpr int GREEDY_TILE = 1024;      // kept boxes staged per tile (large-segment kernel)

struct SegInfo {
    int n;               // boxes entering suppression (after the max_nms cut)
    int nblk;            // ceil(n / 64)
    long long mask_off;  // first mask word of this segment, -1: does not fit the mask budget
    long long unit_begin;  // first mask tile of this segment
};

struct PlanHeader {
    long long total_units;
    unsigned int done;  // CTAs of the sort kernel that have finished (reset by the last one)
    unsigned int pad;
};

__device__ __forceinline__ uint32_t score_key_desc(float s) {
    uint32_t b = __float_as_uint(s);
    uint32_t asc = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return ~asc;
}

// Register-blocked bitonic sort of one segment (2048 / 4096 / 8192 padded keys, E = 2 / 4 / 8 consecutive keys per thread).
// The network is the plain one (for k = 2 .. pad, for j = k/2 .. 1: compare-exchange i with i ^ j, ascending where
// (i & k) == 0), but only the strides that cross warps (j >= 32 E) go through shared memory: strides below E are
// compare-exchanges between a thread's own registers, strides E .. 16 E shuffles between lanes.  8192 keys: 15 of the 91
// stages (and 20 instead of 91 barriers) touch shared memory.  Shared-memory slots are skewed by one every 16 keys so
// that a thread's E consecutive 64-bit keys do not fall on two banks for the whole warp.
__device__ __forceinline__ int sort_slot(int i) { return i + (i >> 4); }

// PACKED: the item index rides in the low bits of the key (score << 32 | tie << log2(PAD) | index), possible when every
// tie of the segment is below 2^(32 - log2(PAD)) -- head row numbers and merge rows always are: two shuffles per key and
// stage instead of three, no index select, shared-memory stages on the keys alone.  The order is the same: ties are
// unique inside a segment, so the index bits never decide.
template <int E, bool PACKED>
__device__ __forceinline__ void sort_reg_stages(unsigned long long (&key)[E], uint32_t (&idx)[E], int t, int k, int j_start) {
    for (int j = j_start; j >= E; j >>= 1) {  // between lanes
        const int lm = j / E;
#pragma unroll
        for (int r = 0; r < E; r++) {
            const int i = t * E + r;
            const unsigned long long ok = __shfl_xor_sync(0xffffffffu, key[r], lm);
            const bool take_min = ((i & j) == 0) == ((i & k) == 0);
            // branch-free (the lanes of a warp disagree about nearly every exchange); keys are unique but for the padding,
            // whose index is the same on both sides
            const unsigned long long lo = ok < key[r] ? ok : key[r], hi = ok < key[r] ? key[r] : ok;
            const unsigned long long nk = take_min ? lo : hi;
            if (!PACKED) {
                const uint32_t oi = __shfl_xor_sync(0xffffffffu, idx[r], lm);
                idx[r] = nk == ok ? oi : idx[r];
            }
            key[r] = nk;
        }
    }
#pragma unroll
    for (int j = E / 2; j > 0; j >>= 1) {  // inside the thread
        if (j > j_start) continue;
#pragma unroll
        for (int r = 0; r < E; r++) {
            if (r & j) continue;
            const bool up = ((t * E + r) & k) == 0;
            const bool sw = (key[r] > key[r | j]) == up;
            const unsigned long long ka = key[r], kb = key[r | j];
            key[r] = sw ? kb : ka;
            key[r | j] = sw ? ka : kb;
            if (!PACKED) {
                const uint32_t ia = idx[r], ib = idx[r | j];
                idx[r] = sw ? ib : ia;
                idx[r | j] = sw ? ia : ib;
            }
        }
    }
}

template <int E, bool PACKED>
__device__ __forceinline__ void sort_segment_body(unsigned long long (&key)[E], uint32_t (&idx)[E], unsigned long long *keys,
                                                  uint32_t *idxs) {
    constexpr int PAD = E * SORT_THREADS;
    const int t = threadIdx.x;
    for (int k = 2; k <= PAD; k <<= 1) {
        int j = k >> 1;
        if (j >= 32 * E) {
#pragma unroll
            for (int r = 0; r < E; r++) {
                keys[sort_slot(t * E + r)] = key[r];
                if (!PACKED) idxs[sort_slot(t * E + r)] = idx[r];
            }
            __syncthreads();
            for (; j >= 32 * E; j >>= 1) {
                for (int c = t; c < PAD / 2; c += SORT_THREADS) {
                    const int i = ((c & ~(j - 1)) << 1) | (c & (j - 1)), p = i + j;
                    const bool up = (i & k) == 0;
                    const unsigned long long a = keys[sort_slot(i)], b = keys[sort_slot(p)];
                    if ((a > b) == up) {
                        keys[sort_slot(i)] = b;
                        keys[sort_slot(p)] = a;
                        if (!PACKED) {
                            const uint32_t ia = idxs[sort_slot(i)], ib = idxs[sort_slot(p)];
                            idxs[sort_slot(i)] = ib;
                            idxs[sort_slot(p)] = ia;
                        }
                    }
                }
                __syncthreads();
            }
#pragma unroll
            for (int r = 0; r < E; r++) {
                key[r] = keys[sort_slot(t * E + r)];
                if (!PACKED) idx[r] = idxs[sort_slot(t * E + r)];
            }
        }
        sort_reg_stages<E, PACKED>(key, idx, t, k, j);
    }
}

template <int E>
__device__ __forceinline__ void sort_segment_regs(const float *__restrict__ seg_rows, int row_stride, int box_off, int score_off,
                                                  const int32_t *__restrict__ seg_tie, int n, int K,
                                                  unsigned long long *keys, uint32_t *idxs, uint32_t *__restrict__ out_idx,
                                                  float4 *__restrict__ out_boxes) {
    constexpr int PAD = E * SORT_THREADS;
    constexpr int IDX_BITS = E == 2 ? 11 : (E == 4 ? 12 : 13);
    static_assert((1 << IDX_BITS) == PAD, "index bits");
    const int t = threadIdx.x;
    unsigned long long key[E];
    uint32_t idx[E], tie_v[E];
    uint32_t sk[E];
    bool wide = false;  // a tie that does not leave room for the index bits
#pragma unroll
    for (int r = 0; r < E; r++) {
        const int i = t * E + r;
        sk[r] = 0xffffffffu;
        tie_v[r] = 0u;
        idx[r] = 0xffffffffu;
        if (i < n) {
            const float sc = seg_rows[(long long)i * row_stride + score_off];
            tie_v[r] = seg_tie ? (uint32_t)seg_tie[i] : (uint32_t)i;
            sk[r] = score_key_desc(sc);
            idx[r] = (uint32_t)i;
            wide |= tie_v[r] >= (1u << (32 - IDX_BITS));
        }
    }
    const bool packed = !__syncthreads_or(wide ? 1 : 0);
#pragma unroll
    for (int r = 0; r < E; r++) {
        const int i = t * E + r;
        if (i < n) key[r] = ((unsigned long long)sk[r] << 32) | (packed ? ((tie_v[r] << IDX_BITS) | (uint32_t)i) : tie_v[r]);
        else key[r] = ~0ull;
    }
    if (packed) {
        sort_segment_body<E, true>(key, idx, keys, idxs);
#pragma unroll
        for (int r = 0; r < E; r++) idx[r] = (uint32_t)key[r] & (uint32_t)(PAD - 1);
    } else {
        sort_segment_body<E, false>(key, idx, keys, idxs);
    }
#pragma unroll
    for (int r = 0; r < E; r++) {
        const int kk = t * E + r;
        if (kk < K) {
            const float *rw = seg_rows + (long long)idx[r] * row_stride + box_off;
            out_idx[kk] = idx[r];
            out_boxes[kk] = make_float4(rw[0], rw[1], rw[2], rw[3]);
        }
    }
}

__global__ void __launch_bounds__(SORT_THREADS)
    nms_sort_kernel(const float *__restrict__ rows, int row_stride, int box_off, int score_off,
                    const int32_t *__restrict__ tie, const int32_t *__restrict__ count, int n_seg, int cap, int cap_e,
                    int max_nms, unsigned long long *gkeys, uint32_t *gidx, long long gpad,
                    uint32_t *__restrict__ sorted_idx, float4 *__restrict__ sorted_boxes, SegInfo *seginfo,
                    PlanHeader *hdr, long long mask_budget_words, int skip_le, int reg_sort) {
    extern __shared__ __align__(16) uint8_t sm[];
    __shared__ int s_last;
    __shared__ long long s_carry[2];
    __shared__ long long s_warp[2][32];
    const int g = blockIdx.x, tid = threadIdx.x;
    int n = count[g];
    n = max(0, min(n, cap));
    if (n <= skip_le) n = 0;  // handled by nms_small_kernel
    int pad = 64;
    while (pad < n) pad <<= 1;
    unsigned long long *keys;
    uint32_t *idx;
    if (pad <= SORT_SMEM_ITEMS) {
        keys = reinterpret_cast<unsigned long long *>(sm);
        idx = reinterpret_cast<uint32_t *>(sm + (size_t)SORT_SMEM_SLOTS * 8);
    } else {
        keys = gkeys + (long long)g * gpad;
        idx = gidx + (long long)g * gpad;
    }
    const float *seg_rows = rows + (long long)g * cap * row_stride;
    const int K = (max_nms > 0) ? min(n, max_nms) : n;  // utils/general.py:578-579
    if (reg_sort && pad >= 2 * SORT_THREADS && pad <= SORT_SMEM_ITEMS) {
        const int32_t *seg_tie = tie ? tie + (long long)g * cap : nullptr;
        uint32_t *o_idx = sorted_idx + (long long)g * cap_e;
        float4 *o_box = sorted_boxes + (long long)g * cap_e;
        uint32_t *idxs = reinterpret_cast<uint32_t *>(sm + (size_t)SORT_SMEM_SLOTS * 8);
        if (pad == 2 * SORT_THREADS) sort_segment_regs<2>(seg_rows, row_stride, box_off, score_off, seg_tie, n, K, keys, idxs, o_idx, o_box);
        else if (pad == 4 * SORT_THREADS) sort_segment_regs<4>(seg_rows, row_stride, box_off, score_off, seg_tie, n, K, keys, idxs, o_idx, o_box);
        else sort_segment_regs<8>(seg_rows, row_stride, box_off, score_off, seg_tie, n, K, keys, idxs, o_idx, o_box);
    } else {
        for (int i = tid; i < pad; i += SORT_THREADS) {
            if (i < n) {
                float s = seg_rows[(long long)i * row_stride + score_off];
                uint32_t t = tie ? (uint32_t)tie[(long long)g * cap + i] : (uint32_t)i;
                keys[i] = ((unsigned long long)score_key_desc(s) << 32) | t;
                idx[i] = (uint32_t)i;
            } else {
                keys[i] = ~0ull;
                idx[i] = 0xffffffffu;
            }
        }
        __syncthreads();
        if (n > 1) {
            for (int k = 2; k <= pad; k <<= 1) {
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int t = tid; t < (pad >> 1); t += SORT_THREADS) {
                        int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                        int p = i + j;
                        bool up = (i & k) == 0;
                        unsigned long long a = keys[i], b = keys[p];
                        if ((a > b) == up) {
                            keys[i] = b;
                            keys[p] = a;
                            uint32_t ia = idx[i], ib = idx[p];
                            idx[i] = ib;
                            idx[p] = ia;
                        }
                    }
                    __syncthreads();
                }
            }
        }
        for (int k = tid; k < K; k += SORT_THREADS) {
            uint32_t ii = idx[k];
            const float *r = seg_rows + (long long)ii * row_stride + box_off;
            sorted_idx[(long long)g * cap_e + k] = ii;
            sorted_boxes[(long long)g * cap_e + k] = make_float4(r[0], r[1], r[2], r[3]);
        }
    }
    if (tid == 0) {
        seginfo[g].n = K;
        seginfo[g].nblk = (K + 63) >> 6;
    }
    // ---- last CTA computes the plan (exclusive prefix of mask words and mask tiles) ---------
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        unsigned int t = atomicAdd(&hdr->done, 1u);
        s_last = (t == (unsigned)n_seg - 1u);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (tid == 0) {
        s_carry[0] = 0;
        s_carry[1] = 0;
    }
    __syncthreads();
    const int lane = tid & 31, warp = tid >> 5;
    for (int base = 0; base < n_seg; base += SORT_THREADS) {
        int s = base + tid;
        long long v[2] = {0, 0};
        if (s < n_seg) {
            volatile SegInfo *si = seginfo + s;
            long long kk = si->n, nb = si->nblk;
            v[0] = kk * nb;              // mask words: one row of nblk words per box
            v[1] = nb * ((nb + 3) >> 2);  // mask tiles: 64 rows x 256 columns
        }
        long long inc[2] = {v[0], v[1]};
#pragma unroll
        for (int q = 0; q < 2; q++) {
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                long long t = __shfl_up_sync(0xffffffffu, inc[q], d);
                if (lane >= d) inc[q] += t;
            }
            if (lane == 31) s_warp[q][warp] = inc[q];
        }
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int q = 0; q < 2; q++) {
                long long w = s_warp[q][lane];
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    long long t = __shfl_up_sync(0xffffffffu, w, d);
                    if (lane >= d) w += t;
                }
                s_warp[q][lane] = w;  // inclusive over warps
            }
        }
        __syncthreads();
        long long excl[2];
#pragma unroll
        for (int q = 0; q < 2; q++) excl[q] = s_carry[q] + (warp ? s_warp[q][warp - 1] : 0) + inc[q] - v[q];
        if (s < n_seg) {
            seginfo[s].mask_off = (excl[0] + v[0] <= mask_budget_words) ? excl[0] : -1;
            seginfo[s].unit_begin = excl[1];
        }
        __syncthreads();
        if (tid == SORT_THREADS - 1) {
            s_carry[0] += s_warp[0][31];
            s_carry[1] += s_warp[1][31];
        }
        __syncthreads();
    }
    if (tid == 0) {
        seginfo[n_seg].n = 0;
        seginfo[n_seg].nblk = 0;
        seginfo[n_seg].mask_off = -1;
        seginfo[n_seg].unit_begin = s_carry[1];
        hdr->total_units = s_carry[1];
        hdr->done = 0;
    }
}

// IoU test of box b (later in score order) against box a.  Disjoint boxes are rejected before any
// arithmetic: inter == 0 gives iou == 0 (or NaN when the union is empty too), which never exceeds a
// threshold >= 0 -- exact, not an approximation.  `quick` is false for thresholds where that does not hold.
template <int MODE>
__device__ __forceinline__ bool iou_gt(const float4 a, const float4 b, float thr, bool inclusive, bool quick) {
    float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
    float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
    if (MODE == FDMS_IOU_NCNN) {
        // cpp/yolov7-face-ncnn/src/yolov7face.cpp:43-74: +1 pixel convention on every extent, `ovr >= thresh`
        float w = fmaxf(0.f, f_add(f_sub(xx2, xx1), 1.f)), h = fmaxf(0.f, f_add(f_sub(yy2, yy1), 1.f));
        if (quick && (!(w > 0.f) || !(h > 0.f))) return false;
        float inter = f_mul(w, h);
        float area_a = f_mul(f_add(f_sub(a.z, a.x), 1.f), f_add(f_sub(a.w, a.y), 1.f));
        float area_b = f_mul(f_add(f_sub(b.z, b.x), 1.f), f_add(f_sub(b.w, b.y), 1.f));
        float iou = f_div(inter, f_sub(f_add(area_a, area_b), inter));
        return iou >= thr;
    }
    if (quick && (!(xx2 > xx1) || !(yy2 > yy1))) return false;
    float w = fmaxf(f_sub(xx2, xx1), 0.f), h = fmaxf(f_sub(yy2, yy1), 0.f);
    float inter = f_mul(w, h);
    float area_a = f_mul(f_sub(a.z, a.x), f_sub(a.w, a.y));
    float bw = f_sub(b.z, b.x), bh = f_sub(b.w, b.y);
    float den;
    if (MODE == FDMS_IOU_TV_CUDA) den = f_sub(__fmaf_rn(bw, bh, area_a), inter);
    else den = f_sub(f_add(area_a, f_mul(bw, bh)), inter);
    float iou = f_div(inter, den);
    return inclusive ? (iou >= thr) : (iou > thr);  // NaN: false either way
}

// Pass 1 of every pair loop: do two boxes overlap at all?  This is iou_gt's own quick-reject test, evaluated
// branch-free for a batch of boxes; only the survivors take the full IoU (pass 2).  The lanes of a warp hold
// different boxes, so a fused loop would run the long path whenever ANY lane overlaps the broadcast box.
template <int MODE>
__device__ __forceinline__ bool boxes_overlap(const float4 a, const float4 b) {
    if (MODE == FDMS_IOU_NCNN)  // extents carry +1: boxes up to one pixel apart still intersect
        return (f_add(f_sub(fminf(a.z, b.z), fmaxf(a.x, b.x)), 1.f) > 0.f) &&
               (f_add(f_sub(fminf(a.w, b.w), fmaxf(a.y, b.y)), 1.f) > 0.f);
    return (fminf(a.z, b.z) > fmaxf(a.x, b.x)) && (fminf(a.w, b.w) > fmaxf(a.y, b.y));
}

template <int MODE>
__global__ void __launch_bounds__(MASK_THREADS)
    nms_mask_kernel(const SegInfo *__restrict__ seginfo, const PlanHeader *__restrict__ hdr, int n_seg, int cap_e,
                    const float4 *__restrict__ sorted_boxes, unsigned long long *__restrict__ mask, float thr,
                    int inclusive) {
    __shared__ float4 s_col[MASK_THREADS];
    const bool quick = thr > 0.f || (thr == 0.f && !inclusive);
    const long long total = hdr->total_units;
    const int tid = threadIdx.x;
    int g = 0;
    long long g_begin = 0, g_end = seginfo[1].unit_begin;
    for (long long u = blockIdx.x; u < total; u += gridDim.x) {
        while (u >= g_end) {  // units are visited in increasing order: walk the segment cursor
            g++;
            g_begin = g_end;
            g_end = seginfo[g + 1].unit_begin;
        }
        const SegInfo si = seginfo[g];
        if (si.mask_off < 0) continue;
        const int ncg = (si.nblk + 3) >> 2;
        const int local = (int)(u - g_begin);
        const int rb = local / ncg, cg = local - rb * ncg;
        const int cb0 = cg * 4;
        if (cb0 + 3 < rb) continue;  // entirely below the diagonal
        const float4 *boxes = sorted_boxes + (long long)g * cap_e;
        __syncthreads();  // previous tile's readers are done with s_col
        {
            int col = cb0 * 64 + tid;
            s_col[tid] = col < si.n ? boxes[col] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncthreads();
        const int row = rb * 64 + (tid & 63);
        const int q = tid >> 6, cblk = cb0 + q;
        if (row < si.n && cblk < si.nblk && cblk >= rb) {
            const float4 a = boxes[row];
            unsigned long long bits = 0;
            const int col0 = cblk * 64;
            const int jn = min(64, si.n - col0);
            const int j0 = (cblk == rb) ? (tid & 63) + 1 : 0;  // only later boxes can be suppressed
            unsigned long long ov = 0ull;
#pragma unroll 8
            for (int j = 0; j < 64; j++) ov |= boxes_overlap<MODE>(a, s_col[q * 64 + j]) ? (1ull << j) : 0ull;
            if (!quick) ov = ~0ull;
            ov &= (jn >= 64 ? ~0ull : ((1ull << jn) - 1ull)) & (j0 >= 64 ? 0ull : (~0ull << j0));
            while (ov) {
                const int j = __ffsll((long long)ov) - 1;
                ov &= ov - 1ull;
                if (iou_gt<MODE>(a, s_col[q * 64 + j], thr, inclusive != 0, false)) bits |= 1ull << j;
            }
            mask[si.mask_off + (long long)row * si.nblk + cblk] = bits;
        }
    }
}

__global__ void __launch_bounds__(SWEEP_THREADS)
    nms_sweep_kernel(const SegInfo *__restrict__ seginfo, const unsigned long long *__restrict__ mask, int cap_e,
                     const uint32_t *__restrict__ sorted_idx, int32_t *__restrict__ keep, int32_t *__restrict__ n_keep,
                     int cap, const int32_t *__restrict__ count, int skip_le) {
    __shared__ unsigned long long s_removed[SWEEP_MAX_NBLK];
    __shared__ unsigned long long s_diag[64];
    __shared__ unsigned long long s_kept;
    __shared__ int s_nkeep;
    const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    if (max(0, min(count[g], cap)) <= skip_le) return;  // handled by nms_small_kernel
    const SegInfo si = seginfo[g];
    if (si.mask_off < 0 && si.n > 0) {  // did not fit the mask budget: report, never guess
        if (tid == 0) n_keep[g] = -1;
        return;
    }
    const int K = si.n, nblk = si.nblk;
    for (int w = tid; w < nblk; w += SWEEP_THREADS) s_removed[w] = 0ull;
    if (tid == 0) s_nkeep = 0;
    __syncthreads();
    const unsigned long long *M = mask + si.mask_off;
    const uint32_t *sidx = sorted_idx + (long long)g * cap_e;
    int32_t *kout = keep + (long long)g * cap;
    for (int c = 0; c < nblk; c++) {
        if (tid < 32) {
            const int r_lo = c * 64 + lane, r_hi = r_lo + 32;
            s_diag[lane] = r_lo < K ? M[(long long)r_lo * nblk + c] : 0ull;
            s_diag[lane + 32] = r_hi < K ? M[(long long)r_hi * nblk + c] : 0ull;
            __syncwarp();
            const int left = K - c * 64;
            const unsigned long long valid = left >= 64 ? ~0ull : ((1ull << left) - 1ull);
            // one step per KEPT box: the diagonal words only hold later columns
            unsigned long long alive = ~s_removed[c] & valid, kept = 0ull;
            while (alive) {
                const int q = __ffsll((long long)alive) - 1;
                kept |= 1ull << q;
                alive &= ~(s_diag[q] | (1ull << q));
            }
            const int base = s_nkeep;
            if ((kept >> lane) & 1ull) kout[base + __popcll(kept & ((1ull << lane) - 1ull))] = (int32_t)sidx[r_lo];
            if ((kept >> (lane + 32)) & 1ull)
                kout[base + __popcll(kept & ((1ull << (lane + 32)) - 1ull))] = (int32_t)sidx[r_hi];
            __syncwarp();
            if (lane == 0) {
                s_kept = kept;
                s_nkeep = base + __popcll(kept);
            }
        }
        __syncthreads();
        const unsigned long long kept = s_kept;
        if (kept != 0ull && c + 1 < nblk) {
            for (int w = c + 1 + tid; w < nblk; w += SWEEP_THREADS) {
                unsigned long long acc = 0ull;
                const unsigned long long *col = M + (long long)c * 64 * nblk + w;
                unsigned long long kb = kept;
#pragma unroll 4
                for (int i = 0; i < 64; i++) {
                    if ((kb >> i) & 1ull) acc |= col[(long long)i * nblk];
                }
                s_removed[w] |= acc;
            }
        }
        __syncthreads();
    }
    if (tid == 0) n_keep[g] = s_nkeep;
}

// Steps (2)+(3) of a blocked greedy step, shared by nms_small_kernel and nms_greedy_kernel.
// blk[0..nb) = the step's candidates in score order (shared memory); s_sup = candidates already suppressed by
// boxes kept in earlier steps.  (2) column masks: colmask[i] = earlier candidates j < i of this block that
// would suppress i (pairs with a pre-suppressed member are skipped -- they cannot matter).  (3) resolution to the
// greedy result by warp 0, chunk by chunk (see below).  Result: s_K = bitset of kept candidates.  All CTA threads
// must call this (it contains barriers).
template <int MODE>
__device__ __forceinline__ void greedy_block(const float4 *blk, int nb, float thr, bool inclusive, bool quick,
                                             unsigned long long (*s_colmask)[4], unsigned long long *s_sup,
                                             unsigned long long *s_K) {
    const int tid = threadIdx.x;
    __syncthreads();  // s_sup complete
    {
        // Work units = (chunk of 32 candidates) x (16 earlier candidates), lower triangle only, dealt round-robin to
        // the 32 warps: every warp gets the same number of units (+-1) instead of the triangular imbalance of one
        // thread per (candidate, 64-column group).  lane = candidate; the earlier box is a broadcast read.
        uint16_t (*cm16)[16] = reinterpret_cast<uint16_t (*)[16]>(s_colmask);  // [i][16]: same bits, little endian
        const int warp = tid >> 5, lane = tid & 31;
        const int nt = (nb + 31) >> 5, n_units = nt * (nt + 1);  // chunk ti has 2 * ti + 2 units
        for (int u = warp; u < n_units; u += GREEDY_THREADS / 32) {
            int ti = 0, tj = u;
            while (tj >= 2 * ti + 2) {
                tj -= 2 * ti + 2;
                ti++;
            }
            const int ci = ti * 32 + lane, j0 = tj * 16;
            unsigned bits = 0u;
            if (ci < nb && !((s_sup[ci >> 6] >> (ci & 63)) & 1ull)) {
                const float4 bi = blk[ci];
                const int jn = min(min(16, ci - j0), nb - j0);  // only earlier candidates
                // pass 1, branch-free: which of the 16 boxes overlap this one at all (the exact quick-reject test);
                // pass 2: the full IoU only for those.  Lanes of a warp are different candidates, so a fused loop
                // would run the long path whenever ANY lane overlaps the broadcast box.
                unsigned ov = 0u;
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    // blk[j0 + j] is in bounds of the block array for every j0; masked below
                    ov |= boxes_overlap<MODE>(blk[j0 + j], bi) ? (1u << j) : 0u;
                }
                if (!quick) ov = 0xffffu;  // thresholds below zero: every pair takes the exact path
                ov &= jn > 0 ? ((1u << jn) - 1u) : 0u;
                ov &= ~(unsigned)(s_sup[j0 >> 6] >> (j0 & 63));  // pre-suppressed boxes suppress nothing
                while (ov) {
                    const int j = __ffs(ov) - 1;
                    ov &= ov - 1u;
                    if (iou_gt<MODE>(blk[j0 + j], bi, thr, inclusive, false)) bits |= 1u << j;
                }
            }
            cm16[ci][tj] = (uint16_t)bits;
            if (tj == 2 * ti + 1)  // last unit of the chunk: clear the words right of the diagonal
                for (int z = tj + 1; z < 16; z++) cm16[ci][z] = 0;
        }
    }
    __syncthreads();  // column masks complete
    // (3) warp 0 resolves the block 32 candidates at a time.  A chunk's candidates are first tested against
    // everything kept before the chunk (one AND per mask word); the dependencies INSIDE the chunk are settled with
    // ballots: an undecided lane is removed once a kept earlier lane suppresses it and kept once no undecided or
    // kept earlier lane does.  The lowest undecided lane always decides, typical chains settle in 2-4 rounds, and a
    // round is two ballots -- no barrier, no shared-memory traffic.
    if (tid < 32) {
        const int lane = tid;
        unsigned long long K0 = 0ull, K1 = 0ull, K2 = 0ull, K3 = 0ull;
        for (int ch = 0; ch * 32 < nb; ch++) {
            const int ci = ch * 32 + lane, w = ch >> 1, sh = (ch & 1) * 32;
            const unsigned long long c0 = s_colmask[ci][0], c1 = s_colmask[ci][1], c2 = s_colmask[ci][2],
                                     c3 = s_colmask[ci][3];
            const bool alive = ci < nb && !((s_sup[w] >> (ci & 63)) & 1ull);
            const bool prior = ((c0 & K0) | (c1 & K1) | (c2 & K2) | (c3 & K3)) != 0ull;
            const unsigned long long cw = w == 0 ? c0 : (w == 1 ? c1 : (w == 2 ? c2 : c3));
            const unsigned m = (unsigned)(cw >> sh);  // earlier lanes of this chunk that would suppress this lane
            unsigned U = __ballot_sync(0xffffffffu, alive && !prior), Kc = 0u;
            while (U) {
                const bool mine = (U >> lane) & 1u;
                const bool rem = mine && (m & Kc) != 0u;
                const bool kp = mine && !rem && (m & U) == 0u;
                const unsigned kb = __ballot_sync(0xffffffffu, kp), rb = __ballot_sync(0xffffffffu, rem);
                Kc |= kb;
                U &= ~(kb | rb);
            }
            const unsigned long long add = (unsigned long long)Kc << sh;
            if (w == 0) K0 |= add;
            else if (w == 1) K1 |= add;
            else if (w == 2) K2 |= add;
            else K3 |= add;
        }
        if (lane == 0) {
            s_K[0] = K0;
            s_K[1] = K1;
            s_K[2] = K2;
            s_K[3] = K3;
        }
    }
    __syncthreads();
}

// Fused single-CTA NMS for segments of at most SMALL_LIMIT (1024) boxes, everything in shared memory:
//   sort       : each warp sorts 32 keys in registers (keys are unique: score desc, tie asc); the rank of a key is
//                found by binary search in the other runs; the box is scattered to that rank;
//   suppression: the same blocked greedy sweep as nms_greedy_kernel (256 candidates per step, kept boxes
//                compacted in place at the front of the box array).
// No global scratch, one launch, a handful of barriers per 256 boxes.
// NC: the rows are kernel inputs (read-only path); !NC: they were written by this CTA before a barrier (the fused
// merge below), so they are read with ld.global.cg.  n = boxes of segment g (1 .. SMALL_LIMIT); returns the number
// of kept boxes (CTA-uniform; the keep list is in global memory behind a barrier when the function returns).
template <bool NC>
__device__ __forceinline__ float ld_row(const float *p) {
    return NC ? __ldg(p) : __ldcg(p);
}

template <int MODE, bool NC>
__device__ __forceinline__ int nms_small_body(const float *rows, int row_stride, int box_off, int score_off,
                                              const int32_t *tie, int n, int g, int cap, int max_nms, float thr,
                                              int inclusive, int32_t *keep) {
    __shared__ unsigned long long s_keys[SMALL_LIMIT];
    __shared__ float4 s_box[SMALL_LIMIT];
    __shared__ uint32_t s_idx[SMALL_LIMIT];
    __shared__ int s_rank[SMALL_LIMIT];
    __shared__ uint16_t s_item[SMALL_LIMIT];
    __shared__ unsigned long long s_rowmask[GREEDY_BS][4];
    __shared__ unsigned long long s_sup[4], s_keptbits[4];
    __shared__ int s_nkeep;
    const int tid = threadIdx.x;
    const bool quick = thr > 0.f || (thr == 0.f && !inclusive);
    const float *seg_rows = rows + (long long)g * cap * row_stride;
    // ---- sort: 32-key runs in registers, then rank by binary search ---------------------------------------------
    // keys are unique (tie = row index or item index; callers must not pass duplicate ties inside a segment).  Every
    // warp sorts the keys of its 32 items with a shuffle bitonic network (no shared memory, no barrier); the rank of
    // a key is its position in its own run plus, for every other run, the number of smaller keys there (a 6-probe
    // binary search).  n^2/32 warp-level compares become n * (n/32) * 6 probes; all 1024 threads take part: the
    // 1024 / pad threads of an element split the runs and add their partial counts.
    const int warp = tid >> 5, lane = tid & 31;
    const int nr = (n + 31) >> 5;  // runs
    int pad = 32;
    while (pad < n) pad <<= 1;
    const int parts = SMALL_THREADS / pad;
    if (warp < nr) {
        unsigned long long k = ~0ull;  // padding sorts last and is smaller than nothing
        uint32_t id = (uint32_t)tid;
        if (tid < n) {
            const float *r = seg_rows + (long long)tid * row_stride;
            const uint32_t t = tie ? (uint32_t)__ldg(tie + (long long)g * cap + tid) : (uint32_t)tid;
            k = ((unsigned long long)score_key_desc(ld_row<NC>(r + score_off)) << 32) | t;
        }
#pragma unroll
        for (int sz = 2; sz <= 32; sz <<= 1) {
#pragma unroll
            for (int st = sz >> 1; st > 0; st >>= 1) {
                const unsigned long long ok = __shfl_xor_sync(0xffffffffu, k, st);
                const uint32_t oid = __shfl_xor_sync(0xffffffffu, id, st);
                const bool take_min = ((lane & sz) == 0) == ((lane & st) == 0);
                if (take_min ? (ok < k) : (ok > k)) {
                    k = ok;
                    id = oid;
                }
            }
        }
        s_keys[tid] = k;
        s_item[tid] = (uint16_t)id;
    }
    if (tid < pad) s_rank[tid] = 0;
    if (tid == 0) s_nkeep = 0;
    __syncthreads();
    {
        const int e = tid & (pad - 1), part = tid / pad;
        if (e < nr * 32) {
            const unsigned long long mykey = s_keys[e];
            if (mykey != ~0ull) {
                const int run = e >> 5;
                int cnt = part == 0 ? (e & 31) : 0;
#pragma unroll 2
                for (int q = part; q < nr; q += parts) {
                    if (q == run) continue;
                    const unsigned long long *A = s_keys + q * 32;
                    int c = 0;
#pragma unroll
                    for (int st = 16; st > 0; st >>= 1) c += (A[c + st - 1] < mykey) ? st : 0;
                    c += (A[c] < mykey) ? 1 : 0;
                    cnt += c;
                }
                if (parts == 1) s_rank[e] = cnt;
                else if (cnt) atomicAdd(&s_rank[e], cnt);
            }
        }
    }
    __syncthreads();
    if (tid < nr * 32 && s_keys[tid] != ~0ull) {
        const int item = s_item[tid], rank = s_rank[tid];
        const float *r = seg_rows + (long long)item * row_stride + box_off;
        s_box[rank] = make_float4(ld_row<NC>(r), ld_row<NC>(r + 1), ld_row<NC>(r + 2), ld_row<NC>(r + 3));
        s_idx[rank] = (uint32_t)item;
    }
    __syncthreads();
    const int K = (max_nms > 0) ? min(n, max_nms) : n;  // utils/general.py:578-579
    // ---- blocked greedy suppression ---------------------------------------------------------------------------
    int32_t *kout = keep + (long long)g * cap;
    const int i = tid & (GREEDY_BS - 1), grp = tid >> 8;
    for (int base = 0; base < K; base += GREEDY_BS) {
        const int nb = min(GREEDY_BS, K - base), nk = s_nkeep;
        if (tid < 4) s_sup[tid] = 0ull;
        __syncthreads();
        const float4 bi = i < nb ? s_box[base + i] : make_float4(0.f, 0.f, 0.f, 0.f);
        // (1) against the boxes kept in earlier steps (compacted at the front of s_box)
        bool sup = false;
        if (i < nb) {
            for (int qb = grp * 32; qb < nk && !sup; qb += 128) {  // the 4 thread groups take 32 kept boxes each
                unsigned ov = 0u;
#pragma unroll 8
                for (int j = 0; j < 32; j++) ov |= boxes_overlap<MODE>(s_box[qb + j], bi) ? (1u << j) : 0u;  // < SMALL_LIMIT
                if (!quick) ov = ~0u;
                ov &= (nk - qb >= 32) ? ~0u : ((1u << (nk - qb)) - 1u);
                while (ov && !sup) {
                    const int j = __ffs(ov) - 1;
                    ov &= ov - 1u;
                    sup = iou_gt<MODE>(s_box[qb + j], bi, thr, inclusive != 0, false);
                }
            }
        }
        {   // one native 32-bit atomic per warp (a warp = 32 consecutive candidates); 64-bit shared-memory OR is a
            // CAS loop and serialises badly when most of a block is suppressed
            const unsigned sb = __ballot_sync(0xffffffffu, sup);
            if ((tid & 31) == 0 && sb) atomicOr(reinterpret_cast<unsigned *>(s_sup) + (i >> 5), sb);
        }
        // (2) + (3): column masks and parallel greedy resolution
        greedy_block<MODE>(s_box + base, nb, thr, inclusive != 0, quick, s_rowmask, s_sup, s_keptbits);
        // (4) append: kept boxes move to the front of s_box (never past the read position), indices go out
        float4 kbox = bi;
        uint32_t kidx = 0;
        int pos = -1;
        if (tid < nb) {
            const int w = tid >> 6, b = tid & 63;
            if ((s_keptbits[w] >> b) & 1ull) {
                pos = nk + __popcll(s_keptbits[w] & ((1ull << b) - 1ull));
                for (int v = 0; v < w; v++) pos += __popcll(s_keptbits[v]);
                kidx = s_idx[base + tid];
            }
        }
        __syncthreads();  // every reader of s_box[base ..] is done before the in-place compaction
        if (pos >= 0) {
            s_box[pos] = kbox;
            kout[pos] = (int32_t)kidx;
        }
        if (tid == 0)
            s_nkeep = nk + __popcll(s_keptbits[0]) + __popcll(s_keptbits[1]) + __popcll(s_keptbits[2]) +
                      __popcll(s_keptbits[3]);
        __syncthreads();
    }
    return s_nkeep;
}

template <int MODE>
__global__ void FDMS_SMALL_BOUNDS
    nms_small_kernel(const float *__restrict__ rows, int row_stride, int box_off, int score_off,
                     const int32_t *__restrict__ tie, const int32_t *__restrict__ count, int cap, int max_nms,
                     int small_limit, float thr, int inclusive, int32_t *__restrict__ keep,
                     int32_t *__restrict__ n_keep) {
    const int g = blockIdx.x;
    fdms_pdl_trigger();
    fdms_pdl_wait();  // the candidates come from the kernel before this one
    const int n = max(0, min(count[g], cap));
    if (n > small_limit) {  // the sort + greedy / bit-matrix path takes it (it overwrites the marker)
        if (threadIdx.x == 0) n_keep[g] = -2;
        return;
    }
    int nk = 0;
    if (n > 0) nk = nms_small_body<MODE, true>(rows, row_stride, box_off, score_off, tie, n, g, cap, max_nms, thr, inclusive, keep);
    if (threadIdx.x == 0) n_keep[g] = nk;
}

// The cross-scale merge of an image in ONE CTA when it has at most SMALL_LIMIT rows (the common case): merge rows
// (rescale, round, weights; merge_rows.cuh) -> NMS with the TV_CPU arithmetic (msfd:234 runs on CPU tensors) ->
// detections[keep].  Replaces merge_prepare_kernel + nms_small_kernel + merge_emit_kernel: two launches and two
// kernel tails less on a chain that is pure latency.  Images with more rows only get their rows here.
__global__ void FDMS_SMALL_BOUNDS
    merge_small_kernel(const __grid_constant__ MergeParams P, const __grid_constant__ EmitParams E, int small_limit,
                       float thr, int inclusive) {
    const int img = blockIdx.x, tid = threadIdx.x;
    fdms_pdl_wait();  // the keep lists come from the kernel before this one
    const int total = merge_prepare_image(P, img, tid, SMALL_THREADS);
    if (total > small_limit) {  // nms_sort_kernel + nms_greedy / mask + merge_emit_kernel take this image
        if (tid == 0) E.n_out[img] = 0;
        return;
    }
    __syncthreads();  // the rows above were written by other threads of this CTA
    int nk = 0;
    if (total > 0)
        nk = nms_small_body<FDMS_IOU_TV_CPU, false>(P.mrows, MROW, 0, 4, nullptr, total, img, P.mcap, 0, thr, inclusive,
                                                    const_cast<int32_t *>(E.keep2));
    if (tid == 0) const_cast<int32_t *>(E.n_keep2)[img] = nk;
    merge_emit_image(E, img, nk, tid, SMALL_THREADS);  // the keep list is behind the body's last barrier
}

// Blocked greedy NMS for larger segments, one CTA per segment, after nms_sort_kernel.  Boxes arrive in score
// order in global scratch; the boxes kept so far are compacted IN PLACE at the front of the same array (the
// write position never passes the read position) and streamed through shared memory in tiles.  Per step 256
// candidates are (1) tested against the kept list with early exit, (2) masked against each other (256 x 256
// bits), (3) resolved by warp 0 in 32-candidate chunks with ballots, (4) appended.  Work ~ K * n_kept / 2 pair tests
// instead of K^2 / 2, and no K^2 bit-matrix.
template <int MODE>
__global__ void __launch_bounds__(GREEDY_THREADS)
    nms_greedy_kernel(const SegInfo *__restrict__ seginfo, int cap_e, float4 *sorted_boxes,
                      const uint32_t *__restrict__ sorted_idx, float thr, int inclusive, int32_t *__restrict__ keep,
                      int32_t *__restrict__ n_keep, int cap, const int32_t *__restrict__ count, int skip_le) {
    __shared__ float4 s_blk[GREEDY_BS];
    __shared__ float4 s_tile[GREEDY_TILE];
    __shared__ unsigned long long s_rowmask[GREEDY_BS][4];
    __shared__ unsigned long long s_sup[4], s_keptbits[4];
    __shared__ int s_nkeep;
    const int g = blockIdx.x, tid = threadIdx.x;
    if (max(0, min(count[g], cap)) <= skip_le) return;  // handled by nms_small_kernel
    const bool quick = thr > 0.f || (thr == 0.f && !inclusive);
    const int K = seginfo[g].n;
    float4 *boxes = sorted_boxes + (long long)g * cap_e;
    const uint32_t *sidx = sorted_idx + (long long)g * cap_e;
    int32_t *kout = keep + (long long)g * cap;
    if (tid == 0) s_nkeep = 0;
    __syncthreads();
    const int i = tid & (GREEDY_BS - 1), grp = tid >> 8;  // 4 groups of 256 threads
    for (int base = 0; base < K; base += GREEDY_BS) {
        const int nb = min(GREEDY_BS, K - base), nk = s_nkeep;
        if (tid < GREEDY_BS) s_blk[tid] = tid < nb ? boxes[base + tid] : make_float4(0.f, 0.f, 0.f, 0.f);
        if (tid < 4) s_sup[tid] = 0ull;
        __syncthreads();
        const float4 bi = s_blk[i];
        // (1) against the kept list, tile by tile: one candidate per thread, 32 branch-free overlap tests in flight.
        // (Tried and dropped, measured on the dense-crowd workload: a WARP per candidate, lane = kept box, leaving the list
        // at the first chunk in which any lane suppresses the candidate -- fewer pair tests, but a dependent LDS / vote
        // chain per chunk and the exact IoU taken in nearly every chunk of a crowd: NMS 0.99 -> 1.29 ms, merge 0.76 -> 1.08.)
        bool sup = false;
        for (int t0 = 0; t0 < nk; t0 += GREEDY_TILE) {
            const int tn = min(GREEDY_TILE, nk - t0);
            if (t0) __syncthreads();  // readers of the previous tile are done
            if (tid < tn) s_tile[tid] = boxes[t0 + tid];
            __syncthreads();
            if (i < nb && !sup) {
                for (int qb = grp * 32; qb < tn && !sup; qb += 128) {
                    unsigned ov = 0u;
#pragma unroll 8
                    for (int j = 0; j < 32; j++) ov |= boxes_overlap<MODE>(s_tile[(qb + j) & (GREEDY_TILE - 1)], bi) ? (1u << j) : 0u;
                    if (!quick) ov = ~0u;
                    ov &= (tn - qb >= 32) ? ~0u : ((1u << (tn - qb)) - 1u);
                    while (ov && !sup) {
                        const int j = __ffs(ov) - 1;
                        ov &= ov - 1u;
                        sup = iou_gt<MODE>(s_tile[qb + j], bi, thr, inclusive != 0, false);
                    }
                }
            }
        }
        {
            const unsigned sb = __ballot_sync(0xffffffffu, sup);
            if ((tid & 31) == 0 && sb) atomicOr(reinterpret_cast<unsigned *>(s_sup) + (i >> 5), sb);
        }
        // (2) + (3): column masks and parallel greedy resolution
        greedy_block<MODE>(s_blk, nb, thr, inclusive != 0, quick, s_rowmask, s_sup, s_keptbits);
        // (4) append the kept boxes (in place) and their item indices
        if (tid < nb) {
            const int w = tid >> 6, b = tid & 63;
            if ((s_keptbits[w] >> b) & 1ull) {
                int pos = nk + __popcll(s_keptbits[w] & ((1ull << b) - 1ull));
                for (int v = 0; v < w; v++) pos += __popcll(s_keptbits[v]);
                boxes[pos] = s_blk[tid];
                kout[pos] = (int32_t)sidx[base + tid];
            }
        }
        if (tid == 0)
            s_nkeep = nk + __popcll(s_keptbits[0]) + __popcll(s_keptbits[1]) + __popcll(s_keptbits[2]) +
                      __popcll(s_keptbits[3]);
        __syncthreads();
    }
    if (tid == 0) n_keep[g] = s_nkeep;
}

// Large segments, one CTA per segment, after nms_sort_kernel: blocked greedy NMS in which a KEPT box marks the later
// candidates it suppresses, found through a size-stratified grid of the segment's candidates in shared memory.  A dense
// crowd needs ~100 pair tests per kept box instead of ~1000 per candidate against the whole kept list (ncu of the list
// kernel on the dense-crowd workload: 81 % ALU-pipe utilisation, bound by the pair tests), and the tests of a step are
// spread evenly over the CTA.
//   grid      : built once per segment (counting sort, CSR).  Level l holds the candidates whose longer side m lies in
//               (c0 2^(l-1), c0 2^l] (level 0: m <= c0) in a uniform grid of cell size c0 2^l over the segment's
//               extent, c0 = extent / (128 x 64), 8 levels; a candidate is filed under the cell of its CENTRE.  Cells
//               of a grid row are consecutive in the CSR order, so a row span is one contiguous run of entries.
//   step      : the next 256 candidates that no kept box has marked (scan of the mark bits, compaction) go through the
//               intra-block masks + ballot resolver (greedy_block); the kept ones are appended to the keep list, then
//               every kept box is taken by a warp that walks the candidates it can suppress and marks them.
//   which     : IoU <= min(w_small / w_large, h_small / h_large) <= m_small / m_large, so a kept box with longer side
//               m_k can only suppress candidates with m in [t m_k, m_k / t] (t = 0.99 x threshold: the bound holds up
//               to a few ulp in the float arithmetic of iou_gt, the margin is 1 %) -- the levels of that range; on
//               level l (candidate extents <= c_l) the candidate's centre must lie within bin_radius() of the kept
//               box's centre along each axis (bin_window: for a threshold of 0.5 about the kept box itself, 1-3 cells
//               a side).  One lane per (level, row) fetches the row's run, the runs are flattened over the 32 lanes
//               (prefix sum + shuffle search).  The cell index is a monotone function of the coordinate, so no
//               suppressing pair is missed.  Exact: the pair test itself is unchanged.
//   no grid   : thresholds for which disjoint or tiny boxes count (<= 0), an extent that is not finite or is tiny
//               against the coordinates' magnitude: every kept box walks ALL later candidates (grid_on = 0).
//   (Tried first, same workload: the KEPT boxes in a single-level grid of linked lists, candidates walking them --
//   merge 2.6 ms instead of 0.7, the CTA waiting at barriers for the few lanes whose large candidates covered 30+
//   cells; then kept boxes in the stratified grid -- NMS 0.99 -> 0.45 ms, merge 0.71 -> 0.49, still 53-74 % barrier
//   stalls at 8.7 active threads per instruction: one lane per candidate walking dependent lists cannot be balanced.)
constexpr int BIN_LEVELS = 8;
constexpr int BIN_GX0 = 128, BIN_GY0 = 64;
constexpr int BIN_CELLS = 10923;  // sum over levels of ceil(128 / 2^l) * ceil(64 / 2^l)
constexpr int BIN_MAX_K = 65535;  // entries are 16-bit positions in score order
constexpr uint32_t BIN_PER_THREAD = (BIN_CELLS + GREEDY_THREADS - 1) / GREEDY_THREADS;

struct BinShared {
    float4 blk[GREEDY_BS];
    unsigned long long rowmask[GREEDY_BS][4];
    unsigned long long sup[4], keptbits[4];
    uint32_t cell_end[BIN_CELLS + 1];  // counts, then exclusive starts, then (after the scatter) the END of each cell
    uint32_t marks[(BIN_MAX_K + 32) / 32];
    uint16_t cpos[GREEDY_BS];  // block member -> position in score order
    uint16_t kl[GREEDY_BS];    // kept members of this step
    int warp_cnt[32];
    float red[5][32];
    int nkeep, consumed, grid_on, next_kept;
    float x0, y0, inv_c0;
    float lvl_inv[BIN_LEVELS], lvl_c[BIN_LEVELS];
    int lvl_gx[BIN_LEVELS], lvl_gy[BIN_LEVELS], lvl_off[BIN_LEVELS];
    uint16_t entries[1];  // K entries (dynamic)
};

// smallest l with m <= c0 2^l, clamped to the levels there are; monotone in q = m / c0 (q > 0)
__device__ __forceinline__ int bin_level(float q) {
    const int l = ((__float_as_int(q) - 1) >> 23) - 126;
    return min(max(l, 0), BIN_LEVELS - 1);
}

// longer side of a box in the kernel's own convention; 0 when the box has no extent in some dimension (or NaN): such a
// box overlaps nothing (boxes_overlap is false for every partner), it is neither filed nor marks anything
template <int MODE>
__device__ __forceinline__ float box_long_side(const float4 b) {
    const float e = MODE == FDMS_IOU_NCNN ? 1.f : 0.f;  // the ncnn flavour's extents carry +1 pixel
    const float w = f_add(f_sub(b.z, b.x), e), h = f_add(f_sub(b.w, b.y), e);
    return (w > 0.f && h > 0.f) ? fmaxf(w, h) : 0.f;
}

__device__ __forceinline__ int bin_cell(float v, float origin, float inv_c, int n) {
    return __float2int_rd(fminf(fmaxf(f_mul(f_sub(v, origin), inv_c), 0.f), (float)(n - 1)));
}

// Cells of level l that can hold the centre of a candidate the kept box (centre xk yk, extent wk hk) suppresses.
// Per axis, with w the kept extent and wc <= c_l the candidate's: IoU <= (overlap along the axis) / max(w, wc), and the
// overlap is at most (w + wc) / 2 - |distance of the centres|, so IoU > t needs
//     |distance| < f(wc) = (w + wc) / 2 - t max(w, wc),
// which over wc in (0, c_l] is largest at wc = c_l or at wc = min(w, c_l) (f rises with slope 1/2 up to w and with
// 1/2 - t after it).  + 1 % of the cell for the float arithmetic.  An empty window has cx0 > cx1 or cy0 > cy1.
struct BinWindow {
    int cx0, cx1, cy0, cy1, gx, off;
};

__device__ __forceinline__ float bin_radius(float w, float cl, float t) {
    const float wa = fminf(w, cl);
    const float f1 = f_sub(f_mul(0.5f, f_add(w, cl)), f_mul(t, fmaxf(w, cl)));
    const float f2 = f_sub(f_mul(0.5f, f_add(w, wa)), f_mul(t, w));  // max(w, wa) = w
    return f_add(fmaxf(f1, f2), f_mul(0.01f, cl));
}

__device__ __forceinline__ BinWindow bin_window(const BinShared &S, int l, float xk, float yk, float wk, float hk, float t,
                                                float x0, float y0) {
    BinWindow r;
    const float cl = S.lvl_c[l], inv = S.lvl_inv[l];
    const int gy = S.lvl_gy[l];
    r.gx = S.lvl_gx[l];
    r.off = S.lvl_off[l];
    const float rx = bin_radius(wk, cl, t), ry = bin_radius(hk, cl, t);
    r.cx0 = bin_cell(f_sub(xk, rx), x0, inv, r.gx);
    r.cx1 = rx >= 0.f ? bin_cell(f_add(xk, rx), x0, inv, r.gx) : -1;
    r.cy0 = bin_cell(f_sub(yk, ry), y0, inv, gy);
    r.cy1 = ry >= 0.f ? bin_cell(f_add(yk, ry), y0, inv, gy) : -1;
    return r;
}

// Every lane brings one run [b, e) of CSR entries (b == e: none); the items of all 32 runs are dealt to the lanes in
// turn (prefix sum + shuffle search) and each is tested against the kept box: later, unmarked candidates it suppresses
// get their mark.  Called by whole warps.
template <int MODE>
__device__ __forceinline__ void bin_mark_runs(BinShared &S, const float4 *__restrict__ boxes, const float4 kb, int kpos,
                                              int base, uint32_t b, uint32_t e, float thr, bool inclusive, int lane) {
    uint32_t incl = e - b;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    const int shift = (int)(b - (incl - (e - b)));  // entry = item + shift inside this lane's run
    for (uint32_t it = lane; it < ((total + 31u) & ~31u); it += 32) {
        int lo = 0;  // the run holding item `it`: the first lane whose inclusive count exceeds it
#pragma unroll
        for (int st = 16; st > 0; st >>= 1) {
            const uint32_t v = __shfl_sync(0xffffffffu, incl, lo + st - 1);
            if (v <= it) lo += st;
        }
        const int sh = __shfl_sync(0xffffffffu, shift, lo);
        if (it < total) {
            const int k = S.entries[(int)it + sh];
            if (k > kpos && k >= base && !((S.marks[k >> 5] >> (k & 31)) & 1u)) {
                const float4 cb = boxes[k];
                if (boxes_overlap<MODE>(kb, cb) && iou_gt<MODE>(kb, cb, thr, inclusive, false))
                    atomicOr(&S.marks[k >> 5], 1u << (k & 31));
            }
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(GREEDY_THREADS)
    nms_binned_kernel(const SegInfo *__restrict__ seginfo, int cap_e, const float4 *__restrict__ sorted_boxes,
                      const uint32_t *__restrict__ sorted_idx, float thr, int inclusive, int32_t *__restrict__ keep,
                      int32_t *__restrict__ n_keep, int cap, const int32_t *__restrict__ count, int skip_le,
                      int bin_on) {
    extern __shared__ __align__(16) uint8_t bin_sm[];
    BinShared &S = *reinterpret_cast<BinShared *>(bin_sm);
    const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (max(0, min(count[g], cap)) <= skip_le) return;  // handled by nms_small_kernel
    const bool quick = thr > 0.f || (thr == 0.f && !inclusive);
    const int K = seginfo[g].n;
    const float4 *boxes = sorted_boxes + (long long)g * cap_e;
    const uint32_t *sidx = sorted_idx + (long long)g * cap_e;
    int32_t *kout = keep + (long long)g * cap;
    const float ext = MODE == FDMS_IOU_NCNN ? 1.f : 0.f;
    // ---- the segment's extent -> level grids --------------------------------------------------------------------
    {
        float mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY, amax = 0.f;
        for (int k = tid; k < K; k += GREEDY_THREADS) {
            const float4 b = boxes[k];
            mnx = fminf(mnx, b.x);
            mny = fminf(mny, b.y);
            mxx = fmaxf(mxx, b.z + ext);
            mxy = fmaxf(mxy, b.w + ext);
            amax = fmaxf(amax, fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w))));
            if (!(fabsf(b.x) < 1e30f && fabsf(b.y) < 1e30f && fabsf(b.z) < 1e30f && fabsf(b.w) < 1e30f)) amax = INFINITY;  // NaN too
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, d));
            mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, d));
            mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, d));
            mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, d));
            amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, d));
        }
        if (lane == 0) {
            S.red[0][warp] = mnx;
            S.red[1][warp] = mny;
            S.red[2][warp] = mxx;
            S.red[3][warp] = mxy;
            S.red[4][warp] = amax;
        }
        for (int c = tid; c <= BIN_CELLS; c += GREEDY_THREADS) S.cell_end[c] = 0u;
        for (int w = tid; w < (K + 31) / 32; w += GREEDY_THREADS) S.marks[w] = 0u;
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < GREEDY_THREADS / 32; w++) {
                mnx = fminf(mnx, S.red[0][w]);
                mny = fminf(mny, S.red[1][w]);
                mxx = fmaxf(mxx, S.red[2][w]);
                mxy = fmaxf(mxy, S.red[3][w]);
                amax = fmaxf(amax, S.red[4][w]);
            }
            const float ex = mxx - mnx, ey = mxy - mny;
            // a little more than extent / cells: the last cell index stays below the grid size without the clamp
            const float c0 = fmaxf(ex / (float)(BIN_GX0 - 1), ey / (float)(BIN_GY0 - 1));
            // the 2 % slack of the search window must dwarf the rounding of a coordinate (2^-23 relative)
            const bool ok = bin_on && quick && thr < 1e30f && amax < 1e30f && c0 > 0.f && c0 > amax * 1e-4f;
            S.grid_on = ok ? 1 : 0;
            S.x0 = mnx;
            S.y0 = mny;
            S.inv_c0 = ok ? 1.f / c0 : 0.f;
            int off = 0;
            for (int l = 0; l < BIN_LEVELS; l++) {
                const float cl = c0 * (float)(1 << l);
                S.lvl_c[l] = cl;
                S.lvl_inv[l] = ok ? 1.f / cl : 0.f;
                S.lvl_gx[l] = ok ? min((BIN_GX0 + (1 << l) - 1) >> l, (int)(ex / cl) + 1) : 1;
                S.lvl_gy[l] = ok ? min((BIN_GY0 + (1 << l) - 1) >> l, (int)(ey / cl) + 1) : 1;
                S.lvl_off[l] = off;
                off += ((BIN_GX0 + (1 << l) - 1) >> l) * ((BIN_GY0 + (1 << l) - 1) >> l);
            }
            S.nkeep = 0;
        }
        __syncthreads();
    }
    const bool grid_on = S.grid_on != 0;
    const float x0 = S.x0, y0 = S.y0, inv_c0 = S.inv_c0;
    // ---- counting sort of the candidates by cell (CSR) ------------------------------------------------------------
    if (grid_on) {
        for (int pass = 0; pass < 2; pass++) {
            for (int k = tid; k < K; k += GREEDY_THREADS) {
                const float4 b = boxes[k];
                const float m = box_long_side<MODE>(b);
                if (m > 0.f) {
                    const int l = bin_level(f_mul(m, inv_c0)), gx = S.lvl_gx[l];
                    const int cx = bin_cell(f_mul(f_add(b.x, f_add(b.z, ext)), 0.5f), x0, S.lvl_inv[l], gx);
                    const int cy = bin_cell(f_mul(f_add(b.y, f_add(b.w, ext)), 0.5f), y0, S.lvl_inv[l], S.lvl_gy[l]);
                    const int cell = S.lvl_off[l] + cy * gx + cx;
                    const uint32_t at = atomicAdd(&S.cell_end[cell], 1u);
                    if (pass) S.entries[at] = (uint16_t)k;
                }
            }
            __syncthreads();
            if (pass == 0) {  // counts -> exclusive starts
                uint32_t v[BIN_PER_THREAD], sum = 0u;
                const int c_first = tid * (int)BIN_PER_THREAD;
#pragma unroll
                for (int q = 0; q < (int)BIN_PER_THREAD; q++) {
                    v[q] = c_first + q < BIN_CELLS ? S.cell_end[c_first + q] : 0u;
                    sum += v[q];
                }
                uint32_t incl = sum;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += t;
                }
                if (lane == 31) S.warp_cnt[warp] = (int)incl;
                __syncthreads();
                uint32_t wsum = (uint32_t)S.warp_cnt[lane];
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, wsum, d);
                    if (lane >= d) wsum += t;
                }
                const uint32_t wbefore = __shfl_sync(0xffffffffu, wsum, max(warp - 1, 0));
                uint32_t run = (warp ? wbefore : 0u) + incl - sum;
#pragma unroll
                for (int q = 0; q < (int)BIN_PER_THREAD; q++) {
                    if (c_first + q < BIN_CELLS) S.cell_end[c_first + q] = run;
                    run += v[q];
                }
                __syncthreads();
            }
        }
    }
    const float t_lo = fminf(0.99f * thr, 1.f), t_hi = 1.f / fmaxf(t_lo, 1e-30f);
    for (int base = 0; base < K;) {
        const int nk = S.nkeep;
        if (tid < 4) S.sup[tid] = 0ull;
        if (tid == 4) S.next_kept = 0;
        // ---- the next GREEDY_BS unmarked candidates, in score order ----------------------------------------------
        int nb = 0;
        while (nb < GREEDY_BS && base < K) {
            const int nw = min(GREEDY_THREADS, K - base), k = base + tid;
            const bool alive = tid < nw && !((S.marks[k >> 5] >> (k & 31)) & 1u);
            const unsigned bal = __ballot_sync(0xffffffffu, alive);
            if (lane == 0) S.warp_cnt[warp] = __popc(bal);
            __syncthreads();
            int incl = S.warp_cnt[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            const int before = __shfl_sync(0xffffffffu, incl, max(warp - 1, 0));
            const int rank = nb + (warp ? before : 0) + __popc(bal & ((1u << lane) - 1u));
            if (alive && rank < GREEDY_BS) {
                S.blk[rank] = boxes[k];
                S.cpos[rank] = (uint16_t)k;
                if (rank == GREEDY_BS - 1) S.consumed = tid + 1;
            }
            __syncthreads();  // warp_cnt is reused by the next window; consumed is visible
            if (nb + total >= GREEDY_BS) {
                base += S.consumed;
                nb = GREEDY_BS;
            } else {
                base += nw;
                nb += total;
            }
        }
        // (2) + (3): column masks and parallel greedy resolution over the block (begins with a barrier)
        greedy_block<MODE>(S.blk, nb, thr, inclusive != 0, quick, S.rowmask, S.sup, S.keptbits);
        // (4) append the kept members
        const int k0 = __popcll(S.keptbits[0]), k1 = __popcll(S.keptbits[1]), k2 = __popcll(S.keptbits[2]),
                  k3 = __popcll(S.keptbits[3]);
        const int n_new = k0 + k1 + k2 + k3;
        if (tid < nb) {
            const int w = tid >> 6, b = tid & 63;
            if ((S.keptbits[w] >> b) & 1ull) {
                const int j = __popcll(S.keptbits[w] & ((1ull << b) - 1ull)) + (w > 0 ? k0 : 0) + (w > 1 ? k1 : 0) + (w > 2 ? k2 : 0);
                kout[nk + j] = (int32_t)sidx[S.cpos[tid]];
                S.kl[j] = (uint16_t)tid;
            }
        }
        __syncthreads();
        // (5) every kept box marks the later candidates it suppresses: one warp per kept box
        if (base < K) {
            // kept boxes are handed to the warps one at a time (their neighbourhoods differ in size by an order of magnitude)
            for (;;) {
                int j = 0;
                if (lane == 0) j = atomicAdd(&S.next_kept, 1);
                j = __shfl_sync(0xffffffffu, j, 0);
                if (j >= n_new) break;
                const int member = S.kl[j];
                const float4 kb = S.blk[member];
                const int kpos = S.cpos[member];
                if (!grid_on) {
                    for (int k = max(base, kpos + 1) + lane; k < K; k += 32) {
                        if ((S.marks[k >> 5] >> (k & 31)) & 1u) continue;
                        const float4 cb = boxes[k];
                        if ((!quick || boxes_overlap<MODE>(kb, cb)) && iou_gt<MODE>(kb, cb, thr, inclusive != 0, false))
                            atomicOr(&S.marks[k >> 5], 1u << (k & 31));
                    }
                    continue;
                }
                const float m = box_long_side<MODE>(kb);
                if (!(m > 0.f)) continue;
                const float q = f_mul(m, inv_c0);
                const int l_lo = bin_level(f_mul(q, t_lo)), l_hi = bin_level(f_mul(q, t_hi));
                const float wk = f_sub(f_add(kb.z, ext), kb.x), hk = f_sub(f_add(kb.w, ext), kb.y);
                const float xk = f_mul(f_add(kb.x, f_add(kb.z, ext)), 0.5f), yk = f_mul(f_add(kb.y, f_add(kb.w, ext)), 0.5f);
                // the runs of one (level, row) per lane: 8 levels x 4 rows in one go; a level that needs more rows (tiny
                // thresholds) sends the box through the row-chunk loop below instead
                BinWindow wnd;
                const int l_mine = l_lo + (lane >> 2);
                wnd = bin_window(S, min(l_mine, BIN_LEVELS - 1), xk, yk, wk, hk, t_lo, x0, y0);
                const bool l_valid = l_mine <= l_hi && wnd.cx0 <= wnd.cx1 && wnd.cy0 <= wnd.cy1;
                if (!__any_sync(0xffffffffu, l_valid && wnd.cy1 - wnd.cy0 >= 4)) {
                    const int cy = wnd.cy0 + (lane & 3);
                    uint32_t b = 0u, e = 0u;
                    if (l_valid && cy <= wnd.cy1) {
                        const int ca = wnd.off + cy * wnd.gx + wnd.cx0, cz = wnd.off + cy * wnd.gx + wnd.cx1;
                        b = ca ? S.cell_end[ca - 1] : 0u;
                        e = S.cell_end[cz];
                    }
                    bin_mark_runs<MODE>(S, boxes, kb, kpos, base, b, e, thr, inclusive != 0, lane);
                    continue;
                }
                for (int l = l_lo; l <= l_hi; l++) {
                    wnd = bin_window(S, l, xk, yk, wk, hk, t_lo, x0, y0);
                    if (wnd.cx0 > wnd.cx1) continue;
                    for (int r0 = wnd.cy0; r0 <= wnd.cy1; r0 += 32) {
                        const int cy = r0 + lane;
                        uint32_t b = 0u, e = 0u;
                        if (cy <= wnd.cy1) {
                            const int ca = wnd.off + cy * wnd.gx + wnd.cx0, cz = wnd.off + cy * wnd.gx + wnd.cx1;
                            b = ca ? S.cell_end[ca - 1] : 0u;
                            e = S.cell_end[cz];
                        }
                        bin_mark_runs<MODE>(S, boxes, kb, kpos, base, b, e, thr, inclusive != 0, lane);
                    }
                }
            }
        }
        if (tid == 0) S.nkeep = nk + n_new;
        __syncthreads();
    }
    if (tid == 0) n_keep[g] = S.nkeep;
}

__global__ void gather_rows_kernel(const float *__restrict__ rows, int row_stride, int n_cols,
                                   const int32_t *__restrict__ keep, const int32_t *__restrict__ n_keep, int cap,
                                   float *__restrict__ out, int out_cap) {
    const int g = blockIdx.y;
    const int nk = min(max(n_keep[g], 0), out_cap);
    const long long total = (long long)nk * n_cols;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        int k = (int)(e / n_cols), c = (int)(e - (long long)k * n_cols);
        int i = keep[(long long)g * cap + k];
        out[((long long)g * out_cap + k) * n_cols + c] = rows[((long long)g * cap + i) * row_stride + c];
    }
}

}  // namespace

// Folds `(double)iou > thr` (TV_CPU) or `iou > (float)thr` (TV_CUDA) into a float compare.
void fdms_fold_threshold(double thr, int iou_mode, float *thr_f, int *inclusive) {
    float t = (float)thr;
    if (iou_mode == FDMS_IOU_TV_CUDA) {
        *thr_f = t;
        *inclusive = 0;
    } else if (iou_mode == FDMS_IOU_NCNN) {  // float ovr >= float thresh
        *thr_f = t;
        *inclusive = 1;
    } else {
        // smallest float f with (double)f > thr: t itself if t > thr, else the next float up
        *thr_f = t;
        *inclusive = ((double)t > thr) ? 1 : 0;
    }
}

int fdms_merge_small_limit(const fdms_ctx *ctx) {
    return ctx->nms_small_limit < SMALL_LIMIT ? ctx->nms_small_limit : SMALL_LIMIT;
}

int fdms_merge_small_launch(fdms_ctx *ctx, const MergeParams &P, const EmitParams &E, int n_img, float thr_f,
                            int inclusive, cudaStream_t st) {
    if (ctx->pdl) {
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.gridDim = dim3(n_img);
        cfg.blockDim = dim3(SMALL_THREADS);
        cfg.stream = st;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        FDMS_CUDA(ctx, cudaLaunchKernelEx(&cfg, merge_small_kernel, P, E, fdms_merge_small_limit(ctx), thr_f, inclusive));
    } else {
        merge_small_kernel<<<n_img, SMALL_THREADS, 0, st>>>(P, E, fdms_merge_small_limit(ctx), thr_f, inclusive);
    }
    FDMS_LAUNCH_CHECK(ctx, "merge_small_kernel");
    return FDMS_OK;
}

int fdms_nms_impl(fdms_ctx *ctx, const float *rows, int row_stride, int box_off, int score_off, const int32_t *tie,
                  const int32_t *count, int n_seg, int cap, double iou_thres, int iou_mode, int max_nms, int32_t *keep,
                  int32_t *n_keep, cudaStream_t st) {
    if (!rows || !count || !keep || !n_keep || n_seg <= 0 || cap <= 0 || row_stride <= 0 || box_off < 0 ||
        score_off < 0 || (iou_mode != FDMS_IOU_TV_CPU && iou_mode != FDMS_IOU_TV_CUDA && iou_mode != FDMS_IOU_NCNN))
        return fdms_fail(ctx, FDMS_ERR_INVALID, "fdms_nms: bad arguments");
    const int cap_e = (max_nms > 0 && max_nms < cap) ? max_nms : cap;
    float thr_f;
    int inclusive;
    fdms_fold_threshold(iou_thres, iou_mode, &thr_f, &inclusive);
    // ---- segments of at most small_limit boxes: one fused launch -------------------------------------------
    const int small_limit = ctx->nms_small_limit < SMALL_LIMIT ? ctx->nms_small_limit : SMALL_LIMIT;
    if (small_limit > 0 && !ctx->nms_small_done) {
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.gridDim = dim3(n_seg);
        cfg.blockDim = dim3(SMALL_THREADS);
        cfg.stream = st;
        cfg.attrs = at;
        cfg.numAttrs = ctx->pdl ? 1 : 0;
#define FDMS_NMS_SMALL(M)                                                                                                    \
    FDMS_CUDA(ctx, cudaLaunchKernelEx(&cfg, nms_small_kernel<M>, rows, row_stride, box_off, score_off, tie, count, cap, max_nms, \
                                      small_limit, thr_f, inclusive, keep, n_keep))
        if (iou_mode == FDMS_IOU_TV_CUDA) FDMS_NMS_SMALL(FDMS_IOU_TV_CUDA);
        else if (iou_mode == FDMS_IOU_NCNN) FDMS_NMS_SMALL(FDMS_IOU_NCNN);
        else FDMS_NMS_SMALL(FDMS_IOU_TV_CPU);
#undef FDMS_NMS_SMALL
        FDMS_LAUNCH_CHECK(ctx, "nms_small_kernel");
    }
    if (small_limit > 0 && (cap <= small_limit || ctx->post_small_only || ctx->small_only_sticky)) return FDMS_OK;  // no segment can be larger / the caller checks
    const int skip_le = small_limit > 0 ? small_limit : -1;
    // ---- larger segments: sort -> mask -> sweep ----------------------------------------------------------------
    if ((cap_e + 63) / 64 > SWEEP_MAX_NBLK)
        return fdms_fail(ctx, FDMS_ERR_UNSUPPORTED, "fdms_nms: more than %d boxes per segment", SWEEP_MAX_NBLK * 64);
    long long pad = 64;
    while (pad < cap) pad <<= 1;
    const bool global_sort = pad > SORT_SMEM_ITEMS;
    if (global_sort) {
        if (int rc = fdms_reserve(ctx, ctx->sort_a, (size_t)n_seg * pad * 8)) return rc;
        if (int rc = fdms_reserve(ctx, ctx->idx_a, (size_t)n_seg * pad * 4)) return rc;
    }
    if (int rc = fdms_reserve(ctx, ctx->idx_b, (size_t)n_seg * cap_e * 4)) return rc;
    if (int rc = fdms_reserve(ctx, ctx->boxes, (size_t)n_seg * cap_e * 16)) return rc;
    const bool fresh_plan = ctx->seginfo.bytes < sizeof(PlanHeader) + (size_t)(n_seg + 1) * sizeof(SegInfo);
    if (int rc = fdms_reserve(ctx, ctx->seginfo, sizeof(PlanHeader) + (size_t)(n_seg + 1) * sizeof(SegInfo))) return rc;
    if (fresh_plan) FDMS_CUDA(ctx, cudaMemsetAsync(ctx->seginfo.p, 0, ctx->seginfo.bytes, st));
    // many segments: one CTA per segment doing the work-efficient blocked greedy sweep; few segments: the
    // bit-matrix path spreads the K^2/2 pair tests over all SMs
    const bool greedy = n_seg >= ctx->nms_greedy_min_seg;
    long long budget_words = 0;
    if (!greedy) {
        const long long nblk_e = (cap_e + 63) / 64;
        size_t mask_upper = (size_t)n_seg * (size_t)cap_e * (size_t)nblk_e * 8;
        size_t mask_bytes = mask_upper < ctx->mask_budget ? mask_upper : ctx->mask_budget;
        if (int rc = fdms_reserve(ctx, ctx->mask, mask_bytes)) return rc;
        budget_words = (long long)(ctx->mask.bytes / 8);
    }

    PlanHeader *hdr = reinterpret_cast<PlanHeader *>(ctx->seginfo.p);
    SegInfo *seginfo = reinterpret_cast<SegInfo *>(hdr + 1);
    if (!ctx->nms_attr) {
        FDMS_CUDA(ctx, cudaFuncSetAttribute(nms_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            SORT_SMEM_SLOTS * 12));
        ctx->nms_attr = true;
    }
    nms_sort_kernel<<<n_seg, SORT_THREADS, SORT_SMEM_SLOTS * 12, st>>>(
        rows, row_stride, box_off, score_off, tie, count, n_seg, cap, cap_e, max_nms,
        reinterpret_cast<unsigned long long *>(ctx->sort_a.p), reinterpret_cast<uint32_t *>(ctx->idx_a.p), pad,
        reinterpret_cast<uint32_t *>(ctx->idx_b.p), reinterpret_cast<float4 *>(ctx->boxes.p), seginfo, hdr,
        budget_words, skip_le, ctx->nms_reg_sort);
    FDMS_LAUNCH_CHECK(ctx, "nms_sort_kernel");
    if (greedy && ctx->nms_bin != 2 && cap_e <= BIN_MAX_K) {  // FDMS_NMS_BIN=2: the list kernel (A/B); 0: no grid
        const size_t bin_smem = sizeof(BinShared) + (size_t)cap_e * sizeof(uint16_t);
        if (!ctx->nms_bin_attr) {
            const int mx = (int)(sizeof(BinShared) + (size_t)BIN_MAX_K * sizeof(uint16_t));
            FDMS_CUDA(ctx, cudaFuncSetAttribute(nms_binned_kernel<FDMS_IOU_TV_CPU>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
            FDMS_CUDA(ctx, cudaFuncSetAttribute(nms_binned_kernel<FDMS_IOU_TV_CUDA>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
            FDMS_CUDA(ctx, cudaFuncSetAttribute(nms_binned_kernel<FDMS_IOU_NCNN>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
            ctx->nms_bin_attr = true;
        }
#define FDMS_NMS_BINNED(M)                                                                                                \
    nms_binned_kernel<M><<<n_seg, GREEDY_THREADS, bin_smem, st>>>(                                                            \
        seginfo, cap_e, reinterpret_cast<const float4 *>(ctx->boxes.p), reinterpret_cast<const uint32_t *>(ctx->idx_b.p), thr_f, \
        inclusive, keep, n_keep, cap, count, skip_le, ctx->nms_bin)
        if (iou_mode == FDMS_IOU_TV_CUDA) FDMS_NMS_BINNED(FDMS_IOU_TV_CUDA);
        else if (iou_mode == FDMS_IOU_NCNN) FDMS_NMS_BINNED(FDMS_IOU_NCNN);
        else FDMS_NMS_BINNED(FDMS_IOU_TV_CPU);
#undef FDMS_NMS_BINNED
        FDMS_LAUNCH_CHECK(ctx, "nms_binned_kernel");
        return FDMS_OK;
    }
    if (greedy) {
#define FDMS_NMS_GREEDY(M)                                                                                         \
    nms_greedy_kernel<M><<<n_seg, GREEDY_THREADS, 0, st>>>(seginfo, cap_e, reinterpret_cast<float4 *>(ctx->boxes.p),  \
                                                           reinterpret_cast<const uint32_t *>(ctx->idx_b.p), thr_f,   \
                                                           inclusive, keep, n_keep, cap, count, skip_le)
        if (iou_mode == FDMS_IOU_TV_CUDA) FDMS_NMS_GREEDY(FDMS_IOU_TV_CUDA);
        else if (iou_mode == FDMS_IOU_NCNN) FDMS_NMS_GREEDY(FDMS_IOU_NCNN);
        else FDMS_NMS_GREEDY(FDMS_IOU_TV_CPU);
#undef FDMS_NMS_GREEDY
        FDMS_LAUNCH_CHECK(ctx, "nms_greedy_kernel");
        return FDMS_OK;
    }
    const int mask_grid = ctx->sm_count * 8;
#define FDMS_NMS_MASK(M)                                                                                                \
    nms_mask_kernel<M><<<mask_grid, MASK_THREADS, 0, st>>>(seginfo, hdr, n_seg, cap_e,                                      \
                                                           reinterpret_cast<const float4 *>(ctx->boxes.p),                 \
                                                           reinterpret_cast<unsigned long long *>(ctx->mask.p), thr_f, inclusive)
    if (iou_mode == FDMS_IOU_TV_CUDA) FDMS_NMS_MASK(FDMS_IOU_TV_CUDA);
    else if (iou_mode == FDMS_IOU_NCNN) FDMS_NMS_MASK(FDMS_IOU_NCNN);
    else FDMS_NMS_MASK(FDMS_IOU_TV_CPU);
#undef FDMS_NMS_MASK
    FDMS_LAUNCH_CHECK(ctx, "nms_mask_kernel");
    nms_sweep_kernel<<<n_seg, SWEEP_THREADS, 0, st>>>(seginfo, reinterpret_cast<const unsigned long long *>(ctx->mask.p),
                                                      cap_e, reinterpret_cast<const uint32_t *>(ctx->idx_b.p), keep,
                                                      n_keep, cap, count, skip_le);
    FDMS_LAUNCH_CHECK(ctx, "nms_sweep_kernel");
    return FDMS_OK;
}

extern "C" {

FDMS_API int fdms_nms(fdms_ctx *ctx, const float *rows, int row_stride, int box_off, int score_off,
                      const int32_t *tie, const int32_t *count, int n_seg, int cap, double iou_thres, int iou_mode,
                      int max_nms, int32_t *keep, int32_t *n_keep, fdms_stream stream) {
    FDMS_BIND(ctx);
    return fdms_nms_impl(ctx, rows, row_stride, box_off, score_off, tie, count, n_seg, cap, iou_thres, iou_mode,
                         max_nms, keep, n_keep, (cudaStream_t)stream);
}

FDMS_API int fdms_gather_rows(fdms_ctx *ctx, const float *rows, int row_stride, int n_cols, const int32_t *keep,
                              const int32_t *n_keep, int n_seg, int cap, float *out, int out_cap,
                              fdms_stream stream) {
    FDMS_BIND(ctx);
    if (!rows || !keep || !n_keep || !out || n_seg <= 0 || n_seg > 65535 || cap <= 0 || out_cap <= 0 || n_cols <= 0 ||
        n_cols > row_stride)
        return fdms_fail(ctx, FDMS_ERR_INVALID, "fdms_gather_rows: bad arguments");
    int bx = (int)ceil_div64((int64_t)out_cap * n_cols, 256);
    if (bx > 64) bx = 64;
    dim3 grid(bx, n_seg);
    gather_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(rows, row_stride, n_cols, keep, n_keep, cap, out, out_cap);
    FDMS_LAUNCH_CHECK(ctx, "gather_rows_kernel");
    return FDMS_OK;
}

}  // extern "C"
