// Batched NMS with torchvision.ops.nms semantics (call sites utils/general.py:584 and
// multi_scale_face_detector.py:234), all segments (image x scale) of a batch per launch and
// no host synchronisation between the stages:
//
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
// IoU arithmetic, selected by iou_mode:
//   TV_CPU : inter / ((area_i + area_j) - inter), suppressed iff (double)iou > thr (double)
//   TV_CUDA: inter / (fma(w_j, h_j, area_i) - inter), suppressed iff iou > (float)thr
// The double comparison is folded on the host into a float threshold plus an inclusive flag.
#include <math.h>

#include "common.cuh"
#include "nms.cuh"

namespace {

constexpr int SORT_THREADS = 1024;
constexpr int SORT_SMEM_ITEMS = 8192;  // 8 B key + 4 B index each = 96 KB
constexpr int MASK_THREADS = 256;      // 64 rows x 4 column blocks
constexpr int SWEEP_THREADS = 512;
constexpr int SWEEP_MAX_NBLK = 4096;   // 262144 boxes per segment
constexpr int SMALL_THREADS = 1024;
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

__global__ void __launch_bounds__(SORT_THREADS)
    nms_sort_kernel(const float *__restrict__ rows, int row_stride, int box_off, int score_off,
                    const int32_t *__restrict__ tie, const int32_t *__restrict__ count, int n_seg, int cap, int cap_e,
                    int max_nms, unsigned long long *gkeys, uint32_t *gidx, long long gpad,
                    uint32_t *__restrict__ sorted_idx, float4 *__restrict__ sorted_boxes, SegInfo *seginfo,
                    PlanHeader *hdr, long long mask_budget_words, int skip_le) {
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
        idx = reinterpret_cast<uint32_t *>(sm + (size_t)SORT_SMEM_ITEMS * 8);
    } else {
        keys = gkeys + (long long)g * gpad;
        idx = gidx + (long long)g * gpad;
    }
    const float *seg_rows = rows + (long long)g * cap * row_stride;
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
    const int K = (max_nms > 0) ? min(n, max_nms) : n;  // utils/general.py:578-579
    for (int k = tid; k < K; k += SORT_THREADS) {
        uint32_t ii = idx[k];
        const float *r = seg_rows + (long long)ii * row_stride + box_off;
        sorted_idx[(long long)g * cap_e + k] = ii;
        sorted_boxes[(long long)g * cap_e + k] = make_float4(r[0], r[1], r[2], r[3]);
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
__device__ __forceinline__ bool iou_gt(const float4 a, float area_a, const float4 b, float thr, bool inclusive,
                                       bool quick) {
    float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
    float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
    if (quick && (!(xx2 > xx1) || !(yy2 > yy1))) return false;
    float w = fmaxf(f_sub(xx2, xx1), 0.f), h = fmaxf(f_sub(yy2, yy1), 0.f);
    float inter = f_mul(w, h);
    float bw = f_sub(b.z, b.x), bh = f_sub(b.w, b.y);
    float den;
    if (MODE == FDMS_IOU_TV_CUDA) den = f_sub(__fmaf_rn(bw, bh, area_a), inter);
    else den = f_sub(f_add(area_a, f_mul(bw, bh)), inter);
    float iou = f_div(inter, den);
    return inclusive ? (iou >= thr) : (iou > thr);  // NaN: false either way
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
            const float area_a = f_mul(f_sub(a.z, a.x), f_sub(a.w, a.y));
            unsigned long long bits = 0;
            const int col0 = cblk * 64;
            const int jn = min(64, si.n - col0);
            int j0 = (cblk == rb) ? (tid & 63) + 1 : 0;  // only later boxes can be suppressed
#pragma unroll 4
            for (int j = j0; j < jn; j++)
                if (iou_gt<MODE>(a, area_a, s_col[q * 64 + j], thr, inclusive != 0, quick)) bits |= 1ull << j;
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
// would suppress i (pairs with a pre-suppressed member are skipped -- they cannot matter).  (3) parallel
// resolution to the greedy fixpoint: an undecided candidate is removed as soon as a KEPT earlier candidate
// suppresses it, and kept as soon as no undecided or kept earlier candidate does; the first undecided one
// always qualifies, and typical suppression graphs settle in a handful of rounds (vs one serial step per kept
// box).  Result: s_K = bitset of kept candidates.  All CTA threads must call this (it contains barriers).
template <int MODE>
__device__ __forceinline__ void greedy_block(const float4 *blk, int nb, float thr, bool inclusive, bool quick,
                                             unsigned long long (*s_colmask)[4], unsigned long long *s_sup,
                                             unsigned long long *s_U, unsigned long long *s_K) {
    const int tid = threadIdx.x;
    const int i = tid & (GREEDY_BS - 1), grp = tid >> 8;
    const int iw = i >> 6;
    const unsigned long long ibit = 1ull << (i & 63);
    __syncthreads();  // s_sup complete
    {
        unsigned long long bits = 0ull;
        if (i < nb && !(s_sup[iw] & ibit)) {
            const float4 bi = blk[i];
            const unsigned long long skip = s_sup[grp];
            const int j0 = grp * 64, j1 = min(grp * 64 + 64, i);
            for (int j = j0; j < j1; j++) {
                if ((skip >> (j & 63)) & 1ull) continue;
                const float4 a = blk[j];
                if (iou_gt<MODE>(a, f_mul(f_sub(a.z, a.x), f_sub(a.w, a.y)), bi, thr, inclusive, quick))
                    bits |= 1ull << (j & 63);
            }
        }
        s_colmask[i][grp] = bits;
    }
    if (tid < 4) {
        const int left = nb - tid * 64;
        const unsigned long long valid = left >= 64 ? ~0ull : (left > 0 ? ((1ull << left) - 1ull) : 0ull);
        s_U[tid] = ~s_sup[tid] & valid;
        s_K[tid] = 0ull;
    }
    __syncthreads();
    for (;;) {
        const unsigned long long u0 = s_U[0], u1 = s_U[1], u2 = s_U[2], u3 = s_U[3];
        if ((u0 | u1 | u2 | u3) == 0ull) break;  // uniform
        bool undecided = false, by_kept = false, by_undecided = false;
        if (tid < GREEDY_BS) {
            const unsigned long long uw = iw == 0 ? u0 : (iw == 1 ? u1 : (iw == 2 ? u2 : u3));
            undecided = (uw & ibit) != 0ull;
            if (undecided) {
                const unsigned long long c0 = s_colmask[i][0], c1 = s_colmask[i][1], c2 = s_colmask[i][2],
                                         c3 = s_colmask[i][3];
                by_kept = ((c0 & s_K[0]) | (c1 & s_K[1]) | (c2 & s_K[2]) | (c3 & s_K[3])) != 0ull;
                by_undecided = ((c0 & u0) | (c1 & u1) | (c2 & u2) | (c3 & u3)) != 0ull;
            }
        }
        __syncthreads();  // every thread has read U and K of this round
        if (undecided) {
            if (by_kept) {
                atomicAnd(&s_U[iw], ~ibit);
            } else if (!by_undecided) {
                atomicOr(&s_K[iw], ibit);
                atomicAnd(&s_U[iw], ~ibit);
            }
        }
        __syncthreads();
    }
}

// Fused single-CTA NMS for segments of at most SMALL_LIMIT (1024) boxes, everything in shared memory:
//   rank sort  : thread i counts the keys smaller than its own (keys are unique: score desc, tie asc) -- no
//                barriers, broadcast shared-memory reads -- and scatters its box to that rank;
//   suppression: the same blocked greedy sweep as nms_greedy_kernel (256 candidates per step, kept boxes
//                compacted in place at the front of the box array).
// No global scratch, one launch, ~3 barriers per 256 boxes.
template <int MODE>
__global__ void __launch_bounds__(SMALL_THREADS)
    nms_small_kernel(const float *__restrict__ rows, int row_stride, int box_off, int score_off,
                     const int32_t *__restrict__ tie, const int32_t *__restrict__ count, int cap, int max_nms,
                     int small_limit, float thr, int inclusive, int32_t *__restrict__ keep,
                     int32_t *__restrict__ n_keep) {
    __shared__ unsigned long long s_keys[SMALL_LIMIT];
    __shared__ float4 s_box[SMALL_LIMIT];
    __shared__ uint32_t s_idx[SMALL_LIMIT];
    __shared__ int s_rank[SMALL_LIMIT];
    __shared__ unsigned long long s_rowmask[GREEDY_BS][4];
    __shared__ unsigned long long s_sup[4], s_keptbits[4], s_U[4];
    __shared__ int s_nkeep;
    const int g = blockIdx.x, tid = threadIdx.x;
    const bool quick = thr > 0.f || (thr == 0.f && !inclusive);
    const int n = max(0, min(count[g], cap));
    if (n > small_limit) return;  // the sort + greedy / bit-matrix path takes it
    if (n == 0) {
        if (tid == 0) n_keep[g] = 0;
        return;
    }
    const float *seg_rows = rows + (long long)g * cap * row_stride;
    // ---- rank sort ----------------------------------------------------------------------------------------
    // keys are unique (tie = row index or item index; callers must not pass duplicate ties inside a segment), so
    // the rank of an item is the number of smaller keys.  All 1024 threads take part: item i = tid % pad, and the
    // 1024 / pad threads of an item split the key range and add their partial counts.
    int pad = 32;
    while (pad < n) pad <<= 1;
    const int parts = SMALL_THREADS / pad;
    const int item = tid & (pad - 1), part = tid / pad;
    if (tid < n) {
        const float *r = seg_rows + (long long)tid * row_stride;
        const uint32_t t = tie ? (uint32_t)tie[(long long)g * cap + tid] : (uint32_t)tid;
        s_keys[tid] = ((unsigned long long)score_key_desc(r[score_off]) << 32) | t;
    }
    if (tid < pad) s_rank[tid] = 0;
    if (tid == 0) s_nkeep = 0;
    __syncthreads();
    if (item < n) {
        const unsigned long long mykey = s_keys[item];
        const int per = (n + parts - 1) / parts, j0 = part * per, j1 = min(n, j0 + per);
        int cnt = 0;
#pragma unroll 4
        for (int j = j0; j < j1; j++) cnt += s_keys[j] < mykey;
        if (parts == 1) s_rank[item] = cnt;
        else if (cnt) atomicAdd(&s_rank[item], cnt);
    }
    __syncthreads();
    if (tid < n) {
        const float *r = seg_rows + (long long)tid * row_stride + box_off;
        const int rank = s_rank[tid];
        s_box[rank] = make_float4(r[0], r[1], r[2], r[3]);
        s_idx[rank] = (uint32_t)tid;
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
        if (i < nb) {
            bool sup = false;
            for (int q = grp; q < nk; q += 4) {
                const float4 a = s_box[q];
                sup |= iou_gt<MODE>(a, f_mul(f_sub(a.z, a.x), f_sub(a.w, a.y)), bi, thr, inclusive != 0, quick);
            }
            if (sup) atomicOr(&s_sup[i >> 6], 1ull << (i & 63));
        }
        // (2) + (3): column masks and parallel greedy resolution
        greedy_block<MODE>(s_box + base, nb, thr, inclusive != 0, quick, s_rowmask, s_sup, s_U, s_keptbits);
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
    if (tid == 0) n_keep[g] = s_nkeep;
}

// Blocked greedy NMS for larger segments, one CTA per segment, after nms_sort_kernel.  Boxes arrive in score
// order in global scratch; the boxes kept so far are compacted IN PLACE at the front of the same array (the
// write position never passes the read position) and streamed through shared memory in tiles.  Per step 256
// candidates are (1) tested against the kept list with early exit, (2) masked against each other (256 x 256
// bits), (3) resolved by warp 0 with one step per KEPT box, (4) appended.  Work ~ K * n_kept / 2 pair tests
// instead of K^2 / 2, and no K^2 bit-matrix.
template <int MODE>
__global__ void __launch_bounds__(GREEDY_THREADS)
    nms_greedy_kernel(const SegInfo *__restrict__ seginfo, int cap_e, float4 *sorted_boxes,
                      const uint32_t *__restrict__ sorted_idx, float thr, int inclusive, int32_t *__restrict__ keep,
                      int32_t *__restrict__ n_keep, int cap, const int32_t *__restrict__ count, int skip_le) {
    __shared__ float4 s_blk[GREEDY_BS];
    __shared__ float4 s_tile[GREEDY_TILE];
    __shared__ unsigned long long s_rowmask[GREEDY_BS][4];
    __shared__ unsigned long long s_sup[4], s_keptbits[4], s_U[4];
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
        // (1) against the kept list, tile by tile
        bool sup = false;
        for (int t0 = 0; t0 < nk; t0 += GREEDY_TILE) {
            const int tn = min(GREEDY_TILE, nk - t0);
            if (t0) __syncthreads();  // readers of the previous tile are done
            if (tid < tn) s_tile[tid] = boxes[t0 + tid];
            __syncthreads();
            if (i < nb && !sup) {
                for (int q = grp; q < tn; q += 4) {
                    const float4 a = s_tile[q];
                    if (iou_gt<MODE>(a, f_mul(f_sub(a.z, a.x), f_sub(a.w, a.y)), bi, thr, inclusive != 0, quick)) {
                        sup = true;
                        break;
                    }
                }
            }
        }
        if (sup) atomicOr(&s_sup[i >> 6], 1ull << (i & 63));
        // (2) + (3): column masks and parallel greedy resolution
        greedy_block<MODE>(s_blk, nb, thr, inclusive != 0, quick, s_rowmask, s_sup, s_U, s_keptbits);
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
    } else {
        // smallest float f with (double)f > thr: t itself if t > thr, else the next float up
        *thr_f = t;
        *inclusive = ((double)t > thr) ? 1 : 0;
    }
}

int fdms_nms_impl(fdms_ctx *ctx, const float *rows, int row_stride, int box_off, int score_off, const int32_t *tie,
                  const int32_t *count, int n_seg, int cap, double iou_thres, int iou_mode, int max_nms, int32_t *keep,
                  int32_t *n_keep, cudaStream_t st) {
    if (!rows || !count || !keep || !n_keep || n_seg <= 0 || cap <= 0 || row_stride <= 0 || box_off < 0 ||
        score_off < 0 || (iou_mode != FDMS_IOU_TV_CPU && iou_mode != FDMS_IOU_TV_CUDA))
        return fdms_fail(ctx, FDMS_ERR_INVALID, "fdms_nms: bad arguments");
    const int cap_e = (max_nms > 0 && max_nms < cap) ? max_nms : cap;
    float thr_f;
    int inclusive;
    fdms_fold_threshold(iou_thres, iou_mode, &thr_f, &inclusive);
    // ---- segments of at most small_limit boxes: one fused launch -------------------------------------------
    const int small_limit = ctx->nms_small_limit < SMALL_LIMIT ? ctx->nms_small_limit : SMALL_LIMIT;
    if (small_limit > 0) {
        if (iou_mode == FDMS_IOU_TV_CUDA)
            nms_small_kernel<FDMS_IOU_TV_CUDA><<<n_seg, SMALL_THREADS, 0, st>>>(
                rows, row_stride, box_off, score_off, tie, count, cap, max_nms, small_limit, thr_f, inclusive, keep, n_keep);
        else
            nms_small_kernel<FDMS_IOU_TV_CPU><<<n_seg, SMALL_THREADS, 0, st>>>(
                rows, row_stride, box_off, score_off, tie, count, cap, max_nms, small_limit, thr_f, inclusive, keep, n_keep);
        FDMS_LAUNCH_CHECK(ctx, "nms_small_kernel");
        if (cap <= small_limit) return FDMS_OK;  // no segment can be larger
    }
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
                                            SORT_SMEM_ITEMS * 12));
        ctx->nms_attr = true;
    }
    nms_sort_kernel<<<n_seg, SORT_THREADS, SORT_SMEM_ITEMS * 12, st>>>(
        rows, row_stride, box_off, score_off, tie, count, n_seg, cap, cap_e, max_nms,
        reinterpret_cast<unsigned long long *>(ctx->sort_a.p), reinterpret_cast<uint32_t *>(ctx->idx_a.p), pad,
        reinterpret_cast<uint32_t *>(ctx->idx_b.p), reinterpret_cast<float4 *>(ctx->boxes.p), seginfo, hdr,
        budget_words, skip_le);
    FDMS_LAUNCH_CHECK(ctx, "nms_sort_kernel");
    if (greedy) {
        if (iou_mode == FDMS_IOU_TV_CUDA)
            nms_greedy_kernel<FDMS_IOU_TV_CUDA><<<n_seg, GREEDY_THREADS, 0, st>>>(
                seginfo, cap_e, reinterpret_cast<float4 *>(ctx->boxes.p), reinterpret_cast<const uint32_t *>(ctx->idx_b.p),
                thr_f, inclusive, keep, n_keep, cap, count, skip_le);
        else
            nms_greedy_kernel<FDMS_IOU_TV_CPU><<<n_seg, GREEDY_THREADS, 0, st>>>(
                seginfo, cap_e, reinterpret_cast<float4 *>(ctx->boxes.p), reinterpret_cast<const uint32_t *>(ctx->idx_b.p),
                thr_f, inclusive, keep, n_keep, cap, count, skip_le);
        FDMS_LAUNCH_CHECK(ctx, "nms_greedy_kernel");
        return FDMS_OK;
    }
    const int mask_grid = ctx->sm_count * 8;
    if (iou_mode == FDMS_IOU_TV_CUDA)
        nms_mask_kernel<FDMS_IOU_TV_CUDA><<<mask_grid, MASK_THREADS, 0, st>>>(
            seginfo, hdr, n_seg, cap_e, reinterpret_cast<const float4 *>(ctx->boxes.p),
            reinterpret_cast<unsigned long long *>(ctx->mask.p), thr_f, inclusive);
    else
        nms_mask_kernel<FDMS_IOU_TV_CPU><<<mask_grid, MASK_THREADS, 0, st>>>(
            seginfo, hdr, n_seg, cap_e, reinterpret_cast<const float4 *>(ctx->boxes.p),
            reinterpret_cast<unsigned long long *>(ctx->mask.p), thr_f, inclusive);
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
    if (int rc = fdms_bind(ctx)) return rc;
    return fdms_nms_impl(ctx, rows, row_stride, box_off, score_off, tie, count, n_seg, cap, iou_thres, iou_mode,
                         max_nms, keep, n_keep, (cudaStream_t)stream);
}

FDMS_API int fdms_gather_rows(fdms_ctx *ctx, const float *rows, int row_stride, int n_cols, const int32_t *keep,
                              const int32_t *n_keep, int n_seg, int cap, float *out, int out_cap,
                              fdms_stream stream) {
    if (int rc = fdms_bind(ctx)) return rc;
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
