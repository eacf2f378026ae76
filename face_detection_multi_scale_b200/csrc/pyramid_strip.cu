// Strip kernel of the fused letterbox pyramid: ONE pass over the source feeds every exact-factor scale.
//
// Replaces, for the scales it covers, utils/datasets.py:873-903 (letterbox) and multi_scale_face_detector.py:96-105
// (BGR->RGB, CHW, half, /255) -- same arithmetic as pyramid.cu (OpenCV INTER_LINEAR u8 fixed point, bit exact), other
// data movement.  The tile kernel fetches the source once per SCALE through L1/L2 (1080p -> 640/960/1280: 7.4 x the
// source bytes in L1 sector requests, every thread window re-touching its neighbours' sectors); here the exact factors
// k:1 and 3:2 are used to cut the source into strips that no scale shares with a neighbour:
//
//   period L = lcm of the scales' source-row periods (k rows -> 1 output row; 3 rows -> 2 output rows for 3:2);
//   a strip = L source rows x CW source columns; 1080p -> 640 / 960 / 1280 gives L = 6 (2 + 3 + 4 output rows).
//
// A persistent CTA per SM runs a ring of shared-memory stages.  One producer warp takes strip numbers from a global
// counter and issues one TMA bulk copy (cp.async.bulk -> UBLKCP) per needed source row, completing on the stage's
// full mbarrier; the consumer warps (16) take (scale, row unit, 32 column groups) work units from the stage's counter, read
// their taps from shared memory (conflict-free for the 3:2 and 2:1 windows), and write 8 adjacent fp16 columns of a
// plane with one 128-bit store per channel and row -- a warp writes 512 contiguous bytes.  Every source byte crosses
// L2 -> SM exactly once and no thread ever waits on a global load.  A warp that has drained a stage arrives on its empty
// mbarrier and moves on, so warps drift across stages and the uneven unit costs average out.  Letterbox border rows
// are separate, store-only work items.
//
// Arithmetic (identical to the general path by construction, checked against it and the oracle):
//   3:2   OpenCV's weights are 1536/512 and 512/1536 on both axes (f = 0.25 / 0.75 exactly), i.e. 512 * {3,1} / {1,3}:
//         h = S0*a0 + S1*a1 = 512*H with H = 3*S0 + S1 (one or two IDP.4A on the raw words, the byte positions are
//         compile-time constants), (b*(h>>4))>>16 = (kb*H)>>2, so t = ((3*H0)>>2) + (H1>>2) + 2 for even output rows
//         and (H1>>2) + ((3*H2)>>2) + 2 for odd ones; the factor 3 of the outer rows is folded into the IDP weights.
//   k even  weights 1024/1024: t = a + b + c + d + 2 over the centre 2 x 2 block (IDP.4A chains).
//   k odd   weights 2048/0: the centre pixel; PRMT drops the byte into the 2^23 magic number.
//   pixel = t >> 2; /255 and the fp16 rounding as in pyramid.cu (exact for all 256 byte values).
#include "pyramid.cuh"

namespace {

constexpr int ST_MAX_CONSUMER_WARPS = 31;
constexpr int ST_MAX_THREADS = 32 * (ST_MAX_CONSUMER_WARPS + 1);  // + the producer warp (the last warp of the CTA)
constexpr int ST_MAX_STAGES = 8;
constexpr int ST_MAX_SLOTS = 16;   // source rows kept in a stage
constexpr int ST_MAX_RU = 16;      // row units of a scale per strip
constexpr int ST_MAX_L = 48;       // source rows of a strip
constexpr int ST_MAX_WU = 160;     // warp units of a strip
enum { ST_POINT = 0, ST_HALF = 1, ST_R32 = 2 };

struct StripScale {
    __half *out;
    long long img_stride, cs;  // element strides (row stride = dst_w)
    int type, k;
    int dst_w, dst_h, top, bottom;  // letterbox rows above / below the image
    int rows_per_period;       // output rows produced per strip
    int n_ru;                  // row units per strip (3:2: pairs of rows)
    int groups;                // 8-column groups per strip
    int group_bytes;           // source bytes of a group (24 k, or 36)
    int n_units;               // n_ru * groups
    int wu_begin;              // first warp unit of this scale in a strip's list
    int border_wu;             // warp units of this scale's border item
    int slot0[ST_MAX_RU];      // byte offset, inside a stage, of the first source row of every row unit
    int out_chunk_cols;        // output columns per strip
    int ru_stride;             // output elements between the first rows of consecutive row units
    long long plane[3];        // element offset of the plane that source channel c goes to (BGR -> RGB swaps 0 and 2)
};

struct StripParams {
    const uint8_t *src;
    long long img_stride, pitch;
    int n_img, periods, chunks, L;
    int n_slots, row_bytes, stage_bytes, n_stages;
    int slot_row[ST_MAX_SLOTS];  // source row of each slot inside the period
    int n_scales;
    StripScale sc[FDMS_MAX_SCALES];
    int n_border_scales, border_scale[FDMS_MAX_SCALES];
    int n_border_items, n_items;  // border items come first
    int wu_per_strip;
    unsigned wu_info[ST_MAX_WU];  // per warp unit of a strip: scale | row unit << 8 | first column group << 16
    int flip, border;
    unsigned *counters;  // [0] next item, [1] CTAs done
};

struct StripShared {
    alignas(8) uint64_t full[ST_MAX_STAGES];
    alignas(8) uint64_t empty[ST_MAX_STAGES];
    // per stage, written by the producer: the item, the hand-out counter of its work units and, per scale, the element
    // offset of the first output row of the item (the consumers do no item arithmetic)
    __half *out_ptr[ST_MAX_STAGES][FDMS_MAX_SCALES][3];  // [channel]: plane pointer at the item's first row and column
    int item[ST_MAX_STAGES];
    int next_wu[ST_MAX_STAGES];
};

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Plain loads (not asm): they stay ordered behind the mbarrier wait ("memory" clobber) and the words a unit never uses
// are eliminated; the pointers derive from the dynamic shared array, so these are LDS.
__device__ __forceinline__ uint32_t lds32(const uint8_t *p) { return *reinterpret_cast<const uint32_t *>(p); }
__device__ __forceinline__ uint2 lds64(const uint8_t *p) { return *reinterpret_cast<const uint2 *>(p); }
__device__ __forceinline__ uint4 lds128(const uint8_t *p) { return *reinterpret_cast<const uint4 *>(p); }

// rn16(rn32(v/255)) == rn16(rn32(v * rn32(1/255))) for all 256 byte values (tests/test_host_logic.py); `bits` are the
// float bits of 2^23 + v.
__device__ __forceinline__ float unit_scale(unsigned bits) {
    const float c = 0.003921568859368563f;
    return __fmaf_rn(__uint_as_float(bits), c, -8388608.0f * c);
}
__device__ __forceinline__ unsigned pack2(unsigned bits_a, unsigned bits_b) {
    const __half2 h = __floats2half2_rn(unit_scale(bits_a), unit_scale(bits_b));
    return *reinterpret_cast<const unsigned *>(&h);
}
__device__ __forceinline__ unsigned magic_of_t(unsigned t) { return (t >> 2) + 0x4B000000u; }  // t = 4 * pixel + rounding
__device__ __forceinline__ void store8(__half *p, const unsigned *bits) {  // 8 adjacent columns of one plane row
    uint4 v;
    v.x = pack2(bits[0], bits[1]);
    v.y = pack2(bits[2], bits[3]);
    v.z = pack2(bits[4], bits[5]);
    v.w = pack2(bits[6], bits[7]);
    *reinterpret_cast<uint4 *>(p) = v;
}

// byte `o` (compile-time) of a window held as 32-bit words: w0 * B[o] + w1 * B[o + 3] + acc with IDP.4A
template <int O, int W0, int W1>
__device__ __forceinline__ unsigned tap2(const uint32_t *w, unsigned acc) {
    constexpr int i0 = O >> 2, s0 = (O & 3) * 8, i1 = (O + 3) >> 2, s1 = ((O + 3) & 3) * 8;
    if (i0 == i1) return __dp4a(w[i0], (unsigned)((W0 << s0) | (W1 << s1)), acc);
    return __dp4a(w[i1], (unsigned)(W1 << s1), __dp4a(w[i0], (unsigned)(W0 << s0), acc));
}

// ---- 3:2 ---------------------------------------------------------------------------------------------------------
// One unit = 8 output columns x 2 output rows from 12 source pixels (36 bytes = 9 aligned words) x 3 source rows.
template <int C, int J>
__device__ __forceinline__ void r32_col(const uint32_t *r0, const uint32_t *r1, const uint32_t *r2, unsigned *te, unsigned *to) {
    constexpr int M = J >> 1, ODD = J & 1;
    constexpr int O = 3 * (3 * M + ODD) + C;        // byte of the left tap, channel C
    constexpr int A = ODD ? 1 : 3, B = ODD ? 3 : 1;  // horizontal weights / 512
    const unsigned q0 = tap2<O, 3 * A, 3 * B>(r0, 0u) >> 2;  // (3 * H0) >> 2
    const unsigned q1 = tap2<O, A, B>(r1, 0u) >> 2;          // H1 >> 2
    const unsigned q2 = tap2<O, 3 * A, 3 * B>(r2, 0u) >> 2;  // (3 * H2) >> 2
    te[J] = magic_of_t(q0 + q1 + 2u);
    to[J] = magic_of_t(q1 + q2 + 2u);
}
template <int C>
__device__ __forceinline__ void r32_channel(const uint32_t *r0, const uint32_t *r1, const uint32_t *r2, __half *pe, __half *po) {
    unsigned te[8], to[8];
    r32_col<C, 0>(r0, r1, r2, te, to);
    r32_col<C, 1>(r0, r1, r2, te, to);
    r32_col<C, 2>(r0, r1, r2, te, to);
    r32_col<C, 3>(r0, r1, r2, te, to);
    r32_col<C, 4>(r0, r1, r2, te, to);
    r32_col<C, 5>(r0, r1, r2, te, to);
    r32_col<C, 6>(r0, r1, r2, te, to);
    r32_col<C, 7>(r0, r1, r2, te, to);
    store8(pe, te);
    store8(po, to);
}
__device__ __forceinline__ void unit_r32(const uint8_t *s0, int row_bytes, __half *const *pc, int ys) {
    uint32_t r0[9], r1[9], r2[9];
#pragma unroll
    for (int i = 0; i < 9; i++) {
        r0[i] = lds32(s0 + 4 * i);
        r1[i] = lds32(s0 + row_bytes + 4 * i);
        r2[i] = lds32(s0 + 2 * row_bytes + 4 * i);
    }
    r32_channel<0>(r0, r1, r2, pc[0], pc[0] + ys);
    r32_channel<1>(r0, r1, r2, pc[1], pc[1] + ys);
    r32_channel<2>(r0, r1, r2, pc[2], pc[2] + ys);
}

// ---- even factors: the centre 2 x 2 block ---------------------------------------------------------------------------
// One unit = 8 output columns x 1 row; source window of 24 K bytes per row (6 K aligned words), rows K/2-1 and K/2.
template <int K, int C, int J>
__device__ __forceinline__ unsigned half_col(const uint32_t *a, const uint32_t *b) {
    constexpr int O = 3 * (K * J + K / 2 - 1) + C;
    return magic_of_t(tap2<O, 1, 1>(b, tap2<O, 1, 1>(a, 2u)));
}
template <int K>
__device__ __forceinline__ void unit_half(const uint8_t *s0, int row_bytes, __half *const *pc) {
    constexpr int NW = 6 * K;
    uint32_t a[NW], b[NW];
#pragma unroll
    for (int i = 0; i < NW; i += 4) {
        const uint4 va = lds128(s0 + 4 * i), vb = lds128(s0 + row_bytes + 4 * i);
        a[i] = va.x; a[i + 1] = va.y; a[i + 2] = va.z; a[i + 3] = va.w;
        b[i] = vb.x; b[i + 1] = vb.y; b[i + 2] = vb.z; b[i + 3] = vb.w;
    }
#define FDMS_HALF_CH(C)                                                                                              \
    {                                                                                                                \
        unsigned t[8] = {half_col<K, C, 0>(a, b), half_col<K, C, 1>(a, b), half_col<K, C, 2>(a, b), half_col<K, C, 3>(a, b), \
                         half_col<K, C, 4>(a, b), half_col<K, C, 5>(a, b), half_col<K, C, 6>(a, b), half_col<K, C, 7>(a, b)}; \
        store8(pc[C], t);                                                                                            \
    }
    FDMS_HALF_CH(0)
    FDMS_HALF_CH(1)
    FDMS_HALF_CH(2)
#undef FDMS_HALF_CH
}

// ---- odd factors: the centre pixel -----------------------------------------------------------------------------------
template <int K, int C, int J>
__device__ __forceinline__ unsigned point_col(const uint32_t *a) {
    constexpr int O = 3 * (K * J + K / 2) + C;
    return __byte_perm(a[O >> 2], 0x4B000000u, 0x7650 + (O & 3));
}
template <int K>
__device__ __forceinline__ void unit_point(const uint8_t *s0, __half *const *pc) {
    constexpr int NW = 6 * K;
    uint32_t a[NW];
#pragma unroll
    for (int i = 0; i < NW; i += 2) {  // the window starts on 8 bytes (24 K per group)
        const uint2 v = lds64(s0 + 4 * i);
        a[i] = v.x;
        a[i + 1] = v.y;
    }
#define FDMS_POINT_CH(C)                                                                                             \
    {                                                                                                                \
        unsigned t[8] = {point_col<K, C, 0>(a), point_col<K, C, 1>(a), point_col<K, C, 2>(a), point_col<K, C, 3>(a),  \
                         point_col<K, C, 4>(a), point_col<K, C, 5>(a), point_col<K, C, 6>(a), point_col<K, C, 7>(a)}; \
        store8(pc[C], t);                                                                                            \
    }
    FDMS_POINT_CH(0)
    FDMS_POINT_CH(1)
    FDMS_POINT_CH(2)
#undef FDMS_POINT_CH
}

__device__ __forceinline__ void strip_body(const StripParams &P) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ StripShared sh;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NS = P.n_stages;
    const int n_consumers = (int)(blockDim.x >> 5) - 1;

    if (tid == 0) {
        for (int s = 0; s < NS; s++) {
            mbar_init(&sh.full[s], 1);
            mbar_init(&sh.empty[s], (uint32_t)n_consumers);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == n_consumers) {
        // ---------------- producer warp ----------------
        for (int it = 0;; it++) {
            const int s = it % NS, use = it / NS;
            if (use > 0) mbar_wait(&sh.empty[s], (use - 1) & 1);  // every consumer warp has released the stage
            int item = 0;
            if (lane == 0) item = (int)atomicAdd(P.counters, 1u);
            item = __shfl_sync(0xffffffffu, item, 0);
            if (item >= P.n_items) {
                if (lane == 0) {
                    sh.item[s] = -1;
                    mbar_arrive(&sh.full[s]);
                }
                break;
            }
            const bool strip = item >= P.n_border_items;
            int img, py = 0, cx = 0;
            if (strip) {
                const int q = item - P.n_border_items;
                cx = q % P.chunks;
                const int r = q / P.chunks;
                py = r % P.periods;
                img = r / P.periods;
            } else {
                img = item / P.n_border_scales;
            }
            if (lane < P.n_scales) {
                const StripScale &S = P.sc[lane];
                __half *base = S.out + (long long)img * S.img_stride +
                               (strip ? (long long)(S.top + py * S.rows_per_period) * S.dst_w + cx * S.out_chunk_cols : 0LL);
                sh.out_ptr[s][lane][0] = base + S.plane[0];
                sh.out_ptr[s][lane][1] = base + S.plane[1];
                sh.out_ptr[s][lane][2] = base + S.plane[2];
            }
            __syncwarp();
            if (lane == 0) {
                sh.item[s] = item;
                sh.next_wu[s] = 0;
                if (strip) mbar_expect_tx(&sh.full[s], (uint32_t)(P.n_slots * P.row_bytes));
                else mbar_arrive(&sh.full[s]);
            }
            __syncwarp();
            if (strip && lane < P.n_slots) {
                const uint8_t *g = P.src + (long long)img * P.img_stride + (long long)(py * P.L + P.slot_row[lane]) * P.pitch +
                                   (long long)cx * P.row_bytes;
                bulk_g2s(smem + (size_t)s * P.stage_bytes + (size_t)lane * P.row_bytes, g, (uint32_t)P.row_bytes, &sh.full[s]);
            }
        }
        if (lane == 0) {  // the last CTA to finish rearms the counters for the next launch
            __threadfence();
            const unsigned done = atomicAdd(P.counters + 1, 1u);
            if (done == gridDim.x - 1) {
                P.counters[0] = 0u;
                P.counters[1] = 0u;
            }
        }
        return;
    }

    // ---------------- consumer warps ----------------
    const unsigned border_bits = 0x4B000000u + (unsigned)P.border;
    for (int it = 0;; it++) {
        const int s = it % NS;
        mbar_wait(&sh.full[s], (it / NS) & 1);
        const int item = sh.item[s];
        if (item < 0) break;
        if (item < P.n_border_items) {
            // letterbox rows above and below the image of one (image, scale): constant rows, store only
            const int bsi = P.border_scale[item % P.n_border_scales];
            const StripScale &S = P.sc[bsi];
            __half *const base = sh.out_ptr[s][bsi][P.flip ? 2 : 0];  // plane 0
            const int gpr = S.dst_w >> 3, n_rows = S.top + S.bottom;
            const unsigned v = pack2(border_bits, border_bits);
            const uint4 vv = make_uint4(v, v, v, v);
            for (;;) {
                int wu = 0;
                if (lane == 0) wu = atomicAdd(&sh.next_wu[s], 1);
                wu = __shfl_sync(0xffffffffu, wu, 0);
                if (wu >= S.border_wu) break;
                const int u = wu * 32 + lane;
                if (u < n_rows * gpr) {
                    const int rr = u / gpr, g = u - rr * gpr;
                    const int y = rr < S.top ? rr : S.dst_h - (n_rows - rr);
                    __half *p = base + y * S.dst_w + 8 * g;
                    *reinterpret_cast<uint4 *>(p) = vv;
                    *reinterpret_cast<uint4 *>(p + S.cs) = vv;
                    *reinterpret_cast<uint4 *>(p + 2 * S.cs) = vv;
                }
            }
        } else {
            const uint8_t *stage = smem + (size_t)s * P.stage_bytes;
            // the next unit number is taken while the current unit runs (the shared-memory atomic's latency is hidden)
            int wu = 0;
            if (lane == 0) wu = atomicAdd(&sh.next_wu[s], 1);
            wu = __shfl_sync(0xffffffffu, wu, 0);
            while (wu < P.wu_per_strip) {
                int nxt = 0;
                if (lane == 0) nxt = atomicAdd(&sh.next_wu[s], 1);
                const unsigned info = P.wu_info[wu];
                const int si = (int)(info & 0xffu), ru = (int)((info >> 8) & 0xffu), g = (int)(info >> 16) + lane;
                const StripScale &S = P.sc[si];
                if (g < S.groups) {
                    const int eo = ru * S.ru_stride + 8 * g;
                    __half *pc[3] = {sh.out_ptr[s][si][0] + eo, sh.out_ptr[s][si][1] + eo, sh.out_ptr[s][si][2] + eo};
                    const uint8_t *s0 = stage + (S.slot0[ru] + g * S.group_bytes);
                    if (S.type == ST_R32) unit_r32(s0, P.row_bytes, pc, S.dst_w);
                    else if (S.type == ST_HALF) {
                        if (S.k == 2) unit_half<2>(s0, P.row_bytes, pc);
                        else if (S.k == 4) unit_half<4>(s0, P.row_bytes, pc);
                        else unit_half<6>(s0, P.row_bytes, pc);
                    } else {
                        if (S.k == 1) unit_point<1>(s0, pc);
                        else if (S.k == 3) unit_point<3>(s0, pc);
                        else unit_point<5>(s0, pc);
                    }
                }
                wu = __shfl_sync(0xffffffffu, nxt, 0);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.empty[s]);
    }
}

// Two register allocations of the same body: 64 registers (the fastest alone) and 56 (a few spilled bytes, but
// 17 warps then leave half of the SM's register file to the 1024-thread NMS / merge CTAs of the other stream, which
// can co-reside instead of queueing behind a persistent kernel).  FDMS_PYR_STRIP_REGS picks.
__global__ void __launch_bounds__(ST_MAX_THREADS, 1) pyramid_strip_kernel(const __grid_constant__ StripParams P) { strip_body(P); }
__global__ void __maxnreg__(56) pyramid_strip_kernel_r56(const __grid_constant__ StripParams P) { strip_body(P); }

int gcd_i(int a, int b) { return b ? gcd_i(b, a % b) : a; }
int lcm_i(int a, int b) { return a / gcd_i(a, b) * b; }

}  // namespace

int fdms_pyramid_strip(fdms_ctx *ctx, const PyrParams &P, int n_img, unsigned *handled, cudaStream_t st) {
    *handled = 0u;
    if (!ctx->pyr_strip || !P.vec_ok || P.virt_h != P.src_h || P.virt_w != P.src_w || ((3LL * P.src_w) & 15)) return FDMS_OK;
    StripParams Q;
    memset(&Q, 0, sizeof(Q));
    int L = 1, col_unit = 16;  // source columns: a strip starts on a 16-byte boundary (16 px = 48 B)
    int idx[FDMS_MAX_SCALES], n = 0;
    for (int s = 0; s < P.n_scales; s++) {
        const PyrScale &S = P.sc[s];
        const bool plane_ok = S.xs == 1 && S.ys == S.dst_w && (S.dst_w & 7) == 0 && (S.cs & 7) == 0 && (S.img_stride & 7) == 0 &&
                              ((uintptr_t)S.out & 15) == 0 && S.left == 0 && S.unpad_w == S.dst_w;
        if (!S.direct || !plane_ok) continue;
        int type, k, period, span;
        if (S.mode == PM_RATIO32) {
            type = ST_R32; k = 0; period = 3; span = 12;
            if (2 * P.src_h != 3 * S.unpad_h || 2 * P.src_w != 3 * S.unpad_w) continue;
        } else {
            k = S.ky;
            if (S.kx != S.ky || k < 1 || k > 6 || k * S.unpad_h != P.src_h || k * S.unpad_w != P.src_w) continue;
            if (S.mode == PM_POINT && (k & 1)) type = ST_POINT;
            else if (S.mode == PM_HALF && !(k & 1)) type = ST_HALF;
            else continue;
            period = k; span = 8 * k;
        }
        if (lcm_i(L, period) > 12) continue;
        L = lcm_i(L, period);
        col_unit = lcm_i(col_unit, span);
        StripScale &T = Q.sc[n];
        T.out = reinterpret_cast<__half *>(S.out);
        T.img_stride = S.img_stride;
        T.cs = S.cs;
        T.type = type;
        T.k = k;
        T.dst_w = S.dst_w;
        T.dst_h = S.dst_h;
        T.top = S.top;
        T.bottom = S.dst_h - S.top - S.unpad_h;
        idx[n++] = s;
    }
    // a launch the strip kernel covers only partly would read the source twice (strip kernel + tile kernel), which costs
    // more than it saves (measured on 4K -> 640/960/1280/1536): all scales or none (FDMS_PYR_STRIP=2 allows the mix: tests)
    if (n == 0 || (n != P.n_scales && ctx->pyr_strip < 2) || P.src_h % L || P.src_w % col_unit) return FDMS_OK;
    // source rows of a period that any scale reads -> shared-memory slots
    auto needed = [&](int r) {
        for (int q = 0; q < n; q++) {
            const StripScale &T = Q.sc[q];
            if (T.type == ST_R32 || (T.type == ST_HALF && (r % T.k == T.k / 2 - 1 || r % T.k == T.k / 2)) ||
                (T.type == ST_POINT && r % T.k == T.k / 2))
                return true;
        }
        return false;
    };
    int slots_per_period = 0, ru_max = 1;
    for (int r = 0; r < L; r++) slots_per_period += needed(r) ? 1 : 0;
    for (int q = 0; q < n; q++) ru_max = ru_max > (Q.sc[q].type == ST_R32 ? L / 3 : L / Q.sc[q].k) ? ru_max : (Q.sc[q].type == ST_R32 ? L / 3 : L / Q.sc[q].k);
    // strip = NP periods x CW columns: the widest strip, then the most periods, that keep a stage within the budget
    int cw = 0;
    for (int c = col_unit; c <= P.src_w; c += col_unit)
        if (P.src_w % c == 0 && (long long)slots_per_period * c * 3 <= ctx->pyr_strip_stage_bytes) cw = c;
    if (!cw) return FDMS_OK;
    int np = 1;
    for (int c = 2; c * L <= ST_MAX_L; c++)
        if ((P.src_h / L) % c == 0 && c * slots_per_period <= ST_MAX_SLOTS && c * ru_max <= ST_MAX_RU &&
            (long long)c * slots_per_period * cw * 3 <= ctx->pyr_strip_stage_bytes)
            np = c;
    L *= np;
    int slot_of[ST_MAX_L];
    for (int r = 0; r < L; r++) {
        slot_of[r] = Q.n_slots;
        if (needed(r)) Q.slot_row[Q.n_slots++] = r;
    }
    Q.row_bytes = cw * 3;
    Q.stage_bytes = Q.n_slots * Q.row_bytes;
    Q.n_stages = (int)(ctx->pyr_strip_smem / Q.stage_bytes);
    if (Q.n_stages > ST_MAX_STAGES) Q.n_stages = ST_MAX_STAGES;
    if (Q.n_stages < 2) return FDMS_OK;
    Q.src = P.src;
    Q.img_stride = P.img_stride;
    Q.pitch = P.pitch;
    Q.n_img = n_img;
    Q.L = L;
    Q.periods = P.src_h / L;
    Q.chunks = P.src_w / cw;
    Q.n_scales = n;
    Q.flip = P.flip;
    Q.border = P.border;
    // units: the expensive scales first, so that the dynamic hand-out ends on cheap ones
    int order[FDMS_MAX_SCALES];
    for (int q = 0; q < n; q++) order[q] = q;
    auto cost = [&](const StripScale &T) { return T.type == ST_R32 ? 3 : (T.type == ST_HALF ? 2 : 1); };
    for (int a = 0; a < n; a++)
        for (int b = a + 1; b < n; b++)
            if (cost(Q.sc[order[b]]) > cost(Q.sc[order[a]])) { int t = order[a]; order[a] = order[b]; order[b] = t; }
    StripScale tmp[FDMS_MAX_SCALES];
    int idx2[FDMS_MAX_SCALES];
    for (int q = 0; q < n; q++) { tmp[q] = Q.sc[order[q]]; idx2[q] = idx[order[q]]; }
    int wu = 0;
    for (int q = 0; q < n; q++) {
        StripScale &T = Q.sc[q];
        T = tmp[q];
        if (T.type == ST_R32) {
            T.n_ru = L / 3;
            T.rows_per_period = 2 * L / 3;
            T.out_chunk_cols = 2 * cw / 3;
            T.group_bytes = 36;
            for (int u = 0; u < T.n_ru; u++) T.slot0[u] = slot_of[3 * u] * cw * 3;
        } else {
            T.n_ru = L / T.k;
            T.rows_per_period = L / T.k;
            T.out_chunk_cols = cw / T.k;
            T.group_bytes = 24 * T.k;
            for (int u = 0; u < T.n_ru; u++) T.slot0[u] = slot_of[T.k * u + (T.type == ST_HALF ? T.k / 2 - 1 : T.k / 2)] * cw * 3;
        }
        T.groups = T.out_chunk_cols / 8;
        T.n_units = T.n_ru * T.groups;
        T.ru_stride = T.dst_w * (T.type == ST_R32 ? 2 : 1);
        T.plane[0] = P.flip ? 2 * T.cs : 0;
        T.plane[1] = T.cs;
        T.plane[2] = P.flip ? 0 : 2 * T.cs;
        T.wu_begin = wu;
        for (int u = 0; u < T.n_ru; u++)
            for (int g0 = 0; g0 < T.groups; g0 += 32) {
                if (wu >= ST_MAX_WU || T.groups >= 65536) { *handled = 0u; return FDMS_OK; }
                Q.wu_info[wu++] = (unsigned)q | ((unsigned)u << 8) | ((unsigned)g0 << 16);
            }
        T.border_wu = ceil_div((T.top + T.bottom) * (T.dst_w / 8), 32);
        if (T.top + T.bottom > 0) Q.border_scale[Q.n_border_scales++] = q;
        *handled |= 1u << idx2[q];
    }
    Q.wu_per_strip = wu;
    Q.n_border_items = n_img * Q.n_border_scales;
    const long long items = (long long)Q.n_border_items + (long long)n_img * Q.periods * Q.chunks;
    if (items > 0x7fffffff) { *handled = 0u; return FDMS_OK; }
    Q.n_items = (int)items;
    if (int rc = fdms_reserve(ctx, ctx->pyr_cnt, 256)) return rc;
    if (!ctx->pyr_cnt_ready) {
        FDMS_CUDA(ctx, cudaMemsetAsync(ctx->pyr_cnt.p, 0, 256, st));
        ctx->pyr_cnt_ready = true;
    }
    Q.counters = reinterpret_cast<unsigned *>(ctx->pyr_cnt.p);
    int smem = Q.n_stages * Q.stage_bytes;
    if (ctx->pyr_strip_reserve > smem) smem = ctx->pyr_strip_reserve > 200 * 1024 ? 200 * 1024 : ctx->pyr_strip_reserve;
    if (!ctx->pyr_strip_attr) {
        FDMS_CUDA(ctx, cudaFuncSetAttribute(pyramid_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        FDMS_CUDA(ctx, cudaFuncSetAttribute(pyramid_strip_kernel_r56, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        ctx->pyr_strip_attr = true;
    }
    int grid = ctx->pyr_strip_ctas > 0 && ctx->pyr_strip_ctas < ctx->sm_count ? ctx->pyr_strip_ctas : ctx->sm_count;
    if (grid > Q.n_items) grid = Q.n_items;
    int warps = ctx->pyr_strip_warps < 1 ? 1 : (ctx->pyr_strip_warps > ST_MAX_CONSUMER_WARPS ? ST_MAX_CONSUMER_WARPS : ctx->pyr_strip_warps);
    if (ctx->pyr_strip_regs <= 56) pyramid_strip_kernel_r56<<<grid, 32 * (warps + 1), smem, st>>>(Q);
    else pyramid_strip_kernel<<<grid, 32 * (warps + 1), smem, st>>>(Q);
    FDMS_LAUNCH_CHECK(ctx, "pyramid_strip_kernel");
    return FDMS_OK;
}
