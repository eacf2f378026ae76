// Device JPEG ingest: what follows the entropy decode of cv2.imread's JPEG path
// (multi_scale_face_detector.py:84 -> libjpeg-turbo), bit for bit, writing BGR straight into the pyramid's source tensor.
//
//   host  (jpeg_host.cu) : marker parser + Huffman decode, one image per thread -> quantised DCT coefficients
//   idct  (this file)    : dequantisation + the "islow" integer IDCT of jidctint.c (CONST_BITS 13, PASS1_BITS 2), range
//                          limit -> component sample planes
//   colour (this file)   : "fancy" (triangle filter) chroma upsampling of jdsample.c -- h2v2: (3*near + far) per
//                          axis, +8 / +7 rounding alternating by column; h2v1: (3*near + far + 1|2) >> 2 -- with libjpeg's
//                          edge handling (the row above the first / below the last real sample row is that row itself,
//                          the columns likewise), then jdcolor.c's YCbCr -> RGB with 16-bit fixed-point tables, BGR out.
// Pinned against cv2.imread itself: tests/test_jpeg_ingest.py (CPU: the same arithmetic restated in numpy on the host
// decoder's coefficients) and tests/test_gpu_jpeg.py (these kernels), array_equal on 4:4:4 / 4:2:2 / 4:2:0 / grey
// streams, odd sizes, restart intervals, optimised tables, and the reference's own photographs.
#include <thread>
#include <vector>

#include "common.cuh"
#include "jpeg_host.cuh"

extern "C" {

// Host-only probe of a JPEG byte stream (no context, no device): geometry of the coefficient buffer the device path uses.
FDMS_API int fdms_jpeg_probe(const uint8_t *data, size_t len, fdms_jpeg_desc *out) {
    if (!data || !out) return FDMS_ERR_INVALID;
    JpegInfo info;
    memset(out, 0, sizeof(*out));
    if (fdms_jpeg_parse(data, len, &info) != 0) return FDMS_ERR_INVALID;
    out->width = info.width;
    out->height = info.height;
    out->ncomp = info.ncomp;
    out->supported = info.supported;
    strncpy(out->reason, info.reason ? info.reason : "", sizeof(out->reason) - 1);
    if (!info.supported) return FDMS_OK;
    for (int c = 0; c < info.ncomp; c++) {
        out->h[c] = info.comp[c].h;
        out->v[c] = info.comp[c].v;
        out->blocks_w[c] = info.comp[c].blocks_w;
        out->blocks_h[c] = info.comp[c].blocks_h;
        out->down_w[c] = info.comp[c].down_w;
        out->down_h[c] = info.comp[c].down_h;
        out->coef_off[c] = info.comp[c].coef_off;
        memcpy(out->qt[c], info.qt[info.comp[c].tq], sizeof(out->qt[c]));
    }
    out->coef_count = info.coef_count;
    return FDMS_OK;
}

// Host-only entropy decode into `coef` (fdms_jpeg_desc.coef_count int16): test / tooling entry point of the host half.
FDMS_API int fdms_jpeg_decode_host(const uint8_t *data, size_t len, int16_t *coef) {
    if (!data || !coef) return FDMS_ERR_INVALID;
    JpegInfo info;
    if (fdms_jpeg_parse(data, len, &info) != 0 || !info.supported) return FDMS_ERR_UNSUPPORTED;
    return fdms_jpeg_decode_coefficients(data, len, &info, coef) == 0 ? FDMS_OK : FDMS_ERR_INVALID;
}

}  // extern "C"

namespace {

struct JpegDevImage {
    int ncomp, hs, vs;           // components; luma sampling factors (1x1, 2x1, 2x2)
    int blocks_w[3], blocks_h[3], down_w[3], down_h[3];
    int block_begin[4];          // prefix of blocks over the components
    long long coef_off[3];       // int16 offset of a component's blocks in the batch's coefficient buffer
    long long plane_off[3];      // byte offset of a component's sample plane in the batch's plane buffer
    int plane_stride[3];
    int skip;                    // 1: not decoded here (unsupported / corrupt): the caller fills the image
    uint16_t qt[3][64];
};

constexpr int IDCT_THREADS = 256;                 // 8 threads per block -> 32 blocks per CTA
constexpr int IDCT_BLOCKS = IDCT_THREADS / 8;
constexpr int WS_STRIDE = 72;                     // ints per block in shared memory (64 + 8: the 4 blocks of a warp on distinct banks)

// jidctint.c (islow): CONST_BITS 13, PASS1_BITS 2
#define FIX_0_298631336 2446
#define FIX_0_390180644 3196
#define FIX_0_541196100 4433
#define FIX_0_765366865 6270
#define FIX_0_899976223 7373
#define FIX_1_175875602 9633
#define FIX_1_501321110 12299
#define FIX_1_847759065 15137
#define FIX_1_961570560 16069
#define FIX_2_053119869 16819
#define FIX_2_562915447 20995
#define FIX_3_072711026 25172

// One 8-point pass of the islow IDCT: x[0..7] in, o[0..7] = descaled outputs (shift 11 after the column pass, 18 after the rows).
template <int SHIFT>
__device__ __forceinline__ void idct8(const int *x, int *o) {
    int z2 = x[2], z3 = x[6];
    int z1 = (z2 + z3) * FIX_0_541196100;
    int tmp2 = z1 + z3 * (-FIX_1_847759065);
    int tmp3 = z1 + z2 * FIX_0_765366865;
    int tmp0 = (x[0] + x[4]) << 13;
    int tmp1 = (x[0] - x[4]) << 13;
    const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = x[7];
    tmp1 = x[5];
    tmp2 = x[3];
    tmp3 = x[1];
    z1 = tmp0 + tmp3;
    z2 = tmp1 + tmp2;
    z3 = tmp0 + tmp2;
    int z4 = tmp1 + tmp3;
    const int z5 = (z3 + z4) * FIX_1_175875602;
    tmp0 *= FIX_0_298631336;
    tmp1 *= FIX_2_053119869;
    tmp2 *= FIX_3_072711026;
    tmp3 *= FIX_1_501321110;
    z1 *= -FIX_0_899976223;
    z2 *= -FIX_2_562915447;
    z3 *= -FIX_1_961570560;
    z4 *= -FIX_0_390180644;
    z3 += z5;
    z4 += z5;
    tmp0 += z1 + z3;
    tmp1 += z2 + z4;
    tmp2 += z2 + z3;
    tmp3 += z1 + z4;
    constexpr int R = 1 << (SHIFT - 1);
    o[0] = (tmp10 + tmp3 + R) >> SHIFT;
    o[7] = (tmp10 - tmp3 + R) >> SHIFT;
    o[1] = (tmp11 + tmp2 + R) >> SHIFT;
    o[6] = (tmp11 - tmp2 + R) >> SHIFT;
    o[2] = (tmp12 + tmp1 + R) >> SHIFT;
    o[5] = (tmp12 - tmp1 + R) >> SHIFT;
    o[3] = (tmp13 + tmp0 + R) >> SHIFT;
    o[4] = (tmp13 - tmp0 + R) >> SHIFT;
}

// libjpeg's post-IDCT range limit: table lookup at (x & RANGE_MASK) with the table centred on 128
__device__ __forceinline__ unsigned range_limit(int x) {
    const int i = x & 1023;
    return i < 128 ? (unsigned)(i + 128) : (i < 512 ? 255u : (i < 896 ? 0u : (unsigned)(i - 896)));
}

// Dequantisation + islow IDCT of every block of every component of every image: 8 threads per block (a column each, then a
// row each), the 8 x 8 intermediate through shared memory.  grid = (block groups of the largest image, image).
__global__ void __launch_bounds__(IDCT_THREADS) jpeg_idct_kernel(const JpegDevImage *__restrict__ imgs, const int16_t *__restrict__ coef,
                                                                 uint8_t *__restrict__ planes) {
    __shared__ int ws[IDCT_BLOCKS * WS_STRIDE];
    const JpegDevImage &I = imgs[blockIdx.y];
    if (I.skip) return;
    const int lb = threadIdx.x >> 3, t = threadIdx.x & 7;
    const int b = blockIdx.x * IDCT_BLOCKS + lb;
    const bool live = b < I.block_begin[I.ncomp];
    int c = 0;
    if (live)
        for (int k = 1; k < I.ncomp; k++)
            if (b >= I.block_begin[k]) c = k;
    int *w = ws + lb * WS_STRIDE;
    int bx = 0, by = 0;
    if (live) {
        const int bl = b - I.block_begin[c];
        by = bl / I.blocks_w[c];
        bx = bl - by * I.blocks_w[c];
        const int16_t *src = coef + I.coef_off[c] + (long long)bl * 64;
        int x[8], o[8];
#pragma unroll
        for (int r = 0; r < 8; r++) x[r] = (int)src[r * 8 + t] * (int)I.qt[c][r * 8 + t];  // column t
        idct8<11>(x, o);
#pragma unroll
        for (int r = 0; r < 8; r++) w[r * 8 + t] = o[r];
    }
    __syncwarp();
    if (live) {
        int x[8], o[8];
#pragma unroll
        for (int k = 0; k < 8; k++) x[k] = w[t * 8 + k];  // row t
        idct8<18>(x, o);
        uint2 v;
        v.x = range_limit(o[0]) | (range_limit(o[1]) << 8) | (range_limit(o[2]) << 16) | (range_limit(o[3]) << 24);
        v.y = range_limit(o[4]) | (range_limit(o[5]) << 8) | (range_limit(o[6]) << 16) | (range_limit(o[7]) << 24);
        uint8_t *dst = planes + I.plane_off[c] + (long long)(by * 8 + t) * I.plane_stride[c] + bx * 8;
        *reinterpret_cast<uint2 *>(dst) = v;
    }
}

// jdcolor.c ycc_rgb_convert: 16-bit fixed-point tables, FIX(x) = (int)(x * 65536 + 0.5)
__device__ __forceinline__ void ycc_to_bgr(int y, int cb, int cr, uint8_t *out) {
    cb -= 128;
    cr -= 128;
    const int r = y + ((91881 * cr + 32768) >> 16);                    // FIX(1.40200)
    const int b = y + ((116130 * cb + 32768) >> 16);                   // FIX(1.77200)
    const int g = y + ((-22554 * cb + 32768 + (-46802) * cr) >> 16);   // -FIX(0.34414), -FIX(0.71414)
    out[0] = (uint8_t)min(max(b, 0), 255);
    out[1] = (uint8_t)min(max(g, 0), 255);
    out[2] = (uint8_t)min(max(r, 0), 255);
}

// Fancy upsampling (jdsample.c) of one chroma sample at output position (X, Y) + colour conversion, one pixel per thread.
// Edges: libjpeg duplicates the first / last REAL sample row as the context row and special-cases the first / last
// column, which is the same as clamping the far tap's index.
__global__ void __launch_bounds__(256) jpeg_color_kernel(const JpegDevImage *__restrict__ imgs, const uint8_t *__restrict__ planes,
                                                         uint8_t *__restrict__ dst, long long img_stride, long long pitch, int h, int w) {
    const JpegDevImage &I = imgs[blockIdx.z];
    if (I.skip) return;
    const int X = blockIdx.x * blockDim.x + threadIdx.x, Y = blockIdx.y;
    if (X >= w || Y >= h) return;
    uint8_t *out = dst + (long long)blockIdx.z * img_stride + (long long)Y * pitch + 3LL * X;
    const int yv = planes[I.plane_off[0] + (long long)Y * I.plane_stride[0] + X];
    if (I.ncomp == 1) {
        out[0] = out[1] = out[2] = (uint8_t)yv;
        return;
    }
    int cbv, crv;
    const uint8_t *pb = planes + I.plane_off[1], *pr = planes + I.plane_off[2];
    const int sb = I.plane_stride[1], sr = I.plane_stride[2];
    if (I.hs == 1) {  // 4:4:4
        cbv = pb[(long long)Y * sb + X];
        crv = pr[(long long)Y * sr + X];
    } else {
        const int dw = I.down_w[1], dh = I.down_h[1];
        const int c0 = X >> 1, c1 = min(max(c0 + ((X & 1) ? 1 : -1), 0), dw - 1);
        if (I.vs == 1) {  // 4:2:2, h2v1: (3 * near + far + 1 | 2) >> 2
            const int rnd = (X & 1) ? 2 : 1;
            cbv = (3 * pb[(long long)Y * sb + c0] + pb[(long long)Y * sb + c1] + rnd) >> 2;
            crv = (3 * pr[(long long)Y * sr + c0] + pr[(long long)Y * sr + c1] + rnd) >> 2;
        } else {  // 4:2:0, h2v2: columns of (3 * near row + far row), then (3 * near + far + 8 | 7) >> 4
            const int r0 = Y >> 1, r1 = min(max(r0 + ((Y & 1) ? 1 : -1), 0), dh - 1);
            const int rnd = (X & 1) ? 7 : 8;
            const int tb = 3 * pb[(long long)r0 * sb + c0] + pb[(long long)r1 * sb + c0];
            const int ob = 3 * pb[(long long)r0 * sb + c1] + pb[(long long)r1 * sb + c1];
            const int tr = 3 * pr[(long long)r0 * sr + c0] + pr[(long long)r1 * sr + c0];
            const int orr = 3 * pr[(long long)r0 * sr + c1] + pr[(long long)r1 * sr + c1];
            cbv = (3 * tb + ob + rnd) >> 4;
            crv = (3 * tr + orr + rnd) >> 4;
        }
    }
    ycc_to_bgr(yv, cbv, crv, out);
}

}  // namespace

extern "C" FDMS_API int fdms_jpeg_decode_batch(fdms_ctx *ctx, const uint8_t *const *data, const size_t *len, int n, int n_threads,
                                               uint8_t *dst, int64_t img_stride, int64_t pitch, int h, int w, int32_t *status,
                                               fdms_stream stream) {
    FDMS_BIND(ctx);
    if (!data || !len || !dst || !status || n <= 0 || n > 65535 || h <= 0 || w <= 0 || pitch < 3LL * w || img_stride < pitch * h)
        return fdms_fail(ctx, FDMS_ERR_INVALID, "fdms_jpeg_decode_batch: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    // ---- parse (cheap, serial), lay the batch out
    std::vector<JpegInfo> info((size_t)n);
    std::vector<JpegDevImage> dev((size_t)n);
    long long coef_total = 0, plane_total = 0;
    int max_blocks = 0, n_dev = 0;
    for (int i = 0; i < n; i++) {
        JpegDevImage &D = dev[(size_t)i];
        memset(&D, 0, sizeof(D));
        D.skip = 1;
        status[i] = 1;
        if (!data[i] || fdms_jpeg_parse(data[i], len[i], &info[(size_t)i]) != 0) {
            status[i] = FDMS_ERR_INVALID;
            continue;
        }
        const JpegInfo &J = info[(size_t)i];
        if (!J.supported) continue;  // status 1: the caller decodes this one on the host
        if (J.width != w || J.height != h) {
            status[i] = FDMS_ERR_INVALID;
            continue;
        }
        D.ncomp = J.ncomp;
        D.hs = J.hmax;
        D.vs = J.vmax;
        int blocks = 0;
        for (int c = 0; c < J.ncomp; c++) {
            const JpegComp &k = J.comp[c];
            D.blocks_w[c] = k.blocks_w;
            D.blocks_h[c] = k.blocks_h;
            D.down_w[c] = k.down_w;
            D.down_h[c] = k.down_h;
            D.block_begin[c] = blocks;
            blocks += k.blocks_w * k.blocks_h;
            D.coef_off[c] = coef_total + k.coef_off;
            D.plane_off[c] = plane_total;
            D.plane_stride[c] = k.blocks_w * 8;
            plane_total += (long long)k.blocks_w * 8 * k.blocks_h * 8;
            memcpy(D.qt[c], J.qt[k.tq], sizeof(D.qt[c]));
        }
        for (int c = J.ncomp; c <= 3; c++) D.block_begin[c] = blocks;
        coef_total += J.coef_count;
        if (blocks > max_blocks) max_blocks = blocks;
        D.skip = 0;
        status[i] = 0;
        n_dev++;
    }
    if (n_dev == 0) return FDMS_OK;
    // ---- buffers: pinned staging (coefficients + descriptors), device coefficients / planes / descriptors
    const size_t coef_bytes = (size_t)coef_total * 2, desc_bytes = (size_t)n * sizeof(JpegDevImage);
    if (ctx->jpg_pending) {  // the previous batch's upload still reads the staging buffer
        FDMS_CUDA(ctx, cudaEventSynchronize(ctx->jpg_event));
        ctx->jpg_pending = false;
    }
    if (coef_bytes + desc_bytes > ctx->jpg_host_bytes) {
        if (ctx->jpg_host) cudaFreeHost(ctx->jpg_host);
        ctx->jpg_host = nullptr;
        ctx->jpg_host_bytes = 0;
        const size_t want = (coef_bytes + desc_bytes) + (coef_bytes + desc_bytes) / 8 + 4096;
        cudaError_t e = cudaHostAlloc(&ctx->jpg_host, want, cudaHostAllocDefault);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fdms_fail(ctx, FDMS_ERR_NOMEM, "cudaHostAlloc(%zu) failed: %s", want, cudaGetErrorString(e));
        }
        ctx->jpg_host_bytes = want;
    }
    if (!ctx->jpg_event) FDMS_CUDA(ctx, cudaEventCreateWithFlags(&ctx->jpg_event, cudaEventDisableTiming));
    if (int rc = fdms_reserve(ctx, ctx->jpg_coef, coef_bytes + desc_bytes)) return rc;
    if (int rc = fdms_reserve(ctx, ctx->jpg_planes, (size_t)plane_total)) return rc;
    int16_t *h_coef = reinterpret_cast<int16_t *>(ctx->jpg_host);
    // ---- entropy decode on host threads, one image at a time per thread
    {
        std::vector<int> todo;
        for (int i = 0; i < n; i++)
            if (!dev[(size_t)i].skip) todo.push_back(i);
        int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
        if (nt < 1) nt = 1;
        if (nt > (int)todo.size()) nt = (int)todo.size();
        auto work = [&](int tid) {
            for (size_t k = (size_t)tid; k < todo.size(); k += (size_t)nt) {
                const int i = todo[k];
                const JpegInfo &J = info[(size_t)i];
                int16_t *out = h_coef + (dev[(size_t)i].coef_off[0] - J.comp[0].coef_off);
                if (fdms_jpeg_decode_coefficients(data[i], len[i], &J, out) != 0) {
                    status[i] = FDMS_ERR_INVALID;  // corrupt scan: the caller decides (cv2 may still salvage it)
                    dev[(size_t)i].skip = 1;
                }
            }
        };
        if (nt == 1) {
            work(0);
        } else {
            std::vector<std::thread> pool;
            for (int t = 0; t < nt; t++) pool.emplace_back(work, t);
            for (auto &t : pool) t.join();
        }
    }
    memcpy(reinterpret_cast<uint8_t *>(ctx->jpg_host) + coef_bytes, dev.data(), desc_bytes);
    FDMS_CUDA(ctx, cudaMemcpyAsync(ctx->jpg_coef.p, ctx->jpg_host, coef_bytes + desc_bytes, cudaMemcpyHostToDevice, st));
    FDMS_CUDA(ctx, cudaEventRecord(ctx->jpg_event, st));
    ctx->jpg_pending = true;
    const int16_t *d_coef = reinterpret_cast<const int16_t *>(ctx->jpg_coef.p);
    const JpegDevImage *d_desc = reinterpret_cast<const JpegDevImage *>(reinterpret_cast<const uint8_t *>(ctx->jpg_coef.p) + coef_bytes);
    jpeg_idct_kernel<<<dim3(ceil_div(max_blocks, IDCT_BLOCKS), n), IDCT_THREADS, 0, st>>>(d_desc, d_coef,
                                                                                         reinterpret_cast<uint8_t *>(ctx->jpg_planes.p));
    FDMS_LAUNCH_CHECK(ctx, "jpeg_idct_kernel");
    jpeg_color_kernel<<<dim3(ceil_div(w, 256), h, n), 256, 0, st>>>(d_desc, reinterpret_cast<const uint8_t *>(ctx->jpg_planes.p), dst,
                                                                   img_stride, pitch, h, w);
    FDMS_LAUNCH_CHECK(ctx, "jpeg_color_kernel");
    return FDMS_OK;
}
