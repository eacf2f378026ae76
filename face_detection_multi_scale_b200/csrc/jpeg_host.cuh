// Host half of the device JPEG ingest: marker parser + baseline Huffman decoder.
//
// The reference decodes every file on the host, once per scale (cv2.imread, multi_scale_face_detector.py:84 ->
// libjpeg-turbo: Huffman decode, dequantisation + islow IDCT, "fancy" chroma upsampling, YCbCr -> BGR).  Only the entropy
// decode is inherently serial; here it stays on the host (one image per thread) and hands the quantised DCT
// coefficients to the device, where jpeg.cu does the rest bit for bit like libjpeg-turbo.  Supported: baseline /
// extended sequential 8-bit JPEG (SOF0 / SOF1), one interleaved scan, 1 or 3 components (YCbCr), luma sampling 1x1 /
// 2x1 / 2x2 with 1x1 chroma, restart intervals, no EXIF rotation.  Anything else is reported as unsupported and the caller
// falls back to cv2.imread.
#pragma once
#include <stddef.h>
#include <stdint.h>

struct JpegComp {
    int id, h, v, tq, td, ta;
    int blocks_w, blocks_h;      // whole blocks the scan holds for this component (MCU-padded)
    int down_w, down_h;          // real samples: ceil(image * samp / max_samp)
    long long coef_off;          // offset (in int16) of this component's blocks inside the coefficient buffer
};

struct JpegInfo {
    int width, height, ncomp;
    int hmax, vmax, mcus_x, mcus_y;
    int restart_interval;
    int supported;               // 0: fall back to the host decoder (reason below)
    const char *reason;
    JpegComp comp[3];
    uint16_t qt[4][64];          // quantisation tables in NATURAL order
    int qt_present[4];
    long long coef_count;        // int16 coefficients in all
    size_t scan_offset;          // first byte of the entropy-coded segment
    // Huffman tables as sent (DHT): class 0 = DC, 1 = AC
    uint8_t bits[2][4][17];
    uint8_t vals[2][4][256];
    int ht_present[2][4];
};

// Parses the headers up to the start of the scan.  Returns 0 on success (info->supported says whether the device path
// can take the file), negative on a malformed stream.
int fdms_jpeg_parse(const uint8_t *data, size_t len, JpegInfo *info);

// Entropy-decodes the scan into `coef` (info->coef_count int16, zero-initialised by the callee): per component a
// (blocks_h, blocks_w, 64) array in natural (row-major, de-zigzagged) order.  Returns 0 on success.
int fdms_jpeg_decode_coefficients(const uint8_t *data, size_t len, const JpegInfo *info, int16_t *coef);
