// Host half of the device JPEG ingest: marker parser + baseline Huffman decoder (see jpeg_host.cuh).
// ITU-T T.81: B.2 (markers), F.2.2 (sequential Huffman decoding), Annex C (table generation), Figure A.6 (zig-zag).
#include "jpeg_host.cuh"

#include <string.h>

namespace {

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

inline int be16(const uint8_t *p) { return (p[0] << 8) | p[1]; }

// EXIF orientation (tag 0x0112) of an APP1 payload, 1 when absent.  cv2.imread rotates / mirrors for 2..8.
int exif_orientation(const uint8_t *p, size_t n) {
    if (n < 14 || memcmp(p, "Exif\0\0", 6) != 0) return 1;
    const uint8_t *t = p + 6;
    const size_t tn = n - 6;
    const bool le = t[0] == 'I' && t[1] == 'I';
    if (!le && !(t[0] == 'M' && t[1] == 'M')) return 1;
    auto u16 = [&](size_t o) -> unsigned { return le ? (unsigned)(t[o] | (t[o + 1] << 8)) : (unsigned)((t[o] << 8) | t[o + 1]); };
    auto u32 = [&](size_t o) -> unsigned {
        return le ? (unsigned)(t[o] | (t[o + 1] << 8) | (t[o + 2] << 16) | ((unsigned)t[o + 3] << 24))
                  : (unsigned)(((unsigned)t[o] << 24) | (t[o + 1] << 16) | (t[o + 2] << 8) | t[o + 3]);
    };
    if (tn < 8) return 1;
    size_t ifd = u32(4);
    if (ifd + 2 > tn) return 1;
    const unsigned cnt = u16(ifd);
    for (unsigned i = 0; i < cnt; i++) {
        const size_t e = ifd + 2 + 12 * (size_t)i;
        if (e + 12 > tn) break;
        if (u16(e) == 0x0112) return (int)u16(e + 8);
    }
    return 1;
}

struct HuffTable {
    // 9-bit lookahead: entry = (length << 8) | symbol, 0 when the code is longer than 9 bits
    uint16_t look[512];
    int maxcode[18];   // largest code of each length, -1 if none; [17] = sentinel
    int valoffset[17]; // index of the first value of a length minus its first code
    uint8_t vals[256];
    // AC tables only: 9-bit lookahead that also covers the magnitude bits when code + magnitude fit in 9 bits (most
    // coefficients of a photograph): (value << 8) | (run << 4) | bits consumed, 0 = take the general path
    int16_t fast_ac[512];
    void build_fast_ac() {
        for (int i = 0; i < 512; i++) {
            fast_ac[i] = 0;
            const unsigned e = look[i];
            if (!e) continue;
            const int len = (int)(e >> 8), rs = (int)(e & 0xFF), run = rs >> 4, mag = rs & 15;
            if (mag == 0 || len + mag > 9) continue;
            int k = ((i << len) & 511) >> (9 - mag);
            if (k < (1 << (mag - 1))) k += (int)((~0u) << mag) + 1;  // extend
            if (k >= -128 && k <= 127) fast_ac[i] = (int16_t)((k * 256) + (run * 16) + (len + mag));
        }
    }
    bool build(const uint8_t *bits, const uint8_t *v) {
        int huffsize[257], huffcode[257];
        int p = 0;
        for (int l = 1; l <= 16; l++) {
            if (p + bits[l] > 256) return false;
            for (int i = 0; i < bits[l]; i++) huffsize[p++] = l;
        }
        const int n = p;
        huffsize[p] = 0;
        int code = 0, si = huffsize[0];
        p = 0;
        while (huffsize[p]) {
            while (huffsize[p] == si) huffcode[p++] = code++;
            if (code > (1 << si)) return false;
            code <<= 1;
            si++;
        }
        p = 0;
        for (int l = 1; l <= 16; l++) {
            if (bits[l]) {
                valoffset[l] = p - huffcode[p];
                p += bits[l];
                maxcode[l] = huffcode[p - 1];
            } else {
                maxcode[l] = -1;
            }
        }
        maxcode[17] = 0xFFFFF;
        memset(look, 0, sizeof(look));
        p = 0;
        for (int l = 1; l <= 9; l++)
            for (int i = 0; i < bits[l]; i++, p++) {
                const int first = huffcode[p] << (9 - l);
                for (int c = 0; c < (1 << (9 - l)); c++) look[first + c] = (uint16_t)((l << 8) | v[p]);
            }
        memcpy(vals, v, (size_t)n);
        return true;
    }
};

struct BitReader {
    const uint8_t *p, *end;
    uint64_t buf = 0;
    int bits = 0;
    bool hit_marker = false;
    void fill() {  // top up to more than 56 bits; after a marker / the end the stream continues with zero bits
        if (!hit_marker && end - p >= 8) {  // fast path: eight bytes at once unless one of them is 0xFF
            uint64_t w;
            memcpy(&w, p, 8);
            w = __builtin_bswap64(w);
            const uint64_t x = ~w;
            if (!((x - 0x0101010101010101ull) & ~x & 0x8080808080808080ull)) {
                const int nb = (64 - bits) >> 3;
                const uint64_t take = nb == 8 ? w : (w >> (64 - 8 * nb)) << (64 - 8 * nb);
                buf |= take >> bits;
                bits += 8 * nb;
                p += nb;
                return;
            }
        }
        while (bits <= 56) {
            unsigned b = 0;
            if (!hit_marker && p < end) {
                b = *p++;
                if (b == 0xFF) {
                    if (p < end && *p == 0x00) {
                        p++;
                    } else {  // a marker: do not consume it
                        p--;
                        hit_marker = true;
                        b = 0;
                    }
                }
            } else {
                hit_marker = true;
            }
            buf |= (uint64_t)b << (56 - bits);
            bits += 8;
        }
    }
    inline unsigned peek(int n) { return (unsigned)(buf >> (64 - n)); }
    inline void drop(int n) {
        buf <<= n;
        bits -= n;
    }
    inline int get(int n) {  // 1 <= n <= 16
        if (bits < n) fill();
        const unsigned v = peek(n);
        drop(n);
        return (int)v;
    }
    inline int decode(const HuffTable &t) {
        if (bits < 16) fill();
        const unsigned e = t.look[peek(9)];
        if (e) {
            drop(e >> 8);
            return (int)(e & 0xFF);
        }
        int l = 10;
        int code = (int)peek(10);
        while (code > t.maxcode[l]) {
            l++;
            if (l > 16) return -1;
            code = (int)peek(l);
        }
        drop(l);
        return t.vals[(code + t.valoffset[l]) & 0xFF];
    }
    void reset_to(const uint8_t *q) {
        p = q;
        buf = 0;
        bits = 0;
        hit_marker = false;
    }
};

inline int extend(int v, int s) { return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }

}  // namespace

int fdms_jpeg_parse(const uint8_t *d, size_t len, JpegInfo *info) {
    memset(info, 0, sizeof(*info));
    info->supported = 1;
    info->reason = "";
    auto unsupported = [&](const char *why) {
        if (info->supported) {
            info->supported = 0;
            info->reason = why;
        }
    };
    if (len < 4 || d[0] != 0xFF || d[1] != 0xD8) return -1;
    size_t pos = 2;
    bool have_sof = false;
    int adobe_transform = -1;
    while (pos + 4 <= len) {
        if (d[pos] != 0xFF) return -1;
        while (pos < len && d[pos] == 0xFF) pos++;  // fill bytes
        if (pos >= len) return -1;
        const int m = d[pos++];
        if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;
        if (m == 0xD9) return -1;  // EOI before any scan
        if (pos + 2 > len) return -1;
        const int seg = be16(d + pos);
        if (seg < 2 || pos + seg > len) return -1;
        const uint8_t *s = d + pos + 2;
        const int n = seg - 2;
        if (m == 0xDB) {
            int o = 0;
            while (o < n) {
                const int pq = s[o] >> 4, tq = s[o] & 15;
                o++;
                if (tq > 3 || o + 64 * (pq ? 2 : 1) > n) return -1;
                for (int i = 0; i < 64; i++) {
                    const int v = pq ? be16(s + o + 2 * i) : s[o + i];
                    info->qt[tq][kZigzag[i]] = (uint16_t)v;
                }
                info->qt_present[tq] = 1;
                o += 64 * (pq ? 2 : 1);
            }
        } else if (m == 0xC0 || m == 0xC1) {
            if (n < 6 || have_sof) return -1;
            have_sof = true;
            if (s[0] != 8) unsupported("sample precision other than 8 bits");
            info->height = be16(s + 1);
            info->width = be16(s + 3);
            info->ncomp = s[5];
            if (info->height <= 0 || info->width <= 0) unsupported("image size not in the frame header");
            if (info->ncomp != 1 && info->ncomp != 3) {
                unsupported("component count other than 1 or 3");
                info->ncomp = info->ncomp > 3 ? 3 : info->ncomp;
            }
            if (n < 6 + 3 * (int)s[5]) return -1;
            for (int c = 0; c < info->ncomp; c++) {
                info->comp[c].id = s[6 + 3 * c];
                info->comp[c].h = s[7 + 3 * c] >> 4;
                info->comp[c].v = s[7 + 3 * c] & 15;
                info->comp[c].tq = s[8 + 3 * c];
                if (info->comp[c].tq > 3) return -1;
            }
        } else if (m == 0xC2 || m == 0xC3 || (m >= 0xC5 && m <= 0xC7) || (m >= 0xC9 && m <= 0xCB) || (m >= 0xCD && m <= 0xCF)) {
            have_sof = true;
            unsupported("progressive / lossless / arithmetic-coded JPEG");
            if (n >= 6) {
                info->height = be16(s + 1);
                info->width = be16(s + 3);
            }
            return 0;
        } else if (m == 0xC4) {
            int o = 0;
            while (o + 17 <= n) {
                const int tc = s[o] >> 4, th = s[o] & 15;
                if (tc > 1 || th > 3) return -1;
                int total = 0;
                info->bits[tc][th][0] = 0;
                for (int i = 1; i <= 16; i++) {
                    info->bits[tc][th][i] = s[o + i];
                    total += s[o + i];
                }
                o += 17;
                if (total > 256 || o + total > n) return -1;
                memcpy(info->vals[tc][th], s + o, (size_t)total);
                info->ht_present[tc][th] = 1;
                o += total;
            }
        } else if (m == 0xDD) {
            if (n < 2) return -1;
            info->restart_interval = be16(s);
        } else if (m == 0xE1) {
            if (exif_orientation(s, (size_t)n) > 1) unsupported("EXIF orientation other than 1 (cv2.imread rotates)");
        } else if (m == 0xEE) {
            if (n >= 12 && memcmp(s, "Adobe", 5) == 0) adobe_transform = s[11];
        } else if (m == 0xDA) {
            if (!have_sof) return -1;
            if (n < 1 || s[0] != info->ncomp || n < 1 + 2 * s[0] + 3) {
                unsupported("non-interleaved / multi-scan JPEG");
                return 0;
            }
            for (int c = 0; c < info->ncomp; c++) {
                int idx = -1;
                for (int k = 0; k < info->ncomp; k++)
                    if (info->comp[k].id == s[1 + 2 * c]) idx = k;
                if (idx != c) {
                    unsupported("scan components out of frame order");
                    return 0;
                }
                info->comp[c].td = s[2 + 2 * c] >> 4;
                info->comp[c].ta = s[2 + 2 * c] & 15;
                if (info->comp[c].td > 3 || info->comp[c].ta > 3) return -1;
            }
            const uint8_t *t = s + 1 + 2 * info->ncomp;
            if (t[0] != 0 || t[1] != 63 || t[2] != 0) unsupported("spectral selection / successive approximation");
            info->scan_offset = pos + seg;
            break;
        }
        pos += seg;
    }
    if (!have_sof || !info->scan_offset) return -1;
    if (!info->supported) return 0;
    // colour space as libjpeg guesses it (jdapimin.c default_decompress_parms)
    if (info->ncomp == 3) {
        if (adobe_transform == 0) unsupported("Adobe RGB / untransformed components");
        else if (adobe_transform < 0 && info->comp[0].id == 'R' && info->comp[1].id == 'G' && info->comp[2].id == 'B')
            unsupported("RGB components");
    }
    // sampling
    if (info->ncomp == 1) {
        info->comp[0].h = info->comp[0].v = 1;  // a single-component scan is never interleaved: one block per MCU
    } else {
        const JpegComp &y = info->comp[0];
        const bool ok = info->comp[1].h == 1 && info->comp[1].v == 1 && info->comp[2].h == 1 && info->comp[2].v == 1 &&
                        ((y.h == 1 && y.v == 1) || (y.h == 2 && y.v == 1) || (y.h == 2 && y.v == 2));
        if (!ok) unsupported("chroma subsampling other than 4:4:4 / 4:2:2 / 4:2:0");
    }
    if (!info->supported) return 0;
    info->hmax = info->comp[0].h;
    info->vmax = info->comp[0].v;
    info->mcus_x = (info->width + 8 * info->hmax - 1) / (8 * info->hmax);
    info->mcus_y = (info->height + 8 * info->vmax - 1) / (8 * info->vmax);
    long long off = 0;
    for (int c = 0; c < info->ncomp; c++) {
        JpegComp &k = info->comp[c];
        k.blocks_w = info->mcus_x * k.h;
        k.blocks_h = info->mcus_y * k.v;
        k.down_w = (info->width * k.h + info->hmax - 1) / info->hmax;
        k.down_h = (info->height * k.v + info->vmax - 1) / info->vmax;
        k.coef_off = off;
        off += (long long)k.blocks_w * k.blocks_h * 64;
        if (!info->qt_present[k.tq]) return -1;
        if (!info->ht_present[0][k.td] || !info->ht_present[1][k.ta]) return -1;
    }
    info->coef_count = off;
    return 0;
}

int fdms_jpeg_decode_coefficients(const uint8_t *d, size_t len, const JpegInfo *info, int16_t *coef) {
    if (!info->supported || info->scan_offset >= len) return -1;
    HuffTable dc[4], ac[4];
    bool built_dc[4] = {false, false, false, false}, built_ac[4] = {false, false, false, false};
    for (int c = 0; c < info->ncomp; c++) {
        const int td = info->comp[c].td, ta = info->comp[c].ta;
        if (!built_dc[td]) {
            if (!dc[td].build(info->bits[0][td], info->vals[0][td])) return -1;
            built_dc[td] = true;
        }
        if (!built_ac[ta]) {
            if (!ac[ta].build(info->bits[1][ta], info->vals[1][ta])) return -1;
            ac[ta].build_fast_ac();
            built_ac[ta] = true;
        }
    }
    BitReader br;
    br.reset_to(d + info->scan_offset);
    br.end = d + len;
    int pred[3] = {0, 0, 0};
    int restarts_left = info->restart_interval, next_rst = 0;
    for (int my = 0; my < info->mcus_y; my++) {
        for (int mx = 0; mx < info->mcus_x; mx++) {
            if (info->restart_interval && restarts_left == 0) {
                // byte-align, expect RSTn (T.81 F.2.2.5 / E.2.4)
                // the reader may have pulled bytes of the padding into its buffer, never the marker itself: it lies ahead
                const uint8_t *q = br.p;
                while (q + 1 < br.end && !(q[0] == 0xFF && q[1] >= 0xD0 && q[1] <= 0xD7)) q++;
                if (q + 1 >= br.end) return -1;
                if ((q[1] & 7) != next_rst) return -1;
                next_rst = (next_rst + 1) & 7;
                br.reset_to(q + 2);
                pred[0] = pred[1] = pred[2] = 0;
                restarts_left = info->restart_interval;
            }
            for (int c = 0; c < info->ncomp; c++) {
                const JpegComp &k = info->comp[c];
                const HuffTable &tdc = dc[k.td], &tac = ac[k.ta];
                for (int v = 0; v < k.v; v++)
                    for (int h = 0; h < k.h; h++) {
                        int16_t *blk = coef + k.coef_off + ((long long)(my * k.v + v) * k.blocks_w + (mx * k.h + h)) * 64;
                        memset(blk, 0, 128);
                        int s = br.decode(tdc);
                        if (s < 0 || s > 11) return -1;
                        if (s) pred[c] += extend(br.get(s), s);
                        blk[0] = (int16_t)pred[c];
                        for (int i = 1; i < 64;) {
                            if (br.bits < 32) br.fill();  // a coefficient takes at most 16 + 11 bits
                            const int f = tac.fast_ac[br.peek(9)];
                            if (f) {  // code and magnitude in one look-up
                                i += (f >> 4) & 15;
                                if (i > 63) return -1;
                                br.drop(f & 15);
                                blk[kZigzag[i]] = (int16_t)(f >> 8);
                                i++;
                                continue;
                            }
                            const int rs = br.decode(tac);
                            if (rs < 0) return -1;
                            const int r = rs >> 4;
                            s = rs & 15;
                            if (s) {
                                i += r;
                                if (i > 63) return -1;
                                blk[kZigzag[i]] = (int16_t)extend(br.get(s), s);
                                i++;
                            } else {
                                if (r != 15) break;
                                i += 16;
                            }
                        }
                    }
            }
            if (info->restart_interval) restarts_left--;
        }
    }
    return 0;
}
