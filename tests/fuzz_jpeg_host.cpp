// Mutation fuzzer for the host half of the JPEG ingest (csrc/jpeg_host.cu: marker parser + Huffman decoder), built with
// AddressSanitizer / UBSan by tests/test_jpeg_ingest.py::test_host_decoder_survives_mutated_streams.  Seeds are small valid
// files of every supported flavour; each iteration mutates one (random bytes, header bytes / bits, truncation, injected
// markers, deleted runs), copies it into an exact-size heap block (so that any over-read is seen) and runs parse + decode.
// usage: fuzz <iterations> <seed files...>   (1 000 000 iterations ran clean when this was written)
#include "jpeg_host.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <random>
int main(int argc, char **argv) {
    std::vector<std::vector<uint8_t>> seeds;
    for (int a = 2; a < argc; a++) {
        FILE *f = fopen(argv[a], "rb"); if (!f) continue; fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
        std::vector<uint8_t> d(n); if (fread(d.data(), 1, n, f) != (size_t)n) return 2; fclose(f); seeds.push_back(d);
    }
    int iters = atoi(argv[1]);
    std::mt19937_64 rng(12345);
    long ok = 0, unsup = 0, bad_parse = 0, bad_dec = 0;
    for (int it = 0; it < iters; it++) {
        std::vector<uint8_t> d = seeds[rng() % seeds.size()];
        int mode = rng() % 6;
        size_t hdr = d.size() < 700 ? d.size() : 700;
        if (mode == 0) { int k = 1 + rng() % 8; for (int j = 0; j < k; j++) d[rng() % d.size()] = (uint8_t)rng(); }          // random bytes anywhere
        else if (mode == 1) { int k = 1 + rng() % 4; for (int j = 0; j < k; j++) d[rng() % hdr] = (uint8_t)rng(); }            // header bytes
        else if (mode == 2) { d.resize(1 + rng() % d.size()); }                                                                 // truncation
        else if (mode == 3) { int k = 1 + rng() % 4; for (int j = 0; j < k; j++) d[rng() % hdr] ^= (uint8_t)(1u << (rng() % 8)); } // header bit flips
        else if (mode == 4) { size_t p = rng() % d.size(); d[p] = 0xFF; if (p + 1 < d.size()) d[p + 1] = (uint8_t)(0xC0 + rng() % 0x40); } // inject markers
        else { size_t p = rng() % d.size(), n = 1 + rng() % 64; if (p + n < d.size()) d.erase(d.begin() + p, d.begin() + p + n); }  // delete a run
        // exact-size heap copy so that ASan sees any over-read
        uint8_t *buf = (uint8_t *)malloc(d.size() ? d.size() : 1);
        memcpy(buf, d.data(), d.size());
        JpegInfo info;
        int rc = fdms_jpeg_parse(buf, d.size(), &info);
        if (rc != 0) { bad_parse++; free(buf); continue; }
        if (!info.supported) { unsup++; free(buf); continue; }
        if (info.coef_count <= 0 || info.coef_count > (1LL << 28)) { unsup++; free(buf); continue; }
        int16_t *coef = (int16_t *)malloc((size_t)info.coef_count * 2);
        rc = fdms_jpeg_decode_coefficients(buf, d.size(), &info, coef);
        if (rc == 0) ok++; else bad_dec++;
        free(coef); free(buf);
    }
    printf("iters %d: decoded %ld, unsupported %ld, parse errors %ld, decode errors %ld\n", iters, ok, unsup, bad_parse, bad_dec);
    return 0;
}
