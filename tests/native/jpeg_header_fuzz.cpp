// ASan + UBSan fuzz of the JPEG header parser incl. the OpenCV-style EXIF reader (lilliput_b200/csrc/jpeg_parse.cpp:
// jpeg_parse_header, exif_orientation_opencv) on mutated files, exact-size heap input.  CPU only.  Seeds: the files of
// tests/test_host_exif.py written out one per file.  Build like png_icc_fuzz.cpp (nvcc -x cu, same sanitizer flags),
// link with the sanitized jpeg_parse object.  Round 1: 600 000 mutants of 66 seeds, 399 202 headers accepted, no report (re-run on the last parser of the round).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "kernels.cuh"
static uint64_t s = 0x9E3779B97F4A7C15ull;
static uint32_t rnd() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 11); }
int main(int argc, char** argv) {
    long iters = atol(argv[1]);
    std::vector<std::vector<uint8_t>> seeds;
    for (int i = 2; i < argc; i++) { FILE* f = fopen(argv[i], "rb"); if (!f) continue; std::vector<uint8_t> v(70000); v.resize(fread(v.data(), 1, v.size(), f)); fclose(f); seeds.push_back(v); }
    unsigned long sink = 0, ok = 0;
    for (long it = 0; it < iters; it++) {
        std::vector<uint8_t> d = seeds[rnd() % seeds.size()];
        int m = rnd() % 4;
        if (m == 0) for (int k = 1 + rnd() % 4; k--;) d[rnd() % d.size()] = (uint8_t)rnd();
        else if (m == 1) d.resize(rnd() % (d.size() + 1));
        else if (m == 2) { size_t a = rnd() % d.size(), n = rnd() % 32; if (a + n <= d.size()) d.erase(d.begin() + a, d.begin() + a + n); }
        else for (int k = 1 + rnd() % 4; k--;) { size_t a = 2 + rnd() % 90; if (a < d.size()) d[a] = (uint8_t)rnd(); }
        uint8_t* in = (uint8_t*)malloc(d.size() ? d.size() : 1); memcpy(in, d.data(), d.size());
        lp::JpegHeader h;
        if (lp::jpeg_parse_header(in, d.size(), &h) == 0) { ok++; sink += h.width + h.orientation; }
        free(in);
    }
    printf("%ld iterations, %lu headers accepted, sink %lu\n", iters, ok, sink);
}
