// ASan + UBSan fuzz of the host-side iCCP / cICP extraction (lilliput_b200/csrc/png_parse.cpp: png_extract_icc, its
// zlib inflater, png_extract_cicp) and of the PNG header parser (png_parse, incl. the eXIf orientation reader) on mutated PNGs, exact-size heap buffers on both sides.  CPU only.  Seeds: any PNG files, e.g. the
// cases of tests/test_host_icc.py written out one per file.  Build and run:
//   nvcc -O1 -g -std=c++17 -x cu -Xcompiler -fsanitize=address,-fsanitize=undefined,-fno-sanitize-recover=undefined \
//        -Iinclude -Ililliput_b200/csrc -c lilliput_b200/csrc/png_parse.cpp -o /tmp/png_parse_asan.o
//   nvcc -O1 -g -std=c++17 -x cu -Xcompiler -fsanitize=address,-fsanitize=undefined -Iinclude -Ililliput_b200/csrc \
//        tests/native/png_icc_fuzz.cpp /tmp/png_parse_asan.o /tmp/jpeg_parse_asan.o -o /tmp/png_icc_fuzz   (jpeg_parse.cpp built like png_parse.cpp: the EXIF reader lives there)
//   /tmp/png_icc_fuzz 400000 seeds/*
// Round 1: 400 000 mutants of 100 seed files (iCCP and cICP cases), 8 159 of them yielding a profile, no report.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <string>
#include "kernels.cuh"  // lp::PngHeader, png_parse, png_extract_icc, png_extract_cicp
static uint64_t s = 88172645463325252ull;
static uint32_t rnd() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 11); }
int main(int argc, char** argv) {
    long iters = atol(argv[1]);
    std::vector<std::vector<uint8_t>> seeds;
    for (int i = 2; i < argc; i++) {
        FILE* f = fopen(argv[i], "rb"); if (!f) continue;
        std::vector<uint8_t> v(70000); v.resize(fread(v.data(), 1, v.size(), f)); fclose(f); seeds.push_back(v);
    }
    unsigned long sink = 0, hits = 0;
    for (long it = 0; it < iters; it++) {
        std::vector<uint8_t> d = seeds[rnd() % seeds.size()];
        int m = rnd() % 4;
        if (m == 0) for (int k = 1 + rnd() % 4; k--;) d[rnd() % d.size()] = (uint8_t)rnd();
        else if (m == 1) d.resize(rnd() % (d.size() + 1));
        else if (m == 2) { size_t a = rnd() % d.size(), n = rnd() % 32; if (a + n <= d.size()) d.erase(d.begin() + a, d.begin() + a + n); }
        else for (int k = 1 + rnd() % 3; k--;) { size_t a = 8 + rnd() % 140; if (a < d.size()) d[a] = (uint8_t)rnd(); }
        // exact-size heap copies so ASan sees any read or write past either buffer
        uint8_t* in = (uint8_t*)malloc(d.size() ? d.size() : 1); memcpy(in, d.data(), d.size());
        size_t cap = (size_t[]){0, 1, 131, 132, 600, 4096, 32768}[rnd() % 7];
        uint8_t* out = (uint8_t*)malloc(cap ? cap : 1);
        int n = lp::png_extract_icc(in, d.size(), out, cap);
        if (n < 0 || (size_t)n > cap) { printf("bad length %d cap %zu\n", n, cap); return 1; }
        for (int i = 0; i < n; i++) sink += out[i];
        uint8_t* four = (uint8_t*)malloc(4);
        if (lp::png_extract_cicp(in, d.size(), four)) sink += four[0] + four[3];
        free(four);
        lp::PngHeader ph;
        if (lp::png_parse(in, d.size(), &ph) == 0) sink += (unsigned)ph.width + ph.orientation + ph.idat.size();
        hits += n > 0;
        free(in); free(out);
    }
    printf("%ld iterations, %lu profiles returned, sink %lu\n", iters, hits, sink);
    return 0;
}
