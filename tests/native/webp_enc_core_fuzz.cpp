// Probe (not part of the pytest suites): the WebP encoder cores on random frames (noise, saturated, flat; 1..150 px
// a side; BGR / BGRA / alpha planes; every quality; roomy and tiny output buffers) under ASan + UBSan.
//   g++ -O1 -g -std=c++17 -fsanitize=address,undefined -fno-sanitize=signed-integer-overflow,shift-base \
//       tests/native/webp_enc_core_fuzz.cpp -o /tmp/asan/enc_fuzz && /tmp/asan/enc_fuzz 3000
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
#include "../../oracle/oracle_webp.cpp"  // host drivers over vp8_enc_core.h / vp8l_enc_core.h
int main(int argc, char** argv) {
    long iters = argc > 1 ? atol(argv[1]) : 2000;
    std::mt19937 rng(7);
    long ok8 = 0, fail8 = 0, okl = 0, faill = 0;
    size_t worst = 0;
    for (long it = 0; it < iters; it++) {
        const int w = 1 + rng() % 150, h = 1 + rng() % 150, ch = (rng() & 1) ? 3 : 4;
        const int kind = rng() % 3;
        std::vector<uint8_t> img((size_t)w * h * ch);
        for (auto& b : img) b = kind == 0 ? (uint8_t)rng() : kind == 1 ? (uint8_t)((rng() & 1) * 255) : (uint8_t)(128 + (int)(rng() % 9) - 4);
        const size_t cap = (rng() % 8 == 0) ? rng() % 512 : (size_t)w * h * 8 + 65536;
        std::vector<uint8_t> out(cap);
        const int q = rng() % 101;
        long n = vp8_cpu_encode(img.data(), (size_t)w * ch, w, h, ch, q, -1, out.data(), out.size());
        if (n > 0) { ok8++; if ((size_t)n > cap) abort(); const size_t mbs = (size_t)((w + 15) / 16) * ((h + 15) / 16); if ((size_t)n / mbs > worst) worst = (size_t)n / mbs; }
        else { fail8++; if (cap >= (size_t)w * h * 8) { printf("VP8 encode failed with a roomy buffer: %dx%d ch%d q%d kind%d\n", w, h, ch, q, kind); } }
        long m = vp8l_cpu_encode(img.data(), (size_t)w * ch, w, h, ch, out.data(), out.size());
        if (m > 0) okl++; else faill++;
        if (ch == 4) {
            std::vector<uint8_t> plane((size_t)w * h);
            for (size_t i = 0; i < plane.size(); i++) plane[i] = img[i * 4 + 3];
            vp8l_cpu_encode(plane.data(), (size_t)w, w, h, 1, out.data(), out.size());
        }
    }
    printf("VP8 %ld ok / %ld failed (worst %zu bytes per macroblock), VP8L %ld / %ld\n", ok8, fail8, worst, okl, faill);
}
