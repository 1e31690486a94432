// Probe: N threads call lp_transform (and the raw region ops) concurrently over tests/native/fake_cudart.cpp with the
// library built under ThreadSanitizer -- lilliput is called from many goroutines, i.e. many OS threads, each with its
// own handles; what they share inside the library (per-thread streams, the tap-table cache, global limits) must be
// race-free.     bash tests/native/host_threads_fake_gpu.sh
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <thread>
#include <vector>

#include "lilliput_b200.h"

static std::vector<uint8_t> read_file(const char* p) {
    std::vector<uint8_t> v;
    FILE* f = fopen(p, "rb");
    if (!f) return v;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    v.resize(n > 0 ? (size_t)n : 0);
    if (n > 0 && fread(v.data(), 1, (size_t)n, f) != (size_t)n) v.clear();
    fclose(f);
    return v;
}

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    const int nthreads = atoi(argv[1]);
    const long iters = atol(argv[2]);
    std::vector<std::vector<uint8_t>> seeds;
    for (int i = 3; i < argc; i++) {
        auto v = read_file(argv[i]);
        if (v.size() >= 16) seeds.push_back(std::move(v));
    }
    std::atomic<long> ok{0}, failed{0};
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++)
        th.emplace_back([&, t] {
            std::mt19937 rng(100 + t);
            const char* exts[] = {".jpeg", ".png", ".webp", ".gif"};
            const int q_jpeg[] = {1, 85}, q_png[] = {16, 3}, q_webp[] = {64, 80};
            std::vector<uint8_t> dst(1 << 20);
            for (long it = 0; it < iters; it++) {
                const std::vector<uint8_t>& d = seeds[rng() % seeds.size()];
                lp_image_options o;
                memset(&o, 0, sizeof o);
                const int e = (int)(rng() % 4);
                o.file_type = exts[e];
                o.width = 1 + (int)(rng() % 64);
                o.height = 1 + (int)(rng() % 64);
                o.resize_method = (int)(rng() % 3);
                o.encode_options = e == 0 ? q_jpeg : e == 1 ? q_png : q_webp;
                o.encode_options_len = e == 3 ? 0 : 2;
                o.encode_timeout_ns = 600ll * 1000000000ll;
                size_t n = 0;
                (lp_transform(d.data(), d.size(), &o, dst.data(), dst.size(), &n, 1024) == 0 ? ok : failed)++;
            }
        });
    for (auto& x : th) x.join();
    printf("done: %d threads, %ld ok, %ld failed\n", nthreads, ok.load(), failed.load());
    return 0;
}
