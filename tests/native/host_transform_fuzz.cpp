// Probe (not part of the pytest suites): mutated files through lp_transform with the host policy layer
// (lilliput_b200/host/lilliput_host.cpp) built under AddressSanitizer + UBSan and linked over the
// REFERENCE's shims (oracle/_ref objects), so whole decode -> fit -> encode flows run on the CPU and
// every buffer the policy layer owns is checked.  Build container only (needs /root/reference).
//
//   bash tests/native/host_transform_fuzz.sh [iterations]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
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

static void mutate(std::mt19937& rng, std::vector<uint8_t>& b) {
    if (b.size() < 16) return;
    const int mode = (int)(rng() % 6);
    if (mode == 0) {
        b.resize(8 + rng() % (b.size() - 8));
        return;
    }
    if (mode == 5) {
        const size_t n = 1 + rng() % 64, a = rng() % (b.size() - 1), c = rng() % (b.size() - 1);
        for (size_t k = 0; k < n && a + k < b.size() && c + k < b.size(); k++) b[c + k] = b[a + k];
        return;
    }
    const int n = 1 + (int)(rng() % 6);
    for (int k = 0; k < n; k++) {
        const size_t i = rng() % b.size();
        switch (mode) {
            case 1: b[i] ^= (uint8_t)(1u << (rng() % 8)); break;
            case 2: b[i] = (uint8_t)rng(); break;
            case 3: b[i] = 0xFF; break;
            default: b[i] = 0; break;
        }
    }
}

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    const long iters = atol(argv[1]);
    std::vector<std::vector<uint8_t>> seeds;
    for (int i = 2; i < argc; i++) {
        auto v = read_file(argv[i]);
        if (v.size() >= 16) seeds.push_back(std::move(v));
    }
    if (seeds.empty()) return 2;
    std::mt19937 rng(getenv("LP_FUZZ_SEED") ? (unsigned)atol(getenv("LP_FUZZ_SEED")) : 99u);
    const char* exts[] = {".jpeg", ".png", ".webp", ".gif"};
    const int q_jpeg[] = {1 /* JpegQuality */, 85}, q_png[] = {16 /* PngCompression */, 3}, q_webp[] = {64 /* WebpQuality */, 80};
    long ok = 0, err[16] = {0};
    for (long it = 0; it < iters + (long)seeds.size(); it++) {
        std::vector<uint8_t> d = seeds[it < (long)seeds.size() ? (size_t)it : rng() % seeds.size()];
        if (it >= (long)seeds.size()) {
            const int rounds = 1 + (int)(rng() % 3);
            for (int r = 0; r < rounds; r++) mutate(rng, d);
        }
        d.shrink_to_fit();
        lp_image_options o;
        memset(&o, 0, sizeof o);
        const int e = (int)(rng() % 4);
        o.file_type = exts[e];
        o.width = 1 + (int)(rng() % 96);
        o.height = 1 + (int)(rng() % 96);
        o.resize_method = (int)(rng() % 3);
        o.normalize_orientation = (int)(rng() & 1);
        o.encode_options = e == 0 ? q_jpeg : e == 1 ? q_png : q_webp;
        o.encode_options_len = e == 3 ? 0 : 2;
        o.max_encode_frames = (rng() % 4 == 0) ? 1 + (int)(rng() % 3) : 0;
        o.max_encode_duration_ns = (rng() % 8 == 0) ? 100000000ll : 0;
        o.encode_timeout_ns = (rng() % 8 == 0) ? 0 : 600ll * 1000000000ll;
        o.disable_animated_output = (rng() % 8 == 0);
        const size_t cap = (rng() % 16 == 0) ? 64 + rng() % 4096 : (1u << 20);
        std::vector<uint8_t> dst(cap);
        size_t n = 0;
        const int rc = lp_transform(d.data(), d.size(), &o, dst.data(), dst.size(), &n, 1024);
        if (rc == 0) {
            if (n > cap) abort();
            ok++;
        } else if (rc < 0 && rc > -16) {
            err[-rc]++;
        }
        if ((it + 1) % 5000 == 0) fprintf(stderr, "%ld iterations\n", it + 1);
    }
    printf("done: %ld inputs, %ld transformed;", iters, ok);
    for (int k = 1; k < 16; k++)
        if (err[k]) printf(" status -%d x %ld", k, err[k]);
    printf("\n");
    return 0;
}
