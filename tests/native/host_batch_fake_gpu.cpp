// Probe: the host side of the batch ABI (chunk schedule, header parsing, table-set dedup, scratch layout, every
// cudaMemcpy size) under ASan + UBSan over tests/native/fake_cudart.cpp -- "device" memory is host memory, kernels
// do nothing.  Inputs: one 3-component baseline JPEG given on the command line, replicated with mutations (also in
// the header: wrong sizes, broken tables, restart markers are per-item errors or other code paths, not crashes).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "lilliput_b200.h"
#include "kernels.cuh"

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
    if (argc < 3) return 2;
    const long rounds = atol(argv[1]);
    std::mt19937 rng(11);
    long items = 0, item_err = 0, call_err = 0;
    for (int fi = 2; fi < argc; fi++) {
        const std::vector<uint8_t> seed = read_file(argv[fi]);
        lp::JpegHeader h;
        if (seed.size() < 64 || lp::jpeg_parse_header(seed.data(), seed.size(), &h) || !h.supported || h.ncomp != 3) continue;
        for (long r = 0; r < rounds; r++) {
            const int n = 1 + (int)(rng() % 40);
            lp_batch_config cfg;
            memset(&cfg, 0, sizeof cfg);
            cfg.max_images = n + (int)(rng() % 3);
            cfg.src_width = h.width;
            cfg.src_height = h.height;
            cfg.dst_width = 1 + (int)(rng() % (unsigned)h.width);
            cfg.dst_height = 1 + (int)(rng() % (unsigned)h.height);
            cfg.resize_method = (rng() & 1) ? LP_OPS_FIT : LP_OPS_RESIZE;
            cfg.jpeg_quality = 1 + (int)(rng() % 100);
            cfg.out_cap = 256 + rng() % 65536;
            cfg.chunk = (rng() % 3 == 0) ? 1 + (int)(rng() % 16) : 0;
            std::vector<std::vector<uint8_t>> files(n);
            size_t total = 0;
            for (auto& f : files) {
                f = seed;
                const int muts = (int)(rng() % 4);
                for (int m = 0; m < muts; m++) {
                    const int mode = (int)(rng() % 4);
                    if (mode == 0 && f.size() > 32) f.resize(16 + rng() % (f.size() - 16));
                    else f[rng() % f.size()] = mode == 1 ? (uint8_t)rng() : mode == 2 ? 0xFF : 0;
                }
                if (rng() % 97 == 0) f.clear();  // an empty input
                f.shrink_to_fit();
                total += f.size();
            }
            cfg.max_in_bytes = total + (rng() % 2 ? 0 : 4096);
            lp_batch* b = lp_batch_create(&cfg);
            if (!b) { call_err++; continue; }
            std::vector<const uint8_t*> in(n);
            std::vector<size_t> in_len(n), out_len(n);
            std::vector<std::vector<uint8_t>> outs(n, std::vector<uint8_t>(cfg.out_cap));
            std::vector<uint8_t*> out(n);
            std::vector<int> status(n);
            for (int i = 0; i < n; i++) { in[i] = files[i].data(); in_len[i] = files[i].size(); out[i] = outs[i].data(); }
            for (int pass = 0; pass < 2; pass++) {
                int rc;
                if ((rng() & 1) == 0) {
                    rc = lp_batch_transform(b, in.data(), in_len.data(), n, out.data(), out_len.data(), status.data());
                } else {
                    rc = lp_batch_stage(b, in.data(), in_len.data(), n, status.data());
                    float ms[LP_STAGE_COUNT];
                    if (!rc) rc = lp_batch_run(b, ms);
                    if (!rc) rc = lp_batch_fetch(b, out.data(), out_len.data(), status.data());
                }
                if (rc) call_err++;
                else
                    for (int i = 0; i < n; i++) { items++; item_err += status[i] != 0; if (out_len[i] > cfg.out_cap) abort(); }
            }
            lp_batch_destroy(b);
        }
    }
    printf("done: %ld items (%ld per-item errors), %ld failed calls\n", items, item_err, call_err);
    return 0;
}
