// Probe: random sequences of per-image ABI calls with random (often wrong) arguments -- regions off the matrix,
// negative offsets, zero sizes, strides that do not fit, channel mismatches, tiny destination buffers -- over
// tests/native/fake_cudart.cpp under ASan + UBSan.  Every call must come back with a value or an error code.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "lp_giflib.h"
#include "lp_opencv.h"
#include "lp_webp.h"
#include "lilliput_b200.h"

#include <exception>
#include <csignal>
#include <unistd.h>
extern "C" void __sanitizer_print_stack_trace(void);

int main(int argc, char** argv) {
    std::set_terminate([] {  // an exception that crossed the C ABI: show where it was thrown
        __sanitizer_print_stack_trace();
        abort();
    });
    signal(SIGALRM, [](int) {  // a call that does not come back: show where it is
        __sanitizer_print_stack_trace();
        _exit(3);
    });
    const long iters = argc > 1 ? atol(argv[1]) : 100000;
    std::mt19937 rng(argc > 2 ? (unsigned)atol(argv[2]) : 3u);
    struct Slot {
        opencv_mat m = nullptr;
        std::vector<uint8_t> store;  // backing memory of a create_from_data mat
    };
    std::vector<Slot> pool(8);
    std::vector<std::vector<uint8_t>> files;  // optional seed files (argv[3..]): sources for the decoder ops
    for (int i = 3; i < argc; i++) {
        FILE* f = fopen(argv[i], "rb");
        if (!f) continue;
        std::vector<uint8_t> v;
        uint8_t buf[4096];
        size_t n;
        while ((n = fread(buf, 1, sizeof buf, f)) > 0) v.insert(v.end(), buf, buf + n);
        fclose(f);
        if (v.size() >= 16) files.push_back(std::move(v));
    }
    auto dim = [&]() -> int {
        switch (rng() % 10) {
            case 0: return 0;
            case 1: return -(int)(rng() % 5);
            case 2: return 1;
            case 3: return 500 + (int)(rng() % 64);
            default: return 1 + (int)(rng() % 96);
        }
    };
    const int types[] = {0 /* 8UC1 */, 16 /* 8UC3 */, 24 /* 8UC4 */, 8 /* 8UC2 */, 2 /* 16U */, 18};
    unsigned long long sink = 0;
    for (long it = 0; it < iters; it++) {
        Slot& a = pool[rng() % pool.size()];
        Slot& b = pool[rng() % pool.size()];
        alarm(120);
        const unsigned op = rng() % (files.empty() ? 16 : 19);
        if (getenv("LP_TRACE")) fprintf(stderr, "op %u a=%d b=%d\n", op, (int)(&a - pool.data()), (int)(&b - pool.data()));
        switch (op) {
            case 0: {  // (re)create owning
                if (a.m) opencv_mat_release(a.m);
                a.store.clear();
                a.m = opencv_mat_create(dim(), dim(), types[rng() % 6]);
                break;
            }
            case 1: {  // (re)create over caller memory, sometimes too small
                if (a.m) opencv_mat_release(a.m);
                const int w = dim(), h = dim(), t = types[rng() % 3];
                const size_t need = (size_t)(w > 0 ? w : 0) * (h > 0 ? h : 0) * (t == 0 ? 1 : t == 16 ? 3 : 4);
                a.store.assign((rng() % 4 == 0) ? need / 2 : need + rng() % 32, (uint8_t)rng());
                a.m = opencv_mat_create_from_data(w, h, t, a.store.data(), a.store.size());
                break;
            }
            case 2: if (a.m) { sink += (unsigned)opencv_mat_get_width(a.m) + (unsigned)opencv_mat_get_height(a.m); sink += (uintptr_t)opencv_mat_get_data(a.m) & 1; } break;
            case 3: if (a.m) opencv_mat_reset(a.m); break;
            case 4: if (a.m) opencv_mat_set_color(a.m, (int)(rng() % 300) - 20, (int)(rng() % 256), (int)(rng() % 256), (int)(rng() % 256)); break;
            case 5: if (a.m) sink += (unsigned)opencv_mat_clear_to_transparent(a.m, dim() - 8, dim() - 8, dim(), dim()); break;
            case 6: if (a.m && b.m) sink += (unsigned)opencv_copy_to_region(a.m, b.m, dim() - 8, dim() - 8, dim(), dim()); break;
            case 7: if (a.m && b.m) sink += (unsigned)opencv_copy_to_region_with_alpha(a.m, b.m, dim() - 8, dim() - 8, dim(), dim()); break;
            case 8: {  // crop -> resize into b -> release the view
                if (!a.m || !b.m) break;
                opencv_mat v = opencv_mat_crop(a.m, dim() - 8, dim() - 8, dim(), dim());
                if (v) {
                    const int w = dim(), h = dim();
                    if (w > 0 && h > 0 && w <= 512 && h <= 512) opencv_mat_resize(v, b.m, w, h, (int)(rng() % 5));
                    opencv_mat_release(v);
                }
                break;
            }
            case 9: if (a.m) opencv_mat_orientation_transform((CVImageOrientation)(rng() % 10), a.m); break;
            case 10: if (a.m) { const size_t st = rng() % 4 == 0 ? 0 : rng() % 2048; const bool ok = opencv_mat_set_row_stride(a.m, st); if (getenv("LP_TRACE")) fprintf(stderr, "   stride %zu -> %d\n", st, (int)ok); sink += ok; } break;
            case 11: {  // encode a.m as JPEG / PNG into a (maybe tiny) buffer
                if (!a.m) break;
                std::vector<uint8_t> dst(rng() % 4 == 0 ? 16 + rng() % 200 : 1 << 18);
                opencv_mat d = opencv_mat_create_empty_from_data((int)dst.size(), dst.data());
                if (!d) break;
                const char* exts[] = {".jpeg", ".png", ".jpg", ".bmp", ""};
                if (opencv_encoder e = opencv_encoder_create(exts[rng() % 5], d)) {
                    const int opt[] = {1, (int)(rng() % 120) - 10, 16, (int)(rng() % 12) - 1, 2, (int)(rng() % 2)};
                    sink += opencv_encoder_write(e, a.m, opt, 2 * (rng() % 4));
                    sink += (unsigned)opencv_mat_get_height(d);
                    opencv_encoder_release(e);
                }
                opencv_mat_release(d);
                break;
            }
            case 12: {  // WebP encoder
                if (!a.m) break;
                std::vector<uint8_t> dst(rng() % 4 == 0 ? 16 + rng() % 200 : 1 << 18);
                if (webp_encoder e = webp_encoder_create(dst.data(), dst.size(), nullptr, 0, 0xFFFFFFFFu, 0)) {
                    const int opt[] = {64, (int)(rng() % 130) - 10};
                    const int frames = 1 + (int)(rng() % 3);
                    for (int f = 0; f < frames; f++) sink += webp_encoder_write(e, a.m, opt, 2, 40, 0, 0, 0, 0);
                    sink += webp_encoder_flush(e);
                    webp_encoder_release(e);
                }
                break;
            }
            case 13: {  // GIF encoder without a decoder to copy from is refused or harmless
                std::vector<uint8_t> dst(1 << 16);
                if (giflib_encoder e = giflib_encoder_create(dst.data(), dst.size())) {
                    sink += (unsigned)giflib_encoder_get_output_length(e);
                    giflib_encoder_release(e);
                }
                break;
            }
            case 14: sink += (unsigned)opencv_type_depth((int)(rng() % 64)) + (unsigned)opencv_type_channels((int)(rng() % 64)) +
                             (unsigned)opencv_type_convert_depth((int)(rng() % 64), (int)(rng() % 8)); break;
            case 16: case 17: case 18: {  // decode a (possibly damaged) file into whatever mat slot `a` holds
                if (!a.m) break;
                std::vector<uint8_t> d = files[rng() % files.size()];
                if (rng() % 3 == 0) d[rng() % d.size()] ^= (uint8_t)(1u << (rng() % 8));
                if (rng() % 7 == 0) d.resize(16 + rng() % (d.size() - 15));
                opencv_mat src = opencv_mat_create_from_data((int)d.size(), 1, 0, d.data(), d.size());
                if (!src) break;
                if (op == 16) {
                    if (opencv_decoder dec = opencv_decoder_create(src)) {
                        if (opencv_decoder_read_header(dec)) sink += opencv_decoder_read_data(dec, a.m);
                        opencv_decoder_release(dec);
                    }
                } else if (op == 17) {
                    if (webp_decoder w = webp_decoder_create(src)) {
                        for (int k = 0; k < 4; k++) {
                            sink += webp_decoder_decode(w, a.m);
                            if (!webp_decoder_has_more_frames(w)) break;
                            webp_decoder_advance_frame(w);
                        }
                        webp_decoder_release(w);
                    }
                } else if (giflib_decoder g = giflib_decoder_create(src)) {
                    for (int k = 0; k < 4; k++) {
                        if (giflib_decoder_decode_frame_header(g) != giflib_decoder_have_next_frame) break;
                        if (rng() & 1) { if (!giflib_decoder_decode_frame(g, a.m)) break; }
                        else if (giflib_decoder_skip_frame(g) != giflib_decoder_have_next_frame) break;
                    }
                    giflib_decoder_release(g);
                }
                opencv_mat_release(src);
                break;
            }
            default: if (a.m) { opencv_mat_release(a.m); a.m = nullptr; a.store.clear(); } break;
        }
        if (getenv("LP_TRACE") && a.m) fprintf(stderr, "   -> a: %d x %d\n", opencv_mat_get_width(a.m), opencv_mat_get_height(a.m));
        if ((it + 1) % 50000 == 0) fprintf(stderr, "%ld calls\n", it + 1);
    }
    for (auto& s : pool)
        if (s.m) opencv_mat_release(s.m);
    printf("done: %ld calls (sink %llu)\n", iters, sink);
    return 0;
}
