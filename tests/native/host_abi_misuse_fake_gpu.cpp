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
        const unsigned op = rng() % 16;
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
