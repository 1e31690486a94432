// ref_shim.cpp -- glue that turns the reference's own C++ shims (compiled from
// /root/reference where they lie) plus the host mirror into oracle/_ref/
// libref_oracle.so.  TEST INFRASTRUCTURE ONLY: this library is the parity
// checker and the CPU baseline (bench.py --impl reference); the product never
// loads it.
//
// It adds only what the reference's Mat-is-the-Go-buffer model makes trivial:
// the two additive coherence calls are no-ops here.
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include <opencv2/core.hpp>

#include "lilliput_b200.h"
#include "lp_opencv.h"

extern "C" int lp_mat_sync_host(opencv_mat) { return 0; }
extern "C" void lp_mat_mark_host_dirty(opencv_mat) {}
extern "C" const char* lp_backend_name(void) { return "reference"; }
// Framebuffer.TonemapToSDR over the reference's own tone-mapper (color_info.cpp, compiled from /root/reference)
extern "C" void tonemap_rgb_8u_inplace(uint8_t* pixels, int width, int height, int channels, uint8_t transfer, uint8_t primaries);
extern "C" int lp_mat_tonemap_to_sdr(opencv_mat mat, int transfer, int primaries) {
    cv::Mat* m = static_cast<cv::Mat*>(mat);
    if (!m || m->empty() || m->depth() != CV_8U || !m->isContinuous()) return 0;
    tonemap_rgb_8u_inplace(m->data, m->cols, m->rows, m->channels(), (uint8_t)transfer, (uint8_t)primaries);
    return 0;
}

// CPU baseline: `threads` workers, each with its own ImageOps (thread_local in
// lp_transform) and cv::setNumThreads(1), run Transform over the
// n inputs round-robin until `min_seconds` elapsed AND every worker did at
// least one image.  Returns images/second; *images_done gets the total.
extern "C" double ref_bench_transform(const uint8_t* const* in, const size_t* in_len, int n,
                                      const lp_image_options* opt, int max_size, int threads,
                                      double min_seconds, size_t out_cap, long* images_done,
                                      int* first_error) {
    cv::setNumThreads(1);
    std::atomic<long> done{0};
    std::atomic<int> err{0};
    std::atomic<int> next{0};
    std::atomic<bool> stop{false};
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) {
        pool.emplace_back([&]() {
            std::vector<uint8_t> out(out_cap);
            while (!stop.load(std::memory_order_relaxed)) {
                int i = next.fetch_add(1) % n;
                size_t len = 0;
                int rc = lp_transform(in[i], in_len[i], opt, out.data(), out.size(), &len, max_size);
                if (rc != 0) {
                    int z = 0;
                    err.compare_exchange_strong(z, rc);
                    break;
                }
                done.fetch_add(1);
            }
        });
    }
    for (;;) {
        std::this_thread::sleep_for(std::chrono::milliseconds(20));
        double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if ((el >= min_seconds && done.load() >= threads) || err.load() != 0) break;
    }
    stop.store(true);
    for (auto& th : pool) th.join();
    double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (images_done) *images_done = done.load();
    if (first_error) *first_error = err.load();
    return (double)done.load() / el;
}

// Exactly `total` Transforms spread over `threads` workers (round-robin over the n inputs);
// returns elapsed seconds (<0 on error).  One bench.py "step" of the reference arm.  The clock starts when
// every worker exists and has run one untimed Transform (its thread-local ImageOps and framebuffers are
// allocated and touched, as in a long-running service: SURVEY 8(d) "framebuffers reused like NewImageOps"),
// so thread start-up and first-touch page faults are outside the timed region.
extern "C" double ref_transform_many(const uint8_t* const* in, const size_t* in_len, int n,
                                     const lp_image_options* opt, int max_size, int threads,
                                     size_t out_cap, long total, int* first_error) {
    cv::setNumThreads(1);
    std::atomic<long> next{0};
    std::atomic<int> err{0};
    std::atomic<int> ready{0};
    std::atomic<bool> go{false};
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) {
        pool.emplace_back([&, t]() {
            std::vector<uint8_t> out(out_cap);
            {
                size_t len = 0;
                (void)lp_transform(in[t % n], in_len[t % n], opt, out.data(), out.size(), &len, max_size);
            }
            ready.fetch_add(1);
            while (!go.load(std::memory_order_acquire)) std::this_thread::yield();
            for (;;) {
                long k = next.fetch_add(1);
                if (k >= total || err.load()) break;
                int i = (int)(k % n);
                size_t len = 0;
                int rc = lp_transform(in[i], in_len[i], opt, out.data(), out.size(), &len, max_size);
                if (rc != 0) {
                    int z = 0;
                    err.compare_exchange_strong(z, rc);
                    break;
                }
            }
        });
    }
    while (ready.load() < threads) std::this_thread::sleep_for(std::chrono::milliseconds(1));
    auto t0 = std::chrono::steady_clock::now();
    go.store(true, std::memory_order_release);
    for (auto& th : pool) th.join();
    double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (first_error) *first_error = err.load();
    return err.load() ? -1.0 : el;
}
