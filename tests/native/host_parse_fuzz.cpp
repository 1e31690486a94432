// Probe (not part of the pytest suites): throws mutated files at every HOST-side parser behind the C ABI --
// JPEG / PNG headers and ICC / cICP extraction, the WebP container walk, the GIF record walk -- with the
// library built under AddressSanitizer + UBSan.  No GPU is needed: nothing here decodes pixels.
//
//   bash tests/native/host_parse_fuzz.sh [iterations]     (builds /tmp/asan/liblp_asan.so and this file)
//
// Seeds are files on the command line (the script dumps the golden fixtures to /tmp/asan/seeds).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "lp_giflib.h"
#include "lp_opencv.h"
#include "lp_webp.h"
#include "lilliput_b200.h"
#include "kernels.cuh"  // internal host parsers behind the decode calls (multi-scan walk, table builders)

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
    const int mode = (int)(rng() % 7);
    if (mode == 0) {
        b.resize(8 + rng() % (b.size() - 8));
        return;
    }
    if (mode == 5) {  // a big-endian length / dimension field blown up
        const size_t i = rng() % (b.size() - 4);
        const uint32_t v = (rng() & 1) ? 0xFFFFFFFFu : (uint32_t)rng();
        b[i] = (uint8_t)(v >> 24); b[i + 1] = (uint8_t)(v >> 16); b[i + 2] = (uint8_t)(v >> 8); b[i + 3] = (uint8_t)v;
        return;
    }
    if (mode == 6) {  // a slice duplicated over another place
        const size_t n = 1 + rng() % 64, a = rng() % (b.size() - 1), c = rng() % (b.size() - 1);
        for (size_t k = 0; k < n && a + k < b.size() && c + k < b.size(); k++) b[c + k] = b[a + k];
        return;
    }
    const int n = 1 + (int)(rng() % 8);
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

static unsigned long long g_sink = 0;
static long g_jpeg = 0, g_scans = 0, g_png = 0, g_cv = 0, g_cv_hdr = 0, g_webp = 0, g_gif = 0, g_gif_frames = 0;

static void exercise(std::vector<uint8_t>& d) {
    uint8_t icc[4096];
    uint8_t a, b, c, e;
    g_sink += (unsigned)opencv_decoder_get_jpeg_icc(d.data(), d.size(), icc, sizeof icc);
    g_sink += (unsigned)opencv_decoder_get_png_icc(d.data(), d.size(), icc, sizeof icc);
    g_sink += (unsigned)opencv_decoder_get_png_cicp(d.data(), d.size(), &a, &b, &c, &e);
    {
        // what opencv_decoder_read_data runs on the host before any launch (abi_opencv.cu decode_jpeg_into)
        lp::JpegHeader h;
        if (lp::jpeg_parse_header(d.data(), d.size(), &h) == 0 && (h.supported || h.multiscan)) {
            g_jpeg++;
            static lp::JpegHuffSet hs;
            lp::jpeg_build_huff_set(h, &hs);
            g_sink += hs.look[0][0] + hs.long_prefix[0][0];
            for (int t = 0; t < 4; t++)  // the device indexes its lookahead table with these
                for (int j = 0; j < lp::kHuffLongPrefixes; j++)
                    if (hs.long_prefix[t][j] != 0xFFFF && hs.long_prefix[t][j] >= (1u << lp::kHuffAcLookBits)) abort();
            if (h.multiscan) {
                static std::vector<lp::JpegScanDesc> scans(256);
                static std::vector<lp::JpegHuffSet> sets(64);
                int nscans = 0, nsets = 0;
                if (lp::jpeg_parse_scans(d.data(), d.size(), h, scans.data(), (int)scans.size(), &nscans, sets.data(),
                                         (int)sets.size(), &nsets) == 0) {
                    g_scans += nscans;
                    for (int k = 0; k < nscans; k++) {  // every segment must lie inside the file
                        if ((size_t)scans[k].data_off + scans[k].data_len > d.size()) abort();
                        if (scans[k].table_set < 0 || scans[k].table_set >= nsets) abort();
                        // everything the multi-scan kernel indexes with
                        const lp::JpegScanDesc& sc = scans[k];
                        if (sc.ns < 1 || sc.ns > 3 || sc.Ss < 0 || sc.Ss > sc.Se || sc.Se > 63 || sc.Al < 0 || sc.Al > 13 ||
                            sc.Ah < 0 || sc.Ah > 15)
                            abort();
                        for (int q = 0; q < sc.ns; q++)
                            if (sc.ci[q] < 0 || sc.ci[q] >= h.ncomp || sc.td[q] < 0 || sc.td[q] > 3 || sc.ta[q] < 0 || sc.ta[q] > 3)
                                abort();
                    }
                }
            } else if (h.scan_offset + h.scan_length > d.size()) {
                abort();
            }
            lp::JpegDecodeItem it;
            memset(&it, 0, sizeof it);
            it.width = h.width; it.height = h.height; it.ncomp = h.ncomp;
            it.mcus_x = h.mcus_x; it.mcus_y = h.mcus_y;
            for (int c2 = 0; c2 < h.ncomp; c2++) { it.h[c2] = h.comp[c2].h; it.v[c2] = h.comp[c2].v; }
            uint32_t plane_bytes = 0;
            if ((long long)h.width * h.height <= (1ll << 26))
                g_sink += lp::jpeg_item_set_window(&it, 0, 0, h.width, h.height, false, &plane_bytes);
        }
        lp::PngHeader ph;
        if (lp::png_parse(d.data(), d.size(), &ph) == 0) {
            g_png++;
            // what the PNG kernels take on trust
            const bool depth_ok = ph.bit_depth == 1 || ph.bit_depth == 2 || ph.bit_depth == 4 || ph.bit_depth == 8 || ph.bit_depth == 16;
            const bool type_ok = ph.color_type == 0 || ph.color_type == 2 || ph.color_type == 3 || ph.color_type == 4 || ph.color_type == 6;
            if (ph.width < 1 || ph.height < 1 || !depth_ok || !type_ok || ph.npal < 0 || ph.npal > 256 || ph.ntrns < 0 ||
                ph.ntrns > 256 || ph.src_channels < 1 || ph.src_channels > 4 || ph.out_channels < 1 || ph.out_channels > 4 ||
                ph.bpp < 1 || ph.bpp > 8)
                abort();
            if (ph.row_bytes != ((size_t)ph.width * ph.src_channels * ph.bit_depth + 7) / 8) abort();
            size_t total = 0;
            for (const auto& sgm : ph.idat) {
                if (sgm.offset > d.size() || sgm.length > d.size() - sgm.offset) abort();
                total += sgm.length;
            }
            if (total != ph.idat_total) abort();
        }
    }
    {
        std::vector<uint8_t> grown(d.size() + 64);
        memcpy(grown.data(), d.data(), d.size());
        g_sink += opencv_png_insert_cicp(grown.data(), d.size(), grown.size(), 9, 16, 0, 1);
    }
    opencv_mat m = opencv_mat_create_from_data((int)d.size(), 1, 0 /* CV_8U */, d.data(), d.size());
    if (!m) return;
    if (opencv_decoder dec = opencv_decoder_create(m)) {
        g_cv++;
        if (opencv_decoder_read_header(dec)) {
            g_cv_hdr++;
            g_sink += (unsigned)opencv_decoder_get_width(dec) + (unsigned)opencv_decoder_get_height(dec) +
                      (unsigned)opencv_decoder_get_pixel_type(dec) + (unsigned)opencv_decoder_get_orientation(dec);
            const char* s = opencv_decoder_get_description(dec);
            g_sink += s ? strlen(s) : 0;
        }
        opencv_decoder_release(dec);
    }
    if (webp_decoder w = webp_decoder_create(m)) {
        g_webp++;
        g_sink += (unsigned)webp_decoder_get_width(w) + (unsigned)webp_decoder_get_height(w) +
                  (unsigned)webp_decoder_get_pixel_type(w) + (unsigned)webp_decoder_get_num_frames(w) +
                  (unsigned)webp_decoder_get_total_duration(w) + webp_decoder_get_bg_color(w) +
                  webp_decoder_get_loop_count(w);
        g_sink += webp_decoder_get_icc(w, icc, sizeof icc);
        for (int k = 0; k < 64 && webp_decoder_has_more_frames(w); k++) {
            webp_decoder_advance_frame(w);
            g_sink += (unsigned)webp_decoder_get_prev_frame_delay(w) + (unsigned)webp_decoder_get_prev_frame_x_offset(w);
        }
        webp_decoder_release(w);
    }
    if (giflib_decoder g = giflib_decoder_create(m)) {
        g_gif++;
        struct GifAnimationInfo info = giflib_decoder_get_animation_info(g);
        g_sink += (unsigned)info.frame_count + (unsigned)info.loop_count;
        g_sink += (unsigned)giflib_decoder_get_width(g) + (unsigned)giflib_decoder_get_height(g) +
                  (unsigned)giflib_decoder_get_num_frames(g);
        for (int k = 0; k < 64; k++) {
            if (giflib_decoder_decode_frame_header(g) != giflib_decoder_have_next_frame) break;
            g_sink += (unsigned)giflib_decoder_get_frame_width(g) + (unsigned)giflib_decoder_get_frame_height(g) +
                      (unsigned)giflib_decoder_get_prev_frame_delay(g) + (unsigned)giflib_decoder_get_prev_frame_disposal(g);
            g_gif_frames++;
            if (giflib_decoder_skip_frame(g) != giflib_decoder_have_next_frame) break;
        }
        giflib_decoder_release(g);
    }
    opencv_mat_release(m);
}

int main(int argc, char** argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s iterations seed-file...\n", argv[0]);
        return 2;
    }
    const long iters = atol(argv[1]);
    std::vector<std::vector<uint8_t>> seeds;
    for (int i = 2; i < argc; i++) {
        auto v = read_file(argv[i]);
        if (v.size() >= 16) seeds.push_back(std::move(v));
    }
    if (seeds.empty()) return 2;
    std::mt19937 rng(getenv("LP_FUZZ_SEED") ? (unsigned)atol(getenv("LP_FUZZ_SEED")) : 20260923u);
    for (auto& s : seeds) {  // the seeds themselves first
        auto d = s;
        exercise(d);
    }
    for (long it = 0; it < iters; it++) {
        // exact-size heap copy: the sanitizer sees any read past the end of the input
        std::vector<uint8_t> d = seeds[rng() % seeds.size()];
        const int rounds = 1 + (int)(rng() % 3);
        for (int r = 0; r < rounds; r++) mutate(rng, d);
        d.shrink_to_fit();
        exercise(d);
        if ((it + 1) % 20000 == 0) fprintf(stderr, "%ld iterations\n", it + 1);
    }
    printf("done: %ld mutated inputs over %zu seeds; accepted: %ld JPEG headers (%ld extra scans), %ld PNG headers, %ld JPEG/PNG decoders (%ld headers), %ld WebP containers, "
           "%ld GIFs (%ld frame headers) (sink %llu)\n",
           iters, seeds.size(), g_jpeg, g_scans, g_png, g_cv, g_cv_hdr, g_webp, g_gif, g_gif_frames, g_sink);
    return 0;
}
