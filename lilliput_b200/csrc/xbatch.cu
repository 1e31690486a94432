// xbatch.cu -- the heterogeneous batch entry point (include/lilliput_b200.h: lp_xbatch_*): N independent
// images of ANY supported format and size through ImageOps.Transform with one set of options, the
// per-image work packed into grid launches.
//
// Per-item semantics are those of lp_transform (= lilliput's NewDecoder + ImageOps.Transform,
// ref lilliput.go:129-164, ops.go:352-444).  What differs is the schedule:
//   1. headers of all items are parsed on the host by a few threads (format sniff as lilliput.go:129-164);
//   2. items are grouped by (decoder kind, source geometry) -- a group shares its Fit crop / output size
//      (ref ops.go:243-255, opencv.go:331-363), so every stage of a group is ONE launch over all its images:
//        JPEG  -> the lp_batch pipeline (batch.cu): parallel Huffman, IDCT, colour, resize, encode
//        PNG   -> IDAT gather + warp-parallel inflate + defilter + convert (png_decode.cu), resize
//        WebP  -> VP8 key frames, one frame per warp (webp_decode.cu), resize
//        GIF   -> every frame of every animation: LZW (one warp per frame), per-pixel compositor over the
//                 frame sequence (gif_decode.cu), resize of every composited canvas
//      and the sinks: JPEG (jpeg_encode.cu), lossy WebP still / animation (webp_encode.cu);
//   3. anything the grid path does not cover (progressive JPEG, EXIF-rotated sources, ICC profiles to carry,
//      lossless WebP output, PNG / GIF output ...) and any item whose grid stage fails goes through
//      lp_transform on a worker thread -- still this library's device kernels, one image per call -- so the
//      status and bytes of EVERY item are what lp_transform would have returned.
// Two worker lanes, each with half of the device arena and its own stream, process chunks of groups
// concurrently, so one lane's PCIe copies and host-side container work overlap the other lane's kernels.
// Nothing is exchanged between images, lanes or GPUs.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <tuple>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"
#include "lilliput_host.hpp"
#include "lp_opencv.h"

using namespace lp;

namespace lp {
lp_batch* batch_create_in(const lp_batch_config* cfg, uint8_t* dev_arena, size_t dev_bytes, uint8_t* host_arena,
                          size_t host_bytes);
}  // namespace lp

namespace {

enum Kind { K_FALLBACK = 0, K_JPEG = 1, K_PNG = 2, K_WEBP = 3, K_GIF = 4 };
enum Sink { S_NONE = 0, S_JPEG = 1, S_WEBP = 2 };

struct XItem {
    Kind kind = K_FALLBACK;
    int w = 0, h = 0, ch = 0;       // decoded frame
    int ow = 0, oh = 0;             // output size
    int cx = 0, cy = 0, cw = 0, chh = 0;  // crop rectangle fed to the resize
    int jpeg_sampling = 0;          // (h0<<12)|(v0<<8)|... groups JPEGs of one component layout
    std::unique_ptr<PngHeader> png;
    WebpStillInfo webp;
    GifAnimPlan* gif = nullptr;
    int gif_frames = 0;
};

struct Task {
    Kind kind;
    std::vector<int> idx;  // items (for GIF: animations)
};

struct Lane {
    int id = 0;
    cudaStream_t st = nullptr;
    uint8_t* dev = nullptr;
    size_t dev_bytes = 0;
    uint8_t* host = nullptr;  // pinned
    size_t host_bytes = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    double ms_decode = 0, ms_resize = 0, ms_encode = 0;
    size_t h2d = 0, d2h = 0;
    long launches = 0;
};

struct Bump {
    uint8_t* base;
    size_t cap, used = 0;
    template <class T>
    T* take(size_t bytes) {
        const size_t need = round_up(bytes, (size_t)256);
        if (used + need > cap) return nullptr;
        T* p = reinterpret_cast<T*>(base + used);
        used += need;
        return p;
    }
};

}  // namespace

struct lp_xbatch {
    lp_xbatch_config cfg;
    int device = 0;
    int threads = 4;
    Lane lanes[2];
    uint8_t* arena = nullptr;
    size_t arena_bytes = 0;
    uint8_t* host_arena = nullptr;
    size_t host_bytes = 0;
    lp_xbatch_stats stats;
    // per call
    const uint8_t* const* in = nullptr;
    const size_t* in_len = nullptr;
    uint8_t* const* out = nullptr;
    size_t out_cap = 0;
    size_t* out_len = nullptr;
    int* status = nullptr;
    lp_image_options opt;
    Sink sink = S_NONE;
    int quality = 0;
    std::vector<XItem> items;
    std::vector<int> fallback;
    std::mutex fb_mu;
};

static void push_fallback(lp_xbatch* X, int i) {
    std::lock_guard<std::mutex> g(X->fb_mu);
    X->fallback.push_back(i);
}

// ------------------------------------------------------------------ parse

static int option_value(const lp_image_options& o, int key, int dflt) {
    int v = dflt;
    for (size_t i = 0; i + 1 < o.encode_options_len; i += 2)
        if (o.encode_options[i] == key) v = o.encode_options[i + 1];
    return v;
}

// Output size and crop of a still of w x h (ref ops.go:449-470, 243-255; opencv.go:331-363).  false: not a
// resize the grid path does (NoResize).
static bool plan_geometry(const lp_image_options& o, XItem* it) {
    if (o.resize_method == LP_OPS_FIT) {
        lilliput::calculateExpectedSize(it->w, it->h, o.width, o.height, &it->ow, &it->oh);
        if (it->ow < 1 || it->oh < 1) return false;
        lilliput::fitCropRect(it->w, it->h, it->ow, it->oh, &it->cx, &it->cy, &it->cw, &it->chh);
        return true;
    }
    if (o.resize_method == LP_OPS_RESIZE) {
        it->ow = o.width;
        it->oh = o.height;
        if (it->ow < 1 || it->oh < 1) return false;
        it->cx = it->cy = 0;
        it->cw = it->w;
        it->chh = it->h;
        return true;
    }
    return false;
}

static void parse_item(lp_xbatch* X, int i) {
    XItem& it = X->items[i];
    const uint8_t* d = X->in[i];
    const size_t n = X->in_len[i];
    it.kind = K_FALLBACK;
    if (!d || n < 16 || X->sink == S_NONE) return;
    const int max_side = X->cfg.max_size > 0 ? X->cfg.max_size : 8192;
    static const uint8_t png_sig[8] = {0x89, 0x50, 0x4E, 0x47, 0x0D, 0x0A, 0x1A, 0x0A};
    if (d[0] == 0xFF && d[1] == 0xD8) {
        if (X->sink != S_JPEG) return;  // JPEG -> WebP carries the ICC profile and is not a measured path: per image
        JpegHeader h;
        if (jpeg_parse_header(d, n, &h) != LP_OK || !h.supported || h.multiscan) return;
        if (h.ncomp != 3) return;
        if (h.orientation >= 2 && h.orientation <= 8) return;
        if (h.width > max_side || h.height > max_side) return;
        it.w = h.width;
        it.h = h.height;
        it.ch = 3;
        it.jpeg_sampling = 0;
        for (int c = 0; c < 3; c++) it.jpeg_sampling = (it.jpeg_sampling << 8) | (h.comp[c].h << 4) | h.comp[c].v;
        if (!plan_geometry(X->opt, &it)) return;
        it.kind = K_JPEG;
        return;
    }
    if (!memcmp(d, png_sig, 8)) {
        std::unique_ptr<PngHeader> h(new PngHeader);
        if (png_parse(d, n, h.get()) != LP_OK) return;
        if (h->orientation != 1 || h->idat.empty() || h->idat_total < 2) return;
        if (h->width > max_side || h->height > max_side) return;
        if (h->bit_depth == 16) return;  // reported as a 16-bit type: the 8-bit Framebuffer path has its own rules
        uint8_t cicp[4];
        if (png_extract_cicp(d, n, cicp)) return;                      // cICP handling stays with Transform
        if (X->sink == S_WEBP) {
            std::vector<uint8_t> icc(32768);
            if (png_extract_icc(d, n, icc.data(), icc.size()) > 0) return;  // profile to carry into the WebP
        }
        // the IDAT payloads must form one forward run of the file (they do in every valid PNG)
        for (size_t k = 1; k < h->idat.size(); k++)
            if (h->idat[k].offset < h->idat[k - 1].offset + h->idat[k - 1].length) return;
        if (h->idat.back().offset + h->idat.back().length > n) return;
        it.w = h->width;
        it.h = h->height;
        it.ch = h->out_channels;
        if (it.ch != 3 && it.ch != 4) return;  // gray PNGs: per image
        if (!plan_geometry(X->opt, &it)) return;
        it.png = std::move(h);
        it.kind = K_PNG;
        return;
    }
    if (!memcmp(d, "RIFF", 4) && !memcmp(d + 8, "WEBP", 4)) {
        WebpStillInfo w;
        if (!webp_still_info(d, n, &w) || !w.simple_lossy) return;
        if (w.width > max_side || w.height > max_side) return;
        it.webp = w;
        it.w = w.width;
        it.h = w.height;
        it.ch = 3;
        if (!plan_geometry(X->opt, &it)) return;
        it.kind = K_WEBP;
        return;
    }
    if (!memcmp(d, "GIF8", 4)) {
        if (X->sink != S_WEBP || X->opt.disable_animated_output || X->opt.max_encode_frames != 0 ||
            X->opt.max_encode_duration_ns != 0)
            return;
        GifAnimPlan* p = gif_plan_parse(d, n, 4096);
        if (!p) return;
        int w = 0, h = 0, nf = 0;
        gif_plan_info(p, &w, &h, &nf, nullptr, nullptr);
        it.w = w;
        it.h = h;
        it.ch = 4;
        if (nf < 2 || w > max_side || h > max_side || !plan_geometry(X->opt, &it)) {  // stills and odd files: per image
            gif_plan_free(p);
            return;
        }
        it.gif = p;
        it.gif_frames = nf;
        it.kind = K_GIF;
        return;
    }
}

// ------------------------------------------------------------------ sinks

// resized frames (n x ow x oh x ch, `stride` apart) -> encoded files in the callers' buffers
static void sink_encode(lp_xbatch* X, Lane& L, Bump& bump, const std::vector<int>& idx, const uint8_t* d_frames,
                        size_t stride, int ow, int oh, int ch, std::vector<int>* failed) {
    const int n = (int)idx.size();
    if (X->sink == S_JPEG) {
        const size_t slot = round_up(std::min(X->out_cap, std::max((size_t)65536, (size_t)ow * oh * ch)), (size_t)256);
        uint8_t* d_out = bump.take<uint8_t>((size_t)n * slot);
        uint32_t* d_len = bump.take<uint32_t>((size_t)n * 4);
        uint8_t* d_packed = bump.take<uint8_t>((size_t)n * slot + 16);
        auto* d_off = bump.take<unsigned long long>((size_t)(n + 1) * 8);
        void* scratch = bump.take<uint8_t>(jpeg_encode_scratch_bytes(ow, oh, ch, n, slot));
        if (!d_out || !d_len || !d_packed || !d_off || !scratch) {
            failed->insert(failed->end(), idx.begin(), idx.end());
            return;
        }
        const bool dbg = getenv("LP_DEBUG") != nullptr;
        const auto tq0 = std::chrono::steady_clock::now();
        JpegEncodeBatch e;
        e.frames = d_frames;
        e.frame_img_stride = stride;
        e.frame_row_stride = (size_t)ow * ch;
        e.width = ow;
        e.height = oh;
        e.channels = ch;
        e.quality = X->quality;
        e.n = n;
        e.out = d_out;
        e.out_cap = slot;
        e.out_len = d_len;
        e.scratch = scratch;
        int rc = jpeg_encode_launch(e, L.st, nullptr);
        if (!rc) rc = compact_launch(d_out, slot, d_len, (uint32_t)slot, n, d_packed, d_off, L.st);
        std::vector<unsigned long long> off((size_t)n + 1);
        std::vector<uint32_t> len((size_t)n);
        if (!rc && (cudaMemcpyAsync(off.data(), d_off, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, L.st) != cudaSuccess ||
                    cudaMemcpyAsync(len.data(), d_len, (size_t)n * 4, cudaMemcpyDeviceToHost, L.st) != cudaSuccess ||
                    cudaStreamSynchronize(L.st) != cudaSuccess))
            rc = LP_ERR_CUDA;
        const size_t total = rc ? 0 : (size_t)off[n];
        if (!rc && total > L.host_bytes) rc = LP_ERR_CUDA;
        if (!rc && total &&
            (cudaMemcpyAsync(L.host, d_packed, total, cudaMemcpyDeviceToHost, L.st) != cudaSuccess ||
             cudaStreamSynchronize(L.st) != cudaSuccess))
            rc = LP_ERR_CUDA;
        if (rc) {
            cudaGetLastError();
            failed->insert(failed->end(), idx.begin(), idx.end());
            return;
        }
        L.d2h += total + (size_t)n * 12;
        const auto tq1 = std::chrono::steady_clock::now();
        if (dbg)
            fprintf(stderr, "[lilliput_b200] jpeg sink: n=%d %dx%dx%d slot=%zu total=%zu: launches + D2H %.2f ms\n", n, ow, oh, ch, slot,
                    total, std::chrono::duration<double, std::milli>(tq1 - tq0).count());
        for (int k = 0; k < n; k++) {
            const int i = idx[k];
            if (len[k] == 0 || len[k] > slot || len[k] > X->out_cap) {  // did not fit the slot: let Transform decide
                failed->push_back(i);
                continue;
            }
            memcpy(X->out[i], L.host + off[k], len[k]);
            X->out_len[i] = len[k];
            X->status[i] = LP_OK;
        }
        return;
    }
    // lossy WebP stills
    std::vector<WebpEncodedFrame> frames;
    int rc = webp_encode_lossy_batch(d_frames, stride, (size_t)ow * ch, ow, oh, ch, n, X->quality, &frames, L.st);
    if (rc) {
        cudaGetLastError();
        failed->insert(failed->end(), idx.begin(), idx.end());
        return;
    }
    std::vector<uint8_t> file;
    for (int k = 0; k < n; k++) {
        const int i = idx[k];
        if (frames[k].image.empty()) {
            failed->push_back(i);
            continue;
        }
        webp_assemble(&frames[k], 1, nullptr, 0, 0xFFFFFFFFu, 0, &file);
        L.d2h += frames[k].image.size() + frames[k].alph.size();
        if (file.size() > X->out_cap) {  // ref webp.cpp:546-551 -> size 0 -> ErrInvalidImage (webp.go:249-251)
            X->status[i] = LP_ERR_INVALID_IMAGE;
            X->out_len[i] = 0;
            continue;
        }
        memcpy(X->out[i], file.data(), file.size());
        X->out_len[i] = file.size();
        X->status[i] = LP_OK;
    }
}

static void lane_time(Lane& L, int from, int to, double* acc) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, L.ev[from], L.ev[to]) == cudaSuccess) *acc += ms;
}

// ------------------------------------------------------------------ PNG / WebP tasks
// A task holds items of ONE decoder kind but any geometry, ordered so that equal geometries are adjacent: the
// entropy stage (inflate / boolean decoder -- the long pole, latency-bound per stream) is one launch over the whole
// task; the geometry-bound stages (resize, encode) are one launch per run of equal geometry.

struct Run {
    int k0, k1;  // positions inside the task
};
static std::vector<Run> runs_of(const lp_xbatch* X, const std::vector<int>& idx) {
    std::vector<Run> r;
    for (int k = 0; k < (int)idx.size();) {
        const XItem& a = X->items[idx[k]];
        int e = k + 1;
        while (e < (int)idx.size()) {
            const XItem& c = X->items[idx[e]];
            if (c.w != a.w || c.h != a.h || c.ch != a.ch) break;
            e++;
        }
        r.push_back(Run{k, e});
        k = e;
    }
    return r;
}

// resize of every run + sinks.  frame_off / frame_stride describe the decoded frames; ok[k] = decode succeeded.
static void resize_and_encode(lp_xbatch* X, Lane& L, Bump& bump, const std::vector<int>& idx, const std::vector<Run>& runs,
                              const uint8_t* d_frames, const std::vector<uint64_t>& frame_off, const std::vector<char>& ok,
                              std::vector<int>* failed) {
    struct Out { uint8_t* d; size_t stride; };
    std::vector<Out> outs(runs.size());
    bool good = true;
    cudaEventRecord(L.ev[1], L.st);
    for (size_t r = 0; r < runs.size() && good; r++) {
        const XItem& g = X->items[idx[runs[r].k0]];
        const int n = runs[r].k1 - runs[r].k0;
        const size_t fs = round_up((size_t)g.w * g.h * g.ch, (size_t)256);
        outs[r].stride = round_up((size_t)g.ow * g.oh * g.ch, (size_t)256);
        outs[r].d = bump.take<uint8_t>((size_t)n * outs[r].stride + 256);
        if (!outs[r].d) { good = false; break; }
        ResizeArgs a{d_frames + frame_off[runs[r].k0], fs, (size_t)g.w * g.ch, g.ch, g.cx, g.cy, g.cw, g.chh, outs[r].d,
                     outs[r].stride, (size_t)g.ow * g.ch, g.ow, g.oh, n, 3};
        good = resize_launch(a, L.st) == LP_OK;
    }
    cudaEventRecord(L.ev[2], L.st);
    if (good) good = cudaStreamSynchronize(L.st) == cudaSuccess;
    if (!good) {
        cudaGetLastError();
        failed->insert(failed->end(), idx.begin(), idx.end());
        return;
    }
    lane_time(L, 0, 1, &L.ms_decode);
    lane_time(L, 1, 2, &L.ms_resize);
    cudaEventRecord(L.ev[2], L.st);
    for (size_t r = 0; r < runs.size(); r++) {
        const XItem& g = X->items[idx[runs[r].k0]];
        std::vector<int> sub;
        bool all = true;
        for (int k = runs[r].k0; k < runs[r].k1; k++) {
            all = all && ok[k];
            sub.push_back(idx[k]);
        }
        if (!all) {  // rare: a corrupt stream in the run -- the run's items take the per-image path, which reports the precise error
            failed->insert(failed->end(), sub.begin(), sub.end());
            continue;
        }
        sink_encode(X, L, bump, sub, outs[r].d, outs[r].stride, g.ow, g.oh, g.ch, failed);
    }
    cudaEventRecord(L.ev[3], L.st);
    cudaEventSynchronize(L.ev[3]);
    lane_time(L, 2, 3, &L.ms_encode);
}

// resize of items [k0, k1) of a task (their frames at d_frames + frame_off[k]) into the task's output area
static bool resize_range(lp_xbatch* X, Lane& L, const std::vector<int>& idx, int k0, int k1, const uint8_t* d_frames,
                         const std::vector<uint64_t>& frame_off, uint8_t* d_out, const std::vector<uint64_t>& out_off) {
    for (int k = k0; k < k1;) {
        const XItem& g = X->items[idx[k]];
        int e = k + 1;
        while (e < k1) {
            const XItem& c = X->items[idx[e]];
            if (c.w != g.w || c.h != g.h || c.ch != g.ch) break;
            e++;
        }
        const size_t fs = round_up((size_t)g.w * g.h * g.ch, (size_t)256), os = round_up((size_t)g.ow * g.oh * g.ch, (size_t)256);
        ResizeArgs a{d_frames + frame_off[k], fs, (size_t)g.w * g.ch, g.ch, g.cx, g.cy, g.cw, g.chh, d_out + out_off[k], os,
                     (size_t)g.ow * g.ch, g.ow, g.oh, e - k, 3};
        if (resize_launch(a, L.st) != LP_OK) return false;
        k = e;
    }
    return true;
}

// sinks over the runs of equal geometry of a whole task (resized frames at d_out + out_off[k])
static void encode_runs(lp_xbatch* X, Lane& L, Bump& bump, const std::vector<int>& idx, const uint8_t* d_out,
                        const std::vector<uint64_t>& out_off, const std::vector<char>& ok, std::vector<int>* failed) {
    cudaEventRecord(L.ev[2], L.st);
    for (const Run& r : runs_of(X, idx)) {
        const XItem& g = X->items[idx[r.k0]];
        std::vector<int> sub;
        bool all = true;
        for (int k = r.k0; k < r.k1; k++) {
            all = all && ok[k];
            sub.push_back(idx[k]);
        }
        if (!all) {  // rare: a corrupt stream in the run -- its items take the per-image path, which reports the precise error
            failed->insert(failed->end(), sub.begin(), sub.end());
            continue;
        }
        sink_encode(X, L, bump, sub, d_out + out_off[r.k0], round_up((size_t)g.ow * g.oh * g.ch, (size_t)256), g.ow, g.oh, g.ch, failed);
    }
    cudaEventRecord(L.ev[3], L.st);
    cudaEventSynchronize(L.ev[3]);
    lane_time(L, 2, 3, &L.ms_encode);
}

// PNG task.  Inflate wants every stream in flight at once and needs only the compressed stream and the scanlines
// (~1.5 x the pixel bytes); the packed frames are needed only between defilter and resize.  So: ONE inflate launch over
// the whole task, then defilter -> resize over windows of frames that reuse one buffer, then the sinks over the task.
static void run_png(lp_xbatch* X, Lane& L, const std::vector<int>& idx) {
    const int n = (int)idx.size();
    std::vector<int> failed;
    Bump bump{L.dev, L.dev_bytes};
    std::vector<PngDecodeItem> items((size_t)n);
    std::vector<SegCopy> segs;
    std::vector<uint64_t> file_off((size_t)n), frame_off((size_t)n), out_off((size_t)n);
    std::vector<int> win_first;  // first item of every frame window
    size_t in_bytes = 0, raw_bytes = 0, out_bytes = 0, win_bytes = 0, win_max = 0;
    const size_t kWindow = std::min<size_t>((size_t)12 << 30, L.dev_bytes / 5);  // frames buffer
    int max_w = 0, max_h = 0;
    for (int k = 0; k < n; k++) {
        const PngHeader& ph = *X->items[idx[k]].png;
        const size_t span = ph.idat.back().offset + ph.idat.back().length - ph.idat.front().offset;
        file_off[k] = in_bytes;
        in_bytes += round_up(span + 16, (size_t)16);
    }
    size_t zg = in_bytes;  // gathered streams follow the uploaded file spans
    for (int k = 0; k < n; k++) {
        const XItem& xi = X->items[idx[k]];
        const PngHeader& ph = *xi.png;
        PngDecodeItem& it = items[k];
        memset(&it, 0, sizeof(it));
        it.z_len = (uint32_t)ph.idat_total;
        it.width = ph.width;
        it.height = ph.height;
        it.bit_depth = ph.bit_depth;
        it.color_type = ph.color_type;
        it.src_channels = ph.src_channels;
        it.out_channels = ph.out_channels;
        it.bpp = ph.bpp;
        it.row_bytes = (uint32_t)ph.row_bytes;
        it.frame_stride = (uint32_t)((size_t)xi.w * xi.ch);
        it.interlace = ph.interlace ? 1 : 0;
        png_item_set_passes(&it);
        it.npal = ph.npal;
        it.ntrns = ph.ntrns;
        it.has_trns = ph.has_trns;
        memcpy(it.trns_rgb, ph.trns_rgb, sizeof(it.trns_rgb));
        memcpy(it.palette, ph.palette, sizeof(it.palette));
        memcpy(it.trns, ph.trns, sizeof(it.trns));
        if (ph.idat.size() == 1) {
            it.z_off = file_off[k];
        } else {
            it.z_off = zg;
            size_t o = zg;
            for (const PngSegment& sg : ph.idat) {
                segs.push_back(SegCopy{file_off[k] + (sg.offset - ph.idat.front().offset), o, (uint32_t)sg.length, 0});
                o += sg.length;
            }
            zg += round_up(ph.idat_total + 16, (size_t)16);
        }
        it.raw_off = raw_bytes;
        raw_bytes += round_up((size_t)it.raw_total + 64, (size_t)256);
        const size_t fb = round_up((size_t)xi.w * xi.h * xi.ch, (size_t)256);
        if (k == 0 || win_bytes + fb > kWindow) {  // open a new frame window
            win_first.push_back(k);
            win_bytes = 0;
        }
        frame_off[k] = win_bytes;
        it.frame_off = win_bytes;
        win_bytes += fb;
        win_max = std::max(win_max, win_bytes);
        out_off[k] = out_bytes;
        out_bytes += round_up((size_t)xi.ow * xi.oh * xi.ch, (size_t)256);
        max_w = std::max(max_w, xi.w);
        max_h = std::max(max_h, xi.h);
    }
    win_first.push_back(n);
    uint8_t* d_in = bump.take<uint8_t>(zg + 4096);
    PngDecodeItem* d_items = bump.take<PngDecodeItem>((size_t)n * sizeof(PngDecodeItem));
    SegCopy* d_segs = bump.take<SegCopy>(segs.size() * sizeof(SegCopy) + 16);
    uint8_t* d_raw = bump.take<uint8_t>(raw_bytes + 256);
    uint8_t* d_frames = bump.take<uint8_t>(win_max + 256);
    uint8_t* d_out = bump.take<uint8_t>(out_bytes + 256);
    if (!d_in || !d_items || !d_segs || !d_raw || !d_frames || !d_out) {
        for (int i : idx) push_fallback(X, i);
        return;
    }
    bool ok = true;
    for (int k = 0; k < n && ok; k++) {
        const PngHeader& ph = *X->items[idx[k]].png;
        const size_t span = ph.idat.back().offset + ph.idat.back().length - ph.idat.front().offset;
        ok = cudaMemcpyAsync(d_in + file_off[k], X->in[idx[k]] + ph.idat.front().offset, span, cudaMemcpyHostToDevice,
                             L.st) == cudaSuccess;
        L.h2d += span;
    }
    ok = ok && cudaMemcpyAsync(d_items, items.data(), (size_t)n * sizeof(PngDecodeItem), cudaMemcpyHostToDevice, L.st) == cudaSuccess;
    if (ok && !segs.empty())
        ok = cudaMemcpyAsync(d_segs, segs.data(), segs.size() * sizeof(SegCopy), cudaMemcpyHostToDevice, L.st) == cudaSuccess;
    cudaEventRecord(L.ev[0], L.st);
    if (ok && !segs.empty()) ok = seg_copy_launch(d_segs, (int)segs.size(), d_in, L.st) == LP_OK;
    PngDecodeBatch b;
    b.items = d_items;
    b.z = d_in;
    b.raw = d_raw;
    b.frames = d_frames;
    b.n = n;
    b.max_width = max_w;
    b.max_height = max_h;
    if (ok) ok = png_inflate_launch(b, L.st) == LP_OK;
    for (size_t wdx = 0; wdx + 1 < win_first.size() && ok; wdx++) {
        const int k0 = win_first[wdx], k1 = win_first[wdx + 1];
        ok = png_unfilter_launch(b, k0, k1 - k0, L.st) == LP_OK;
        if (ok) ok = resize_range(X, L, idx, k0, k1, d_frames, frame_off, d_out, out_off);
    }
    cudaEventRecord(L.ev[1], L.st);
    if (ok) ok = cudaMemcpyAsync(items.data(), d_items, (size_t)n * sizeof(PngDecodeItem), cudaMemcpyDeviceToHost, L.st) == cudaSuccess;
    if (ok) ok = cudaStreamSynchronize(L.st) == cudaSuccess;
    if (!ok) {
        cudaGetLastError();
        for (int i : idx) push_fallback(X, i);
        return;
    }
    lane_time(L, 0, 1, &L.ms_decode);  // (the resize launches sit between the defilter launches here: counted as decode)
    std::vector<char> good((size_t)n);
    for (int k = 0; k < n; k++) good[k] = items[k].status == 0;
    encode_runs(X, L, bump, idx, d_out, out_off, good, &failed);
    for (int i : failed) push_fallback(X, i);
}

static void run_webp(lp_xbatch* X, Lane& L, const std::vector<int>& idx) {
    const int n = (int)idx.size();
    const std::vector<Run> runs = runs_of(X, idx);
    Bump bump{L.dev, L.dev_bytes};
    std::vector<uint64_t> off((size_t)n), frame_off((size_t)n);
    std::vector<uint32_t> len((size_t)n);
    std::vector<int> ws((size_t)n), hs((size_t)n);
    size_t in_bytes = 0, frame_bytes = 0;
    for (int k = 0; k < n; k++) {
        const XItem& xi = X->items[idx[k]];
        off[k] = in_bytes;
        len[k] = (uint32_t)xi.webp.vp8_len;
        in_bytes += round_up((size_t)len[k] + 64, (size_t)16);
        frame_off[k] = frame_bytes;
        frame_bytes += round_up((size_t)xi.w * xi.h * 3, (size_t)256);
        ws[k] = xi.w;
        hs[k] = xi.h;
    }
    uint8_t* d_in = bump.take<uint8_t>(in_bytes + 4096);
    uint8_t* d_frames = bump.take<uint8_t>(frame_bytes + 256);
    std::vector<int> st((size_t)n, 0), failed;
    bool ok = d_in && d_frames;
    for (int k = 0; k < n && ok; k++) {
        ok = cudaMemcpyAsync(d_in + off[k], X->in[idx[k]] + X->items[idx[k]].webp.vp8_off, len[k], cudaMemcpyHostToDevice,
                             L.st) == cudaSuccess;
        L.h2d += len[k];
    }
    cudaEventRecord(L.ev[0], L.st);
    if (ok) ok = webp_vp8_decode_batch(d_in, off.data(), len.data(), n, ws.data(), hs.data(), d_frames, frame_off.data(), st.data(), L.st) == LP_OK;
    if (!ok) {
        cudaGetLastError();
        for (int i : idx) push_fallback(X, i);
        return;
    }
    std::vector<char> good((size_t)n);
    for (int k = 0; k < n; k++) good[k] = st[k] == 0;
    resize_and_encode(X, L, bump, idx, runs, d_frames, frame_off, good, &failed);
    for (int i : failed) push_fallback(X, i);
}

// ------------------------------------------------------------------ GIF groups (animations -> animated WebP)

static void run_gif(lp_xbatch* X, Lane& L, const std::vector<int>& idx) {
    const int na = (int)idx.size();
    const XItem& g = X->items[idx[0]];
    const int w = g.w, h = g.h, ch = 4;
    Bump bump{L.dev, L.dev_bytes};
    const size_t canvas_stride = round_up((size_t)w * h * 4, (size_t)256);
    const size_t out_stride = round_up((size_t)g.ow * g.oh * ch, (size_t)256);
    std::vector<int> first((size_t)na + 1, 0);
    size_t scratch_bytes = 0;
    std::vector<GifAnimPlan*> plans((size_t)na);
    std::vector<const uint8_t*> files((size_t)na);
    std::vector<size_t> flen((size_t)na);
    for (int a = 0; a < na; a++) {
        const XItem& it = X->items[idx[a]];
        first[a + 1] = first[a] + it.gif_frames;
        plans[a] = it.gif;
        files[a] = X->in[idx[a]];
        flen[a] = X->in_len[idx[a]];
        scratch_bytes += gif_plan_device_bytes(it.gif);
        L.h2d += flen[a];
    }
    const int nf = first[na];
    uint8_t* d_scratch = bump.take<uint8_t>(scratch_bytes + 4096);
    uint8_t* d_canvases = bump.take<uint8_t>((size_t)nf * canvas_stride + 256);
    uint8_t* d_resized = bump.take<uint8_t>((size_t)nf * out_stride + 256);
    std::vector<int> st((size_t)na, 0);
    bool ok = d_scratch && d_canvases && d_resized;
    cudaEventRecord(L.ev[0], L.st);
    if (ok)
        ok = gif_decode_batch(plans.data(), files.data(), flen.data(), na, d_scratch, scratch_bytes + 4096, d_canvases,
                              canvas_stride, first.data(), st.data(), L.st) == LP_OK;
    cudaEventRecord(L.ev[1], L.st);
    if (ok) {
        ResizeArgs r{d_canvases, canvas_stride, (size_t)w * 4, 4, g.cx, g.cy, g.cw, g.chh, d_resized, out_stride,
                     (size_t)g.ow * 4, g.ow, g.oh, nf, 3};
        ok = resize_launch(r, L.st) == LP_OK;
    }
    cudaEventRecord(L.ev[2], L.st);
    if (ok) ok = cudaStreamSynchronize(L.st) == cudaSuccess;
    if (!ok) {
        cudaGetLastError();
        for (int i : idx) push_fallback(X, i);
        return;
    }
    lane_time(L, 0, 1, &L.ms_decode);
    lane_time(L, 1, 2, &L.ms_resize);
    cudaEventRecord(L.ev[2], L.st);
    std::vector<WebpEncodedFrame> frames;
    int rc = webp_encode_lossy_batch(d_resized, out_stride, (size_t)g.ow * 4, g.ow, g.oh, 4, nf, X->quality, &frames, L.st);
    cudaEventRecord(L.ev[3], L.st);
    cudaEventSynchronize(L.ev[3]);
    lane_time(L, 2, 3, &L.ms_encode);
    if (rc) {
        cudaGetLastError();
        for (int i : idx) push_fallback(X, i);
        return;
    }
    std::vector<uint8_t> file;
    for (int a = 0; a < na; a++) {
        const int i = idx[a];
        bool good = st[a] == 0;
        for (int f = first[a]; f < first[a + 1] && good; f++) good = !frames[f].image.empty();
        if (!good) {
            push_fallback(X, i);
            continue;
        }
        uint32_t bg = 0xFFFFFFFFu;
        int loops = 0;
        gif_plan_info(X->items[i].gif, nullptr, nullptr, nullptr, &bg, &loops);
        for (int f = first[a]; f < first[a + 1]; f++) {
            frames[f].duration = gif_plan_delay_ms(X->items[i].gif, f - first[a]);
            L.d2h += frames[f].image.size() + frames[f].alph.size();
        }
        webp_assemble(&frames[first[a]], first[a + 1] - first[a], nullptr, 0, bg, (uint32_t)loops, &file);
        if (file.size() > X->out_cap) {
            X->status[i] = LP_ERR_INVALID_IMAGE;
            X->out_len[i] = 0;
            continue;
        }
        memcpy(X->out[i], file.data(), file.size());
        X->out_len[i] = file.size();
        X->status[i] = LP_OK;
    }
}

// ------------------------------------------------------------------ JPEG groups (the lp_batch pipeline)

static void run_jpeg(lp_xbatch* X, Lane& L, const std::vector<int>& idx) {
    const int n = (int)idx.size();
    const XItem& g = X->items[idx[0]];
    size_t in_bytes = 0;
    for (int i : idx) in_bytes += X->in_len[i];
    lp_batch_config c;
    memset(&c, 0, sizeof(c));
    c.device = X->device;
    c.max_images = n;
    c.src_width = g.w;
    c.src_height = g.h;
    c.dst_width = X->opt.width;
    c.dst_height = X->opt.height;
    c.resize_method = X->opt.resize_method;
    c.jpeg_quality = X->quality;
    c.max_in_bytes = in_bytes + (1 << 20);
    // slot per output: never larger than the callers' buffers (lp_batch copies a whole result into out[i])
    c.out_cap = std::min(X->out_cap, round_up(std::max((size_t)65536, (size_t)g.ow * g.oh * 3), (size_t)256));
    if (c.out_cap >= 256) c.out_cap = c.out_cap / 256 * 256;
    // images per chunk: whole Huffman waves while the per-chunk scratch fits the lane's arena
    const size_t mcus = (size_t)ceil_div(g.w, 8) * ceil_div(g.h, 8);
    const size_t per_img = (mcus * 3 + 64) * (128 + 64 + 2) + (size_t)g.w * g.h * 3 + 65536;
    const size_t fixed = 2 * in_bytes + (size_t)n * (c.out_cap + (size_t)g.ow * g.oh * 3 + 4096) + (64u << 20);
    const int slots = jpeg_huff_parallel_slots();
    long fit = L.dev_bytes > fixed ? (long)((L.dev_bytes - fixed) / per_img) : 0;
    if (fit < 1) {
        for (int i : idx) push_fallback(X, i);
        return;
    }
    int chunk = (int)std::min<long>(fit, 3L * std::max(slots, 1));
    if (slots > 0 && chunk > slots) chunk = chunk / slots * slots;
    c.chunk = std::max(1, std::min(chunk, n));
    lp_batch* b = batch_create_in(&c, L.dev, L.dev_bytes, L.host, L.host_bytes);
    if (!b) {
        for (int i : idx) push_fallback(X, i);
        return;
    }
    std::vector<const uint8_t*> in((size_t)n);
    std::vector<size_t> len((size_t)n), olen((size_t)n, 0);
    std::vector<uint8_t*> out((size_t)n);
    std::vector<int> st((size_t)n, 0);
    for (int k = 0; k < n; k++) {
        in[k] = X->in[idx[k]];
        len[k] = X->in_len[idx[k]];
        out[k] = X->out[idx[k]];
    }
    cudaEventRecord(L.ev[0], L.st);
    const int rc = lp_batch_transform(b, in.data(), len.data(), n, out.data(), olen.data(), st.data());
    L.h2d += in_bytes;
    for (int k = 0; k < n; k++) {
        const int i = idx[k];
        if (rc || st[k] != LP_OK) {
            push_fallback(X, i);  // whatever the batch pipeline would not take: Transform decides
            continue;
        }
        X->out_len[i] = olen[k];
        X->status[i] = LP_OK;
        L.d2h += olen[k];
    }
    lp_batch_destroy(b);
}

// ------------------------------------------------------------------ the call

extern "C" lp_xbatch* lp_xbatch_create(const lp_xbatch_config* cfg) {
    if (!cfg) return nullptr;
    if (ensure_device()) return nullptr;
    int prev = 0;
    cudaGetDevice(&prev);
    if (cudaSetDevice(cfg->device) != cudaSuccess) return nullptr;
    lp_xbatch* X = new lp_xbatch;
    X->cfg = *cfg;
    X->device = cfg->device;
    memset(&X->stats, 0, sizeof(X->stats));
    unsigned hc = std::thread::hardware_concurrency();
    X->threads = cfg->host_threads > 0 ? cfg->host_threads : (int)std::min(16u, std::max(2u, hc));
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    size_t arena = cfg->arena_bytes ? cfg->arena_bytes : (size_t)(free_b * 0.72);
    arena = arena / 2 / 4096 * 4096 * 2;
    X->host_bytes = (size_t)2 << 30;
    bool ok = cudaMalloc(&X->arena, arena) == cudaSuccess && cudaMallocHost(&X->host_arena, X->host_bytes) == cudaSuccess;
    X->arena_bytes = arena;
    for (int l = 0; l < 2 && ok; l++) {
        Lane& L = X->lanes[l];
        L.id = l;
        L.dev = X->arena + (size_t)l * (arena / 2);
        L.dev_bytes = arena / 2;
        L.host = X->host_arena + (size_t)l * (X->host_bytes / 2);
        L.host_bytes = X->host_bytes / 2;
        ok = cudaStreamCreateWithFlags(&L.st, cudaStreamNonBlocking) == cudaSuccess;
        for (int e = 0; e < 4 && ok; e++) ok = cudaEventCreate(&L.ev[e]) == cudaSuccess;
    }
    cudaSetDevice(prev);
    if (!ok) {
        fprintf(stderr, "[lilliput_b200] lp_xbatch_create: device arena (%zu B) or pinned staging allocation failed\n", arena);
        cudaGetLastError();
        lp_xbatch_destroy(X);
        return nullptr;
    }
    return X;
}

extern "C" void lp_xbatch_destroy(lp_xbatch* X) {
    if (!X) return;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(X->device);
    cudaDeviceSynchronize();
    for (Lane& L : X->lanes) {
        for (auto& e : L.ev)
            if (e) cudaEventDestroy(e);
        if (L.st) cudaStreamDestroy(L.st);
    }
    if (X->arena) cudaFree(X->arena);
    if (X->host_arena) cudaFreeHost(X->host_arena);
    cudaSetDevice(prev);
    delete X;
}

extern "C" void lp_xbatch_get_stats(const lp_xbatch* X, lp_xbatch_stats* out) {
    if (X && out) *out = X->stats;
}

template <class F>
static void parallel_for(int n, int threads, F&& fn) {
    if (n <= 0) return;
    threads = std::max(1, std::min(threads, n));
    if (threads == 1) {
        for (int i = 0; i < n; i++) fn(i);
        return;
    }
    std::atomic<int> next{0};
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
        pool.emplace_back([&]() {
            for (;;) {
                const int i = next.fetch_add(1);
                if (i >= n) break;
                fn(i);
            }
        });
    for (auto& th : pool) th.join();
}

// device bytes one item of a group needs inside a lane's arena (upper bound, sinks included)
static size_t item_device_bytes(const lp_xbatch* X, const XItem& it, int i) {
    const size_t outb = (size_t)it.ow * it.oh * 4 * 3 + (256u << 10);
    switch (it.kind) {
        case K_PNG: {
            // compressed span (+ its gathered copy) + inflated scanlines + resized output; the packed frames live in a
            // window buffer shared by the task (a fifth of the lane's arena, reserved by split_by_memory)
            const size_t raw = ((size_t)it.w * (it.ch == 4 ? 4 : 3) + 2) * it.h * (it.png && it.png->interlace ? 2 : 1);
            return 2 * X->in_len[i] + raw + outb + 8192;
        }
        case K_WEBP:
            return X->in_len[i] + (size_t)it.w * it.h * 3 + outb + 4096;
        case K_GIF:
            return gif_plan_device_bytes(it.gif) + (size_t)it.gif_frames * ((size_t)it.w * it.h * 4 + outb) + 8192;
        default:
            return 0;
    }
}

extern "C" int lp_xbatch_transform(lp_xbatch* X, const uint8_t* const* in, const size_t* in_len, int n,
                                   const lp_image_options* opt, uint8_t* const* out, size_t out_cap, size_t* out_len,
                                   int* status) {
    if (!X || n < 0 || !opt || (n > 0 && (!in || !in_len || !out || !out_len || !status))) return LP_ERR_BAD_ARGUMENT;
    int prev = 0;
    cudaGetDevice(&prev);
    LP_CUDA_OK(cudaSetDevice(X->device));
    const auto t0 = std::chrono::steady_clock::now();
    X->in = in;
    X->in_len = in_len;
    X->out = out;
    X->out_cap = out_cap;
    X->out_len = out_len;
    X->status = status;
    X->opt = *opt;
    X->items.clear();
    X->items.resize((size_t)n);
    X->fallback.clear();
    memset(&X->stats, 0, sizeof(X->stats));
    for (Lane& L : X->lanes) {
        L.ms_decode = L.ms_resize = L.ms_encode = 0;
        L.h2d = L.d2h = 0;
        L.launches = 0;
    }
    for (int i = 0; i < n; i++) {
        status[i] = LP_ERR_UNSUPPORTED;  // every item is overwritten by its group or by the fallback
        out_len[i] = 0;
    }
    // which sink the options ask for (ref lilliput.go:176-195 NewEncoder by extension)
    std::string ext = opt->file_type ? opt->file_type : "";
    for (auto& c : ext) c = (char)tolower((unsigned char)c);
    X->sink = S_NONE;
    if (ext == ".jpeg" || ext == ".jpg") {
        X->sink = S_JPEG;
        X->quality = option_value(*opt, CV_IMWRITE_JPEG_QUALITY, 95);  // OpenCV's default
        if (option_value(*opt, 2 /* JpegProgressive */, 0)) X->sink = S_NONE;
    } else if (ext == ".webp") {
        const int q = option_value(*opt, CV_IMWRITE_WEBP_QUALITY, 100);
        X->quality = q < 1 ? 1 : q;
        X->sink = q > 100 ? S_NONE : S_WEBP;  // lossless output: per image
    }
    parallel_for(n, X->threads, [&](int i) { parse_item(X, i); });
    const auto t1 = std::chrono::steady_clock::now();
    // groups -> tasks that fit a lane
    std::map<std::tuple<int, int, int, int, int>, std::vector<int>> groups;
    for (int i = 0; i < n; i++) {
        const XItem& it = X->items[i];
        if (it.kind == K_FALLBACK) X->fallback.push_back(i);
        else groups[std::make_tuple((int)it.kind, it.w, it.h, it.ch, it.jpeg_sampling)].push_back(i);
    }
    std::vector<Task> tasks;
    std::vector<double> cost;  // rough device time: the longest tasks start first
    const size_t lane_cap = X->lanes[0].dev_bytes;
    // PNG and WebP: all geometries of a kind in as few tasks as the arena allows (the map keeps equal
    // geometries adjacent); at least two, so both lanes work.  GIF: by canvas size.  JPEG: by geometry.
    std::map<int, std::vector<int>> merged;
    for (auto& kv : groups) {
        const Kind kind = (Kind)std::get<0>(kv.first);
        if (kind == K_PNG || kind == K_WEBP) merged[(int)kind].insert(merged[(int)kind].end(), kv.second.begin(), kv.second.end());
    }
    auto split_by_memory = [&](Kind kind, const std::vector<int>& g) {
        size_t total = 0;
        for (int i : g) total += item_device_bytes(X, X->items[i], i);
        const size_t half = total / 2 + 1;
        Task cur{kind, {}};
        const size_t reserve = kind == K_PNG ? lane_cap / 5 + (64u << 20) : (64u << 20);  // PNG: the frame window
        size_t used = reserve;
        for (int i : g) {
            const size_t need = item_device_bytes(X, X->items[i], i) +
                                (kind == K_PNG ? 0 : 0);
            if (kind == K_PNG && round_up((size_t)X->items[i].w * X->items[i].h * X->items[i].ch, (size_t)256) > std::min<size_t>((size_t)12 << 30, lane_cap / 5)) {
                X->fallback.push_back(i);  // one frame larger than the window
                continue;
            }
            if (need + reserve > lane_cap) {
                X->fallback.push_back(i);
                continue;
            }
            if (!cur.idx.empty() && (used + need > lane_cap || used > half + reserve)) {
                tasks.push_back(cur);
                cost.push_back((double)used);
                cur.idx.clear();
                used = reserve;
            }
            cur.idx.push_back(i);
            used += need;
        }
        if (!cur.idx.empty()) {
            tasks.push_back(cur);
            cost.push_back((double)used);
        }
    };
    for (auto& kv : merged) split_by_memory((Kind)kv.first, kv.second);
    for (auto& kv : groups) {
        const Kind kind = (Kind)std::get<0>(kv.first);
        const std::vector<int>& g = kv.second;
        if (kind == K_PNG || kind == K_WEBP) continue;
        if (kind == K_JPEG) {
            // host staging bounds a JPEG task: out slots + item mirrors come from the lane's pinned arena
            const XItem& it0 = X->items[g[0]];
            const size_t slot = round_up(std::min(out_cap, std::max((size_t)65536, (size_t)it0.ow * it0.oh * 3)), (size_t)256) +
                                sizeof(JpegDecodeItem) + 64;
            const size_t per = std::max<size_t>(1, X->lanes[0].host_bytes / slot);
            const size_t want = std::min(per, std::max<size_t>(1, (g.size() + 1) / 2));  // two tasks at least: both lanes work
            for (size_t a = 0; a < g.size(); a += want) {
                tasks.push_back(Task{kind, std::vector<int>(g.begin() + a, g.begin() + std::min(g.size(), a + want))});
                cost.push_back((double)tasks.back().idx.size() * it0.w * it0.h * 0.05);
            }
            continue;
        }
        split_by_memory(kind, g);
    }
    {   // longest first
        std::vector<int> order(tasks.size());
        for (size_t t = 0; t < order.size(); t++) order[t] = (int)t;
        std::sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
        std::vector<Task> sorted;
        for (int t : order) sorted.push_back(std::move(tasks[t]));
        tasks.swap(sorted);
    }
    X->stats.groups = (int)groups.size();
    // two lanes drain the task list
    std::atomic<int> next{0};
    const long launches0 = g_launches;
    std::atomic<long> lane_launches{0};
    auto lane_main = [&](int l) {
        cudaSetDevice(X->device);
        Lane& L = X->lanes[l];
        const long mine0 = g_launches;
        for (;;) {
            const int t = next.fetch_add(1);
            if (t >= (int)tasks.size()) break;
            const Task& task = tasks[t];
            switch (task.kind) {
                case K_JPEG: run_jpeg(X, L, task.idx); break;
                case K_PNG: run_png(X, L, task.idx); break;
                case K_WEBP: run_webp(X, L, task.idx); break;
                case K_GIF: run_gif(X, L, task.idx); break;
                default: for (int i : task.idx) push_fallback(X, i); break;
            }
        }
        lane_launches += g_launches - mine0;
    };
    {
        std::thread other(lane_main, 1);
        lane_main(0);
        other.join();
    }
    (void)launches0;
    const auto t2 = std::chrono::steady_clock::now();
    // everything else, one image per call, a few host threads (each has its own stream)
    const int max_size = X->cfg.max_size > 0 ? X->cfg.max_size : 8192;
    std::atomic<long> fb_launches{0};
    {
        std::vector<int> fb = X->fallback;
        parallel_for((int)fb.size(), std::min(X->threads, 8), [&](int k) {
            cudaSetDevice(X->device);
            const long l0 = g_launches;
            const int i = fb[k];
            size_t len = 0;
            status[i] = lp_transform(in[i], in_len[i], opt, out[i], out_cap, &len, max_size);
            out_len[i] = status[i] == LP_OK ? len : 0;
            fb_launches += g_launches - l0;
        });
        X->stats.fallback_items = (int)fb.size();
    }
    for (XItem& it : X->items)
        if (it.gif) {
            gif_plan_free(it.gif);
            it.gif = nullptr;
        }
    const auto t3 = std::chrono::steady_clock::now();
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double, std::milli>(b - a).count();
    };
    X->stats.grid_items = n - X->stats.fallback_items;
    X->stats.ms_parse = ms(t0, t1);
    X->stats.ms_grid = ms(t1, t2);
    X->stats.ms_fallback = ms(t2, t3);
    X->stats.ms_total = ms(t0, t3);
    for (Lane& L : X->lanes) {
        X->stats.ms_decode += L.ms_decode;
        X->stats.ms_resize += L.ms_resize;
        X->stats.ms_encode += L.ms_encode;
        X->stats.h2d_bytes += L.h2d;
        X->stats.d2h_bytes += L.d2h;
        X->stats.launches += (int)L.launches;
        X->stats.ms_busy_max_lane = std::max(X->stats.ms_busy_max_lane, L.ms_decode + L.ms_resize + L.ms_encode);
    }
    X->stats.launches += (int)(lane_launches.load() + fb_launches.load());
    cudaSetDevice(prev);
    return LP_OK;
}

// ------------------------------------------------------------------ library-level multi-GPU dispatch
// images are independent units with zero exchange -- the batch is cut into one contiguous block per
// GPU, balanced by compressed bytes, and each block runs through that GPU's own lp_xbatch (own arena, pinned
// staging, streams) on its own host thread; results land in the caller's arrays by index.  No collective, no
// peer traffic.  (bench.py's torchrun harness is the process-per-GPU form of the same sharding.)

struct lp_multi {
    std::vector<lp_xbatch*> ctx;
    std::vector<int> devices;
    std::vector<int> first;  // block boundaries of the last call (ctx.size() + 1)
};

extern "C" lp_multi* lp_multi_create(const int* devices, int n_devices, const lp_xbatch_config* tmpl) {
    if (n_devices < 1 || !devices) return nullptr;
    lp_multi* m = new lp_multi;
    for (int g = 0; g < n_devices; g++) {
        lp_xbatch_config c;
        memset(&c, 0, sizeof(c));
        if (tmpl) c = *tmpl;
        c.device = devices[g];
        lp_xbatch* x = lp_xbatch_create(&c);
        if (!x) {
            for (lp_xbatch* y : m->ctx) lp_xbatch_destroy(y);
            delete m;
            return nullptr;
        }
        m->ctx.push_back(x);
        m->devices.push_back(devices[g]);
    }
    return m;
}

extern "C" void lp_multi_destroy(lp_multi* m) {
    if (!m) return;
    for (lp_xbatch* x : m->ctx) lp_xbatch_destroy(x);
    delete m;
}

extern "C" int lp_multi_device_count(const lp_multi* m) { return m ? (int)m->ctx.size() : 0; }

// Host-only: cut n items into `parts` contiguous blocks with (nearly) equal compressed bytes ("contiguous
// blocks balanced by compressed bytes"); first[p] .. first[p + 1] is block p, first has parts + 1 entries.
extern "C" void lp_shard_blocks(const size_t* in_len, int n, int parts, int* first) {
    if (parts < 1 || !first) return;
    for (int p = 0; p <= parts; p++) first[p] = n < 0 ? 0 : n;
    first[0] = 0;
    if (n <= 0 || !in_len) return;
    size_t total = 0;
    for (int i = 0; i < n; i++) total += in_len[i] + 4096;  // (+ a per-image constant: tiny files still cost a launch slot)
    size_t acc = 0;
    int g = 1;
    for (int i = 0; i < n && g < parts; i++) {
        acc += in_len[i] + 4096;
        while (g < parts && acc * (size_t)parts >= total * (size_t)g) first[g++] = i + 1;
    }
}

extern "C" int lp_multi_transform(lp_multi* m, const uint8_t* const* in, const size_t* in_len, int n,
                                  const lp_image_options* opt, uint8_t* const* out, size_t out_cap, size_t* out_len,
                                  int* status) {
    if (!m || n < 0 || !opt || (n > 0 && (!in || !in_len || !out || !out_len || !status))) return LP_ERR_BAD_ARGUMENT;
    const int G = (int)m->ctx.size();
    m->first.assign((size_t)G + 1, n);
    lp_shard_blocks(in_len, n, G, m->first.data());
    std::vector<int> rc((size_t)G, LP_OK);
    std::vector<std::thread> pool;
    for (int g = 0; g < G; g++) {
        const int i0 = m->first[g], cnt = m->first[g + 1] - m->first[g];
        if (cnt <= 0) continue;
        pool.emplace_back([=, &rc]() {
            rc[g] = lp_xbatch_transform(m->ctx[g], in + i0, in_len + i0, cnt, opt, out + i0, out_cap, out_len + i0, status + i0);
        });
    }
    for (auto& t : pool) t.join();
    for (int g = 0; g < G; g++)
        if (rc[g]) return rc[g];
    return LP_OK;
}

// per-device statistics of the last call
extern "C" void lp_multi_get_stats(const lp_multi* m, int device_index, lp_xbatch_stats* out) {
    if (m && out && device_index >= 0 && device_index < (int)m->ctx.size()) lp_xbatch_get_stats(m->ctx[device_index], out);
}
