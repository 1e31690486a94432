// webp_decode.cu -- lilliput's WebP decoder surface (include/lp_webp.h = ref webp.hpp:35-51,74-75)
// on sm_100a: host RIFF walk, device VP8 key-frame decode, device upsample + colour conversion.
//
// Replaces: webp_decoder_* (ref webp.cpp:61-370), i.e. libwebpmux's chunk walk
// (WebPMuxCreate / GetFeatures / GetFrame / GetCanvasSize / GetAnimationParams / GetChunk "ICCP")
// and libwebp's WebPDecodeBGRInto / WebPDecodeBGRAInto.  The VP8 decoding logic lives in
// vp8_core.h (shared with the CPU test harness); this file holds the kernels and the ABI.
//
// VP8 on a GPU: one frame's mode bits and coefficient tokens are two serial arithmetic-coded
// streams whose contexts chain through the picture, so the unit of parallelism is the frame: one
// warp per frame.  Lane 0 walks the bitstream and reconstructs macroblocks into the frame's YUV
// planes in HBM; the loop filter then runs warp-wide (32 lanes = the 16 luma + 8 + 8 chroma sample
// positions of one macroblock edge); a second, fully parallel kernel does libwebp's "fancy"
// chroma upsampling and the fixed-point YUV->BGR(A) conversion per output pixel.
#include <cstring>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"
#include "lp_webp.h"

#define LP_VP8_FN static __device__
#define LP_VP8_HD static __host__ __device__
#define LP_VP8_TABLE static __device__ const
#include "vp8_core.h"

namespace lp {

// ------------------------------------------------------------------ kernels

struct Vp8Item {
    const uint8_t* data;  // VP8 payload (frame tag onwards), device
    uint32_t size;
    uint8_t* work;        // vp8::work_bytes(mb_w, mb_h)
    int* status;          // 0 ok
    int mb_w, mb_h;       // from the host's look at the 10-byte frame header
};

// One macroblock's edges, warp-wide.  Lanes 0..15 take the luma sample positions of an edge,
// 16..23 the U and 24..31 the V positions.  Order of edges follows RFC 6386 s.15.
__device__ void filter_macroblock_warp(const vp8::FrameHdr& h, vp8::Work& w, int mb_x, int mb_y, int lane) {
    const uint32_t fi = w.finfo[mb_y * h.mb_w + mb_x];
    const int limit = fi & 255, ilevel = (fi >> 8) & 255, hev_t = (fi >> 16) & 255, inner = fi >> 24;
    if (limit == 0) return;
    const int ys = h.mb_w * 16, cs = h.mb_w * 8;
    const int simple = h.filter_type == 1;
    const bool luma = lane < 16;
    uint8_t* base;
    int stride, pos;
    if (luma) {
        base = w.y + (size_t)mb_y * 16 * ys + mb_x * 16;
        stride = ys;
        pos = lane;
    } else {
        base = (lane < 24 ? w.u : w.v) + (size_t)mb_y * 8 * cs + mb_x * 8;
        stride = cs;
        pos = lane & 7;
    }
    const bool active = luma || !simple;
    // vertical edges (filter crosses columns): sample position = row `pos`
    if (mb_x > 0) {
        if (active) {
            uint8_t* p = base + (size_t)pos * stride;
            if (simple) vp8::filter_pos_simple(p, 1, limit + 4);
            else vp8::filter_pos_normal(p, 1, limit + 4, ilevel, hev_t, 1);
        }
        __syncwarp();
    }
    if (inner) {
        for (int k = 4; k < 16; k += 4) {
            if (active && (luma || k == 4)) {
                uint8_t* p = base + (size_t)pos * stride + k;
                if (simple) vp8::filter_pos_simple(p, 1, limit);
                else vp8::filter_pos_normal(p, 1, limit, ilevel, hev_t, 0);
            }
            __syncwarp();
        }
    }
    // horizontal edges (filter crosses rows): sample position = column `pos`
    if (mb_y > 0) {
        if (active) {
            uint8_t* p = base + pos;
            if (simple) vp8::filter_pos_simple(p, stride, limit + 4);
            else vp8::filter_pos_normal(p, stride, limit + 4, ilevel, hev_t, 1);
        }
        __syncwarp();
    }
    if (inner) {
        for (int k = 4; k < 16; k += 4) {
            if (active && (luma || k == 4)) {
                uint8_t* p = base + (size_t)k * stride + pos;
                if (simple) vp8::filter_pos_simple(p, stride, limit);
                else vp8::filter_pos_normal(p, stride, limit, ilevel, hev_t, 0);
            }
            __syncwarp();
        }
    }
}

constexpr int kVp8WarpsPerBlock = 4;

__global__ void __launch_bounds__(kVp8WarpsPerBlock * 32) vp8_decode_kernel(const Vp8Item* items, int n) {
    __shared__ vp8::FrameHdr s_hdr[kVp8WarpsPerBlock];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int idx = blockIdx.x * kVp8WarpsPerBlock + wid;
    if (idx >= n) return;
    const Vp8Item it = items[idx];
    vp8::FrameHdr& h = s_hdr[wid];
    vp8::Work w;
    vp8::work_carve(it.work, it.mb_w, it.mb_h, w);
    int st = 0;
    if (lane == 0) {
        vp8::BoolDec br;
        st = vp8::parse_frame_header(it.data, it.size, h, br, w.proba);
        if (!st && (h.mb_w != it.mb_w || h.mb_h != it.mb_h)) st = vp8::VP8_BAD;
        if (!st) st = vp8::decode_macroblocks(it.data, h, br, w);
        *it.status = st;
    }
    st = __shfl_sync(0xffffffffu, st, 0);
    __syncwarp();
    if (st || h.filter_type == 0) return;
    for (int mb_y = 0; mb_y < h.mb_h; mb_y++)
        for (int mb_x = 0; mb_x < h.mb_w; mb_x++) filter_macroblock_warp(h, w, mb_x, mb_y, lane);
}

struct Vp8Output {
    const uint8_t *y, *u, *v;
    const uint8_t* alpha;  // width x height plane or null
    int ys, cs, width, height;
    uint8_t* dst;
    size_t dst_step;
    int channels;  // 3 or 4
};

__global__ void vp8_output_kernel(Vp8Output o) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= o.width) return;
    const int u = vp8::upsample_at(o.u, o.cs, o.width, o.height, x, y);
    const int v = vp8::upsample_at(o.v, o.cs, o.width, o.height, x, y);
    uint8_t bgr[3];
    vp8::yuv_to_bgr(o.y[(size_t)y * o.ys + x], u, v, bgr);
    uint8_t* d = o.dst + (size_t)y * o.dst_step + (size_t)x * o.channels;
    d[0] = bgr[0];
    d[1] = bgr[1];
    d[2] = bgr[2];
    if (o.channels == 4) d[3] = o.alpha ? o.alpha[(size_t)y * o.width + x] : 255;
}

// ------------------------------------------------------------------ container walk (host)
// RIFF layout per the WebP container specification; checks follow what WebPMuxCreate /
// MuxValidate reject (ref webp.cpp:65-69 treats a NULL mux as "not a WebP").

static inline uint32_t le16(const uint8_t* p) { return p[0] | (p[1] << 8); }
static inline uint32_t le24(const uint8_t* p) { return p[0] | (p[1] << 8) | ((uint32_t)p[2] << 16); }
static inline uint32_t le32(const uint8_t* p) { return le24(p) | ((uint32_t)p[3] << 24); }

enum { kFlagAnim = 0x02, kFlagXmp = 0x04, kFlagExif = 0x08, kFlagAlpha = 0x10, kFlagIcc = 0x20 };
constexpr uint32_t kMaxChunkPayload = ~0u - 8 - 1;

struct WebpFrame {
    size_t img_off = 0, img_len = 0;  // VP8 / VP8L payload
    bool lossless = false;
    size_t alph_off = 0, alph_len = 0;
    bool has_alph = false;
    int x_off = 0, y_off = 0, width = 0, height = 0;
    int duration = 0, dispose = 0, blend = 0;
    bool has_alpha = false;  // ALPH chunk, or VP8L header alpha bit
};

struct WebpContainer {
    int canvas_w = 0, canvas_h = 0;
    uint32_t flags = 0;
    bool has_vp8x = false, has_anim_chunk = false;
    uint32_t bgcolor = 0xFFFFFFFFu, loop_count = 0;
    size_t icc_off = 0, icc_len = 0;
    bool has_icc = false, has_exif = false, has_xmp = false;
    std::vector<WebpFrame> frames;
};

// Validates an image payload and fills the frame's size (VP8GetInfo / VP8LGetInfo equivalents).
static bool image_info(const uint8_t* p, size_t n, bool lossless, WebpFrame* f) {
    if (!lossless) {
        if (n < 10) return false;
        const uint32_t tag = le24(p);
        if (tag & 1) return false;                       // not a key frame
        if (((tag >> 1) & 7) > 3) return false;          // unknown profile
        if (!((tag >> 4) & 1)) return false;             // invisible frame
        if ((tag >> 5) >= n) return false;               // partition_length beyond the chunk
        if (p[3] != 0x9d || p[4] != 0x01 || p[5] != 0x2a) return false;
        f->width = le16(p + 6) & 0x3fff;
        f->height = le16(p + 8) & 0x3fff;
        return f->width > 0 && f->height > 0;
    }
    if (n < 5 || p[0] != 0x2f) return false;
    const uint32_t bits = le32(p + 1);
    f->width = (int)(bits & 0x3fff) + 1;
    f->height = (int)((bits >> 14) & 0x3fff) + 1;
    f->has_alpha = (bits >> 28) & 1;
    return ((bits >> 29) & 7) == 0;  // version
}

// Walks the sub-chunks that make up one image (ALPH? then VP8 / VP8L) in [pos, end).
static bool parse_image_chunks(const uint8_t* b, size_t pos, size_t end, WebpFrame* f, bool* got_image) {
    *got_image = false;
    while (pos + 8 <= end) {
        const uint32_t n = le32(b + pos + 4);
        if (n > kMaxChunkPayload) return false;
        const size_t padded = 8 + (((size_t)n + 1) & ~(size_t)1);
        if (pos + padded > end && pos + 8 + n > end) return false;
        const uint8_t* tag = b + pos;
        if (!memcmp(tag, "ALPH", 4)) {
            if (f->has_alph || *got_image) return false;
            f->has_alph = true;
            f->alph_off = pos + 8;
            f->alph_len = n;
        } else if (!memcmp(tag, "VP8 ", 4) || !memcmp(tag, "VP8L", 4)) {
            if (*got_image) return false;
            f->lossless = tag[3] == 'L';
            f->img_off = pos + 8;
            f->img_len = n;
            if (!image_info(b + f->img_off, f->img_len, f->lossless, f)) return false;
            if (f->has_alph && !f->lossless) f->has_alpha = true;
            *got_image = true;
        }
        pos += padded;
    }
    return true;
}

static bool webp_parse(const uint8_t* b, size_t size, WebpContainer* c) {
    if (size < 20 || memcmp(b, "RIFF", 4) || memcmp(b + 8, "WEBP", 4)) return false;
    uint32_t riff = le32(b + 4);
    if (riff > kMaxChunkPayload) return false;
    riff = (riff + 1) & ~1u;
    if (riff < 8 || riff > size) return false;
    if (size > (size_t)riff + 8) size = (size_t)riff + 8;
    size_t pos = 12;
    WebpFrame still;       // image chunks at the top level (non-animated file)
    bool still_open = false, still_done = false;
    while (pos + 8 <= size) {
        const uint8_t* tag = b + pos;
        const uint32_t n = le32(b + pos + 4);
        if (n > kMaxChunkPayload) return false;
        const size_t padded = 8 + (((size_t)n + 1) & ~(size_t)1);
        if (padded > (size_t)riff) return false;
        if (pos + padded > size) return false;  // truncated chunk
        const uint8_t* d = b + pos + 8;
        if (!memcmp(tag, "VP8X", 4)) {
            if (c->has_vp8x || n < 10) return false;
            c->has_vp8x = true;
            c->flags = d[0];
            c->canvas_w = (int)le24(d + 4) + 1;
            c->canvas_h = (int)le24(d + 7) + 1;
        } else if (!memcmp(tag, "ICCP", 4)) {
            if (c->has_icc) return false;
            c->has_icc = true;
            c->icc_off = pos + 8;
            c->icc_len = n;
        } else if (!memcmp(tag, "EXIF", 4)) {
            if (c->has_exif) return false;
            c->has_exif = true;
        } else if (!memcmp(tag, "XMP ", 4)) {
            if (c->has_xmp) return false;
            c->has_xmp = true;
        } else if (!memcmp(tag, "ANIM", 4)) {
            if (c->has_anim_chunk || n < 6) return false;
            c->has_anim_chunk = true;
            c->bgcolor = le32(d);
            c->loop_count = le16(d + 4);
        } else if (!memcmp(tag, "ANMF", 4)) {
            if (still_open || n < 16) return false;
            WebpFrame f;
            f.x_off = 2 * (int)le24(d);
            f.y_off = 2 * (int)le24(d + 3);
            const int fw = (int)le24(d + 6) + 1, fh = (int)le24(d + 9) + 1;
            f.duration = (int)le24(d + 12);
            f.dispose = d[15] & 1;          // 1 = dispose to background
            f.blend = (d[15] >> 1) & 1;     // 1 = do not blend
            bool got = false;
            if (!parse_image_chunks(b, pos + 8 + 16, pos + 8 + n, &f, &got) || !got) return false;
            if (f.width != fw || f.height != fh) return false;
            c->frames.push_back(f);
        } else if (!memcmp(tag, "ALPH", 4)) {
            if (still_open || still_done) return false;
            still_open = true;
            still.has_alph = true;
            still.alph_off = pos + 8;
            still.alph_len = n;
        } else if (!memcmp(tag, "VP8 ", 4) || !memcmp(tag, "VP8L", 4)) {
            if (still_done) return false;
            still.lossless = tag[3] == 'L';
            still.img_off = pos + 8;
            still.img_len = n;
            if (!image_info(d, n, still.lossless, &still)) return false;
            if (still.has_alph && !still.lossless) still.has_alpha = true;
            still_open = false;
            still_done = true;
        } else {
            if (still_open) return false;  // an ALPH chunk must be followed by its image
        }
        pos += padded;
    }
    if (still_open) return false;
    if (still_done) {
        if (!c->frames.empty()) return false;  // ANMF frames and a bare image do not mix
        still.duration = 1;                    // what WebPMuxGetFrame reports for a non-animated image
        c->frames.push_back(still);
    }
    if (c->frames.empty()) return false;
    // MuxValidate: feature flags and chunks must agree
    const bool anim = c->flags & kFlagAnim;
    if (!c->has_vp8x) {
        if (c->frames.size() != 1 || c->frames[0].has_alph || c->has_icc || c->has_anim_chunk || c->has_exif ||
            c->has_xmp)
            return false;
        c->canvas_w = c->frames[0].width;
        c->canvas_h = c->frames[0].height;
        c->flags = c->frames[0].has_alpha ? kFlagAlpha : 0;
    } else {
        if (((c->flags & kFlagIcc) != 0) != c->has_icc) return false;
        if (((c->flags & kFlagExif) != 0) != c->has_exif) return false;
        if (((c->flags & kFlagXmp) != 0) != c->has_xmp) return false;
        if (anim != c->has_anim_chunk) return false;
        if (!anim && (c->frames.size() != 1 || !still_done)) return false;
        if (anim && still_done) return false;
        bool any_alpha = false;
        for (const WebpFrame& f : c->frames) any_alpha |= f.has_alpha;
        if (any_alpha && !(c->flags & kFlagAlpha)) return false;
        for (const WebpFrame& f : c->frames)
            if (f.x_off + f.width > c->canvas_w || f.y_off + f.height > c->canvas_h) return false;
    }
    return true;
}

// The mat handle is defined in abi_opencv.cu.
const uint8_t* mat_host_bytes(const void* mat, size_t* len);
int mat_bind_device_frame(void* mat, int cols, int rows, int type, uint8_t** dev, size_t* step);
void mat_mark_device_written(void* mat);

}  // namespace lp

using namespace lp;

struct webp_decoder_struct {
    const uint8_t* bytes = nullptr;
    size_t len = 0;
    WebpContainer c;
    bool has_alpha = false, has_animation = false;
    int total_duration = 0;
    int current_frame_index = 1;
    int prev_delay = 0, prev_x = 0, prev_y = 0, prev_dispose = 0, prev_blend = 0;
    bool prev_has_alpha = false;
    // device scratch, grown on demand
    uint8_t* d_in = nullptr;
    uint8_t* d_work = nullptr;
    Vp8Item* d_item = nullptr;
    int* d_status = nullptr;
    size_t in_cap = 0, work_cap = 0;
};

struct webp_encoder_struct {
    int unused;
};

extern "C" {

// ref webp.cpp:61-139
webp_decoder webp_decoder_create(const opencv_mat buf) {
    if (!buf) return nullptr;
    size_t len = 0;
    const uint8_t* bytes = mat_host_bytes(buf, &len);
    if (!bytes) return nullptr;
    auto* d = new webp_decoder_struct;
    d->bytes = bytes;
    d->len = len;
    if (!webp_parse(bytes, len, &d->c)) {
        delete d;
        return nullptr;
    }
    d->has_alpha = (d->c.flags & kFlagAlpha) != 0;
    for (const WebpFrame& f : d->c.frames) d->total_duration += f.duration;
    d->c.bgcolor = (d->c.flags & kFlagAnim) ? d->c.bgcolor : 0xFFFFFFFFu;
    if (d->c.flags & kFlagAnim) {
        d->has_animation = true;
    } else {
        d->total_duration = 0;
        d->c.loop_count = 0;
    }
    return d;
}

int webp_decoder_get_width(const webp_decoder d) { return d->c.canvas_w; }
int webp_decoder_get_height(const webp_decoder d) { return d->c.canvas_h; }
int webp_decoder_get_pixel_type(const webp_decoder d) { return d->has_alpha ? CV_8UC4 : CV_8UC3; }
int webp_decoder_get_num_frames(const webp_decoder d) { return d ? (int)d->c.frames.size() : 0; }
int webp_decoder_get_total_duration(const webp_decoder d) { return d ? d->total_duration : 0; }
int webp_decoder_get_prev_frame_delay(const webp_decoder d) { return d->prev_delay; }
int webp_decoder_get_prev_frame_dispose(const webp_decoder d) { return d->prev_dispose; }
int webp_decoder_get_prev_frame_blend(const webp_decoder d) { return d->prev_blend; }
int webp_decoder_get_prev_frame_x_offset(const webp_decoder d) { return d->prev_x; }
int webp_decoder_get_prev_frame_y_offset(const webp_decoder d) { return d->prev_y; }
bool webp_decoder_get_prev_frame_has_alpha(const webp_decoder d) { return d->prev_has_alpha; }
uint32_t webp_decoder_get_bg_color(const webp_decoder d) { return d->c.bgcolor; }
uint32_t webp_decoder_get_loop_count(const webp_decoder d) { return d->c.loop_count; }

// ref webp.cpp:251-262
size_t webp_decoder_get_icc(const webp_decoder d, void* dst, size_t dst_len) {
    if (!d->c.has_icc || d->c.icc_len == 0 || d->c.icc_len > dst_len) return 0;
    memcpy(dst, d->bytes + d->c.icc_off, d->c.icc_len);
    return d->c.icc_len;
}

// ref webp.cpp:269-281
int webp_decoder_has_more_frames(webp_decoder d) { return d->current_frame_index < (int)d->c.frames.size(); }
void webp_decoder_advance_frame(webp_decoder d) { d->current_frame_index++; }

void webp_decoder_release(webp_decoder d) {
    if (!d) return;
    cudaStream_t st = thread_stream();
    if (d->d_in) cudaFreeAsync(d->d_in, st);
    if (d->d_work) cudaFreeAsync(d->d_work, st);
    if (d->d_item) cudaFreeAsync(d->d_item, st);
    if (d->d_status) cudaFreeAsync(d->d_status, st);
    cudaStreamSynchronize(st);
    delete d;
}

// ref webp.cpp:291-359
bool webp_decoder_decode(const webp_decoder d, opencv_mat mat) {
    if (!d || !mat) return false;
    if (d->current_frame_index < 1 || d->current_frame_index > (int)d->c.frames.size()) return false;
    const WebpFrame& f = d->c.frames[d->current_frame_index - 1];
    const int type = webp_decoder_get_pixel_type(d);
    uint8_t* frame_dev = nullptr;
    size_t frame_step = 0;
    if (mat_bind_device_frame(mat, f.width, f.height, type, &frame_dev, &frame_step) != 0) return false;
    d->prev_delay = f.duration;
    d->prev_x = f.x_off;
    d->prev_y = f.y_off;
    d->prev_dispose = f.dispose;
    d->prev_blend = f.blend;
    d->prev_has_alpha = f.has_alpha;
    if (f.lossless) return false;                 // VP8L: not decoded on the device yet
    if (f.has_alph && type == CV_8UC4) return false;  // ALPH plane: not decoded yet

    cudaStream_t st = thread_stream();
    const int mb_w = (f.width + 15) >> 4, mb_h = (f.height + 15) >> 4;
    const size_t need_work = vp8::work_bytes(mb_w, mb_h);
    if (!d->d_item) {
        if (cudaMallocAsync(&d->d_item, sizeof(Vp8Item), st) != cudaSuccess) return false;
        if (cudaMallocAsync(&d->d_status, sizeof(int), st) != cudaSuccess) return false;
    }
    if (f.img_len + 16 > d->in_cap) {
        if (d->d_in) cudaFreeAsync(d->d_in, st);
        d->in_cap = f.img_len + 4096;
        if (cudaMallocAsync(&d->d_in, d->in_cap, st) != cudaSuccess) return false;
    }
    if (need_work > d->work_cap) {
        if (d->d_work) cudaFreeAsync(d->d_work, st);
        d->work_cap = need_work;
        if (cudaMallocAsync(&d->d_work, d->work_cap, st) != cudaSuccess) return false;
    }
    cudaMemcpyAsync(d->d_in, d->bytes + f.img_off, f.img_len, cudaMemcpyHostToDevice, st);
    Vp8Item item{d->d_in, (uint32_t)f.img_len, d->d_work, d->d_status, mb_w, mb_h};
    cudaMemcpyAsync(d->d_item, &item, sizeof(item), cudaMemcpyHostToDevice, st);
    vp8_decode_kernel<<<1, kVp8WarpsPerBlock * 32, 0, st>>>(d->d_item, 1);
    g_launches++;
    vp8::Work w;
    vp8::work_carve(d->d_work, mb_w, mb_h, w);
    Vp8Output o{w.y, w.u, w.v, nullptr, mb_w * 16, mb_w * 8, f.width, f.height, frame_dev, frame_step, type == CV_8UC4 ? 4 : 3};
    dim3 grid(ceil_div(f.width, 128), f.height);
    vp8_output_kernel<<<grid, 128, 0, st>>>(o);
    g_launches++;
    int status = 0;
    cudaMemcpyAsync(&status, d->d_status, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) return false;
    if (status != 0) return false;
    mat_mark_device_written(mat);
    return true;
}

// ---- encoder: not implemented (see include/lp_webp.h) ---------------------------------------
webp_encoder webp_encoder_create(void*, size_t, const void*, size_t, uint32_t, int) { return nullptr; }
size_t webp_encoder_write(webp_encoder, const opencv_mat, const int*, size_t, int, int, int, int, int) { return 0; }
void webp_encoder_release(webp_encoder) {}
size_t webp_encoder_flush(webp_encoder) { return 0; }

}  // extern "C"
