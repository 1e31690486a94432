// webp_decode.cu -- lilliput's WebP decoder surface (include/lp_webp.h = ref webp.hpp:35-51,74-75)
// on sm_100a: host RIFF walk, device VP8 / VP8L / ALPH decode, device upsample + colour conversion.
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
#define LP_VP8_INL static __device__ __forceinline__
#define LP_VP8_HD static __host__ __device__
#define LP_VP8_TABLE static __device__ const
#include "vp8_core.h"
#include "vp8l_core.h"

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

// Per-warp shared state of the decode kernel.
struct Vp8WarpSmem {
    vp8::FrameHdr hdr;
    vp8::MbInfo mb;
    alignas(16) int16_t coeffs[25 * 16];
    alignas(16) uint8_t yb[vp8::YB_SIZE];
    alignas(16) uint8_t ub[vp8::CB_SIZE];
    alignas(16) uint8_t vb[vp8::CB_SIZE];
};

// 16x16 / 8x8 prediction spread over lanes: `dst` block of `size`, this lane fills `n` pixels
// starting at (row, col).  The DC sum arrives pre-reduced.
__device__ __forceinline__ void pred_span(uint8_t* dst, int size, int mode, int row, int col, int n, int dc) {
    uint8_t* d = dst + row * vp8::BPS + col;
    if (mode == vp8::DC_PRED) {
        for (int i = 0; i < n; i++) d[i] = (uint8_t)dc;
    } else if (mode == vp8::TM_PRED) {
        const uint8_t* top = dst - vp8::BPS;
        const int l = dst[row * vp8::BPS - 1] - top[-1];
        for (int i = 0; i < n; i++) d[i] = vp8::clip8(top[col + i] + l);
    } else if (mode == vp8::V_PRED) {
        const uint8_t* top = dst - vp8::BPS;
        for (int i = 0; i < n; i++) d[i] = top[col + i];
    } else {
        const uint8_t l = dst[row * vp8::BPS - 1];
        for (int i = 0; i < n; i++) d[i] = l;
    }
}

// Sum over a group of `width` consecutive lanes (width = 16 or 32), result in every lane of it.
__device__ __forceinline__ int group_sum(int v, int width) {
    for (int o = width >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Reconstructs one parsed macroblock with the whole warp (same arithmetic as vp8::reconstruct_mb).
__device__ void reconstruct_mb_warp(Vp8WarpSmem& sm, vp8::Work& w, int mb_x, int mb_y, int lane) {
    using namespace vp8;
    const FrameHdr& h = sm.hdr;
    const MbInfo& mb = sm.mb;
    const int mb_w = h.mb_w, ys = mb_w * 16, cs = mb_w * 8;
    uint8_t* yd = sm.yb + BPS + 8;
    uint8_t* ud = sm.ub + BPS + 8;
    uint8_t* vd = sm.vb + BPS + 8;
    uint8_t* py = w.y + (size_t)mb_y * 16 * ys + mb_x * 16;
    uint8_t* pu = w.u + (size_t)mb_y * 8 * cs + mb_x * 8;
    uint8_t* pv = w.v + (size_t)mb_y * 8 * cs + mb_x * 8;
    const bool have_top = mb_y > 0, have_left = mb_x > 0;
    // ---- borders (s.12.2) ----
    if (lane < 16) {
        yd[lane * BPS - 1] = have_left ? py[lane * ys - 1] : 129;
    } else {
        const int j = lane & 7;
        uint8_t* cd = lane < 24 ? ud : vd;
        const uint8_t* pc = lane < 24 ? pu : pv;
        cd[j * BPS - 1] = have_left ? pc[j * cs - 1] : 129;
    }
    if (lane < 21) {  // luma: top-left, 16 above, 4 above-right
        const int i = lane - 1;
        uint8_t v = 127;
        if (have_top) {
            if (i < 0) v = have_left ? py[-1 - ys] : 129;
            else if (i < 16 || mb_x < mb_w - 1) v = py[i - ys];
            else v = py[15 - ys];
        }
        yd[i - BPS] = v;
    }
    {
        const int i = (lane & 15) - 1;  // chroma: top-left + 8 above, U on lanes 0..8, V on 16..24
        if (i < 8) {
            uint8_t* cd = lane < 16 ? ud : vd;
            const uint8_t* pc = lane < 16 ? pu : pv;
            uint8_t v = 127;
            if (have_top) v = i < 0 ? (have_left ? pc[-1 - cs] : 129) : pc[i - cs];
            cd[i - BPS] = v;
        }
    }
    __syncwarp();
    // ---- luma prediction ----
    if (!mb.is_i4x4) {
        int dc = 0x80;
        if (mb.ymode == DC_PRED) {
            int v = 0;
            if (lane < 16) v = have_top ? yd[lane - BPS] : 0;
            else v = have_left ? yd[(lane - 16) * BPS - 1] : 0;
            const int s = group_sum(v, 32);
            if (have_top && have_left) dc = (s + 16) >> 5;
            else if (have_top || have_left) dc = (s + 8) >> 4;
        }
        pred_span(yd, 16, mb.ymode, lane >> 1, (lane & 1) * 8, 8, dc);
    } else if (lane == 0) {
        for (int r = 1; r < 4; r++)
            for (int i = 16; i < 20; i++) yd[(4 * r - 1) * BPS + i] = yd[i - BPS];
        for (int n = 0; n < 16; n++) {
            uint8_t* d = yd + (n >> 2) * 4 * BPS + (n & 3) * 4;
            pred_4x4(d, BPS, mb.modes[n]);
            if ((mb.nz_blocks >> n) & 1) inverse_dct_add(sm.coeffs + n * 16, d, BPS);
        }
    }
    // ---- chroma prediction: U on lanes 0..15, V on 16..31, 4 pixels each ----
    {
        uint8_t* cd = lane < 16 ? ud : vd;
        const int l = lane & 15;
        int dc = 0x80;
        if (mb.uvmode == DC_PRED) {
            int v = 0;
            if (l < 8) v = have_top ? cd[l - BPS] : 0;
            else v = have_left ? cd[(l - 8) * BPS - 1] : 0;
            const int s = group_sum(v, 16);
            if (have_top && have_left) dc = (s + 8) >> 4;
            else if (have_top || have_left) dc = (s + 4) >> 3;
        }
        pred_span(cd, 8, mb.uvmode, l >> 1, (l & 1) * 4, 4, dc);
    }
    __syncwarp();
    // ---- residuals: one 4x4 block per lane (i4x4 luma was added in order above) ----
    if (lane < 24 && ((mb.nz_blocks >> lane) & 1) && !(mb.is_i4x4 && lane < 16)) {
        uint8_t* d;
        if (lane < 16) d = yd + (lane >> 2) * 4 * BPS + (lane & 3) * 4;
        else {
            const int n = lane & 3;
            d = (lane < 20 ? ud : vd) + (n >> 1) * 4 * BPS + (n & 1) * 4;
        }
        inverse_dct_add(sm.coeffs + lane * 16, d, BPS);
    }
    __syncwarp();
    // ---- store: luma row per lane 0..15, chroma row per lane 16..31 ----
    if (lane < 16) {
        const uint2 a = *reinterpret_cast<const uint2*>(yd + lane * BPS);
        const uint2 b = *reinterpret_cast<const uint2*>(yd + lane * BPS + 8);
        *reinterpret_cast<uint4*>(py + (size_t)lane * ys) = make_uint4(a.x, a.y, b.x, b.y);
    } else {
        const int j = lane & 7;
        const uint8_t* cd = lane < 24 ? ud : vd;
        uint8_t* pc = lane < 24 ? pu : pv;
        *reinterpret_cast<uint2*>(pc + (size_t)j * cs) = *reinterpret_cast<const uint2*>(cd + j * BPS);
    }
    __syncwarp();
}

__global__ void __launch_bounds__(kVp8WarpsPerBlock * 32) vp8_decode_kernel(const Vp8Item* items, int n) {
    __shared__ Vp8WarpSmem s_warp[kVp8WarpsPerBlock];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int idx = blockIdx.x * kVp8WarpsPerBlock + wid;
    if (idx >= n) return;
    const Vp8Item it = items[idx];
    Vp8WarpSmem& sm = s_warp[wid];
    vp8::FrameHdr& h = sm.hdr;
    vp8::Work w;
    vp8::work_carve(it.work, it.mb_w, it.mb_h, w);
    int st = 0;
    vp8::BoolDec br;        // first partition (lane 0)
    vp8::BoolDec parts[8];  // token partitions (lane 0)
    if (lane == 0) {
        st = vp8::parse_frame_header(it.data, it.size, h, br, w.proba);
        if (!st && (h.mb_w != it.mb_w || h.mb_h != it.mb_h)) st = vp8::VP8_BAD;
        if (!st)
            for (int p = 0; p < h.num_parts; p++) vp8::bd_init(parts[p], it.data + h.part_off[p], h.part_len[p]);
        if (st) *it.status = st;  // zeroed by the launcher; an ALPH error must survive
    }
    st = __shfl_sync(0xffffffffu, st, 0);
    __syncwarp();
    if (st) return;
    const int mb_w = h.mb_w, mb_h = h.mb_h;
    for (int i = lane; i < mb_w * 4; i += 32) w.top_modes[i] = vp8::B_DC;
    for (int i = lane; i < mb_w * 9; i += 32) w.top_nz[i] = 0;
    __syncwarp();
    uint32_t* cz = reinterpret_cast<uint32_t*>(sm.coeffs);
    for (int mb_y = 0; mb_y < mb_h; mb_y++) {
        vp8::RowCtx rc;
        vp8::BoolDec tbr;
        if (lane == 0) {
            vp8::row_ctx_reset(rc);
            tbr = parts[mb_y & (h.num_parts - 1)];
        }
        for (int mb_x = 0; mb_x < mb_w; mb_x++) {
            for (int i = lane; i < 25 * 8; i += 32) cz[i] = 0;
            __syncwarp();
            if (lane == 0) {
                const int skip = vp8::parse_mb_modes(h, br, w.top_modes + mb_x * 4, rc, sm.mb);
                vp8::parse_mb_residuals(h, tbr, w.proba, w.top_nz + mb_x * 9, rc, skip, sm.mb, sm.coeffs);
                w.finfo[mb_y * mb_w + mb_x] = sm.mb.finfo;
            }
            __syncwarp();
            reconstruct_mb_warp(sm, w, mb_x, mb_y, lane);
        }
        if (lane == 0) parts[mb_y & (h.num_parts - 1)] = tbr;
    }
    __syncwarp();
    if (h.filter_type == 0) return;
    for (int mb_y = 0; mb_y < mb_h; mb_y++)
        for (int mb_x = 0; mb_x < mb_w; mb_x++) filter_macroblock_warp(h, w, mb_x, mb_y, lane);
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

// VP8L (lossless) frames and ALPH planes: an LZ77 + prefix-coded stream is one serial chain, so
// one thread walks it (vp8l_core.h) inside a bump arena in HBM; the colour-order conversion to the
// mat is a separate, parallel kernel.
struct Vp8lItem {
    const uint8_t* data;   // "VP8L" chunk payload, or "ALPH" chunk payload
    uint32_t size;
    int width, height;
    uint8_t* arena;
    size_t arena_cap;
    int is_alph;
    uint8_t* alpha_out;    // is_alph: width*height plane
    uint32_t** px_out;     // !is_alph: where to leave the pointer to the final ARGB pixels
    int* status;
};

__global__ void vp8l_decode_kernel(Vp8lItem it) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    vp8l::Arena a{it.arena, it.arena_cap, 0};
    int rc;
    if (it.is_alph) {
        rc = vp8l::decode_alph(it.data, it.size, it.width, it.height, a, it.alpha_out);
    } else {
        uint32_t* px = nullptr;
        rc = vp8l::decode_vp8l(it.data, it.size, it.width, it.height, a, &px);
        *it.px_out = px;
    }
    if (rc) *it.status = rc;
}

__global__ void argb_output_kernel(uint32_t* const* px_ptr, int width, int height, uint8_t* dst, size_t dst_step,
                                   int channels) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const uint32_t* px = *px_ptr;
    if (x >= width || !px) return;
    const uint32_t v = px[(size_t)y * width + x];
    uint8_t* d = dst + (size_t)y * dst_step + (size_t)x * channels;
    d[0] = (uint8_t)v;
    d[1] = (uint8_t)(v >> 8);
    d[2] = (uint8_t)(v >> 16);
    if (channels == 4) d[3] = (uint8_t)(v >> 24);
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

// ------------------------------------------------------------------ batch helpers (xbatch.cu)

// What the batch path needs to know about a file: is it ONE lossy key frame without alpha, profile or
// animation (then its VP8 payload can join a grid launch), and where that payload lies.
bool webp_still_info(const uint8_t* data, size_t len, WebpStillInfo* out) {
    WebpContainer c;
    if (!webp_parse(data, len, &c)) return false;
    if (c.frames.size() != 1) return true;
    const WebpFrame& f = c.frames[0];
    out->width = c.canvas_w;
    out->height = c.canvas_h;
    out->vp8_off = f.img_off;
    out->vp8_len = f.img_len;
    out->simple_lossy = !f.lossless && !f.has_alph && !c.has_icc && !(c.flags & kFlagAnim) && !(c.flags & kFlagAlpha) &&
                        f.width == c.canvas_w && f.height == c.canvas_h && f.img_off + f.img_len <= len;
    return true;
}

__global__ void vp8_output_batch_kernel(const uint8_t* work, size_t work_stride, int mb_w, int mb_h, int width, int height,
                                        uint8_t* frames, size_t frame_stride) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= width) return;
    const uint8_t* base = work + (size_t)blockIdx.z * work_stride;
    const size_t ypl = (size_t)mb_w * 16 * mb_h * 16;
    const uint8_t* yp = base;
    const uint8_t* up = base + ypl;
    const uint8_t* vp = up + ypl / 4;
    const int u = vp8::upsample_at(up, mb_w * 8, width, height, x, y);
    const int v = vp8::upsample_at(vp, mb_w * 8, width, height, x, y);
    uint8_t bgr[3];
    vp8::yuv_to_bgr(yp[(size_t)y * (mb_w * 16) + x], u, v, bgr);
    uint8_t* d = frames + (size_t)blockIdx.z * frame_stride + ((size_t)y * width + x) * 3;
    d[0] = bgr[0];
    d[1] = bgr[1];
    d[2] = bgr[2];
}

// n VP8 key frames (any sizes, equal sizes adjacent): one warp per frame (vp8_decode_kernel) in ONE launch,
// then every pixel of every frame through the fancy upsampler + colour conversion, one launch per run of
// equal geometry.  Packed BGR frames at d_frames + frame_off[i].
int webp_vp8_decode_batch(const uint8_t* d_in, const uint64_t* in_off, const uint32_t* in_len, int n, const int* width,
                          const int* height, uint8_t* d_frames, const uint64_t* frame_off, int* h_status, cudaStream_t st) {
    if (n <= 0) return LP_OK;
    std::vector<size_t> work_off((size_t)n + 1, 0);
    for (int i = 0; i < n; i++) work_off[i + 1] = work_off[i] + vp8::work_bytes((width[i] + 15) >> 4, (height[i] + 15) >> 4);
    uint8_t* scratch = nullptr;
    const size_t items_b = round_up((size_t)n * sizeof(Vp8Item), (size_t)256), stat_b = round_up((size_t)n * 4, (size_t)256);
    if (cudaMallocAsync(&scratch, items_b + stat_b + work_off[n], st) != cudaSuccess) {
        cudaGetLastError();
        return LP_ERR_CUDA;
    }
    Vp8Item* d_items = reinterpret_cast<Vp8Item*>(scratch);
    int* d_status = reinterpret_cast<int*>(scratch + items_b);
    uint8_t* d_work = scratch + items_b + stat_b;
    std::vector<Vp8Item> items((size_t)n);
    for (int i = 0; i < n; i++)
        items[i] = Vp8Item{d_in + in_off[i], in_len[i], d_work + work_off[i], d_status + i, (width[i] + 15) >> 4, (height[i] + 15) >> 4};
    int rc = LP_OK;
    cudaMemsetAsync(d_status, 0, (size_t)n * 4, st);
    if (cudaMemcpyAsync(d_items, items.data(), (size_t)n * sizeof(Vp8Item), cudaMemcpyHostToDevice, st) != cudaSuccess) rc = LP_ERR_CUDA;
    if (!rc) {
        vp8_decode_kernel<<<ceil_div(n, kVp8WarpsPerBlock), kVp8WarpsPerBlock * 32, 0, st>>>(d_items, n);
        g_launches++;
        for (int i0 = 0; i0 < n;) {
            // frames of one geometry lie back to back, round_up(w * h * 3, 256) apart (the caller's layout)
            const size_t fstride = round_up((size_t)width[i0] * height[i0] * 3, (size_t)256);
            int i1 = i0 + 1;
            while (i1 < n && width[i1] == width[i0] && height[i1] == height[i0] &&
                   frame_off[i1] == frame_off[i0] + (uint64_t)(i1 - i0) * fstride)
                i1++;
            const int mb_w = (width[i0] + 15) >> 4, mb_h = (height[i0] + 15) >> 4;
            dim3 grid(ceil_div(width[i0], 128), height[i0], i1 - i0);
            vp8_output_batch_kernel<<<grid, 128, 0, st>>>(d_work + work_off[i0], vp8::work_bytes(mb_w, mb_h), mb_w, mb_h, width[i0],
                                                          height[i0], d_frames + frame_off[i0], fstride);
            g_launches++;
            i0 = i1;
        }
        if (cudaGetLastError() != cudaSuccess) rc = LP_ERR_CUDA;
    }
    if (!rc && (cudaMemcpyAsync(h_status, d_status, (size_t)n * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                cudaStreamSynchronize(st) != cudaSuccess))  // (also keeps `items` alive until the copy has been consumed)
        rc = LP_ERR_CUDA;
    cudaFreeAsync(scratch, st);
    return rc;
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
    uint8_t* d_arena = nullptr;   // VP8L / ALPH bump arena
    uint8_t* d_alpha = nullptr;   // decoded ALPH plane
    uint8_t* d_alph_in = nullptr; // ALPH chunk bytes
    uint32_t** d_px = nullptr;    // where the VP8L kernel leaves its result pointer
    size_t in_cap = 0, work_cap = 0, arena_cap = 0, alpha_cap = 0, alph_in_cap = 0;
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
    if (d->d_arena) cudaFreeAsync(d->d_arena, st);
    if (d->d_alpha) cudaFreeAsync(d->d_alpha, st);
    if (d->d_alph_in) cudaFreeAsync(d->d_alph_in, st);
    if (d->d_px) cudaFreeAsync(d->d_px, st);
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
    cudaStream_t st = thread_stream();
    const int channels = type == CV_8UC4 ? 4 : 3;
    auto grow = [&](uint8_t** p, size_t* cap, size_t need) -> bool {
        if (need <= *cap) return true;
        if (*p) cudaFreeAsync(*p, st);
        *cap = need;
        return cudaMallocAsync(p, need, st) == cudaSuccess;
    };
    if (!d->d_item) {
        if (cudaMallocAsync(&d->d_item, sizeof(Vp8Item), st) != cudaSuccess) return false;
        if (cudaMallocAsync(&d->d_status, sizeof(int), st) != cudaSuccess) return false;
        if (cudaMallocAsync(&d->d_px, sizeof(uint32_t*), st) != cudaSuccess) return false;
    }
    cudaMemsetAsync(d->d_status, 0, sizeof(int), st);
    cudaMemsetAsync(d->d_px, 0, sizeof(uint32_t*), st);
    if (!grow(&d->d_in, &d->in_cap, f.img_len + 4096)) return false;
    cudaMemcpyAsync(d->d_in, d->bytes + f.img_off, f.img_len, cudaMemcpyHostToDevice, st);
    const size_t npix = (size_t)f.width * f.height;
    const size_t arena_need = npix * 12 + (16u << 20);
    const bool need_alph = f.has_alph && !f.lossless && channels == 4;
    if (f.lossless || need_alph) {
        if (!grow(&d->d_arena, &d->arena_cap, arena_need)) return false;
    }
    if (f.lossless) {
        Vp8lItem li{d->d_in, (uint32_t)f.img_len, f.width, f.height, d->d_arena, d->arena_cap, 0, nullptr, d->d_px, d->d_status};
        vp8l_decode_kernel<<<1, 32, 0, st>>>(li);
        g_launches++;
        dim3 grid(ceil_div(f.width, 128), f.height);
        argb_output_kernel<<<grid, 128, 0, st>>>(d->d_px, f.width, f.height, frame_dev, frame_step, channels);
        g_launches++;
    } else {
        const int mb_w = (f.width + 15) >> 4, mb_h = (f.height + 15) >> 4;
        if (!grow(&d->d_work, &d->work_cap, vp8::work_bytes(mb_w, mb_h))) return false;
        if (need_alph) {
            if (!grow(&d->d_alph_in, &d->alph_in_cap, f.alph_len + 4096)) return false;
            if (!grow(&d->d_alpha, &d->alpha_cap, npix + 256)) return false;
            cudaMemcpyAsync(d->d_alph_in, d->bytes + f.alph_off, f.alph_len, cudaMemcpyHostToDevice, st);
            Vp8lItem ai{d->d_alph_in, (uint32_t)f.alph_len, f.width, f.height, d->d_arena, d->arena_cap, 1, d->d_alpha, nullptr, d->d_status};
            vp8l_decode_kernel<<<1, 32, 0, st>>>(ai);
            g_launches++;
        }
        Vp8Item item{d->d_in, (uint32_t)f.img_len, d->d_work, d->d_status, mb_w, mb_h};
        cudaMemcpyAsync(d->d_item, &item, sizeof(item), cudaMemcpyHostToDevice, st);
        vp8_decode_kernel<<<1, kVp8WarpsPerBlock * 32, 0, st>>>(d->d_item, 1);
        g_launches++;
        vp8::Work w;
        vp8::work_carve(d->d_work, mb_w, mb_h, w);
        Vp8Output o{w.y, w.u, w.v, need_alph ? d->d_alpha : nullptr, mb_w * 16, mb_w * 8, f.width, f.height, frame_dev, frame_step, channels};
        dim3 grid(ceil_div(f.width, 128), f.height);
        vp8_output_kernel<<<grid, 128, 0, st>>>(o);
        g_launches++;
    }
    int status = 0;
    cudaMemcpyAsync(&status, d->d_status, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) return false;
    if (status != 0) return false;
    mat_mark_device_written(mat);
    return true;
}

}  // extern "C"
