// vp8_cpu.cpp -- TEST INFRASTRUCTURE.  Compiles lilliput_b200/csrc/vp8_core.h for the host so
// the VP8 decoding logic the device kernels run can be checked bit-for-bit against the
// reference's libwebp (through oracle/_ref) on a machine without a GPU.  Not part of the product.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../lilliput_b200/csrc/vp8_core.h"

extern "C" int vp8_cpu_info(const uint8_t* d, size_t n, int* w, int* h) {
    if (n < 10 || d[3] != 0x9d || d[4] != 0x01 || d[5] != 0x2a) return 1;
    *w = ((d[7] << 8) | d[6]) & 0x3fff;
    *h = ((d[9] << 8) | d[8]) & 0x3fff;
    return 0;
}

// Decodes a VP8 key-frame payload to interleaved BGR.  `stage` 0 = final, 1 = skip loop filter.
extern "C" int vp8_cpu_decode_bgr(const uint8_t* d, size_t n, uint8_t* out, int stride, int stage,
                                  uint8_t* yuv_out) {
    vp8::FrameHdr h;
    vp8::BoolDec br;
    uint8_t proba[1056];
    if (vp8::parse_frame_header(d, n, h, br, proba)) return 1;
    std::vector<uint8_t> mem(vp8::work_bytes(h.mb_w, h.mb_h));
    vp8::Work w;
    vp8::work_carve(mem.data(), h.mb_w, h.mb_h, w);
    memcpy(w.proba, proba, 1056);
    if (vp8::decode_macroblocks(d, h, br, w)) return 2;
    if (stage == 0 && h.filter_type > 0)
        for (int y = 0; y < h.mb_h; y++)
            for (int x = 0; x < h.mb_w; x++) vp8::filter_macroblock(h, w, x, y);
    const int ys = h.mb_w * 16, cs = h.mb_w * 8;
    if (yuv_out) memcpy(yuv_out, w.y, (size_t)ys * h.mb_h * 16 * 3 / 2);
    for (int y = 0; y < h.height; y++)
        for (int x = 0; x < h.width; x++) {
            const int u = vp8::upsample_at(w.u, cs, h.width, h.height, x, y);
            const int v = vp8::upsample_at(w.v, cs, h.width, h.height, x, y);
            vp8::yuv_to_bgr(w.y[(size_t)y * ys + x], u, v, out + (size_t)y * stride + 3 * x);
        }
    return 0;
}

// ---- VP8L (lossless) and ALPH, same idea ---------------------------------------------------
#include "../../lilliput_b200/csrc/vp8l_core.h"

// Decodes a "VP8L" chunk payload to BGRA (channels = 4) or BGR (3).
extern "C" int vp8l_cpu_decode(const uint8_t* d, size_t n, int w, int h, uint8_t* out, int channels) {
    std::vector<uint8_t> mem((size_t)w * h * 16 + (32u << 20));
    vp8l::Arena a{mem.data(), mem.size(), 0};
    uint32_t* px = nullptr;
    const int rc = vp8l::decode_vp8l(d, n, w, h, a, &px);
    if (rc) return rc;
    for (size_t i = 0; i < (size_t)w * h; i++) {
        out[i * channels + 0] = (uint8_t)px[i];
        out[i * channels + 1] = (uint8_t)(px[i] >> 8);
        out[i * channels + 2] = (uint8_t)(px[i] >> 16);
        if (channels == 4) out[i * 4 + 3] = (uint8_t)(px[i] >> 24);
    }
    return 0;
}

// Decodes an "ALPH" chunk payload to a w*h alpha plane.
extern "C" int alph_cpu_decode(const uint8_t* d, size_t n, int w, int h, uint8_t* alpha) {
    std::vector<uint8_t> mem((size_t)w * h * 16 + (32u << 20));
    vp8l::Arena a{mem.data(), mem.size(), 0};
    return vp8l::decode_alph(d, n, w, h, a, alpha);
}
