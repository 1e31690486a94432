// vp8l_enc_core.h -- a WebP lossless (VP8L) ENCODER small enough to run per pixel on the device.
// Stream layout (WebP Lossless Bitstream Specification):
//   header | subtract-green transform | predictor transform (one mode, 12 = clamp(L + T - TL), for
//   the whole picture, carried by a sub-image whose pixels cost zero bits) | no colour cache | no
//   meta prefix image | five prefix codes built from the residual histograms | every pixel as four
//   literals (green, red, blue, alpha).
// No LZ77, no colour cache: each pixel's bits depend only on its own residual, so the bit length
// of every pixel is known independently, a prefix sum places it, and all pixels are packed in
// parallel.  Lossless by construction: any conforming decoder returns the input pixels
// (tests decode the output with the reference's libwebp and with vp8l_core.h and compare).
//
// Stands where WebPEncodeLosslessBGR / BGRA do for the reference (ref webp.cpp:711-720); the
// reference's files are smaller (libwebp searches LZ77 + colour cache + per-block transforms) --
// file bytes are NOT comparable, decoded pixels are.
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifndef LP_L_HD
#define LP_L_HD static inline
#endif

namespace vp8lenc {

enum { kPredMode = 12, kPredBits = 9 };  // 512x512 blocks: the mode sub-image is a handful of pixels

// channels 3 / 4: a BGR(A) frame.  channels 1: a plane that travels in the GREEN channel (how an ALPH
// chunk carries alpha), the other channels constant.
LP_L_HD uint32_t load_argb(const uint8_t* frame, size_t step, int channels, int x, int y) {
    const uint8_t* p = frame + (size_t)y * step + (size_t)x * channels;
    if (channels == 1) return 0xff000000u | ((uint32_t)p[0] << 8);
    const uint32_t a = channels == 4 ? p[3] : 255u;
    return (a << 24) | ((uint32_t)p[2] << 16) | ((uint32_t)p[1] << 8) | p[0];
}
LP_L_HD uint32_t subtract_green(uint32_t argb) {
    const uint32_t g = (argb >> 8) & 255;
    const uint32_t r = (((argb >> 16) & 255) - g) & 255, b = ((argb & 255) - g) & 255;
    return (argb & 0xff00ff00u) | (r << 16) | b;
}
LP_L_HD uint32_t sub_px(uint32_t a, uint32_t b) {  // per-byte a - b mod 256
    const uint32_t ag = 0x00ff00ffu + (a & 0xff00ff00u) - (b & 0xff00ff00u);
    const uint32_t rb = 0xff00ff00u + (a & 0x00ff00ffu) - (b & 0x00ff00ffu);
    return (ag & 0xff00ff00u) | (rb & 0x00ff00ffu);
}
LP_L_HD uint32_t clamp_add_sub_full(uint32_t L, uint32_t T, uint32_t TL) {
    uint32_t r = 0;
    for (int s = 0; s < 32; s += 8) {
        int v = (int)((L >> s) & 255) + (int)((T >> s) & 255) - (int)((TL >> s) & 255);
        v = v < 0 ? 0 : v > 255 ? 255 : v;
        r |= (uint32_t)v << s;
    }
    return r;
}
// Residual of pixel (x, y): green-subtracted value minus its prediction from the green-subtracted
// ORIGINAL neighbours (lossless: the decoder's reconstruction equals the original).
LP_L_HD uint32_t source_px(const uint8_t* frame, size_t step, int channels, int x, int y) {
    const uint32_t v = load_argb(frame, step, channels, x, y);
    return channels == 1 ? v : subtract_green(v);  // a lone plane has nothing to decorrelate
}
LP_L_HD uint32_t residual_at(const uint8_t* frame, size_t step, int channels, int x, int y) {
    const uint32_t cur = source_px(frame, step, channels, x, y);
    uint32_t pred;
    if (y == 0) pred = x == 0 ? 0xff000000u : source_px(frame, step, channels, x - 1, 0);
    else if (x == 0) pred = source_px(frame, step, channels, 0, y - 1);
    else
        pred = clamp_add_sub_full(source_px(frame, step, channels, x - 1, y), source_px(frame, step, channels, x, y - 1),
                                  source_px(frame, step, channels, x - 1, y - 1));
    return sub_px(cur, pred);
}

// Prefix codes of the four literal alphabets, ready for LSB-first packing (code bits reversed).
struct CodeTable {
    uint16_t code[4][256];  // [0] green, [1] red, [2] blue, [3] alpha
    uint8_t len[4][256];
};
// The up-to-60 bits of one pixel, in stream order.
LP_L_HD void pixel_bits(uint32_t resid, const CodeTable& t, uint64_t* bits, int* nbits) {
    const int g = (resid >> 8) & 255, r = (resid >> 16) & 255, b = resid & 255, a = resid >> 24;
    uint64_t v = t.code[0][g];
    int n = t.len[0][g];
    v |= (uint64_t)t.code[1][r] << n;
    n += t.len[1][r];
    v |= (uint64_t)t.code[2][b] << n;
    n += t.len[2][b];
    v |= (uint64_t)t.code[3][a] << n;
    n += t.len[3][a];
    *bits = v;
    *nbits = n;
}

}  // namespace vp8lenc

// ---------------------------------------------------------------- host-only part
#include <algorithm>
#include <vector>

namespace vp8lenc {

struct BitWriter {  // least-significant bit first (spec s.2)
    std::vector<uint8_t> bytes;
    uint64_t acc = 0;
    int n = 0;
    size_t nbits = 0;
    void put(uint32_t v, int k) {
        acc |= (uint64_t)(v & (k >= 32 ? 0xffffffffu : ((1u << k) - 1))) << n;
        n += k;
        nbits += (size_t)k;
        while (n >= 8) {
            bytes.push_back((uint8_t)acc);
            acc >>= 8;
            n -= 8;
        }
    }
    void flush() {
        if (n > 0) bytes.push_back((uint8_t)acc);
        acc = 0;
        n = 0;
    }
};

// Length-limited (15) Huffman code lengths from a histogram: plain Huffman, and if the tree is too
// deep the small counts are lifted and the tree rebuilt (the result stays a complete prefix code).
static inline void build_lengths(const uint32_t* hist, int n, uint8_t* lens) {
    std::vector<uint32_t> cnt(hist, hist + n);
    for (;;) {
        struct Node { uint64_t w; int l, r; };
        std::vector<Node> nodes;
        std::vector<int> live;
        for (int i = 0; i < n; i++) {
            lens[i] = 0;
            if (cnt[i]) {
                nodes.push_back({cnt[i], -1 - i, 0});
                live.push_back((int)nodes.size() - 1);
            }
        }
        if (live.size() < 2) {
            if (live.size() == 1) lens[-1 - nodes[live[0]].l] = 1;
            return;
        }
        auto cmp = [&](int a, int b) { return nodes[a].w > nodes[b].w || (nodes[a].w == nodes[b].w && a > b); };
        std::make_heap(live.begin(), live.end(), cmp);
        while (live.size() > 1) {
            std::pop_heap(live.begin(), live.end(), cmp);
            const int a = live.back();
            live.pop_back();
            std::pop_heap(live.begin(), live.end(), cmp);
            const int b = live.back();
            live.pop_back();
            nodes.push_back({nodes[a].w + nodes[b].w, a, b});
            live.push_back((int)nodes.size() - 1);
            std::push_heap(live.begin(), live.end(), cmp);
        }
        // depths
        int maxd = 0;
        std::vector<std::pair<int, int>> stack{{live[0], 0}};
        while (!stack.empty()) {
            auto [id, d] = stack.back();
            stack.pop_back();
            if (nodes[id].l < 0) {
                lens[-1 - nodes[id].l] = (uint8_t)std::min(d, 255);
                maxd = std::max(maxd, d);
            } else {
                stack.push_back({nodes[id].l, d + 1});
                stack.push_back({nodes[id].r, d + 1});
            }
        }
        if (maxd <= 15) return;
        for (auto& c : cnt)
            if (c) c = (c >> 1) + 1;  // flatten the distribution and try again
        uint32_t mx = *std::max_element(cnt.begin(), cnt.end());
        if (mx <= 2)
            for (auto& c : cnt)
                if (c) c = 1;
    }
}

// Canonical codes (MSB-first per the spec) bit-reversed for an LSB-first writer.
static inline void canonical_reversed(const uint8_t* lens, int n, uint16_t* codes) {
    int count[16] = {0}, next[16] = {0};
    for (int i = 0; i < n; i++) count[lens[i]]++;
    count[0] = 0;
    int code = 0;
    for (int l = 1; l < 16; l++) {
        code = (code + count[l - 1]) << 1;
        next[l] = code;
    }
    for (int i = 0; i < n; i++) {
        codes[i] = 0;
        const int l = lens[i];
        if (!l) continue;
        const int c = next[l]++;
        int rev = 0;
        for (int b = 0; b < l; b++) rev |= ((c >> b) & 1) << (l - 1 - b);
        codes[i] = (uint16_t)rev;
    }
}

// One prefix code definition (spec s.6.2): the simple form for <= 1 used symbol, else the normal form
// with a flat 4-bit code-length code (every length 0..15 costs 4 bits, no run-length symbols).
static inline void write_code(BitWriter& w, const uint8_t* lens, int alphabet) {
    int used = 0, last = 0;
    for (int i = 0; i < alphabet; i++)
        if (lens[i]) {
            used++;
            last = i;
        }
    if (used <= 1) {
        w.put(1, 1);                    // simple code
        w.put(0, 1);                    // one symbol
        w.put(last > 1 ? 1 : 0, 1);     // symbol width: 8 bits or 1 bit
        w.put((uint32_t)last, last > 1 ? 8 : 1);
        return;
    }
    static const uint8_t order[19] = {17, 18, 0, 1, 2, 3, 4, 5, 16, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15};
    w.put(0, 1);        // normal code
    w.put(19 - 4, 4);   // all 19 code-length-code lengths follow
    for (int i = 0; i < 19; i++) w.put(order[i] < 16 ? 4 : 0, 3);
    w.put(0, 1);        // lengths for the whole alphabet
    for (int i = 0; i < alphabet; i++) {
        const int v = lens[i];  // canonical 4-bit code of value v is v itself, sent MSB first
        w.put(((v & 1) << 3) | ((v & 2) << 1) | ((v & 4) >> 1) | ((v & 8) >> 3), 4);
    }
}

// Everything that precedes the pixel data.  hist = 4 x 256 residual histograms (green, red, blue,
// alpha).  with_header: the 5-byte VP8L header (absent for an ALPH payload, which is also the case
// that skips subtract-green: see source_px).  Fills `table`.
static inline void write_stream_head(BitWriter& w, int width, int height, bool has_alpha, bool with_header,
                                     bool use_subtract_green, const uint32_t* hist, CodeTable* table) {
    if (with_header) {
        w.put(0x2f, 8);
        w.put((uint32_t)(width - 1), 14);
        w.put((uint32_t)(height - 1), 14);
        w.put(has_alpha ? 1 : 0, 1);
        w.put(0, 3);
    }
    if (use_subtract_green) {
        w.put(1, 1);  // transform: subtract green
        w.put(2, 2);
    }
    w.put(1, 1);  // transform: predictor, block size 1 << kPredBits
    w.put(0, 2);
    w.put(kPredBits - 2, 3);
    {   // the mode sub-image: every pixel 0xff00<mode>00, five single-symbol codes, zero bits per pixel
        w.put(0, 1);  // no colour cache
        uint8_t one[280];
        const int sym[5] = {kPredMode, 0, 0, 255, 0};
        const int alpha_sz[5] = {280, 256, 256, 256, 40};
        for (int k = 0; k < 5; k++) {
            for (int i = 0; i < alpha_sz[k]; i++) one[i] = 0;
            one[sym[k]] = 1;
            write_code(w, one, alpha_sz[k]);
        }
    }
    w.put(0, 1);  // no more transforms
    w.put(0, 1);  // no colour cache
    w.put(0, 1);  // no meta prefix image
    uint8_t lens[280];
    for (int k = 0; k < 4; k++) {
        const int alphabet = k == 0 ? 280 : 256;
        for (int i = 0; i < alphabet; i++) lens[i] = 0;
        build_lengths(hist + k * 256, 256, lens);
        int used = 0;
        for (int i = 0; i < 256; i++) used += lens[i] != 0;
        write_code(w, lens, alphabet);
        if (used <= 1)
            for (int i = 0; i < 256; i++) lens[i] = 0;  // a single-symbol code is read with zero bits
        canonical_reversed(lens, 256, table->code[k]);
        for (int i = 0; i < 256; i++) table->len[k][i] = lens[i];
    }
    {   // distance code: never used (no backward references), still has to be defined
        uint8_t d[40] = {0};
        d[0] = 1;
        write_code(w, d, 40);
    }
}

}  // namespace vp8lenc
