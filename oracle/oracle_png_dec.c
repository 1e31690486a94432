/*
 * oracle_png_dec.c -- CPU restatement of PNG decoding as the reference performs it.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Reference call site: opencv_decoder_read_data (ref opencv.cpp:166-171) ->
 * cv::ImageDecoder::readData -> OpenCV 4.11 grfmt_png.cpp -> libpng 1.6.47 + zlib-ng 2.3.3
 * (ref deps/build-deps-linux.sh:191,209; sources not in tree).  Restated from the PNG
 * specification (W3C PNG, RFC 1950 zlib, RFC 1951 DEFLATE) plus the transform set OpenCV asks
 * libpng for when the destination Mat is 8-bit with the channel count lilliput's Framebuffer
 * gives it (ref opencv.go:250-267): strip_16 (high byte), palette_to_rgb, tRNS_to_alpha,
 * expand_gray_1_2_4_to_8, gray_to_rgb for gray+alpha, BGR order.  Output channel count:
 * gray -> 1; RGB / palette -> 3, or 4 when a tRNS chunk is present; gray+alpha / RGBA -> 4.
 * Lossless, so any conformant inflate + defilter gives identical pixels.
 * Interlaced (Adam7) images are not handled here (returns -2), matching the device path.
 */
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

typedef struct { const uint8_t* p; size_t n, pos; uint64_t acc; int cnt; } InBits;

static inline void ib_fill(InBits* b) {
    while (b->cnt <= 56) {
        uint64_t byte = b->pos < b->n ? b->p[b->pos] : 0;
        b->pos++;
        b->acc |= byte << b->cnt;
        b->cnt += 8;
    }
}
static inline unsigned ib_get(InBits* b, int n) {
    if (n == 0) return 0;
    if (b->cnt < n) ib_fill(b);
    unsigned v = (unsigned)(b->acc & ((1ull << n) - 1));
    b->acc >>= n;
    b->cnt -= n;
    return v;
}

typedef struct { uint16_t count[16], sym[320]; } Huff;

static int huff_build(Huff* h, const uint8_t* len, int n) {
    uint16_t offs[16];
    memset(h->count, 0, sizeof(h->count));
    for (int i = 0; i < n; i++) h->count[len[i]]++;
    h->count[0] = 0;
    int left = 1;
    for (int l = 1; l < 16; l++) {
        left <<= 1;
        left -= h->count[l];
        if (left < 0) return -1;
    }
    offs[1] = 0;
    for (int l = 1; l < 15; l++) offs[l + 1] = offs[l] + h->count[l];
    for (int i = 0; i < n; i++)
        if (len[i]) h->sym[offs[len[i]]++] = (uint16_t)i;
    return 0;
}
static int huff_decode(InBits* b, const Huff* h) {
    int code = 0, first = 0, index = 0;
    for (int l = 1; l < 16; l++) {
        code |= (int)ib_get(b, 1);
        int c = h->count[l];
        if (code - c < first) return h->sym[index + (code - first)];
        index += c;
        first += c;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}

static const uint16_t LBASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t LEXT[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t DBASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t DEXT[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

/* RFC 1950 + 1951.  Returns bytes produced, or -1. */
static long zlib_inflate(const uint8_t* in, size_t n, uint8_t* out, size_t cap) {
    if (n < 2 || (in[0] & 15) != 8 || ((in[0] << 8) | in[1]) % 31) return -1;
    if (in[1] & 0x20) return -1; /* preset dictionary */
    InBits b = {in + 2, n - 2, 0, 0, 0};
    size_t o = 0;
    int last;
    do {
        last = (int)ib_get(&b, 1);
        int type = (int)ib_get(&b, 2);
        if (type == 0) {
            ib_get(&b, b.cnt & 7); /* to the byte boundary */
            unsigned len = ib_get(&b, 16), nlen = ib_get(&b, 16);
            if ((len ^ 0xFFFF) != nlen) return -1;
            for (unsigned i = 0; i < len; i++) {
                if (o >= cap) return -1;
                out[o++] = (uint8_t)ib_get(&b, 8);
            }
        } else if (type == 1 || type == 2) {
            Huff hl, hd;
            uint8_t lens[320];
            if (type == 1) {
                int i = 0;
                for (; i < 144; i++) lens[i] = 8;
                for (; i < 256; i++) lens[i] = 9;
                for (; i < 280; i++) lens[i] = 7;
                for (; i < 288; i++) lens[i] = 8;
                huff_build(&hl, lens, 288);
                for (i = 0; i < 30; i++) lens[i] = 5;
                huff_build(&hd, lens, 30);
            } else {
                static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
                int nl = (int)ib_get(&b, 5) + 257, nd = (int)ib_get(&b, 5) + 1, nc = (int)ib_get(&b, 4) + 4;
                if (nl > 286 || nd > 30) return -1;
                uint8_t cl[19];
                memset(cl, 0, sizeof(cl));
                for (int i = 0; i < nc; i++) cl[order[i]] = (uint8_t)ib_get(&b, 3);
                Huff hc;
                if (huff_build(&hc, cl, 19)) return -1;
                int i = 0;
                while (i < nl + nd) {
                    int s = huff_decode(&b, &hc);
                    if (s < 0) return -1;
                    if (s < 16) lens[i++] = (uint8_t)s;
                    else {
                        int rep, v = 0;
                        if (s == 16) { if (!i) return -1; v = lens[i - 1]; rep = 3 + (int)ib_get(&b, 2); }
                        else if (s == 17) rep = 3 + (int)ib_get(&b, 3);
                        else rep = 11 + (int)ib_get(&b, 7);
                        if (i + rep > nl + nd) return -1;
                        while (rep--) lens[i++] = (uint8_t)v;
                    }
                }
                if (huff_build(&hl, lens, nl)) return -1;
                huff_build(&hd, lens + nl, nd); /* incomplete distance codes are legal */
            }
            for (;;) {
                int s = huff_decode(&b, &hl);
                if (s < 0) return -1;
                if (s < 256) {
                    if (o >= cap) return -1;
                    out[o++] = (uint8_t)s;
                } else if (s == 256) break;
                else {
                    s -= 257;
                    if (s >= 29) return -1;
                    unsigned len = LBASE[s] + ib_get(&b, LEXT[s]);
                    int ds = huff_decode(&b, &hd);
                    if (ds < 0 || ds >= 30) return -1;
                    unsigned dist = DBASE[ds] + ib_get(&b, DEXT[ds]);
                    if (dist > o || o + len > cap) return -1;
                    for (unsigned i = 0; i < len; i++, o++) out[o] = out[o - dist];
                }
            }
        } else return -1;
    } while (!last);
    return (long)o;
}

static inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
static inline int paeth(int a, int b, int c) {
    int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

int oracle_png_decode(const uint8_t* in, size_t len, uint8_t* out, size_t out_cap, int* width,
                      int* height, int* channels, int* depth) {
    static const uint8_t sig[8] = {0x89, 0x50, 0x4E, 0x47, 0x0D, 0x0A, 0x1A, 0x0A};
    if (len < 8 + 25 || memcmp(in, sig, 8)) return -3;
    int W = 0, H = 0, bd = 0, ct = 0, interlace = 0, have_trns = 0, ntrns = 0, npal = 0;
    uint8_t pal[256][3], trns[256];
    uint16_t trns_rgb[3] = {0, 0, 0};
    uint8_t* z = malloc(len);
    size_t zn = 0, pos = 8;
    int rc = 0;
    while (pos + 12 <= len) {
        uint32_t n = be32(in + pos);
        const uint8_t* t = in + pos + 4;
        const uint8_t* d = in + pos + 8;
        if (pos + 12 + (size_t)n > len) { rc = -1; break; }
        if (!memcmp(t, "IHDR", 4) && n >= 13) {
            W = (int)be32(d); H = (int)be32(d + 4); bd = d[8]; ct = d[9]; interlace = d[12];
        } else if (!memcmp(t, "PLTE", 4)) {
            npal = (int)(n / 3);
            if (npal > 256) npal = 256;
            memcpy(pal, d, (size_t)npal * 3);
        } else if (!memcmp(t, "tRNS", 4)) {
            have_trns = 1;
            if (ct == 3) { ntrns = n > 256 ? 256 : (int)n; memcpy(trns, d, ntrns); }
            else if (ct == 2 && n >= 6) for (int i = 0; i < 3; i++) trns_rgb[i] = (uint16_t)((d[2 * i] << 8) | d[2 * i + 1]);
        } else if (!memcmp(t, "IDAT", 4)) {
            memcpy(z + zn, d, n);
            zn += n;
        } else if (!memcmp(t, "IEND", 4)) break;
        pos += 12 + (size_t)n;
    }
    if (rc || W < 1 || H < 1) { free(z); return rc ? rc : -3; }
    int src_ch = ct == 0 ? 1 : ct == 2 ? 3 : ct == 3 ? 1 : ct == 4 ? 2 : ct == 6 ? 4 : 0;
    if (!src_ch || (bd != 1 && bd != 2 && bd != 4 && bd != 8 && bd != 16)) { free(z); return -3; }
    int och = (ct == 0) ? 1 : (ct == 4 || ct == 6) ? 4 : (have_trns ? 4 : 3);
    if (width) *width = W;
    if (height) *height = H;
    if (channels) *channels = och;
    if (depth) *depth = bd;
    if (!out) { free(z); return 0; }
    if ((size_t)W * H * och > out_cap) { free(z); return -4; }
    /* Adam7 (PNG spec s.8.2): seven reduced images, each filtered on its own; pass geometry below.
     * A non-interlaced image is the single "pass" {0,0,1,1}. */
    static const int PX0[7] = {0, 4, 0, 2, 0, 1, 0}, PY0[7] = {0, 0, 4, 0, 2, 0, 1};
    static const int PDX[7] = {8, 8, 4, 4, 2, 2, 1}, PDY[7] = {8, 8, 8, 4, 4, 2, 2};
    const int npass = interlace ? 7 : 1;
    size_t bpp_bits = (size_t)src_ch * bd, bpp = bpp_bits >= 8 ? bpp_bits / 8 : 1;
    size_t total = 0;
    for (int ps = 0; ps < npass; ps++) {
        int x0 = interlace ? PX0[ps] : 0, y0 = interlace ? PY0[ps] : 0, dx = interlace ? PDX[ps] : 1, dy = interlace ? PDY[ps] : 1;
        int pw = W > x0 ? (W - x0 + dx - 1) / dx : 0, ph = H > y0 ? (H - y0 + dy - 1) / dy : 0;
        if (pw && ph) total += ((pw * bpp_bits + 7) / 8 + 1) * (size_t)ph;
    }
    uint8_t* raw = malloc(total + 8);
    long got = zlib_inflate(z, zn, raw, total);
    free(z);
    if (got < (long)total) { free(raw); return -1; }
    size_t full_stride = (W * bpp_bits + 7) / 8;
    uint8_t* prev = calloc(full_stride + 1, 1);
    size_t off = 0;
    for (int ps = 0; ps < npass; ps++) {
        int x0 = interlace ? PX0[ps] : 0, y0 = interlace ? PY0[ps] : 0, dx = interlace ? PDX[ps] : 1, dy = interlace ? PDY[ps] : 1;
        int pw = W > x0 ? (W - x0 + dx - 1) / dx : 0, ph = H > y0 ? (H - y0 + dy - 1) / dy : 0;
        if (!pw || !ph) continue;
        size_t stride = (pw * bpp_bits + 7) / 8;
        memset(prev, 0, stride);
    for (int py = 0; py < ph; py++) {
        uint8_t* r = raw + off + (size_t)py * (stride + 1);
        int f = r[0];
        uint8_t* c = r + 1;
        for (size_t x = 0; x < stride; x++) {
            int a = x >= bpp ? c[x - bpp] : 0, b = prev[x], cc = x >= bpp ? prev[x - bpp] : 0, v = c[x];
            switch (f) {
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: v += paeth(a, b, cc); break;
                default: break;
            }
            c[x] = (uint8_t)v;
        }
        memcpy(prev, c, stride);
        uint8_t* orow = out + (size_t)(y0 + py * dy) * W * och;
        for (int px = 0; px < pw; px++) {
            uint8_t* o = orow + (size_t)(x0 + px * dx) * och - (size_t)px * och; /* o[px*och+k] lands on the pixel */
            int x = px;
            unsigned s[4] = {0, 0, 0, 0}; /* samples at full precision */
            for (int k = 0; k < src_ch; k++) {
                size_t bit = ((size_t)x * src_ch + k) * bd;
                if (bd == 16) s[k] = (unsigned)((c[bit / 8] << 8) | c[bit / 8 + 1]);
                else if (bd == 8) s[k] = c[bit / 8];
                else s[k] = (c[bit / 8] >> (8 - bd - (bit & 7))) & ((1u << bd) - 1);
            }
            if (ct == 3) {
                unsigned idx = s[0] < (unsigned)npal ? s[0] : 0;
                o[x * och + 0] = pal[idx][2]; o[x * och + 1] = pal[idx][1]; o[x * och + 2] = pal[idx][0];
                if (och == 4) o[x * 4 + 3] = s[0] < (unsigned)ntrns ? trns[s[0]] : 255;
                if (s[0] >= (unsigned)npal) { o[x * och] = o[x * och + 1] = o[x * och + 2] = 0; }
            } else if (ct == 0) {
                o[x] = bd == 16 ? (uint8_t)(s[0] >> 8) : bd == 8 ? (uint8_t)s[0] : (uint8_t)(s[0] * (255u / ((1u << bd) - 1)));
            } else if (ct == 4) {
                uint8_t g = bd == 16 ? (uint8_t)(s[0] >> 8) : (uint8_t)s[0];
                o[x * 4] = o[x * 4 + 1] = o[x * 4 + 2] = g;
                o[x * 4 + 3] = bd == 16 ? (uint8_t)(s[1] >> 8) : (uint8_t)s[1];
            } else { /* RGB / RGBA -> BGR(A) */
                int sh = bd == 16 ? 8 : 0;
                o[x * och + 0] = (uint8_t)(s[2] >> sh); o[x * och + 1] = (uint8_t)(s[1] >> sh); o[x * och + 2] = (uint8_t)(s[0] >> sh);
                if (ct == 6) o[x * 4 + 3] = (uint8_t)(s[3] >> sh);
                else if (och == 4) o[x * 4 + 3] = (s[0] == trns_rgb[0] && s[1] == trns_rgb[1] && s[2] == trns_rgb[2]) ? 0 : 255;
            }
        }
    }
    off += (stride + 1) * (size_t)ph;
    }
    free(prev);
    free(raw);
    return 0;
}
