// png_parse.cpp -- host-side PNG container parsing (W3C PNG: signature, chunk walk, IHDR / PLTE /
// tRNS / IDAT / iCCP) and a small zlib inflater used ONLY for the iCCP profile blob (metadata, a
// few KB).  Pixel data is inflated, defiltered and converted on the device (png_decode.cu).
//
// Stands where cv::ImageDecoder::readHeader does for PNG inputs (ref opencv.cpp:126-164) and
// where opencv_decoder_get_png_icc uses libpng's png_get_iCCP (ref opencv.cpp:315-345).
#include <cstring>
#include <vector>

#include "kernels.cuh"
#include "lilliput_b200.h"

namespace lp {

static inline uint32_t be32(const uint8_t* p) {
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
}
static const uint8_t kPngSig[8] = {0x89, 0x50, 0x4E, 0x47, 0x0D, 0x0A, 0x1A, 0x0A};

int png_parse(const uint8_t* in, size_t len, PngHeader* out) {
    PngHeader& h = *out;
    h = PngHeader();
    if (len < 8 + 25 || memcmp(in, kPngSig, 8) != 0) return LP_ERR_INVALID_IMAGE;
    size_t pos = 8;
    bool have_ihdr = false;
    while (pos + 12 <= len) {
        const uint32_t n = be32(in + pos);
        const uint8_t* type = in + pos + 4;
        const uint8_t* d = in + pos + 8;
        if (pos + 12 + (size_t)n > len) break;  // truncated chunk: keep what was seen
        if (!memcmp(type, "IHDR", 4) && n >= 13) {
            h.width = (int)be32(d);
            h.height = (int)be32(d + 4);
            h.bit_depth = d[8];
            h.color_type = d[9];
            h.interlace = d[12];
            have_ihdr = true;
        } else if (!memcmp(type, "PLTE", 4)) {
            h.npal = (int)std::min<uint32_t>(n / 3, 256);
            memcpy(h.palette, d, (size_t)h.npal * 3);
        } else if (!memcmp(type, "tRNS", 4)) {
            h.has_trns = true;
            if (h.color_type == 3) {
                h.ntrns = (int)std::min<uint32_t>(n, 256);
                memcpy(h.trns, d, h.ntrns);
            } else if (h.color_type == 2 && n >= 6) {
                for (int i = 0; i < 3; i++) h.trns_rgb[i] = (uint16_t)((d[2 * i] << 8) | d[2 * i + 1]);
            }
        } else if (!memcmp(type, "IDAT", 4)) {
            h.idat.push_back({pos + 8, (size_t)n});
            h.idat_total += n;
        } else if (!memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + (size_t)n;
    }
    if (!have_ihdr || h.width < 1 || h.height < 1) return LP_ERR_INVALID_IMAGE;
    switch (h.color_type) {
        case 0: h.src_channels = 1; break;
        case 2: h.src_channels = 3; break;
        case 3: h.src_channels = 1; break;
        case 4: h.src_channels = 2; break;
        case 6: h.src_channels = 4; break;
        default: return LP_ERR_INVALID_IMAGE;
    }
    const int bd = h.bit_depth;
    if (bd != 1 && bd != 2 && bd != 4 && bd != 8 && bd != 16) return LP_ERR_INVALID_IMAGE;
    // OpenCV's PNG reader: gray -> 1 channel; gray+alpha / RGBA -> 4; RGB / palette -> 3, or 4 with tRNS
    h.out_channels = h.color_type == 0 ? 1 : (h.color_type == 4 || h.color_type == 6) ? 4 : (h.has_trns ? 4 : 3);
    const size_t bits = (size_t)h.src_channels * bd;
    h.row_bytes = ((size_t)h.width * bits + 7) / 8;
    h.bpp = bits >= 8 ? (int)(bits / 8) : 1;
    return LP_OK;
}

// ---- tiny RFC 1950/1951 inflater for metadata blobs (host) ---------------------------------

namespace {
struct Bits {
    const uint8_t* p;
    size_t n, pos = 0;
    uint64_t acc = 0;
    int cnt = 0;
    unsigned get(int k) {
        while (cnt < k) {
            acc |= (uint64_t)(pos < n ? p[pos] : 0) << cnt;
            pos++;
            cnt += 8;
        }
        unsigned v = (unsigned)(acc & ((1ull << k) - 1));
        acc >>= k;
        cnt -= k;
        return k ? v : 0;
    }
};
struct Canon {
    uint16_t count[16], sym[320];
    bool build(const uint8_t* len, int n) {
        memset(count, 0, sizeof(count));
        for (int i = 0; i < n; i++) count[len[i]]++;
        count[0] = 0;
        int left = 1;
        for (int l = 1; l < 16; l++) {
            left = (left << 1) - count[l];
            if (left < 0) return false;
        }
        uint16_t offs[16];
        offs[1] = 0;
        for (int l = 1; l < 15; l++) offs[l + 1] = offs[l] + count[l];
        for (int i = 0; i < n; i++)
            if (len[i]) sym[offs[len[i]]++] = (uint16_t)i;
        return true;
    }
    int decode(Bits& b) const {
        int code = 0, first = 0, index = 0;
        for (int l = 1; l < 16; l++) {
            code |= (int)b.get(1);
            int c = count[l];
            if (code - c < first) return sym[index + (code - first)];
            index += c;
            first = (first + c) << 1;
            code <<= 1;
        }
        return -1;
    }
};
const uint16_t kLBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t kLExt[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t kDExt[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
}  // namespace

// Inflates a zlib stream into `out` (at most cap bytes).  Returns bytes produced or -1.
long host_zlib_inflate(const uint8_t* in, size_t n, uint8_t* out, size_t cap) {
    if (n < 2 || (in[0] & 15) != 8 || (((unsigned)in[0] << 8) | in[1]) % 31 || (in[1] & 0x20)) return -1;
    Bits b{in + 2, n - 2};
    size_t o = 0;
    int last;
    do {
        last = (int)b.get(1);
        const int type = (int)b.get(2);
        if (type == 0) {
            b.get(b.cnt & 7);
            unsigned len = b.get(16), nlen = b.get(16);
            if ((len ^ 0xFFFF) != nlen || o + len > cap) return -1;
            for (unsigned i = 0; i < len; i++) out[o++] = (uint8_t)b.get(8);
        } else if (type == 1 || type == 2) {
            Canon hl, hd;
            uint8_t lens[320];
            if (type == 1) {
                int i = 0;
                for (; i < 144; i++) lens[i] = 8;
                for (; i < 256; i++) lens[i] = 9;
                for (; i < 280; i++) lens[i] = 7;
                for (; i < 288; i++) lens[i] = 8;
                hl.build(lens, 288);
                for (i = 0; i < 30; i++) lens[i] = 5;
                hd.build(lens, 30);
            } else {
                static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
                int nl = (int)b.get(5) + 257, nd = (int)b.get(5) + 1, nc = (int)b.get(4) + 4;
                if (nl > 286 || nd > 30) return -1;
                uint8_t cl[19] = {0};
                for (int i = 0; i < nc; i++) cl[order[i]] = (uint8_t)b.get(3);
                Canon hc;
                if (!hc.build(cl, 19)) return -1;
                int i = 0;
                while (i < nl + nd) {
                    int s = hc.decode(b);
                    if (s < 0) return -1;
                    if (s < 16) { lens[i++] = (uint8_t)s; continue; }
                    int rep, v = 0;
                    if (s == 16) { if (!i) return -1; v = lens[i - 1]; rep = 3 + (int)b.get(2); }
                    else if (s == 17) rep = 3 + (int)b.get(3);
                    else rep = 11 + (int)b.get(7);
                    if (i + rep > nl + nd) return -1;
                    while (rep--) lens[i++] = (uint8_t)v;
                }
                if (!hl.build(lens, nl)) return -1;
                hd.build(lens + nl, nd);
            }
            for (;;) {
                int s = hl.decode(b);
                if (s < 0) return -1;
                if (s < 256) { if (o >= cap) return -1; out[o++] = (uint8_t)s; }
                else if (s == 256) break;
                else {
                    s -= 257;
                    if (s >= 29) return -1;
                    unsigned len = kLBase[s] + b.get(kLExt[s]);
                    int ds = hd.decode(b);
                    if (ds < 0 || ds >= 30) return -1;
                    unsigned dist = kDBase[ds] + b.get(kDExt[ds]);
                    if (dist > o || o + len > cap) return -1;
                    for (unsigned i = 0; i < len; i++, o++) out[o] = out[o - dist];
                }
            }
        } else {
            return -1;
        }
    } while (!last);
    return (long)o;
}

// iCCP: keyword\0 method(0) zlib(profile).  Returns profile length copied into dest, or 0.
int png_extract_icc(const uint8_t* in, size_t len, uint8_t* dest, size_t dest_len) {
    if (len < 8 || memcmp(in, kPngSig, 8) != 0) return 0;
    size_t pos = 8;
    while (pos + 12 <= len) {
        const uint32_t n = be32(in + pos);
        const uint8_t* type = in + pos + 4;
        const uint8_t* d = in + pos + 8;
        if (pos + 12 + (size_t)n > len) return 0;
        if (!memcmp(type, "IDAT", 4) || !memcmp(type, "IEND", 4)) return 0;  // iCCP precedes IDAT
        if (!memcmp(type, "iCCP", 4)) {
            size_t k = 0;
            while (k < n && k < 80 && d[k]) k++;
            if (k + 2 >= n || d[k + 1] != 0) return 0;
            std::vector<uint8_t> buf(dest_len + 1);
            long got = host_zlib_inflate(d + k + 2, n - k - 2, buf.data(), buf.size());
            if (got <= 0 || (size_t)got > dest_len) return 0;
            memcpy(dest, buf.data(), (size_t)got);
            return (int)got;
        }
        pos += 12 + (size_t)n;
    }
    return 0;
}

}  // namespace lp
