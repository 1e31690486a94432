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

static bool png_chunk_crc_ok(const uint8_t* type, uint32_t n);

int png_parse(const uint8_t* in, size_t len, PngHeader* out) {
    PngHeader& h = *out;
    h = PngHeader();
    if (len < 8 + 25 || memcmp(in, kPngSig, 8) != 0) return LP_ERR_INVALID_IMAGE;
    size_t pos = 8;
    bool have_ihdr = false, have_exif = false, have_plte = false;
    while (pos + 12 <= len) {
        const uint32_t n = be32(in + pos);
        const uint8_t* type = in + pos + 4;
        const uint8_t* d = in + pos + 8;
        if (pos + 12 + (size_t)n > len) {
            // OpenCV's PngDecoder::readHeader pulls in whole chunks up to and including the first IDAT: a file that
            // ends before that chunk is complete is refused at the header.  Behind it: keep what was seen, the
            // decode reports the short stream.
            if (h.idat.empty()) return LP_ERR_INVALID_IMAGE;
            break;
        }
        // Header chunks are taken the way libpng 1.6.47 takes them for the reference (png_read_info under OpenCV's
        // PngDecoder::readHeader; pinned by tests/test_host_png_header.py): what it refuses, readHeader refuses.
        if (h.idat.empty()) {  // everything in front of the image data is read by png_read_info
            for (int i = 0; i < 4; i++)
                if (!((type[i] >= 'A' && type[i] <= 'Z') || (type[i] >= 'a' && type[i] <= 'z'))) return LP_ERR_INVALID_IMAGE;
            if (type[2] & 0x20) return LP_ERR_INVALID_IMAGE;  // reserved bit: "bad header (invalid type)"
            if (!(type[0] & 0x20) && memcmp(type, "IHDR", 4) && memcmp(type, "PLTE", 4) && memcmp(type, "IDAT", 4))
                return LP_ERR_INVALID_IMAGE;  // IEND out of place, or a critical chunk nobody knows
        }
        if (!h.idat.empty() && memcmp(type, "IDAT", 4) && memcmp(type, "IEND", 4)) {
            // behind the first IDAT only more image data and the end matter to this parser (the header is complete)
        } else if (!memcmp(type, "IHDR", 4)) {
            if (have_ihdr || pos != 8 || n != 13 || !png_chunk_crc_ok(type, n)) return LP_ERR_INVALID_IMAGE;
            const uint32_t w = be32(d), hh = be32(d + 4);
            const int bd = d[8], ct = d[9];
            if (w == 0 || hh == 0 || w > 1000000u || hh > 1000000u) return LP_ERR_INVALID_IMAGE;  // png_check_IHDR, default user limits
            if (bd != 1 && bd != 2 && bd != 4 && bd != 8 && bd != 16) return LP_ERR_INVALID_IMAGE;
            if (ct == 1 || ct == 5 || ct > 6) return LP_ERR_INVALID_IMAGE;
            if ((ct == 3 && bd > 8) || ((ct == 2 || ct == 4 || ct == 6) && bd < 8)) return LP_ERR_INVALID_IMAGE;
            if (d[10] != 0 || d[11] != 0 || d[12] > 1) return LP_ERR_INVALID_IMAGE;
            h.width = (int)w;
            h.height = (int)hh;
            h.bit_depth = bd;
            h.color_type = ct;
            h.interlace = d[12];
            have_ihdr = true;
        } else if (!have_ihdr) {
            return LP_ERR_INVALID_IMAGE;  // "Missing IHDR before ..."
        } else if (!memcmp(type, "PLTE", 4)) {
            if (!h.idat.empty() || !(h.color_type & 2)) {
                // after the image data, or in a grayscale image: skipped
            } else if (h.color_type == 3) {
                if (have_plte || n > 3 * 256 || n % 3 || n == 0 || !png_chunk_crc_ok(type, n)) return LP_ERR_INVALID_IMAGE;
                have_plte = true;
                h.npal = (int)std::min<uint32_t>(n / 3, 1u << h.bit_depth);  // entries past 2^depth are dropped
                memcpy(h.palette, d, (size_t)h.npal * 3);
            } else if (!have_plte && !(n > 3 * 256 || n % 3)) {
                if (n == 0) return LP_ERR_INVALID_IMAGE;  // "Invalid palette"
                have_plte = true;  // a suggested palette: nothing is decoded with it, but tRNS may not follow... (see below)
            }
        } else if (!memcmp(type, "tRNS", 4)) {
            // kept only in front of the image data, once, with a good CRC and the length its colour type asks for
            if (h.idat.empty() && !h.has_trns && png_chunk_crc_ok(type, n)) {
                if (h.color_type == 3) {
                    if (have_plte && n >= 1 && n <= (uint32_t)h.npal && n <= 256) {
                        h.has_trns = true;
                        h.ntrns = (int)n;
                        memcpy(h.trns, d, h.ntrns);
                    }
                } else if (h.color_type == 2) {
                    if (n == 6) {
                        h.has_trns = true;
                        for (int i = 0; i < 3; i++) h.trns_rgb[i] = (uint16_t)((d[2 * i] << 8) | d[2 * i + 1]);
                    }
                } else if (h.color_type == 0) {
                    if (n == 2) h.has_trns = true;
                }
            }
        } else if (!memcmp(type, "eXIf", 4)) {
            // OpenCV's PNG reader hands libpng's eXIf block (png_get_eXIf_1 after png_read_info, so only a chunk in
            // front of the first IDAT) to the same ExifReader as a JPEG's APP1.  libpng keeps the first chunk whose
            // CRC is good and which starts with a proper byte-order mark ("II" / "MM").
            if (!have_exif && h.idat.empty() && n >= 2 && d[0] == d[1] && (d[0] == 'I' || d[0] == 'M') &&
                png_chunk_crc_ok(type, n)) {
                have_exif = true;
                int o = 0;
                if (exif_orientation_opencv(d, n, &o)) h.orientation = o;
            }
        } else if (!memcmp(type, "IDAT", 4)) {
            if (h.idat.empty() && h.color_type == 3 && !have_plte) return LP_ERR_INVALID_IMAGE;  // "Missing PLTE before IDAT"
            h.idat.push_back({pos + 8, (size_t)n});
            h.idat_total += n;
        } else if (!memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + (size_t)n;
    }
    if (!have_ihdr || h.width < 1 || h.height < 1) return LP_ERR_INVALID_IMAGE;
    if (h.idat.empty()) return LP_ERR_INVALID_IMAGE;  // the data ended before a first IDAT chunk was complete
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

// Inflates a zlib stream into `out`.  Stops as soon as `cap` bytes exist and returns cap -- the caller asks for
// exactly the bytes it needs, the way libpng's png_inflate_read does, and like zlib it then looks at nothing after
// them (no end-of-stream, no Adler-32).  Returns fewer than cap when the stream ends first, -1 when it is malformed
// or runs out of input.
long host_zlib_inflate(const uint8_t* in, size_t n, uint8_t* out, size_t cap) {
    if (n < 2 || (in[0] & 15) != 8 || (in[0] >> 4) > 7 || (((unsigned)in[0] << 8) | in[1]) % 31 || (in[1] & 0x20)) return -1;
    const unsigned window = 1u << ((in[0] >> 4) + 8);
    if (cap == 0) return 0;
    Bits b{in + 2, n - 2};
    size_t o = 0;
    int last;
    do {
        last = (int)b.get(1);
        const int type = (int)b.get(2);
        if (b.pos > b.n) return -1;
        if (type == 0) {
            b.get(b.cnt & 7);
            unsigned len = b.get(16), nlen = b.get(16);
            if (b.pos > b.n || (len ^ 0xFFFF) != nlen) return -1;
            for (unsigned i = 0; i < len; i++) {
                out[o++] = (uint8_t)b.get(8);
                if (b.pos > b.n) return -1;
                if (o == cap) return (long)o;
            }
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
                    if (s < 0 || b.pos > b.n) return -1;
                    if (s < 16) { lens[i++] = (uint8_t)s; continue; }
                    int rep, v = 0;
                    if (s == 16) { if (!i) return -1; v = lens[i - 1]; rep = 3 + (int)b.get(2); }
                    else if (s == 17) rep = 3 + (int)b.get(3);
                    else rep = 11 + (int)b.get(7);
                    if (i + rep > nl + nd) return -1;
                    while (rep--) lens[i++] = (uint8_t)v;
                }
                if (b.pos > b.n || lens[256] == 0) return -1;  // zlib: "missing end-of-block"
                if (!hl.build(lens, nl)) return -1;
                hd.build(lens + nl, nd);
            }
            for (;;) {
                int s = hl.decode(b);
                if (s < 0 || b.pos > b.n) return -1;
                if (s < 256) {
                    out[o++] = (uint8_t)s;
                    if (o == cap) return (long)o;
                } else if (s == 256) {
                    break;
                } else {
                    s -= 257;
                    if (s >= 29) return -1;
                    unsigned len = kLBase[s] + b.get(kLExt[s]);
                    int ds = hd.decode(b);
                    if (ds < 0 || ds >= 30) return -1;
                    unsigned dist = kDBase[ds] + b.get(kDExt[ds]);
                    if (b.pos > b.n || dist > o || dist > window) return -1;
                    for (unsigned i = 0; i < len; i++, o++) {
                        out[o] = out[o - dist];
                        if (o + 1 == cap) return (long)cap;
                    }
                }
            }
        } else {
            return -1;
        }
    } while (!last);
    return (long)o;
}

// ---- iCCP, with the acceptance rules of the reference's libpng 1.6.47 (png_handle_iCCP, png_icc_check_length /
//      _header / _tag_table): what that library refuses to store, opencv_decoder_get_png_icc reports as "no
//      profile" (ref opencv.cpp:315-345), and so must this.

static inline uint32_t fourcc(const char* s) { return be32(reinterpret_cast<const uint8_t*>(s)); }

// CRC-32 of a chunk's type + body against the stored value (a CRC error in a critical chunk is a png_error).
static bool png_chunk_crc_ok(const uint8_t* type, uint32_t n) {
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < 4 + (size_t)n; i++) {
        c ^= type[i];
        for (int k = 0; k < 8; k++) c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1)));
    }
    return ~c == be32(type + 4 + n);
}

// One iCCP chunk body -> the validated profile, or false.
static bool png_read_iccp_chunk(const uint8_t* d, uint32_t n, int color_type, std::vector<uint8_t>* out) {
    const uint32_t read_length = n < 81 ? n : 81;  // keyword + separator + method: at most 81 bytes
    if (n - read_length < 11) return false;        // "too short": libpng wants 11 bytes of zlib stream AFTER those 81
    uint32_t k = 0;
    while (k < 80 && k < read_length && d[k]) k++;
    if (k < 1 || k > 79) return false;                          // "bad keyword"
    if (!(k + 1 < read_length && d[k + 1] == 0)) return false;  // "bad compression method"
    const uint8_t* z = d + k + 2;
    const size_t zn = n - k - 2;
    uint8_t head[132];
    if (host_zlib_inflate(z, zn, head, sizeof head) != (long)sizeof head) return false;  // profile truncated
    const uint32_t plen = be32(head);
    if (plen < 132 || plen > 8000000u) return false;            // "too short" / over PNG_USER_CHUNK_MALLOC_MAX
    if (head[8] > 3 && (plen & 3)) return false;                // "invalid length" (ICC v4+: multiple of 4)
    const uint32_t tags = be32(head + 128);
    if (tags > 357913930u || (uint64_t)plen < 132 + 12ull * tags) return false;  // "tag count too large"
    if (be32(head + 64) >= 0xffff) return false;                // "invalid rendering intent"
    if (be32(head + 36) != fourcc("acsp")) return false;        // "invalid signature"
    const uint32_t space = be32(head + 16);
    if (space == fourcc("RGB ")) {
        if (!(color_type & 2)) return false;                    // RGB profile on a grayscale PNG
    } else if (space == fourcc("GRAY")) {
        if (color_type & 2) return false;                       // gray profile on a colour PNG
    } else {
        return false;                                           // "invalid ICC profile color space"
    }
    const uint32_t cls = be32(head + 12);
    if (cls == fourcc("abst") || cls == fourcc("link")) return false;
    const uint32_t pcs = be32(head + 20);
    if (pcs != fourcc("XYZ ") && pcs != fourcc("Lab ")) return false;  // "unexpected ICC PCS encoding"
    std::vector<uint8_t> prof(plen);
    if (host_zlib_inflate(z, zn, prof.data(), plen) != (long)plen) return false;  // truncated; extra data is allowed
    for (uint32_t t = 0; t < tags; t++) {
        const uint8_t* tag = prof.data() + 132 + 12 * (size_t)t;
        const uint32_t start = be32(tag + 4), length = be32(tag + 8);
        if (start > plen || length > plen - start) return false;  // "ICC profile tag outside profile"
    }
    out->swap(prof);
    return true;
}

// The part of a PNG that png_read_info reads -- everything up to the first IDAT header -- walked with libpng
// 1.6.47's rules for what ends the read with an error.  `visit(type, body, n, colour_type, after_plte)` is called
// for every ancillary chunk.  Returns false when png_read_info would have failed (png_error -> longjmp), which
// makes every getter built on it answer "nothing".
template <class F>
static bool png_walk_info(const uint8_t* in, size_t len, F&& visit) {
    if (len < 8 || memcmp(in, kPngSig, 8) != 0) return false;
    size_t pos = 8;
    int color_type = -1;
    bool seen_plte = false;
    for (;;) {
        if (pos + 8 > len) return false;  // ran off the data before IDAT: libpng's read callback errors out
        const uint32_t n = be32(in + pos);
        const uint8_t* type = in + pos + 4;
        const uint8_t* d = in + pos + 8;
        if (n > 0x7fffffffu) return false;  // "PNG unsigned integer out of range"
        for (int i = 0; i < 4; i++)
            if (!((type[i] >= 'A' && type[i] <= 'Z') || (type[i] >= 'a' && type[i] <= 'z'))) return false;  // "invalid chunk type"
        if (type[2] & 0x20) return false;  // reserved bit set: "bad header (invalid type)"
        if (color_type < 0) {  // the first chunk has to be a well-formed IHDR (png_check_IHDR)
            if (memcmp(type, "IHDR", 4) != 0 || n != 13 || pos + 12 + 13 > len) return false;
            if (!png_chunk_crc_ok(type, n)) return false;
            const uint32_t w = be32(d), h = be32(d + 4);
            const int bd = d[8], ct = d[9];
            if (w == 0 || h == 0 || w > 1000000u || h > 1000000u) return false;  // libpng's default user limits
            if (bd != 1 && bd != 2 && bd != 4 && bd != 8 && bd != 16) return false;
            if (ct == 1 || ct == 5 || ct > 6) return false;
            if ((ct == 3 && bd > 8) || ((ct == 2 || ct == 4 || ct == 6) && bd < 8)) return false;
            if (d[10] != 0 || d[11] != 0 || d[12] > 1) return false;
            color_type = ct;
            pos += 25;
            continue;
        }
        if (!memcmp(type, "IDAT", 4)) return true;  // png_read_info stops at the first IDAT header
        if (pos + 12 + (size_t)n > len) return false;
        if (!memcmp(type, "IEND", 4) || !memcmp(type, "IHDR", 4)) return false;  // out of place: png_chunk_error
        if (!memcmp(type, "PLTE", 4)) {  // png_handle_PLTE, as libpng 1.6.47 behaves
            if (!(color_type & 2)) {
                // "ignored in grayscale PNG": skipped whatever it holds, and it does not count as a palette
            } else if (color_type == 3) {
                // the palette of a palette image is critical: duplicate, bad length, empty or CRC error all end the read
                if (seen_plte || n > 3 * 256 || n % 3 || n == 0 || !png_chunk_crc_ok(type, n)) return false;
                seen_plte = true;
            } else if (seen_plte) {
                // a suggested palette in an RGB(A) image is treated like an ancillary chunk: "duplicate" is skipped
            } else if (n > 3 * 256 || n % 3) {
                // "invalid": skipped, and a later PLTE is still the first
            } else {
                if (n == 0) return false;  // png_set_PLTE: "Invalid palette" is a png_error for every colour type
                seen_plte = true;          // a CRC error here is only a warning and the chunk still counts
            }
        } else if (!(type[0] & 0x20)) {
            return false;  // an unknown critical chunk ends the read with an error
        } else {
            visit(type, d, n, color_type, seen_plte);
        }
        pos += 12 + (size_t)n;
    }
}

// What png_read_info + png_get_iCCP leave in the caller's hands (ref opencv.cpp:315-345): the first profile libpng
// would have stored.  Returns the profile length copied into dest, or 0.
int png_extract_icc(const uint8_t* in, size_t len, uint8_t* dest, size_t dest_len) {
    std::vector<uint8_t> profile;
    const bool ok = png_walk_info(in, len, [&](const uint8_t* type, const uint8_t* d, uint32_t n, int color_type, bool after_plte) {
        // after PLTE the chunk is out of place; once a profile has been stored a further chunk is a duplicate:
        // both are skipped.  A chunk that was refused does not stop a later one from being taken.
        if (!memcmp(type, "iCCP", 4) && profile.empty() && !after_plte) png_read_iccp_chunk(d, n, color_type, &profile);
    });
    if (!ok || profile.empty() || profile.size() > dest_len) return 0;
    memcpy(dest, profile.data(), profile.size());
    return (int)profile.size();
}

// png_read_info + png_get_cICP (ref opencv.cpp:347-395): the four code points of the first cICP chunk libpng
// accepts.  A chunk that was read (4 bytes, CRC good) counts as THE cICP chunk even when png_set_cICP then refuses
// it for a non-zero matrix; a later one is a duplicate.  Returns 1 and fills out[4], or 0.
int png_extract_cicp(const uint8_t* in, size_t len, uint8_t* out) {
    bool seen = false, found = false;
    const bool ok = png_walk_info(in, len, [&](const uint8_t* type, const uint8_t* d, uint32_t n, int, bool after_plte) {
        if (memcmp(type, "cICP", 4) != 0 || seen || after_plte) return;
        if (n != 4 || !png_chunk_crc_ok(type, n)) return;
        seen = true;
        if (d[2] != 0) return;  // "Invalid cICP matrix coefficients": RGB data only
        memcpy(out, d, 4);
        found = true;
    });
    return ok && found ? 1 : 0;
}

}  // namespace lp
