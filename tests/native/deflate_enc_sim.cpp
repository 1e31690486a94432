// Host build of lilliput_b200/csrc/deflate_enc_core.h (the PNG encoder's DEFLATE writer) for the CPU suite:
// a whole zlib stream made the way png_encode.cu makes it -- 78 01, one chunk after the other, a final empty fixed
// block, Adler-32 -- so tests/test_deflate_enc_core.py can hand it to zlib's inflate.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../lilliput_b200/csrc/deflate_enc_core.h"

extern "C" long defenc_compress(const uint8_t* src, long n, int level, uint8_t* out, long cap) {
    std::vector<uint16_t> tok(defenc::kTokCap), prev(defenc::kChunk);
    std::vector<uint8_t> chunk(defenc::kChunkOut);
    defenc::Work* w = new defenc::Work;
    long o = 0;
    if (cap < 2) return -1;
    out[o++] = 0x78;
    out[o++] = 0x01;
    uint32_t s1 = 1, s2 = 0;
    for (long at = 0; at < n; at += defenc::kChunk) {
        const int len = (int)(n - at < defenc::kChunk ? n - at : defenc::kChunk);
        const size_t got = defenc::write_chunk(src + at, len, level, *w, prev.data(), tok.data(), chunk.data());
        if (got > (size_t)defenc::kChunkOut || o + (long)got > cap) { delete w; return -1; }
        memcpy(out + o, chunk.data(), got);
        o += (long)got;
        for (int i = 0; i < len; i++) {
            s1 = (s1 + src[at + i]) % 65521u;
            s2 = (s2 + s1) % 65521u;
        }
    }
    delete w;
    if (o + 6 > cap) return -1;
    out[o++] = 0x03;  // BFINAL = 1, BTYPE = 01, end of block
    out[o++] = 0x00;
    out[o++] = (uint8_t)(s2 >> 8);
    out[o++] = (uint8_t)s2;
    out[o++] = (uint8_t)(s1 >> 8);
    out[o++] = (uint8_t)s1;
    return o;
}
