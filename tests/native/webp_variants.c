/* webp_variants.c -- TEST INFRASTRUCTURE (dev container only).  Encodes an image with the
 * reference's vendored libwebp under a sweep of encoder configurations, so that the VP8 streams
 * exercise every syntax element the decoder has to handle (simple / normal loop filter,
 * sharpness, 1..8 token partitions, 1..4 segments, skip flags, all intra modes), and decodes each
 * with the same libwebp for the expected pixels.  Built by tests/golden/make_golden_webp.py
 * against /root/reference/deps -- never shipped, never run on the GPU box.
 *
 *   int lpv_encode(const uint8_t* bgr, int w, int h, int stride, int has_alpha, float quality,
 *                  int method, int filter_type, int filter_strength, int sharpness, int partitions,
 *                  int segments, int sns, int alpha_compression, int lossless,
 *                  uint8_t* out, size_t cap);
 *   int lpv_decode(const uint8_t* webp, size_t n, uint8_t* out, int stride, int channels);
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <webp/decode.h>
#include <webp/encode.h>

int lpv_encode(const uint8_t* px, int w, int h, int stride, int has_alpha, float quality, int method,
               int filter_type, int filter_strength, int sharpness, int partitions, int segments, int sns,
               int alpha_compression, int lossless, uint8_t* out, size_t cap) {
    WebPConfig cfg;
    WebPPicture pic;
    WebPMemoryWriter wr;
    if (!WebPConfigPreset(&cfg, WEBP_PRESET_DEFAULT, quality)) return -1;
    cfg.method = method;
    cfg.filter_type = filter_type;
    cfg.filter_strength = filter_strength;
    cfg.filter_sharpness = sharpness;
    cfg.partitions = partitions;
    cfg.segments = segments;
    cfg.sns_strength = sns;
    cfg.alpha_compression = alpha_compression;
    cfg.lossless = lossless;
    if (!WebPValidateConfig(&cfg)) return -2;
    if (!WebPPictureInit(&pic)) return -3;
    pic.width = w;
    pic.height = h;
    pic.use_argb = lossless;
    if (!(has_alpha ? WebPPictureImportBGRA(&pic, px, stride) : WebPPictureImportBGR(&pic, px, stride))) return -4;
    WebPMemoryWriterInit(&wr);
    pic.writer = WebPMemoryWrite;
    pic.custom_ptr = &wr;
    int ok = WebPEncode(&cfg, &pic);
    WebPPictureFree(&pic);
    if (!ok || wr.size > cap) {
        WebPMemoryWriterClear(&wr);
        return -5;
    }
    memcpy(out, wr.mem, wr.size);
    int n = (int)wr.size;
    WebPMemoryWriterClear(&wr);
    return n;
}

int lpv_decode(const uint8_t* webp, size_t n, uint8_t* out, int stride, int channels) {
    int w, h;
    if (!WebPGetInfo(webp, n, &w, &h)) return -1;
    uint8_t* r = channels == 4 ? WebPDecodeBGRAInto(webp, n, out, (size_t)stride * h, stride)
                               : WebPDecodeBGRInto(webp, n, out, (size_t)stride * h, stride);
    return r ? 0 : -2;
}
