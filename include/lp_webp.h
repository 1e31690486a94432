/*
 * lp_webp.h -- the cgo surface of lilliput's WebP adapter as exported by liblilliput_b200.so.
 * Same names, signatures and return conventions as the reference's webp.hpp (cited per symbol).
 *
 * Decode: the RIFF container (VP8 / VP8L / VP8X / ICCP / ANIM / ANMF / ALPH) is walked on the
 * host, standing where libwebpmux does for the reference (ref webp.cpp:61-139); the frames are
 * decoded on the device, bit-exact to WebPDecodeBGRInto / WebPDecodeBGRAInto (ref webp.cpp:336-351):
 *   VP8  (lossy key frames): boolean-coded modes and tokens, inverse transforms, intra prediction,
 *        loop filter, libwebp's "fancy" upsampler and fixed-point YUV->BGR;
 *   VP8L (lossless): prefix codes + meta prefix image, colour cache, LZ77, the four transforms;
 *   ALPH (alpha plane of a lossy frame): raw or VP8L-coded, with its prediction filters.
 *
 * Encode (ref webp.cpp:388-783): lossless (quality > 100) and lossy stills, alpha through an ALPH
 * chunk, ICC, and animations as full-canvas ANMF frames, all encoded on the device by this
 * library's own VP8L / VP8 encoders: lossless output is pixel-exact; lossy output is a valid stream at
 * libwebp's quality->quantiser mapping, not libwebp's bytes.
 */
#ifndef LP_WEBP_H
#define LP_WEBP_H

#include "lp_opencv.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ref webp.hpp:13-23 */
enum WebpEncoderOptions {
    WEBP_METHOD = 1000,
    WEBP_FILTER_STRENGTH = 1001,
    WEBP_FILTER_TYPE = 1002,
    WEBP_AUTOFILTER = 1003,
    WEBP_PARTITIONS = 1004,
    WEBP_SEGMENTS = 1005,
    WEBP_PREPROCESSING = 1006,
    WEBP_THREAD_LEVEL = 1007,
    WEBP_PALETTE = 1008
};

/* ref webp.hpp:28-29 */
typedef struct webp_decoder_struct* webp_decoder;
typedef struct webp_encoder_struct* webp_encoder;

/* ref webp.hpp:35-51 */
webp_decoder webp_decoder_create(const opencv_mat buf);
int webp_decoder_get_width(const webp_decoder d);
int webp_decoder_get_height(const webp_decoder d);
int webp_decoder_get_pixel_type(const webp_decoder d);
int webp_decoder_get_num_frames(const webp_decoder d);
int webp_decoder_get_total_duration(const webp_decoder d);
int webp_decoder_get_prev_frame_delay(const webp_decoder d);
int webp_decoder_get_prev_frame_dispose(const webp_decoder d);
int webp_decoder_get_prev_frame_blend(const webp_decoder d);
int webp_decoder_get_prev_frame_x_offset(const webp_decoder d);
int webp_decoder_get_prev_frame_y_offset(const webp_decoder d);
bool webp_decoder_get_prev_frame_has_alpha(const webp_decoder d);
uint32_t webp_decoder_get_bg_color(const webp_decoder d);
uint32_t webp_decoder_get_loop_count(const webp_decoder d);
size_t webp_decoder_get_icc(const webp_decoder d, void* buf, size_t buf_len);
void webp_decoder_release(webp_decoder d);
bool webp_decoder_decode(webp_decoder d, opencv_mat mat);

/* ref webp.hpp:56-73 */
webp_encoder webp_encoder_create(void* buf, size_t buf_len, const void* icc, size_t icc_len,
                                 uint32_t bgcolor, int loop_count);
size_t webp_encoder_write(webp_encoder e, const opencv_mat src, const int* opt, size_t opt_len,
                          int delay, int blend, int dispose, int x_offset, int y_offset);
void webp_encoder_release(webp_encoder e);
size_t webp_encoder_flush(webp_encoder e);
/* ref webp.hpp:74-75 */
void webp_decoder_advance_frame(webp_decoder d);
int webp_decoder_has_more_frames(webp_decoder d);

#ifdef __cplusplus
}
#endif

#endif
