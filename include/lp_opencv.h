/*
 * lp_opencv.h -- the per-image cgo surface of lilliput's OpenCV adapter, as
 * exported by liblilliput_b200.so.
 *
 * Every symbol below has the same name, argument list and return convention
 * as the declaration it replaces in the reference's opencv.hpp (cited per
 * function as "ref opencv.hpp:LINE"), so lilliput's opencv.go / ops.go bind to
 * it unchanged; only the `#cgo LDFLAGS` line of cgo.go differs (see
 * INTEGRATION.md).  Unlike the reference header this file does NOT include
 * any OpenCV header: the handles are opaque and the few OpenCV constants the
 * Go side reads (CV_8U, CV_8UC3, CV_8UC4) are restated here.
 *
 * Behavioural contract (what differs behind the boundary): a `opencv_mat`
 * that wraps caller memory (`opencv_mat_create_from_data`) is mirrored in
 * HBM; decode / orient / resize / blend / clear / encode run as sm_100a CUDA
 * kernels on the mirror, and the host bytes are refreshed lazily
 * (`opencv_mat_get_data` and `lp_mat_sync_host`).  There is no CPU fallback:
 * if no CUDA device can be initialised every pixel-touching entry point fails
 * (NULL / false / OPENCV_ERROR_UNKNOWN) and logs to stderr.
 */
#ifndef LP_OPENCV_H
#define LP_OPENCV_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ref opencv.hpp:17-26 */
typedef enum CVImageOrientation {
    CV_IMAGE_ORIENTATION_TL = 1,
    CV_IMAGE_ORIENTATION_TR = 2,
    CV_IMAGE_ORIENTATION_BR = 3,
    CV_IMAGE_ORIENTATION_BL = 4,
    CV_IMAGE_ORIENTATION_LT = 5,
    CV_IMAGE_ORIENTATION_RT = 6,
    CV_IMAGE_ORIENTATION_RB = 7,
    CV_IMAGE_ORIENTATION_LB = 8
} CVImageOrientation;

/* ref opencv.hpp:33-36 (keys of the flat k,v option array) */
#define CV_IMWRITE_JPEG_QUALITY 1
#define CV_IMWRITE_JPEG_PROGRESSIVE 2
#define CV_IMWRITE_PNG_COMPRESSION 16
#define CV_IMWRITE_WEBP_QUALITY 64

/* OpenCV type codes read by opencv.go (opencv2/core/hal/interface.h):
 * type = depth + ((channels - 1) << 3), depth 0 = 8U, 2 = 16U. */
#ifndef CV_8U
#define CV_8U 0
#define CV_16U 2
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_8UC4 24
#endif

/* ref opencv.hpp:53-55; values are OpenCV's INTER_* (1, 2, 3). */
extern const int CV_INTER_AREA;
extern const int CV_INTER_LINEAR;
extern const int CV_INTER_CUBIC;

/* ref opencv.hpp:57-59 */
typedef void* opencv_mat;
typedef void* opencv_decoder;
typedef void* opencv_encoder;

/* ref opencv.hpp:61-63 */
int opencv_type_depth(int type);
int opencv_type_channels(int type);
int opencv_type_convert_depth(int type, int depth);

/* ref opencv.hpp:65-74.  `buf` is a 1-row CV_8U mat over the compressed
 * bytes (borrowed).  Returns NULL when the signature is not JPEG / PNG. */
opencv_decoder opencv_decoder_create(const opencv_mat buf);
const char* opencv_decoder_get_description(const opencv_decoder d);
void opencv_decoder_release(opencv_decoder d);
bool opencv_decoder_set_source(opencv_decoder d, const opencv_mat buf);
bool opencv_decoder_read_header(opencv_decoder d);
int opencv_decoder_get_width(const opencv_decoder d);
int opencv_decoder_get_height(const opencv_decoder d);
int opencv_decoder_get_pixel_type(const opencv_decoder d);
int opencv_decoder_get_orientation(const opencv_decoder d);
bool opencv_decoder_read_data(opencv_decoder d, opencv_mat dst);

/* ref opencv.hpp:75-86 */
int opencv_copy_to_region_with_alpha(opencv_mat src, opencv_mat dst, int xOffset, int yOffset,
                                     int width, int height);
int opencv_copy_to_region(opencv_mat src, opencv_mat dst, int xOffset, int yOffset, int width,
                          int height);
/* ref opencv.hpp:87-94 */
void opencv_mat_set_color(opencv_mat, int red, int green, int blue, int alpha);
void opencv_mat_reset(opencv_mat mat);
int opencv_mat_clear_to_transparent(opencv_mat mat, int xOffset, int yOffset, int width,
                                    int height);

/* ref opencv.hpp:95-113 */
opencv_mat opencv_mat_create(int width, int height, int type);
opencv_mat opencv_mat_create_from_data(int width, int height, int type, void* data,
                                       size_t data_len);
opencv_mat opencv_mat_create_empty_from_data(int length, void* data);
bool opencv_mat_set_row_stride(opencv_mat mat, size_t stride);
void opencv_mat_release(opencv_mat mat);
void opencv_mat_resize(const opencv_mat src, opencv_mat dst, int width, int height,
                       int interpolation);
opencv_mat opencv_mat_crop(const opencv_mat src, int x, int y, int width, int height);
void opencv_mat_orientation_transform(CVImageOrientation orientation, opencv_mat mat);
int opencv_mat_get_width(const opencv_mat mat);
int opencv_mat_get_height(const opencv_mat mat);
void* opencv_mat_get_data(const opencv_mat mat);

/* ref opencv.hpp:115-117 */
opencv_encoder opencv_encoder_create(const char* ext, opencv_mat dst);
void opencv_encoder_release(opencv_encoder e);
bool opencv_encoder_write(opencv_encoder e, const opencv_mat src, const int* opt, size_t opt_len);

/* ref opencv.hpp:118-132 (host-only container parsing; no pixels) */
int opencv_decoder_get_jpeg_icc(void* src, size_t src_len, void* dest, size_t dest_len);
int opencv_decoder_get_png_icc(void* src, size_t src_len, void* dest, size_t dest_len);
int opencv_decoder_get_png_cicp(void* src, size_t src_len, uint8_t* primaries, uint8_t* transfer,
                                uint8_t* matrix, uint8_t* full_range);
size_t opencv_png_insert_cicp(void* png, size_t png_len, size_t png_cap, uint8_t primaries,
                              uint8_t transfer, uint8_t matrix, uint8_t full_range);

/* ref opencv.hpp:135-145 */
#define OPENCV_SUCCESS 0
#define OPENCV_ERROR_INVALID_CHANNEL_COUNT 1
#define OPENCV_ERROR_OUT_OF_BOUNDS 2
#define OPENCV_ERROR_NULL_MATRIX 3
#define OPENCV_ERROR_RESIZE_FAILED 4
#define OPENCV_ERROR_COPY_FAILED 5
#define OPENCV_ERROR_CONVERSION_FAILED 6
#define OPENCV_ERROR_ALPHA_BLENDING_FAILED 7
#define OPENCV_ERROR_FINAL_CONVERSION_FAILED 8
#define OPENCV_ERROR_INVALID_DIMENSIONS 9
#define OPENCV_ERROR_UNKNOWN 10

/*
 * Additive (not in the reference): make the caller-visible host bytes of a
 * device-mirrored mat current.  The reference never needs this because its
 * Mat *is* the Go buffer; lilliput's Go code touches pixel bytes directly in
 * only three places (opencv.go:224 Clear, opencv.go:802 TonemapToSDR and the
 * PSNR benchmark), each of which would call this first.  Returns 0 on success.
 */
int lp_mat_sync_host(opencv_mat mat);
/* Additive: Framebuffer.TonemapToSDR (ref opencv.go:791-810): PQ / HLG pixels of an 8-bit BGR(A) mat -> SDR BT.709
 * in place (ref color_info.cpp:112-270).  The Go side calls tonemap_rgb_8u_inplace on its own buffer; with the
 * pixels in HBM the call goes through the mat instead.  transfer / primaries are the cICP code points. */
int lp_mat_tonemap_to_sdr(opencv_mat mat, int transfer, int primaries);
/* Additive: tell the library the host bytes were modified by the caller. */
void lp_mat_mark_host_dirty(opencv_mat mat);

#ifdef __cplusplus
}
#endif
#endif
