// kernels.cuh -- internal launch interfaces between the translation units of
// liblilliput_b200.  Everything here takes DEVICE pointers and a stream.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <vector>

namespace lp {

// ---- resize.cu ---------------------------------------------------------------------------
struct ResizeArgs {
    const uint8_t* src;
    size_t src_img_stride, src_row_stride;
    int channels;
    int crop_x, crop_y, crop_w, crop_h;
    uint8_t* dst;
    size_t dst_img_stride, dst_row_stride;
    int dst_w, dst_h;
    int n;
    int interpolation;  // 1 = INTER_LINEAR, 2 = INTER_CUBIC, 3 = INTER_AREA
};
int resize_launch(const ResizeArgs& a, cudaStream_t st);

// ---- jpeg_parse.cpp (host) -----------------------------------------------------------------
struct JpegComp {
    int id, h, v, tq, td, ta;
};
struct JpegHeader {
    int width = 0, height = 0, ncomp = 0;
    JpegComp comp[3];
    int maxh = 1, maxv = 1;
    uint16_t qt[4][64];  // natural order
    bool qt_present[4] = {false, false, false, false};
    uint8_t huff_bits[2][4][17];  // [class][id][len]
    uint8_t huff_vals[2][4][256];
    bool huff_present[2][4] = {{false, false, false, false}, {false, false, false, false}};
    int restart_interval = 0;
    int orientation = 1;
    size_t scan_offset = 0;  // first entropy-coded byte
    size_t scan_length = 0;  // bytes available from scan_offset (upper bound)
    bool progressive = false;
    bool supported = false;  // baseline/extended sequential Huffman, 8-bit, 1 or 3 comps, one scan
    bool multiscan = false;  // progressive, or sequential with one scan per component: serial device path
    int mcus_x = 0, mcus_y = 0;
};
// Parses markers up to and including the first SOS.  Returns 0, or a negative lp_status.
int jpeg_parse_header(const uint8_t* data, size_t len, JpegHeader* out);

// One scan of a multi-scan file (progressive: T.81 Annex G; or non-interleaved sequential).
struct JpegScanDesc {
    uint32_t data_off, data_len;  // entropy-coded segment inside the uploaded FILE
    int32_t ns;                   // components in the scan
    int32_t ci[3], td[3], ta[3];  // frame component index, DC / AC table ids
    int32_t Ss, Se, Ah, Al;       // spectral band and successive-approximation bit positions
    int32_t restart_interval;
    int32_t table_set;            // Huffman tables in force at this scan
    int32_t progressive;
};

// ---- jpeg_decode.cu ------------------------------------------------------------------------
// Device-side description of one image to decode (array of these lives in HBM).
struct JpegDecodeItem {
    uint64_t scan_off;    // offset of the entropy-coded segment in the batch scan buffer
    uint32_t scan_len;    // bytes
    uint32_t table_set;   // index into the Huffman table-set array
    uint64_t coef_off;    // int16 offset of this image's coefficient blocks
    uint64_t plane_off;   // byte offset of this image's component planes
    uint64_t frame_off;   // byte offset of this image's packed output frame
    int32_t width, height, ncomp;
    int32_t mcus_x, mcus_y, restart_interval;
    int32_t h[3], v[3];   // sampling factors
    int32_t bw[3], bh[3]; // blocks per component plane (padded to the MCU grid)
    int32_t dw[3], dh[3]; // true downsampled component size in samples
    uint32_t block_off[3];  // first block of component c inside the image's coef area
    uint32_t plane_rel[3];  // byte offset of component c's plane inside the image's plane area
    uint16_t qt[3][64];     // per-component quantisation table, natural order
    int32_t td[3], ta[3];
    int32_t status;         // written by the decode kernel: 0 ok, <0 corrupt
    int32_t frame_channels; // 1 or 3
    // parallel Huffman path (jpeg_huff_parallel.cu)
    uint64_t clean_off;     // byte offset of this image's unstuffed bit string
    uint64_t state_off;     // SubState offset (2 * nsub entries reserved)
    uint64_t dcdiff_off;    // int16 offset of this image's DC-difference array (MCU order)
    uint32_t clean_len;     // written by jpeg_unstuff_kernel
    uint32_t pad_;
    // Region of interest.  Only MCUs [roi_mx0, roi_mx0+roi_mcx) x [roi_my0, roi_my0+roi_mcy) get
    // coefficients / planes (bw, bh, block_off, plane_rel describe THAT grid); the packed frame holds
    // the pixel window [win_x0, win_x0+win_w) x [win_y0, win_y0+win_h) with rows win_stride apart.
    // A full decode has roi = every MCU and win = the whole image.
    int32_t roi_mx0, roi_my0, roi_mcx, roi_mcy;
    int32_t win_x0, win_y0, win_w, win_h;
    uint32_t win_stride;
    uint32_t pad2_;
};

// Fills the ROI / window / layout fields of `it` (whose width, height, ncomp, h, v, mcus_* are set)
// for the pixel window [x0,x1) x [y0,y1).  align16 rounds the window's x range outwards to 16 px so
// the vectorised colour kernel can use 16-byte stores.  Returns blocks in the ROI.
uint32_t jpeg_item_set_window(JpegDecodeItem* it, int x0, int y0, int x1, int y1, bool align16,
                              uint32_t* plane_bytes);

// Huffman decode tables for one image (or many images sharing them), device format.
constexpr int kHuffAcLookBits = 12;    // AC lookahead of the parallel decoder (jpeg_huff_parallel.cu)
constexpr int kHuffLongPrefixes = 16;  // second-level tables per AC table for codes longer than that
struct JpegHuffSet {
    // [class*4+id]: 9-bit lookahead: (len<<8)|symbol, 0 when the code is longer than 9 bits
    uint16_t look[8][512];
    int32_t maxcode[8][18];  // canonical decode for long codes; maxcode[17] = sentinel
    int32_t valoffset[8][17];
    uint8_t vals[8][256];
    // AC codes longer than kHuffAcLookBits, two-level: long_prefix[id][j] = their first kHuffAcLookBits
    // bits (0xFFFF = unused slot), long_sub[id][j][next 4 bits] = (len<<8)|symbol, 0 = not a codeword.
    // Canonical codes put every long code behind a handful of all-ones prefixes (8 for the Annex K
    // tables); prefixes that do not fit here are left to the bit-by-bit walk.
    uint16_t long_prefix[4][kHuffLongPrefixes];
    uint16_t long_sub[4][kHuffLongPrefixes][16];
};
void jpeg_build_huff_set(const JpegHeader& h, JpegHuffSet* out);
// Walks every SOS of a multi-scan file, snapshotting the Huffman tables / restart interval in force.
// `sets`/`scans` are caller arrays (max_sets / max_scans entries).  Returns 0 or a negative lp_status.
int jpeg_parse_scans(const uint8_t* data, size_t len, const JpegHeader& h, JpegScanDesc* scans, int max_scans,
                     int* nscans, JpegHuffSet* sets, int max_sets, int* nsets);

struct JpegDecodeBatch {
    JpegDecodeItem* items;      // device
    const JpegHuffSet* tables;  // device
    const uint8_t* scan;        // device, concatenated entropy-coded segments
    int16_t* coef;              // device, zeroed by the launcher
    uint8_t* planes;            // device
    uint8_t* frames;            // device, packed BGR / gray frames
    int n;
    size_t coef_elems_total;    // for the memset
    int max_blocks_per_image;
    int max_width, max_height;
    // parallel Huffman scratch (all device); used when use_parallel_huffman is set
    bool use_parallel_huffman = false;
    uint8_t* clean = nullptr;
    void* states = nullptr;      // SubState[]
    uint32_t* nslots = nullptr;
    int16_t* dcdiff = nullptr;   // DC differences of ALL blocks of an image, MCU order
    // multi-scan files (n == 1): every scan decoded in order by one thread; `scan` holds the whole file
    const JpegScanDesc* scans = nullptr;
    int nscans = 0;
    // images with restart markers inside a parallel-Huffman launch: (image, interval) work list (device), one
    // thread per restart interval; their marker offsets live in `nslots` at the image's state_off, the number of
    // intervals the host expects in clean_len
    const uint2* rst_work = nullptr;
    int n_rst_work = 0;
};
// Scratch sizing for the parallel Huffman path, per image with `scan_len` entropy-coded bytes.
inline size_t huff_clean_bytes(size_t scan_len) { return ((scan_len + 48 + 15) / 16) * 16; }
inline size_t huff_nsub(size_t scan_len) { return scan_len * 8 / 1024 + 2; }
struct JpegHuffParallelArgs {
    JpegDecodeItem* items;
    const JpegHuffSet* tables;
    const uint8_t* scan;
    uint8_t* clean;
    void* states;
    uint32_t* nslots;
    int16_t* coef;
    int16_t* dcdiff;
    int n;
};
int jpeg_huff_parallel_launch(const JpegHuffParallelArgs& a, cudaStream_t st);
// Images the sync kernel can keep resident at once (SMs x CTAs per SM); chunk sizes that are a
// multiple of this avoid a mostly-empty last wave.
int jpeg_huff_parallel_slots();
// Diagnostics: clock64 cycles per phase of the sync kernels summed over CTAs (see jpeg_huff_parallel.cu)
int jpeg_huff_phase_clocks(unsigned long long out[8], int reset);
// Launches: memset(coef) -> huffman decode -> idct -> upsample+colour.
int jpeg_decode_launch(const JpegDecodeBatch& b, cudaStream_t st, cudaEvent_t ev_after_huff);

// ---- png_parse.cpp (host) / png_decode.cu ------------------------------------------------------
struct PngSegment {
    size_t offset, length;  // an IDAT payload inside the file
};
struct PngHeader {
    int width = 0, height = 0, bit_depth = 0, color_type = 0, interlace = 0;
    int src_channels = 0, out_channels = 0, bpp = 1;
    size_t row_bytes = 0;  // filtered scanline without the filter-type byte
    int npal = 0, ntrns = 0;
    bool has_trns = false;
    uint8_t palette[256 * 3] = {0};
    uint8_t trns[256] = {0};
    uint16_t trns_rgb[3] = {0, 0, 0};
    std::vector<PngSegment> idat;
    size_t idat_total = 0;
    int orientation = 1;  // from an eXIf chunk in front of the first IDAT (OpenCV reads it like a JPEG's EXIF block)
};
// EXIF orientation of a TIFF block, read the way the reference's OpenCV reads it (jpeg_parse.cpp)
bool exif_orientation_opencv(const uint8_t* tiff, size_t n, int* value);
int png_parse(const uint8_t* data, size_t len, PngHeader* out);
int png_extract_icc(const uint8_t* data, size_t len, uint8_t* dest, size_t dest_len);
// cICP code points (primaries, transfer, matrix, full range) of the chunk libpng would report; 1 if found
int png_extract_cicp(const uint8_t* data, size_t len, uint8_t* out4);

// One PNG to decode (array in HBM).  zoff: the concatenated IDAT payload (one zlib stream);
// raw: (row_bytes+1)*height bytes of filtered scanlines, defiltered in place; frame: packed output.
struct PngDecodeItem {
    uint64_t z_off, raw_off, frame_off;
    uint32_t z_len;
    int32_t width, height, bit_depth, color_type, src_channels, out_channels, bpp;
    uint32_t row_bytes, frame_stride;
    int32_t npal, ntrns, has_trns;
    uint16_t trns_rgb[3];
    uint16_t pad_;
    uint8_t palette[256 * 3];
    uint8_t trns[256];
    int32_t status;  // 0 ok, <0 corrupt stream
    uint32_t produced;
    // Adam7 (PNG spec s.8.2): the inflated stream is up to seven reduced images back to back, each
    // with its own scanline length.  A non-interlaced image is one "pass" covering everything.
    int32_t interlace, npass;
    uint32_t raw_total;               // inflated bytes expected
    uint32_t pass_off[7], pass_rb[7];  // byte offset / scanline bytes (without the filter byte)
    int32_t pass_w[7], pass_h[7];
};
// Fills npass / raw_total / pass_* from width, height, bit depth, channels, interlace.
void png_item_set_passes(PngDecodeItem* it);
struct PngDecodeBatch {
    PngDecodeItem* items;  // device
    const uint8_t* z;      // device: zlib streams
    uint8_t* raw;          // device
    uint8_t* frames;       // device
    int n;
    int max_width, max_height;
};
// inflate -> defilter -> convert to packed Gray / BGR / BGRA u8.
int png_decode_launch(const PngDecodeBatch& b, cudaStream_t st);
// the two halves, for callers that keep every stream's scanlines but only a window of frames (xbatch.cu)
int png_inflate_launch(const PngDecodeBatch& b, cudaStream_t st);
int png_unfilter_launch(const PngDecodeBatch& b, int first, int count, cudaStream_t st);

// ---- png_encode.cu ---------------------------------------------------------------------------
// One packed device frame -> a complete PNG file in host memory (filter + deflate on the device).
int png_encode_frame(const uint8_t* frame, size_t row_stride, int width, int height, int channels, int level,
                     bool adaptive_filters, std::vector<uint8_t>* out, cudaStream_t st);

// ---- jpeg_encode.cu ------------------------------------------------------------------------
struct JpegEncodeBatch {
    const uint8_t* frames;  // device packed frames
    size_t frame_img_stride, frame_row_stride;
    int width, height, channels;  // shared by the batch
    int quality;
    int n;
    uint8_t* out;            // device, n * out_cap
    size_t out_cap;
    uint32_t* out_len;       // device, n (0 on overflow)
    // scratch (device), sized by jpeg_encode_scratch_bytes
    void* scratch;
};
size_t jpeg_encode_scratch_bytes(int width, int height, int channels, int n, size_t out_cap);
int jpeg_encode_launch(const JpegEncodeBatch& b, cudaStream_t st, cudaEvent_t ev_after_transform);

// ---- webp_encode.cu -------------------------------------------------------------------------
struct WebpEncodedFrame {
    std::vector<uint8_t> image;  // "VP8 " or "VP8L" payload
    std::vector<uint8_t> alph;   // "ALPH" payload (lossy frames with transparency)
    bool lossless = false, has_alpha = false;
    int width = 0, height = 0, duration = 0;
};
// n packed device frames of one geometry (frame i at d_frames + i*img_stride) -> lossy VP8 payloads (+ ALPH for
// frames with transparency), ref webp.cpp:711-729 / 650-700 per frame.
int webp_encode_lossy_batch(const uint8_t* d_frames, size_t img_stride, size_t row_step, int width, int height,
                            int channels, int n, int quality, std::vector<WebpEncodedFrame>* out, cudaStream_t st);
void webp_assemble(const WebpEncodedFrame* frames, int n, const uint8_t* icc, size_t icc_len, uint32_t bgcolor,
                   uint32_t loop_count, std::vector<uint8_t>* file);

// ---- batch helpers of webp_decode.cu / gif_decode.cu (used by xbatch.cu) ------------------------
struct WebpStillInfo {
    int width = 0, height = 0;
    size_t vp8_off = 0, vp8_len = 0;  // "VP8 " payload inside the file
    bool simple_lossy = false;        // one VP8 key frame, no ALPH / ICCP / animation
};
bool webp_still_info(const uint8_t* data, size_t len, WebpStillInfo* out);
int webp_vp8_decode_batch(const uint8_t* d_in, const uint64_t* in_off, const uint32_t* in_len, int n, const int* width,
                          const int* height, uint8_t* d_frames, const uint64_t* frame_off, int* h_status, cudaStream_t st);
struct GifAnimPlan;
GifAnimPlan* gif_plan_parse(const uint8_t* data, size_t len, int max_frames);
void gif_plan_free(GifAnimPlan* p);
void gif_plan_info(const GifAnimPlan* p, int* width, int* height, int* nframes, uint32_t* bgcolor, int* loop_count);
int gif_plan_delay_ms(const GifAnimPlan* p, int frame);
size_t gif_plan_device_bytes(const GifAnimPlan* p);
// decodes + composites every frame of `n` animations of one canvas size: animation a, frame f lands at
// d_canvases + (first_frame[a] + f) * canvas_stride (BGRA); h_status per animation
int gif_decode_batch(GifAnimPlan* const* plans, const uint8_t* const* files, const size_t* file_len, int n,
                     uint8_t* d_scratch, size_t scratch_bytes, uint8_t* d_canvases, size_t canvas_stride,
                     const int* first_frame, int* h_status, cudaStream_t st);

// ---- pixel_ops.cu --------------------------------------------------------------------------
int orient_launch(const uint8_t* src, int w, int h, int channels, int orientation, uint8_t* dst,
                  cudaStream_t st);
int copy_region_launch(const uint8_t* src, size_t src_step, int src_ch, uint8_t* dst,
                       size_t dst_step, int dst_ch, int w, int h, cudaStream_t st);
// ---- tonemap.cu ----------------------------------------------------------------------------
int tonemap_to_sdr_launch(uint8_t* d_px, size_t step, int channels, int w, int h, int transfer, int primaries, cudaStream_t st);

int compact_launch(const uint8_t* src, size_t stride, const uint32_t* len, uint32_t cap, int n, uint8_t* dst,
                   unsigned long long* off /* n + 1 */, cudaStream_t st);
struct SegCopy {
    uint64_t src, dst;  // byte offsets from one base pointer
    uint32_t len, pad_;
};
int seg_copy_launch(const SegCopy* d_segs, int n, uint8_t* base, cudaStream_t st);
int blend_region_launch(const uint8_t* src, size_t src_step, int src_ch, uint8_t* dst,
                        size_t dst_step, int dst_ch, int w, int h, cudaStream_t st);

}  // namespace lp
