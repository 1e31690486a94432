// jpeg_decode.cu -- baseline JPEG decode on sm_100a: Huffman scan -> dequant + 8x8 IDCT ->
// fancy chroma upsampling + YCbCr->BGR into a packed device frame.
//
// Replaces: opencv_decoder_read_data (ref opencv.cpp:166-171), i.e. what
// cv::ImageDecoder::readData asks of libjpeg-turbo 3.1.0 with its defaults (ISLOW IDCT,
// fancy upsampling, JCS_EXT_BGR).  Arithmetic contract: SURVEY.md Appendix E.2; results are
// bit-identical to the reference on baseline streams (tests/test_jpeg_decode_gpu.py).
//
// Kernels (each one grid launch over the whole batch):
//   jpeg_huff_decode_kernel   entropy-coded segment -> quantised coefficients (int16, natural
//                             order, zero-initialised buffer).  Bit-serial by nature; this
//                             first version runs one stream per thread.
//   jpeg_idct_kernel          one thread per 8x8 block: dequantise, two 1-D passes in
//                             registers, 8-byte row stores into the component plane.
//   jpeg_upsample_color_kernel  triangle upsampling + fixed-point colour conversion, packed BGR.
#include "common.cuh"
#include "kernels.cuh"

namespace lp {

__constant__ uint8_t c_zigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,
                                     12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
                                     35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
                                     58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// ------------------------------------------------------------------ ROI layout (host)

uint32_t jpeg_item_set_window(JpegDecodeItem* it, int x0, int y0, int x1, int y1, bool align16,
                              uint32_t* plane_bytes_out) {
    int maxh = 1, maxv = 1;
    for (int c = 0; c < it->ncomp; c++) {
        maxh = it->h[c] > maxh ? it->h[c] : maxh;
        maxv = it->v[c] > maxv ? it->v[c] : maxv;
    }
    x0 = x0 < 0 ? 0 : x0;
    y0 = y0 < 0 ? 0 : y0;
    x1 = x1 > it->width ? it->width : x1;
    y1 = y1 > it->height ? it->height : y1;
    if (align16) {
        x0 &= ~15;
        x1 = (x1 + 15) & ~15;
        if (x1 > it->width) x1 = it->width;
    }
    it->win_x0 = x0;
    it->win_y0 = y0;
    it->win_w = x1 - x0;
    it->win_h = y1 - y0;
    const int fc = it->ncomp == 1 ? 1 : 3;
    it->win_stride = (uint32_t)(((size_t)it->win_w * fc + 15) / 16 * 16);
    if (x0 == 0 && x1 == it->width) it->win_stride = (uint32_t)it->win_w * fc;  // whole rows: packed
    // fancy upsampling reads one chroma sample beyond the window on every side = 2 luma pixels at 2x
    const int mw = 8 * maxh, mh = 8 * maxv;
    int px0 = x0 - 2 * maxh, px1 = x1 - 1 + 2 * maxh, py0 = y0 - 2 * maxv, py1 = y1 - 1 + 2 * maxv;
    px0 = px0 < 0 ? 0 : px0;
    py0 = py0 < 0 ? 0 : py0;
    px1 = px1 > it->width - 1 ? it->width - 1 : px1;
    py1 = py1 > it->height - 1 ? it->height - 1 : py1;
    it->roi_mx0 = px0 / mw;
    it->roi_my0 = py0 / mh;
    it->roi_mcx = px1 / mw - it->roi_mx0 + 1;
    it->roi_mcy = py1 / mh - it->roi_my0 + 1;
    uint32_t blocks = 0, plane_bytes = 0;
    for (int c = 0; c < it->ncomp; c++) {
        it->bw[c] = it->roi_mcx * it->h[c];
        it->bh[c] = it->roi_mcy * it->v[c];
        it->block_off[c] = blocks;
        it->plane_rel[c] = plane_bytes;
        blocks += (uint32_t)it->bw[c] * it->bh[c];
        plane_bytes += (uint32_t)it->bw[c] * it->bh[c] * 64;
    }
    if (plane_bytes_out) *plane_bytes_out = plane_bytes;
    return blocks;
}

// Coefficient block of component c at full-image block position (X, Y); nullptr outside the ROI.
__device__ __forceinline__ int16_t* roi_block(const JpegDecodeItem& it, int16_t* coef, int c, int X, int Y) {
    const int rx = X - it.roi_mx0 * it.h[c], ry = Y - it.roi_my0 * it.v[c];
    if (rx < 0 || ry < 0 || rx >= it.bw[c] || ry >= it.bh[c]) return nullptr;
    return coef + it.coef_off + ((size_t)it.block_off[c] + (size_t)ry * it.bw[c] + rx) * 64;
}

// ------------------------------------------------------------------ entropy decode (serial)

struct BitReader {
    const uint8_t* p;
    const uint8_t* end;
    uint64_t acc;
    int nbits;
    bool marker;
};

__device__ __forceinline__ void br_fill(BitReader& b) {
    while (b.nbits <= 56) {
        uint32_t byte = 0;
        if (!b.marker && b.p < b.end) {
            byte = *b.p;
            if (byte == 0xFF) {
                const uint8_t* q = b.p + 1;
                while (q < b.end && *q == 0xFF) q++;
                if (q < b.end && *q == 0x00) {
                    b.p = q + 1;  // stuffed FF
                } else {
                    b.marker = true;  // real marker: feed zeros from here on
                    byte = 0;
                }
            } else {
                b.p++;
            }
        }
        b.acc |= (uint64_t)byte << (56 - b.nbits);
        b.nbits += 8;
    }
}

__device__ __forceinline__ int huff_symbol(BitReader& b, const JpegHuffSet* hs, int t) {
    if (b.nbits < 32) br_fill(b);
    const uint32_t peek = (uint32_t)(b.acc >> 48);
    const uint32_t e = hs->look[t][peek >> 7];
    if (e) {
        const int l = e >> 8;
        b.acc <<= l;
        b.nbits -= l;
        return e & 0xFF;
    }
    int l = 10;
    int code = (int)(peek >> 6);
    while (l <= 16 && code > hs->maxcode[t][l]) {
        l++;
        code = (int)(peek >> (16 - l));
    }
    if (l > 16) return -1;
    b.acc <<= l;
    b.nbits -= l;
    return hs->vals[t][(code + hs->valoffset[t][l]) & 0xFF];
}

__device__ __forceinline__ int receive_extend(BitReader& b, int n) {
    if (b.nbits < 32) br_fill(b);
    const int v = (int)(b.acc >> (64 - n));
    b.acc <<= n;
    b.nbits -= n;
    return v < (1 << (n - 1)) ? v - (1 << n) + 1 : v;
}

__global__ void jpeg_huff_decode_kernel(JpegDecodeItem* items, const JpegHuffSet* tables,
                                        const uint8_t* scan, int16_t* coef, int n) {
    __shared__ uint8_t zz[64];
    for (int k = threadIdx.x; k < 64; k += blockDim.x) zz[k] = c_zigzag[k];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    JpegDecodeItem& it = items[i];
    const JpegHuffSet* hs = tables + it.table_set;
    BitReader b{scan + it.scan_off, scan + it.scan_off + it.scan_len, 0, 0, false};
    int pred[3] = {0, 0, 0};
    int todo = it.restart_interval;
    int status = 0;
    const int nc = it.ncomp;
    for (int my = 0; my < it.mcus_y && status == 0; my++) {
        for (int mx = 0; mx < it.mcus_x && status == 0; mx++) {
            if (it.restart_interval && todo == 0) {
                // byte-align, skip to just past the next RSTn, reset predictors
                b.acc = 0;
                b.nbits = 0;
                const uint8_t* q = b.p;
                while (q + 1 < b.end && !(q[0] == 0xFF && q[1] >= 0xD0 && q[1] <= 0xD7)) q++;
                if (q + 1 >= b.end) {
                    status = -3;
                    break;
                }
                b.p = q + 2;
                b.marker = false;
                pred[0] = pred[1] = pred[2] = 0;
                todo = it.restart_interval;
            }
            for (int c = 0; c < nc && status == 0; c++) {
                const int td = it.td[c], ta = 4 + it.ta[c];
                for (int by = 0; by < it.v[c] && status == 0; by++) {
                    for (int bx = 0; bx < it.h[c]; bx++) {
                        const int X = mx * it.h[c] + bx, Y = my * it.v[c] + by;
                        int16_t* blk = roi_block(it, coef, c, X, Y);  // nullptr: decode but do not store
                        int s = huff_symbol(b, hs, td);
                        if (s < 0 || s > 15) { status = -3; break; }
                        if (s) pred[c] += receive_extend(b, s);
                        if (blk) blk[0] = (int16_t)pred[c];
                        for (int k = 1; k < 64;) {
                            const int rs = huff_symbol(b, hs, ta);
                            if (rs < 0) { status = -3; break; }
                            const int r = rs >> 4, sz = rs & 15;
                            if (sz == 0) {
                                if (r != 15) break;
                                k += 16;
                                continue;
                            }
                            k += r;
                            if (k > 63) { status = -3; break; }
                            const int val = receive_extend(b, sz);
                            if (blk) blk[zz[k]] = (int16_t)val;
                            k++;
                        }
                        if (status) break;
                    }
                }
            }
            if (it.restart_interval) todo--;
        }
    }
    it.status = status;
}

// ------------------------------------------------------------------ restart-interval-parallel decode
// A scan with restart markers (DRI) is a sequence of independently decodable intervals: every RSTn marker is
// byte aligned, the DC predictors restart at 0 behind it (T.81 E.1.4, F.2.2.x; libjpeg-turbo jdhuff.c
// process_restart).  So: one pass finds the markers of every image (jpeg_rst_scan_kernel), then ONE THREAD PER
// INTERVAL decodes its MCUs (jpeg_rst_decode_kernel) -- with DRI = one MCU row a 1080p batch of 4096 images is
// 278 528 independent streams.  Coefficients go to the same scan-order layout the self-synchronising decoder of
// non-DRI streams writes ([roi MCU][block in MCU], DC values final), so both kinds share one IDCT launch.

// positions (byte offsets behind the marker) of the RSTn markers of one image, in order; rst_off[0] = 0
__global__ void __launch_bounds__(256) jpeg_rst_scan_kernel(JpegDecodeItem* items, const uint8_t* scan, uint32_t* rst_all) {
    __shared__ uint32_t warp_sums[8];
    __shared__ uint32_t s_carry;
    JpegDecodeItem& it = items[blockIdx.x];
    if (it.status != 0 || it.restart_interval == 0) return;
    const uint8_t* s = scan + it.scan_off;
    const uint32_t len = it.scan_len;
    uint32_t* out = rst_all + it.state_off;  // (state_off doubles as the interval table offset of DRI images)
    const uint32_t cap = it.clean_len;       // intervals expected (set by the host); out has cap + 1 slots
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) {
        s_carry = 1;
        out[0] = 0;
    }
    __syncthreads();
    for (uint32_t base = 0; base < len; base += 256 * 16) {
        // 16 consecutive bytes per thread; marker = FF followed by D0..D7
        const uint32_t b0 = base + (uint32_t)tid * 16;
        uint32_t found = 0;
        for (uint32_t k = 0; k < 16; k++) {
            const uint32_t i = b0 + k;
            if (i + 1 < len && s[i] == 0xFF && (s[i + 1] & 0xF8) == 0xD0) found |= 1u << k;
        }
        const uint32_t cnt = __popc(found);
        uint32_t inc = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) warp_sums[wid] = inc;
        __syncthreads();
        uint32_t before = s_carry;
        for (int w = 0; w < wid; w++) before += warp_sums[w];
        uint32_t at = before + inc - cnt;
        for (uint32_t k = 0; k < 16; k++)
            if (found & (1u << k)) {
                if (at <= cap) out[at] = b0 + k + 2;
                at++;
            }
        __syncthreads();
        if (tid == 255) s_carry = before + inc;
        __syncthreads();
    }
    if (tid == 0) it.pad_ = s_carry;  // intervals found (markers + 1)
}

__global__ void __launch_bounds__(128)
    jpeg_rst_decode_kernel(JpegDecodeItem* items, const JpegHuffSet* tables, const uint8_t* scan, const uint32_t* rst_all,
                           const uint2* work /* (image, interval) */, int nwork, int16_t* coef) {
    __shared__ uint8_t zz[64];
    for (int k = threadIdx.x; k < 64; k += blockDim.x) zz[k] = c_zigzag[k];
    __syncthreads();
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nwork) return;
    JpegDecodeItem& it = items[work[w].x];
    if (it.status != 0) return;
    const uint32_t k = work[w].y, nint = it.clean_len;
    if (it.pad_ != nint) {  // fewer or more markers than intervals: damaged stream (the per-image path resynchronises as libjpeg does)
        if (k == 0) it.status = -3;
        return;
    }
    const uint32_t* off = rst_all + it.state_off;
    const JpegHuffSet* hs = tables + it.table_set;
    const uint8_t* base = scan + it.scan_off;
    const uint32_t begin = off[k], end = k + 1 < nint ? off[k + 1] - 2 : it.scan_len;
    if (k > 0 && (base[begin - 1] & 7u) != ((k - 1) & 7u)) {  // RSTn must count modulo 8
        it.status = -3;
        return;
    }
    BitReader b{base + begin, base + end, 0, 0, false};
    int nb = 0;
    for (int c = 0; c < it.ncomp; c++) nb += it.h[c] * it.v[c];
    const uint32_t total_mcus = (uint32_t)it.mcus_x * it.mcus_y;
    uint32_t mcu = k * (uint32_t)it.restart_interval;
    const uint32_t mcu_end = min(total_mcus, mcu + (uint32_t)it.restart_interval);
    int pred[3] = {0, 0, 0};
    int status = 0;
    int16_t* coef_base = coef + it.coef_off;
    for (; mcu < mcu_end && !status; mcu++) {
        const int mx = (int)(mcu % (uint32_t)it.mcus_x), my = (int)(mcu / (uint32_t)it.mcus_x);
        const int rx = mx - it.roi_mx0, ry = my - it.roi_my0;
        const bool inside = (unsigned)rx < (unsigned)it.roi_mcx && (unsigned)ry < (unsigned)it.roi_mcy;
        int16_t* blk = coef_base + ((size_t)ry * it.roi_mcx + rx) * ((size_t)nb * 64);
        for (int c = 0; c < it.ncomp && !status; c++) {
            const int td = it.td[c], ta = 4 + it.ta[c];
            for (int j = 0; j < it.h[c] * it.v[c]; j++, blk += 64) {
                int s = huff_symbol(b, hs, td);
                if (s < 0 || s > 15) { status = -3; break; }
                if (s) pred[c] += receive_extend(b, s);
                if (inside) blk[0] = (int16_t)pred[c];
                for (int q = 1; q < 64;) {
                    const int rs = huff_symbol(b, hs, ta);
                    if (rs < 0) { status = -3; break; }
                    const int r = rs >> 4, sz = rs & 15;
                    if (sz == 0) {
                        if (r != 15) break;
                        q += 16;
                        continue;
                    }
                    q += r;
                    if (q > 63) { status = -3; break; }
                    const int val = receive_extend(b, sz);
                    if (inside) blk[zz[q]] = (int16_t)val;
                    q++;
                }
                if (status) break;
            }
        }
    }
    if (status) it.status = status;
}

int jpeg_rst_launch(JpegDecodeItem* items, const JpegHuffSet* tables, const uint8_t* scan, uint32_t* rst_all, const uint2* d_work,
                    int nwork, int n_images, int16_t* coef, cudaStream_t st) {
    if (nwork <= 0) return LP_OK;
    jpeg_rst_scan_kernel<<<n_images, 256, 0, st>>>(items, scan, rst_all);
    jpeg_rst_decode_kernel<<<ceil_div(nwork, 128), 128, 0, st>>>(items, tables, scan, rst_all, d_work, nwork, coef);
    g_launches += 2;
    LP_CUDA_OK(cudaGetLastError());
    return LP_OK;
}

// ------------------------------------------------------------------ multi-scan files (serial)
// Progressive JPEG (T.81 Annex G; libjpeg-turbo's jdphuff.c is what the reference runs) and
// sequential files with one scan per component.  Every scan refines the same coefficient array,
// and inside a scan the end-of-band runs chain across blocks, so one thread walks the scans in
// order.  Restated in oracle/oracle_jpeg_dec.c (prog_*), which is pinned on the reference.

__device__ __forceinline__ int br_bits(BitReader& b, int n) {
    if (n == 0) return 0;
    if (b.nbits < 32) br_fill(b);
    const int v = (int)(b.acc >> (64 - n));
    b.acc <<= n;
    b.nbits -= n;
    return v;
}

struct ProgState {
    int Ss, Se, Al;
    unsigned eobrun;
};

__device__ int prog_ac_first(BitReader& b, const JpegHuffSet* hs, int ta, ProgState& ps, int16_t* blk, const uint8_t* zz) {
    if (ps.eobrun > 0) {
        ps.eobrun--;
        return 0;
    }
    for (int k = ps.Ss; k <= ps.Se; k++) {
        const int rs = huff_symbol(b, hs, ta);
        if (rs < 0) return -3;
        int r = rs >> 4;
        const int n = rs & 15;
        if (n) {
            k += r;
            if (k > 63) return -3;
            blk[zz[k]] = (int16_t)((unsigned)receive_extend(b, n) << ps.Al);
        } else if (r == 15) {
            k += 15;
        } else {
            ps.eobrun = 1u << r;
            if (r) ps.eobrun += (unsigned)br_bits(b, r);
            ps.eobrun--;
            break;
        }
    }
    return 0;
}

__device__ int prog_ac_refine(BitReader& b, const JpegHuffSet* hs, int ta, ProgState& ps, int16_t* blk, const uint8_t* zz) {
    const int p1 = 1 << ps.Al, m1 = -(1 << ps.Al);
    int k = ps.Ss;
    if (ps.eobrun == 0) {
        for (; k <= ps.Se; k++) {
            const int rs = huff_symbol(b, hs, ta);
            if (rs < 0) return -3;
            int r = rs >> 4;
            const int n = rs & 15;
            int val = 0;
            if (n) {
                if (n != 1) return -3;
                val = br_bits(b, 1) ? p1 : m1;
            } else if (r != 15) {
                ps.eobrun = 1u << r;
                if (r) ps.eobrun += (unsigned)br_bits(b, r);
                break;
            }
            do {
                int16_t* co = blk + zz[k];
                if (*co != 0) {
                    if (br_bits(b, 1)) {
                        if ((*co & p1) == 0) *co = (int16_t)(*co + (*co >= 0 ? p1 : m1));
                    }
                } else {
                    if (--r < 0) break;
                }
                k++;
            } while (k <= ps.Se);
            if (val) {
                if (k > 63) return -3;
                blk[zz[k]] = (int16_t)val;
            }
        }
    }
    if (ps.eobrun > 0) {
        for (; k <= ps.Se; k++) {
            int16_t* co = blk + zz[k];
            if (*co != 0 && br_bits(b, 1)) {
                if ((*co & p1) == 0) *co = (int16_t)(*co + (*co >= 0 ? p1 : m1));
            }
        }
        ps.eobrun--;
    }
    return 0;
}

__global__ void jpeg_multiscan_kernel(JpegDecodeItem* item, const JpegScanDesc* scans, int nscans,
                                      const JpegHuffSet* sets, const uint8_t* file, int16_t* coef) {
    __shared__ uint8_t zz[64];
    for (int k = threadIdx.x; k < 64; k += blockDim.x) zz[k] = c_zigzag[k];
    __syncthreads();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    JpegDecodeItem& it = *item;
    int status = 0;
    for (int s = 0; s < nscans && status == 0; s++) {
        const JpegScanDesc sc = scans[s];
        const JpegHuffSet* hs = sets + sc.table_set;
        BitReader b{file + sc.data_off, file + sc.data_off + sc.data_len, 0, 0, false};
        ProgState ps{sc.Ss, sc.Se, sc.Al, 0u};
        int pred[3] = {0, 0, 0};
        int mcux, mcuy;
        if (sc.ns == 1) {  // non-interleaved: one block per MCU over the component's true block grid
            mcux = (it.dw[sc.ci[0]] + 7) / 8;
            mcuy = (it.dh[sc.ci[0]] + 7) / 8;
        } else {
            mcux = it.mcus_x;
            mcuy = it.mcus_y;
        }
        int todo = sc.restart_interval;
        for (int my = 0; my < mcuy && status == 0; my++) {
            for (int mx = 0; mx < mcux && status == 0; mx++) {
                if (sc.restart_interval && todo == 0) {
                    b.acc = 0;
                    b.nbits = 0;
                    const uint8_t* q = b.p;
                    while (q + 1 < b.end && !(q[0] == 0xFF && q[1] >= 0xD0 && q[1] <= 0xD7)) q++;
                    if (q + 1 >= b.end) {
                        status = -3;
                        break;
                    }
                    b.p = q + 2;
                    b.marker = false;
                    pred[0] = pred[1] = pred[2] = 0;
                    ps.eobrun = 0;
                    todo = sc.restart_interval;
                }
                for (int i = 0; i < sc.ns && status == 0; i++) {
                    const int c = sc.ci[i];
                    const int bh = sc.ns == 1 ? 1 : it.h[c], bv = sc.ns == 1 ? 1 : it.v[c];
                    const int td = sc.td[i], ta = 4 + sc.ta[i];
                    for (int by = 0; by < bv && status == 0; by++) {
                        for (int bx = 0; bx < bh && status == 0; bx++) {
                            int16_t* blk = roi_block(it, coef, c, mx * bh + bx, my * bv + by);
                            if (!blk) {
                                status = -3;
                                break;
                            }
                            if (!sc.progressive) {  // sequential block: DC difference + AC run/size pairs
                                const int sz = huff_symbol(b, hs, td);
                                if (sz < 0 || sz > 15) { status = -3; break; }
                                if (sz) pred[i] += receive_extend(b, sz);
                                blk[0] = (int16_t)pred[i];
                                for (int k = 1; k < 64;) {
                                    const int rs = huff_symbol(b, hs, ta);
                                    if (rs < 0) { status = -3; break; }
                                    const int r = rs >> 4, n = rs & 15;
                                    if (n == 0) {
                                        if (r != 15) break;
                                        k += 16;
                                        continue;
                                    }
                                    k += r;
                                    if (k > 63) { status = -3; break; }
                                    blk[zz[k]] = (int16_t)receive_extend(b, n);
                                    k++;
                                }
                            } else if (sc.Ss == 0) {
                                if (sc.Ah == 0) {  // DC first pass
                                    const int sz = huff_symbol(b, hs, td);
                                    if (sz < 0 || sz > 15) { status = -3; break; }
                                    if (sz) pred[i] += receive_extend(b, sz);
                                    blk[0] = (int16_t)((unsigned)pred[i] << sc.Al);
                                } else if (br_bits(b, 1)) {  // DC refinement
                                    blk[0] |= (int16_t)(1 << sc.Al);
                                }
                            } else {
                                status = sc.Ah == 0 ? prog_ac_first(b, hs, ta, ps, blk, zz)
                                                    : prog_ac_refine(b, hs, ta, ps, blk, zz);
                            }
                        }
                    }
                }
                if (sc.restart_interval) todo--;
            }
        }
    }
    it.status = status;
}

// ------------------------------------------------------------------ dequant + ISLOW IDCT

#define LP_FIX_0_298631336 2446
#define LP_FIX_0_390180644 3196
#define LP_FIX_0_541196100 4433
#define LP_FIX_0_765366865 6270
#define LP_FIX_0_899976223 7373
#define LP_FIX_1_175875602 9633
#define LP_FIX_1_501321110 12299
#define LP_FIX_1_847759065 15137
#define LP_FIX_1_961570560 16069
#define LP_FIX_2_053119869 16819
#define LP_FIX_2_562915447 20995
#define LP_FIX_3_072711026 25172

__device__ __forceinline__ int sat_s16(int v) {
    int r;
    asm("cvt.sat.s16.s32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}

template <int SHIFT, bool SAT16>
__device__ __forceinline__ void idct8(int& v0, int& v1, int& v2, int& v3, int& v4, int& v5, int& v6,
                                      int& v7) {
    // even part
    int z1 = (v2 + v6) * LP_FIX_0_541196100;
    const int tmp2 = z1 - v6 * LP_FIX_1_847759065;
    const int tmp3 = z1 + v2 * LP_FIX_0_765366865;
    const int tmp0 = (v0 + v4) << 13;
    const int tmp1 = (v0 - v4) << 13;
    const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    // odd part
    int t0 = v7, t1 = v5, t2 = v3, t3 = v1;
    z1 = t0 + t3;
    int z2 = t1 + t2, z3 = t0 + t2, z4 = t1 + t3;
    const int z5 = (z3 + z4) * LP_FIX_1_175875602;
    t0 *= LP_FIX_0_298631336;
    t1 *= LP_FIX_2_053119869;
    t2 *= LP_FIX_3_072711026;
    t3 *= LP_FIX_1_501321110;
    z1 *= -LP_FIX_0_899976223;
    z2 *= -LP_FIX_2_562915447;
    z3 = z3 * -LP_FIX_1_961570560 + z5;
    z4 = z4 * -LP_FIX_0_390180644 + z5;
    t0 += z1 + z3;
    t1 += z2 + z4;
    t2 += z2 + z3;
    t3 += z1 + z4;
    constexpr int R = 1 << (SHIFT - 1);
    v0 = (tmp10 + t3 + R) >> SHIFT;
    v7 = (tmp10 - t3 + R) >> SHIFT;
    v1 = (tmp11 + t2 + R) >> SHIFT;
    v6 = (tmp11 - t2 + R) >> SHIFT;
    v2 = (tmp12 + t1 + R) >> SHIFT;
    v5 = (tmp12 - t1 + R) >> SHIFT;
    v3 = (tmp13 + t0 + R) >> SHIFT;
    v4 = (tmp13 - t0 + R) >> SHIFT;
    if (SAT16) {  // int16 saturation between the passes (what the SIMD IDCT's packssdw does): one cvt.sat each
        v0 = sat_s16(v0); v1 = sat_s16(v1); v2 = sat_s16(v2); v3 = sat_s16(v3);
        v4 = sat_s16(v4); v5 = sat_s16(v5); v6 = sat_s16(v6); v7 = sat_s16(v7);
    }
}

// Four samples: clamp to [-128, 127], + 128, pack.  cvt.pack.sat.s8.s32 saturates and packs two values
// per instruction; the level shift is an XOR of the sign bits.
__device__ __forceinline__ uint32_t pack_px(int a, int b, int c, int d) {
    uint32_t hi, w;
    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(d), "r"(c), "r"(0));
    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(w) : "r"(b), "r"(a), "r"(hi));
    return w ^ 0x80808080u;
}

__global__ void __launch_bounds__(128)
    jpeg_idct_kernel(const JpegDecodeItem* items, const int16_t* coef, uint8_t* planes, int mcu_order) {
    const JpegDecodeItem& it = items[blockIdx.y];
    if (it.status != 0) return;
    // the image's quantisation tables, once per CTA (read back 8 entries per load below)
    __shared__ __align__(16) uint16_t s_qt[3][64];
    for (int i = threadIdx.x; i < 3 * 64; i += blockDim.x) s_qt[i >> 6][i & 63] = it.qt[i >> 6][i & 63];
    __syncthreads();
    const int blk = blockIdx.x * blockDim.x + threadIdx.x;
    int c = 0;
    if (it.ncomp == 3) c = blk >= (int)it.block_off[2] ? 2 : (blk >= (int)it.block_off[1] ? 1 : 0);
    const int nblk = (int)it.block_off[it.ncomp - 1] + it.bw[it.ncomp - 1] * it.bh[it.ncomp - 1];
    if (blk >= nblk) return;
    const int rel = blk - (int)it.block_off[c];
    const int X = rel % it.bw[c], Y = rel / it.bw[c];
    // The serial and multi-scan entropy decoders store blocks per component in raster order (index =
    // blk); the parallel decoder stores them in scan order (jpeg_huff_parallel.cu): block (X % h, Y % v)
    // of component c inside ROI MCU (X / h, Y / v).
    size_t sblk = (size_t)blk;
    if (mcu_order) {
        const int h = it.h[c], vv = it.v[c];
        int nb = 0, kfirst = 0;
        for (int k = 0; k < it.ncomp; k++) {
            if (k < c) kfirst += it.h[k] * it.v[k];
            nb += it.h[k] * it.v[k];
        }
        sblk = ((size_t)(Y / vv) * it.roi_mcx + X / h) * nb + kfirst + (Y % vv) * h + X % h;
    }
    const uint4* src = reinterpret_cast<const uint4*>(coef + it.coef_off + sblk * 64);
    const uint4* q4 = reinterpret_cast<const uint4*>(s_qt[c]);
    int v[64];
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const uint4 u = __ldg(src + r);
        const uint4 qq = q4[r];
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
        const uint32_t qw[4] = {qq.x, qq.y, qq.z, qq.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            // 16-bit wrapping multiply (pmullw), as libjpeg-turbo's SIMD dequantisation does
            const int lo = (int16_t)(w[k] & 0xffff), hi = (int16_t)(w[k] >> 16);
            v[r * 8 + 2 * k] = (int16_t)(lo * (int)(qw[k] & 0xffff));
            v[r * 8 + 2 * k + 1] = (int16_t)(hi * (int)(qw[k] >> 16));
        }
    }
#pragma unroll
    for (int x = 0; x < 8; x++)
        idct8<11, true>(v[x], v[8 + x], v[16 + x], v[24 + x], v[32 + x], v[40 + x], v[48 + x], v[56 + x]);
    const int stride = it.bw[c] * 8;
    uint8_t* dst = planes + it.plane_off + it.plane_rel[c] + (size_t)Y * 8 * stride + X * 8;
#pragma unroll
    for (int r = 0; r < 8; r++) {
        idct8<18, false>(v[r * 8], v[r * 8 + 1], v[r * 8 + 2], v[r * 8 + 3], v[r * 8 + 4], v[r * 8 + 5],
                         v[r * 8 + 6], v[r * 8 + 7]);
        uint2 o;
        o.x = pack_px(v[r * 8], v[r * 8 + 1], v[r * 8 + 2], v[r * 8 + 3]);
        o.y = pack_px(v[r * 8 + 4], v[r * 8 + 5], v[r * 8 + 6], v[r * 8 + 7]);
        *reinterpret_cast<uint2*>(dst + (size_t)r * stride) = o;
    }
}

// ------------------------------------------------------------------ upsample + colour

// Value of component c at full-resolution pixel (x, y): libjpeg-turbo jdsample.c.
__device__ __forceinline__ int upsampled(const JpegDecodeItem& it, const uint8_t* planes, int c, int maxh,
                                         int maxv, int x, int y) {
    const int stride = it.bw[c] * 8;
    // plane origin shifted so that full-image component coordinates index it directly
    const uint8_t* pl = planes + it.plane_off + it.plane_rel[c] -
                        ((size_t)(it.roi_my0 * it.v[c] * 8) * stride + (size_t)(it.roi_mx0 * it.h[c] * 8));
    const int hr = maxh / it.h[c], vr = maxv / it.v[c];
    const int cw = it.dw[c], ch = it.dh[c];
    if (hr == 1 && vr == 1) return pl[(size_t)y * stride + x];
    if (hr == 2 && vr == 2) {
        const int cy = y >> 1, i = x >> 1;
        const int fy = min(max((y & 1) ? cy + 1 : cy - 1, 0), ch - 1);
        const uint8_t* s0 = pl + (size_t)cy * stride;
        const uint8_t* s1 = pl + (size_t)fy * stride;
        const int cs = 3 * s0[i] + s1[i];
        if (x & 1) {
            if (i == cw - 1) return (cs * 4 + 7) >> 4;
            return (cs * 3 + 3 * s0[i + 1] + s1[i + 1] + 7) >> 4;
        }
        if (i == 0) return (cs * 4 + 8) >> 4;
        return (cs * 3 + 3 * s0[i - 1] + s1[i - 1] + 8) >> 4;
    }
    if (hr == 2 && vr == 1) {
        const uint8_t* s = pl + (size_t)y * stride;
        const int i = x >> 1;
        if (x & 1) return (i == cw - 1) ? s[i] : (3 * s[i] + s[i + 1] + 2) >> 2;
        return (i == 0) ? s[0] : (3 * s[i] + s[i - 1] + 1) >> 2;
    }
    if (hr == 1 && vr == 2) {
        const int cy = y >> 1;
        const int fy = min(max((y & 1) ? cy + 1 : cy - 1, 0), ch - 1);
        return (3 * pl[(size_t)cy * stride + x] + pl[(size_t)fy * stride + x] + ((y & 1) ? 2 : 1)) >> 2;
    }
    return pl[(size_t)(y / vr) * stride + x / hr];  // int_upsample (replication)
}

// Thread = 16 consecutive output pixels of one row.  4:2:0 frames whose rows are 16-byte aligned take
// the vector path (one 16-byte Y load, four 8-byte chroma loads, three 16-byte stores); everything
// else (other samplings, grayscale, odd widths) goes pixel by pixel through upsampled().
__global__ void __launch_bounds__(128)
    jpeg_upsample_color_kernel(const JpegDecodeItem* items, const uint8_t* planes, uint8_t* frames) {
    const JpegDecodeItem& it = items[blockIdx.y];
    if (it.status != 0) return;
    // flat index over (row, 16-pixel segment): a window row of 68 segments would leave half of a 128-thread
    // block idle if blocks were tied to rows (ncu round 2: 23 of 32 lanes active)
    const unsigned segs = (unsigned)(it.win_w + 15) >> 4;
    const unsigned f = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned row = f / segs;
    if (row >= (unsigned)it.win_h) return;
    const int x0 = it.win_x0 + (int)(f - row * segs) * 16;  // full-image coordinates
    const int y = it.win_y0 + (int)row;
    const int xe = it.win_x0 + it.win_w;
    uint8_t* out = frames + it.frame_off + (size_t)(y - it.win_y0) * it.win_stride;  // this window row
    if (it.ncomp == 1) {
        const int stride = it.bw[0] * 8;
        const uint8_t* pl = planes + it.plane_off + it.plane_rel[0] + (size_t)(y - it.roi_my0 * 8) * stride -
                            (size_t)(it.roi_mx0 * 8);
        for (int x = x0; x < min(x0 + 16, xe); x++) out[x - it.win_x0] = pl[x];
        return;
    }
    const bool fast = it.h[0] == 2 && it.v[0] == 2 && it.h[1] == 1 && it.v[1] == 1 && it.h[2] == 1 &&
                      it.v[2] == 1 && (it.width & 15) == 0 && (it.win_x0 & 15) == 0 && (it.win_w & 15) == 0 &&
                      (it.win_stride & 15) == 0 && (it.frame_off & 15) == 0;
    if (!fast) {
        const int maxh = max(it.h[0], max(it.h[1], it.h[2])), maxv = max(it.v[0], max(it.v[1], it.v[2]));
        for (int x = x0; x < min(x0 + 16, xe); x++) {
            const int Y = upsampled(it, planes, 0, maxh, maxv, x, y);
            const int cb = upsampled(it, planes, 1, maxh, maxv, x, y) - 128;
            const int cr = upsampled(it, planes, 2, maxh, maxv, x, y) - 128;
            const int r = Y + ((91881 * cr + 32768) >> 16);
            const int g = Y + ((-22554 * cb + 32768 - 46802 * cr) >> 16);
            const int b = Y + ((116130 * cb + 32768) >> 16);
            uint8_t* o = out + (size_t)(x - it.win_x0) * 3;
            o[0] = (uint8_t)min(max(b, 0), 255);
            o[1] = (uint8_t)min(max(g, 0), 255);
            o[2] = (uint8_t)min(max(r, 0), 255);
        }
        return;
    }
    const uint8_t* base = planes + it.plane_off;
    const int sy = it.bw[0] * 8, sc = it.bw[1] * 8;
    const int lx = it.roi_mx0 * 16, ly = it.roi_my0 * 16;  // luma / chroma plane origins (4:2:0)
    const int cx_ = it.roi_mx0 * 8, cy_ = it.roi_my0 * 8;
    const uint4 yv = *reinterpret_cast<const uint4*>(base + it.plane_rel[0] + (size_t)(y - ly) * sy + (x0 - lx));
    const uint32_t yw[4] = {yv.x, yv.y, yv.z, yv.w};
    const int cw = it.dw[1], chh = it.dh[1];
    const int cy = y >> 1;
    const int fy = min(max((y & 1) ? cy + 1 : cy - 1, 0), chh - 1);
    const int i0 = x0 >> 1;
    const int il = max(i0 - 1, 0), ir = min(i0 + 8, cw - 1);
    int cs[2][10];  // 3*near + far for chroma columns i0-1 .. i0+8, Cb and Cr
#pragma unroll
    for (int c = 0; c < 2; c++) {
        const uint8_t* pn = base + it.plane_rel[1 + c] + (size_t)(cy - cy_) * sc - cx_;
        const uint8_t* pf = base + it.plane_rel[1 + c] + (size_t)(fy - cy_) * sc - cx_;
        const uint2 n8 = *reinterpret_cast<const uint2*>(pn + i0);
        const uint2 f8 = *reinterpret_cast<const uint2*>(pf + i0);
        cs[c][0] = 3 * pn[il] + pf[il];
        cs[c][9] = 3 * pn[ir] + pf[ir];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            cs[c][1 + k] = 3 * (int)((n8.x >> (8 * k)) & 0xff) + (int)((f8.x >> (8 * k)) & 0xff);
            cs[c][5 + k] = 3 * (int)((n8.y >> (8 * k)) & 0xff) + (int)((f8.y >> (8 * k)) & 0xff);
        }
    }
    // 3 * (3*near + far) per chroma column once, so a pixel's chroma is add + add + shift; the -128 of
    // Cb / Cr is folded into the colour-conversion constants; clamps are one VIMNMX.RELU each; bytes are
    // assembled with PRMT (3 per output word) instead of shift + or per byte.
    int c3[2][10];
#pragma unroll
    for (int c = 0; c < 2; c++)
#pragma unroll
        for (int j = 1; j < 9; j++) c3[c][j] = 3 * cs[c][j];
    uint32_t ow[12];
#pragma unroll
    for (int q = 0; q < 4; q++) {  // 4 pixels -> 3 output words
        uint32_t B[4], G[4], R[4];
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const int k = 4 * q + t;
            const int j = 1 + (k >> 1);
            int cb, cr;
            if (k & 1) {
                cb = (c3[0][j] + cs[0][j + 1] + 7) >> 4;
                cr = (c3[1][j] + cs[1][j + 1] + 7) >> 4;
            } else {
                cb = (c3[0][j] + cs[0][j - 1] + 8) >> 4;
                cr = (c3[1][j] + cs[1][j - 1] + 8) >> 4;
            }
            const int Y = (int)__byte_perm(yw[q], 0, 0x4440 + t);  // byte t of the word, zero-extended
            // (c * (x - 128) + 32768) >> 16 with the -128 folded into the addend
            R[t] = (uint32_t)__vimin_s32_relu(Y + ((91881 * cr + (32768 - 91881 * 128)) >> 16), 255);
            G[t] = (uint32_t)__vimin_s32_relu(Y + ((-22554 * cb - 46802 * cr + (32768 + (22554 + 46802) * 128)) >> 16), 255);
            B[t] = (uint32_t)__vimin_s32_relu(Y + ((116130 * cb + (32768 - 116130 * 128)) >> 16), 255);
        }
        // bytes: B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3
        ow[3 * q + 0] = __byte_perm(__byte_perm(B[0], G[0], 0x0040), __byte_perm(R[0], B[1], 0x0040), 0x5410);
        ow[3 * q + 1] = __byte_perm(__byte_perm(G[1], R[1], 0x0040), __byte_perm(B[2], G[2], 0x0040), 0x5410);
        ow[3 * q + 2] = __byte_perm(__byte_perm(R[2], B[3], 0x0040), __byte_perm(G[3], R[3], 0x0040), 0x5410);
    }
    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)(x0 - it.win_x0) * 3);
    dst[0] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    dst[1] = make_uint4(ow[4], ow[5], ow[6], ow[7]);
    dst[2] = make_uint4(ow[8], ow[9], ow[10], ow[11]);
}

// ------------------------------------------------------------------ launcher

int jpeg_decode_launch(const JpegDecodeBatch& b, cudaStream_t st, cudaEvent_t ev_after_huff) {
    if (b.n <= 0) return LP_OK;
    LP_CUDA_OK(cudaMemsetAsync(b.coef, 0, b.coef_elems_total * sizeof(int16_t), st));
    if (b.scans) {
        jpeg_multiscan_kernel<<<1, 32, 0, st>>>(b.items, b.scans, b.nscans, b.tables, b.scan, b.coef);
        g_launches++;
        LP_CUDA_OK(cudaGetLastError());
    } else if (b.use_parallel_huffman) {
        JpegHuffParallelArgs a{b.items, b.tables, b.scan, b.clean, b.states, b.nslots, b.coef, b.dcdiff, b.n};
        int rc = jpeg_huff_parallel_launch(a, st);  // skips the images that carry restart markers
        if (rc) return rc;
        rc = jpeg_rst_launch(b.items, b.tables, b.scan, b.nslots, b.rst_work, b.n_rst_work, b.n, b.coef, st);
        if (rc) return rc;
    } else {
        const int threads = 32;
        jpeg_huff_decode_kernel<<<ceil_div(b.n, threads), threads, 0, st>>>(b.items, b.tables, b.scan,
                                                                           b.coef, b.n);
        g_launches++;
        LP_CUDA_OK(cudaGetLastError());
    }
    if (ev_after_huff) LP_CUDA_OK(cudaEventRecord(ev_after_huff, st));
    {
        dim3 grid(ceil_div(b.max_blocks_per_image, 128), b.n);
        const int mcu_order = !b.scans && b.use_parallel_huffman;
        jpeg_idct_kernel<<<grid, 128, 0, st>>>(b.items, b.coef, b.planes, mcu_order);
        g_launches++;
        LP_CUDA_OK(cudaGetLastError());
    }
    {
        const long segs = (long)ceil_div(b.max_width, 16) * b.max_height;
        dim3 grid((unsigned)((segs + 127) / 128), b.n);
        jpeg_upsample_color_kernel<<<grid, 128, 0, st>>>(b.items, b.planes, b.frames);
        g_launches++;
        LP_CUDA_OK(cudaGetLastError());
    }
    return LP_OK;
}

}  // namespace lp
