// png_decode.cu -- PNG pixel decode on sm_100a: zlib inflate -> scanline defilter -> packed
// Gray / BGR / BGRA u8, as cv::ImageDecoder::readData produces it for lilliput's 8-bit Framebuffer.
//
// Replaces: opencv_decoder_read_data for PNG inputs (ref opencv.cpp:166-171 -> OpenCV grfmt_png ->
// libpng 1.6.47 + zlib-ng 2.3.3).  Lossless: the result is bit-identical to the reference
// (tests/test_gpu_png.py).  Transform set: 16-bit -> high byte,
// palette -> BGR (+A with tRNS), gray 1/2/4 -> 8 by replication, gray+alpha -> B=G=R,A, RGB(A) -> BGR(A).
//
// A DEFLATE stream is one serial bit string with back-references, so the unit of parallelism is the
// image: one warp per image.
//   png_inflate_kernel   lane 0 walks the Huffman symbols (10-bit lookahead tables in shared memory);
//                        LZ77 copies and stored blocks are done by all 32 lanes.
//   png_defilter_kernel  32 scanlines at a time, lane r one scanline, skewed by one pixel per lane so
//                        that "up" comes from lane r-1 by shuffle and "left"/"upper-left" stay in
//                        registers (Sub/Up/Average/Paeth, modulo 256), in place.
//   png_convert_kernel   one thread per output pixel.
#include "common.cuh"
#include "kernels.cuh"

namespace lp {

constexpr int kPngWarps = 4;  // warps (= images) per CTA
constexpr int kLitBits = 10, kDistBits = 8;

struct InflateShared {
    uint16_t lit[1 << kLitBits];    // (symbol << 4) | length, 0 = longer than kLitBits
    uint16_t dist[1 << kDistBits];
    uint16_t lcount[16], dcount[16];
    uint16_t lsym[288], dsym[32];
    uint8_t lens[320];
};

struct LsbBits {
    const uint8_t* p;
    const uint8_t* end;
    uint64_t acc;
    int cnt;
    // 32 bits per refill: two aligned word loads + a funnel shift instead of a chain of byte loads
    // (the refill sits on the critical path of every symbol of a serial decoder)
    __device__ __forceinline__ void fill() {
        if (cnt > 32) return;
        uint32_t w;
        if (p + 4 <= end) {
            const uintptr_t a = reinterpret_cast<uintptr_t>(p);
            const uint32_t* q = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
            w = __funnelshift_r(q[0], q[1], 8 * (int)(a & 3));  // q[1] may lie past `end`: inside the padded buffer
        } else {
            w = 0;
            for (int k = 0; k < 4; k++) w |= (uint32_t)(p + k < end ? p[k] : 0) << (8 * k);
        }
        acc |= (uint64_t)w << cnt;
        p += 4;
        cnt += 32;
    }
    __device__ __forceinline__ uint32_t get(int n) {
        if (cnt < n) fill();
        const uint32_t v = (uint32_t)(acc & ((1ull << n) - 1));
        acc >>= n;
        cnt -= n;
        return v;
    }
};

// Canonical code -> lookup tables (lane 0).  Returns false for an over-subscribed code.
__device__ bool build_tables(const uint8_t* lens, int n, uint16_t* count, uint16_t* sym, uint16_t* look,
                             int look_bits) {
    for (int l = 0; l < 16; l++) count[l] = 0;
    for (int i = 0; i < n; i++) count[lens[i]]++;
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
        if (lens[i]) sym[offs[lens[i]]++] = (uint16_t)i;
    for (int i = 0; i < (1 << look_bits); i++) look[i] = 0;
    // codes in canonical order; the stream carries them MSB first inside an LSB-first bit string,
    // so the lookahead index is the bit-reversed code
    int code = 0, k = 0;
    for (int l = 1; l <= look_bits; l++) {
        for (int c = 0; c < count[l]; c++, k++, code++) {
            const uint32_t rev = __brev((uint32_t)code) >> (32 - l);
            for (uint32_t j = rev; j < (1u << look_bits); j += 1u << l) look[j] = (uint16_t)((sym[k] << 4) | l);
        }
        code <<= 1;
    }
    return true;
}

__device__ __forceinline__ int decode_sym(LsbBits& b, const uint16_t* look, int look_bits,
                                          const uint16_t* count, const uint16_t* sym) {
    if (b.cnt < 32) b.fill();
    const uint32_t e = look[b.acc & ((1u << look_bits) - 1)];
    if (e) {
        b.acc >>= (e & 15);
        b.cnt -= (e & 15);
        return (int)(e >> 4);
    }
    int code = 0, first = 0, index = 0;  // canonical walk, one bit at a time
    for (int l = 1; l < 16; l++) {
        code |= (int)(b.acc & 1);
        b.acc >>= 1;
        b.cnt--;
        const int c = count[l];
        if (code - c < first) return sym[index + (code - first)];
        index += c;
        first = (first + c) << 1;
        code <<= 1;
    }
    return -1;
}

__constant__ uint16_t c_lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t c_lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t c_dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__constant__ uint8_t c_dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__constant__ uint8_t c_clorder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

__global__ void __launch_bounds__(kPngWarps * 32)
    png_inflate_kernel(PngDecodeItem* items, const uint8_t* zall, uint8_t* rawall, int n) {
    __shared__ InflateShared sh_all[kPngWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int img = blockIdx.x * kPngWarps + warp;
    if (img >= n) return;
    PngDecodeItem& it = items[img];
    if (it.status != 0) return;
    InflateShared& sh = sh_all[warp];
    const uint8_t* z = zall + it.z_off;
    uint8_t* out = rawall + it.raw_off;
    const uint32_t cap = it.raw_total;
    uint32_t o = 0;
    int status = 0;
    LsbBits b{z + 2, z + it.z_len, 0, 0};
    if (it.z_len < 2 || (z[0] & 15) != 8 || (z[1] & 0x20)) status = -3;
    // Commands broadcast from lane 0 to the warp: 0 = literal run done / continue, 1 = copy, 2 = stop
    int last = 0;
    while (status == 0 && !last) {
        int type = 0;
        if (lane == 0) {
            last = (int)b.get(1);
            type = (int)b.get(2);
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        type = __shfl_sync(0xffffffffu, type, 0);
        if (type == 0) {  // stored block: all lanes copy
            uint32_t len = 0;
            const uint8_t* src = nullptr;
            if (lane == 0) {
                b.get(b.cnt & 7);
                // un-read whole bytes still in the accumulator
                b.p -= b.cnt >> 3;
                b.acc = 0;
                b.cnt = 0;
                const uint32_t l0 = b.p + 4 <= b.end ? (b.p[0] | (b.p[1] << 8)) : 0;
                const uint32_t nl = b.p + 4 <= b.end ? (b.p[2] | (b.p[3] << 8)) : 1;
                if ((l0 ^ 0xFFFF) != nl || b.p + 4 + l0 > b.end || o + l0 > cap) status = -3;
                len = l0;
                src = b.p + 4;
                b.p += 4 + l0;
            }
            status = __shfl_sync(0xffffffffu, status, 0);
            len = __shfl_sync(0xffffffffu, len, 0);
            src = (const uint8_t*)__shfl_sync(0xffffffffu, (unsigned long long)src, 0);
            if (status) break;
            for (uint32_t i = lane; i < len; i += 32) out[o + i] = src[i];
            o += len;
            __syncwarp();
            continue;
        }
        if (type == 3) {
            status = -3;
            break;
        }
        if (lane == 0) {
            if (type == 1) {
                int i = 0;
                for (; i < 144; i++) sh.lens[i] = 8;
                for (; i < 256; i++) sh.lens[i] = 9;
                for (; i < 280; i++) sh.lens[i] = 7;
                for (; i < 288; i++) sh.lens[i] = 8;
                build_tables(sh.lens, 288, sh.lcount, sh.lsym, sh.lit, kLitBits);
                for (i = 0; i < 30; i++) sh.lens[i] = 5;
                build_tables(sh.lens, 30, sh.dcount, sh.dsym, sh.dist, kDistBits);
            } else {
                const int nl = (int)b.get(5) + 257, nd = (int)b.get(5) + 1, nc = (int)b.get(4) + 4;
                if (nl > 286 || nd > 30) status = -3;
                uint8_t cl[19];
                for (int i = 0; i < 19; i++) cl[i] = 0;
                for (int i = 0; i < nc && !status; i++) cl[c_clorder[i]] = (uint8_t)b.get(3);
                // code-length code: reuse the distance tables as scratch
                if (!status && !build_tables(cl, 19, sh.dcount, sh.dsym, sh.dist, 7)) status = -3;
                int i = 0;
                while (!status && i < nl + nd) {
                    const int s = decode_sym(b, sh.dist, 7, sh.dcount, sh.dsym);
                    if (s < 0) { status = -3; break; }
                    if (s < 16) { sh.lens[i++] = (uint8_t)s; continue; }
                    int rep, v = 0;
                    if (s == 16) {
                        if (!i) { status = -3; break; }
                        v = sh.lens[i - 1];
                        rep = 3 + (int)b.get(2);
                    } else if (s == 17) rep = 3 + (int)b.get(3);
                    else rep = 11 + (int)b.get(7);
                    if (i + rep > nl + nd) { status = -3; break; }
                    while (rep--) sh.lens[i++] = (uint8_t)v;
                }
                if (!status && !build_tables(sh.lens, nl, sh.lcount, sh.lsym, sh.lit, kLitBits)) status = -3;
                if (!status) build_tables(sh.lens + nl, nd, sh.dcount, sh.dsym, sh.dist, kDistBits);
            }
        }
        status = __shfl_sync(0xffffffffu, status, 0);
        if (status) break;
        __syncwarp();
        // symbols: lane 0 decodes; a match is handed to the whole warp
        for (;;) {
            uint32_t len = 0, dist = 0;
            int cmd = 0;  // 1 = copy, 2 = end of block, 3 = error
            if (lane == 0) {
                for (;;) {
                    const int s = decode_sym(b, sh.lit, kLitBits, sh.lcount, sh.lsym);
                    if (s < 0) { cmd = 3; break; }
                    if (s < 256) {
                        if (o >= cap) { cmd = 3; break; }
                        out[o++] = (uint8_t)s;
                        continue;
                    }
                    if (s == 256) { cmd = 2; break; }
                    const int ls = s - 257;
                    if (ls >= 29) { cmd = 3; break; }
                    len = c_lbase[ls] + b.get(c_lext[ls]);
                    const int ds = decode_sym(b, sh.dist, kDistBits, sh.dcount, sh.dsym);
                    if (ds < 0 || ds >= 30) { cmd = 3; break; }
                    dist = c_dbase[ds] + b.get(c_dext[ds]);
                    if (dist > o || o + len > cap) { cmd = 3; break; }
                    cmd = 1;
                    break;
                }
            }
            cmd = __shfl_sync(0xffffffffu, cmd, 0);
            if (cmd == 2) break;
            if (cmd == 3) { status = -3; break; }
            len = __shfl_sync(0xffffffffu, len, 0);
            dist = __shfl_sync(0xffffffffu, dist, 0);
            o = __shfl_sync(0xffffffffu, o, 0);
            // out[o+i] = out[o+i-dist]; with dist < len the pattern repeats with period dist, so every
            // source byte already exists: out[o - dist + (i mod dist)]
            __syncwarp();
            for (uint32_t i = lane; i < len; i += 32) out[o + i] = out[o - dist + (dist >= len ? i : i % dist)];
            __syncwarp();
            o += len;
        }
        o = __shfl_sync(0xffffffffu, o, 0);
        if (status) break;
    }
    o = __shfl_sync(0xffffffffu, o, 0);
    if (lane == 0) {
        if (!status && o < cap) status = -3;  // fewer scanline bytes than the header promises
        it.status = status;
        it.produced = o;
    }
}

// ------------------------------------------------------------------ defilter

__device__ __forceinline__ int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// BPP = bytes per complete pixel (1,2,3,4,6,8).  Lane r owns scanline y0 + r and, at step t, pixel
// t - r: the pixel above it was produced by lane r-1 one step earlier.
template <int BPP>
__device__ void defilter_image(uint8_t* raw, uint32_t row_bytes, int height) {
    const int lane = threadIdx.x & 31;
    const uint32_t pitch = row_bytes + 1;
    const int npx = (int)((row_bytes + BPP - 1) / BPP);
    for (int y0 = 0; y0 < height; y0 += 32) {
        const int y = y0 + lane;
        const bool live = y < height;
        uint8_t* row = raw + (size_t)(live ? y : 0) * pitch;
        const int ft = live ? row[0] : 0;
        const uint8_t* above = (y0 > 0) ? raw + (size_t)(y0 - 1) * pitch + 1 : nullptr;  // lane 0's "up"
        uint8_t left[BPP], up[BPP], upleft[BPP], cur[BPP];
#pragma unroll
        for (int k = 0; k < BPP; k++) left[k] = up[k] = upleft[k] = cur[k] = 0;
        for (int t = 0; t < npx + 31; t++) {
            // what lane r-1 produced in the previous step is the pixel above this lane's pixel
            uint32_t lo = 0, hi = 0;
#pragma unroll
            for (int k = 0; k < BPP; k++) {
                if (k < 4) lo |= (uint32_t)cur[k] << (8 * k);
                else hi |= (uint32_t)cur[k] << (8 * (k - 4));
            }
            uint32_t ulo = __shfl_up_sync(0xffffffffu, lo, 1), uhi = BPP > 4 ? __shfl_up_sync(0xffffffffu, hi, 1) : 0;
            const int x = t - lane;
            const bool act = live && x >= 0 && x < npx;
#pragma unroll
            for (int k = 0; k < BPP; k++) {
                upleft[k] = up[k];
                if (lane == 0) {
                    const uint32_t bi = (uint32_t)x * BPP + k;
                    up[k] = (above && act && bi < row_bytes) ? above[bi] : 0;
                } else {
                    up[k] = (uint8_t)((k < 4 ? ulo >> (8 * k) : uhi >> (8 * (k - 4))) & 0xff);
                }
            }
            if (act) {
#pragma unroll
                for (int k = 0; k < BPP; k++) {
                    const uint32_t bi = (uint32_t)x * BPP + k;
                    if (bi < row_bytes) {
                        int v = row[1 + bi];
                        const int a = left[k], b = up[k], c = upleft[k];
                        v += ft == 1 ? a : ft == 2 ? b : ft == 3 ? ((a + b) >> 1) : ft == 4 ? paeth(a, b, c) : 0;
                        cur[k] = (uint8_t)v;
                        row[1 + bi] = (uint8_t)v;
                    }
                }
#pragma unroll
                for (int k = 0; k < BPP; k++) left[k] = cur[k];
            } else if (x < 0) {
#pragma unroll
                for (int k = 0; k < BPP; k++) up[k] = 0;  // not started: nothing above-left yet
            }
        }
        __syncwarp();
        __threadfence_block();
    }
}

__global__ void __launch_bounds__(kPngWarps * 32)
    png_defilter_kernel(PngDecodeItem* items, uint8_t* rawall, int n) {
    const int warp = threadIdx.x >> 5;
    const int img = blockIdx.x * kPngWarps + warp;
    if (img >= n) return;
    PngDecodeItem& it = items[img];
    if (it.status != 0) return;
    for (int ps = 0; ps < it.npass; ps++) {  // Adam7: every reduced image is filtered on its own
        if (it.pass_w[ps] == 0 || it.pass_h[ps] == 0) continue;
        uint8_t* raw = rawall + it.raw_off + it.pass_off[ps];
        const uint32_t rb = it.pass_rb[ps];
        const int ph = it.pass_h[ps];
        switch (it.bpp) {
            case 1: defilter_image<1>(raw, rb, ph); break;
            case 2: defilter_image<2>(raw, rb, ph); break;
            case 3: defilter_image<3>(raw, rb, ph); break;
            case 4: defilter_image<4>(raw, rb, ph); break;
            case 6: defilter_image<6>(raw, rb, ph); break;
            default: defilter_image<8>(raw, rb, ph); break;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------ convert

__global__ void png_convert_kernel(const PngDecodeItem* items, const uint8_t* rawall, uint8_t* frames) {
    const PngDecodeItem& it = items[blockIdx.z];
    if (it.status != 0) return;
    const int ox = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y;
    if (ox >= it.width || oy >= it.height) return;
    int x = ox, y = oy, ps = 0;
    if (it.interlace) {  // which reduced image holds (ox, oy), and where
        int x0, y0, sx, sy;
        if (oy & 1) { ps = 6; x0 = 0; y0 = 1; sx = 0; sy = 1; }
        else if (ox & 1) { ps = 5; x0 = 1; y0 = 0; sx = 1; sy = 1; }
        else if ((oy & 3) == 2) { ps = 4; x0 = 0; y0 = 2; sx = 1; sy = 2; }
        else if ((ox & 3) == 2) { ps = 3; x0 = 2; y0 = 0; sx = 2; sy = 2; }
        else if ((oy & 7) == 4) { ps = 2; x0 = 0; y0 = 4; sx = 2; sy = 3; }
        else if ((ox & 7) == 4) { ps = 1; x0 = 4; y0 = 0; sx = 3; sy = 3; }
        else { ps = 0; x0 = 0; y0 = 0; sx = 3; sy = 3; }
        x = (ox - x0) >> sx;
        y = (oy - y0) >> sy;
    }
    const uint8_t* c = rawall + it.raw_off + it.pass_off[ps] + (size_t)y * (it.pass_rb[ps] + 1) + 1;
    uint8_t* o = frames + it.frame_off + (size_t)oy * it.frame_stride + (size_t)ox * it.out_channels;
    const int bd = it.bit_depth, sc = it.src_channels, ct = it.color_type, och = it.out_channels;
    uint32_t s[4] = {0, 0, 0, 0};
    for (int k = 0; k < sc; k++) {
        const size_t bit = ((size_t)x * sc + k) * bd;
        if (bd == 16) s[k] = ((uint32_t)c[bit >> 3] << 8) | c[(bit >> 3) + 1];
        else if (bd == 8) s[k] = c[bit >> 3];
        else s[k] = (c[bit >> 3] >> (8 - bd - (bit & 7))) & ((1u << bd) - 1);
    }
    if (ct == 3) {
        const bool ok = s[0] < (uint32_t)it.npal;
        o[0] = ok ? it.palette[s[0] * 3 + 2] : 0;
        o[1] = ok ? it.palette[s[0] * 3 + 1] : 0;
        o[2] = ok ? it.palette[s[0] * 3 + 0] : 0;
        if (och == 4) o[3] = s[0] < (uint32_t)it.ntrns ? it.trns[s[0]] : 255;
    } else if (ct == 0) {
        o[0] = bd == 16 ? (uint8_t)(s[0] >> 8) : bd == 8 ? (uint8_t)s[0] : (uint8_t)(s[0] * (255u / ((1u << bd) - 1)));
    } else if (ct == 4) {
        const uint8_t g = bd == 16 ? (uint8_t)(s[0] >> 8) : (uint8_t)s[0];
        o[0] = o[1] = o[2] = g;
        o[3] = bd == 16 ? (uint8_t)(s[1] >> 8) : (uint8_t)s[1];
    } else {
        const int sh = bd == 16 ? 8 : 0;
        o[0] = (uint8_t)(s[2] >> sh);
        o[1] = (uint8_t)(s[1] >> sh);
        o[2] = (uint8_t)(s[0] >> sh);
        if (ct == 6) o[3] = (uint8_t)(s[3] >> sh);
        else if (och == 4)
            o[3] = (s[0] == it.trns_rgb[0] && s[1] == it.trns_rgb[1] && s[2] == it.trns_rgb[2]) ? 0 : 255;
    }
}

void png_item_set_passes(PngDecodeItem* it) {
    static const int X0[7] = {0, 4, 0, 2, 0, 1, 0}, Y0[7] = {0, 0, 4, 0, 2, 0, 1};
    static const int DX[7] = {8, 8, 4, 4, 2, 2, 1}, DY[7] = {8, 8, 8, 4, 4, 2, 2};
    const size_t bits = (size_t)it->src_channels * it->bit_depth;
    it->npass = it->interlace ? 7 : 1;
    uint32_t off = 0;
    for (int ps = 0; ps < it->npass; ps++) {
        const int x0 = it->interlace ? X0[ps] : 0, y0 = it->interlace ? Y0[ps] : 0;
        const int dx = it->interlace ? DX[ps] : 1, dy = it->interlace ? DY[ps] : 1;
        const int pw = it->width > x0 ? (it->width - x0 + dx - 1) / dx : 0;
        const int ph = it->height > y0 ? (it->height - y0 + dy - 1) / dy : 0;
        it->pass_w[ps] = pw;
        it->pass_h[ps] = ph;
        it->pass_off[ps] = off;
        it->pass_rb[ps] = (uint32_t)(((size_t)pw * bits + 7) / 8);
        if (pw && ph) off += (it->pass_rb[ps] + 1) * (uint32_t)ph;
    }
    it->raw_total = off;
}

int png_decode_launch(const PngDecodeBatch& b, cudaStream_t st) {
    if (b.n <= 0) return LP_OK;
    const int ctas = ceil_div(b.n, kPngWarps);
    png_inflate_kernel<<<ctas, kPngWarps * 32, 0, st>>>(b.items, b.z, b.raw, b.n);
    g_launches++;
    LP_CUDA_OK(cudaGetLastError());
    png_defilter_kernel<<<ctas, kPngWarps * 32, 0, st>>>(b.items, b.raw, b.n);
    g_launches++;
    LP_CUDA_OK(cudaGetLastError());
    dim3 grid(ceil_div(b.max_width, 128), b.max_height, b.n);
    png_convert_kernel<<<grid, 128, 0, st>>>(b.items, b.raw, b.frames);
    g_launches++;
    LP_CUDA_OK(cudaGetLastError());
    return LP_OK;
}

}  // namespace lp
