// png_encode.cu -- PNG encode on sm_100a: scanline filtering + DEFLATE (hash-chain LZ77 whose search effort follows
// PngCompression, per-chunk dynamic Huffman codes, independent 32 KB chunks joined by sync-flush blocks:
// deflate_enc_core.h) + container assembly.
//
// Replaces: opencv_encoder_write for ".png" (ref opencv.cpp:185-194 -> cv::ImageEncoder::write ->
// OpenCV grfmt_png -> libpng 1.6.47 + zlib-ng 2.3.3).  PNG is lossless and the contract for this
// path is DECODED-PIXEL equality with identical IHDR policy (BGR -> colour type 2, BGRA -> 6,
// Gray -> 0, 8-bit, non-interlaced, no ancillary chunks), not byte-identical files: reproducing
// zlib-ng's match finder bit for bit is neither possible nor useful.
// Filter policy follows what OpenCV asks libpng for: with IMWRITE_PNG_COMPRESSION given, libpng's
// adaptive minimum-sum-of-absolute-differences heuristic over None/Sub/Up/Average/Paeth; without
// it, Sub on every row.  Level 0 emits stored blocks.
//
// Kernels:
//   png_filter_kernel    warp per scanline: the five candidate sums, pick, write [type][bytes].
//   png_deflate_kernel   warp per 32 KB chunk: lane 0 runs defenc::write_chunk (LZ77 tokens into global scratch,
//                        symbol statistics -> length-limited dynamic Huffman codes in shared memory, or fixed
//                        codes / a stored block when smaller); every chunk ends byte-aligned with an empty stored
//                        block, so chunks are independent and concatenate by memcpy.  All lanes
//                        compute the chunk's Adler-32 partial sums.
//   png_pack_kernel      CTA per image: prefix sum of chunk sizes, compaction into one zlib stream.
// The host adds the signature, IHDR, IDAT framing, CRC-32 and IEND (a few hundred bytes of work
// per image next to the D2H copy it has to do anyway).
#include <cstring>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"

#define LP_DEF_FN static __device__
#define LP_DEF_TABLE static __device__ const
#include "deflate_enc_core.h"

namespace lp {

constexpr int kChunk = defenc::kChunk;        // uncompressed bytes per DEFLATE chunk
constexpr int kChunkOut = defenc::kChunkOut;  // worst case: stored fallback
constexpr int kDefWarps = 2;                  // (defenc::Work is ~15 KB of shared memory per chunk in flight)

// ------------------------------------------------------------------ filtering

__device__ __forceinline__ int paeth_pred(int a, int b, int c) {
    const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// Sample k of pixel x in PNG order (RGB[A] / Gray) from a packed BGR[A] / Gray frame row.
__device__ __forceinline__ int png_sample(const uint8_t* row, int x, int k, int C) {
    if (x < 0 || !row) return 0;
    const int c = (C >= 3 && k < 3) ? 2 - k : k;  // swap B and R
    return row[(size_t)x * C + c];
}

__global__ void __launch_bounds__(128)
    png_filter_kernel(const uint8_t* frame, size_t row_stride, int W, int H, int C, int adaptive, uint8_t* filt) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= H) return;
    const int y = warp;
    const uint8_t* cur = frame + (size_t)y * row_stride;
    const uint8_t* up = y > 0 ? frame + (size_t)(y - 1) * row_stride : nullptr;
    const int nbytes = W * C;
    int best = 1;  // Sub (OpenCV's no-parameter default)
    if (adaptive) {
        uint32_t sum[5] = {0, 0, 0, 0, 0};
        for (int i = lane; i < nbytes; i += 32) {
            const int x = i / C, k = i % C;
            const int v = png_sample(cur, x, k, C), a = png_sample(cur, x - 1, k, C);
            const int b = png_sample(up, x, k, C), c = png_sample(up, x - 1, k, C);
            const int f[5] = {v, v - a, v - b, v - ((a + b) >> 1), v - paeth_pred(a, b, c)};
#pragma unroll
            for (int t = 0; t < 5; t++) {
                const int u = f[t] & 0xff;
                sum[t] += u < 128 ? u : 256 - u;  // libpng: |signed byte|
            }
        }
#pragma unroll
        for (int t = 0; t < 5; t++)
#pragma unroll
            for (int o = 16; o; o >>= 1) sum[t] += __shfl_xor_sync(0xffffffffu, sum[t], o);
        best = 0;
#pragma unroll
        for (int t = 1; t < 5; t++)
            if (sum[t] < sum[best]) best = t;  // strict <: the earliest filter wins ties, as in libpng
    }
    uint8_t* out = filt + (size_t)y * (nbytes + 1);
    if (lane == 0) out[0] = (uint8_t)best;
    for (int i = lane; i < nbytes; i += 32) {
        const int x = i / C, k = i % C;
        const int v = png_sample(cur, x, k, C), a = png_sample(cur, x - 1, k, C);
        const int b = png_sample(up, x, k, C), c = png_sample(up, x - 1, k, C);
        const int p = best == 0 ? 0 : best == 1 ? a : best == 2 ? b : best == 3 ? ((a + b) >> 1) : paeth_pred(a, b, c);
        out[1 + i] = (uint8_t)(v - p);
    }
}

// ------------------------------------------------------------------ DEFLATE

// chunk_len[i] = compressed bytes of chunk i (each chunk owns kChunkOut bytes of `comp`);
// adler[i] = {sum of bytes, position-weighted sum} mod 65521 for the host-side combine.
// scratch: per chunk kChunk uint16 of hash-chain links + defenc::kTokCap uint16 of tokens.
__global__ void __launch_bounds__(kDefWarps * 32)
    png_deflate_kernel(const uint8_t* filt, size_t total, int nchunks, int level, uint8_t* comp, uint32_t* chunk_len,
                       uint2* adler, uint16_t* scratch) {
    __shared__ defenc::Work work[kDefWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunk = blockIdx.x * kDefWarps + warp;
    if (chunk >= nchunks) return;
    const uint8_t* src = filt + (size_t)chunk * kChunk;
    const int n = (int)min((size_t)kChunk, total - (size_t)chunk * kChunk);
    uint8_t* dst = comp + (size_t)chunk * kChunkOut;
    // Adler-32 partials: A = sum b_i, B = sum (n - i) b_i
    uint64_t A = 0, B = 0;
    for (int i = lane; i < n; i += 32) {
        A += src[i];
        B += (uint64_t)(n - i) * src[i];
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        A += __shfl_xor_sync(0xffffffffu, A, o);
        B += __shfl_xor_sync(0xffffffffu, B, o);
    }
    __syncwarp();
    if (lane != 0) return;
    adler[chunk] = make_uint2((uint32_t)(A % 65521u), (uint32_t)(B % 65521u));
    uint16_t* prev = scratch + (size_t)chunk * (kChunk + defenc::kTokCap);
    chunk_len[chunk] = (uint32_t)defenc::write_chunk(src, n, level, work[warp], prev, prev + kChunk, dst);
}

// Concatenate the chunks of one image: out = 78 01 | chunks... | 03 00 (final empty fixed block).
__global__ void __launch_bounds__(256)
    png_pack_kernel(const uint8_t* comp, const uint32_t* chunk_len, int nchunks, uint8_t* out, size_t out_cap,
                    uint32_t* total_out) {
    __shared__ uint32_t s_off;
    if (threadIdx.x == 0) s_off = 2;
    __syncthreads();
    // chunks are few thousand at most: serial offsets by thread 0 per group, parallel copies
    for (int c = 0; c < nchunks; c++) {
        const uint32_t len = chunk_len[c];
        const uint32_t off = s_off;
        if ((size_t)off + len + 2 <= out_cap)
            for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) out[off + i] = comp[(size_t)c * kChunkOut + i];
        __syncthreads();
        if (threadIdx.x == 0) s_off = off + len;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const uint32_t off = s_off;
        if ((size_t)off + 2 <= out_cap) {
            out[0] = 0x78;
            out[1] = 0x01;
            out[off] = 0x03;  // BFINAL = 1, BTYPE = 01, EOB (7 zero bits)
            out[off + 1] = 0x00;
            *total_out = off + 2;
        } else {
            *total_out = 0;  // did not fit
        }
    }
}

// ------------------------------------------------------------------ host side

static uint32_t crc32_update(uint32_t c, const uint8_t* p, size_t n) {
    struct Table {
        uint32_t t[256];
        Table() {
            for (uint32_t i = 0; i < 256; i++) {
                uint32_t v = i;
                for (int k = 0; k < 8; k++) v = (v >> 1) ^ (0xEDB88320u & (0u - (v & 1)));
                t[i] = v;
            }
        }
    };
    static const Table tab;  // thread-safe static initialisation
    const uint32_t* table = tab.t;
    c = ~c;
    for (size_t i = 0; i < n; i++) c = table[(c ^ p[i]) & 0xff] ^ (c >> 8);
    return ~c;
}
static void put_be32(std::vector<uint8_t>& v, uint32_t x) {
    v.push_back((uint8_t)(x >> 24)); v.push_back((uint8_t)(x >> 16)); v.push_back((uint8_t)(x >> 8)); v.push_back((uint8_t)x);
}

// Encodes one packed device frame to a PNG file in `out` (host).  Returns LP_OK / error.
int png_encode_frame(const uint8_t* frame, size_t row_stride, int W, int H, int C, int level, bool adaptive,
                     std::vector<uint8_t>* out, cudaStream_t st) {
    if (W < 1 || H < 1 || (C != 1 && C != 3 && C != 4)) return LP_ERR_BAD_ARGUMENT;
    const size_t raw = ((size_t)W * C + 1) * H;
    const int nchunks = (int)ceil_div(raw, (size_t)kChunk);
    const size_t comp_bytes = (size_t)nchunks * kChunkOut;
    const size_t z_cap = comp_bytes + 16;
    uint8_t* buf = nullptr;
    const size_t off_comp = round_up(raw + 16, (size_t)256);
    const size_t off_len = off_comp + round_up(comp_bytes, (size_t)256);
    const size_t off_adler = off_len + round_up((size_t)nchunks * 4 + 4, (size_t)256);
    const size_t off_z = off_adler + round_up((size_t)nchunks * 8, (size_t)256);
    const size_t off_scratch = off_z + round_up(z_cap, (size_t)256);
    const size_t scratch_bytes = level == 0 ? 0 : (size_t)nchunks * (kChunk + defenc::kTokCap) * sizeof(uint16_t);
    LP_CUDA_OK(cudaMallocAsync(&buf, off_scratch + scratch_bytes, st));
    uint8_t* d_filt = buf;
    uint8_t* d_comp = buf + off_comp;
    uint32_t* d_len = reinterpret_cast<uint32_t*>(buf + off_len);
    uint2* d_adler = reinterpret_cast<uint2*>(buf + off_adler);
    uint8_t* d_z = buf + off_z;
    uint32_t* d_total = d_len + nchunks;
    png_filter_kernel<<<(unsigned)ceil_div((long long)H * 32, 128LL), 128, 0, st>>>(frame, row_stride, W, H, C, adaptive ? 1 : 0, d_filt);
    g_launches++;
    const int lvl = level < 0 ? 1 : level > 9 ? 9 : level;  // zlib's range; OpenCV's own default is 1
    png_deflate_kernel<<<ceil_div(nchunks, kDefWarps), kDefWarps * 32, 0, st>>>(d_filt, raw, nchunks, lvl, d_comp, d_len, d_adler,
                                                                              reinterpret_cast<uint16_t*>(buf + off_scratch));
    g_launches++;
    png_pack_kernel<<<1, 256, 0, st>>>(d_comp, d_len, nchunks, d_z, z_cap, d_total);
    g_launches++;
    LP_CUDA_OK(cudaGetLastError());
    uint32_t total = 0;
    std::vector<uint2> adl(nchunks);
    LP_CUDA_OK(cudaMemcpyAsync(&total, d_total, 4, cudaMemcpyDeviceToHost, st));
    LP_CUDA_OK(cudaMemcpyAsync(adl.data(), d_adler, (size_t)nchunks * 8, cudaMemcpyDeviceToHost, st));
    LP_CUDA_OK(cudaStreamSynchronize(st));
    if (total == 0 || total > z_cap) {  // (total > z_cap cannot come from the pack kernel; never read past d_z)
        cudaFreeAsync(buf, st);
        return LP_ERR_CUDA;
    }
    // container: signature, IHDR, one IDAT, IEND
    std::vector<uint8_t>& o = *out;
    o.clear();
    o.reserve((size_t)total + 80);
    static const uint8_t sig[8] = {0x89, 0x50, 0x4E, 0x47, 0x0D, 0x0A, 0x1A, 0x0A};
    o.insert(o.end(), sig, sig + 8);
    put_be32(o, 13);
    const size_t ihdr_at = o.size();
    o.insert(o.end(), {'I', 'H', 'D', 'R'});
    put_be32(o, (uint32_t)W);
    put_be32(o, (uint32_t)H);
    o.push_back(8);
    o.push_back(C == 1 ? 0 : C == 3 ? 2 : 6);
    o.push_back(0); o.push_back(0); o.push_back(0);
    put_be32(o, crc32_update(0, o.data() + ihdr_at, 17));
    put_be32(o, total + 4);  // zlib stream + Adler-32
    const size_t idat_at = o.size();
    o.insert(o.end(), {'I', 'D', 'A', 'T'});
    o.resize(o.size() + total);
    LP_CUDA_OK(cudaMemcpyAsync(o.data() + idat_at + 4, d_z, total, cudaMemcpyDeviceToHost, st));
    // Adler-32 of the filtered scanlines from the per-chunk partials
    uint64_t s1 = 1, s2 = 0;
    for (int c = 0; c < nchunks; c++) {
        const uint64_t len = std::min((size_t)kChunk, raw - (size_t)c * kChunk);
        s2 = (s2 + (len % 65521) * s1 + adl[c].y) % 65521;
        s1 = (s1 + adl[c].x) % 65521;
    }
    LP_CUDA_OK(cudaStreamSynchronize(st));
    put_be32(o, (uint32_t)((s2 << 16) | s1));
    put_be32(o, crc32_update(0, o.data() + idat_at, (size_t)total + 8));
    put_be32(o, 0);
    const size_t iend_at = o.size();
    o.insert(o.end(), {'I', 'E', 'N', 'D'});
    put_be32(o, crc32_update(0, o.data() + iend_at, 4));
    cudaFreeAsync(buf, st);
    return LP_OK;
}

}  // namespace lp
