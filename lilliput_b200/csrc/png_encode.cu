// png_encode.cu -- PNG encode on sm_100a: scanline filtering + DEFLATE (fixed Huffman, hashed
// LZ77, independent 32 KB chunks joined by sync-flush blocks) + container assembly.
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
//   png_deflate_kernel   warp per 32 KB chunk: lane 0 runs a greedy hash-chain-free LZ77 (one
//                        candidate per 4-byte hash, 4096-entry table in shared memory) and emits
//                        fixed-Huffman codes; every chunk ends byte-aligned with an empty stored
//                        block, so chunks are independent and concatenate by memcpy.  All lanes
//                        compute the chunk's Adler-32 partial sums.
//   png_pack_kernel      CTA per image: prefix sum of chunk sizes, compaction into one zlib stream.
// The host adds the signature, IHDR, IDAT framing, CRC-32 and IEND (a few hundred bytes of work
// per image next to the D2H copy it has to do anyway).
#include <cstring>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"

namespace lp {

constexpr int kChunk = 32768;            // uncompressed bytes per DEFLATE chunk
constexpr int kChunkOut = kChunk + 64;   // worst case: stored fallback
constexpr int kHashBits = 12;
constexpr int kDefWarps = 4;

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

// ------------------------------------------------------------------ DEFLATE, fixed Huffman

struct BitOut {
    uint8_t* p;
    uint64_t acc;
    int cnt;
    __device__ __forceinline__ void put(uint32_t v, int n) {  // LSB first
        acc |= (uint64_t)v << cnt;
        cnt += n;
        while (cnt >= 8) {
            *p++ = (uint8_t)acc;
            acc >>= 8;
            cnt -= 8;
        }
    }
    __device__ __forceinline__ void align() {
        if (cnt) {
            *p++ = (uint8_t)acc;
            acc = 0;
            cnt = 0;
        }
    }
};

// Fixed literal/length code of RFC 1951 3.2.6, bit-reversed for LSB-first packing.
__device__ __forceinline__ void put_litlen(BitOut& b, int sym) {
    uint32_t code;
    int len;
    if (sym < 144) { code = 0x30 + sym; len = 8; }
    else if (sym < 256) { code = 0x190 + (sym - 144); len = 9; }
    else if (sym < 280) { code = sym - 256; len = 7; }
    else { code = 0xC0 + (sym - 280); len = 8; }
    b.put(__brev(code) >> (32 - len), len);
}

__constant__ uint16_t c_len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t c_len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t c_dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__constant__ uint8_t c_dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

__device__ __forceinline__ void put_match(BitOut& b, int len, int dist) {
    int ls = 28;
    while (c_len_base[ls] > len) ls--;
    put_litlen(b, 257 + ls);
    if (c_len_extra[ls]) b.put((uint32_t)(len - c_len_base[ls]), c_len_extra[ls]);
    int ds = 29;
    while (c_dist_base[ds] > dist) ds--;
    b.put(__brev((uint32_t)ds) >> 27, 5);
    if (c_dist_extra[ds]) b.put((uint32_t)(dist - c_dist_base[ds]), c_dist_extra[ds]);
}

// chunk_len[i] = compressed bytes of chunk i (each chunk owns kChunkOut bytes of `comp`);
// adler[i] = {sum of bytes, position-weighted sum} mod 65521 for the host-side combine.
__global__ void __launch_bounds__(kDefWarps * 32)
    png_deflate_kernel(const uint8_t* filt, size_t total, int nchunks, int stored_only, uint8_t* comp,
                       uint32_t* chunk_len, uint2* adler) {
    __shared__ uint16_t hash_all[kDefWarps][1 << kHashBits];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunk = blockIdx.x * kDefWarps + warp;
    if (chunk >= nchunks) return;
    const uint8_t* src = filt + (size_t)chunk * kChunk;
    const int n = (int)min((size_t)kChunk, total - (size_t)chunk * kChunk);
    uint8_t* dst = comp + (size_t)chunk * kChunkOut;
    uint16_t* hash = hash_all[warp];
    for (int i = lane; i < (1 << kHashBits); i += 32) hash[i] = 0xFFFF;
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
    BitOut b{dst, 0, 0};
    bool stored = stored_only != 0;
    if (!stored) {
        b.put(2, 3);  // BFINAL = 0, BTYPE = 01 (fixed Huffman)
        int i = 0;
        while (i < n) {
            int best_len = 0, best_dist = 0;
            if (i + 4 <= n) {
                const uint32_t w = src[i] | (src[i + 1] << 8) | (src[i + 2] << 16) | ((uint32_t)src[i + 3] << 24);
                const uint32_t h = (w * 2654435761u) >> (32 - kHashBits);
                const int cand = hash[h];
                hash[h] = (uint16_t)i;
                if (cand != 0xFFFF && i - cand <= 32768 && cand < i) {
                    int l = 0;
                    const int maxl = min(258, n - i);
                    while (l < maxl && src[cand + l] == src[i + l]) l++;
                    if (l >= 4) {
                        best_len = l;
                        best_dist = i - cand;
                    }
                }
            }
            if (best_len) {
                put_match(b, best_len, best_dist);
                i += best_len;
            } else {
                put_litlen(b, src[i]);
                i++;
            }
            if ((int)(b.p - dst) > n + 16) {  // expanding: give up, store the chunk instead
                stored = true;
                break;
            }
        }
        if (!stored) {
            put_litlen(b, 256);  // end of block
            b.put(0, 3);         // empty stored block = sync flush: byte-aligns the chunk
            b.align();
            *b.p++ = 0x00; *b.p++ = 0x00; *b.p++ = 0xFF; *b.p++ = 0xFF;
        }
    }
    if (stored) {
        b = BitOut{dst, 0, 0};
        *b.p++ = 0x00;  // BFINAL = 0, BTYPE = 00, padding
        *b.p++ = (uint8_t)n; *b.p++ = (uint8_t)(n >> 8);
        *b.p++ = (uint8_t)~n; *b.p++ = (uint8_t)((~n) >> 8);
        for (int i = 0; i < n; i++) *b.p++ = src[i];
    }
    chunk_len[chunk] = (uint32_t)(b.p - dst);
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
    LP_CUDA_OK(cudaMallocAsync(&buf, off_z + z_cap, st));
    uint8_t* d_filt = buf;
    uint8_t* d_comp = buf + off_comp;
    uint32_t* d_len = reinterpret_cast<uint32_t*>(buf + off_len);
    uint2* d_adler = reinterpret_cast<uint2*>(buf + off_adler);
    uint8_t* d_z = buf + off_z;
    uint32_t* d_total = d_len + nchunks;
    png_filter_kernel<<<(unsigned)ceil_div((long long)H * 32, 128LL), 128, 0, st>>>(frame, row_stride, W, H, C, adaptive ? 1 : 0, d_filt);
    g_launches++;
    png_deflate_kernel<<<ceil_div(nchunks, kDefWarps), kDefWarps * 32, 0, st>>>(d_filt, raw, nchunks, level == 0,
                                                                              d_comp, d_len, d_adler);
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
