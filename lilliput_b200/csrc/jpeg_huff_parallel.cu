// jpeg_huff_parallel.cu -- intra-image parallel Huffman decoding of a baseline JPEG scan that has
// NO restart markers (BASELINE config 2's primary corpus), one CTA per image.
//
// Replaces the bit-serial part of libjpeg-turbo's jdhuff.c that cv::ImageDecoder::readData runs for
// the reference (ref opencv.cpp:166-171).  Results (quantised coefficients) are identical to a
// sequential decode; only the schedule differs.
//
// JPEG's Huffman code self-synchronises: a decoder started at an arbitrary bit falls back onto
// true codeword boundaries (and the true position inside the MCU) after a few dozen symbols.  So
//   1. jpeg_unstuff_kernel   removes FF00 byte stuffing and finds the end of the entropy-coded
//                            segment, giving a plain bit string (block scan + compaction);
//   2. jpeg_huff_sync_kernel cuts it into 1024-bit subsequences; every thread decodes its own from a
//                            guessed state, then re-decodes from its left neighbour's exit state
//                            until no exit state changes (a fixed point that is exact by induction
//                            from subsequence 0).  A prefix sum of the per-subsequence coefficient
//                            counts gives every subsequence its absolute output position, and one
//                            last decode writes AC coefficients and DC differences;
//   3. the same kernel then turns DC differences into DC values with a per-component prefix sum.
// (Scheme after Weissenberger & Schmidt, "Accelerating JPEG Decompression on GPUs", restated from
// the published description.)
#include "common.cuh"
#include "kernels.cuh"

namespace lp {

constexpr int kSubBits = 1024;  // subsequence length in bits (128 bytes)
constexpr int kHuffThreads = 512;

__constant__ uint8_t c_zigzag_p[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,
                                       12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
                                       35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
                                       58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// ------------------------------------------------------------------ block scan helper

template <int THREADS>
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* total, uint32_t* warp_sums) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t s = lane < THREADS / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        if (lane < THREADS / 32) warp_sums[lane] = s;
    }
    __syncthreads();
    const uint32_t base = wid ? warp_sums[wid - 1] : 0;
    *total = warp_sums[THREADS / 32 - 1];
    __syncthreads();
    return base + inc - v;
}

// ------------------------------------------------------------------ 1. unstuff

// clean[] gets the entropy-coded bytes with every "FF 00" reduced to "FF"; the segment ends at the
// first FF that is followed by anything else (a marker).  clean_len[img] = bytes written.
__global__ void __launch_bounds__(kHuffThreads)
    jpeg_unstuff_kernel(JpegDecodeItem* items, const uint8_t* scan, uint8_t* clean) {
    __shared__ uint32_t warp_sums[kHuffThreads / 32];
    __shared__ uint32_t s_end, s_carry;
    JpegDecodeItem& it = items[blockIdx.x];
    const uint8_t* src = scan + it.scan_off;
    const uint32_t len = it.scan_len;
    uint8_t* dst = clean + it.clean_off;
    const int tid = threadIdx.x;
    if (tid == 0) {
        s_end = len;
        s_carry = 0;
    }
    __syncthreads();
    if (it.status != 0) {
        if (tid == 0) it.clean_len = 0;
        return;
    }
    // pass 1: first marker = FF followed by a byte that is not 00 (a trailing lone FF also ends it)
    uint32_t my_end = len;
    for (uint32_t i = tid; i < len; i += kHuffThreads) {
        if (src[i] == 0xFF) {
            const uint32_t nx = (i + 1 < len) ? src[i + 1] : 0xD9;
            if (nx != 0x00) {
                my_end = i;
                break;  // positions only grow along this thread's stride
            }
        }
    }
    atomicMin(&s_end, my_end);
    __syncthreads();
    const uint32_t end = s_end;
    // pass 2: compaction, 8 bytes per thread per round
    constexpr uint32_t kPer = 8;
    for (uint32_t base = 0; base < end; base += kHuffThreads * kPer) {
        const uint32_t b0 = base + tid * kPer;
        uint8_t v[kPer];
        uint32_t keep = 0, cnt = 0;
        uint8_t prev = (b0 > 0 && b0 <= end) ? src[b0 - 1] : 0;
#pragma unroll
        for (uint32_t k = 0; k < kPer; k++) {
            const uint32_t i = b0 + k;
            const uint8_t c = i < end ? src[i] : 0;
            v[k] = c;
            const bool drop = (i >= end) || (c == 0x00 && prev == 0xFF);
            if (!drop) {
                keep |= 1u << k;
                cnt++;
            }
            prev = c;
        }
        uint32_t total;
        const uint32_t ex = block_excl_scan<kHuffThreads>(cnt, &total, warp_sums);
        uint32_t o = s_carry + ex;
#pragma unroll
        for (uint32_t k = 0; k < kPer; k++)
            if (keep & (1u << k)) dst[o++] = v[k];
        __syncthreads();
        if (tid == 0) s_carry += total;
        __syncthreads();
    }
    if (tid == 0) it.clean_len = s_carry;
    // zero padding so the 8-byte window loads past the end read defined data
    for (uint32_t k = tid; k < 16; k += kHuffThreads) dst[s_carry + k] = 0;
}

// ------------------------------------------------------------------ 2. sync + write

struct SubState {
    uint32_t p;      // absolute bit position of the first symbol that starts after this subsequence
    uint32_t phase;  // (block-in-MCU << 6) | zig-zag index expected at p
};

struct HuffShared {
    uint16_t look[8][512];
    int32_t maxcode[8][18];
    int32_t valoffset[8][17];
    uint8_t vals[8][256];
    uint8_t zz[64];
    uint8_t blk_dc[16], blk_ac[16];  // table index per block-in-MCU
};

// 32 bits of the clean stream starting at bit position p.
__device__ __forceinline__ uint32_t peek32(const uint8_t* s, uint32_t p) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(s) + (p >> 5);
    const uint32_t a = __byte_perm(w[0], 0, 0x0123), b = __byte_perm(w[1], 0, 0x0123);
    return __funnelshift_l(b, a, p & 31);
}

// Decode symbols that START in [p, limit).  Returns the exit state and the number of coefficient
// slots consumed.  WRITE: store AC coefficients / DC differences at absolute slot `pos`.
template <bool WRITE>
__device__ __forceinline__ void decode_span(const HuffShared& hs, const uint8_t* s, uint32_t& p, uint32_t limit,
                                            uint32_t& phase, uint32_t& nslots, int nb,
                                            // WRITE-only state:
                                            uint64_t pos, uint64_t total_slots, const JpegDecodeItem* it,
                                            int16_t* coef, int16_t* dcdiff, int* status) {
    uint32_t blk = phase >> 6, z = phase & 63;
    uint32_t n = 0;
    // WRITE: current block's destination
    int16_t* dstblk = nullptr;
    auto locate = [&](uint64_t slot) {
        // absolute block counter -> (mcu, k) -> plane block
        const uint64_t babs = slot >> 6;
        const uint32_t mcu = (uint32_t)(babs / (uint32_t)nb), k = (uint32_t)(babs % (uint32_t)nb);
        int c = 0, kk = (int)k;
        while (c < it->ncomp - 1 && kk >= it->h[c] * it->v[c]) {
            kk -= it->h[c] * it->v[c];
            c++;
        }
        const int bx = kk % it->h[c], by = kk / it->h[c];
        const int mx = (int)(mcu % (uint32_t)it->mcus_x), my = (int)(mcu / (uint32_t)it->mcus_x);
        const int X = mx * it->h[c] + bx, Y = my * it->v[c] + by;
        dstblk = coef + it->coef_off + ((size_t)it->block_off[c] + (size_t)Y * it->bw[c] + X) * 64;
    };
    if (WRITE && pos < total_slots) locate(pos);
    while (p < limit) {
        if (WRITE && pos + n >= total_slots) break;  // all MCUs done: the rest is padding
        const uint32_t w = peek32(s, p);
        const int t = z == 0 ? hs.blk_dc[blk] : hs.blk_ac[blk];
        uint32_t e = hs.look[t][w >> 23];
        int len, sym;
        if (e) {
            len = e >> 8;
            sym = e & 0xFF;
        } else {
            len = 10;
            int code = (int)(w >> 22);
            while (len <= 16 && code > hs.maxcode[t][len]) {
                len++;
                code = (int)(w >> (32 - len));
            }
            if (len > 16) {  // not a codeword: only possible on a wrong guess (or a corrupt stream)
                if (WRITE) {
                    *status = -3;
                    break;
                }
                p += 1;
                continue;
            }
            sym = hs.vals[t][(code + hs.valoffset[t][len]) & 0xFF];
        }
        if (z == 0) {
            const int sz = sym & 15;
            if (WRITE) {
                int diff = 0;
                if (sz) {
                    const int v = (int)((w << len) >> (32 - sz));
                    diff = v < (1 << (sz - 1)) ? v - (1 << sz) + 1 : v;
                }
                dcdiff[(pos + n) >> 6] = (int16_t)diff;
            }
            p += len + sz;
            z = 1;
            n += 1;
        } else {
            const int r = sym >> 4, sz = sym & 15;
            if (sz == 0) {
                p += len;
                if (r == 15) {  // ZRL
                    if (z + 16 > 63) {  // runs past the block (libjpeg just ends the block here)
                        n += 64 - z;
                        z = 0;
                    } else {
                        z += 16;
                        n += 16;
                    }
                } else {  // EOB
                    n += 64 - z;
                    z = 0;
                }
            } else {
                uint32_t zn = z + r;
                if (zn > 63) {  // wrong guess (or corrupt data): close the block
                    if (WRITE) *status = -3;
                    p += len + sz;
                    n += 64 - z;
                    z = 0;
                } else {
                    if (WRITE) {
                        const int v = (int)((w << len) >> (32 - sz));
                        dstblk[hs.zz[zn]] = (int16_t)(v < (1 << (sz - 1)) ? v - (1 << sz) + 1 : v);
                    }
                    p += len + sz;
                    n += r + 1;
                    z = zn + 1;
                    if (z == 64) z = 0;
                }
            }
            if (z == 0) {
                blk = blk + 1 == (uint32_t)nb ? 0 : blk + 1;
                if (WRITE && pos + n < total_slots) locate(pos + n);
            }
        }
    }
    phase = (blk << 6) | z;
    nslots = n;
}

__global__ void __launch_bounds__(kHuffThreads)
    jpeg_huff_sync_kernel(JpegDecodeItem* items, const JpegHuffSet* tables, const uint8_t* clean,
                          SubState* states_all, uint32_t* nslots_all, int16_t* coef, int16_t* dcdiff_all) {
    __shared__ HuffShared hs;
    __shared__ uint32_t warp_sums[kHuffThreads / 32];
    __shared__ uint32_t s_carry;
    __shared__ int s_changed, s_status;
    JpegDecodeItem& it = items[blockIdx.x];
    const int tid = threadIdx.x;
    if (it.status != 0) return;
    // ---- tables to shared memory
    {
        const JpegHuffSet* g = tables + it.table_set;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(g);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&hs);
        static_assert(sizeof(JpegHuffSet) % 4 == 0, "table layout");
        for (int i = tid; i < (int)(sizeof(JpegHuffSet) / 4); i += kHuffThreads) dst[i] = src[i];
        if (tid < 64) hs.zz[tid] = c_zigzag_p[tid];
        if (tid == 0) {
            int k = 0;
            for (int c = 0; c < it.ncomp; c++)
                for (int j = 0; j < it.h[c] * it.v[c]; j++, k++) {
                    hs.blk_dc[k] = (uint8_t)it.td[c];
                    hs.blk_ac[k] = (uint8_t)(4 + it.ta[c]);
                }
            s_changed = 0;
            s_status = 0;
            s_carry = 0;
        }
    }
    __syncthreads();
    int nb = 0;
    for (int c = 0; c < it.ncomp; c++) nb += it.h[c] * it.v[c];
    const uint8_t* s = clean + it.clean_off;
    const uint32_t total_bits = it.clean_len * 8u;
    const uint32_t nsub = (total_bits + kSubBits - 1) / kSubBits;
    SubState* st = states_all + it.state_off;   // two buffers of nsub each
    uint32_t* ns = nslots_all + it.state_off / 2;  // nsub entries
    const uint64_t total_slots = (uint64_t)it.mcus_x * it.mcus_y * nb * 64;
    int16_t* dcdiff = dcdiff_all + it.dcdiff_off;

    // ---- pass 0: every subsequence from a guessed state (exact only for subsequence 0)
    for (uint32_t i = tid; i < nsub; i += kHuffThreads) {
        uint32_t p = i * kSubBits, phase = 0, n = 0;
        const uint32_t limit = min((i + 1) * kSubBits, total_bits);
        decode_span<false>(hs, s, p, limit, phase, n, nb, 0, 0, nullptr, nullptr, nullptr, nullptr);
        st[i] = SubState{p, phase};
        ns[i] = n;
    }
    __syncthreads();
    // ---- synchronisation: re-decode from the left neighbour's exit state until nothing changes
    SubState* cur = st;
    SubState* nxt = st + nsub;
    for (uint32_t iter = 0; iter < nsub; iter++) {
        for (uint32_t i = tid; i < nsub; i += kHuffThreads) {
            SubState out = cur[i];
            if (i > 0) {
                const SubState in = cur[i - 1];
                uint32_t p = in.p, phase = in.phase, n = 0;
                const uint32_t limit = min((i + 1) * kSubBits, total_bits);
                if (p < limit) decode_span<false>(hs, s, p, limit, phase, n, nb, 0, 0, nullptr, nullptr, nullptr, nullptr);
                if (p != out.p || phase != out.phase || n != ns[i]) {
                    out = SubState{p, phase};
                    ns[i] = n;
                    s_changed = 1;
                }
            }
            nxt[i] = out;
        }
        __syncthreads();
        const int changed = s_changed;
        __syncthreads();
        if (tid == 0) s_changed = 0;
        SubState* tmp = cur;
        cur = nxt;
        nxt = tmp;
        if (!changed) break;
        __syncthreads();
    }
    __syncthreads();
    // ---- prefix sum of slot counts, then the writing decode
    for (uint32_t base = 0; base < nsub; base += kHuffThreads) {
        const uint32_t i = base + tid;
        const uint32_t v = i < nsub ? ns[i] : 0;
        uint32_t total;
        const uint32_t ex = block_excl_scan<kHuffThreads>(v, &total, warp_sums);
        // slot counts of one image fit 32 bits up to 64 Mpixel-components; positions kept in 64
        const uint64_t pos = (uint64_t)s_carry + ex;
        if (i < nsub) {
            uint32_t p = i == 0 ? 0u : cur[i - 1].p;
            uint32_t phase = i == 0 ? 0u : cur[i - 1].phase;
            uint32_t n = 0;
            const uint32_t limit = min((i + 1) * kSubBits, total_bits);
            int status = 0;
            // the slot position implied by the prefix sum must agree with the carried phase
            if ((uint32_t)(pos % ((uint64_t)nb * 64)) != ((phase >> 6) * 64 + (phase & 63)) && pos < total_slots)
                status = -3;
            if (!status && p < limit)
                decode_span<true>(hs, s, p, limit, phase, n, nb, pos, total_slots, &it, coef, dcdiff, &status);
            if (status) s_status = status;
        }
        __syncthreads();
        if (tid == 0) s_carry += total;
        __syncthreads();
    }
    if (tid == 0) {
        // every MCU must have been produced
        if ((uint64_t)s_carry < total_slots) s_status = -3;
        if (s_status) it.status = s_status;
    }
    __syncthreads();
    if (s_status) return;
    // ---- 3. DC differences -> DC values: per component, prefix sum in MCU (scan) order
    int koff = 0;
    for (int c = 0; c < it.ncomp; c++) {
        const int bpc = it.h[c] * it.v[c];
        const uint32_t nblk = (uint32_t)it.mcus_x * it.mcus_y * bpc;
        if (tid == 0) s_carry = 0;
        __syncthreads();
        for (uint32_t base = 0; base < nblk; base += kHuffThreads) {
            const uint32_t j = base + tid;
            int d = 0;
            uint32_t mcu = 0, kk = 0;
            if (j < nblk) {
                mcu = j / bpc;
                kk = j % bpc;
                d = dcdiff[(size_t)mcu * nb + koff + kk];
            }
            // signed inclusive scan via unsigned wraparound arithmetic
            uint32_t total;
            const uint32_t ex = block_excl_scan<kHuffThreads>((uint32_t)d, &total, warp_sums);
            if (j < nblk) {
                const int dc = (int)(s_carry + ex + (uint32_t)d);
                const int bx = kk % it.h[c], by = kk / it.h[c];
                const int mx = (int)(mcu % (uint32_t)it.mcus_x), my = (int)(mcu / (uint32_t)it.mcus_x);
                const int X = mx * it.h[c] + bx, Y = my * it.v[c] + by;
                coef[it.coef_off + ((size_t)it.block_off[c] + (size_t)Y * it.bw[c] + X) * 64] = (int16_t)dc;
            }
            __syncthreads();
            if (tid == 0) s_carry += total;
            __syncthreads();
        }
        koff += bpc;
    }
}

// ------------------------------------------------------------------ launcher

int jpeg_huff_parallel_launch(const JpegHuffParallelArgs& a, cudaStream_t st) {
    if (a.n <= 0) return LP_OK;
    jpeg_unstuff_kernel<<<a.n, kHuffThreads, 0, st>>>(a.items, a.scan, a.clean);
    g_launches++;
    LP_CUDA_OK(cudaGetLastError());
    jpeg_huff_sync_kernel<<<a.n, kHuffThreads, 0, st>>>(a.items, a.tables, a.clean,
                                                       reinterpret_cast<SubState*>(a.states), a.nslots, a.coef,
                                                       a.dcdiff);
    g_launches++;
    LP_CUDA_OK(cudaGetLastError());
    return LP_OK;
}

}  // namespace lp
