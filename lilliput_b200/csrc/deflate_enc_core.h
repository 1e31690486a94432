// deflate_enc_core.h -- the DEFLATE (RFC 1951) block writer behind the PNG encoder: greedy hashed LZ77 over one
// 32 KB chunk, per-chunk DYNAMIC Huffman codes (length-limited canonical codes built from the chunk's own symbol
// statistics, code lengths sent through the code-length code with run-length symbols 16 / 17 / 18), fixed codes or a
// stored block when those are smaller.
//
// Replaces, for opencv_encoder_write(".png") (ref opencv.cpp:173-194 -> cv::ImageEncoder -> libpng 1.6.47 -> zlib
// deflate): what libpng gets from zlib at compression levels 1..9 -- LZ77 + dynamic Huffman blocks.  PNG is lossless:
// the contract is decoded-pixel equality, not zlib's exact bytes (its match finder's choices are not reproducible and
// carry no meaning); the tests bound the size against libpng's file instead.
//
// One lane runs a chunk (chunks are independent: every chunk ends in a sync flush, so a file has thousands of them in
// flight).  The same source compiles for the host (LP_DEF_FN = static inline): tests/test_deflate_enc_core.py runs it
// against zlib's inflate on the CPU, the GPU suite shows the device writes the same bytes.
#pragma once
#include <cstddef>
#include <cstdint>

#ifndef LP_DEF_FN
#define LP_DEF_FN static inline
#endif
#ifndef LP_DEF_TABLE
#define LP_DEF_TABLE static const
#endif

namespace defenc {

constexpr int kChunk = 32768;           // uncompressed bytes per chunk
constexpr int kChunkOut = kChunk + 64;  // worst case: the stored form
constexpr int kHashBits = 12;
constexpr int kLit = 286, kDist = 30, kCl = 19;
constexpr int kTokCap = kChunk + 8;     // uint16 units: a literal takes one, a match two

LP_DEF_TABLE uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
LP_DEF_TABLE uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
LP_DEF_TABLE uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
LP_DEF_TABLE uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
LP_DEF_TABLE uint8_t kClOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

struct BitOut {  // LSB-first
    uint8_t* p;
    uint64_t acc;
    int cnt;
};
LP_DEF_FN void bo_put(BitOut& b, uint32_t v, int n) {
    b.acc |= (uint64_t)v << b.cnt;
    b.cnt += n;
    while (b.cnt >= 8) {
        *b.p++ = (uint8_t)b.acc;
        b.acc >>= 8;
        b.cnt -= 8;
    }
}
LP_DEF_FN void bo_align(BitOut& b) {
    if (b.cnt) {
        *b.p++ = (uint8_t)b.acc;
        b.acc = 0;
        b.cnt = 0;
    }
}
LP_DEF_FN uint32_t bit_reverse(uint32_t v, int n) {
    uint32_t r = 0;
    for (int i = 0; i < n; i++) r |= ((v >> i) & 1u) << (n - 1 - i);
    return r;
}

// Per-chunk working memory (shared memory on the device).
struct Work {
    uint32_t freq_l[kLit], freq_d[kDist], freq_c[kCl];
    uint8_t len_l[kLit], len_d[kDist], len_c[kCl];
    uint16_t code_l[kLit], code_d[kDist], code_c[kCl];  // bit-reversed, ready for LSB-first output
    // Huffman construction scratch (sized for the literal / length alphabet)
    uint32_t node_freq[2 * kLit];
    uint16_t node_parent[2 * kLit];
    uint16_t order[kLit];
    uint8_t cl_sym[kLit + kDist];    // the code-length sequence in code-length-code symbols
    uint8_t cl_extra[kLit + kDist];  // their extra-bit values
    int ncl, hlit, hdist, hclen;
    uint16_t hash[1 << kHashBits];
};

LP_DEF_FN int length_symbol(int len) {
    int s = 28;
    while (kLenBase[s] > len) s--;
    return s;
}
LP_DEF_FN int distance_symbol(int dist) {
    int s = 29;
    while (kDistBase[s] > dist) s--;
    return s;
}

// What a PngCompression level buys (zlib's own levels trade search effort the same way): how many earlier positions
// with the same 4-byte hash a position is compared against, and whether a match is deferred when the next position has a
// longer one (lazy evaluation).
LP_DEF_FN int level_chain(int level) { return level <= 2 ? 1 : level <= 5 ? 6 : level <= 7 ? 24 : 64; }
LP_DEF_FN int level_lazy(int level) { return level >= 4; }

LP_DEF_FN uint32_t hash4(const uint8_t* p) {
    const uint32_t v = p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24);
    return (v * 2654435761u) >> (32 - kHashBits);
}
// longest match for position i among the first `max_chain` entries of its hash chain (length >= 4 or 0)
LP_DEF_FN int longest_match(const uint8_t* src, int n, int i, int first, const uint16_t* prev, int max_chain, int* dist) {
    int best = 0;
    const int maxl = n - i < 258 ? n - i : 258;
    for (int cand = first, k = 0; cand != 0xFFFF && k < max_chain; cand = prev[cand], k++) {
        if (src[cand + best] != src[i + best]) continue;  // cannot beat the best so far
        int l = 0;
        while (l < maxl && src[cand + l] == src[i + l]) l++;
        if (l > best) {
            best = l;
            *dist = i - cand;
            if (l == maxl) break;
        }
    }
    return best >= 4 ? best : 0;
}

// LZ77 over one chunk: hash chains (head in `w.hash`, links in `prev`, one entry per position), greedy or lazy.
// Tokens: a literal = its byte value; a match = 0x8000 | (len - 3), then dist - 1.  Also fills the symbol statistics.
// Returns the number of uint16 units written.
LP_DEF_FN int tokenize(const uint8_t* src, int n, int level, Work& w, uint16_t* prev, uint16_t* tok) {
    for (int i = 0; i < (1 << kHashBits); i++) w.hash[i] = 0xFFFF;
    for (int i = 0; i < kLit; i++) w.freq_l[i] = 0;
    for (int i = 0; i < kDist; i++) w.freq_d[i] = 0;
    const int max_chain = level_chain(level), lazy = level_lazy(level);
    int nt = 0, i = 0;
    int inserted = 0;  // positions [0, inserted) are in the chains
    // pending match of the previous position (lazy evaluation)
    int have_prev = 0, prev_len = 0, prev_dist = 0;
    while (i < n) {
        int len = 0, dist = 0;
        if (i + 4 <= n) {
            const uint32_t h = hash4(src + i);
            len = longest_match(src, n, i, w.hash[h], prev, max_chain, &dist);
            if (inserted <= i) {
                prev[i] = w.hash[h];
                w.hash[h] = (uint16_t)i;
                inserted = i + 1;
            }
        }
        if (have_prev) {
            if (len > prev_len) {  // the later match wins: the previous position goes out as a literal
                tok[nt++] = src[i - 1];
                w.freq_l[src[i - 1]]++;
            } else {               // keep the earlier match (it started at i - 1)
                tok[nt++] = (uint16_t)(0x8000 | (prev_len - 3));
                tok[nt++] = (uint16_t)(prev_dist - 1);
                w.freq_l[257 + length_symbol(prev_len)]++;
                w.freq_d[distance_symbol(prev_dist)]++;
                const int end = i - 1 + prev_len;
                for (int j = inserted; j < end && j + 4 <= n; j++) {
                    const uint32_t hj = hash4(src + j);
                    prev[j] = w.hash[hj];
                    w.hash[hj] = (uint16_t)j;
                }
                if (end > inserted) inserted = end;
                i = end;
                have_prev = 0;
                continue;
            }
            have_prev = 0;
        }
        if (len) {
            if (lazy && len < 32 && i + 1 < n) {  // look one position ahead before committing
                have_prev = 1;
                prev_len = len;
                prev_dist = dist;
                i++;
                continue;
            }
            tok[nt++] = (uint16_t)(0x8000 | (len - 3));
            tok[nt++] = (uint16_t)(dist - 1);
            w.freq_l[257 + length_symbol(len)]++;
            w.freq_d[distance_symbol(dist)]++;
            const int end = i + len;
            for (int j = inserted; j < end && j + 4 <= n; j++) {
                const uint32_t hj = hash4(src + j);
                prev[j] = w.hash[hj];
                w.hash[hj] = (uint16_t)j;
            }
            if (end > inserted) inserted = end;
            i = end;
        } else {
            tok[nt++] = src[i];
            w.freq_l[src[i]]++;
            i++;
        }
    }
    if (have_prev) {  // (cannot happen: a deferred match is resolved at the next position, which exists)
        tok[nt++] = src[n - 1];
        w.freq_l[src[n - 1]]++;
    }
    w.freq_l[256]++;  // end of block
    return nt;
}

// Length-limited Huffman code lengths for `n` symbols from `freq` (two-queue merge over the frequency-sorted leaves;
// when the tree comes out deeper than `limit` the frequencies are halved -- kept >= 1 -- and the tree rebuilt: the
// flatter distribution converges to a balanced tree, which fits).  At least two symbols get a code, as zlib does, so
// the code is never empty or a lone zero-bit code.
LP_DEF_FN void build_lengths(const uint32_t* freq, int n, int limit, uint8_t* lens, Work& w) {
    // working copy of the leaf frequencies in node_freq[0, n): the caller's statistics stay exact
    int used = 0;
    for (int i = 0; i < n; i++) {
        w.node_freq[i] = freq[i];
        used += freq[i] != 0;
    }
    for (int i = 0; used < 2 && i < n; i++)
        if (!w.node_freq[i]) {
            w.node_freq[i] = 1;
            used++;
        }
    for (;;) {
        int m = 0;
        for (int i = 0; i < n; i++)
            if (w.node_freq[i]) {  // insertion sort by (frequency, symbol)
                int j = m++;
                while (j > 0 && w.node_freq[w.order[j - 1]] > w.node_freq[i]) {
                    w.order[j] = w.order[j - 1];
                    j--;
                }
                w.order[j] = (uint16_t)i;
            }
        // leaves are nodes [0, n) (by symbol), internal nodes [n, n + m - 1) in creation (= frequency) order
        int q1 = 0, q2 = n, next = n;
        const int last = n + m - 1;
        while (next < last) {
            int pick[2];
            for (int k = 0; k < 2; k++) {
                const bool leaf = q1 < m && (q2 >= next || w.node_freq[w.order[q1]] <= w.node_freq[q2]);
                pick[k] = leaf ? w.order[q1++] : q2++;
            }
            w.node_freq[next] = w.node_freq[pick[0]] + w.node_freq[pick[1]];
            w.node_parent[pick[0]] = (uint16_t)next;
            w.node_parent[pick[1]] = (uint16_t)next;
            next++;
        }
        const int root = last - 1;
        int deepest = 0;
        for (int i = 0; i < n; i++) {
            int d = 0;
            if (w.node_freq[i])
                for (int v = i; v != root; v = w.node_parent[v]) d++;
            lens[i] = (uint8_t)d;
            if (d > deepest) deepest = d;
        }
        if (deepest <= limit) return;
        for (int i = 0; i < n; i++)
            if (w.node_freq[i]) w.node_freq[i] = (w.node_freq[i] + 1) >> 1;
    }
}

// canonical codes (RFC 1951 3.2.2), stored bit-reversed
LP_DEF_FN void make_codes(const uint8_t* lens, int n, uint16_t* codes) {
    int bl_count[16], next_code[16];
    for (int b = 0; b < 16; b++) bl_count[b] = 0;
    for (int i = 0; i < n; i++) bl_count[lens[i]]++;
    bl_count[0] = 0;
    int code = 0;
    for (int b = 1; b < 16; b++) {
        code = (code + bl_count[b - 1]) << 1;
        next_code[b] = code;
    }
    for (int i = 0; i < n; i++)
        if (lens[i]) codes[i] = (uint16_t)bit_reverse((uint32_t)next_code[lens[i]]++, lens[i]);
        else codes[i] = 0;
}

// Builds the three codes of a dynamic block from the statistics in `w`; returns the size of the block in bits
// (header + symbols + extra bits).
LP_DEF_FN uint64_t prepare_dynamic(Work& w) {
    build_lengths(w.freq_l, kLit, 15, w.len_l, w);
    build_lengths(w.freq_d, kDist, 15, w.len_d, w);
    make_codes(w.len_l, kLit, w.code_l);
    make_codes(w.len_d, kDist, w.code_d);
    int hlit = kLit, hdist = kDist;
    while (hlit > 257 && w.len_l[hlit - 1] == 0) hlit--;
    while (hdist > 1 && w.len_d[hdist - 1] == 0) hdist--;
    w.hlit = hlit;
    w.hdist = hdist;
    // the code lengths as one sequence, run-length coded: 16 = repeat previous 3..6, 17 = 3..10 zeros, 18 = 11..138 zeros
    for (int i = 0; i < kCl; i++) w.freq_c[i] = 0;
    const int total = hlit + hdist;
    int ncl = 0, i = 0;
    while (i < total) {
        const int v = i < hlit ? w.len_l[i] : w.len_d[i - hlit];
        int run = 1;
        while (i + run < total && (i + run < hlit ? w.len_l[i + run] : w.len_d[i + run - hlit]) == v) run++;
        if (v == 0 && run >= 3) {
            const int r = run > 138 ? 138 : run;
            w.cl_sym[ncl] = r <= 10 ? 17 : 18;
            w.cl_extra[ncl] = (uint8_t)(r <= 10 ? r - 3 : r - 11);
            w.freq_c[w.cl_sym[ncl]]++;
            ncl++;
            i += r;
        } else if (v != 0 && run >= 4) {  // the value once, then "repeat previous"
            w.cl_sym[ncl] = (uint8_t)v;
            w.cl_extra[ncl] = 0;
            w.freq_c[v]++;
            ncl++;
            const int r = run - 1 > 6 ? 6 : run - 1;
            w.cl_sym[ncl] = 16;
            w.cl_extra[ncl] = (uint8_t)(r - 3);
            w.freq_c[16]++;
            ncl++;
            i += 1 + r;
        } else {
            w.cl_sym[ncl] = (uint8_t)v;
            w.cl_extra[ncl] = 0;
            w.freq_c[v]++;
            ncl++;
            i++;
        }
    }
    w.ncl = ncl;
    const uint32_t* fc = w.freq_c;
    build_lengths(w.freq_c, kCl, 7, w.len_c, w);
    make_codes(w.len_c, kCl, w.code_c);
    int hclen = kCl;
    while (hclen > 4 && w.len_c[kClOrder[hclen - 1]] == 0) hclen--;
    w.hclen = hclen;
    uint64_t bits = 3 + 5 + 5 + 4 + 3 * (uint64_t)hclen;
    for (int k = 0; k < kCl; k++) bits += (uint64_t)fc[k] * w.len_c[k];
    bits += 2 * (uint64_t)fc[16] + 3 * (uint64_t)fc[17] + 7 * (uint64_t)fc[18];
    for (int s = 0; s < kLit; s++) bits += (uint64_t)w.freq_l[s] * (w.len_l[s] + (s >= 257 ? kLenExtra[s - 257] : 0));
    for (int s = 0; s < kDist; s++) bits += (uint64_t)w.freq_d[s] * (w.len_d[s] + kDistExtra[s]);
    return bits;
}

LP_DEF_FN int fixed_litlen_bits(int s) { return s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8; }
LP_DEF_FN uint32_t fixed_litlen_code(int s) {  // RFC 1951 3.2.6, bit-reversed
    const uint32_t c = s < 144 ? 0x30 + s : s < 256 ? 0x190 + (s - 144) : s < 280 ? s - 256 : 0xC0 + (s - 280);
    return bit_reverse(c, fixed_litlen_bits(s));
}

// One chunk: `n` bytes of `src` -> a byte-aligned sequence of DEFLATE blocks (none of them final) that ends in an
// empty stored block (a sync flush), so chunks concatenate.  stored_only: level 0.  Returns the bytes written (<= kChunkOut).
LP_DEF_FN size_t write_chunk(const uint8_t* src, int n, int level, Work& w, uint16_t* prev, uint16_t* tok, uint8_t* dst) {
    bool stored = level == 0;
    BitOut b{dst, 0, 0};
    if (!stored) {
        const int nt = tokenize(src, n, level, w, prev, tok);
        uint64_t fixed_bits = 3;
        for (int s = 0; s < kLit; s++) fixed_bits += (uint64_t)w.freq_l[s] * (fixed_litlen_bits(s) + (s >= 257 ? kLenExtra[s - 257] : 0));
        for (int s = 0; s < kDist; s++) fixed_bits += (uint64_t)w.freq_d[s] * (5 + kDistExtra[s]);
        const uint64_t dyn_bits = prepare_dynamic(w);
        const bool dynamic = dyn_bits < fixed_bits;
        const uint64_t best = dynamic ? dyn_bits : fixed_bits;
        if (best + 7 + 32 > (uint64_t)n * 8 + 40) {
            stored = true;  // the data does not compress: a stored block is smaller
        } else {
            if (dynamic) {
                bo_put(b, 4, 3);  // BFINAL = 0, BTYPE = 10
                bo_put(b, (uint32_t)(w.hlit - 257), 5);
                bo_put(b, (uint32_t)(w.hdist - 1), 5);
                bo_put(b, (uint32_t)(w.hclen - 4), 4);
                for (int k = 0; k < w.hclen; k++) bo_put(b, w.len_c[kClOrder[k]], 3);
                for (int k = 0; k < w.ncl; k++) {
                    const int s = w.cl_sym[k];
                    bo_put(b, w.code_c[s], w.len_c[s]);
                    if (s == 16) bo_put(b, w.cl_extra[k], 2);
                    else if (s == 17) bo_put(b, w.cl_extra[k], 3);
                    else if (s == 18) bo_put(b, w.cl_extra[k], 7);
                }
            } else {
                bo_put(b, 2, 3);  // BFINAL = 0, BTYPE = 01
            }
            for (int k = 0; k < nt; k++) {
                const uint16_t t = tok[k];
                if (t < 0x8000) {
                    if (dynamic) bo_put(b, w.code_l[t], w.len_l[t]);
                    else bo_put(b, fixed_litlen_code(t), fixed_litlen_bits(t));
                    continue;
                }
                const int len = (t & 0x7FFF) + 3, dist = tok[++k] + 1;
                const int ls = length_symbol(len), ds = distance_symbol(dist);
                if (dynamic) bo_put(b, w.code_l[257 + ls], w.len_l[257 + ls]);
                else bo_put(b, fixed_litlen_code(257 + ls), fixed_litlen_bits(257 + ls));
                if (kLenExtra[ls]) bo_put(b, (uint32_t)(len - kLenBase[ls]), kLenExtra[ls]);
                if (dynamic) bo_put(b, w.code_d[ds], w.len_d[ds]);
                else bo_put(b, bit_reverse((uint32_t)ds, 5), 5);
                if (kDistExtra[ds]) bo_put(b, (uint32_t)(dist - kDistBase[ds]), kDistExtra[ds]);
            }
            if (dynamic) bo_put(b, w.code_l[256], w.len_l[256]);
            else bo_put(b, fixed_litlen_code(256), 7);
            bo_put(b, 0, 3);  // empty stored block = sync flush: byte-aligns the chunk
            bo_align(b);
            *b.p++ = 0x00;
            *b.p++ = 0x00;
            *b.p++ = 0xFF;
            *b.p++ = 0xFF;
        }
    }
    if (stored) {
        b = BitOut{dst, 0, 0};
        *b.p++ = 0x00;  // BFINAL = 0, BTYPE = 00, padding
        *b.p++ = (uint8_t)n;
        *b.p++ = (uint8_t)(n >> 8);
        *b.p++ = (uint8_t)~n;
        *b.p++ = (uint8_t)((~n) >> 8);
        for (int i = 0; i < n; i++) *b.p++ = src[i];
    }
    return (size_t)(b.p - dst);
}

}  // namespace defenc
