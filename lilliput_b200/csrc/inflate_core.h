// inflate_core.h -- warp-parallel DEFLATE (RFC 1951) decoding of ONE zlib stream by ONE warp.
//
// Replaces the bit-serial inflate that libpng/zlib-ng run for the reference's PNG decode
// (ref opencv.cpp:166-171 -> cv::PngDecoder::readData -> png_read_image -> inflate).  The result is the
// inflated byte string, identical to zlib's; only the schedule differs.
//
// A DEFLATE block is one serial bit string, but its prefix code self-synchronises: a decoder started at
// a wrong bit falls onto true symbol boundaries within a few symbols, and -- unlike JPEG -- the decoder
// has no other state than the bit position.  So, per "window" of 32 x kSubBits bits of one block:
//   pass A   lane i decodes the symbols that START in [P + i*S, P + (i+1)*S) from a guessed start
//            (lane 0's start is exact), counting output bytes and matches;
//   pass B   lane i re-decodes from lane i-1's exit position whenever that differs from the start it
//            used, until nothing changes (exact by induction from lane 0; lanes behind an end-of-block
//            or an invalid code are ignored);
//   pass C   a prefix sum of the byte counts gives every lane its output position; literals are written
//            to a shared-memory ring in parallel, matches are listed (in stream order, by a second
//            prefix sum) and then copied one after another by the whole warp.
// The ring is flushed to global memory in 16-byte vectors; matches that reach behind the ring read the
// flushed bytes back.  Windows whose output would overrun half the ring are cut short (the write pass
// takes a byte budget), so match-heavy streams degrade to "few symbols per window", never to an error.
//
// The same source compiles for the device (one warp = 32 threads, collectives = shuffles / ballots) and
// for the host (LP_INF_HOST: lanes simulated by loops), so the CPU test-suite runs the exact control
// flow of the kernel against zlib (tests/test_inflate_core.py); the GPU tests then only have to show
// that the real warp reproduces the simulation.
#pragma once
#include <stdint.h>
#include <string.h>

#ifdef LP_INF_HOST
#define LP_INF_FN static inline
#define LP_INF_LANES(l) for (int l = 0; l < 32; l++)
#else
#define LP_INF_FN static __device__ __forceinline__
#define LP_INF_LANES(l) for (int l = (int)(threadIdx.x & 31), _lp_once = 0; !_lp_once; _lp_once = 1)
#endif

namespace lpinf {

#ifndef LP_INF_SUB
#define LP_INF_SUB 320
#endif
#ifndef LP_INF_WARM
#define LP_INF_WARM 64
#endif
constexpr uint32_t kSubBits = LP_INF_SUB;      // bits per lane and window
constexpr uint32_t kWinBits = 32 * kSubBits;
constexpr uint32_t kInWords = kWinBits / 32 + 8;  // window + the bits a symbol / a refill may read past it
#ifndef LP_INF_RING
#define LP_INF_RING 4096
#endif
constexpr uint32_t kRing = LP_INF_RING, kRingMask = kRing - 1;
constexpr uint32_t kCapT = kRing / 2;          // output bytes per window
constexpr uint32_t kMaxMatches = kCapT / 3 + 8;  // a match is at least 3 bytes
constexpr int kLitBits = 10, kDistBits = 8;
constexpr uint32_t kCkBits = 64;                       // a checkpoint every kCkBits bits of a subsequence
constexpr int kCk = (int)(kSubBits / kCkBits);         // slots 1 .. kCk-1 are used (slot 0 would be the entry itself)

// lane-private variable: one register on the device, 32 slots in the host simulation
#ifdef LP_INF_HOST
template <class T>
struct LaneVar {
    T v[32];
    T& operator[](int l) { return v[l]; }
    const T& operator[](int l) const { return v[l]; }
};
#else
template <class T>
struct LaneVar {
    T v;
    __device__ __forceinline__ T& operator[](int) { return v; }
    __device__ __forceinline__ const T& operator[](int) const { return v; }
};
#endif

// ---- warp collectives -------------------------------------------------------------------------------
LP_INF_FN void wsync() {
#ifndef LP_INF_HOST
    __syncwarp();
#endif
}
LP_INF_FN void shift_up(LaneVar<uint32_t>& out, const LaneVar<uint32_t>& in, uint32_t lane0) {
#ifdef LP_INF_HOST
    uint32_t prev = lane0;
    for (int l = 0; l < 32; l++) {
        const uint32_t cur = in[l];
        out[l] = prev;
        prev = cur;
    }
#else
    const uint32_t t = __shfl_up_sync(0xffffffffu, in.v, 1);
    out.v = (threadIdx.x & 31) ? t : lane0;
#endif
}
LP_INF_FN uint32_t ballot(const LaneVar<uint32_t>& p) {
#ifdef LP_INF_HOST
    uint32_t m = 0;
    for (int l = 0; l < 32; l++) m |= (p[l] ? 1u : 0u) << l;
    return m;
#else
    return __ballot_sync(0xffffffffu, p.v != 0);
#endif
}
LP_INF_FN uint32_t excl_scan(LaneVar<uint32_t>& out, const LaneVar<uint32_t>& in) {  // returns the total
#ifdef LP_INF_HOST
    uint32_t s = 0;
    for (int l = 0; l < 32; l++) {
        const uint32_t v = in[l];
        out[l] = s;
        s += v;
    }
    return s;
#else
    const int lane = threadIdx.x & 31;
    uint32_t inc = in.v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    out.v = inc - in.v;
    return __shfl_sync(0xffffffffu, inc, 31);
#endif
}
LP_INF_FN uint32_t bcast(const LaneVar<uint32_t>& v, int src) {
#ifdef LP_INF_HOST
    return v[src];
#else
    return __shfl_sync(0xffffffffu, v.v, src);
#endif
}
// lanes holding the same value as this lane (bit mask)
LP_INF_FN void match_any(LaneVar<uint32_t>& out, const LaneVar<uint32_t>& in) {
#ifdef LP_INF_HOST
    for (int l = 0; l < 32; l++) {
        uint32_t m = 0;
        for (int k = 0; k < 32; k++) m |= (in[k] == in[l] ? 1u : 0u) << k;
        out[l] = m;
    }
#else
    out.v = __match_any_sync(0xffffffffu, in.v);
#endif
}
LP_INF_FN uint32_t popc32(uint32_t x) {
#ifdef LP_INF_HOST
    return (uint32_t)__builtin_popcount(x);
#else
    return (uint32_t)__popc(x);
#endif
}
LP_INF_FN uint32_t ffs32(uint32_t x) {  // 1-based, 0 for x == 0
#ifdef LP_INF_HOST
    return (uint32_t)__builtin_ffs((int)x);
#else
    return (uint32_t)__ffs((int)x);
#endif
}
LP_INF_FN uint32_t brev32(uint32_t x) {
#ifdef LP_INF_HOST
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    return (x >> 16) | (x << 16);
#else
    return __brev(x);
#endif
}

// ---- per-warp working set (shared memory on the device) -----------------------------------------------
struct Match {
    uint32_t q;   // absolute output position of the copy's first byte
    uint32_t ld;  // (length << 16) | distance
};

// Lookup entries (32 bit):  bits 0..3 code length (0 = not a code of <= lookahead bits)
//   literal/length table: bits 4..5 kind (0 literal, 1 end of block, 2 length, 3 invalid symbol 286/287),
//                         bits 8..16 literal value or length base, bits 20..22 extra-bit count
//   distance table:       bits 4..5 kind (0 distance, 3 invalid symbol 30/31), bits 8..22 base, bits 24..27 extra bits
struct WarpShared {
    alignas(16) uint8_t ring[kRing];
    uint32_t inbuf[kInWords];
    uint32_t lit[1 << kLitBits];
    uint32_t dist[1 << kDistBits];
    uint32_t cnt32[2][16];     // codes per length (lit, dist)
    uint16_t first[2][17];     // canonical first code per length, for the long-code walk
    uint16_t index[2][17];     // first sorted-symbol index per length
    uint16_t lsym[288], dsym[32];
    uint8_t lens[320];
    uint16_t cl[128];          // code-length code lookahead (7 bits): (symbol << 4) | length
    // per lane and checkpoint j: the first symbol start at or behind (nominal start + j * kCkBits) of the lane's
    // latest decode, with the matches / bytes decoded in front of it
    uint32_t ck_pos_nm[32][kCk];  // (position << 16) | matches; position 0xFFFF = none
    uint32_t ck_cnt[32][kCk];
    // the matches of the window being written, in stream order, compact: (position in the window << 20) |
    // (min(length - 3, 127) << 13) | distance; length field 127 = look the full record up in Stream::mlist
    // (lengths >= 130 and distances >= 8192).  Reading them back from global memory cost ~300 cycles per match.
    uint32_t mrec[kMaxMatches];
    uint16_t near[kMaxMatches];  // indices (into mrec) of the matches that depend on window bytes, in stream order
};

LP_INF_FN uint32_t lit_entry(uint32_t sym, uint32_t len) {
    if (sym < 256) return len | (sym << 8);
    if (sym == 256) return len | (1u << 4);
    if (sym > 285) return len | (3u << 4);
    const uint32_t ls = sym - 257;
    uint32_t extra, base;
    if (ls < 8) { extra = 0; base = 3 + ls; }
    else if (ls == 28) { extra = 0; base = 258; }
    else { extra = (ls - 4) >> 2; base = 3 + ((4 + (ls & 3)) << extra); }
    return len | (2u << 4) | (base << 8) | (extra << 20);
}
LP_INF_FN uint32_t dist_entry(uint32_t sym, uint32_t len) {
    if (sym > 29) return len | (3u << 4);
    uint32_t extra, base;
    if (sym < 4) { extra = 0; base = 1 + sym; }
    else { extra = (sym - 2) >> 1; base = 1 + ((2 + (sym & 1)) << extra); }
    return len | (base << 8) | (extra << 24);
}

// Canonical code (RFC 1951 3.2.2) of `n` code lengths -> lookahead table + long-code walk data.
// t = 0: literal/length, t = 1: distance.  Returns 0, or 1 for an over-subscribed code (zlib: "invalid
// code lengths set" / "invalid distances set"; incomplete codes are accepted like the earlier kernel did).
LP_INF_FN int build_table(WarpShared& ws, int t, const uint8_t* lens, int n) {
    uint32_t* look = t ? ws.dist : ws.lit;
    uint16_t* sorted = t ? ws.dsym : ws.lsym;
    const int look_bits = t ? kDistBits : kLitBits;
    LP_INF_LANES(l) {
        if (l < 16) ws.cnt32[t][l] = 0;
    }
    wsync();
    // codes per length: lane l owns length l (n <= 288)
    LP_INF_LANES(l) {
        if (l >= 1 && l < 16) {
            uint32_t c = 0;
            for (int i = 0; i < n; i++) c += lens[i] == l;
            ws.cnt32[t][l] = c;
        }
    }
    wsync();
    int over = 0;
    LP_INF_LANES(l) {
        if (l == 0) {
            int left = 1;
            uint32_t first = 0, index = 0;
            for (int len = 1; len < 16; len++) {
                left = (left << 1) - (int)ws.cnt32[t][len];
                if (left < 0) over = 1;
                ws.first[t][len] = (uint16_t)first;
                ws.index[t][len] = (uint16_t)index;
                index += ws.cnt32[t][len];
                first = (first + ws.cnt32[t][len]) << 1;
            }
            ws.first[t][16] = (uint16_t)over;  // broadcast slot
        }
    }
    wsync();
    over = ws.first[t][16];
    if (over) return 1;
    // symbols sorted by (length, symbol): lane l places the symbols of length l+1 ... spread by length so
    // that no two lanes write the same region (15 lengths, <= 288 symbols)
    LP_INF_LANES(l) {
        if (l >= 1 && l < 16) {
            uint32_t at = ws.index[t][l];
            for (int i = 0; i < n; i++)
                if (lens[i] == l) sorted[at++] = (uint16_t)i;
        }
    }
    wsync();
    // lookahead entries, one canonical walk per entry (entries are spread over the lanes)
    LP_INF_LANES(l) {
        for (uint32_t j = (uint32_t)l; j < (1u << look_bits); j += 32) {
            uint32_t code = 0, e = 0;
            for (int len = 1; len <= look_bits; len++) {
                code |= (j >> (len - 1)) & 1u;
                const uint32_t c = ws.cnt32[t][len], f = ws.first[t][len];
                if (code - f < c) {  // (unsigned: code >= f always holds on a valid walk)
                    const uint32_t sym = sorted[ws.index[t][len] + (code - f)];
                    e = t ? dist_entry(sym, (uint32_t)len) : lit_entry(sym, (uint32_t)len);
                    break;
                }
                code <<= 1;
            }
            look[j] = e;
        }
    }
    wsync();
    return 0;
}

// LSB-first bit reader over the window buffer: `pos` counts bits from bit 0 of inbuf[0].
struct Bits {
    uint64_t acc;
    uint32_t n;    // valid bits in acc
    uint32_t wp;   // next word of inbuf
    uint32_t pos;  // bit position of acc's bit 0
};
LP_INF_FN void bits_init(Bits& b, const uint32_t* inbuf, uint32_t pos) {
    b.wp = pos >> 5;
    b.acc = (uint64_t)inbuf[b.wp] >> (pos & 31);
    b.n = 32 - (pos & 31);
    b.wp++;
    b.pos = pos;
}
LP_INF_FN void bits_fill(Bits& b, const uint32_t* inbuf) {  // afterwards n >= 32
    if (b.n < 32) {
        b.acc |= (uint64_t)inbuf[b.wp < kInWords ? b.wp : kInWords - 1] << b.n;
        b.wp++;
        b.n += 32;
    }
}
LP_INF_FN void bits_drop(Bits& b, uint32_t k) {
    b.acc >>= k;
    b.n -= k;
    b.pos += k;
}
LP_INF_FN uint32_t bits_get(Bits& b, const uint32_t* inbuf, uint32_t k) {  // k <= 16
    bits_fill(b, inbuf);
    const uint32_t v = (uint32_t)b.acc & ((1u << k) - 1u);
    bits_drop(b, k);
    return v;
}

// 64 stream bits from bit position `pos` of the window buffer as two 32-bit halves: three word loads and two
// funnel shifts, no state carried between symbols (every field of a symbol is then a shift + mask of lo / hi).
LP_INF_FN uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) {  // bits [sh, sh + 32) of hi:lo, sh < 32
#ifdef LP_INF_HOST
    return (uint32_t)((((uint64_t)hi << 32) | lo) >> sh);
#else
    return __funnelshift_r(lo, hi, sh);
#endif
}

// oldest output position the ring still holds once a window that starts at o and produces wT bytes is written
LP_INF_FN uint32_t ring_lo_d0(uint32_t o, uint32_t wT) { return o + wT > kRing ? o + wT - kRing : 0; }

// A code longer than the lookahead of table t in `bits` (>= 15 valid bits): the canonical walk continues behind
// the lookahead (a code of <= look_bits bits would have been in the table), at most 15 - look_bits steps.
// Returns the entry (code length in bits 0..3) or 0 when the bits are no code word.
LP_INF_FN uint32_t long_code(const WarpShared& ws, int t, uint32_t bits) {
    const int look_bits = t ? kDistBits : kLitBits;
    uint32_t code = brev32(bits << (32 - look_bits));  // the first look_bits bits, most significant first
    for (int len = look_bits + 1; len < 16; len++) {
        code = (code << 1) | ((bits >> (len - 1)) & 1u);
        const uint32_t c = ws.cnt32[t][len], f = ws.first[t][len];
        if (code - f < c) {
            const uint32_t sym = (t ? ws.dsym : ws.lsym)[ws.index[t][len] + (code - f)];
            return t ? dist_entry(sym, (uint32_t)len) : lit_entry(sym, (uint32_t)len);
        }
    }
    return 0;
}

struct Span {
    uint32_t exit;  // bit position (window-relative) of the first symbol not consumed
    uint32_t cnt;   // output bytes
    uint32_t nm;    // matches
    uint32_t flag;  // 0, 1 = stopped behind an end-of-block symbol, 2 = invalid data
};
enum { kFlagNone = 0, kFlagEob = 1, kFlagBad = 2 };

// Decode the symbols that start in [start, end).  WRITE: literals go to the ring at absolute output
// position q, matches are appended to mlist; stops in front of the first symbol that would take the
// output past `budget` bytes (exit then points at that symbol, flag stays 0).
// REC (count passes): at every checkpoint boundary the lane crosses, the first symbol start behind it and the
// counts in front of it are recorded.  CMP (re-decodes of pass B): if the previous decode of this lane recorded the
// SAME position at a checkpoint, the two decodes have merged -- everything behind is identical -- so the decode
// stops there and the totals follow by arithmetic (r.exit / r.flag are then left untouched by the caller).
// Returns true when it merged.  ck slots are addressed relative to `nominal` (the subsequence start).
template <bool WRITE, bool REC, bool CMP>
LP_INF_FN bool decode_span(WarpShared& ws, uint32_t start, uint32_t end, uint32_t q, uint32_t budget,
                           Match* mlist, Span& r, uint32_t nominal, int lane, uint32_t old_cnt, uint32_t old_nm,
                           uint32_t* mrec = nullptr, uint32_t obase = 0) {
    uint32_t pos = start, cnt = 0, nm = 0, flag = kFlagNone;
    uint32_t cw = 0xFFFFFFF0u, w0 = 0, w1 = 0, w2 = 0;  // window words in registers (cw = index of w0)
    uint32_t next_ck = nominal + kCkBits;
    int j = 0;
    while (pos < end) {
        if (REC && pos >= next_ck) {
            while (pos >= next_ck) {
                j++;
                next_ck += kCkBits;
            }
            if (j < kCk) {
                const uint32_t rel16 = (pos - nominal) & 0xFFFFu;
                if (CMP) {
                    const uint32_t o_pn = ws.ck_pos_nm[lane][j], o_c = ws.ck_cnt[lane][j];
                    if ((o_pn >> 16) == rel16) {
                        // merged: totals = this decode up to here + the old decode from here on
                        const uint32_t dc = cnt - o_c, dn = nm - (o_pn & 0xFFFFu);
                        for (int jj = j; jj < kCk; jj++) {
                            const uint32_t pn = ws.ck_pos_nm[lane][jj];
                            if ((pn >> 16) == 0xFFFFu) continue;
                            ws.ck_pos_nm[lane][jj] = (pn & 0xFFFF0000u) | ((pn + dn) & 0xFFFFu);
                            ws.ck_cnt[lane][jj] += dc;
                        }
                        r.cnt = old_cnt + dc;
                        r.nm = old_nm + dn;
                        r.exit = (uint32_t)j;  // (statistics only: the caller keeps its old exit)
                        return true;
                    }
                }
                ws.ck_pos_nm[lane][j] = (rel16 << 16) | (nm & 0xFFFFu);
                ws.ck_cnt[lane][j] = cnt;
            }
        }
        // the three window words the symbol can touch stay in registers; a symbol is 1..48 bits, so they slide by at
        // most two words (shared-memory requests, not instructions, are what limits this kernel when many streams
        // share an SM: ~0.5 loads per symbol here instead of 3)
        const uint32_t wi = pos >> 5, sh = pos & 31u;
        if (wi != cw) {
            if (wi == cw + 1) {
                w0 = w1;
                w1 = w2;
                w2 = ws.inbuf[wi + 2];
            } else {
                w0 = ws.inbuf[wi];
                w1 = ws.inbuf[wi + 1];
                w2 = ws.inbuf[wi + 2];
            }
            cw = wi;
        }
        const uint32_t lo = funnel_r(w0, w1, sh), hi = funnel_r(w1, w2, sh);
        uint32_t e = ws.lit[lo & ((1u << kLitBits) - 1u)];
        if (!(e & 15u)) e = long_code(ws, 0, lo);
        const uint32_t kind = (e >> 4) & 3u;
        uint32_t used = e & 15u;
        if (e == 0 || kind == 3) {
            flag = kFlagBad;
            break;
        }
        if (kind == 0) {
            if (WRITE) {
                if (cnt + 1 > budget) break;
                ws.ring[(q + cnt) & kRingMask] = (uint8_t)(e >> 8);
            }
            cnt++;
            pos += used;
            continue;
        }
        if (kind == 1) {
            flag = kFlagEob;
            pos += used;
            break;
        }
        // length (+ extra bits), then distance code (+ extra bits): at most 15 + 5 + 15 + 13 = 48 bits of lo / hi
        const uint32_t lx = (e >> 20) & 7u;
        const uint32_t len = ((e >> 8) & 0x1FFu) + (funnel_r(lo, hi, used) & ((1u << lx) - 1u));
        used += lx;  // <= 20
        const uint32_t rest = funnel_r(lo, hi, used);  // 32 bits behind the length: the distance code and its extra bits
        uint32_t d = ws.dist[rest & ((1u << kDistBits) - 1u)];
        if (!(d & 15u)) d = long_code(ws, 1, rest);
        if (d == 0 || ((d >> 4) & 3u) == 3) {
            flag = kFlagBad;
            break;
        }
        const uint32_t dl = d & 15u, dx = (d >> 24) & 15u;
        const uint32_t dist = ((d >> 8) & 0x7FFFu) + ((rest >> dl) & ((1u << dx) - 1u));
        used += dl + dx;
        if (WRITE) {
            if (cnt + len > budget) break;
            if (dist > q + cnt) {  // zlib: "invalid distance too far back"
                flag = kFlagBad;
                break;
            }
            mlist[nm].q = q + cnt;
            mlist[nm].ld = (len << 16) | dist;
            {
                const bool esc = len - 3 >= 127u || dist >= 8192u;
                mrec[nm] = ((q + cnt - obase) << 20) | ((esc ? 127u : len - 3) << 13) | (esc ? 0u : dist);
            }
        }
        cnt += len;
        nm++;
        pos += used;
    }
    if (REC)
        for (int jj = j + 1; jj < kCk; jj++) ws.ck_pos_nm[lane][jj] = 0xFFFF0000u;  // not reached by this decode
    r.exit = pos;
    r.cnt = cnt;
    r.nm = nm;
    r.flag = flag;
    return false;
}

// ---- the stream ------------------------------------------------------------------------------------------

// Optional counters (LP_INF_STATS): the host simulation counts windows / rounds / re-decodes, the device
// build (png_decode.cu with -DLP_INF_STATS) adds clock64() per phase.  [0] blocks [1] windows [2] rounds
// [3] lane re-decodes in pass B [4] partial windows [5] matches; device clocks: [8] header+tables
// [9] load_window [10] pass A [11] pass B [12] pass C [13] match copy [14] flush
#ifdef LP_INF_STATS
#ifdef LP_INF_HOST
static unsigned long long g_stats[16];
#define LP_INF_COUNT(i, v) (g_stats[i] += (v))
#define LP_INF_CLOCK(i) do { } while (0)
#else
__device__ unsigned long long g_stats[16];
#define LP_INF_COUNT(i, v) do { if ((threadIdx.x & 31) == 0) atomicAdd(&g_stats[i], (unsigned long long)(v)); } while (0)
#define LP_INF_CLOCK(i) do { const long long _now = clock64(); if ((threadIdx.x & 31) == 0) atomicAdd(&g_stats[i], (unsigned long long)(_now - _t0)); _t0 = _now; } while (0)
#endif
#else
#define LP_INF_COUNT(i, v) do { } while (0)
#define LP_INF_CLOCK(i) do { } while (0)
#endif

struct Stream {
    const uint8_t* z;   // zlib stream (global memory)
    uint32_t z_len;
    uint8_t* out;       // inflated bytes (global memory)
    uint32_t cap;       // bytes expected
    Match* mlist;       // kMaxMatches entries (global scratch of this warp)
};

// inbuf <- the window that starts at absolute bit P; returns the window-relative bit offset of P (0..31+).
LP_INF_FN uint32_t load_window(WarpShared& ws, const Stream& s, uint64_t P) {
    const uint64_t byte0 = P >> 3;
    const uintptr_t a = reinterpret_cast<uintptr_t>(s.z) + (uintptr_t)byte0;
    const uintptr_t a0 = a & ~(uintptr_t)3;
    const uintptr_t zend = reinterpret_cast<uintptr_t>(s.z) + s.z_len;
    LP_INF_LANES(l) {
        for (uint32_t w = (uint32_t)l; w < kInWords; w += 32) {
            const uintptr_t wa = a0 + 4u * (uintptr_t)w;
            uint32_t v = 0;
            if (wa + 4 <= zend && wa >= reinterpret_cast<uintptr_t>(s.z)) {
                v = *reinterpret_cast<const uint32_t*>(wa);
            } else {
                for (int k = 0; k < 4; k++) {
                    const uintptr_t ba = wa + (uintptr_t)k;
                    if (ba >= reinterpret_cast<uintptr_t>(s.z) && ba < zend) v |= (uint32_t)(*reinterpret_cast<const uint8_t*>(ba)) << (8 * k);
                }
            }
            ws.inbuf[w] = v;
        }
    }
    wsync();
    return (uint32_t)((a - a0) * 8 + (P & 7));
}

// ring -> out for [flushed, upto), both multiples of 16 when the vector path is usable
LP_INF_FN void flush_ring(WarpShared& ws, const Stream& s, uint32_t flushed, uint32_t upto, bool vec) {
    if (vec) {
        LP_INF_LANES(l) {
            for (uint32_t p = flushed + 16u * (uint32_t)l; p < upto; p += 16u * 32u) {
#ifdef LP_INF_HOST
                memcpy(s.out + p, ws.ring + (p & kRingMask), 16);
#else
                *reinterpret_cast<uint4*>(s.out + p) = *reinterpret_cast<const uint4*>(ws.ring + (p & kRingMask));
#endif
            }
        }
    } else {
        LP_INF_LANES(l) {
            for (uint32_t p = flushed + (uint32_t)l; p < upto; p += 32u) s.out[p] = ws.ring[p & kRingMask];
        }
    }
    wsync();
}

// Whole stream.  Returns 0 or -3 (corrupt); *produced = bytes written to s.out.
LP_INF_FN int inflate_stream(WarpShared& ws, const Stream& s, uint32_t* produced) {
    uint32_t o = 0, flushed = 0;
    *produced = 0;
    if (s.z_len < 2) return -3;
    const bool vec = (reinterpret_cast<uintptr_t>(s.out) & 15) == 0;
    {
        uint32_t h0 = 0, h1 = 0;
        h0 = s.z[0];
        h1 = s.z[1];
        if ((h0 & 15u) != 8 || (h1 & 0x20u)) return -3;  // not deflate / preset dictionary
    }
    uint64_t P = 16;  // absolute bit position in the stream
    const uint64_t total_bits = (uint64_t)s.z_len * 8;
    int last = 0;
#if defined(LP_INF_STATS) && !defined(LP_INF_HOST)
    long long _t0 = clock64();
#endif
    while (!last) {
        if (P + 3 > total_bits) return -3;
        LP_INF_COUNT(0, 1);
        uint32_t rel = load_window(ws, s, P);
        LP_INF_CLOCK(9);
        // ---- block header (lane 0 reads, the warp learns the result through shared memory)
        uint32_t type = 0, hdr_err = 0, hdr_bits = 0;
        {
            LaneVar<uint32_t> a, bb, cc, dd;
            LP_INF_LANES(l) {
                a[l] = bb[l] = cc[l] = dd[l] = 0;
                if (l == 0) {
                    Bits b;
                    bits_init(b, ws.inbuf, rel);
                    const uint32_t lastv = bits_get(b, ws.inbuf, 1);
                    const uint32_t ty = bits_get(b, ws.inbuf, 2);
                    uint32_t err = 0;
                    if (ty == 1) {
                        int i = 0;
                        for (; i < 144; i++) ws.lens[i] = 8;
                        for (; i < 256; i++) ws.lens[i] = 9;
                        for (; i < 280; i++) ws.lens[i] = 7;
                        for (; i < 288; i++) ws.lens[i] = 8;
                        for (i = 0; i < 32; i++) ws.lens[288 + i] = 5;
                        cc[l] = 288 | (32u << 16);  // fixed code: 288 literal/length + 32 distance symbols (30, 31 invalid)
                    } else if (ty == 2) {
                        const uint32_t nl = bits_get(b, ws.inbuf, 5) + 257, nd = bits_get(b, ws.inbuf, 5) + 1;
                        const uint32_t nc = bits_get(b, ws.inbuf, 4) + 4;
                        if (nl > 286 || nd > 30) err = 1;
                        uint8_t cl[19];
                        for (int i = 0; i < 19; i++) cl[i] = 0;
                        const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
                        for (uint32_t i = 0; i < nc && !err; i++) cl[order[i]] = (uint8_t)bits_get(b, ws.inbuf, 3);
                        // code-length code: 7-bit lookahead (codes are at most 7 bits)
                        uint32_t count[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                        for (int i = 0; i < 19; i++) count[cl[i]]++;
                        count[0] = 0;
                        int left = 1;
                        for (int len = 1; len < 8; len++) {
                            left = (left << 1) - (int)count[len];
                            if (left < 0) err = 1;
                        }
                        for (int i = 0; i < 128; i++) ws.cl[i] = 0;
                        if (!err) {
                            uint32_t code = 0;
                            for (int len = 1; len < 8; len++) {
                                for (int sym = 0; sym < 19; sym++) {
                                    if (cl[sym] != len) continue;
                                    const uint32_t rev = brev32(code) >> (32 - len);
                                    for (uint32_t j = rev; j < 128; j += 1u << len) ws.cl[j] = (uint16_t)((sym << 4) | len);
                                    code++;
                                }
                                code <<= 1;
                            }
                        }
                        uint32_t i = 0;
                        while (!err && i < nl + nd) {
                            bits_fill(b, ws.inbuf);
                            const uint32_t e = ws.cl[(uint32_t)b.acc & 127u];
                            if (!e) { err = 1; break; }
                            bits_drop(b, e & 15u);
                            const uint32_t sy = e >> 4;
                            if (sy < 16) { ws.lens[i++] = (uint8_t)sy; continue; }
                            uint32_t rep, v = 0;
                            if (sy == 16) {
                                if (!i) { err = 1; break; }
                                v = ws.lens[i - 1];
                                rep = 3 + bits_get(b, ws.inbuf, 2);
                            } else if (sy == 17) rep = 3 + bits_get(b, ws.inbuf, 3);
                            else rep = 11 + bits_get(b, ws.inbuf, 7);
                            if (i + rep > nl + nd) { err = 1; break; }
                            while (rep--) ws.lens[i++] = (uint8_t)v;
                        }
                        if (!err && ws.lens[256] == 0) err = 1;  // zlib: "missing end-of-block"
                        cc[l] = nl | (nd << 16);
                        if (b.pos > kInWords * 32 - 64) err = 1;  // header ran out of the window (cannot happen: <= ~4600 bits)
                    } else if (ty == 3) {
                        err = 1;
                    }
                    a[l] = lastv | (ty << 1) | (err << 3);
                    bb[l] = b.pos - rel;
                }
            }
            const uint32_t av = bcast(a, 0);
            last = (int)(av & 1u);
            type = (av >> 1) & 3u;
            hdr_err = av >> 3;
            hdr_bits = bcast(bb, 0);
            const uint32_t nn = bcast(cc, 0);
            wsync();
            if (hdr_err) return -3;
            P += hdr_bits;
            if (type != 0) {
                const int nl = (int)(nn & 0xFFFF), nd = (int)(nn >> 16);
                if (build_table(ws, 0, ws.lens, nl)) return -3;
                if (build_table(ws, 1, ws.lens + nl, nd)) return -3;
            }
        }
        LP_INF_CLOCK(8);
        if (type == 0) {
            // ---- stored block: LEN / NLEN at the next byte boundary, then raw bytes
            const uint64_t pb = (P + 7) >> 3;
            if (pb + 4 > s.z_len) return -3;
            const uint32_t len = (uint32_t)s.z[pb] | ((uint32_t)s.z[pb + 1] << 8);
            const uint32_t nlen = (uint32_t)s.z[pb + 2] | ((uint32_t)s.z[pb + 3] << 8);
            if ((len ^ 0xFFFFu) != nlen || pb + 4 + len > s.z_len || len > s.cap - o) return -3;
            const uint8_t* src = s.z + pb + 4;
            uint32_t done = 0;
            while (done < len) {
                const uint32_t n = len - done < kCapT ? len - done : kCapT;
                LP_INF_LANES(l) {
                    for (uint32_t i = (uint32_t)l; i < n; i += 32) ws.ring[(o + i) & kRingMask] = src[done + i];
                }
                wsync();
                o += n;
                done += n;
                const uint32_t upto = vec ? (o & ~15u) : o;
                flush_ring(ws, s, flushed, upto, vec);
                flushed = upto;
            }
            P = (pb + 4 + len) * 8;
            continue;
        }
        // ---- compressed block: windows of 32 subsequences until the end-of-block symbol
        bool eob = false;
        while (!eob) {
            if (P >= total_bits) return -3;  // the block runs past the end of the stream
            LP_INF_COUNT(1, 1);
            rel = load_window(ws, s, P);
            LP_INF_CLOCK(9);
            LaneVar<uint32_t> start, exitp, cnt, nm, flag, want, changed, term;
            Span sp;
            // pass A: every lane from the start of its subsequence (exact for lane 0 only), recording checkpoints
            LP_INF_LANES(l) {
                const uint32_t nominal = rel + (uint32_t)l * kSubBits;
                start[l] = nominal;
                decode_span<false, true, false>(ws, nominal, nominal + kSubBits, 0, 0, nullptr, sp, nominal, l, 0, 0);
                exitp[l] = sp.exit; cnt[l] = sp.cnt; nm[l] = sp.nm; flag[l] = sp.flag;
            }
            LP_INF_CLOCK(10);
            // pass B: fixed point.  A lane whose left neighbour's exit differs from the entry it used decodes again
            // from there -- but only until it meets a checkpoint of its previous decode (usually the first or second)
            for (int round = 0; round < 34; round++) {
                shift_up(want, exitp, rel);
                LP_INF_LANES(l) { term[l] = flag[l] != kFlagNone; }
                const uint32_t tm = ballot(term);
                LP_INF_LANES(l) {
                    const bool dead = (tm & ((1u << l) - 1u)) != 0;
                    changed[l] = (!dead && want[l] != start[l]) ? 1u : 0u;
                }
                if (!ballot(changed)) break;
                LP_INF_COUNT(2, 1);
#if defined(LP_INF_STATS) && defined(LP_INF_HOST)
                uint32_t round_len = 0;
#endif
                LP_INF_LANES(l) {
                    if (changed[l]) {
                        const uint32_t nominal = rel + (uint32_t)l * kSubBits;
                        start[l] = want[l];
                        if (want[l] >= nominal + kSubBits) {  // the neighbour's last symbol covers this whole subsequence
                            exitp[l] = want[l]; cnt[l] = 0; nm[l] = 0; flag[l] = kFlagNone;
                            for (int jj = 1; jj < kCk; jj++) ws.ck_pos_nm[l][jj] = 0xFFFF0000u;
                        } else {
                            sp.cnt = sp.nm = 0;
                            const bool merged = decode_span<false, true, true>(ws, want[l], nominal + kSubBits, 0, 0, nullptr, sp,
                                                                               nominal, l, cnt[l], nm[l]);
                            cnt[l] = sp.cnt; nm[l] = sp.nm;
#if defined(LP_INF_STATS) && defined(LP_INF_HOST)
                            { const uint32_t len_ = merged ? sp.exit : (uint32_t)kCk; if (len_ > round_len) round_len = len_; }
#endif
                            if (!merged) {
                                LP_INF_COUNT(3, 1);
                                exitp[l] = sp.exit; flag[l] = sp.flag;
                            }
                        }
                    }
                }
#if defined(LP_INF_STATS) && defined(LP_INF_HOST)
                g_stats[7] += round_len;
#endif
            }
            LP_INF_CLOCK(11);
            LP_INF_LANES(l) { term[l] = flag[l] != kFlagNone; }
            const uint32_t tm = ballot(term);
            const int k = tm ? (int)ffs32(tm) - 1 : 32;  // first lane that ends the block (or fails)
            // lanes behind k carry nothing
            LP_INF_LANES(l) {
                if (l > k) { cnt[l] = 0; nm[l] = 0; }
            }
            LaneVar<uint32_t> coff, moff;
            const uint32_t T = excl_scan(coff, cnt);
            excl_scan(moff, nm);
            if (T > s.cap - o) {
                // more data than the image has room for: libpng stops reading at the last row; an earlier kernel
                // version and the tests treat it as corrupt
                return -3;
            }
            // pass C: write.  Lanes whose bytes end within the window budget write everything; the first lane
            // that would cross it writes what fits; lanes behind it (and behind k) write nothing.
            LaneVar<uint32_t> wexit, wcnt, wnm, wflag, full;
            LP_INF_LANES(l) { full[l] = (l <= k && coff[l] + cnt[l] <= kCapT) ? 1u : 0u; }
            const uint32_t fm = ballot(full);
            const int nfull = (int)ffs32(~fm) - 1 < 0 ? 32 : (int)ffs32(~fm) - 1;  // leading full lanes
            LP_INF_LANES(l) {
                wexit[l] = exitp[l]; wcnt[l] = 0; wnm[l] = 0; wflag[l] = kFlagNone;
                const bool active = l <= nfull && l <= k && l < 32;
                if (active) {
                    const uint32_t budget = l < nfull ? 0xFFFFFFFFu : (kCapT > coff[l] ? kCapT - coff[l] : 0u);
                    decode_span<true, false, false>(ws, start[l], rel + (uint32_t)(l + 1) * kSubBits, o + coff[l], budget,
                                                    s.mlist + moff[l], sp, 0, l, 0, 0, ws.mrec + moff[l], o);
                    wexit[l] = sp.exit; wcnt[l] = sp.cnt; wnm[l] = sp.nm; wflag[l] = sp.flag;
                }
            }
            wsync();
            {
                LP_INF_LANES(l) { term[l] = wflag[l] == kFlagBad; }
                if (ballot(term)) return -3;
            }
            if (k < 32 && bcast(flag, k) == kFlagBad && nfull > k) return -3;  // an invalid code in the true stream
            // what was actually written: lanes [0, last_w], the last one perhaps partially
            const int last_w = nfull < 32 ? (nfull <= k ? nfull : k) : 31;
            const int last_lane = last_w > 31 ? 31 : last_w;
            const uint32_t new_rel = bcast(wexit, last_lane);
            const uint32_t wT = bcast(coff, last_lane) + bcast(wcnt, last_lane);
            const uint32_t wM = bcast(moff, last_lane) + bcast(wnm, last_lane);
            eob = (last_lane == k) && bcast(wflag, last_lane) == kFlagEob;
            LP_INF_CLOCK(12);
            LP_INF_COUNT(4, nfull < 32 && nfull <= k ? 1 : 0);
            LP_INF_COUNT(5, wM);
            // matches: every lane copies its own (they lie in its own output segment, in order).  A copy may
            // start once its source bytes are final: below the high-water mark H (everything in front of the
            // first lane that is still blocked), or inside the lane's own finished prefix.  The lowest
            // unfinished lane is never blocked, so every round finishes at least one more lane; PNG data
            // mostly refers a pixel or a scanline back, i.e. to the lane itself or in front of the window.
            const uint32_t ring_lo = o + wT > kRing ? o + wT - kRing : 0;  // oldest position the ring still holds
            // phase D0: matches whose source lies entirely in front of the window (three quarters of them in filtered
            // PNG data: the pixel above is a scanline back) depend on nothing in the window -- they are spread evenly
            // over the lanes, whatever subsequence they came from.  Sources behind the ring are read from the flushed
            // output, eight byte-loads in flight at a time.
            uint32_t n_near = 0;
            if (wM) {
                for (uint32_t base = 0; base < wM; base += 32) {
                    LaneVar<uint32_t> isnear;
                    LP_INF_LANES(l) {
                        const uint32_t j = base + (uint32_t)l;
                        isnear[l] = 0;
                        if (j < wM) {
                            const uint32_t rec = ws.mrec[j];
                            Match m;
                            if (((rec >> 13) & 127u) == 127u) {
                                m = s.mlist[j];
                            } else {
                                m.q = o + (rec >> 20);
                                m.ld = ((((rec >> 13) & 127u) + 3u) << 16) | (rec & 8191u);
                            }
                            const uint32_t len = m.ld >> 16, dist = m.ld & 0xFFFFu;
                            if (!(dist >= len && (m.q - o) + len <= dist)) {
                                isnear[l] = 1;  // depends on window bytes: phase D1
                            } else {
                                const uint32_t src0 = m.q - dist;
                                for (uint32_t i0 = 0; i0 < len; i0 += 8) {
                                    uint8_t b[8];
#ifndef LP_INF_HOST
#pragma unroll
#endif
                                    for (uint32_t k = 0; k < 8; k++) {
                                        const uint32_t sq = src0 + i0 + k;
                                        b[k] = i0 + k < len ? (sq >= ring_lo ? ws.ring[sq & kRingMask] : s.out[sq]) : (uint8_t)0;
                                    }
#ifndef LP_INF_HOST
#pragma unroll
#endif
                                    for (uint32_t k = 0; k < 8; k++)
                                        if (i0 + k < len) ws.ring[(m.q + i0 + k) & kRingMask] = b[k];
                                }
                            }
                        }
                    }
                    const uint32_t nm_ = ballot(isnear);
                    LP_INF_LANES(l) {
                        if (isnear[l]) ws.near[n_near + popc32(nm_ & ((1u << l) - 1u))] = (uint16_t)(base + (uint32_t)l);
                    }
                    n_near += popc32(nm_);
                }
                wsync();
            }
#ifndef LP_INF_D1_ROUNDS
            // phase D1, serial form (measured 7 % faster in a 1024-stream batch than the per-lane rounds kept below under
            // LP_INF_D1_ROUNDS: the chains leave 2-3 lanes active either way, and this form has nothing to check): the matches that depend on window bytes, in stream order, by one lane -- every
            // source is final when its match is reached, so there is nothing to check
            if (n_near) {
                LP_INF_LANES(l) {
                    if (l == 0) {
                        for (uint32_t k = 0; k < n_near; k++) {
                            const uint32_t j = ws.near[k], rec = ws.mrec[j];
                            Match m;
                            if (((rec >> 13) & 127u) == 127u) {
                                m = s.mlist[j];
                            } else {
                                m.q = o + (rec >> 20);
                                m.ld = ((((rec >> 13) & 127u) + 3u) << 16) | (rec & 8191u);
                            }
                            const uint32_t len = m.ld >> 16, dist = m.ld & 0xFFFFu, src0 = m.q - dist;
                            uint32_t kk = 0;
                            for (uint32_t i = 0; i < len; i++) {
                                const uint32_t sq = src0 + kk;
                                ws.ring[(m.q + i) & kRingMask] = sq >= ring_lo ? ws.ring[sq & kRingMask] : s.out[sq];
                                if (++kk == dist) kk = 0;
                            }
                        }
                    }
                }
                wsync();
            }
#else
            if (n_near) {
                LaneVar<uint32_t> mi, mend, own_start, blocked_q, notdone;
                LP_INF_LANES(l) {
                    const bool wrote = l <= last_lane;
                    mi[l] = moff[l];
                    mend[l] = wrote ? moff[l] + wnm[l] : moff[l];
                    own_start[l] = o + coff[l];
                    blocked_q[l] = 0;
                }
                uint32_t H = o;
                for (int round = 0; round < 34; round++) {
#if defined(LP_INF_STATS) && defined(LP_INF_HOST)
                    uint32_t round_max = 0;
#endif
                    LP_INF_LANES(l) {
#if defined(LP_INF_STATS) && defined(LP_INF_HOST)
                        uint32_t done_here = 0;
#endif
                        while (mi[l] < mend[l]) {
                            const uint32_t rec = ws.mrec[mi[l]];
                            Match m;
                            if (((rec >> 13) & 127u) == 127u) {
                                m = s.mlist[mi[l]];
                            } else {
                                m.q = o + (rec >> 20);
                                m.ld = ((((rec >> 13) & 127u) + 3u) << 16) | (rec & 8191u);
                            }
                            const uint32_t len = m.ld >> 16, dist = m.ld & 0xFFFFu;
                            if (dist >= len && (m.q - o) + len <= dist) {  // done in phase D0
                                mi[l]++;
                                continue;
                            }
                            const uint32_t src0 = m.q - dist, L = dist < len ? dist : len;
                            const bool ok = src0 + L <= H || src0 >= own_start[l] || own_start[l] <= H;
                            if (!ok) {
                                blocked_q[l] = m.q;
#if defined(LP_INF_STATS) && defined(LP_INF_HOST)
                                g_stats[9]++;
#endif
                                break;
                            }
#if defined(LP_INF_STATS) && defined(LP_INF_HOST)
                            done_here++;
                            if (src0 < ring_lo) { g_stats[8]++; g_stats[10] += len; }
                            if (src0 + L <= o) g_stats[11]++;
#endif
                            const uint32_t sp_ = src0 & kRingMask, dp_ = m.q & kRingMask;
                            if (src0 >= ring_lo && sp_ + len <= kRing && dp_ + len <= kRing) {
                                // both ends in the ring without wrap-around (the usual case): four bytes per step, loads
                                // first, so they are in flight together instead of one dependent load-store pair per byte
                                if (dist >= 4 || dist >= len) {
                                    // (dist >= 4: a step of four never reads a byte it has not yet written)
                                    uint32_t i = 0;
                                    for (; i + 4 <= len; i += 4) {
                                        const uint8_t b0 = ws.ring[sp_ + i], b1 = ws.ring[sp_ + i + 1], b2 = ws.ring[sp_ + i + 2],
                                                      b3 = ws.ring[sp_ + i + 3];
                                        ws.ring[dp_ + i] = b0; ws.ring[dp_ + i + 1] = b1; ws.ring[dp_ + i + 2] = b2; ws.ring[dp_ + i + 3] = b3;
                                    }
                                    if (i < len) {
                                        const uint32_t r = len - i;  // 1..3
                                        const uint8_t b0 = ws.ring[sp_ + i], b1 = r > 1 ? ws.ring[sp_ + i + 1] : 0,
                                                      b2 = r > 2 ? ws.ring[sp_ + i + 2] : 0;
                                        ws.ring[dp_ + i] = b0;
                                        if (r > 1) ws.ring[dp_ + i + 1] = b1;
                                        if (r > 2) ws.ring[dp_ + i + 2] = b2;
                                    }
                                } else {
                                    // run with a period of 1..3 bytes: the pattern is read once, then only stored
                                    const uint8_t p0 = ws.ring[sp_], p1 = dist > 1 ? ws.ring[sp_ + 1] : p0,
                                                  p2 = dist > 2 ? ws.ring[sp_ + 2] : (dist > 1 ? p0 : p0);
                                    uint32_t k = 0;
                                    for (uint32_t i = 0; i < len; i++) {
                                        ws.ring[dp_ + i] = k == 0 ? p0 : k == 1 ? p1 : p2;
                                        if (++k == dist) k = 0;
                                    }
                                }
                            } else {
                                uint32_t k = 0;  // i mod dist, kept incrementally
                                for (uint32_t i = 0; i < len; i++) {
                                    const uint32_t sq = src0 + k;
                                    const uint8_t v = sq >= ring_lo ? ws.ring[sq & kRingMask] : s.out[sq];
                                    ws.ring[(m.q + i) & kRingMask] = v;
                                    if (++k == dist) k = 0;
                                }
                            }
                            mi[l]++;
                        }
#if defined(LP_INF_STATS) && defined(LP_INF_HOST)
                        if (done_here > round_max) round_max = done_here;
#endif
                        notdone[l] = mi[l] < mend[l] ? 1u : 0u;
                    }
                    wsync();
#if defined(LP_INF_STATS) && defined(LP_INF_HOST)
                    g_stats[12] += round_max;
#endif
                    const uint32_t nd = ballot(notdone);
                    if (!nd) break;
                    LP_INF_COUNT(6, 1);
                    H = bcast(blocked_q, (int)ffs32(nd) - 1);
                }
            }
#endif
            LP_INF_CLOCK(13);
            o += wT;
            P += new_rel - rel;
            const uint32_t upto = vec ? (o & ~15u) : o;
            flush_ring(ws, s, flushed, upto, vec);
            flushed = upto;
            LP_INF_CLOCK(14);
        }
    }
    // tail bytes the vector flush left behind
    LP_INF_LANES(l) {
        for (uint32_t p = flushed + (uint32_t)l; p < o; p += 32u) s.out[p] = ws.ring[p & kRingMask];
    }
    wsync();
    *produced = o;
    return 0;
}

}  // namespace lpinf
