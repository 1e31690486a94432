/* oracle_gif.c -- TEST INFRASTRUCTURE (oracle/): CPU restatement of lilliput's GIF adapter.
 *
 * Decode: giflib 5.2.2's record walk + LZW (DGifGetRecordType / DGifGetExtension / DGifGetImageDesc /
 * DGifGetLine, pinned in deps/build-deps-linux.sh:260) and the reference's own full-canvas compositor
 * (ref giflib.cpp:349-568, background colour :595-636, forced transparent index :548-565).
 * Encode: the reference's palette mapping with its order-dependent 15-bit memo (ref giflib.cpp:934-1098),
 * transparency removal (:896-917), giflib's EGifCompressLine / EGifCompressOutput / EGifBufferedOutput
 * and the container writers (EGifPutScreenDesc / PutImageDesc / PutExtension*, ref giflib.cpp:803-860,
 * 1100-1222).  Serial and simple on purpose.  Pinned on the reference itself: frames and whole GIF->GIF
 * files in tests/golden (made through oracle/_ref) must come out bit / byte identical
 * (tests/test_oracle_gif.py).  Never linked into the product. */
#include <limits.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int function;
    int len;
    uint8_t bytes[255];
} Ext;

typedef struct {
    int disposal, delay, transparent, user_input;
} Gcb;

typedef struct oracle_gif {
    const uint8_t* p;
    size_t n, pos;
    int sw, sh, bg_index, gct_colors;
    const uint8_t* gct;
    uint8_t packed, aspect;
    /* current image */
    int left, top, width, height, interlace, lct_colors, min_code;
    const uint8_t* lct;
    /* extension blocks since the last frame (gif->ExtensionBlocks) */
    Ext* ext;
    int next, cap_ext;
    int clear_ext;
    int have_first;
    int prev_disposal, prev_left, prev_top, prev_width, prev_height;
    uint8_t bg[4]; /* B G R A */
    uint8_t* prev; /* restore-previous snapshot */
    uint8_t* idx;
    size_t idx_cap;
} oracle_gif;

static void gcb_default(Gcb* g) { g->disposal = 0; g->delay = 0; g->transparent = -1; g->user_input = 0; }
/* giflib_get_frame_gcb (ref giflib.cpp:248-270): every well-formed (4-byte) graphic control block, in order */
static Gcb frame_gcb(const Ext* e, int n) {
    Gcb g;
    gcb_default(&g);
    for (int i = 0; i < n; i++)
        if (e[i].function == 0xF9 && e[i].len == 4) {
            const uint8_t* b = e[i].bytes;
            g.disposal = (b[0] >> 2) & 7;
            g.user_input = (b[0] >> 1) & 1;
            g.delay = b[1] | (b[2] << 8);
            g.transparent = (b[0] & 1) ? b[3] : -1;
        }
    return g;
}
/* giflib_set_frame_gcb (ref giflib.cpp:272-291) */
static void set_frame_gcb(Ext* e, int n, const Gcb* g) {
    for (int i = 0; i < n; i++)
        if (e[i].function == 0xF9 && e[i].len >= 4) {
            e[i].bytes[0] = (uint8_t)((g->transparent != -1 ? 1 : 0) | (g->user_input ? 2 : 0) | ((g->disposal & 7) << 2));
            e[i].bytes[1] = (uint8_t)(g->delay & 0xFF);
            e[i].bytes[2] = (uint8_t)((g->delay >> 8) & 0xFF);
            e[i].bytes[3] = (uint8_t)g->transparent;
        }
}

oracle_gif* oracle_gif_open(const uint8_t* in, size_t len) {
    if (len < 13 || (memcmp(in, "GIF87a", 6) && memcmp(in, "GIF89a", 6))) return NULL;
    oracle_gif* d = calloc(1, sizeof(*d));
    d->p = in;
    d->n = len;
    d->sw = in[6] | (in[7] << 8);
    d->sh = in[8] | (in[9] << 8);
    d->packed = in[10];
    d->bg_index = in[11];
    d->aspect = in[12];
    d->pos = 13;
    if (d->packed & 0x80) {
        d->gct_colors = 1 << ((d->packed & 7) + 1);
        if (d->pos + (size_t)d->gct_colors * 3 > len) { free(d); return NULL; }
        d->gct = in + d->pos;
        d->pos += (size_t)d->gct_colors * 3;
    }
    if (d->sw <= 0 || d->sh <= 0) { free(d); return NULL; }
    d->prev = calloc((size_t)d->sw * d->sh, 4);
    d->bg[0] = d->bg[1] = d->bg[2] = d->bg[3] = 255;
    return d;
}
void oracle_gif_close(oracle_gif* d) {
    if (!d) return;
    free(d->ext);
    free(d->prev);
    free(d->idx);
    free(d);
}
int oracle_gif_width(const oracle_gif* d) { return d->sw; }
int oracle_gif_height(const oracle_gif* d) { return d->sh; }

static int sub_block(oracle_gif* d, const uint8_t** data, int* len) {
    if (d->pos >= d->n) return 0;
    *len = d->p[d->pos++];
    *data = d->p + d->pos;
    if (d->pos + (size_t)*len > d->n) return 0;
    d->pos += (size_t)*len;
    return 1;
}
static void push_ext(oracle_gif* d, int function, const uint8_t* b, int len) {
    if (d->next == d->cap_ext) {
        d->cap_ext = d->cap_ext ? d->cap_ext * 2 : 8;
        d->ext = realloc(d->ext, (size_t)d->cap_ext * sizeof(Ext));
    }
    d->ext[d->next].function = function;
    d->ext[d->next].len = len;
    memcpy(d->ext[d->next].bytes, b, (size_t)len);
    d->next++;
}

/* Advances to the next image descriptor.  1 = frame header read, 0 = end of file, -1 = error. */
static int next_header(oracle_gif* d) {
    if (d->clear_ext) { d->next = 0; d->clear_ext = 0; }
    for (;;) {
        if (d->pos >= d->n) return -1;
        const uint8_t rec = d->p[d->pos++];
        if (rec == 0x3B) return 0;
        if (rec == 0x21) { /* extension: label, then sub-blocks (ref giflib.cpp:209-246) */
            if (d->pos >= d->n) return -1;
            const int label = d->p[d->pos++];
            const uint8_t* data;
            int len, first = 1;
            if (!sub_block(d, &data, &len)) return -1;
            while (len) {
                push_ext(d, first ? label : 0, data, len);
                first = 0;
                if (!sub_block(d, &data, &len)) return -1;
            }
            continue;
        }
        if (rec != 0x2C) return -1;
        if (d->pos + 9 > d->n) return -1;
        const uint8_t* q = d->p + d->pos;
        d->left = q[0] | (q[1] << 8);
        d->top = q[2] | (q[3] << 8);
        d->width = q[4] | (q[5] << 8);
        d->height = q[6] | (q[7] << 8);
        d->interlace = (q[8] & 0x40) != 0;
        d->pos += 9;
        d->lct = NULL;
        d->lct_colors = 0;
        if (q[8] & 0x80) {
            d->lct_colors = 1 << ((q[8] & 7) + 1);
            if (d->pos + (size_t)d->lct_colors * 3 > d->n) return -1;
            d->lct = d->p + d->pos;
            d->pos += (size_t)d->lct_colors * 3;
        }
        if (d->pos >= d->n) return -1;
        d->min_code = d->p[d->pos++];
        if (d->min_code > 8) return -1;
        return 1;
    }
}

/* giflib's DGifDecompressLine / DGifDecompressInput: LZW with deferred code-size growth. */
static int lzw_decode(const uint8_t* z, size_t zn, int min_code, uint8_t* out, size_t npix) {
    uint16_t prefix[4096];
    uint8_t suffix[4096], stack[4097];
    const int clear = 1 << min_code, eof = clear + 1;
    int bits = min_code + 1, maxcode1 = 1 << bits, running = clear + 2, top = clear + 2, last = -1;
    uint64_t acc = 0;
    int cnt = 0;
    size_t zp = 0, o = 0;
    while (o < npix) {
        while (cnt < bits) {
            if (zp >= zn) return -1;
            acc |= (uint64_t)z[zp++] << cnt;
            cnt += 8;
        }
        const int code = (int)(acc & ((1u << bits) - 1));
        acc >>= bits;
        cnt -= bits;
        if (running < 4097 && ++running > maxcode1 && bits < 12) {
            maxcode1 <<= 1;
            bits++;
        }
        if (code == eof) return -1;
        if (code == clear) {
            bits = min_code + 1;
            maxcode1 = 1 << bits;
            running = clear + 2;
            top = clear + 2;
            last = -1;
            continue;
        }
        const int create = running - 2; /* the entry this code defines, if there is a previous string */
        int sp = 0, cur = code;
        if (code >= clear) {
            if (last >= 0 && code == create && code == top) { /* KwKwK: previous string + its own first symbol */
                cur = last;
                while (cur >= clear && cur < 4096 && sp < 4096) cur = prefix[cur];
                stack[sp++] = (uint8_t)cur;
                cur = last;
            } else if (!(code > eof && code < top)) {
                return -1;
            }
        }
        while (cur >= clear) {
            if (cur >= 4096 || sp > 4095) return -1;
            stack[sp++] = suffix[cur];
            cur = prefix[cur];
        }
        stack[sp++] = (uint8_t)cur;
        if (last >= 0 && create < 4096 && create == top) { /* giflib defines an entry once (Prefix == NO_SUCH_CODE) */
            prefix[create] = (uint16_t)last;
            suffix[create] = (uint8_t)cur; /* first symbol of the string just emitted */
            top++;
        }
        last = code;
        while (sp > 0 && o < npix) out[o++] = stack[--sp];
    }
    return 0;
}

/* Decodes the next frame onto `canvas` (sw*sh BGRA, kept by the caller between calls).
 * Returns 1 = frame, 0 = end of file, -1 = error.  delay in 1/100 s; disposal = giflib DisposalMode. */
int oracle_gif_next_frame(oracle_gif* d, uint8_t* canvas, int* delay_cs, int* disposal) {
    const int r = next_header(d);
    if (r <= 0) return r;
    if (d->width <= 0 || d->height <= 0) return -1;
    /* LZW sub-blocks */
    size_t zn = 0, zcap = 1 << 16;
    uint8_t* z = malloc(zcap);
    for (;;) {
        const uint8_t* data;
        int len;
        if (!sub_block(d, &data, &len)) { free(z); return -1; }
        if (!len) break;
        if (zn + (size_t)len > zcap) { zcap = (zn + (size_t)len) * 2; z = realloc(z, zcap); }
        memcpy(z + zn, data, (size_t)len);
        zn += (size_t)len;
    }
    const size_t npix = (size_t)d->width * d->height;
    if (npix > d->idx_cap) { d->idx = realloc(d->idx, npix); d->idx_cap = npix; }
    const int lz = lzw_decode(z, zn, d->min_code, d->idx, npix);
    free(z);
    if (lz) return -1;
    Gcb g = frame_gcb(d->ext, d->next);
    const uint8_t* colors = d->lct ? d->lct : d->gct;
    const int ncolors = d->lct ? d->lct_colors : d->gct_colors;
    if (!colors) return -1;
    const int cw = d->sw, ch = d->sh;
    if (!d->have_first) { /* ref giflib.cpp:595-636 */
        const int valid = d->gct && d->bg_index >= 0 && d->bg_index < d->gct_colors;
        d->bg[2] = valid ? d->gct[d->bg_index * 3] : 255;
        d->bg[1] = valid ? d->gct[d->bg_index * 3 + 1] : 255;
        d->bg[0] = valid ? d->gct[d->bg_index * 3 + 2] : 255;
        d->bg[3] = g.transparent != -1 ? 0 : 255;
        for (size_t i = 0; i < (size_t)cw * ch; i++) memcpy(canvas + 4 * i, d->bg, 4);
    } else { /* dispose the previous frame's (clipped) rectangle, then snapshot (ref giflib.cpp:383-470) */
        int pl = d->prev_left, pt = d->prev_top, pw = d->prev_width, ph = d->prev_height;
        if (pl < 0) { pw += pl; pl = 0; }
        if (pt < 0) { ph += pt; pt = 0; }
        if (pl + pw > cw) pw = cw - pl;
        if (pt + ph > ch) ph = ch - pt;
        for (int y = pt; y < pt + ph; y++)
            for (int x = pl; x < pl + pw; x++) {
                uint8_t* px = canvas + 4 * ((size_t)y * cw + x);
                if (d->prev_disposal == 2) memcpy(px, d->bg, 4);
                else if (d->prev_disposal == 3) memcpy(px, d->prev + 4 * ((size_t)y * cw + x), 4);
            }
        memcpy(d->prev, canvas, (size_t)cw * ch * 4);
    }
    /* draw: non-transparent, in-palette pixels, clipped to the canvas; interlaced rows in 4 passes */
    static const int off[4] = {0, 4, 2, 1}, jmp[4] = {8, 8, 4, 2};
    size_t row = 0;
    for (int ps = 0; ps < (d->interlace ? 4 : 1); ps++)
        for (int fy = d->interlace ? off[ps] : 0; fy < d->height; fy += d->interlace ? jmp[ps] : 1, row++) {
            const int y = d->top + fy;
            if (y < 0 || y >= ch) continue;
            for (int fx = 0; fx < d->width; fx++) {
                const int x = d->left + fx;
                if (x < 0 || x >= cw) continue;
                const int i = d->idx[row * d->width + fx];
                if (i == g.transparent || i >= ncolors) continue;
                uint8_t* px = canvas + 4 * ((size_t)y * cw + x);
                px[0] = colors[i * 3 + 2];
                px[1] = colors[i * 3 + 1];
                px[2] = colors[i * 3];
                px[3] = 255;
            }
        }
    /* a partial frame without a transparent index gets one forced (ref giflib.cpp:548-565) */
    if ((d->height < ch || d->width < cw || d->left != 0 || d->top != 0) && g.transparent == -1) {
        Gcb f = g;
        f.transparent = ncolors - 1;
        set_frame_gcb(d->ext, d->next, &f);
    }
    if (delay_cs) *delay_cs = g.delay;
    if (disposal) *disposal = g.disposal;
    d->prev_disposal = g.disposal;
    d->prev_left = d->left;
    d->prev_top = d->top;
    d->prev_width = d->width;
    d->prev_height = d->height;
    d->have_first = 1;
    d->clear_ext = 1;
    return 1;
}

/* giflib_decoder_skip_frame (ref giflib.cpp:570-590): header + data skipped, nothing composited, and the
 * extension blocks are NOT marked for clearing -- they pile up until the next decoded frame (or the flush). */
int oracle_gif_skip_frame(oracle_gif* d) {
    const int r = next_header(d);
    if (r <= 0) return r;
    for (;;) {
        const uint8_t* data;
        int len;
        if (!sub_block(d, &data, &len)) return -1;
        if (!len) break;
    }
    return 1;
}

/* ------------------------------------------------------------------ encoder */

typedef struct oracle_gif_enc {
    uint8_t* dst;
    size_t cap, off;
    int sw, sh, bg, has_gct, ngct;
    uint8_t gct[768];
    int wrote_first, prev_disposal, have_prev_colors, nprev;
    uint8_t prev_colors[768];
    uint8_t* prev_frame;
    uint8_t present[1 << 15], index[1 << 15]; /* the 15-bit crushed-colour memo */
    int failed;
} oracle_gif_enc;

static void put(oracle_gif_enc* e, const void* p, size_t n) {
    if (e->off + n > e->cap) { e->failed = 1; return; }
    memcpy(e->dst + e->off, p, n);
    e->off += n;
}
static void put8(oracle_gif_enc* e, int v) { uint8_t b = (uint8_t)v; put(e, &b, 1); }
static void put16(oracle_gif_enc* e, int v) { put8(e, v & 0xff); put8(e, (v >> 8) & 0xff); }

oracle_gif_enc* oracle_gif_enc_open(const oracle_gif* d, int w, int h, uint8_t* dst, size_t cap) {
    oracle_gif_enc* e = calloc(1, sizeof(*e));
    e->dst = dst;
    e->cap = cap;
    e->sw = w;
    e->sh = h;
    e->has_gct = d->gct != NULL;
    e->bg = (e->has_gct && d->bg_index >= 0 && d->bg_index < d->gct_colors) ? d->bg_index : 0;
    if (e->has_gct) { e->ngct = d->gct_colors; memcpy(e->gct, d->gct, (size_t)d->gct_colors * 3); }
    e->prev_frame = calloc((size_t)w * h, 4);
    put(e, "GIF89a", 6);
    put16(e, w);
    put16(e, h);
    put8(e, e->has_gct ? d->packed : ((d->packed & 0x70) | 0x07));
    put8(e, e->bg);
    put8(e, d->aspect);
    if (e->has_gct) put(e, e->gct, (size_t)e->ngct * 3);
    return e;
}
void oracle_gif_enc_close(oracle_gif_enc* e) {
    if (!e) return;
    free(e->prev_frame);
    free(e);
}

static void write_exts(oracle_gif_enc* e, const Ext* x, int n) {
    for (int i = 0; i < n; i++) {
        if (x[i].function != 0) { put8(e, 0x21); put8(e, x[i].function); }
        put8(e, x[i].len);
        put(e, x[i].bytes, (size_t)x[i].len);
        if (i == n - 1 || x[i + 1].function != 0) put8(e, 0);
    }
}
static int rgb_dist(int r0, int g0, int b0, int r1, int g1, int b1) { return abs(r0 - r1) + abs(g0 - g1) + abs(b0 - b1); }

/* One full-canvas BGRA frame (w x h, packed).  Returns 1, or 0 on failure. */
int oracle_gif_enc_frame(oracle_gif_enc* e, const oracle_gif* d, const uint8_t* bgra, int w, int h) {
    if (w > e->sw || h > e->sh) return 0;
    const int has_local = d->lct != NULL;
    Ext* ext = malloc((size_t)(d->next ? d->next : 1) * sizeof(Ext));
    memcpy(ext, d->ext, (size_t)d->next * sizeof(Ext));
    Gcb g = frame_gcb(ext, d->next);
    if (g.transparent != -1 && e->has_gct && !has_local && g.transparent == e->bg && d->bg[3] == 255) {
        g.transparent = -1; /* ref giflib.cpp:896-917 */
        set_frame_gcb(ext, d->next, &g);
    }
    const uint8_t* colors = has_local ? d->lct : (e->has_gct ? e->gct : NULL);
    const int ncolors = has_local ? d->lct_colors : e->ngct;
    if (!colors) { free(ext); return 0; }
    int clear_memo = 1;
    if (e->wrote_first && e->have_prev_colors && e->nprev == ncolors) clear_memo = memcmp(e->prev_colors, colors, (size_t)ncolors * 3) != 0;
    if (clear_memo) memset(e->present, 0, sizeof(e->present));
    const int T = g.transparent, have_t = T != -1;
    const int prev_valid = e->wrote_first && (e->prev_disposal == 0 || e->prev_disposal == 1);
    uint8_t* px = malloc((size_t)w * h);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const uint8_t* s = bgra + 4 * ((size_t)y * w + x);
            const int B = s[0], G = s[1], R = s[2], A = s[3];
            if (A < 128 && have_t) { px[(size_t)y * w + x] = (uint8_t)T; continue; }
            const int crushed = ((R >> 3) << 10) | ((G >> 3) << 5) | (B >> 3);
            int least = INT_MAX, best = 0;
            if (!e->present[crushed]) {
                const int extreme = (R > 240 && G > 240 && B > 240) || (R < 15 && G < 15 && B < 15);
                const int Rc = extreme ? R : (R & 0xf8) | 4, Gc = extreme ? G : (G & 0xf8) | 4, Bc = extreme ? B : (B & 0xf8) | 4;
                for (int i = 0; i < ncolors; i++) {
                    if (i == T) continue;
                    const int dd = rgb_dist(Rc, Gc, Bc, colors[i * 3], colors[i * 3 + 1], colors[i * 3 + 2]);
                    if (dd < least) { least = dd; best = i; }
                }
                e->present[crushed] = 1;
                e->index[crushed] = (uint8_t)best;
            } else {
                best = e->index[crushed];
                least = rgb_dist(R, G, B, colors[best * 3], colors[best * 3 + 1], colors[best * 3 + 2]);
            }
            if (prev_valid && have_t) {
                const uint8_t* l = e->prev_frame + 4 * ((size_t)y * e->sw + x);
                if (rgb_dist(R, G, B, l[2], l[1], l[0]) < least) best = T;
            }
            px[(size_t)y * w + x] = (uint8_t)best;
        }
    for (int y = 0; y < h; y++) memcpy(e->prev_frame + 4 * (size_t)y * e->sw, bgra + 4 * (size_t)y * w, (size_t)w * 4);
    memcpy(e->prev_colors, colors, (size_t)ncolors * 3);
    e->nprev = ncolors;
    e->have_prev_colors = 1;
    e->prev_disposal = g.disposal;
    /* container */
    write_exts(e, ext, d->next);
    free(ext);
    int bpp = 1;
    while ((1 << bpp) < ncolors) bpp++;
    put8(e, 0x2C);
    put16(e, 0);
    put16(e, 0);
    put16(e, w);
    put16(e, h);
    put8(e, (has_local ? 0x80 : 0) | (d->interlace ? 0x40 : 0) | (has_local ? bpp - 1 : 0));
    if (has_local) put(e, d->lct, (size_t)d->lct_colors * 3);
    const int cb = bpp < 2 ? 2 : bpp;
    put8(e, cb);
    /* giflib EGifCompressLine / Output / BufferedOutput */
    {
        const int clear = 1 << cb, eof = clear + 1;
        int running_code = eof + 1, running_bits = cb + 1, max_code1 = 1 << running_bits;
        uint32_t dword = 0;
        int shift = 0, blk_n = 0;
        size_t blk_start = 0;
        int32_t* table = malloc(sizeof(int32_t) * (1 << 20)); /* (prefix << 8 | pixel) -> code, -1 = none */
        memset(table, 0xFF, sizeof(int32_t) * (1 << 20));
#define PUT_BYTE(b)                                                  \
    do {                                                             \
        if (blk_n == 0) { blk_start = e->off; put8(e, 0); }          \
        put8(e, (b));                                                \
        if (++blk_n == 255) { if (!e->failed) e->dst[blk_start] = 255; blk_n = 0; } \
    } while (0)
#define PUT_CODE(c)                                                                         \
    do {                                                                                    \
        dword |= (uint32_t)(c) << shift;                                                    \
        shift += running_bits;                                                              \
        while (shift >= 8) { PUT_BYTE(dword & 0xff); dword >>= 8; shift -= 8; }             \
        if (running_code >= max_code1 && (c) <= 4095) max_code1 = 1 << ++running_bits;      \
    } while (0)
        PUT_CODE(clear);
        int crnt = -1;
        static const int off[4] = {0, 4, 2, 1}, jmp[4] = {8, 8, 4, 2};
        for (int ps = 0; ps < (d->interlace ? 4 : 1); ps++)
            for (int y = d->interlace ? off[ps] : 0; y < h; y += d->interlace ? jmp[ps] : 1)
                for (int x = 0; x < w; x++) {
                    const int pixel = px[(size_t)y * w + x] & ((1 << cb) - 1);
                    if (crnt < 0) { crnt = pixel; continue; }
                    const int key = (crnt << 8) | pixel;
                    if (table[key] >= 0) { crnt = table[key]; continue; }
                    PUT_CODE(crnt);
                    crnt = pixel;
                    if (running_code >= 4095) {
                        PUT_CODE(clear);
                        running_code = eof + 1;
                        running_bits = cb + 1;
                        max_code1 = 1 << running_bits;
                        memset(table, 0xFF, sizeof(int32_t) * (1 << 20));
                    } else {
                        table[key] = running_code++;
                    }
                }
        PUT_CODE(crnt);
        PUT_CODE(eof);
        while (shift > 0) { PUT_BYTE(dword & 0xff); dword >>= 8; shift -= 8; }
        if (blk_n > 0 && !e->failed) e->dst[blk_start] = (uint8_t)blk_n;
        put8(e, 0);
        free(table);
#undef PUT_BYTE
#undef PUT_CODE
    }
    free(px);
    e->wrote_first = 1;
    return !e->failed;
}

/* Trailing extension blocks + the GIF trailer; returns the file size or 0. */
size_t oracle_gif_enc_finish(oracle_gif_enc* e, const oracle_gif* d) {
    write_exts(e, d->ext, d->next);
    put8(e, 0x3B);
    return e->failed ? 0 : e->off;
}
