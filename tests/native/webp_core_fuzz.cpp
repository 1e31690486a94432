// Probe (not part of the pytest suites): the product's WebP codec cores (vp8_core.h, vp8l_core.h -- the code the
// device kernels run, compiled here for the host exactly as oracle/oracle_webp.cpp does) under AddressSanitizer +
// UBSan on mutated VP8 / VP8L / ALPH payloads.  An out-of-bounds access found here is one the kernel would make.
// Inputs get the slack the device gives them (webp_decode.cu: img_len + 4096 bytes, zero-filled here) minus a
// margin, the VP8L arena is sized with the device's formula.
//
//   bash tests/native/webp_core_fuzz.sh [iterations]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../../lilliput_b200/csrc/vp8_core.h"
#include "../../lilliput_b200/csrc/vp8l_core.h"

struct Payload {
    int kind;  // 0 VP8, 1 VP8L, 2 ALPH
    int w, h;  // ALPH: of the frame it belongs to
    std::vector<uint8_t> bytes;
};

static uint32_t le32(const uint8_t* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }

static void walk(const uint8_t* p, size_t n, std::vector<Payload>& out) {
    size_t pos = 0;
    int fw = 0, fh = 0;
    std::vector<uint8_t> pending_alph;
    while (pos + 8 <= n) {
        const uint32_t sz = le32(p + pos + 4);
        const uint8_t* body = p + pos + 8;
        if (pos + 8 + sz > n) break;
        if (!memcmp(p + pos, "ANMF", 4) && sz >= 16) {
            walk(body + 16, sz - 16, out);
        } else if (!memcmp(p + pos, "ALPH", 4)) {
            pending_alph.assign(body, body + sz);
        } else if (!memcmp(p + pos, "VP8 ", 4) && sz >= 10) {
            fw = ((body[7] << 8) | body[6]) & 0x3fff;
            fh = ((body[9] << 8) | body[8]) & 0x3fff;
            out.push_back({0, fw, fh, std::vector<uint8_t>(body, body + sz)});
            if (!pending_alph.empty()) out.push_back({2, fw, fh, pending_alph});
            pending_alph.clear();
        } else if (!memcmp(p + pos, "VP8L", 4) && sz >= 5) {
            const uint32_t bits = le32(body + 1);
            out.push_back({1, (int)(bits & 0x3fff) + 1, (int)((bits >> 14) & 0x3fff) + 1, std::vector<uint8_t>(body, body + sz)});
        }
        pos += 8 + sz + (sz & 1);
    }
}

static std::vector<uint8_t> read_file(const char* p) {
    std::vector<uint8_t> v;
    FILE* f = fopen(p, "rb");
    if (!f) return v;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    v.resize(n > 0 ? (size_t)n : 0);
    if (n > 0 && fread(v.data(), 1, (size_t)n, f) != (size_t)n) v.clear();
    fclose(f);
    return v;
}

static void mutate(std::mt19937& rng, std::vector<uint8_t>& b) {
    if (b.size() < 12) return;
    const int mode = (int)(rng() % 6);
    if (mode == 0) {
        b.resize(6 + rng() % (b.size() - 6));
        return;
    }
    if (mode == 5) {
        const size_t n = 1 + rng() % 32, a = rng() % (b.size() - 1), c = rng() % (b.size() - 1);
        for (size_t k = 0; k < n && a + k < b.size() && c + k < b.size(); k++) b[c + k] = b[a + k];
        return;
    }
    const int n = 1 + (int)(rng() % 6);
    for (int k = 0; k < n; k++) {
        const size_t i = rng() % b.size();
        switch (mode) {
            case 1: b[i] ^= (uint8_t)(1u << (rng() % 8)); break;
            case 2: b[i] = (uint8_t)rng(); break;
            case 3: b[i] = 0xFF; break;
            default: b[i] = 0; break;
        }
    }
}

constexpr size_t kSlack = 64;  // the device gives 4096; anything that needs more than this is worth a look

static long g_ok[3], g_err[3];

static void run(const Payload& p, const std::vector<uint8_t>& bytes) {
    std::vector<uint8_t> in(bytes.size() + kSlack, 0);
    memcpy(in.data(), bytes.data(), bytes.size());
    const size_t n = bytes.size();
    if (p.kind == 0) {
        vp8::FrameHdr h;
        vp8::BoolDec br;
        uint8_t proba[1056];
        if (vp8::parse_frame_header(in.data(), n, h, br, proba)) { g_err[0]++; return; }
        if ((size_t)h.mb_w * h.mb_h > 4096) { g_err[0]++; return; }  // a mutated header asking for a huge frame: just slow
        std::vector<uint8_t> mem(vp8::work_bytes(h.mb_w, h.mb_h));
        vp8::Work w;
        vp8::work_carve(mem.data(), h.mb_w, h.mb_h, w);
        memcpy(w.proba, proba, 1056);
        if (vp8::decode_macroblocks(in.data(), h, br, w)) { g_err[0]++; return; }
        if (h.filter_type > 0)
            for (int y = 0; y < h.mb_h; y++)
                for (int x = 0; x < h.mb_w; x++) vp8::filter_macroblock(h, w, x, y);
        g_ok[0]++;
        return;
    }
    // dimensions come from the container (VP8L: the header repeats them and the decoder checks)
    int w = p.w, h = p.h;
    if (p.kind == 1 && n >= 5) {
        const uint32_t bits = le32(in.data() + 1);
        w = (int)(bits & 0x3fff) + 1;
        h = (int)((bits >> 14) & 0x3fff) + 1;
    }
    if ((size_t)w * h > (1u << 20)) { g_err[p.kind]++; return; }
    const size_t npix = (size_t)w * h;
    static std::vector<uint8_t> mem;
    mem.assign(npix * 12 + (16u << 20), 0);  // webp_decode.cu arena_need
    vp8l::Arena a{mem.data(), mem.size(), 0};
    int rc;
    if (p.kind == 1) {
        uint32_t* px = nullptr;
        rc = vp8l::decode_vp8l(in.data(), n, w, h, a, &px);
        if (!rc) {  // the output kernel reads every pixel
            uint32_t s = 0;
            for (size_t i = 0; i < npix; i++) s += px[i];
            if (s == 0x12345678u) printf(" ");
        }
    } else {
        std::vector<uint8_t> alpha(npix);
        rc = vp8l::decode_alph(in.data(), n, w, h, a, alpha.data());
    }
    (rc ? g_err : g_ok)[p.kind]++;
}

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    const long iters = atol(argv[1]);
    std::vector<Payload> seeds;
    for (int i = 2; i < argc; i++) {
        auto v = read_file(argv[i]);
        if (v.size() > 12 && !memcmp(v.data(), "RIFF", 4) && !memcmp(v.data() + 8, "WEBP", 4)) walk(v.data() + 12, v.size() - 12, seeds);
    }
    if (seeds.empty()) return 2;
    size_t kinds[3] = {0, 0, 0};
    for (auto& s : seeds) kinds[s.kind]++;
    fprintf(stderr, "%zu payloads: %zu VP8, %zu VP8L, %zu ALPH\n", seeds.size(), kinds[0], kinds[1], kinds[2]);
    for (auto& s : seeds) run(s, s.bytes);
    std::mt19937 rng(getenv("LP_FUZZ_SEED") ? (unsigned)atol(getenv("LP_FUZZ_SEED")) : 4242u);
    for (long it = 0; it < iters; it++) {
        const Payload& s = seeds[rng() % seeds.size()];
        if (s.bytes.size() > 300000) continue;
        std::vector<uint8_t> b = s.bytes;
        const int rounds = 1 + (int)(rng() % 3);
        for (int r = 0; r < rounds; r++) mutate(rng, b);
        run(s, b);
        if ((it + 1) % 10000 == 0) fprintf(stderr, "%ld iterations\n", it + 1);
    }
    printf("done: VP8 %ld decoded / %ld refused, VP8L %ld / %ld, ALPH %ld / %ld\n", g_ok[0], g_err[0], g_ok[1], g_err[1],
           g_ok[2], g_err[2]);
    return 0;
}
