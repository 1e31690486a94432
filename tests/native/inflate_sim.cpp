// inflate_sim.cpp -- the product's warp-parallel inflate (lilliput_b200/csrc/inflate_core.h) compiled
// for the HOST with its 32 lanes simulated by loops, so the CPU suite can run the kernel's exact control
// flow against zlib (tests/test_inflate_core.py).  Test infrastructure; not linked into the product.
#define LP_INF_HOST 1
#include "../../lilliput_b200/csrc/inflate_core.h"

#include <cstdlib>
#include <vector>

extern "C" int lp_inflate_sim(const uint8_t* z, uint32_t z_len, uint8_t* out, uint32_t cap, uint32_t* produced) {
    static thread_local lpinf::WarpShared ws;
    std::vector<lpinf::Match> ml(lpinf::kMaxMatches + 64);
    // 16-byte aligned output exercises the vector flush; callers pass aligned or unaligned buffers
    lpinf::Stream s{z, z_len, out, cap, ml.data()};
    return lpinf::inflate_stream(ws, s, produced);
}

#ifdef LP_INF_STATS
extern "C" void lp_inflate_sim_stats(unsigned long long* out16, int reset) {
    for (int i = 0; i < 16; i++) { out16[i] = lpinf::g_stats[i]; if (reset) lpinf::g_stats[i] = 0; }
}
#endif
