#!/usr/bin/env python
"""Prints kVp8Pred4[10][16] for lilliput_b200/csrc/vp8_enc_core.h: the 4x4 intra predictors of RFC 6386 s.12.3 as data --
for every (mode, pixel) which edge samples are averaged -- read out of vp8::pred_4x4's own source (vp8_core.h), so that
the device can evaluate all ten modes of a block without a ten-way divergent switch.  Entry = i0 | i1 << 4 | i2 << 8 |
kind << 12 over the edge array e[13] = {L, K, J, I, X, A, B, C, D, E, F, G, H}; kind 0 = (e[i0] + 2 e[i1] + e[i2] + 2) >> 2,
1 = (e[i0] + e[i1] + 1) >> 1, 2 = e[i0], 3 = computed by the caller (B_DC, B_TM).
tests/test_webp_encode_core.py checks the table against vp8::pred_4x4 on random edges."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
E = {n: i for i, n in enumerate("LKJIXABCDEFGH")}
MODES = ["B_DC", "B_TM", "B_VE", "B_HE", "B_RD", "B_VR", "B_LD", "B_VL", "B_HD", "B_HU"]


def main():
    src = open(os.path.join(ROOT, "lilliput_b200", "csrc", "vp8_core.h")).read()
    body = src[src.index("LP_VP8_FN void pred_4x4("):src.index("#undef LP_DST")]
    tab = {m: [None] * 16 for m in MODES}
    tab["B_DC"] = tab["B_TM"] = [3 << 12] * 16
    top = "XABCDE"
    left = "XIJKLL"
    for x in range(4):
        for y in range(4):
            tab["B_VE"][y * 4 + x] = E[top[x]] | E[top[x + 1]] << 4 | E[top[x + 2]] << 8
            tab["B_HE"][y * 4 + x] = E[left[y]] | E[left[y + 1]] << 4 | E[left[y + 2]] << 8
    cases = re.split(r"case (B_\w+):|default:\s*// (B_\w+)", body)
    i = 1
    while i < len(cases):
        mode = cases[i] or cases[i + 1]
        text = cases[i + 2]
        i += 3
        if mode in ("B_DC", "B_TM", "B_VE", "B_HE"):
            continue
        for stmt in text.split(";"):
            dsts = re.findall(r"LP_DST\((\d), (\d)\)", stmt)
            if not dsts:
                continue
            m3 = re.search(r"LP_AVG3\((\w), (\w), (\w)\)", stmt)
            m2 = re.search(r"LP_AVG2\((\w), (\w)\)", stmt)
            if m3:
                v = E[m3.group(1)] | E[m3.group(2)] << 4 | E[m3.group(3)] << 8
            elif m2:
                v = E[m2.group(1)] | E[m2.group(2)] << 4 | 1 << 12
            else:
                v = E[re.search(r"\(uint8_t\)(\w)", stmt).group(1)] | 2 << 12
            for x, y in dsts:
                tab[mode][int(y) * 4 + int(x)] = v
    print("LP_VP8_TABLE uint16_t kVp8Pred4[10][16] = {")
    for m in MODES:
        assert all(v is not None for v in tab[m]), m
        print("    {" + ", ".join("0x%04x" % v for v in tab[m]) + "},  // " + m)
    print("};")


if __name__ == "__main__":
    main()
