#!/usr/bin/env python3
"""GPU probe: per-image PNG decode (opencv_decoder_read_data) of one 3840x2160 RGBA PNG written with zlib
level 6 (adaptive-ish filters), checked against the pixels and timed; plus the product-written PNG."""
import json, os, sys, time, zlib, struct
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lilliput_b200 import abi
from lilliput_b200.synth import synth_image

def png_bytes(img, level=6, ftype=1):
    h, w, c = img.shape
    raw = img.astype(np.int16)
    if ftype == 1:   # Sub
        f = raw.copy(); f[:, 1:] -= raw[:, :-1]
    elif ftype == 2: # Up
        f = raw.copy(); f[1:] -= raw[:-1]
    else:
        f = raw
    f = (f & 255).astype(np.uint8).reshape(h, w * c)
    rows = np.concatenate([np.full((h, 1), ftype, np.uint8), f], axis=1).tobytes()
    def ch(t, d): return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xffffffff)
    ct = {1: 0, 3: 2, 4: 6}[c]
    return b"\x89PNG\r\n\x1a\n" + ch(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, ct, 0, 0, 0)) + ch(b"IDAT", zlib.compress(rows, level)) + ch(b"IEND", b"")

lib = abi.load_cuda()
import os
cases = ((3840, 2160, 4),) if os.environ.get("PNG_PROBE_ONE") else ((3840, 2160, 4), (1920, 1080, 3))
for (w, h, c) in cases:
    img = synth_image(3, w, h, c)           # BGR(A)
    rgb = img[:, :, [2, 1, 0, 3]] if c == 4 else img[:, :, ::-1]
    for ft in ((1,) if os.environ.get("PNG_PROBE_ONE") else (1, 2, 0)):
        p = png_bytes(np.ascontiguousarray(rgb), 6, ft)
        out = lib.decode(p)
        ok = bool(np.array_equal(out, img))
        t = time.perf_counter()
        for _ in range(3): lib.decode(p)
        ms = (time.perf_counter() - t) / 3 * 1e3
        print(json.dumps({"png": f"{w}x{h}x{c}", "filter": ft, "bytes": len(p), "ok": ok, "ms_per_decode": round(ms, 2)}), flush=True)
if os.environ.get("LP_CUDA_LIB", "").endswith("_stats.so"):
    import ctypes as C
    st = (C.c_ulonglong * 16)()
    lib.l.lp_png_inflate_stats(st, 0)
    names = ["blocks", "windows", "rounds", "redecodes", "partial", "matches", "match_rounds", "", "clk_header_tables", "clk_load_window", "clk_passA", "clk_passB", "clk_passC", "clk_matches", "clk_flush", ""]
    print(json.dumps({n: int(v) for n, v in zip(names, st) if n}))
