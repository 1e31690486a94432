"""CPU: GIF animation metadata (giflib_decoder_get_animation_info behind lp_gif_get_info; host-only in both
libraries) -- product against the live reference (giflib 5.2.2 driven by giflib.cpp:1308-1431) on the golden GIFs,
on hand-built files without any graphics control block, and on seeded mutants.  Loop count, frame count, total
duration and the background colour have to agree whenever the file holds no malformed graphics control block: for
a GCB whose sub-block is not 4 bytes long the reference reads an uninitialised struct (DGifExtensionToGCB fails and
its result is not checked), which no re-implementation can follow."""
import random

import numpy as np
import pytest

from lilliput_b200 import abi


def _info(lib, b):
    try:
        return lib.gif_info(b)
    except abi.LilliputError:
        return None


def _has_malformed_gcb(b):
    """Walks the blocks the way the prescan does; True if some 0x21 0xF9 extension's first sub-block is not 4 long."""
    if len(b) < 13:
        return False
    i = 13 + (3 * (2 << (b[10] & 7)) if b[10] & 0x80 else 0)
    while i < len(b):
        c = b[i]
        if c == 0x21 and i + 2 < len(b):
            if b[i + 1] == 0xF9 and b[i + 2] != 4:
                return True
            i += 2
            while i < len(b) and b[i] != 0:
                i += 1 + b[i]
            i += 1
        elif c == 0x2C and i + 9 < len(b):
            p = b[i + 9]
            i += 10 + (3 * (2 << (p & 7)) if p & 0x80 else 0) + 1
            while i < len(b) and b[i] != 0:
                i += 1 + b[i]
            i += 1
        elif c == 0x3B:
            return False
        else:
            i += 1
    return False


def _gif(gcb=None, trailer=True, bg=1, gct=True, frames=1):
    """A 2x2 GIF89a built by hand: optional global colour table (4 entries), optional GCB per frame."""
    out = b"GIF89a" + bytes([2, 0, 2, 0, (0x81 if gct else 0x00), bg, 0])
    if gct:
        out += bytes([10, 20, 30, 200, 100, 50, 1, 2, 3, 250, 251, 252])
    for _ in range(frames):
        if gcb is not None:
            out += b"\x21\xf9\x04" + gcb + b"\x00"
        out += b"\x2c" + bytes([0, 0, 0, 0, 2, 0, 2, 0, 0]) + b"\x02\x03\x84\x8f\x05\x00"   # 2-bit LZW data
    return out + (b"\x3b" if trailer else b"")


def test_gif_info_matches_the_reference(ref_lib, golden):
    product = abi.load_cuda()
    cases = {k: golden[k].tobytes() for k in golden.files
             if k.startswith("gif_") and golden[k].dtype == np.uint8 and golden[k].ndim == 1}
    cases.update({
        "no_gcb_with_trailer": _gif(),                       # terminator: defaults stay (white, alpha 0)
        "no_gcb_no_trailer": _gif(trailer=False),            # walk ends on a read error: colour table entry, alpha 0
        "no_gcb_no_trailer_no_table": _gif(trailer=False, gct=False),
        "no_gcb_bg_outside_table": _gif(trailer=False, bg=9),
        "opaque_gcb": _gif(gcb=bytes([0, 5, 0, 0])),
        "transparent_gcb": _gif(gcb=bytes([1, 5, 0, 2])),
        "transparent_gcb_bg_outside_table": _gif(gcb=bytes([1, 5, 0, 2]), bg=200),
        "three_frames_short_delays": _gif(gcb=bytes([0, 1, 0, 0]), frames=3),
        "three_frames_no_gcb": _gif(frames=3),
        "header_only": _gif()[:13],
        "cut_inside_colour_table": _gif()[:20],
    })
    for name, data in cases.items():
        assert _info(product, data) == _info(ref_lib, data), name
    assert _info(product, cases["no_gcb_with_trailer"])["background_color"] == 0x00FFFFFF
    assert _info(product, cases["no_gcb_no_trailer"])["background_color"] == 0x00C86432     # entry 1, alpha 0
    assert _info(product, cases["opaque_gcb"])["background_color"] == 0xFFC86432
    assert _info(product, cases["three_frames_short_delays"])["duration_ms"] == 10 + 20 + 20

    seeds = [v for v in cases.values() if len(v) < 120000]
    rnd = random.Random(1)
    compared = 0
    for it in range(4000):
        b = bytearray(rnd.choice(seeds))
        mode = rnd.randrange(3)
        if mode == 0:
            for _ in range(rnd.randrange(1, 4)):
                b[rnd.randrange(0, min(len(b), 900))] = rnd.randrange(256)
        elif mode == 1:
            b = b[:rnd.randrange(6, len(b))]
        else:
            for _ in range(rnd.randrange(1, 4)):
                b[rnd.randrange(len(b))] = rnd.randrange(256)
        b = bytes(b)
        if _has_malformed_gcb(b):
            continue
        if len(b) >= 10 and (b[6] | b[7] << 8) * (b[8] | b[9] << 8) * 4 > 1 << 28:
            continue    # the reference allocates a canvas-sized scratch at create: whether that succeeds is the host's business
        compared += 1
        assert _info(product, b) == _info(ref_lib, b), (it, mode)
    assert compared > 3000
