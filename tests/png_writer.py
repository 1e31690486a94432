"""A small PNG writer for tests (test infrastructure): any colour type / bit depth, optional Adam7
interlacing, per-row filter types chosen by the caller -- things Pillow / OpenCV cannot write."""
import struct
import zlib

import numpy as np

PASSES = [(0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)]


def _chunk(tag, data):
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def _pack_rows(samples, bit_depth):
    """samples: [h, w*channels] integer array -> list of packed byte rows."""
    h, n = samples.shape
    if bit_depth == 8:
        return [samples[y].astype(np.uint8).tobytes() for y in range(h)]
    if bit_depth == 16:
        return [samples[y].astype(">u2").tobytes() for y in range(h)]
    rows = []
    per = 8 // bit_depth
    for y in range(h):
        pad = (-n) % per
        v = np.concatenate([samples[y], np.zeros(pad, samples.dtype)]).reshape(-1, per).astype(np.uint32)
        acc = np.zeros(v.shape[0], np.uint32)
        for k in range(per):
            acc |= v[:, k] << (8 - bit_depth * (k + 1))
        rows.append(acc.astype(np.uint8).tobytes())
    return rows


def _filter(rows, bpp, ftypes):
    out = bytearray()
    prev = bytes(len(rows[0])) if rows else b""
    for y, row in enumerate(rows):
        f = ftypes[y % len(ftypes)]
        cur = bytearray(len(row))
        for x in range(len(row)):
            a = row[x - bpp] if x >= bpp else 0
            b = prev[x]
            c = prev[x - bpp] if x >= bpp else 0
            if f == 0:
                p = 0
            elif f == 1:
                p = a
            elif f == 2:
                p = b
            elif f == 3:
                p = (a + b) >> 1
            else:
                pa, pb, pc = abs(b - c), abs(a - c), abs(a + b - 2 * c)
                p = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
            cur[x] = (row[x] - p) & 255
        out.append(f)
        out += cur
        prev = row
    return bytes(out)


def write_png(samples, color_type, bit_depth, interlace=False, ftypes=(0, 1, 2, 3, 4), palette=None, trns=None, level=6):
    """samples: [h, w, channels] integers at full precision (palette images: indices, channels = 1)."""
    samples = np.asarray(samples)
    h, w, ch = samples.shape
    bits = ch * bit_depth
    bpp = max(1, bits // 8)
    raw = b""
    for (x0, y0, dx, dy) in (PASSES if interlace else [(0, 0, 1, 1)]):
        sub = samples[y0::dy, x0::dx]
        if sub.shape[0] == 0 or sub.shape[1] == 0:
            continue
        rows = _pack_rows(sub.reshape(sub.shape[0], -1), bit_depth)
        raw += _filter(rows, bpp, ftypes)
    out = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, bit_depth, color_type, 0, 0, int(interlace)))
    if palette is not None:
        out += _chunk(b"PLTE", bytes(np.asarray(palette, np.uint8).reshape(-1)))
    if trns is not None:
        out += _chunk(b"tRNS", bytes(trns))
    z = zlib.compress(raw, level)
    half = len(z) // 2
    out += _chunk(b"IDAT", z[:half]) + _chunk(b"IDAT", z[half:]) + _chunk(b"IEND", b"")
    return out
