"""CPU: what opencv_decoder_read_header reports for a PNG (width, height, pixel type, orientation, or a refusal) --
product (png_parse.cpp, host-only) against the live reference (libpng 1.6.47 under OpenCV's PngDecoder) on the
golden files and on seeded, structured mutants: IHDR fields rewritten (CRC fixed), chunks renamed, tRNS / PLTE
chunks of every length inserted in every position, files cut short anywhere.  Full agreement is asserted, refusals included.  (Chunk LENGTH
fields are left alone: a damaged length makes the reference spend seconds per file.)"""
import random
import struct
import zlib

import numpy as np

from lilliput_b200 import abi


def _hdr(lib, b):
    try:
        return lib.header(b)
    except abi.LilliputError:
        return None


def _chunk(t, body, crc=None):
    return struct.pack(">I", len(body)) + t + body + struct.pack(">I", zlib.crc32(t + body) if crc is None else crc)


def _fix_crc(b, at):
    n = struct.unpack(">I", b[at:at + 4])[0]
    b[at + 8 + n:at + 12 + n] = struct.pack(">I", zlib.crc32(bytes(b[at + 4:at + 8 + n])))


def _chunks(b):
    i, out = 8, []
    while i + 12 <= len(b):
        n = struct.unpack(">I", b[i:i + 4])[0]
        out.append((i, bytes(b[i + 4:i + 8]), n))
        i += 12 + n
    return out


def test_png_header_matches_the_reference(ref_lib, golden):
    product = abi.load_cuda()
    seeds = {k: golden[k].tobytes() for k in golden.files
             if k.startswith("png_") and golden[k].dtype == np.uint8 and golden[k].ndim == 1 and golden[k].size < 200000}
    assert len(seeds) >= 15
    for name, data in seeds.items():
        assert _hdr(product, data) == _hdr(ref_lib, data) is not None, name
    pool = sorted(seeds.items())
    rnd = random.Random(4)
    refused = taken = 0
    for it in range(6000):
        name, s = rnd.choice(pool)
        b = bytearray(s)
        mode = rnd.randrange(6)
        if mode == 5:        # the file simply ends somewhere
            b = b[:rnd.randrange(8, len(b))]
        elif mode == 0:      # bit depth / colour type / compression / filter / interlace
            f = rnd.choice([24, 25, 26, 27, 28])
            b[f] = rnd.choice([0, 1, 2, 3, 4, 5, 6, 7, 8, 16, 32]) if f < 26 else rnd.choice([0, 1, 2])
            _fix_crc(b, 8)
        elif mode == 1:      # an ancillary / palette chunk becomes another
            cands = [c for c in _chunks(b) if c[1] in (b"tRNS", b"PLTE", b"gAMA", b"sBIT", b"bKGD", b"pHYs")]
            if cands:
                at = rnd.choice(cands)[0]
                b[at + 4:at + 8] = rnd.choice([b"tRNS", b"PLTE", b"bKGD", b"sBIT", b"hIST", b"IDAT"])
                _fix_crc(b, at)
        elif mode == 2:      # size
            b[16:20] = struct.pack(">I", rnd.choice([0, 1, 2, 7, 33, 1000, 1000000, 1000001]))
            if rnd.random() < 0.5:
                b[20:24] = struct.pack(">I", rnd.choice([0, 1, 3, 64, 1000001]))
            _fix_crc(b, 8)
        elif mode == 3:      # a tRNS or PLTE chunk of some length dropped in front of / behind some chunk
            kind = rnd.choice([b"tRNS", b"tRNS", b"PLTE"])
            n = rnd.choice([0, 1, 2, 3, 5, 6, 7, 9, 12, 48, 255, 256, 257, 768, 771])
            body = bytes(rnd.randrange(256) for _ in range(n))
            new = _chunk(kind, body, crc=rnd.choice([None, None, None, 12345]))
            at = rnd.choice(_chunks(b)[1:])[0]
            b[at:at] = new
        else:                # a critical chunk loses its CRC, or an unknown chunk appears
            if rnd.random() < 0.5:
                at = rnd.choice([c for c in _chunks(b) if c[1] in (b"IHDR", b"PLTE", b"tRNS", b"gAMA")])[0]
                n = struct.unpack(">I", b[at:at + 4])[0]
                b[at + 8 + n] ^= 0x55
            else:
                at = rnd.choice(_chunks(b)[1:])[0]
                b[at:at] = _chunk(rnd.choice([b"zzZz", b"ZZZZ", b"prVt", b"gAvA", b"IHDR", b"IEND"]), b"\x01\x02")
        p, r = _hdr(product, bytes(b)), _hdr(ref_lib, bytes(b))
        assert p == r, (it, mode, name)
        taken += r is not None
        refused += r is None
    assert taken > 1500 and refused > 1000
