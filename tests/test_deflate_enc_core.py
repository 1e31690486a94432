"""CPU: the DEFLATE writer of the PNG encoder (lilliput_b200/csrc/deflate_enc_core.h: hash-chain LZ77, dynamic Huffman
blocks, independent 32 KB chunks joined by sync flushes), compiled for the host by tests/native/deflate_enc_sim.cpp.

It replaces what libpng gets from zlib behind opencv_encoder_write(".png") (ref opencv.cpp:173-194); the contract is
lossless-ness, checked here with zlib's own inflate on every kind of input, and a size close to zlib's at the same
PngCompression level, checked on filtered scanlines of PNG files written by the reference's libpng (via cv2)."""
import ctypes as C
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

from lilliput_b200.synth import synth_image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def defenc(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("defenc") / "libdefenc.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so,
                           os.path.join(ROOT, "tests", "native", "deflate_enc_sim.cpp")])
    l = C.CDLL(so)
    l.defenc_compress.restype = C.c_long
    l.defenc_compress.argtypes = [C.c_char_p, C.c_long, C.c_int, C.c_void_p, C.c_long]

    def compress(data: bytes, level: int = 6) -> bytes:
        cap = len(data) + len(data) // 1000 * 70 + 1024
        out = (C.c_uint8 * cap)()
        n = l.defenc_compress(data, len(data), level, out, cap)
        assert n > 0
        return bytes(out[:n])
    return compress


def _inputs():
    rng = np.random.default_rng(1)
    yield "empty", b""
    yield "one byte", b"a"
    yield "three bytes", b"abc"
    yield "zeros", bytes(100000)
    yield "random", rng.integers(0, 256, 70000, dtype=np.uint8).tobytes()
    yield "text", b"the quick brown fox jumps over the lazy dog " * 3000
    yield "two symbols", rng.integers(0, 2, 50000, dtype=np.uint8).tobytes()
    yield "one symbol then noise", bytes(40000) + rng.integers(0, 256, 3000, dtype=np.uint8).tobytes()
    yield "geometric", np.clip(rng.geometric(0.15, 120000) - 1, 0, 255).astype(np.uint8).tobytes()
    yield "residuals", (rng.normal(0, 6, 200000).round().astype(np.int64) & 255).astype(np.uint8).tobytes()
    yield "chunk + 1", bytes(32768) + b"x"
    yield "exactly a chunk", rng.integers(0, 3, 32768, dtype=np.uint8).tobytes()
    yield "long matches", (bytes(range(256)) * 4 + b"#") * 120
    # a skewed alphabet whose plain Huffman tree is deeper than 15 levels: Fibonacci-like counts
    fib = [1, 1]
    while len(fib) < 24:
        fib.append(fib[-1] + fib[-2])
    yield "deep tree", bytes(np.repeat(np.arange(24, dtype=np.uint8), np.minimum(fib, 30000)))
    yield "all byte values once", bytes(range(256))


@pytest.mark.parametrize("level", [0, 1, 3, 6, 9])
def test_every_stream_inflates_with_zlib(defenc, level):
    for name, data in _inputs():
        z = defenc(data, level)
        assert zlib.decompress(z) == data, (name, level)
        if level == 0:
            assert len(z) <= len(data) + 5 * (len(data) // 32768 + 1) + 8
        else:
            assert len(z) <= len(data) + len(data) // 1000 + 64, (name, len(z), len(data))  # never expands past stored blocks


def test_random_lengths_and_contents(defenc):
    rng = np.random.default_rng(2)
    for t in range(60):
        n = int(rng.integers(0, 90000))
        kind = t % 4
        if kind == 0:
            d = rng.integers(0, 256, n, dtype=np.uint8)
        elif kind == 1:
            d = rng.integers(0, 5, n, dtype=np.uint8)
        elif kind == 2:
            d = (rng.normal(0, 3, n).round().astype(np.int64) & 255).astype(np.uint8)
        else:
            d = np.tile(rng.integers(0, 256, max(1, n // 50), dtype=np.uint8), 51)[:n]
        data = d.tobytes()
        assert zlib.decompress(defenc(data, int(rng.integers(1, 10)))) == data


def _idat(png: bytes) -> bytes:
    o, z = 8, b""
    while o < len(png):
        ln, = struct.unpack(">I", png[o:o + 4])
        if png[o + 4:o + 8] == b"IDAT":
            z += png[o + 8:o + 8 + ln]
        o += 12 + ln
    return z


def test_size_against_libpng_at_the_same_level(defenc):
    """Filtered scanlines as the reference's libpng produces them (adaptive filter heuristic), compressed by zlib there
    and by this writer here: within 1.08 x of zlib's size at levels 1, 3 and 6 (measured 0.86-1.04) and 1.12 x at level 9
    (independent 32 KB chunks and a 4-byte minimum match cost most where zlib's longest searches pay: 1.10 on noise-free
    synthetic content)."""
    cv2 = pytest.importorskip("cv2")
    for seed, w, h, ch, noise in [(21, 512, 512, 3, 6.0), (54, 512, 512, 3, 0.0), (6, 300, 200, 4, 10.0), (7, 640, 480, 1, 2.0)]:
        img = synth_image(seed, w, h, ch, noise=noise)
        for level in (1, 3, 6, 9):
            ok, png = cv2.imencode(".png", img, [cv2.IMWRITE_PNG_COMPRESSION, level])
            assert ok
            z = _idat(png.tobytes())
            raw = zlib.decompress(z)
            mine = defenc(raw, level)
            assert zlib.decompress(mine) == raw
            assert len(mine) <= (1.12 if level == 9 else 1.08) * len(z), (seed, level, len(mine), len(z))


def test_higher_levels_search_harder(defenc):
    data = (b"abcdefgh" * 7 + b"0123456789" * 3 + bytes(range(200))) * 300
    sizes = [len(defenc(data, lv)) for lv in (1, 3, 6, 9)]
    assert sizes[0] >= sizes[1] >= sizes[2] >= sizes[3] and sizes[3] < sizes[0]
