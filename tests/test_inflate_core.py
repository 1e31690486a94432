"""CPU: the product's warp-parallel DEFLATE decoder (lilliput_b200/csrc/inflate_core.h -- the code png_inflate_kernel
runs, one warp per zlib stream) compiled for the host with its 32 lanes simulated by loops
(tests/native/inflate_sim.cpp), against zlib: every block type, compression level and strategy, small windows,
sync / full flush points, streams that end early, buffers that are too small, and bit-flipped streams (never a
crash; where zlib accepts a mutant the bytes agree).  The GPU suite (test_gpu_png.py, test_png_adam7.py) then shows
the real warp reproduces this simulation on the PNG corpus."""
import ctypes as C
import os
import subprocess
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("inflate_sim") / "libinflate_sim.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so,
                           os.path.join(ROOT, "tests", "native", "inflate_sim.cpp")])
    lib = C.CDLL(so)
    lib.lp_inflate_sim.argtypes = [C.c_char_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]

    def run(z, n, off=0):
        buf = np.zeros(n + 96, np.uint8)
        al = (-buf.ctypes.data) % 16 + off          # off = 0: the 16-byte vector flush; otherwise the byte path
        prod = C.c_uint32(0)
        rc = lib.lp_inflate_sim(z, len(z), buf.ctypes.data + al, n, C.byref(prod))
        return rc, prod.value, bytes(buf[al:al + prod.value])
    return run


def _datasets():
    rng = np.random.default_rng(1)
    yield "empty", b""
    yield "one", b"a"
    yield "zeros", bytes(100000)
    yield "noise", rng.integers(0, 256, 120000, dtype=np.uint8).tobytes()
    yield "noise6", np.clip(rng.normal(128, 6, 200000), 0, 255).astype(np.uint8).tobytes()
    yield "text", b"the quick brown fox jumps over the lazy dog " * 3000
    yield "small", np.clip(rng.normal(0, 2, 250000), -128, 127).astype(np.int8).tobytes()
    yield "mixed", bytes(50000) + rng.integers(0, 256, 70000, dtype=np.uint8).tobytes() + b"abc" * 30000
    yield "runs", b"".join(bytes([int(v)]) * int(k) for v, k in zip(rng.integers(0, 256, 2000), rng.integers(1, 600, 2000)))
    yield "rgba", np.dstack([np.clip(rng.normal(128, 6, (150, 300, 3)), 0, 255).astype(np.uint8),
                             np.full((150, 300, 1), 255, np.uint8)]).tobytes()


def test_inflate_matches_zlib(sim):
    for name, d in _datasets():
        for level in (0, 1, 6, 9):
            for strat in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_RLE, zlib.Z_HUFFMAN_ONLY, zlib.Z_FILTERED):
                for wbits in (15, 9):
                    co = zlib.compressobj(level, zlib.DEFLATED, wbits, 8, strat)
                    z = co.compress(d) + co.flush()
                    for off in (0, 3):
                        rc, n, out = sim(z, len(d), off)
                        assert rc == 0 and out == d, (name, level, strat, wbits, off, rc, n)


def test_inflate_flush_points_truncation_and_capacity(sim):
    rng = np.random.default_rng(2)
    d = np.clip(rng.normal(128, 6, 200000), 0, 255).astype(np.uint8).tobytes() + b"xyz" * 20000
    co = zlib.compressobj(6)
    z = b""
    for i in range(0, len(d), 7777):
        z += co.compress(d[i:i + 7777]) + co.flush(zlib.Z_SYNC_FLUSH if (i // 7777) % 2 else zlib.Z_FULL_FLUSH)
    z += co.flush()
    rc, n, out = sim(z, len(d))
    assert rc == 0 and out == d
    assert sim(z, len(d) - 5)[0] == -3                 # more data than the image has room for
    rc, n, out = sim(z, len(d) + 5)                    # fewer bytes than expected: the kernel wrapper reports that
    assert rc == 0 and n == len(d) and out == d
    assert sim(z[:len(z) // 2], len(d))[0] == -3       # the stream ends inside a block
    assert sim(b"\x78", 10)[0] == -3 and sim(b"\x79\x9c\x03\x00", 10)[0] == -3


def test_inflate_mutants_never_disagree_with_zlib(sim):
    rng = np.random.default_rng(3)
    d = np.clip(rng.normal(128, 9, 60000), 0, 255).astype(np.uint8).tobytes() + b"lilliput " * 3000
    z = zlib.compress(d, 6)
    for _ in range(400):
        zz = bytearray(z)
        for _ in range(int(rng.integers(1, 4))):
            zz[int(rng.integers(2, len(zz)))] ^= 1 << int(rng.integers(0, 8))
        rc, n, out = sim(bytes(zz), len(d))
        try:
            ref = zlib.decompressobj().decompress(bytes(zz))
        except zlib.error:
            continue
        if rc == 0:
            m = min(len(ref), len(out))
            assert out[:m] == ref[:m]
