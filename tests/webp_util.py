"""Helpers shared by the WebP tests (test infrastructure)."""
import ctypes
import os
import struct
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_BUILD = os.path.join(ROOT, "tests", "native", "_build")


def webp_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "webp_golden.npz"))


def chunks_of(webp: bytes):
    """[(tag, payload)] of a RIFF/WEBP file's top-level chunks."""
    pos, out = 12, []
    while pos + 8 <= len(webp):
        n = struct.unpack("<I", webp[pos + 4:pos + 8])[0]
        out.append((webp[pos:pos + 4], webp[pos + 8:pos + 8 + n]))
        pos += 8 + n + (n & 1)
    return out


def vp8_cpu_lib():
    """tests/native/vp8_cpu.cpp (the device's VP8 logic compiled for the host) as a ctypes handle."""
    os.makedirs(_BUILD, exist_ok=True)
    so = os.path.join(_BUILD, "libvp8cpu.so")
    srcs = [os.path.join(ROOT, "tests", "native", "vp8_cpu.cpp"),
            os.path.join(ROOT, "lilliput_b200", "csrc", "vp8_core.h"),
            os.path.join(ROOT, "lilliput_b200", "csrc", "vp8_tables.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", so, srcs[0]])
    return ctypes.CDLL(so)


def vp8_cpu_decode(lib, payload: bytes) -> np.ndarray:
    arr = np.frombuffer(payload, np.uint8)
    w, h = ctypes.c_int(), ctypes.c_int()
    p = arr.ctypes.data_as(ctypes.c_void_p)
    assert lib.vp8_cpu_info(p, ctypes.c_size_t(arr.size), ctypes.byref(w), ctypes.byref(h)) == 0
    out = np.zeros((h.value, w.value, 3), np.uint8)
    rc = lib.vp8_cpu_decode_bgr(p, ctypes.c_size_t(arr.size), out.ctypes.data_as(ctypes.c_void_p), w.value * 3, 0, None)
    assert rc == 0, rc
    return out
