"""ctypes binding of oracle/liboracle.so (the C restatement).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg; never by the product path (lilliput_b200/).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
INTER_LINEAR, INTER_AREA = 1, 3


def build(force: bool = False) -> None:
    srcs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(LIB) or any(
            os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        l = C.CDLL(LIB)
        l.oracle_resize.restype = C.c_int
        l.oracle_resize.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int]
        l.oracle_jpeg_decode.restype = C.c_int
        l.oracle_jpeg_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t] + \
            [C.POINTER(C.c_int)] * 4
        l.oracle_jpeg_encode.restype = C.c_size_t
        l.oracle_jpeg_encode.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                         C.c_int, C.c_void_p, C.c_size_t]
        l.oracle_png_decode.restype = C.c_int
        l.oracle_png_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t] + \
            [C.POINTER(C.c_int)] * 4
        l.oracle_orient.restype = C.c_int
        l.oracle_orient.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                    C.POINTER(C.c_int), C.POINTER(C.c_int)]
        l.oracle_blend_over.restype = C.c_int
        l.oracle_blend_over.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t,
                                        C.c_int, C.c_int, C.c_int]
        l.oracle_fit_rect.argtypes = [C.c_int] * 4 + [C.POINTER(C.c_int)] * 4
        l.oracle_expected_size.argtypes = [C.c_int] * 4 + [C.POINTER(C.c_int)] * 2
        l.oracle_area_taps.restype = C.c_int
        l.oracle_area_taps.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_int]
        _lib = l
    return _lib


def _cn(img):
    return 1 if img.ndim == 2 else img.shape[2]


def resize(img: np.ndarray, w: int, h: int, crop=None, interpolation=INTER_AREA) -> np.ndarray:
    img = np.ascontiguousarray(img, dtype=np.uint8)
    cn = _cn(img)
    cx, cy, cw, ch = crop if crop else (0, 0, img.shape[1], img.shape[0])
    dst = np.empty((h, w, cn) if cn > 1 else (h, w), dtype=np.uint8)
    rc = lib().oracle_resize(img.ctypes.data, img.shape[1] * cn, cn, cx, cy, cw, ch,
                             dst.ctypes.data, w * cn, w, h, interpolation)
    if rc != 0:
        raise RuntimeError(f"oracle_resize rc={rc}")
    return dst


def fit_rect(sw, sh, dw, dh):
    a = [C.c_int() for _ in range(4)]
    lib().oracle_fit_rect(sw, sh, dw, dh, *[C.byref(x) for x in a])
    return tuple(x.value for x in a)


def expected_size(ow, oh, rw, rh):
    a, b = C.c_int(), C.c_int()
    lib().oracle_expected_size(ow, oh, rw, rh, C.byref(a), C.byref(b))
    return a.value, b.value


def fit(img: np.ndarray, w: int, h: int) -> np.ndarray:
    """Framebuffer.Fit (ref opencv.go:326-374)."""
    l, t, wc, hc = fit_rect(img.shape[1], img.shape[0], w, h)
    return resize(img, w, h, crop=(l, t, wc, hc))


def jpeg_decode(data: bytes):
    src = np.frombuffer(data, dtype=np.uint8)
    w, h, c, o = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    rc = lib().oracle_jpeg_decode(src.ctypes.data, src.size, None, 0, C.byref(w), C.byref(h),
                                  C.byref(c), C.byref(o))
    if rc != 0:
        raise RuntimeError(f"oracle_jpeg_decode header rc={rc}")
    out = np.empty(h.value * w.value * c.value, dtype=np.uint8)
    rc = lib().oracle_jpeg_decode(src.ctypes.data, src.size, out.ctypes.data, out.size,
                                  C.byref(w), C.byref(h), C.byref(c), C.byref(o))
    if rc != 0:
        raise RuntimeError(f"oracle_jpeg_decode rc={rc}")
    shape = (h.value, w.value, c.value) if c.value > 1 else (h.value, w.value)
    return out.reshape(shape), o.value


def png_decode(data: bytes):
    src = np.frombuffer(data, dtype=np.uint8)
    w, h, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    rc = lib().oracle_png_decode(src.ctypes.data, src.size, None, 0, C.byref(w), C.byref(h),
                                 C.byref(c), C.byref(d))
    if rc != 0:
        raise RuntimeError(f"oracle_png_decode header rc={rc}")
    out = np.empty(h.value * w.value * c.value, dtype=np.uint8)
    rc = lib().oracle_png_decode(src.ctypes.data, src.size, out.ctypes.data, out.size,
                                 C.byref(w), C.byref(h), C.byref(c), C.byref(d))
    if rc != 0:
        raise RuntimeError(f"oracle_png_decode rc={rc}")
    shape = (h.value, w.value, c.value) if c.value > 1 else (h.value, w.value)
    return out.reshape(shape)


def jpeg_encode(img: np.ndarray, quality: int = 95) -> bytes:
    img = np.ascontiguousarray(img, dtype=np.uint8)
    cn = _cn(img)
    cap = img.size * 2 + (1 << 16)
    out = np.empty(cap, dtype=np.uint8)
    n = lib().oracle_jpeg_encode(img.ctypes.data, img.shape[1] * cn, img.shape[1], img.shape[0],
                                 cn, quality, out.ctypes.data, cap)
    if n == 0:
        raise RuntimeError("oracle_jpeg_encode failed")
    return out[:n].tobytes()


def orient(img: np.ndarray, orientation: int) -> np.ndarray:
    img = np.ascontiguousarray(img, dtype=np.uint8)
    cn = _cn(img)
    dst = np.empty(img.size, dtype=np.uint8)
    ow, oh = C.c_int(), C.c_int()
    lib().oracle_orient(img.ctypes.data, img.shape[1], img.shape[0], cn, orientation,
                        dst.ctypes.data, C.byref(ow), C.byref(oh))
    return dst.reshape((oh.value, ow.value, cn) if cn > 1 else (oh.value, ow.value))


def blend_over(src: np.ndarray, dst: np.ndarray) -> np.ndarray:
    src = np.ascontiguousarray(src, dtype=np.uint8)
    out = np.array(dst, dtype=np.uint8, order="C", copy=True)
    rc = lib().oracle_blend_over(src.ctypes.data, src.shape[1] * src.shape[2], src.shape[2],
                                 out.ctypes.data, out.shape[1] * out.shape[2], out.shape[2],
                                 out.shape[1], out.shape[0])
    if rc != 0:
        raise RuntimeError("oracle_blend_over rc=%d" % rc)
    return out


# --- GIF (oracle_gif.c): giflib record walk + LZW + the reference's compositor, and its encoder ----
def _gif_sigs():
    l = lib()
    l.oracle_gif_open.restype = C.c_void_p
    l.oracle_gif_open.argtypes = [C.c_void_p, C.c_size_t]
    l.oracle_gif_close.argtypes = [C.c_void_p]
    l.oracle_gif_width.argtypes = [C.c_void_p]
    l.oracle_gif_height.argtypes = [C.c_void_p]
    l.oracle_gif_next_frame.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    l.oracle_gif_skip_frame.argtypes = [C.c_void_p]
    l.oracle_gif_enc_open.restype = C.c_void_p
    l.oracle_gif_enc_open.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    l.oracle_gif_enc_close.argtypes = [C.c_void_p]
    l.oracle_gif_enc_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    l.oracle_gif_enc_finish.restype = C.c_size_t
    l.oracle_gif_enc_finish.argtypes = [C.c_void_p, C.c_void_p]
    return l


def gif_frames(data: bytes, max_frames: int = 1 << 16):
    """Every composited full-canvas BGRA frame: (frames, delays in 1/100 s, giflib disposal modes, rc)."""
    l = _gif_sigs()
    src = np.frombuffer(data, dtype=np.uint8)
    d = l.oracle_gif_open(src.ctypes.data, src.size)
    if not d:
        raise RuntimeError("oracle_gif_open failed")
    try:
        w, h = l.oracle_gif_width(d), l.oracle_gif_height(d)
        canvas = np.zeros((h, w, 4), dtype=np.uint8)
        frames, delays, disps, rc = [], [], [], 0
        for _ in range(max_frames):
            dl, dp = C.c_int(), C.c_int()
            r = l.oracle_gif_next_frame(d, canvas.ctypes.data, C.byref(dl), C.byref(dp))
            if r <= 0:
                rc = r
                break
            frames.append(canvas.copy())
            delays.append(dl.value)
            disps.append(dp.value)
        return frames, delays, disps, rc
    finally:
        l.oracle_gif_close(d)


def gif_transcode(data: bytes, per_frame=None, max_frames: int = 0, cap: int = 8 << 20) -> bytes:
    """GIF -> GIF as ImageOps.Transform drives it: decode + composite each frame, per_frame(BGRA) (fit /
    resize; identity when None), encode with the reference's palette mapping and giflib's LZW."""
    l = _gif_sigs()
    src = np.frombuffer(data, dtype=np.uint8)
    d = l.oracle_gif_open(src.ctypes.data, src.size)
    if not d:
        raise RuntimeError("oracle_gif_open failed")
    dst = np.zeros(cap, dtype=np.uint8)
    e = None
    try:
        w, h = l.oracle_gif_width(d), l.oracle_gif_height(d)
        canvas = np.zeros((h, w, 4), dtype=np.uint8)
        n = 0
        while True:
            if max_frames and n >= max_frames:
                # ops.go:384-390: past the frame cap the remaining frames are skipped (decoded to the end)
                while l.oracle_gif_skip_frame(d) > 0:  # gifDecoder.SkipFrame until EOF
                    pass
                break
            r = l.oracle_gif_next_frame(d, canvas.ctypes.data, None, None)
            if r < 0:
                raise RuntimeError("oracle gif decode failed")
            if r == 0:
                break
            f = np.ascontiguousarray(per_frame(canvas) if per_frame else canvas)
            if e is None:
                e = l.oracle_gif_enc_open(d, f.shape[1], f.shape[0], dst.ctypes.data, dst.size)
            if not l.oracle_gif_enc_frame(e, d, f.ctypes.data, f.shape[1], f.shape[0]):
                raise RuntimeError("oracle gif encode failed")
            n += 1
        size = l.oracle_gif_enc_finish(e, d) if e else 0
        return dst[:size].tobytes()
    finally:
        if e:
            l.oracle_gif_enc_close(e)
        l.oracle_gif_close(d)
