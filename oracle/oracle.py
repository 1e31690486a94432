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
INTER_LINEAR, INTER_CUBIC, INTER_AREA = 1, 2, 3


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



# ---- HDR tone map (numpy restatement; TEST INFRASTRUCTURE) -----------------------------------------------------
# Follows the reference's tonemap_rgb_8u_inplace -> tonemap_rgb_to_sdr (color_info.cpp:112-270): u8 / 255 -> PQ
# (color_info.cpp:82-95) or HLG (:98-110) EOTF -> cv::TonemapReinhard(1.0, 0.6, 0.2, 0.3) -> primaries matrix
# (:160-197) -> x 255, round, saturate.  cv::TonemapReinhard lives in the vendored OpenCV 4.11 photo module (binary
# only): its published algorithm (min-max normalise, grey / log statistics, key -> map_key, intensity = exp(-i),
# per-channel adaptation, second normalisation) is restated here and PINNED on oracle/_ref, which now links the
# reference's own color_info.cpp: tests/test_oracle_tonemap.py finds at most +-1 LSB on ~0.01 % of the samples (the
# reference's SIMD evaluation order in the last fp32 ulp), on fresh inputs and on tests/golden/tonemap_golden.npz.
_TM_MATS = {9: [1.6605, -0.5876, -0.0728, -0.1246, 1.1329, -0.0083, -0.0182, -0.1006, 1.1187],
            12: [1.2249, -0.2247, -0.0002, -0.0420, 1.0419, 0.0001, -0.0197, 0.0754, 0.9443],
            11: [1.2249, -0.2247, -0.0002, -0.0420, 1.0419, 0.0001, -0.0197, 0.0754, 0.9443],
            6: [1.0440, -0.0440, 0.0, -0.0, 1.0, 0.0, 0.0, 0.0, 1.0],
            10: [1.0569715, -0.2039770, 0.0556301, 0.0415551, 1.8759675, -0.9692436, -0.4986108, -1.5373832, 3.2409699]}


def tonemap_to_sdr(img: np.ndarray, transfer: int, primaries: int) -> np.ndarray:
    f32 = np.float32

    def norm(x):
        mn, mx = float(x.min()), float(x.max())
        if mx - mn > 2.220446049250313e-16:
            alpha = 1.0 / (mx - mn)
            return (x * f32(alpha) + f32(-mn * alpha)).astype(f32)
        return x.copy()
    x = img[:, :, :3].astype(f32) * f32(1.0 / 255)
    if transfer == 16:
        xp = np.power(x, f32(1.0) / f32(78.84375), dtype=f32)
        lin = np.power(np.maximum(xp - f32(0.8359375), f32(0)) / (f32(18.8515625) - f32(18.6875) * xp),
                       f32(1.0) / f32(0.1593017578125), dtype=f32)
    elif transfer == 18:
        lin = np.where(x <= f32(0.5), x * x / f32(3.0),
                       (np.exp((x - f32(0.55991073)) / f32(0.17883277), dtype=f32) + f32(0.28466892)) / f32(12.0)).astype(f32)
    else:
        lin = x
    im = norm(lin)
    gray = (im[:, :, 0] * f32(0.299) + im[:, :, 1] * f32(0.587) + im[:, :, 2] * f32(0.114)).astype(f32)
    logi = np.log(np.maximum(gray, f32(1e-4)), dtype=f32)
    log_mean = f32(logi.astype(np.float64).sum() / logi.size)
    log_min, log_max = float(logi.min()), float(logi.max())
    with np.errstate(invalid="ignore", divide="ignore"):
        key = f32(np.float64(log_max - float(log_mean)) / np.float64(log_max - log_min))  # 0/0 = NaN on a flat frame, as in C++
    map_key = f32(0.3) + f32(0.7) * f32(np.power(key, f32(1.4)))
    inten = f32(np.exp(-f32(0.6)))
    ca, la = f32(0.3), f32(0.2)
    gray_mean = f32(gray.astype(np.float64).mean())
    out = np.empty_like(im)
    for i in range(3):
        glob = ca * f32(im[:, :, i].astype(np.float64).mean()) + (f32(1) - ca) * gray_mean
        adapt = la * (ca * im[:, :, i] + (f32(1) - ca) * gray) + (f32(1) - la) * glob
        with np.errstate(invalid="ignore"):
            adapt = np.power(inten * adapt, map_key, dtype=f32)
        out[:, :, i] = im[:, :, i] * (f32(1.0) / (adapt + im[:, :, i]))
    out = norm(out)
    if primaries in _TM_MATS:
        m = np.array(_TM_MATS[primaries], f32).reshape(3, 3)
        out = (out.reshape(-1, 3) @ m.T).reshape(out.shape).astype(f32)
    if transfer == 8:
        with np.errstate(invalid="ignore"):
            out = np.power(out, f32(1.0 / 2.2), dtype=f32)
    res = np.clip(np.rint(np.nan_to_num(out * f32(255.0), nan=0.0)), 0, 255).astype(np.uint8)
    r = np.ascontiguousarray(img).copy()
    r[:, :, :3] = res
    return r
