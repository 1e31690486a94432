"""Generates tests/golden/webp_golden.npz from the REFERENCE ITSELF: WebP streams made with the
libwebp the reference vendors (tests/native/webp_variants.c sweeps the encoder configuration so the
VP8 streams cover simple/normal loop filters, sharpness, 1..8 token partitions, 1..4 segments ...),
hand-assembled RIFF containers (VP8X / ICCP / ANIM / ANMF), a few of the reference's own test
fixtures, and -- for every stream -- what lilliput's webp_decoder_* returns for it through
oracle/_ref (frame pixels as SHA-256, frame metadata, container info, status).

Run in the build container:  python tests/golden/make_golden_webp.py
"""
import ctypes
import hashlib
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lilliput_b200 import abi  # noqa: E402
from lilliput_b200.synth import synth_image  # noqa: E402

DEPS = "/root/reference/deps/linux/amd64"
TESTDATA = "/root/reference/testdata"


def build_variants():
    so = os.path.join(tempfile.mkdtemp(), "libwebpvar.so")
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tests/native/webp_variants.c"),
                           f"-I{DEPS}/include", f"{DEPS}/lib/libwebp.a", f"{DEPS}/lib/libsharpyuv.a", "-lm", "-lpthread"])
    return ctypes.CDLL(so)


VAR = build_variants()


def enc(img, q=75, method=4, ft=1, fs=60, sharp=0, parts=0, segs=4, sns=50, ac=1, lossless=0):
    img = np.ascontiguousarray(img)
    h, w, c = img.shape
    out = np.zeros(w * h * 4 + 8192, np.uint8)
    n = VAR.lpv_encode(img.ctypes.data_as(ctypes.c_void_p), w, h, w * c, int(c == 4), ctypes.c_float(q), method, ft, fs,
                       sharp, parts, segs, sns, ac, lossless, out.ctypes.data_as(ctypes.c_void_p),
                       ctypes.c_size_t(len(out)))
    assert n > 0, n
    return out[:n].tobytes()


def chunk(tag, payload):
    return tag + struct.pack("<I", len(payload)) + payload + (b"\0" if len(payload) & 1 else b"")


def riff(chunks):
    body = b"WEBP" + b"".join(chunks)
    return b"RIFF" + struct.pack("<I", len(body)) + body


def chunks_of(webp):
    pos, out = 12, []
    while pos + 8 <= len(webp):
        n = struct.unpack("<I", webp[pos + 4:pos + 8])[0]
        out.append((webp[pos:pos + 4], webp[pos + 8:pos + 8 + n]))
        pos += 8 + n + (n & 1)
    return out


def le24(v):
    return struct.pack("<I", v)[:3]


def vp8x(flags, w, h):
    return chunk(b"VP8X", bytes([flags, 0, 0, 0]) + le24(w - 1) + le24(h - 1))


def anmf(x, y, w, h, dur, dispose, blend_off, image_chunks):
    hdr = le24(x // 2) + le24(y // 2) + le24(w - 1) + le24(h - 1) + le24(dur) + bytes([(dispose & 1) | ((blend_off & 1) << 1)])
    return chunk(b"ANMF", hdr + b"".join(chunk(t, p) for t, p in image_chunks))


def lossy_cases():
    cases = {}
    sweep = [  # (seed, w, h, noise, kwargs)
        (101, 256, 256, 6.0, dict(q=80)),
        (102, 800, 297, 6.0, dict(q=75)),
        (103, 17, 33, 6.0, dict(q=90)),
        (104, 1, 1, 6.0, dict(q=50)),
        (105, 16, 16, 6.0, dict(q=50)),
        (106, 15, 7, 30.0, dict(q=99)),
        (107, 255, 257, 40.0, dict(q=30, method=6)),
        (108, 640, 360, 6.0, dict(q=60, ft=0, fs=50)),            # simple loop filter
        (109, 321, 199, 20.0, dict(q=60, ft=0, fs=100, sharp=3)),
        (110, 321, 199, 20.0, dict(q=60, ft=1, fs=100, sharp=7)),
        (111, 400, 300, 12.0, dict(q=70, parts=1)),               # 2 token partitions
        (112, 400, 300, 12.0, dict(q=70, parts=2, segs=2)),       # 4 partitions
        (113, 400, 300, 12.0, dict(q=70, parts=3, segs=1)),       # 8 partitions, no segmentation
        (114, 200, 500, 12.0, dict(q=40, segs=3, sns=100)),
        (115, 333, 111, 60.0, dict(q=5, method=0)),               # heavy quantisation, fast mode decisions
        (116, 333, 111, 60.0, dict(q=100, method=6)),
        (117, 128, 128, 6.0, dict(q=50, fs=0)),                   # loop filter off
        (118, 1920, 1080, 6.0, dict(q=80)),                       # BASELINE geometry
        (119, 511, 385, 25.0, dict(q=85, parts=3)),
        (120, 97, 61, 40.0, dict(q=20, method=2, sharp=5, fs=80)),
    ]
    for seed, w, h, noise, kw in sweep:
        img = synth_image(seed, w, h, 3, noise=noise)
        cases[f"lossy{seed}"] = enc(img, **kw)
    rnd = np.random.default_rng(5).integers(0, 256, (70, 90, 3), dtype=np.uint8)
    cases["lossy_noise"] = enc(rnd, q=90)
    return cases


def container_cases():
    cases = {}
    a = synth_image(201, 120, 80, 3)
    vp8 = chunks_of(enc(a, q=70))[0][1]
    icc = bytes(range(256)) * 3 + b"x"  # odd length: exercises chunk padding
    cases["vp8x_icc"] = riff([vp8x(0x20, 120, 80), chunk(b"ICCP", icc), chunk(b"VP8 ", vp8)])
    cases["vp8x_exif"] = riff([vp8x(0x08, 120, 80), chunk(b"VP8 ", vp8), chunk(b"EXIF", b"Exif\0\0II*\0")])
    cases["vp8x_plain"] = riff([vp8x(0, 120, 80), chunk(b"VP8 ", vp8), chunk(b"ZZZZ", b"unknown chunk")])
    cases["bad_exif_flag"] = riff([vp8x(0, 120, 80), chunk(b"VP8 ", vp8), chunk(b"EXIF", b"Exif\0\0II*\0")])
    cases["vp8x_alpha_flag_no_alph"] = riff([vp8x(0x10, 120, 80), chunk(b"VP8 ", vp8)])
    # animation of opaque lossy frames at different offsets / sizes / dispose+blend flags
    frames = []
    for i, (w, h, x, y, dur, disp, nob) in enumerate([(160, 120, 0, 0, 100, 0, 0), (64, 48, 32, 20, 40, 1, 0),
                                                      (50, 50, 110, 70, 70, 0, 1), (160, 120, 0, 0, 0, 1, 1)]):
        p = chunks_of(enc(synth_image(210 + i, w, h, 3), q=60))[0][1]
        frames.append(anmf(x, y, w, h, dur, disp, nob, [(b"VP8 ", p)]))
    anim = chunk(b"ANIM", struct.pack("<IH", 0x80402010, 3))
    cases["anim_lossy"] = riff([vp8x(0x02, 160, 120), anim] + frames)
    cases["anim_lossy_alpha_flag"] = riff([vp8x(0x12, 160, 120), anim] + frames)
    # alpha + lossless: decoded by the reference; the device path reports what it cannot do yet
    rgba = synth_image(220, 96, 64, 4)
    cases["lossy_alpha"] = enc(rgba, q=80)
    cases["lossy_alpha_raw"] = enc(rgba, q=80, ac=0)
    cases["lossless_rgb"] = enc(synth_image(221, 96, 64, 3), lossless=1)
    cases["lossless_rgba"] = enc(rgba, lossless=1)
    # malformed containers: webp_decoder_create must refuse them
    good = cases["vp8x_icc"]
    cases["bad_truncated"] = good[:len(good) // 2]
    cases["bad_magic"] = b"RIFF" + good[4:8] + b"WEBQ" + good[12:]
    cases["bad_riff_size"] = b"RIFF" + struct.pack("<I", len(good) + 100) + good[8:]
    cases["bad_no_image"] = riff([vp8x(0, 120, 80)])
    cases["bad_keyframe_bit"] = riff([chunk(b"VP8 ", bytes([vp8[0] | 1]) + vp8[1:])])
    cases["bad_icc_flag_missing_chunk"] = riff([vp8x(0x20, 120, 80), chunk(b"VP8 ", vp8)])
    return cases


def fixture_cases():
    out = {}
    for name in ["tears_of_steel_no_icc", "tears_of_steel_icc", "party-discord", "animated-webp-supported"]:
        out["fixture_" + name] = open(os.path.join(TESTDATA, name + ".webp"), "rb").read()
    return out


def main():
    ref = abi.load_reference()
    out = {}
    names = []
    cases = {}
    cases.update(lossy_cases())
    cases.update(container_cases())
    cases.update(fixture_cases())
    for name, data in cases.items():
        info, frames, metas, rc = ref.webp_frames(data)
        names.append(name)
        out[f"webp_{name}"] = np.frombuffer(data, np.uint8)
        out[f"webprc_{name}"] = np.array([rc, len(frames)], np.int64)
        if info is None:
            print(f"{name:40s} rejected rc={rc}")
            continue
        out[f"webpinfo_{name}"] = np.array([info[k] for k in ("width", "height", "pixel_type", "num_frames", "total_duration",
                                                               "loop_count", "bg_color", "icc_len")], np.int64)
        out[f"webpmeta_{name}"] = np.array([[f.shape[1], f.shape[0], f.shape[2], m["x"], m["y"], m["delay"], m["dispose"],
                                             m["blend"]] for f, m in zip(frames, metas)], np.int64).reshape(-1, 8)
        out[f"webpsha_{name}"] = np.array([hashlib.sha256(f.tobytes()).hexdigest() for f in frames])
        if frames and frames[0].size <= 64 * 1024:
            out[f"webpframe0_{name}"] = frames[0]
        print(f"{name:40s} {len(data):8d} B  rc={rc} frames={len(frames)} info={info}")
    out["webp_names"] = np.array(names)
    path = os.path.join(ROOT, "tests", "golden", "webp_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
