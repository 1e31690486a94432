"""Generates tests/golden/jpeg_multiscan_golden.npz from the REFERENCE ITSELF (oracle/_ref):
progressive JPEGs (libjpeg's default scan script via OpenCV and Pillow writers: DC first/refine,
AC first/refine, 4:4:4 / 4:2:2 / 4:2:0 / 4:1:1, gray, restart intervals, optimised tables) and
sequential files with one scan per component, each with the pixels the reference decodes.

Run in the build container:  python tests/golden/make_golden_jpeg_multiscan.py
"""
import hashlib
import io
import os
import sys

import cv2
import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lilliput_b200 import abi  # noqa: E402
from lilliput_b200.synth import synth_image  # noqa: E402


def cases():
    out = {}
    k = 0
    for i, (w, h, ch) in enumerate([(64, 48, 3), (257, 131, 3), (33, 17, 1), (1, 1, 3), (7, 200, 3), (800, 297, 3)]):
        img = synth_image(400 + i, w, h, ch, noise=12.0)
        if ch == 1:
            img = img.reshape(h, w)
        for q, samp, rst in [(30, None, 0), (90, 0x111111, 0), (75, 0x211111, 7), (60, 0x221111, 0), (85, 0x411111, 3)]:
            if ch == 1 and samp not in (None, 0x111111):
                continue
            opts = [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_PROGRESSIVE, 1]
            if samp and ch == 3:
                opts += [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, samp]
            if rst:
                opts += [cv2.IMWRITE_JPEG_RST_INTERVAL, rst]
            ok, enc = cv2.imencode(".jpg", img, opts)
            assert ok
            out[f"prog{k:02d}_{w}x{h}c{ch}q{q}"] = enc.tobytes()
            k += 1
    rgb = synth_image(500, 511, 385, 3, noise=20.0)[:, :, ::-1]
    for name, kw in [("pil_prog_opt", dict(progressive=True, optimize=True, quality=50, subsampling=1)),
                     ("pil_prog_444", dict(progressive=True, quality=95, subsampling=0))]:
        bio = io.BytesIO()
        Image.fromarray(rgb).save(bio, "JPEG", **kw)
        out[name] = bio.getvalue()
    return out


def main():
    ref = abi.load_reference()
    out, names = {}, []
    for name, data in cases().items():
        px = ref.decode(data)
        names.append(name)
        out[f"jpg_{name}"] = np.frombuffer(data, np.uint8)
        out[f"sha_{name}"] = np.array(hashlib.sha256(px.tobytes()).hexdigest())
        out[f"shape_{name}"] = np.array(px.shape, np.int64)
        if px.size <= 48 * 1024:
            out[f"px_{name}"] = px
        print(f"{name:32s} {len(data):7d} B -> {px.shape}")
    out["names"] = np.array(names)
    path = os.path.join(ROOT, "tests", "golden", "jpeg_multiscan_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
