"""Generates tests/golden/png_adam7_golden.npz from the REFERENCE ITSELF (oracle/_ref): Adam7-
interlaced PNGs of every colour type / bit depth (written by tests/png_writer.py, since neither
Pillow nor OpenCV writes interlaced files), each with the pixels the reference decodes.

Run in the build container:  python tests/golden/make_golden_png_adam7.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lilliput_b200 import abi  # noqa: E402
from tests.png_writer import write_png  # noqa: E402


def cases():
    rng = np.random.default_rng(4)
    out = {}
    for (w, h) in [(1, 1), (2, 3), (5, 5), (8, 8), (9, 17), (33, 20), (100, 37)]:
        for ct, ch in [(0, 1), (2, 3), (3, 1), (4, 2), (6, 4)]:
            for bd in ([1, 4, 8, 16] if ct == 0 else [2, 8] if ct == 3 else [8, 16]):
                pal = trns = None
                if ct == 3:
                    npal = min(256, 1 << bd)
                    pal = rng.integers(0, 256, (npal, 3))
                    smp = rng.integers(0, npal, (h, w, 1))
                    if (w + h) % 2:
                        trns = bytes(rng.integers(0, 256, npal // 2 + 1, dtype=np.uint8))
                else:
                    smp = rng.integers(0, (1 << bd), (h, w, ch))
                out[f"a7_{w}x{h}_ct{ct}_bd{bd}"] = write_png(smp, ct, bd, interlace=True, palette=pal, trns=trns)
    # a larger natural image, RGB8, interlaced
    from lilliput_b200.synth import synth_image
    img = synth_image(77, 301, 211, 3)[:, :, ::-1]
    out["a7_301x211_rgb"] = write_png(img, 2, 8, interlace=True, ftypes=(4, 1, 3))
    return out


def main():
    ref = abi.load_reference()
    out, names = {}, []
    for name, data in cases().items():
        px = ref.decode(data)
        names.append(name)
        out[f"png_{name}"] = np.frombuffer(data, np.uint8)
        out[f"sha_{name}"] = np.array(hashlib.sha256(px.tobytes()).hexdigest())
        out[f"shape_{name}"] = np.array(px.shape, np.int64)
    out["names"] = np.array(names)
    path = os.path.join(ROOT, "tests", "golden", "png_adam7_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), len(names), "cases")


if __name__ == "__main__":
    main()
