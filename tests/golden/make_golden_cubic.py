"""Golden vectors for cv::resize(INTER_CUBIC) as the REFERENCE answers it (oracle/_ref = the reference's own
opencv.cpp over its vendored OpenCV + IPP; ref opencv.cpp:20, 196-208).  Run in the build container:
    python tests/golden/make_golden_cubic.py
Writes tests/golden/cubic_golden.npz: for every case the seeded input and the reference's output."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lilliput_b200 import abi  # noqa: E402

CASES = [  # (sw, sh, dw, dh, channels)
    (64, 48, 256, 192, 3), (100, 75, 256, 256, 3), (31, 17, 97, 203, 4), (120, 90, 640, 480, 1), (40, 40, 41, 39, 3),
    (200, 150, 128, 96, 3), (4, 4, 64, 64, 4), (5, 9, 7, 200, 1),
    # sources under 4 px on an axis: OpenCV's own fixed-point bicubic instead of IPP
    (3, 20, 31, 47, 3), (20, 3, 40, 9, 4), (1, 9, 31, 47, 1), (2, 2, 33, 17, 3), (9, 1, 64, 5, 3),
]


def main():
    ref = abi.load_reference()
    out = {}
    for i, (sw, sh, dw, dh, ch) in enumerate(CASES):
        rng = np.random.default_rng(4000 + i)
        img = rng.integers(0, 256, (sh, sw, ch), dtype=np.uint8)
        if i % 3 == 2:
            img = (img // 85 * 85).astype(np.uint8)  # flat levels: overshoot clamps at 0 / 255
        img = np.ascontiguousarray(img if ch > 1 else img.reshape(sh, sw))
        out[f"cubic_{i}_src"] = img
        out[f"cubic_{i}_dst"] = ref.resize(img, dw, dh, interpolation=2)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cubic_golden.npz"), **out)
    print("wrote", len(CASES), "cases")


if __name__ == "__main__":
    main()
