"""Generates tests/golden/tonemap_golden.npz from oracle/_ref (the reference's color_info.cpp + vendored OpenCV):
inputs and the reference's tone-mapped outputs for PQ / HLG x a few primaries.  Run where /root/reference exists."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lilliput_b200 import abi  # noqa: E402
from lilliput_b200.synth import synth_image  # noqa: E402

ref = abi.load_reference()
out = {}
for name, (seed, w, h, c, noise) in {"a": (61, 96, 64, 3, 8.0), "b": (62, 50, 70, 4, 20.0), "c": (63, 128, 40, 3, 3.0)}.items():
    img = synth_image(seed, w, h, c, noise=noise)
    out["src_" + name] = img
    for tr in (16, 18):
        for pr in (9, 12, 1):
            out[f"out_{name}_{tr}_{pr}"] = ref.tonemap(img, tr, pr)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "tonemap_golden.npz"), **out)
print("wrote", len(out), "arrays")
