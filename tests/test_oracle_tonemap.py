"""CPU: the HDR tone-map restatement (oracle.tonemap_to_sdr, following ref color_info.cpp:112-270 and the published
cv::TonemapReinhard algorithm) pinned on the reference: golden vectors made from oracle/_ref -- which links the
reference's own color_info.cpp -- by tests/golden/make_tonemap_golden.py, and, where _ref is present, fresh inputs
against the live library.  Tolerance, stated: at most +-1 LSB per 8-bit sample on at most 0.1 % of the samples (fp32
evaluation order of the vendored OpenCV's SIMD loops in the last ulp); everything else identical."""
import os

import numpy as np
import pytest

from lilliput_b200.synth import synth_image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def close_enough(a, b):
    d = np.abs(a.astype(int) - b.astype(int))
    return d.max() <= 1 and (d > 0).mean() <= 1e-3


def test_tonemap_oracle_matches_golden(oracle):
    g = np.load(os.path.join(ROOT, "tests", "golden", "tonemap_golden.npz"))
    n = 0
    for k in g.files:
        if not k.startswith("out_"):
            continue
        _, name, tr, pr = k.split("_")
        src = g["src_" + name]
        assert close_enough(oracle.tonemap_to_sdr(src, int(tr), int(pr)), g[k]), k
        if src.shape[2] == 4:
            assert np.array_equal(oracle.tonemap_to_sdr(src, int(tr), int(pr))[:, :, 3], src[:, :, 3])  # alpha untouched
        n += 1
    assert n >= 12


def test_tonemap_oracle_matches_live_reference(oracle, ref_lib):
    for seed, (w, h, c) in enumerate([(160, 120, 3), (97, 61, 4), (33, 200, 3)]):
        img = synth_image(900 + seed, w, h, c, noise=10.0)
        for tr in (16, 18):
            for pr in (9, 12, 11, 6, 10, 1, 2):
                assert close_enough(oracle.tonemap_to_sdr(img, tr, pr), ref_lib.tonemap(img, tr, pr)), (seed, tr, pr)
    flat = np.full((20, 30, 3), 77, np.uint8)   # max == min: the normalisation is skipped
    for tr in (16, 18):
        assert close_enough(oracle.tonemap_to_sdr(flat, tr, 9), ref_lib.tonemap(flat, tr, 9))
