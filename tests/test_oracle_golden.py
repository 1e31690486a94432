"""CPU: pins the restated oracle (oracle/*.c) against golden vectors produced by the reference
itself (tests/golden/make_golden.py) and, when oracle/_ref is built, against the reference live."""
import hashlib

import numpy as np
import pytest

from lilliput_b200.synth import synth_image
from tests.cases import BLEND_CASES, JPEG_CASES, ORIENT_SRC, RESIZE_CASES


@pytest.mark.parametrize("case", RESIZE_CASES, ids=lambda c: f"seed{c[0]}")
def test_resize_matches_golden(oracle, golden, case):
    seed, sw, sh, ch, crop, dw, dh, interp = case
    img = synth_image(seed, sw, sh, ch, noise=12.0)
    got = oracle.resize(img, dw, dh, crop=crop, interpolation=interp)
    assert np.array_equal(got, golden[f"resize_{seed}"])  # bit-exact


@pytest.mark.parametrize("case", JPEG_CASES, ids=lambda c: f"seed{c[0]}")
def test_jpeg_encode_matches_golden(oracle, golden, case):
    seed, w, h, ch, q = case
    img = synth_image(seed, w, h, ch, noise=8.0)
    got = oracle.jpeg_encode(img, q)
    assert got == golden[f"jpeg_{seed}"].tobytes()  # byte-identical file
    sha = dict(s.split(":") for s in golden["jpeg_sha"])
    assert hashlib.sha256(got).hexdigest() == sha[str(seed)]


@pytest.mark.parametrize("case", JPEG_CASES, ids=lambda c: f"seed{c[0]}")
def test_jpeg_decode_matches_golden(oracle, golden, case):
    seed = case[0]
    got, orient = oracle.jpeg_decode(golden[f"jpeg_{seed}"].tobytes())
    assert orient == 1
    assert np.array_equal(got, golden[f"jpegdec_{seed}"])


@pytest.mark.parametrize("name", ["444", "422", "440", "411", "420"])
@pytest.mark.parametrize("rst", [0, 3])
def test_jpeg_decode_sampling_and_restart(oracle, golden, name, rst):
    got, _ = oracle.jpeg_decode(golden[f"jpegvar_{name}_{rst}"].tobytes())
    assert np.array_equal(got, golden[f"jpegvardec_{name}_{rst}"])


def test_orientation_golden_table(oracle, golden):
    # 3x2 image 012/345
    expect = {1: [[0, 1, 2], [3, 4, 5]], 2: [[2, 1, 0], [5, 4, 3]], 3: [[5, 4, 3], [2, 1, 0]],
              4: [[3, 4, 5], [0, 1, 2]], 5: [[0, 3], [1, 4], [2, 5]], 6: [[3, 0], [4, 1], [5, 2]],
              7: [[5, 2], [4, 1], [3, 0]], 8: [[2, 5], [1, 4], [0, 3]]}
    for o in range(1, 9):
        got = oracle.orient(ORIENT_SRC, o)
        assert got.tolist() == expect[o]
        assert np.array_equal(got, golden[f"orient_{o}"])
    img = synth_image(42, 37, 23, 3, noise=10.0)
    for o in range(1, 9):
        assert np.array_equal(oracle.orient(img, o), golden[f"orient3_{o}"])


def test_blend_golden_vectors(oracle):
    # SURVEY.md Appendix D, captured from the reference
    expect = [(0, 0, 0, 0), (100, 110, 120, 255), (10, 20, 30, 255), (55, 65, 75, 255),
              (10, 20, 30, 128), (92, 64, 56, 160), (1, 1, 1, 254)]
    for (s, d), e in zip(BLEND_CASES, expect):
        got = oracle.blend_over(np.array([[s]], dtype=np.uint8), np.array([[d]], dtype=np.uint8))
        assert tuple(got[0, 0]) == e
    got = oracle.blend_over(np.array([[(10, 20, 30, 128)]], dtype=np.uint8),
                            np.array([[(100, 110, 120)]], dtype=np.uint8))
    assert tuple(got[0, 0]) == (55, 65, 75)


def test_policy_helpers(oracle):
    # Fit crop rectangles quoted in SURVEY.md 8a (configs 1-4)
    assert oracle.fit_rect(1920, 1080, 256, 256) == (420, 0, 1080, 1080)
    assert oracle.fit_rect(3840, 2160, 512, 512) == (840, 0, 2160, 2160)
    assert oracle.fit_rect(1280, 720, 256, 256) == (280, 0, 720, 720)
    assert oracle.fit_rect(800, 297, 256, 256) == (251, 0, 297, 297)
    # calculateExpectedSize (ref ops.go:243-255)
    assert oracle.expected_size(800, 297, 256, 256) == (256, 256)
    assert oracle.expected_size(100, 80, 256, 256) == (80, 80)
    assert oracle.expected_size(100, 80, 300, 200) == (100, 80)
    assert oracle.expected_size(100, 80, 300, 50) == (300, 50)


def test_config1_end_to_end_through_oracle(oracle, golden):
    """ferry_sunset.jpg -> Fit 256x256 JPEG q85: decode + crop/resize + encode restated."""
    data = golden["c1_input"].tobytes()
    img, orient = oracle.jpeg_decode(data)
    assert img.shape == (297, 800, 3) and orient == 1
    out = oracle.jpeg_encode(oracle.fit(img, 256, 256), 85)
    assert len(out) == 11651
    assert out == golden["c1_output"].tobytes()


def test_oracle_against_live_reference(oracle, ref_lib):
    """Where oracle/_ref exists, compare on fresh random cases (not in the golden file)."""
    rng = np.random.default_rng(1234)
    for _ in range(6):
        sw, sh = int(rng.integers(40, 700)), int(rng.integers(40, 700))
        dw, dh = int(rng.integers(8, 400)), int(rng.integers(8, 400))
        ch = int(rng.choice([1, 3, 4]))
        img = rng.integers(0, 256, (sh, sw, ch) if ch > 1 else (sh, sw), dtype=np.uint8)
        for interp in (3, 1):
            assert np.array_equal(oracle.resize(img, dw, dh, interpolation=interp),
                                  ref_lib.resize(img, dw, dh, interpolation=interp))
    from lilliput_b200 import abi
    for seed in range(3):
        w, h = int(rng.integers(1, 300)), int(rng.integers(1, 300))
        img = synth_image(900 + seed, w, h, 3, noise=10.0)
        q = int(rng.integers(1, 101))
        enc = ref_lib.encode(".jpeg", img, {abi.JpegQuality: q})
        assert oracle.jpeg_encode(img, q) == enc
        assert np.array_equal(oracle.jpeg_decode(enc)[0], ref_lib.decode(enc))


from tests.cases import PNG_NAMES  # noqa: E402


@pytest.mark.parametrize("name", PNG_NAMES)
def test_png_decode_matches_golden(oracle, golden, name):
    """PNG decode is lossless: every colour type / bit depth / filter / block type the fixtures
    cover must equal what the reference decoded (Gray, BGR or BGRA u8; 16-bit -> high byte)."""
    got = oracle.png_decode(golden[f"png_{name}"].tobytes())
    assert np.array_equal(got, golden[f"pngdec_{name}"])
