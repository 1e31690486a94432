"""Randomised differential tests on the device: the product library (through the C ABI) against the
oracle on the same random inputs the build container checks against the reference itself
(tests/test_oracle_live_reference.py).  Bit-exact."""
import numpy as np
import pytest

from lilliput_b200 import abi
from tests.test_oracle_live_reference import _gif_files, _jpeg_files, _png_files, _rand_img

pytestmark = pytest.mark.gpu


def test_device_resize_fit_orient_random(cuda_lib, oracle):
    rng = np.random.default_rng(31)
    for _ in range(40):
        sw, sh = int(rng.integers(1, 700)), int(rng.integers(1, 500))
        dw, dh = int(rng.integers(1, 300)), int(rng.integers(1, 300))
        ch = int(rng.choice([1, 3, 4]))
        img = _rand_img(rng, sh, sw, ch)
        for interp in (abi.INTER_AREA, 1):
            assert np.array_equal(cuda_lib.resize(img, dw, dh, interpolation=interp),
                                  oracle.resize(img, dw, dh, interpolation=interp)), (sw, sh, dw, dh, ch, interp)
        fw, fh = min(dw, sw), min(dh, sh)
        assert np.array_equal(cuda_lib.fit(img, fw, fh), oracle.fit(img, fw, fh)), (sw, sh, fw, fh, ch)
    for _ in range(4):
        w, h, ch = int(rng.integers(1, 70)), int(rng.integers(1, 70)), int(rng.choice([1, 3, 4]))
        img = _rand_img(rng, h, w, ch)
        for o in range(1, 9):
            assert np.array_equal(cuda_lib.orient(img, o), oracle.orient(img, o)), (w, h, ch, o)


def test_device_jpeg_decode_random_files(cuda_lib, oracle):
    rng = np.random.default_rng(9)
    for label, data in _jpeg_files(rng):
        assert np.array_equal(cuda_lib.decode(data), oracle.jpeg_decode(data)[0]), label


def test_device_jpeg_encode_random_images(cuda_lib, oracle):
    rng = np.random.default_rng(10)
    for _ in range(20):
        w, h, ch = int(rng.integers(1, 400)), int(rng.integers(1, 300)), int(rng.choice([1, 3, 4]))
        img = _rand_img(rng, h, w, ch)
        q = int(rng.integers(1, 101))
        assert cuda_lib.encode(".jpeg", img, {abi.JpegQuality: q}) == oracle.jpeg_encode(img, q), (w, h, ch, q)


def test_device_png_decode_random_files(cuda_lib, oracle):
    rng = np.random.default_rng(11)
    for label, data in _png_files(rng):
        assert np.array_equal(cuda_lib.decode(data), oracle.png_decode(data)), label


def test_device_gif_decode_random_files(cuda_lib, oracle):
    rng = np.random.default_rng(12)
    for label, data in _gif_files(rng):
        gf, gd, gp, _ = oracle.gif_frames(data)
        ef, ed, ep, _ = cuda_lib.gif_frames(data)
        assert len(gf) == len(ef) and len(gf) >= 1, label
        assert [d * 10 for d in gd] == list(ed), label
        assert [{2: 1, 3: 2}.get(d, 0) for d in gp] == list(ep), label
        for k in range(len(gf)):
            assert np.array_equal(gf[k], ef[k]), (label, k)
