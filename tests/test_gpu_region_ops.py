"""GPU: the animated-compositing helpers of SURVEY 8(a) R10, driven DIRECTLY through the C ABI --
opencv_copy_to_region, opencv_copy_to_region_with_alpha, opencv_mat_clear_to_transparent
(ref opencv.cpp:680-752, 556-667, 508-543) -- and compared bit for bit with
  * the Appendix-D golden blend vectors (BLEND_CASES) and the C restatement (oracle.blend_over), and
  * the LIVE reference (oracle/_ref: the reference's own opencv.cpp over the vendored OpenCV) on random BGRA / BGR
    pairs: both alphas zero (0/0 -> NaN -> 0), every 3<->4 channel combination, offsets, and the size-mismatch path
    that resizes the source with INTER_LINEAR first.
The same ctypes driver runs against both libraries: they export the same symbols.
"""
import ctypes as C

import numpy as np
import pytest

from lilliput_b200 import abi
from tests.cases import BLEND_CASES

pytestmark = pytest.mark.gpu


class Mats:
    """Thin driver over the opencv_mat_* ABI of one library."""

    def __init__(self, lib):
        l = self.l = lib.l
        l.opencv_mat_create_from_data.restype = C.c_void_p
        l.opencv_mat_create_from_data.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
        l.opencv_mat_release.argtypes = [C.c_void_p]
        for name in ("opencv_copy_to_region", "opencv_copy_to_region_with_alpha"):
            f = getattr(l, name)
            f.restype = C.c_int
            f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        l.opencv_mat_clear_to_transparent.restype = C.c_int
        l.opencv_mat_clear_to_transparent.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        l.lp_mat_sync_host.restype = C.c_int
        l.lp_mat_sync_host.argtypes = [C.c_void_p]

    def wrap(self, a):
        h, w = a.shape[:2]
        ch = a.shape[2]
        m = self.l.opencv_mat_create_from_data(w, h, (ch - 1) << 3, a.ctypes.data, a.size)
        assert m
        return m

    def run(self, op, src, dst, x, y, w, h):
        """op(src -> region of dst); returns (rc, dst after the call)."""
        src = np.ascontiguousarray(src)
        dst = np.ascontiguousarray(dst).copy()
        ms, md = self.wrap(src), self.wrap(dst)
        rc = getattr(self.l, op)(ms, md, x, y, w, h)
        assert self.l.lp_mat_sync_host(md) == 0
        self.l.opencv_mat_release(ms)
        self.l.opencv_mat_release(md)
        return rc, dst

    def clear(self, dst, x, y, w, h):
        dst = np.ascontiguousarray(dst).copy()
        md = self.wrap(dst)
        rc = self.l.opencv_mat_clear_to_transparent(md, x, y, w, h)
        assert self.l.lp_mat_sync_host(md) == 0
        self.l.opencv_mat_release(md)
        return rc, dst


@pytest.fixture(scope="module")
def dev(cuda_lib):
    return Mats(cuda_lib)


@pytest.fixture(scope="module")
def ref(ref_lib):
    return Mats(ref_lib)


def test_blend_golden_vectors_on_the_device(dev, golden, oracle):
    """the seven (src, dst) BGRA pairs, one pixel each, plus the C restatement on a tile."""
    want = golden["blend_out"] if "blend_out" in golden.files else None
    for k, (s, d) in enumerate(BLEND_CASES):
        src = np.array([[s]], dtype=np.uint8)
        dst = np.array([[d]], dtype=np.uint8)
        rc, got = dev.run("opencv_copy_to_region_with_alpha", src, dst, 0, 0, 1, 1)
        assert rc == 0
        assert np.array_equal(got, oracle.blend_over(src, dst)), f"case {k}: {s} over {d}"
        if want is not None:
            assert np.array_equal(got[0, 0], want[k])
    rng = np.random.default_rng(5)
    src = rng.integers(0, 256, (37, 53, 4), dtype=np.uint8)
    dst = rng.integers(0, 256, (37, 53, 4), dtype=np.uint8)
    src[:8, :, 3] = 0
    dst[:4, :, 3] = 0                      # both alphas zero in the first rows: 0/0 -> NaN -> 0
    src[8:12, :, 3] = 255
    rc, got = dev.run("opencv_copy_to_region_with_alpha", src, dst, 0, 0, 53, 37)
    assert rc == 0 and np.array_equal(got, oracle.blend_over(src, dst))


@pytest.mark.parametrize("op", ["opencv_copy_to_region", "opencv_copy_to_region_with_alpha"])
def test_region_ops_match_the_live_reference(dev, ref, op):
    rng = np.random.default_rng(11 if op.endswith("alpha") else 12)
    n_checked = 0
    for trial in range(60):
        dch = int(rng.choice([3, 4]))
        sch = int(rng.choice([3, 4]))
        dh, dw = int(rng.integers(8, 70)), int(rng.integers(8, 90))
        w, h = int(rng.integers(1, dw + 1)), int(rng.integers(1, dh + 1))
        x, y = int(rng.integers(0, dw - w + 1)), int(rng.integers(0, dh - h + 1))
        if trial % 3 == 0:                 # source of another size: both sides resize it with INTER_LINEAR first
            sh, sw = int(rng.integers(1, 60)), int(rng.integers(1, 60))
        else:
            sh, sw = h, w
        src = rng.integers(0, 256, (sh, sw, sch), dtype=np.uint8)
        dst = rng.integers(0, 256, (dh, dw, dch), dtype=np.uint8)
        if sch == 4:
            src[rng.random((sh, sw)) < 0.3, 3] = 0
            src[rng.random((sh, sw)) < 0.2, 3] = 255
        if dch == 4:
            dst[rng.random((dh, dw)) < 0.3, 3] = 0
        rc_r, want = ref.run(op, src, dst, x, y, w, h)
        rc_d, got = dev.run(op, src, dst, x, y, w, h)
        assert rc_d == rc_r, f"trial {trial}: return code {rc_d} vs reference {rc_r}"
        if rc_r == 0:
            assert np.array_equal(got, want), f"trial {trial}: {sch}ch {sw}x{sh} -> {dch}ch region {w}x{h}@{x},{y}"
            n_checked += 1
    assert n_checked >= 40


def test_region_ops_reject_what_the_reference_rejects(dev, ref):
    rng = np.random.default_rng(13)
    src = rng.integers(0, 256, (10, 10, 4), dtype=np.uint8)
    dst = rng.integers(0, 256, (20, 30, 4), dtype=np.uint8)
    for (x, y, w, h) in [(-1, 0, 5, 5), (0, -2, 5, 5), (26, 0, 5, 5), (0, 16, 5, 5), (0, 0, 0, 5), (0, 0, 31, 5), (0, 0, 5, 21)]:
        for op in ("opencv_copy_to_region", "opencv_copy_to_region_with_alpha"):
            rc_r, want = ref.run(op, src, dst, x, y, w, h)
            rc_d, got = dev.run(op, src, dst, x, y, w, h)
            assert rc_d == rc_r, (op, x, y, w, h)
            assert np.array_equal(got, want)
        rc_r, want = ref.clear(dst, x, y, w, h)
        rc_d, got = dev.clear(dst, x, y, w, h)
        assert rc_d == rc_r and np.array_equal(got, want), ("clear", x, y, w, h)


def test_clear_to_transparent_matches_the_live_reference(dev, ref):
    rng = np.random.default_rng(14)
    for trial in range(30):
        ch = int(rng.choice([3, 4]))
        dh, dw = int(rng.integers(4, 60)), int(rng.integers(4, 80))
        w, h = int(rng.integers(1, dw + 1)), int(rng.integers(1, dh + 1))
        x, y = int(rng.integers(0, dw - w + 1)), int(rng.integers(0, dh - h + 1))
        dst = rng.integers(1, 256, (dh, dw, ch), dtype=np.uint8)
        rc_r, want = ref.clear(dst, x, y, w, h)
        rc_d, got = dev.clear(dst, x, y, w, h)
        assert rc_d == rc_r and np.array_equal(got, want), f"trial {trial}"
