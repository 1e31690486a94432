"""The resize tap-table cache and the JPEG encoder constant cache are bounded; past the bound the tables are built per
call and live in stream order around the launch.  The bounds are read once per process, so the overflow path is
driven in a child process with both caches limited to ONE entry: every result must still equal the oracle's."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import numpy as np
from lilliput_b200 import abi
from lilliput_b200.synth import synth_image
from oracle import oracle
lib = abi.load_cuda()
rng = np.random.default_rng(3)
for k in range(12):
    sw, sh = int(rng.integers(40, 400)), int(rng.integers(40, 300))
    dw, dh = int(rng.integers(8, sw)), int(rng.integers(8, sh))
    ch = int(rng.choice([1, 3, 4]))
    img = synth_image(50 + k, sw, sh, ch, noise=9.0)
    img = img.reshape(sh, sw, ch) if ch > 1 else img.reshape(sh, sw)
    got = lib.resize(img, dw, dh)
    assert np.array_equal(got, oracle.resize(img, dw, dh)), ("resize", sw, sh, dw, dh, ch)
    q = int(rng.integers(1, 101))
    assert lib.encode(".jpeg", got, {abi.JpegQuality: q}) == oracle.jpeg_encode(got, q), ("encode", dw, dh, ch, q)
# and again in the same order: the one cached entry is long gone for all but the first
img = synth_image(99, 300, 200, 3, noise=5.0)
for _ in range(3):
    assert np.array_equal(lib.fit(img, 64, 64), oracle.fit(img, 64, 64))
print("bounded-cache child ok")
"""


@pytest.mark.gpu
def test_results_do_not_depend_on_the_cache_bounds():
    env = dict(os.environ, LP_RESIZE_TAB_CAP="1", LP_JPEG_ENC_CONST_CAP="1", PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-c", CHILD], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "bounded-cache child ok" in r.stdout
