"""Adam7-interlaced PNG decode: the oracle restatement (CPU) and the device path (GPU) against
pixels the reference itself decoded (tests/golden/png_adam7_golden.npz).  Lossless: bit-exact."""
import hashlib
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(ROOT, "tests", "golden", "png_adam7_golden.npz"))
NAMES = [str(n) for n in G["names"]]


def _check(px, name):
    assert list(px.shape) == [int(v) for v in G[f"shape_{name}"]]
    assert hashlib.sha256(np.ascontiguousarray(px).tobytes()).hexdigest() == str(G[f"sha_{name}"])


@pytest.mark.parametrize("name", NAMES)
def test_oracle_adam7_matches_reference(oracle, name):
    px = oracle.png_decode(G[f"png_{name}"].tobytes())
    _check(px[0] if isinstance(px, tuple) else px, name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_device_adam7_matches_reference(cuda_lib, name):
    _check(cuda_lib.decode(G[f"png_{name}"].tobytes()), name)
