"""Progressive / multi-scan JPEG decode: the oracle restatement (CPU) and the device kernel (GPU)
against pixels the reference itself decoded (tests/golden/jpeg_multiscan_golden.npz).  Bit-exact."""
import hashlib
import os

import numpy as np
import pytest

from lilliput_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(ROOT, "tests", "golden", "jpeg_multiscan_golden.npz"))
NAMES = [str(n) for n in G["names"]]


def _check(px, name):
    assert list(px.shape) == [int(v) for v in G[f"shape_{name}"]]
    assert hashlib.sha256(np.ascontiguousarray(px).tobytes()).hexdigest() == str(G[f"sha_{name}"])
    if f"px_{name}" in G.files:
        assert np.array_equal(px, G[f"px_{name}"])


@pytest.mark.parametrize("name", NAMES)
def test_oracle_progressive_matches_reference(oracle, name):
    px, _ = oracle.jpeg_decode(G[f"jpg_{name}"].tobytes())
    _check(px, name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_device_progressive_matches_reference(cuda_lib, name):
    _check(cuda_lib.decode(G[f"jpg_{name}"].tobytes()), name)


@pytest.mark.gpu
def test_progressive_source_through_transform(cuda_lib, oracle):
    """Progressive JPEG -> Fit -> baseline JPEG through lp_transform, against the oracle pipeline."""
    name = next(n for n in NAMES if "800x297" in n)
    data = G[f"jpg_{name}"].tobytes()
    src, _ = oracle.jpeg_decode(data)
    out = cuda_lib.transform(data, abi.ImageOptions(FileType=".jpeg", Width=256, Height=256,
                                                    ResizeMethod=abi.ImageOpsFit,
                                                    EncodeOptions={abi.JpegQuality: 85}))
    assert out == oracle.jpeg_encode(oracle.fit(src, 256, 256), 85)
