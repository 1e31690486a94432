"""GPU: GIF -> GIF ImageOps.Transform (device LZW decode + compositor, fit / resize, device palette
mapping + device LZW encode) against the BYTES the reference library wrote for the same call
(tests/golden/gif_encode_golden.npz, made through oracle/_ref).  GIF encoding is deterministic
integer work, so the bar is byte-identical files."""
import hashlib
import os

import numpy as np
import pytest

from lilliput_b200 import abi
from tests.golden.make_golden_gif_encode import CASES, TIMEOUT_NS

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(ROOT, "tests", "golden", "gif_encode_golden.npz"))


@pytest.mark.parametrize("fixture,label,kw", CASES, ids=[f"{c[0]}__{c[1]}" for c in CASES])
def test_gif_to_gif_bytes_match_reference(cuda_lib, golden, fixture, label, kw):
    name = f"{fixture}__{label}"
    data = golden[f"gif_{fixture}"].tobytes()
    out = cuda_lib.transform(data, abi.ImageOptions(FileType=".gif", EncodeTimeout_ns=TIMEOUT_NS, **kw))
    if hashlib.sha256(out).hexdigest() != str(G[f"sha_{name}"]):
        dump = os.path.join(ROOT, "gpurun_out")
        os.makedirs(dump, exist_ok=True)
        open(os.path.join(dump, f"gifenc_{name}.gif"), "wb").write(out)
        msg = f"{len(out)} B vs {int(G[f'len_{name}'])} B"
        if f"out_{name}" in G.files:
            want = G[f"out_{name}"].tobytes()
            first = next((i for i in range(min(len(out), len(want))) if out[i] != want[i]), min(len(out), len(want)))
            msg += f", first difference at byte {first}"
        pytest.fail(msg)


def test_gif_encoder_reports_a_full_destination(cuda_lib, golden):
    """encode_func refuses writes past dst_len (ref giflib.cpp:762-771) -> the frame fails -> ErrInvalidImage."""
    data = golden["gif_no-loop"].tobytes()
    with pytest.raises(abi.LilliputError):
        cuda_lib.transform(data, abi.ImageOptions(FileType=".gif", Width=64, Height=64, ResizeMethod=abi.ImageOpsFit,
                                                  EncodeTimeout_ns=TIMEOUT_NS), dst_cap=2000)
