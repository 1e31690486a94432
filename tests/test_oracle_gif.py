"""CPU: the GIF restatement in oracle/oracle_gif.c pinned on the reference: composited frames against
the reference-made golden frames, whole GIF -> GIF files against the bytes the reference library
wrote (tests/golden/gif_encode_golden.npz).  Bit / byte identical."""
import hashlib
import os

import numpy as np
import pytest

from lilliput_b200 import abi
from tests.cases import GIF_NAMES
from tests.golden.make_golden_gif_encode import CASES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GE = np.load(os.path.join(ROOT, "tests", "golden", "gif_encode_golden.npz"))


@pytest.mark.parametrize("name", GIF_NAMES)
def test_oracle_gif_frames_match_reference(oracle, golden, name):
    frames, delays, disposals, rc = oracle.gif_frames(golden[f"gif_{name}"].tobytes())
    meta = golden[f"gifmeta_{name}"]
    assert len(frames) == int(meta[7])
    assert [d * 10 for d in delays] == [int(v) for v in golden[f"gifdelay_{name}"]]
    # giflib disposal 2 -> GIF_DISPOSE_BACKGROUND (1), 3 -> GIF_DISPOSE_PREVIOUS (2), else none (ref giflib.cpp:187-199)
    assert [{2: 1, 3: 2}.get(d, 0) for d in disposals] == [int(v) for v in golden[f"gifdisp_{name}"]]
    assert [hashlib.sha256(f.tobytes()).hexdigest() for f in frames] == list(golden[f"gifsha_{name}"])


@pytest.mark.parametrize("fixture,label,kw", CASES, ids=[f"{c[0]}__{c[1]}" for c in CASES])
def test_oracle_gif_to_gif_bytes_match_reference(oracle, golden, fixture, label, kw):
    name = f"{fixture}__{label}"
    data = golden[f"gif_{fixture}"].tobytes()
    w, h, method = kw["Width"], kw["Height"], kw["ResizeMethod"]
    frames0, _, _, _ = oracle.gif_frames(data, max_frames=1)
    if method == abi.ImageOpsFit:  # ops.go:243-255: Fit requests larger than the source are trimmed
        w, h = oracle.expected_size(frames0[0].shape[1], frames0[0].shape[0], w, h)
    per_frame = None
    if method == abi.ImageOpsFit:
        per_frame = lambda f: oracle.fit(f, w, h)  # noqa: E731
    elif method == abi.ImageOpsResize:
        per_frame = lambda f: oracle.resize(f, w, h)  # noqa: E731
    out = oracle.gif_transcode(data, per_frame, max_frames=kw.get("MaxEncodeFrames", 0))
    assert len(out) == int(GE[f"len_{name}"])
    assert hashlib.sha256(out).hexdigest() == str(GE[f"sha_{name}"])
