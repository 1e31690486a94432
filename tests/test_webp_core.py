"""CPU: the VP8 key-frame decoding logic the device kernels run (lilliput_b200/csrc/vp8_core.h),
compiled for the host by oracle/oracle_webp.cpp, against frames decoded by the reference's own
libwebp (tests/golden/webp_golden.npz, made by make_golden_webp.py through oracle/_ref).
Bit-exact: VP8 decoding is integer arithmetic end to end, upsampler and colour matrix included."""
import hashlib

import numpy as np
import pytest

from tests.webp_util import (alph_cpu_decode, chunks_of, frames_of, vp8_cpu_decode, vp8_cpu_lib, vp8l_cpu_decode,
                             webp_golden)

G = webp_golden()
LOSSY = [n for n in G["webp_names"] if n.startswith("lossy") and "alpha" not in n] + \
        ["vp8x_icc", "vp8x_plain", "fixture_tears_of_steel_no_icc", "fixture_tears_of_steel_icc"]


@pytest.fixture(scope="module")
def cpu():
    return vp8_cpu_lib()


@pytest.mark.parametrize("name", LOSSY)
def test_vp8_core_matches_reference_frames(cpu, name):
    payload = dict(chunks_of(G[f"webp_{name}"].tobytes()))[b"VP8 "]
    got = vp8_cpu_decode(cpu, payload)
    info = G[f"webpinfo_{name}"]
    assert got.shape == (int(info[1]), int(info[0]), 3)
    assert hashlib.sha256(got.tobytes()).hexdigest() == str(G[f"webpsha_{name}"][0])
    if f"webpframe0_{name}" in G.files:
        assert np.array_equal(got, G[f"webpframe0_{name}"])


def test_vp8_core_against_reference_library_sweep(cpu, ref_lib):
    """Where oracle/_ref exists (build container): fresh streams from OpenCV's WebP writer at a
    spread of sizes and qualities, decoded by the reference and by the core."""
    cv2 = pytest.importorskip("cv2")
    from lilliput_b200.synth import synth_image
    for i, (w, h, q) in enumerate([(64, 48, 90), (100, 75, 30), (257, 131, 75), (333, 500, 95), (17, 9, 10), (2, 2, 50)]):
        ok, enc = cv2.imencode(".webp", synth_image(300 + i, w, h, 3, noise=15.0), [cv2.IMWRITE_WEBP_QUALITY, q])
        assert ok
        data = enc.tobytes()
        _, frames, _, rc = ref_lib.webp_frames(data)
        assert rc == 0
        got = vp8_cpu_decode(cpu, dict(chunks_of(data))[b"VP8 "])
        assert np.array_equal(got, frames[0])


MIXED = ["lossless_rgb", "lossless_rgba", "lossy_alpha", "lossy_alpha_raw", "fixture_party-discord",
         "fixture_animated-webp-supported"]


@pytest.mark.parametrize("name", MIXED)
def test_vp8l_and_alph_core_match_reference_frames(cpu, name):
    """VP8L frames, and lossy frames with an ALPH plane (raw or VP8L-coded, filtered), frame by
    frame: colour from the VP8 / VP8L core, alpha from the ALPH core, against the reference's output."""
    data = G[f"webp_{name}"].tobytes()
    meta = G[f"webpmeta_{name}"]
    sha = G[f"webpsha_{name}"]
    frames = frames_of(data)
    assert len(frames) == len(sha)
    for i, (tag, payload, alph) in enumerate(frames):
        w, h, ch = [int(v) for v in meta[i][:3]]
        if tag == b"VP8L":
            got = vp8l_cpu_decode(cpu, payload, w, h, ch)
        else:
            bgr = vp8_cpu_decode(cpu, payload)
            if ch == 4:
                a = alph_cpu_decode(cpu, alph, w, h) if alph is not None else np.full((h, w), 255, np.uint8)
                got = np.dstack([bgr, a])
            else:
                got = bgr
        assert got.shape == (h, w, ch)
        assert hashlib.sha256(np.ascontiguousarray(got).tobytes()).hexdigest() == str(sha[i]), f"frame {i}"
