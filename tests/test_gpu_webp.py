"""GPU: WebP decode (host RIFF walk + device VP8 / VP8L / ALPH decode + device upsample/colour) through
the webp_decoder_* ABI vs what the reference's webp.cpp returns for the same bytes (golden frames
made through oracle/_ref).  Bit-exact, frame by frame, metadata included."""
import hashlib

import numpy as np
import pytest

from lilliput_b200 import abi
from tests.webp_util import webp_golden

pytestmark = pytest.mark.gpu
G = webp_golden()
NAMES = [str(n) for n in G["webp_names"]]

@pytest.mark.parametrize("name", NAMES)
def test_webp_frames_match_reference(cuda_lib, name):
    data = G[f"webp_{name}"].tobytes()
    rc_ref, n_ref = [int(v) for v in G[f"webprc_{name}"]]
    info, frames, metas, rc = cuda_lib.webp_frames(data)
    if f"webpinfo_{name}" not in G.files:  # the reference refuses the container
        assert info is None and rc == rc_ref
        return
    keys = ("width", "height", "pixel_type", "num_frames", "total_duration", "loop_count", "bg_color", "icc_len")
    assert [info[k] for k in keys] == [int(v) for v in G[f"webpinfo_{name}"]]
    assert rc == rc_ref and len(frames) == n_ref
    meta = G[f"webpmeta_{name}"]
    for i, (f, m) in enumerate(zip(frames, metas)):
        assert [f.shape[1], f.shape[0], f.shape[2], m["x"], m["y"], m["delay"], m["dispose"], m["blend"]] == \
            [int(v) for v in meta[i]]
        assert hashlib.sha256(f.tobytes()).hexdigest() == str(G[f"webpsha_{name}"][i]), f"frame {i}"
    if f"webpframe0_{name}" in G.files:
        assert np.array_equal(frames[0], G[f"webpframe0_{name}"])


def test_webp_to_jpeg_and_png_transform(cuda_lib, oracle):
    """WebP -> Fit / Resize -> JPEG / PNG through lp_transform (ops.go:352-444 with webpDecoder)."""
    data = G["webp_lossy102"].tobytes()  # 800x297
    _, frames, _, _ = cuda_lib.webp_frames(data)
    src = frames[0]
    out = cuda_lib.transform(data, abi.ImageOptions(FileType=".jpeg", Width=256, Height=256,
                                                    ResizeMethod=abi.ImageOpsFit,
                                                    EncodeOptions={abi.JpegQuality: 85}))
    assert out == oracle.jpeg_encode(oracle.fit(src, 256, 256), 85)
    out = cuda_lib.transform(data, abi.ImageOptions(FileType=".png", Width=100, Height=50,
                                                    ResizeMethod=abi.ImageOpsResize,
                                                    EncodeOptions={abi.PngCompression: 7}))
    assert np.array_equal(oracle.png_decode(out), oracle.resize(src, 100, 50))


def test_webp_with_icc_transforms_like_the_reference(cuda_lib, oracle):
    """tears_of_steel_icc.webp (VP8X + ICCP + VP8): the profile is readable (info icc_len, checked
    above) and the JPEG written from it is byte-identical to the reference's, which does not embed
    the WebP's profile in a JPEG (measured through oracle/_ref)."""
    data = G["webp_fixture_tears_of_steel_icc"].tobytes()
    _, frames, _, _ = cuda_lib.webp_frames(data)
    out = cuda_lib.transform(data, abi.ImageOptions(FileType=".jpeg", Width=320, Height=200,
                                                    ResizeMethod=abi.ImageOpsFit,
                                                    EncodeOptions={abi.JpegQuality: 85}))
    assert out == oracle.jpeg_encode(oracle.fit(frames[0], 320, 200), 85)
    assert b"ICC_PROFILE\0" not in out[:4096]


def test_webp_header_fields(cuda_lib):
    w, h, ptype, _ = cuda_lib.header(G["webp_fixture_tears_of_steel_no_icc"].tobytes())[:4]
    assert (w, h) == (1920, 800)


@pytest.mark.parametrize("name", ["fixture_party-discord", "fixture_animated-webp-supported", "anim_lossy",
                                  "lossy_alpha", "lossless_rgba"])
def test_animated_and_alpha_webp_transform_matches_reference_library(cuda_lib, ref_lib, oracle, name):
    """Whole ImageOps.Transform with a WebP source (ops.go:352-444: animated sources composite into
    the canvas first, single-frame encoders stop after frame 1) against the reference library
    itself: JPEG byte-identical, PNG pixel-identical."""
    data = G[f"webp_{name}"].tobytes()
    jopt = abi.ImageOptions(FileType=".jpeg", Width=24, Height=16, ResizeMethod=abi.ImageOpsFit,
                            EncodeOptions={abi.JpegQuality: 90})
    assert cuda_lib.transform(data, jopt) == ref_lib.transform(data, jopt)
    popt = abi.ImageOptions(FileType=".png", Width=20, Height=12, ResizeMethod=abi.ImageOpsResize,
                            EncodeOptions={abi.PngCompression: 7})
    assert np.array_equal(oracle.png_decode(cuda_lib.transform(data, popt)),
                          oracle.png_decode(ref_lib.transform(data, popt)))
