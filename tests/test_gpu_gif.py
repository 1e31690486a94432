"""GPU: GIF decode (host container walk + device LZW + device compositor) through gifDecoder's
DecodeTo loop vs frames decoded by the reference itself.  Lossless: bit-exact, frame by frame."""
import hashlib

import numpy as np
import pytest

from lilliput_b200 import abi
from tests.cases import GIF_NAMES

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", GIF_NAMES)
def test_gif_frames_match_reference(cuda_lib, golden, name):
    data = golden[f"gif_{name}"].tobytes()
    meta = golden[f"gifmeta_{name}"]
    info = cuda_lib.gif_info(data)
    assert [info["width"], info["height"], info["frame_count"], info["loop_count"], info["duration_ms"],
            info["background_color"]] == [int(v) for v in meta[:6]]
    frames, delays, disposals, rc = cuda_lib.gif_frames(data)
    assert rc == int(meta[6]) and len(frames) == int(meta[7])
    assert delays == [int(v) for v in golden[f"gifdelay_{name}"]]
    assert disposals == [int(v) for v in golden[f"gifdisp_{name}"]]
    sha = [hashlib.sha256(f.tobytes()).hexdigest() for f in frames]
    assert sha == list(golden[f"gifsha_{name}"])
    if f"gifframes_{name}" in golden.files:
        assert np.array_equal(frames, golden[f"gifframes_{name}"])


def test_reference_metadata_table(cuda_lib, golden):
    """giflib_test.go:201-241 / webp_test.go:424-454: loop counts, frame counts, durations."""
    expect = {"party-discord": (0, 16, 480), "no-loop": (1, 44, 4400), "duplicate_number_of_loops": (2, 2, 0),
              "dispose_bgnd": (0, 5, 5000)}
    for name, (loops, frames, dur) in expect.items():
        info = cuda_lib.gif_info(golden[f"gif_{name}"].tobytes())
        assert (info["loop_count"], info["frame_count"], info["duration_ms"]) == (loops, frames, dur)


def test_gif_first_frame_to_jpeg_and_png(cuda_lib, oracle, golden):
    """GIF -> Fit -> JPEG / PNG through lp_transform: animated source, single-frame encoders return
    after frame 1 (ops.go:416-418); the animated branch composites into a canvas-sized buffer first."""
    data = golden["gif_party-discord"].tobytes()
    first = golden["gifframes_party-discord"][0]  # 18x28 BGRA full canvas
    out = cuda_lib.transform(data, abi.ImageOptions(FileType=".png", Width=16, Height=16,
                                                    ResizeMethod=abi.ImageOpsFit,
                                                    EncodeOptions={abi.PngCompression: 7}))
    assert np.array_equal(oracle.png_decode(out), oracle.fit(first, 16, 16))
    out = cuda_lib.transform(data, abi.ImageOptions(FileType=".jpeg", Width=10, Height=12,
                                                    ResizeMethod=abi.ImageOpsResize,
                                                    EncodeOptions={abi.JpegQuality: 90}))
    assert out == oracle.jpeg_encode(oracle.resize(first, 10, 12), 90)


def test_gif_errors(cuda_lib, golden):
    data = golden["gif_no-loop"].tobytes()
    frames, _, _, rc = cuda_lib.gif_frames(data[: len(data) // 2])  # truncated mid-stream
    assert rc != 0 and len(frames) >= 1
    with pytest.raises(abi.LilliputError):
        cuda_lib.gif_info(b"GIF89a\x00")
