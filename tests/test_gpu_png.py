"""GPU: PNG decode (device inflate + defilter + convert) through the C ABI vs the reference-made
golden vectors and the oracle.  Lossless path: bit-exact."""
import io

import numpy as np
import pytest

from lilliput_b200 import abi
from lilliput_b200.synth import synth_image
from tests.cases import PNG_NAMES

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", PNG_NAMES)
def test_png_decode_matches_reference_golden(cuda_lib, golden, name):
    data = golden[f"png_{name}"].tobytes()
    exp = golden[f"pngdec_{name}"]
    w, h, t, o = cuda_lib.header(data)
    assert (h, w) == exp.shape[:2] and o == 1
    assert ((t >> 3) & 63) + 1 == (1 if exp.ndim == 2 else exp.shape[2])
    assert np.array_equal(cuda_lib.decode(data), exp)


def test_png_random_sizes_vs_oracle(cuda_lib, oracle):
    from PIL import Image
    rng = np.random.default_rng(11)
    for i in range(12):
        w, h = int(rng.integers(1, 300)), int(rng.integers(1, 200))
        ch = int(rng.choice([1, 3, 4]))
        img = synth_image(700 + i, w, h, ch, noise=float(rng.choice([0.0, 3.0, 20.0])))
        mode_img = Image.fromarray(img if ch != 3 else img[:, :, ::-1].copy()) if ch != 4 else \
            Image.fromarray(img[:, :, [2, 1, 0, 3]].copy())
        bio = io.BytesIO()
        mode_img.save(bio, "PNG", compress_level=int(rng.integers(0, 10)))
        data = bio.getvalue()
        assert np.array_equal(cuda_lib.decode(data), oracle.png_decode(data)), (w, h, ch)


def test_png_to_jpeg_transform(cuda_lib, golden, oracle):
    """PNG -> Fit -> JPEG through lp_transform (the PNG share of BASELINE config 5)."""
    data = golden["png_fixture_ferry"].tobytes()
    opt = abi.ImageOptions(FileType=".jpeg", Width=256, Height=256, ResizeMethod=abi.ImageOpsFit,
                           NormalizeOrientation=True, EncodeOptions={abi.JpegQuality: 85})
    out = cuda_lib.transform(data, opt)
    dec = oracle.png_decode(data)
    assert out == oracle.jpeg_encode(oracle.fit(dec, 256, 256), 85)
    # RGBA source: alpha is dropped by the JPEG encoder
    data = golden["png_filters"].tobytes()
    out = cuda_lib.transform(data, abi.ImageOptions(FileType=".jpeg", Width=100, Height=60,
                                                    ResizeMethod=abi.ImageOpsResize,
                                                    EncodeOptions={abi.JpegQuality: 90}))
    dec = oracle.png_decode(data)
    assert out == oracle.jpeg_encode(oracle.resize(dec, 100, 60), 90)


def test_png_errors(cuda_lib, golden):
    data = golden["png_rgb"].tobytes()
    with pytest.raises(abi.LilliputError) as e:
        cuda_lib.decode(data[: len(data) // 2] + b"\x00" * 40)  # truncated inside the first IDAT
    # refused at the header, like the reference (OpenCV's readHeader wants the first IDAT chunk whole;
    # tests/test_host_png_header.py compares the two on files cut short anywhere)
    assert e.value.code == -1  # ErrInvalidImage


@pytest.mark.parametrize("case", [(97, 61, 3, 7), (64, 64, 4, 1), (256, 256, 3, 7), (33, 17, 1, 6),
                                  (500, 300, 4, 9), (40, 30, 3, 0), (1, 1, 3, 7), (640, 360, 3, None)])
def test_png_encode_decodes_to_identical_pixels(cuda_lib, oracle, ref_lib, case):
    """PNG output contract = decoded-pixel equality + IHDR policy (colour type 0/2/6, 8-bit), not
    byte-identical files.  The file must decode with an independent decoder (oracle, PIL) to the
    exact input pixels.  Size: LZ77 whose search effort follows PngCompression + per-chunk dynamic Huffman codes
    (deflate_enc_core.h): within 1.10 x of the reference's file at the same level (0.86-1.04 measured on the CPU build
    of the same code, tests/test_deflate_enc_core.py)."""
    import io
    from PIL import Image
    w, h, ch, level = case
    img = synth_image(800 + w, w, h, ch, noise=4.0)
    opts = {} if level is None else {abi.PngCompression: level}
    data = cuda_lib.encode(".png", img, opts)
    assert data[:8] == b"\x89PNG\r\n\x1a\n" and data[12:16] == b"IHDR"
    assert int.from_bytes(data[16:20], "big") == w and int.from_bytes(data[20:24], "big") == h
    assert data[24] == 8 and data[25] == {1: 0, 3: 2, 4: 6}[ch] and data[28] == 0   # depth, type, no interlace
    assert np.array_equal(oracle.png_decode(data), img)
    pil = np.array(Image.open(io.BytesIO(data)))
    exp = img if ch == 1 else (img[:, :, ::-1] if ch == 3 else img[:, :, [2, 1, 0, 3]])
    assert np.array_equal(pil, exp)
    assert np.array_equal(cuda_lib.decode(data), img)   # and through the device decoder
    ref_size = len(ref_lib.encode(".png", img, opts))
    if level != 0 and w * h > 4096:
        assert len(data) <= 1.10 * ref_size + 256, (len(data), ref_size)


def test_png_to_png_transform(cuda_lib, oracle, golden):
    """PNG -> Fit -> PNG (lossless path): decoded output == oracle fit of the decoded input."""
    data = golden["png_filters"].tobytes()   # 320x200 BGRA
    opt = abi.ImageOptions(FileType=".png", Width=64, Height=64, ResizeMethod=abi.ImageOpsFit,
                           EncodeOptions={abi.PngCompression: 7})
    out = cuda_lib.transform(data, opt)
    exp = oracle.fit(oracle.png_decode(data), 64, 64)
    assert np.array_equal(oracle.png_decode(out), exp)
    with pytest.raises(abi.LilliputError) as e:   # too-small destination -> ErrBufTooSmall
        cuda_lib.transform(data, opt, dst_cap=64)
    assert e.value.code == -3


@pytest.mark.gpu
def test_png_encoder_stream_is_the_serial_cores(cuda_lib, tmp_path):
    """The device runs deflate_enc_core.h one lane per chunk; the CPU suite runs the same source against zlib.  Same
    filtered scanlines in -> byte-identical zlib stream out, at every effort level."""
    import ctypes as C
    import os
    import struct
    import subprocess
    import zlib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = str(tmp_path / "libdefenc.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so,
                           os.path.join(root, "tests", "native", "deflate_enc_sim.cpp")])
    l = C.CDLL(so)
    l.defenc_compress.restype = C.c_long
    l.defenc_compress.argtypes = [C.c_char_p, C.c_long, C.c_int, C.c_void_p, C.c_long]
    for (w, h, ch, level, noise) in [(300, 200, 3, 1, 4.0), (256, 256, 4, 6, 0.0), (640, 360, 3, 9, 8.0), (97, 61, 1, 3, 2.0),
                                     (50, 40, 3, 0, 4.0)]:
        img = synth_image(900 + w, w, h, ch, noise=noise)
        png = cuda_lib.encode(".png", img, {abi.PngCompression: level})
        o, z = 8, b""
        while o < len(png):
            ln, = struct.unpack(">I", png[o:o + 4])
            if png[o + 4:o + 8] == b"IDAT":
                z += png[o + 8:o + 8 + ln]
            o += 12 + ln
        raw = zlib.decompress(z)
        assert len(raw) == (w * ch + 1) * h
        cap = len(raw) + 4096
        out = (C.c_uint8 * cap)()
        n = l.defenc_compress(raw, len(raw), level, out, cap)
        assert n > 0 and bytes(out[:n]) == z, (w, h, ch, level)
