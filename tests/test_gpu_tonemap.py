"""GPU: Framebuffer.TonemapToSDR on the device (csrc/tonemap.cu, ref color_info.cpp:112-270) against the oracle
restatement and the live reference, and an HDR-tagged PNG through the whole Transform on both libraries.
Tolerance (floating point, stated): +-1 LSB per 8-bit sample on at most 0.2 % of the samples."""
import struct
import zlib

import numpy as np
import pytest

from lilliput_b200 import abi
from lilliput_b200.synth import synth_image

pytestmark = pytest.mark.gpu


def close_enough(a, b, frac=2e-3):
    d = np.abs(a.astype(int) - b.astype(int))
    return d.max() <= 1 and (d > 0).mean() <= frac


@pytest.mark.parametrize("transfer", [16, 18])
def test_device_tonemap_matches_oracle_and_reference(cuda_lib, ref_lib, oracle, transfer):
    for seed, (w, h, c) in enumerate([(160, 120, 3), (97, 61, 4), (1920, 1080, 3), (3, 2, 3)]):
        img = synth_image(910 + seed, w, h, c, noise=10.0)
        for pr in (9, 12, 6, 10, 1):
            got = cuda_lib.tonemap(img, transfer, pr)
            assert close_enough(got, oracle.tonemap_to_sdr(img, transfer, pr)), (seed, pr)
            if w <= 200:
                assert close_enough(got, ref_lib.tonemap(img, transfer, pr)), (seed, pr)
            if c == 4:
                assert np.array_equal(got[:, :, 3], img[:, :, 3])


def _png_with_cicp(img_bgr, primaries, transfer):
    h, w, _ = img_bgr.shape
    rows = np.concatenate([np.zeros((h, 1), np.uint8), img_bgr[:, :, ::-1].reshape(h, -1)], axis=1).tobytes()

    def ch(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xffffffff)
    return (b"\x89PNG\r\n\x1a\n" + ch(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) +
            ch(b"cICP", bytes([primaries, transfer, 0, 1])) + ch(b"IDAT", zlib.compress(rows, 6)) + ch(b"IEND", b""))


def test_hdr_png_is_tone_mapped_by_transform(cuda_lib, ref_lib, oracle):
    """ops.go:154-165, 511-517: a PNG whose cICP chunk says PQ / BT.2020 is tone-mapped right after the decode; the
    output carries no cICP.  Both libraries, PNG output (lossless), pixels within the tone-map tolerance."""
    img = synth_image(930, 200, 150, 3, noise=8.0)
    png = _png_with_cicp(img, 9, 16)
    opt = abi.ImageOptions(FileType=".png", Width=0, Height=0, ResizeMethod=abi.ImageOpsNoResize,
                           EncodeOptions={abi.PngCompression: 3})
    a = oracle.png_decode(cuda_lib.transform(png, opt))
    b = oracle.png_decode(ref_lib.transform(png, opt))
    a = a[0] if isinstance(a, tuple) else a
    b = b[0] if isinstance(b, tuple) else b
    assert close_enough(a, b)
    assert close_enough(a, oracle.tonemap_to_sdr(img, 16, 9))
    assert not np.array_equal(a, img)
