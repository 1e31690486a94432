"""GPU: the CUDA path through the C ABI against the oracle and the reference-made golden vectors.
Integer / byte work is compared bit-exact; the fp32 area resize is ALSO bit-exact (same FMA chain
order), so no tolerance appears anywhere in this file."""
import numpy as np
import pytest

from lilliput_b200 import abi
from lilliput_b200.synth import synth_image
from tests.cases import BLEND_CASES, JPEG_CASES, ORIENT_SRC, RESIZE_CASES

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", RESIZE_CASES, ids=lambda c: f"seed{c[0]}")
def test_resize_matches_reference_golden(cuda_lib, golden, oracle, case):
    seed, sw, sh, ch, crop, dw, dh, interp = case
    img = synth_image(seed, sw, sh, ch, noise=12.0)
    got = cuda_lib.resize(img, dw, dh, crop=crop, interpolation=interp)
    assert np.array_equal(got, golden[f"resize_{seed}"])
    assert np.array_equal(got, oracle.resize(img, dw, dh, crop=crop, interpolation=interp))


def test_resize_random_geometries_vs_oracle(cuda_lib, oracle):
    rng = np.random.default_rng(77)
    for _ in range(40):
        sw, sh = int(rng.integers(2, 900)), int(rng.integers(2, 900))
        dw, dh = int(rng.integers(1, 520)), int(rng.integers(1, 520))
        ch = int(rng.choice([1, 3, 4]))
        img = rng.integers(0, 256, (sh, sw, ch) if ch > 1 else (sh, sw), dtype=np.uint8)
        for interp in (3, 1):
            got = cuda_lib.resize(img, dw, dh, interpolation=interp)
            exp = oracle.resize(img, dw, dh, interpolation=interp)
            assert np.array_equal(got, exp), (sw, sh, dw, dh, ch, interp)


def test_resize_edge_cases(cuda_lib, oracle):
    rng = np.random.default_rng(5)
    # 1-pixel sources/destinations, large tap counts (fallback kernel), tile edges (257, 511)
    for (sw, sh, dw, dh, ch) in [(1, 1, 1, 1, 3), (5, 5, 1, 1, 3), (1, 9, 1, 3, 4), (4000, 30, 100, 7, 3),
                                 (2600, 40, 257, 9, 3), (3000, 64, 511, 16, 4), (64, 3000, 16, 90, 1),
                                 (1080, 1080, 256, 256, 3), (2160, 2160, 512, 512, 4)]:
        img = rng.integers(0, 256, (sh, sw, ch) if ch > 1 else (sh, sw), dtype=np.uint8)
        assert np.array_equal(cuda_lib.resize(img, dw, dh), oracle.resize(img, dw, dh)), (sw, sh, dw, dh)


def test_fit_matches_oracle(cuda_lib, oracle):
    img = synth_image(5, 1920, 1080, 3)
    assert np.array_equal(cuda_lib.fit(img, 256, 256), oracle.fit(img, 256, 256))
    img = synth_image(6, 480, 854, 4)
    assert np.array_equal(cuda_lib.fit(img, 100, 300), oracle.fit(img, 100, 300))


@pytest.mark.parametrize("case", JPEG_CASES, ids=lambda c: f"seed{c[0]}")
def test_jpeg_decode_matches_reference_golden(cuda_lib, golden, case):
    seed = case[0]
    got = cuda_lib.decode(golden[f"jpeg_{seed}"].tobytes())
    assert np.array_equal(got, golden[f"jpegdec_{seed}"])


@pytest.mark.parametrize("name", ["444", "422", "440", "411", "420"])
@pytest.mark.parametrize("rst", [0, 3])
def test_jpeg_decode_sampling_and_restart(cuda_lib, golden, name, rst):
    got = cuda_lib.decode(golden[f"jpegvar_{name}_{rst}"].tobytes())
    assert np.array_equal(got, golden[f"jpegvardec_{name}_{rst}"])


@pytest.mark.parametrize("case", JPEG_CASES, ids=lambda c: f"seed{c[0]}")
def test_jpeg_encode_byte_identical(cuda_lib, golden, case):
    seed, w, h, ch, q = case
    img = synth_image(seed, w, h, ch, noise=8.0)
    got = cuda_lib.encode(".jpeg", img, {abi.JpegQuality: q})
    assert got == golden[f"jpeg_{seed}"].tobytes()


def test_jpeg_roundtrip_large(cuda_lib, oracle):
    """1920x1080 (config 2 size): encode on device == oracle encode; decode on device == oracle."""
    img = synth_image(1000, 1920, 1080, 3)
    enc = cuda_lib.encode(".jpeg", img, {abi.JpegQuality: 90})
    assert enc == oracle.jpeg_encode(img, 90)
    dec = cuda_lib.decode(enc)
    assert np.array_equal(dec, oracle.jpeg_decode(enc)[0])


def test_orientation(cuda_lib, golden):
    for o in range(1, 9):
        assert np.array_equal(cuda_lib.orient(ORIENT_SRC, o), golden[f"orient_{o}"])
    img = synth_image(42, 37, 23, 3, noise=10.0)
    for o in range(1, 9):
        assert np.array_equal(cuda_lib.orient(img, o), golden[f"orient3_{o}"])


def test_transform_config1(cuda_lib, golden):
    """BASELINE config 1: ferry_sunset.jpg -> Fit 256x256 JPEG q85 through lp_transform."""
    opt = abi.ImageOptions(FileType=".jpeg", Width=256, Height=256, ResizeMethod=abi.ImageOpsFit,
                           NormalizeOrientation=True, EncodeOptions={abi.JpegQuality: 85})
    out = cuda_lib.transform(golden["c1_input"].tobytes(), opt)
    assert len(out) == 11651
    assert out == golden["c1_output"].tobytes()


def test_transform_exif_orientation(cuda_lib, golden):
    opt = abi.ImageOptions(FileType=".jpeg", Width=64, Height=64, ResizeMethod=abi.ImageOpsFit,
                           NormalizeOrientation=True, EncodeOptions={abi.JpegQuality: 85})
    assert np.array_equal(cuda_lib.decode(golden["c6_input"].tobytes()), golden["c6_decoded"])
    assert cuda_lib.transform(golden["c6_input"].tobytes(), opt) == golden["c6_output"].tobytes()


def test_transform_errors(cuda_lib):
    opt = abi.ImageOptions(FileType=".jpeg", Width=8, Height=8, ResizeMethod=abi.ImageOpsFit)
    with pytest.raises(abi.LilliputError) as e:
        cuda_lib.transform(b"not an image at all", opt)
    assert e.value.code == -1  # ErrInvalidImage
    img = synth_image(3, 64, 64, 3)
    data = cuda_lib.encode(".jpeg", img, {abi.JpegQuality: 85})
    with pytest.raises(abi.LilliputError) as e:  # destination too small -> ErrBufTooSmall
        cuda_lib.transform(data, abi.ImageOptions(FileType=".jpeg", Width=32, Height=32,
                                                  ResizeMethod=abi.ImageOpsFit), dst_cap=100)
    assert e.value.code == -3
    with pytest.raises(abi.LilliputError) as e:  # frame larger than ImageOps buffers
        cuda_lib.transform(data, abi.ImageOptions(FileType=".jpeg", Width=32, Height=32,
                                                  ResizeMethod=abi.ImageOpsFit), max_size=32)
    assert e.value.code == -3
