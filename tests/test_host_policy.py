"""The host policy layer (lilliput_b200/host/lilliput_host.cpp = ImageOps.Transform and its helpers,
ref ops.go:352-444) checked rule by rule.  The same C++ is linked twice: over the reference's own shims
(oracle/_ref, runs on the CPU -- the variant that runs in the build container) and over the CUDA library
(`-m gpu`).  Expectations are the reference's rules as written in ops.go, cited per test.
"""
import io
import os

import numpy as np
import pytest

from lilliput_b200 import abi

PIL_Image = pytest.importorskip("PIL.Image")


@pytest.fixture(params=["reference_shims", pytest.param("cuda", marks=pytest.mark.gpu)])
def lib(request):
    if request.param == "cuda":
        return request.getfixturevalue("cuda_lib")
    return request.getfixturevalue("ref_lib")


def _gif(n=5, w=48, h=36, duration_ms=70, loop=0):
    rng = np.random.default_rng(5)
    frames = []
    for k in range(n):
        a = np.full((h, w, 3), 30 + 40 * k, np.uint8)
        a[4 + 3 * k:14 + 3 * k, 6 + 5 * k:20 + 5 * k] = rng.integers(0, 256, 3, dtype=np.uint8)
        frames.append(PIL_Image.fromarray(a).quantize(32))
    bio = io.BytesIO()
    frames[0].save(bio, "GIF", save_all=True, append_images=frames[1:], duration=duration_ms, loop=loop, optimize=False)
    return bio.getvalue()


def _opts(**kw):
    kw.setdefault("EncodeTimeout_ns", 600 * 10**9)
    return abi.ImageOptions(**kw)


def _dims(lib, data):
    w, h, _, _ = lib.header(data)
    return w, h


# ---- calculateExpectedSize (ref ops.go:243-255), through Fit on a still image (ref ops.go:170-206)

@pytest.mark.parametrize("req,exp", [
    ((256, 256), (256, 256)),    # ordinary request
    ((400, 400), (297, 297)),    # square request larger than min(src): min x min
    ((1000, 900), (800, 297)),   # both larger, not square: source size
    ((1000, 200), (1000, 200)),  # only one larger: the request stands (Fit then crops to the request's aspect)
    ((300, 200), (300, 200)),
])
def test_expected_size_rules(lib, golden, req, exp):
    data = golden["c1_input"].tobytes()  # 800 x 297
    out = lib.transform(data, _opts(FileType=".jpeg", Width=req[0], Height=req[1], ResizeMethod=abi.ImageOpsFit,
                                    EncodeOptions={abi.JpegQuality: 85}))
    assert _dims(lib, out) == exp


def test_noresize_still_skips_fit(lib, golden, oracle):
    """ImageOpsNoResize on a non-animated source encodes the decoded frame as it is (ref ops.go:450-452)."""
    data = golden["c1_input"].tobytes()
    out = lib.transform(data, _opts(FileType=".png", ResizeMethod=abi.ImageOpsNoResize, Width=10, Height=10))
    assert np.array_equal(oracle.png_decode(out), oracle.jpeg_decode(data)[0])


def test_resize_stretches_to_the_request(lib, golden, oracle):
    """ImageOpsResize ignores the aspect ratio and calculateExpectedSize (ref ops.go:208-236)."""
    data = golden["c1_input"].tobytes()
    out = lib.transform(data, _opts(FileType=".png", Width=1000, Height=50, ResizeMethod=abi.ImageOpsResize))
    src = oracle.jpeg_decode(data)[0]
    assert np.array_equal(oracle.png_decode(out), oracle.resize(src, 1000, 50))


def test_orientation_is_applied_with_or_without_normalize(lib, golden, oracle):
    """normalizeOrientation runs on every iteration (ref ops.go:392); the flag only tells inputCanvasSize that
    the axes were swapped (ref ops.go:474-479) -- for a still Fit the pixels are the same either way."""
    data = golden["c6_input"].tobytes()
    src, orientation = oracle.jpeg_decode(data)
    assert orientation not in (0, 1)
    turned = oracle.orient(src, orientation)
    for flag in (False, True):
        out = lib.transform(data, _opts(FileType=".png", ResizeMethod=abi.ImageOpsNoResize, NormalizeOrientation=flag))
        assert np.array_equal(oracle.png_decode(out), turned), flag


# ---- animated sources: frame caps, single-frame output, timeouts (ref ops.go:384-390, 423-437)

def test_animated_gif_to_gif_keeps_every_frame(lib):
    out = lib.transform(_gif(5), _opts(FileType=".gif", Width=24, Height=18, ResizeMethod=abi.ImageOpsFit))
    info = lib.gif_info(out)
    assert (info["width"], info["height"], info["frame_count"]) == (24, 18, 5)


def test_disable_animated_output_encodes_one_frame(lib):
    """DisableAnimatedOutput: one frame, then the encoder is flushed (ref ops.go:423-425)."""
    out = lib.transform(_gif(5), _opts(FileType=".gif", Width=24, Height=18, ResizeMethod=abi.ImageOpsFit,
                                       DisableAnimatedOutput=True))
    assert lib.gif_info(out)["frame_count"] == 1


@pytest.mark.parametrize("cap", [1, 2, 4])
def test_max_encode_frames_skips_the_rest(lib, cap):
    """MaxEncodeFrames: after `cap` frames the decoder is skipped to EOF and the encoder flushed (ref ops.go:427-433)."""
    out = lib.transform(_gif(5), _opts(FileType=".gif", Width=24, Height=18, ResizeMethod=abi.ImageOpsFit,
                                       MaxEncodeFrames=cap))
    assert lib.gif_info(out)["frame_count"] == cap


def test_max_encode_duration_stops_before_the_frame_that_exceeds_it(lib):
    """The running duration is checked BEFORE a frame is transformed (ref ops.go:384-390): 5 frames of 70 ms and a
    200 ms cap keep the frames whose cumulative duration is <= 200 ms, i.e. two."""
    out = lib.transform(_gif(5, duration_ms=70), _opts(FileType=".gif", Width=24, Height=18,
                                                       ResizeMethod=abi.ImageOpsFit,
                                                       MaxEncodeDuration_ns=200 * 10**6))
    assert lib.gif_info(out)["frame_count"] == 2


def test_frame_cap_on_a_decoder_that_cannot_skip(lib):
    """WebP sources cannot SkipFrame: a frame cap that bites fails with ErrSkipNotSupported (ref ops.go:336-349,
    webp.go SkipFrame)."""
    webp = np.load(os.path.join(os.path.dirname(__file__), "golden", "webp_golden.npz"))["webp_anim_lossy"].tobytes()
    with pytest.raises(abi.LilliputError) as e:
        lib.transform(webp, _opts(FileType=".webp", Width=16, Height=16, ResizeMethod=abi.ImageOpsFit,
                                  MaxEncodeFrames=1, EncodeOptions={abi.WebpQuality: 80}))
    assert e.value.code == abi.LP_ERR_SKIP_NOT_SUPPORTED


def test_zero_encode_timeout_fails_after_the_first_frame_of_an_animation(lib):
    """encodeTimeoutTime = now + 0: a multi-frame encode returns ErrEncodeTimeout after its first frame
    (ref ops.go:368,435-437); a still image never reaches that check (content is returned first)."""
    with pytest.raises(abi.LilliputError) as e:
        lib.transform(_gif(3), abi.ImageOptions(FileType=".gif", Width=24, Height=18, ResizeMethod=abi.ImageOpsFit))
    assert e.value.code == abi.LP_ERR_ENCODE_TIMEOUT


def test_zero_encode_timeout_is_harmless_for_a_still(lib, golden):
    out = lib.transform(golden["c1_input"].tobytes(),
                        abi.ImageOptions(FileType=".jpeg", Width=64, Height=64, ResizeMethod=abi.ImageOpsFit,
                                         EncodeOptions={abi.JpegQuality: 85}))
    assert _dims(lib, out) == (64, 64)


# ---- errors (ref lilliput.go:19-27)

def test_garbage_is_an_invalid_image(lib):
    with pytest.raises(abi.LilliputError) as e:
        lib.transform(b"\x00" * 64, _opts(FileType=".jpeg", Width=8, Height=8, ResizeMethod=abi.ImageOpsFit))
    assert e.value.code == abi.LP_ERR_INVALID_IMAGE


def test_small_destination_is_buf_too_small(lib, golden):
    with pytest.raises(abi.LilliputError) as e:
        lib.transform(golden["c1_input"].tobytes(),
                      _opts(FileType=".jpeg", Width=256, Height=256, ResizeMethod=abi.ImageOpsFit,
                            EncodeOptions={abi.JpegQuality: 85}), dst_cap=512)
    assert e.value.code == abi.LP_ERR_BUF_TOO_SMALL


# ---- cICP colour signalling of a PNG source (ref ops.go:306-332, 511-517; the reference's own tests:
#      png_cicp_test.go:89-155).  SDR tags ride through to a PNG output; PQ / HLG tags are never re-emitted
#      (the tone-mapping the reference applies to those pixels is out of scope -- DESIGN.md s.8).

def _with_cicp(png, primaries, transfer):
    """injectPNGCICP (ref png_cicp_test.go:15-35): a cICP chunk directly after IHDR."""
    import struct
    import zlib
    body = bytes([primaries, transfer, 0, 1])
    chunk = struct.pack(">I", 4) + b"cICP" + body + struct.pack(">I", zlib.crc32(b"cICP" + body))
    at = 8 + 12 + struct.unpack(">I", png[8:12])[0]
    return png[:at] + chunk + png[at:]


def _png_chunks(png):
    """pngChunkTypes (ref png_cicp_test.go:37-46) plus the chunk bodies."""
    import struct
    out, i = [], 8
    while i + 8 <= len(png):
        n = struct.unpack(">I", png[i:i + 4])[0]
        out.append((png[i + 4:i + 8], png[i + 8:i + 8 + n]))
        i += n + 12
    return out


def _png_source(golden):
    return golden["png_rgb"].tobytes()


def _transform_png(lib, data):
    """transformPNG (ref png_cicp_test.go:57-84): Fit at the source's own size, PNG compression 7."""
    w, h = _dims(lib, data)
    return lib.transform(data, _opts(FileType=".png", Width=w, Height=h, ResizeMethod=abi.ImageOpsFit,
                                     EncodeOptions={abi.PngCompression: 7}))


def test_sdr_cicp_round_trips_to_png(lib, golden, oracle):
    """TestPNGSDRCICPRoundTrips: Display-P3 primaries + sRGB transfer is signalling only -- the chunk survives,
    right after IHDR, and the pixels are the untagged transform's."""
    src = _png_source(golden)
    plain = _transform_png(lib, src)
    tagged = _transform_png(lib, _with_cicp(src, 12, 13))
    chunks = _png_chunks(tagged)
    assert chunks[0][0] == b"IHDR" and chunks[1] == (b"cICP", bytes([12, 13, 0, 1]))
    assert len(tagged) == len(plain) + 16
    assert tagged[:33] + tagged[49:] == plain                       # nothing else in the file moved
    assert np.array_equal(oracle.png_decode(tagged), oracle.png_decode(plain))


def test_png_without_cicp_gains_none(lib, golden):
    """TestPNGWithoutCICPUnchanged."""
    assert b"cICP" not in [t for t, _ in _png_chunks(_transform_png(lib, _png_source(golden)))]


@pytest.mark.parametrize("transfer", [16, 18])
def test_hdr_cicp_tag_is_not_re_emitted(lib, golden, transfer):
    """The second half of TestPNGHDRCICPIsTonemapped: a PQ (16) or HLG (18) tag no longer describes the output
    and is dropped (ref ops.go:513-517)."""
    out = _transform_png(lib, _with_cicp(_png_source(golden), 9, transfer))
    assert b"cICP" not in [t for t, _ in _png_chunks(out)]


def test_sdr_cicp_is_png_output_only(lib, golden):
    """applyOutputCICP is a no-op for anything that does not start with the PNG signature (ref ops.go:314-316)."""
    src = _with_cicp(_png_source(golden), 12, 13)
    w, h = _dims(lib, src)
    o = _opts(FileType=".jpeg", Width=w, Height=h, ResizeMethod=abi.ImageOpsFit, EncodeOptions={abi.JpegQuality: 85})
    assert lib.transform(src, o) == lib.transform(_png_source(golden), o)


def test_sdr_cicp_needs_sixteen_spare_bytes(lib, golden):
    """The insert is skipped when the caller's buffer has no room for the chunk (ref opencv.cpp:421-424):
    the transform still succeeds, untagged."""
    src = _png_source(golden)
    plain = _transform_png(lib, src)
    w, h = _dims(lib, src)
    o = _opts(FileType=".png", Width=w, Height=h, ResizeMethod=abi.ImageOpsFit, EncodeOptions={abi.PngCompression: 7})
    out = lib.transform(_with_cicp(src, 12, 13), o, dst_cap=len(plain) + 15)
    assert out == plain


# ---- a PNG can carry an EXIF orientation too (eXIf chunk, read by OpenCV's PNG decoder like a JPEG's APP1), and
#      Transform applies it on every iteration (ref ops.go:392)

def test_png_exif_orientation_is_applied(lib, golden, oracle):
    import struct
    import zlib
    src = _png_source(golden)
    tiff = b"II*\0" + struct.pack("<IH", 8, 1) + struct.pack("<HHI", 0x0112, 3, 1) + struct.pack("<HH", 6, 0) + bytes(4)
    chunk = struct.pack(">I", len(tiff)) + b"eXIf" + tiff + struct.pack(">I", zlib.crc32(b"eXIf" + tiff))
    at = 8 + 12 + struct.unpack(">I", src[8:12])[0]
    turned = src[:at] + chunk + src[at:]
    w, h, _, o = lib.header(turned)
    assert (w, h, o) == (*_dims(lib, src), 6)
    plain = oracle.png_decode(lib.transform(src, _opts(FileType=".png", ResizeMethod=abi.ImageOpsNoResize)))
    out = oracle.png_decode(lib.transform(turned, _opts(FileType=".png", ResizeMethod=abi.ImageOpsNoResize)))
    assert out.shape[:2] == (w, h)
    assert np.array_equal(out, np.rot90(plain, k=-1))          # orientation 6 = 90 degrees clockwise


@pytest.mark.parametrize("value", [0, 9, 300, 65535])
def test_out_of_range_exif_orientation_is_reported_and_does_nothing(lib, oracle, value):
    """The reader passes the EXIF word through unvalidated (tests/test_host_exif.py); OrientationTransform then
    has no case for it (ref opencv.cpp:217-221) and the frame is encoded as decoded."""
    import struct
    from lilliput_b200.synth import synth_image
    base = oracle.jpeg_encode(synth_image(3, 40, 24, 3), 85)
    tiff = b"II*\0" + struct.pack("<IH", 8, 1) + struct.pack("<HHI", 0x0112, 3, 1) + struct.pack("<HH", value, 0) + bytes(4)
    seg = b"Exif\0\0" + tiff
    tagged = base[:2] + b"\xff\xe1" + struct.pack(">H", len(seg) + 2) + seg + base[2:]
    assert lib.header(tagged) == (40, 24, 16, value)
    o = _opts(FileType=".png", ResizeMethod=abi.ImageOpsNoResize, NormalizeOrientation=True)
    assert lib.transform(tagged, o) == lib.transform(base, o)


@pytest.mark.parametrize("damage", ["lost_quantisation_table", "table_selector_out_of_range"])
def test_jpeg_with_a_missing_table_reads_its_header_and_fails_to_decode(lib, oracle, damage):
    """libjpeg checks tables when the decode starts, not when the header is read: Header() succeeds and DecodeTo
    is what fails (ErrDecodingFailed, ref opencv.go:829-831) -- not ErrInvalidImage from the header."""
    from lilliput_b200.synth import synth_image
    b = bytearray(oracle.jpeg_encode(synth_image(3, 40, 24, 3), 85))
    if damage == "lost_quantisation_table":
        b[b.find(b"\xff\xdb")] = 0x3F                  # the marker's FF is gone: the segment is skipped as garbage
    else:
        b[b.find(b"\xff\xc0") + 12] = 151              # Tq of the first component
    assert lib.header(bytes(b)) == (40, 24, 16, 1)
    with pytest.raises(abi.LilliputError) as e:
        lib.decode(bytes(b))
    assert e.value.code == -2
    with pytest.raises(abi.LilliputError) as e:
        lib.transform(bytes(b), _opts(FileType=".jpeg", Width=16, Height=16, ResizeMethod=abi.ImageOpsFit))
    assert e.value.code == -2


@pytest.mark.parametrize("lost", ["first", "all"])
def test_jpeg_without_huffman_tables_decodes_with_the_annex_k_tables(lib, oracle, lost):
    """Motion-JPEG frames are written without DHT segments; libjpeg-turbo fills the undefined slots 0 and 1 with the
    Annex K tables when the decode starts (jdhuff.c: jinit_huff_decoder -> std_huff_tables), so such a file -- or
    one whose DHT was lost to damage -- decodes in the reference.  The source here was encoded with those tables."""
    from lilliput_b200.synth import synth_image
    base = oracle.jpeg_encode(synth_image(3, 40, 24, 3), 85)
    b = bytearray(base)
    at = b.find(b"\xff\xc4")
    while at >= 0:
        b[at] = 0x3F                                   # no longer a marker: skipped as garbage
        at = b.find(b"\xff\xc4", at + 1) if lost == "all" else -1
    assert lib.header(bytes(b)) == (40, 24, 16, 1)
    assert np.array_equal(lib.decode(bytes(b)), oracle.jpeg_decode(base)[0])
