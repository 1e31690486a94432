"""CPU: the product library loads without a GPU and exports every symbol the headers declare
(no compute call is made here), and the host-only container helpers behave like the reference."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "lilliput_b200", "liblilliput_b200.so")


def _declared():
    names = set()
    for h in ("lp_opencv.h", "lilliput_b200.h", "lp_giflib.h", "lp_webp.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b((?:opencv|lp|giflib|webp)_[a-z0-9_]+)\s*\(", src))
    return names


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB), "run __graft_entry__.build() first"
    out = subprocess.check_output(["nm", "-D", "--defined-only", LIB]).decode()
    have = {l.split()[-1] for l in out.splitlines()}
    missing = sorted(n for n in _declared() if n not in have)
    assert missing == []
    for g in ("CV_INTER_AREA", "CV_INTER_LINEAR", "CV_INTER_CUBIC"):  # ref opencv.hpp:53-55
        assert g in have
    # the 36 functions of the reference's opencv.hpp:61-132
    assert len([n for n in _declared() if n.startswith("opencv_")]) == 36
    # giflib.hpp:33-52 (19 functions) and webp.hpp:35-75 (23 functions)
    assert len([n for n in _declared() if n.startswith("giflib_")]) == 19
    assert len([n for n in _declared() if n.startswith("webp_")]) == 23


def test_webp_container_is_parsed_on_the_host():
    """webp_decoder_create / get_* need no device: header fields of a golden stream, and a refusal."""
    from lilliput_b200 import abi
    from tests.webp_util import webp_golden
    g = webp_golden()
    lib = abi.load_cuda()
    info, _, _, rc = lib.webp_frames(g["webp_anim_lossy"].tobytes(), decode=False)
    assert rc == 0 and [info[k] for k in ("width", "height", "pixel_type", "num_frames", "total_duration",
                                          "loop_count", "bg_color")] == [int(v) for v in g["webpinfo_anim_lossy"][:7]]
    info, _, _, rc = lib.webp_frames(g["webp_bad_truncated"].tobytes(), decode=False)
    assert info is None and rc == abi.LP_ERR_INVALID_IMAGE


def test_library_loads_and_host_only_entry_points_work():
    l = C.CDLL(LIB)
    l.lp_backend_name.restype = C.c_char_p
    assert l.lp_backend_name() == b"cuda-sm100a"
    assert C.c_int.in_dll(l, "CV_INTER_AREA").value == 3
    assert C.c_int.in_dll(l, "CV_INTER_LINEAR").value == 1
    # type helpers (ref opencv.cpp:83-96): CV_8UC3 = 16, CV_16UC4 = 26
    assert l.opencv_type_depth(16) == 8 and l.opencv_type_channels(16) == 3
    assert l.opencv_type_depth(26) == 16 and l.opencv_type_channels(26) == 4
    assert l.opencv_type_convert_depth(26, 0) == 24
    # mat over caller memory: NULL when the buffer is too small (ref opencv.cpp:29-32)
    l.opencv_mat_create_from_data.restype = C.c_void_p
    l.opencv_mat_create_from_data.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    l.opencv_mat_release.argtypes = [C.c_void_p]
    buf = np.zeros(100, dtype=np.uint8)
    assert l.opencv_mat_create_from_data(10, 10, 16, buf.ctypes.data, buf.size) is None
    m = l.opencv_mat_create_from_data(5, 5, 16, buf.ctypes.data, buf.size)
    assert m
    l.opencv_mat_get_width.argtypes = [C.c_void_p]
    l.opencv_mat_get_data.argtypes = [C.c_void_p]
    l.opencv_mat_get_data.restype = C.c_void_p
    assert l.opencv_mat_get_width(m) == 5 and l.opencv_mat_get_data(m) == buf.ctypes.data
    l.opencv_mat_release(m)
    # decoder sniffing: unknown signature -> NULL (ErrInvalidImage), JPEG header parse is host-only
    l.opencv_decoder_create.restype = C.c_void_p
    l.opencv_decoder_create.argtypes = [C.c_void_p]
    junk = np.frombuffer(b"definitely not an image", dtype=np.uint8).copy()
    jm = l.opencv_mat_create_from_data(junk.size, 1, 0, junk.ctypes.data, junk.size)
    assert l.opencv_decoder_create(jm) is None
    l.opencv_mat_release(jm)


def test_jpeg_header_and_cicp_helpers(golden):
    l = C.CDLL(LIB)
    data = golden["c6_input"].copy()  # EXIF orientation 6
    l.opencv_mat_create_from_data.restype = C.c_void_p
    l.opencv_mat_create_from_data.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    l.opencv_decoder_create.restype = C.c_void_p
    l.opencv_decoder_create.argtypes = [C.c_void_p]
    for f in ("opencv_decoder_read_header", "opencv_decoder_get_width", "opencv_decoder_get_height",
              "opencv_decoder_get_orientation", "opencv_decoder_get_pixel_type", "opencv_decoder_release"):
        getattr(l, f).argtypes = [C.c_void_p]
    l.opencv_decoder_read_header.restype = C.c_bool
    l.opencv_decoder_get_description.restype = C.c_char_p
    l.opencv_decoder_get_description.argtypes = [C.c_void_p]
    m = l.opencv_mat_create_from_data(data.size, 1, 0, data.ctypes.data, data.size)
    d = l.opencv_decoder_create(m)
    assert d and l.opencv_decoder_get_description(d) == b"JPEG"
    assert l.opencv_decoder_read_header(d)
    h, w = golden["c6_decoded"].shape[:2]
    assert (l.opencv_decoder_get_width(d), l.opencv_decoder_get_height(d)) == (w, h)
    assert l.opencv_decoder_get_orientation(d) == 6
    assert l.opencv_decoder_get_pixel_type(d) == 16
    l.opencv_decoder_release(d)
    # cICP insert + read back (ref opencv.cpp:413-464, png_cicp_test.go)
    # a well-formed 1x1 gray PNG: the reader follows libpng (tests/test_host_icc.py) and reports nothing for a
    # file whose header part png_read_info would refuse
    import struct
    import zlib

    def chunk(t, body):
        return struct.pack(">I", len(body)) + t + body + struct.pack(">I", zlib.crc32(t + body))
    png = bytearray(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", 1, 1, 8, 0, 0, 0, 0)) +
                    chunk(b"IDAT", zlib.compress(b"\x00\x7f")) + chunk(b"IEND", b""))
    cap = len(png) + 32
    arr = (C.c_uint8 * cap)(*png)
    l.opencv_png_insert_cicp.restype = C.c_size_t
    l.opencv_png_insert_cicp.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t] + [C.c_uint8] * 4
    n = l.opencv_png_insert_cicp(arr, len(png), cap, 12, 13, 0, 1)
    assert n == len(png) + 16
    out = bytes(arr[:n])
    assert out[33:37] == b"\x00\x00\x00\x04" and out[37:41] == b"cICP" and out[41:45] == bytes([12, 13, 0, 1])
    import zlib
    assert int.from_bytes(out[45:49], "big") == zlib.crc32(out[37:45])
    vals = [C.c_uint8() for _ in range(4)]
    l.opencv_decoder_get_png_cicp.argtypes = [C.c_void_p, C.c_size_t] + [C.POINTER(C.c_uint8)] * 4
    assert l.opencv_decoder_get_png_cicp(arr, n, *[C.byref(v) for v in vals]) == 1
    assert [v.value for v in vals] == [12, 13, 0, 1]
    assert l.opencv_png_insert_cicp(arr, n, n + 3, 1, 1, 1, 1) == n  # no room: unchanged


def test_hostile_headers_are_refused_not_fatal():
    """Found by tests/fuzz_probe.py: (1) a frame header this decoder does not take (12-bit) followed by a
    scan header used to divide by a zero sampling factor; (2) a header declaring a gigantic image made the
    stage helpers allocate by the header's word.  Both are host-side and must answer with an error code."""
    from lilliput_b200 import abi
    lib = abi.load_cuda()
    g = np.load(os.path.join(ROOT, "tests", "golden", "jpeg_multiscan_golden.npz"))
    data = bytearray(g["jpg_" + str(g["names"][0])].tobytes())
    i = data.find(b"\xff\xc2")
    assert i > 0
    data[i + 4] = 12  # sample precision
    try:
        lib.header(bytes(data))
    except abi.LilliputError:
        pass
    from tests.png_writer import write_png
    png = bytearray(write_png(np.zeros((2, 2, 3), np.int64), 2, 8))
    png[16:24] = (60000).to_bytes(4, "big") + (60000).to_bytes(4, "big")  # IHDR width, height ...
    import zlib
    png[29:33] = zlib.crc32(bytes(png[12:29])).to_bytes(4, "big")         # ... with its CRC put right (a bad one is ErrInvalidImage)
    with pytest.raises(abi.LilliputError) as e:
        lib.decode(bytes(png))
    assert e.value.code == abi.LP_ERR_BUF_TOO_SMALL
