"""CPU: the reference's own unit tests for its host-side container sniffers (opencv_test.go:9-220 --
TestAPNG, TestContentLength_*, TestPNGWalk_*), replayed byte array by byte array against the C++ mirror
(lilliput_host.cpp: detectAPNG, detectContentLength, the PNG chunk walker; ref opencv.go:467-637) through
lp_detect_apng / lp_detect_content_length / lp_png_chunk_types.  The product library answers these on the
host -- no device call -- so they run here without a GPU; the same mirror linked over the reference's shims
(oracle/_ref) is checked beside it.  Also the two cICP helpers of the ABI (ref opencv.cpp:397-464)."""
import ctypes as C
import struct
import zlib

import pytest

from lilliput_b200 import abi

PNG_MAGIC = bytes([0x89, 0x50, 0x4E, 0x47, 0x0D, 0x0A, 0x1A, 0x0A])
IHDR0 = bytes([0, 0, 0, 0]) + b"IHDR" + bytes([0, 0, 0, 0])            # size, type, crc (opencv_test.go:10-14)
FAKE4 = bytes([0, 0, 0, 4, 1, 2, 3, 4, 8, 9, 8, 9, 0, 0, 0, 0])        # size 4, type (not real), data, crc


@pytest.fixture(params=["product", "reference_shims"])
def sniff(request):
    lib = abi.load_cuda() if request.param == "product" else request.getfixturevalue("ref_lib")
    l = lib.l
    for f in (l.lp_detect_apng, l.lp_detect_content_length):
        f.restype, f.argtypes = C.c_int, [C.c_char_p, C.c_size_t]
    l.lp_png_chunk_types.restype = C.c_int
    l.lp_png_chunk_types.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_int]

    class S:
        @staticmethod
        def apng(b):
            return bool(l.lp_detect_apng(b, len(b)))

        @staticmethod
        def content_length(b):
            return l.lp_detect_content_length(b, len(b))

        @staticmethod
        def chunks(b, cap=16):
            out = C.create_string_buffer(4 * cap)
            n = l.lp_png_chunk_types(b, len(b), out, cap)
            return None if n < 0 else [out.raw[4 * i:4 * i + 4] for i in range(min(n, cap))]
    return S


def test_apng(sniff):                                                   # opencv_test.go:9-35
    png = PNG_MAGIC + IHDR0
    assert not sniff.apng(png)
    for chunk in (b"acTL", b"fcTL", b"fdAT"):
        assert sniff.apng(png + bytes(4) + chunk + bytes(4))
    assert not sniff.apng(IHDR0)                                        # not a PNG at all
    assert not sniff.apng(b"")


def test_content_length_png_extra_data(sniff):                          # opencv_test.go:37-62
    png = PNG_MAGIC + IHDR0 + FAKE4 + bytes([0, 0, 0, 0, 7, 7, 7, 7, 0, 0, 0, 0])
    assert sniff.content_length(png) == len(png)
    png += bytes([56, 56])
    assert sniff.content_length(png) == len(png)


def test_content_length_png_iend(sniff):                                # opencv_test.go:64-88
    png = PNG_MAGIC + IHDR0 + bytes(4) + b"IEND" + bytes(4)
    assert sniff.content_length(png + FAKE4) == len(png)
    # an IEND whose declared body runs past the data: clamped to the data (opencv.go:521-524)
    cut = PNG_MAGIC + IHDR0 + bytes([0, 0, 1, 0]) + b"IEND" + bytes(4)
    assert sniff.content_length(cut) == len(cut)


def test_content_length_jpeg_extra_data(sniff):                         # opencv_test.go:90-111
    jpeg = bytes([0xFF, 0xD8,
                  0xFF, 0xE7, 0x00, 0x04, 0xFF, 0xD9,                   # made-up segment (holds an EOI look-alike)
                  0xFF, 0xDA, 0x00, 0x04, 0x00, 0x00,                   # SOS
                  0x00, 0x01, 0xD9, 0xFF, 0xD5, 0xD5,                   # ECS data (FF D5 = RST5 continues it)
                  0xFF, 0xD9])
    assert sniff.content_length(jpeg) == len(jpeg)
    assert sniff.content_length(jpeg + bytes([0xFF, 0xC2, 0x00, 0x02])) == len(jpeg)


def test_content_length_jpeg_entropy_coding(sniff):                     # opencv_test.go:113-127
    jpeg = bytes([0xFF, 0xD8,
                  0xFF, 0xE7, 0x00, 0x04, 0xFF, 0xD9,
                  0xFF, 0xDA, 0x00, 0x02,
                  0x02, 0x01, 0xFF, 0x00, 0xD9,                         # stuffed FF 00 inside the ECS
                  0xFF, 0xFF,                                           # padding
                  0xFF, 0xD9,
                  0x01])                                                # extra
    assert sniff.content_length(jpeg) == len(jpeg) - 1


def test_content_length_unrecognized_and_truncated(sniff):              # opencv_test.go:129-135
    assert sniff.content_length(bytes(128)) == 128
    assert sniff.content_length(b"") == 0
    # a JPEG that ends inside a sized segment, and one that ends on a lone FF inside the scan: full length
    assert sniff.content_length(bytes([0xFF, 0xD8, 0xFF, 0xE0, 0x00])) == 5
    cut = bytes([0xFF, 0xD8, 0xFF, 0xDA, 0x00, 0x02, 0x11, 0x22, 0xFF])
    assert sniff.content_length(cut) == len(cut)


def test_png_walk_extra_data(sniff):                                    # opencv_test.go:158-179
    png = PNG_MAGIC + IHDR0 + FAKE4
    for _ in range(11):                                                 # min chunk size is 12
        png += bytes(1)
        assert sniff.chunks(png) == [b"IHDR", bytes([1, 2, 3, 4])]
    assert sniff.chunks(png + bytes(1)) == [b"IHDR", bytes([1, 2, 3, 4]), bytes(4)]


def test_png_walk_bad_size(sniff):                                      # opencv_test.go:181-198
    png = PNG_MAGIC + IHDR0 + bytes([0, 128, 0, 4, 1, 2, 3, 4, 8, 9, 8, 9, 0, 0, 0, 0])
    assert sniff.chunks(png) == [b"IHDR", bytes([1, 2, 3, 4])]
    huge = PNG_MAGIC + bytes([0xFF, 0xFF, 0xFF, 0xFF]) + b"IHDR" + bytes(4) + FAKE4   # 4 GiB body: walk just ends
    assert sniff.chunks(huge) == [b"IHDR"]


def test_png_walk_not_png(sniff):                                       # opencv_test.go:200-211
    assert sniff.chunks(IHDR0) is None


def test_png_walk_no_chunks(sniff):                                     # opencv_test.go:213-221
    png = PNG_MAGIC
    for _ in range(12):
        assert sniff.chunks(png) == []
        png += bytes(1)


def test_png_walk_reports_count_past_cap(sniff):
    png = PNG_MAGIC + IHDR0 * 5
    assert sniff.chunks(png, cap=2) == [b"IHDR", b"IHDR"]


# ---- cICP helpers of the per-image ABI (host-only in both libraries)

def _chunk(t, body):
    return struct.pack(">I", len(body)) + t + body + struct.pack(">I", zlib.crc32(t + body))


def _tiny_png(extra=b""):
    ihdr = _chunk(b"IHDR", struct.pack(">IIBBBBB", 1, 1, 8, 0, 0, 0, 0))
    idat = _chunk(b"IDAT", zlib.compress(b"\x00\x7f"))
    return PNG_MAGIC + ihdr + extra + idat + _chunk(b"IEND", b"")


@pytest.fixture(params=["product", "reference_shims"])
def cicp(request):
    lib = abi.load_cuda() if request.param == "product" else request.getfixturevalue("ref_lib")
    l = lib.l
    l.opencv_decoder_get_png_cicp.restype = C.c_int
    l.opencv_decoder_get_png_cicp.argtypes = [C.c_char_p, C.c_size_t] + [C.POINTER(C.c_uint8)] * 4
    l.opencv_png_insert_cicp.restype = C.c_size_t
    l.opencv_png_insert_cicp.argtypes = [C.c_char_p, C.c_size_t, C.c_size_t] + [C.c_uint8] * 4

    class K:
        @staticmethod
        def get(b):
            v = [C.c_uint8(0) for _ in range(4)]
            found = l.opencv_decoder_get_png_cicp(b, len(b), *[C.byref(x) for x in v])
            return tuple(x.value for x in v) if found else None

        @staticmethod
        def insert(b, cap, tag):
            buf = C.create_string_buffer(b, cap)
            n = l.opencv_png_insert_cicp(buf, len(b), cap, *tag)
            return buf.raw[:n]
    return K


def test_cicp_is_read_before_idat_only(cicp):                           # ref opencv.cpp:397-411
    tag = _chunk(b"cICP", bytes([12, 13, 0, 1]))
    assert cicp.get(_tiny_png()) is None
    assert cicp.get(_tiny_png(tag)) == (12, 13, 0, 1)
    late = _tiny_png()
    iend = late.rindex(b"IEND") - 4
    assert cicp.get(late[:iend] + tag + late[iend:]) is None            # after IDAT: not part of the header
    assert cicp.get(b"not a png at all") is None


def test_cicp_insert_goes_right_after_ihdr_with_a_valid_crc(cicp):      # ref opencv.cpp:413-464
    png = _tiny_png()
    out = cicp.insert(png, len(png) + 16, (9, 16, 0, 1))
    assert len(out) == len(png) + 16
    assert out[:33] == png[:33] and out[49:] == png[33:]
    assert out[33:49] == _chunk(b"cICP", bytes([9, 16, 0, 1]))
    assert cicp.get(out) == (9, 16, 0, 1)
    assert cicp.insert(png, len(png) + 15, (9, 16, 0, 1)) == png         # no room: left as it was
    assert cicp.insert(b"\x00" * 40, 64, (1, 1, 0, 1)) == b"\x00" * 40   # not a PNG: left as it was
