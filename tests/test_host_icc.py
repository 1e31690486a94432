"""CPU: ICC extraction behind the C ABI -- opencv_decoder_get_jpeg_icc / opencv_decoder_get_png_icc (host-only in
both libraries) -- product library against the reference's own shims (oracle/_ref: libjpeg-turbo's
jpeg_read_icc_profile, libpng 1.6.47's png_read_info + png_get_iCCP; ref opencv.cpp:253-345), byte for byte, on
hand-built files and on seeded random mutants.  The reference's TestICC (opencv_test.go:222-293) pins presence /
absence on four fixtures; what a decoder library does with a malformed APP2 sequence or iCCP chunk is pinned only
by running it, which is what this file does.  Every rule the product follows (png_parse.cpp: png_extract_icc)
was put there because one of these cases disagreed."""
import ctypes as C
import random
import struct
import zlib

import numpy as np
import pytest

from lilliput_b200 import abi
from lilliput_b200.synth import synth_image

MAGIC = bytes([0x89, 0x50, 0x4E, 0x47, 0x0D, 0x0A, 0x1A, 0x0A])


def _bind(lib):
    l = lib.l
    for f in (l.opencv_decoder_get_jpeg_icc, l.opencv_decoder_get_png_icc):
        f.restype, f.argtypes = C.c_int, [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]

    def call(fn, b, cap):
        out = C.create_string_buffer(max(cap, 1))
        n = fn(b, len(b), out, cap)
        return out.raw[:n] if n > 0 else n
    return (lambda b, cap=32768: call(l.opencv_decoder_get_jpeg_icc, b, cap),
            lambda b, cap=32768: call(l.opencv_decoder_get_png_icc, b, cap))


@pytest.fixture(scope="module")
def both(ref_lib):
    return _bind(abi.load_cuda()), _bind(ref_lib)


def _quiet(capfd):
    capfd.readouterr()          # libjpeg / libpng report every refusal on stderr


# ------------------------------------------------------------------------------------------- JPEG

def _app2(seq, cnt, body, tag=b"ICC_PROFILE\0"):
    payload = tag + bytes([seq, cnt]) + body
    return b"\xff\xe2" + struct.pack(">H", len(payload) + 2) + payload


def _jpeg_cases(oracle, golden):
    base = oracle.jpeg_encode(synth_image(3, 40, 24, 3), 85)
    prof = bytes(range(256)) * 3
    sos = base.index(b"\xff\xda")

    def w(*segs, at=2):
        return base[:at] + b"".join(segs) + base[at:]
    return {
        "fixture_with_icc": golden["c1_input"].tobytes(),
        "none": base,
        "one": w(_app2(1, 1, prof)),
        "three": w(_app2(1, 3, prof[:100]), _app2(2, 3, prof[100:300]), _app2(3, 3, prof[300:])),
        "out_of_order": w(_app2(3, 3, prof[300:]), _app2(1, 3, prof[:100]), _app2(2, 3, prof[100:300])),
        "missing_chunk": w(_app2(1, 3, prof[:100]), _app2(3, 3, prof[300:])),
        "duplicate_seq": w(_app2(1, 2, prof[:100]), _app2(1, 2, prof[:100])),
        "count_mismatch": w(_app2(1, 2, prof[:100]), _app2(2, 3, prof[100:])),
        "seq_zero": w(_app2(0, 1, prof)),
        "seq_past_count": w(_app2(2, 1, prof)),
        "empty_body": w(_app2(1, 1, b"")),
        "empty_first_of_two": w(_app2(1, 2, b""), _app2(2, 2, prof)),
        "other_app2": w(_app2(1, 1, prof, tag=b"FPXR\0\0\0\0\0\0\0\0")),
        "short_app2": w(b"\xff\xe2\x00\x06ICC_"),
        "late_in_header": w(_app2(1, 1, prof), at=sos),
        "larger_than_dest": w(_app2(1, 1, bytes(40000))),
        "two_segments_65k": w(_app2(1, 2, bytes(65519)), _app2(2, 2, bytes(100))),
        "cut_before_sos": w(_app2(1, 1, prof))[:sos + len(prof) // 2],
        "cut_inside_icc": w(_app2(1, 1, prof))[:200],
        "ff_fill_bytes": w(b"\xff\xff\xff" + _app2(1, 1, prof)),
        "not_jpeg": MAGIC + bytes(100),
        "soi_only": b"\xff\xd8",
        "255_chunks": w(*[_app2(i, 255, bytes([i])) for i in range(1, 256)]),
    }


def test_jpeg_icc_matches_the_reference(both, oracle, golden, capfd):
    (pj, _), (rj, _) = both
    cases = _jpeg_cases(oracle, golden)
    for name, data in cases.items():
        for cap in (32768, 768, 767, 100):
            assert pj(data, cap) == rj(data, cap), (name, cap)
    got = pj(cases["fixture_with_icc"])
    assert isinstance(got, bytes) and len(got) > 128 and got[36:40] == b"acsp"     # the reference's TestICC: present
    assert pj(cases["none"]) == 0                                                 # ... and absent
    assert pj(cases["out_of_order"]) == bytes(range(256)) * 3                     # reassembled by sequence number
    _quiet(capfd)


# -------------------------------------------------------------------------------------------- PNG

def _chunk(t, body, crc=None):
    return struct.pack(">I", len(body)) + t + body + struct.pack(">I", zlib.crc32(t + body) if crc is None else crc)


def _png(*extra, after=(), ctype=2):
    ihdr = _chunk(b"IHDR", struct.pack(">IIBBBBB", 2, 2, 8, ctype, 0, 0, 0))
    raw = b"".join(b"\x00" + bytes(2 * {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype]) for _ in range(2))
    return MAGIC + ihdr + b"".join(extra) + _chunk(b"IDAT", zlib.compress(raw)) + b"".join(after) + _chunk(b"IEND", b"")


def _iccp(profile, name=b"icc", method=0, level=6, raw=None):
    return _chunk(b"iCCP", name + b"\0" + bytes([method]) + (zlib.compress(profile, level) if raw is None else raw))


def _header(n, space=b"RGB ", cls=b"mntr", pcs=b"XYZ ", version=4, intent=0, sig=b"acsp", tags=0):
    p = bytearray(n)
    p[0:4] = struct.pack(">I", n)
    p[8] = version
    p[12:16], p[16:20], p[20:24], p[36:40] = cls, space, pcs, sig
    p[64:68] = struct.pack(">I", intent)
    p[68:80] = bytes([0, 0, 0xf6, 0xd6, 0, 1, 0, 0, 0, 0, 0xd3, 0x2d])           # D50
    p[128:132] = struct.pack(">I", tags)
    return bytes(p)


def _fill(p, n=400, relen=True):
    """Random tail: keeps the compressed chunk above libpng's 81 + 11 byte floor."""
    r = random.Random(len(p))
    q = bytearray(p) + bytes(r.randrange(256) for _ in range(n))
    if relen:
        q[0:4] = struct.pack(">I", len(q))
    return bytes(q)


def _png_cases(golden):
    good = _fill(_header(132))
    gray = _fill(_header(132, space=b"GRAY"))
    tagged = _fill(_header(132, tags=2) + b"desc" + struct.pack(">II", 156, 40) + b"cprt" + struct.pack(">II", 196, 60), 124)
    tag_out = _fill(_header(132, tags=1) + b"desc" + struct.pack(">II", 500, 400), 124)
    c = {
        "fixture_with_icc": golden["png_fixture_ferry"].tobytes(),
        "no_iccp": golden["png_rgb"].tobytes(),
        "good": _png(_iccp(good)),
        "stored_deflate": _png(_iccp(good, level=0)),
        "tiny_chunk_under_92_bytes": _png(_iccp(_header(132))),                   # valid profile, "too short" chunk
        "zeros_4k_level9": _png(_iccp(_header(4096), level=9)),
        "larger_than_dest": _png(_iccp(_fill(_header(132), 40000))),
        "junk": _png(_iccp(bytes(range(200)) * 3)),
        "shorter_than_header": _png(_iccp(good[:100] + bytes(300), level=0)),
        "extra_data_after_profile": _png(_iccp(good + bytes(8))),                 # allowed: declared length is returned
        "declared_longer_than_stream": _png(_iccp(_fill(_header(132), 400, relen=False)[:-1] + b"\0")[:0] + _iccp(
            struct.pack(">I", 600) + good[4:])),
        "bad_method": _png(_iccp(good, method=1)),
        "bad_zlib_header": _png(_iccp(good, raw=b"\x78\x9c\xff\xff\xff\xff" + bytes(100))),
        "window_too_large": _png(_iccp(good, raw=b"\x88\x1c" + zlib.compress(good)[2:])),
        "cut_zlib": _png(_iccp(good, raw=zlib.compress(good)[:-40])),
        "adler_wrong": _png(_iccp(good, raw=zlib.compress(good)[:-4] + b"\0\0\0\0")),
        "empty_keyword": _png(_iccp(good, name=b"")),
        "keyword_80": _png(_iccp(good, name=b"n" * 80)),
        "keyword_79": _png(_iccp(good, name=b"n" * 79)),
        "after_idat": _png(after=(_iccp(good),)),
        "after_plte": _png(_chunk(b"PLTE", bytes(9)), _iccp(good), ctype=3),
        "before_plte": _png(_iccp(good), _chunk(b"PLTE", bytes(9)), ctype=3),
        "after_plte_in_rgb": _png(_chunk(b"PLTE", bytes(9)), _iccp(good)),
        "after_ignored_plte_in_gray": _png(_chunk(b"PLTE", bytes(9)), _iccp(gray), ctype=0),
        "after_invalid_plte_in_rgb": _png(_chunk(b"PLTE", bytes(10)), _iccp(good)),
        "invalid_plte_in_palette_image": _png(_iccp(good), _chunk(b"PLTE", bytes(10)), ctype=3),
        "empty_plte_after": _png(_iccp(good), _chunk(b"PLTE", b"")),
        "plte_crc_error_after": _png(_iccp(good), _chunk(b"PLTE", bytes(9), crc=1)),
        "two_plte_after": _png(_iccp(good), _chunk(b"PLTE", bytes(9)), _chunk(b"PLTE", bytes(9))),
        "two_iccp_first_wins": _png(_iccp(good), _iccp(_fill(_header(132), 200))),
        "two_iccp_first_invalid": _png(_iccp(bytes(range(200)) * 3), _iccp(good)),
        "iccp_crc_error": _png(_chunk(b"iCCP", b"icc\0\0" + zlib.compress(good), crc=0)),
        "ihdr_crc_error": MAGIC + _chunk(b"IHDR", struct.pack(">IIBBBBB", 2, 2, 8, 2, 0, 0, 0), crc=7) + _png(_iccp(good))[33:],
        "rgb_profile_on_gray": _png(_iccp(good), ctype=0),
        "gray_profile_on_gray_alpha": _png(_iccp(gray), ctype=4),
        "gray_profile_on_rgb": _png(_iccp(gray)),
        "rgb_profile_on_palette": _png(_iccp(good), _chunk(b"PLTE", bytes(9)), ctype=3),
        "cmyk_profile": _png(_iccp(_fill(_header(132, space=b"CMYK")))),
        "abstract_class": _png(_iccp(_fill(_header(132, cls=b"abst")))),
        "link_class": _png(_iccp(_fill(_header(132, cls=b"link")))),
        "named_colour_class": _png(_iccp(_fill(_header(132, cls=b"nmcl")))),
        "unknown_class": _png(_iccp(_fill(_header(132, cls=b"zzzz")))),
        "lab_pcs": _png(_iccp(_fill(_header(132, pcs=b"Lab ")))),
        "bad_pcs": _png(_iccp(_fill(_header(132, pcs=b"Luv ")))),
        "bad_signature": _png(_iccp(_fill(_header(132, sig=b"acsq")))),
        "intent_ffff": _png(_iccp(_fill(_header(132, intent=0xffff)))),
        "intent_7": _png(_iccp(_fill(_header(132, intent=7)))),
        "v4_length_not_multiple_of_4": _png(_iccp(_fill(_header(132), 401))),
        "v2_length_not_multiple_of_4": _png(_iccp(_fill(_header(132, version=2), 401))),
        "tag_table": _png(_iccp(tagged)),
        "tag_outside_profile": _png(_iccp(tag_out)),
        "tag_count_huge": _png(_iccp(_fill(_header(132, tags=0x20000000)))),
        "srgb_chunk_first": _png(_chunk(b"sRGB", b"\0"), _iccp(good)),
        "gama_chrm_first": _png(_chunk(b"gAMA", struct.pack(">I", 45455)), _chunk(b"cHRM", bytes(32)), _iccp(good)),
        "ancillary_crc_errors_around": _png(_chunk(b"tEXt", b"k\0v", crc=5), _iccp(good), _chunk(b"gAMA", bytes(4), crc=5)),
        "unknown_critical_chunk": _png(_iccp(good), _chunk(b"ZZZZ", bytes(4))),
        "reserved_bit_chunk_name": _png(_iccp(good), _chunk(b"gAvA", bytes(4))),
        "non_letter_chunk_name": _png(_iccp(good), _chunk(b"gA1A", bytes(4))),
        "iend_before_idat": MAGIC + _chunk(b"IHDR", struct.pack(">IIBBBBB", 2, 2, 8, 2, 0, 0, 0)) + _iccp(good) + _chunk(b"IEND", b""),
        "ends_before_idat": _png(_iccp(good))[:33 + len(_iccp(good))],
        "ends_inside_idat_header": _png(_iccp(good))[:33 + len(_iccp(good)) + 6],
        "ends_inside_idat_body": _png(_iccp(good))[:33 + len(_iccp(good)) + 10],   # png_read_info never reads it
        "ends_inside_iccp": _png(_iccp(good))[:90],
        "zero_width": MAGIC + _chunk(b"IHDR", struct.pack(">IIBBBBB", 0, 2, 8, 2, 0, 0, 0)) + _png(_iccp(good))[33:],
        "width_over_user_limit": MAGIC + _chunk(b"IHDR", struct.pack(">IIBBBBB", 1000001, 2, 8, 2, 0, 0, 0)) + _png(_iccp(good))[33:],
        "bad_bit_depth": MAGIC + _chunk(b"IHDR", struct.pack(">IIBBBBB", 2, 2, 3, 2, 0, 0, 0)) + _png(_iccp(good))[33:],
        "rgb_4_bit": MAGIC + _chunk(b"IHDR", struct.pack(">IIBBBBB", 2, 2, 4, 2, 0, 0, 0)) + _png(_iccp(good))[33:],
        "interlace_2": MAGIC + _chunk(b"IHDR", struct.pack(">IIBBBBB", 2, 2, 8, 2, 0, 0, 2)) + _png(_iccp(good))[33:],
        "first_chunk_not_ihdr": MAGIC + _iccp(good) + _png()[8:],
        "chunk_length_over_2g": MAGIC + _png()[8:33] + struct.pack(">I", 0x80000000) + b"tEXt" + bytes(40),
        "not_png": b"\xff\xd8" + bytes(50),
        "empty": b"",
    }
    return c, good


def test_png_icc_matches_the_reference(both, golden, capfd):
    (_, pp), (_, rp) = both
    cases, good = _png_cases(golden)
    for name, data in cases.items():
        for cap in (32768, 1 << 20, len(good), len(good) - 1):
            assert pp(data, cap) == rp(data, cap), (name, cap)
    # a few answers spelled out, so that the two sides cannot agree on being wrong in the same way
    assert pp(cases["good"]) == good and pp(cases["stored_deflate"]) == good
    assert pp(cases["extra_data_after_profile"]) == good
    assert pp(cases["adler_wrong"]) == good
    assert pp(cases["after_ignored_plte_in_gray"]) != 0 and pp(cases["after_plte_in_rgb"]) == 0
    assert pp(cases["tiny_chunk_under_92_bytes"]) == 0 and pp(cases["keyword_79"]) == good and pp(cases["keyword_80"]) == 0
    assert pp(cases["ends_inside_idat_body"]) == good and pp(cases["ends_before_idat"]) == 0
    assert len(pp(cases["fixture_with_icc"])) > 128 and pp(cases["no_iccp"]) == 0   # the reference's TestICC
    _quiet(capfd)


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_png_icc_random_mutants_match_the_reference(both, golden, capfd, seed):
    """3 x 2500 seeded mutants of the files above: byte flips in the framing, truncation, profile-header edits
    recompressed at random levels, flips anywhere, bytes cut out of the iCCP body."""
    (_, pp), (_, rp) = both
    cases, _ = _png_cases(golden)
    seeds = [v for k, v in sorted(cases.items()) if len(v) > 60 and k not in ("larger_than_dest", "no_iccp")]
    rnd = random.Random(seed)
    accepted = 0
    for it in range(2500):
        b = bytearray(rnd.choice(seeds))
        mode = rnd.randrange(5)
        at = b.find(b"iCCP")
        if mode == 0:
            for _ in range(rnd.randrange(1, 4)):
                b[rnd.randrange(8, min(len(b), 130))] = rnd.randrange(256)
        elif mode == 1:
            b = b[:rnd.randrange(8, len(b))]
        elif mode == 2 and at >= 4:
            n = struct.unpack(">I", b[at - 4:at])[0]
            body = bytes(b[at + 4:at + 4 + n])
            k = body.find(0)
            try:
                prof = bytearray(zlib.decompress(body[k + 2:]))
            except Exception:
                continue
            if k < 0 or len(prof) < 4:
                continue
            for _ in range(rnd.randrange(1, 4)):
                prof[rnd.randrange(0, min(len(prof), 200))] = rnd.randrange(256)
            if rnd.random() < 0.5:
                prof[0:4] = struct.pack(">I", max(0, len(prof) + rnd.randrange(-8, 9)))
            b = b[:at - 4] + _chunk(b"iCCP", body[:k + 2] + zlib.compress(bytes(prof), rnd.choice([0, 1, 6, 9]))) + b[at + 8 + n:]
        elif mode == 3:
            for _ in range(rnd.randrange(1, 6)):
                b[rnd.randrange(len(b))] = rnd.randrange(256)
        elif at >= 0:
            c = rnd.randrange(at + 4, min(len(b), at + 300))
            del b[c:c + rnd.randrange(1, 20)]
        b = bytes(b)
        cap = rnd.choice([32768, 32768, 600, 132])
        want = rp(b, cap)
        assert pp(b, cap) == want, (seed, it, mode, cap)
        accepted += isinstance(want, bytes)
    assert accepted > 150          # the campaign is not all refusals
    _quiet(capfd)


# ------------------------------------------------------------------------------------------- cICP

def _bind_cicp(lib):
    l = lib.l
    l.opencv_decoder_get_png_cicp.restype = C.c_int
    l.opencv_decoder_get_png_cicp.argtypes = [C.c_char_p, C.c_size_t] + [C.POINTER(C.c_uint8)] * 4

    def get(b):
        v = [C.c_uint8(0xEE) for _ in range(4)]
        found = l.opencv_decoder_get_png_cicp(b, len(b), *[C.byref(x) for x in v])
        return tuple(x.value for x in v) if found else None
    return get


def _cicp(a, b, c, d, n=4, crc=None):
    return _chunk(b"cICP", (bytes([a, b, c, d]) + bytes(4))[:n], crc=crc)


def _cicp_cases():
    good = _fill(_header(132))
    ihdr = struct.pack(">IIBBBBB", 2, 2, 8, 2, 0, 0, 0)
    return {
        "plain": _png(),
        "sdr_p3": _png(_cicp(12, 13, 0, 1)),
        "pq_2020": _png(_cicp(9, 16, 0, 1)),
        "hlg": _png(_cicp(9, 18, 0, 0)),
        "matrix_nonzero_is_refused": _png(_cicp(1, 13, 5, 1)),
        "range_2": _png(_cicp(1, 13, 0, 2)),
        "three_bytes": _png(_cicp(1, 13, 0, 1, n=3)),
        "five_bytes": _png(_cicp(1, 13, 0, 1, n=5)),
        "all_zero": _png(_cicp(0, 0, 0, 0)),
        "all_ff": _png(_cicp(255, 255, 0, 255)),
        "two_first_wins": _png(_cicp(12, 13, 0, 1), _cicp(9, 16, 0, 1)),
        "two_first_refused_matrix": _png(_cicp(1, 13, 5, 1), _cicp(9, 16, 0, 1)),
        "two_first_crc_error": _png(_cicp(1, 13, 0, 1, crc=3), _cicp(9, 16, 0, 1)),
        "two_first_wrong_length": _png(_cicp(1, 13, 0, 1, n=5), _cicp(9, 16, 0, 1)),
        "after_plte_rgb": _png(_chunk(b"PLTE", bytes(9)), _cicp(12, 13, 0, 1)),
        "after_plte_palette": _png(_chunk(b"PLTE", bytes(9)), _cicp(12, 13, 0, 1), ctype=3),
        "before_plte_palette": _png(_cicp(12, 13, 0, 1), _chunk(b"PLTE", bytes(9)), ctype=3),
        "after_ignored_plte_gray": _png(_chunk(b"PLTE", bytes(9)), _cicp(12, 13, 0, 1), ctype=0),
        "after_idat": _png(after=(_cicp(12, 13, 0, 1),)),
        "crc_error": _png(_cicp(12, 13, 0, 1, crc=3)),
        "with_iccp_after": _png(_cicp(12, 13, 0, 1), _iccp(good)),
        "with_iccp_before": _png(_iccp(good), _cicp(12, 13, 0, 1)),
        "with_srgb": _png(_chunk(b"sRGB", b"\0"), _cicp(12, 13, 0, 1)),
        "with_mdcv_clli": _png(_cicp(9, 16, 0, 1), _chunk(b"mDCV", bytes(24)), _chunk(b"cLLI", bytes(8))),
        "gray_image": _png(_cicp(12, 13, 0, 1), ctype=0),
        "ends_inside_cicp": _png(_cicp(12, 13, 0, 1))[:40],
        "ends_before_idat": _png(_cicp(12, 13, 0, 1))[:49],
        "ends_in_idat_body": _png(_cicp(12, 13, 0, 1))[:60],
        "ihdr_crc_error": MAGIC + _chunk(b"IHDR", ihdr, crc=1) + _png(_cicp(12, 13, 0, 1))[33:],
        "unknown_critical_after": _png(_cicp(12, 13, 0, 1), _chunk(b"ZZZZ", b"")),
        "not_png": bytes(64),
    }


def test_png_cicp_matches_the_reference(ref_lib, capfd):
    p, r = _bind_cicp(abi.load_cuda()), _bind_cicp(ref_lib)
    cases = _cicp_cases()
    for name, data in cases.items():
        assert p(data) == r(data), name
    assert p(cases["sdr_p3"]) == (12, 13, 0, 1) and p(cases["two_first_wins"]) == (12, 13, 0, 1)
    assert p(cases["matrix_nonzero_is_refused"]) is None and p(cases["two_first_refused_matrix"]) is None
    assert p(cases["two_first_crc_error"]) == (9, 16, 0, 1)
    assert p(cases["after_plte_rgb"]) is None and p(cases["after_ignored_plte_gray"]) == (12, 13, 0, 1)
    assert p(cases["ends_before_idat"]) is None and p(cases["ends_in_idat_body"]) == (12, 13, 0, 1)
    rnd = random.Random(5)
    seeds = list(cases.values())
    accepted = 0
    for it in range(6000):
        b = bytearray(rnd.choice(seeds))
        mode = rnd.randrange(3)
        if mode == 0:
            for _ in range(rnd.randrange(1, 4)):
                b[rnd.randrange(8, min(len(b), 80))] = rnd.randrange(256)
        elif mode == 1:
            b = b[:rnd.randrange(8, len(b))]
        else:
            for _ in range(rnd.randrange(1, 5)):
                b[rnd.randrange(len(b))] = rnd.randrange(256)
        want = r(bytes(b))
        assert p(bytes(b)) == want, (it, mode)
        accepted += want is not None
    assert accepted > 300
    _quiet(capfd)
