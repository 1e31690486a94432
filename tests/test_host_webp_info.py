"""CPU: the WebP container as webp_decoder_create reads it (canvas size, pixel type, frame count, total duration,
loop count, background colour, ICC length; host-only in both libraries) -- product against the live reference
(libwebp 1.5.0's WebPMux behind webp.cpp:61-134) on every golden stream and on seeded mutants.  Whenever both take
the file the eight fields agree.  Which damaged containers are refused is libwebp's business and only partly
mirrored: about 1 % of the mutants are taken by one side only (ANMF offsets past the canvas, damaged VP8 frame
headers), either way round."""
import random

import numpy as np

from lilliput_b200 import abi
from tests.webp_util import webp_golden


def _info(lib, b):
    info, _, _, rc = lib.webp_frames(b, decode=False)
    return tuple(info.items()) if info else None


def test_webp_container_fields_match_the_reference(ref_lib):
    product = abi.load_cuda()
    g = webp_golden()
    seeds = {k: g[k].tobytes() for k in g.files
             if k.startswith("webp_") and g[k].dtype == np.uint8 and g[k].ndim == 1 and 16 <= g[k].size < 60000}
    assert len(seeds) >= 30
    for name, data in seeds.items():
        assert _info(product, data) == _info(ref_lib, data), name
    pool = list(seeds.values())
    rnd = random.Random(2)
    both = one_sided = 0
    for it in range(4000):
        b = bytearray(rnd.choice(pool))
        mode = rnd.randrange(3)
        if mode == 0:
            for _ in range(rnd.randrange(1, 4)):
                b[rnd.randrange(0, min(len(b), 120))] = rnd.randrange(256)
        elif mode == 1:
            b = b[:rnd.randrange(12, len(b))]
        else:
            for _ in range(rnd.randrange(1, 4)):
                b[rnd.randrange(len(b))] = rnd.randrange(256)
        p, r = _info(product, bytes(b)), _info(ref_lib, bytes(b))
        if p is not None and r is not None:
            both += 1
            assert p == r, (it, mode)
        else:
            one_sided += (p is None) != (r is None)
    assert both > 1000
    assert one_sided < 4000 * 0.03          # the acceptance sets stay close (1.2 % when this was written)


def _icc(lib, b):
    import ctypes as C
    l = lib.l
    l.opencv_mat_create_from_data.restype = C.c_void_p
    l.opencv_mat_create_from_data.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    l.webp_decoder_create.restype = C.c_void_p
    l.webp_decoder_create.argtypes = [C.c_void_p]
    l.webp_decoder_get_icc.restype = C.c_size_t
    l.webp_decoder_get_icc.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    l.webp_decoder_release.argtypes = [C.c_void_p]
    l.opencv_mat_release.argtypes = [C.c_void_p]
    a = np.frombuffer(b, dtype=np.uint8).copy()
    m = l.opencv_mat_create_from_data(a.size, 1, 0, a.ctypes.data, a.size)
    d = l.webp_decoder_create(m)
    if not d:
        l.opencv_mat_release(m)
        return None
    big, small = C.create_string_buffer(32768), C.create_string_buffer(16)
    n, n_small = l.webp_decoder_get_icc(d, big, 32768), l.webp_decoder_get_icc(d, small, 16)
    l.webp_decoder_release(d)
    l.opencv_mat_release(m)
    return big.raw[:n], n_small


def test_webp_icc_bytes_match_the_reference(ref_lib):
    """webp_decoder_get_icc (ref webp.cpp:136-160): the ICCP chunk's bytes, and what a 16-byte buffer gets."""
    product = abi.load_cuda()
    g = webp_golden()
    with_icc = 0
    for k in g.files:
        a = g[k]
        if k.startswith("webp_") and a.dtype == np.uint8 and a.ndim == 1 and a.size >= 16:
            want = _icc(ref_lib, a.tobytes())
            assert _icc(product, a.tobytes()) == want, k
            with_icc += bool(want and len(want[0]))
    assert with_icc >= 2
