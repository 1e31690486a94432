"""Diagnostic (not a test): time tuning variants of resize_area_kernel<3,6> at config-2 geometry.
Each variant runs in its own process because LP_RESIZE_VARIANT is read once."""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import numpy as np
    from lilliput_b200 import abi
    lib = abi.load_cuda(); l = lib.l
    n = 512
    W, H = 1920, 1080
    l.lp_dev_alloc.restype = C.c_void_p; l.lp_dev_alloc.argtypes = [C.c_size_t]
    src = l.lp_dev_alloc(n * W * H * 3); dst = l.lp_dev_alloc(n * 256 * 256 * 3)
    host = np.random.default_rng(0).integers(0, 256, W * H * 3, dtype=np.uint8)
    l.lp_memcpy_h2d.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    for i in range(0, n, 64):
        l.lp_memcpy_h2d(src + i * W * H * 3, host.ctypes.data, host.size)
    ms = C.c_float(0)
    l.lp_resize_area_time_dev.restype = C.c_int
    l.lp_resize_area_time_dev.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]
    rc = l.lp_resize_area_time_dev(src, W * H * 3, W * 3, 3, 420, 0, 1080, 1080, dst, 256 * 256 * 3, 256 * 3, 256, 256, n, 20, C.byref(ms))
    gb = n * (1080 * 1080 * 3 + 256 * 256 * 3) / 1e9
    print(f"variant={os.environ.get('LP_RESIZE_VARIANT','default'):>8s} rc={rc} ms={ms.value:.4f} GB/s={gb / (ms.value * 1e-3):8.1f} frac={gb / (ms.value * 1e-3) / 6583.5:.3f}")
else:
    for v, rpb in [("", ""), ("122", "16"), ("122", "4"), ("124", "16"), ("112", "16")]:
        env = dict(os.environ)
        if v: env["LP_RESIZE_VARIANT"] = v
        if rpb: env["LP_RESIZE_RPB"] = rpb
        print("rpb", rpb or "8", end=" ", flush=True)
        subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env)
