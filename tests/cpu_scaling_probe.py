"""Diagnostic (not a test): reference CPU path throughput vs thread count on this host."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lilliput_b200 import abi
from lilliput_b200.synth import synth_image
from oracle import oracle
ref = abi.load_reference()
l = ref.l
l.ref_transform_many.restype = C.c_double
l.ref_transform_many.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_long, C.POINTER(C.c_int)]
files = [np.frombuffer(oracle.jpeg_encode(synth_image(1000 + i, 1920, 1080, 3), 90), dtype=np.uint8) for i in range(16)]
n = len(files)
ptrs = (C.c_void_p * n)(*[f.ctypes.data for f in files]); lens = (C.c_size_t * n)(*[f.size for f in files])
opt = abi.ImageOptions(FileType=".jpeg", Width=256, Height=256, ResizeMethod=abi.ImageOpsFit, NormalizeOrientation=True, EncodeOptions={abi.JpegQuality: 85})._c()
print("cpu.max:", open("/sys/fs/cgroup/cpu.max").read().strip() if os.path.exists("/sys/fs/cgroup/cpu.max") else "n/a", "affinity:", len(os.sched_getaffinity(0)))
for t in (1, 4, 16, 32, 64, 128):
    err = C.c_int(0)
    l.ref_transform_many(ptrs, lens, n, C.byref(opt), 2048, t, 1 << 20, t * 2, C.byref(err))  # warm
    total = t * 40
    el = l.ref_transform_many(ptrs, lens, n, C.byref(opt), 2048, t, 1 << 20, total, C.byref(err))
    print(f"threads={t:4d} images={total:5d} s={el:.2f} img/s={total/el:8.1f} per-thread={total/el/t:6.1f}")
