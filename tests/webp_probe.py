"""Probe (not a test): wall time of the per-image WebP decode path on the device."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lilliput_b200 import abi
from tests.webp_util import webp_golden
g = webp_golden()
lib = abi.load_cuda()
for name in ["lossy118", "fixture_tears_of_steel_no_icc", "lossy102", "lossy101"]:
    data = g[f"webp_{name}"].tobytes()
    lib.webp_frames(data)
    t = time.perf_counter()
    n = 5
    for _ in range(n):
        lib.webp_frames(data)
    dt = (time.perf_counter() - t) / n
    print(f"{name}: {len(data)} B  {dt*1e3:.2f} ms/decode", flush=True)
