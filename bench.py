#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline: images/sec, 1920x1080 JPEG -> Fit 256x256 JPEG q85.

One "step" = one pass of the hot path (decode -> Fit/area resize -> encode) over one batch of
4096 synthetic 1080p baseline JPEGs per GPU (BASELINE config 2).  Reported on one JSON line:

  value      whole-job images/s with the compressed inputs already resident in HBM (lp_batch_run,
             CUDA-event timed inside the library on its own stream, max over ranks)
  e2e        the same metric through the reference-facing batch call lp_batch_transform with HOST
             (pinned) buffers: header parse + H2D + kernels + D2H inside the timed region
  roofline   the area-resize kernel: algorithmic bytes per launch / CUDA-event launch time,
             against MEASURED_PEAKS.json's HBM copy bandwidth
  cpu_baseline  the reference's own CPU path (oracle/_ref = lilliput's opencv.cpp shims + vendored
             libs, driven by the ops.go mirror) on the box's host cores, bounded sample, N=1 only

`--impl reference` times only that CPU path (all host threads) and prints the same line shape.
Multi-GPU: one process per GPU under torchrun; images are sharded by index, no collective on the
data path (weak scaling: 4096 images per GPU).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SRC_W, SRC_H, DST, Q_IN, Q_OUT = 1920, 1080, 256, 90, 85
CROP = 1080
RESIZE_BYTES_PER_IMAGE = CROP * CROP * 3 + DST * DST * 3  # 3 695 808 B


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="images per GPU per step")
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5],
                    help="BASELINE.json config: 2 = headline (1080p JPEG -> 256x256 JPEG), 3 = 4K RGBA PNG -> 512x512 "
                         "WebP, 4 = 128-frame 720p GIF -> 256x256 animated WebP, 5 = mixed JPEG/PNG/WebP -> 256x256 JPEG")
    ap.add_argument("--distinct", type=int, default=0, help="configs 3-5: distinct files generated (replicated to the batch)")
    ap.add_argument("--variant", default="default", choices=["default", "cv2", "optimized", "dri", "pcg64"],
                    help="config 2 corpus: default = this library's encoder (byte-identical to the reference's); cv2 = "
                         "OpenCV / libjpeg-turbo written files; optimized = per-image optimised Huffman tables (a DHT "
                         "per file); dri = restart interval of one MCU row; pcg64 = SURVEY 8(d)'s generator to the letter "
                         "(numpy PCG64(1000 + i) content made on the host CPUs, files written by libjpeg-turbo; ~0.7 s "
                         "per image and core, so use it with a smaller --batch).  Secondary, labelled lines.")
    return ap.parse_args()


# ------------------------------------------------------------------------------ corpus

def make_corpus_cv2(lib, device, n, seed0, variant):
    """Same pixel content as make_corpus, files written by OpenCV's libjpeg-turbo build (cv2.imencode) in a thread
    pool: q90, 4:2:0, standard tables ("cv2"), IMWRITE_JPEG_OPTIMIZE ("optimized") or a restart interval of one MCU
    row = 120 MCUs ("dri")."""
    import cv2
    import torch
    from concurrent.futures import ThreadPoolExecutor
    from lilliput_b200.corpus import synth_frames_gpu
    flags = [cv2.IMWRITE_JPEG_QUALITY, Q_IN]
    if variant == "optimized":
        flags += [cv2.IMWRITE_JPEG_OPTIMIZE, 1]
    if variant == "dri":
        flags += [cv2.IMWRITE_JPEG_RST_INTERVAL, SRC_W // 16]
    dev = torch.device("cuda", device)
    blobs = []

    def enc(im):
        ok, b = cv2.imencode(".jpg", im, flags)
        assert ok
        return np.asarray(b).reshape(-1)
    with ThreadPoolExecutor(max_workers=max(2, usable_cpus())) as ex:
        for g0 in range(0, n, 64):
            cnt = min(64, n - g0)
            frames = synth_frames_gpu(dev, cnt, SRC_W, SRC_H, 3, seed0 + g0)
            blobs += list(ex.map(enc, [frames[i] for i in range(cnt)]))
    lens = [int(b.size) for b in blobs]
    total = int(sum(lens))
    lib.l.lp_host_alloc_pinned.restype = C.c_void_p
    lib.l.lp_host_alloc_pinned.argtypes = [C.c_size_t]
    base = lib.l.lp_host_alloc_pinned(total + 64)
    arena = np.ctypeslib.as_array(C.cast(base, C.POINTER(C.c_uint8)), shape=(total + 64,))
    offs, o = [], 0
    for b in blobs:
        arena[o:o + b.size] = b
        offs.append(o)
        o += b.size
    return base, arena, offs, lens


def make_corpus_pcg64(lib, n, first):
    from lilliput_b200.corpus import corpus_config2_pcg64
    blobs = corpus_config2_pcg64(n, first=first, w=SRC_W, h=SRC_H, seed0=1000, quality=Q_IN, workers=usable_cpus())
    lens = [int(b.size) for b in blobs]
    total = int(sum(lens))
    lib.l.lp_host_alloc_pinned.restype = C.c_void_p
    lib.l.lp_host_alloc_pinned.argtypes = [C.c_size_t]
    base = lib.l.lp_host_alloc_pinned(total + 64)
    arena = np.ctypeslib.as_array(C.cast(base, C.POINTER(C.c_uint8)), shape=(total + 64,))
    offs, o = [], 0
    for b in blobs:
        arena[o:o + b.size] = b
        offs.append(o)
        o += b.size
    return base, arena, offs, lens


def make_corpus(lib, device, n, seed0, variant="default"):
    if variant == "pcg64":
        return make_corpus_pcg64(lib, n, seed0 - 1000)  # seed0 = 1000 + first image index of this rank's shard
    if variant != "default":
        return make_corpus_cv2(lib, device, n, seed0, variant)
    return _make_corpus_default(lib, device, n, seed0)


def _make_corpus_default(lib, device, n, seed0):
    """n distinct synthetic 1080p baseline JPEGs (q90, 4:2:0, no restart markers), made on the GPU:
    pixel content from torch (seeded per image), encoded by the library's own encoder, whose
    output is byte-identical to the reference encoder (tests/test_gpu_parity.py).  Returns a pinned
    host arena plus (offset, length) per image."""
    import torch
    l = lib.l
    l.lp_jpeg_encode_dev.restype = C.c_int
    l.lp_jpeg_encode_dev.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    dev = torch.device("cuda", device)
    cap = 1 << 20
    group = 64
    yy = torch.arange(SRC_H, device=dev, dtype=torch.float32).view(1, SRC_H, 1, 1)
    xx = torch.arange(SRC_W, device=dev, dtype=torch.float32).view(1, 1, SRC_W, 1)
    blobs, lens = [], []
    out = torch.empty((group, cap), dtype=torch.uint8, device=dev)
    out_len = torch.zeros(group, dtype=torch.int32, device=dev)
    for g0 in range(0, n, group):
        cnt = min(group, n - g0)
        gen = torch.Generator(device=dev)
        gen.manual_seed(seed0 + g0)
        img = torch.full((cnt, SRC_H, SRC_W, 3), 128.0, device=dev)
        for _ in range(6):  # low-frequency field: sum of random 2-D cosines per channel
            amp = (40 + 50 * torch.rand((cnt, 1, 1, 3), generator=gen, device=dev)) / 6.0
            f = (0.5 + 5.5 * torch.rand((cnt, 1, 1, 3, 2), generator=gen, device=dev)) * (2 * np.pi / SRC_W)
            ph = 2 * np.pi * torch.rand((cnt, 1, 1, 3), generator=gen, device=dev)
            img += amp * torch.cos(f[..., 0] * xx + f[..., 1] * yy + ph)
        for _ in range(8):  # filled rectangles: hard edges
            cx = torch.randint(0, SRC_W, (cnt,), generator=gen, device=dev)
            cy = torch.randint(0, SRC_H, (cnt,), generator=gen, device=dev)
            rw = torch.randint(48, 320, (cnt,), generator=gen, device=dev)
            rh = torch.randint(27, 180, (cnt,), generator=gen, device=dev)
            col = 255 * torch.rand((cnt, 1, 1, 3), generator=gen, device=dev)
            m = ((xx.squeeze(-1) - cx.view(-1, 1, 1)).abs() < rw.view(-1, 1, 1)) & \
                ((yy.squeeze(-1) - cy.view(-1, 1, 1)).abs() < rh.view(-1, 1, 1))
            img = torch.where(m.unsqueeze(-1), col, img)
        img += 6.0 * torch.randn(img.shape, generator=gen, device=dev)
        frames = img.round_().clamp_(0, 255).to(torch.uint8).contiguous()
        del img
        rc = l.lp_jpeg_encode_dev(frames.data_ptr(), SRC_H * SRC_W * 3, SRC_W * 3, SRC_W, SRC_H, 3, Q_IN,
                                  cnt, out.data_ptr(), cap, out_len.data_ptr(),
                                  torch.cuda.current_stream(dev).cuda_stream)
        if rc:
            raise RuntimeError(f"lp_jpeg_encode_dev failed: {rc}")
        torch.cuda.synchronize(dev)
        ln = out_len[:cnt].cpu().numpy()
        assert (ln > 0).all()
        host = out[:cnt].cpu().numpy()
        for i in range(cnt):
            blobs.append(host[i, : ln[i]].copy())
            lens.append(int(ln[i]))
    total = int(sum(lens))
    l.lp_host_alloc_pinned.restype = C.c_void_p
    l.lp_host_alloc_pinned.argtypes = [C.c_size_t]
    base = l.lp_host_alloc_pinned(total + 64)
    if not base:
        raise RuntimeError("pinned allocation failed")
    arena = np.ctypeslib.as_array(C.cast(base, C.POINTER(C.c_uint8)), shape=(total + 64,))
    offs, o = [], 0
    for b in blobs:
        arena[o:o + b.size] = b
        offs.append(o)
        o += b.size
    return base, arena, offs, lens


# ------------------------------------------------------------------------------ clocks

class ClockSampler:
    """SM clock + throttle reasons DURING the timed regions.  A timed region is a few hundred ms, shorter
    than nvidia-smi takes to start, so the sampler is started BEFORE the warm-up and every sample carries
    a host time stamp; stop() keeps the samples that fall inside the windows opened with begin()/end().
    Primary source: NVML polled every 20 ms from a thread; fallback: an `nvidia-smi -lms 100` child."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"),
            (0x4, "sw_power_cap"))

    def __init__(self, index, uuid=None):
        self.index = index
        self.uuid = uuid[4:] if uuid and str(uuid).startswith("GPU-") else uuid   # bare UUID; "GPU-" is added where used
        self.nvml_rows = []       # (t, sm_mhz, reasons bit mask)
        self.smi_rows = []        # (t, sm_mhz, max_mhz, [reason names])
        self.windows = []
        self.max_mhz = None
        self.proc = None
        self.threads = []
        self.halt = threading.Event()
        self.source = None

    # -- sources
    def _nvml_open(self):
        import pynvml
        pynvml.nvmlInit()
        h = None
        if self.uuid:                                        # CUDA_VISIBLE_DEVICES may renumber: go by UUID
            try:
                h = pynvml.nvmlDeviceGetHandleByUUID(f"GPU-{self.uuid}")
            except Exception:
                h = None
        if h is None:
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        reasons(h)                                           # probe once: raises if unsupported
        return pynvml, h, reasons

    def _nvml_poll(self, pynvml, h, reasons):
        while not self.halt.is_set():
            try:
                self.nvml_rows.append((time.perf_counter(), int(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)),
                                       int(reasons(h))))
            except Exception:
                pass
            self.halt.wait(0.02)

    def _smi_read(self):
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            if r and r[0].isdigit():
                names = [n for k, (_, n) in enumerate(self.BITS) if len(r) > 2 + k and r[2 + k] == "Active"]
                self.smi_rows.append((time.perf_counter(), int(r[0]),
                                      int(r[1]) if len(r) > 1 and r[1].isdigit() else None, names))

    def start(self):
        try:
            pynvml, h, reasons = self._nvml_open()
            t = threading.Thread(target=self._nvml_poll, args=(pynvml, h, reasons), daemon=True)
            t.start()
            self.threads.append(t)
        except Exception:
            pass
        try:
            # the plain index is what nvidia-smi counts by unless CUDA_VISIBLE_DEVICES renumbers the devices
            by_uuid = self.uuid and os.environ.get("CUDA_VISIBLE_DEVICES")
            sel = f"--id={'GPU-' + str(self.uuid) if by_uuid else self.index}"
            self.proc = subprocess.Popen(
                ["nvidia-smi", sel, f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            t = threading.Thread(target=self._smi_read, daemon=True)
            t.start()
            self.threads.append(t)
        except Exception:
            self.proc = None

    def begin(self):
        self.windows.append([time.perf_counter(), None])

    def end(self):
        self.windows[-1][1] = time.perf_counter()

    def _inside(self, t):
        return any(a <= t <= (b if b is not None else t) for a, b in self.windows)

    def stop(self):
        self.halt.set()
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        for t in self.threads:
            t.join(timeout=2)
        return self.summary(self.nvml_rows, self.smi_rows)

    def summary(self, nvml_rows, smi_rows):
        window = "timed regions"
        rows = [r for r in nvml_rows if self._inside(r[0])]
        if rows:
            source, sm, mx = "nvml 20 ms", [r[1] for r in rows], self.max_mhz
            mask = 0
            for r in rows:
                mask |= r[2]
            reasons = [n for b, n in self.BITS if mask & b]
        else:
            rows = [r for r in smi_rows if self._inside(r[0])]
            if not rows and smi_rows and self.windows:
                # nothing landed inside (region shorter than the sampling period): the warm-up runs the same
                # kernels back to back, so its samples are the next best thing -- and the JSON says so
                t_end = max(b or a for a, b in self.windows)
                rows = [r for r in smi_rows if r[0] <= t_end]
                window = "warm-up + timed regions"
            source, sm = "nvidia-smi -lms 100", [r[1] for r in rows]
            mxs = [r[2] for r in rows if r[2]]
            mx = max(mxs) if mxs else None
            reasons = [n for _, n in self.BITS if any(n in r[3] for r in rows)]
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"], "samples": 0}
        return {"sm_mhz": int(np.median(sm)), "sm_max_mhz": mx, "reasons": reasons, "samples": len(sm),
                "source": source, "window": window}


# ------------------------------------------------------------------------------ CPU reference arm

def ref_opts():
    from lilliput_b200 import abi
    return abi.ImageOptions(FileType=".jpeg", Width=DST, Height=DST, ResizeMethod=abi.ImageOpsFit,
                            NormalizeOrientation=True, EncodeOptions={abi.JpegQuality: Q_OUT})


def cpu_reference_run(base, offs, lens, total_images, threads):
    """`total_images` Transforms through the REFERENCE's own CPU implementation (oracle/_ref).
    Returns elapsed seconds.  This is the one place bench.py executes anything under oracle/."""
    from lilliput_b200 import abi
    ref = abi.load_reference()
    l = ref.l
    l.ref_transform_many.restype = C.c_double
    l.ref_transform_many.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                     C.c_size_t, C.c_long, C.POINTER(C.c_int)]
    n = len(offs)
    ptrs = (C.c_void_p * n)(*[base + o for o in offs])
    ln = (C.c_size_t * n)(*lens)
    opt = ref_opts()._c()
    err = C.c_int(0)
    el = l.ref_transform_many(ptrs, ln, n, C.byref(opt), 2048, threads, 1 << 20, total_images, C.byref(err))
    if el < 0:
        raise RuntimeError(f"reference transform failed: {err.value}")
    return el


def usable_cpus():
    """Host threads this container may actually use: the cgroup CPU quota if one is set."""
    n = os.cpu_count() or 1
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(np.ceil(int(quota) / int(period)))))
    except Exception:
        pass
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    return n


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"



# ------------------------------------------------------------------------------ configs 3, 4, 5 (lp_xbatch_*)

XCFG = {
    3: dict(metric="images_per_sec_4k_rgba_png_to_512x512_webp_q85", batch=2048, distinct=16, unit="images/s",
            workload="config3: batch 2048 synthetic 3840x2160 RGBA PNG (zlib 6) -> Fit 512x512 WebP q85 lossy + alpha, per GPU",
            opt=dict(FileType=".webp", Width=512, Height=512, q_key="WebpQuality", q=85), out_cap=1 << 20,
            resize_bytes=2160 * 2160 * 4 + 512 * 512 * 4),
    4: dict(metric="animations_per_sec_128f_720p_gif_to_256x256_animated_webp_q85", batch=256, distinct=4, unit="animations/s",
            workload="config4: 256 synthetic 128-frame 1280x720 GIF -> Fit 256x256 animated WebP q85, per GPU",
            opt=dict(FileType=".webp", Width=256, Height=256, q_key="WebpQuality", q=85), out_cap=8 << 20,
            resize_bytes=128 * (720 * 720 * 4 + 256 * 256 * 4)),
    5: dict(metric="images_per_sec_mixed_jpeg_png_webp_480p_4k_to_256x256_jpeg_q85", batch=8192, distinct=2, unit="images/s",
            workload="config5: mixed batch (60% JPEG, 25% PNG, 15% WebP; 854x480 .. 3840x2160) -> Fit 256x256 JPEG q85, "
                     "8192 images per GPU (65536 over 8 GPUs, sharded by image index)",
            opt=dict(FileType=".jpeg", Width=256, Height=256, q_key="JpegQuality", q=85), out_cap=1 << 17,
            resize_bytes=None),
}


def x_options(cfg):
    from lilliput_b200 import abi
    o = cfg["opt"]
    return abi.ImageOptions(FileType=o["FileType"], Width=o["Width"], Height=o["Height"], ResizeMethod=abi.ImageOpsFit,
                            NormalizeOrientation=True, EncodeOptions={getattr(abi, o["q_key"]): o["q"]},
                            EncodeTimeout_ns=600 * 10**9)


def x_corpus(config, dev, n, distinct, rank):
    """(list of distinct files, index of the file behind each of the n batch items)."""
    from lilliput_b200 import corpus
    if config == 3:
        files = corpus.corpus_config3(dev, distinct, seed0=2000 + 1000 * rank)
        return files, [i % len(files) for i in range(n)]
    if config == 4:
        files = corpus.corpus_config4(dev, distinct, seed0=3000 + 1000 * rank)
        return files, [i % len(files) for i in range(n)]
    cells = corpus.corpus_config5(dev, variants=distinct, seed0=5000 + 1000 * rank)
    files, where = [], {}
    for key, lst in cells.items():
        where[key] = [len(files) + k for k in range(len(lst))]
        files += lst
    idx = []
    for i in range(n):
        g = rank * n + i  # global image index: shard by contiguous block of the 65536
        cell = corpus.c5_kind(g)
        idx.append(where[cell][(g // 100) % len(where[cell])])
    return files, idx


def x_reference_run(cfg, base, offs, lens, total, threads, max_size=8192):
    from lilliput_b200 import abi
    ref = abi.load_reference()
    l = ref.l
    l.ref_transform_many.restype = C.c_double
    l.ref_transform_many.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                     C.c_size_t, C.c_long, C.POINTER(C.c_int)]
    n = len(offs)
    ptrs = (C.c_void_p * n)(*[base + o for o in offs])
    ln = (C.c_size_t * n)(*lens)
    opt = x_options(cfg)._c()
    err = C.c_int(0)
    el = l.ref_transform_many(ptrs, ln, n, C.byref(opt), max_size, threads, cfg["out_cap"], total, C.byref(err))
    if el < 0:
        raise RuntimeError(f"reference transform failed: {err.value}")
    return el


def main_x(args):
    """BASELINE configs 3 / 4 / 5 through lp_xbatch_transform (host buffers in, host buffers out)."""
    cfg = XCFG[args.config]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference" and rank != 0:
        return 0
    import torch
    from lilliput_b200 import abi
    from lilliput_b200.shard import max_over_ranks
    dist = None
    if world > 1 and args.impl != "reference":
        import torch.distributed as dist_mod
        dist = dist_mod
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = abi.load_cuda()
    lib.l.lp_set_device.argtypes = [C.c_int]
    lib.l.lp_set_device(local_rank)
    cores = usable_cpus()
    threads = min(os.cpu_count() or 1, 2 * cores)
    n = args.batch if args.batch != 4096 else cfg["batch"]
    distinct = args.distinct or cfg["distinct"]
    t_setup = time.time()
    files, idx = x_corpus(args.config, dev, n, distinct, rank)
    torch.cuda.empty_cache()
    lens_d = [int(f.size) for f in files]
    total = int(sum(lens_d))
    lib.l.lp_host_alloc_pinned.restype = C.c_void_p
    lib.l.lp_host_alloc_pinned.argtypes = [C.c_size_t]
    base = lib.l.lp_host_alloc_pinned(total + 64)
    arena = np.ctypeslib.as_array(C.cast(base, C.POINTER(C.c_uint8)), shape=(total + 64,))
    offs_d, o = [], 0
    for f in files:
        arena[o:o + f.size] = f
        offs_d.append(o)
        o += f.size
    setup_s = time.time() - t_setup
    units = n  # images (configs 3, 5) or animations (config 4)

    if args.impl == "reference":
        per_step = max(threads * 6, 96) if args.config != 4 else max(threads * 2, 16)
        for _ in range(min(args.warmup, 1)):
            x_reference_run(cfg, base, offs_d, lens_d, threads, threads)
        t = 0.0
        for _ in range(args.steps):
            t += x_reference_run(cfg, base, offs_d, lens_d, per_step, threads)
        v = per_step * args.steps / t
        line = {"impl": "reference", "metric": cfg["metric"], "value": round(v, 3), "unit": cfg["unit"], "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1000 * t / args.steps, 3),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": cfg["workload"], "units_per_step": per_step, "unique_files": len(files)},
                "cpu_baseline": {"value": round(v, 3), "unit": cfg["unit"], "cores": cores, "kind": "reference",
                                 "sample": f"{per_step} Transforms per step over {len(files)} distinct inputs, {threads} threads "
                                           f"on {cores} usable CPUs, workers + framebuffers warmed before the clock, "
                                           f"cv::setNumThreads(1), {cpu_model()}"},
                "e2e": {"value": round(v, 3), "unit": cfg["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    xb = abi.XBatch(lib, local_rank)
    ptrs = (C.c_void_p * n)(*[base + offs_d[k] for k in idx])
    ln = (C.c_size_t * n)(*[lens_d[k] for k in idx])
    out_cap = cfg["out_cap"]
    out_base = lib.l.lp_host_alloc_pinned(n * out_cap)
    out_ptrs = (C.c_void_p * n)(*[out_base + i * out_cap for i in range(n)])
    out_lens = (C.c_size_t * n)()
    status = (C.c_int * n)()
    copt = x_options(cfg)._c()

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    try:
        gpu_uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        gpu_uuid = None
    sampler = ClockSampler(local_rank, gpu_uuid)
    sampler.start()
    for _ in range(args.warmup):
        rc = xb.transform_into(ptrs, ln, n, copt, out_ptrs, out_cap, out_lens, status)
        assert rc == 0
    bad = [(i, status[i]) for i in range(n) if status[i] != 0]
    assert not bad, f"per-item failures: {bad[:8]}"
    barrier()
    sampler.begin()
    agg = {}
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rc = xb.transform_into(ptrs, ln, n, copt, out_ptrs, out_cap, out_lens, status)
        assert rc == 0
        for k, v in xb.stats().items():
            agg[k] = agg.get(k, 0) + v
    barrier()
    e2e_s = time.perf_counter() - t0
    sampler.end()
    clocks = sampler.stop()
    assert all(status[i] == 0 for i in range(n))
    kern_s = agg["ms_busy_max_lane"] / 1000.0
    kern_max, e2e_max = max_over_ranks([kern_s, e2e_s], dist, device="cuda")
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        st = args.steps
        in_bytes = int(sum(lens_d[k] for k in idx))
        line = {
            "metric": cfg["metric"], "value": round(world * units * st / kern_max, 2), "unit": cfg["unit"], "n_gpus": world,
            "steps": st, "warmup": args.warmup, "ms_per_step": round(1000 * kern_max / st, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": cfg["workload"], "units_per_gpu_per_step": units, "unique_files": len(files),
                       "sharding": "by image index, no collective",
                       "timing": "value: CUDA-event time of the grid stages (decode + resize + encode; inputs already in HBM, "
                                 "H2D / header parsing excluded) of the busier of the two concurrently running lanes, "
                                 "max over ranks; e2e: wall clock around lp_xbatch_transform with pinned host buffers "
                                 "in and out; stage_ms_per_step sums both lanes",
                       "l2": "inputs %.2f GB per step exceed the 126 MB L2; no flush needed" % (in_bytes / 1e9),
                       "stage_ms_per_step": {k: round(agg[k] / st, 3) for k in ("ms_parse", "ms_grid", "ms_fallback", "ms_total",
                                                                                 "ms_decode", "ms_resize", "ms_encode", "ms_busy_max_lane")},
                       "grid_items": int(agg["grid_items"] / st), "fallback_items": int(agg["fallback_items"] / st),
                       "groups": int(agg["groups"] / st), "setup_s": round(setup_s, 1)},
            "e2e": {"value": round(world * units * st / e2e_max, 2), "unit": cfg["unit"],
                    "h2d_bytes_per_step": int(agg["h2d_bytes"] / st), "d2h_bytes_per_step": int(agg["d2h_bytes"] / st),
                    "ms_per_step": round(1000 * e2e_max / st, 3)},
            "gpu_launches": int(agg["launches"]),
            "clocks": clocks,
        }
        if cfg["resize_bytes"] and agg["ms_resize"] > 0:
            ach = units * st * cfg["resize_bytes"] / (agg["ms_resize"] * 1e-3) / 1e9
            line["roofline"] = {"kernel": "resize_area_kernel (all launches of the step)", "bound": "hbm", "achieved": round(ach, 1),
                                "peak": hbm_peak, "unit": "GB/s", "frac": round(ach / hbm_peak, 4), "traffic": None,
                                "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650",
                                "bytes_per_unit": cfg["resize_bytes"]}
        if world == 1 and not args.no_cpu_baseline and os.path.exists(abi.REF_LIB):
            # bounded sample: one call of at least --cpu-seconds of CPU work (the pool and its framebuffers are warmed
            # before the clock starts)
            total_n = max(4 * threads, 32)
            el = x_reference_run(cfg, base, offs_d, lens_d, total_n, threads)
            while el < args.cpu_seconds:   # grow the call until ONE call lasts --cpu-seconds (short calls under-report:
                total_n = int(total_n * min(8.0, max(1.5, 1.25 * args.cpu_seconds / max(el, 1e-3))))  # few images per thread)
                el = x_reference_run(cfg, base, offs_d, lens_d, total_n, threads)
            line["cpu_baseline"] = {"value": round(total_n / el, 3), "unit": cfg["unit"], "cores": cores, "kind": "reference",
                                    "sample": f"{total_n} Transforms over the {len(files)} distinct inputs in {el:.1f} s, {threads} "
                                              f"threads on {cores} usable CPUs, workers warmed before the clock, "
                                              f"cv::setNumThreads(1), {cpu_model()}"}
        print(json.dumps(line))
    xb.close()
    if dist:
        dist.destroy_process_group()
    return 0


# ------------------------------------------------------------------------------ main

def resize_traffic(images_per_launch):
    """dram__bytes_read.sum + dram__bytes_write.sum of the resize kernel per launch, from the committed
    `ncu --set full` capture (profiles/r02_resize_traffic.json: measured per image on a 1332-image launch,
    1.014x the algorithmic bytes), scaled to this run's images per launch.  None if the file is missing."""
    path = os.path.join(ROOT, "profiles", "r02_resize_traffic.json")
    if not os.path.exists(path):
        path = os.path.join(ROOT, "profiles", "r01_resize_traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        return int(json.load(f)["traffic_bytes_per_image"] * images_per_launch)


def main():
    args = parse_args()
    if args.config != 2:
        return main_x(args)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference" and rank != 0:
        return 0
    import torch
    from lilliput_b200 import abi

    dist = None
    if world > 1 and args.impl != "reference":
        import torch.distributed as dist_mod
        dist = dist_mod
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    lib = abi.load_cuda()  # no fallback: raises if the .so or the GPU is missing
    lib.l.lp_set_device.argtypes = [C.c_int]
    lib.l.lp_set_device(local_rank)
    cores = usable_cpus()
    # the vendored libs gain a little from 2 threads per granted CPU; more only adds contention
    threads = min(os.cpu_count() or 1, 2 * cores)

    if args.impl == "reference":
        # the same corpus as the GPU arm (its 4096 distinct inputs), workers warmed before each clock
        sample_n = args.batch
        from lilliput_b200.shard import corpus_seed
        base, arena, offs, lens = make_corpus(lib, local_rank, sample_n, corpus_seed(1000, 0, sample_n), args.variant)
        # a step = whole passes over the corpus, as many as it takes to last ~4 s: a single 4096-image pass is under a
        # second on a many-core host, and so short a call under-reports the CPU path (pool start-up, ragged tail)
        pass_s = cpu_reference_run(base, offs, lens, sample_n, threads)
        per_step = sample_n * int(min(64, max(1, np.ceil(4.0 / max(pass_s, 1e-3)))))
        for _ in range(max(0, args.warmup - 1)):
            cpu_reference_run(base, offs, lens, per_step, threads)
        t = 0.0
        for _ in range(args.steps):
            t += cpu_reference_run(base, offs, lens, per_step, threads)
        v = per_step * args.steps / t
        line = {
            "impl": "reference", "metric": "images_per_sec_1080p_jpeg_to_256x256_jpeg_q85", "value": round(v, 2),
            "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(1000 * t / args.steps, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "config2: synthetic 1920x1080 baseline JPEG q90 -> Fit 256x256 JPEG q85",
                       "images_per_step": per_step, "unique_images": sample_n, "images_per_gpu_per_step": per_step},
            "cpu_baseline": {"value": round(v, 2), "unit": "images/s", "cores": cores, "kind": "reference",
                             "sample": f"{per_step} Transforms per step ({per_step // sample_n} passes over the {sample_n} distinct inputs of the GPU arm's corpus), "
                                       f"workers + framebuffers warmed before the clock, "
                                       f"{threads} threads on {cores} usable CPUs (cgroup quota; host has "
                                       f"{os.cpu_count()} hw threads), cv::setNumThreads(1), {cpu_model()}"},
            "e2e": {"value": round(v, 2), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    n = args.batch
    t_setup = time.time()
    from lilliput_b200.shard import corpus_seed, max_over_ranks
    base, arena, offs, lens = make_corpus(lib, local_rank, n, corpus_seed(1000, rank, n), args.variant)
    in_bytes = int(sum(lens))
    out_cap = 65536
    b = abi.Batch(lib, local_rank, n, SRC_W, SRC_H, DST, DST, Q_OUT, max_in_bytes=in_bytes + (1 << 20),
                  out_cap=out_cap, chunk=args.chunk)
    ptrs = (C.c_void_p * n)(*[base + o for o in offs])
    ln = (C.c_size_t * n)(*lens)
    lib.l.lp_host_alloc_pinned.restype = C.c_void_p
    out_base = lib.l.lp_host_alloc_pinned(n * out_cap)
    out_ptrs = (C.c_void_p * n)(*[out_base + i * out_cap for i in range(n)])
    out_lens = (C.c_size_t * n)()
    status = (C.c_int * n)()
    setup_s = time.time() - t_setup

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident: inputs staged in HBM once, kernels only in the timed region
    st = b.stage([(base + o, l_) for o, l_ in zip(offs, lens)])
    assert all(s == 0 for s in st), "corpus rejected by the header parser"
    try:
        gpu_uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        gpu_uuid = None
    sampler = ClockSampler(local_rank, gpu_uuid)
    sampler.start()                      # before the warm-up: nvidia-smi needs longer to start than a step takes
    for _ in range(args.warmup):
        b.run()
    barrier()
    phase = (C.c_ulonglong * 8)()
    has_phase = hasattr(lib.l, "lp_huff_phase_clocks")
    if has_phase:
        lib.l.lp_huff_phase_clocks.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
        lib.l.lp_huff_phase_clocks(phase, 1)   # clear the entropy kernel's per-phase cycle counters
    sampler.begin()
    stage_sum = {}
    dev_ms = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ms = b.run()
        dev_ms += ms["total"]
        for k, v in ms.items():
            stage_sum[k] = stage_sum.get(k, 0.0) + v
    barrier()
    wall_s = time.perf_counter() - t0
    sampler.end()
    huff_phase = None
    if has_phase and lib.l.lp_huff_phase_clocks(phase, 0) == 0 and phase[5]:
        tot = float(sum(phase[k] for k in range(5))) or 1.0
        huff_phase = {k: round(phase[i] / tot, 4) for i, k in enumerate(("tables", "guess", "sync", "write", "dc"))}
        huff_phase["cycles_per_image"] = int(tot / phase[5])
    launches = b.last_launches() * args.steps
    outs, fst = b.fetch(n)
    assert all(s == 0 for s in fst), "device pipeline reported per-image failures"
    out_bytes = int(sum(len(o) for o in outs))

    # ---------------- end to end: host buffers in, host buffers out, every step
    for _ in range(max(1, args.warmup // 2)):
        rc = b.transform_into(ptrs, ln, n, out_ptrs, out_lens, status)
        assert rc == 0
    barrier()
    sampler.begin()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rc = b.transform_into(ptrs, ln, n, out_ptrs, out_lens, status)
        assert rc == 0
    barrier()
    e2e_s = time.perf_counter() - t0
    sampler.end()
    clocks = sampler.stop()
    assert all(status[i] == 0 for i in range(n))

    dev_s, e2e_max, wall_max = max_over_ranks([dev_ms / 1000.0, e2e_s, wall_s], dist, device="cuda")

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650"
        lib.l.lp_batch_chunk.argtypes = [C.c_void_p]
        chunk = lib.l.lp_batch_chunk(b.h)
        nchunks = max(1, -(-n // chunk))
        mean_r, max_r = C.c_double(0), C.c_int(0)
        lib.l.lp_batch_sync_rounds.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)]
        lib.l.lp_batch_sync_rounds(b.h, C.byref(mean_r), C.byref(max_r))
        resize_ms_per_launch = stage_sum["resize"] / (args.steps * nchunks)
        per_launch_images = n / nchunks
        achieved = per_launch_images * RESIZE_BYTES_PER_IMAGE / (resize_ms_per_launch * 1e-3) / 1e9
        lib.l.lp_batch_d2h_overhead_per_image.restype = C.c_size_t
        d2h_over = int(lib.l.lp_batch_d2h_overhead_per_image())
        value = world * n * args.steps / dev_s
        e2e_v = world * n * args.steps / e2e_max
        line = {
            "metric": "images_per_sec_1080p_jpeg_to_256x256_jpeg_q85", "value": round(value, 1),
            "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(1000 * dev_s / args.steps, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "config2: batch 4096 synthetic 1920x1080 baseline JPEG q90 (4:2:0, no DRI) -> "
                                   "Fit 256x256 JPEG q85, per GPU" + ("" if args.variant == "default" else
                                                                      f" [SECONDARY corpus variant: {args.variant}]"),
                       "corpus": {"default": "torch content, this library's encoder (byte-identical to the reference's), standard tables",
                                  "cv2": "torch content, files written by cv2 (libjpeg-turbo), standard tables",
                                  "optimized": "torch content, cv2 with per-image optimised Huffman tables (one DHT set per file)",
                                  "dri": "torch content, cv2 with a restart interval of one MCU row (120 MCUs)",
                                  "pcg64": "SURVEY 8(d) generator: numpy PCG64(1000 + i) content on the host, files written by "
                                           "libjpeg-turbo (cv2), standard tables"}[args.variant],
                       "images_per_gpu_per_step": n, "sharding": "by image index, no collective",
                       "l2": "inputs (%.2f GB compressed, 25 GB decoded per step) exceed the 126 MB L2; no flush needed"
                             % (in_bytes / 1e9),
                       "timing": "CUDA events on the library stream (first launch -> last kernel), max over ranks",
                       "wall_ms_per_step": round(1000 * wall_max / args.steps, 3),
                       "stage_ms_per_step": {k: round(v / args.steps, 3) for k, v in stage_sum.items()},
                       "chunk_images": chunk,
                       "huffman_sync_rounds": {"mean": round(mean_r.value, 2), "max": max_r.value},
                       "huffman_phase_share": huff_phase,
                       "setup_s": round(setup_s, 1)},
            "e2e": {"value": round(e2e_v, 1), "unit": "images/s", "h2d_bytes_per_step": in_bytes,
                    "d2h_bytes_per_step": out_bytes + n * d2h_over, "encoded_bytes_per_step": out_bytes,
                    "d2h_note": "encoded bytes are compacted on the device and written straight into the pinned, "
                                "device-mapped output buffer; per image also 4 B length, 8 B offset and the item mirror (status)",
                    "ms_per_step": round(1000 * e2e_max / args.steps, 3)},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"kernel": "resize_area_kernel<3,6>", "bound": "hbm", "achieved": round(achieved, 1),
                         "peak": hbm_peak, "unit": "GB/s", "frac": round(achieved / hbm_peak, 4),
                         "peak_source": peak_src, "traffic": resize_traffic(per_launch_images),
                         "bytes_per_launch": int(per_launch_images * RESIZE_BYTES_PER_IMAGE),
                         "ms_per_launch": round(resize_ms_per_launch, 4)},
        }
        if world == 1 and not args.no_cpu_baseline and os.path.exists(abi.REF_LIB):
            sample_n = n
            total = max(32 * threads, 1024)
            el = cpu_reference_run(base, offs[:sample_n], lens[:sample_n], total, threads)
            while el < args.cpu_seconds:   # one call of at least --cpu-seconds of CPU work, workers warmed before the clock
                total = int(total * min(8.0, max(1.5, 1.25 * args.cpu_seconds / max(el, 1e-3))))
                el = cpu_reference_run(base, offs[:sample_n], lens[:sample_n], total, threads)
            line["cpu_baseline"] = {"value": round(total / el, 2), "unit": "images/s", "cores": cores,
                                    "kind": "reference",
                                    "sample": f"{total} Transforms over the first {sample_n} inputs of the same "
                                              f"corpus in {el:.1f} s, {threads} threads on {cores} usable CPUs "
                                              f"(cgroup quota; host has {os.cpu_count()} hw threads), "
                                              f"cv::setNumThreads(1), {cpu_model()}"}
        print(json.dumps(line))
    b.close()
    if dist:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
