"""Synthetic corpora for bench.py's BASELINE configs 3, 4 and 5.

Pixel content is made on the GPU with torch (low-frequency cosine field + filled rectangles + Gaussian noise, seeded
per image -- the recipe of lilliput_b200/synth.py at a speed that makes 4K corpora practical); the FILES are written
by independent encoders, none of them this library's: OpenCV 4.13's libjpeg-turbo / libpng / libwebp builds (cv2) and
Pillow's GIF writer -- the same codec families the reference links, so the streams carry what real encoders emit
(8 KiB IDAT chunks, adaptive PNG filters, libwebp's 4x4 modes / segments / probability updates, GIF sub-blocks).
bench.py replicates a bounded number of distinct files up to the batch size and says so in `config`.
"""
from __future__ import annotations

import io
from concurrent.futures import ThreadPoolExecutor

import numpy as np


def synth_frames_gpu(dev, n, w, h, channels, seed, noise=6.0, shift=None):
    """n frames [n, h, w, channels] uint8 on the host.  channels 4: alpha = radial falloff x soft random mask.
    shift: per-frame x translation of the field in pixels (animations), else independent images."""
    import torch
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed))
    yy = torch.arange(h, device=dev, dtype=torch.float32).view(1, h, 1, 1)
    xx = torch.arange(w, device=dev, dtype=torch.float32).view(1, 1, w, 1)
    cnt = 1 if shift is not None else n
    img = torch.full((cnt, h, w, 3), 128.0, device=dev)
    dx = torch.zeros((n, 1, 1, 1), device=dev) if shift is None else torch.tensor(shift, device=dev, dtype=torch.float32).view(n, 1, 1, 1)
    if shift is not None:
        img = torch.full((n, h, w, 3), 128.0, device=dev)
    for _ in range(6):
        amp = (40 + 50 * torch.rand((cnt, 1, 1, 3), generator=gen, device=dev)) / 6.0
        f = (0.5 + 5.5 * torch.rand((cnt, 1, 1, 3, 2), generator=gen, device=dev)) * (2 * np.pi / w)
        ph = 2 * np.pi * torch.rand((cnt, 1, 1, 3), generator=gen, device=dev)
        img += amp * torch.cos(f[..., 0] * (xx + dx) + f[..., 1] * yy + ph)
    for k in range(8):
        cx = torch.randint(0, w, (cnt,), generator=gen, device=dev).view(-1, 1, 1).float()
        cy = torch.randint(0, h, (cnt,), generator=gen, device=dev).view(-1, 1, 1).float()
        rw = torch.randint(max(2, w // 40), max(3, w // 6), (cnt,), generator=gen, device=dev).view(-1, 1, 1)
        rh = torch.randint(max(2, h // 40), max(3, h // 6), (cnt,), generator=gen, device=dev).view(-1, 1, 1)
        col = 255 * torch.rand((cnt, 1, 1, 3), generator=gen, device=dev)
        if shift is not None and k < 3:  # moving sprites
            cx = cx + dx.view(n, 1, 1) * (1.5 + k)
            cx = torch.remainder(cx, w)
        m = ((xx.squeeze(-1) - cx).abs() < rw) & ((yy.squeeze(-1) - cy).abs() < rh)
        img = torch.where(m.unsqueeze(-1), col, img)
    if noise > 0:
        img += noise * torch.randn(img.shape, generator=gen, device=dev)
    out = img.round_().clamp_(0, 255).to(torch.uint8)
    if channels == 4:
        r = torch.hypot((xx - w / 2) / (w / 2), (yy - h / 2) / (h / 2)).squeeze(-1)
        fx = (0.5 + 2.5 * torch.rand((out.shape[0], 1, 1), generator=gen, device=dev)) * (2 * np.pi / w)
        mask = 0.75 + 0.25 * torch.cos(fx * xx.squeeze(-1) + 0.7 * fx * yy.squeeze(-1))
        a = (255 * (1.2 - r) * mask).clamp_(0, 255).to(torch.uint8).unsqueeze(-1)
        out = torch.cat([out, a.expand(out.shape[0], h, w, 1)], dim=-1)
    return out.contiguous().cpu().numpy()


def _pool_map(fn, items, threads=8):
    with ThreadPoolExecutor(max_workers=threads) as ex:
        return list(ex.map(fn, items))


def encode_png(img, level=6):
    import cv2
    ok, b = cv2.imencode(".png", img, [cv2.IMWRITE_PNG_COMPRESSION, level])
    assert ok
    return np.asarray(b).reshape(-1)


def encode_jpeg(img, quality=90):
    import cv2
    ok, b = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, quality])
    assert ok
    return np.asarray(b).reshape(-1)


def encode_webp(img, quality=85):
    import cv2
    ok, b = cv2.imencode(".webp", img, [cv2.IMWRITE_WEBP_QUALITY, quality])
    assert ok
    return np.asarray(b).reshape(-1)


def encode_gif(frames_bgr, delay_cs=4, loop=0):
    """Animated GIF: 256-colour global palette from frame 0 (median cut), full frames, no local palettes."""
    from PIL import Image
    first = Image.fromarray(np.ascontiguousarray(frames_bgr[0][:, :, ::-1])).quantize(256, method=Image.Quantize.MEDIANCUT, dither=Image.Dither.NONE)
    ims = [first]
    for f in frames_bgr[1:]:
        ims.append(Image.fromarray(np.ascontiguousarray(f[:, :, ::-1])).quantize(palette=first, dither=Image.Dither.NONE))
    bio = io.BytesIO()
    ims[0].save(bio, "GIF", save_all=True, append_images=ims[1:], duration=delay_cs * 10, loop=loop, disposal=1, optimize=False)
    return np.frombuffer(bio.getvalue(), dtype=np.uint8)


def corpus_config3(dev, distinct, seed0=2000, w=3840, h=2160):
    """config 3: RGBA PNG, zlib level 6, libpng's own filter choice."""
    files = []
    for g0 in range(0, distinct, 4):
        cnt = min(4, distinct - g0)
        frames = synth_frames_gpu(dev, cnt, w, h, 4, seed0 + g0)
        files += _pool_map(encode_png, [frames[i] for i in range(cnt)], threads=4)
    return files


def corpus_config4(dev, distinct, seed0=3000, w=1280, h=720, nframes=128):
    """config 4: 128-frame 1280x720 GIFs, field translated 4 px / frame + moving sprites, global palette."""
    def one(k):
        frames = synth_frames_gpu(dev, nframes, w, h, 3, seed0 + k, noise=0.0, shift=[4.0 * t for t in range(nframes)])
        return encode_gif(frames)
    return [one(k) for k in range(distinct)]


C5_SIZES = [(854, 480), (1280, 720), (1920, 1080), (2560, 1440), (3840, 2160)]


def c5_kind(i):
    """format / size of image i of config 5: i mod 20 -> 0-11 JPEG, 12-14 PNG RGB, 15-16 PNG RGBA,
    17-19 WebP lossy; (i div 20) mod 5 -> size."""
    f = i % 20
    fmt = "jpeg" if f < 12 else "png3" if f < 15 else "png4" if f < 17 else "webp"
    return fmt, C5_SIZES[(i // 20) % 5]


def corpus_config5(dev, variants=2, seed0=5000):
    """config 5: `variants` distinct files per (format, size) cell; bench.py maps image i to cell c5_kind(i)."""
    cells = {}
    for si, (w, h) in enumerate(C5_SIZES):
        for fmt in ("jpeg", "png3", "png4", "webp"):
            ch = 4 if fmt == "png4" else 3
            frames = synth_frames_gpu(dev, variants, w, h, ch, seed0 + 100 * si + {"jpeg": 0, "png3": 1, "png4": 2, "webp": 3}[fmt])
            enc = {"jpeg": lambda im: encode_jpeg(im, 90), "png3": encode_png, "png4": encode_png,
                   "webp": lambda im: encode_webp(im, 85)}[fmt]
            cells[(fmt, (w, h))] = _pool_map(enc, [frames[i] for i in range(variants)], threads=variants)
    return cells


# ------------------------------------------------------------------ SURVEY 8(d) C2, to the letter (CPU generator)

def pcg64_frame(i, w=1920, h=1080, seed0=1000):
    """Image i of SURVEY.md 8(d)'s config-2 corpus: rng = numpy PCG64(seed0 + i); low-frequency field (six random 2-D
    cosines per channel, amplitudes 40..90 in total), eight filled rectangles / ellipses, N(0, 6) noise, clipped to u8."""
    rng = np.random.Generator(np.random.PCG64(seed0 + i))
    yy = np.arange(h, dtype=np.float64)
    xx = np.arange(w, dtype=np.float64)
    img = np.full((h, w, 3), 128.0, dtype=np.float32)
    for _ in range(6):
        amp = (40 + 50 * rng.random(3)) / 6.0
        f = (0.5 + 5.5 * rng.random((3, 2))) * (2 * np.pi / w)
        ph = 2 * np.pi * rng.random(3)
        for c in range(3):  # cos(fx x + fy y + ph) as two outer products: no transcendental per pixel
            ax, ay = f[c, 0] * xx + ph[c], f[c, 1] * yy
            img[:, :, c] += (amp[c] * (np.outer(np.cos(ay), np.cos(ax)) - np.outer(np.sin(ay), np.sin(ax)))).astype(np.float32)
    for k in range(8):
        cx, cy = int(rng.integers(0, w)), int(rng.integers(0, h))
        rw, rh = int(rng.integers(48, 320)), int(rng.integers(27, 180))
        col = (255 * rng.random(3)).astype(np.float32)
        x0, x1, y0, y1 = max(cx - rw, 0), min(cx + rw, w), max(cy - rh, 0), min(cy + rh, h)
        if k % 2:  # ellipse
            sub_y = (np.arange(y0, y1, dtype=np.float32).reshape(-1, 1) - cy) / rh
            sub_x = (np.arange(x0, x1, dtype=np.float32).reshape(1, -1) - cx) / rw
            m = (sub_x * sub_x + sub_y * sub_y) < 1.0
            img[y0:y1, x0:x1][m] = col
        else:
            img[y0:y1, x0:x1] = col
    img += rng.standard_normal(img.shape, dtype=np.float32) * np.float32(6.0)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def _pcg64_jpeg(args):
    import cv2
    i, w, h, seed0, q = args
    cv2.setNumThreads(1)
    ok, b = cv2.imencode(".jpg", pcg64_frame(i, w, h, seed0), [cv2.IMWRITE_JPEG_QUALITY, q])  # 4:2:0, Annex K tables, no DRI
    assert ok
    return np.asarray(b).reshape(-1).copy()


def corpus_config2_pcg64(n, first=0, w=1920, h=1080, seed0=1000, quality=90, workers=8):
    """n files of the C2 corpus, images first .. first + n - 1, written by libjpeg-turbo (cv2) in `workers` processes."""
    from concurrent.futures import ProcessPoolExecutor
    jobs = [(first + k, w, h, seed0, quality) for k in range(n)]
    with ProcessPoolExecutor(max_workers=max(1, workers)) as ex:
        return list(ex.map(_pcg64_jpeg, jobs, chunksize=4))
