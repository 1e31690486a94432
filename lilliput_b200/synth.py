"""Deterministic synthetic image content (SURVEY.md 8(d) corpus recipe).

Image i of a corpus is a low-frequency field (sum of random 2-D cosines per channel) plus a
few filled rectangles/ellipses (edges) plus Gaussian noise, clipped to u8, generated from
``numpy.random.Generator(PCG64(seed))``.  Host-side numpy only: bench.py uses it to make the
pixel content, which is then JPEG-encoded (by the library under test or by the oracle).
"""
from __future__ import annotations

import numpy as np


def synth_image(seed: int, width: int, height: int, channels: int = 3, noise: float = 6.0,
                n_shapes: int = 8) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(seed))
    y = np.arange(height, dtype=np.float32)[:, None]
    x = np.arange(width, dtype=np.float32)[None, :]
    planes = []
    for _ in range(max(channels, 1)):
        f = np.full((height, width), 128.0, dtype=np.float32)
        for _ in range(6):
            amp = rng.uniform(40, 90) / 6.0
            fx, fy = rng.uniform(0.5, 6.0, 2) * 2 * np.pi / max(width, height)
            ph = rng.uniform(0, 2 * np.pi)
            f += amp * np.cos(fx * x + fy * y + ph)
        planes.append(f)
    img = np.stack(planes, axis=-1)
    for _ in range(n_shapes):
        cx, cy = rng.integers(0, width), rng.integers(0, height)
        rw, rh = rng.integers(max(2, width // 40), max(3, width // 6)), rng.integers(
            max(2, height // 40), max(3, height // 6))
        col = rng.uniform(0, 255, img.shape[2]).astype(np.float32)
        x0, x1 = max(0, cx - rw), min(width, cx + rw)
        y0, y1 = max(0, cy - rh), min(height, cy + rh)
        if x1 <= x0 or y1 <= y0:
            continue
        if rng.random() < 0.5:
            img[y0:y1, x0:x1] = col
        else:
            yy = (np.arange(y0, y1, dtype=np.float32)[:, None] - cy) / rh
            xx = (np.arange(x0, x1, dtype=np.float32)[None, :] - cx) / rw
            m = (xx * xx + yy * yy) <= 1.0
            img[y0:y1, x0:x1][m] = col
    if noise > 0:
        img += rng.normal(0, noise, img.shape).astype(np.float32)
    out = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    if channels == 1:
        return out[:, :, 0]
    if channels == 4:
        # alpha: radial falloff
        r = np.hypot((x - width / 2) / (width / 2), (y - height / 2) / (height / 2))
        out[:, :, 3] = np.clip(255 * (1.2 - r), 0, 255).astype(np.uint8)
    return out
