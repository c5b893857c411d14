"""Seeded synthetic frames for the parity tests and bench.py (recipes of SURVEY.md §8d).
Pure numpy, deterministic for a given (seed, size); values are uint8.
"""
import numpy as np


def frame_shapes(seed, ny, nx, n_shapes=64, noise=8):
    """C2 / C5 recipe: sum of random filled rectangles and discs (intensity U[0,255]) + i.i.d.
    noise U[0,noise); grey uint8 [ny, nx]."""
    rng = np.random.default_rng(seed)
    img = np.zeros((ny, nx), np.float32)
    yy, xx = None, None
    for i in range(n_shapes):
        v = rng.uniform(0, 255)
        cx, cy = rng.integers(0, nx), rng.integers(0, ny)
        w, h = rng.integers(nx // 32 + 2, nx // 4 + 3), rng.integers(ny // 32 + 2, ny // 4 + 3)
        x0, x1 = max(cx - w // 2, 0), min(cx + w // 2 + 1, nx)
        y0, y1 = max(cy - h // 2, 0), min(cy + h // 2 + 1, ny)
        if i % 2 == 0:
            img[y0:y1, x0:x1] = v
        else:
            r = min(w, h) / 2.0
            ys = np.arange(y0, y1, dtype=np.float32)[:, None] - cy
            xs = np.arange(x0, x1, dtype=np.float32)[None, :] - cx
            m = ys * ys + xs * xs <= r * r
            sub = img[y0:y1, x0:x1]
            sub[m] = v
    img += rng.uniform(0, noise, size=(ny, nx)).astype(np.float32)
    return np.clip(img, 0, 255).astype(np.uint8)


def frame_rgb(seed, ny, nx, noise=16):
    """C3 recipe: per-channel smooth sinusoid mix + rectangles + noise U[0,noise); uint8 [ny, nx, 3]."""
    rng = np.random.default_rng(seed)
    y = np.arange(ny, dtype=np.float32)[:, None]
    x = np.arange(nx, dtype=np.float32)[None, :]
    out = np.zeros((ny, nx, 3), np.float32)
    for c in range(3):
        acc = np.full((ny, nx), 110.0, np.float32)
        for _ in range(4):
            fx, fy = rng.uniform(0.002, 0.06, 2)
            ph = rng.uniform(0, 6.28)
            acc += rng.uniform(10, 35) * np.sin(fx * x + fy * y + ph)
        out[..., c] = acc
    for _ in range(24):
        cx, cy = rng.integers(0, nx), rng.integers(0, ny)
        w, h = rng.integers(nx // 40 + 2, nx // 6 + 3), rng.integers(ny // 40 + 2, ny // 6 + 3)
        x0, x1 = max(cx - w // 2, 0), min(cx + w // 2 + 1, nx)
        y0, y1 = max(cy - h // 2, 0), min(cy + h // 2 + 1, ny)
        out[y0:y1, x0:x1, :] = rng.uniform(0, 255, 3).astype(np.float32)
    out += rng.uniform(0, noise, size=(ny, nx, 3)).astype(np.float32)
    return np.clip(out, 0, 255).astype(np.uint8)


def frame_blobs(seed, ny, nx, n_blobs=None, noise=7):
    """C4 recipe: Gaussian blobs (sigma in [3,23), amplitude +-[40,120)) on level 96 + noise,
    grey replicated to 3 channels; uint8 [ny, nx, 3].  n_blobs defaults to 4000 per 4K frame,
    scaled by area."""
    rng = np.random.default_rng(seed)
    if n_blobs is None:
        n_blobs = max(8, int(round(4000 * (ny * nx) / (3840.0 * 2160.0))))
    img = np.full((ny, nx), 96.0, np.float32)
    for _ in range(n_blobs):
        s = rng.uniform(3, 23)
        a = rng.uniform(40, 120) * (1 if rng.random() < 0.5 else -1)
        cx, cy = rng.uniform(0, nx), rng.uniform(0, ny)
        r = int(4 * s) + 1
        x0, x1 = max(int(cx) - r, 0), min(int(cx) + r + 1, nx)
        y0, y1 = max(int(cy) - r, 0), min(int(cy) + r + 1, ny)
        if x1 <= x0 or y1 <= y0:
            continue
        ys = np.arange(y0, y1, dtype=np.float32)[:, None] - cy
        xs = np.arange(x0, x1, dtype=np.float32)[None, :] - cx
        img[y0:y1, x0:x1] += a * np.exp(-(ys * ys + xs * xs) / (2 * s * s))
    img += rng.uniform(0, noise, size=(ny, nx)).astype(np.float32)
    g = np.clip(img, 0, 255).astype(np.uint8)
    return np.repeat(g[:, :, None], 3, axis=2)


def batch(fn, seed0, n, ny, nx, distinct=4):
    """n frames: `distinct` generated frames, the rest are cyclic shifts of them (cheap, still all
    different) — used by bench.py to fill large batches."""
    base = [fn(seed0 + i, ny, nx) for i in range(min(distinct, n))]
    out = []
    for i in range(n):
        b = base[i % len(base)]
        k = i // len(base)
        out.append(np.roll(b, (37 * k, 101 * k), axis=(0, 1)) if k else b)
    return np.stack(out)
