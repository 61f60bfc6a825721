"""Deterministic synthetic frames for parity tests and bench.py.

Procedurally composited high-contrast face patches on a blurred-noise background; the patch
is built to fire haarcascade_frontalface_alt and to push windows through every stage, so that
stage-exit depth maps are exercised at all depths.  Pure numpy (no cv2) so that the generator
behaves identically wherever it runs.
"""
from __future__ import annotations

import numpy as np


def _blur(img: np.ndarray, sigma: float) -> np.ndarray:
    """Separable Gaussian blur, reflect-101 borders, float64 accumulate."""
    if sigma <= 0:
        return img.astype(np.float64)
    r = max(1, int(np.ceil(3 * sigma)))
    k = np.exp(-0.5 * (np.arange(-r, r + 1) / sigma) ** 2)
    k /= k.sum()
    out = img.astype(np.float64)
    for axis in (0, 1):
        pad = [(0, 0), (0, 0)]
        pad[axis] = (r, r)
        p = np.pad(out, pad, mode="reflect")
        acc = np.zeros_like(out)
        n = out.shape[axis]
        for i, kv in enumerate(k):
            sl = [slice(None), slice(None)]
            sl[axis] = slice(i, i + n)
            acc += kv * p[tuple(sl)]
        out = acc
    return out


def _ellipse(xx, yy, cx, cy, rx, ry):
    return ((xx - cx) / rx) ** 2 + ((yy - cy) / ry) ** 2 <= 1.0


def face_patch(s: int) -> np.ndarray:
    """s x s float patch (SURVEY.md §8d recipe)."""
    yy, xx = (np.mgrid[0:s, 0:s] + 0.5) / s
    p = np.full((s, s), 90.0)
    p[_ellipse(xx, yy, .5, .52, .38, .48)] = 185
    p[yy < .18] = 70
    for ex in (.33, .67):
        p[_ellipse(xx, yy, ex, .39, .14, .08)] = 140
        p[_ellipse(xx, yy, ex, .31, .13, .025)] = 75
        p[_ellipse(xx, yy, ex, .40, .10, .045)] = 45
    p[(np.abs(xx - .5) < .045) & (yy > .36) & (yy < .62)] = 205
    for nx in (.45, .55):
        p[_ellipse(xx, yy, nx, .64, .03, .02)] = 95
    for cx in (.3, .7):
        p[_ellipse(xx, yy, cx, .58, .10, .08)] = 200
    p[_ellipse(xx, yy, .5, .77, .16, .035)] = 80
    p[yy > .9] = 130
    return _blur(p, s / 60.0)


def frame(width: int, height: int, k: int, seed: int, channels: int = 3, tint: bool = True,
          smin: float = 0.10, smax: float = 1 / 3) -> np.ndarray:
    """One BGR (channels=3) or BGRA (channels=4) frame with k faces."""
    rng = np.random.default_rng(seed)
    bg = np.clip(110 + 25 * rng.standard_normal((height, width)), 0, 255)
    g = _blur(bg, 1.5)
    for _ in range(k):
        s = int(rng.uniform(height * smin, height * smax))
        x = int(rng.integers(0, max(1, width - s)))
        y = int(rng.integers(0, max(1, height - s)))
        g[y:y + s, x:x + s] = face_patch(s)
    g8 = np.clip(np.rint(g), 0, 255).astype(np.uint8)
    out = np.empty((height, width, channels), np.uint8)
    tints = (5, 0, -5) if tint else (0, 0, 0)
    for c in range(3):
        out[..., c] = np.clip(g8.astype(np.int16) + tints[c], 0, 255).astype(np.uint8)
    if channels == 4:
        out[..., 3] = 255
    return out


def tracker_sequence(width: int, height: int, nframes: int, seed: int = 4, nsq: int = 3,
                     noise: int = 0) -> list[np.ndarray]:
    """cfg4: static background + bright squares translating 4 px/frame, BGRA."""
    rng = np.random.default_rng(seed)
    bg = np.clip(np.rint(_blur(np.clip(110 + 25 * rng.standard_normal((height, width)), 0, 255), 1.5)), 0, 255)
    sq = [(int(rng.integers(40, 121)), int(rng.integers(0, width // 2)), int(rng.integers(0, height - 130)),
           int(rng.choice([-4, 4])) if i else 4) for i in range(nsq)]
    frames = []
    for f in range(nframes):
        g = bg.copy()
        for (s, x0, y0, vx) in sq:
            x = (x0 + vx * f) % (width - s)
            g[y0:y0 + s, x:x + s] = 235
        if noise:
            idx = rng.integers(0, width * height, noise)
            g.reshape(-1)[idx] = 255
        out = np.empty((height, width, 4), np.uint8)
        out[..., 0] = out[..., 1] = out[..., 2] = g.astype(np.uint8)
        out[..., 3] = 255
        frames.append(out)
    return frames


def to_yuv420(bgr: np.ndarray, fmt: str = "I420") -> np.ndarray:
    """Pack a BGR frame (even width and height) as a 4:2:0 buffer of h*3/2 rows of w bytes — the layout a video decoder
    hands to GStreamer and cv2.cvtColor(.., COLOR_YUV2BGR_<fmt>) reads.  BT.601 studio range, chroma = mean of the 2x2
    block.  Input generator only: parity of the ingest path is defined on the YUV bytes, not on this conversion."""
    h, w, _ = bgr.shape
    assert w % 2 == 0 and h % 2 == 0
    b, g, r = (bgr[..., c].astype(np.float64) for c in range(3))
    y = 16 + (65.481 * r + 128.553 * g + 24.966 * b) / 255.0
    u = 128 + (-37.797 * r - 74.203 * g + 112.0 * b) / 255.0
    v = 128 + (112.0 * r - 93.786 * g - 18.214 * b) / 255.0
    sub = lambda a: a.reshape(h // 2, 2, w // 2, 2).mean(axis=(1, 3))
    q = lambda a: np.clip(np.rint(a), 0, 255).astype(np.uint8)
    y8, u8, v8 = q(y), q(sub(u)), q(sub(v))
    out = np.empty(w * h * 3 // 2, np.uint8)
    out[:w * h] = y8.reshape(-1)
    c = out[w * h:]
    if fmt == "I420":
        c[:w * h // 4] = u8.reshape(-1); c[w * h // 4:] = v8.reshape(-1)
    elif fmt == "YV12":
        c[:w * h // 4] = v8.reshape(-1); c[w * h // 4:] = u8.reshape(-1)
    elif fmt == "NV12":
        c[0::2] = u8.reshape(-1); c[1::2] = v8.reshape(-1)
    elif fmt == "NV21":
        c[0::2] = v8.reshape(-1); c[1::2] = u8.reshape(-1)
    else:
        raise ValueError(fmt)
    return out.reshape(h * 3 // 2, w)


def yuv420_planes(buf: np.ndarray, w: int, h: int, fmt: str = "I420"):
    """Plane views of a packed 4:2:0 buffer: (y, u, v) in I420 meaning for I420 / YV12, (y, uv) for NV12 / NV21."""
    flat = buf.reshape(-1)
    y = flat[:w * h].reshape(h, w)
    if fmt in ("NV12", "NV21"):
        return y, flat[w * h:].reshape(h // 2, w)
    a = flat[w * h:w * h + w * h // 4].reshape(h // 2, w // 2)
    b = flat[w * h + w * h // 4:].reshape(h // 2, w // 2)
    return (y, a, b) if fmt == "I420" else (y, b, a)
