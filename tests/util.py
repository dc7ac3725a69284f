"""Shared generators for the parity tests (seeded; BASELINE.md / SURVEY 8d)."""
import numpy as np


def rgba_random(w, h, seed, alpha="opaque"):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    if alpha == "opaque":
        a[..., 3] = 255
    elif alpha == "premul":      # valid premultiplied: c <= a
        al = a[..., 3:4].astype(np.uint16)
        a[..., :3] = (a[..., :3].astype(np.uint16) * al // 255).astype(np.uint8)
    elif alpha == "raw":         # arbitrary bytes, R may exceed A
        pass
    return a


def rgba_gradient(w, h):
    x = np.arange(w, dtype=np.int64)[None, :]
    y = np.arange(h, dtype=np.int64)[:, None]
    a = np.empty((h, w, 4), np.uint8)
    a[..., 0] = (x * 255 // max(w - 1, 1)).astype(np.uint8)
    a[..., 1] = (y * 255 // max(h - 1, 1)).astype(np.uint8)
    a[..., 2] = ((x + y) & 255).astype(np.uint8)
    a[..., 3] = 255
    return a


def synthetic_glyphs(W, H, seed, n=6, anchor=None):
    """Seeded random alpha masks in ~20x30 boxes, neighbours overlapping."""
    from collections import namedtuple
    G = namedtuple("G", "x0 y0 x1 y1 mask mp_x mp_y")
    rng = np.random.default_rng(seed)
    ax, ay = anchor if anchor else (max(W - 20 - 18 * n, 0), max(H - 60, 0))
    out = []
    for k in range(n):
        gw, gh = int(rng.integers(12, 24)), int(rng.integers(20, 34))
        m = rng.integers(0, 256, (gh, gw), dtype=np.uint8)
        m[rng.random((gh, gw)) < 0.35] = 0
        gx, gy = ax + 15 * k, ay + int(rng.integers(0, 8))     # 15 px pitch < width: overlap
        x0, y0 = max(gx, 0), max(gy, 0)
        x1, y1 = min(gx + gw, W), min(gy + gh, H)
        if x0 >= x1 or y0 >= y1:
            continue
        # freetype passes mp = (0, dr.Min.Y - glyphRect.Min.Y)
        out.append(G(x0, y0, x1, y1, m, 0, y0 - gy))
    return out


def diff_stats(a, b):
    d = np.abs(a.astype(np.int16) - b.astype(np.int16))
    return int(d.max()) if d.size else 0, float((d > 0).mean()) if d.size else 0.0
