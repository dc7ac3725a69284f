#!/usr/bin/env python
"""Generate tests/golden/golden_v1.npz -- known answers that pin the oracle.

WHAT THESE ARE (and are not).  The reference (sj-shoff/ImageProcessor) is pure Go with
un-vendored module dependencies; this image has no Go toolchain, the reference ships no
tests, fixtures or golden images, so outputs of the real reference cannot be produced
here: parity with the Go binary stays UNPINNED (oracle/ip_oracle.h).  What this
script commits instead are answers computed by code that shares NOTHING with oracle/ or
with the CUDA path:

  resample   torch.nn.functional.interpolate(mode="bilinear", antialias=True) in float64
             -- PyTorch's own C++ implementation of the separable area-scaled tent filter
             (the same filter as golang.org/x/image/draw.BiLinear, SURVEY.md Spec R) --
             followed by the x/image quantiser uint8(ftou(v) >> 8) written here in numpy.
             Sources are fed as the 16-bit premultiplied samples Spec R's scaleX_<type>
             produce (numpy integer code below, independent of the oracle's C).
  blend      stdlib image/draw drawGlyphOver evaluated with Python's arbitrary-precision
             integers reduced mod 2^32 at the points Go's uint32 arithmetic wraps.
  geometry   literal (w,h) -> (nw,nh) pairs worked out by hand from resize.go:63-72 and
             thumbnail.go:52-63,115-127 (SURVEY.md 8c iv).

Inputs are regenerated from seeds at test time (tests/test_golden.py), only the
expected outputs are stored.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.golden.cases import RESAMPLE_CASES, BLEND_CASES, make_source, make_blend_case  # noqa: E402


# ---- Spec R, source adaptors: 16-bit premultiplied samples (H, W, 4) as float64 -------
def samples16(kind, planes):
    if kind in ("rgba", "rgba_premul"):
        return planes[0].astype(np.float64) * 257.0
    if kind == "nrgba":
        p = planes[0].astype(np.uint64)
        a16 = p[..., 3:4] * 0x101
        c16 = p[..., :3] * a16 // 0xff
        return np.concatenate([c16, a16], axis=2).astype(np.float64)
    if kind == "gray":
        y = planes[0].astype(np.float64) * 257.0
        return np.stack([y, y, y, np.full_like(y, 65535.0)], axis=2)
    if kind.startswith("ycbcr"):
        y, cb, cr = planes
        h, w = y.shape
        yy, xx = np.mgrid[0:h, 0:w]
        if kind == "ycbcr444":
            ci = (yy, xx)
        elif kind == "ycbcr422":
            ci = (yy, xx // 2)
        elif kind == "ycbcr420":
            ci = (yy // 2, xx // 2)
        else:
            ci = (yy // 2, xx)
        yy1 = y.astype(np.int64) * 0x10101
        cb1 = cb[ci].astype(np.int64) - 128
        cr1 = cr[ci].astype(np.int64) - 128
        r = np.clip((yy1 + 91881 * cr1) >> 8, 0, 0xffff)
        g = np.clip((yy1 - 22554 * cb1 - 46802 * cr1) >> 8, 0, 0xffff)
        b = np.clip((yy1 + 116130 * cb1) >> 8, 0, 0xffff)
        return np.stack([r, g, b, np.full_like(r, 0xffff)], axis=2).astype(np.float64)
    raise ValueError(kind)


def to_rgba8(s16):
    """The 1:1 first pass of cropAndResize: uint8(min(c16, a16) >> 8) (Spec R Src store)."""
    s = s16.astype(np.int64)
    a = s[..., 3:4]
    c = np.minimum(s[..., :3], a)
    return (np.concatenate([c, a], axis=2) >> 8).astype(np.uint8)


def quantise(v):
    """x/image: premultiplied clamp, ftou = int32(0xffff*f + 0.5) clamped, then >> 8."""
    a = v[..., 3:4]
    v = np.concatenate([np.minimum(v[..., :3], a), a], axis=2)
    u = np.clip(np.floor(65535.0 * v + 0.5), 0, 65535).astype(np.int64)
    return (u >> 8).astype(np.uint8)


def torch_scale(s16, dw, dh):
    t = torch.from_numpy(np.ascontiguousarray(s16.transpose(2, 0, 1))[None]) / 65535.0
    o = F.interpolate(t.double(), size=(dh, dw), mode="bilinear", antialias=True, align_corners=False)
    return quantise(o[0].permute(1, 2, 0).numpy())


# ---- Spec W, drawGlyphOver with explicit uint32 wrap --------------------------------
M32 = (1 << 32) - 1


def glyph_over(dst, color, glyphs):
    """dst (H,W,4) uint8 modified in place; glyphs = [(x0,y0,x1,y1,mask,mp_x,mp_y)] in string order."""
    sr, sg, sb, sa = (c * 0x101 for c in color)
    m = 0xffff
    for x0, y0, x1, y1, mask, mp_x, mp_y in glyphs:
        for y in range(y0, y1):
            for x in range(x0, x1):
                ma = int(mask[y - y0 + mp_y, x - x0 + mp_x])
                if ma == 0:
                    continue
                ma |= ma << 8
                a = (((m - (sa * ma & M32) // m) & M32) * 0x101) & M32
                for ch, s in enumerate((sr, sg, sb, sa)):
                    d = int(dst[y, x, ch])
                    v = ((d * a & M32) + (s * ma & M32)) & M32
                    dst[y, x, ch] = ((v // m) >> 8) & 0xff


def main():
    out = {}
    for name, kind, w, h, seed, ops in RESAMPLE_CASES:
        planes = make_source(kind, w, h, seed)
        s16 = samples16(kind, planes)
        for op in ops:
            if op[0] == "resize":
                _, dw, dh = op
                out[f"{name}/resize_{dw}x{dh}"] = torch_scale(s16, dw, dh)
            elif op[0] == "thumb":
                _, size = op
                cs = min(w, h)
                cx, cy = ((w - h) // 2, 0) if w > h else (0, (h - w) // 2)
                crop8 = to_rgba8(s16[cy:cy + cs, cx:cx + cs])
                out[f"{name}/thumb_{size}"] = torch_scale(crop8.astype(np.float64) * 257.0, size, size)
    for name, w, h, seed, color, n in BLEND_CASES:
        dst, glyphs = make_blend_case(w, h, seed, n)
        glyph_over(dst, color, glyphs)
        out[f"{name}/blend"] = dst
    # exhaustive (dst byte x mask byte) table for the reference's default colour, white @ 127
    tab = np.zeros((256, 256, 4), np.uint8)
    for d in range(256):
        row = np.full((1, 256, 4), d, np.uint8)
        glyph_over(row, (255, 255, 255, 127), [(0, 0, 256, 1, np.arange(256, dtype=np.uint8)[None, :], 0, 0)])
        tab[d] = row[0]
    out["blend_table_white127"] = tab
    path = os.path.join(HERE, "golden_v1.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
