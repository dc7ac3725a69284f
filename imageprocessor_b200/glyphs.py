"""Host-side glyph masks for the watermark (a stand-in, and labelled as one).

In the reference the masks come from github.com/golang/freetype rasterising the
embedded Go Regular TTF (operations/watermark.go:29-38,98-108,151); in production
the Go host keeps doing exactly that and hands the per-rune *image.Alpha masks to
ipg_submit, so mask parity holds by construction.  Neither that
rasteriser nor that font exists in this image, so the Python harness rasterises
with PIL/FreeType and whatever scalable font PIL ships.  Only the *blend* of the
masks is under parity test; mask parity is unpinned.

The layout below mirrors freetype.Context.DrawString: pen starts at the anchor
(baseline), one DrawMask per rune in string order, rect clipped to the image,
mask point mp = (0, dr.Min.Y - glyphRect.Min.Y).
"""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np

from .engine import GlyphMask

MARGIN = 20  # watermark.go:121


def text_height_px(font_size: float) -> int:
    """watermark.go:116,118: fixed.Int26_6(fontSize*64*1.2).Ceil()"""
    fx = int(font_size * 64 * 1.2)
    return (fx + 63) >> 6


def anchor(position: str, W: int, H: int, width_px: int, height_px: int) -> Tuple[int, int]:
    """watermark.go:121-148 (Go int division truncates toward zero)."""
    def div2(v: int) -> int:
        return int(v / 2) if v < 0 else v // 2
    m = MARGIN
    table = {
        "top-left": (m, m + height_px),
        "top-right": (W - width_px - m, m + height_px),
        "top-center": (div2(W - width_px), m + height_px),
        "bottom-left": (m, H - m),
        "bottom-right": (W - width_px - m, H - m),
        "bottom-center": (div2(W - width_px), H - m),
        "center": (div2(W - width_px), div2(H + height_px)),
    }
    return table.get(position, table["bottom-right"])


def parse_color(s: str, opacity: float) -> Tuple[Tuple[int, int, int, int], bool]:
    """watermark.go:159-186 + the caller's black fallback (:94-97).
    Returns (color.RGBA bytes, ok)."""
    def atoi(p: str):
        q = p[1:] if p[:1] in "+-" else p
        if not q or not all("0" <= ch <= "9" for ch in q):
            return None
        v = int(p)
        return v if -(1 << 63) <= v < (1 << 63) else None

    a_op = int(255 * opacity) & 0xFF
    parts = s.replace(" ", "").split(",")
    if len(parts) not in (3, 4):
        return (0, 0, 0, a_op), False
    rgb = [atoi(p) for p in parts[:3]]
    if any(v is None for v in rgb):
        return (0, 0, 0, a_op), False
    r, g, b = (min(max(v, 0), 255) for v in rgb)
    a = a_op
    if len(parts) == 4:
        av = atoi(parts[3])
        if av is not None:
            a = min(max(av, 0), 255)
    return (r, g, b, a), True


def rasterize_runes(text: str, font_size: float):
    """[(advance_px_float, x_off, y_off, mask)] per rune, offsets relative to the pen
    on the baseline (y down).  PIL/FreeType stand-in for truetype.Face + raster."""
    from PIL import ImageFont
    font = ImageFont.load_default(size=font_size)
    out = []
    for ch in text:
        adv = float(font.getlength(ch))
        m, off = font.getmask2(ch, mode="L", anchor="ls")
        w, h = m.size
        if w == 0 or h == 0:
            out.append((adv, 0, 0, np.zeros((0, 0), np.uint8)))
            continue
        mask = np.frombuffer(bytes(m), np.uint8).reshape(h, w).copy()
        out.append((adv, int(off[0]), int(off[1]), mask))
    return out


def layout_watermark(W: int, H: int, text: str, position: str = "bottom-right",
                     font_size: float = 36.0) -> List[GlyphMask]:
    """Per-rune DrawMask list for addTextWatermark's DrawString call."""
    runes = rasterize_runes(text, font_size)
    width_px = int(math.ceil(sum(r[0] for r in runes)))
    height_px = text_height_px(font_size)
    px, py = anchor(position, W, H, width_px, height_px)
    pen = float(px)
    glyphs: List[GlyphMask] = []
    for adv, xo, yo, mask in runes:
        if mask.size:
            gx, gy = int(math.floor(pen)) + xo, py + yo
            gh, gw = mask.shape
            x0, y0, x1, y1 = max(gx, 0), max(gy, 0), min(gx + gw, W), min(gy + gh, H)
            if x0 < x1 and y0 < y1:
                # freetype: mp = (0, dr.Min.Y - glyphRect.Min.Y); keep the rect inside the mask
                x1 = min(x1, x0 + gw)
                glyphs.append(GlyphMask(x0, y0, x1, y1, mask, 0, y0 - gy))
        pen += adv
    return glyphs
