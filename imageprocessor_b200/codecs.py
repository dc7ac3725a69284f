"""Host codec stand-ins for the in-process harness: what Go's image.Decode / jpeg.Encode / png.Encode do on the
reference's host (internal/usecase/processor/image_processor.go:47, operations/resize.go:78-91,
operations/watermark.go:66-79).  north_star keeps the codecs on the host, timed and reported separately; no Go
toolchain exists here, so PIL (libjpeg-turbo, libpng) stands in and the numbers are labelled as such.

What matters for parity is the CONCRETE raster type image.Decode hands the operations, because x/image's Scale and
stdlib draw.Draw switch on it (Spec R/W):

    JPEG  colour, 4:2:0 / 4:2:2 / 4:4:4 ...  -> *image.YCbCr with that subsample ratio   (planar; never RGB)
    JPEG  grayscale                           -> *image.Gray
    PNG   8-bit truecolour                    -> *image.RGBA  (opaque)
    PNG   8-bit truecolour + alpha            -> *image.NRGBA (straight alpha)
    PNG   8-bit gray                          -> *image.Gray
    PNG   16-bit gray / colour / + alpha      -> *image.Gray16 / *image.RGBA64 / *image.NRGBA64 (generic RGBA64At path)
    PNG / GIF paletted                        -> *image.Paletted: Scale and draw.Draw reach it through At(x,y).RGBA(),
                                                 which for color.RGBA / color.NRGBA palette entries is byte * 0x101 resp.
                                                 the 16-bit premultiply of scaleX_NRGBA -- bit-identical to expanding the
                                                 palette into an RGBA8 (no translucent entry) or NRGBA8 raster first
                                                 (expand_paletted below; tests/test_paletted.py proves the identity
                                                 against the oracle's generic At().RGBA() path).

libjpeg-turbo hands back chroma already upsampled; the 4:2:0 planes here are rebuilt by taking every second sample,
which is not the file's stored chroma (the IDCT output before "fancy" upsampling) but is a valid 4:2:0 raster of the
same picture.  Parity is defined on identical decoded planes (north_star), never across decoders.
"""
from __future__ import annotations

import io
from typing import Optional, Tuple

import numpy as np

from . import _lib as L
from .engine import Image


def decode(data: bytes) -> Tuple[Image, str]:
    """image.Decode stand-in: (raster of the concrete type Go would produce, format name)."""
    from PIL import Image as PI
    if data[:8] == b"\x89PNG\r\n\x1a\n" and len(data) > 26 and data[24] == 16:
        return _decode_png16(data, data[25]), "png"
    im = PI.open(io.BytesIO(data))
    fmt = (im.format or "").lower()
    if fmt == "jpeg":
        if im.mode == "L":
            return Image.from_gray(np.asarray(im)), "jpeg"
        im.draft("YCbCr", im.size)
        if im.mode != "YCbCr":
            im = im.convert("YCbCr")
        ycc = np.asarray(im)
        ss = getattr(im, "layer", None)
        y = np.ascontiguousarray(ycc[..., 0])
        sub_h, sub_v = 2, 2
        if ss and len(ss) >= 1:           # [(id, hsamp, vsamp, qtable), ...]: luma sampling factors give the ratio
            sub_h, sub_v = ss[0][1], ss[0][2]
        if sub_h == 2 and sub_v == 2:
            lay, cb, cr = L.YCBCR420, ycc[::2, ::2, 1], ycc[::2, ::2, 2]
        elif sub_h == 2 and sub_v == 1:
            lay, cb, cr = L.YCBCR422, ycc[:, ::2, 1], ycc[:, ::2, 2]
        elif sub_h == 1 and sub_v == 2:
            lay, cb, cr = L.YCBCR440, ycc[::2, :, 1], ycc[::2, :, 2]
        else:
            lay, cb, cr = L.YCBCR444, ycc[..., 1], ycc[..., 2]
        return Image.from_ycbcr(y, np.ascontiguousarray(cb), np.ascontiguousarray(cr), lay), "jpeg"
    if im.mode == "P":
        pal = np.array(im.getpalette("RGBA"), np.uint8).reshape(-1, 4)
        return expand_paletted(np.asarray(im), pal), fmt or "png"
    if im.mode == "L":
        return Image.from_gray(np.asarray(im)), fmt or "png"
    if im.mode == "RGBA":
        return Image.from_rgba(np.asarray(im), L.NRGBA8), fmt or "png"
    rgb = np.asarray(im.convert("RGB"))
    a = np.empty(rgb.shape[:2] + (4,), np.uint8)
    a[..., :3] = rgb
    a[..., 3] = 255
    return Image.from_rgba(a, L.RGBA8, opaque_hint=True), fmt or "png"


def _decode_png16(data: bytes, color_type: int) -> Image:
    """A 16-bit PNG as Go's image/png decodes it: gray -> *image.Gray16, truecolour -> *image.RGBA64 (opaque),
    gray+alpha / truecolour+alpha -> *image.NRGBA64.  (PIL narrows 16-bit colour to 8 bits; OpenCV keeps it.)"""
    import cv2
    a = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_UNCHANGED)
    if a is None or a.dtype != np.uint16:
        raise ValueError("png: cannot decode 16-bit image")
    if a.ndim == 2:                       # gray
        return Image.from_deep(np.ascontiguousarray(a), L.GRAY16)
    h, w, c = a.shape
    out = np.empty((h, w, 4), np.uint16)
    if c == 2:                            # gray + alpha
        out[..., 0] = out[..., 1] = out[..., 2] = a[..., 0]
        out[..., 3] = a[..., 1]
        return Image.from_deep(out, L.NRGBA64)
    out[..., 0], out[..., 1], out[..., 2] = a[..., 2], a[..., 1], a[..., 0]     # OpenCV hands back BGR(A)
    if c == 4:
        out[..., 3] = a[..., 3]
        return Image.from_deep(out, L.NRGBA64)
    out[..., 3] = 0xFFFF
    return Image.from_deep(out, L.RGBA64)


def expand_paletted(indices: np.ndarray, palette_rgba: np.ndarray) -> Image:
    """*image.Paletted -> the RGBA8 / NRGBA8 raster whose scaleX_RGBA / scaleX_NRGBA samples equal the generic
    path's Palette[i].RGBA() (GIF and opaque PNG palettes hold color.RGBA entries; a PNG tRNS chunk makes them
    color.NRGBA).  Fully transparent GIF entries are color.RGBA{0,0,0,0}: valid premultiplied, stay RGBA8."""
    pal = np.ascontiguousarray(palette_rgba, np.uint8)
    px = pal[indices]
    translucent = bool(((pal[:, 3] != 255) & ~((pal == 0).all(axis=1))).any())
    return Image.from_rgba(np.ascontiguousarray(px), L.NRGBA8 if translucent else L.RGBA8)


def encode(rgba: np.ndarray, fmt: str, quality: int = 85, png_level: int = 1) -> bytes:
    """jpeg.Encode(q85) / png.Encode / gif.Encode stand-in for an *image.RGBA result.  png_level: Go's png.Encode
    uses DefaultCompression (zlib 6); the harness defaults to 1 so that the stand-in does not dominate, and says so."""
    from PIL import Image as PI
    im = PI.fromarray(rgba, "RGBA")
    buf = io.BytesIO()
    if fmt == "jpeg":
        im.convert("RGB").save(buf, "JPEG", quality=quality)
    elif fmt == "png":
        im.save(buf, "PNG", compress_level=png_level)
    else:
        im.convert("RGB").quantize(256).save(buf, "GIF")
    return buf.getvalue()


def synth_picture(w: int, h: int, seed: int) -> np.ndarray:
    """Photo-like synthetic RGB (smooth separable gradients + seeded low-amplitude noise), (h, w, 3) uint8.
    Built from 1-D profiles and integer adds so that a 48 MP picture costs a fraction of a second."""
    rng = np.random.default_rng(seed)
    x = np.linspace(0.0, 1.0, w, dtype=np.float32)
    y = np.linspace(0.0, 1.0, h, dtype=np.float32)
    ph = rng.uniform(0, 6.28, 6)
    noise = rng.integers(-6, 7, (h, w), dtype=np.int8).astype(np.int16)
    out = np.empty((h, w, 3), np.uint8)
    for c in range(3):
        gx = (64.0 + 50.0 * np.sin(ph[c] + 3.1 * x * (c + 1)) + 30.0 * x).astype(np.int16)
        gy = (64.0 + 50.0 * np.cos(ph[3 + c] + 2.3 * y * (3 - c)) + 30.0 * y).astype(np.int16)
        v = gx[None, :] + gy[:, None]
        v += noise
        np.clip(v, 0, 255, out=v)
        out[..., c] = v
    return out


def synth_file(w: int, h: int, seed: int, kind: str) -> bytes:
    """An encoded source file: kind 'jpeg' (4:2:0, q90), 'png' (truecolour) or 'png-alpha' (truecolour + alpha)."""
    from PIL import Image as PI
    rgb = synth_picture(w, h, seed)
    buf = io.BytesIO()
    if kind == "jpeg":
        PI.fromarray(rgb, "RGB").save(buf, "JPEG", quality=90, subsampling="4:2:0")
    elif kind == "png":
        PI.fromarray(rgb, "RGB").save(buf, "PNG", compress_level=1)
    else:
        a = np.empty((h, w, 4), np.uint8)
        a[..., :3] = rgb
        rng = np.random.default_rng(seed + 7)    # alpha falls from 255 on the left to 255 - 3 * (0..63) on the right
        a[..., 3] = 255 - rng.integers(0, 64, (h, w), dtype=np.uint8) * (np.arange(w, dtype=np.int32) * 4 // max(w, 1)).astype(np.uint8)[None, :]
        PI.fromarray(a, "RGBA").save(buf, "PNG", compress_level=1)
    return buf.getvalue()
