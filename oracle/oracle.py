"""ctypes front-end of the CPU oracle (oracle/ip_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(imageprocessor_b200) never imports this module.

PARITY UNPINNED: see oracle/ip_oracle.h -- the reference has no tests/goldens
and cannot be built here (Go, un-vendored deps, no toolchain).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libip_oracle.so")

RGBA8, NRGBA8, GRAY8, YCBCR444, YCBCR422, YCBCR420, YCBCR440, RGBA64, NRGBA64, GRAY16, PALETTED_RGBA, PALETTED_NRGBA = range(12)
OP_OVER, OP_SRC = 0, 1


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc, -ffp-contract=off)."""
    srcs = [os.path.join(_HERE, n) for n in ("ip_oracle.c", "ip_jpeg_oracle.c", "ip_oracle.h")]
    if (force or not os.path.exists(_LIB_PATH)
            or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(s) for s in srcs)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libip_oracle.so"])
    return _LIB_PATH


class _Image(C.Structure):
    _fields_ = [("layout", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
                ("plane", C.c_void_p * 3), ("stride", C.c_int32 * 3)]


class _Glyph(C.Structure):
    _fields_ = [("x0", C.c_int32), ("y0", C.c_int32), ("x1", C.c_int32), ("y1", C.c_int32),
                ("mp_x", C.c_int32), ("mp_y", C.c_int32), ("mask_stride", C.c_int32),
                ("mask_w", C.c_int32), ("mask_h", C.c_int32), ("mask", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        ip = C.POINTER(C.c_int)
        L.ipo_keep_aspect_dims.argtypes = [C.c_int] * 4 + [ip, ip]
        L.ipo_thumb_fit_dims.argtypes = [C.c_int] * 3 + [ip, ip]
        L.ipo_crop_square.argtypes = [C.c_int] * 2 + [ip, ip, ip]
        L.ipo_watermark_anchor.argtypes = [C.c_char_p] + [C.c_int] * 4 + [ip, ip]
        L.ipo_watermark_height_px.argtypes = [C.c_double]
        L.ipo_parse_color.argtypes = [C.c_char_p, C.c_double, C.POINTER(C.c_uint8)]
        L.ipo_distrib.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ipo_scale_bilinear.argtypes = [C.POINTER(_Image)] + [C.c_int] * 4 + [C.c_void_p] + [C.c_int] * 4
        L.ipo_resize_image.argtypes = [C.POINTER(_Image), C.c_int, C.c_int, C.c_void_p]
        L.ipo_crop_and_resize.argtypes = [C.POINTER(_Image), C.c_int, C.c_void_p]
        L.ipo_draw_src.argtypes = [C.POINTER(_Image), C.c_void_p, C.c_int]
        L.ipo_glyph_over.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint8), C.POINTER(_Glyph)]
        L.ipo_watermark.argtypes = [C.POINTER(_Image), C.c_void_p, C.c_int, C.POINTER(C.c_uint8),
                                    C.POINTER(_Glyph), C.c_int]
        L.ipo_rgba_to_ycbcr420.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ipo_jpeg_encode_rgba.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
        L.ipo_jpeg_encode_rgba.restype = C.c_size_t
        L.ipo_jpeg_encode_ycbcr.argtypes = [C.POINTER(_Image), C.c_int, C.c_void_p, C.c_size_t]
        L.ipo_jpeg_encode_ycbcr.restype = C.c_size_t
        L.ipo_jpeg_encode_gray.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
        L.ipo_jpeg_encode_gray.restype = C.c_size_t
        L.ipo_bench_batch.argtypes = [C.POINTER(_Image), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.POINTER(C.c_uint8), C.POINTER(_Glyph),
                                      C.c_int, C.POINTER(C.c_uint64)]
        L.ipo_bench_batch.restype = C.c_double
        _lib = L
    return _lib


@dataclass
class Raster:
    """A decoded image as Go's image.Decode would hand it to the ops."""
    layout: int
    width: int
    height: int
    planes: Tuple[np.ndarray, ...]  # C-contiguous uint8 2-D (rows x stride bytes)

    @staticmethod
    def rgba(a: np.ndarray, layout: int = RGBA8) -> "Raster":
        a = np.ascontiguousarray(a, dtype=np.uint8)
        assert a.ndim == 3 and a.shape[2] == 4
        return Raster(layout, a.shape[1], a.shape[0], (a.reshape(a.shape[0], -1),))

    @staticmethod
    def gray(a: np.ndarray) -> "Raster":
        a = np.ascontiguousarray(a, dtype=np.uint8)
        return Raster(GRAY8, a.shape[1], a.shape[0], (a,))

    @staticmethod
    def deep(a: np.ndarray, layout: int) -> "Raster":
        """*image.RGBA64 / *image.NRGBA64 (h, w, 4) or *image.Gray16 (h, w) from uint16 values: stored big-endian as Go does."""
        be = np.ascontiguousarray(a.astype(">u2"))
        h, w = a.shape[:2]
        return Raster(layout, w, h, (be.view(np.uint8).reshape(h, -1),))

    @staticmethod
    def paletted(indices: np.ndarray, palette_rgba: np.ndarray, nrgba_entries: bool) -> "Raster":
        """*image.Paletted: index bytes + up to 256 palette entries (color.RGBA, or color.NRGBA when nrgba_entries)."""
        idx = np.ascontiguousarray(indices, dtype=np.uint8)
        pal = np.zeros((256, 4), np.uint8)
        pal[:len(palette_rgba)] = palette_rgba
        return Raster(PALETTED_NRGBA if nrgba_entries else PALETTED_RGBA, idx.shape[1], idx.shape[0], (idx, pal.reshape(1, -1)))

    @staticmethod
    def ycbcr(y: np.ndarray, cb: np.ndarray, cr: np.ndarray, layout: int) -> "Raster":
        y, cb, cr = (np.ascontiguousarray(p, dtype=np.uint8) for p in (y, cb, cr))
        return Raster(layout, y.shape[1], y.shape[0], (y, cb, cr))

    def c(self) -> _Image:
        im = _Image()
        im.layout, im.width, im.height = self.layout, self.width, self.height
        for k, p in enumerate(self.planes):
            im.plane[k] = p.ctypes.data
            im.stride[k] = p.strides[0]
        return im


def chroma_shape(layout: int, w: int, h: int) -> Tuple[int, int]:
    """(rows, cols) of the Cb/Cr planes as image.NewYCbCr sizes them."""
    if layout == YCBCR444:
        return h, w
    if layout == YCBCR422:
        return h, (w + 1) // 2
    if layout == YCBCR420:
        return (h + 1) // 2, (w + 1) // 2
    if layout == YCBCR440:
        return (h + 1) // 2, w
    raise ValueError(layout)


def keep_aspect_dims(ow, oh, w, h):
    a, b = C.c_int(), C.c_int()
    lib().ipo_keep_aspect_dims(ow, oh, w, h, a, b)
    return a.value, b.value


def thumb_fit_dims(ow, oh, size):
    a, b = C.c_int(), C.c_int()
    lib().ipo_thumb_fit_dims(ow, oh, size, a, b)
    return a.value, b.value


def crop_square(ow, oh):
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    lib().ipo_crop_square(ow, oh, a, b, c)
    return a.value, b.value, c.value


def watermark_anchor(position: str, W, H, width_px, height_px):
    a, b = C.c_int(), C.c_int()
    lib().ipo_watermark_anchor(position.encode(), W, H, width_px, height_px, a, b)
    return a.value, b.value


def watermark_height_px(font_size: float) -> int:
    return lib().ipo_watermark_height_px(font_size)


def parse_color(s: str, opacity: float):
    out = (C.c_uint8 * 4)()
    rc = lib().ipo_parse_color(s.encode(), opacity, out)
    return rc, tuple(out)


def distrib(dw: int, sw: int):
    n = lib().ipo_distrib(dw, sw, None, None, None, None)
    starts = np.zeros(dw + 1, np.int32)
    coords = np.zeros(max(n, 1), np.int32)
    weights = np.zeros(max(n, 1), np.float64)
    inv = np.zeros(dw, np.float64)
    lib().ipo_distrib(dw, sw, starts.ctypes.data, coords.ctypes.data, weights.ctypes.data,
                      inv.ctypes.data)
    return starts, coords[:n], weights[:n], inv


def scale_bilinear(src: Raster, rect, dw, dh, op=OP_OVER, dst: Optional[np.ndarray] = None):
    sx0, sy0, sw, sh = rect
    if dst is None:
        dst = np.zeros((dh, dw, 4), np.uint8)
    im = src.c()
    rc = lib().ipo_scale_bilinear(C.byref(im), sx0, sy0, sw, sh, dst.ctypes.data, dst.strides[0],
                                  dw, dh, op)
    if rc:
        raise ValueError("ipo_scale_bilinear failed")
    return dst


def resize_image(src: Raster, dw, dh) -> np.ndarray:
    dst = np.empty((dh, dw, 4), np.uint8)
    im = src.c()
    if lib().ipo_resize_image(C.byref(im), dw, dh, dst.ctypes.data):
        raise ValueError("ipo_resize_image failed")
    return dst


def crop_and_resize(src: Raster, size) -> np.ndarray:
    dst = np.empty((size, size, 4), np.uint8)
    im = src.c()
    if lib().ipo_crop_and_resize(C.byref(im), size, dst.ctypes.data):
        raise ValueError("ipo_crop_and_resize failed")
    return dst


def draw_src(src: Raster) -> np.ndarray:
    dst = np.zeros((src.height, src.width, 4), np.uint8)
    im = src.c()
    if lib().ipo_draw_src(C.byref(im), dst.ctypes.data, dst.strides[0]):
        raise ValueError("ipo_draw_src failed")
    return dst


def rgba_to_ycbcr420(rgba: np.ndarray):
    """(Y, Cb, Cr) planes Go's image/jpeg writer derives from an *image.RGBA (writer.go rgbaToYCbCr + scale)."""
    a = np.ascontiguousarray(rgba, dtype=np.uint8)
    h, w = a.shape[:2]
    y = np.empty((h, w), np.uint8)
    cb, cr = np.empty(((h + 1) // 2, (w + 1) // 2), np.uint8), np.empty(((h + 1) // 2, (w + 1) // 2), np.uint8)
    if lib().ipo_rgba_to_ycbcr420(a.ctypes.data, a.strides[0], w, h, y.ctypes.data, cb.ctypes.data, cr.ctypes.data):
        raise ValueError("ipo_rgba_to_ycbcr420 failed")
    return y, cb, cr


def _jpeg_result(n: int, out: np.ndarray) -> bytes:
    if n == 0 or n == C.c_size_t(-1).value:
        raise ValueError("jpeg oracle: bad argument or output buffer too small")
    return out[:n].tobytes()


def jpeg_encode_rgba(rgba: np.ndarray, quality: int = 85) -> bytes:
    """jpeg.Encode(w, m *image.RGBA, &jpeg.Options{Quality: quality}) -- the whole file (Go 1.24 writer.go)."""
    a = np.ascontiguousarray(rgba, dtype=np.uint8)
    h, w = a.shape[:2]
    out = np.empty(w * h * 8 + 4096, np.uint8)
    return _jpeg_result(lib().ipo_jpeg_encode_rgba(a.ctypes.data, a.strides[0], w, h, quality, out.ctypes.data, out.size), out)


def jpeg_encode_ycbcr(src: "Raster", quality: int = 85) -> bytes:
    """jpeg.Encode of an *image.YCbCr (any subsample ratio): the writer's yCbCrToYCbCr path."""
    im = src.c()
    out = np.empty(src.width * src.height * 8 + 4096, np.uint8)
    return _jpeg_result(lib().ipo_jpeg_encode_ycbcr(C.byref(im), quality, out.ctypes.data, out.size), out)


def jpeg_encode_gray(gray: np.ndarray, quality: int = 85) -> bytes:
    """jpeg.Encode of an *image.Gray."""
    a = np.ascontiguousarray(gray, dtype=np.uint8)
    h, w = a.shape
    out = np.empty(w * h * 8 + 4096, np.uint8)
    return _jpeg_result(lib().ipo_jpeg_encode_gray(a.ctypes.data, a.strides[0], w, h, quality, out.ctypes.data, out.size), out)


@dataclass
class Glyph:
    """One DrawMask call of freetype's DrawString: dst rect, mask, mask point."""
    x0: int
    y0: int
    x1: int
    y1: int
    mask: np.ndarray  # uint8 2-D (*image.Alpha)
    mp_x: int = 0
    mp_y: int = 0


def _glyph_array(glyphs: Sequence[Glyph]):
    arr = (_Glyph * max(len(glyphs), 1))()
    keep = []
    for k, g in enumerate(glyphs):
        m = np.ascontiguousarray(g.mask, dtype=np.uint8)
        keep.append(m)
        arr[k].x0, arr[k].y0, arr[k].x1, arr[k].y1 = g.x0, g.y0, g.x1, g.y1
        arr[k].mp_x, arr[k].mp_y = g.mp_x, g.mp_y
        arr[k].mask_stride = m.strides[0]
        arr[k].mask_w, arr[k].mask_h = m.shape[1], m.shape[0]
        arr[k].mask = m.ctypes.data
    return arr, keep


def glyph_over(dst: np.ndarray, rgba, g: Glyph) -> None:
    arr, keep = _glyph_array([g])
    col = (C.c_uint8 * 4)(*rgba)
    lib().ipo_glyph_over(dst.ctypes.data, dst.strides[0], col, C.byref(arr[0]))


def watermark(src: Raster, rgba, glyphs: Sequence[Glyph]) -> np.ndarray:
    dst = np.zeros((src.height, src.width, 4), np.uint8)
    arr, keep = _glyph_array(glyphs)
    col = (C.c_uint8 * 4)(*rgba)
    im = src.c()
    if lib().ipo_watermark(C.byref(im), dst.ctypes.data, dst.strides[0], col, arr, len(glyphs)):
        raise ValueError("ipo_watermark failed")
    return dst


def bench_batch(rasters: List[Raster], n_threads: int, ops: int, rw=1024, rh=768, keep_aspect=True,
                thumb_size=200, rgba=(255, 255, 255, 127), glyphs: Sequence[Glyph] = ()):
    """Time the restated reference CPU path; returns (seconds, checksum)."""
    n = len(rasters)
    imgs = (_Image * n)(*[r.c() for r in rasters])
    arr, keep = _glyph_array(glyphs)
    col = (C.c_uint8 * 4)(*rgba)
    cs = C.c_uint64()
    secs = lib().ipo_bench_batch(imgs, n, n_threads, ops, rw, rh, int(keep_aspect), thumb_size,
                                 col, arr, len(glyphs), C.byref(cs))
    return secs, cs.value


def drawstring_layout(face, text, size, W, H, px, py):
    """golang/freetype freetype.go (*Context).DrawString + glyph() restated for the checker side (26.6 pen, the
    (rune, quarter-pixel) mask cache, per-rune clipped rectangle with mask point (0, dy))."""
    pen_x, pen_y = px * 64, py * 64
    cache, out = {}, []
    for ch in text:
        r = ord(ch)
        ix, fx, iy, fy = pen_x >> 6, pen_x & 63, pen_y >> 6, pen_y & 63
        idx = face.index(r)
        slot = ((idx % 256) * 4 + fx // 16) * 1 + fy // 64
        if slot not in cache or cache[slot][0] != idx:
            cache[slot] = (idx, face.mask(r, size, fx, fy))
        adv, ox, oy, m = cache[slot][1]
        if m.size:
            gx0, gy0 = ix + ox, iy + oy
            x0, y0, x1, y1 = max(gx0, 0), max(gy0, 0), min(gx0 + m.shape[1], W), min(gy0 + m.shape[0], H)
            if x0 < x1 and y0 < y1:
                out.append((x0, y0, min(x1, x0 + m.shape[1]), y1, m, 0, y0 - gy0))
        pen_x += adv
    return out
