"""ctypes front-end of the host layer (include/ipgpu_host.h): the reference's
processor.ImageProcessor.Process behind the same call shape, for the in-process harness.

What the reference keeps on the host stays on the host here too and is supplied as
callbacks: the encoders (PIL stands in for Go's image/jpeg, image/png, image/gif), the
file repository (a dict stands in for MinIO, internal/repository/image/cloud/minio), and
the truetype face + rasteriser (PIL/FreeType stands in for golang/freetype + Go Regular;
mask parity is unpinned, see glyphs.py).  The raster work runs in libipgpu.so only.
"""
from __future__ import annotations

import ctypes as C
import io
import json
import threading
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib as L
from .engine import Engine, Image


class IphGlyph(C.Structure):
    _fields_ = [("advance_26_6", C.c_int32), ("off_x", C.c_int32), ("off_y", C.c_int32),
                ("mask_w", C.c_int32), ("mask_h", C.c_int32), ("mask_stride", C.c_int32),
                ("mask", C.c_void_p)]


ENCODE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int,
                        C.POINTER(C.c_void_p), C.POINTER(C.c_size_t))
RELEASE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p)
SAVE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t, C.c_char_p)
ADVANCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint32, C.c_double, C.POINTER(C.c_int32))
MASK_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint32, C.c_double, C.c_int, C.c_int, C.POINTER(IphGlyph))
KERN_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_double)
INDEX_FN = C.CFUNCTYPE(C.c_uint32, C.c_void_p, C.c_uint32)


class Callbacks(C.Structure):
    _fields_ = [("user", C.c_void_p), ("encode", ENCODE_FN), ("release", RELEASE_FN), ("save_processed", SAVE_FN),
                ("glyph_advance", ADVANCE_FN), ("glyph_mask", MASK_FN), ("kern", KERN_FN), ("glyph_index", INDEX_FN)]


def _host_lib():
    lib = L.load()
    if getattr(lib, "_iph_typed", False):
        return lib
    vp = C.c_void_p
    lib.iph_processor_new.argtypes = [vp, C.POINTER(Callbacks)]
    lib.iph_processor_new.restype = vp
    lib.iph_processor_free.argtypes = [vp]
    lib.iph_processor_free.restype = None
    lib.iph_set_device_jpeg.argtypes = [vp, C.c_int]
    lib.iph_process.argtypes = [vp, C.c_char_p, C.POINTER(L.ImageDesc), C.c_char_p, C.c_char_p, C.POINTER(vp)]
    lib.iph_process_batch.argtypes = [vp, C.c_int, C.POINTER(C.c_char_p), C.POINTER(L.ImageDesc), C.POINTER(C.c_char_p),
                                      C.POINTER(vp), C.POINTER(C.c_int), C.POINTER(vp)]
    lib.iph_free.argtypes = [vp]
    lib.iph_free.restype = None
    lib.iph_last_error.restype = C.c_char_p
    lib.iph_parse_color.argtypes = [C.c_char_p, C.c_double, C.POINTER(C.c_uint8)]
    lib.iph_watermark_height_px.argtypes = [C.c_double]
    lib.iph_watermark_anchor.argtypes = [C.c_char_p] + [C.c_int] * 4 + [C.POINTER(C.c_int)] * 2
    lib.iph_watermark_anchor.restype = None
    lib.iph_generate_path.argtypes = [C.c_char_p] * 4
    lib.iph_generate_path.restype = vp
    lib.iph_content_type.argtypes = [C.c_char_p]
    lib.iph_content_type.restype = C.c_char_p
    lib._iph_typed = True
    return lib


# ---- pure host helpers --------------------------------------------------------------
def parse_color(s: str, opacity: float) -> Tuple[int, Tuple[int, int, int, int]]:
    out = (C.c_uint8 * 4)()
    rc = _host_lib().iph_parse_color(s.encode(), opacity, out)
    return rc, tuple(out)


def watermark_height_px(font_size: float) -> int:
    return _host_lib().iph_watermark_height_px(font_size)


def watermark_anchor(position: str, W: int, H: int, width_px: int, height_px: int) -> Tuple[int, int]:
    x, y = C.c_int(), C.c_int()
    _host_lib().iph_watermark_anchor(position.encode(), W, H, width_px, height_px, x, y)
    return x.value, y.value


def generate_path(image_id: str, operation: str, fmt: str, params: dict) -> str:
    lib = _host_lib()
    p = lib.iph_generate_path(image_id.encode(), operation.encode(), fmt.encode(), json.dumps(params).encode())
    s = C.string_at(p).decode()
    lib.iph_free(p)
    return s


def content_type(path: str) -> str:
    return _host_lib().iph_content_type(path.encode()).decode()


# ---- stand-ins for what the Go host keeps ---------------------------------------------
class MemoryFileRepo:
    """fileRepository.SaveProcessed (processor/contract.go:8-14) into a dict."""

    def __init__(self):
        self.objects: Dict[str, Tuple[bytes, str]] = {}
        self.fail_on: Optional[str] = None

    def save_processed(self, path: str, data: bytes, content_type: str) -> int:
        if self.fail_on and self.fail_on in path:
            return -1
        self.objects[path] = (data, content_type)
        return 0


def pil_encode(rgba: np.ndarray, fmt: str, quality: int) -> bytes:
    """jpeg.Encode(q) / png.Encode / gif.Encode stand-in (PIL; not byte-identical to Go's codecs)."""
    from PIL import Image as PI
    im = PI.fromarray(rgba, "RGBA")
    buf = io.BytesIO()
    if fmt == "jpeg":
        im.convert("RGB").save(buf, "JPEG", quality=quality)
    elif fmt == "png":
        im.save(buf, "PNG")
    else:
        im.convert("P").save(buf, "GIF")
    return buf.getvalue()


def raw_encode(rgba: np.ndarray, fmt: str, quality: int) -> bytes:
    """Lossless container for parity tests: header + raw RGBA, so the stored object can be
    compared byte for byte with the oracle."""
    h, w = rgba.shape[:2]
    return b"RAW0" + fmt.encode().ljust(8, b"\0") + np.array([w, h], np.int32).tobytes() + rgba.tobytes()


def raw_decode(data: bytes) -> Tuple[str, np.ndarray]:
    assert data[:4] == b"RAW0"
    fmt = data[4:12].rstrip(b"\0").decode()
    w, h = np.frombuffer(data[12:20], np.int32)
    return fmt, np.frombuffer(data[20:], np.uint8).reshape(int(h), int(w), 4)


class PilFace:
    """truetype.Face + freetype rasteriser stand-in (see glyphs.py for what that means)."""

    def __init__(self):
        self._fonts = {}

    def font(self, size: float):
        from PIL import ImageFont
        if size not in self._fonts:
            self._fonts[size] = ImageFont.load_default(size=size)
        return self._fonts[size]

    def advance_26_6(self, rune: int, size: float) -> Optional[int]:
        return int(round(self.font(size).getlength(chr(rune)) * 64))

    def index(self, rune: int) -> int:
        """Font.Index(rune) stand-in: PIL does not expose the cmap, so the rune is its own index."""
        return rune

    def mask(self, rune: int, size: float, fx: int, fy: int):
        """(advance_26_6, off_x, off_y, mask) with offsets relative to the integer pen, y down."""
        f = self.font(size)
        adv = int(round(f.getlength(chr(rune)) * 64))
        m, off = f.getmask2(chr(rune), mode="L", anchor="ls")
        w, h = m.size
        if w == 0 or h == 0:
            return adv, 0, 0, np.zeros((0, 0), np.uint8)
        return adv, int(off[0]), int(off[1]), np.frombuffer(bytes(m), np.uint8).reshape(h, w).copy()


class ImageProcessor:
    """processor.ImageProcessor (image_processor.go:21-37): Process(task, decoded image)."""

    def __init__(self, engine: Optional[Engine], file_repo: MemoryFileRepo, encode=pil_encode, face: Optional[PilFace] = None,
                 device_jpeg: bool = False):
        """device_jpeg: JPEG-bound results are encoded on the device exactly as Go's jpeg.Encode(q85) would (opt-in,
        iph_set_device_jpeg); `encode` then only sees PNG / GIF targets."""
        self._lib = _host_lib()
        self.file_repo = file_repo
        self.encode = encode
        self.face = face or PilFace()
        self._live = {}
        self._tls = threading.local()   # Process is re-entrant (WORKER_CONCURRENCY threads): the mask handed to C lives per thread

        def _encode(user, rgba, w, h, stride, fmt, quality, out, out_len):
            try:
                if w <= 0 or h <= 0:
                    arr = np.zeros((max(h, 0), max(w, 0), 4), np.uint8)
                else:
                    arr = np.ctypeslib.as_array((C.c_uint8 * (stride * h)).from_address(rgba)).reshape(h, stride)[:, :w * 4].reshape(h, w, 4)
                data = self.encode(arr, fmt.decode(), quality)
                buf = C.create_string_buffer(data, len(data))
                addr = C.addressof(buf)
                self._live[addr] = buf
                out[0] = addr
                out_len[0] = len(data)
                return 0
            except Exception:
                return -1

        def _release(user, buf):
            self._live.pop(buf, None)

        def _save(user, path, data, size, ctype):
            return self.file_repo.save_processed(path.decode(), C.string_at(data, size), ctype.decode())

        def _advance(user, rune, size, out):
            a = self.face.advance_26_6(rune, size)
            if a is None:
                return 1
            out[0] = a
            return 0

        def _mask(user, rune, size, fx, fy, out):
            try:
                adv, ox, oy, m = self.face.mask(rune, size, fx, fy)
                keep = self._tls.mask_keep = np.ascontiguousarray(m)
                g = out[0]
                g.advance_26_6, g.off_x, g.off_y = adv, ox, oy
                g.mask_h, g.mask_w = (m.shape if m.size else (0, 0))
                g.mask_stride = m.shape[1] if m.size else 0
                g.mask = keep.ctypes.data if m.size else None
                return 0
            except Exception:
                return -1

        def _index(user, rune):
            return self.face.index(rune)

        self._cbs = Callbacks(None, ENCODE_FN(_encode), RELEASE_FN(_release), SAVE_FN(_save), ADVANCE_FN(_advance),
                              MASK_FN(_mask), KERN_FN(0), INDEX_FN(_index) if hasattr(self.face, "index") else INDEX_FN(0))
        self._engine = engine
        self._p = self._lib.iph_processor_new(engine._ctx if engine else None, C.byref(self._cbs))
        if not self._p:
            raise MemoryError("iph_processor_new failed")
        if device_jpeg:
            self._lib.iph_set_device_jpeg(self._p, 1)

    def close(self):
        if self._p:
            self._lib.iph_processor_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def process(self, task: dict | str, image: Optional[Image], decoded_format: str = "jpeg",
                decode_error: Optional[str] = None) -> Tuple[dict, Optional[str]]:
        """Returns (ProcessingResult as dict, error string or None) -- the (result, err) pair of Process."""
        tj = task if isinstance(task, str) else json.dumps(task)
        desc = image.desc() if image is not None else None
        out = C.c_void_p()
        rc = self._lib.iph_process(self._p, tj.encode(), C.byref(desc) if desc is not None else None,
                                   decoded_format.encode(), decode_error.encode() if decode_error else None, C.byref(out))
        res = json.loads(C.string_at(out.value).decode()) if out.value else None
        raw = C.string_at(out.value).decode() if out.value else None
        if out.value:
            self._lib.iph_free(out)
        self.last_result_json = raw
        return res, (self._lib.iph_last_error().decode() if rc else None)

    def process_batch(self, tasks: Sequence[dict | str], images: Sequence[Image], decoded_formats: Sequence[str]):
        """The batching processWorker: [(result dict, error or None)] per message."""
        n = len(tasks)
        tj = (C.c_char_p * n)(*[(t if isinstance(t, str) else json.dumps(t)).encode() for t in tasks])
        descs = (L.ImageDesc * n)(*[im.desc() for im in images])
        fmts = (C.c_char_p * n)(*[f.encode() for f in decoded_formats])
        outs = (C.c_void_p * n)()
        rcs = (C.c_int * n)()
        errs = (C.c_void_p * n)()
        self._lib.iph_process_batch(self._p, n, tj, descs, fmts, outs, rcs, errs)
        res = []
        for i in range(n):
            r = json.loads(C.string_at(outs[i]).decode()) if outs[i] else None
            e = C.string_at(errs[i]).decode() if errs[i] else None
            if outs[i]:
                self._lib.iph_free(outs[i])
            if errs[i]:
                self._lib.iph_free(errs[i])
            res.append((r, e))
        return res
