"""Thin object wrapper over the C ABI: contexts, pinned buffers, submit/wait.

This is plumbing for the Python harness (tests, bench, the worker mirror); the
compute is entirely in libipgpu.so.  A Go host binds the same entry points via
cgo.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib as L


@dataclass
class Image:
    """A decoded raster as Go's image.Decode hands it to the ops
    (internal/usecase/processor/image_processor.go:47)."""
    layout: int
    width: int
    height: int
    planes: Tuple[np.ndarray, ...]   # uint8, 2-D (rows x stride bytes), C-contiguous rows
    opaque_hint: bool = False
    memspace: int = L.MEM_HOST
    device_ptrs: Tuple[int, ...] = ()   # when memspace == MEM_DEVICE
    device_strides: Tuple[int, ...] = ()

    @staticmethod
    def from_rgba(a: np.ndarray, layout: int = L.RGBA8, opaque_hint: bool = False) -> "Image":
        assert a.dtype == np.uint8 and a.ndim == 3 and a.shape[2] == 4
        if a.strides[2] != 1 or a.strides[1] != 4:
            a = np.ascontiguousarray(a)
        return Image(layout, a.shape[1], a.shape[0], (a,), opaque_hint)

    @staticmethod
    def from_gray(a: np.ndarray) -> "Image":
        assert a.dtype == np.uint8 and a.ndim == 2
        if a.strides[1] != 1:
            a = np.ascontiguousarray(a)
        return Image(L.GRAY8, a.shape[1], a.shape[0], (a,), True)

    @staticmethod
    def from_deep(a: np.ndarray, layout: int) -> "Image":
        """*image.RGBA64 / *image.NRGBA64 from (h, w, 4) uint16 values or *image.Gray16 from (h, w): stored big-endian,
        8 (2) bytes per pixel, as Go's Pix holds them."""
        assert a.dtype == np.uint16 and ((a.ndim == 3 and a.shape[2] == 4) or a.ndim == 2)
        be = np.ascontiguousarray(a.astype(">u2"))
        return Image(layout, a.shape[1], a.shape[0], (be.view(np.uint8).reshape(a.shape[0], -1),), layout == L.GRAY16)

    @staticmethod
    def from_ycbcr(y: np.ndarray, cb: np.ndarray, cr: np.ndarray, layout: int) -> "Image":
        planes = tuple(p if p.strides[1] == 1 else np.ascontiguousarray(p) for p in (y, cb, cr))
        return Image(layout, y.shape[1], y.shape[0], planes, True)

    @staticmethod
    def on_device(layout: int, width: int, height: int, ptrs: Sequence[int], strides: Sequence[int],
                  opaque_hint: bool = False) -> "Image":
        return Image(layout, width, height, (), opaque_hint, L.MEM_DEVICE, tuple(ptrs), tuple(strides))

    def desc(self) -> L.ImageDesc:
        d = L.ImageDesc()
        d.layout, d.memspace, d.width, d.height = self.layout, self.memspace, self.width, self.height
        d.opaque_hint = int(self.opaque_hint)
        if self.memspace == L.MEM_DEVICE:
            for k, (p, s) in enumerate(zip(self.device_ptrs, self.device_strides)):
                d.plane[k] = p
                d.stride[k] = s
        else:
            for k, p in enumerate(self.planes):
                d.plane[k] = p.ctypes.data
                d.stride[k] = p.strides[0]
        return d


@dataclass
class GlyphMask:
    """One DrawMask call of freetype's DrawString: destination rect (clipped to
    the image), alpha mask, mask point (watermark.go:151)."""
    x0: int
    y0: int
    x1: int
    y1: int
    mask: np.ndarray
    mp_x: int = 0
    mp_y: int = 0


@dataclass
class OpSpec:
    kind: int
    dst_w: int
    dst_h: int
    rect: Tuple[int, int, int, int] = (0, 0, 0, 0)
    color: Tuple[int, int, int, int] = (0, 0, 0, 0)
    glyphs: Sequence[GlyphMask] = ()
    dst: Optional[np.ndarray] = None      # host destination (h, w, 4) uint8; allocated if None
    dst_device: Optional[Tuple[int, int]] = None  # (device pointer, stride) instead of dst
    flags: int = 0                        # ipg_op_flags (OPF_WATERMARK_PATCH_ONLY: dst must already hold the source pixels)
    # planar YCbCr 4:2:0 result (ipg_op.dst_layout, for results that will be JPEG-encoded): three PinnedBuffer-backed
    # uint8 planes (Y h x w, Cb and Cr (h+1)//2 x (w+1)//2); the RGBA destination fields are then unused
    dst_ycbcr420: Optional[Tuple[np.ndarray, np.ndarray, np.ndarray]] = None
    # the result as a JPEG file encoded on the device (ipg_op.dst_layout = IPG_LAYOUT_JPEG): quality as jpeg.Options.Quality
    # (the reference uses 85).  The output is then a `JpegResult`; `jpeg_buffer` is an optional PinnedBuffer-backed uint8
    # array to receive the file (its last 8 bytes: the length), allocated when None; `jpeg_capacity` sizes that allocation
    # (default w * h + 64 KiB)
    jpeg_quality: Optional[int] = None
    jpeg_buffer: Optional[np.ndarray] = None
    jpeg_capacity: Optional[int] = None

    @staticmethod
    def resize(dw: int, dh: int, **kw) -> "OpSpec":
        return OpSpec(L.OP_RESIZE, dw, dh, **kw)

    @staticmethod
    def thumb_crop(rect: Tuple[int, int, int, int], size: int, **kw) -> "OpSpec":
        return OpSpec(L.OP_THUMB_CROP, size, size, rect=rect, **kw)

    @staticmethod
    def watermark(w: int, h: int, color, glyphs: Sequence[GlyphMask], **kw) -> "OpSpec":
        return OpSpec(L.OP_WATERMARK, w, h, color=tuple(color), glyphs=glyphs, **kw)


class JpegResult:
    """A result encoded on the device: `data` is the file once the ticket was waited for.  The engine writes the file and
    its length after ipg_submit returned, so both live in ipg_alloc_pinned memory: `buffer` is a uint8 view of a
    PinnedBuffer whose last 8 bytes (8-byte aligned) hold the length."""

    def __init__(self, buffer: np.ndarray, owner=None):
        assert buffer.dtype == np.uint8 and buffer.ndim == 1 and buffer.size >= 1032 and buffer.ctypes.data % 8 == 0
        self.capacity = (buffer.size - 8) & ~7
        self.buffer = buffer
        self._len = buffer[self.capacity:self.capacity + 8].view(np.uint64)
        self._len[0] = 0
        self._owner = owner   # a PinnedBuffer this result allocated itself (released by detach())
        self._bytes = None

    @property
    def nbytes(self) -> int:
        return len(self._bytes) if self._bytes is not None else int(self._len[0])

    @property
    def data(self) -> bytes:
        return self._bytes if self._bytes is not None else self.buffer[:self.nbytes].tobytes()

    def detach(self):
        """A result that allocated its own pinned buffer copies the file out and frees the buffer (Engine.wait calls
        this): the bytes then outlive the engine."""
        if self._owner is not None:
            self._bytes = self.buffer[:int(self._len[0])].tobytes()
            self.buffer = self._len = None
            self._owner.free()
            self._owner = None


@dataclass
class Ticket:
    id: int
    outputs: List[Optional[np.ndarray]]
    _keep: list = field(default_factory=list, repr=False)


class PinnedBuffer:
    """Host memory from ipg_alloc_pinned, exposed as a numpy uint8 array."""

    def __init__(self, engine: "Engine", nbytes: int):
        self._engine = engine
        self.ptr = L.load().ipg_alloc_pinned(engine._ctx, nbytes)
        if not self.ptr:
            raise L.IpgError(L.ERR_NOMEM, L.last_error())
        self.nbytes = nbytes
        self.array = np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(self.ptr))

    def free(self):
        if self.ptr and self._engine._ctx:
            L.load().ipg_free_pinned(self._engine._ctx, self.ptr)
        self.ptr = None
        self.array = None


class Engine:
    """Owns an ipg_ctx.  Raises if libipgpu.so is missing or no CUDA device exists."""

    def __init__(self, devices: Optional[Sequence[int]] = None, precision: int = L.PRECISION_EXACT,
                 lanes_per_device: int = 3, max_batch: int = 16, batch_window_us: int = 200,
                 lane_device_bytes: int = 0, lane_pinned_bytes: int = 0, fuse_targets: int = 0):
        lib = L.load()
        cfg = L.Config()
        cfg.struct_size = C.sizeof(L.Config)
        cfg.precision = precision
        cfg.lanes_per_device = lanes_per_device
        cfg.max_batch = max_batch
        cfg.batch_window_us = batch_window_us
        cfg.fuse_targets = fuse_targets
        cfg.lane_device_bytes = lane_device_bytes
        cfg.lane_pinned_bytes = lane_pinned_bytes
        ctx = C.c_void_p()
        if devices:
            arr = (C.c_int * len(devices))(*devices)
            rc = lib.ipg_init(arr, len(devices), C.byref(cfg), C.byref(ctx))
        else:
            rc = lib.ipg_init(None, 0, C.byref(cfg), C.byref(ctx))
        self._ctx = None
        L.check(rc)
        self._ctx = ctx
        self._lib = lib

    # -- lifecycle -------------------------------------------------------------
    def close(self):
        if self._ctx:
            self._lib.ipg_destroy(self._ctx)
            self._ctx = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def device_count(self) -> int:
        return self._lib.ipg_device_count(self._ctx)

    # -- memory ----------------------------------------------------------------
    def alloc_pinned(self, nbytes: int) -> PinnedBuffer:
        return PinnedBuffer(self, nbytes)

    def alloc_device(self, device: int, nbytes: int) -> int:
        p = self._lib.ipg_alloc_device(self._ctx, device, nbytes)
        if not p:
            raise L.IpgError(L.ERR_NOMEM, L.last_error())
        return p

    def free_device(self, device: int, ptr: int) -> None:
        self._lib.ipg_free_device(self._ctx, device, ptr)

    def to_device(self, device: int, dst: int, src: np.ndarray) -> None:
        src = np.ascontiguousarray(src)
        L.check(self._lib.ipg_copy_to_device(self._ctx, device, dst, src.ctypes.data, src.nbytes))

    def from_device(self, device: int, dst: np.ndarray, src: int) -> None:
        assert dst.flags["C_CONTIGUOUS"]
        L.check(self._lib.ipg_copy_from_device(self._ctx, device, dst.ctypes.data, src, dst.nbytes))

    # -- hot path --------------------------------------------------------------
    def submit(self, image: Image, ops: Sequence[OpSpec], device: Optional[int] = None) -> Ticket:
        n = len(ops)
        arr = (L.Op * max(n, 1))()
        keep: list = [image]
        outs: List[Optional[np.ndarray]] = []
        for k, o in enumerate(ops):
            c = arr[k]
            c.kind, c.dst_w, c.dst_h, c.flags = o.kind, o.dst_w, o.dst_h, o.flags
            c.rect_x, c.rect_y, c.rect_w, c.rect_h = o.rect
            for j in range(4):
                c.color[j] = o.color[j]
            if o.glyphs:
                ga = (L.Glyph * len(o.glyphs))()
                for j, g in enumerate(o.glyphs):
                    m = g.mask if (g.mask.dtype == np.uint8 and g.mask.strides[1] == 1) else np.ascontiguousarray(g.mask, np.uint8)
                    keep.append(m)
                    ga[j].x0, ga[j].y0, ga[j].x1, ga[j].y1 = g.x0, g.y0, g.x1, g.y1
                    ga[j].mp_x, ga[j].mp_y = g.mp_x, g.mp_y
                    ga[j].mask_w, ga[j].mask_h, ga[j].mask_stride = m.shape[1], m.shape[0], m.strides[0]
                    ga[j].mask = m.ctypes.data
                keep.append(ga)
                c.n_glyphs = len(o.glyphs)
                c.glyphs = ga
            if o.jpeg_quality is not None:
                if o.jpeg_buffer is None:   # convenience for tests: one pinned allocation per result (a worker reuses its buffers)
                    own = self.alloc_pinned((o.jpeg_capacity or (o.dst_w * o.dst_h + (64 << 10))) + 16)
                    res = JpegResult(own.array, own)
                else:
                    res = JpegResult(o.jpeg_buffer)
                c.dst_layout = L.JPEG
                c.dst, c.dst_capacity, c.jpeg_quality = res.buffer.ctypes.data, res.capacity, o.jpeg_quality
                c.dst_len = C.cast(res.buffer.ctypes.data + res.capacity, C.POINTER(C.c_uint64))
                c.dst_memspace = L.MEM_HOST
                keep.append(res)
                outs.append(res)
            elif o.dst_ycbcr420 is not None:
                yp, cbp, crp = o.dst_ycbcr420
                assert yp.shape == (o.dst_h, o.dst_w) and cbp.shape == crp.shape == ((o.dst_h + 1) // 2, (o.dst_w + 1) // 2)
                c.dst_layout = L.YCBCR420
                c.dst, c.dst_stride = yp.ctypes.data, yp.strides[0]
                c.dst_cb, c.dst_cr, c.dst_cstride = cbp.ctypes.data, crp.ctypes.data, cbp.strides[0]
                assert crp.strides[0] == cbp.strides[0]
                c.dst_memspace = L.MEM_HOST
                outs.append(o.dst_ycbcr420)
            elif o.dst_device is not None:
                c.dst, c.dst_stride = o.dst_device
                c.dst_memspace = L.MEM_DEVICE
                outs.append(None)
            else:
                dst = o.dst
                if dst is None:
                    dst = np.empty((max(o.dst_h, 0), max(o.dst_w, 0), 4), np.uint8)
                assert dst.dtype == np.uint8 and dst.shape[:2] == (max(o.dst_h, 0), max(o.dst_w, 0))
                c.dst = dst.ctypes.data if dst.size else None
                c.dst_stride = dst.strides[0] if dst.size else 0
                c.dst_memspace = L.MEM_HOST
                outs.append(dst)
        d = image.desc()
        tid = C.c_uint64()
        if device is None:
            rc = self._lib.ipg_submit(self._ctx, C.byref(d), arr, n, C.byref(tid))
        else:
            rc = self._lib.ipg_submit_on(self._ctx, device, C.byref(d), arr, n, C.byref(tid))
        L.check(rc)
        keep.append(arr)
        return Ticket(tid.value, outs, keep)

    def wait(self, ticket: Ticket, timeout_ms: int = -1) -> List[Optional[np.ndarray]]:
        rc = self._lib.ipg_wait(self._ctx, ticket.id, timeout_ms)
        if rc != L.ERR_TIMEOUT:
            for o in ticket.outputs:
                if isinstance(o, JpegResult):
                    o.detach()
        L.check(rc)
        ticket._keep.clear()
        return ticket.outputs

    def run(self, image: Image, ops: Sequence[OpSpec], device: Optional[int] = None):
        return self.wait(self.submit(image, ops, device))

    def flush(self) -> None:
        L.check(self._lib.ipg_flush(self._ctx))

    def reset_stats(self) -> None:
        L.check(self._lib.ipg_reset_stats(self._ctx))

    def stats(self) -> dict:
        s = L.Stats()
        L.check(self._lib.ipg_get_stats(self._ctx, C.byref(s)))
        return s.as_dict()


# -- reference geometry (double/int arithmetic lives in the C library) ----------
def keep_aspect_dims(ow: int, oh: int, w: int, h: int) -> Tuple[int, int]:
    a, b = C.c_int(), C.c_int()
    L.load().ipg_keep_aspect_dims(ow, oh, w, h, a, b)
    return a.value, b.value


def thumb_fit_dims(ow: int, oh: int, size: int) -> Tuple[int, int]:
    a, b = C.c_int(), C.c_int()
    L.load().ipg_thumb_fit_dims(ow, oh, size, a, b)
    return a.value, b.value


def crop_square(ow: int, oh: int) -> Tuple[int, int, int]:
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    L.load().ipg_crop_square(ow, oh, a, b, c)
    return a.value, b.value, c.value
