"""ctypes binding of libipgpu.so (C ABI: include/ipgpu.h).

The library is the product; there is no Python or CPU implementation behind this
module.  Importing works anywhere (so tooling can inspect the ABI), but creating an
engine without the built library or without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IPG_LIB_PATH") or os.path.join(_HERE, "libipgpu.so")  # env: kernel experiments only

# ipg_status
OK, ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_TIMEOUT, ERR_NO_DEVICE, ERR_SHUTDOWN, ERR_INTERNAL = 0, -1, -2, -3, -4, -5, -6, -7
# ipg_layout
RGBA8, NRGBA8, GRAY8, YCBCR444, YCBCR422, YCBCR420, YCBCR440, RGBA64, NRGBA64, GRAY16 = range(10)
JPEG = 16  # destination only: the result as the JPEG file Go's jpeg.Encode would write, encoded on the device
# ipg_memspace
MEM_HOST, MEM_DEVICE = 0, 1
# ipg_precision
PRECISION_EXACT, PRECISION_FAST, PRECISION_REFERENCE = 0, 1, 2
# ipg_op_kind
OP_RESIZE, OP_THUMB_CROP, OP_WATERMARK = 1, 2, 3
# ipg_op_flags
OPF_WATERMARK_PATCH_ONLY = 1


class IpgError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"ipgpu error {code}: {message}")
        self.code = code
        self.message = message


class Config(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("precision", C.c_int32), ("lanes_per_device", C.c_int32),
                ("max_batch", C.c_int32), ("batch_window_us", C.c_int32), ("fuse_targets", C.c_int32),
                ("lane_device_bytes", C.c_uint64), ("lane_pinned_bytes", C.c_uint64)]


class ImageDesc(C.Structure):
    _fields_ = [("layout", C.c_int32), ("memspace", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
                ("plane", C.c_void_p * 3), ("stride", C.c_int32 * 3), ("opaque_hint", C.c_int32)]


class Glyph(C.Structure):
    _fields_ = [("x0", C.c_int32), ("y0", C.c_int32), ("x1", C.c_int32), ("y1", C.c_int32),
                ("mp_x", C.c_int32), ("mp_y", C.c_int32), ("mask_w", C.c_int32), ("mask_h", C.c_int32),
                ("mask_stride", C.c_int32), ("reserved0", C.c_int32), ("mask", C.c_void_p)]


class Op(C.Structure):
    _fields_ = [("kind", C.c_int32), ("dst_w", C.c_int32), ("dst_h", C.c_int32),
                ("rect_x", C.c_int32), ("rect_y", C.c_int32), ("rect_w", C.c_int32), ("rect_h", C.c_int32),
                ("color", C.c_uint8 * 4), ("n_glyphs", C.c_int32), ("glyphs", C.POINTER(Glyph)),
                ("dst", C.c_void_p), ("dst_stride", C.c_int32), ("dst_memspace", C.c_int32), ("flags", C.c_int32),
                ("dst_layout", C.c_int32), ("dst_cb", C.c_void_p), ("dst_cr", C.c_void_p), ("dst_cstride", C.c_int32),
                ("jpeg_quality", C.c_int32), ("dst_capacity", C.c_uint64), ("dst_len", C.POINTER(C.c_uint64))]


class Stats(C.Structure):
    _fields_ = [("tickets_done", C.c_uint64), ("batches", C.c_uint64), ("kernels_launched", C.c_uint64),
                ("bytes_h2d", C.c_uint64), ("bytes_d2h", C.c_uint64), ("exact_fixups", C.c_uint64),
                ("exact_fallbacks", C.c_uint64), ("staged_copies", C.c_uint64),
                ("kernel_ms", C.c_double), ("stream_kernel_ms", C.c_double),
                ("fix_kernel_ms", C.c_double), ("other_kernel_ms", C.c_double),
                ("kernel_span_ms", C.c_double), ("batch_span_ms", C.c_double),
                ("stream_fast_kernel_ms", C.c_double), ("fast_jobs", C.c_uint64),
                ("h2d_ms", C.c_double), ("d2h_ms", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# every symbol include/ipgpu.h declares
EXPORTS = [
    "ipg_init", "ipg_destroy", "ipg_device_count", "ipg_last_error", "ipg_abi_version",
    "ipg_alloc_pinned", "ipg_free_pinned", "ipg_alloc_device", "ipg_free_device",
    "ipg_copy_to_device", "ipg_copy_from_device", "ipg_submit", "ipg_submit_on", "ipg_wait",
    "ipg_flush", "ipg_get_stats", "ipg_reset_stats", "ipg_keep_aspect_dims", "ipg_thumb_fit_dims", "ipg_crop_square",
]
# ... and include/ipgpu_host.h
HOST_EXPORTS = [
    "iph_processor_new", "iph_processor_free", "iph_set_device_jpeg", "iph_process", "iph_process_batch", "iph_free", "iph_last_error",
    "iph_parse_color", "iph_watermark_height_px", "iph_watermark_anchor", "iph_generate_path", "iph_content_type",
]

_lib = None


def load():
    """dlopen libipgpu.so and type its entry points. Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C imageprocessor_b200/csrc` "
            "(or __graft_entry__.build()). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, ip = C.c_void_p, C.POINTER(C.c_int)
    L.ipg_init.argtypes = [ip, C.c_int, C.POINTER(Config), C.POINTER(vp)]
    L.ipg_init.restype = C.c_int
    L.ipg_destroy.argtypes = [vp]
    L.ipg_destroy.restype = None
    L.ipg_device_count.argtypes = [vp]
    L.ipg_last_error.restype = C.c_char_p
    L.ipg_abi_version.restype = C.c_int
    L.ipg_alloc_pinned.argtypes = [vp, C.c_size_t]
    L.ipg_alloc_pinned.restype = vp
    L.ipg_free_pinned.argtypes = [vp, vp]
    L.ipg_free_pinned.restype = None
    L.ipg_alloc_device.argtypes = [vp, C.c_int, C.c_size_t]
    L.ipg_alloc_device.restype = vp
    L.ipg_free_device.argtypes = [vp, C.c_int, vp]
    L.ipg_free_device.restype = None
    L.ipg_copy_to_device.argtypes = [vp, C.c_int, vp, vp, C.c_size_t]
    L.ipg_copy_from_device.argtypes = [vp, C.c_int, vp, vp, C.c_size_t]
    L.ipg_submit.argtypes = [vp, C.POINTER(ImageDesc), C.POINTER(Op), C.c_int, C.POINTER(C.c_uint64)]
    L.ipg_submit_on.argtypes = [vp, C.c_int, C.POINTER(ImageDesc), C.POINTER(Op), C.c_int, C.POINTER(C.c_uint64)]
    L.ipg_wait.argtypes = [vp, C.c_uint64, C.c_int]
    L.ipg_flush.argtypes = [vp]
    L.ipg_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.ipg_reset_stats.argtypes = [vp]
    L.ipg_keep_aspect_dims.argtypes = [C.c_int] * 4 + [ip, ip]
    L.ipg_keep_aspect_dims.restype = None
    L.ipg_thumb_fit_dims.argtypes = [C.c_int] * 3 + [ip, ip]
    L.ipg_thumb_fit_dims.restype = None
    L.ipg_crop_square.argtypes = [C.c_int] * 2 + [ip, ip, ip]
    L.ipg_crop_square.restype = None
    _lib = L
    return L


def last_error() -> str:
    return (load().ipg_last_error() or b"").decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != OK:
        raise IpgError(rc, last_error())
