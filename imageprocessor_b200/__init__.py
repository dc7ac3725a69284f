"""imageprocessor_b200 -- B200-native (sm_100a) raster engine for the hot path of
sj-shoff/ImageProcessor's worker: resize -> 1024x768 keep-aspect, 200x200 thumbnail,
text-watermark alpha-over, behind the reference's own processor.Process call.

The compute lives in libipgpu.so (hand-written CUDA, C ABI in include/ipgpu.h);
this package is the Python-side harness over it.  No CPU fallback exists.
"""
from . import _lib
from ._lib import (IpgError, RGBA8, NRGBA8, GRAY8, YCBCR444, YCBCR422, YCBCR420, YCBCR440, RGBA64, NRGBA64, GRAY16, JPEG,
                   PRECISION_EXACT, PRECISION_FAST, PRECISION_REFERENCE,
                   OP_RESIZE, OP_THUMB_CROP, OP_WATERMARK, OPF_WATERMARK_PATCH_ONLY, MEM_HOST, MEM_DEVICE)
from .engine import (Engine, Image, OpSpec, GlyphMask, Ticket, PinnedBuffer, JpegResult,
                     keep_aspect_dims, thumb_fit_dims, crop_square)

__all__ = [n for n in dir() if not n.startswith("_")]
