"""The drop-in boundary: libipgpu.so loads on a CPU-only box and exports every symbol the
headers under include/ declare (no compute is called here)."""
import ctypes as C
import glob
import os
import re

import imageprocessor_b200 as ip
from imageprocessor_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = []
    for h in sorted(glob.glob(os.path.join(ROOT, "include", "*.h"))):
        text = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names += re.findall(r"IPG_API\s+[^;(]*?\b(ip[gh]_\w+)\s*\(", text)
    return names


def test_library_exports_every_declared_symbol():
    lib = L.load()
    names = declared_symbols()
    assert len(names) >= 30 and len(set(names)) == len(names)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"libipgpu.so lacks {missing}"
    assert sorted(names) == sorted(L.EXPORTS + L.HOST_EXPORTS), "python binding list out of sync with the headers"
    assert lib.ipg_abi_version() == 3


def test_no_cpu_fallback_without_a_device():
    """Without a CUDA device ipg_init must fail loudly (IPG_ERR_NO_DEVICE); on a GPU box it succeeds."""
    lib = L.load()
    ctx = C.c_void_p()
    rc = lib.ipg_init(None, 0, None, C.byref(ctx))
    if rc == 0:
        lib.ipg_destroy(ctx)
    else:
        assert rc in (L.ERR_NO_DEVICE, L.ERR_CUDA) and not ctx.value
        assert b"no CPU fallback" in lib.ipg_last_error() or b"CUDA" in lib.ipg_last_error()


def test_struct_layouts_match_the_header():
    # sizes the C side was compiled with (x86-64 SysV): guards the ctypes mirrors in _lib.py
    assert C.sizeof(L.Config) == 40 and C.sizeof(L.ImageDesc) == 56 and C.sizeof(L.Glyph) == 48
    assert C.sizeof(L.Op) == 112 and C.sizeof(L.Stats) == 144


def test_geometry_helpers_follow_the_reference():
    assert ip.keep_aspect_dims(4000, 3000, 1024, 768) == (1024, 768)
    assert ip.keep_aspect_dims(1002, 751, 1024, 768) == (1023, 767)
    assert ip.thumb_fit_dims(4000, 3000, 200) == (266, 200)
    assert ip.crop_square(7680, 4320) == (1680, 0, 4320)
