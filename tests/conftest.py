import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_device_count():
    """Devices ipg_init can open; 0 only when the library itself answers IPG_ERR_NO_DEVICE.  None when the library
    cannot be loaded or fails some other way: then nothing is skipped and the tests fail loudly."""
    import ctypes as C
    try:
        from imageprocessor_b200 import _lib as L
        lib = L.load()
    except Exception:
        return None
    ctx = C.c_void_p()
    rc = lib.ipg_init(None, 0, None, C.byref(ctx))
    if rc == L.ERR_NO_DEVICE:
        return 0
    if rc != 0:
        return None
    n = lib.ipg_device_count(ctx)
    lib.ipg_destroy(ctx)
    return n


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a box without a CUDA device skips the gpu-marked tests instead of failing them.
    (On the GPU box nothing is skipped: there the product path must run, and it fails loudly if libipgpu.so is missing.)"""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items:
        return
    n = _cuda_device_count()
    if n is None:
        return
    for it in gpu_items:
        need = it.get_closest_marker("gpu").kwargs.get("min_devices", 1)
        if n < need:
            it.add_marker(pytest.mark.skip(reason=f"needs {need} CUDA device(s), ipg_init sees {n}"))


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def engines():
    """One engine per precision mode, shared by the GPU tests."""
    import imageprocessor_b200 as ip
    made = {}

    def get(precision=ip.PRECISION_EXACT, **kw):
        key = (precision, tuple(sorted(kw.items())))
        if key not in made:
            made[key] = ip.Engine(devices=[0], precision=precision, **kw)
        return made[key]

    yield get
    for e in made.values():
        e.close()
