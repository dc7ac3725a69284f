"""CPU check of the device JPEG writer's arithmetic: tests/support/jpeg_emu.cpp compiles the product's own
imageprocessor_b200/csrc/jpeg_core.h + jpeg_host.cpp (the per-thread functions of jpeg.cu and the host-built tables /
header) for the host and walks them in the kernels' decomposition.  Its files must equal the oracle's restatement of
Go's jpeg.Encode byte for byte.  The kernels proper (scans, warp collectives, staging) are the `-m gpu` tests' job
(tests/test_jpeg_gpu.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from tests.test_jpeg_oracle import smooth

SUP = os.path.join(os.path.dirname(os.path.abspath(__file__)), "support")


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-C", SUP, "-s", "libjpeg_emu.so"])
    L = C.CDLL(os.path.join(SUP, "libjpeg_emu.so"))
    L.jpeg_emu_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t]
    L.jpeg_emu_encode.restype = C.c_long

    def encode(rgba, quality=85, out_cap=None, scan_cap=None):
        a = np.ascontiguousarray(rgba, dtype=np.uint8)
        h, w = a.shape[:2]
        out = np.zeros(out_cap if out_cap is not None else w * h * 8 + 4096, np.uint8)
        n = L.jpeg_emu_encode(a.ctypes.data, a.strides[0], w, h, quality, out.ctypes.data, out.size,
                              scan_cap if scan_cap is not None else w * h * 8 + 4096)
        return None if n < 0 else out[:n].tobytes()
    return encode


def photo(rng, h, w, amp=12.0):
    return np.dstack([smooth(rng, h, w, amp) for _ in range(3)] + [np.full((h, w), 255, np.uint8)])


@pytest.mark.parametrize("size", [(1, 1), (7, 3), (16, 16), (17, 16), (16, 17), (33, 47), (200, 200), (1024, 768), (1023, 767), (640, 8)])
@pytest.mark.parametrize("quality", [85, 100, 10])
def test_emulated_kernels_equal_the_oracle(emu, size, quality):
    w, h = size
    rng = np.random.default_rng(w * 31 + h + quality)
    for img in (photo(rng, h, w), rng.integers(0, 256, (h, w, 4), dtype=np.uint8)):
        assert emu(img, quality) == O.jpeg_encode_rgba(img, quality)


def test_constant_and_extreme_images(emu):
    """All-0xff scan bytes (stuffing on every byte is impossible, but long 0xff runs are: white at q100 gives none, noise
    does), saturated chroma (the RGBToYCbCr clamps), black / white / primaries."""
    for colour in ((0, 0, 0), (255, 255, 255), (255, 0, 0), (0, 255, 0), (0, 0, 255), (255, 255, 0), (1, 254, 127)):
        img = np.zeros((40, 56, 4), np.uint8)
        img[..., :3] = colour
        img[..., 3] = 255
        assert emu(img) == O.jpeg_encode_rgba(img)
    rng = np.random.default_rng(1)
    noise = rng.integers(0, 2, (64, 64, 4), dtype=np.uint8) * 255  # +-full-swing edges: the largest coefficients
    for q in (100, 85, 1):
        assert emu(noise, q) == O.jpeg_encode_rgba(noise, q)


def test_premultiplied_alpha_is_ignored_like_the_go_writer(emu):
    rng = np.random.default_rng(2)
    img = photo(rng, 48, 64)
    a = rng.integers(0, 256, (48, 64), dtype=np.uint8)
    img[..., :3] = (img[..., :3].astype(np.uint16) * a[..., None] // 255).astype(np.uint8)
    img[..., 3] = a
    assert emu(img) == O.jpeg_encode_rgba(img)


def test_stuffing_crosses_chunk_and_lane_boundaries(emu):
    """Noise at quality 100 yields thousands of 0xff scan bytes spread over many 512-byte chunks."""
    rng = np.random.default_rng(4)
    img = rng.integers(0, 256, (128, 256, 4), dtype=np.uint8)
    f = emu(img, 100)
    assert f == O.jpeg_encode_rgba(img, 100)
    assert f.count(b"\xff\x00") > 100


def test_capacity_is_reported_not_overrun(emu):
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (64, 64, 4), dtype=np.uint8)
    ref = O.jpeg_encode_rgba(img, 85)
    assert emu(img, 85, out_cap=len(ref)) == ref
    assert emu(img, 85, out_cap=len(ref) - 1) is None
    assert emu(img, 85, scan_cap=512) is None


def test_exact_division_by_reciprocal():
    """jpeg_quant divides by 8 * quant with a multiply-high: exact for every numerator the DCT can produce (|a| + half < 2^18)
    and every divisor a quality setting can produce (8 .. 2040)."""
    n = np.arange(0, 1 << 18, dtype=np.uint64)
    for d in range(8, 2041, 8):
        recip = (1 << 32) // d + 1
        assert np.array_equal((n * np.uint64(recip)) >> np.uint64(32), n // np.uint64(d)), d


def test_random_small_images_property(emu):
    """Seeded sweep over sizes 1..40 x 1..40 and qualities 1..100: every file equals the oracle's (MCU padding in both
    axes, single-MCU images, scans shorter than one stuffing chunk)."""
    rng = np.random.default_rng(2024)
    for _ in range(150):
        w, h, q = int(rng.integers(1, 41)), int(rng.integers(1, 41)), int(rng.integers(1, 101))
        kind = int(rng.integers(0, 3))
        if kind == 0:
            img = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        elif kind == 1:
            img = photo(rng, h, w, 3.0)
        else:   # flat with one impulse: a long run of zeros then a coefficient (ZRL codes)
            img = np.full((h, w, 4), int(rng.integers(0, 256)), np.uint8)
            img[int(rng.integers(0, h)), int(rng.integers(0, w)), :3] = rng.integers(0, 256, 3)
        assert emu(img, q) == O.jpeg_encode_rgba(img, q), (w, h, q, kind)


def test_widest_image_the_writer_accepts(emu):
    """65535 pixels per side is the most Go's writer takes (it refuses 1 << 16): 4096 MCUs in one row, the last one padded."""
    rng = np.random.default_rng(6)
    img = np.repeat(rng.integers(0, 256, (3, 65535 // 15 + 1, 4), dtype=np.uint8), 15, axis=1)[:, :65535]
    f = emu(img, 85)
    assert f == O.jpeg_encode_rgba(img, 85)
    assert f[0x88:0x88 + 9] == bytes([0xff, 0xc0, 0, 17, 8, 0, 3, 0xff, 0xff])   # SOF0 at byte 136: height 3, width 65535
