"""CPU tests of the product's host planner (imageprocessor_b200/csrc/plan.cpp).

tests/support/plan_emu.cpp walks the planner's tables exactly as k_stream does
(same ownership rules, fp32 fmaf order, quantiser, ambiguity window).  Checked here:
  * the axis tables equal the oracle's newDistrib restatement bit for bit;
  * every output pixel is produced exactly once (tile/band ownership is a partition);
  * "certified fp32": every byte NOT flagged for the fp64 fix-up already equals the
    fp64 oracle, flagged bytes are within 1, and few pixels are flagged.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests.util import rgba_random

HERE = os.path.dirname(os.path.abspath(__file__))
SUP = os.path.join(HERE, "support")


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-C", SUP, "-s", "libplan_emu.so"])
    L = C.CDLL(os.path.join(SUP, "libplan_emu.so"))
    L.planemu_axis.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 5
    L.planemu_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    return L


def run_emu(L, a, specs, two_stage, bands_hint=4):
    h, w = a.shape[:2]
    n = len(specs)
    sp = np.array(specs, np.int32).reshape(-1)
    ts = np.array(two_stage, np.int32)
    dsts = [np.zeros((s[5], s[4], 4), np.uint8) for s in specs]
    flags = [np.zeros((s[5], s[4]), np.uint8) for s in specs]
    dp = (C.c_void_p * n)(*[d.ctypes.data for d in dsts])
    fp = (C.c_void_p * n)(*[f.ctypes.data for f in flags])
    info = np.zeros(5, np.int32)
    rc = L.planemu_run(a.ctypes.data, a.strides[0], w, h, n, sp.ctypes.data, ts.ctypes.data, dp, fp,
                       bands_hint, info.ctypes.data)
    return rc, dsts, flags, info


@pytest.mark.parametrize("dn,sn", [(1024, 4000), (768, 3000), (200, 3000), (1023, 1002), (767, 147), (7, 7), (1, 5000), (300, 301)])
def test_axis_tables_equal_oracle(emu, oracle, dn, sn):
    st, co, w, inv = oracle.distrib(dn, sn)
    off = np.zeros(dn + 1, np.int32); first = np.zeros(dn, np.int32)
    ww = np.zeros(len(w), np.float64); iv = np.zeros(dn, np.float64); mt = C.c_int()
    n = emu.planemu_axis(dn, sn, off.ctypes.data, first.ctypes.data, ww.ctypes.data, iv.ctypes.data, C.byref(mt))
    assert n == len(w)
    assert np.array_equal(off, st)
    assert np.array_equal(first, co[st[:-1]])
    assert np.array_equal(ww.view(np.uint64), w.view(np.uint64))          # bit-identical doubles
    assert np.array_equal(iv.view(np.uint64), inv.view(np.uint64))
    assert mt.value == int(np.diff(st).max())
    assert abs((w[st[0]:st[1]] * inv[0]).sum() - 1) < 1e-12


@pytest.mark.parametrize("w,h,rw,rh,size,bands", [
    (400, 300, 102, 76, 20, 1), (1000, 750, 256, 192, 50, 4), (1203, 899, 300, 224, 64, 7),
    (2000, 1500, 1024, 768, 200, 3), (600, 800, 1024, 768, 200, 2), (1601, 1201, 640, 480, 100, 5),
])
@pytest.mark.parametrize("alpha", ["opaque", "raw"])
def test_stream_plan_certified_fp32(emu, oracle, w, h, rw, rh, size, bands, alpha):
    a = rgba_random(w, h, w * 31 + h, alpha)
    nw, nh = oracle.keep_aspect_dims(w, h, rw, rh)
    cx, cy, cs = oracle.crop_square(w, h)
    specs = [(0, 0, w, h, nw, nh), (cx, cy, cs, cs, size, size)]
    rc, dsts, flags, info = run_emu(emu, a, specs, [0, 1], bands)
    if nh > h:   # vertical upscale cannot stream; the engine falls back to k_exact
        assert rc == -1
        return
    assert rc == 0
    R = oracle.Raster.rgba(a)
    refs = [oracle.resize_image(R, nw, nh), oracle.crop_and_resize(R, size)]
    for d, f, ref in zip(dsts, flags, refs):
        assert np.all((f == 16) | (f == 1)), "each output pixel written exactly once"
        amb = f == 1
        diff = np.abs(d.astype(int) - ref.astype(int)).max(axis=2)
        assert diff[~amb].max(initial=0) == 0, "an unflagged byte differs from the fp64 oracle"
        assert diff.max(initial=0) <= 1
        assert amb.mean() < 0.03


def test_stream_plan_rejects_what_it_cannot_do(emu):
    a = rgba_random(64, 48, 1)
    rc, *_ = run_emu(emu, a, [(0, 0, 64, 48, 200, 150)], [0])     # upscale
    assert rc == -1
    a = rgba_random(40000, 4, 2)
    rc, *_ = run_emu(emu, a, [(0, 0, 40000, 4, 40, 1)], [0])      # 1000-tap rows: wider than a slab
    assert rc == -1


def test_single_target_and_identity(emu, oracle):
    a = rgba_random(500, 333, 9, "raw")
    rc, dsts, flags, info = run_emu(emu, a, [(0, 0, 500, 333, 500, 333)], [0], 3)   # 1:1
    assert rc == 0
    ref = oracle.resize_image(oracle.Raster.rgba(a), 500, 333)
    amb = flags[0] == 1
    assert np.array_equal(dsts[0][~amb], ref[~amb])


def test_stream_plan_random_geometries(emu, oracle):
    """Seeded sweep over odd sizes, aspect ratios and scales (the mixed-size stream of BASELINE configs[4],
    shrunk so the oracle stays fast): whatever the planner accepts must be a partition of the outputs and
    certified; whatever it rejects must be a vertical upscale or a support wider than a slab."""
    rng = np.random.default_rng(4242)
    accepted = 0
    for case in range(40):
        w, h = int(rng.integers(33, 1400)), int(rng.integers(33, 1100))
        rw, rh = int(rng.integers(16, 1100)), int(rng.integers(16, 800))
        size = int(rng.integers(8, 220))
        bands = int(rng.integers(1, 9))
        a = rgba_random(w, h, 9000 + case, "raw" if case % 3 == 0 else "opaque")
        nw, nh = oracle.keep_aspect_dims(w, h, rw, rh)
        if nw <= 0 or nh <= 0:
            continue
        cx, cy, cs = oracle.crop_square(w, h)
        R = oracle.Raster.rgba(a)
        for spec, two_stage, ref in (((0, 0, w, h, nw, nh), 0, lambda: oracle.resize_image(R, nw, nh)),
                                     ((cx, cy, cs, cs, size, size), 1, lambda: oracle.crop_and_resize(R, size))):
            rc, dsts, flags, info = run_emu(emu, a, [spec], [two_stage], bands)
            if rc == -1:
                assert spec[5] > spec[3] or spec[2] / spec[4] > 200, f"planner rejected a streamable geometry {spec}"
                continue
            assert rc == 0, f"emulator error {rc} for {spec} in {w}x{h}"
            accepted += 1
            want = ref()
            f = flags[0]
            assert np.all((f == 16) | (f == 1)), f"outputs not written exactly once for {spec}"
            amb = f == 1
            diff = np.abs(dsts[0].astype(int) - want.astype(int)).max(axis=2)
            assert diff[~amb].max(initial=0) == 0 and diff.max(initial=0) <= 1, f"certification broken for {spec} in {w}x{h}"
    assert accepted >= 40


# ---- the 16-bit-sample kernels (k_stream_planar: planar YCbCr, Gray, NRGBA) ---------------------------------
def samples16(oracle, kind, w, h, seed):
    """A random source of the given kind as (oracle Raster, H x W x 4 uint16 samples): the 16-bit premultiplied
    values x/image's pass 1 reads per tap (draw/impl.go scaleX_NRGBA / scaleX_Gray / scaleX_YCbCr*), in numpy."""
    rng = np.random.default_rng(seed)
    s = np.empty((h, w, 4), np.uint16)
    if kind == "nrgba":
        a = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        a16 = a[..., 3].astype(np.uint32) * 0x101
        for q in range(3):
            s[..., q] = a[..., q].astype(np.uint32) * a16 // 0xFF
        s[..., 3] = a16
        return oracle.Raster.rgba(a, oracle.NRGBA8), s
    if kind == "gray":
        g = rng.integers(0, 256, (h, w), dtype=np.uint8)
        s[..., :3] = (g.astype(np.uint16) * 0x101)[..., None]
        s[..., 3] = 0xFFFF
        return oracle.Raster.gray(g), s
    lay = {"420": oracle.YCBCR420, "422": oracle.YCBCR422, "444": oracle.YCBCR444, "440": oracle.YCBCR440}[kind]
    ch, cw = oracle.chroma_shape(lay, w, h)
    y = rng.integers(0, 256, (h, w), dtype=np.uint8)
    cb, cr = rng.integers(0, 256, (ch, cw), dtype=np.uint8), rng.integers(0, 256, (ch, cw), dtype=np.uint8)
    yi = np.arange(h)[:, None] >> (1 if kind in ("420", "440") else 0)     # nearest chroma sample, as Go indexes it
    xi = np.arange(w)[None, :] >> (1 if kind in ("420", "422") else 0)
    yy1 = y.astype(np.int64) * 0x10101
    cb1, cr1 = cb[yi, xi].astype(np.int64) - 128, cr[yi, xi].astype(np.int64) - 128
    s[..., 0] = np.clip((yy1 + 91881 * cr1) >> 8, 0, 0xFFFF)
    s[..., 1] = np.clip((yy1 - 22554 * cb1 - 46802 * cr1) >> 8, 0, 0xFFFF)
    s[..., 2] = np.clip((yy1 + 116130 * cb1) >> 8, 0, 0xFFFF)
    s[..., 3] = 0xFFFF
    return oracle.Raster.ycbcr(y, cb, cr, lay), s


def run_emu16(L, s, spec, two_stage, bands_hint=4):
    h, w = s.shape[:2]
    L.planemu_run16.argtypes = L.planemu_run.argtypes
    sp = np.array(spec, np.int32)
    ts = np.array([two_stage], np.int32)
    dst, flag = np.zeros((spec[5], spec[4], 4), np.uint8), np.zeros((spec[5], spec[4]), np.uint8)
    dp, fp = (C.c_void_p * 1)(dst.ctypes.data), (C.c_void_p * 1)(flag.ctypes.data)
    info = np.zeros(5, np.int32)
    rc = L.planemu_run16(s.ctypes.data, w * 4, w, h, 1, sp.ctypes.data, ts.ctypes.data, dp, fp, bands_hint, info.ctypes.data)
    return rc, dst, flag


@pytest.mark.parametrize("kind", ["nrgba", "gray", "420", "422", "444", "440"])
def test_sixteen_bit_sample_kernels_certified_fp32(emu, oracle, kind):
    """The same certificate for k_stream_planar: over 16-bit samples with unscaled vertical weights, every byte the
    fp32 pass does not flag equals the oracle's float64 result for the real source (NRGBA, Gray, four YCbCr
    subsamplings), flagged ones are within 1 -- resize and the two-stage crop thumbnail, several geometries."""
    for i, (w, h, rw, rh, size, bands) in enumerate([(400, 300, 102, 76, 20, 1), (1203, 899, 300, 224, 64, 7),
                                                     (1601, 1201, 640, 480, 100, 5), (997, 1403, 1024, 768, 200, 3)]):
        R, s = samples16(oracle, kind, w, h, 700 + i)
        nw, nh = oracle.keep_aspect_dims(w, h, rw, rh)
        cx, cy, cs = oracle.crop_square(w, h)
        for spec, two, ref in (((0, 0, w, h, nw, nh), 0, lambda: oracle.resize_image(R, nw, nh)),
                               ((cx, cy, cs, cs, size, size), 1, lambda: oracle.crop_and_resize(R, size))):
            rc, d, f = run_emu16(emu, s, spec, two, bands)
            assert rc == 0, f"emulator error {rc} for {kind} {spec}"
            assert np.all((f == 16) | (f == 1)), "each output pixel written exactly once"
            amb = f == 1
            diff = np.abs(d.astype(int) - ref().astype(int)).max(axis=2)
            assert diff[~amb].max(initial=0) == 0, f"{kind} {spec}: an unflagged byte differs from the fp64 oracle"
            assert diff.max(initial=0) <= 1 and amb.mean() < 0.03
