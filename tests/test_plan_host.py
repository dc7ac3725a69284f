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
    info = np.zeros(8, np.int32)
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
    info = np.zeros(8, np.int32)
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


# ---- the fp32 certificate, attacked (DESIGN.md 2.1 has the derivation) ---------------------------
def exact_T(oracle, a8, spec):
    """256 * v + 128 for every output channel in float64 from dense weight matrices (error ~1e-9 units): the real
    value the fp32 kernel approximates, in the units of the certificate (1/256 of a 16-bit step)."""
    rx, ry, rw, rh, dw, dh = spec

    def dense(dn, sn):
        st, co, w, inv = oracle.distrib(dn, sn)
        M = np.zeros((dn, sn), np.float64)
        for o in range(dn):
            M[o, co[st[o]:st[o + 1]]] = w[st[o]:st[o + 1]] * inv[o]
        return M, int(np.diff(st).max())

    Mx, tx = dense(dw, rw)
    My, ty = dense(dh, rh)
    crop = a8[ry:ry + rh, rx:rx + rw].astype(np.float64) * 257.0
    T = np.empty((dh, dw, 4))
    for ch in range(4):
        T[..., ch] = (My @ crop[..., ch] @ Mx.T) * 256.0 + 128.0
    return T, tx, ty


def run_emu_capture(L, a, spec, two_stage, bands=3):
    cap = np.zeros((spec[5], spec[4], 4), np.float32)
    L.planemu_capture.argtypes = [C.c_void_p, C.c_void_p]
    L.planemu_capture(cap.ctypes.data, None)
    try:
        rc, dsts, flags, info = run_emu(L, a, [spec], [two_stage], bands)
    finally:
        L.planemu_capture(None, None)
    return rc, dsts[0], flags[0], info, cap


def adversarial_images(w, h):
    yield "random", rgba_random(w, h, 5)
    a = np.full((h, w, 4), 255, np.uint8)
    yield "all-255 (largest partial sums: every rounding at its half-ulp maximum)", a
    for period, phase in [(2, 0), (2, 1), (3, 1), (5, 2), (7, 3)]:
        a = np.zeros((h, w, 4), np.uint8)
        a[..., 3] = 255
        a[(np.arange(h) % period) == phase, :, :3] = 255
        yield f"row stripes {period}/{phase}", a
        a = np.zeros((h, w, 4), np.uint8)
        a[..., 3] = 255
        a[:, (np.arange(w) % period) == phase, :3] = 255
        yield f"column stripes {period}/{phase}", a
    yy, xx = np.mgrid[0:h, 0:w]
    a = np.zeros((h, w, 4), np.uint8)
    a[..., 3] = 255
    a[..., :3] = (((yy + xx) & 1) * 255)[..., None]
    yield "checkerboard", a


@pytest.mark.parametrize("w,h,spec", [
    (400, 300, (0, 0, 400, 300, 102, 76)),        # the 4:1 resize shape (8 x 8 taps)
    (1000, 750, (125, 0, 750, 750, 50, 50)),      # the 15:1 thumbnail shape (29-31 taps, split over 4 threads)
    (640, 480, (0, 0, 640, 480, 512, 384)),       # mild 1.25:1 downscale (several outputs per lane)
    (900, 880, (10, 0, 880, 880, 20, 20)),        # 44:1 -- the 8K thumbnail's support (87-89 taps is past it; 44 here)
])
def test_fp32_error_stays_inside_the_proven_bound(emu, oracle, w, h, spec):
    """max |T_fp32 - T_exact| over adversarial images must stay below the derived bound, with the flag rule
    (window D = bound + floor + margin) catching every byte that differs from the float64 oracle."""
    worst = 0.0
    for name, a in adversarial_images(w, h):
        rc, d, f, info, cap = run_emu_capture(emu, a, spec, 0)
        assert rc == 0, name
        T, tx, ty = exact_T(oracle, a, spec)
        D = int(info[5])
        err = np.abs(cap.astype(np.float64) - T)
        # the clamp to alpha acts on r, g, b only when they exceed alpha: opaque inputs here, so it never binds
        worst = max(worst, float(err.max()))
        assert err.max() <= D - 2, f"{name}: |T_fp32 - T_exact| = {err.max():.2f} exceeds the proven bound {D - 2} (D = {D})"
        rx, ry, rw, rh, dw, dh = spec
        ref = oracle.scale_bilinear(oracle.Raster.rgba(a), (rx, ry, rw, rh), dw, dh)
        amb = f == 1
        assert np.array_equal(d[~amb], ref[~amb]), f"{name}: an unflagged byte differs from the float64 oracle"
        assert np.abs(d.astype(int) - ref.astype(int)).max() <= 1
    print(f"{spec}: taps {tx}x{ty}, D = {D}, worst observed |dT| = {worst:.2f}")
    assert D == emu.planemu_fix_d(tx, ty, 1) or D > emu.planemu_fix_d(tx, ty, 1)   # parts only ever widen the window


def test_constants_sit_mid_cell_and_are_never_flagged(emu, oracle):
    """byte * 0x101 puts an exact constant at T = c * 65792 + 128: 128 units from either quantiser step, so a constant
    region is never ambiguous while D < 128 -- every value 0..255, inside bands 3 supports tall."""
    w, band = 256, 40
    a = np.zeros((256 * band, w, 4), np.uint8)
    a[..., 3] = 255
    for c in range(256):
        a[c * band:(c + 1) * band, :, :3] = c
    spec = (0, 0, w, 256 * band, 64, 256 * band // 4)
    rc, d, f, info, cap = run_emu_capture(emu, a, spec, 0, 8)
    assert rc == 0 and info[5] < 128
    for c in range(256):
        rows = slice(c * band // 4 + 3, (c + 1) * band // 4 - 3)       # output rows whose support lies inside band c
        assert (d[rows, :, :3] == c).all() and (d[rows, :, 3] == 255).all()
        assert (f[rows] == 16).all(), f"constant {c} was flagged"
        assert np.all(np.abs(cap[rows, :, 0] - (c * 65792 + 128)) <= info[5] - 2)


def test_values_engineered_onto_a_quantiser_step_are_flagged(emu, oracle):
    """For a handful of output pixels the source bytes under their support are searched until the EXACT value sits
    within a few hundredths of a unit of a quantiser step (256 v + 128 = 65536 m): the worst case for the certificate.
    Every one of them must be flagged, and no unflagged byte anywhere may differ from the float64 oracle."""
    w, h, spec = 400, 300, (0, 0, 400, 300, 102, 76)
    a = rgba_random(w, h, 77)
    stx, cox, wx, invx = oracle.distrib(102, 400)
    sty, coy, wy, invy = oracle.distrib(76, 300)
    rng = np.random.default_rng(1)
    targets = [(ox, oy) for oy in range(5, 70, 9) for ox in range(6, 96, 11)]
    hit = []
    for (ox, oy) in targets:
        xs, ys = cox[stx[ox]:stx[ox + 1]], coy[sty[oy]:sty[oy + 1]]
        c = 257.0 * 256.0 * np.outer(wy[sty[oy]:sty[oy + 1]] * invy[oy], wx[stx[ox]:stx[ox + 1]] * invx[ox])  # units per byte step
        ch = int(rng.integers(0, 3))
        blk = a[ys[0]:ys[-1] + 1, xs[0]:xs[-1] + 1, ch].astype(np.int64)
        T = float((c * blk).sum()) + 128.0
        goal = round(T / 65536.0) * 65536.0
        goal = min(max(goal, 65536.0), 255 * 65536.0)
        r = goal - T
        # best single or pairwise +-1 move per round (differences of two weights reach far below the smallest weight)
        cf, bf = c.reshape(-1), blk.reshape(-1)
        n = cf.size
        for it in range(600):
            if abs(r) < 0.02:
                break
            if abs(r) > 2.0 * cf.max():                          # coarse phase: the largest whole step any one tap allows
                step = np.clip(np.round(r / cf), -bf, 255 - bf).astype(np.int64)
                q = int(np.argmin(np.abs(r - cf * step)))
                if step[q] != 0:
                    bf[q] += step[q]
                    r -= cf[q] * step[q]
                    continue
            up, dn = bf < 255, bf > 0
            cand = []                                            # (delta of T, i, di, j, dj)
            one = np.concatenate([np.where(up, cf, np.inf), np.where(dn, -cf, np.inf)])
            k = int(np.argmin(np.abs(r - one)))
            cand.append((one[k], k % n, 1 if k < n else -1, -1, 0))
            for di, mi in ((1, up), (-1, dn)):
                for dj, mj in ((1, up), (-1, dn)):
                    M = di * cf[:, None] + dj * cf[None, :]
                    M = np.where(mi[:, None] & mj[None, :], M, np.inf)
                    np.fill_diagonal(M, np.inf)
                    k = int(np.argmin(np.abs(r - M)))
                    cand.append((M.reshape(-1)[k], k // n, di, k % n, dj))
            dT, i, di, j, dj = min(cand, key=lambda t: abs(r - t[0]))
            if not np.isfinite(dT) or abs(r - dT) >= abs(r):
                q = int(rng.integers(0, n))                                  # stuck: nudge one tap and go on
                bf[q] = min(max(bf[q] + int(rng.integers(-2, 3)), 0), 255)
                r = goal - (float((cf * bf).sum()) + 128.0)
                continue
            bf[i] += di
            if j >= 0:
                bf[j] += dj
            r -= dT
        blk = bf.reshape(blk.shape)
        a[ys[0]:ys[-1] + 1, xs[0]:xs[-1] + 1, ch] = blk.astype(np.uint8)
        hit.append((ox, oy, ch, abs(r)))
    assert sum(1 for t in hit if t[3] < 0.5) >= len(hit) * 0.6, "the search failed to reach the step"
    rc, d, f, info, cap = run_emu_capture(emu, a, spec, 0)
    assert rc == 0
    T, _, _ = exact_T(oracle, a, spec)
    ref = oracle.resize_image(oracle.Raster.rgba(a), 102, 76)
    amb = f == 1
    assert np.array_equal(d[~amb], ref[~amb]), "an unflagged byte differs from the float64 oracle"
    for (ox, oy, ch, r) in hit:
        if r < 0.5:
            dist = abs(((T[oy, ox, ch] + 32768.0) % 65536.0) - 32768.0)
            assert dist < 1.0 and amb[oy, ox], f"pixel ({ox},{oy}) sits {dist:.3f} units from a step and was not flagged"


def test_8k_thumbnail_support_certified(emu, oracle):
    """The 8K crop thumbnail: 4320 -> 200, 44 taps per axis, one output split over 4 threads -- the widest support of
    the BASELINE configs (the 48 MP one, 61 taps, runs on the GPU in tests/test_full_size.py)."""
    w, h = 2592, 2160       # half-scale 8K with the same 21.6:1 ratio and tap counts: crop 2160^2 -> 100
    a = rgba_random(w, h, 8)
    spec = (216, 0, 2160, 2160, 100, 100)
    rc, d, f, info, cap = run_emu_capture(emu, a, spec, 1, 4)
    assert rc == 0
    T, tx, ty = exact_T(oracle, a, spec)
    assert tx >= 44 and ty >= 44
    assert np.abs(cap.astype(np.float64) - T).max() <= info[5] - 2
    ref = oracle.crop_and_resize(oracle.Raster.rgba(a), 100)
    amb = f == 1
    assert np.array_equal(d[~amb], ref[~amb]) and np.abs(d.astype(int) - ref.astype(int)).max() <= 1


# ---- k_direct: the small-support kernel (vertical upscales, mild downscales) ------------------------------------------
def run_direct(L, a, spec, two_stage, sixteen=False):
    h, w = a.shape[:2]
    dst = np.zeros((spec[5], spec[4], 4), np.uint8)
    flag = np.zeros((spec[5], spec[4]), np.uint8)
    cap = np.zeros((spec[5], spec[4], 4), np.float32)
    info = np.zeros(4, np.int32)
    sp = np.array(spec, np.int32)
    fn = L.planemu_direct16 if sixteen else L.planemu_direct
    fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    rc = fn(a.ctypes.data, a.strides[0] // a.itemsize, w, h, sp.ctypes.data, two_stage, dst.ctypes.data, flag.ctypes.data,
            cap.ctypes.data, info.ctypes.data)
    return rc, dst, flag, cap, info


@pytest.mark.parametrize("w,h,dw,dh", [(640, 480, 1024, 768), (147, 147, 767, 767), (1152, 864, 1024, 768), (1300, 975, 1024, 768),
                                       (1002, 751, 1023, 767), (400, 30, 102, 76), (64, 48, 64, 48), (3, 2, 9, 7)])
@pytest.mark.parametrize("alpha", ["opaque", "raw"])
def test_direct_kernel_certified_fp32(emu, oracle, w, h, dw, dh, alpha):
    """Upscales (what the reference does to small images), 1:1 and mild downscales: unflagged bytes equal the float64
    oracle, flagged ones are within 1, the measured error stays inside the proven bound."""
    a = rgba_random(w, h, w * 7 + h, alpha)
    rc, d, f, cap, info = run_direct(emu, a, (0, 0, w, h, dw, dh), 0)
    assert rc == 0
    ref = oracle.resize_image(oracle.Raster.rgba(a), dw, dh)
    amb = f == 1
    diff = np.abs(d.astype(int) - ref.astype(int)).max(axis=2)
    assert diff[~amb].max(initial=0) == 0, "an unflagged byte differs from the fp64 oracle"
    assert diff.max(initial=0) <= 1 and amb.mean() < 0.03
    if alpha == "opaque":       # (with alpha < 255 the premultiplied clamp may bind: the bound is about the unclamped sums)
        T, tx, ty = exact_T(oracle, a, (0, 0, w, h, dw, dh))
        assert np.abs(cap.astype(np.float64) - T).max() <= info[0] - 2
        assert info[0] == emu.planemu_fix_d(int(info[1]), int(info[2]), 1)


def test_direct_kernel_two_stage_sixteen_bit_and_rejections(emu, oracle):
    # the crop thumbnail of a small image (cropAndResize's 8-bit crop stage, then an upscale)
    a = rgba_random(180, 120, 3, "raw")
    cx, cy, cs = oracle.crop_square(180, 120)
    rc, d, f, cap, info = run_direct(emu, a, (cx, cy, cs, cs, 200, 200), 1)
    assert rc == 0
    ref = oracle.crop_and_resize(oracle.Raster.rgba(a), 200)
    assert np.array_equal(d[f != 1], ref[f != 1]) and np.abs(d.astype(int) - ref.astype(int)).max() <= 1
    # 16-bit samples (what an NRGBA / YCbCr / Gray source feeds): upscale of a 4:2:0 source
    R, s = samples16(oracle, "420", 333, 222, 5)
    rc, d, f, cap, info = run_direct(emu, s, (0, 0, 333, 222, 1024, 682), 0, sixteen=True)
    assert rc == 0
    ref = oracle.resize_image(R, 1024, 682)
    assert np.array_equal(d[f != 1], ref[f != 1]) and np.abs(d.astype(int) - ref.astype(int)).max() <= 1
    # supports wider than DIRECT_MAX_TAPS on a streamable geometry are not this kernel's: the planner says so
    rc, *_ = run_direct(emu, rgba_random(400, 300, 1), (0, 0, 400, 300, 102, 76), 0)
    assert rc == -1


# ---- the integer-moment vertical form of wide 8-bit targets (GroupRecI; k_stream's IDP.2A loop) -----------------------------
@pytest.fixture()
def vint_emu(emu):
    emu.planemu_set_vint(1)
    yield emu
    emu.planemu_set_vint(0)


@pytest.mark.parametrize("w,h,size,bands", [
    (1000, 750, 50, 3),      # 15:1, integer centres (the 12 MP thumbnail's shape)
    (1000, 750, 50, 1),      # ... one band
    (1356, 2203, 100, 7),    # 13.56:1 portrait, fractional centres, many bands
    (905, 640, 64, 2),       # 10:1
    (2592, 2160, 100, 4),    # 21.6:1, the 8K thumbnail's ratio: segments of 21-22 rows flushed in two pieces
    (3004, 3000, 100, 5),    # 30:1, the 48 MP thumbnail's ratio
    (1283, 1279, 40, 6),     # 31.98:1, the longest segments the form takes
])
def test_integer_moment_form_certified(vint_emu, oracle, w, h, size, bands):
    """Every byte the integer-moment form does not flag equals the float64 oracle, flagged ones are within 1, the
    measured |T - T_exact| stays inside the window the planner derived for this form (smaller than the fp32 chain's),
    and fewer pixels are flagged than by the fp32 form."""
    a = rgba_random(w, h, w * 3 + h)
    cx, cy, cs = oracle.crop_square(w, h)
    spec = (cx, cy, cs, cs, size, size)
    rc, d, f, info, cap = run_emu_capture(vint_emu, a, spec, 1, bands)
    assert rc == 0 and info[7] == 1, "the planner did not offer the integer-moment form"
    T, tx, ty = exact_T(oracle, a, spec)
    err = np.abs(cap.astype(np.float64) - T)[..., :3].max()
    assert err <= info[5] - 2, f"|T - T_exact| = {err:.2f} exceeds the derived bound (D = {info[5]})"
    assert info[5] < vint_emu.planemu_fix_d(tx, ty, 4)
    ref = oracle.crop_and_resize(oracle.Raster.rgba(a), size)
    assert np.all((f == 16) | (f == 1)), "each output pixel written exactly once"
    amb = f == 1
    assert np.array_equal(d[~amb], ref[~amb]) and np.abs(d.astype(int) - ref.astype(int)).max() <= 1
    vint_emu.planemu_set_vint(0)
    rc0, d0, f0, info0, _ = run_emu_capture(vint_emu, a, spec, 1, bands)
    assert rc0 == 0 and (f0 == 1).sum() >= amb.sum()


def test_integer_moment_form_attacked(vint_emu, oracle):
    """The adversarial images of the fp32 certificate through the integer form (15:1 and 21.6:1): all-255 puts every
    moment at its maximum (M0 = 16 * 255 fills its 12 bits), stripes at every phase move the mass to either end of a
    segment; a non-opaque source is not this form's (the emulator, like the engine's redo, takes the fp32 one)."""
    for (w, h, spec) in [(1000, 750, (125, 0, 750, 750, 50, 50)), (1300, 1080, (110, 0, 1080, 1080, 50, 50))]:
        worst = 0.0
        for name, a in adversarial_images(w, h):
            rc, d, f, info, cap = run_emu_capture(vint_emu, a, spec, 0)
            assert rc == 0 and info[7] == 1, name
            T, _, _ = exact_T(oracle, a, spec)
            err = float(np.abs(cap.astype(np.float64) - T)[..., :3].max())
            worst = max(worst, err)
            assert err <= info[5] - 2, f"{name}: |T - T_exact| = {err:.2f} exceeds the derived bound (D = {info[5]})"
            ref = oracle.scale_bilinear(oracle.Raster.rgba(a), spec[:4], spec[4], spec[5])
            amb = f == 1
            assert np.array_equal(d[~amb], ref[~amb]), f"{name}: an unflagged byte differs from the float64 oracle"
        print(f"{spec}: D = {info[5]}, worst observed |dT| = {worst:.2f}")
    a = rgba_random(1000, 750, 3, "raw")
    rc, d, f, info, cap = run_emu_capture(vint_emu, a, (125, 0, 750, 750, 50, 50), 1)
    ref = oracle.crop_and_resize(oracle.Raster.rgba(a), 50)
    assert rc == 0 and np.array_equal(d[f != 1], ref[f != 1])


def test_integer_moment_form_is_refused_where_it_does_not_apply(vint_emu, oracle):
    """Local (narrow-support) targets, scales above 32:1 and 16-bit sample sources keep the fp32 form."""
    for (w, h, spec) in [(400, 300, (0, 0, 400, 300, 102, 76)),        # 4:1 resize: local
                         (700, 660, (20, 0, 660, 660, 20, 20)),        # 33:1
                         (640, 480, (80, 0, 480, 480, 200, 200))]:     # 2.4:1 thumbnail of a small image
        rc, dsts, flags, info = run_emu(vint_emu, rgba_random(w, h, 1), [spec], [1])
        assert rc == 0 and info[7] == 0, spec
    R, s = samples16(oracle, "420", 1000, 750, 5)
    rc, d, f = run_emu16(vint_emu, s, (125, 0, 750, 750, 50, 50), 1)
    assert rc == 0 and np.array_equal(d[f != 1], oracle.crop_and_resize(R, 50)[f != 1])


def test_integer_moment_form_random_geometries(vint_emu, oracle):
    """Seeded sweep over crop sizes, offsets, scales 8.5:1 ... 33:1 and band counts: whatever geometry the planner gives
    the integer form must be a partition of the outputs, keep its records consistent (one end per group, pieces of at
    most 16 rows: the emulator returns an error otherwise) and be certified; above 32:1 it must fall back."""
    rng = np.random.default_rng(2024)
    taken = 0
    for case in range(24):
        size = int(rng.integers(16, 72))
        scale = float(rng.uniform(8.6, 33.5))
        cs = int(round(size * scale)) + int(rng.integers(-3, 4))
        w, h = cs + int(rng.integers(0, 260)), cs + int(rng.integers(0, 40))
        if case % 2:
            w, h = h, w
        a = rgba_random(w, h, 500 + case)
        if case % 6 == 5:
            a[..., :3] = 255
        cx, cy, cs2 = oracle.crop_square(w, h)
        spec = (cx, cy, cs2, cs2, size, size)
        rc, d, f, info, cap = run_emu_capture(vint_emu, a, spec, 1, int(rng.integers(1, 9)))
        assert rc == 0, f"emulator error {rc} for {spec} in {w}x{h}"
        s = cs2 / size
        assert info[7] == (1 if s <= 32.0 else 0), f"{spec}: scale {s:.2f}"
        taken += int(info[7])
        T, _, _ = exact_T(oracle, a, spec)
        assert np.abs(cap.astype(np.float64) - T)[..., :3].max() <= info[5] - 2
        ref = oracle.crop_and_resize(oracle.Raster.rgba(a), size)
        assert np.all((f == 16) | (f == 1))
        assert np.array_equal(d[f != 1], ref[f != 1]) and np.abs(d.astype(int) - ref.astype(int)).max() <= 1, spec
    assert taken >= 18
