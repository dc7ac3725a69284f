"""Parity of the CUDA path (through the C ABI) with the CPU oracle.

Bar (BASELINE.json north_star): watermark bit-exact; resize/thumbnail max |diff| <= 1
per uint8 channel.  The default EXACT mode is held to the stronger bar: byte-identical.
"""
import numpy as np
import pytest

import imageprocessor_b200 as ip
from tests.util import rgba_random, rgba_gradient, synthetic_glyphs, diff_stats

pytestmark = pytest.mark.gpu


def _ops_resize_thumb(w, h, rw=1024, rh=768, size=200):
    nw, nh = ip.keep_aspect_dims(w, h, rw, rh)
    cx, cy, cs = ip.crop_square(w, h)
    return [ip.OpSpec.resize(nw, nh), ip.OpSpec.thumb_crop((cx, cy, cs, cs), size)], (nw, nh)


@pytest.mark.parametrize("w,h,rw,rh,size", [
    (400, 300, 102, 76, 20), (1000, 750, 256, 192, 50), (401, 303, 100, 75, 33),
    (1999, 1201, 512, 384, 64), (640, 480, 1024, 768, 200), (300, 400, 1024, 768, 200),
])
@pytest.mark.parametrize("alpha", ["opaque", "premul", "raw"])
def test_resize_thumb_exact_mode_is_bit_exact(engines, oracle, w, h, rw, rh, size, alpha):
    a = rgba_random(w, h, 7 + w + h, alpha)
    ops, (nw, nh) = _ops_resize_thumb(w, h, rw, rh, size)
    out = engines(ip.PRECISION_EXACT).run(ip.Image.from_rgba(a), ops)
    R = oracle.Raster.rgba(a)
    assert diff_stats(out[0], oracle.resize_image(R, nw, nh)) == (0, 0.0)
    assert diff_stats(out[1], oracle.crop_and_resize(R, size)) == (0, 0.0)


@pytest.mark.parametrize("w,h", [(1000, 750), (1999, 1201), (2400, 1800)])
def test_resize_thumb_fast_mode_within_one(engines, oracle, w, h):
    a = rgba_random(w, h, 11)
    ops, (nw, nh) = _ops_resize_thumb(w, h, 1024, 768, 200)
    out = engines(ip.PRECISION_FAST).run(ip.Image.from_rgba(a, opaque_hint=True), ops)
    R = oracle.Raster.rgba(a)
    m0, f0 = diff_stats(out[0], oracle.resize_image(R, nw, nh))
    m1, f1 = diff_stats(out[1], oracle.crop_and_resize(R, 200))
    print(f"fast-mode mismatch fraction: resize {f0:.2e} thumb {f1:.2e}")
    assert m0 <= 1 and m1 <= 1          # tolerance stated by north_star
    assert f0 < 2e-3 and f1 < 2e-3


def test_reference_mode_is_bit_exact(engines, oracle):
    a = rgba_random(700, 500, 3, "premul")
    ops, (nw, nh) = _ops_resize_thumb(700, 500, 256, 256, 40)
    out = engines(ip.PRECISION_REFERENCE).run(ip.Image.from_rgba(a), ops)
    R = oracle.Raster.rgba(a)
    assert diff_stats(out[0], oracle.resize_image(R, nw, nh)) == (0, 0.0)
    assert diff_stats(out[1], oracle.crop_and_resize(R, 40)) == (0, 0.0)


@pytest.mark.parametrize("w,h", [(640, 480), (1001, 701)])
def test_watermark_bit_exact(engines, oracle, w, h):
    a = rgba_random(w, h, 5, "premul")
    gl = synthetic_glyphs(w, h, 99, n=8)
    col = (255, 255, 255, 127)
    ops = [ip.OpSpec.watermark(w, h, col, [ip.GlyphMask(*g) for g in gl])]
    out = engines(ip.PRECISION_EXACT).run(ip.Image.from_rgba(a), ops)
    ref = oracle.watermark(oracle.Raster.rgba(a), col, [oracle.Glyph(*g) for g in gl])
    assert np.array_equal(out[0], ref)
    assert not np.array_equal(out[0], a)      # the blend did something


@pytest.mark.parametrize("fuse", [1, 2, 3])
def test_full_pipeline_one_pass(engines, oracle, fuse):
    w, h = 1600, 1200
    a = rgba_gradient(w, h)
    gl = synthetic_glyphs(w, h, 1, n=10)
    col = (255, 255, 255, 127)
    ops, (nw, nh) = _ops_resize_thumb(w, h)
    ops.append(ip.OpSpec.watermark(w, h, col, [ip.GlyphMask(*g) for g in gl]))
    e = engines(ip.PRECISION_EXACT, fuse_targets=fuse)      # one pass per target (default) / both targets fused
    k0 = e.stats()["kernels_launched"]
    out = e.run(ip.Image.from_rgba(a), ops)
    assert e.stats()["kernels_launched"] > k0
    R = oracle.Raster.rgba(a)
    assert np.array_equal(out[0], oracle.resize_image(R, nw, nh))
    assert np.array_equal(out[1], oracle.crop_and_resize(R, 200))
    assert np.array_equal(out[2], oracle.watermark(R, col, [oracle.Glyph(*g) for g in gl]))


def _ycbcr(oracle, w, h, layout, seed):
    rng = np.random.default_rng(seed)
    ch, cw = oracle.chroma_shape(layout, w, h)
    return (rng.integers(0, 256, (h, w), dtype=np.uint8), rng.integers(0, 256, (ch, cw), dtype=np.uint8),
            rng.integers(0, 256, (ch, cw), dtype=np.uint8))


@pytest.mark.parametrize("layout", [ip.YCBCR444, ip.YCBCR422, ip.YCBCR420, ip.YCBCR440])
def test_ycbcr_sources(engines, oracle, layout):
    w, h = 403, 301
    y, cb, cr = _ycbcr(oracle, w, h, layout, 17)
    gl = synthetic_glyphs(w, h, 4, n=5)
    col = (10, 200, 30, 200)
    ops, (nw, nh) = _ops_resize_thumb(w, h, 128, 128, 32)
    ops.append(ip.OpSpec.watermark(w, h, col, [ip.GlyphMask(*g) for g in gl]))
    out = engines(ip.PRECISION_EXACT).run(ip.Image.from_ycbcr(y, cb, cr, layout), ops)
    R = oracle.Raster.ycbcr(y, cb, cr, layout)
    assert np.array_equal(out[0], oracle.resize_image(R, nw, nh))
    assert np.array_equal(out[1], oracle.crop_and_resize(R, 32))
    assert np.array_equal(out[2], oracle.watermark(R, col, [oracle.Glyph(*g) for g in gl]))


def test_nrgba_and_gray_sources(engines, oracle):
    w, h = 333, 222
    a = rgba_random(w, h, 21, "raw")
    g = np.random.default_rng(22).integers(0, 256, (h, w), dtype=np.uint8)
    gl = synthetic_glyphs(w, h, 4, n=4)
    col = (0, 0, 0, 127)
    for img, R in ((ip.Image.from_rgba(a, ip.NRGBA8), oracle.Raster.rgba(a, oracle.NRGBA8)),
                   (ip.Image.from_gray(g), oracle.Raster.gray(g))):
        ops, (nw, nh) = _ops_resize_thumb(w, h, 100, 100, 25)
        ops.append(ip.OpSpec.watermark(w, h, col, [ip.GlyphMask(*x) for x in gl]))
        out = engines(ip.PRECISION_EXACT).run(img, ops)
        assert np.array_equal(out[0], oracle.resize_image(R, nw, nh))
        assert np.array_equal(out[1], oracle.crop_and_resize(R, 25))
        assert np.array_equal(out[2], oracle.watermark(R, col, [oracle.Glyph(*x) for x in gl]))


def test_many_tickets_batched(engines, oracle):
    e = engines(ip.PRECISION_EXACT)
    imgs = [rgba_random(800 + 8 * k, 600 + 4 * k, 100 + k) for k in range(12)]
    tickets, want = [], []
    for a in imgs:
        h, w = a.shape[:2]
        ops, (nw, nh) = _ops_resize_thumb(w, h, 320, 240, 48)
        tickets.append(e.submit(ip.Image.from_rgba(a), ops))
        want.append((nw, nh))
    for a, t, (nw, nh) in zip(imgs, tickets, want):
        out = e.wait(t)
        R = oracle.Raster.rgba(a)
        assert np.array_equal(out[0], oracle.resize_image(R, nw, nh))
        assert np.array_equal(out[1], oracle.crop_and_resize(R, 48))


def test_pinned_and_device_buffers(engines, oracle):
    e = engines(ip.PRECISION_EXACT)
    w, h = 1024, 768
    a = rgba_random(w, h, 8)
    pin = e.alloc_pinned(a.nbytes)
    pin.array[:] = a.reshape(-1)
    pa = pin.array.reshape(h, w, 4)
    out_pin = e.alloc_pinned(256 * 192 * 4)
    dst = out_pin.array.reshape(192, 256, 4)
    s0 = e.stats()["staged_copies"]
    e.run(ip.Image.from_rgba(pa), [ip.OpSpec.resize(256, 192, dst=dst)])
    assert e.stats()["staged_copies"] == s0          # zero-copy: no staging memcpy
    ref = oracle.resize_image(oracle.Raster.rgba(a), 256, 192)
    assert np.array_equal(dst, ref)
    # device-resident source and destination
    dsrc = e.alloc_device(0, a.nbytes)
    ddst = e.alloc_device(0, 256 * 192 * 4)
    e.to_device(0, dsrc, a)
    img = ip.Image.on_device(ip.RGBA8, w, h, [dsrc], [w * 4])
    e.run(img, [ip.OpSpec.resize(256, 192, dst_device=(ddst, 256 * 4))], device=0)
    back = np.empty((192, 256, 4), np.uint8)
    e.from_device(0, back, ddst)
    assert np.array_equal(back, ref)
    e.free_device(0, dsrc)
    e.free_device(0, ddst)
    pin.free()
    out_pin.free()


def test_error_paths(engines):
    e = engines(ip.PRECISION_EXACT)
    a = rgba_random(64, 64, 1)
    with pytest.raises(ip.IpgError) as ei:
        e.run(ip.Image.from_rgba(a), [ip.OpSpec.thumb_crop((10, 10, 100, 100), 8)])
    assert ei.value.code == ip._lib.ERR_INVALID
    with pytest.raises(ip.IpgError):
        e.run(ip.Image.from_rgba(a), [ip.OpSpec(99, 8, 8)])
    # zero-sized output (e.g. keep-aspect of an extreme strip) is an empty image, not an error
    out = e.run(ip.Image.from_rgba(a), [ip.OpSpec.resize(0, 5)])
    assert out[0].size == 0


def test_mixed_size_stream(engines, oracle):
    """BASELINE configs[4] in miniature: a seeded stream of odd sizes and aspect ratios (portrait, square,
    panoramic, tiny), every image through all three operations, tickets in flight together."""
    e = engines(ip.PRECISION_EXACT)
    rng = np.random.default_rng(77)
    shapes = [(int(rng.integers(40, 1900)), int(rng.integers(40, 1300))) for _ in range(14)]
    shapes += [(1, 1), (3, 2), (2048, 17), (19, 1500), (1024, 768), (200, 200)]
    work = []
    for k, (w, h) in enumerate(shapes):
        a = rgba_random(w, h, 3000 + k, ["opaque", "premul", "raw"][k % 3])
        nw, nh = ip.keep_aspect_dims(w, h, 1024, 768)
        cx, cy, cs = ip.crop_square(w, h)
        gl = synthetic_glyphs(w, h, k, n=4)
        col = (255, 255, 255, 127)
        ops = [ip.OpSpec.resize(nw, nh), ip.OpSpec.thumb_crop((cx, cy, cs, cs), 200),
               ip.OpSpec.watermark(w, h, col, [ip.GlyphMask(*g) for g in gl])]
        work.append((a, (nw, nh), gl, col, e.submit(ip.Image.from_rgba(a), ops)))
    for a, (nw, nh), gl, col, t in work:
        out = e.wait(t)
        R = oracle.Raster.rgba(a)
        h, w = a.shape[:2]
        if nw > 0 and nh > 0:
            assert np.array_equal(out[0], oracle.resize_image(R, nw, nh)), f"resize {w}x{h} -> {nw}x{nh}"
        assert np.array_equal(out[1], oracle.crop_and_resize(R, 200)), f"thumb {w}x{h}"
        assert np.array_equal(out[2], oracle.watermark(R, col, [oracle.Glyph(*g) for g in gl])), f"watermark {w}x{h}"


def test_concurrent_submitters_and_timeouts(engines, oracle):
    """WORKER_CONCURRENCY goroutines call Process at once (worker.go:90-96): 8 threads submit and wait on one
    context concurrently; every result is right, tickets are single-use, a zero timeout reports IPG_ERR_TIMEOUT
    and leaves the ticket valid."""
    import threading
    e = engines(ip.PRECISION_EXACT)
    errors, lock = [], threading.Lock()

    def worker(k):
        try:
            rng = np.random.default_rng(400 + k)
            for it in range(6):
                w, h = int(rng.integers(200, 1500)), int(rng.integers(200, 1100))
                a = rgba_random(w, h, 1000 * k + it, ["opaque", "premul"][it % 2])
                ops, (nw, nh) = _ops_resize_thumb(w, h, 512, 384, 64)
                t = e.submit(ip.Image.from_rgba(a), ops)
                out = e.wait(t)
                R = oracle.Raster.rgba(a)
                ok = np.array_equal(out[0], oracle.resize_image(R, nw, nh)) and np.array_equal(out[1], oracle.crop_and_resize(R, 64))
                if not ok:
                    raise AssertionError(f"thread {k} image {it} ({w}x{h}) differs")
        except Exception as ex:  # noqa: BLE001
            with lock:
                errors.append(repr(ex))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    # ticket semantics
    a = rgba_random(3000, 2000, 9)
    t = e.submit(ip.Image.from_rgba(a), [ip.OpSpec.resize(750, 500)])
    try:
        e.wait(t, timeout_ms=0)          # may or may not have finished yet
        finished_early = True
    except ip.IpgError as ex:
        assert ex.code == ip._lib.ERR_TIMEOUT
        finished_early = False
    if not finished_early:
        out = e.wait(t)                  # still valid after the timeout
        assert np.array_equal(out[0], oracle.resize_image(oracle.Raster.rgba(a), 750, 500))
    with pytest.raises(ip.IpgError) as ei:
        e.wait(t)                        # consumed
    assert ei.value.code == ip._lib.ERR_INVALID


@pytest.mark.parametrize("slack", ["tiny", "exact", "plus4", "plus240"])
def test_fix_list_overflow_paths(oracle, slack, monkeypatch):
    """The fp64 fix list is sized generously; these engines are built with a tiny one (IPG_FIX_CAPACITY) to drive
    the paths a normal run never takes: the list overflows (every target is redone whole in fp64), and the list
    holds the entries but has no room at its back end to re-queue the wide-support ones (the warp that meets one
    finishes it in place).  The capacities are chosen around the number of pixels this source flags (measured first
    with the default list; it depends on the certificate's window D), a few hundred of them wide-support."""
    a = rgba_random(1600, 1200, 0)
    nw, nh = ip.keep_aspect_dims(1600, 1200, 1024, 768)
    ops = lambda: [ip.OpSpec.resize(nw, nh), ip.OpSpec.thumb_crop((200, 0, 1200, 1200), 200)]  # noqa: E731
    with ip.Engine(devices=[0], precision=ip.PRECISION_EXACT) as e:
        e.run(ip.Image.from_rgba(a), ops())
        n_flagged = e.stats()["exact_fixups"]
    assert 200 < n_flagged < 5000
    capacity = {"tiny": 64, "exact": n_flagged, "plus4": n_flagged + 4, "plus240": n_flagged + 240}[slack]
    monkeypatch.setenv("IPG_FIX_CAPACITY", str(capacity))
    with ip.Engine(devices=[0], precision=ip.PRECISION_EXACT) as e:
        out = e.run(ip.Image.from_rgba(a), ops())
        n_fix = e.stats()["exact_fixups"]
    R = oracle.Raster.rgba(a)
    assert n_fix == n_flagged
    assert np.array_equal(out[0], oracle.resize_image(R, nw, nh))
    assert np.array_equal(out[1], oracle.crop_and_resize(R, 200))


@pytest.mark.parametrize("kind", ["nrgba", "gray", "420", "422", "444", "440"])
def test_sixteen_bit_sample_layouts_odd_sizes(engines, oracle, kind):
    """k_stream_planar (planar YCbCr, Gray, NRGBA) over awkward geometries: widths that are not multiples of the
    4-pixel thread granule or of the 16-byte bulk-copy granule, single rows and columns, extreme aspect ratios,
    mild and strong downscales, the centre crop at odd offsets.  Whatever the planner cannot stream falls back to the fp64
    kernel; the bytes must be the oracle's either way."""
    rng = np.random.default_rng({"nrgba": 1, "gray": 2, "420": 3, "422": 4, "444": 5, "440": 6}[kind])
    sizes = [(1, 1), (2, 3), (5, 4), (17, 2048), (2048, 17), (333, 222), (1001, 999), (1280, 960), (1919, 1081), (2500, 64)]
    e = engines(ip.PRECISION_EXACT)
    for w, h in sizes:
        if kind == "nrgba":
            a = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
            img, R = ip.Image.from_rgba(a, ip.NRGBA8), oracle.Raster.rgba(a, oracle.NRGBA8)
        elif kind == "gray":
            g = rng.integers(0, 256, (h, w), dtype=np.uint8)
            img, R = ip.Image.from_gray(g), oracle.Raster.gray(g)
        else:
            lay = {"420": ip.YCBCR420, "422": ip.YCBCR422, "444": ip.YCBCR444, "440": ip.YCBCR440}[kind]
            y, cb, cr = _ycbcr(oracle, w, h, lay, int(rng.integers(1 << 30)))
            img, R = ip.Image.from_ycbcr(y, cb, cr, lay), oracle.Raster.ycbcr(y, cb, cr, lay)
        dw, dh = max(1, int(w / rng.uniform(1.0, 6.0))), max(1, int(h / rng.uniform(1.0, 6.0)))
        cx, cy, cs = ip.crop_square(w, h)       # the reference's centre square (thumbnail.go:115-127)
        size = int(rng.integers(1, 96))
        out = e.run(img, [ip.OpSpec.resize(dw, dh), ip.OpSpec.thumb_crop((cx, cy, cs, cs), size)])
        assert np.array_equal(out[0], oracle.resize_image(R, dw, dh)), f"{kind} resize {w}x{h} -> {dw}x{dh}"
        assert np.array_equal(out[1], oracle.crop_and_resize(R, size)), f"{kind} thumb {w}x{h} crop {cx},{cy},{cs} -> {size}"


@pytest.mark.gpu
def test_submit_returns_nomem_instead_of_blocking_when_staging_is_exhausted(monkeypatch):
    """ipg_submit must answer with a status when pageable outputs that were never waited for fill the
    pinned staging pool (staging of a pageable destination is released by ipg_wait) -- not block for ever."""
    import time
    import imageprocessor_b200 as ip
    from imageprocessor_b200 import _lib as L
    monkeypatch.setenv("IPG_STAGING_TIMEOUT_MS", "300")
    w, h = 1024, 1024                                   # 4 MiB source + 4 MiB watermark output per ticket, both pageable
    a = rgba_random(w, h, 1)
    with ip.Engine(devices=[0], lanes_per_device=1, lane_pinned_bytes=16 << 20) as e:
        held, t0, err = [], time.time(), None
        try:
            for _ in range(16):                          # 16 x 4 MiB of outputs never fit 16 MiB of staging
                held.append(e.submit(ip.Image.from_rgba(a), [ip.OpSpec.watermark(w, h, (255, 255, 255, 127), [])]))
        except L.IpgError as ex:
            err = ex
        assert err is not None and err.code == L.ERR_NOMEM and "not waited for" in str(err)
        assert time.time() - t0 < 30
        for t in held:                                   # the tickets that were accepted still complete, bit-exact
            assert np.array_equal(e.wait(t)[0], a)
        # and the pool is whole again
        assert np.array_equal(e.run(ip.Image.from_rgba(a), [ip.OpSpec.watermark(w, h, (255, 255, 255, 127), [])])[0], a)


@pytest.mark.parametrize("w,h,where", [(1600, 1200, "pinned-alias"), (4000, 3000, "pageable-copy"), (1001, 777, "device")])
def test_watermark_patch_only_equals_the_full_frame_result(engines, oracle, w, h, where):
    """IPG_OPF_WATERMARK_PATCH_ONLY: for an *image.RGBA source draw.Draw(Src) is a copy, so only the glyph union box is
    produced and written; with dst holding the source pixels (the source buffer itself, or a copy) the result must be the
    oracle's full watermarked frame, and nothing outside the box may be touched.  Other ops of the same ticket still see
    the ORIGINAL pixels although the watermark lands in the very buffer they were uploaded from."""
    from imageprocessor_b200 import _lib as L
    e = engines(ip.PRECISION_EXACT)
    a = rgba_random(w, h, 55, "premul")
    gl = synthetic_glyphs(w, h, 9, n=7)
    col = (255, 255, 255, 127)
    R = oracle.Raster.rgba(a)
    want = oracle.watermark(R, col, [oracle.Glyph(*g) for g in gl])
    glyphs = [ip.GlyphMask(*g) for g in gl]
    nw, nh = ip.keep_aspect_dims(w, h, 1024, 768)
    d0 = e.stats()["bytes_d2h"]
    if where == "pinned-alias":
        buf = e.alloc_pinned(a.nbytes)
        src = buf.array.reshape(h, w, 4)
        src[...] = a
        out = e.run(ip.Image.from_rgba(src), [ip.OpSpec.watermark(w, h, col, glyphs, dst=src, flags=L.OPF_WATERMARK_PATCH_ONLY),
                                              ip.OpSpec.resize(nw, nh)])
        got = src.copy()
        assert np.array_equal(out[1], oracle.resize_image(R, nw, nh)), "the resize must see the original, not the patched buffer"
        buf.free()
    elif where == "pageable-copy":
        dst = a.copy()
        dst[0, 0] = (1, 2, 3, 4)                       # a sentinel outside the box: the engine must not touch it
        e.run(ip.Image.from_rgba(a), [ip.OpSpec.watermark(w, h, col, glyphs, dst=dst, flags=L.OPF_WATERMARK_PATCH_ONLY)])
        assert tuple(dst[0, 0]) == (1, 2, 3, 4)
        dst[0, 0] = a[0, 0]
        got = dst
    else:
        p_src, p_dst = e.alloc_device(0, a.nbytes), e.alloc_device(0, a.nbytes)
        e.to_device(0, p_src, a)
        e.to_device(0, p_dst, a)
        e.run(ip.Image.on_device(L.RGBA8, w, h, [p_src], [w * 4]),
              [ip.OpSpec.watermark(w, h, col, glyphs, dst_device=(p_dst, w * 4), flags=L.OPF_WATERMARK_PATCH_ONLY)], device=0)
        got = np.empty_like(a)
        e.from_device(0, got, p_dst)
        e.free_device(0, p_src)
        e.free_device(0, p_dst)
    assert np.array_equal(got, want)
    if where != "device":
        x0, y0 = min(g.x0 for g in gl), min(g.y0 for g in gl)
        x1, y1 = max(g.x1 for g in gl), max(g.y1 for g in gl)
        moved = e.stats()["bytes_d2h"] - d0
        assert moved < (x1 - x0) * (y1 - y0) * 4 + nw * nh * 4 + 4096, "only the box (and the resize) may cross PCIe"


def test_watermark_patch_only_is_ignored_for_converting_layouts(engines, oracle):
    """For NRGBA / YCbCr / Gray sources draw.Draw(Src) is a conversion: the flag is ignored and the full frame is produced."""
    from imageprocessor_b200 import _lib as L
    e = engines(ip.PRECISION_EXACT)
    w, h = 640, 480
    a = rgba_random(w, h, 3, "raw")
    gl = synthetic_glyphs(w, h, 2)
    col = (10, 200, 30, 200)
    out = e.run(ip.Image.from_rgba(a, ip.NRGBA8), [ip.OpSpec.watermark(w, h, col, [ip.GlyphMask(*g) for g in gl],
                                                                      flags=L.OPF_WATERMARK_PATCH_ONLY)])
    assert np.array_equal(out[0], oracle.watermark(oracle.Raster.rgba(a, oracle.NRGBA8), col, [oracle.Glyph(*g) for g in gl]))


@pytest.mark.parametrize("kind", ["nrgba", "gray", "420", "422", "444", "440"])
def test_watermark_frame_of_converting_layouts_fused_and_standalone(engines, oracle, kind):
    """draw.Draw(Src) of a non-RGBA source is a conversion.  With a streaming resize in the ticket it rides on that pass
    (k_stream_planar stores the high bytes of the 16-bit samples it computes); without one -- or when the resize cannot
    stream (upscale) -- the vectorised k_watermark does it.  Odd widths, unaligned tails, bands and tiles; every byte
    of the frame, the glyph blend on top, and the resize / thumbnail of the same ticket against the oracle."""
    rng = np.random.default_rng({"nrgba": 11, "gray": 12, "420": 13, "422": 14, "444": 15, "440": 16}[kind])
    e = engines(ip.PRECISION_EXACT)
    col = (255, 255, 255, 127)
    for (w, h, with_resize) in [(1603, 1201, True), (2049, 777, True), (4000, 3000, True), (333, 222, True), (1001, 999, False),
                                (40, 30, True), (5000, 129, True)]:
        if kind == "nrgba":
            a = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
            img, R = ip.Image.from_rgba(a, ip.NRGBA8), oracle.Raster.rgba(a, oracle.NRGBA8)
        elif kind == "gray":
            g = rng.integers(0, 256, (h, w), dtype=np.uint8)
            img, R = ip.Image.from_gray(g), oracle.Raster.gray(g)
        else:
            lay = {"420": ip.YCBCR420, "422": ip.YCBCR422, "444": ip.YCBCR444, "440": ip.YCBCR440}[kind]
            y, cb, cr = _ycbcr(oracle, w, h, lay, int(rng.integers(1 << 30)))
            img, R = ip.Image.from_ycbcr(y, cb, cr, lay), oracle.Raster.ycbcr(y, cb, cr, lay)
        gl = synthetic_glyphs(w, h, 21, n=5)
        ops = [ip.OpSpec.watermark(w, h, col, [ip.GlyphMask(*g) for g in gl])]
        nw, nh = ip.keep_aspect_dims(w, h, 1024, 768)
        cx, cy, cs = ip.crop_square(w, h)
        if with_resize:
            ops += [ip.OpSpec.resize(nw, nh), ip.OpSpec.thumb_crop((cx, cy, cs, cs), 64)]
        out = e.run(img, ops)
        assert np.array_equal(out[0], oracle.watermark(R, col, [oracle.Glyph(*g) for g in gl])), f"{kind} watermark {w}x{h}"
        if with_resize:
            assert np.array_equal(out[1], oracle.resize_image(R, nw, nh)), f"{kind} resize {w}x{h}"
            assert np.array_equal(out[2], oracle.crop_and_resize(R, 64)), f"{kind} thumb {w}x{h}"


def _pinned_planes(e, w, h):
    cw, ch = (w + 1) // 2, (h + 1) // 2
    bufs = [e.alloc_pinned(w * h), e.alloc_pinned(cw * ch), e.alloc_pinned(cw * ch)]
    planes = (bufs[0].array.reshape(h, w), bufs[1].array.reshape(ch, cw), bufs[2].array.reshape(ch, cw))
    for p in planes:
        p[...] = 0x5A
    return bufs, planes


@pytest.mark.parametrize("w,h", [(1600, 1200), (1001, 777), (33, 17), (4000, 3000)])
def test_ycbcr420_destination_is_what_gos_jpeg_writer_derives(engines, oracle, w, h):
    """ipg_op.dst_layout = YCBCR420 (first step): each result comes back as the planar 4:2:0 image Go's
    image/jpeg writer computes from the *image.RGBA result before its DCT (integer RGBToYCbCr, 2x2 chroma mean with
    (sum + 2) >> 2, edge replication) -- compared here with the oracle's restatement applied to the oracle's RGBA result,
    for all three ops, odd sizes included, RGBA and planar sources."""
    e = engines(ip.PRECISION_EXACT, lane_device_bytes=2 << 30)
    col = (255, 255, 255, 127)
    for src_kind in ("rgba", "420"):
        if src_kind == "rgba":
            a = rgba_random(w, h, 91, "premul")
            img, R = ip.Image.from_rgba(a), oracle.Raster.rgba(a)
        else:
            y, cb, cr = _ycbcr(oracle, w, h, ip.YCBCR420, 92)
            img, R = ip.Image.from_ycbcr(y, cb, cr, ip.YCBCR420), oracle.Raster.ycbcr(y, cb, cr, oracle.YCBCR420)
        nw, nh = ip.keep_aspect_dims(w, h, 1024, 768)
        cx, cy, cs = ip.crop_square(w, h)
        gl = synthetic_glyphs(w, h, 5)
        keep = []
        specs = []
        for (dw, dh) in ((nw, nh), (64, 64), (w, h)):
            bufs, planes = _pinned_planes(e, dw, dh)
            keep += bufs
            specs.append(planes)
        out = e.run(img, [ip.OpSpec.resize(nw, nh, dst_ycbcr420=specs[0]),
                          ip.OpSpec.thumb_crop((cx, cy, cs, cs), 64, dst_ycbcr420=specs[1]),
                          ip.OpSpec.watermark(w, h, col, [ip.GlyphMask(*g) for g in gl], dst_ycbcr420=specs[2])])
        want = [oracle.resize_image(R, nw, nh), oracle.crop_and_resize(R, 64), oracle.watermark(R, col, [oracle.Glyph(*g) for g in gl])]
        for k, (planes, rgba) in enumerate(zip(out, want)):
            ey, ecb, ecr = oracle.rgba_to_ycbcr420(rgba)
            assert np.array_equal(planes[0], ey), f"{src_kind} op {k}: Y"
            assert np.array_equal(planes[1], ecb) and np.array_equal(planes[2], ecr), f"{src_kind} op {k}: chroma"
        for b in keep:
            b.free()


def test_ycbcr420_destination_argument_errors(engines):
    from imageprocessor_b200 import _lib as L
    e = engines(ip.PRECISION_EXACT)
    a = rgba_random(64, 48, 1)
    y, cb, cr = np.empty((48, 64), np.uint8), np.empty((24, 32), np.uint8), np.empty((24, 32), np.uint8)
    with pytest.raises(ip.IpgError) as ei:      # pageable planes: there is no staging path for this layout
        e.run(ip.Image.from_rgba(a), [ip.OpSpec.resize(64, 48, dst_ycbcr420=(y, cb, cr))])
    assert ei.value.code == L.ERR_INVALID and "pinned" in str(ei.value)


@pytest.mark.parametrize("w,h,size,alpha,merge", [
    (1000, 750, 50, "opaque", "1"),      # 15:1, integer centres, one band
    (1356, 2203, 100, "opaque", "1"),    # 13.56:1 portrait: fractional centres, several bands
    (2600, 1951, 200, "opaque", "1"),    # 9.755:1
    (3204, 2401, 160, "opaque", "0"),    # 15.006:1 through the separate wide-target launch (k_stream<1,WM,2>)
    (2592, 2160, 100, "opaque", "1"),    # 21.6:1, the 8K thumbnail's ratio: segments of 21-22 rows, flushed in two pieces
    (3004, 3000, 100, "opaque", "1"),    # 30:1, the 48 MP thumbnail's ratio
    (1500, 1300, 100, "premul", "1"),    # alpha < 255: the lean kernel raises the redo flag, the fp32 form redoes the job
])
def test_integer_moment_thumbnail_is_bit_exact(oracle, monkeypatch, w, h, size, alpha, merge):
    """The integer-moment vertical pass of wide 8-bit targets (GroupRecI, k_stream's IDP.2A loop; 8.5:1 ... 32:1
    thumbnails): byte-identical to the oracle, like the fp32 form it replaces (IPG_VINT=0), with the watermark copy
    riding on the thumbnail pass (thumbnail first in the op list) and a resize beside it.  The two forms flag different
    pixels (their certificates differ), which is how the test knows the integer form ran."""
    a = rgba_random(w, h, w + h, alpha)
    nw, nh = ip.keep_aspect_dims(w, h, 640, 480)
    cx, cy, cs = ip.crop_square(w, h)
    gl = synthetic_glyphs(w, h, 3)
    col = (255, 255, 255, 127)
    R = oracle.Raster.rgba(a)
    want = [oracle.crop_and_resize(R, size), oracle.watermark(R, col, [oracle.Glyph(*g) for g in gl]), oracle.resize_image(R, nw, nh)]
    monkeypatch.setenv("IPG_MERGE_LEAN", merge)
    fixups = {}
    for vint in ("1", "0"):
        monkeypatch.setenv("IPG_VINT", vint)
        with ip.Engine(devices=[0], precision=ip.PRECISION_EXACT) as e:
            out = e.run(ip.Image.from_rgba(a), [ip.OpSpec.thumb_crop((cx, cy, cs, cs), size), ip.OpSpec.watermark(w, h, col, [ip.GlyphMask(*g) for g in gl]),
                                                ip.OpSpec.resize(nw, nh)])
            fixups[vint] = e.stats()["exact_fixups"]
        for k, name in enumerate(("thumbnail", "watermark", "resize")):
            assert np.array_equal(out[k], want[k]), f"IPG_VINT={vint}: {name} differs from the oracle"
    if alpha == "opaque":
        assert fixups["1"] != fixups["0"], "the integer-moment form did not run"
