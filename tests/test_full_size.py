"""BASELINE.json's full sizes on the GPU: the 12 MP and 8K configurations against the oracle
(a few seconds of CPU each), plus size-independent properties of the path."""
import numpy as np
import pytest

import imageprocessor_b200 as ip
from imageprocessor_b200 import glyphs as G
from tests.util import rgba_random, rgba_gradient

pytestmark = pytest.mark.gpu


def full_pipeline(e, a, text="© ImageProcessor"):
    h, w = a.shape[:2]
    nw, nh = ip.keep_aspect_dims(w, h, 1024, 768)
    cx, cy, cs = ip.crop_square(w, h)
    gl = G.layout_watermark(w, h, text)
    col, _ = G.parse_color("255,255,255", 0.5)
    out = e.run(ip.Image.from_rgba(a), [ip.OpSpec.resize(nw, nh), ip.OpSpec.thumb_crop((cx, cy, cs, cs), 200),
                                          ip.OpSpec.watermark(w, h, col, gl)])
    return out, (nw, nh), gl, col


@pytest.mark.parametrize("w,h,gen,fuse", [(4000, 3000, "random", 1), (4000, 3000, "random", 3), (4000, 3000, "gradient", 2),
                                          (7680, 4320, "random", 1), (3000, 4000, "random", 3)])
def test_baseline_sizes_bit_exact(engines, oracle, w, h, gen, fuse):
    a = rgba_random(w, h, 1000) if gen == "random" else rgba_gradient(w, h)
    e = engines(ip.PRECISION_EXACT, lane_device_bytes=2 << 30, fuse_targets=fuse)
    f0 = e.stats()["exact_fallbacks"]
    out, (nw, nh), gl, col = full_pipeline(e, a)
    assert e.stats()["exact_fallbacks"] == f0, "the streaming kernel must take these geometries"
    assert (nw, nh) == {(4000, 3000): (1024, 768), (7680, 4320): (1024, 576), (3000, 4000): (576, 768)}[(w, h)]
    R = oracle.Raster.rgba(a)
    assert np.array_equal(out[0], oracle.resize_image(R, nw, nh))
    assert np.array_equal(out[1], oracle.crop_and_resize(R, 200))
    og = [oracle.Glyph(g.x0, g.y0, g.x1, g.y1, g.mask, g.mp_x, g.mp_y) for g in gl]
    assert np.array_equal(out[2], oracle.watermark(R, col, og))


def test_fast_mode_12mp_within_one_and_mismatch_fraction(engines, oracle):
    a = rgba_random(4000, 3000, 1001)
    out, (nw, nh), *_ = full_pipeline(engines(ip.PRECISION_FAST), a)
    R = oracle.Raster.rgba(a)
    for got, want in ((out[0], oracle.resize_image(R, nw, nh)), (out[1], oracle.crop_and_resize(R, 200))):
        d = np.abs(got.astype(np.int16) - want.astype(np.int16))
        frac = float((d > 0).mean())
        print(f"fast mode mismatch fraction {frac:.2e}")
        assert d.max() <= 1 and frac < 1e-3          # north_star tolerance: max |diff| <= 1, fraction reported


def test_properties_at_full_size(engines):
    e = engines(ip.PRECISION_EXACT, lane_device_bytes=2 << 30)
    # a constant image stays that constant through both resamplers; the watermark only touches its glyph box
    a = np.empty((3000, 4000, 4), np.uint8)
    a[...] = (37, 201, 90, 255)
    out, (nw, nh), gl, col = full_pipeline(e, a)
    assert (out[0] == a[0, 0]).all() and (out[1] == a[0, 0]).all()
    x0, y0 = min(g.x0 for g in gl), min(g.y0 for g in gl)
    x1, y1 = max(g.x1 for g in gl), max(g.y1 for g in gl)
    outside = np.ones(a.shape[:2], bool)
    outside[y0:y1, x0:x1] = False
    assert np.array_equal(out[2][outside], a[outside]) and not np.array_equal(out[2][y0:y1, x0:x1], a[y0:y1, x0:x1])
    # 1:1 resize is the identity on valid premultiplied input; transparent images stay transparent
    b = rgba_random(1024, 768, 5, "premul")
    assert np.array_equal(e.run(ip.Image.from_rgba(b), [ip.OpSpec.resize(1024, 768)])[0], b)
    z = np.zeros((1200, 1600, 4), np.uint8)
    o = e.run(ip.Image.from_rgba(z), [ip.OpSpec.resize(1024, 768), ip.OpSpec.thumb_crop((200, 0, 1200, 1200), 200)])
    assert not o[0].any() and not o[1].any()
    # linearity of the filter up to the quantiser: resize(255 - a) == 255 - resize(a) within 1 on opaque input
    c = rgba_random(2000, 1500, 6)
    inv = c.copy()
    inv[..., :3] = 255 - c[..., :3]
    r1 = e.run(ip.Image.from_rgba(c), [ip.OpSpec.resize(512, 384)])[0].astype(np.int16)
    r2 = e.run(ip.Image.from_rgba(inv), [ip.OpSpec.resize(512, 384)])[0].astype(np.int16)
    assert np.abs((255 - r1[..., :3]) - r2[..., :3]).max() <= 1


@pytest.mark.parametrize("layout", ["420", "444", "422", "440"])
def test_ycbcr_12mp_streams_and_is_bit_exact(engines, oracle, layout):
    """A 12 MP planar source (what image.Decode hands over for a JPEG): resize + crop thumbnail run on the
    planar streaming kernel (no whole-image fp64 fallback) and equal the oracle's per-tap 16-bit conversion."""
    lay = {"420": ip.YCBCR420, "444": ip.YCBCR444, "422": ip.YCBCR422, "440": ip.YCBCR440}[layout]
    olay = {"420": oracle.YCBCR420, "444": oracle.YCBCR444, "422": oracle.YCBCR422, "440": oracle.YCBCR440}[layout]
    w, h = 4000, 3000
    rng = np.random.default_rng(99)
    ch, cw = oracle.chroma_shape(olay, w, h)
    y = rng.integers(0, 256, (h, w), dtype=np.uint8)
    cb, cr = rng.integers(0, 256, (ch, cw), dtype=np.uint8), rng.integers(0, 256, (ch, cw), dtype=np.uint8)
    e = engines(ip.PRECISION_EXACT, lane_device_bytes=2 << 30)
    f0 = e.stats()["exact_fallbacks"]
    out = e.run(ip.Image.from_ycbcr(y, cb, cr, lay), [ip.OpSpec.resize(1024, 768), ip.OpSpec.thumb_crop((500, 0, 3000, 3000), 200)])
    assert e.stats()["exact_fallbacks"] == f0
    R = oracle.Raster.ycbcr(y, cb, cr, olay)
    assert np.array_equal(out[0], oracle.resize_image(R, 1024, 768))
    assert np.array_equal(out[1], oracle.crop_and_resize(R, 200))


@pytest.mark.parametrize("w,h", [(4001, 3003), (3001, 2003), (1920, 1080)])
def test_gray_streams_and_is_bit_exact(engines, oracle, w, h):
    g = np.random.default_rng(5).integers(0, 256, (h, w), dtype=np.uint8)
    e = engines(ip.PRECISION_EXACT, lane_device_bytes=2 << 30)
    f0 = e.stats()["exact_fallbacks"]
    nw, nh = ip.keep_aspect_dims(w, h, 1024, 768)
    cx, cy, cs = ip.crop_square(w, h)
    out = e.run(ip.Image.from_gray(g), [ip.OpSpec.resize(nw, nh), ip.OpSpec.thumb_crop((cx, cy, cs, cs), 200)])
    assert e.stats()["exact_fallbacks"] == f0
    R = oracle.Raster.gray(g)
    assert np.array_equal(out[0], oracle.resize_image(R, nw, nh))
    assert np.array_equal(out[1], oracle.crop_and_resize(R, 200))


@pytest.mark.parametrize("w,h,alpha", [(4000, 3000, "raw"), (3001, 2003, "raw"), (1920, 1080, "edges"), (640, 480, "raw")])
def test_nrgba_streams_and_is_bit_exact(engines, oracle, w, h, alpha):
    """*image.NRGBA (a PNG with an alpha channel): resize + crop thumbnail run on the 16-bit-sample streaming
    kernel (no whole-image fp64 fallback) and equal the oracle's per-tap integer premultiply
    (x/image scaleX_NRGBA: a16 = A * 0x101, c16 = C * a16 / 0xff)."""
    a = rgba_random(w, h, 31, "raw")
    if alpha == "edges":  # mostly opaque or transparent, as real cut-outs are
        m = np.random.default_rng(7).integers(0, 8, (h, w))
        a[..., 3] = np.where(m < 3, 0, np.where(m < 7, 255, a[..., 3]))
    e = engines(ip.PRECISION_EXACT, lane_device_bytes=2 << 30)
    f0 = e.stats()["exact_fallbacks"]
    nw, nh = ip.keep_aspect_dims(w, h, 1024, 768)
    cx, cy, cs = ip.crop_square(w, h)
    ops = [ip.OpSpec.thumb_crop((cx, cy, cs, cs), 200)]
    if nh <= h:  # (a vertical upscale does not stream: whole-image fp64, tested in test_gpu_parity)
        ops.insert(0, ip.OpSpec.resize(nw, nh))
    out = e.run(ip.Image.from_rgba(a, ip.NRGBA8), ops)
    assert e.stats()["exact_fallbacks"] == f0
    R = oracle.Raster.rgba(a, oracle.NRGBA8)
    if nh <= h:
        assert np.array_equal(out[0], oracle.resize_image(R, nw, nh))
    assert np.array_equal(out[-1], oracle.crop_and_resize(R, 200))


# ---- VERDICT r1 "untested at full size": per-pixel alpha at 12 MP, and the top of config 5's range (48 MP) ----------
@pytest.mark.parametrize("alpha", ["premul", "raw", "one_pixel"])
def test_alpha_paths_at_12mp_bit_exact(engines, oracle, alpha):
    """12 MP *image.RGBA with alpha < 255: the lean kernel raises the job's redo flag and the general instantiation
    (per-pixel alpha lanes, premultiplied clamp of the thumbnail's 8-bit crop stage) redoes it.  'one_pixel': a single
    translucent pixel in the last band -- the flag must still be raised and the whole job redone."""
    w, h = 4000, 3000
    a = rgba_random(w, h, 4242, "opaque" if alpha == "one_pixel" else alpha)
    if alpha == "one_pixel":
        a[2991, 3977] = (10, 20, 30, 200)
    e = engines(ip.PRECISION_EXACT, lane_device_bytes=2 << 30)
    f0 = e.stats()["exact_fallbacks"]
    out, (nw, nh), gl, col = full_pipeline(e, a)
    assert e.stats()["exact_fallbacks"] == f0
    R = oracle.Raster.rgba(a)
    assert np.array_equal(out[0], oracle.resize_image(R, nw, nh))
    assert np.array_equal(out[1], oracle.crop_and_resize(R, 200))
    og = [oracle.Glyph(g.x0, g.y0, g.x1, g.y1, g.mask, g.mp_x, g.mp_y) for g in gl]
    assert np.array_equal(out[2], oracle.watermark(R, col, og))


@pytest.mark.parametrize("layout", ["rgba", "rgba_alpha", "nrgba", "420", "444", "gray"])
def test_48mp_every_layout_bit_exact(engines, oracle, layout):
    """8000x6000 (config 5's largest size; 192 MB in + 192 MB watermark out for RGBA), full pipeline, every source type."""
    w, h = 8000, 6000
    rng = np.random.default_rng(48)
    e = engines(ip.PRECISION_EXACT, lane_device_bytes=3 << 30, lane_pinned_bytes=1 << 30)
    nw, nh = ip.keep_aspect_dims(w, h, 1024, 768)
    cx, cy, cs = ip.crop_square(w, h)
    gl = G.layout_watermark(w, h, "© ImageProcessor")
    col, _ = G.parse_color("255,255,255", 0.5)
    if layout in ("rgba", "rgba_alpha", "nrgba"):
        a = rgba_random(w, h, 48, {"rgba": "opaque", "rgba_alpha": "premul", "nrgba": "raw"}[layout])
        img = ip.Image.from_rgba(a, ip.NRGBA8 if layout == "nrgba" else ip.RGBA8)
        R = oracle.Raster.rgba(a, oracle.NRGBA8 if layout == "nrgba" else oracle.RGBA8)
    elif layout == "gray":
        g8 = rng.integers(0, 256, (h, w), dtype=np.uint8)
        img, R = ip.Image.from_gray(g8), oracle.Raster.gray(g8)
    else:
        lay, olay = {"420": (ip.YCBCR420, oracle.YCBCR420), "444": (ip.YCBCR444, oracle.YCBCR444)}[layout]
        ch, cw = oracle.chroma_shape(olay, w, h)
        y = rng.integers(0, 256, (h, w), dtype=np.uint8)
        cb, cr = rng.integers(0, 256, (ch, cw), dtype=np.uint8), rng.integers(0, 256, (ch, cw), dtype=np.uint8)
        img, R = ip.Image.from_ycbcr(y, cb, cr, lay), oracle.Raster.ycbcr(y, cb, cr, olay)
    f0 = e.stats()["exact_fallbacks"]
    out = e.run(img, [ip.OpSpec.resize(nw, nh), ip.OpSpec.thumb_crop((cx, cy, cs, cs), 200), ip.OpSpec.watermark(w, h, col, gl)])
    assert e.stats()["exact_fallbacks"] == f0, "48 MP must stream (no whole-image fp64 fallback)"
    assert np.array_equal(out[0], oracle.resize_image(R, nw, nh))
    assert np.array_equal(out[1], oracle.crop_and_resize(R, 200))
    og = [oracle.Glyph(g.x0, g.y0, g.x1, g.y1, g.mask, g.mp_x, g.mp_y) for g in gl]
    assert np.array_equal(out[2], oracle.watermark(R, col, og))
