"""GIF / paletted and 16-bit sources.

*image.Paletted reaches x/image's Scale and stdlib draw.Draw through the generic At(x, y).RGBA() path.  The product
does not carry a paletted layout: the host expands the palette into an RGBA8 (entries are color.RGBA: GIF, PNG without
tRNS) or NRGBA8 (color.NRGBA entries: PNG with tRNS) raster first (imageprocessor_b200/codecs.py expand_paletted; a Go
host would do the same loop).  not gpu: the oracle's generic paletted path equals its typed path on the expanded
raster, byte for byte -- the identity the integration relies on.  gpu: the engine on the expanded raster and on the
16-bit layouts (RGBA64 / NRGBA64 / Gray16, big-endian Pix) against the oracle's generic path.
"""
import io

import numpy as np
import pytest

import imageprocessor_b200 as ip
from imageprocessor_b200 import codecs
from tests.util import synthetic_glyphs


def _palette(rng, kind):
    pal = rng.integers(0, 256, (256, 4), dtype=np.uint8)
    if kind == "opaque":
        pal[:, 3] = 255
    elif kind == "gif-transparent":          # GIF: one fully transparent entry, stored as color.RGBA{0,0,0,0}
        pal[:, 3] = 255
        pal[rng.integers(0, 256)] = 0
    return pal                               # "trns": arbitrary straight alpha per entry (PNG tRNS -> color.NRGBA)


@pytest.mark.parametrize("kind", ["opaque", "gif-transparent", "trns"])
def test_generic_paletted_path_equals_typed_path_on_expanded_raster(oracle, kind):
    O = oracle
    rng = np.random.default_rng({"opaque": 1, "gif-transparent": 2, "trns": 3}[kind])
    for (w, h, dw, dh, size) in [(160, 120, 64, 48, 50), (333, 222, 1024, 682, 200), (97, 131, 31, 57, 16)]:
        pal = _palette(rng, kind)
        idx = rng.integers(0, 256, (h, w), dtype=np.uint8)
        nrgba = kind == "trns"
        P = O.Raster.paletted(idx, pal, nrgba)
        E = O.Raster.rgba(pal[idx], O.NRGBA8 if nrgba else O.RGBA8)
        assert np.array_equal(O.resize_image(P, dw, dh), O.resize_image(E, dw, dh))
        assert np.array_equal(O.crop_and_resize(P, size), O.crop_and_resize(E, size))
        assert np.array_equal(O.draw_src(P), O.draw_src(E))
        img = codecs.expand_paletted(idx, pal) if not nrgba else None
        if img is not None:                  # the product-side helper picks the layout from the palette itself
            assert img.layout == ip.RGBA8 and np.array_equal(img.planes[0], pal[idx])
    img = codecs.expand_paletted(np.zeros((2, 2), np.uint8), np.array([[10, 20, 30, 128]], np.uint8))
    assert img.layout == ip.NRGBA8           # a translucent entry can only be color.NRGBA (PNG tRNS)


def test_decode_gif_and_paletted_png_and_sixteen_bit_png():
    from PIL import Image as PI
    import cv2
    rgb = codecs.synth_picture(64, 40, 4)
    buf = io.BytesIO()
    PI.fromarray(rgb, "RGB").quantize(16).save(buf, "PNG")
    img, fmt = codecs.decode(buf.getvalue())
    assert fmt == "png" and img.layout == ip.RGBA8 and img.planes[0].shape == (40, 64, 4)
    g16 = (np.arange(40 * 64, dtype=np.uint32).reshape(40, 64) * 16 % 65536).astype(np.uint16)
    ok, enc = cv2.imencode(".png", g16)
    img, fmt = codecs.decode(enc.tobytes())
    assert fmt == "png" and img.layout == ip.GRAY16 and img.planes[0].shape == (40, 128)
    assert img.planes[0][0, 2] == (g16[0, 1] >> 8) and img.planes[0][0, 3] == (g16[0, 1] & 0xFF)    # big-endian, as Go's Pix
    bgra = np.random.default_rng(1).integers(0, 65536, (40, 64, 4)).astype(np.uint16)
    ok, enc = cv2.imencode(".png", bgra)
    img, fmt = codecs.decode(enc.tobytes())
    assert img.layout == ip.NRGBA64 and img.planes[0].shape == (40, 512)
    ok, enc = cv2.imencode(".png", bgra[..., :3])
    img, fmt = codecs.decode(enc.tobytes())
    assert img.layout == ip.RGBA64 and img.planes[0][0, 6] == 0xFF and img.planes[0][0, 7] == 0xFF


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["gif-transparent", "trns"])
def test_engine_on_expanded_palette_matches_generic_oracle_path(engines, oracle, kind):
    O = oracle
    rng = np.random.default_rng(8)
    w, h = 1200, 900
    pal = _palette(rng, kind)
    idx = rng.integers(0, 256, (h, w), dtype=np.uint8)
    nrgba = kind == "trns"
    P = O.Raster.paletted(idx, pal, nrgba)
    img = ip.Image.from_rgba(np.ascontiguousarray(pal[idx]), ip.NRGBA8 if nrgba else ip.RGBA8)
    nw, nh = ip.keep_aspect_dims(w, h, 1024, 768)
    cx, cy, cs = ip.crop_square(w, h)
    gl = synthetic_glyphs(w, h, 4)
    col = (255, 255, 255, 127)
    out = engines(ip.PRECISION_EXACT).run(img, [ip.OpSpec.resize(nw, nh), ip.OpSpec.thumb_crop((cx, cy, cs, cs), 200),
                                                ip.OpSpec.watermark(w, h, col, [ip.GlyphMask(*g) for g in gl])])
    assert np.array_equal(out[0], O.resize_image(P, nw, nh))
    assert np.array_equal(out[1], O.crop_and_resize(P, 200))
    assert np.array_equal(out[2], O.watermark(P, col, [O.Glyph(*g) for g in gl]))


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["rgba64", "nrgba64", "gray16"])
def test_sixteen_bit_layouts(engines, oracle, layout):
    """16-bit PNG types through the engine: downscales run whole-image in float64, upscales in the fp32 k_direct with
    its certificate; the crop thumbnail passes the 8-bit crop stage; the watermark frame is uint8(RGBA() >> 8)."""
    O = oracle
    rng = np.random.default_rng({"rgba64": 1, "nrgba64": 2, "gray16": 3}[layout])
    e = engines(ip.PRECISION_EXACT)
    col = (20, 250, 130, 180)
    for (w, h, dw, dh) in [(640, 480, 256, 192), (300, 200, 1024, 682), (801, 603, 801, 603), (64, 512, 33, 700)]:
        if layout == "gray16":
            a = rng.integers(0, 65536, (h, w)).astype(np.uint16)
        else:
            a = rng.integers(0, 65536, (h, w, 4)).astype(np.uint16)
            if layout == "rgba64":           # valid premultiplied: c <= a
                a[..., :3] = (a[..., :3].astype(np.uint64) * a[..., 3:4] // 65535).astype(np.uint16)
        lay, olay = {"rgba64": (ip.RGBA64, O.RGBA64), "nrgba64": (ip.NRGBA64, O.NRGBA64), "gray16": (ip.GRAY16, O.GRAY16)}[layout]
        img, R = ip.Image.from_deep(a, lay), O.Raster.deep(a, olay)
        cx, cy, cs = ip.crop_square(w, h)
        gl = synthetic_glyphs(w, h, 6, n=3)
        out = e.run(img, [ip.OpSpec.resize(dw, dh), ip.OpSpec.thumb_crop((cx, cy, cs, cs), 48),
                          ip.OpSpec.watermark(w, h, col, [ip.GlyphMask(*g) for g in gl])])
        assert np.array_equal(out[0], O.resize_image(R, dw, dh)), f"{layout} resize {w}x{h} -> {dw}x{dh}"
        assert np.array_equal(out[1], O.crop_and_resize(R, 48)), f"{layout} thumb {w}x{h}"
        assert np.array_equal(out[2], O.watermark(R, col, [O.Glyph(*g) for g in gl])), f"{layout} watermark {w}x{h}"
