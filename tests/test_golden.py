"""Known-answer tests against tests/golden/golden_v1.npz.

The fixtures were computed by code independent of both oracle/ and the CUDA path
(PyTorch's antialiased bilinear in float64 + the x/image quantiser; Python big-int
drawGlyphOver; hand-derived geometry) -- see tests/golden/make_golden.py for what they
pin and what they cannot (the Go binary itself: parity with it stays unpinned).

  not gpu : the oracle reproduces every fixture byte for byte
  gpu     : so does libipgpu.so through the C ABI (EXACT mode)
"""
import os

import numpy as np
import pytest

from tests.golden.cases import (RESAMPLE_CASES, BLEND_CASES, make_source, make_blend_case)

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz"))


def _oracle_raster(O, kind, planes):
    if kind in ("rgba", "rgba_premul"):
        return O.Raster.rgba(planes[0])
    if kind == "nrgba":
        return O.Raster.rgba(planes[0], O.NRGBA8)
    if kind == "gray":
        return O.Raster.gray(planes[0])
    lay = {"ycbcr444": O.YCBCR444, "ycbcr422": O.YCBCR422, "ycbcr420": O.YCBCR420, "ycbcr440": O.YCBCR440}[kind]
    return O.Raster.ycbcr(*planes, lay)


def _ip_image(ip, kind, planes):
    if kind in ("rgba", "rgba_premul"):
        return ip.Image.from_rgba(planes[0])
    if kind == "nrgba":
        return ip.Image.from_rgba(planes[0], ip.NRGBA8)
    if kind == "gray":
        return ip.Image.from_gray(planes[0])
    lay = {"ycbcr444": ip.YCBCR444, "ycbcr422": ip.YCBCR422, "ycbcr420": ip.YCBCR420, "ycbcr440": ip.YCBCR440}[kind]
    return ip.Image.from_ycbcr(*planes, lay)


@pytest.mark.parametrize("case", RESAMPLE_CASES, ids=[c[0] for c in RESAMPLE_CASES])
def test_oracle_resample_matches_golden(oracle, case):
    name, kind, w, h, seed, ops = case
    R = _oracle_raster(oracle, kind, make_source(kind, w, h, seed))
    for op in ops:
        if op[0] == "resize":
            got, want = oracle.resize_image(R, op[1], op[2]), GOLD[f"{name}/resize_{op[1]}x{op[2]}"]
        else:
            got, want = oracle.crop_and_resize(R, op[1]), GOLD[f"{name}/thumb_{op[1]}"]
        assert np.array_equal(got, want), f"{name} {op}: {(got != want).sum()} bytes differ"


@pytest.mark.parametrize("case", BLEND_CASES, ids=[c[0] for c in BLEND_CASES])
def test_oracle_blend_matches_golden(oracle, case):
    name, w, h, seed, color, n = case
    dst, glyphs = make_blend_case(w, h, seed, n)
    got = oracle.watermark(oracle.Raster.rgba(dst), color, [oracle.Glyph(*g) for g in glyphs])
    assert np.array_equal(got, GOLD[f"{name}/blend"])


def test_oracle_blend_table_white127(oracle):
    """All 65,536 (dst byte, mask byte) pairs for the default colour (SURVEY.md 8c iii)."""
    tab = GOLD["blend_table_white127"]
    dst = np.repeat(np.arange(256, dtype=np.uint8)[:, None, None], 256, axis=1).repeat(4, axis=2).copy()
    mask = np.repeat(np.arange(256, dtype=np.uint8)[None, :], 256, axis=0).copy()
    got = oracle.watermark(oracle.Raster.rgba(dst), (255, 255, 255, 127), [oracle.Glyph(0, 0, 256, 256, mask)])
    assert np.array_equal(got, tab)
    # spot values quoted in SURVEY.md Spec W (computed there from the formula by hand)
    for m, rgb, a in zip((1, 64, 128, 192, 255), (1, 64, 128, 192, 255), (0, 31, 63, 95, 127)):
        assert tuple(tab[0, m]) == (rgb, rgb, rgb, a)
    for m, rgb in zip((1, 64, 128, 192, 255), (0, 32, 64, 96, 128)):
        assert tuple(tab[255, m][:3]) == (rgb, rgb, rgb) and tab[255, m][3] == 255
    assert [int(tab[200, m][0]) for m in (1, 64, 128, 192, 255)] == [201, 239, 23, 62, 100]


def test_geometry_known_answers(oracle):
    """resize.go:63-72 / thumbnail.go:52-63,115-127 worked by hand (SURVEY.md 8c iv)."""
    ka = oracle.keep_aspect_dims
    assert ka(4000, 3000, 1024, 768) == (1024, 768)
    assert ka(7680, 4320, 1024, 768) == (1024, 576)
    assert ka(3000, 4000, 1024, 768) == (576, 768)
    assert ka(1002, 751, 1024, 768) == (1023, 767)      # double truncation, not rounding
    assert ka(147, 147, 1024, 768) == (767, 767)
    assert oracle.thumb_fit_dims(4000, 3000, 200) == (266, 200)
    assert oracle.thumb_fit_dims(3000, 4000, 200) == (200, 266)
    assert oracle.crop_square(4000, 3000) == (500, 0, 3000)
    assert oracle.crop_square(3000, 4001) == (0, 500, 3000)
    assert oracle.crop_square(7680, 4320) == (1680, 0, 4320)
    assert oracle.watermark_height_px(36.0) == 44
    assert oracle.parse_color("255,255,255", 0.5) == (0, (255, 255, 255, 127))


# ---- the CUDA path against the same fixtures ------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("case", RESAMPLE_CASES, ids=[c[0] for c in RESAMPLE_CASES])
def test_cuda_resample_matches_golden(engines, case):
    import imageprocessor_b200 as ip
    name, kind, w, h, seed, ops = case
    img = _ip_image(ip, kind, make_source(kind, w, h, seed))
    specs, want = [], []
    for op in ops:
        if op[0] == "resize":
            specs.append(ip.OpSpec.resize(op[1], op[2]))
            want.append(GOLD[f"{name}/resize_{op[1]}x{op[2]}"])
        else:
            cx, cy, cs = ip.crop_square(w, h)
            specs.append(ip.OpSpec.thumb_crop((cx, cy, cs, cs), op[1]))
            want.append(GOLD[f"{name}/thumb_{op[1]}"])
    out = engines(ip.PRECISION_EXACT).run(img, specs)
    for o, wv, op in zip(out, want, ops):
        assert np.array_equal(o, wv), f"{name} {op}: {(o != wv).sum()} bytes differ"


@pytest.mark.gpu
@pytest.mark.parametrize("case", BLEND_CASES, ids=[c[0] for c in BLEND_CASES])
def test_cuda_blend_matches_golden(engines, case):
    import imageprocessor_b200 as ip
    name, w, h, seed, color, n = case
    dst, glyphs = make_blend_case(w, h, seed, n)
    out = engines(ip.PRECISION_EXACT).run(ip.Image.from_rgba(dst), [
        ip.OpSpec.watermark(w, h, color, [ip.GlyphMask(*g) for g in glyphs])])
    assert np.array_equal(out[0], GOLD[f"{name}/blend"])


@pytest.mark.gpu
def test_cuda_blend_table_white127(engines):
    import imageprocessor_b200 as ip
    dst = np.repeat(np.arange(256, dtype=np.uint8)[:, None, None], 256, axis=1).repeat(4, axis=2).copy()
    mask = np.repeat(np.arange(256, dtype=np.uint8)[None, :], 256, axis=0).copy()
    out = engines(ip.PRECISION_EXACT).run(ip.Image.from_rgba(dst), [
        ip.OpSpec.watermark(256, 256, (255, 255, 255, 127), [ip.GlyphMask(0, 0, 256, 256, mask)])])
    assert np.array_equal(out[0], GOLD["blend_table_white127"])
