"""Pins oracle/ip_jpeg_oracle.c (the restatement of Go's image/jpeg writer) with an independent implementation.

libjpeg-turbo (through PIL) implements the same baseline process: Annex K quantisation tables scaled by the same quality
rule, the jfdctint forward DCT Go's fdct.go translates, round-half-away quantisation and the Annex K Huffman tables.  What
differs is outside the coefficient path -- libjpeg's own colour conversion, its h2v2 chroma downsample (alternating
bias where Go adds 2) and its file header (JFIF APP0, one DQT / DHT segment per table).  So the comparison feeds both
the same Y, Cb, Cr samples with chroma constant over every 2 x 2 group (then both downsamplers return that constant) and
requires the entropy-coded segment and the table payloads to be equal byte for byte.

One more genuine difference bounds the sizes compared: when the image is an odd number of 8 x 8 luma blocks wide or high,
libjpeg fills the missing block of the last MCU with a "dummy block" (DC of its neighbour, no AC: jccoefct.c), whereas Go
builds it from edge-replicated PIXELS like any other block (writer.go rgbaToYCbCr clamps sx, sy).  Partial blocks inside
the image are edge-replicated by both.  So the colour cases use sizes with an even block count per axis; the odd ones
(200 x 200, 4000 x 3000: 375 block rows) are covered by the oracle's own structure and by the decode test below.
"""
import io

import numpy as np
import pytest
from PIL import Image

from oracle import oracle as O


def segments(data: bytes):
    """[(marker, payload)] of a JPEG file; the SOS entry's payload is header + entropy-coded data up to EOI."""
    assert data[:2] == b"\xff\xd8"
    out, i = [], 2
    while i < len(data):
        assert data[i] == 0xFF, hex(i)
        m = data[i + 1]
        if m == 0xD9:
            out.append((m, b""))
            break
        n = (data[i + 2] << 8) | data[i + 3]
        if m == 0xDA:
            assert data[-2:] == b"\xff\xd9"
            out.append((m, data[i + 2:-2]))
            out.append((0xD9, b""))
            break
        out.append((m, data[i + 4:i + 2 + n]))
        i += 2 + n
    return out


def tables(segs, marker, unit):
    """DQT / DHT tables keyed by their id byte, whether the file holds one segment per table or one for all."""
    t = {}
    for m, p in segs:
        if m != marker:
            continue
        i = 0
        while i < len(p):
            n = unit(p, i)
            t[p[i]] = p[i + 1:i + n]
            i += n
    return t


def dqt_unit(p, i):
    assert p[i] >> 4 == 0  # 8-bit tables
    return 65


def dht_unit(p, i):
    return 17 + sum(p[i + 1:i + 17])


def scan(segs):
    p = [p for m, p in segs if m == 0xDA][0]
    n = (p[0] << 8) | p[1]
    return p[:n], p[n:]


def pil_jpeg(img: Image.Image, quality: int, subsampling: int) -> bytes:
    b = io.BytesIO()
    img.save(b, "JPEG", quality=quality, subsampling=subsampling, optimize=False, progressive=False)
    return b.getvalue()


def smooth(rng, h, w, amp=40.0):
    """Photo-like plane: low-frequency gradients plus noise (so that runs of zeros, ZRL and EOB codes all occur)."""
    y, x = np.mgrid[0:h, 0:w]
    v = 128 + 90 * np.sin(x / 37.0 + rng.uniform(0, 6)) * np.cos(y / 23.0 + rng.uniform(0, 6)) + rng.normal(0, amp, (h, w))
    return np.clip(v, 0, 255).astype(np.uint8)


@pytest.mark.parametrize("quality", [85, 50, 20, 100, 1])
@pytest.mark.parametrize("size", [(64, 48), (250, 122), (123, 77), (16, 16), (9, 9), (1024, 768)])
def test_color_scan_equals_libjpeg_turbo(size, quality):
    w, h = size
    rng = np.random.default_rng(w * 1000 + h + quality)
    cw, ch = (w + 1) // 2, (h + 1) // 2
    y = smooth(rng, h, w) if quality != 100 else rng.integers(0, 256, (h, w), dtype=np.uint8)
    cb, cr = smooth(rng, ch, cw, 15.0), smooth(rng, ch, cw, 15.0)
    ours = O.jpeg_encode_ycbcr(O.Raster.ycbcr(y, cb, cr, O.YCBCR420), quality)
    up = lambda c: np.repeat(np.repeat(c, 2, 0), 2, 1)[:h, :w]
    theirs = pil_jpeg(Image.merge("YCbCr", [Image.fromarray(p) for p in (y, up(cb), up(cr))]), quality, 2)
    so, st = segments(ours), segments(theirs)
    assert tables(so, 0xDB, dqt_unit) == tables(st, 0xDB, dqt_unit)
    assert tables(so, 0xC4, dht_unit) == tables(st, 0xC4, dht_unit)
    sof_o = [p for m, p in so if m == 0xC0][0]
    sof_t = [p for m, p in st if m == 0xC0][0]
    assert sof_o == sof_t  # precision, size, 3 components 0x22 / 0x11 / 0x11, table selectors 0 / 1 / 1
    ho, eo = scan(so)
    ht, et = scan(st)
    assert ho == ht
    assert eo == et, f"entropy-coded segment differs ({len(eo)} vs {len(et)} bytes)"


@pytest.mark.parametrize("quality", [85, 30])
@pytest.mark.parametrize("size", [(64, 48), (41, 23), (8, 8), (3, 5)])
def test_gray_scan_equals_libjpeg_turbo(size, quality):
    w, h = size
    rng = np.random.default_rng(w * 77 + h + quality)
    g = smooth(rng, h, w)
    so, st = segments(O.jpeg_encode_gray(g, quality)), segments(pil_jpeg(Image.fromarray(g), quality, 0))
    assert tables(so, 0xDB, dqt_unit)[0] == tables(st, 0xDB, dqt_unit)[0]
    assert scan(so) == scan(st)


def test_file_layout_is_the_go_writers():
    """SOI, one DQT with both tables, SOF0, one DHT with the four tables, SOS, data, EOI -- and nothing else
    (Go's writer emits no JFIF / APPn segment)."""
    rng = np.random.default_rng(3)
    rgba = np.dstack([smooth(rng, 40, 56, 5.0) for _ in range(4)])
    data = O.jpeg_encode_rgba(rgba, 85)
    segs = segments(data)
    assert [m for m, _ in segs] == [0xDB, 0xC0, 0xC4, 0xDA, 0xD9]
    assert len(segs[0][1]) == 2 * 65 and len(segs[2][1]) == 2 * (17 + 12) + 2 * (17 + 162)
    assert data[2:6] == b"\xff\xdb\x00\x84" and segs[1][1] == bytes([8, 0, 40, 0, 56, 3, 1, 0x22, 0, 2, 0x11, 1, 3, 0x11, 1])
    img = np.asarray(Image.open(io.BytesIO(data)).convert("RGB")).astype(np.int32)
    assert np.abs(img - rgba[..., :3]).mean() < 8  # and it decodes to the image


def test_rgba_path_equals_ycbcr420_path_without_mcu_padding():
    """On sizes that are whole MCUs jpeg.Encode(RGBA) == jpeg.Encode(the 4:2:0 image ipo_rgba_to_ycbcr420 derives): pins the
    RGBA entry (rgbaToYCbCr + scale) to the YCbCr entry the libjpeg comparison covers.  With MCU padding they differ by
    construction (the RGBA path replicates the edge PIXEL's chroma, the YCbCr path the edge chroma SAMPLE) -- shown too."""
    rng = np.random.default_rng(9)
    for w, h in ((64, 48), (16, 16), (160, 32)):
        rgba = np.dstack([smooth(rng, h, w) for _ in range(4)])
        y, cb, cr = O.rgba_to_ycbcr420(rgba)
        assert O.jpeg_encode_rgba(rgba, 85) == O.jpeg_encode_ycbcr(O.Raster.ycbcr(y, cb, cr, O.YCBCR420), 85)
    rgba = np.dstack([smooth(rng, 24, 40) for _ in range(4)])  # 24 rows: the last MCU row is half padding
    y, cb, cr = O.rgba_to_ycbcr420(rgba)
    assert O.jpeg_encode_rgba(rgba, 85) != O.jpeg_encode_ycbcr(O.Raster.ycbcr(y, cb, cr, O.YCBCR420), 85)


def test_decoded_quality_matches_pil_encoder():
    """End to end sanity on a photo-like RGBA image: decoded PSNR within 0.5 dB of PIL's own q85 4:2:0 file."""
    rng = np.random.default_rng(5)
    rgb = np.dstack([smooth(rng, 240, 320, 6.0) for _ in range(3)])
    rgba = np.dstack([rgb, np.full((240, 320), 255, np.uint8)])
    ours = np.asarray(Image.open(io.BytesIO(O.jpeg_encode_rgba(rgba, 85))).convert("RGB")).astype(np.float64)
    theirs = np.asarray(Image.open(io.BytesIO(pil_jpeg(Image.fromarray(rgb), 85, 2))).convert("RGB")).astype(np.float64)
    psnr = lambda a: 10 * np.log10(255 ** 2 / np.mean((a - rgb) ** 2))
    assert abs(psnr(ours) - psnr(theirs)) < 0.5 and psnr(ours) > 30
