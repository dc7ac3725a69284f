"""Device-side JPEG writer (ipg_op.dst_layout = IPG_LAYOUT_JPEG) against the oracle's restatement of Go's
jpeg.Encode(buf, result, &jpeg.Options{Quality: 85}) -- operations/resize.go:78-91, watermark.go:66-79.

The file the engine returns must equal, byte for byte, the oracle's encoding of the oracle's raster result: the raster
kernels are bit-exact and the writer is integer arithmetic, so there is no tolerance anywhere."""
import io

import numpy as np
import pytest
from PIL import Image as PILImage

import imageprocessor_b200 as ip
from tests.test_jpeg_oracle import smooth
from tests.util import rgba_random, synthetic_glyphs

pytestmark = pytest.mark.gpu


def photo(rng, h, w, amp=10.0):
    return np.dstack([smooth(rng, h, w, amp) for _ in range(3)] + [np.full((h, w), 255, np.uint8)])


def three_ops(w, h, gl, col, **kw):
    nw, nh = ip.keep_aspect_dims(w, h, 1024, 768)
    cx, cy, cs = ip.crop_square(w, h)
    return [ip.OpSpec.resize(nw, nh, **kw), ip.OpSpec.thumb_crop((cx, cy, cs, cs), 200, **kw),
            ip.OpSpec.watermark(w, h, col, [ip.GlyphMask(*g) for g in gl], **kw)]


def oracle_files(O, a, gl, col, quality=85):
    h, w = a.shape[:2]
    R = O.Raster.rgba(a)
    nw, nh = ip.keep_aspect_dims(w, h, 1024, 768)
    og = [O.Glyph(g.x0, g.y0, g.x1, g.y1, g.mask, g.mp_x, g.mp_y) for g in gl]
    return [O.jpeg_encode_rgba(O.resize_image(R, nw, nh), quality), O.jpeg_encode_rgba(O.crop_and_resize(R, 200), quality),
            O.jpeg_encode_rgba(O.watermark(R, col, og), quality)]


@pytest.mark.parametrize("size", [(1600, 1200), (1999, 1201), (640, 480), (333, 517)])
def test_three_results_as_go_jpeg_files(engines, oracle, size):
    w, h = size
    a = photo(np.random.default_rng(w + h), h, w)
    gl, col = synthetic_glyphs(w, h, 3), (255, 255, 255, 127)
    out = engines().run(ip.Image.from_rgba(a), three_ops(w, h, gl, col, jpeg_quality=85))
    want = oracle_files(oracle, a, gl, col)
    for k, name in enumerate(("resize", "thumbnail", "watermark")):
        assert out[k].nbytes == len(want[k]), f"{name}: {out[k].nbytes} bytes, the oracle's file has {len(want[k])}"
        assert out[k].data == want[k], f"{name} file differs from jpeg.Encode of the oracle result"
    img = np.asarray(PILImage.open(io.BytesIO(out[2].data)).convert("RGB")).astype(np.int32)
    assert img.shape[:2] == (h, w) and np.abs(img - a[..., :3]).mean() < 12  # and a stock decoder reads it back


@pytest.mark.parametrize("quality", [100, 50, 1])
def test_noise_every_quality_and_heavy_stuffing(engines, oracle, quality):
    """Random bytes: the largest scans (thousands of 0xff bytes to stuff at quality 100), every run / size code."""
    w, h = 517, 389
    a = rgba_random(w, h, 11)
    res = engines().run(ip.Image.from_rgba(a), [ip.OpSpec.watermark(w, h, (0, 0, 0, 255), [], jpeg_quality=quality, jpeg_capacity=w * h * 4)])[0]
    want = oracle.jpeg_encode_rgba(a, quality)
    assert res.data == want
    if quality == 100:
        assert want.count(b"\xff\x00") > 50


def test_tiny_and_mcu_edge_sizes(engines, oracle):
    e = engines()
    rng = np.random.default_rng(5)
    for w, h in ((1, 1), (15, 17), (16, 16), (17, 16), (31, 33), (48, 8), (8, 48)):
        a = photo(rng, h, w)
        res = e.run(ip.Image.from_rgba(a), [ip.OpSpec.watermark(w, h, (0, 0, 0, 255), [], jpeg_quality=85, jpeg_capacity=8192)])[0]
        assert res.data == oracle.jpeg_encode_rgba(a, 85), (w, h)


def test_widest_image_the_writer_accepts_and_the_first_it_refuses(engines, oracle):
    """65535 pixels per side is the most Go's writer takes; 65536 is refused with its own message before anything runs."""
    e = engines()
    rng = np.random.default_rng(6)
    a = np.repeat(rng.integers(0, 256, (16, 65535 // 15 + 1, 4), dtype=np.uint8), 15, axis=1)[:, :65535].copy()
    a[..., 3] = 255
    res = e.run(ip.Image.from_rgba(a), [ip.OpSpec.watermark(65535, 16, (0, 0, 0, 255), [], jpeg_quality=85, jpeg_capacity=65535 * 16 * 2)])[0]
    assert res.data == oracle.jpeg_encode_rgba(a, 85)
    b = np.zeros((8, 65536, 4), np.uint8)
    with pytest.raises(ip.IpgError) as ei:
        e.submit(ip.Image.from_rgba(b), [ip.OpSpec.watermark(65536, 8, (0, 0, 0, 255), [], jpeg_quality=85)])
    assert "too large to encode" in ei.value.message


def test_alpha_results_encode_their_premultiplied_bytes(engines, oracle):
    """A PNG with alpha resized and written as JPEG: Go's writer reads the premultiplied R, G, B of the *image.RGBA."""
    w, h = 1203, 907
    a = rgba_random(w, h, 21, alpha="premul")
    nw, nh = ip.keep_aspect_dims(w, h, 1024, 768)
    res = engines().run(ip.Image.from_rgba(a), [ip.OpSpec.resize(nw, nh, jpeg_quality=85, jpeg_capacity=nw * nh * 3)])[0]
    assert res.data == oracle.jpeg_encode_rgba(oracle.resize_image(oracle.Raster.rgba(a), nw, nh), 85)


def test_jpeg_beside_rgba_and_planar_results_in_one_ticket(engines, oracle):
    """Mixed destinations in one submission, and a YCbCr 4:2:0 source (what a decoded JPEG is)."""
    w, h = 1280, 960
    rng = np.random.default_rng(8)
    y, cb, cr = smooth(rng, h, w, 8.0), smooth(rng, h // 2, w // 2, 4.0), smooth(rng, h // 2, w // 2, 4.0)
    src = ip.Image.from_ycbcr(y, cb, cr, ip.YCBCR420)
    nw, nh = ip.keep_aspect_dims(w, h, 1024, 768)
    cx, cy, cs = ip.crop_square(w, h)
    out = engines().run(src, [ip.OpSpec.resize(nw, nh, jpeg_quality=85), ip.OpSpec.thumb_crop((cx, cy, cs, cs), 200),
                              ip.OpSpec.thumb_crop((cx, cy, cs, cs), 200, jpeg_quality=85)])
    R = oracle.Raster.ycbcr(y, cb, cr, oracle.YCBCR420)
    assert out[0].data == oracle.jpeg_encode_rgba(oracle.resize_image(R, nw, nh), 85)
    t = oracle.crop_and_resize(R, 200)
    assert np.array_equal(out[1], t) and out[2].data == oracle.jpeg_encode_rgba(t, 85)


def test_capacity_too_small_fails_that_ticket_only(engines, oracle):
    e = engines()
    w, h = 640, 480
    a = rgba_random(w, h, 2)
    small = e.submit(ip.Image.from_rgba(a), [ip.OpSpec.watermark(w, h, (0, 0, 0, 255), [], jpeg_quality=85, jpeg_capacity=4096)])
    good = e.submit(ip.Image.from_rgba(a), [ip.OpSpec.watermark(w, h, (0, 0, 0, 255), [], jpeg_quality=85, jpeg_capacity=w * h * 3)])
    with pytest.raises(ip.IpgError) as ei:
        e.wait(small)
    assert ei.value.code == ip._lib.ERR_NOMEM and "JPEG" in ei.value.message
    assert e.wait(good)[0].data == oracle.jpeg_encode_rgba(a, 85)


def test_pinned_destination_and_many_tickets_in_flight(engines, oracle):
    """The shape the worker uses: pinned file buffers, several tickets coalesced into one batch."""
    e = engines()
    w, h = 800, 600
    rng = np.random.default_rng(13)
    imgs = [photo(rng, h, w) for _ in range(6)]
    bufs = [e.alloc_pinned(w * h) for _ in imgs]
    ts = [e.submit(ip.Image.from_rgba(a), [ip.OpSpec.watermark(w, h, (0, 0, 0, 255), [], jpeg_quality=85, jpeg_buffer=b.array)])
          for a, b in zip(imgs, bufs)]
    for a, t in zip(imgs, ts):
        assert e.wait(t)[0].data == oracle.jpeg_encode_rgba(a, 85)
    for b in bufs:
        b.free()


def test_twelve_megapixel_watermark_file(engines, oracle):
    """BASELINE's 12 MP frame (4000 x 3000: 375 block rows, so the last MCU row is half padding): all three files."""
    w, h = 4000, 3000
    a = photo(np.random.default_rng(1), h, w, 6.0)
    gl, col = synthetic_glyphs(w, h, 3), (255, 255, 255, 127)
    out = engines(lane_device_bytes=2 << 30).run(ip.Image.from_rgba(a), three_ops(w, h, gl, col, jpeg_quality=85))
    want = oracle_files(oracle, a, gl, col)
    for k in range(3):
        assert out[k].data == want[k]


def test_pageable_destination_is_refused(engines):
    """The engine writes the file and its length after ipg_submit returned: it takes them in its own pinned memory only."""
    w, h = 64, 48
    a = rgba_random(w, h, 3)
    buf = np.zeros(1 << 16, np.uint8)   # ordinary heap memory
    with pytest.raises(ip.IpgError) as ei:
        engines().submit(ip.Image.from_rgba(a), [ip.OpSpec.watermark(w, h, (0, 0, 0, 255), [], jpeg_quality=85, jpeg_buffer=buf)])
    assert ei.value.code == ip._lib.ERR_INVALID and "ipg_alloc_pinned" in ei.value.message
