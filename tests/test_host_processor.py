"""The host layer (include/ipgpu_host.h) against the reference's own behaviour:
processor.ImageProcessor.Process (internal/usecase/processor/image_processor.go:39-182),
the three operations' parameter handling (operations/resize.go:26-59, thumbnail.go:25-46,
watermark.go:40-60,159-190) and the task/result schema (internal/domain/task.go:3-23).

not gpu: everything that happens before/after the raster call.
gpu    : Process end to end; stored objects compared byte for byte with the oracle.
"""
import json

import numpy as np
import pytest

from imageprocessor_b200 import processor as P
from tests.util import rgba_random

CANONICAL_OPS = [   # internal/http-server/handler/image/image.go:222-277
    {"Type": "thumbnail", "Parameters": {"size": 200, "crop_to_fit": True}},
    {"Type": "resize", "Parameters": {"width": 1024, "height": 768, "keep_aspect": True}},
    {"Type": "watermark", "Parameters": {"text": "© ImageProcessor", "opacity": 0.5, "position": "bottom-right"}},
]


def task(ops, fmt="", image_id="img-1"):
    # usecase/image/image.go:85-93: no json tags, so the keys are the Go field names
    return {"ID": "task-1", "ImageID": image_id, "OriginalPath": "original/2024/01/01/1.jpg", "Bucket": "images",
            "Operations": ops, "Format": fmt}


# ---- host-only ------------------------------------------------------------------------
def test_parse_color_matches_reference_rules():
    assert P.parse_color("255,255,255", 0.5) == (0, (255, 255, 255, 127))      # uint8(255*0.5) truncates
    assert P.parse_color(" 10, 300 , -5 ", 1.0) == (0, (10, 255, 0, 255))      # spaces stripped, clamped
    assert P.parse_color("1,2,3,40", 0.5) == (0, (1, 2, 3, 40))                # explicit alpha wins
    assert P.parse_color("1,2,3,x", 0.25) == (0, (1, 2, 3, 63))                # bad alpha -> opacity
    assert P.parse_color("red", 0.5) == (-1, (0, 0, 0, 127))                   # caller falls back to black
    assert P.parse_color("1,2", 0.5)[0] == -1 and P.parse_color("1,2,3,4,5", 0.5)[0] == -1
    assert P.parse_color("1.5,2,3", 0.5)[0] == -1                              # strconv.Atoi rejects floats


def test_watermark_geometry():
    assert P.watermark_height_px(36) == 44 and P.watermark_height_px(12) == 15
    W, H, w, h = 4000, 3000, 300, 44
    assert P.watermark_anchor("bottom-right", W, H, w, h) == (3680, 2980)
    assert P.watermark_anchor("top-left", W, H, w, h) == (20, 64)
    assert P.watermark_anchor("top-right", W, H, w, h) == (3680, 64)
    assert P.watermark_anchor("top-center", W, H, w, h) == (1850, 64)
    assert P.watermark_anchor("bottom-left", W, H, w, h) == (20, 2980)
    assert P.watermark_anchor("bottom-center", W, H, w, h) == (1850, 2980)
    assert P.watermark_anchor("center", W, H, w, h) == (1850, 1522)
    assert P.watermark_anchor("nonsense", W, H, w, h) == (3680, 2980)
    assert P.watermark_anchor("center", 100, 30, 301, 44) == (-100, 37)         # Go '/' truncates toward zero


def test_generate_path_and_content_type():
    gp = P.generate_path
    assert gp("abc", "resize", "jpeg", {"width": 1024, "height": 768}) == "processed/resize/abc/1024x768.jpeg"
    assert gp("abc", "resize", "png", {}) == "processed/resize/abc/0x0.png"                 # requested, not actual, dims
    assert gp("abc", "thumbnail", "jpeg", {"size": 150}) == "processed/thumbnails/abc/150.jpeg"
    assert gp("abc", "thumbnail", "gif", {}) == "processed/thumbnails/abc/200.gif"          # DefaultThumbnailSize
    assert gp("abc", "watermark", "jpeg", {}) == "processed/watermarked/abc/watermarked.jpeg"
    assert gp("abc", "Rotate", "png", {}) == "processed/rotate/abc/processed.png"
    ct = P.content_type
    assert ct("a/b.jpg") == ct("a/b.JPEG") == "image/jpeg" and ct("x.png") == "image/png" and ct("x.gif") == "image/gif"
    assert ct("x.webp") == "image/webp" and ct("x.bmp") == "image/bmp" and ct("x.tif") == "image/tiff"
    assert ct("noext") == "image/jpeg" and ct("dir.png/file") == "image/jpeg"


def test_process_parameter_errors_without_a_device():
    """Errors raised before any raster work: same strings, same partial result as the reference."""
    from imageprocessor_b200 import Image
    repo = P.MemoryFileRepo()
    ip_ = P.ImageProcessor(None, repo)
    img = Image.from_rgba(rgba_random(64, 48, 1))
    res, err = ip_.process(task([{"Type": "resize", "Parameters": {"height": 10}}]), img)
    assert err == "operation resize failed: failed to process operation resize: width parameter is required and must be a number"
    assert res == {"ID": "task-1", "ImageID": "img-1", "Status": "failed", "ProcessedPaths": {},
                   "Error": "Operation resize failed: failed to process operation resize: width parameter is required and must be a number"}
    res, err = ip_.process(task([{"Type": "resize", "Parameters": {"width": 10, "height": 0}}]), img)
    assert err.endswith("width and height must be positive numbers")
    res, err = ip_.process(task([{"Type": "thumbnail", "Parameters": {"size": -3}}]), img)
    assert err.endswith("size must be a positive number")
    res, err = ip_.process(task([{"Type": "rotate", "Parameters": {}}]), img)
    assert err == "operation rotate failed: unsupported operation type: rotate"          # not wrapped by applyOperation
    assert res["Error"] == "Operation rotate failed: unsupported operation type: rotate"
    res, err = ip_.process(task(CANONICAL_OPS), None, decode_error="image: unknown format")
    assert err == "failed to decode image: image: unknown format" and res["Status"] == "failed"
    assert res["Error"] == "Failed to decode image: image: unknown format"
    res, err = ip_.process("{not json", img)
    assert err.startswith("failed to unmarshal task")
    res, err = ip_.process(task([]), img)                                             # no operations: completed, nothing saved
    assert err is None and res["Status"] == "completed" and res["ProcessedPaths"] == {}
    # raster work without an engine must fail loudly, never fall back
    res, err = ip_.process(task(CANONICAL_OPS[:1]), img)
    assert "no raster engine" in err and res["Status"] == "failed"
    assert repo.objects == {}
    ip_.close()


def test_result_json_is_encoding_json_shaped():
    from imageprocessor_b200 import Image
    ip_ = P.ImageProcessor(None, P.MemoryFileRepo())
    t = task([{"Type": "rotate", "Parameters": {}}], image_id='a<b>&"c ')
    ip_.process(t, Image.from_rgba(rgba_random(8, 8, 1)))
    raw = ip_.last_result_json
    assert raw.startswith('{"ID":"task-1","ImageID":"a\\u003cb\\u003e\\u0026\\"c\\u2028","Status":"failed","ProcessedPaths":{},"Error":')
    assert json.loads(raw)["ImageID"] == 'a<b>&"c '
    ip_.close()


def test_oversized_output_is_refused_before_any_allocation():
    """{"width": 60000, "height": 70000} from the broker must not reach the pinned allocator."""
    from imageprocessor_b200 import Image
    ip_ = P.ImageProcessor(None, P.MemoryFileRepo())
    img = Image.from_rgba(rgba_random(64, 48, 1))
    res, err = ip_.process(task([{"Type": "resize", "Parameters": {"width": 60000, "height": 70000}}]), img)
    assert "exceed the raster engine's limit of 65536" in err and res["Status"] == "failed"
    res, err = ip_.process(task([{"Type": "thumbnail", "Parameters": {"size": 1e9, "crop_to_fit": True}}]), img)
    assert "exceed the raster engine's limit" in err
    # keep_aspect brings 60000x70000 down to the source's aspect: 64x48 * min(937.5, 1458.3) = 60000x45000 -> accepted
    # by the bound (and then stopped only by the missing engine)
    res, err = ip_.process(task([{"Type": "resize", "Parameters": {"width": 60000, "height": 70000, "keep_aspect": True}}]), img)
    assert "no raster engine" in err
    ip_.close()


class CountingFace(P.PilFace):
    """Counts rasterise calls; maps runes to glyph indices through a table (default: the rune)."""

    def __init__(self, index_of=None):
        super().__init__()
        self.calls = []
        self.index_of = index_of or {}

    def index(self, rune):
        return self.index_of.get(rune, rune)

    def mask(self, rune, size, fx, fy):
        self.calls.append((rune, fx, fy))
        return super().mask(rune, size, fx, fy)


def _watermark_calls(text, face):
    from imageprocessor_b200 import Image
    ip_ = P.ImageProcessor(None, P.MemoryFileRepo(), face=face)
    ip_.process(task([{"Type": "watermark", "Parameters": {"text": text}}]), Image.from_rgba(rgba_random(900, 300, 1)))
    ip_.close()
    return face.calls


def test_glyph_mask_cache_is_freetypes_direct_mapped_table():
    """freetype.Context.glyph(): slot = (index % 256) * 4 + fx / 16; hit iff the slot holds the same index; a miss
    rasterises at this occurrence's sub-pixel offset and overwrites the slot."""
    # integer advances keep fx == 0 for every rune: one bucket per glyph index
    class IntFace(CountingFace):
        def advance_26_6(self, rune, size):
            return 20 * 64

        def mask(self, rune, size, fx, fy):
            adv, ox, oy, m = super().mask(rune, size, fx, fy)
            return 20 * 64, ox, oy, m
    # same rune again in the same bucket: one rasterisation
    assert [c[0] for c in _watermark_calls("aaaa", IntFace())] == [ord("a")]
    # two runes with ONE glyph index (e.g. both missing -> .notdef) share the first one's mask
    assert [c[0] for c in _watermark_calls("abab", IntFace({ord("a"): 7, ord("b"): 7}))] == [ord("a")]
    # indices equal mod 256 collide in the slot: each occurrence evicts the other and is rasterised again
    calls = _watermark_calls("abab", IntFace({ord("a"): 7, ord("b"): 7 + 256}))
    assert [c[0] for c in calls] == [ord("a"), ord("b"), ord("a"), ord("b")]
    # distinct slots: once each
    assert [c[0] for c in _watermark_calls("abab", IntFace({ord("a"): 7, ord("b"): 8}))] == [ord("a"), ord("b")]


# ---- end to end on the GPU --------------------------------------------------------------
from oracle.oracle import drawstring_layout as go_drawstring_layout  # noqa: E402


@pytest.mark.gpu
def test_process_canonical_task_matches_oracle(engines, oracle):
    import imageprocessor_b200 as ip
    w, h = 1600, 1200
    a = rgba_random(w, h, 77)
    repo = P.MemoryFileRepo()
    proc = P.ImageProcessor(engines(ip.PRECISION_EXACT), repo, encode=P.raw_encode)
    res, err = proc.process(task(CANONICAL_OPS), ip.Image.from_rgba(a), "jpeg")
    assert err is None
    assert res == {"ID": "task-1", "ImageID": "img-1", "Status": "completed", "Error": "", "ProcessedPaths": {
        "thumbnail": "processed/thumbnails/img-1/200.jpeg",
        "resize": "processed/resize/img-1/1024x768.jpeg",
        "watermark": "processed/watermarked/img-1/watermarked.jpeg"}}
    assert {k: v[1] for k, v in repo.objects.items()} == {p: "image/jpeg" for p in res["ProcessedPaths"].values()}
    R = oracle.Raster.rgba(a)
    nw, nh = oracle.keep_aspect_dims(w, h, 1024, 768)
    fmt, got = P.raw_decode(repo.objects["processed/resize/img-1/1024x768.jpeg"][0])
    assert fmt == "jpeg" and np.array_equal(got, oracle.resize_image(R, nw, nh))
    _, got = P.raw_decode(repo.objects["processed/thumbnails/img-1/200.jpeg"][0])
    assert np.array_equal(got, oracle.crop_and_resize(R, 200))
    # watermark: same glyph source on both sides (PilFace), blend bit-exact
    face = proc.face
    text = "© ImageProcessor"
    width_px = (sum(face.advance_26_6(ord(c), 36.0) for c in text) + 63) >> 6
    px, py = oracle.watermark_anchor("bottom-right", w, h, width_px, oracle.watermark_height_px(36.0))
    gl = go_drawstring_layout(face, text, 36.0, w, h, px, py)
    _, got = P.raw_decode(repo.objects["processed/watermarked/img-1/watermarked.jpeg"][0])
    want = oracle.watermark(R, (255, 255, 255, 127), [oracle.Glyph(*g) for g in gl])
    assert np.array_equal(got, want) and not np.array_equal(got, a)
    proc.close()


@pytest.mark.gpu
def test_process_formats_fit_thumbnail_and_failure_isolation(engines, oracle):
    import imageprocessor_b200 as ip
    a = rgba_random(900, 600, 5)
    repo = P.MemoryFileRepo()
    proc = P.ImageProcessor(engines(ip.PRECISION_EXACT), repo, encode=P.raw_encode)
    img = ip.Image.from_rgba(a)
    # png source, no target format: keeps png; non-crop thumbnail is fit-short-side (thumbnail.go:52-63)
    ops = [{"Type": "thumbnail", "Parameters": {"size": 100}}, {"Type": "resize", "Parameters": {"width": 300, "height": 300}}]
    res, err = proc.process(task(ops), img, "png")
    assert err is None and res["ProcessedPaths"] == {"thumbnail": "processed/thumbnails/img-1/100.png",
                                                      "resize": "processed/resize/img-1/300x300.png"}
    fmt, got = P.raw_decode(repo.objects["processed/thumbnails/img-1/100.png"][0])
    assert fmt == "png" and got.shape[:2] == (100, 150)
    assert np.array_equal(got, oracle.resize_image(oracle.Raster.rgba(a), 150, 100))
    _, got = P.raw_decode(repo.objects["processed/resize/img-1/300x300.png"][0])
    assert np.array_equal(got, oracle.resize_image(oracle.Raster.rgba(a), 300, 300))   # keep_aspect absent: stretched
    # webp target -> jpeg; watermark of a gif -> jpeg (resize.go:88-90, watermark.go:73-75)
    res, err = proc.process(task([{"Type": "resize", "Parameters": {"width": 90, "height": 60}}], fmt="webp"), img, "png")
    assert res["ProcessedPaths"]["resize"].endswith("90x60.jpeg")
    res, err = proc.process(task([{"Type": "watermark", "Parameters": {}}, {"Type": "thumbnail", "Parameters": {}}]), img, "gif")
    assert res["ProcessedPaths"] == {"watermark": "processed/watermarked/img-1/watermarked.jpeg",
                                     "thumbnail": "processed/thumbnails/img-1/200.gif"}
    # Process stops at the first failing operation; what came before stays saved (image_processor.go:64-75)
    repo.objects.clear()
    ops = [{"Type": "thumbnail", "Parameters": {"size": 32, "crop_to_fit": True}}, {"Type": "resize", "Parameters": {"width": "x"}},
           {"Type": "watermark", "Parameters": {}}]
    res, err = proc.process(task(ops), img, "jpeg")
    assert err.startswith("operation resize failed") and res["Status"] == "failed"
    assert res["ProcessedPaths"] == {"thumbnail": "processed/thumbnails/img-1/32.jpeg"} and list(repo.objects) == ["processed/thumbnails/img-1/32.jpeg"]
    # SaveProcessed failure
    repo.fail_on = "resize"
    res, err = proc.process(task([{"Type": "resize", "Parameters": {"width": 10, "height": 10}}]), img, "jpeg")
    assert err.startswith("failed to save processed image") and res["Error"].startswith("Failed to save processed image")
    repo.fail_on = None
    proc.close()


@pytest.mark.gpu
def test_process_batch_is_the_batching_worker_loop(engines, oracle):
    import imageprocessor_b200 as ip
    # its own engine with a long batching window: the coalescing asserted below must not depend on how fast
    # this (Python) submitter happens to run on a loaded box
    e = engines(ip.PRECISION_EXACT, batch_window_us=20000)
    repo = P.MemoryFileRepo()
    proc = P.ImageProcessor(e, repo, encode=P.raw_encode)
    imgs = [rgba_random(640 + 16 * k, 480 + 8 * k, 200 + k) for k in range(6)]
    tasks = [task(CANONICAL_OPS[:2], image_id=f"im{k}") for k in range(6)]
    tasks[3] = task([{"Type": "resize", "Parameters": {}}], image_id="im3")        # one bad message among good ones
    proc.process_batch(tasks, [ip.Image.from_rgba(a) for a in imgs], ["jpeg"] * 6)   # warms the pinned output buffers
    repo.objects.clear()
    b0 = e.stats()["batches"]
    out = proc.process_batch(tasks, [ip.Image.from_rgba(a) for a in imgs], ["jpeg"] * 6)
    assert e.stats()["batches"] - b0 <= 3         # tickets were coalesced, not one launch sequence per image
    for k, ((res, err), a) in enumerate(zip(out, imgs)):
        if k == 3:
            assert err and res["Status"] == "failed" and res["ProcessedPaths"] == {}
            continue
        assert err is None and res["Status"] == "completed"
        h, w = a.shape[:2]
        nw, nh = oracle.keep_aspect_dims(w, h, 1024, 768)
        _, got = P.raw_decode(repo.objects[f"processed/resize/im{k}/1024x768.jpeg"][0])
        assert np.array_equal(got, oracle.resize_image(oracle.Raster.rgba(a), nw, nh))
        _, got = P.raw_decode(repo.objects[f"processed/thumbnails/im{k}/200.jpeg"][0])
        assert np.array_equal(got, oracle.crop_and_resize(oracle.Raster.rgba(a), 200))
    proc.close()


@pytest.mark.gpu
def test_process_with_device_jpeg_saves_the_files_go_would_encode(engines, oracle, monkeypatch):
    """iph_set_device_jpeg: JPEG-bound results skip the encode callback and are saved as the device writer produced them --
    byte for byte jpeg.Encode(q85) of the oracle's raster result; PNG targets still go through the callback."""
    import imageprocessor_b200 as ip
    w, h = 1600, 1200
    a = rgba_random(w, h, 78)
    repo = P.MemoryFileRepo()
    seen = []

    def encode(rgba, fmt, quality):
        seen.append(fmt)
        return P.raw_encode(rgba, fmt, quality)

    proc = P.ImageProcessor(engines(ip.PRECISION_EXACT), repo, encode=encode, device_jpeg=True)
    res, err = proc.process(task(CANONICAL_OPS), ip.Image.from_rgba(a), "jpeg")
    assert err is None and res["Status"] == "completed" and seen == []
    R = oracle.Raster.rgba(a)
    nw, nh = oracle.keep_aspect_dims(w, h, 1024, 768)
    assert repo.objects["processed/resize/img-1/1024x768.jpeg"] == (oracle.jpeg_encode_rgba(oracle.resize_image(R, nw, nh), 85), "image/jpeg")
    assert repo.objects["processed/thumbnails/img-1/200.jpeg"][0] == oracle.jpeg_encode_rgba(oracle.crop_and_resize(R, 200), 85)
    face, text = proc.face, "© ImageProcessor"
    width_px = (sum(face.advance_26_6(ord(c), 36.0) for c in text) + 63) >> 6
    px, py = oracle.watermark_anchor("bottom-right", w, h, width_px, oracle.watermark_height_px(36.0))
    gl = go_drawstring_layout(face, text, 36.0, w, h, px, py)
    want = oracle.watermark(R, (255, 255, 255, 127), [oracle.Glyph(*g) for g in gl])
    assert repo.objects["processed/watermarked/img-1/watermarked.jpeg"][0] == oracle.jpeg_encode_rgba(want, 85)
    # a PNG target keeps the host encoder
    res, err = proc.process(task([{"Type": "resize", "Parameters": {"width": 300, "height": 200}}], fmt="png"), ip.Image.from_rgba(a), "jpeg")
    assert err is None and seen == ["png"]
    proc.close()
    # a file that does not fit the first buffer (forced: 1/16 byte per pixel) is retried with room for any scan
    monkeypatch.setenv("IPH_JPEG_FIRST_DIV", "16")
    proc = P.ImageProcessor(engines(ip.PRECISION_EXACT), repo, encode=encode, device_jpeg=True)
    res, err = proc.process(task([{"Type": "watermark", "Parameters": {}}]), ip.Image.from_rgba(a), "jpeg")
    assert err is None and repo.objects["processed/watermarked/img-1/watermarked.jpeg"][0] == oracle.jpeg_encode_rgba(want, 85)
    proc.close()
