"""The streaming worker (imageprocessor_b200/worker.py): the reference's processWorker / processMessage loop
(internal/worker/worker.go:112-149,165-234) over the GPU processor, and the host codec stand-ins that feed it the
concrete raster types image.Decode would (imageprocessor_b200/codecs.py).

not gpu: decode types, the loop's bookkeeping and error isolation without an engine.
gpu    : a mixed PNG+JPEG stream (BASELINE configs[4] in miniature) end to end, stored objects byte for byte vs the oracle.
"""
import io

import numpy as np
import pytest

import imageprocessor_b200 as ip
from imageprocessor_b200 import codecs, processor as P
from imageprocessor_b200.worker import StreamingWorker

OPS = [{"Type": "thumbnail", "Parameters": {"size": 200, "crop_to_fit": True}},
       {"Type": "resize", "Parameters": {"width": 1024, "height": 768, "keep_aspect": True}},
       {"Type": "watermark", "Parameters": {"text": "© ImageProcessor", "opacity": 0.5, "position": "bottom-right"}}]


def task(i, ops=OPS, fmt=""):
    return {"ID": f"task-{i}", "ImageID": f"img-{i}", "OriginalPath": f"original/{i}", "Bucket": "images", "Operations": ops,
            "Format": fmt}


def test_decode_hands_over_the_types_image_decode_would():
    from PIL import Image as PI
    img, fmt = codecs.decode(codecs.synth_file(641, 481, 1, "jpeg"))
    assert fmt == "jpeg" and img.layout == ip.YCBCR420 and img.opaque_hint
    assert img.planes[0].shape == (481, 641) and img.planes[1].shape == (241, 321) == img.planes[2].shape  # image.NewYCbCr sizes
    img, fmt = codecs.decode(codecs.synth_file(100, 50, 2, "png"))
    assert fmt == "png" and img.layout == ip.RGBA8 and img.opaque_hint and (img.planes[0][..., 3] == 255).all()
    img, fmt = codecs.decode(codecs.synth_file(100, 50, 3, "png-alpha"))
    assert fmt == "png" and img.layout == ip.NRGBA8 and not img.opaque_hint
    buf = io.BytesIO()
    PI.fromarray(codecs.synth_picture(64, 40, 4)[..., 0], "L").save(buf, "JPEG")
    img, fmt = codecs.decode(buf.getvalue())
    assert fmt == "jpeg" and img.layout == ip.GRAY8
    buf = io.BytesIO()
    PI.fromarray(codecs.synth_picture(64, 40, 4), "RGB").save(buf, "JPEG", subsampling="4:4:4")
    assert codecs.decode(buf.getvalue())[0].layout == ip.YCBCR444
    buf = io.BytesIO()
    PI.fromarray(codecs.synth_picture(64, 40, 4), "RGB").save(buf, "JPEG", subsampling="4:2:2")
    img = codecs.decode(buf.getvalue())[0]
    assert img.layout == ip.YCBCR422 and img.planes[1].shape == (40, 32)
    buf = io.BytesIO()
    PI.fromarray(codecs.synth_picture(64, 40, 4), "RGB").quantize(32).save(buf, "GIF")
    img, fmt = codecs.decode(buf.getvalue())
    assert fmt == "gif" and img.layout == ip.RGBA8 and img.planes[0].shape == (40, 64, 4)


def test_worker_loop_without_an_engine_isolates_failures():
    """Every message gets its own (result, error); a corrupt file fails like image.Decode does, the others still run up to
    the raster call, which must fail loudly (no CPU fallback)."""
    proc = P.ImageProcessor(None, P.MemoryFileRepo())
    wk = StreamingWorker(proc, concurrency=3)
    files = [codecs.synth_file(320, 240, i, ("jpeg", "png", "png-alpha")[i % 3]) for i in range(7)]
    files[3] = b"not an image"
    st = wk.run([(task(i), files[i]) for i in range(7)])
    assert st.messages == 7 and st.failed == 7 and len(st.results) == 7
    assert [r.task_id for r in st.results] == [f"task-{i}" for i in range(7)]          # results keep message order
    assert st.results[3].error.startswith("failed to decode image:") and st.results[3].pixels == 0
    for i in (0, 1, 2, 4, 5, 6):
        assert "no raster engine" in st.results[i].error and st.results[i].pixels == 320 * 240
    # a task without operations completes and is counted as such
    st = wk.run([(task(0, ops=[]), files[0])])
    assert st.failed == 0 and st.results[0].result["Status"] == "completed"
    s = st.summary()
    assert s["messages"] == 1 and set(s["thread_seconds"]) == {"decode", "raster_submit_to_wait", "encode", "save"}
    proc.close()


def test_stream_spec_is_seeded_and_covers_the_range():
    import bench
    a, b = bench.c5_stream_spec(42, 32), bench.c5_stream_spec(42, 32)
    assert a == b and a != bench.c5_stream_spec(43, 32)
    mp = [w * h / 1e6 for (w, h, _, _) in a]
    assert max(mp) == 48.0 and min(mp) < 0.31 and all(0.07 <= m <= 64.0 for m in mp)
    kinds = [k for (_, _, k, _) in a]
    assert kinds.count("jpeg") == 16 and kinds.count("png") == 8 and kinds.count("png-alpha") == 8


@pytest.mark.gpu
def test_mixed_stream_end_to_end_matches_oracle(oracle):
    """BASELINE configs[4] in miniature: PNG + JPEG files of mixed sizes through W worker threads; every stored object is
    compared with the oracle run on the SAME decoded planes (parity is defined on identical decoded inputs)."""
    O = oracle
    spec = [(1600, 1200, "jpeg"), (801, 601, "png"), (2048, 1365, "png-alpha"), (640, 480, "jpeg"), (3000, 2000, "png"),
            (1000, 1500, "jpeg"), (1920, 1080, "png-alpha"), (4000, 3000, "jpeg"), (500, 500, "png"), (2400, 1350, "jpeg")]
    files = [codecs.synth_file(w, h, 100 + i, k) for i, (w, h, k) in enumerate(spec)]
    repo = P.MemoryFileRepo()
    with ip.Engine(devices=[0], lanes_per_device=3, max_batch=8, batch_window_us=200) as eng:
        proc = P.ImageProcessor(eng, repo, encode=P.raw_encode)
        wk = StreamingWorker(proc, concurrency=4)
        st = wk.run([(task(i), files[i]) for i in range(len(files))])
        face = proc.face
        assert st.failed == 0, [r.error for r in st.results]
        engine_stats = eng.stats()
        proc.close()
    assert engine_stats["tickets_done"] == len(files) and engine_stats["kernels_launched"] > 0
    assert len(repo.objects) == 3 * len(files)
    for i, (w, h, kind) in enumerate(spec):
        img, fmt = codecs.decode(files[i])
        if img.layout in (ip.RGBA8, ip.NRGBA8):
            R = O.Raster.rgba(img.planes[0].reshape(h, w, 4), img.layout)
        else:
            R = O.Raster.ycbcr(*img.planes, img.layout)
        res = st.results[i].result
        assert res["Status"] == "completed" and set(res["ProcessedPaths"]) == {"thumbnail", "resize", "watermark"}
        ext = "jpeg" if kind == "jpeg" else "png"
        assert res["ProcessedPaths"]["resize"] == f"processed/resize/img-{i}/1024x768.{ext}"
        nw, nh = ip.keep_aspect_dims(w, h, 1024, 768)
        assert np.array_equal(P.raw_decode(repo.objects[res["ProcessedPaths"]["resize"]][0])[1], O.resize_image(R, nw, nh))
        assert np.array_equal(P.raw_decode(repo.objects[res["ProcessedPaths"]["thumbnail"]][0])[1], O.crop_and_resize(R, 200))
        wpx = (sum(face.advance_26_6(ord(ch), 36.0) for ch in "© ImageProcessor") + 63) >> 6
        px, py = P.watermark_anchor("bottom-right", w, h, wpx, P.watermark_height_px(36.0))
        gl = O.drawstring_layout(face, "© ImageProcessor", 36.0, w, h, px, py)
        want = O.watermark(R, (255, 255, 255, 127), [O.Glyph(*g) for g in gl])
        assert np.array_equal(P.raw_decode(repo.objects[res["ProcessedPaths"]["watermark"]][0])[1], want)
