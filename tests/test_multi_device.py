"""One context driving every visible GPU (ipg_init(NULL, 0, ...)): the layout north_star names -- one worker
thread + stream set per device inside ONE process, tickets routed by ipg_submit -- which is what the cgo shim in
INTEGRATION.md uses.  Needs >= 2 devices (run with `gpurun --gpus 2`); skipped below that.

Why this exists: cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and the occupancy-derived grids are per device;
round 1 cached them process-wide, so every k_stream launch on a second device would have failed.
"""
import numpy as np
import pytest

from tests.util import rgba_random, synthetic_glyphs


def _ops_and_expected(ip, O, a, layout="rgba"):
    h, w = a.shape[:2]
    nw, nh = ip.keep_aspect_dims(w, h, 1024, 768)
    cx, cy, cs = ip.crop_square(w, h)
    gl = synthetic_glyphs(w, h, 5)
    col = (255, 255, 255, 127)
    ops = [ip.OpSpec.resize(nw, nh), ip.OpSpec.thumb_crop((cx, cy, cs, cs), 200),
           ip.OpSpec.watermark(w, h, col, [ip.GlyphMask(*g) for g in gl]),
           ip.OpSpec.thumb_crop((cx, cy, cs, cs), 200, jpeg_quality=85)]   # ... and one result as a device-encoded JPEG file
    R = O.Raster.rgba(a, O.NRGBA8 if layout == "nrgba" else O.RGBA8)
    ogl = [O.Glyph(*g) for g in gl]
    thumb = O.crop_and_resize(R, 200)
    return ops, [O.resize_image(R, nw, nh), thumb, O.watermark(R, col, ogl), O.jpeg_encode_rgba(thumb, 85)]


@pytest.mark.gpu(min_devices=2)
def test_one_context_all_devices_auto_routing_and_submit_on(oracle):
    import imageprocessor_b200 as ip
    from imageprocessor_b200 import _lib as L
    O = oracle
    with ip.Engine(devices=None, max_batch=4, batch_window_us=0) as e:
        n_dev = e.device_count
        assert n_dev >= 2
        cases = []
        sizes = [(1600, 1200), (2048, 1536), (1999, 1201), (1280, 960), (3000, 2000), (1024, 1024)]
        for k, (w, h) in enumerate(sizes * 2):
            alpha = "opaque" if k % 3 else "premul"          # lean kernel, redo flag and general (alpha) paths
            layout = "nrgba" if k % 5 == 4 else "rgba"       # ... and k_stream_planar<NRGBA> + k_watermark
            a = rgba_random(w, h, 300 + k, alpha="raw" if layout == "nrgba" else alpha)
            ops, exp = _ops_and_expected(ip, O, a, layout)
            img = ip.Image.from_rgba(a, L.NRGBA8 if layout == "nrgba" else L.RGBA8)
            cases.append((img, ops, exp))
        # (1) auto routing: submit everything, then wait (least-outstanding-bytes spreads it over the devices)
        tickets = [e.submit(img, ops) for img, ops, _ in cases]
        for t, (_, _, exp) in zip(tickets, cases):
            for got, want in zip(e.wait(t), exp):
                assert got.data == want if isinstance(want, bytes) else np.array_equal(got, want)
        # (2) every device explicitly, every kernel family on each
        for dev in range(n_dev):
            tickets = [e.submit(img, ops, device=dev) for img, ops, _ in cases[:6]]
            for t, (_, _, exp) in zip(tickets, cases[:6]):
                for got, want in zip(e.wait(t), exp):
                    assert (got.data == want if isinstance(want, bytes) else np.array_equal(got, want)), f"device {dev}"
        st = e.stats()
        assert st["tickets_done"] == len(cases) + 6 * n_dev and st["kernels_launched"] > 0


@pytest.mark.gpu(min_devices=2)
def test_device_resident_buffers_on_second_device(oracle):
    """IPG_MEM_DEVICE source and destinations on device 1 of a two-device context."""
    import imageprocessor_b200 as ip
    from imageprocessor_b200 import _lib as L
    O = oracle
    w, h = 2000, 1500
    a = rgba_random(w, h, 9)
    nw, nh = ip.keep_aspect_dims(w, h, 1024, 768)
    with ip.Engine(devices=None) as e:
        dev = e.device_count - 1
        src = e.alloc_device(dev, a.nbytes)
        dst = e.alloc_device(dev, nw * nh * 4)
        e.to_device(dev, src, a)
        img = ip.Image.on_device(L.RGBA8, w, h, [src], [w * 4])
        e.run(img, [ip.OpSpec.resize(nw, nh, dst_device=(dst, nw * 4))], device=dev)
        out = np.empty((nh, nw, 4), np.uint8)
        e.from_device(dev, out, dst)
        e.free_device(dev, src)
        e.free_device(dev, dst)
    assert np.array_equal(out, O.resize_image(O.Raster.rgba(a), nw, nh))
