#!/usr/bin/env python
"""Diagnostic for the mixed-size stream (BASELINE configs[4]): where does the raster time go?
Runs the bench's seeded stream, already decoded, (a) through the raw engine API with pinned buffers, one submit per
image then wait-all, (b) through the streaming worker without encoders.  Prints wall time and the engine's per-class
kernel times; IPG_TRACE=1 adds the per-batch device timeline."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import imageprocessor_b200 as ip  # noqa: E402
from imageprocessor_b200 import codecs, glyphs as G, processor as P  # noqa: E402
from imageprocessor_b200.worker import StreamingWorker  # noqa: E402


def main():
    n = int(os.environ.get("N", 32))
    spec = bench.c5_stream_spec(42, n)
    decoded = [codecs.decode(codecs.synth_file(w, h, sd, k)) for (w, h, k, sd) in spec]
    eng = ip.Engine(devices=[0], lanes_per_device=4, max_batch=8, batch_window_us=200, lane_device_bytes=2 << 30,
                    lane_pinned_bytes=1 << 30)
    col = (255, 255, 255, 127)

    def ops_for(img):
        w, h = img.width, img.height
        nw, nh = ip.keep_aspect_dims(w, h, 1024, 768)
        cx, cy, cs = ip.crop_square(w, h)
        return [ip.OpSpec.thumb_crop((cx, cy, cs, cs), 200), ip.OpSpec.resize(nw, nh),
                ip.OpSpec.watermark(w, h, col, G.layout_watermark(w, h, "© ImageProcessor"))]

    all_ops = [ops_for(img) for img, _ in decoded]
    for rep in range(2):
        eng.reset_stats()
        t0 = time.perf_counter()
        tk = [eng.submit(img, ops) for (img, _), ops in zip(decoded, all_ops)]
        t1 = time.perf_counter()
        for t in tk:
            eng.wait(t)
        t2 = time.perf_counter()
        st = eng.stats()
        print(f"raw engine, pageable buffers: submit {t1 - t0:.3f} s, total {t2 - t0:.3f} s; stream {st['stream_kernel_ms']:.1f} ms "
              f"fix {st['fix_kernel_ms']:.1f} ms other {st['other_kernel_ms']:.1f} ms, batches {st['batches']}, fallbacks {st['exact_fallbacks']}, "
              f"fixups {st['exact_fixups']}, batch span {st['batch_span_ms']:.1f} ms")
    # per image, alone
    for i, ((img, _), ops) in enumerate(zip(decoded, all_ops)):
        eng.reset_stats()
        t0 = time.perf_counter()
        eng.run(img, ops)
        dt = time.perf_counter() - t0
        st = eng.stats()
        print(f"  {i:2d} {spec[i][2]:9s} {img.width}x{img.height}: {1e3 * dt:7.1f} ms wall; stream {st['stream_kernel_ms']:.2f} fix {st['fix_kernel_ms']:.2f} "
              f"other {st['other_kernel_ms']:.2f} ms, fixups {st['exact_fixups']}, fallbacks {st['exact_fallbacks']}")
    ops_json = [{"Type": "thumbnail", "Parameters": {"size": 200, "crop_to_fit": True}},
                {"Type": "resize", "Parameters": {"width": 1024, "height": 768, "keep_aspect": True}},
                {"Type": "watermark", "Parameters": {}}]
    proc = P.ImageProcessor(eng, P.MemoryFileRepo(), encode=lambda a, f, q: b"")
    wk = StreamingWorker(proc, 16, decode=lambda item: item)
    msgs = [({"ID": str(i), "ImageID": str(i), "Operations": ops_json, "Format": ""}, decoded[i]) for i in range(n)]
    for rep in range(3):
        eng.reset_stats()
        s = wk.run(msgs)
        st = eng.stats()
        print(f"worker x16, no codecs: wall {s.wall_s:.3f} s ({n / s.wall_s:.1f} img/s); stream {st['stream_kernel_ms']:.1f} fix {st['fix_kernel_ms']:.1f} "
              f"other {st['other_kernel_ms']:.1f} ms, batches {st['batches']}, batch span {st['batch_span_ms']:.1f} ms, failed {s.failed}")
    proc.close()
    eng.close()


if __name__ == "__main__":
    main()
