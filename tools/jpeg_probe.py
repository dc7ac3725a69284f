#!/usr/bin/env python
"""Timing probe of the device-side JPEG writer: N 12 MP RGBA sources resident in HBM, resize + thumbnail + watermark with
every result returned as a JPEG file (pinned host buffers), against the same ops returning RGBA8.  Prints the engine's
per-section device times (CUDA events) and the bytes that crossed PCIe.  --noise: random bytes (worst case for the
entropy coder) instead of a photo-like image."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import imageprocessor_b200 as ip
from imageprocessor_b200 import glyphs as G


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=32)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--w", type=int, default=4000)
    ap.add_argument("--h", type=int, default=3000)
    ap.add_argument("--noise", type=int, default=0)
    ap.add_argument("--verify", type=int, default=1)
    a = ap.parse_args()
    import torch
    dev = torch.device("cuda", 0)
    W, H = a.w, a.h
    g = torch.Generator(device=dev)
    srcs = []
    yy, xx = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
    for i in range(a.images):
        g.manual_seed(1000 + i)
        if a.noise:
            t = torch.randint(0, 256, (H, W, 4), dtype=torch.uint8, device=dev, generator=g)
        else:  # smooth gradients + mild noise: compresses like a photograph
            t = torch.empty((H, W, 4), dtype=torch.uint8, device=dev)
            for c in range(3):
                v = 128 + 90 * torch.sin(xx / (37.0 + 5 * c) + i) * torch.cos(yy / (23.0 + 3 * c)) + 8 * torch.randn((H, W), device=dev, generator=g)
                t[..., c] = v.clamp(0, 255).to(torch.uint8)
        t[..., 3] = 255
        srcs.append(t)
    del yy, xx
    torch.cuda.synchronize()
    nw, nh = ip.keep_aspect_dims(W, H, 1024, 768)
    cx, cy, cs = ip.crop_square(W, H)
    gl = G.layout_watermark(W, H, "© ImageProcessor")
    col, _ = G.parse_color("255,255,255", 0.5)
    eng = ip.Engine(devices=[0], lanes_per_device=2, max_batch=a.images, batch_window_us=2000, lane_device_bytes=12 << 30)
    cap = W * H * (3 if a.noise else 1)
    bufs = [[eng.alloc_pinned(nw * nh * 3 + 16), eng.alloc_pinned(200 * 200 * 3 + 4096 + 16), eng.alloc_pinned(cap + 16)] for _ in range(a.images)]
    rgba = [[eng.alloc_pinned(nw * nh * 4), eng.alloc_pinned(200 * 200 * 4), eng.alloc_pinned(W * H * 4)] for _ in range(a.images)]
    out = {}
    for mode in ("jpeg", "rgba"):
        for step in range(a.steps + 1):
            if step == 1:
                eng.reset_stats()
                t0 = time.perf_counter()
            ts = []
            for i, t in enumerate(srcs):
                im = ip.Image.on_device(ip.RGBA8, W, H, [t.data_ptr()], [W * 4], opaque_hint=False)
                if mode == "jpeg":
                    kw = [dict(jpeg_quality=85, jpeg_buffer=b.array) for b in bufs[i]]
                else:
                    kw = [dict(dst=rgba[i][0].array.reshape(nh, nw, 4)), dict(dst=rgba[i][1].array.reshape(200, 200, 4)),
                          dict(dst=rgba[i][2].array.reshape(H, W, 4))]
                ts.append(eng.submit(im, [ip.OpSpec.resize(nw, nh, **kw[0]), ip.OpSpec.thumb_crop((cx, cy, cs, cs), 200, **kw[1]),
                                          ip.OpSpec.watermark(W, H, col, gl, **kw[2])], device=0))
            res = [eng.wait(t) for t in ts]
        wall = time.perf_counter() - t0
        st = eng.stats()
        n = a.images * a.steps
        out[mode] = {"images_per_s_wall": n / wall, "us_per_image": {k: 1e3 * st[k] / n for k in ("stream_kernel_ms", "fix_kernel_ms", "other_kernel_ms", "d2h_ms")},
                     "d2h_MB_per_image": st["bytes_d2h"] / n / 1e6, "kernels": st["kernels_launched"]}
        if mode == "jpeg":
            out[mode]["file_bytes"] = [int(r.nbytes) for r in res[0]]
            if a.verify:
                from oracle import oracle as O
                src = srcs[0].cpu().numpy()
                R = O.Raster.rgba(src)
                og = [O.Glyph(q.x0, q.y0, q.x1, q.y1, q.mask, q.mp_x, q.mp_y) for q in gl]
                want = [O.jpeg_encode_rgba(O.resize_image(R, nw, nh), 85), O.jpeg_encode_rgba(O.crop_and_resize(R, 200), 85),
                        O.jpeg_encode_rgba(O.watermark(R, col, og), 85)]
                out[mode]["verified_image0_all_files_byte_identical"] = [res[0][k].data == want[k] for k in range(3)]
    print(json.dumps(out))
    eng.close()


if __name__ == "__main__":
    main()
