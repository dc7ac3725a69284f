#!/usr/bin/env python
"""Small fixed workload for ncu / timing experiments: N 12 MP RGBA images resident in
HBM, resize+thumb(+watermark) through the C ABI, one lane so launches do not overlap.
Prints per-kernel-class device time from the engine's CUDA events."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import imageprocessor_b200 as ip
from imageprocessor_b200 import glyphs as G


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=32)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--ops", default="rtw")
    ap.add_argument("--precision", type=int, default=0)
    ap.add_argument("--lanes", type=int, default=1)
    ap.add_argument("--max-batch", type=int, default=32)
    ap.add_argument("--w", type=int, default=4000)
    ap.add_argument("--h", type=int, default=3000)
    ap.add_argument("--opaque-hint", type=int, default=0)
    ap.add_argument("--fuse", type=int, default=0)
    ap.add_argument("--layout", default="rgba", choices=["rgba", "nrgba", "ycbcr420", "ycbcr444"])
    a = ap.parse_args()
    import torch
    dev = torch.device("cuda", 0)
    W, H = a.w, a.h
    srcs = []
    g = torch.Generator(device=dev)
    planar = a.layout not in ("rgba", "nrgba")
    cw, ch = ((W + 1) // 2, (H + 1) // 2) if a.layout == "ycbcr420" else (W, H)
    for i in range(a.images):
        g.manual_seed(1000 + i)
        if planar:
            srcs.append((torch.randint(0, 256, (H, W), dtype=torch.uint8, device=dev, generator=g),
                         torch.randint(0, 256, (ch, cw), dtype=torch.uint8, device=dev, generator=g),
                         torch.randint(0, 256, (ch, cw), dtype=torch.uint8, device=dev, generator=g)))
            continue
        t = torch.randint(0, 256, (H, W, 4), dtype=torch.uint8, device=dev, generator=g)
        if a.layout == "rgba":
            t[..., 3] = 255
        srcs.append(t)
    nw, nh = ip.keep_aspect_dims(W, H, 1024, 768)
    cx, cy, cs = ip.crop_square(W, H)
    o_r = torch.empty((a.images, nh, nw, 4), dtype=torch.uint8, device=dev)
    o_t = torch.empty((a.images, 200, 200, 4), dtype=torch.uint8, device=dev)
    o_w = torch.empty((a.images, H, W, 4), dtype=torch.uint8, device=dev) if "w" in a.ops else None
    torch.cuda.synchronize()
    gl = G.layout_watermark(W, H, "© ImageProcessor")
    col, _ = G.parse_color("255,255,255", 0.5)
    eng = ip.Engine(devices=[0], precision=a.precision, lanes_per_device=a.lanes, max_batch=a.max_batch,
                    batch_window_us=2000, fuse_targets=a.fuse)
    for step in range(a.steps + 1):
        if step == 1:
            eng.reset_stats()
        tk = []
        for i in range(a.images):
            ops = []
            if "r" in a.ops:
                ops.append(ip.OpSpec.resize(nw, nh, dst_device=(o_r[i].data_ptr(), nw * 4)))
            if "t" in a.ops:
                ops.append(ip.OpSpec.thumb_crop((cx, cy, cs, cs), 200, dst_device=(o_t[i].data_ptr(), 800)))
            if "w" in a.ops:
                ops.append(ip.OpSpec.watermark(W, H, col, gl, dst_device=(o_w[i].data_ptr(), W * 4)))
            if planar:
                lay = ip.YCBCR420 if a.layout == "ycbcr420" else ip.YCBCR444
                img = ip.Image.on_device(lay, W, H, [p.data_ptr() for p in srcs[i]], [W, cw, cw], opaque_hint=True)
            else:
                img = ip.Image.on_device(ip.NRGBA8 if a.layout == "nrgba" else ip.RGBA8, W, H, [srcs[i].data_ptr()], [W * 4],
                                         opaque_hint=bool(a.opaque_hint))
            tk.append(eng.submit(img, ops, device=0))
        for t in tk:
            eng.wait(t)
    st = eng.stats()
    n = a.images * a.steps
    src_bytes = (W * H + 2 * cw * ch) if planar else W * H * 4
    bytes_img = src_bytes + (W * H * 4 if "w" in a.ops else 0) + (nw * nh * 4 if "r" in a.ops else 0) + (160000 if "t" in a.ops else 0)
    out = {"images": n, "ops": a.ops, "batches": st["batches"], "launches": st["kernels_launched"],
           "stream_us_per_image": 1e3 * st["stream_kernel_ms"] / n, "fix_us_per_image": 1e3 * st["fix_kernel_ms"] / n,
           "other_us_per_image": 1e3 * st["other_kernel_ms"] / n, "fixups_per_image": st["exact_fixups"] / n,
           "stream_GBps": bytes_img * n / (st["stream_kernel_ms"] * 1e-3) / 1e9 if st["stream_kernel_ms"] else None,
           "span_us_per_image": 1e3 * st["kernel_span_ms"] / n}
    print(json.dumps(out))
    eng.close()


if __name__ == "__main__":
    main()
