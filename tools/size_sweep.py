#!/usr/bin/env python
"""Device time per image against source size: opaque RGBA8 sources resident in HBM, resize 1024x768
keep-aspect + thumbnail 200 crop (+ watermark), one launch sequence per step, CUDA-event kernel times from the engine.
Writes profiles/<tag>_size_sweep.json when --out is given."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import imageprocessor_b200 as ip  # noqa: E402
from imageprocessor_b200 import glyphs as G  # noqa: E402

# a 4:3 ladder (same output shape 1024x768 for every rung at or above it, so the rungs compare), then the 16:9 / 3:2 sizes
# DESIGN.md quoted in round 1
SIZES = [(640, 480), (1152, 864), (1632, 1224), (2048, 1536), (2832, 2124), (4000, 3000), (6656, 4992), (8000, 6000),
         (1920, 1080), (2560, 1440), (3000, 2000), (7680, 4320)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--ops", default="rt,rtw")
    a = ap.parse_args()
    import torch
    dev = torch.device("cuda", 0)
    eng = ip.Engine(devices=[0], lanes_per_device=1, max_batch=64, batch_window_us=5000, lane_device_bytes=2 << 30)
    col, _ = G.parse_color("255,255,255", 0.5)
    rows = []
    for (W, H) in SIZES:
        n = int(min(64, max(8, 1.2e9 // (W * H * 4))))
        g = torch.Generator(device=dev)
        srcs = []
        for i in range(n):
            g.manual_seed(i)
            t = torch.randint(0, 256, (H, W, 4), dtype=torch.uint8, device=dev, generator=g)
            t[..., 3] = 255
            srcs.append(t)
        nw, nh = ip.keep_aspect_dims(W, H, 1024, 768)
        cx, cy, cs = ip.crop_square(W, H)
        o_r = torch.empty((n, nh, nw, 4), dtype=torch.uint8, device=dev)
        o_t = torch.empty((n, 200, 200, 4), dtype=torch.uint8, device=dev)
        o_w = torch.empty((n, H, W, 4), dtype=torch.uint8, device=dev)
        gl = G.layout_watermark(W, H, "© ImageProcessor")
        rec = {"size": f"{W}x{H}", "megapixels": W * H / 1e6, "images_per_launch": n, "resize_to": f"{nw}x{nh}"}
        for ops in a.ops.split(","):
            for step in range(4):
                if step == 1:
                    eng.reset_stats()
                tk = []
                for i in range(n):
                    o = [ip.OpSpec.resize(nw, nh, dst_device=(o_r[i].data_ptr(), nw * 4)),
                         ip.OpSpec.thumb_crop((cx, cy, cs, cs), 200, dst_device=(o_t[i].data_ptr(), 800))]
                    if "w" in ops:
                        o.append(ip.OpSpec.watermark(W, H, col, gl, dst_device=(o_w[i].data_ptr(), W * 4)))
                    tk.append(eng.submit(ip.Image.on_device(ip.RGBA8, W, H, [srcs[i].data_ptr()], [W * 4]), o, device=0))
                for t in tk:
                    eng.wait(t)
            st = eng.stats()
            m = 3 * n
            byt = W * H * 4 + nw * nh * 4 + 160000 + (W * H * 4 if "w" in ops else 0)
            us = 1e3 * st["kernel_ms"] / m
            rec[ops] = {"kernel_us_per_image": us, "stream_us": 1e3 * st["stream_kernel_ms"] / m, "fix_us": 1e3 * st["fix_kernel_ms"] / m,
                        "other_us": 1e3 * st["other_kernel_ms"] / m, "algorithmic_MB": byt / 1e6, "GBps": byt / us / 1e3,
                        "whole_image_fp64_fallbacks_per_image": st["exact_fallbacks"] / m, "fixups_per_image": st["exact_fixups"] / m}
        rows.append(rec)
        print(json.dumps(rec), flush=True)
        del srcs, o_r, o_t, o_w
        torch.cuda.empty_cache()
    eng.close()
    if a.out:
        json.dump({"tool": "tools/size_sweep.py", "what": "device time per image (CUDA events around the kernels), opaque RGBA8, device-resident",
                   "rows": rows}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
