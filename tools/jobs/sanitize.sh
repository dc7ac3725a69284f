#!/bin/bash
# GPU-box job: compute-sanitizer memcheck over a small mixed workload (lean, general/alpha, fused, odd sizes, YCbCr).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python - <<'PY' 2>&1 | tail -25
import numpy as np, imageprocessor_b200 as ip
from tests.util import rgba_random, synthetic_glyphs
for fuse in (1, 2, 3):
    with ip.Engine(devices=[0], fuse_targets=fuse, lanes_per_device=2) as e:
        for (w, h, alpha) in [(1600, 1200, "opaque"), (1203, 899, "premul"), (641, 479, "raw"), (2048, 17, "opaque"), (33, 1500, "opaque")]:
            a = rgba_random(w, h, 5, alpha)
            nw, nh = ip.keep_aspect_dims(w, h, 1024, 768)
            cx, cy, cs = ip.crop_square(w, h)
            gl = synthetic_glyphs(w, h, 3, n=4)
            ops = [ip.OpSpec.resize(nw, nh), ip.OpSpec.thumb_crop((cx, cy, cs, cs), 200),
                   ip.OpSpec.watermark(w, h, (255, 255, 255, 127), [ip.GlyphMask(*g) for g in gl])]
            out = e.run(ip.Image.from_rgba(a), ops)
        y = np.random.default_rng(1).integers(0, 256, (301, 403), dtype=np.uint8)
        c = np.random.default_rng(2).integers(0, 256, (151, 202), dtype=np.uint8)
        e.run(ip.Image.from_ycbcr(y, c, c.copy(), ip.YCBCR420), [ip.OpSpec.resize(100, 75), ip.OpSpec.watermark(403, 301, (0, 0, 0, 127), [])])
    print("fuse", fuse, "ok")
PY
echo "sanitizer rc=${PIPESTATUS[0]}"
