#!/usr/bin/env python
"""PCIe-path probe: host-pinned sources/destinations through the C ABI with chosen ops."""
import argparse, ctypes as C, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import imageprocessor_b200 as ip
from imageprocessor_b200 import _lib as L, glyphs as G

ap = argparse.ArgumentParser()
ap.add_argument("--ops", default="rtw"); ap.add_argument("--images", type=int, default=128)
ap.add_argument("--slots", type=int, default=48); ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--lanes", type=int, default=4); ap.add_argument("--steps", type=int, default=3); ap.add_argument("--precision", type=int, default=0)
ap.add_argument("--layout", default="rgba", choices=["rgba", "ycbcr420"])
a = ap.parse_args()
W, H = 4000, 3000
eng = ip.Engine(devices=[0], precision=a.precision, lanes_per_device=a.lanes, max_batch=a.batch, batch_window_us=100)
lib, ctx = L.load(), eng._ctx
nw, nh = ip.keep_aspect_dims(W, H, 1024, 768); cx, cy, cs = ip.crop_square(W, H)
gl = G.layout_watermark(W, H, "© ImageProcessor"); col, _ = G.parse_color("255,255,255", 0.5)
garr = (L.Glyph * len(gl))(); keep = []
for j, g in enumerate(gl):
    m = np.ascontiguousarray(g.mask); keep.append(m)
    garr[j].x0, garr[j].y0, garr[j].x1, garr[j].y1, garr[j].mp_x, garr[j].mp_y = g.x0, g.y0, g.x1, g.y1, g.mp_x, g.mp_y
    garr[j].mask_w, garr[j].mask_h, garr[j].mask_stride, garr[j].mask = m.shape[1], m.shape[0], m.strides[0], m.ctypes.data
rng = np.random.default_rng(0)
base = rng.integers(0, 256, (H, W, 4), dtype=np.uint8); base[..., 3] = 255
descs, opss, pins = [], [], []
for s in range(a.slots):
    d = L.ImageDesc(); d.memspace, d.width, d.height = L.MEM_HOST, W, H
    if a.layout == "rgba":
        pi = eng.alloc_pinned(W * H * 4); pi.array[:] = base.reshape(-1); pins.append(pi)
        d.layout = L.RGBA8; d.plane[0] = pi.ptr; d.stride[0] = W * 4
    else:  # what image.Decode returns for a 4:2:0 JPEG: three planes, 1.5 bytes per pixel over PCIe
        d.layout = L.YCBCR420; d.opaque_hint = 1
        for k, (pw, ph) in enumerate(((W, H), (W // 2, H // 2), (W // 2, H // 2))):
            pi = eng.alloc_pinned(pw * ph); pi.array[:] = base.reshape(-1)[:pw * ph]; pins.append(pi)
            d.plane[k] = pi.ptr; d.stride[k] = pw
    descs.append(d)
    ops = (L.Op * 3)(); n = 0
    if "r" in a.ops:
        p = eng.alloc_pinned(nw * nh * 4); pins.append(p)
        ops[n].kind, ops[n].dst_w, ops[n].dst_h, ops[n].dst, ops[n].dst_stride = L.OP_RESIZE, nw, nh, p.ptr, nw * 4; n += 1
    if "t" in a.ops:
        p = eng.alloc_pinned(160000); pins.append(p)
        ops[n].kind, ops[n].dst_w, ops[n].dst_h, ops[n].dst, ops[n].dst_stride = L.OP_THUMB_CROP, 200, 200, p.ptr, 800
        ops[n].rect_x, ops[n].rect_y, ops[n].rect_w, ops[n].rect_h = cx, cy, cs, cs; n += 1
    if "w" in a.ops:
        p = eng.alloc_pinned(W * H * 4); pins.append(p)
        ops[n].kind, ops[n].dst_w, ops[n].dst_h, ops[n].dst, ops[n].dst_stride = L.OP_WATERMARK, W, H, p.ptr, W * 4
        for k in range(4): ops[n].color[k] = col[k]
        ops[n].n_glyphs, ops[n].glyphs = len(gl), garr; n += 1
    opss.append((ops, n))
tids = (C.c_uint64 * a.images)()
refs = [C.cast(C.byref(tids, 8 * i), C.POINTER(C.c_uint64)) for i in range(a.images)]
def step():
    for i in range(a.images):
        s = i % a.slots
        if i >= a.slots: L.check(lib.ipg_wait(ctx, tids[i - a.slots], -1))
        L.check(lib.ipg_submit_on(ctx, 0, C.byref(descs[s]), opss[s][0], opss[s][1], refs[i]))
    for i in range(max(a.images - a.slots, 0), a.images): L.check(lib.ipg_wait(ctx, tids[i], -1))
step(); eng.reset_stats(); t0 = time.perf_counter()
for _ in range(a.steps): step()
eng.flush(); dt = time.perf_counter() - t0; st = eng.stats()
print(json.dumps({"layout": a.layout, "ops": a.ops, "batch": a.batch, "lanes": a.lanes, "slots": a.slots, "img_per_s": a.images * a.steps / dt,
                  "h2d_GBps": st["bytes_h2d"] / dt / 1e9, "d2h_GBps": st["bytes_d2h"] / dt / 1e9, "batches": st["batches"]}))
eng.close()
