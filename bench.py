#!/usr/bin/env python
"""bench.py -- images/sec for resize(1024x768 keep-aspect) + thumbnail(200 crop) +
watermark on a batch of 12 MP RGBA images (BASELINE.json metric / configs[1..2]).

  python bench.py [--gpus N] [--steps K] [--warmup W]          # this repo's CUDA path
  python bench.py --impl reference [...]                        # restated reference CPU path
  torchrun ... bench.py --gpus N ...                            # one rank per GPU, images shard by rank

One step = one pass of the hot path over the whole batch (256 images per GPU).
`value`  : device-resident inputs/outputs, device clock from the first kernel start to
           the last kernel end of the K timed steps (CUDA events on the launching streams).
`e2e`    : the same K steps through the C ABI with pinned HOST buffers: H2D of every
           source and D2H of every result inside the timed region.
`roofline`: k_stream's algorithmic bytes per launch / its mean launch duration, against
           the measured HBM copy peak (MEASURED_PEAKS.json).
`cpu_baseline`: oracle/ (C restatement of the reference algorithm) on the host cores.
Nothing here reads /root/reference.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W_IMG, H_IMG = 4000, 3000
RW, RH, THUMB = 1024, 768, 200
WM_TEXT, WM_OPACITY, WM_COLOR = "© ImageProcessor", 0.5, "255,255,255"
METRIC = "images/sec resize+thumb+watermark, 12MP batch"
# algorithmic bytes per 12 MP RGBA image, source read once
BYTES_PER_IMAGE = W_IMG * H_IMG * 4 + RW * RH * 4 + THUMB * THUMB * 4 + W_IMG * H_IMG * 4
# ... and of the two passes the engine runs per image: the lean k_stream instantiation fuses
# resize + watermark copy (source read once), the general one does the thumbnail over the crop square
BYTES_LEAN_PASS = W_IMG * H_IMG * 4 + RW * RH * 4 + W_IMG * H_IMG * 4
BYTES_THUMB_PASS = H_IMG * H_IMG * 4 + THUMB * THUMB * 4


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """dram bytes per k_stream launch from the committed ncu --set full capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 8:
                    continue
                try:
                    sm.append(float(p[0])); mx.append(float(p[1]))
                except ValueError:
                    continue
                for n, v in zip(names, p[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


# ---------------------------------------------------------------------------------
def make_host_image(seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, (H_IMG, W_IMG, 4), dtype=np.uint8)
    a[..., 3] = 255
    return a


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path: Go cannot be built here, so
    this is oracle/ (C restatement, float64 scalar, fresh temp buffers, one image per
    thread) on all host cores, on a bounded sample per step."""
    if rank != 0:
        return
    from oracle import oracle as O
    from imageprocessor_b200 import glyphs as G
    O.build()
    cores = os.cpu_count() or 1
    n = env_int("IPG_BENCH_REF_IMAGES", 4 * cores)
    imgs = [make_host_image(1000 + i) for i in range(min(n, 4))]
    rasters = [O.Raster.rgba(imgs[i % len(imgs)]) for i in range(n)]
    gl = [O.Glyph(g.x0, g.y0, g.x1, g.y1, g.mask, g.mp_x, g.mp_y)
          for g in G.layout_watermark(W_IMG, H_IMG, WM_TEXT)]
    col, _ = G.parse_color(WM_COLOR, WM_OPACITY)
    for _ in range(min(args.warmup, 1)):
        O.bench_batch(rasters[:cores], cores, 7, RW, RH, True, THUMB, col, gl)
    secs = 0.0
    for _ in range(args.steps):
        s, _ = O.bench_batch(rasters, cores, 7, RW, RH, True, THUMB, col, gl)
        secs += s
    value = n * args.steps / secs
    sample = f"{n} images per step (one per host thread), {args.steps} steps"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus, args.images),   # the CUDA arm's config; the bounded sample is in cpu_baseline.sample
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "restated reference CPU path (C), not the Go build"},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def bind_near_gpu(index):
    """Best effort, N > 1 only: keep this rank's pinned host buffers on the NUMA node of its GPU (memory policy
    first -- it works even when the allowed cores are all on the other socket -- then the node's cores), so 8
    ranks do not push their H2D/D2H traffic across the socket interconnect."""
    info = {}
    try:
        bus = subprocess.check_output(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                                      text=True, timeout=20).strip().lower()
        dom, rest = bus.split(":", 1)
        dev = f"/sys/bus/pci/devices/{dom[-4:]}:{rest}"
        node = int(open(dev + "/numa_node").read().strip())
        info["gpu_numa_node"] = node
        if node >= 0:
            import ctypes
            libc = ctypes.CDLL(None, use_errno=True)
            mask = ctypes.c_ulong(1 << node)
            MPOL_PREFERRED = 1
            rc = libc.syscall(238, MPOL_PREFERRED, ctypes.byref(mask), ctypes.c_ulong(64))  # set_mempolicy
            info["mempolicy"] = "preferred" if rc == 0 else f"errno {ctypes.get_errno()}"
        cpus = set()
        for part in open(dev + "/local_cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        info["allowed_cpus"], info["local_allowed_cpus"] = len(allowed), len(use)
        if use and use != allowed:
            os.sched_setaffinity(0, use)
            info["affinity"] = "narrowed"
    except Exception as e:  # noqa: BLE001
        info["error"] = f"{type(e).__name__}: {e}"
    return info


def workload_config(n_gpus, images_per_gpu):
    return {
        "workload": "BASELINE configs[2]: batch of 12MP (4000x3000) synthetic RGBA images, "
                    "resize 1024x768 keep-aspect + thumbnail 200 crop + watermark '(c) ImageProcessor'",
        "images_per_gpu_per_step": images_per_gpu, "image": f"{W_IMG}x{H_IMG} RGBA8", "ops": "resize+thumb+watermark",
        "sharding": f"by image, {n_gpus} rank(s), no collective",
        "l2_policy": "inputs larger than L2 (each step streams >= 12 GB of distinct sources per GPU)",
        "precision": "EXACT (fp32 stream + fp64 fix-up: byte-identical to the fp64 reference algorithm)",
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--images", type=int, default=env_int("IPG_BENCH_IMAGES", 256), help="images per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = env_int("RANK", 0)
    world = env_int("WORLD_SIZE", 1)
    local_rank = env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import imageprocessor_b200 as ip
    from imageprocessor_b200 import _lib as L
    from imageprocessor_b200 import glyphs as G

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_near_gpu(local_rank) if world > 1 else "single rank: not bound"
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from imageprocessor_b200 import sharding as S

    def barrier():
        S.barrier()
        torch.cuda.synchronize()

    max_over_ranks, sum_over_ranks = S.reduce_max, S.reduce_sum

    n_img = args.images
    lib = L.load()
    eng = ip.Engine(devices=[local_rank], precision=ip.PRECISION_EXACT, lanes_per_device=3,
                    max_batch=env_int("IPG_BENCH_MAX_BATCH", 128), batch_window_us=100)
    ctx = eng._ctx

    # ---- synthetic sources, resident in HBM (seeded per image; A=255)
    gen = torch.Generator(device=dev)
    srcs = []
    for i in range(n_img):
        gen.manual_seed(1000 + rank * n_img + i)
        t = torch.randint(0, 256, (H_IMG, W_IMG, 4), dtype=torch.uint8, device=dev, generator=gen)
        t[..., 3] = 255
        srcs.append(t)
    nw, nh = ip.keep_aspect_dims(W_IMG, H_IMG, RW, RH)
    cx, cy, cs = ip.crop_square(W_IMG, H_IMG)
    out_r = torch.empty((n_img, nh, nw, 4), dtype=torch.uint8, device=dev)
    out_t = torch.empty((n_img, THUMB, THUMB, 4), dtype=torch.uint8, device=dev)
    out_w = torch.empty((n_img, H_IMG, W_IMG, 4), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()

    glyph_list = G.layout_watermark(W_IMG, H_IMG, WM_TEXT, "bottom-right", 36.0)
    color, _ = G.parse_color(WM_COLOR, WM_OPACITY)
    garr = (L.Glyph * len(glyph_list))()
    gkeep = []
    for j, g in enumerate(glyph_list):
        m = np.ascontiguousarray(g.mask, np.uint8)
        gkeep.append(m)
        garr[j].x0, garr[j].y0, garr[j].x1, garr[j].y1 = g.x0, g.y0, g.x1, g.y1
        garr[j].mp_x, garr[j].mp_y = g.mp_x, g.mp_y
        garr[j].mask_w, garr[j].mask_h, garr[j].mask_stride = m.shape[1], m.shape[0], m.strides[0]
        garr[j].mask = m.ctypes.data

    def make_ops(dst_r, dst_t, dst_w, memspace):
        ops = (L.Op * 3)()
        ops[0].kind, ops[0].dst_w, ops[0].dst_h = L.OP_RESIZE, nw, nh
        ops[0].dst, ops[0].dst_stride, ops[0].dst_memspace = dst_r, nw * 4, memspace
        ops[1].kind, ops[1].dst_w, ops[1].dst_h = L.OP_THUMB_CROP, THUMB, THUMB
        ops[1].rect_x, ops[1].rect_y, ops[1].rect_w, ops[1].rect_h = cx, cy, cs, cs
        ops[1].dst, ops[1].dst_stride, ops[1].dst_memspace = dst_t, THUMB * 4, memspace
        ops[2].kind, ops[2].dst_w, ops[2].dst_h = L.OP_WATERMARK, W_IMG, H_IMG
        for k in range(4):
            ops[2].color[k] = color[k]
        ops[2].n_glyphs, ops[2].glyphs = len(glyph_list), garr
        ops[2].dst, ops[2].dst_stride, ops[2].dst_memspace = dst_w, W_IMG * 4, memspace
        return ops

    def make_desc(ptr, memspace):
        d = L.ImageDesc()
        d.layout, d.memspace, d.width, d.height = L.RGBA8, memspace, W_IMG, H_IMG
        d.plane[0] = ptr
        d.stride[0] = W_IMG * 4
        d.opaque_hint = 0   # *image.RGBA: alpha unknown to the caller, as in the reference
        return d

    dev_descs = [make_desc(srcs[i].data_ptr(), L.MEM_DEVICE) for i in range(n_img)]
    dev_ops = [make_ops(out_r[i].data_ptr(), out_t[i].data_ptr(), out_w[i].data_ptr(), L.MEM_DEVICE) for i in range(n_img)]
    # two ticket sets: a worker keeps submitting, so step k+1 is submitted before step k is waited for (the
    # batcher never starves between steps); every step's submissions and completions lie inside the timed region
    tids = [(C.c_uint64 * n_img)() for _ in range(2)]
    tid_ref = [[C.cast(C.byref(t, 8 * i), C.POINTER(C.c_uint64)) for i in range(n_img)] for t in tids]
    submit_on, wait = lib.ipg_submit_on, lib.ipg_wait

    def submit_step(k):
        refs = tid_ref[k & 1]
        for i in range(n_img):
            rc = submit_on(ctx, 0, C.byref(dev_descs[i]), dev_ops[i], 3, refs[i])
            if rc:
                L.check(rc)

    def wait_step(k):
        t = tids[k & 1]
        for i in range(n_img):
            rc = wait(ctx, t[i], -1)
            if rc:
                L.check(rc)

    def run_steps(n):
        for k in range(n):
            submit_step(k)
            if k:
                wait_step(k - 1)
        if n:
            wait_step(n - 1)

    # ---- device-resident throughput
    run_steps(args.warmup)
    barrier()
    eng.reset_stats()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t0 = time.perf_counter()
    run_steps(args.steps)
    eng.flush()
    barrier()
    wall_dev = time.perf_counter() - t0
    clocks = sampler.stop()
    st = eng.stats()
    span_s = max_over_ranks(st["kernel_span_ms"] / 1e3)
    total_images = sum_over_ranks(float(n_img * args.steps))
    value = total_images / span_s
    launches = int(sum_over_ranks(float(st["kernels_launched"])))
    stream_ms_per_image = st["stream_kernel_ms"] / (n_img * args.steps)
    peak, peak_src = measured_peak()
    traffic = ncu_traffic()

    # ---- the dominant kernel.  In the timed run above every batch is one k_stream<1,true,4> launch that holds
    # both passes of every image (resize + watermark copy; thumbnail); its duration is st["stream_fast_kernel_ms"].
    # Each pass alone is timed in a short extra run with the merge and the stream overlap switched off.
    lean_ms_run = st["stream_fast_kernel_ms"]
    launches_run = max(st["batches"], 1)
    imgs_per_launch = n_img * args.steps / launches_run
    merged_GBps = BYTES_PER_IMAGE * n_img * args.steps / (lean_ms_run * 1e-3) / 1e9
    os.environ["IPG_NO_OVERLAP"] = "1"
    os.environ["IPG_MERGE_LEAN"] = "0"
    eng_iso = ip.Engine(devices=[local_rank], precision=ip.PRECISION_EXACT, lanes_per_device=1,
                        max_batch=env_int("IPG_BENCH_MAX_BATCH", 128), batch_window_us=2000)
    del os.environ["IPG_NO_OVERLAP"], os.environ["IPG_MERGE_LEAN"]
    n_iso = min(n_img, 64)
    for rep in range(3):           # 2 warm-up passes, 1 timed
        if rep == 2:
            eng_iso.reset_stats()
        for i in range(n_iso):
            L.check(submit_on(eng_iso._ctx, 0, C.byref(dev_descs[i]), dev_ops[i], 3, tid_ref[0][i]))
        for i in range(n_iso):
            L.check(wait(eng_iso._ctx, tids[0][i], -1))
    iso = eng_iso.stats()
    eng_iso.close()
    pass_a_ms, both_ms = iso["stream_fast_kernel_ms"], iso["stream_kernel_ms"]
    pass_a_GBps = BYTES_LEAN_PASS * n_iso / (pass_a_ms * 1e-3) / 1e9
    pass_b_GBps = BYTES_THUMB_PASS * n_iso / (max(both_ms - pass_a_ms, 1e-9) * 1e-3) / 1e9
    roofline = {
        "bound": "hbm",
        "kernel": "k_stream<1,true,4> (lean): per image a resize + watermark-copy pass and a thumbnail pass, all in one launch",
        "achieved": merged_GBps, "peak": peak, "unit": "GB/s", "frac": merged_GBps / peak, "peak_source": peak_src,
        "frac_of_nominal_8TBs": merged_GBps / 8000.0,
        "algorithmic_bytes_per_image": BYTES_PER_IMAGE,
        "algorithmic_note": "SURVEY.md 8(d) pipeline-min figure: the source counted ONCE although the thumbnail pass reads its "
                            "crop square a second time (DESIGN.md 4.1.3: two lean passes beat every fused form measured)",
        "images_per_launch": imgs_per_launch, "ms_per_launch": lean_ms_run / launches_run,
        "how": "CUDA events on the launching stream around the kernel, summed over the K timed steps",
        "traffic": traffic["dram_bytes_per_image"] * imgs_per_launch if traffic else None,
        "traffic_note": (traffic.get("note") + f"; {traffic['dram_bytes_per_image'] / 1e6:.1f} MB per image x images_per_launch")
                        if traffic else "no ncu --set full capture committed yet",
        "kernel_share_of_step": lean_ms_run / max(st["kernel_ms"], 1e-9),
        "passes_timed_alone": {
            "how": f"second engine, IPG_MERGE_LEAN=0 IPG_NO_OVERLAP=1, {n_iso} device-resident images: each pass is its own launch",
            "resize+watermark_copy": {"kernel": "k_stream<1,true,1>", "algorithmic_bytes_per_image": BYTES_LEAN_PASS,
                                      "achieved": pass_a_GBps, "frac": pass_a_GBps / peak, "us_per_image": 1e3 * pass_a_ms / n_iso},
            "thumbnail": {"kernel": "k_stream<1,false,2>", "algorithmic_bytes_per_image": BYTES_THUMB_PASS,
                          "achieved": pass_b_GBps, "frac": pass_b_GBps / peak, "us_per_image": 1e3 * (both_ms - pass_a_ms) / n_iso}},
        "stream_us_per_image": 1e3 * stream_ms_per_image,
        "fix_kernel_ms_per_step": st["fix_kernel_ms"] / args.steps,
        "exact_fixups_per_image": st["exact_fixups"] / max(n_img * args.steps, 1),
    }

    # ---- verification of the timed outputs against the oracle (rank 0, one image)
    verified = None
    if rank == 0 and not args.no_verify:
        from oracle import oracle as O
        a = srcs[0].cpu().numpy()
        R = O.Raster.rgba(a)
        ok_r = np.array_equal(out_r[0].cpu().numpy(), O.resize_image(R, nw, nh))
        ok_t = np.array_equal(out_t[0].cpu().numpy(), O.crop_and_resize(R, THUMB))
        ogl = [O.Glyph(g.x0, g.y0, g.x1, g.y1, g.mask, g.mp_x, g.mp_y) for g in glyph_list]
        ok_w = np.array_equal(out_w[0].cpu().numpy(), O.watermark(R, color, ogl))
        verified = {"resize_bit_exact": bool(ok_r), "thumb_bit_exact": bool(ok_t), "watermark_bit_exact": bool(ok_w)}

    # ---- end to end through the C ABI with pinned host buffers
    e2e = None
    if not args.no_e2e:
        # a separate context tuned for the PCIe-bound path: smaller batches, more lanes, so
        # the H2D of one batch overlaps the D2H of another
        eng.close()
        eng = ip.Engine(devices=[local_rank], precision=ip.PRECISION_EXACT, lanes_per_device=4,
                        max_batch=env_int("IPG_BENCH_E2E_BATCH", 8), batch_window_us=100)
        ctx = eng._ctx
        n_slots = min(env_int("IPG_BENCH_HOST_SLOTS", 64), n_img)
        src_bytes = W_IMG * H_IMG * 4
        pins = []
        h_descs, h_ops = [], []
        for s in range(n_slots):
            p_in = eng.alloc_pinned(src_bytes)
            p_in.array[:] = srcs[s].cpu().numpy().reshape(-1)
            p_r, p_t, p_w = eng.alloc_pinned(nw * nh * 4), eng.alloc_pinned(THUMB * THUMB * 4), eng.alloc_pinned(src_bytes)
            pins += [p_in, p_r, p_t, p_w]
            h_descs.append(make_desc(p_in.ptr, L.MEM_HOST))
            h_ops.append(make_ops(p_r.ptr, p_t.ptr, p_w.ptr, L.MEM_HOST))
        submit = lib.ipg_submit_on

        # a ring of n_slots pinned slots: slot s is re-submitted as soon as its previous ticket is done.  The K steps
        # run as one stream of K * n_img submissions (a worker does not drain its pipeline between batches); every
        # submission, copy and completion lies inside the timed region.
        slot_tid = (C.c_uint64 * n_slots)()
        slot_ref = [C.cast(C.byref(slot_tid, 8 * s), C.POINTER(C.c_uint64)) for s in range(n_slots)]

        def run_host(n_steps):
            total = n_steps * n_img
            for j in range(total):
                s = j % n_slots
                if j >= n_slots:          # bounded in flight: slot s is free once its previous ticket is done
                    rc = wait(ctx, slot_tid[s], -1)
                    if rc:
                        L.check(rc)
                rc = submit(ctx, 0, C.byref(h_descs[s]), h_ops[s], 3, slot_ref[s])
                if rc:
                    L.check(rc)
            for s in range(min(n_slots, total)):
                rc = wait(ctx, slot_tid[s], -1)
                if rc:
                    L.check(rc)

        del out_w
        torch.cuda.empty_cache()
        run_host(max(1, min(args.warmup, 3)))
        barrier()
        eng.reset_stats()
        t0 = time.perf_counter()
        run_host(args.steps)
        eng.flush()
        barrier()
        wall = max_over_ranks(time.perf_counter() - t0)
        st2 = eng.stats()
        e2e = {
            "value": total_images / wall, "unit": "images/s",
            "h2d_bytes_per_step": int(st2["bytes_h2d"] / args.steps), "d2h_bytes_per_step": int(st2["bytes_d2h"] / args.steps),
            "timing": "host wall clock around K steps submitted as one stream (barrier + flush both sides), max over ranks",
            "device_span_s": st2["batch_span_ms"] / 1e3,
            "h2d_GBps": st2["bytes_h2d"] / wall / 1e9, "d2h_GBps": st2["bytes_d2h"] / wall / 1e9,
            "host_buffers": f"{n_slots} pinned slots per rank (ipg_alloc_pinned), zero staging copies: {st2['staged_copies'] == 0}",
            "verified_slot0": None,
        }
        if rank == 0 and not args.no_verify:
            from oracle import oracle as O
            a = pins[0].array.reshape(H_IMG, W_IMG, 4)
            e2e["verified_slot0"] = bool(np.array_equal(pins[1].array.reshape(nh, nw, 4),
                                                        O.resize_image(O.Raster.rgba(a), nw, nh)))
        for p in pins:
            p.free()

    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        cores = os.cpu_count() or 1
        n_s = env_int("IPG_BENCH_CPU_IMAGES", 8 * cores)
        host_imgs = [srcs[i % n_img].cpu().numpy() for i in range(min(n_s, 4))]
        rasters = [O.Raster.rgba(host_imgs[i % len(host_imgs)]) for i in range(n_s)]
        ogl = [O.Glyph(g.x0, g.y0, g.x1, g.y1, g.mask, g.mp_x, g.mp_y) for g in glyph_list]
        secs, _ = O.bench_batch(rasters, cores, 7, RW, RH, True, THUMB, color, ogl)
        cpu = {"value": n_s / secs, "unit": "images/s", "cores": cores, "kind": "port",
               "sample": f"{n_s} images of the same workload, one per host thread, one pass ({secs:.1f} s)",
               "note": "restated reference CPU path (C, float64 scalar), not the Go build"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * span_s / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world, n_img),
            "timing": "CUDA events on the launching streams: first kernel start -> last kernel end over the K steps, max over ranks; "
                      "step k+1 is submitted before step k is waited for (a worker keeps submitting), all inside the timed region",
            "wall_ms_per_step": 1e3 * wall_dev / args.steps,
            "hbm_GBps": value * BYTES_PER_IMAGE / 1e9 / world,
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
            "cpu_baseline": cpu, "e2e": e2e, "verified": verified,
            "host": {"nproc": os.cpu_count(), "numa_binding_rank0": numa},
        }
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
