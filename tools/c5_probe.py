#!/usr/bin/env python
"""bench.py's configs.c5 (mixed 0.3-48 MP stream through the streaming worker) on its own, one GPU: prints the four
legs' images/s.  For A/B runs of engine knobs (IPG_VINT=0 python tools/c5_probe.py)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import imageprocessor_b200 as ip

c = bench.config_c5(ip, 0, 1, 0, lambda: None, lambda x: x, lambda x: x)
out = {}
for k in ("end_to_end_with_codecs", "end_to_end_device_jpeg_encode", "end_to_end_device_jpeg_encode_all_targets_jpeg",
          "raster_only_decoded_inputs_no_encode"):
    a = c[k]
    out[k] = {"images_per_s": round(a["images_per_s"], 1), "wall_s": round(a["wall_s"], 3),
              "kernel_ms": round(a["rank0_engine"]["kernel_ms"], 2), "batches": a["rank0_engine"]["batches"]}
out["verified"] = c["verified"]["all_bit_exact"]
print(json.dumps(out))
