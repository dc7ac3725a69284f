#!/bin/bash
# GPU-box job (1 GPU): the whole gpu test suite, then the default bench line and the reference arm.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r2c}
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
( time timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err ) 2>&1 | grep real; echo bench rc=$?
tail -3 gpurun_out/bench_$TAG.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("value", round(d["value"]), "frac", round(d["roofline"]["frac"], 3), "e2e", round(e["value"]), "pcie", round(e["pcie"]["frac_of_pcie"], 3), d["verified"], e["verified_slot0"])
for k in ("watermark_patched_in_place", "results_as_ycbcr420", "results_as_jpeg_files"):
    v = e[k]; print(" ", k, round(v["value"]), "down GB/s", round(v["d2h_GBps_aggregate"], 1), [v[q] for q in v if q.startswith("verified")], v.get("file_bytes_slot0"))
c = d.get("configs") or {}
if "error" in c: print(c["error"], c.get("traceback"))
if "c5" in c:
    for k in ("end_to_end_with_codecs", "end_to_end_device_jpeg_encode", "end_to_end_device_jpeg_encode_all_targets_jpeg", "raster_only_decoded_inputs_no_encode"):
        a = c["c5"][k]; print("  c5", k, "img/s", round(a["images_per_s"], 1), "MP/s", round(a["megapixels_per_s"]), "ms/img", {q: round(x, 1) for q, x in a["rank0_mean_ms_per_image"].items()}, "failed", a["rank0_failed"])
    print("  c5 verified", c["c5"]["verified"]["all_bit_exact"], c["c5"]["verified"].get("device_jpeg_files_byte_identical"))
print("cpu", d["cpu_baseline"]["value"] if d.get("cpu_baseline") else None)
PY
