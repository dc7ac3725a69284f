#!/bin/bash
# GPU-box job (1 GPU): the default bench line (all sub-records), then the ncu launch list of the short bench command.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r2e}
( time timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err ) 2>&1 | grep real; echo bench rc=$?
tail -3 gpurun_out/bench_$TAG.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
e = d["e2e"]; r = d["roofline"]
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "frac", round(r["frac"], 3), "alone", {k: round(v["frac"], 3) for k, v in r["passes_timed_alone"].items() if isinstance(v, dict)},
      "stream us/img", round(r["stream_us_per_image"], 2), "fix ms/step", round(r["fix_kernel_ms_per_step"], 3), "fixups/img", round(r["exact_fixups_per_image"], 1))
print("e2e", round(e["value"]), "pcie", round(e["pcie"]["frac_of_pcie"], 3), d["verified"], e["verified_slot0"])
for k in ("watermark_patched_in_place", "results_as_ycbcr420", "results_as_jpeg_files"):
    v = e[k]; print(" ", k, round(v["value"]), [v[q] for q in v if q.startswith("verified")])
c = d.get("configs") or {}
if "error" in c: print(c["error"], c.get("traceback"))
if "c4" in c: print("  c4", json.dumps(c["c4"])[:600])
if "c5" in c:
    for k in ("end_to_end_with_codecs", "end_to_end_device_jpeg_encode", "end_to_end_device_jpeg_encode_all_targets_jpeg", "raster_only_decoded_inputs_no_encode"):
        a = c["c5"][k]; print("  c5", k, "img/s", round(a["images_per_s"], 1), "failed", a["rank0_failed"])
    print("  c5 verified", c["c5"]["verified"]["all_bit_exact"], c["c5"]["verified"].get("device_jpeg_files_byte_identical"))
print("cpu", d["cpu_baseline"]["value"] if d.get("cpu_baseline") else None, "clocks", d["clocks"])
PY
BENCH="python bench.py --steps 1 --warmup 3 --images 64 --no-e2e --no-cpu-baseline --no-verify --no-configs"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv $BENCH > gpurun_out/ncu_launches_$TAG.log 2>&1; echo launches rc=$?
