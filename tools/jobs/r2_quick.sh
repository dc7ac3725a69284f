#!/bin/bash
# GPU-box job (1 GPU): selected gpu tests ($1 = -k expression, default all) then a quick bench line.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q ${1:+-k "$1"} 2>&1 | tail -8
timeout 600 python bench.py --steps ${STEPS:-10} --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo bench rc=$?; tail -3 gpurun_out/bench_quick.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_quick.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "frac", round(r["frac"], 3), "thumb alone", round(r["passes_timed_alone"]["thumbnail"]["frac"], 3),
      "passA alone", round(r["passes_timed_alone"]["resize+watermark_copy"]["frac"], 3), "fix ms/step", round(r["fix_kernel_ms_per_step"], 3), d["verified"])
e = d["e2e"]
print("e2e", round(e["value"]), "pcie", e["pcie"]["achieved"], e["pcie"]["peak"], e["pcie"]["frac_of_pcie"], e["verified_slot0"])
print("in place", e.get("watermark_patched_in_place"))
PY
