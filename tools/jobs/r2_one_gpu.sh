#!/bin/bash
# GPU-box job (1 GPU): -m gpu suite, then the full bench line (with the c1/c4/c5 sub-records) and the reference arm.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nproc; free -g | head -2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; echo bench rc=$?
tail -5 gpurun_out/bench_r2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r2.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks", "verified")})
r = d["roofline"]; print("roofline", r["achieved"], r["frac"], r["passes_timed_alone"]["thumbnail"]["frac"], r["fix_kernel_ms_per_step"], r["exact_fixups_per_image"])
print("e2e", d["e2e"])
print("cpu", d["cpu_baseline"])
print(json.dumps(d.get("configs"), indent=1)[:6000])
PY
