#!/bin/bash
# GPU-box job (gpurun --gpus 2): the whole -m gpu suite incl. the one-context-all-devices tests, then the
# host<->device scaling microbenchmark at N = 1, 2.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
nproc; free -g | head -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 tools/micro/pcie_scale --out gpurun_out/pcie_scaling_2gpu.json; echo pcie rc=$?
python - <<'PY'
import json
d = json.load(open("gpurun_out/pcie_scaling_2gpu.json"))
print({k: d[k] for k in d if k != "results"})
for r in d["results"]:
    if r.get("ok"):
        print(f'{r["model"][:7]:8s}{r["alloc"][:28]:30s}{r["dir"]:5s} N={r["n_devices"]} chunk={r["chunk_mb"]:.0f}  up {r["h2d_GBps_aggregate"]:6.1f}  down {r["d2h_GBps_aggregate"]:6.1f}  per-dev {r["per_device_GBps_min"]:.1f}..{r["per_device_GBps_max"]:.1f}')
    else:
        print(r)
PY
