#!/bin/bash
# GPU-box job (gpurun --gpus NG): host<->device ceiling of the box at N = 1..NG, then the bench at NG GPUs in both
# process models (one rank per GPU under torchrun; one process driving every device), then the multi-device tests.
NG=${1:-2}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc; free -g | head -2 | tail -1
nvidia-smi topo -m > gpurun_out/topo_${NG}gpu.txt 2>&1
timeout 300 tools/micro/pcie_scale --total-mb 2048 --out gpurun_out/pcie_scaling_${NG}gpu.json; echo pcie rc=$?
python - <<PY
import json
d = json.load(open("gpurun_out/pcie_scaling_${NG}gpu.json"))
for r in d["results"]:
    if r.get("ok") and r["alloc"] in ("cudaHostAlloc", "mmapHUGETLB2M+cudaHostRegister") and r["chunk_mb"] > 40:
        print(f'{r["model"][:7]:8s}{r["alloc"][:14]:15s}{r["dir"]:5s} N={r["n_devices"]}  up {r["h2d_GBps_aggregate"]:6.1f}  down {r["d2h_GBps_aggregate"]:6.1f}  per-dev {r["per_device_GBps_min"]:.1f}..{r["per_device_GBps_max"]:.1f}')
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $NG --steps 5 --warmup 3 > gpurun_out/bench_torchrun_${NG}gpu.json 2> gpurun_out/bench_torchrun_${NG}gpu.err; echo torchrun rc=$?
tail -3 gpurun_out/bench_torchrun_${NG}gpu.err
timeout 600 python bench.py --single-process --gpus $NG --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_single_${NG}gpu.json 2> gpurun_out/bench_single_${NG}gpu.err; echo single rc=$?
tail -3 gpurun_out/bench_single_${NG}gpu.err
python - <<PY
import json
for name in ("torchrun", "single"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{name}_${NG}gpu.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(name, "no line", e); continue
    e = d["e2e"]
    print(name, "value", round(d["value"]), "e2e", round(e["value"]), "up/down agg GB/s", round(e["h2d_GBps_aggregate"], 1), round(e["d2h_GBps_aggregate"], 1), "verified", d["verified"], e["verified_slot0"])
    c = d.get("configs") or {}
    if "error" in c: print(c["error"], c.get("traceback"))
    if "c4" in c: print(" c4 value", round(c["c4"]["value"]), "e2e", round(c["c4"]["e2e"]["value"]), c["c4"]["verified"], c["c4"]["verified_e2e_slot0_all_outputs"])
    if "c5" in c:
        a, b = c["c5"]["end_to_end_with_codecs"], c["c5"]["raster_only_decoded_inputs_no_encode"]
        print(" c5 with codecs img/s", round(a["images_per_s"], 1), "MP/s", round(a["megapixels_per_s"]), "| raster only img/s", round(b["images_per_s"], 1), "MP/s", round(b["megapixels_per_s"]), c["c5"]["verified"]["all_bit_exact"] if c["c5"]["verified"] else None)
PY
timeout 300 python -m pytest tests/test_multi_device.py -m gpu -x -q 2>&1 | tail -3
