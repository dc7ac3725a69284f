#!/bin/bash
# GPU-box job (gpurun --gpus 2): the one-context-all-devices tests, then the final code's bench under torchrun at 2 GPUs
# (every sub-record and opt-in e2e variant) and in the one-process-all-devices model.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_multi_device.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_final_2gpu.json 2> gpurun_out/bench_final_2gpu.err; echo torchrun rc=$?
tail -2 gpurun_out/bench_final_2gpu.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_final_2gpu.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("value", round(d["value"]), "e2e", round(e["value"]), "pcie", round(e["pcie"]["frac_of_pcie"], 3), d["verified"], e["verified_slot0"])
for k in ("watermark_patched_in_place", "results_as_ycbcr420", "results_as_jpeg_files"):
    v = e[k]; print(" ", k, round(v["value"]), [v[q] for q in v if q.startswith("verified")])
c = d.get("configs") or {}
if "error" in c: print(c["error"], c.get("traceback"))
if "c5" in c:
    for k in ("end_to_end_with_codecs", "end_to_end_device_jpeg_encode", "end_to_end_device_jpeg_encode_all_targets_jpeg", "raster_only_decoded_inputs_no_encode"):
        a = c["c5"][k]; print("  c5", k, "img/s", round(a["images_per_s"], 1), "failed", a["rank0_failed"])
    print("  c5 verified", c["c5"]["verified"]["all_bit_exact"], c["c5"]["verified"].get("device_jpeg_files_byte_identical"))
PY
timeout 600 python bench.py --gpus 2 --single-process --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_single_2gpu_final.json 2> gpurun_out/bench_single_2gpu_final.err; echo single rc=$?
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_single_2gpu_final.json").read().strip().splitlines()[-1])
print("single-process value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "jpeg", round(d["e2e"]["results_as_jpeg_files"]["value"]), d["e2e"]["results_as_jpeg_files"]["verified_slot0_all_files_byte_identical"])
PY
