#!/bin/bash
# GPU-box job (gpurun --gpus NG): the final code's bench under torchrun at NG GPUs (with the config sub-records and the
# opt-in e2e variants), then the one-context-all-devices tests.
NG=${1:-8}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29519 \
    bench.py --gpus $NG --steps 5 --warmup 3 > gpurun_out/bench_final_${NG}gpu.json 2> gpurun_out/bench_final_${NG}gpu.err; echo torchrun rc=$?
tail -3 gpurun_out/bench_final_${NG}gpu.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_final_${NG}gpu.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("value", round(d["value"]), "e2e", round(e["value"]), "pcie", round(e["pcie"]["achieved"], 1), "/", round(e["pcie"]["peak"], 1), "=", round(e["pcie"]["frac_of_pcie"], 3), d["verified"], e["verified_slot0"])
print(" in place", round(e["watermark_patched_in_place"]["value"]), "up", round(e["watermark_patched_in_place"]["h2d_GBps_aggregate"], 1), e["watermark_patched_in_place"]["verified_first_submission_all_outputs"])
print(" ycbcr420", round(e["results_as_ycbcr420"]["value"]), "up", round(e["results_as_ycbcr420"]["h2d_GBps_aggregate"], 1), "down", round(e["results_as_ycbcr420"]["d2h_GBps_aggregate"], 1), e["results_as_ycbcr420"]["verified_slot0_all_planes"])
c = d.get("configs") or {}
if "error" in c: print(c["error"], c.get("traceback"))
if "c4" in c: print(" c4 value", round(c["c4"]["value"]), "e2e", round(c["c4"]["e2e"]["value"]), c["c4"]["verified"], c["c4"]["verified_e2e_slot0_all_outputs"])
if "c5" in c:
    a, b = c["c5"]["end_to_end_with_codecs"], c["c5"]["raster_only_decoded_inputs_no_encode"]
    print(" c5 with codecs img/s", round(a["images_per_s"], 1), "MP/s", round(a["megapixels_per_s"]), "| raster only img/s", round(b["images_per_s"], 1), "MP/s", round(b["megapixels_per_s"]), c["c5"]["verified"]["all_bit_exact"] if c["c5"]["verified"] else None)
PY
timeout 300 python -m pytest tests/test_multi_device.py -m gpu -x -q 2>&1 | tail -3
