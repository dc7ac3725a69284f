#!/bin/bash
# GPU-box job: the integer-moment thumbnail pass (IPG_VINT) against the fp32 form -- its parity test first, then A/B timing of
# the merged lean launch (profile_step, device-resident: 64 x 12 MP r+t+w / r+t / t, 24 x 8K r+t+w / t), then the whole gpu suite.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "integer_moment" 2>&1 | tail -5
show='import json,sys; d=json.loads(sys.stdin.readline()); print(round(d["stream_us_per_image"],2), "us/img", round(d["stream_GBps"]), "GB/s fix", round(d["fix_us_per_image"],2), "fixups/img", round(d["fixups_per_image"],1))'
for v in 0 1; do
  for ops in rtw rt t; do
    echo -n "IPG_VINT=$v 12MP $ops: "
    IPG_VINT=$v timeout 120 python tools/profile_step.py --images 64 --steps 4 --ops $ops --lanes 1 --max-batch 64 | python -c "$show"
  done
  for ops in rtw t; do
    echo -n "IPG_VINT=$v 8K $ops: "
    IPG_VINT=$v timeout 120 python tools/profile_step.py --images 24 --steps 3 --ops $ops --lanes 1 --max-batch 24 --w 7680 --h 4320 | python -c "$show"
  done
done 2>&1 | tee gpurun_out/vint_ab.log
if [ "$1" != "quick" ]; then timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5; fi
