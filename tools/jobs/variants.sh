#!/bin/bash
# GPU-box job: time several builds of the library (kernel experiments) on the same op mixes.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
for lib in "$@"; do
  for ops in r rt rtw; do
    echo -n "$lib $ops: "
    IPG_LIB_PATH=$PWD/imageprocessor_b200/$lib timeout 60 python tools/profile_step.py --images 32 --steps 3 --ops $ops --lanes 1 | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(round(d['stream_us_per_image'],2), 'us/img', round(d['stream_GBps']), 'GB/s fix', round(d['fix_us_per_image'],2))"
  done
done
