#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for lib in "$@"; do
  for ops in r rw; do
    echo -n "$lib $ops: "
    IPG_LIB_PATH=$PWD/imageprocessor_b200/$lib timeout 120 python tools/profile_step.py --images 32 --steps 3 --ops $ops --lanes 1 --precision 1 | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(round(d['stream_us_per_image'],2), 'us/img', round(d['stream_GBps']), 'GB/s')"
  done
done
