#!/bin/bash
# GPU-box job: A/B of library builds on the merged lean launch (profile_step, 64 x 12 MP device-resident, r+t+w and r+t), three runs each
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for rep in 1 2; do
for lib in "$@"; do
  for ops in rtw rt; do
    echo -n "$lib $ops: "
    IPG_LIB_PATH=$PWD/imageprocessor_b200/$lib timeout 120 python tools/profile_step.py --images 64 --steps 4 --ops $ops --lanes 1 --max-batch 64 | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(round(d['stream_us_per_image'],2), 'us/img', round(d['stream_GBps']), 'GB/s fix', round(d['fix_us_per_image'],2), 'fixups', d.get('exact_fixups'))"
  done
done
done
