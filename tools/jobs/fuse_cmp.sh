#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for f in 0 1; do for ops in rt rtw; do
  echo -n "12MP fuse=$f $ops: "
  timeout 120 python tools/profile_step.py --images 32 --steps 3 --ops $ops --lanes 1 --fuse $f | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(round(d['stream_us_per_image'],2), 'us/img', round(d['stream_GBps']), 'GB/s fix', round(d['fix_us_per_image'],2))"
done; done
for f in 0 1; do
  echo -n "8K fuse=$f rtw: "
  timeout 300 python tools/profile_step.py --images 12 --steps 2 --ops rtw --lanes 1 --w 7680 --h 4320 --fuse $f | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(round(d['stream_us_per_image'],2), 'us/img', round(d['stream_GBps']), 'GB/s fix', round(d['fix_us_per_image'],2))"
done
