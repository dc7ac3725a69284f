#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
IPG_MERGE_LEAN=1 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for m in 0 1 0 1; do for ops in rt rtw; do
  echo -n "12MP merge=$m $ops: "
  IPG_MERGE_LEAN=$m timeout 120 python tools/profile_step.py --images 32 --steps 3 --ops $ops --lanes 1 | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(round(d['stream_us_per_image'],2), 'us/img', round(d['stream_GBps']), 'GB/s fix', round(d['fix_us_per_image'],2))"
done; done
for m in 0 1; do
  echo -n "8K merge=$m rtw: "
  IPG_MERGE_LEAN=$m timeout 300 python tools/profile_step.py --images 12 --steps 2 --ops rtw --lanes 1 --w 7680 --h 4320 | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(round(d['stream_us_per_image'],2), 'us/img', round(d['stream_GBps']), 'GB/s fix', round(d['fix_us_per_image'],2))"
done
