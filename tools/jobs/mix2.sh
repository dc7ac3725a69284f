#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for ov in 0 1; do for ops in rt rtw; do
  echo -n "no_overlap=$ov $ops: "
  IPG_NO_OVERLAP=$ov timeout 120 python tools/profile_step.py --images 32 --steps 3 --ops $ops --lanes 1 | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(round(d['stream_us_per_image'],2), 'us/img', round(d['stream_GBps']), 'GB/s fix', round(d['fix_us_per_image'],2))"
done; done
