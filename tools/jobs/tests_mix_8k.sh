#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for ops in rw t rtw; do
  echo -n "12MP $ops: "
  timeout 120 python tools/profile_step.py --images 32 --steps 3 --ops $ops --lanes 1 | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(round(d['stream_us_per_image'],2), 'us/img', round(d['stream_GBps']), 'GB/s fix', round(d['fix_us_per_image'],2))"
done
bash tools/jobs/k8.sh
