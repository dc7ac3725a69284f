#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for ops in r t rw rtw; do
  echo -n "8K $ops: "
  timeout 300 python tools/profile_step.py --images 12 --steps 2 --ops $ops --lanes 1 --w 7680 --h 4320 | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(round(d['stream_us_per_image'],2), 'us/img', round(d['stream_GBps']), 'GB/s fix', round(d['fix_us_per_image'],2), 'launches', d['launches'])"
done
