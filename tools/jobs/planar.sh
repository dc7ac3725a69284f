#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for lay in ycbcr420 ycbcr444; do for ops in r t rt; do
  echo -n "12MP $lay $ops: "
  timeout 120 python tools/profile_step.py --images 32 --steps 3 --ops $ops --lanes 1 --layout $lay | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(round(d['stream_us_per_image'],2), 'us/img', round(d['stream_GBps']), 'GB/s fix', round(d['fix_us_per_image'],2), 'other', round(d['other_us_per_image'],2))"
done; done
