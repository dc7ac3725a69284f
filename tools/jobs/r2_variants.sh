#!/bin/bash
# GPU-box job: time several builds of the library (kernel experiments) on the same op mixes; then the default build's
# bench with the batch-tail overlap on and off.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for lib in "$@"; do
  for ops in r rt rtw; do
    echo -n "$lib $ops: "
    IPG_LIB_PATH=$PWD/imageprocessor_b200/$lib timeout 120 python tools/profile_step.py --images 64 --steps 3 --ops $ops --lanes 1 --max-batch 64 | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(round(d['stream_us_per_image'],2), 'us/img', round(d['stream_GBps']), 'GB/s fix', round(d['fix_us_per_image'],2))"
  done
  echo -n "$lib 1920x1080 rt: "
  IPG_LIB_PATH=$PWD/imageprocessor_b200/$lib timeout 120 python tools/profile_step.py --images 64 --steps 3 --ops rt --w 1920 --h 1080 --lanes 1 --max-batch 64 | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(round(d['stream_us_per_image'],2), 'us/img')"
  echo -n "$lib ycbcr420 rt: "
  IPG_LIB_PATH=$PWD/imageprocessor_b200/$lib timeout 120 python tools/profile_step.py --images 32 --steps 3 --ops rt --layout ycbcr420 --lanes 1 | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(round(d['stream_us_per_image'],2), 'us/img fix', round(d['fix_us_per_image'],2))"
done
for ov in 0 1; do
  echo -n "bench overlap_tail=$ov: "
  IPG_OVERLAP_TAIL=$ov timeout 300 python bench.py --steps 10 --warmup 3 --no-configs --no-cpu-baseline --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.readline()); r=d['roofline']; print('value', round(d['value']), 'frac', round(r['frac'],3), 'thumb', round(r['passes_timed_alone']['thumbnail']['frac'],3), 'passA', round(r['passes_timed_alone']['resize+watermark_copy']['frac'],3), d['verified'])"
done
