#!/bin/bash
# GPU-box job: bench `value` leg at several band-count targets (CTAs a batch should at least give).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for t in "$@"; do
  echo -n "IPG_BAND_CTAS $t: "
  IPG_BAND_CTAS=$t timeout 200 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); r=d['roofline']; print(round(d['value']), 'img/s', round(d['ms_per_step'],3), 'ms/step  k_stream frac', round(r['frac'],3), 'us/img', round(r.get('stream_us_per_image',0),2), 'launches', d['gpu_launches'])"
done
