#!/bin/bash
# GPU-box job: A/B of library builds over a few source sizes (profile_step, device-resident, one lane): 12 MP r+t+w / r+t,
# 8K r+t+w, and the mid sizes whose horizontal pass runs in the table forms (6 MP, 1080p, 1 MP r+t).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
show='import json,sys; d=json.loads(sys.stdin.readline()); print(round(d["stream_us_per_image"],2), "us/img", "span", round(d["span_us_per_image"],2), "fix", round(d["fix_us_per_image"],2))'
IFS=';' read -ra CFGS <<< "${CFGS:-4000 3000 rtw 64;4000 3000 rt 64;7680 4320 rtw 24;2832 2124 rt 64;1920 1080 rt 64;1152 864 rt 64}"
for lib in "$@"; do
  for cfg in "${CFGS[@]}"; do
    set -- $cfg
    echo -n "$lib $1x$2 $3: "
    IPG_LIB_PATH=$PWD/imageprocessor_b200/$lib timeout 120 python tools/profile_step.py --images $4 --steps 4 --ops $3 --lanes 1 --max-batch $4 --w $1 --h $2 | python -c "$show"
  done
done 2>&1 | tee gpurun_out/ab2.log
