#!/bin/bash
# GPU-box job (1 GPU): A/B of two library builds on the merged launch (12 MP r+t+w and r+t), then the whole gpu suite, the
# default bench line and the ncu launch list with the faster one (IPG_LIB_PATH), so that one call decides and validates.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r2f}; NEW=${2:-libipgpu.so}; OLD=${3:-libipgpu_prev.so}
t() { IPG_LIB_PATH=$PWD/imageprocessor_b200/$1 timeout 120 python tools/profile_step.py --images 64 --steps 4 --ops $2 --lanes 1 --max-batch 64 | python -c "import json,sys; print(json.loads(sys.stdin.readline())['stream_us_per_image'])"; }
n1=$(t $NEW rtw); o1=$(t $OLD rtw); n2=$(t $NEW rt); o2=$(t $OLD rt)
echo "new $NEW rtw $n1 rt $n2 | old $OLD rtw $o1 rt $o2" | tee gpurun_out/pick_$TAG.log
WIN=$(python -c "print('$NEW' if ($n1 + $n2) <= ($o1 + $o2) * 1.002 else '$OLD')")
echo "winner $WIN" | tee -a gpurun_out/pick_$TAG.log
export IPG_LIB_PATH=$PWD/imageprocessor_b200/$WIN
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
bash tools/jobs/r2_final_bench.sh $TAG
