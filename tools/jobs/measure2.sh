#!/bin/bash
# GPU-box job: the round's measurement set.  $1 = tag
cd "${GRAFT_REPO_ROOT:-/root/repo}"
TAG=${1:-r1}
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo ref rc=$?
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo bench rc=$?
tail -3 gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json
# launch list of a short run of the same program (cold-cache, serialised: compare shares)
python bench.py --steps 1 --warmup 3 --images 64 --no-e2e --no-cpu-baseline --no-verify > gpurun_out/bench_short_$TAG.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --images 64 --no-e2e --no-cpu-baseline --no-verify > gpurun_out/ncu_launches_$TAG.log 2>&1
echo launches rc=$?
# full captures (ncu --set full) of the dominant kernel and of its main pass alone
bash tools/jobs/ncu_two.sh $TAG
