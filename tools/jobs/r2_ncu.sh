#!/bin/bash
# GPU-box job (1 GPU): quick bench line, then the round's ncu evidence: launch list of the bench command and
# ncu --set full captures of (a) the merged lean kernel, (b) the thumbnail pass alone, (c) a mid-size source, (d) planar 4:2:0.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/bench_quick_$TAG.json 2> gpurun_out/bench_quick_$TAG.err; echo bench rc=$?
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_quick_$TAG.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "frac", round(r["frac"], 3), "thumb alone", round(r["passes_timed_alone"]["thumbnail"]["frac"], 3),
      "passA alone", round(r["passes_timed_alone"]["resize+watermark_copy"]["frac"], 3), "fix ms/step", round(r["fix_kernel_ms_per_step"], 3))
print("e2e", round(d["e2e"]["value"]), d["e2e"]["pcie"]["achieved"], d["e2e"]["pcie"]["peak"], d["e2e"]["pcie"]["frac_of_pcie"], d["verified"], d["e2e"]["verified_slot0"])
PY
BENCH="python bench.py --steps 1 --warmup 3 --images 64 --no-e2e --no-cpu-baseline --no-verify --no-configs"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv $BENCH > gpurun_out/ncu_launches_$TAG.log 2>&1; echo launches rc=$?
KREGEX='k_stream<\(int\)1, \(bool\)1, \(int\)4>' KSKIP=1 bash tools/jobs/ncu_one.sh rtw lean_$TAG | tail -2
IPG_MERGE_LEAN=0 KREGEX='k_stream<\(int\)1, \(bool\)0, \(int\)2>' KSKIP=1 bash tools/jobs/ncu_one.sh rt thumb_$TAG | tail -2
python tools/profile_step.py --images 64 --steps 1 --ops rt --w 1920 --h 1080 --max-batch 64 > gpurun_out/plain_mid_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_stream<\(int\)1, \(bool\)0, \(int\)4>' -s 1 -c 1 -o gpurun_out/prof_mid_$TAG -f \
    python tools/profile_step.py --images 64 --steps 1 --ops rt --w 1920 --h 1080 --max-batch 64 > gpurun_out/ncu_mid_$TAG.log 2>&1; echo mid rc=$?; cat gpurun_out/plain_mid_$TAG.log
for lay in ycbcr420 nrgba; do
python tools/profile_step.py --images 16 --steps 1 --ops rtw --layout $lay > gpurun_out/plain_${lay}_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_stream_planar" -s 1 -c 1 -o gpurun_out/prof_${lay}_$TAG -f \
    python tools/profile_step.py --images 16 --steps 1 --ops rtw --layout $lay > gpurun_out/ncu_${lay}_$TAG.log 2>&1; echo $lay rc=$?; cat gpurun_out/plain_${lay}_$TAG.log
done
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_exact_fix" -s 2 -c 2 -o gpurun_out/prof_fix_$TAG -f \
    python tools/profile_step.py --images 16 --steps 1 --ops rtw > gpurun_out/ncu_fix_$TAG.log 2>&1; echo fix rc=$?
python tools/size_sweep.py --out gpurun_out/size_sweep_$TAG.json > gpurun_out/size_sweep_$TAG.log 2>&1; echo sweep rc=$?
ls -la gpurun_out/*_$TAG.ncu-rep
