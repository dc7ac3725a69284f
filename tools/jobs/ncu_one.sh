#!/bin/bash
# GPU-box job: one ncu --set full capture of k_stream for an op mix ($1, default rtw) -> gpurun_out/prof_$2.ncu-rep
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OPS=${1:-rtw}; TAG=${2:-x}
python tools/profile_step.py --images 16 --steps 1 --ops $OPS > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:${KREGEX:-k_stream}" -s ${KSKIP:-1} -c 1 -o gpurun_out/prof_$TAG -f \
    python tools/profile_step.py --images 16 --steps 1 --ops $OPS > gpurun_out/ncu_$TAG.log 2>&1
echo rc=$?; cat gpurun_out/plain_$TAG.log; tail -3 gpurun_out/ncu_$TAG.log
