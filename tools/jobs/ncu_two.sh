#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
TAG=${1:-r1}
KREGEX='k_stream<\(int\)1, \(bool\)1, \(bool\)1>' KSKIP=1 bash tools/jobs/ncu_one.sh rtw lean_$TAG | tail -3
KREGEX='k_stream<\(int\)1, \(bool\)0, \(bool\)0>' KSKIP=1 bash tools/jobs/ncu_one.sh rtw thumb_$TAG | tail -3
ls -la gpurun_out/*.ncu-rep | tail -3
