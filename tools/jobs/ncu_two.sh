#!/bin/bash
# GPU-box job: ncu --set full of (1) the merged lean kernel as the bench runs it, (2) the resize + watermark pass alone.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
TAG=${1:-r1}
KREGEX='k_stream<\(int\)1, \(bool\)1, \(int\)4>' KSKIP=1 bash tools/jobs/ncu_one.sh rtw lean_$TAG | tail -3
IPG_MERGE_LEAN=0 KREGEX='k_stream<\(int\)1, \(bool\)1, \(int\)1>' KSKIP=1 bash tools/jobs/ncu_one.sh rtw passA_$TAG | tail -3
ls -la gpurun_out/*.ncu-rep | tail -3
