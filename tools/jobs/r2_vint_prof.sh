#!/bin/bash
# GPU-box job: issue rates of the integer ops (tools/micro/int_rate) and one ncu --set full capture of the merged lean launch
# with the integer-moment thumbnail pass (IPG_VINT=1) and of the thumbnail pass alone.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
./tools/micro/int_rate | tee gpurun_out/int_rate.json
IPG_VINT=1 KREGEX='k_stream<\(int\)1, \(bool\)1, \(int\)4>' KSKIP=1 bash tools/jobs/ncu_one.sh rtw lean_vint | tail -2
IPG_VINT=1 KREGEX='k_stream<\(int\)1, \(bool\)0, \(int\)4>' KSKIP=1 bash tools/jobs/ncu_one.sh t thumb_vint | tail -2
ls -la gpurun_out/*vint*.ncu-rep
