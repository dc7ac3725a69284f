#!/bin/bash
# GPU-box job: configs.c5 alone, with and without the integer-moment thumbnail pass
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for v in 1 0 1; do echo -n "IPG_VINT=$v "; IPG_VINT=$v timeout 200 python tools/c5_probe.py 2>&1 | tail -1; done | tee gpurun_out/c5_ab.log
