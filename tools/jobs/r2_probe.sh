#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python tools/stream_probe.py 2>&1 | tail -60
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
