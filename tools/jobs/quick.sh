#!/bin/bash
# GPU-box job: parity tests, then per-kernel device times for the three op mixes.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for ops in r rt rtw; do
  timeout 120 python tools/profile_step.py --images 32 --steps 3 --ops $ops --lanes 1
done
