#!/bin/bash
# GPU-box job: parity tests, then device times of the NRGBA streaming path.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for ops in r rt; do
  echo -n "nrgba $ops: "; timeout 100 python tools/profile_step.py --images 16 --steps 2 --ops $ops --layout nrgba
done
