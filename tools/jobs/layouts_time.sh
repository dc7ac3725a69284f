#!/bin/bash
# GPU-box job: parity suite, then device times of the 16-bit-sample streaming kernels.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for lay in ycbcr420 ycbcr444 nrgba; do
  for ops in r rt; do echo -n "$lay $ops: "; timeout 100 python tools/profile_step.py --images 32 --steps 2 --ops $ops --layout $lay | cut -c1-180; done
done
