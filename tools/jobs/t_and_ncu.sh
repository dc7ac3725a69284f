#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 120 python tools/profile_step.py --images 32 --steps 3 --ops t --lanes 1
timeout 120 python tools/profile_step.py --images 32 --steps 3 --ops w --lanes 1
bash tools/jobs/ncu_one.sh ${1:-rt} ${2:-rt_v9a}
