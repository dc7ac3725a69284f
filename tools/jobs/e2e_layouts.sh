#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for lay in rgba ycbcr420; do for ops in rt rtw; do
  timeout 300 python tools/e2e_probe.py --layout $lay --ops $ops --images 192 --steps 3
done; done
