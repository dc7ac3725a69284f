#!/bin/bash
# GPU-box job: parity suite, smoke, the bench line, and the 8K device times (resize+watermark alone, all three ops).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo bench rc=$?
for ops in rw rtw; do
  echo -n "8K $ops: "; timeout 100 python tools/profile_step.py --images 16 --steps 2 --ops $ops --w 7680 --h 4320 --max-batch 16
done
