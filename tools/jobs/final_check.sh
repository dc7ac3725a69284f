#!/bin/bash
# GPU-box job: parity suite, smoke, the bench line, and device times of the 16-bit-sample kernels.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo bench rc=$?
for lay in ycbcr420 nrgba; do
  echo -n "$lay rt: "; timeout 100 python tools/profile_step.py --images 32 --steps 3 --ops rt --layout $lay | cut -c1-200
done
