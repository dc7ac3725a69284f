#!/bin/bash
# GPU-box job: ncu --set full capture of k_jpeg_dct (the JPEG writer's dominant kernel), after a plain run of the same command
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python tools/jpeg_probe.py --images 4 --steps 1 --verify 0 > gpurun_out/plain_jpeg_dct.log 2>&1; echo plain rc=$?
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:k_jpeg_dct -s 1 -c 1 -o gpurun_out/prof_jpeg_dct -f \
    python tools/jpeg_probe.py --images 4 --steps 1 --verify 0 > gpurun_out/ncu_jpeg_dct.log 2>&1; echo dct rc=$?
python -c "import __graft_entry__ as g; g.smoke()"
