#!/bin/bash
# GPU-box job: device JPEG writer -- parity tests, then the timing probe (photo-like and noise), then (arg 1 = full) the whole gpu suite.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_jpeg_gpu.py -m gpu -x -q 2>&1 | tail -15
timeout 300 python tools/jpeg_probe.py --images 16 --steps 2 > gpurun_out/jpeg_probe_photo.json 2> gpurun_out/jpeg_probe_photo.err; echo probe rc=$?; tail -c 1500 gpurun_out/jpeg_probe_photo.json; tail -3 gpurun_out/jpeg_probe_photo.err
timeout 300 python tools/jpeg_probe.py --images 16 --steps 2 --noise 1 --verify 0 > gpurun_out/jpeg_probe_noise.json 2> gpurun_out/jpeg_probe_noise.err; echo probe rc=$?; tail -c 1500 gpurun_out/jpeg_probe_noise.json; tail -3 gpurun_out/jpeg_probe_noise.err
if [ "$1" = "full" ]; then timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5; fi
