#!/bin/bash
# GPU-box job: the JPEG writer's section time (CUDA events, tools/jpeg_probe.py) for several builds of the library
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for lib in "$@"; do
  echo -n "$lib: "
  IPG_LIB_PATH=$PWD/imageprocessor_b200/$lib timeout 200 python tools/jpeg_probe.py --images 16 --steps 3 --verify 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())['jpeg']
print('jpeg section', round(d['us_per_image']['other_kernel_ms'],1), 'us/img; wall', round(d['images_per_s_wall']), 'img/s', d.get('verified_image0_all_files_byte_identical'))"
done
