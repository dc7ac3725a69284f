#!/bin/bash
# GPU-box job: ncu --set full of k_stream_planar for a planar 4:2:0 source and for an NRGBA source (resize pass),
# after a plain run of the same command.  -> gpurun_out/prof_planar420.ncu-rep, prof_nrgba.ncu-rep
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for lay in ycbcr420 nrgba; do
  tag=${lay/ycbcr/planar}
  true && \
  timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_stream_planar" -s 1 -c 1 -o gpurun_out/prof_$tag -f \
      python tools/profile_step.py --images 16 --steps 1 --ops rt --layout $lay > gpurun_out/ncu_$tag.log 2>&1
  echo "$lay rc=$?"
done
