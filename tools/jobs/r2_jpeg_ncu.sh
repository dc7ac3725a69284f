#!/bin/bash
# GPU-box job: per-launch durations of the JPEG writer's kernels (ncu launch list of the probe, 4 images, one step)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_jpeg --csv --log-file gpurun_out/jpeg_launches.csv \
    python tools/jpeg_probe.py --images 4 --steps 1 --verify 0 > gpurun_out/jpeg_ncu.log 2>&1; echo ncu rc=$?
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/jpeg_launches.csv")) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.defaultdict(list)
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); u = r[ui]
    v = v / 1000 if u in ("ns", "nsecond") else v
    agg[r[ki].split("(")[0]].append(v)
for k, v in agg.items(): print(f"{k:40s} n={len(v):3d} mean {sum(v)/len(v):9.1f} us  (4 images x 3 files per launch)")
PY
if [ "$1" = "full" ]; then
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:k_jpeg_dct -s 1 -c 1 -o gpurun_out/prof_jpeg_dct -f \
    python tools/jpeg_probe.py --images 4 --steps 1 --verify 0 > gpurun_out/ncu_jpeg_dct.log 2>&1; echo dct rc=$?
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:k_jpeg_write -s 1 -c 1 -o gpurun_out/prof_jpeg_write -f \
    python tools/jpeg_probe.py --images 4 --steps 1 --verify 0 > gpurun_out/ncu_jpeg_write.log 2>&1; echo write rc=$?
fi
