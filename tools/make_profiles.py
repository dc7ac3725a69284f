#!/usr/bin/env python
"""Turn the scratch captures in gpurun_out/ into the tracked summaries under profiles/.
usage: python tools/make_profiles.py <round tag, e.g. r1> <launch-list tag> <full-capture tag> <images in the full capture>"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rnd, ltag, ftag, n_img = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)

# ---- launch list: keep this library's kernels, summarise shares
rows = [r for r in csv.reader(open(os.path.join(ROOT, "gpurun_out", f"launches_{ltag}.csv"))) if len(r) > 5]
hdr = rows[0]
iN, iV, iG, iB = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
ours = [r for r in rows[1:] if "ipg::" in r[iN]]
with open(os.path.join(out, f"{rnd}_launches.csv"), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "grid", "block", "gpu__time_duration.sum [ns]"])
    for r in ours:
        w.writerow([r[iN].split("(")[0].replace("void ", ""), r[iG], r[iB], r[iV]])
agg = {}
for r in ours:
    k = r[iN].split("(")[0].replace("void ", "")
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += float(r[iV].replace(",", ""))
tot = sum(v[1] for v in agg.values())
with open(os.path.join(out, f"{rnd}_launch_summary.txt"), "w") as f:
    f.write(f"ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches: compare SHARES)\n")
    f.write(f"command: python bench.py --steps 1 --warmup 3 --images 64 --no-e2e --no-cpu-baseline --no-verify --no-configs (tools/jobs/r2_ncu.sh)\n\n")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write(f"{k:32s} launches {n:4d}  total {t/1e3:10.1f} us  share {100*t/tot:5.1f}%\n")

# ---- full capture of the dominant kernel
rep = os.path.join(ROOT, "gpurun_out", f"prof_{ftag}.ncu-rep")
txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_hot.py"), rep, "1.0"], capture_output=True, text=True).stdout
with open(os.path.join(out, f"{rnd}_k_stream_full.txt"), "w") as f:
    f.write(f"ncu --set full --clock-control none --import-source on -k regex:k_stream -s 1 -c 1\n")
    f.write(f"command: python tools/profile_step.py --images {n_img} --steps 1 --ops rtw  (one launch = {n_img} 12 MP images; kernel filter: the merged lean instantiation k_stream<1,true,4>)\n")
    f.write("summary by tools/ncu_hot.py (headline metrics, stall totals, SASS lines with >= 1% of the samples)\n\n")
    f.write(txt)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
h, u, v = r[0], r[1], r[2]
def metric(name):
    i = h.index(name)
    x = float(v[i].replace(",", ""))
    return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u[i], 1)
rd, wr = metric("dram__bytes_read.sum"), metric("dram__bytes_write.sum")
json.dump({"kernel": v[h.index("Kernel Name")], "images_per_launch": n_img,
           "dram_bytes_read_per_launch": rd, "dram_bytes_write_per_launch": wr,
           "dram_bytes_per_launch": rd + wr, "dram_bytes_per_image": (rd + wr) / n_img,
           "note": f"ncu --set full, profiles/{rnd}_k_stream_full.txt; dram__bytes_read.sum + dram__bytes_write.sum"},
          open(os.path.join(out, "traffic.json"), "w"), indent=1)
# optional: further captures "name=tag" -> profiles/<rnd>_<name>_full.txt
for extra in sys.argv[5:]:
    name, tag = extra.split("=")
    t = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_hot.py"), os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep"), "1.0"],
                       capture_output=True, text=True).stdout
    with open(os.path.join(out, f"{rnd}_{name}_full.txt"), "w") as f:
        f.write("ncu --set full --clock-control none --import-source on (same command as the k_stream capture)\n\n" + t)
print(open(os.path.join(out, f"{rnd}_launch_summary.txt")).read())
print(open(os.path.join(out, "traffic.json")).read())
