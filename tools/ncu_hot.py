#!/usr/bin/env python
"""Summarise an .ncu-rep here (no GPU): headline metrics + hottest SASS lines with stall reasons.
usage: python tools/ncu_hot.py gpurun_out/prof_x.ncu-rep [min_sample_pct]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.8
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps', 'launch__waves_per_multiprocessor',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.max',
        'smsp__cycles_active.avg', 'local_load', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum']
for r in rows[2:]:
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:70s} {r[i]} {units[i]}")
    print('---')
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None
data = []
nk = 0
for r in rows:
    if r and r[0] == 'Kernel Name':
        nk += 1
        if nk > 1:
            break
        continue
    if r and r[0] == 'Address':
        hdr = r
        continue
    if hdr:
        data.append(r)
iA, iS, iSrc = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Source')
tot = sum(int(r[iA]) for r in data)
ts = sum(int(r[iS]) for r in data)
print('total warp-inst', tot, 'samples', ts, 'sass lines', len(data))
stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
agg = {}
for r in data:
    for i in stall:
        agg[hdr[i][6:]] = agg.get(hdr[i][6:], 0) + int(r[i])
print('stall totals:', sorted(((v, k) for k, v in agg.items()), reverse=True)[:8])
for k, r in enumerate(data):
    a, s = int(r[iA]), int(r[iS])
    if s / ts * 100 >= thr:
        st = sorted(((int(r[i]), hdr[i][6:]) for i in stall), reverse=True)[:2]
        print(f"{k:5d} {r[iSrc].strip()[:70]:70s} exec {100*a/tot:5.2f}% samp {100*s/ts:5.2f}% {st}")
