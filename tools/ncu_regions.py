#!/usr/bin/env python
"""Per-SASS execution counts from an ncu report's source page, grouped into regions of
equal execution count (a cheap way to see where the issued instructions go)."""
import csv, subprocess, sys
def main(path, detail=None):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]; data = rows[2:]
    iS, iI, iW = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)')
    segs, cur = [], None
    for k, r in enumerate(data):
        c, s = int(r[iI]), int(r[iW])
        if cur and abs(c - cur['c']) <= 0.02 * max(c, cur['c'], 1):
            cur['n'] += 1; cur['sum'] += c; cur['s'] += s; cur['end'] = k
        else:
            if cur: segs.append(cur)
            cur = {'start': k, 'end': k, 'c': c, 'n': 1, 'sum': c, 's': s}
    segs.append(cur)
    tot = sum(x['sum'] for x in segs); tots = sum(x['s'] for x in segs)
    print('total warp instr', tot, 'stall samples', tots, 'sass lines', len(data))
    for x in segs:
        if x['sum'] > 0.004 * tot or x['s'] > 0.01 * tots:
            print(f"[{x['start']:4d}-{x['end']:4d}] n={x['n']:3d} exec/instr={x['c']:>9d} total={x['sum']:>10d} ({100*x['sum']/tot:4.1f}%) samples={x['s']} ({100*x['s']/max(tots,1):4.1f}%)")
    if detail:
        a, b = map(int, detail.split('-'))
        for k in range(a, b + 1):
            r = data[k]
            print(f"{k:4d} {int(r[iI]):>9d} {int(r[iW]):>6d}  {r[iS].strip()[:100]}")
if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
