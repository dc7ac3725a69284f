#!/usr/bin/env python
"""Dump the golden cases' inputs for main.go, or (--compare) check main.go's outputs against
tests/golden/golden_v1.npz and the oracle.  See README.md."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.golden.cases import RESAMPLE_CASES, BLEND_CASES, make_source, make_blend_case  # noqa: E402


def main():
    out = sys.argv[1]
    compare = "--compare" in sys.argv
    os.makedirs(out, exist_ok=True)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))
    index, bad = [], 0
    for name, kind, w, h, seed, ops in RESAMPLE_CASES:
        planes = make_source(kind, w, h, seed)
        files = []
        for i, p in enumerate(planes):
            files.append(f"{name}.p{i}.raw")
            p.tofile(os.path.join(out, files[-1]))
        index.append({"Name": name, "Kind": kind, "W": w, "H": h, "Planes": files, "Ops": [list(o) for o in ops]})
        if compare:
            for op in ops:
                tag = f"resize_{op[1]}x{op[2]}" if op[0] == "resize" else f"thumb_{op[1]}"
                want = gold[f"{name}/{tag}"]
                got = np.fromfile(os.path.join(out, f"{name}.{tag}.out"), np.uint8).reshape(want.shape)
                n = int((got != want).sum())
                bad += n
                print(f"{name}/{tag}: {n} bytes differ from the fixture")
    for name, w, h, seed, color, n in BLEND_CASES:
        dst, glyphs = make_blend_case(w, h, seed, n)
        dst.tofile(os.path.join(out, f"{name}.p0.raw"))
        gl = []
        for k, (x0, y0, x1, y1, m, mpx, mpy) in enumerate(glyphs):
            m.tofile(os.path.join(out, f"{name}.m{k}.raw"))
            gl.append({"X0": x0, "Y0": y0, "X1": x1, "Y1": y1, "MpX": mpx, "MpY": mpy, "MaskW": m.shape[1], "MaskH": m.shape[0],
                       "Mask": f"{name}.m{k}.raw"})
        index.append({"Name": name, "Kind": "rgba", "W": w, "H": h, "Planes": [f"{name}.p0.raw"], "Ops": [["blend"]],
                      "Color": list(color), "Glyphs": gl})
        if compare:
            want = gold[f"{name}/blend"]
            got = np.fromfile(os.path.join(out, f"{name}.blend.out"), np.uint8).reshape(want.shape)
            d = int((got != want).sum())
            bad += d
            print(f"{name}/blend: {d} bytes differ from the fixture")
    json.dump(index, open(os.path.join(out, "index.json"), "w"))
    if compare:
        print("PARITY PINNED: the real reference reproduces every fixture" if bad == 0 else f"{bad} bytes differ: fix oracle/ip_oracle.c")
        sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
