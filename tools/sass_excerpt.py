#!/usr/bin/env python
"""profiles/<round>_k_stream.sass: mnemonic counts of the default merged lean kernel in the built library plus two excerpts --
one source row of the fp32 row loop and one group of the integer-moment loop.  No GPU needed (cuobjdump).
usage: python tools/sass_excerpt.py r2"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rnd = sys.argv[1] if len(sys.argv) > 1 else "r2"
lib = os.path.join(ROOT, "imageprocessor_b200", "libipgpu.so")
fun = "_ZN3ipg8k_streamILi1ELb1ELi4EEEvPKNS_9StreamJobEPKNS_10StreamItemENS_7FixListE"
txt = subprocess.run(["cuobjdump", "-sass", "-fun", fun, lib], capture_output=True, text=True).stdout
ins = [m.group(1).strip() for m in re.finditer(r"/\*[0-9a-f]{4,5}\*/\s+(.*?);", txt)]
cnt = collections.Counter(re.sub(r"^@!?U?P\d\s+", "", i).split()[0].split(".")[0] for i in ins)
idp = [k for k, i in enumerate(ins) if "IDP.2A" in i]
ffma = [k for k, i in enumerate(ins) if i.startswith("FFMA2") or " FFMA2 " in i]
# the fp32 row loop: the longest run of the listing after the integer loop where PRMT / FADD2 / FFMA2 repeat; take the first
# LDS.128 after the integer loop that is followed by 12 PRMT within 60 instructions
row0 = None
for k in range(idp[-1] + 1, len(ins) - 80):
    if ins[k].startswith("LDS.128") and sum(1 for j in ins[k:k + 60] if "PRMT" in j) >= 12 and sum(1 for j in ins[k:k + 70] if "FFMA2" in j) >= 12:
        row0 = k
        break
out = os.path.join(ROOT, "profiles", f"{rnd}_k_stream.sass")
with open(out, "w") as f:
    f.write("cuobjdump -sass -fun 'ipg::k_stream<1, true, 4>' imageprocessor_b200/libipgpu.so   (the default merged lean launch; sm_100a)\n")
    f.write(f"{len(ins)} SASS instructions; mnemonic counts (top 26):\n  " + ", ".join(f"{k} x{v}" for k, v in cnt.most_common(26)) + "\n")
    f.write("evidence: FFMA2/FADD2 = packed fp32 pairs; UBLKCP = cp.async.bulk (1-D TMA, .S.G = global->shared loads, .G.S = shared->global watermark stores);\n"
            "SYNCS = mbarrier (ARRIVE.TRANS64 = expect_tx, PHASECHK.TRANS64.TRYWAIT = try_wait.parity); PRMT = byte -> 2^23+b;\n"
            "IDP.2A = the integer-moment thumbnail pass (one per channel and source row: a [m, 0] / [0, m] 16-bit pair times two bytes of the pixel word); no LDL/STL (no spills).\n")
    f.write(f"the three row loops: integer-moment loop at instructions {idp[0]}..{idp[-1]}, fp32 loops (lane-per-output form, then table forms) after it\n")
    if row0 is not None:
        f.write("\n---- one source row of the fp32 V loop (LDS.128 of 4 pixels, 12 PRMT + 6 FADD2 unpack, 12 FFMA2 into the two accumulator sets, emit test) ----\n")
        f.write("\n".join("    " + i for i in ins[row0:row0 + 48]) + "\n")
    f.write("\n---- one 4-row group of the integer-moment loop (mbarrier wait, 4 LDS.128, 48 IDP.2A, opacity AND, flush test) ----\n")
    f.write("\n".join("    " + i for i in ins[max(idp[0] - 28, 0):idp[47] + 3]) + "\n")
print(open(out).read()[:1500])
