#!/usr/bin/env python
"""tools/sass_profile.py REPORT.ncu-rep [top] : instruction mix and stall samples of the first kernel of an
`ncu --set full --import-source on` capture, by opcode, plus the SASS lines with the most stall samples."""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
want = sys.argv[3] if len(sys.argv) > 3 else ""   # substring of the kernel name (first matching launch)
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                     text=True, check=True).stdout
allrows = list(csv.reader(io.StringIO(txt)))
# the page lists the launches one after the other, each introduced by a "Kernel Name" row
starts = [i for i, r in enumerate(allrows) if r and r[0] == "Kernel Name"]
pick = next((i for i in starts if want in allrows[i][1]), starts[0])
end = next((i for i in starts if i > pick), len(allrows))
rows = allrows[pick:end]
print("kernel:", rows[0][1][:100])
hdr = rows[1]
ia, isrc, ist, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
ops = collections.Counter()
stl = collections.Counter()
wsh = collections.Counter()
wgl = collections.Counter()
iws = hdr.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in hdr else None
iwg = hdr.index("L1 Tag Requests Global") if "L1 Tag Requests Global" in hdr else None
lines = []
for r in rows[2:]:
    if len(r) <= iex or not r[iex].isdigit():
        continue
    src = r[isrc].strip()
    tok = src.split()
    op = tok[1] if tok and tok[0].startswith("@") else (tok[0] if tok else "?")
    op = op.split(".")[0] + ("." + ".".join(op.split(".")[1:2]) if op.startswith(("LD", "ST")) else "")
    ex, st = int(r[iex]), int(r[ist] or 0)
    ops[op] += ex
    stl[op] += st
    if iws is not None and r[iws].isdigit():
        wsh[op] += int(r[iws])
    if iwg is not None and r[iwg].isdigit():
        wgl[op] += int(r[iwg])
    lines.append((st, ex, src))
tot, tst = sum(ops.values()), sum(stl.values())
print(f"instructions executed (warp level): {tot}   stall samples: {tst}")
for op, n in ops.most_common(top):
    print(f"  {op:14s} {n:14d} {100.0 * n / tot:6.2f} %   stalls {100.0 * stl[op] / max(tst, 1):6.2f} %   smem wavefronts {wsh[op]:12d}   global tag requests {wgl[op]:12d}")
print("lines with the most stall samples:")
for st, ex, src in sorted(lines, reverse=True)[:top]:
    print(f"  {100.0 * st / max(tst, 1):6.2f} %  {src[:110]}")
