#!/usr/bin/env python
"""tools/src_profile.py REPORT.ncu-rep [top] : stall samples and executed instructions per CUDA source line
(`ncu --set full --import-source on` capture compiled with -lineinfo; needs ncu, no GPU)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                     text=True, check=True).stdout
fname, hdr, out = "", None, []
for r in csv.reader(io.StringIO(txt)):
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        ist, iex = r.index("Warp Stall Sampling (All Samples)"), r.index("Instructions Executed")
        continue
    if hdr and len(r) > iex and r[2] == "-" and r[0].isdigit():   # a CUDA line (SASS rows carry an address)
        try:
            out.append((int(r[ist] or 0), int(r[iex] or 0), fname, int(r[0]), r[1].strip()[:100]))
        except ValueError:
            pass
tot, tex = sum(o[0] for o in out) or 1, sum(o[1] for o in out) or 1
print(f"stall samples {tot}, warp instructions {tex}")
for st, ex, f, ln, src in sorted(out, reverse=True)[:top]:
    print(f"{100 * st / tot:6.2f}% stall {100 * ex / tex:6.2f}% instr  {f}:{ln}: {src}")
