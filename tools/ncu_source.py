#!/usr/bin/env python
"""Per-SASS-instruction execution counts of one kernel from an .ncu-rep (source page).
usage: tools/ncu_source.py rep kernel-substring [min_count]"""
import csv, subprocess, sys
rep, kname = sys.argv[1], sys.argv[2]
minc = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kname, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, isrc, ie, it, iss = (hdr.index(x) for x in ("Address", "Source", "Instructions Executed", "Avg. Threads Executed", "# Samples"))
iw = hdr.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in hdr else None
base = None
tot = 0
for r in rows[2:]:
    if len(r) <= ie or not r[ia].startswith("0x"):
        break
    a = int(r[ia], 16)
    base = a if base is None else base
    e = int(r[ie]); tot += e
    if e >= minc:
        wf = r[iw] if iw is not None else ""
        print(f"{a - base:05x} {e:>10d} {float(r[it]):5.1f} {r[iss]:>6s} {wf:>9s}  {r[isrc].strip()[:80]}")
print("total warp instructions", tot)
