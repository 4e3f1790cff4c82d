#!/usr/bin/env python
"""Split a kernel's source-page counters into segments delimited by BAR.SYNC: executed
warp instructions, shared-memory wavefronts and stall samples by reason per segment.
usage: tools/ncu_phases.py rep kernel-regex"""
import csv, subprocess, sys
rep, kname = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kname, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
def new():
    return {"inst": 0, "wf": 0, "samples": 0, **{s: 0 for s in stalls}}
seg, segs, first = new(), [], None
for r in rows[2:]:
    if len(r) < len(hdr) or not r[0].startswith("0x"):
        continue
    a = int(r[0], 16); first = a if first is None else first
    seg["inst"] += int(r[col["Instructions Executed"]]); seg["wf"] += int(r[col["L1 Wavefronts Shared"]] or 0)
    seg["samples"] += int(r[col["# Samples"]])
    for s in stalls:
        seg[s] += int(r[col[s]] or 0)
    if "BAR.SYNC" in r[1]:
        segs.append((a - first, seg)); seg = new()
segs.append((-1, seg))
tot = sum(s["samples"] for _, s in segs)
for a, s in segs:
    top = sorted(((s[k], k[6:]) for k in stalls), reverse=True)[:6]
    print(f"..{a:05x}: {s['inst']/1e6:7.1f} M inst  {s['wf']/1e6:6.1f} M wavefronts  {100*s['samples']/max(tot,1):5.1f}% time | " +
          " ".join(f"{k}={v}" for v, k in top if v))
