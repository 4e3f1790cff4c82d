#!/usr/bin/env python3
"""Static SASS digest of rufus_b200/librufus_gpu.so (no GPU needed): per kernel the number of SASS instructions and the
counts of the opcodes that tell what the kernel is made of — UBLKCP (cp.async.bulk, the TMA bulk-copy engine), SYNCS
(mbarrier), shared/global atomics, 128-bit loads/stores, BREV (reverse complement), tensor-core opcodes (none expected:
integer path) — plus registers / static shared memory from cuobjdump -res-usage.
    python tools/sass_digest.py > profiles/<round>_sass_digest.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "rufus_b200", "librufus_gpu.so")
WATCH = ("UBLKCP", "UBLKRED", "UTMALDG", "UTMASTG", "SYNCS", "ATOMS", "ATOMG", "ATOM", "RED", "REDG", "LDG.E.128", "STG.E.128",
         "LDS.128", "STS.128", "LDGSTS", "BREV", "MATCH", "VOTE", "SHFL", "BAR", "IMAD", "LOP3", "SHF", "POPC", "FLO")
TENSOR = re.compile(r"\b(HMMA|IMMA|DMMA|QMMA|OMMA|UTCHMMA|UTCIMMA|UTCQMMA|UTCOMMA|UTCMMA|WGMMA|HGMMA|IGMMA)")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", SO], capture_output=True, text=True).stdout
    usage = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*(.*)", res):
        usage[m.group(1)] = m.group(2).strip()
    demangle = {}
    names = sorted(usage)
    if names:
        out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
        demangle = dict(zip(names, out)) if len(out) == len(names) else {}
    archs = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    kernels[cur][w if w not in ("ATOMS", "ATOMG", "RED", "REDG", "ATOM") else ".".join(op.split(".")[:2])] += 1
            if TENSOR.search(op):
                kernels[cur]["TENSOR"] += 1
    print(f"# SASS digest of rufus_b200/librufus_gpu.so (tools/sass_digest.py: cuobjdump -sass / -res-usage; static counts per kernel)")
    print("# UBLKCP = cp.async.bulk (TMA bulk copy engine), SYNCS = mbarrier; TENSOR = any tensor-core opcode (none: integer path)")
    print(f"# cubin architectures: {archs}\n")
    tensor = 0
    for k, c in sorted(kernels.items(), key=lambda kv: -kv[1]["total"]):
        tensor += c["TENSOR"]
        name = demangle.get(k, k)
        print(name[:150])
        print("   " + usage.get(k, ""))
        print(f"   instructions {c['total']}  " + "  ".join(f"{w}:{n}" for w, n in sorted(c.items()) if w != "total"))
    print(f"\ntensor-core opcodes in the whole library: {tensor}")


if __name__ == "__main__":
    sys.exit(main())
