#!/usr/bin/env python
"""Time tools/bin/Overlap against the reference binary (oracle/_ref/Overlap) on a synthetic pull:
reads of a random region on both strands with substitutions and N (tests/cases.overlap_fuzz_input).
usage: tools/overlap_bench.py [n_reads] [genome_len]   (prints wall seconds and whether the outputs agree)"""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cases import overlap_fuzz_input
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
G = int(sys.argv[2]) if len(sys.argv) > 2 else 30000
text = overlap_fuzz_input(7, n_reads=n, genome_len=G)
args = [".98", "100", "1", "FP", "20", "1", "o", "0"]
with tempfile.TemporaryDirectory() as d:
    out = {}
    for name, exe, thr in (("ours", os.path.join(ROOT, "tools", "bin", "Overlap"), "1"),
                           ("reference T=1", os.path.join(ROOT, "oracle", "_ref", "Overlap"), "1"),
                           ("reference T=16", os.path.join(ROOT, "oracle", "_ref", "Overlap"), "16")):
        if not os.path.exists(exe):
            continue
        w = os.path.join(d, name.replace(" ", "_"))
        os.makedirs(w)
        open(os.path.join(w, "in.fastqd"), "w").write(text)
        t0 = time.time()
        subprocess.run([exe, "in.fastqd", *args, thr], cwd=w, check=True, stdout=subprocess.DEVNULL)
        dt = time.time() - t0
        out[name] = open(os.path.join(w, "o.fastqd")).read()
        print(f"{name:16s} {dt:8.2f} s   contigs {out[name].count(chr(10)) // 6}")
    if "ours" in out and "reference T=1" in out:
        print("ours == reference (T=1):", out["ours"] == out["reference T=1"])
