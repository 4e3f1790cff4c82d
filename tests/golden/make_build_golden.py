#!/usr/bin/env python
"""Golden outputs of the REFERENCE RUFUS.Build (oracle/_ref/RUFUS.Build, built from /root/reference
by oracle/Makefile) on the committed sorted dumps of the testRun and edge cases: subject dump as
`KMER<TAB>COUNT`, control dumps as `KMER COUNT` with -d ' ' (how scripts/RunRUFUS.1000G.sh:66 calls
it). Run in the build container; writes tests/golden/build_golden.json."""
import gzip, json, os, subprocess, tempfile
HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "..", "..", "oracle", "_ref", "RUFUS.Build")
CASES = {"testRun": ("Child", "Mother", "Father"), "edge": ("Subj", "CtlA", "CtlB")}
PARAMS = [(5, 140, 0), (2, 1000, 3), (3, 40, 1)]
out = {"generator": "tests/golden/make_build_golden.py", "cases": []}
for case, names in CASES.items():
    g = json.load(open(os.path.join(HERE, case, "golden.json")))
    with tempfile.TemporaryDirectory() as d:
        for i, s in enumerate(names):
            txt = gzip.open(os.path.join(HERE, case, s + ".dump.sorted.gz")).read().decode()
            open(os.path.join(d, s + ".tab"), "w").write(txt.replace(" ", "\t") if i == 0 else txt)
        for mS, mx, mC in PARAMS:
            cmd = [REF] + sum((["-c", os.path.join(d, s + ".tab")] for s in names[1:]), []) + \
                  ["-s", os.path.join(d, names[0] + ".tab"), "-o", os.path.join(d, "out"), "-hs", str(g["k"]),
                   "-mS", str(mS), "-max", str(mx), "-mC", str(mC), "-t", "1", "-d", " "]
            subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
            lines = open(os.path.join(d, "out")).read().splitlines()
            out["cases"].append({"case": case, "k": g["k"], "mS": mS, "max": mx, "mC": mC, "lines": lines})
            print(case, mS, mx, mC, len(lines))
json.dump(out, open(os.path.join(HERE, "build_golden.json"), "w"))
