#!/usr/bin/env python
"""Golden .Jhash header values (matrix, size, reprobes) from the REFERENCE binary oracle/_ref/jellyfish
(built from /root/reference by oracle/Makefile). Run in the build container; writes jf_headers.json."""
import json, os, subprocess, tempfile
HERE = os.path.dirname(os.path.abspath(__file__))
JF = os.path.join(HERE, "..", "..", "oracle", "_ref", "jellyfish")
cases = []
with tempfile.TemporaryDirectory() as d:
    fa = os.path.join(d, "x.fa")
    open(fa, "w").write(">x\nACGTTGCAAGGCTTAACCGGTTAAGCTAGCTAGGATCGATCGGATTTACAGCAT\n")
    for size_arg, k, canon in (("16M", 25, True), ("256M", 25, True), ("1G", 25, False), ("64M", 31, True),
                               ("20k", 25, True), ("100", 5, False), ("1M", 16, True), ("3000000", 21, True)):
        out = os.path.join(d, "o.jf")
        subprocess.run([JF, "count", "-m", str(k), "-s", size_arg, "-t", "1", "-o", out] + (["-C"] if canon else []) + [fa],
                       check=True)
        raw = open(out, "rb").read()
        hlen = int(raw[:9])
        h = json.loads(raw[9:9 + hlen].rstrip(b"\0"))
        mult = {"k": 1000, "M": 1000 ** 2, "G": 1000 ** 3}
        n = int(size_arg[:-1]) * mult[size_arg[-1]] if size_arg[-1] in mult else int(size_arg)
        cases.append({"size_arg": n, "size_text": size_arg, "k": k, "r": h["matrix1"]["r"], "size": h["size"],
                      "columns": h["matrix1"]["columns"], "max_reprobe": h["max_reprobe"],
                      "n_reprobes": len(h["reprobes"]), "key_len": h["key_len"], "val_len": h["val_len"],
                      "counter_len": h["counter_len"]})
json.dump({"generator": "tests/golden/make_jf_headers.py", "cases": cases}, open(os.path.join(HERE, "jf_headers.json"), "w"))
print(len(cases), "cases")
