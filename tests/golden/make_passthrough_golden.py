#!/usr/bin/env python
"""Golden outputs of the REFERENCE PassThroughSamCheck / PassThroughSamCheck.stranded (oracle/_ref,
built from /root/reference by oracle/Makefile) on a SAM text made from the committed edge-case FASTQ
files: interleaved mates, every 3rd record flagged reverse (0x10) with its sequence reverse-
complemented as an aligner would store it, two chromosomes, some unpaired reads, IUPAC characters.
Run in the build container; writes tests/golden/passthrough/{input.sam, golden.json}."""
import hashlib, json, os, subprocess, sys, tempfile
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from rufus_b200 import fastq
REFBIN = os.path.join(HERE, "..", "..", "oracle", "_ref")
out = os.path.join(HERE, "passthrough")
os.makedirs(out, exist_ok=True)
m1 = fastq.read_fastq(os.path.join(HERE, "edge", "Subj.mate1.fastq"))
m2 = fastq.read_fastq(os.path.join(HERE, "edge", "Subj.mate2.fastq"))
comp = {65: 84, 67: 71, 71: 67, 84: 65, 78: 78}
lines = []
for i in range(len(m1)):
    for mate, r in ((0, m1), (1, m2)):
        if mate == 1 and i % 11 == 5:
            continue                                   # unpaired: mate 2 missing
        name = f"pair{i:04d}"                          # mates share the name (the pairing key)
        seq = bytes(r.seq[r.offs[i]:r.offs[i + 1]])
        qual = bytes(r.qual[r.offs[i]:r.offs[i + 1]])
        flag = 65 if mate == 0 else 129
        if (2 * i + mate) % 3 == 0:
            flag |= 16
            seq = bytes(comp.get(c, c) for c in reversed(seq))
            qual = qual[::-1]
        chrom = "chr1" if i < len(m1) // 2 else "chr2"
        lines.append("\t".join([name, str(flag), chrom, str(100 + i), "60", f"{len(seq)}M", "=", str(300 + i), "0",
                                seq.decode(), qual.decode(), "NM:i:0"]))
sam = "\n".join(lines) + "\n"
open(os.path.join(out, "input.sam"), "w").write(sam)
g = {"generator": "tests/golden/make_passthrough_golden.py"}
with tempfile.TemporaryDirectory() as d:
    r = subprocess.run([os.path.join(REFBIN, "PassThroughSamCheck"), os.path.join(d, "chr1")], input=sam.encode(),
                       capture_output=True, check=True)
    g["plain_fastq_md5"] = hashlib.md5(r.stdout).hexdigest()
    g["plain_chr"] = open(os.path.join(d, "chr1")).read()
    subprocess.run([os.path.join(REFBIN, "PassThroughSamCheck.stranded"), os.path.join(d, "chr2"), os.path.join(d, "st")],
                   input=sam.encode(), capture_output=True, check=True)
    g["stranded_mate1_md5"] = hashlib.md5(open(os.path.join(d, "st.mate1.fastq"), "rb").read()).hexdigest()
    g["stranded_mate2_md5"] = hashlib.md5(open(os.path.join(d, "st.mate2.fastq"), "rb").read()).hexdigest()
    g["stranded_pairs"] = open(os.path.join(d, "st.mate1.fastq")).read().count("\n") // 4
    g["stranded_chr"] = open(os.path.join(d, "chr2")).read()
json.dump(g, open(os.path.join(out, "golden.json"), "w"), indent=1)
print(g)
