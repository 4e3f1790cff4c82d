#!/usr/bin/env python
"""Golden output of the REFERENCE AnnotateOverlap (oracle/_ref, built from /root/reference by
oracle/Makefile): the first 300 reads of the bundled Child mate-1 FASTQ annotated with the committed
HashList of the testRun case (reads stand in for contigs: same tool, same format). Run in the build
container; writes tests/golden/annotate/{reads.fastq, HashList, golden.json}."""
import gzip, hashlib, json, os, subprocess
HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "..", "..", "oracle", "_ref", "AnnotateOverlap")
out = os.path.join(HERE, "annotate")
os.makedirs(out, exist_ok=True)
g = json.load(open(os.path.join(HERE, "testRun", "golden.json")))
lines = gzip.open(os.path.join(HERE, "testRun", "Child.mate1.fastq.gz")).read().decode().splitlines()
names = set(g["pulled_mate1_names_sorted"])
recs = [lines[i:i + 4] for i in range(0, len(lines), 4)]
pick = [r for r in recs if r[0] in names] + recs[:250]        # the pulled reads (they carry mutant k-mers) + others
open(os.path.join(out, "reads.fastq"), "w").write("".join("\n".join(r) + "\n" for r in pick))
open(os.path.join(out, "HashList"), "w").write("\n".join(g["hashlist_sorted"]) + "\n")
r = subprocess.run([REF, "HashList", "reads.fastq", "hash.out"], cwd=out, capture_output=True, check=True)
res = {"generator": "tests/golden/make_annotate_golden.py", "n_reads": len(pick),
       "stdout_md5": hashlib.md5(r.stdout).hexdigest(),
       "hashout_md5": hashlib.md5(open(os.path.join(out, "hash.out"), "rb").read()).hexdigest(),
       "depth_lines": r.stdout.decode().splitlines()[3::4][:len(names)]}
os.remove(os.path.join(out, "hash.out"))
json.dump(res, open(os.path.join(out, "golden.json"), "w"), indent=1)
print(res["n_reads"], res["stdout_md5"], res["hashout_md5"], sum(1 for l in res["depth_lines"] if set(l) != {"!"}))
