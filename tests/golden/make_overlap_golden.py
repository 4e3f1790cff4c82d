#!/usr/bin/env python
"""Golden fixtures for the NEXT row of the scope table (the greedy overlap assembly):
the reference Overlap binary (oracle/_ref/Overlap, built from /root/reference/src/Overlap.cpp by
oracle/Makefile) run three times with the arguments of scripts/Overlap.shorter.sh:141,153,165 on the
read pairs RUFUS.Filter pulls from the bundled Child sample, written as a depth-1 / strand-'+'
FASTQD (the form OverlapSam would hand over; bwa is not available here). Threads = 1: with more
threads the reference's tie-breaking is not deterministic. Writes tests/golden/overlap/."""
import gzip, json, os, shutil, subprocess, tempfile
HERE = os.path.dirname(os.path.abspath(__file__))
REFBIN = os.path.join(HERE, "..", "..", "oracle", "_ref")
out = os.path.join(HERE, "overlap")
os.makedirs(out, exist_ok=True)
g = json.load(open(os.path.join(HERE, "testRun", "golden.json")))
with tempfile.TemporaryDirectory() as d:
    for m in (1, 2):
        open(os.path.join(d, f"m{m}.fq"), "wb").write(gzip.open(os.path.join(HERE, "testRun", f"Child.mate{m}.fastq.gz")).read())
    open(os.path.join(d, "HashList"), "w").write("\n".join(g["hashlist_sorted"]) + "\n")
    subprocess.run([os.path.join(REFBIN, "RUFUS.Filter"), "HashList", "m1.fq", "m2.fq", "pull", "25", "15", "1", "1"],
                   cwd=d, check=True, stdout=subprocess.DEVNULL)
    recs = []
    for m in (1, 2):
        lines = open(os.path.join(d, f"pull.Mutations.Mate{m}.fastq")).read().splitlines()
        recs += [lines[i:i + 4] for i in range(0, len(lines), 4)]
    recs.sort(key=lambda r: r[0])                      # RUFUS.Filter's output order is not deterministic
    with open(os.path.join(d, "in.sam.fastqd"), "w") as f:
        for n, s, p, q in recs:
            f.write(f"{n}\n{s}\n{p}\n{q}\n+\n{' '.join(['1'] * len(s))}\n")
    rounds = [("in.sam.fastqd", ".98", "100", "1", "FP", "20", "1", "r.1", "0", "1"),
              ("r.1.fastqd", ".98", "75", "2", "FP", "20", "1", "r.2", "1", "1"),
              ("r.2.fastqd", ".98", "50", "2", "Child", "20", "1", "r.3", "1", "1")]
    counts = [len(recs)]
    for a in rounds:
        subprocess.run([os.path.join(REFBIN, "Overlap"), *a], cwd=d, check=True, stdout=subprocess.DEVNULL)
        counts.append(sum(1 for _ in open(os.path.join(d, a[7] + ".fastqd"))) // 6)
    for f in ("in.sam.fastqd", "r.1.fastqd", "r.1.fastq", "r.2.fastqd", "r.2.fastq", "r.3.fastqd", "r.3.fastq"):
        shutil.copy(os.path.join(d, f), os.path.join(out, f))
    json.dump({"generator": "tests/golden/make_overlap_golden.py", "rounds": [list(a) for a in rounds],
               "records_per_round": counts}, open(os.path.join(out, "golden.json"), "w"), indent=1)
    print("records per round:", counts)
