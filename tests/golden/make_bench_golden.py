#!/usr/bin/env python
"""Digests of the REFERENCE's output on the very workloads bench.py times.

For world = 1, 2, 4, 8 the bench workload is: synthetic trio (rufus_b200/synth.py, seed 42), genome
5 Mb x world, every rank brings its own slice of 500 k pairs per sample (world = 1: the unsliced stream
= BASELINE configs[1] / SURVEY §8d C2). This script builds the UNION of the rank slices, runs the
reference's own commands on it (oracle/_ref binaries, same recipe as make_golden.py / SURVEY §8c):

  jellyfish count --disk -m 25 -L 2 -C  ->  jellyfish histo -f  ->  jellyfish merge | jellyfish query | awk
  ->  RUFUS.Filter HashList mate1 mate2

and stores md5 digests of the subject histogram file, the sorted HashList and the sorted pulled mate-1
names in tests/golden/bench_digests.json. bench.py recomputes the same digests from the CUDA path's
results after the timed loop and fails (exit code 3) on a mismatch; tests/test_gpu_parity.py does the same.

Runs only in the build container (needs oracle/_ref). Usage: make_bench_golden.py [world ...]
"""
import hashlib, json, os, subprocess, sys, tempfile, shutil, time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REFBIN = os.path.join(ROOT, "oracle", "_ref")
JF = os.path.join(REFBIN, "jellyfish")
OUT = os.path.join(HERE, "bench_digests.json")

from rufus_b200 import synth, workload  # noqa: E402


def md5(b: bytes) -> str:
    return hashlib.md5(b).hexdigest()


def sh(cmd, cwd):
    subprocess.run(["bash", "-c", cmd], check=True, cwd=cwd)


def run(world, seed=workload.SEED, pairs=workload.PAIRS, genome_unit=workload.GENOME_UNIT, k=workload.K):
    t0 = time.time()
    work = tempfile.mkdtemp(prefix="rg_benchgold_")
    try:
        fq = {}
        for s in synth.SAMPLES:
            fq[s] = (os.path.join(work, f"{s}.m1.fq"), os.path.join(work, f"{s}.m2.fq"))
        for rank in range(world):
            trio = workload.make_slice(seed, genome_unit, pairs, rank, world)
            for s in synth.SAMPLES:
                s1, q1, s2, q2 = trio[s]
                pre = workload.name_prefix(s, rank, world)
                with open(fq[s][0], "ab") as f:
                    f.write(synth.fastq_bytes(s1, q1, prefix=pre))
                with open(fq[s][1], "ab") as f:
                    f.write(synth.fastq_bytes(s2, q2, prefix=pre))
            del trio
        size = "256M" if world == 1 else "1G"
        for s in synth.SAMPLES:
            sh(f"cat {fq[s][0]} {fq[s][1]} | {JF} count --disk -m {k} -L 2 -s {size} -t 8 -o {s}.Jhash -C /dev/stdin && "
               f"{JF} histo -f -o {s}.Jhash.histo {s}.Jhash", work)
        sh(f"{JF} merge Child.Jhash Mother.Jhash Father.Jhash > merge.out && "
           f"{JF} query -s <(awk '{{print \">\"$1\"\\n\"$1}}' merge.out) Child.Jhash | "
           f"awk -v var={workload.MIN_COV} ' $2 >= var ' | awk -v var={workload.MAX_COV} ' $2 <= var ' > HashList", work)
        sh(f"{REFBIN}/RUFUS.Filter HashList {fq['Child'][0]} {fq['Child'][1]} out {k} {workload.MINQ} {workload.THRESH} 8 "
           "> /dev/null", work)
        histo = open(os.path.join(work, "Child.Jhash.histo"), "rb").read()
        hl = sorted(open(os.path.join(work, "HashList")).read().splitlines())
        names = sorted(open(os.path.join(work, "out.Mutations.Mate1.fastq")).read().splitlines()[0::4])
        res = {"world": world, "seed": seed, "pairs_per_rank": pairs, "genome": genome_unit * world, "k": k,
               "min_cov": workload.MIN_COV, "max_cov": workload.MAX_COV, "minq": workload.MINQ, "thresh": workload.THRESH,
               "histo_md5": md5(histo), "hashlist_md5": md5(("\n".join(hl) + "\n").encode()), "hashlist_lines": len(hl),
               "pulled_md5": md5(("\n".join(names) + "\n").encode()), "pulled_pairs": len(names),
               "reference_seconds": round(time.time() - t0, 1)}
        return res
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    worlds = [int(a) for a in sys.argv[1:]] or [1]
    db = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for w in worlds:
        r = run(w)
        db[workload.digest_key(w)] = r
        print(w, r, flush=True)
        json.dump(db, open(OUT, "w"), indent=1, sort_keys=True)
