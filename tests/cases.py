"""Shared loaders for the golden cases (inputs + reference outputs)."""
import json
import os

import numpy as np

from conftest import GOLDEN
from rufus_b200 import fastq, synth

BUNDLED = ("Child", "Mother", "Father")
EDGE = ("Subj", "CtlA", "CtlB")


def golden(case):
    return json.load(open(os.path.join(GOLDEN, case, "golden.json")))


def load_case(case):
    """-> (golden dict, {sample: (mate1 Reads, mate2 Reads)}), first sample = subject."""
    g = golden(case)
    reads = {}
    if case == "testRun":
        for s in BUNDLED:
            reads[s] = tuple(fastq.read_fastq(os.path.join(GOLDEN, case, f"{s}.mate{m}.fastq.gz")) for m in (1, 2))
    elif case == "edge":
        for s in EDGE:
            reads[s] = tuple(fastq.read_fastq(os.path.join(GOLDEN, case, f"{s}.mate{m}.fastq")) for m in (1, 2))
    else:
        import hashlib
        sy = g["synth"]
        make = {"trio": synth.make_trio, "pair": synth.make_pair, "quad": synth.make_quad}[sy.get("kind", "trio")]
        trio, _ = make(sy["seed"], sy["genome_len"], sy["n_pairs"])
        for s in trio:
            s1, q1, s2, q2 = trio[s]
            md5 = hashlib.md5(s1.tobytes() + q1.tobytes() + s2.tobytes() + q2.tobytes()).hexdigest()
            assert md5 == sy["input_md5"][s], "synthetic generator drifted from the committed golden inputs"
            r1, r2 = fastq.from_matrix(s1, q1), fastq.from_matrix(s2, q2)
            r1.names = [f"@{s[0]}{i:09d}".encode() for i in range(len(r1))]
            r2.names = list(r1.names)
            reads[s] = (r1, r2)
    return g, reads


def both_mates(pair):
    return fastq.concat(pair[0], pair[1])


def overlap_fuzz_input(seed: int, n_reads: int = 70, genome_len: int = 1500) -> str:
    """a small random .fastqd: reads of a 1.5 kb genome on both strands, 1 % substitutions, 0.5 % N,
    30 % of the reads with per-base depths in {1,2,3,7} (end trimming), strand strings of 1-3 of + - ."""
    import random
    rnd = random.Random(seed)
    comp = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}
    genome = "".join(rnd.choice("ACGT") for _ in range(genome_len))
    recs = []
    for i in range(n_reads):
        L = rnd.randint(60, 160)
        p = rnd.randint(0, len(genome) - L)
        s = list(genome[p:p + L])
        for j in range(L):
            r = rnd.random()
            if r < 0.01:
                s[j] = rnd.choice("ACGT")
            elif r < 0.015:
                s[j] = "N"
        s = "".join(s)
        if rnd.random() < 0.5:
            s = "".join(comp[c] for c in reversed(s))
        q = "".join(rnd.choice("#,5AFJ") for _ in range(L))
        d = [rnd.choice([1, 1, 2, 3, 7]) for _ in range(L)] if rnd.random() < 0.3 else [1] * L
        st = "".join(rnd.choice("+-.") for _ in range(rnd.randint(1, 3)))
        recs.append(f"@r{i}\n{s}\n+\n{q}\n{st}\n{' '.join(map(str, d))}\n")
    return "".join(recs)
