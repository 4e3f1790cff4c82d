"""The benchmark workload (BASELINE.json configs[1], SURVEY §8d "synth-trio v1" C2) and the digests that
pin a run's RESULT to the reference's output on the same inputs.

One definition shared by bench.py, tests/golden/make_bench_golden.py (which runs the reference's own
binaries on it) and the `-m gpu` parity tests:

  world = 1 : seed 42, genome 5 Mb, 500 k pairs x 150 bp per sample (30x), the unsliced read stream;
  world = N : genome 5 Mb x N (so every N is a 30x trio), rank r brings slice r of 500 k pairs per sample
              drawn from the whole genome; the job's result is that of the UNION of the slices.

Digests (what scripts/CheckJellyHashList.sh:12 and RUFUS.Filter.cpp:222-277 leave on disk, order-free):
  histo_md5    md5 of the subject's `jellyfish histo -f` file (10002 rows "count n")
  hashlist_md5 md5 of the sorted HashList lines "KMER COUNT"
  pulled_md5   md5 of the sorted names of the pulled mate-1 reads
"""
from __future__ import annotations

import hashlib
import json
import os

import numpy as np

from . import kmer, synth

SEED, PAIRS, GENOME_UNIT, READ_LEN = 42, 500_000, 5_000_000, 150
K, MIN_COV, MAX_COV, MINQ, THRESH = 25, 5, 150, 15, 1
DIGESTS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "bench_digests.json")


def make_slice(seed: int, genome_unit: int, pairs: int, rank: int, world: int):
    """reads of rank `rank` (dict sample -> seq1, qual1, seq2, qual2)"""
    trio, _ = synth.make_trio(seed, genome_unit * world, pairs, slice_id=None if world == 1 else rank)
    return trio


def name_prefix(sample: str, rank: int, world: int) -> str:
    """FASTQ read-name prefix (synth.fastq_bytes appends the 9-digit pair index)"""
    return sample[0] if world == 1 else f"{sample[0]}{rank}_"


def digest_key(world: int, seed: int = SEED, pairs: int = PAIRS, genome_unit: int = GENOME_UNIT, k: int = K) -> str:
    return f"seed{seed}_g{genome_unit}x{world}_p{pairs}x{world}_k{k}"


def _md5(lines) -> str:
    return hashlib.md5(("\n".join(lines) + "\n").encode()).hexdigest()


def histo_md5(h) -> str:
    return hashlib.md5("".join(f"{i} {int(v)}\n" for i, v in enumerate(h)).encode()).hexdigest()


def hashlist_md5(keys, counts, k: int = K) -> str:
    return _md5(sorted(kmer.dump_text(np.asarray(keys, np.uint64), np.asarray(counts), k, " ")))


def pulled_names(keep, sample: str, rank: int, world: int) -> list[str]:
    pre = name_prefix(sample, rank, world)
    return [f"@{pre}{int(i):09d}" for i in np.flatnonzero(np.asarray(keep))]


def pulled_md5(names) -> str:
    return _md5(sorted(names))


def expected(world: int, **kw):
    """the reference's digests for this workload, or None when no golden was generated for it"""
    try:
        return json.load(open(DIGESTS)).get(digest_key(world, **kw))
    except OSError:
        return None


def check(world: int, histo, keys, counts, names, **kw) -> dict:
    """compare a run's results with the reference digests; returns the `parity` object of the bench line"""
    exp = expected(world, **kw)
    got = {"histo_md5": histo_md5(histo), "hashlist_md5": hashlist_md5(keys, counts, kw.get("k", K)),
           "hashlist_lines": int(len(keys)), "pulled_md5": pulled_md5(names), "pulled_pairs": int(len(names))}
    if exp is None:
        return dict(got, ok=None, note="no reference digest committed for this workload")
    bad = [f for f in got if got[f] != exp[f]]
    return dict(got, ok=not bad, mismatch=bad, reference=digest_key(world, **kw))
