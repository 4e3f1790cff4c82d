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


# ---------------------------------------------------------------------------------------------------------
# The large configs (BASELINE.json configs[2..4]), scaled to what the GPU-minute budget of a
# build round allows: per-GPU sizes are fixed and the genome grows with the GPU count (true coverage at every N).
# Inputs come from the counter-based generator (rg_synth_pairs_dev / rg_synth_pairs_host: same bytes on the GPU
# and on the CPU, include/rufus_gpu.h), 5 % of the genome in repeat families. Full-size CPU counting is out of
# reach, so parity is SLICED: the keys with slice_hash(key) % SLICE_MOD == 0 (a fixed function of the key) of
# every sample's table, and the HashList restricted to that slice, against tests/golden/big_digests.json, which
# tests/golden/make_big_digests.py computes with the CPU restatement on the identical reads.
# ---------------------------------------------------------------------------------------------------------
SLICE_MOD = 64
BIG = {
    # name: kind, samples (generator sample ids, subject first), pairs per GPU per sample, genome per GPU, k, seed, cutoffs
    "c3": {"title": "30x WGS trio, k=25 (BASELINE configs[2] scaled)", "kind": 0, "samples": (0, 1, 2),
           "pairs": (4_000_000, 4_000_000, 4_000_000), "genome_unit": 40_000_000, "k": 25, "seed": 1001,
           "min_cov": (5,), "max_cov": 150},
    "c4": {"title": "tumor 60x vs normal 30x, k=25 (BASELINE configs[3] scaled)", "kind": 1, "samples": (0, 1),
           "pairs": (6_000_000, 3_000_000), "genome_unit": 30_000_000, "k": 25, "seed": 2001,
           "min_cov": (5,), "max_cov": 300},
    "c5": {"title": "proband + 3 controls, 30x each, k=31, min-count sweep (BASELINE configs[4] scaled)", "kind": 0,
           "samples": (0, 1, 2, 3), "pairs": (3_000_000,) * 4, "genome_unit": 30_000_000, "k": 31, "seed": 3001,
           "min_cov": (3, 5, 8, 12, 20), "max_cov": 150},
}
BIG_DIGESTS = os.path.join(os.path.dirname(DIGESTS), "big_digests.json")


def big_spec(cfg: dict, sample_index: int, rank: int, world: int, first: int = 0, n: int | None = None):
    """generator coordinates of pairs [first, first + n) of rank `rank`'s slice of a sample"""
    from . import api
    pairs = cfg["pairs"][sample_index]
    n = pairs - first if n is None else n
    return api.SynthSpec(seed=cfg["seed"], genome_len=cfg["genome_unit"] * world, first_pair=rank * pairs + first,
                         n_pairs=n, read_len=READ_LEN, sample=cfg["samples"][sample_index], kind=cfg["kind"], repeats=1)


def big_key(name: str, world: int) -> str:
    c = BIG[name]
    return f"{name}_seed{c['seed']}_g{c['genome_unit']}x{world}_p{'-'.join(str(p) for p in c['pairs'])}x{world}_k{c['k']}_slice{SLICE_MOD}"


def slice_hash(keys) -> np.ndarray:
    """the library's rg_hash32b, vectorised: a fixed 32-bit function of the key"""
    k = np.asarray(keys, np.uint64)
    m32 = np.uint64(0xFFFFFFFF)
    h = ((k & m32) * np.uint64(0xC2B2AE35) + (k >> np.uint64(32)) * np.uint64(0x27D4EB2F)) & m32
    h ^= h >> np.uint64(15)
    return ((h * np.uint64(0x2C1B3C6D)) & m32).astype(np.uint32)


def table_digest(keys, counts) -> dict:
    """digest of a (key, count) set: order-free (sorted by key)"""
    keys = np.asarray(keys, np.uint64)
    counts = np.asarray(counts, np.uint32)
    o = np.argsort(keys, kind="stable")
    return {"n": int(keys.size), "md5": hashlib.md5(keys[o].tobytes() + counts[o].tobytes()).hexdigest()}


def big_expected(name: str, world: int):
    try:
        return json.load(open(BIG_DIGESTS)).get(big_key(name, world))
    except OSError:
        return None
