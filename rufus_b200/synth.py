"""Seeded synthetic trio reads ("synth-trio v1").

Host-side numpy only: the same bytes feed the reference binaries (golden
generation, CPU baseline), the oracle and the CUDA path, so parity is checked
on identical inputs. Not on the product path.

Model: iid-uniform genome of `genome_len` bases; four parental haplotypes =
genome + private SNVs (1/2000); child = one maternal + one paternal haplotype
+ `n_denovo` de-novo SNVs (the signal RUFUS looks for). Reads: `read_len`-bp
pairs, insert U[300,400), mate 2 reverse-complemented, uniform start, fair-coin
haplotype; substitution error `err` per base with quality '#'(Q2) at error
bases, 1 % ','(Q11), otherwise 'J'(Q41); 'N' at rate `n_rate`.
"""
from __future__ import annotations

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
SAMPLES = ("Child", "Mother", "Father")


def _mutate(hap: np.ndarray, pos: np.ndarray, rng) -> None:
    hap[pos] = (hap[pos] + rng.integers(1, 4, size=pos.size, dtype=np.uint8)) & 3


def make_haplotypes(seed: int, genome_len: int, n_denovo: int | None = None):
    rng = np.random.default_rng(seed)
    genome = rng.integers(0, 4, size=genome_len, dtype=np.uint8)
    haps = {}
    for name in ("M0", "M1", "F0", "F1"):
        h = genome.copy()
        n = max(1, genome_len // 2000)
        _mutate(h, rng.choice(genome_len, size=n, replace=False), rng)
        haps[name] = h
    if n_denovo is None:
        n_denovo = max(20, genome_len // 100000)
    c0, c1 = haps["M0"].copy(), haps["F1"].copy()
    dn = np.sort(rng.choice(np.arange(500, genome_len - 500), size=n_denovo, replace=False))
    _mutate(c0, dn[::2], rng)
    _mutate(c1, dn[1::2], rng)
    return {"Child": (c0, c1), "Mother": (haps["M0"], haps["M1"]), "Father": (haps["F0"], haps["F1"])}, dn


def sample_pairs(haps, n_pairs: int, rng, read_len: int = 150, err: float = 0.002,
                 lowq: float = 0.01, n_rate: float = 2e-4, chunk: int = 1 << 17):
    """Return (seq1, qual1, seq2, qual2): uint8 [n_pairs, read_len] ASCII."""
    G = haps[0].size
    both = np.stack(haps)  # [2, G]
    out = [np.empty((n_pairs, read_len), dtype=np.uint8) for _ in range(4)]
    ar = np.arange(read_len, dtype=np.int64)
    for lo in range(0, n_pairs, chunk):
        n = min(chunk, n_pairs - lo)
        ins = rng.integers(300, 400, size=n)
        start = rng.integers(0, G - 400, size=n)
        hap = rng.integers(0, 2, size=n)
        for mate in (0, 1):
            if mate == 0:
                idx = start[:, None] + ar[None, :]
                codes = both[hap[:, None], idx]
            else:
                idx = (start + ins - 1)[:, None] - ar[None, :]
                codes = 3 - both[hap[:, None], idx]
            r = rng.random(size=(n, read_len), dtype=np.float32)
            is_err = r < err
            sub = rng.integers(1, 4, size=(n, read_len), dtype=np.uint8)
            codes = np.where(is_err, (codes + sub) & 3, codes).astype(np.uint8)
            seq = _ACGT[codes]
            qual = np.full((n, read_len), ord("J"), dtype=np.uint8)
            qual[(r >= err) & (r < err + lowq)] = ord(",")
            qual[is_err] = ord("#")
            is_n = rng.random(size=(n, read_len), dtype=np.float32) < n_rate
            seq[is_n] = ord("N")
            qual[is_n] = ord("#")
            out[2 * mate][lo:lo + n] = seq
            out[2 * mate + 1][lo:lo + n] = qual
    return tuple(out)


def make_trio(seed: int, genome_len: int, n_pairs: int, read_len: int = 150, slice_id: int | None = None, **kw):
    """dict sample -> (seq1, qual1, seq2, qual2); plus the de-novo positions.

    slice_id selects an independent batch of reads from the SAME individuals (one
    slice per GPU rank in the multi-GPU runs); None keeps the original stream."""
    haps, dn = make_haplotypes(seed, genome_len, kw.pop("n_denovo", None))
    trio = {}
    for i, s in enumerate(SAMPLES):
        rng = np.random.default_rng([seed, 1000 + i] + ([] if slice_id is None else [slice_id]))
        trio[s] = sample_pairs(haps[s], n_pairs, rng, read_len, **kw)
    return trio, dn


PAIR_SAMPLES = ("Tumor", "Normal")
QUAD_SAMPLES = ("Proband", "Mother", "Father", "Sibling")


def make_pair(seed: int, genome_len: int, n_pairs_normal: int, read_len: int = 150, tumor_factor: int = 2,
              n_somatic: int | None = None, **kw):
    """Tumor/normal pair (BASELINE configs[3]): the normal is a diploid genome (two
    haplotypes with private SNVs), the tumor is the same individual plus `n_somatic` somatic SNVs,
    sequenced `tumor_factor` times deeper. dict sample -> (seq1, qual1, seq2, qual2), subject first."""
    rng = np.random.default_rng([seed, 77])
    genome = rng.integers(0, 4, size=genome_len, dtype=np.uint8)
    normal = []
    for _ in range(2):
        h = genome.copy()
        _mutate(h, rng.choice(genome_len, size=max(1, genome_len // 1000), replace=False), rng)
        normal.append(h)
    if n_somatic is None:
        n_somatic = max(20, genome_len // 100000)
    som = np.sort(rng.choice(np.arange(500, genome_len - 500), size=n_somatic, replace=False))
    tumor = [normal[0].copy(), normal[1].copy()]
    _mutate(tumor[0], som[::2], rng)
    _mutate(tumor[1], som[1::2], rng)
    out = {}
    for i, (name, haps, n) in enumerate((("Tumor", tumor, tumor_factor * n_pairs_normal), ("Normal", normal, n_pairs_normal))):
        out[name] = sample_pairs(tuple(haps), n, np.random.default_rng([seed, 2000 + i]), read_len, **kw)
    return out, som


def make_quad(seed: int, genome_len: int, n_pairs: int, read_len: int = 150, **kw):
    """Proband + 3 controls (BASELINE configs[4]): the trio of make_haplotypes plus an
    unaffected sibling carrying the two parental haplotypes the proband did not inherit."""
    haps, dn = make_haplotypes(seed, genome_len, kw.pop("n_denovo", None))
    people = {"Proband": haps["Child"], "Mother": haps["Mother"], "Father": haps["Father"],
              "Sibling": (haps["Mother"][1], haps["Father"][0])}
    out = {}
    for i, s in enumerate(QUAD_SAMPLES):
        out[s] = sample_pairs(people[s], n_pairs, np.random.default_rng([seed, 3000 + i]), read_len, **kw)
    return out, dn


def fastq_bytes(seq: np.ndarray, qual: np.ndarray, prefix: str = "r", suffix: str = "") -> bytes:
    """Vectorised fixed-width 4-line FASTQ of a [n, L] read matrix."""
    n, L = seq.shape
    head = np.array([f"@{prefix}{i:09d}{suffix}\n".encode() for i in range(n)], dtype="S")
    hw = head.dtype.itemsize
    rec = np.empty((n, hw + L + 3 + L + 1), dtype=np.uint8)
    rec[:, :hw] = head.view(np.uint8).reshape(n, hw)
    rec[:, hw:hw + L] = seq
    rec[:, hw + L:hw + L + 3] = np.frombuffer(b"\n+\n", dtype=np.uint8)
    rec[:, hw + L + 3:hw + 2 * L + 3] = qual
    rec[:, -1] = ord("\n")
    return rec.tobytes()
