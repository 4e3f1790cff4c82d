"""Host-side FASTQ ingest: 4-line records -> concatenated ASCII + offsets.

Mirrors what the reference feeds its tools: `jellyfish count` consumes only the
sequence lines (mer_overlap_sequence_parser.hpp:179-207) and RUFUS.Filter reads
records four lines at a time (RUFUS.Filter.cpp:162-194).
"""
from __future__ import annotations

import gzip

import numpy as np


class Reads:
    __slots__ = ("names", "seq", "qual", "offs", "plus")

    def __init__(self, names, seq, qual, offs, plus=None):
        self.names, self.seq, self.qual, self.offs, self.plus = names, seq, qual, offs, plus

    def __len__(self):
        return len(self.offs) - 1


def read_fastq(path: str) -> Reads:
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rb") as f:
        lines = f.read().split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    n = len(lines) // 4
    names = lines[0:4 * n:4]
    seqs = lines[1:4 * n:4]
    plus = lines[2:4 * n:4]
    quals = lines[3:4 * n:4]
    lens = np.fromiter((len(s) for s in seqs), dtype=np.uint64, count=n)
    offs = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(lens, out=offs[1:])
    seq = np.frombuffer(b"".join(seqs), dtype=np.uint8).copy()
    qual = np.frombuffer(b"".join(q.ljust(len(s), b"!")[:len(s)] for s, q in zip(seqs, quals)), dtype=np.uint8).copy()
    return Reads(names, seq, qual, offs, plus)


def concat(a: Reads, b: Reads) -> Reads:
    offs = np.concatenate([a.offs, b.offs[1:] + a.offs[-1]])
    return Reads(a.names + b.names, np.concatenate([a.seq, b.seq]), np.concatenate([a.qual, b.qual]), offs)


def from_matrix(seq: np.ndarray, qual: np.ndarray | None = None, prefix: str = "r") -> Reads:
    n, L = seq.shape
    offs = (np.arange(n + 1, dtype=np.uint64) * np.uint64(L))
    return Reads(None, seq.reshape(-1), None if qual is None else qual.reshape(-1), offs)
