"""Host-side k-mer text <-> packed helpers (Jellyfish packing: first base in the
most-significant 2 bits, A0 C1 G2 T3 — mer_dna.hpp:46-63,330-354)."""
from __future__ import annotations

import numpy as np

_B = np.frombuffer(b"ACGT", dtype=np.uint8)


def keys_to_text(keys: np.ndarray, k: int) -> list[str]:
    keys = np.asarray(keys, dtype=np.uint64)
    if keys.size == 0:
        return []
    sh = (2 * (k - 1 - np.arange(k, dtype=np.uint64)))[None, :]
    m = _B[((keys[:, None] >> sh) & np.uint64(3)).astype(np.intp)]
    return [r.tobytes().decode() for r in m]


def text_to_key(s: str) -> int:
    v = 0
    for ch in s:
        v = (v << 2) | "ACGT".index(ch.upper())
    return v


def dump_text(keys: np.ndarray, counts: np.ndarray, k: int, sep: str = " ") -> list[str]:
    return [f"{m}{sep}{c}" for m, c in zip(keys_to_text(keys, k), counts.tolist())]
