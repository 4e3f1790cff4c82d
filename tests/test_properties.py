"""Size-independent properties of the path, checked on the CPU oracle with random inputs (hypothesis).
The same properties hold for the CUDA path through parity with the oracle; the ones that need no
reference output at all (linearity, strand symmetry) are what the full-size GPU runs rely on."""
import os
import sys

import numpy as np
from hypothesis import given, settings, strategies as st

from conftest import ROOT
from rufus_b200 import kmer

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle  # noqa: E402

ALPHABET = "ACGTacgtN"
reads_st = st.lists(st.text(alphabet=ALPHABET, min_size=0, max_size=80), min_size=0, max_size=12)
COMP = str.maketrans("ACGTacgtN", "TGCAtgcaN")


def _pack(reads):
    seq = np.frombuffer("".join(reads).encode(), np.uint8)
    offs = np.cumsum([0] + [len(r) for r in reads]).astype(np.uint64)
    return seq, offs


def _count(reads, k, canonical=True, lower=1):
    seq, offs = _pack(reads)
    keys, counts = oracle.count(seq, offs, k, canonical, lower)
    return dict(zip(keys.tolist(), counts.tolist()))


def _naive(reads, k, canonical):
    """mer_iterator.hpp:59-89 + mer_dna.hpp:235-258 spelled out on strings"""
    out = {}
    for r in reads:
        r = r.upper()
        for i in range(len(r) - k + 1):
            w = r[i:i + k]
            if any(c not in "ACGT" for c in w):
                continue
            if canonical:
                w = min(w, w[::-1].translate(COMP))
            key = kmer.text_to_key(w)
            out[key] = out.get(key, 0) + 1
    return out


@settings(max_examples=150, deadline=None)
@given(reads_st, st.integers(1, 32), st.booleans())
def test_count_equals_the_definition(reads, k, canonical):
    if k == 32 and not canonical:
        k = 31
    assert _count(reads, k, canonical) == _naive(reads, k, canonical)


@settings(max_examples=100, deadline=None)
@given(reads_st, reads_st, st.integers(1, 32))
def test_count_is_linear_in_the_reads(a, b, k):
    ca, cb, cab = _count(a, k), _count(b, k), _count(a + b, k)
    keys = set(ca) | set(cb)
    assert set(cab) == keys and all(cab[x] == ca.get(x, 0) + cb.get(x, 0) for x in keys)


@settings(max_examples=100, deadline=None)
@given(reads_st, st.integers(1, 32))
def test_canonical_count_ignores_the_strand(reads, k):
    rc = [r[::-1].translate(COMP) for r in reads]
    assert _count(reads, k) == _count(rc, k)


@settings(max_examples=100, deadline=None)
@given(st.integers(1, 32), st.data())
def test_key_text_round_trip(k, data):
    s = data.draw(st.text(alphabet="ACGT", min_size=k, max_size=k))
    key = kmer.text_to_key(s)
    assert kmer.keys_to_text(np.array([key], np.uint64), k) == [s]
    # first base in the most significant bits, A0 C1 G2 T3 (mer_dna.hpp:46-63,330-354)
    assert key == sum("ACGT".index(c) << (2 * (k - 1 - i)) for i, c in enumerate(s))


@settings(max_examples=60, deadline=None)
@given(reads_st, st.integers(2, 6), st.integers(1, 32))
def test_lower_bound_only_drops_entries(reads, lower, k):
    full, cut = _count(reads * 3, k), _count(reads * 3, k, True, lower)
    assert cut == {x: c for x, c in full.items() if c >= lower}


@settings(max_examples=60, deadline=None)
@given(reads_st, reads_st, st.integers(5, 20))
def test_hashlist_is_the_set_difference(subj, ctl, k):
    """merge_files.cc:86-154 + CheckJellyHashList.sh:12: subject k-mers (count >= 5, within the cutoffs)
    that no control holds at count >= 2"""
    subj, ctl = subj * 6, ctl * 2
    seq, offs = _pack(subj)
    ts = oracle.count(seq, offs, k, True, 2)
    seq, offs = _pack(ctl)
    tc = oracle.count(seq, offs, k, True, 2)
    hk, hc = oracle.hashlist(ts, [tc], 5, 1000)
    s, c = dict(zip(ts[0].tolist(), ts[1].tolist())), set(tc[0].tolist())
    want = {x: n for x, n in s.items() if n >= 5 and n <= 1000 and x not in c}
    assert dict(zip(hk.tolist(), hc.tolist())) == want
