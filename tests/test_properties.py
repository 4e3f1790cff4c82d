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


# ---- the read pull (src/RUFUS.Filter.cpp:196-262, src/RUFUS.Filter.ss.cpp:164-201) against a
# string-level statement of the streak rule
read_q = st.integers(1, 70).flatmap(lambda n: st.tuples(st.text(alphabet="ACGTN", min_size=n, max_size=n),
                                                        st.text(alphabet="#,5J", min_size=n, max_size=n)))


def _naive_hits(seq, qual, k, minq, mutant, last_base_tested):
    streak = hits = 0
    for i in range(len(seq) if last_base_tested else len(seq) - 1):
        if ord(qual[i]) - 33 < minq or seq[i] == "N":
            streak = 0
        else:
            streak += 1
        if streak >= k and seq[i - k + 1:i + 1] in mutant:
            hits += 1
    return hits


@settings(max_examples=150, deadline=None)
@given(st.lists(st.tuples(read_q, read_q), min_size=1, max_size=8), st.integers(3, 12), st.sampled_from([0, 10, 15, 30]),
       st.integers(1, 3), st.data())
def test_read_pull_equals_the_streak_rule(pairs, k, minq, thresh, data):
    # mutant k-mers: some windows of the reads themselves (either strand), so that there are hits
    windows = [s[i:i + k] for (s, _), (t, _) in pairs for s in (s, t) for i in range(len(s) - k + 1) if "N" not in s[i:i + k]]
    chosen = data.draw(st.lists(st.sampled_from(windows), max_size=6)) if windows else []
    chosen = [w if data.draw(st.booleans()) else w[::-1].translate(COMP) for w in chosen]
    mutant = set(chosen) | {w[::-1].translate(COMP) for w in chosen}
    rset = oracle.rufus_set_from_kmers(chosen) if chosen else np.zeros(0, np.uint64)
    m1 = [p[0] for p in pairs]
    m2 = [p[1] for p in pairs]
    s1, o1 = _pack([s for s, _ in m1]); q1, _ = _pack([q for _, q in m1])
    s2, o2 = _pack([s for s, _ in m2]); q2, _ = _pack([q for _, q in m2])
    keep = oracle.filter_pairs(s1, q1, o1, s2, q2, o2, k, minq, thresh, rset)
    want = [int(_naive_hits(a[0], a[1], k, minq, mutant, False) >= thresh or
                _naive_hits(b[0], b[1], k, minq, mutant, False) >= thresh) for a, b in pairs]
    assert keep.tolist() == want
    hits = oracle.filter_single(s1, q1, o1, k, minq, rset)
    assert hits.tolist() == [_naive_hits(a[0], a[1], k, minq, mutant, True) for a in m1]


@settings(max_examples=80, deadline=None)
@given(st.lists(reads_st, min_size=1, max_size=4), st.integers(4, 16))
def test_merge_emits_keys_of_exactly_one_input(samples, k):
    """modified merge (merge_files.cc:86-152): a k-mer is printed iff exactly one input file holds it
    and its count there is >= 5"""
    tabs = []
    for r in samples:
        seq, offs = _pack(r * 5)
        tabs.append(oracle.count(seq, offs, k, True, 2))
    mk, mc = oracle.merge_unique(tabs)
    want = {}
    for i, (keys, counts) in enumerate(tabs):
        others = set().union(*[set(t[0].tolist()) for j, t in enumerate(tabs) if j != i]) if len(tabs) > 1 else set()
        for x, c in zip(keys.tolist(), counts.tolist()):
            if x not in others and c >= 5:
                want[x] = c
    assert dict(zip(mk.tolist(), mc.tolist())) == want
    assert mk.tolist() == sorted(mk.tolist())


@settings(max_examples=80, deadline=None)
@given(reads_st, st.lists(reads_st, min_size=0, max_size=3), st.integers(4, 16), st.integers(1, 4), st.integers(4, 30),
       st.integers(0, 6))
def test_build_select_is_the_threshold_rule(subj, ctls, k, lo, hi, max_ctl):
    """RUFUS.Build (src/RUFUS.Build.cpp:224-262): subject entries with lo <= count <= hi whose summed
    control count is <= max_ctl"""
    seq, offs = _pack(subj * 3)
    ts = oracle.count(seq, offs, k, True, 1)
    tcs = []
    for r in ctls:
        seq, offs = _pack(r)
        tcs.append(oracle.count(seq, offs, k, True, 1))
    keys, sc, cs = oracle.build_select(ts, tcs, lo, hi, max_ctl)
    lut = [dict(zip(t[0].tolist(), t[1].tolist())) for t in tcs]
    want = [(x, c, sum(d.get(x, 0) for d in lut)) for x, c in zip(ts[0].tolist(), ts[1].tolist())
            if lo <= c <= hi and sum(d.get(x, 0) for d in lut) <= max_ctl]
    assert list(zip(keys.tolist(), sc.tolist(), cs.tolist())) == want
