"""GPU parity: every stage of the hot path through the C ABI vs (a) the golden
outputs of the reference binaries and (b) the CPU oracle on the same inputs.
Bit-exact: integer/byte work only."""
import os
import sys

import numpy as np
import pytest

from cases import both_mates, load_case
from conftest import ROOT
from rufus_b200 import api, kmer, synth, fastq

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle  # noqa: E402

pytestmark = pytest.mark.gpu
# pair_k25 / quad_k31: scaled-down BASELINE configs[3] (tumor/normal, one control) and configs[4]
# (proband + 3 controls, k = 31, min-count sweep)
CASES = ["testRun", "edge", "synth_k25", "synth_k31", "pair_k25", "quad_k31"]
MODES = [api.MODE_DIRECT, api.MODE_TILED]


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


def _count_all(ctx, reads, k, mode):
    tabs = {}
    for s, pair in reads.items():
        r = both_mates(pair)
        tabs[s] = api.KmerTable.count(ctx, r.seq, r.offs, k, True, mode=mode)
    return tabs


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("case", CASES)
def test_pipeline_matches_reference_goldens(ctx, case, mode):
    g, reads = load_case(case)
    k = g["k"]
    names = list(reads)
    tabs = _count_all(ctx, reads, k, mode)
    mins = [g["L"]] + [g["ctlL"]] * (len(names) - 1)
    for s, lo in zip(names, mins):
        keys, counts = tabs[s].dump(min_count=lo)
        gs = g["samples"][s]
        assert keys.size == gs["n_distinct"]
        r = both_mates(reads[s])
        okeys, ocounts = oracle.count(r.seq, r.offs, k, True, lo)
        assert np.array_equal(keys, okeys) and np.array_equal(counts, ocounts)
        h = tabs[s].histo(min_count=lo)
        assert h.size == 10002
        assert {int(i): int(v) for i, v in enumerate(h) if v} == {int(a): b for a, b in gs["histo_nonzero"].items()}
    mk, mc, _ = api.merge_unique([tabs[s] for s in names], mins)
    assert sorted(kmer.dump_text(mk, mc, k, "\t")) == g["merge_sorted"]
    subj = names[0]
    hk, hc = api.subtract(tabs[subj], [tabs[s] for s in names[1:]], subj_min=mins[0], other_min=mins[1:],
                          min_cov=g["min_cov"], max_cov=g["max_cov"])
    hashlist = sorted(kmer.dump_text(hk, hc, k))
    assert hashlist == g["hashlist_sorted"]
    # query == jellyfish query on the subject .Jhash
    q = tabs[subj].query(mk, min_count=mins[0])
    sk, sc = tabs[subj].dump(min_count=mins[0])
    lut = dict(zip(sk.tolist(), sc.tolist()))
    assert q.tolist() == [lut.get(x, 0) for x in mk.tolist()]
    # read pull, both ways of building the Mutations set
    text = ("\n".join(hashlist) + "\n").encode()
    m1, m2 = reads[subj]
    for mset in (api.MutantSet.from_hashlist_text(ctx, text, k, paired=True), api.MutantSet.from_jf_keys(ctx, hk, k)):
        keep = mset.filter_pairs(m1.seq, m1.qual, m1.offs, m2.seq, m2.qual, m2.offs, g["minq"], g["thresh"])
        assert sorted(m1.names[i].decode() for i in np.flatnonzero(keep)) == g["pulled_mate1_names_sorted"]
        hits = mset.filter_single(m1.seq, m1.qual, m1.offs, g["minq"])
        got = sorted(f"{m1.names[i].decode()}:MH{hits[i]}" for i in np.flatnonzero(hits >= g["thresh"]))
        assert got == g["single_names_sorted"]
        mset.free()
    import hashlib
    for mc, sw in g.get("sweep", {}).items():
        hk, hc = api.subtract(tabs[subj], [tabs[s] for s in names[1:]], subj_min=mins[0], other_min=mins[1:],
                              min_cov=int(mc), max_cov=g["max_cov"])
        lines = sorted(kmer.dump_text(hk, hc, k))
        assert len(lines) == sw["hashlist_lines"]
        assert hashlib.md5(("\n".join(lines) + "\n").encode()).hexdigest() == sw["hashlist_md5"]
        mset = api.MutantSet.from_jf_keys(ctx, hk, k)
        keep = mset.filter_pairs(m1.seq, m1.qual, m1.offs, m2.seq, m2.qual, m2.offs, g["minq"], g["thresh"])
        assert sorted(m1.names[i].decode() for i in np.flatnonzero(keep)) == sw["pulled_mate1_names_sorted"]
        mset.free()
    for t in tabs.values():
        t.free()


def test_exclude_table_records_exclude_whatever_their_count(ctx):
    """SURVEY Appendix B.5 (`-e` exclude files, runRufus.sh:925): one more table in the subtraction,
    kept from count 1, against the golden made with the reference's merge + query."""
    g, reads = load_case("edge")
    k, names = g["k"], list(reads)
    tabs = [_count_all(ctx, reads, k, api.MODE_TILED)[s] for s in names]
    seq = np.frombuffer("".join(g["exclude"]["reads"]).encode(), np.uint8)
    offs = np.cumsum([0] + [len(r) for r in g["exclude"]["reads"]]).astype(np.uint64)
    tabs.append(api.KmerTable.count(ctx, seq, offs, k, True, mode=api.MODE_DIRECT))
    mins = [g["L"], g["ctlL"], g["ctlL"], 1]
    mk, mc, _ = api.merge_unique(tabs, mins)
    assert sorted(kmer.dump_text(mk, mc, k, "\t")) == g["exclude"]["merge_sorted"]
    hk, hc = api.subtract(tabs[0], tabs[1:], subj_min=mins[0], other_min=mins[1:], min_cov=g["min_cov"], max_cov=g["max_cov"])
    assert sorted(kmer.dump_text(hk, hc, k)) == g["exclude"]["hashlist_sorted"]
    for t in tabs:
        t.free()


@pytest.mark.parametrize("mode", MODES)
def test_medium_synthetic_vs_oracle(ctx, mode):
    """60k pairs/sample (18 Mbases/sample), two pushes per sample, stream and offsets inputs."""
    trio, _ = synth.make_trio(99, 600_000, 60_000)
    k = 25
    tabs, otabs = {}, {}
    for s in synth.SAMPLES:
        s1, q1, s2, q2 = trio[s]
        t = api.KmerTable(ctx, k, True, est_distinct=6_000_000, mode=mode)
        # mate 1 as a '\n'-separated stream (what the FASTQ sequence lines look like), mate 2 via offsets
        st = np.concatenate([s1, np.full((s1.shape[0], 1), 10, np.uint8)], axis=1)
        t.push_stream(st)
        r2 = fastq.from_matrix(s2)
        t.push_reads(r2.seq, r2.offs)
        tabs[s] = t.end()
        both = np.concatenate([s1, s2])
        otabs[s] = oracle.count(both.reshape(-1), np.arange(both.shape[0] + 1, dtype=np.uint64) * 150, k, True, 1)
        keys, counts = t.dump()
        assert np.array_equal(keys, otabs[s][0]) and np.array_equal(counts, otabs[s][1])
        st = t.stats()
        assert st["distinct"] == keys.size and st["occurrences"] == int(counts.sum())
    names = list(synth.SAMPLES)
    flt = lambda kc, lo: (kc[0][kc[1] >= lo], kc[1][kc[1] >= lo])
    ok, oc = oracle.hashlist(flt(otabs["Child"], 2), [flt(otabs[s], 2) for s in names[1:]], 5, 200)
    hk, hc = api.subtract(tabs["Child"], [tabs[s] for s in names[1:]], min_cov=5, max_cov=200)
    assert np.array_equal(hk, ok) and np.array_equal(hc, oc) and hk.size > 50
    s1, q1, s2, q2 = trio["Child"]
    offs = np.arange(s1.shape[0] + 1, dtype=np.uint64) * 150
    mset = api.MutantSet.from_jf_keys(ctx, hk, k)
    keep = mset.filter_pairs(s1, q1, offs, s2, q2, offs, 15, 1)
    rset = oracle.rufus_set_from_kmers(kmer.keys_to_text(hk, k))
    assert len(mset) == rset.size
    okeep = oracle.filter_pairs(s1, q1, offs, s2, q2, offs, k, 15, 1, rset)
    assert np.array_equal(keep, okeep) and keep.sum() > 20
    for thr in (2, 5):
        assert np.array_equal(mset.filter_pairs(s1, q1, offs, s2, q2, offs, 20, thr),
                              oracle.filter_pairs(s1, q1, offs, s2, q2, offs, k, 20, thr, rset))


@pytest.mark.parametrize("mode", MODES)
def test_table_growth_and_heavy_hitters(ctx, mode):
    """est_distinct far too small -> spill + grow; poly-A reads -> one key with a huge count."""
    rng = np.random.default_rng(3)
    reads = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=(40_000, 100))]
    reads[:5000] = ord("A")
    offs = np.arange(reads.shape[0] + 1, dtype=np.uint64) * 100
    t = api.KmerTable(ctx, 21, True, est_distinct=1000, mode=mode)
    t.push_reads(reads, offs)
    t.end()
    keys, counts = t.dump()
    okeys, ocounts = oracle.count(reads.reshape(-1), offs, 21, True, 1)
    assert np.array_equal(keys, okeys) and np.array_equal(counts, ocounts)
    assert counts.max() >= 5000 * 80
    t.free()


@pytest.mark.parametrize("k,canonical", [(1, True), (5, False), (16, True), (31, False), (32, True)])
def test_k_range_and_noncanonical(ctx, k, canonical):
    rng = np.random.default_rng(k)
    seq = np.frombuffer(b"ACGTNacgt", np.uint8)[rng.integers(0, 9, size=200_000)]
    lens = rng.integers(0, 80, size=4000)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    seq = seq[: int(offs[-1])]
    for mode in MODES:
        t = api.KmerTable.count(ctx, seq, offs, k, canonical, mode=mode)
        keys, counts = t.dump()
        okeys, ocounts = oracle.count(seq, offs, k, canonical, 1)
        assert np.array_equal(keys, okeys) and np.array_equal(counts, ocounts)
        t.free()


def test_empty_inputs(ctx):
    t = api.KmerTable(ctx, 25, True, est_distinct=100)
    t.push_reads(np.zeros(0, np.uint8), np.zeros(1, np.uint64))
    t.push_stream(np.zeros(0, np.uint8))
    t.end()
    assert t.dump()[0].size == 0 and int(t.histo().sum()) == 0
    hk, hc = api.subtract(t, [t])
    assert hk.size == 0
    mset = api.MutantSet.from_jf_keys(ctx, hk, 25)
    z = np.zeros(0, np.uint8)
    o = np.zeros(1, np.uint64)
    assert mset.filter_pairs(z, z, o, z, z, o).size == 0
    # reads shorter than k and empty reads never hit
    seq = np.frombuffer(b"ACGTACGTAC", np.uint8)
    qual = np.full(10, ord("J"), np.uint8)
    o = np.array([0, 0, 4, 4, 10], np.uint64)
    assert mset.filter_single(seq, qual, o).tolist() == [0, 0, 0, 0]


def test_from_pairs_roundtrip(ctx):
    rng = np.random.default_rng(1)
    keys = np.unique(rng.integers(0, 1 << 50, size=100_000, dtype=np.uint64))
    counts = rng.integers(1, 5000, size=keys.size, dtype=np.uint32)
    t = api.KmerTable.from_pairs(ctx, 25, keys, counts)
    dk, dc = t.dump()
    assert np.array_equal(dk, keys) and np.array_equal(dc, counts)
    assert np.array_equal(t.query(keys[:1000]), counts[:1000])
    assert t.query(np.array([(1 << 50) - 1 - int(keys[0])], np.uint64) + np.uint64(0)).size == 1
    t.free()


def test_two_level_partition_matches_single_level(monkeypatch):
    """Tables with more tiles than one sorted partition pass can histogram use two passes (stream ->
    hash-prefix groups -> tiles). RG_GROUPS_LOCAL forces that geometry on a small table: same dump,
    same histogram as the oracle, for the device-stream, host-stream and offsets entry points."""
    monkeypatch.setenv("RG_GROUPS_LOCAL", "5")
    c = api.Context(0)
    try:
        trio, _ = synth.make_trio(23, 200_000, 30_000)
        s1, _, s2, _ = trio["Child"]
        both = np.concatenate([s1, s2])
        offs = np.arange(both.shape[0] + 1, dtype=np.uint64) * 150
        okeys, ocounts = oracle.count(both.reshape(-1), offs, 25, True, 1)
        t = api.KmerTable.count(c, both, offs, 25, True, est_distinct=3_000_000, mode=api.MODE_TILED)
        keys, counts = t.dump(min_count=1)
        assert np.array_equal(keys, okeys) and np.array_equal(counts, ocounts)
        t.free()
        # host stream entry point (pieces + halo) on the same geometry, twice into the same table
        t = api.KmerTable(c, 25, True, 3_000_000, api.MODE_TILED)
        stream = np.concatenate([both, np.full((both.shape[0], 1), 10, np.uint8)], axis=1).reshape(-1)
        t.push_stream(stream)
        t.push_stream(stream)
        t.end()
        keys, counts = t.dump(min_count=1)
        assert np.array_equal(keys, okeys) and np.array_equal(counts, 2 * ocounts)
        t.free()
    finally:
        c.close()


def test_build_select_matches_oracle(ctx):
    """rg_build_select (RUFUS.Build) vs the oracle on a synthetic trio"""
    trio, _ = synth.make_trio(31, 100_000, 10_000)
    tabs, otabs = [], []
    for s in synth.SAMPLES:
        s1, _, s2, _ = trio[s]
        both = np.concatenate([s1, s2])
        offs = np.arange(both.shape[0] + 1, dtype=np.uint64) * 150
        tabs.append(api.KmerTable.count(ctx, both, offs, 25, True, mode=api.MODE_TILED))
        otabs.append(oracle.count(both.reshape(-1), offs, 25, True, 1))
    for mS, mx, mC in ((5, 60, 0), (2, 1000, 3), (1, 3, 100)):
        k1, c1, s1 = api.build_select(tabs[0], tabs[1:], mS, mx, mC)
        k2, c2, s2 = oracle.build_select(otabs[0], otabs[1:], mS, mx, mC)
        assert np.array_equal(k1, k2) and np.array_equal(c1, c2) and np.array_equal(s1, s2), (mS, mx, mC)
    for t in tabs:
        t.free()


def test_annotate_matches_oracle(ctx):
    """rg_annotate (AnnotateOverlap's per-base depth) vs the oracle on reads with planted mutant k-mers, N and low quality"""
    rng = np.random.default_rng(5)
    k, n, L = 25, 200, 120
    seq = rng.integers(0, 4, (n, L)).astype(np.uint8)
    seq = np.frombuffer(b"ACGT", np.uint8)[seq]
    qual = np.full((n, L), ord("I"), np.uint8)
    qual[rng.random((n, L)) < 0.02] = ord("#")          # Q2 < 3
    seq[rng.random((n, L)) < 0.01] = ord("N")
    kmers = []
    for i in range(0, n, 3):
        p = int(rng.integers(0, L - k))
        kmers.append(bytes(seq[i, p:p + k]).decode())
    kmers = [m for m in kmers if "N" not in m]
    # every 2nd planted k-mer is listed as its reverse complement
    comp = str.maketrans("ACGT", "TGCA")
    listed = [m if i % 2 else m.translate(comp)[::-1] for i, m in enumerate(kmers)]
    text = "".join(f"{m} 7\n" for m in listed).encode()
    offs = np.arange(n + 1, dtype=np.uint64) * L
    mset = api.MutantSet.from_hashlist_text(ctx, text, k, paired=True)
    got = mset.annotate(seq, qual, offs, 3)
    want = oracle.annotate(seq, qual, offs, k, oracle.rufus_set_from_kmers(listed), 3)
    assert np.array_equal(got, want) and int(want.max()) > 0
    mset.free()


def test_two_level_partition_growth_and_heavy_hitters(monkeypatch):
    """the same stress on the two-level geometry: slabs of groups AND of tiles overflow (poly-A reads),
    the table starts far too small and grows between pushes"""
    monkeypatch.setenv("RG_GROUPS_LOCAL", "3")
    c = api.Context(0)
    try:
        rng = np.random.default_rng(4)
        reads = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=(60_000, 100))]
        reads[:8000] = ord("A")
        offs = np.arange(reads.shape[0] + 1, dtype=np.uint64) * 100
        t = api.KmerTable(c, 21, True, est_distinct=200_000, mode=api.MODE_TILED)
        half = 30_000
        t.push_reads(reads[:half], offs[:half + 1])
        t.end(trim=False)
        t.push_reads(reads[half:], offs[: reads.shape[0] - half + 1])
        t.end()
        keys, counts = t.dump()
        okeys, ocounts = oracle.count(reads.reshape(-1), offs, 21, True, 1)
        assert np.array_equal(keys, okeys) and np.array_equal(counts, ocounts)
        t.free()
    finally:
        c.close()


@pytest.mark.parametrize("mode", MODES)
def test_split_device_pushes(ctx, mode, monkeypatch):
    """device pushes beyond RG_PUSH_SPLIT bytes are counted in pieces: whole reads for fixed-stride input, a 32-byte
    halo whose already-counted windows are skipped otherwise. Same table as one push."""
    import torch
    monkeypatch.setenv("RG_PUSH_SPLIT", "70000")
    trio, _ = synth.make_trio(41, 60_000, 3_000)
    s1, _, s2, _ = trio["Child"]
    both = np.concatenate([s1, s2])                                        # [6000, 150]
    offs = np.arange(both.shape[0] + 1, dtype=np.uint64) * 150
    okeys, ocounts = oracle.count(both.reshape(-1), offs, 25, True, 1)
    dev = torch.device("cuda", 0)
    # (a) fixed stride, no separators
    t = api.KmerTable(ctx, 25, True, 2_000_000, mode)
    flat = torch.from_numpy(both.reshape(-1).copy()).to(dev)
    t.push_dev(flat, flat.numel(), 150)
    t.end()
    keys, counts = t.dump(min_count=1)
    assert np.array_equal(keys, okeys) and np.array_equal(counts, ocounts)
    t.free()
    # (b) separator-delimited stream (151-byte rows): pieces cut anywhere, windows across the cut counted once
    t = api.KmerTable(ctx, 25, True, 2_000_000, mode)
    rows = np.concatenate([both, np.full((both.shape[0], 1), 10, np.uint8)], axis=1).reshape(-1)
    stream = torch.from_numpy(rows.copy()).to(dev)
    t.push_dev(stream, stream.numel(), 0)
    t.end()
    keys, counts = t.dump(min_count=1)
    assert np.array_equal(keys, okeys) and np.array_equal(counts, ocounts)
    t.free()
    # (c) one long read without separators: every cut is inside a run of k-mers
    t = api.KmerTable(ctx, 25, True, 2_000_000, mode)
    t.push_dev(flat, flat.numel(), 0)
    t.end()
    keys, counts = t.dump(min_count=1)
    k2, c2 = oracle.count(both.reshape(-1), np.array([0, both.size], np.uint64), 25, True, 1)
    assert np.array_equal(keys, k2) and np.array_equal(counts, c2)
    t.free()


def test_bench_workload_matches_reference_digests(ctx):
    """BASELINE configs[1] / SURVEY §8d C2 at FULL size (seed 42, 5 Mb, 500 k pairs x 3 samples), exactly as bench.py
    builds it: histogram, HashList and pulled read names against digests of the reference binaries' output
    (tests/golden/bench_digests.json, tests/golden/make_bench_golden.py). Bit-exact."""
    import torch
    from rufus_b200 import pipeline, workload
    exp = workload.expected(1)
    assert exp is not None and exp["hashlist_lines"] == 1265 and exp["pulled_pairs"] == 682
    trio = workload.make_slice(workload.SEED, workload.GENOME_UNIT, workload.PAIRS, 0, 1)
    dev = torch.device("cuda", 0)
    ctx.use_stream(torch.cuda.current_stream().cuda_stream)
    streams = [torch.from_numpy(np.concatenate([trio[s][0], trio[s][2]]).reshape(-1)).to(dev) for s in synth.SAMPLES]
    subj = tuple(torch.from_numpy(a.reshape(-1)).to(dev) for a in trio["Child"])
    est = int(1.03 * workload.GENOME_UNIT + 6.2 * 2 * workload.PAIRS)
    pipe = pipeline.DevicePipeline(ctx, workload.K, streams, subj, workload.READ_LEN, est, workload.MIN_COV,
                                   workload.MAX_COV, workload.MINQ, workload.THRESH)
    pipe.step()
    pipe.step()          # twice: cleared and re-used tables give the same result
    g = pipe.gathered_results()
    par = workload.check(1, g["histo"], g["keys"], g["counts"], g["names"])
    pipe.free()
    ctx.use_stream(None)
    assert par["ok"], par


def test_synthetic_generator_device_equals_host(ctx):
    """the counter-based read generator (synth-trio v2) writes the same bytes on the GPU and on the CPU, for every
    sample kind, with and without repeat families, and for a slice that does not start at pair 0"""
    import ctypes as C
    import torch
    L = api.lib()
    dev = torch.device("cuda", 0)
    for kind, sample, repeats, first in ((api.SYNTH_TRIO, 0, 1, 0), (api.SYNTH_TRIO, 3, 0, 12345), (api.SYNTH_PAIR, 0, 1, 777),
                                         (api.SYNTH_PAIR, 1, 1, 0)):
        sp = api.SynthSpec(seed=99, genome_len=300_000, first_pair=first, n_pairs=3000, read_len=150, sample=sample,
                           kind=kind, repeats=repeats)
        host = api.synth_pairs_host(sp)
        bufs = [torch.empty(3000 * 150, dtype=torch.uint8, device=dev) for _ in range(4)]
        api._check(L.rg_synth_pairs_dev(ctx.h, C.byref(sp), *[api._ptr(b) for b in bufs]), "rg_synth_pairs_dev")
        ctx.sync()
        for h, d in zip(host, bufs):
            assert np.array_equal(h.reshape(-1), d.cpu().numpy())
    # and the sliced dump is the slice of the full dump
    s1, _, s2, _ = host
    reads = np.concatenate([s1, s2])
    offs = np.arange(reads.shape[0] + 1, dtype=np.uint64) * 150
    t = api.KmerTable.count(ctx, reads, offs, 25, True)
    keys, counts = t.dump(min_count=2)
    dk = torch.empty(keys.size + 1, dtype=torch.int64, device=dev)
    dc = torch.empty(keys.size + 1, dtype=torch.int32, device=dev)
    n = L.rg_dump_slice_dev(t.h, 2, 16, 5, api._ptr(dk), api._ptr(dc), keys.size + 1, None)
    sk, sc2 = dk[:n].cpu().numpy().view(np.uint64), dc[:n].cpu().numpy().view(np.uint32)
    o = np.argsort(sk)
    m = oracle.slice_hash(keys) % 16 == 5
    assert n > 0 and np.array_equal(sk[o], keys[m]) and np.array_equal(sc2[o], counts[m])
    t.free()
