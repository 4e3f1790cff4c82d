"""Pin the CPU oracle (oracle/rufus_oracle.c) to outputs of the reference's own
binaries (tests/golden/*/golden.json, made by tests/golden/make_golden.py)."""
import gzip
import json
import hashlib
import os
import subprocess

import numpy as np
import pytest

from cases import both_mates, load_case
from conftest import GOLDEN, ROOT
from rufus_b200 import kmer

import sys
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle  # noqa: E402

# pair_k25 / quad_k31: scaled-down BASELINE configs[3] (tumor/normal, one control) and configs[4]
# (proband + 3 controls, k = 31, min-count sweep)
CASES = ["testRun", "edge", "synth_k25", "synth_k31", "pair_k25", "quad_k31"]


def _dump_bytes(keys, counts, k):
    lines = sorted(kmer.dump_text(keys, counts, k))
    return ("\n".join(lines) + "\n").encode() if lines else b""


@pytest.mark.parametrize("case", CASES)
def test_count_histo_subtract_filter_match_reference(case):
    g, reads = load_case(case)
    k = g["k"]
    names = list(reads)
    tabs = {}
    for i, s in enumerate(names):
        r = both_mates(reads[s])
        lower = g["L"] if i == 0 else g["ctlL"]
        keys, counts = oracle.count(r.seq, r.offs, k, True, lower)
        tabs[s] = (keys, counts)
        gs = g["samples"][s]
        assert keys.size == gs["n_distinct"]
        assert hashlib.md5(_dump_bytes(keys, counts, k)).hexdigest() == gs["dump_sorted_md5"]
        h, base = oracle.histo(counts)
        assert h.size == gs["histo_rows"] == 10002 and base == 0
        nz = {int(i): int(v) for i, v in enumerate(h) if v}
        assert nz == {int(a): b for a, b in gs["histo_nonzero"].items()}
        text = "".join(f"{i} {v}\n" for i, v in enumerate(h))
        assert hashlib.md5(text.encode()).hexdigest() == gs["histo_md5"]
        dump_gz = os.path.join(GOLDEN, case, s + ".dump.sorted.gz")
        if os.path.exists(dump_gz):
            assert gzip.open(dump_gz).read() == _dump_bytes(keys, counts, k)
    subj = names[0]
    mk, mc = oracle.merge_unique([tabs[s] for s in names])
    assert sorted(kmer.dump_text(mk, mc, k, "\t")) == g["merge_sorted"]
    hk, hc = oracle.hashlist(tabs[subj], [tabs[s] for s in names[1:]], g["min_cov"], g["max_cov"])
    assert sorted(kmer.dump_text(hk, hc, k)) == g["hashlist_sorted"]
    rset = oracle.rufus_set_from_kmers(kmer.keys_to_text(hk, k))
    m1, m2 = reads[subj]
    keep = oracle.filter_pairs(m1.seq, m1.qual, m1.offs, m2.seq, m2.qual, m2.offs, k, g["minq"], g["thresh"], rset)
    got = sorted(m1.names[i].decode() for i in np.flatnonzero(keep))
    assert got == g["pulled_mate1_names_sorted"]
    hits = oracle.filter_single(m1.seq, m1.qual, m1.offs, k, g["minq"], rset)
    got = sorted(f"{m1.names[i].decode()}:MH{hits[i]}" for i in np.flatnonzero(hits >= g["thresh"]))
    assert got == g["single_names_sorted"]
    for mc, sw in g.get("sweep", {}).items():
        hk, hc = oracle.hashlist(tabs[subj], [tabs[s] for s in names[1:]], int(mc), g["max_cov"])
        lines = sorted(kmer.dump_text(hk, hc, k))
        assert len(lines) == sw["hashlist_lines"]
        assert hashlib.md5(("\n".join(lines) + "\n").encode()).hexdigest() == sw["hashlist_md5"]
        rset = oracle.rufus_set_from_kmers(kmer.keys_to_text(hk, k))
        keep = oracle.filter_pairs(m1.seq, m1.qual, m1.offs, m2.seq, m2.qual, m2.offs, k, g["minq"], g["thresh"], rset)
        assert sorted(m1.names[i].decode() for i in np.flatnonzero(keep)) == sw["pulled_mate1_names_sorted"]


def test_exclude_file_records_exclude_whatever_their_count():
    """an `-e` exclude .Jhash is one more merge input; a record seen ONCE excludes
    (the file is counted without -L), while a control k-mer seen once (-L 2) does not."""
    g, reads = load_case("edge")
    k, names = g["k"], list(reads)
    tabs = []
    for i, s in enumerate(names):
        r = both_mates(reads[s])
        tabs.append(oracle.count(r.seq, r.offs, k, True, g["L"] if i == 0 else g["ctlL"]))
    seq = np.frombuffer("".join(g["exclude"]["reads"]).encode(), np.uint8)
    offs = np.cumsum([0] + [len(r) for r in g["exclude"]["reads"]]).astype(np.uint64)
    tabs.append(oracle.count(seq, offs, k, True, 1))
    mk, mc = oracle.merge_unique(tabs)
    assert sorted(kmer.dump_text(mk, mc, k, "\t")) == g["exclude"]["merge_sorted"]
    hk, hc = oracle.hashlist(tabs[0], tabs[1:], g["min_cov"], g["max_cov"])
    assert sorted(kmer.dump_text(hk, hc, k)) == g["exclude"]["hashlist_sorted"]
    assert len(g["exclude"]["hashlist_sorted"]) < len(g["hashlist_sorted"])


def test_rufus_pack_known_answers():
    # Util.cpp:51-84: base i at bits (2i, 2i+1): A 00, C bit(2i+1), G bit(2i), T both.
    L = oracle.lib()
    assert L.ro_rufus_pack(b"A", 1) == 0
    assert L.ro_rufus_pack(b"C", 1) == 2
    assert L.ro_rufus_pack(b"G", 1) == 1
    assert L.ro_rufus_pack(b"T", 1) == 3
    assert L.ro_rufus_pack(b"AC", 2) == 8
    assert L.ro_rufus_pack(b"xT", 2) == 12     # unknown byte == 'A'
    assert L.ro_rufus_pack(b"c", 1) == 0       # lower case is NOT recognised by RUFUS


def test_count_semantics_small():
    seq = np.frombuffer(b"ACGTNacgtAC", dtype=np.uint8)
    offs = np.array([0, 5, 11], dtype=np.uint64)
    keys, counts = oracle.count(seq, offs, 3, True, 1)
    txt = dict(zip(kmer.keys_to_text(keys, 3), counts.tolist()))
    # read1 ACGT|N -> ACG(canon ACG), CGT(canon ACG); read2 acgtAC -> ACG,ACG(cgt),GTA,TAC(canon GTA)
    assert txt == {"ACG": 4, "GTA": 2}
    keys, counts = oracle.count(seq, offs, 3, False, 1)
    txt = dict(zip(kmer.keys_to_text(keys, 3), counts.tolist()))
    assert txt == {"ACG": 2, "CGT": 2, "GTA": 1, "TAC": 1}


def test_build_select_matches_reference_rufus_build():
    """oracle.build_select vs the committed outputs of the reference RUFUS.Build (tests/golden/make_build_golden.py)"""
    import gzip
    import json
    from rufus_b200 import kmer
    gold = json.load(open(os.path.join(GOLDEN, "build_golden.json")))
    names = {"testRun": ("Child", "Mother", "Father"), "edge": ("Subj", "CtlA", "CtlB")}
    tabs = {}
    for c in gold["cases"]:
        case, k = c["case"], c["k"]
        if case not in tabs:
            tabs[case] = []
            for s in names[case]:
                rows = [l.split() for l in gzip.open(os.path.join(GOLDEN, case, s + ".dump.sorted.gz")).read().decode().splitlines()]
                keys = np.array([kmer.text_to_key(r[0]) for r in rows], np.uint64)
                counts = np.array([int(r[1]) for r in rows], np.uint32)
                o = np.argsort(keys, kind="stable")
                tabs[case].append((keys[o], counts[o]))
        sub, ctl = tabs[case][0], tabs[case][1:]
        keys, counts, sums = oracle.build_select(sub, ctl, c["mS"], c["max"], c["mC"])
        text = kmer.keys_to_text(keys, k)
        pack = oracle.lib().ro_rufus_pack
        lines = [f"{pack(t.encode(), len(t))}\t{int(s)}\t{int(n)}\t{t}" for t, n, s in zip(text, counts, sums)]
        assert lines == c["lines"], (case, c["mS"], c["max"], c["mC"])


def test_annotate_matches_reference_annotate_overlap():
    """oracle.annotate vs the committed output of the reference AnnotateOverlap (tests/golden/make_annotate_golden.py)"""
    import json
    from rufus_b200 import fastq
    d = os.path.join(GOLDEN, "annotate")
    g = json.load(open(os.path.join(d, "golden.json")))
    r = fastq.read_fastq(os.path.join(d, "reads.fastq"))
    kmers = [l.split()[0] for l in open(os.path.join(d, "HashList")).read().splitlines()]
    depth = oracle.annotate(r.seq, r.qual, r.offs, len(kmers[0]), oracle.rufus_set_from_kmers(kmers), 3)
    got = ["".join(chr(min(int(x), 93) + 33) for x in depth[int(r.offs[i]):int(r.offs[i + 1])]) for i in range(len(r))]
    assert got[:len(g["depth_lines"])] == g["depth_lines"]
    names = [n.decode() for n in r.names]
    text = "".join(f"{names[i]}:MH0\n{bytes(r.seq[int(r.offs[i]):int(r.offs[i+1])]).decode()}\n+\n{got[i]}\n" for i in range(len(r)))
    assert hashlib.md5(text.encode()).hexdigest() == g["stdout_md5"]


def test_overlap_oracle_matches_reference_overlap():
    """oracle/overlap_oracle.py (greedy assembly, SURVEY §8a O1-O5: the next row to build) vs the outputs of the
    reference Overlap binary for the three rounds of Overlap.shorter.sh:141,153,165 (tests/golden/overlap)"""
    import json
    import overlap_oracle as oo
    d = os.path.join(GOLDEN, "overlap")
    g = json.load(open(os.path.join(d, "golden.json")))
    for a in g["rounds"]:
        text = open(os.path.join(d, a[0])).read()
        fq, fqd, _ = oo.overlap(text, a[1], int(a[2]), int(a[3]), int(a[5]), int(a[6]), int(a[8]), a[5], int(a[9]))
        assert fqd == open(os.path.join(d, a[7] + ".fastqd")).read(), a
        assert fq == open(os.path.join(d, a[7] + ".fastq")).read(), a


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "Overlap")),
                    reason="reference binaries (oracle/_ref) not present")
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_overlap_oracle_differential_vs_reference_binary(seed, tmp_path):
    """random overlapping reads (both strands, substitutions, N, depths > 1 that trigger end trimming, mixed
    strand strings) through the reference binary and the restatement. Depth 0 is left out: the reference
    reverses depth strings with Util::RevQual, which drops NUL bytes, and then indexes past the end."""
    import random
    import subprocess
    import overlap_oracle as oo
    from cases import overlap_fuzz_input
    text = overlap_fuzz_input(seed)
    args = [[".98", "100", "1", "FP", "20", "1", "o", "0", "1"], [".95", "40", "1", "FP", "20", "1", "o", "1", "1"],
            [".9", "30", "2", "X", "15", "2", "o", "2", "1"]][seed % 3]
    open(tmp_path / "in.fastqd", "w").write(text)
    subprocess.run([os.path.join(ROOT, "oracle", "_ref", "Overlap"), "in.fastqd", *args], cwd=tmp_path, check=True,
                   stdout=subprocess.DEVNULL)
    fq, fqd, bad = oo.overlap(text, args[0], int(args[1]), int(args[2]), int(args[4]), int(args[5]), int(args[7]),
                              args[4], 1)
    assert fqd == open(tmp_path / "o.fastqd").read()
    assert fq == open(tmp_path / "o.fastq").read()
    assert bad == open(tmp_path / "o.fastqbad.fastq").read()


def _reads_from_lists(names, seqs, quals):
    from rufus_b200 import fastq
    offs = np.cumsum([0] + [len(x) for x in seqs]).astype(np.uint64)
    return fastq.Reads([("@" + n).encode() for n in names], np.frombuffer("".join(seqs).encode(), np.uint8),
                       np.frombuffer("".join(quals).encode(), np.uint8), offs)


def test_bam_mode_matches_reference_chain():
    """BASELINE configs[0] in BAM mode (tests/golden/testRun_bam): the k-mers of EVERY record the
    generator prints (-F 3328, sequences as stored) are counted; the read pull runs on the re-paired,
    strand-restored mates (PassThroughSamCheck.stranded.cpp:197-279: the second-seen record of a name
    is mate 1, flag 0x10 => reverse complement + reversed quality, leftovers dropped)."""
    import bam_oracle
    g = json.load(open(os.path.join(GOLDEN, "testRun_bam", "golden.json")))
    k, names = g["k"], ("Child", "Mother", "Father")
    tabs, child = [], None
    for i, s in enumerate(names):
        _, lines = bam_oracle.view(os.path.join(GOLDEN, "testRun_bam", s + ".bam"), 3328)
        f = [l.split("\t") for l in lines]
        r = _reads_from_lists([x[0] for x in f], [x[9] for x in f], [x[10] for x in f])
        tabs.append(oracle.count(r.seq, r.offs, k, True, g["L"] if i == 0 else g["ctlL"]))
        gs = g["samples"][s]
        assert tabs[-1][0].size == gs["n_distinct"]
        assert hashlib.md5(_dump_bytes(*tabs[-1], k)).hexdigest() == gs["dump_sorted_md5"]
        if i == 0:
            child = f
    mk, mc = oracle.merge_unique(tabs)
    assert sorted(kmer.dump_text(mk, mc, k, "\t")) == g["merge_sorted"]
    hk, hc = oracle.hashlist(tabs[0], tabs[1:], g["min_cov"], g["max_cov"])
    assert sorted(kmer.dump_text(hk, hc, k)) == g["hashlist_sorted"]
    comp = str.maketrans("ACGTN", "TGCAN")
    parked, m1, m2 = {}, [], []
    for x in child:
        seq, qual = x[9], x[10]
        if int(x[1]) & 16:
            seq, qual = "".join(c for c in seq[::-1] if c in "ACGTN").translate(comp), qual[::-1]
        if x[0] in parked:
            m1.append((x[0], seq, qual))
            m2.append(parked.pop(x[0]))
        else:
            parked[x[0]] = (x[0], seq, qual)
    assert len(m1) == g["repaired_pairs"]
    r1, r2 = (_reads_from_lists(*zip(*m)) for m in (m1, m2))
    rset = oracle.rufus_set_from_kmers(kmer.keys_to_text(hk, k))
    keep = oracle.filter_pairs(r1.seq, r1.qual, r1.offs, r2.seq, r2.qual, r2.offs, k, g["minq"], g["thresh"], rset)
    assert sorted(r1.names[i].decode() for i in np.flatnonzero(keep)) == g["pulled_mate1_names_sorted"]


GENSEQ = os.path.join(ROOT, "oracle", "_ref", "generate_sequence")


def _fasta_records(path):
    seqs, cur = [], []
    for line in open(path, "rb").read().split(b"\n"):
        if line.startswith(b">"):
            if cur:
                seqs.append(b"".join(cur))
            cur = []
        else:
            cur.append(line)
    if cur:
        seqs.append(b"".join(cur))
    return seqs


def _histo_text(counts, lower=1, upper=None):
    """`jellyfish histo` without -f (histo_main.cc:64-84): the non-empty buckets only, low = 1,
    high = 10000, every count above 10000 in bucket 10001"""
    c = counts[counts >= lower]
    if upper is not None:
        c = c[c <= upper]
    h = np.bincount(np.minimum(c, 10001).astype(np.int64), minlength=10002)
    return "".join(f"{i} {v}\n" for i, v in enumerate(h) if v and i >= 1).encode()


@pytest.mark.skipif(not os.path.exists(GENSEQ), reason="oracle/_ref/generate_sequence not built (needs /root/reference)")
def test_jellyfish_known_answers(tmp_path):
    """The reference's OWN known-answer vectors for this path: the md5 sums in the test scripts of the
    vendored Jellyfish 2.2.5 tarball (tests/parallel_hashing.sh:6-19, tests/multi_file.sh:6-9) on the
    inputs of tests/generate_sequence.sh:6-7 (fixed seeds). k = 15, canonical; the m = 40 vectors are
    outside this path (k <= 32)."""
    subprocess.run([GENSEQ, "-o", "seq10m", "-m", "10", "-m", "22", "-s", "3141592653", "10000000"], cwd=tmp_path, check=True,
                   capture_output=True)
    subprocess.run([GENSEQ, "-o", "seq1m", "-s", "1040104553"] + ["1000000"] * 5, cwd=tmp_path, check=True, capture_output=True)

    def count(files):
        seqs = [s for f in files for s in _fasta_records(os.path.join(tmp_path, f))]
        seq = np.frombuffer(b"".join(seqs), np.uint8)
        offs = np.cumsum([0] + [len(s) for s in seqs]).astype(np.uint64)
        return oracle.count(seq, offs, 15, True, 1)[1]

    c = count(["seq10m.fa"])
    assert hashlib.md5(_histo_text(c)).hexdigest() == "864c0b0826854bdc72a85d170549b64b"          # ..._m15_s2M.histo
    assert hashlib.md5(_histo_text(c, 2, 3)).hexdigest() == "94625cd2d59e278f08421a673eb0926a"    # ..._m15_s2M_L2_U3.histo
    c = count(["seq1m_0.fa", "seq1m_1.fa", "seq1m_2.fa", "seq10m.fa", "seq1m_3.fa", "seq1m_4.fa"])
    assert hashlib.md5(_histo_text(c)).hexdigest() == "d93b7678037814c256d1d9120a0e6422"          # multi_file_m15_s2M.histo


def test_slice_counter_is_the_slice_of_the_full_count():
    """ro_slice_* (sliced parity of the large configs) = ro_count restricted to slice_hash(key) % mod == rem, and the
    host side of the synthetic generator is deterministic and sample-specific"""
    from rufus_b200 import api
    sp = api.SynthSpec(seed=7, genome_len=200_000, first_pair=0, n_pairs=4000, read_len=150, sample=0, kind=0, repeats=1)
    s1, q1, s2, q2 = api.synth_pairs_host(sp)
    again = api.synth_pairs_host(sp)
    assert all(np.array_equal(a, b) for a, b in zip((s1, q1, s2, q2), again))
    sp.sample = 1
    assert not np.array_equal(api.synth_pairs_host(sp)[0], s1)
    assert set(np.unique(s1).tolist()) <= set(b"ACGTN") and set(np.unique(q1).tolist()) <= set(b"J,#")
    reads = np.concatenate([s1, s2])
    offs = np.arange(reads.shape[0] + 1, dtype=np.uint64) * 150
    keys, counts = oracle.count(reads.reshape(-1), offs, 25, True, 2)
    sc = oracle.SliceCounter(25, True, 16, 3)
    sc.add(reads[:3000])
    sc.add(reads[3000:])
    sk, scn = sc.finish(2)
    m = oracle.slice_hash(keys) % 16 == 3
    assert sk.size > 100 and np.array_equal(keys[m], sk) and np.array_equal(counts[m], scn)
