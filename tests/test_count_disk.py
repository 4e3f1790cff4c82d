"""`jellyfish count --disk` when the table does not fit (hash_counter.hpp:182-202, count_main.cc:316-338): the table
is written out as intermediate files <out><N>, which are merged at the end — counts summed over the parts, THEN -L/-U
(tar:jellyfish-2.2.5/jellyfish/merge_files.cc:78-79) — and unlinked. On the CPU:
  * the merge (tools/rg_jf_parts.hpp, `jellyfish rg-merge-parts`) against the reference's binary: parts counted by the
    reference from pieces of an input, merged by us = the reference's count of the whole input, record for record;
    the intermediate files the REFERENCE's own --disk run leaves behind (--no-merge), merged by us, hold what its
    count of the whole input holds (dump / histo / query of the reference's binary on our file);
  * the control flow of the executable against the recording stand-in of the C ABI: which tables are written out when,
    with which arguments, and that the final file is the merge of those parts.
The GPU side of the same path is in tests/test_zzzz_count_disk_gpu.py."""
import hashlib
import gzip
import os
import subprocess

import pytest

import abi_recorder
from conftest import GOLDEN, ROOT

JF = os.path.join(ROOT, "tools", "bin", "jellyfish")
REF = os.path.join(ROOT, "oracle", "_ref", "jellyfish")
pytestmark = pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(JF)), reason="needs oracle/_ref/jellyfish and tools/bin/jellyfish")


def sh(cmd, cwd, **kw):
    return subprocess.run(cmd, shell=True, check=True, cwd=cwd, capture_output=True, executable="/bin/bash", **kw)


def records(path):
    raw = open(path, "rb").read()
    return raw[9 + int(raw[:9]):]


@pytest.fixture(scope="module")
def inputs(tmp_path_factory):
    d = tmp_path_factory.mktemp("disk")
    for name, src in (("a.fq", "Child.mate1"), ("b.fq", "Child.mate2"), ("c.fq", "Mother.mate1")):
        open(d / name, "wb").write(gzip.open(os.path.join(GOLDEN, "testRun", src + ".fastq.gz")).read())
    for i, f in enumerate(("a.fq", "b.fq", "c.fq")):          # one part per file, every count kept
        sh(f"{REF} count -m 25 -C -s 16M -t 2 -o part{i} {f}", d)
    sh(f"{REF} count -m 25 -C -L 2 -s 16M -t 2 -o whole a.fq b.fq c.fq", d)
    return d


def test_merge_of_the_parts_is_the_count_of_the_whole(inputs):
    d = str(inputs)
    r = sh(f"{JF} rg-merge-parts -L 2 -o merged part0 part1 part2", d)
    n = len(records(os.path.join(d, "whole"))) // 11
    assert n > 15000 and r.stderr.decode().strip() == f"{n} records"
    assert records(os.path.join(d, "merged")) == records(os.path.join(d, "whole"))
    # -L and -U apply to the SUMMED counts
    sh(f"{REF} count -m 25 -C -L 3 -U 7 -s 16M -t 2 -o whole37 a.fq b.fq c.fq && {JF} rg-merge-parts -L 3 -U 7 -o merged37 part0 part1 part2", d)
    assert records(os.path.join(d, "merged37")) == records(os.path.join(d, "whole37")) and len(records(os.path.join(d, "whole37"))) > 0
    # one part: a copy with the filter applied; the header is the part's
    sh(f"{JF} rg-merge-parts -o single part0", d)
    assert open(os.path.join(d, "single"), "rb").read() == open(os.path.join(d, "part0"), "rb").read()


def test_the_references_own_intermediate_files(inputs):
    """--disk with a table of 20 k slots: the reference spills eight times; --no-merge keeps its parts"""
    d = str(inputs)
    sh(f"{REF} count --disk --no-merge -m 25 -C -L 2 -s 20k -t 2 -o sp a.fq b.fq c.fq", d)
    parts = sorted(p for p in os.listdir(d) if p.startswith("sp") and p[2:].isdigit())
    assert len(parts) >= 3
    sh(f"{JF} rg-merge-parts -L 2 -o spm " + " ".join(parts), d)
    md5 = lambda cmd: hashlib.md5(sh(cmd, d).stdout).hexdigest()
    assert md5(f"{REF} dump -c spm | LC_ALL=C sort") == md5(f"{REF} dump -c whole | LC_ALL=C sort")
    assert md5(f"{REF} histo spm") == md5(f"{REF} histo whole")
    some = sh(f"{REF} dump -c whole | head -40 | cut -d' ' -f1", d).stdout.decode().split()
    q = lambda db: sh(f"{REF} query {db} " + " ".join(some), d).stdout      # (binary search: needs the (pos, key) order)
    assert q("spm") == sh(f"{REF} dump -c whole | head -40", d).stdout


def test_merge_refuses_what_jellyfish_refuses(inputs):
    d = str(inputs)
    sh(f"{REF} count -m 25 -C -s 1M -t 2 -o other a.fq", d)                     # another -s: another matrix (merge_files.cc:193-202)
    r = subprocess.run([JF, "rg-merge-parts", "-o", "x", "part0", "other"], cwd=d, capture_output=True)
    assert r.returncode == 1 and b"was not written with the size" in r.stderr
    raw = open(os.path.join(d, "part1"), "rb").read()
    open(os.path.join(d, "cut"), "wb").write(raw[:-5])
    r = subprocess.run([JF, "rg-merge-parts", "-o", "x", "part0", "cut"], cwd=d, capture_output=True)
    assert r.returncode == 1 and b"truncated record" in r.stderr
    # two records swapped: not a sorted file
    h = 9 + int(raw[:9])
    bad = raw[:h] + raw[h + 11:h + 22] + raw[h:h + 11] + raw[h + 22:]
    open(os.path.join(d, "swapped"), "wb").write(bad)
    r = subprocess.run([JF, "rg-merge-parts", "-o", "x", "swapped"], cwd=d, capture_output=True)
    assert r.returncode == 1 and b"not in (pos, key) order" in r.stderr
    r = subprocess.run([JF, "rg-merge-parts", "-o", "x", "a.fq"], cwd=d, capture_output=True)
    assert r.returncode == 1 and b"not a Jellyfish binary file" in r.stderr


def _count(env, d, extra_env, *args):
    e = dict(env, **extra_env)
    return subprocess.run([JF, "count", "--disk", "-m", "25", "-C", "-L", "2", "-s", "16M", "-t", "1", "-o", "out.Jhash", *args],
                          cwd=d, capture_output=True, env=e, timeout=120)


def test_count_spills_merges_and_unlinks(inputs, tmp_path):
    """The executable against the recorder: 27 MB of sequence arrive in chunks of 8 MB, then what is left. With a limit
    of 10 M distinct k-mers the second chunk may not join the first in one table: the table is written out (every
    count: min 0, 4-byte counts), emptied, counting goes on; the end writes the last table; the parts are merged with
    -L 2 and unlinked."""
    d = str(tmp_path)
    env = abi_recorder.build(tmp_path)
    seq = (b"ACGTTGCA" * 20)[:150]
    with open(os.path.join(d, "big.fa"), "wb") as f:
        for i in range(180_000):
            f.write(b">r\n" + seq + b"\n")
    parts_dir = str(inputs)
    r = _count(env, d, {"RG_COUNT_SPILL_DISTINCT": "10M", "RG_RECORD_PARTS": parts_dir}, "big.fa")
    assert r.returncode == 0, r.stderr
    log = [l for l in r.stderr.decode().splitlines() if l.split()[0] in ("push", "write_jhash", "clear")]
    kinds = [l.split()[0] for l in log]
    assert kinds == ["push", "write_jhash", "clear", "push", "write_jhash", "clear"], log
    w = [l for l in log if l.startswith("write_jhash")]
    assert w[0].split()[1:] == ["out.Jhash0", "min=0", "max=0", "size=16000000", "counter_len=4"]
    assert w[1].split()[1] == "out.Jhash1"
    # the final file = the merge of the two (canned) parts with -L 2; parts unlinked
    sh(f"{JF} rg-merge-parts -L 2 -o expect part0 part1", parts_dir)
    assert records(os.path.join(d, "out.Jhash")) == records(os.path.join(parts_dir, "expect"))
    assert not os.path.exists(os.path.join(d, "out.Jhash0")) and not os.path.exists(os.path.join(d, "out.Jhash1"))
    # the merge also leaves the histogram next to the file (as an unspilled count does): `histo -f` needs no GPU and
    # prints what the reference's histo prints for the merged file
    h = subprocess.run([JF, "histo", "-f", "out.Jhash"], cwd=d, capture_output=True, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert h.returncode == 0 and h.stdout == sh(f"{REF} histo -f expect", parts_dir).stdout and h.stdout.count(b"\n") == 10002
    # --no-merge keeps the parts and writes no final file (count_main.cc:325)
    os.remove(os.path.join(d, "out.Jhash"))
    r = _count(env, d, {"RG_COUNT_SPILL_DISTINCT": "10M", "RG_RECORD_PARTS": parts_dir}, "--no-merge", "big.fa")
    assert r.returncode == 0 and os.path.exists(os.path.join(d, "out.Jhash0")) and os.path.exists(os.path.join(d, "out.Jhash1"))
    assert not os.path.exists(os.path.join(d, "out.Jhash"))


def test_count_that_fits_never_looks(inputs, tmp_path):
    """below the limit nothing changes: no stats, no intermediate file, one .Jhash written with the user's -L"""
    d = str(tmp_path)
    env = abi_recorder.build(tmp_path)
    open(os.path.join(d, "small.fa"), "wb").write(b">r\nACGTACGTACGTACGTACGTACGTACGTACGT\n" * 1000)
    r = _count(env, d, {}, "small.fa")
    assert r.returncode == 0, r.stderr
    log = [l for l in r.stderr.decode().splitlines() if l.split()[0] in ("push", "write_jhash", "clear")]
    assert [l.split()[0] for l in log] == ["push", "write_jhash"] and "out.Jhash min=2 max=0" in log[1]
    # the table holds less than the limit although more bytes than the limit were pushed: stats say so, no spill
    seq = (b"ACGTTGCA" * 20)[:150]
    with open(os.path.join(d, "big.fa"), "wb") as f:
        for i in range(180_000):
            f.write(b">r\n" + seq + b"\n")
    r = _count(env, d, {"RG_COUNT_SPILL_DISTINCT": "25M", "RG_RECORD_DISTINCT": "1000"}, "big.fa")
    assert r.returncode == 0, r.stderr
    log = [l for l in r.stderr.decode().splitlines() if l.split()[0] in ("push", "write_jhash", "clear")]
    assert [l.split()[0] for l in log] == ["push", "push", "write_jhash"], log


def test_merge_differential_fuzz(inputs, tmp_path):
    """random k (other key widths), -C or not, table size, -L/-U, number of parts: pieces counted by the reference's binary and
    merged by us = the reference's count of the whole, record for record"""
    import random
    d, rng = str(tmp_path), random.Random(9)
    recs = open(os.path.join(str(inputs), "a.fq"), "rb").read().split(b"\n@")
    recs = [recs[0]] + [b"@" + r for r in recs[1:]]
    for it in range(4):
        k, nparts, canon = rng.choice([15, 21, 25, 31]), rng.randrange(2, 7), rng.choice(["-C", ""])
        sz, L, U = rng.choice(["1M", "4M", "16M"]), rng.choice([1, 2, 3]), rng.choice([0, 0, 50])
        sub = rng.sample(recs, 1200)
        cuts = sorted(rng.sample(range(1, 1200), nparts - 1)) + [1200]
        files, prev = [], 0
        for i, c in enumerate(cuts):
            open(os.path.join(d, f"f{i}.fq"), "wb").write(b"\n".join(sub[prev:c]) + b"\n")
            files.append(f"f{i}.fq")
            prev = c
        for i, fn in enumerate(files):
            sh(f"{REF} count -m {k} {canon} -s {sz} -t 2 -o p{i} {fn}", d)
        u = f"-U {U}" if U else ""
        sh(f"{REF} count -m {k} {canon} -L {L} {u} -s {sz} -t 2 -o whole " + " ".join(files), d)
        sh(f"{JF} rg-merge-parts -L {L} {u} -o merged " + " ".join(f"p{i}" for i in range(len(files))), d)
        assert records(os.path.join(d, "merged")) == records(os.path.join(d, "whole")), (it, k, nparts, canon, sz, L, U)
        assert len(records(os.path.join(d, "whole"))) > 0
