"""`jellyfish count --disk` with a table limit small enough to force intermediate files (RG_COUNT_SPILL_DISTINCT): the
table is written out and emptied between chunks, the parts are merged at the end (tools/rg_jf_parts.hpp). The merged
.Jhash has to hold the very records of the unlimited run and of the reference's binary on the same input
(hash_counter.hpp:182-202, count_main.cc:316-338: "result is identical to single-table counting").
The host side (merge against the reference's own intermediate files, control flow of the executable) is checked
without a GPU in tests/test_count_disk.py; this file sorts last on purpose."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
JF = os.path.join(ROOT, "tools", "bin", "jellyfish")
REF = os.path.join(ROOT, "oracle", "_ref", "jellyfish")


def records(path):
    raw = open(path, "rb").read()
    return raw[9 + int(raw[:9]):]


def test_spilled_count_equals_the_plain_count(tmp_path):
    d = str(tmp_path)
    rng = np.random.default_rng(7)
    genome = rng.integers(0, 4, 300_000, dtype=np.uint8)
    starts = rng.integers(0, genome.size - 150, 170_000)
    reads = np.frombuffer(b"ACGT", np.uint8)[genome[starts[:, None] + np.arange(150)[None, :]]]
    err = rng.random(reads.shape) < 0.003                      # some singletons, some N
    reads[err] = np.frombuffer(b"ACGTN", np.uint8)[rng.integers(0, 5, int(err.sum()))]
    with open(os.path.join(d, "in.fa"), "wb") as f:            # 25.5 Mbases: one 8 MB chunk, then the rest
        f.write(b"".join(b">r\n" + r.tobytes() + b"\n" for r in reads))
    base = [JF, "count", "--disk", "-m", "25", "-C", "-L", "2", "-s", "16M", "-t", "2"]
    subprocess.run(base + ["-o", "plain.Jhash", "in.fa"], cwd=d, check=True, timeout=300)
    env = dict(os.environ, RG_COUNT_SPILL_DISTINCT="1M")
    subprocess.run(base + ["-o", "spilled.Jhash", "in.fa"], cwd=d, check=True, env=env, timeout=300)
    want = records(os.path.join(d, "plain.Jhash"))
    assert len(want) // 11 > 250_000
    assert records(os.path.join(d, "spilled.Jhash")) == want
    assert not [p for p in os.listdir(d) if p.startswith("spilled.Jhash") and p not in ("spilled.Jhash", "spilled.Jhash.rg_histo")]   # parts unlinked
    # both counts leave the histogram next to the file (histo_kernel for the plain one, the merge for the spilled one)
    histo = lambda f: subprocess.run([JF, "histo", "-f", f], cwd=d, check=True, capture_output=True).stdout
    assert histo("spilled.Jhash") == histo("plain.Jhash") and histo("plain.Jhash").count(b"\n") == 10002
    # the parts themselves: kept by --no-merge, more than one, and no final file
    subprocess.run(base + ["--no-merge", "-o", "kept.Jhash", "in.fa"], cwd=d, check=True, env=env, timeout=300)
    parts = sorted(p for p in os.listdir(d) if p.startswith("kept.Jhash"))
    assert parts == ["kept.Jhash0", "kept.Jhash1"]
    # merged on their own, with another threshold: what the plain count with that threshold holds
    subprocess.run([JF, "rg-merge-parts", "-L", "5", "-U", "200", "-o", "m5.Jhash"] + parts, cwd=d, check=True)
    subprocess.run(base[:6] + ["-L", "5", "-U", "200", "-s", "16M", "-o", "p5.Jhash", "in.fa"], cwd=d, check=True, timeout=300)
    assert records(os.path.join(d, "m5.Jhash")) == records(os.path.join(d, "p5.Jhash"))
    if os.path.exists(REF):                                   # and the reference's binary agrees
        subprocess.run([REF, "count", "-m", "25", "-C", "-L", "2", "-s", "16M", "-t", "4", "-o", "ref.Jhash", "in.fa"], cwd=d,
                       check=True, timeout=600)
        assert records(os.path.join(d, "ref.Jhash")) == want


@pytest.mark.parametrize("tiled", [False, True])
def test_clear_really_empties_the_table(tiled):
    """rg_table_clear between two counts (what the spill does between parts, and what every bench step does): nothing of
    the first input is left for readers or for a SMALL second push, which goes straight into global memory instead of
    through the tile insert that rewrites whole tiles (round-1 advice: the shortcut that only reset a flag)."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    from rufus_b200 import api
    ctx = api.Context(0)
    rng = np.random.default_rng(11)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    big = acgt[rng.integers(0, 4, (20_000, 150))]
    small = acgt[rng.integers(0, 4, (300, 150))]
    offs = lambda a: np.arange(a.shape[0] + 1, dtype=np.uint64) * 150
    t = api.KmerTable(ctx, 25, True, est_distinct=4_000_000, mode=api.MODE_TILED if tiled else api.MODE_DIRECT)
    t.push_reads(big.reshape(-1), offs(big))
    t.end(trim=False)
    assert t.stats()["distinct"] > 2_000_000
    t.clear()
    assert t.stats()["distinct"] == 0 and t.dump()[0].size == 0 and int(t.histo().sum()) == 0
    t.push_reads(small.reshape(-1), offs(small))
    t.end()
    keys, counts = t.dump()
    ok, oc = oracle.count(small.reshape(-1), offs(small), 25, True, 1)
    assert np.array_equal(keys, ok) and np.array_equal(counts, oc)
    t.free()
    ctx.close()
