"""GPU path against the reference's OWN known-answer vectors: the md5 sums of Jellyfish 2.2.5's test
scripts (tests/parallel_hashing.sh:6-19, tests/multi_file.sh:6-9 of the vendored tarball) on the inputs of
tests/generate_sequence.sh:6-7. Needs the reference's generate_sequence (oracle/_ref, built from the
reference tree by oracle/Makefile; it travels to the GPU box with the other reference binaries)."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from rufus_b200 import api

pytestmark = pytest.mark.gpu
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


def _histo_md5(h, lower=1, upper=10001):
    # `jellyfish histo` without -f: the non-empty buckets only
    return hashlib.md5("".join(f"{i} {int(v)}\n" for i, v in enumerate(h) if v and lower <= i <= upper).encode()).hexdigest()


@pytest.mark.skipif(not os.path.exists(GENSEQ), reason="oracle/_ref/generate_sequence not present")
@pytest.mark.parametrize("mode", [api.MODE_TILED, api.MODE_DIRECT])
def test_jellyfish_known_answers_on_gpu(tmp_path, mode):
    subprocess.run([GENSEQ, "-o", "seq10m", "-m", "10", "-m", "22", "-s", "3141592653", "10000000"], cwd=tmp_path, check=True,
                   capture_output=True)
    subprocess.run([GENSEQ, "-o", "seq1m", "-s", "1040104553"] + ["1000000"] * 5, cwd=tmp_path, check=True, capture_output=True)
    ctx = api.Context(0)

    def histo(files):
        seqs = [s for f in files for s in _fasta_records(os.path.join(tmp_path, f))]
        seq = np.frombuffer(b"".join(seqs), np.uint8)
        offs = np.cumsum([0] + [len(s) for s in seqs]).astype(np.uint64)
        t = api.KmerTable.count(ctx, seq, offs, 15, True, mode=mode)
        h = t.histo(min_count=1)
        t.free()
        return h

    h = histo(["seq10m.fa"])
    assert _histo_md5(h) == "864c0b0826854bdc72a85d170549b64b"          # parallel_hashing_m15_s2M.histo
    assert _histo_md5(h, 2, 3) == "94625cd2d59e278f08421a673eb0926a"    # ..._m15_s2M_L2_U3.histo
    h = histo(["seq1m_0.fa", "seq1m_1.fa", "seq1m_2.fa", "seq10m.fa", "seq1m_3.fa", "seq1m_4.fa"])
    assert _histo_md5(h) == "d93b7678037814c256d1d9120a0e6422"          # multi_file_m15_s2M.histo
    ctx.close()
