"""CPU-side parity of the host-only drop-ins (tools/bin/PassThroughSamCheck, .stranded; SURVEY §8a
P1/P2) against the committed outputs of the reference binaries, and against the binaries
themselves where oracle/_ref is present."""
import hashlib
import json
import os
import subprocess

import pytest

from conftest import GOLDEN, ROOT

BIN = os.path.join(ROOT, "tools", "bin")
REFBIN = os.path.join(ROOT, "oracle", "_ref")


def md5(b):
    return hashlib.md5(b).hexdigest()


@pytest.fixture(scope="module")
def sam():
    if not os.path.exists(os.path.join(BIN, "PassThroughSamCheck")):
        pytest.fail("tools/bin is missing: run `make tools`")
    return open(os.path.join(GOLDEN, "passthrough", "input.sam"), "rb").read()


def test_passthrough_plain_matches_reference(sam, tmp_path):
    g = json.load(open(os.path.join(GOLDEN, "passthrough", "golden.json")))
    r = subprocess.run([os.path.join(BIN, "PassThroughSamCheck"), str(tmp_path / "chr")], input=sam,
                       capture_output=True, check=True)
    assert md5(r.stdout) == g["plain_fastq_md5"]
    assert open(tmp_path / "chr").read() == g["plain_chr"]
    if os.path.exists(os.path.join(REFBIN, "PassThroughSamCheck")):
        ref = subprocess.run([os.path.join(REFBIN, "PassThroughSamCheck"), str(tmp_path / "chr_ref")], input=sam,
                             capture_output=True, check=True)
        assert ref.stdout == r.stdout


def test_passthrough_stranded_matches_reference(sam, tmp_path):
    g = json.load(open(os.path.join(GOLDEN, "passthrough", "golden.json")))
    subprocess.run([os.path.join(BIN, "PassThroughSamCheck.stranded"), str(tmp_path / "chr"), str(tmp_path / "st")],
                   input=sam, capture_output=True, check=True)
    m1 = open(tmp_path / "st.mate1.fastq", "rb").read()
    m2 = open(tmp_path / "st.mate2.fastq", "rb").read()
    assert (md5(m1), md5(m2)) == (g["stranded_mate1_md5"], g["stranded_mate2_md5"])
    assert m1.count(b"\n") // 4 == g["stranded_pairs"]
    assert open(tmp_path / "chr").read() == g["stranded_chr"]
