"""`jellyfish count X.bam` (tools/rg_bam_input.hpp: the BAM decoded in process, sequences straight into the pinned
upload buffers) writes the very `.Jhash` the FIFO chain of RunJellyForRUFUS.sh:28-29 gives with our executables
(rg_bamview -F 3328 | PassThroughSamCheck | jellyfish count), which test_gpu_cli.py pins to the reference's
goldens. The host side of this path (the byte stream handed to the counter) is checked without a GPU in
test_input_stream.py; this file only adds the last step and sorts last on purpose."""
import hashlib
import os
import subprocess

import pytest

from cases import golden
from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "tools", "bin")
JF = os.path.join(BIN, "jellyfish")


def sh(cmd, cwd):
    return subprocess.run(cmd, shell=True, check=True, cwd=cwd, capture_output=True, executable="/bin/bash")


def records(path):
    raw = open(path, "rb").read()
    return raw[9 + int(raw[:9]):]


@pytest.mark.parametrize("sample", ["Child", "Father"])
def test_count_of_a_bam_equals_count_of_the_chain(sample, tmp_path):
    g = golden("testRun_bam")
    k, work = g["k"], str(tmp_path)
    bam = os.path.join(GOLDEN, "testRun_bam", sample + ".bam")
    lo = g["L"] if sample == "Child" else g["ctlL"]
    sh(f"{BIN}/rg_bamview view -F 3328 {bam} | {BIN}/PassThroughSamCheck x.chr > chain.fq", work)
    sh(f"timeout 60 {JF} count --disk -m {k} -L {lo} -s 16M -t 4 -o chain.Jhash -C chain.fq", work)
    sh(f"timeout 60 {JF} count --disk -m {k} -L {lo} -s 16M -t 4 -o direct.Jhash -C {bam}", work)
    assert records(os.path.join(work, "direct.Jhash")) == records(os.path.join(work, "chain.Jhash"))
    dump = sh(f"{JF} dump -c direct.Jhash | LC_ALL=C sort", work).stdout
    gs = g["samples"][sample]
    assert dump.count(b"\n") == gs["n_distinct"] and hashlib.md5(dump).hexdigest() == gs["dump_sorted_md5"]
