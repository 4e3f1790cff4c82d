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


def test_read_pull_from_the_bam_equals_the_golden_pull(tmp_path):
    """`RUFUS.Filter HashList X.bam X.bam …` (tools/rg_bam_pairs.hpp: decoded and re-paired in process) pulls the pairs
    the reference pulled from the re-paired FASTQ files of the same BAM (tests/golden/make_golden.py bam_case)."""
    g = golden("testRun_bam")
    work = str(tmp_path)
    bam = os.path.join(GOLDEN, "testRun_bam", "Child.bam")
    open(os.path.join(work, "HashList"), "w").write("\n".join(g["hashlist_sorted"]) + "\n")
    sh(f"timeout 60 {BIN}/RUFUS.Filter HashList {bam} {bam} Child {g['k']} {g['minq']} {g['thresh']} 4", work)
    o1 = open(os.path.join(work, "Child.Mutations.Mate1.fastq")).read().splitlines()
    o2 = open(os.path.join(work, "Child.Mutations.Mate2.fastq")).read().splitlines()
    assert sorted(o1[0::4]) == g["pulled_mate1_names_sorted"]
    assert sorted(o2[0::4]) == g["pulled_mate2_names_sorted"]
    # and it is the very output of the same tool on the chain's re-paired files
    sh(f"{BIN}/rg_bamview view -F 3328 {bam} | {BIN}/PassThroughSamCheck.stranded c.chr t > t.log", work)
    sh(f"timeout 60 {BIN}/RUFUS.Filter HashList t.mate1.fastq t.mate2.fastq Chain {g['k']} {g['minq']} {g['thresh']} 4", work)
    for m in (1, 2):
        assert open(os.path.join(work, f"Child.Mutations.Mate{m}.fastq"), "rb").read() == \
            open(os.path.join(work, f"Chain.Mutations.Mate{m}.fastq"), "rb").read()
