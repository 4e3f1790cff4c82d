"""The host logic of the drop-in executables themselves (tools/jellyfish.cpp, tools/rufus_filter.cpp), run here without a
GPU against the RECORDING stand-in of tests/abi_recorder.py (no computation: it records what the executables hand to
the C ABI). What is pinned: `jellyfish count` pushes exactly the byte stream of the input reader (FASTA/FASTQ serial,
FASTQ on several threads, BAM in process) in growing chunks; RUFUS.Filter hands rg_filter_pairs the same arrays and
writes the same files whether the pairs come from two FASTQ files, two FIFOs fed in lock-step, or the BAM itself."""
import gzip
import os
import subprocess

import pytest

import abi_recorder
from conftest import GOLDEN, ROOT

BIN = os.path.join(ROOT, "tools", "bin")
BAM = os.path.join(GOLDEN, "testRun_bam", "Child.bam")


@pytest.fixture(scope="module")
def rec(tmp_path_factory):
    for exe in ("jellyfish", "RUFUS.Filter", "RUFUS.Filter.single", "rg_input_to_seq"):
        if not os.path.exists(os.path.join(BIN, exe)):
            pytest.fail(f"tools/bin/{exe} is missing: run `make tools`")
    return abi_recorder.build(str(tmp_path_factory.mktemp("recorder")))


def reader_stream(path, *args, env=None):
    return subprocess.run([os.path.join(BIN, "rg_input_to_seq"), *args, str(path)], capture_output=True, check=True, env=env).stdout


def count(rec, work, *inputs, threads=4, extra_env=None):
    env = dict(rec, **(extra_env or {}))
    r = subprocess.run([os.path.join(BIN, "jellyfish"), "count", "--disk", "-m", "25", "-L", "2", "-s", "16M", "-t", str(threads),
                        "-o", "x.Jhash", "-C", *[str(i) for i in inputs]], cwd=work, capture_output=True, env=env)
    assert r.returncode == 0, r.stderr
    return abi_recorder.pushes(rec), [int(l.split()[1]) for l in r.stderr.decode().splitlines() if l.startswith("push ")]


def test_jellyfish_count_pushes_the_reader_stream(rec, tmp_path):
    fq1, fq2 = tmp_path / "m1.fq", tmp_path / "m2.fq"
    fq1.write_bytes(gzip.open(os.path.join(GOLDEN, "testRun", "Child.mate1.fastq.gz")).read())
    fq2.write_bytes(gzip.open(os.path.join(GOLDEN, "testRun", "Child.mate2.fastq.gz")).read())
    want = reader_stream(fq1) + reader_stream(fq2)
    got, sizes = count(rec, tmp_path, fq1, fq2)
    assert got == want and sum(sizes) == len(want)
    # the same files scanned on several threads (thresholds lowered so that these small files qualify)
    got, _ = count(rec, tmp_path, fq1, fq2, extra_env={"RG_PARALLEL_FASTQ_MIN_BYTES": "1", "RG_PARALLEL_FASTQ_SEGMENT": "50000"})
    assert got == want
    # through a FIFO, as RunJellyForRUFUS.sh:28-29 feeds it
    r = subprocess.run(["bash", "-c", f"{BIN}/jellyfish count --disk -m 25 -L 2 -s 16M -t 4 -o x.Jhash -C <(cat {fq1} {fq2})"],
                       cwd=tmp_path, capture_output=True, env=rec)
    assert r.returncode == 0 and abi_recorder.pushes(rec) == want
    # the BAM in process == the chain's FASTQ
    chain = subprocess.run(f"{BIN}/rg_bamview view -F 3328 {BAM} | {BIN}/PassThroughSamCheck /dev/null", shell=True, capture_output=True, check=True).stdout
    (tmp_path / "chain.fq").write_bytes(chain)
    got_bam, _ = count(rec, tmp_path, BAM, threads=3)
    got_chain, _ = count(rec, tmp_path, tmp_path / "chain.fq")
    assert got_bam == got_chain == reader_stream(BAM)


def test_jellyfish_count_chunks_grow_and_cover_a_large_input(rec, tmp_path):
    fq = tmp_path / "big.fq"
    one = gzip.open(os.path.join(GOLDEN, "testRun", "Child.mate1.fastq.gz")).read()
    fq.write_bytes(one * 60)                        # ~40 MB of FASTQ: several chunks (8 MiB, 32 MiB, ...)
    want = reader_stream(fq)
    got, sizes = count(rec, tmp_path, fq, threads=1)
    assert got == want and len(sizes) >= 2 and sizes[0] >= (8 << 20) and sizes[0] < (10 << 20)
    got, _ = count(rec, tmp_path, fq, threads=4, extra_env={"RG_PARALLEL_FASTQ_MIN_BYTES": "1"})
    assert got == want


def run_filter(rec, work, m1, m2, stub, threads=4, shell_prefix=None):
    cmd = f"{BIN}/RUFUS.Filter HashList {m1} {m2} {stub} 25 15 1 {threads}"
    if shell_prefix:
        cmd = shell_prefix + cmd + "; wait"
    r = subprocess.run(["bash", "-c", cmd], cwd=work, capture_output=True, env=rec)
    assert r.returncode == 0, r.stderr
    digests = [l for l in r.stderr.decode().splitlines() if l.startswith("filter_pairs")]
    outs = [open(os.path.join(work, f"{stub}.Mutations.Mate{m}.fastq"), "rb").read() for m in (1, 2)]
    return digests, outs, r.stdout.decode()


def test_rufus_filter_hands_over_the_same_pairs_from_files_fifos_and_bam(rec, tmp_path):
    work = str(tmp_path)
    open(os.path.join(work, "HashList"), "w").write("ACGTACGTACGTACGTACGTACGTA 7\n")
    subprocess.run(f"{BIN}/rg_bamview view -F 3328 {BAM} | {BIN}/PassThroughSamCheck.stranded c.chr t > /dev/null", shell=True, cwd=work, check=True)
    d_files, o_files, log = run_filter(rec, work, "t.mate1.fastq", "t.mate2.fastq", "files")
    assert "Read in 2008 records" in log and len(d_files) == 1 and "n=2008 minq=15 thresh=1" in d_files[0]
    # every pair whose index is not 1 mod 3, verbatim and in order
    recs = open(os.path.join(work, "t.mate1.fastq"), "rb").read().split(b"\n")
    want = b"".join(b"\n".join(recs[4 * i:4 * i + 4]) + b"\n" for i in range(2008) if i % 3 != 1)
    assert o_files[0] == want
    d_bam, o_bam, _ = run_filter(rec, work, BAM, BAM, "bam", threads=3)
    d_dash, o_dash, _ = run_filter(rec, work, BAM, "-", "dash", threads=1)
    assert d_bam == d_files == d_dash and o_bam == o_files == o_dash
    fifo = (f"rm -f f.mate1.fastq f.mate2.fastq; mkfifo f.mate1.fastq f.mate2.fastq; "
            f"{BIN}/rg_bamview view -F 3328 {BAM} | {BIN}/PassThroughSamCheck.stranded f.chr f > /dev/null & timeout 60 ")
    d_fifo, o_fifo, _ = run_filter(rec, work, "f.mate1.fastq", "f.mate2.fastq", "fifo", shell_prefix=fifo)
    assert d_fifo == d_files and o_fifo == o_files


def test_rufus_filter_single_marks_the_names(rec, tmp_path):
    work = str(tmp_path)
    open(os.path.join(work, "HashList"), "w").write("ACGTACGTACGTACGTACGTACGTA 7\n")
    fq = tmp_path / "r.fq"
    fq.write_bytes(b"".join(b"@r%d x\nACGTACGTAC\n+\nIIIIIIIIII\n" % i for i in range(10)))
    r = subprocess.run([os.path.join(BIN, "RUFUS.Filter.single"), "HashList", str(fq), "single", "25", "15", "2", "4"], cwd=work, capture_output=True, env=rec)
    assert r.returncode == 0 and b"n=10 minq=15" in r.stderr
    out = open(os.path.join(work, "single.Mutations.fastq"), "rb").read()
    assert out == b"".join(b"@r%d x:MH%d\nACGTACGTAC\n+\nIIIIIIIIII\n" % (i, i % 4) for i in range(10) if i % 4 >= 2)


def test_a_batch_that_inflates_to_many_megabytes_is_pushed_in_bounded_chunks(rec, tmp_path):
    """200 000 identical reads compress 200:1: one pipeline batch holds all of them (57 MB inflated, 30 MB of bases). The chunk handed
    over must still stay near the chunk size — the multi-GPU count stages it in a device buffer of chunk + 16 MB."""
    import struct
    from bam_build import bgzf
    seq = "ACGTTGCA" * 18 + "ACGTAC"                # 150 bases
    packed = bytes((("=ACMGRSVTWYHKDBN".index(seq[i]) << 4) | "=ACMGRSVTWYHKDBN".index(seq[i + 1])) for i in range(0, 150, 2))
    r = struct.pack("<iiBBHHHiiii", 0, 100, 3, 60, 4680, 1, 99, 150, 0, 400, 300) + b"rd\0" + struct.pack("<I", (150 << 4)) + packed + bytes([30] * 150)
    text = "@HD\tVN:1.6\n"
    body = b"BAM\x01" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", 1) + struct.pack("<i", 5) + b"chr1\0" + struct.pack("<i", 10 ** 6)
    n = 200000
    bam = tmp_path / "same.bam"
    bam.write_bytes(bgzf(body + (struct.pack("<i", len(r)) + r) * n, 65280))
    assert os.path.getsize(bam) < (1 << 20)
    got, sizes = count(rec, tmp_path, bam, threads=2)
    assert got == (seq.encode() + b"N") * n
    assert sizes[0] < (8 << 20) + (5 << 20) and max(sizes) < (32 << 20) + (5 << 20)
