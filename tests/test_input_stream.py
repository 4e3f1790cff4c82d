"""The input side of `jellyfish count` (tools/rg_cli.hpp:read_sequences, tools/rg_bam_input.hpp:read_input), on the
host: tools/bin/rg_input_to_seq prints the byte stream the counter is handed. FASTA/FASTQ: the parsing rules of
mer_overlap_sequence_parser.hpp:124-207 (sequence lines joined, one 'N' after every record). BAM: what the reference's
chain `samtools view -F 3328 X.bam | PassThroughSamCheck` would feed it, produced in process — compared with OUR
chain (rg_bamview | PassThroughSamCheck, itself pinned to the reference tools in test_tools_host.py)."""
import os
import random
import subprocess
import sys

import pytest

from conftest import GOLDEN, ROOT

BIN = os.path.join(ROOT, "tools", "bin")
EXE = os.path.join(BIN, "rg_input_to_seq")


def stream(path, *args, env=None):
    r = subprocess.run([EXE, *args, str(path)], capture_output=True, env=env)
    assert r.returncode == 0, r.stderr
    return r.stdout


def chain(bam, excl=3328):
    p1 = subprocess.run([os.path.join(BIN, "rg_bamview"), "view", "-F", str(excl), str(bam)], capture_output=True, check=True)
    p2 = subprocess.run([os.path.join(BIN, "PassThroughSamCheck"), os.devnull], input=p1.stdout, capture_output=True, check=True)
    return p2.stdout


@pytest.fixture(scope="module", autouse=True)
def exe():
    if not os.path.exists(EXE):
        pytest.fail("tools/bin/rg_input_to_seq is missing: run `make tools`")


def test_fastq_and_fasta_streams(tmp_path):
    fq = tmp_path / "a.fq"
    fq.write_bytes(b"@r1 x\nACGT\n+\nIIII\n@r2\nAC\nGT\nNN\n+r2\nII\nII\nII\n@r3\n\n+\n\n@r4\nacgtn\n+\n@@@@@")
    assert stream(fq) == b"ACGTN" + b"ACGTNNN" + b"N" + b"acgtnN"
    fa = tmp_path / "a.fa"
    fa.write_bytes(b">s1\nACGT\nAC\n>s2\n>s3\nGGNN\n")
    assert stream(fa) == b"ACGTACN" + b"GGNN"          # an empty record adds nothing; a record ending in N gets no second N
    (tmp_path / "empty").write_bytes(b"")
    assert stream(tmp_path / "empty") == b""
    (tmp_path / "bad").write_bytes(b"hello\n")
    assert subprocess.run([EXE, str(tmp_path / "bad")], capture_output=True).returncode != 0
    # chunked: flushing every few bytes changes nothing
    assert stream(fq, "-c", "3") == stream(fq)


@pytest.mark.parametrize("sample", ["Child", "Mother", "Father"])
def test_bam_stream_equals_the_chain(sample, tmp_path):
    bam = os.path.join(GOLDEN, "testRun_bam", sample + ".bam")
    fq = tmp_path / "chain.fq"
    fq.write_bytes(chain(bam))
    want = stream(fq)
    assert want.count(b"N") >= 2000
    for args in ((), ("-t", "4"), ("-t", "3", "-c", "1000")):
        assert stream(bam, *args) == want
    # a different flag filter through the environment
    fq.write_bytes(chain(bam, 16))
    assert stream(bam, env=dict(os.environ, RG_BAM_EXCLUDE_FLAGS="16")) == stream(fq)


@pytest.mark.parametrize("seed", range(4))
def test_bam_stream_fuzz(seed, tmp_path):
    """random records (empty SEQ, IUPAC codes, odd lengths, every flag), tiny BGZF blocks"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from bam_build import build_bam
    rnd = random.Random(900 + seed)
    refs = [("chr1", 100000), ("chr2", 5000)]
    recs = []
    for r in range(rnd.randrange(1, 300)):
        L = rnd.choice([0, 1, 2, 3, 24, 25, 26, 149, 150, 151, 1000])
        seq = "".join(rnd.choice("ACGT" if rnd.random() < 0.95 else "=ACMGRSVTWYHKDBN") for _ in range(L))
        qual = None if rnd.random() < 0.1 else "".join(chr(33 + rnd.randrange(41)) for _ in range(L))
        recs.append((f"q{r}", rnd.choice([0, 16, 83, 99, 147, 163, 256 + 99, 1024 + 83, 2048 + 99, 4]), 0, r, 60,
                     [(L, "M")] if L else [], 0, r + 100, 250, seq, qual, [("NM", "C", 1)] if rnd.random() < 0.7 else []))
    bam = tmp_path / "f.bam"
    bam.write_bytes(build_bam("@HD\tVN:1.6\n", refs, recs, block=rnd.choice([61, 500, 4096, 65280])))
    want = b"".join((rec[9] or "*").encode() + b"N" for rec in recs if not rec[1] & 3328)
    assert stream(bam, "-t", str(rnd.choice([1, 2, 6]))) == want


def test_bam_errors(tmp_path):
    data = open(os.path.join(GOLDEN, "testRun_bam", "Child.bam"), "rb").read()
    (tmp_path / "cut.bam").write_bytes(data[:len(data) // 2])
    r = subprocess.run([EXE, str(tmp_path / "cut.bam")], capture_output=True)
    assert r.returncode == 1 and b"truncated" in r.stderr
    import gzip
    (tmp_path / "plain.gz").write_bytes(gzip.compress(b"@r\nACGT\n+\nIIII\n"))
    r = subprocess.run([EXE, str(tmp_path / "plain.gz")], capture_output=True)
    assert r.returncode == 1 and b"BGZF" in r.stderr      # a gzip file that is not BGZF: refused, not mis-parsed


# ---- the paired input of RUFUS.Filter: two FASTQ files (the chain's re-paired mates) or the BAM itself, re-paired in
# process (tools/rg_bam_pairs.hpp); tools/bin/rg_pairs_check writes the pairs and a digest of the batch arrays
PAIRS = os.environ.get("RG_PAIRS_CHECK", os.path.join(BIN, "rg_pairs_check"))   # (override: a sanitizer build)


def stranded_chain(bam, stub, excl=3328):
    p1 = subprocess.run([os.path.join(BIN, "rg_bamview"), "view", "-F", str(excl), str(bam)], capture_output=True, check=True)
    subprocess.run([os.path.join(BIN, "PassThroughSamCheck.stranded"), os.devnull, str(stub)], input=p1.stdout, capture_output=True, check=True)
    return open(f"{stub}.mate1.fastq", "rb").read(), open(f"{stub}.mate2.fastq", "rb").read()


def pairs_check(*args):
    r = subprocess.run([PAIRS, *[str(a) for a in args]], capture_output=True)
    assert r.returncode == 0, r.stderr
    stub = str(args[-1])
    return open(stub + ".mate1.fastq", "rb").read(), open(stub + ".mate2.fastq", "rb").read(), r.stdout


@pytest.mark.parametrize("sample", ["Child", "Mother"])
def test_bam_pairs_equal_the_repaired_fastq_files_of_the_chain(sample, tmp_path):
    bam = os.path.join(GOLDEN, "testRun_bam", sample + ".bam")
    m1, m2 = stranded_chain(bam, tmp_path / "chain")
    assert m1.count(b"\n") >= 4000
    f1, f2, digest_fastq = pairs_check("-b", "300", tmp_path / "chain.mate1.fastq", tmp_path / "chain.mate2.fastq", tmp_path / "viafastq")
    assert (f1, f2) == (m1, m2)
    for args in (("-b", "300"), ("-b", "300", "-t", "5")):
        b1, b2, digest_bam = pairs_check(*args, bam, tmp_path / "direct")
        assert (b1, b2) == (m1, m2)
        assert digest_bam == digest_fastq          # the arrays handed to rg_filter_pairs, batch by batch


@pytest.mark.parametrize("seed", range(4))
def test_bam_pairs_fuzz(seed, tmp_path):
    """mates in any order and distance, reverse strands, IUPAC codes (dropped on the reverse strand), empty SEQ, absent
    QUAL, odd lengths, secondary/duplicate/supplementary records, reads whose mate never comes"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from bam_build import build_bam
    rnd = random.Random(4000 + seed)
    refs = [("chr1", 100000)]
    recs = []
    for p in range(rnd.randrange(1, 200)):
        mates = []
        for m in range(rnd.choice([1, 2, 2, 2, 3])):
            L = rnd.choice([0, 1, 2, 25, 26, 149, 150, 151])
            seq = "".join(rnd.choice("ACGT" if rnd.random() < 0.9 else "=ACMGRSVTWYHKDBN") for _ in range(L))
            qual = None if rnd.random() < 0.1 else "".join(chr(33 + rnd.randrange(41)) for _ in range(L))
            flag = rnd.choice([65, 81, 97, 113, 129, 145, 161, 177, 99, 147, 83, 163, 256 + 99, 1024 + 147, 2048 + 83])
            mates.append((f"pair{p}", flag, 0, p, 60, [(L, "M")] if L else [], 0, p + 200, 300, seq, qual, [("NM", "C", 0)]))
        recs += mates
    rnd.shuffle(recs)
    bam = tmp_path / "f.bam"
    bam.write_bytes(build_bam("@HD\tVN:1.6\n", refs, recs, block=rnd.choice([61, 700, 65280])))
    m1, m2 = stranded_chain(bam, tmp_path / "chain")
    _, _, digest_fastq = pairs_check("-b", "7", tmp_path / "chain.mate1.fastq", tmp_path / "chain.mate2.fastq", tmp_path / "viafastq")
    b1, b2, digest_bam = pairs_check("-b", "7", "-t", rnd.choice([1, 3]), bam, tmp_path / "direct")
    assert (b1, b2) == (m1, m2)
    assert digest_bam == digest_fastq


def test_bam_pairs_with_thousands_of_parked_reads(tmp_path):
    """all first mates, then all second mates in another order: the parked-read table grows several times and
    every deletion shifts its probe sequences"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from bam_build import build_bam
    rnd = random.Random(77)
    n = 6000
    def rec(p, flag):
        L = rnd.choice([50, 100, 151])
        return (f"r{p % 97}_{p}", flag, 0, p, 60, [(L, "M")], 0, p + 200, 300, "".join(rnd.choice("ACGT") for _ in range(L)),
                "".join(chr(33 + rnd.randrange(41)) for _ in range(L)), [("NM", "C", 0)])
    first = [rec(p, rnd.choice([99, 83])) for p in range(n)]
    second = [rec(p, rnd.choice([147, 163])) for p in range(n)]
    rnd.shuffle(second)
    bam = tmp_path / "f.bam"
    bam.write_bytes(build_bam("@HD\tVN:1.6\n", [("chr1", 10 ** 6)], first + second, block=65280))
    m1, m2 = stranded_chain(bam, tmp_path / "chain")
    assert m1.count(b"\n") == 4 * n
    _, _, digest_fastq = pairs_check("-b", "1000", tmp_path / "chain.mate1.fastq", tmp_path / "chain.mate2.fastq", tmp_path / "viafastq")
    b1, b2, digest_bam = pairs_check("-b", "1000", "-t", "4", bam, tmp_path / "direct")
    assert (b1, b2) == (m1, m2) and digest_bam == digest_fastq


# ---- `jellyfish count -t T reads.fastq`: the scan of a regular FASTQ file on several threads (tools/rg_fastq_parallel.hpp)
def par_env(segment):
    return dict(os.environ, RG_PARALLEL_FASTQ_MIN_BYTES="1", RG_PARALLEL_FASTQ_SEGMENT=str(segment))


@pytest.mark.parametrize("seed", range(6))
def test_parallel_fastq_scan_equals_the_serial_reader(seed, tmp_path):
    """segments of 1-5 kB over records of every length: quality lines starting with '@' or '+', N bases, CRLF, a last
    record without a newline or cut short"""
    rnd = random.Random(6000 + seed)
    eol = b"\r\n" if seed == 4 else b"\n"
    out = bytearray()
    for r in range(rnd.randrange(50, 600)):
        L = rnd.choice([0, 1, 24, 25, 100, 150, 151, 2000])
        seq = bytes(rnd.choice(b"ACGTN" if rnd.random() < 0.05 else b"ACGT") for _ in range(L))
        qual = bytes(rnd.choice(b"@+IJF#,5") for _ in range(L))
        out += b"@read%d extra" % r + eol + seq + eol + (b"+" if rnd.random() < 0.7 else b"+read%d" % r) + eol + qual + eol
    data = bytes(out)
    if seed == 1:
        data = data[:-len(eol)]                    # no trailing newline
    if seed == 2:
        data = data[:len(data) - rnd.randrange(1, 150)]    # cut inside the last record
    fq = tmp_path / "a.fq"
    fq.write_bytes(data)
    want = stream(fq)                              # -t 1: the serial reader
    for seg in (1024, 3000, 5000, 10 ** 6):
        assert stream(fq, "-t", "3", env=par_env(seg)) == want
    assert stream(fq, "-t", "8", "-c", "4096", env=par_env(2048)) == want


def test_parallel_fastq_scan_refuses_wrapped_records_and_leaves_fifos_to_the_serial_reader(tmp_path):
    wrapped = tmp_path / "w.fq"
    wrapped.write_bytes(b"@r1\nACGT\nACGT\n+\nIIII\nIIII\n" * 400)
    want = stream(wrapped)                          # the serial reader joins the lines
    assert want == b"ACGTACGTN" * 400
    r = subprocess.run([EXE, "-t", "4", str(wrapped)], capture_output=True, env=par_env(1024))
    assert r.returncode == 1 and b"run with -t 1" in r.stderr
    # a FIFO is not a regular file: serial reader whatever -t says
    plain = tmp_path / "p.fq"
    plain.write_bytes(b"@r\nACGTA\n+\nIIIII\n" * 1000)
    r = subprocess.run(["bash", "-c", f"{EXE} -t 4 <(cat {plain})"], capture_output=True, env=par_env(1024))
    assert r.returncode == 0 and r.stdout == b"ACGTAN" * 1000
    # FASTA keeps the serial reader too
    fa = tmp_path / "a.fa"
    fa.write_bytes(b">s\nACGT\nACGT\n" * 300)
    assert stream(fa, "-t", "4", env=par_env(1024)) == b"ACGTACGTN" * 300
