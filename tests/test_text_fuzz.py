"""Damaged TEXT inputs (FASTQ, FASTA, SAM) never crash the host tools that parse user data: the input side of `jellyfish
count` (serial reader and the threaded scan of tools/rg_fastq_parallel.hpp, through rg_input_to_seq), the pair reader of
`RUFUS.Filter` (tools/rg_fastq_batch.hpp through rg_pairs_check and rg_lockstep_check), `PassThroughSamCheck(.stranded)`
and `rg_fastq_to_sam`. AddressSanitizer + UBSan builds in a temporary directory; pieces of the bundled trio with random
bytes, truncation, missing newlines / tabs, duplicated slices, 200 kB lines, empty and CRLF lines. Exit code 0 or 1 and
no sanitizer report. (400 mutations x 8 runs of this generator were clean when the test was written; the suite keeps 14.
The BAM layer is tests/test_bam_fuzz.py, the DEFLATE layer tests/test_inflate.py.)"""
import gzip
import os
import random
import subprocess
from concurrent.futures import ThreadPoolExecutor

import pytest

from conftest import GOLDEN, ROOT

TOOLS = os.path.join(ROOT, "tools")
SAN = ["-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-I" + os.path.join(ROOT, "include")]
LINK = ["-L" + os.path.join(ROOT, "rufus_b200"), "-lrufus_gpu", "-Wl,-rpath," + os.path.join(ROOT, "rufus_b200"),
        "-Wl,-rpath-link,/usr/local/cuda/lib64"]


@pytest.fixture(scope="module")
def san(tmp_path_factory):
    if not os.path.exists(os.path.join(ROOT, "rufus_b200", "librufus_gpu.so")):
        pytest.skip("librufus_gpu.so not built (the harnesses link it, none of its functions is called)")
    d = tmp_path_factory.mktemp("san_text")
    jobs = [("seq", ["micro/bam_seq_check.cpp"], [], LINK + ["-lz", "-pthread"]),
            ("pairs", ["micro/bam_pairs_check.cpp"], [], LINK + ["-lz", "-pthread"]),
            ("lock", ["lockstep_check.cpp"], [], LINK),
            ("pts", ["passthrough_sam.cpp"], [], LINK),
            ("ptss", ["passthrough_sam.cpp"], ["-DRG_STRANDED"], LINK),
            ("f2s", ["fastq_to_sam.cpp"], [], [])]

    def cc(job):
        name, src, defs, libs = job
        subprocess.run(["g++"] + SAN + defs + [os.path.join(TOOLS, s) for s in src] + ["-o", str(d / name)] + libs, check=True,
                       capture_output=True)
    with ThreadPoolExecutor(6) as ex:
        list(ex.map(cc, jobs))
    return d


def _mutate(rng, b):
    d = bytearray(b)
    kind = rng.randrange(6)
    if kind == 0:
        for _ in range(rng.randrange(1, 30)):
            d[rng.randrange(len(d))] = rng.randrange(256)
    elif kind == 1:
        d = d[:rng.randrange(len(d))]
    elif kind == 2:
        for _ in range(rng.randrange(1, 10)):
            i = d.find(rng.choice([b"\n", b"\t"]), rng.randrange(len(d)))
            if i >= 0:
                del d[i]
    elif kind == 3:
        a = rng.randrange(len(d))
        d[a:a] = d[a:min(len(d), a + rng.randrange(1, 3000))]
    elif kind == 4:
        a = rng.randrange(len(d))
        d[a:a] = bytes([rng.choice(b"ACGTN@+>\t ")]) * rng.randrange(1, 200000)
    else:
        a = rng.randrange(len(d))
        d[a:a] = rng.choice([b"\n\n\n", b"\r\n", b"\0\0", b"@\n", b">\n", b"+\n"])
    return bytes(d)


def test_damaged_text_never_crashes_the_parsers(san, tmp_path):
    fq1 = gzip.open(os.path.join(GOLDEN, "testRun", "Child.mate1.fastq.gz")).read()[:60000]
    fq2 = gzip.open(os.path.join(GOLDEN, "testRun", "Child.mate2.fastq.gz")).read()[:60000]
    sam = subprocess.run([os.path.join(TOOLS, "bin", "rg_bamview"), "view", "-F", "3328", os.path.join(GOLDEN, "testRun_bam", "Child.bam")],
                         capture_output=True, check=True).stdout[:80000]
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0", RG_PARALLEL_FASTQ_MIN_BYTES="1000", RG_PARALLEL_FASTQ_SEGMENT="7000")
    t, bad = str(tmp_path), []

    def run(tag, cmd, stdin=None):
        r = subprocess.run([str(san / cmd[0])] + cmd[1:], input=stdin, capture_output=True, env=env, timeout=120)
        if r.returncode not in (0, 1) or b"Sanitizer" in r.stderr or b"runtime error" in r.stderr:
            bad.append((tag, cmd, r.returncode, r.stderr[-800:].decode(errors="replace")))

    # undamaged first: the sanitizer builds do what the shipped ones do
    open(t + "/a.fq", "wb").write(fq1[:fq1.rfind(b"\n@") + 1])
    want = subprocess.run([os.path.join(TOOLS, "bin", "rg_input_to_seq"), t + "/a.fq"], capture_output=True, check=True).stdout
    got = subprocess.run([str(san / "seq"), "-t", "3", t + "/a.fq"], capture_output=True, env=env)
    assert got.returncode == 0 and got.stdout == want and want.count(b"N") > 100
    rng = random.Random(20261017)
    for it in range(14):
        a, b, s = _mutate(rng, fq1), _mutate(rng, fq2), _mutate(rng, sam)
        open(t + "/a.fq", "wb").write(a)
        open(t + "/b.fq", "wb").write(b)
        open(t + "/a.fa", "wb").write(b"".join(b">" + l for l in a.split(b"@")[1:3]) + a[:3000].replace(b"@", b">"))
        run(it, ["seq", "-t", "1", t + "/a.fq"])
        run(it, ["seq", "-t", "1", t + "/a.fa"])
        run(it, ["seq", "-t", "3", t + "/a.fq"])                       # threaded scan, 7 kB segments
        run(it, ["pairs", "-n", "-b", "50", t + "/a.fq", t + "/b.fq", t + "/stub"])
        run(it, ["lock", t + "/a.fq", t + "/b.fq", "37"])
        run(it, ["pts", t + "/x.chr"], stdin=s)
        run(it, ["ptss", t + "/x.chr", t + "/st"], stdin=s)
        run(it, ["f2s", t + "/a.fq"])
    assert not bad, bad
