"""Record-level corruption fuzz of the three in-process BAM consumers (tools/rg_bam_pipeline.hpp and its formatters):
`rg_bamview` (SAM text), the input side of `jellyfish count X.bam` (tools/rg_bam_input.hpp through the rg_input_to_seq
harness) and of `RUFUS.Filter HashList X.bam X.bam` (tools/rg_bam_pairs.hpp through rg_pairs_check). A BAM is user input:
the first records of the bundled Child.bam are damaged BEHIND the compression layer (length fields, names, random bytes,
truncation, garbage between records, header lengths), re-wrapped in valid BGZF members by the spec-level writer of
tests/bam_build.py and fed to AddressSanitizer + UBSan builds of the tools (built into a temporary directory). Every run
has to end with exit code 0 or 1 ("corrupt BAM record …") and no sanitizer report: no out-of-bounds read or write, no
hang. (tests/test_inflate.py does the same one layer below, for the DEFLATE decoder. 500 mutations x 3 tools of this
generator ran clean when the test was written; the suite keeps a sample of 24.)"""
import os
import random
import struct
import subprocess
import zlib
from concurrent.futures import ThreadPoolExecutor

import pytest

from bam_build import bgzf
from conftest import GOLDEN, ROOT

TOOLS = os.path.join(ROOT, "tools")
SAN = ["-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-I" + os.path.join(ROOT, "include")]
LINK = ["-L" + os.path.join(ROOT, "rufus_b200"), "-lrufus_gpu", "-Wl,-rpath," + os.path.join(ROOT, "rufus_b200"),
        "-Wl,-rpath-link,/usr/local/cuda/lib64"]


def _unbgzf(b):
    out, p = b"", 0
    while p < len(b):
        xlen = struct.unpack_from("<H", b, p + 10)[0]
        bsize = struct.unpack_from("<H", b, p + 16)[0] + 1
        out += zlib.decompress(b[p + 12 + xlen:p + bsize - 8], -15)
        p += bsize
    return out


@pytest.fixture(scope="module")
def san_tools(tmp_path_factory):
    if not os.path.exists(os.path.join(ROOT, "rufus_b200", "librufus_gpu.so")):
        pytest.skip("librufus_gpu.so not built (the harnesses link it, none of its functions is called)")
    d = tmp_path_factory.mktemp("san")
    jobs = [("bamview", ["bamview.cpp"], ["-lz", "-pthread"]),
            ("seq", ["micro/bam_seq_check.cpp"], LINK + ["-lz", "-pthread"]),
            ("pairs", ["micro/bam_pairs_check.cpp"], LINK + ["-lz", "-pthread"])]

    def cc(job):
        name, src, libs = job
        subprocess.run(["g++"] + SAN + [os.path.join(TOOLS, s) for s in src] + ["-o", str(d / name)] + libs, check=True,
                       capture_output=True)
    with ThreadPoolExecutor(3) as ex:
        list(ex.map(cc, jobs))
    return d


def _records(raw):
    p = 4
    lt = struct.unpack_from("<i", raw, p)[0]
    p += 4 + lt
    nref = struct.unpack_from("<i", raw, p)[0]
    p += 4
    for _ in range(nref):
        p += 4 + struct.unpack_from("<i", raw, p)[0] + 4
    hdr_end, recs = p, []
    while p < len(raw):
        recs.append(p)
        p += 4 + struct.unpack_from("<i", raw, p)[0]
    return lt, hdr_end, recs


def _mutate(rng, small, lt, hdr_end, recs):
    d = bytearray(small)
    kind = rng.randrange(6)
    if kind == 0:      # a length field of one record: block_size, l_read_name/mapq/bin, n_cigar/flag, l_seq, next_ref
        r, field = rng.choice(recs), rng.choice([0, 12, 16, 20, 24])
        val = rng.choice([0, 1, 2, 31, 32, 33, 255, 65535, 0x7FFFFFFF, 0xFFFFFFFF, rng.randrange(1 << 32), rng.randrange(1 << 12)])
        struct.pack_into("<I", d, r + field, val)
    elif kind == 1:    # random bytes, mostly behind the header
        for _ in range(rng.randrange(1, 20)):
            d[rng.randrange(hdr_end if rng.random() < 0.8 else 0, len(d))] = rng.randrange(256)
    elif kind == 2:    # the stream ends anywhere
        d = d[:rng.randrange(0, len(d))]
    elif kind == 3:    # header lengths: l_text, n_ref, the first l_name
        struct.pack_into("<I", d, rng.choice([4, 8 + lt, 8 + lt + 4]), rng.choice([0, 0xFFFFFFFF, 0x7FFFFFFF, rng.randrange(1 << 20)]))
    elif kind == 4:    # l_read_name 0 / 1 / longer than the record
        d[rng.choice(recs) + 12] = rng.choice([0, 1, 255])
    else:              # garbage between two records
        r = rng.choice(recs)
        d[r:r] = bytes(rng.randrange(256) for _ in range(rng.randrange(1, 64)))
    return kind, bytes(d)


def test_damaged_bam_records_never_crash_the_consumers(san_tools, tmp_path):
    raw = _unbgzf(open(os.path.join(GOLDEN, "testRun_bam", "Child.bam"), "rb").read())
    lt, hdr_end, recs = _records(raw)
    small = raw[:recs[150]]
    recs = recs[:150]
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0", UBSAN_OPTIONS="print_stacktrace=1")
    tools = [("bamview", lambda f: ["view", "-F", "3328", f]), ("seq", lambda f: ["-t", "3", f]),
             ("pairs", lambda f: ["-t", "3", "-n", f, str(tmp_path / "stub")])]
    # the undamaged prefix first: the sanitizer builds agree with the shipped tools on good input
    good = str(tmp_path / "good.bam")
    open(good, "wb").write(bgzf(small, 4096))
    want = subprocess.run([os.path.join(TOOLS, "bin", "rg_bamview"), "view", "-F", "3328", good], capture_output=True, check=True).stdout
    got = subprocess.run([str(san_tools / "bamview"), "view", "-F", "3328", good], capture_output=True, env=env)
    assert got.returncode == 0 and got.stdout == want and want.count(b"\n") > 100
    rng, bad, kinds = random.Random(20261017), [], set()
    for it in range(24):
        kind, data = _mutate(rng, small, lt, hdr_end, recs)
        kinds.add(kind)
        f = str(tmp_path / "in.bam")
        open(f, "wb").write(bgzf(data, rng.choice([300, 997, 4096, 65280])))
        for name, args in tools:
            r = subprocess.run([str(san_tools / name)] + args(f), capture_output=True, env=env, timeout=120)
            if r.returncode not in (0, 1) or b"Sanitizer" in r.stderr or b"runtime error" in r.stderr:
                bad.append((it, kind, name, r.returncode, r.stderr[-800:].decode(errors="replace")))
    assert not bad, bad
    assert len(kinds) >= 5
