"""CPU-side parity of the host-only drop-ins (tools/bin/PassThroughSamCheck, .stranded; SURVEY §8a
P1/P2) against the committed outputs of the reference binaries, and against the binaries
themselves where oracle/_ref is present."""
import gzip
import hashlib
import json
import os
import shutil
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


# ---- BAM ingest: tools/bin/rg_bamview (the `samtools view -F 3328 X.bam` of a .generator script,
# runRufus.sh:599) against the SAM-spec restatement oracle/bam_oracle.py and the committed fixtures
import sys
sys.path.insert(0, os.path.join(ROOT, "oracle"))
BAMG = os.path.join(GOLDEN, "bam")
REF_TESTRUN = "/root/reference/testRun"


def _bamview(*args):
    exe = os.path.join(BIN, "rg_bamview")
    if not os.path.exists(exe):
        pytest.fail("tools/bin/rg_bamview is missing: run `make tools`")
    return subprocess.run([exe, *args], capture_output=True)


def test_bamview_synthetic_matches_hand_built_sam():
    """every tag type, arrays, unmapped records, absent SEQ/QUAL/CIGAR, records straddling BGZF blocks"""
    import bam_oracle
    want = open(os.path.join(BAMG, "synthetic.sam"), "rb").read()
    g = json.load(open(os.path.join(BAMG, "golden.json")))
    bam = os.path.join(BAMG, "synthetic.bam")
    hdr, lines = bam_oracle.view(bam)
    assert ("\n".join(lines) + "\n").encode() == want and hdr == g["synthetic"]["header"]
    for threads in ("1", "3", "16"):
        r = _bamview("view", "-@", threads, bam)
        assert r.returncode == 0 and r.stdout == want
    r = _bamview("-h", bam)
    assert r.stdout == g["synthetic"]["header"].encode() + want
    # flag filters: -F drops duplicates (1024), -f keeps only reverse-strand records
    keep = [l for l in want.decode().splitlines() if not int(l.split("\t")[1]) & 1024]
    assert _bamview("-F", "1024", bam).stdout.decode().splitlines() == keep
    assert _bamview("-F1024", bam).stdout.decode().splitlines() == keep
    rev = [l for l in want.decode().splitlines() if int(l.split("\t")[1]) & 16]
    assert _bamview("-f", "16", bam).stdout.decode().splitlines() == rev
    # stdin input
    r = subprocess.run([os.path.join(BIN, "rg_bamview"), "-"], input=open(bam, "rb").read(), capture_output=True)
    assert r.returncode == 0 and r.stdout == want
    # zlib's inflate instead of tools/rg_inflate.hpp (the fallback path for a member the fast decoder rejects)
    r = subprocess.run([os.path.join(BIN, "rg_bamview"), bam], capture_output=True, env=dict(os.environ, RG_BAMVIEW_ZLIB="1"))
    assert r.returncode == 0 and r.stdout == want


def test_bamview_rejects_bad_input(tmp_path):
    bad = tmp_path / "bad.bam"
    bad.write_bytes(b"this is not a BGZF file at all")
    assert _bamview(str(bad)).returncode == 1
    data = open(os.path.join(BAMG, "synthetic.bam"), "rb").read()
    cut = tmp_path / "cut.bam"
    cut.write_bytes(data[:len(data) // 2])
    assert _bamview(str(cut)).returncode == 1                     # truncated block
    flip = bytearray(data)
    flip[len(flip) // 2] ^= 0x55
    (tmp_path / "flip.bam").write_bytes(bytes(flip))
    assert _bamview(str(tmp_path / "flip.bam")).returncode == 1   # inflate or crc failure
    assert _bamview(os.path.join(BAMG, "synthetic.bam"), "chr1:1-100").returncode == 2   # region: unsupported
    assert _bamview(str(tmp_path / "missing.bam")).returncode == 1


def _pairs_md5(m1, m2):
    def load(p):
        l = open(p).read().splitlines()
        return {l[i][1:]: (l[i + 1], l[i + 3]) for i in range(0, len(l), 4)}
    a, b = load(m1), load(m2)
    rows = sorted(n + "\t" + "\t".join("\t".join(x) for x in sorted((a[n], b[n]))) for n in a)
    return md5("\n".join(rows).encode()), len(rows)


@pytest.mark.parametrize("sample", ["Father", "Child", "Mother"])
def test_bam_ingest_chain_reproduces_bundled_fastq(sample, tmp_path):
    """BAM -> rg_bamview -> PassThroughSamCheck.stranded gives exactly the reads of the bundled
    testRun/X.mate{1,2}.fastq (names, original-strand sequences, qualities); the SAM text equals the
    restatement's."""
    import bam_oracle
    g = json.load(open(os.path.join(BAMG, "golden.json")))["bundled"][sample]
    bam = os.path.join(GOLDEN, "testRun_bam", sample + ".bam")     # the reference's bundled BAM files (data fixtures)
    assert md5(open(bam, "rb").read()) == g["bam_md5"]
    for tag, excl in (("F3328", 3328), ("all", 0)):
        r = _bamview("view", "-F", str(excl), bam)
        assert r.returncode == 0
        assert r.stdout.count(b"\n") == g[tag]["records"] and md5(r.stdout) == g[tag]["sam_md5"]
        _, lines = bam_oracle.view(bam, excl)
        assert ("\n".join(lines) + "\n").encode() == r.stdout
    # all primary records (the bundled FASTQ files keep duplicates), re-paired by name
    sam = _bamview("-F", "2304", bam).stdout
    subprocess.run([os.path.join(BIN, "PassThroughSamCheck.stranded"), str(tmp_path / "chr"), str(tmp_path / "o")],
                   input=sam, capture_output=True, check=True)
    got = _pairs_md5(tmp_path / "o.mate1.fastq", tmp_path / "o.mate2.fastq")
    assert got == (g["fastq_pairs_md5"], g["fastq_pairs"])


# ---- FIFO pipelines of runRufus.sh:963-967: one producer writes both mate files in lock-step, the
# consumer reads them in lock-step; neither side may wait for a whole block of ONE file
def _fifo_pipeline(tmp_path, writer, reader_cmd):
    sam = _bamview("view", "-F", "3328", os.path.join(GOLDEN, "testRun_bam", "Child.bam")).stdout
    (tmp_path / "in.sam").write_bytes(sam)
    cmd = (f"mkfifo t.mate1.fastq t.mate2.fastq; {writer} c.chr t < in.sam > t.log & "
           f"timeout 60 {reader_cmd}; rc=$?; wait; exit $rc")
    return subprocess.run(cmd, shell=True, cwd=tmp_path, capture_output=True, executable="/bin/bash")


def _expected_pairs(tmp_path):
    subprocess.run(f"{BIN}/PassThroughSamCheck.stranded c2.chr plain < in.sam > plain.log", shell=True, cwd=tmp_path, check=True)
    n1 = open(tmp_path / "plain.mate1.fastq").read().splitlines()[0::4]
    n2 = open(tmp_path / "plain.mate2.fastq").read().splitlines()[0::4]
    assert len(n1) == json.load(open(os.path.join(GOLDEN, "testRun_bam", "golden.json")))["repaired_pairs"]
    return "".join(f"{a}\t{b}\n" for a, b in zip(n1, n2)).encode()


@pytest.mark.parametrize("batch", ["7", "100000"])
def test_fifo_lockstep_our_writer_our_reader(tmp_path, batch):
    r = _fifo_pipeline(tmp_path, f"{BIN}/PassThroughSamCheck.stranded",
                       f"{BIN}/rg_lockstep_check t.mate1.fastq t.mate2.fastq {batch}")
    assert r.returncode == 0, r.stderr[-500:]
    assert r.stdout == _expected_pairs(tmp_path)


@pytest.mark.skipif(not os.path.exists(os.path.join(REFBIN, "PassThroughSamCheck.stranded")),
                    reason="reference binaries (oracle/_ref) not present")
def test_fifo_lockstep_with_reference_tools(tmp_path):
    # the reference's line-flushing writer feeding our reader ...
    r = _fifo_pipeline(tmp_path, f"{REFBIN}/PassThroughSamCheck.stranded",
                       f"{BIN}/rg_lockstep_check t.mate1.fastq t.mate2.fastq 1000")
    assert r.returncode == 0, r.stderr[-500:]
    assert r.stdout == _expected_pairs(tmp_path)
    # ... and our block-buffering writer feeding the reference's RUFUS.Filter (a CPU program)
    g = json.load(open(os.path.join(GOLDEN, "testRun_bam", "golden.json")))
    (tmp_path / "HashList").write_text("\n".join(g["hashlist_sorted"]) + "\n")
    subprocess.run("rm -f t.mate1.fastq t.mate2.fastq", shell=True, cwd=tmp_path)
    r = _fifo_pipeline(tmp_path, f"{BIN}/PassThroughSamCheck.stranded",
                       f"{REFBIN}/RUFUS.Filter HashList t.mate1.fastq t.mate2.fastq out {g['k']} {g['minq']} {g['thresh']} 2 > f.log")
    assert r.returncode == 0, r.stderr[-500:]
    o1 = open(tmp_path / "out.Mutations.Mate1.fastq").read().splitlines()
    assert sorted(o1[0::4]) == g["pulled_mate1_names_sorted"]


# ---- differential fuzz of the two SAM -> FASTQ drop-ins against the reference binaries (both are CPU
# programs): random well-formed records with repeated names, reverse flags, IUPAC / lower-case bases,
# unpaired reads, several chromosomes
@pytest.mark.skipif(not os.path.exists(os.path.join(REFBIN, "PassThroughSamCheck.stranded")),
                    reason="reference binaries (oracle/_ref) not present")
@pytest.mark.parametrize("seed", range(6))
def test_passthrough_differential_fuzz(seed, tmp_path):
    import random
    rnd = random.Random(1000 + seed)
    lines = []
    names = [f"read{rnd.randrange(60)}:{rnd.randrange(3)}" for _ in range(300)]
    for nm in names:
        L = rnd.choice([0, 1, 5, 36, 100, 151]) if rnd.random() < 0.3 else 100
        seq = "".join(rnd.choice("ACGTNacgtRYKM" if rnd.random() < 0.05 else "ACGT") for _ in range(L)) or "*"
        qual = "".join(chr(33 + rnd.randrange(42)) for _ in range(len(seq)))
        flag = rnd.choice([0, 16, 65, 81, 97, 113, 129, 145, 147, 163, 99, 83, 1024 + 99, 2048 + 16, 4, 77, 141])
        chrom = rnd.choice(["1", "1", "2", "X", "GL000207.1", "*"])
        tags = "\t".join(rnd.sample(["NM:i:0", "MD:Z:100", "AS:i:98", "RG:Z:grp", "XS:i:0"], rnd.randrange(1, 4)))
        lines.append("\t".join([nm, str(flag), chrom, str(rnd.randrange(1, 10 ** 6)), "60", f"{len(seq)}M", "=",
                                str(rnd.randrange(1, 10 ** 6)), str(rnd.randrange(-500, 500)), seq, qual, tags]))
    sam = ("\n".join(lines) + "\n").encode()
    ours = subprocess.run([os.path.join(BIN, "PassThroughSamCheck"), str(tmp_path / "a.chr")], input=sam, capture_output=True)
    ref = subprocess.run([os.path.join(REFBIN, "PassThroughSamCheck"), str(tmp_path / "b.chr")], input=sam, capture_output=True)
    assert ref.returncode == 0 and ours.returncode == 0
    assert ours.stdout == ref.stdout
    assert open(tmp_path / "a.chr").read() == open(tmp_path / "b.chr").read()
    subprocess.run([os.path.join(BIN, "PassThroughSamCheck.stranded"), str(tmp_path / "c.chr"), str(tmp_path / "o")], input=sam,
                   capture_output=True, check=True)
    subprocess.run([os.path.join(REFBIN, "PassThroughSamCheck.stranded"), str(tmp_path / "d.chr"), str(tmp_path / "r")], input=sam,
                   capture_output=True, check=True)
    for m in (1, 2):
        assert open(tmp_path / f"o.mate{m}.fastq", "rb").read() == open(tmp_path / f"r.mate{m}.fastq", "rb").read()
    assert open(tmp_path / "c.chr").read() == open(tmp_path / "d.chr").read()


@pytest.mark.parametrize("seed", range(8))
def test_bamview_fuzz_against_spec_writer_and_restatement(seed, tmp_path):
    """random records (every tag type, arrays, unmapped, absent SEQ/QUAL/CIGAR, long names) written by
    tests/bam_build.py with random BGZF block sizes: rg_bamview == the text the writer states ==
    oracle/bam_oracle.py, with and without flag filters"""
    import random
    import bam_oracle
    from bam_build import build_bam, expected_sam
    rnd = random.Random(77 + seed)
    refs = [(f"chr{i}", 1000 * (i + 1)) for i in range(rnd.randrange(1, 5))]
    text = "@HD\tVN:1.6\n" + "".join(f"@SQ\tSN:{n}\tLN:{l}\n" for n, l in refs)

    def rand_tag(i):
        t = rnd.choice("AcCsSiIfZHB")
        tag = "X" + "abcdefghijklmnop"[i]
        if t == "A":
            return (tag, t, rnd.choice("QZa!~"))
        if t == "Z":
            return (tag, t, "".join(rnd.choice("ab c:,;|-_09") for _ in range(rnd.randrange(0, 20))))
        if t == "H":
            return (tag, t, "".join(rnd.choice("0123456789ABCDEF") for _ in range(2 * rnd.randrange(0, 6))))
        lim = {"c": (-128, 127), "C": (0, 255), "s": (-32768, 32767), "S": (0, 65535), "i": (-2 ** 31, 2 ** 31 - 1),
               "I": (0, 2 ** 32 - 1)}
        if t == "f":
            return (tag, t, rnd.choice([0.0, 1.5, -2.25, 1e-3, 12345.678, 3e10, -1e-7]))
        if t == "B":
            sub = rnd.choice("cCsSiIf")
            n = rnd.randrange(0, 6)
            vals = [rnd.choice([0.5, -1.0, 2.75, 1e6]) if sub == "f" else rnd.randint(*lim[sub]) for _ in range(n)]
            return (tag, t, (sub, vals))
        return (tag, t, rnd.randint(*lim[t]))

    recs = []
    for r in range(rnd.randrange(1, 120)):
        L = rnd.choice([0, 1, 2, 7, 50, 151, 400])
        seq = "".join(rnd.choice("ACGTN" if rnd.random() < 0.9 else "=ACMGRSVTWYHKDBN") for _ in range(L))
        qual = None if rnd.random() < 0.15 else "".join(chr(33 + rnd.randrange(60)) for _ in range(L))
        ref = rnd.randrange(-1, len(refs))
        cig = [] if ref < 0 or rnd.random() < 0.1 else [(rnd.randrange(1, 300), rnd.choice("MIDNSHP=X")) for _ in range(rnd.randrange(1, 6))]
        name = "".join(rnd.choice("abcXYZ0189:/_.") for _ in range(rnd.randrange(1, 200)))
        recs.append((name, rnd.choice([0, 4, 16, 65, 77, 83, 99, 141, 147, 163, 256 + 16, 1024 + 99, 2048 + 83, 3328 + 1]),
                     ref, rnd.randrange(-1, 10 ** 6) if ref >= 0 else -1, rnd.randrange(0, 255), cig,
                     rnd.randrange(-1, len(refs)), rnd.randrange(-1, 10 ** 6), rnd.randrange(-10 ** 5, 10 ** 5), seq, qual,
                     [rand_tag(i) for i in range(rnd.randrange(0, 8))]))
    bam = tmp_path / "f.bam"
    bam.write_bytes(build_bam(text, refs, recs, block=rnd.choice([61, 300, 997, 4096, 65280])))
    want = expected_sam(refs, recs)
    _, lines = bam_oracle.view(str(bam))
    assert lines == want
    for excl, req in ((0, 0), (3328, 0), (4, 0), (0, 16), (1024, 1)):
        keep = [l for l in want if not int(l.split("\t")[1]) & excl and (int(l.split("\t")[1]) & req) == req]
        r = _bamview("-F", str(excl), "-f", str(req), "-@", str(rnd.choice([1, 2, 5])), str(bam))
        assert r.returncode == 0, r.stderr
        assert r.stdout.decode().splitlines() == keep


def test_bamview_header_and_records_larger_than_a_batch(tmp_path):
    """a BAM header of ~0.5 MB and reads of 70 k / 150 k bases in BGZF blocks of 61 bytes: the header and the long
    records straddle several pipeline batches (4096 members each), so they are stitched from the carried bytes of
    the batches before (the path beside the 64 KiB headroom)"""
    import random
    import bam_oracle
    from bam_build import build_bam, expected_sam
    rnd = random.Random(3)
    refs = [(f"contig_{i:05d}_with_a_long_name", 1000 + i) for i in range(6000)]
    text = "@HD\tVN:1.6\n" + "".join(f"@SQ\tSN:{n}\tLN:{l}\n" for n, l in refs)
    recs = []
    for r, L in enumerate([150, 70_000, 151, 150_000, 1, 149]):
        seq = "".join(rnd.choice("ACGT") for _ in range(L))
        qual = "".join(chr(33 + rnd.randrange(41)) for _ in range(L))
        recs.append((f"read{r}", 99 if r % 2 == 0 else 147, r * 7, 100 + r, 60, [(L, "M")], r * 7, 400 + r, 300, seq, qual,
                     [("NM", "C", r), ("RG", "Z", "grp")]))
    bam = tmp_path / "big.bam"
    bam.write_bytes(build_bam(text, refs, recs, block=61))
    want = expected_sam(refs, recs)
    for threads in ("1", "4"):
        r = _bamview("-@", threads, str(bam))
        assert r.returncode == 0, r.stderr
        assert r.stdout.decode().splitlines() == want
    r = _bamview("-h", "-f", "16", str(bam))
    assert r.stdout.decode() == text + "".join(l + "\n" for l in want if int(l.split("\t")[1]) & 16)
    hdr, lines = bam_oracle.view(str(bam))
    assert lines == want and hdr == text


# ---- FASTQ -> SAM text: tools/bin/rg_fastq_to_sam, the drop-in for `perl scripts/FastqToSam.pl` (the generator line
# runRufus.sh:159-161 writes for a FASTQ input)
F2S = os.path.join(BIN, "rg_fastq_to_sam")
F2S_PERL = "/root/reference/scripts/FastqToSam.pl"


def test_fastq_to_sam_matches_the_reference_script_golden():
    if not os.path.exists(F2S):
        pytest.fail("tools/bin/rg_fastq_to_sam is missing: run `make tools`")
    g = os.path.join(GOLDEN, "fastq_to_sam")
    r = subprocess.run([F2S, os.path.join(g, "edge.fastq")], capture_output=True)
    assert r.returncode == 0 and r.stdout == open(os.path.join(g, "edge.sam"), "rb").read()
    assert subprocess.run([F2S, os.path.join(g, "missing.fastq")], capture_output=True).returncode != 0
    # through a FIFO, as `perl FastqToSam.pl <(cat X.fastq)` is used, and on into PassThroughSamCheck:
    # FASTQ -> SAM text -> FASTQ gives back name (first word), sequence and quality of every record
    fq = gzip.open(os.path.join(GOLDEN, "testRun", "Child.mate1.fastq.gz")).read()
    r = subprocess.run(["bash", "-c", f"{F2S} <(cat) | {os.path.join(BIN, 'PassThroughSamCheck')} /dev/null"], input=fq, capture_output=True)
    assert r.returncode == 0
    lines, back = fq.split(b"\n"), r.stdout.split(b"\n")
    assert len(back) == len(lines)
    assert back[1::4] == lines[1::4] and back[3::4] == lines[3::4]
    assert [l.split()[0] for l in lines[0:-1:4]] == back[0:-1:4]


@pytest.mark.skipif(not (os.path.exists(F2S_PERL) and shutil.which("perl")), reason="needs perl and the reference script (build container)")
@pytest.mark.parametrize("seed", range(4))
def test_fastq_to_sam_differential_fuzz_against_the_perl_script(seed, tmp_path):
    import random
    rnd = random.Random(500 + seed)
    out = bytearray()
    for _ in range(rnd.randrange(1, 400)):
        name = b"".join(rnd.choice([b"@", b"r", b"1", b"/", b" ", b"\t", b"\x0b", b"\x0c", b":", b"x"]) for _ in range(rnd.randrange(0, 12)))
        eol = rnd.choice([b"\n", b"\n", b"\r\n"])
        L = rnd.choice([0, 1, 5, 150, 3000])
        seq = bytes(rnd.choice(b"ACGTNacgt") for _ in range(L))
        qual = bytes(33 + rnd.randrange(60) for _ in range(L))
        out += name + eol + seq + eol + b"+" + (name if rnd.random() < 0.3 else b"") + eol + qual + eol
    if rnd.random() < 0.5:
        out = out[:rnd.randrange(len(out) + 1)]          # cut anywhere: short last record, no trailing newline
    fq = tmp_path / "f.fastq"
    fq.write_bytes(bytes(out))
    want = subprocess.run(["perl", F2S_PERL, str(fq)], capture_output=True).stdout
    got = subprocess.run([F2S, str(fq)], capture_output=True)
    assert got.returncode == 0 and got.stdout == want
