"""GPU parity of the drop-in executables (tools/bin: jellyfish, RUFUS.Filter, RUFUS.Filter.single):
the reference's own stage commands (tests/golden/make_golden.py replays them with the
reference binaries) are run with OUR binaries on the same inputs and compared with the committed
goldens; where the reference binaries travelled with the repo (oracle/_ref), the files we write are
also fed to the REAL jellyfish (and vice versa) to pin the .Jhash byte format."""
import gzip
import hashlib
import os
import shutil
import subprocess

import pytest

from cases import golden
from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "tools", "bin")
JF = os.path.join(BIN, "jellyfish")
REFJF = os.path.join(ROOT, "oracle", "_ref", "jellyfish")
SAMPLES = {"testRun": ("Child", "Mother", "Father"), "edge": ("Subj", "CtlA", "CtlB")}


def sh(cmd, cwd, **kw):
    return subprocess.run(cmd, shell=True, check=True, cwd=cwd, capture_output=True, executable="/bin/bash", **kw)


def md5(b):
    return hashlib.md5(b).hexdigest()


def records(path):
    raw = open(path, "rb").read()
    return raw[9 + int(raw[:9]):]


@pytest.fixture(scope="module", params=["testRun", "edge"])
def run(request, tmp_path_factory):
    if not os.path.exists(JF):
        pytest.fail("tools/bin/jellyfish is missing: run `make tools`")
    case = request.param
    g = golden(case)
    work = str(tmp_path_factory.mktemp(case))
    fq = {}
    for s in SAMPLES[case]:
        for m in (1, 2):
            dst = os.path.join(work, f"{s}.mate{m}.fastq")
            src = os.path.join(GOLDEN, case, f"{s}.mate{m}.fastq")
            if os.path.exists(src + ".gz"):
                open(dst, "wb").write(gzip.open(src + ".gz").read())
            else:
                shutil.copy(src, dst)
        fq[s] = (f"{s}.mate1.fastq", f"{s}.mate2.fastq")
    k = g["k"]
    for i, s in enumerate(SAMPLES[case]):
        lo = g["L"] if i == 0 else g["ctlL"]
        sh(f"{JF} count --disk -m {k} -L {lo} -s 16M -t 4 -o {s}.Jhash -C {fq[s][0]} {fq[s][1]}", work)
        sh(f"{JF} histo -f -o {s}.histo {s}.Jhash", work)
    return case, g, work, fq


def test_count_histo_dump_match_goldens(run):
    case, g, work, fq = run
    for s in SAMPLES[case]:
        gs = g["samples"][s]
        assert md5(open(os.path.join(work, s + ".histo"), "rb").read()) == gs["histo_md5"]
        # (that one came from the sidecar `count` leaves next to the .Jhash; the loading path must agree, and a
        # sidecar that describes another file must be ignored)
        assert os.path.exists(os.path.join(work, s + ".Jhash.rg_histo"))
        h2 = sh(f"RG_NO_HISTO_SIDECAR=1 {JF} histo -f {s}.Jhash", work).stdout
        assert md5(h2) == gs["histo_md5"]
        sh(f"cp {s}.Jhash {s}.copy.Jhash && cp {SAMPLES[case][0]}.Jhash.rg_histo {s}.copy.Jhash.rg_histo && sleep 0.01 && touch {s}.copy.Jhash", work)
        assert md5(sh(f"{JF} histo -f {s}.copy.Jhash", work).stdout) == gs["histo_md5"]
        dump = sh(f"{JF} dump -c {s}.Jhash | LC_ALL=C sort", work).stdout
        assert dump.count(b"\n") == gs["n_distinct"]
        assert md5(dump) == gs["dump_sorted_md5"]
        assert dump == gzip.open(os.path.join(GOLDEN, case, s + ".dump.sorted.gz")).read()


def test_merge_query_hashlist_match_goldens(run):
    case, g, work, fq = run
    names = SAMPLES[case]
    merged = sh(f"{JF} merge " + " ".join(s + ".Jhash" for s in names) + " > merge.out; cat merge.out", work).stdout
    assert sorted(merged.decode().splitlines()) == g["merge_sorted"]
    q = sh(f"{JF} query -s <(awk '{{print \">\"$1\"\\n\"$1}}' merge.out) {names[0]}.Jhash", work).stdout.decode()
    hl = [l for l in q.splitlines() if g["min_cov"] <= int(l.split()[1]) <= g["max_cov"]]
    assert sorted(hl) == g["hashlist_sorted"]
    open(os.path.join(work, "HashList"), "w").write("\n".join(hl) + "\n")


def test_exclude_file_in_merge(run):
    """`-e` exclude files are extra `jellyfish merge` inputs (runRufus.sh:925); a record seen once excludes"""
    case, g, work, fq = run
    if "exclude" not in g:
        pytest.skip("no exclude golden for this case")
    names = SAMPLES[case]
    with open(os.path.join(work, "excl.fa"), "w") as f:
        for i, r in enumerate(g["exclude"]["reads"]):
            f.write(f">x{i}\n{r}\n")
    sh(f"{JF} count -m {g['k']} -s 16M -t 2 -C -o excl.Jhash excl.fa", work)
    merged = sh(f"{JF} merge " + " ".join(s + ".Jhash" for s in names) + " excl.Jhash > merge_e.out; cat merge_e.out", work).stdout
    assert sorted(merged.decode().splitlines()) == g["exclude"]["merge_sorted"]
    q = sh(f"{JF} query -s <(awk '{{print \">\"$1\"\\n\"$1}}' merge_e.out) {names[0]}.Jhash", work).stdout.decode()
    hl = [l for l in q.splitlines() if g["min_cov"] <= int(l.split()[1]) <= g["max_cov"]]
    assert sorted(hl) == g["exclude"]["hashlist_sorted"]
    if os.path.exists(REFJF):
        # an exclude file written by the REAL jellyfish is read the same way
        sh(f"{REFJF} count -m {g['k']} -s 16M -t 2 -C -o excl_ref.Jhash excl.fa", work)
        merged2 = sh(f"{JF} merge " + " ".join(s + ".Jhash" for s in names) + " excl_ref.Jhash", work).stdout
        assert sorted(merged2.decode().splitlines()) == g["exclude"]["merge_sorted"]


def test_read_pull_matches_goldens(run):
    case, g, work, fq = run
    if not os.path.exists(os.path.join(work, "HashList")):
        open(os.path.join(work, "HashList"), "w").write("\n".join(g["hashlist_sorted"]) + "\n")
    m1, m2 = fq[SAMPLES[case][0]]
    sh(f"{BIN}/RUFUS.Filter HashList {m1} {m2} out {g['k']} {g['minq']} {g['thresh']} 4", work)
    o1 = open(os.path.join(work, "out.Mutations.Mate1.fastq")).read().splitlines()
    o2 = open(os.path.join(work, "out.Mutations.Mate2.fastq")).read().splitlines()
    assert sorted(o1[0::4]) == g["pulled_mate1_names_sorted"]
    assert sorted(o2[0::4]) == g["pulled_mate2_names_sorted"]
    # the records are the input records, verbatim
    src = open(os.path.join(work, m1)).read().splitlines()
    have = {tuple(src[i:i + 4]) for i in range(0, len(src), 4)}
    assert all(tuple(o1[i:i + 4]) in have for i in range(0, len(o1), 4))
    sh(f"{BIN}/RUFUS.Filter.single HashList {m1} se {g['k']} {g['minq']} {g['thresh']} 4", work)
    se = open(os.path.join(work, "se.Mutations.fastq")).read().splitlines()
    assert sorted(se[0::4]) == g["single_names_sorted"]
    # mate files may be FIFOs (runRufus.sh:971-981): lock-step reading must not deadlock
    sh(f"mkfifo f1 f2; cat {m1} > f1 & cat {m2} > f2 & timeout 120 {BIN}/RUFUS.Filter HashList f1 f2 fifo "
       f"{g['k']} {g['minq']} {g['thresh']} 4; wait", work)
    assert open(os.path.join(work, "fifo.Mutations.Mate1.fastq")).read().splitlines() == o1


@pytest.mark.skipif(not os.path.exists(REFJF), reason="reference binaries (oracle/_ref) not present")
def test_jhash_files_interoperate_with_real_jellyfish(run):
    case, g, work, fq = run
    names = SAMPLES[case]
    for i, s in enumerate(names):
        lo = g["L"] if i == 0 else g["ctlL"]
        # the real tool reads our file: same records in the same order as our own dump
        assert sh(f"{REFJF} dump -c {s}.Jhash", work).stdout == sh(f"{JF} dump -c {s}.Jhash", work).stdout
        assert sh(f"{REFJF} histo -f {s}.Jhash", work).stdout == open(os.path.join(work, s + ".histo"), "rb").read()
        # and the file the real tool writes for the same command has byte-identical records
        sh(f"cat {fq[s][0]} {fq[s][1]} > {s}.both.fq; {REFJF} count --disk -m {g['k']} -L {lo} -s 16M -t 4 "
           f"-o {s}.ref.Jhash -C {s}.both.fq", work)
        assert records(os.path.join(work, s + ".Jhash")) == records(os.path.join(work, s + ".ref.Jhash"))
        # we read the real file
        assert sh(f"{JF} dump -c {s}.ref.Jhash", work).stdout == sh(f"{REFJF} dump -c {s}.ref.Jhash", work).stdout
    ours = " ".join(s + ".Jhash" for s in names)
    mixed = f"{names[0]}.Jhash {names[1]}.ref.Jhash {names[2]}.Jhash"
    ref_merge = sh(f"{REFJF} merge {ours}", work).stdout
    assert ref_merge == sh(f"{JF} merge {ours}", work).stdout          # same lines in the same order
    assert sh(f"{REFJF} merge {mixed}", work).stdout == ref_merge      # real merge accepts a mix of both writers
    assert sh(f"{JF} merge {mixed}", work).stdout == ref_merge
    q = f"query -s <(awk '{{print \">\"$1\"\\n\"$1}}' merge.out) {names[0]}.Jhash"
    sh(f"{JF} merge {ours} > merge.out", work)
    assert sh(f"{REFJF} {q}", work).stdout == sh(f"{JF} {q}", work).stdout


def test_rufus_build_matches_reference_goldens(tmp_path):
    """tools/bin/RUFUS.Build (legacy selection, src/RUFUS.Build.cpp) on the committed sorted dumps vs the
    outputs of the reference binary (tests/golden/build_golden.json), byte for byte incl. order."""
    import json
    gold = json.load(open(os.path.join(GOLDEN, "build_golden.json")))
    work = str(tmp_path)
    done = set()
    for c in gold["cases"]:
        case = c["case"]
        names = SAMPLES[case]
        if case not in done:
            for i, s in enumerate(names):
                txt = gzip.open(os.path.join(GOLDEN, case, s + ".dump.sorted.gz")).read().decode()
                open(os.path.join(work, f"{case}.{s}.tab"), "w").write(txt.replace(" ", "\t") if i == 0 else txt)
            done.add(case)
        ctl = " ".join(f"-c {case}.{s}.tab" for s in names[1:])
        sh(f"{BIN}/RUFUS.Build {ctl} -s {case}.{names[0]}.tab -o out -hs {c['k']} -mS {c['mS']} -max {c['max']} "
           f"-mC {c['mC']} -t 1 -d ' '", work)
        assert open(os.path.join(work, "out")).read().splitlines() == c["lines"], (case, c["mS"], c["max"], c["mC"])
        assert os.path.exists(os.path.join(work, "out.Buld.Log"))


def test_annotate_overlap_matches_reference_golden(tmp_path):
    """tools/bin/AnnotateOverlap (the .hashcount.fastq writer) vs the committed output of the reference binary"""
    import json
    d = os.path.join(GOLDEN, "annotate")
    g = json.load(open(os.path.join(d, "golden.json")))
    r = subprocess.run([os.path.join(BIN, "AnnotateOverlap"), os.path.join(d, "HashList"), os.path.join(d, "reads.fastq"),
                        str(tmp_path / "hash.out")], capture_output=True, check=True)
    assert r.stdout.decode().splitlines()[3::4][:len(g["depth_lines"])] == g["depth_lines"]
    assert md5(r.stdout) == g["stdout_md5"]
    assert md5(open(tmp_path / "hash.out", "rb").read()) == g["hashout_md5"]
    # stdin form (Overlap.shorter.sh pipes the contigs in)
    r2 = subprocess.run([os.path.join(BIN, "AnnotateOverlap"), os.path.join(d, "HashList"), "stdin",
                         str(tmp_path / "hash2.out")], input=open(os.path.join(d, "reads.fastq"), "rb").read(),
                        capture_output=True, check=True)
    assert r2.stdout == r.stdout


def test_overlap_matches_reference_goldens(tmp_path):
    """tools/bin/Overlap (greedy assembly; seed index, seed lookups and Align3 on the GPU) on the three rounds of
    Overlap.shorter.sh:141,153,165 vs the committed outputs of the reference binary (tests/golden/overlap)"""
    import json
    d = os.path.join(GOLDEN, "overlap")
    g = json.load(open(os.path.join(d, "golden.json")))
    work = str(tmp_path)
    shutil.copy(os.path.join(d, "in.sam.fastqd"), work)
    for a in g["rounds"]:
        sh(f"{BIN}/Overlap " + " ".join(a), work)       # each round reads the previous round's OUR output
        for ext in (".fastqd", ".fastq"):
            assert open(os.path.join(work, a[7] + ext)).read() == open(os.path.join(d, a[7] + ext)).read(), (a, ext)


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4, 5])
def test_overlap_fuzz_matches_oracle(tmp_path, seed):
    """random overlapping reads (both strands, substitutions, N, depths > 1, mixed strand strings): our Overlap vs
    the oracle restatement (itself pinned to the reference binary, tests/test_oracle.py), and vs the binary when present"""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import overlap_oracle as oo
    from cases import overlap_fuzz_input
    text = overlap_fuzz_input(seed, n_reads=90)
    args = [[".98", "100", "1", "FP", "20", "1", "o", "0", "1"], [".95", "40", "1", "FP", "20", "1", "o", "1", "1"],
            [".9", "30", "2", "X", "15", "2", "o", "2", "1"]][seed % 3]
    open(tmp_path / "in.fastqd", "w").write(text)
    sh(f"{BIN}/Overlap in.fastqd " + " ".join(args), str(tmp_path))
    fq, fqd, bad = oo.overlap(text, args[0], int(args[1]), int(args[2]), int(args[4]), int(args[5]), int(args[7]), args[4], 1)
    assert open(tmp_path / "o.fastqd").read() == fqd
    assert open(tmp_path / "o.fastq").read() == fq
    assert open(tmp_path / "o.fastqbad.fastq").read() == bad
    ref = os.path.join(ROOT, "oracle", "_ref", "Overlap")
    if os.path.exists(ref):
        os.makedirs(tmp_path / "ref")
        shutil.copy(tmp_path / "in.fastqd", tmp_path / "ref")
        sh(f"{ref} in.fastqd " + " ".join(args), str(tmp_path / "ref"))
        assert open(tmp_path / "ref" / "o.fastqd").read() == fqd


def test_bam_mode_chain_matches_reference(tmp_path):
    """BASELINE configs[0] in BAM mode, with the stage commands of RunJellyForRUFUS.sh:22-37 and
    runRufus.sh:925,966-967 (FIFOs included) and OUR executables end to end: rg_bamview ->
    PassThroughSamCheck -> jellyfish count/histo/merge/query -> PassThroughSamCheck.stranded ->
    RUFUS.Filter. Golden = the reference's binaries downstream of the same view step
    (tests/golden/make_golden.py bam_case)."""
    g = golden("testRun_bam")
    work = str(tmp_path)
    k, names = g["k"], ("Child", "Mother", "Father")
    gen = {s: f"{BIN}/rg_bamview view -F 3328 {GOLDEN}/testRun_bam/{s}.bam" for s in names}
    for i, s in enumerate(names):
        lo = g["L"] if i == 0 else g["ctlL"]
        sh(f"mkfifo {s}.fq; {gen[s]} | {BIN}/PassThroughSamCheck {s}.Jelly.chr > {s}.fq & "
           f"timeout 60 {JF} count --disk -m {k} -L {lo} -s 16M -t 4 -o {s}.Jhash -C {s}.fq; rm {s}.fq; wait", work)
        sh(f"{JF} histo -f -o {s}.histo {s}.Jhash", work)
        gs = g["samples"][s]
        assert md5(open(os.path.join(work, s + ".histo"), "rb").read()) == gs["histo_md5"]
        dump = sh(f"{JF} dump -c {s}.Jhash | LC_ALL=C sort", work).stdout
        assert dump.count(b"\n") == gs["n_distinct"] and md5(dump) == gs["dump_sorted_md5"]
    merged = sh(f"{JF} merge " + " ".join(s + ".Jhash" for s in names) + " > merge.out; cat merge.out", work).stdout
    assert sorted(merged.decode().splitlines()) == g["merge_sorted"]
    q = sh(f"{JF} query -s <(awk '{{print \">\"$1\"\\n\"$1}}' merge.out) Child.Jhash", work).stdout.decode()
    hl = [l for l in q.splitlines() if g["min_cov"] <= int(l.split()[1]) <= g["max_cov"]]
    assert sorted(hl) == g["hashlist_sorted"]
    open(os.path.join(work, "HashList"), "w").write("\n".join(hl) + "\n")
    # the read pull reads the re-paired mates through two FIFOs while the re-pairing is still running
    sh(f"mkfifo t.mate1.fastq t.mate2.fastq; {gen['Child']} | {BIN}/PassThroughSamCheck.stranded Child.filter.chr t > t.log & "
       f"timeout 60 {BIN}/RUFUS.Filter HashList t.mate1.fastq t.mate2.fastq Child {k} {g['minq']} {g['thresh']} 4; wait", work)
    o1 = open(os.path.join(work, "Child.Mutations.Mate1.fastq")).read().splitlines()
    o2 = open(os.path.join(work, "Child.Mutations.Mate2.fastq")).read().splitlines()
    assert sorted(o1[0::4]) == g["pulled_mate1_names_sorted"]
    assert sorted(o2[0::4]) == g["pulled_mate2_names_sorted"]
    # and the re-paired files themselves are the reference's, byte for byte
    sh(f"{gen['Child']} | {BIN}/PassThroughSamCheck.stranded c.chr r > r.log", work)
    assert [md5(open(os.path.join(work, f"r.mate{m}.fastq"), "rb").read()) for m in (1, 2)] == g["repaired_mate_md5"]
    assert open(os.path.join(work, "r.mate1.fastq")).read().count("\n") // 4 == g["repaired_pairs"]
