#!/usr/bin/env python
"""Regenerate tests/golden/* by running the REFERENCE's own binaries.

Runs only in the build container (needs /root/reference and oracle/_ref built by
`make -C oracle ref`). It replays the hot-path commands of the reference
verbatim:

  bash X.generator | PassThroughSamCheck X.chr > X.fq        RunJellyForRUFUS.sh:28
  jellyfish count --disk -m K -L 2 -s SIZE -t T -o X.Jhash -C X.fq      ...:29
  jellyfish histo -f -o X.Jhash.histo X.Jhash                           ...:37
  jellyfish merge S.Jhash C1.Jhash C2.Jhash > S.temp            runRufus.sh:925
  jellyfish query -s <(awk fasta) S.Jhash | awk >=Min | awk <=Max       CheckJellyHashList.sh:12
  RUFUS.Filter HashList m1.fq m2.fq stub K MinQ Thresh Threads  runRufus.sh:981
  RUFUS.Filter.single HashList m1.fq stub K MinQ Thresh Threads

`-s 16M` instead of `-s 8G`: the (k-mer,count) multiset does not depend on the
table size ([probed]); only record order does, which nothing
downstream uses. Outputs are committed as small sorted text fixtures.
"""
import gzip, hashlib, json, os, shutil, subprocess, sys, tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REFBIN = os.path.join(ROOT, "oracle", "_ref")
REF = "/root/reference"
JF = os.path.join(REFBIN, "jellyfish")

from rufus_b200 import synth  # noqa: E402


def sh(cmd, **kw):
    return subprocess.run(["bash", "-c", cmd], check=True, **kw)


def md5(b: bytes) -> str:
    return hashlib.md5(b).hexdigest()


def run_case(work, fq, k, min_cov, max_cov, minq=15, thresh=1, L=2, ctlL=2, size="16M", sweep=(), generators=None):
    """fq: dict sample -> (mate1.fastq path, mate2.fastq path); first sample = subject.
    generators: dict sample -> text of the .generator script (BAM mode); the subject's mate files for
    the read pull are then made from its generator by PassThroughSamCheck.stranded (runRufus.sh:966)."""
    res = {"k": k, "min_cov": min_cov, "max_cov": max_cov, "minq": minq, "thresh": thresh,
           "L": L, "ctlL": ctlL, "samples": {}}
    names = list(fq)
    for i, s in enumerate(names):
        gen = os.path.join(work, s + ".generator")
        with open(gen, "w") as f:
            if generators:
                f.write(generators[s])
            else:
                for p in fq[s]:
                    f.write(f"perl {REF}/scripts/FastqToSam.pl <(cat {p})\n")
        lo = L if i == 0 else ctlL
        sh(f"bash {gen} | {REFBIN}/PassThroughSamCheck {gen}.Jelly.chr > {gen}.fq", cwd=work)
        sh(f"{JF} count --disk -m {k} -L {lo} -s {size} -t 4 -o {gen}.Jhash -C {gen}.fq", cwd=work)
        sh(f"{JF} histo -f -o {gen}.Jhash.histo {gen}.Jhash", cwd=work)
        dump = subprocess.run(f"{JF} dump -c {gen}.Jhash | LC_ALL=C sort", shell=True, check=True,
                              capture_output=True, cwd=work).stdout
        histo = open(gen + ".Jhash.histo").read()
        nz = {int(a): int(b) for a, b in (l.split() for l in histo.splitlines()) if int(b)}
        res["samples"][s] = {"n_distinct": dump.count(b"\n"), "dump_sorted_md5": md5(dump),
                             "histo_md5": md5(histo.encode()), "histo_rows": len(histo.splitlines()),
                             "histo_nonzero": nz}
        res["samples"][s]["_dump"] = dump
    subj = names[0]
    jh = lambda s: os.path.join(work, s + ".generator.Jhash")
    merged = subprocess.run(f"{JF} merge " + " ".join(jh(s) for s in names), shell=True, check=True,
                            capture_output=True, cwd=work).stdout
    open(os.path.join(work, "merge.out"), "wb").write(merged)
    hl = subprocess.run(
        ["bash", "-c", f"{JF} query -s <(cat merge.out | awk '{{print \">\"$1\"\\n\"$1}}') {jh(subj)}"
         f" | awk -v var={min_cov} ' $2 >= var ' | awk -v var={max_cov} ' $2 <= var '"],
        check=True, capture_output=True, cwd=work).stdout
    open(os.path.join(work, "HashList"), "wb").write(hl)
    res["merge_sorted"] = sorted(merged.decode().splitlines())
    res["hashlist_sorted"] = sorted(hl.decode().splitlines())
    if generators:
        gen = os.path.join(work, subj + ".generator")
        sh(f"bash {gen} | {REFBIN}/PassThroughSamCheck.stranded {gen}.filter.chr {gen}.temp > {gen}.temp", cwd=work)
        m1, m2 = gen + ".temp.mate1.fastq", gen + ".temp.mate2.fastq"
        res["repaired_pairs"] = open(m1).read().count("\n") // 4
        res["repaired_mate_md5"] = [md5(open(m1, "rb").read()), md5(open(m2, "rb").read())]
    else:
        m1, m2 = fq[subj]
    sh(f"{REFBIN}/RUFUS.Filter HashList {m1} {m2} {work}/out {k} {minq} {thresh} 4 > /dev/null", cwd=work)
    o1 = open(os.path.join(work, "out.Mutations.Mate1.fastq")).read().splitlines()
    o2 = open(os.path.join(work, "out.Mutations.Mate2.fastq")).read().splitlines()
    res["pulled_mate1_names_sorted"] = sorted(o1[0::4])
    res["pulled_mate2_names_sorted"] = sorted(o2[0::4])
    sh(f"{REFBIN}/RUFUS.Filter.single HashList {m1} {work}/se {k} {minq} {thresh} 4 > /dev/null", cwd=work)
    se = open(os.path.join(work, "se.Mutations.fastq")).read().splitlines()
    res["single_names_sorted"] = sorted(se[0::4])
    # min-count sweep (BASELINE configs[4]): the same merge output cut at other MutantMinCov values
    # (CheckJellyHashList.sh:12 awk filters), each HashList pulled through RUFUS.Filter
    if sweep:
        res["sweep"] = {}
    for mc in sweep:
        hl = subprocess.run(
            ["bash", "-c", f"{JF} query -s <(cat merge.out | awk '{{print \">\"$1\"\\n\"$1}}') {jh(subj)}"
             f" | awk -v var={mc} ' $2 >= var ' | awk -v var={max_cov} ' $2 <= var '"],
            check=True, capture_output=True, cwd=work).stdout
        open(os.path.join(work, f"HashList.{mc}"), "wb").write(hl)
        sh(f"{REFBIN}/RUFUS.Filter HashList.{mc} {m1} {m2} {work}/sw{mc} {k} {minq} {thresh} 4 > /dev/null", cwd=work)
        o1 = open(os.path.join(work, f"sw{mc}.Mutations.Mate1.fastq")).read().splitlines()
        res["sweep"][str(mc)] = {"hashlist_md5": md5(("\n".join(sorted(hl.decode().splitlines())) + "\n").encode()),
                                 "hashlist_lines": len(hl.decode().splitlines()),
                                 "pulled_mate1_names_sorted": sorted(o1[0::4])}
    return res


def strip_dumps(res, outdir, keep):
    for s, d in res["samples"].items():
        dump = d.pop("_dump")
        if keep:
            with gzip.GzipFile(os.path.join(outdir, s + ".dump.sorted.gz"), "wb", mtime=0) as f:
                f.write(dump)


def bundled(work):
    out = os.path.join(HERE, "testRun")
    os.makedirs(out, exist_ok=True)
    fq = {}
    for s in ("Child", "Mother", "Father"):
        ps = []
        for m in (1, 2):
            src = f"{REF}/testRun/{s}.mate{m}.fastq"
            with open(src, "rb") as fi, gzip.GzipFile(os.path.join(out, f"{s}.mate{m}.fastq.gz"), "wb", mtime=0) as fo:
                fo.write(fi.read())
            ps.append(src)
        fq[s] = tuple(ps)
    # MinCov 5 / MaxHashDepth 140 = what ModelDist picks on the child histogram
    res = run_case(work, fq, 25, 5, 140)
    strip_dumps(res, out, keep=True)
    json.dump(res, open(os.path.join(out, "golden.json"), "w"), indent=0, sort_keys=True)
    print("bundled: distinct", {s: d["n_distinct"] for s, d in res["samples"].items()},
          "hashlist", len(res["hashlist_sorted"]), "pulled", len(res["pulled_mate1_names_sorted"]))


def bam_case(work):
    """BASELINE configs[0] as literally written: the bundled testRun/*.bam in BAM mode (runRufus.sh:599
    generator `samtools view -F 3328 X.bam`, RunJellyForRUFUS.sh:28-29, runRufus.sh:966-967). samtools is
    not installed here, so the view step is tools/bin/rg_bamview (pinned separately: tests/golden/bam);
    EVERYTHING downstream of it is the reference's own binaries."""
    out = os.path.join(HERE, "testRun_bam")
    os.makedirs(out, exist_ok=True)
    view = os.path.join(ROOT, "tools", "bin", "rg_bamview")
    gens = {}
    for s in ("Child", "Mother", "Father"):
        dst = os.path.join(out, s + ".bam")
        shutil.copyfile(f"{REF}/testRun/{s}.bam", dst)
        os.chmod(dst, 0o644)
        gens[s] = f"{view} view -F 3328 {dst}\n"
    res = run_case(work, {s: None for s in gens}, 25, 5, 140, generators=gens)
    strip_dumps(res, out, keep=False)
    json.dump(res, open(os.path.join(out, "golden.json"), "w"), indent=0, sort_keys=True)
    print("testRun_bam: distinct", {s: d["n_distinct"] for s, d in res["samples"].items()},
          "hashlist", len(res["hashlist_sorted"]), "re-paired", res["repaired_pairs"],
          "pulled", len(res["pulled_mate1_names_sorted"]), "single", len(res["single_names_sorted"]))


def synth_case(work, name, seed, G, n_pairs, k, min_cov, max_cov, kind="trio", **kw):
    out = os.path.join(HERE, name)
    os.makedirs(out, exist_ok=True)
    make = {"trio": synth.make_trio, "pair": synth.make_pair, "quad": synth.make_quad}[kind]
    trio, _ = make(seed, G, n_pairs)
    fq, insum = {}, {}
    for s in trio:
        s1, q1, s2, q2 = trio[s]
        p1, p2 = os.path.join(work, f"{s}.m1.fq"), os.path.join(work, f"{s}.m2.fq")
        open(p1, "wb").write(synth.fastq_bytes(s1, q1, prefix=s[0]))
        open(p2, "wb").write(synth.fastq_bytes(s2, q2, prefix=s[0]))
        fq[s] = (p1, p2)
        insum[s] = md5(s1.tobytes() + q1.tobytes() + s2.tobytes() + q2.tobytes())
    res = run_case(work, fq, k, min_cov, max_cov, **kw)
    res["synth"] = {"seed": seed, "genome_len": G, "n_pairs": n_pairs, "input_md5": insum}
    if kind != "trio":
        res["synth"]["kind"] = kind
    strip_dumps(res, out, keep=False)
    json.dump(res, open(os.path.join(out, "golden.json"), "w"), indent=0, sort_keys=True)
    print(name, "distinct", {s: d["n_distinct"] for s, d in res["samples"].items()},
          "hashlist", len(res["hashlist_sorted"]), "pulled", len(res["pulled_mate1_names_sorted"]),
          "single", len(res["single_names_sorted"]))


EDGE_READS = [
    # (name, seq, qual) — crafted to hit the quirks of SURVEY Appendix B
    ("lower", "acgtacgtacgtacgtacgtacgtacgtACGTACGTTTGACCA", None),
    ("withN", "ACGTTGCATGCATGCCGTAGCTAGCTAGGNCTAGCTAGGATCGATCGATTTTGACAGCATCGAC", None),
    ("iupac", "ACGTTGCATGCATGCCGTAGCTAGCTAGGRCTAGCTAGGATCGATCGATTTTGACAGCATCGAC", None),
    ("short", "ACGTACGTACGTACGTACGTACGT", None),
    ("exactk", "ACGTTGCATGCATGCCGTAGCTAGC", None),
    ("kplus1", "ACGTTGCATGCATGCCGTAGCTAGCT", None),
    ("polyA", "A" * 80, None),
    ("palin", "ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT", None),
    ("lowq", "GATTACAGATTACAGATTACAGATTACAGATTACAGATTACAGATTACAGATTACA", "mid"),
    ("lastbase", "TTGACCGGTAACGTTAGCCGATCGGATCAGGTTCAAGCTTA", None),
    ("ctl1x", "TGCATCGGATTCAGGCTAAGCTTCGGATACCGATTGCAT", None),
]


def edge_case(work):
    """Tiny hand-made pairs: each read repeated 6x in the subject so counts >= 5, absent in controls."""
    out = os.path.join(HERE, "edge")
    os.makedirs(out, exist_ok=True)
    import random
    rnd = random.Random(5)
    def q(seq, mode):
        if mode == "mid":
            return "".join("J" if (i < 30 or i > 33) else "#" for i in range(len(seq)))
        return "J" * len(seq)
    filler = "".join(rnd.choice("ACGT") for _ in range(60))
    fq = {}
    for s in ("Subj", "CtlA", "CtlB"):
        p1, p2 = os.path.join(out, f"{s}.mate1.fastq"), os.path.join(out, f"{s}.mate2.fastq")
        with open(p1, "w") as f1, open(p2, "w") as f2:
            reps = 6 if s == "Subj" else 0
            for name, seq, mode in EDGE_READS:
                for r in range(reps):
                    f1.write(f"@{name}.{r}/1\n{seq}\n+\n{q(seq, mode)}\n")
                    f2.write(f"@{name}.{r}/2\n{filler}\n+\n{'J' * len(filler)}\n")
            # shared background present in every sample (gets subtracted), 7 copies
            for r in range(7):
                f1.write(f"@bg.{r}/1\n{filler}\n+\n{'J' * len(filler)}\n")
                f2.write(f"@bg.{r}/2\n{filler[::-1]}\n+\n{'J' * len(filler)}\n")
            if s == "CtlA":   # control holds one subject read once (count 1 < ParLowK=2 => "absent")
                f1.write(f"@once/1\n{EDGE_READS[10][1]}\n+\n{'J' * len(EDGE_READS[10][1])}\n")
                f2.write(f"@once/2\n{filler}\n+\n{'J' * len(filler)}\n")
            if s == "CtlB":   # control holds another subject read twice (count 2 => present => subtracted)
                for r in range(2):
                    f1.write(f"@twice.{r}/1\n{EDGE_READS[9][1]}\n+\n{'J' * len(EDGE_READS[9][1])}\n")
                    f2.write(f"@twice.{r}/2\n{filler}\n+\n{'J' * len(filler)}\n")
            # a mate-2-only hit: mate1 is filler, mate2 carries a subject-only sequence
            if s == "Subj":
                m2seq = "CCATGGTTAACCGGATATCGCGATTAGCATCGGACTTACG"
                for r in range(6):
                    f1.write(f"@m2only.{r}/1\n{filler}\n+\n{'J' * len(filler)}\n")
                    f2.write(f"@m2only.{r}/2\n{m2seq}\n+\n{'J' * len(m2seq)}\n")
        fq[s] = (p1, p2)
    res = run_case(work, fq, 25, 5, 1200)
    # an `-e` exclude file (runRufus.sh:925: extra merge inputs): ANY record excludes, whatever its
    # count. Here: two subject-only reads seen ONCE, counted without -L (count-1 records are kept)
    excl_reads = [EDGE_READS[10][1], EDGE_READS[1][1]]
    with open(os.path.join(work, "excl.fa"), "w") as f:
        for i, r in enumerate(excl_reads):
            f.write(f">x{i}\n{r}\n")
    sh(f"{JF} count -m 25 -s 16M -t 2 -C -o excl.Jhash excl.fa", cwd=work)
    jh = lambda s: os.path.join(work, s + ".generator.Jhash")
    merged = subprocess.run(f"{JF} merge " + " ".join(jh(s) for s in fq) + " excl.Jhash", shell=True, check=True,
                            capture_output=True, cwd=work).stdout
    open(os.path.join(work, "merge_e.out"), "wb").write(merged)
    hl = subprocess.run(
        ["bash", "-c", f"{JF} query -s <(cat merge_e.out | awk '{{print \">\"$1\"\\n\"$1}}') {jh('Subj')}"
         f" | awk -v var=5 ' $2 >= var ' | awk -v var=1200 ' $2 <= var '"],
        check=True, capture_output=True, cwd=work).stdout
    res["exclude"] = {"reads": excl_reads, "merge_sorted": sorted(merged.decode().splitlines()),
                      "hashlist_sorted": sorted(hl.decode().splitlines())}
    assert len(res["exclude"]["hashlist_sorted"]) < len(res["hashlist_sorted"])
    strip_dumps(res, out, keep=True)
    json.dump(res, open(os.path.join(out, "golden.json"), "w"), indent=0, sort_keys=True)
    print("edge: distinct", {s: d["n_distinct"] for s, d in res["samples"].items()},
          "hashlist", len(res["hashlist_sorted"]), "pulled", res["pulled_mate1_names_sorted"],
          "single", res["single_names_sorted"])


if __name__ == "__main__":
    only = set(sys.argv[1:])
    for fn, args, kw in ((bundled, (), {}), (bam_case, (), {}), (edge_case, (), {}),
                         (synth_case, ("synth_k25", 7, 50000, 5000, 25, 5, 150), {}),
                         (synth_case, ("synth_k31", 11, 40000, 4000, 31, 5, 150), {}),
                         # scaled-down BASELINE configs[3] (60x tumor vs 30x normal, one control) and
                         # configs[4] (proband + 3 controls, k = 31, min-count sweep)
                         (synth_case, ("pair_k25", 21, 40000, 4000, 25, 5, 300), {"kind": "pair"}),
                         (synth_case, ("quad_k31", 31, 40000, 4000, 31, 5, 150), {"kind": "quad", "sweep": (2, 3, 8, 12)})):
        if only and (args[0] if args else fn.__name__) not in only:
            continue
        work = tempfile.mkdtemp(prefix="rg_golden_")
        try:
            fn(work, *args, **kw)
        finally:
            shutil.rmtree(work, ignore_errors=True)
