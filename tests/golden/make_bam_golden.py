#!/usr/bin/env python
"""Fixtures for the BAM ingest (tools/bin/rg_bamview, oracle/bam_oracle.py). Run in the build container.

  (the bundled BAM files themselves are committed under tests/golden/testRun_bam/ by make_golden.py)
  tests/golden/bam/golden.json       per bundled BAM: record counts and md5 of the SAM text of
                                     oracle/bam_oracle.py with -F 3328 (runRufus.sh:599) and without a
                                     filter; md5 of the re-paired reads, which must equal the
                                     bundled testRun/X.mate{1,2}.fastq (names, sequences, qualities)
  tests/golden/bam/synthetic.bam     written HERE field by field with a small BGZF block size, so that
  tests/golden/bam/synthetic.sam     records straddle blocks: every tag type, arrays, unmapped
                                     records, absent SEQ/QUAL/CIGAR; the expected text is built from
                                     the same source fields, independently of any decoder
"""
import hashlib, json, os, shutil, struct, sys, zlib
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bam_oracle  # noqa: E402
from bam_build import build_bam, expected_sam  # noqa: E402
REF = "/root/reference/testRun"
out = os.path.join(HERE, "bam")


def md5(b):
    return hashlib.md5(b).hexdigest()


def pairs_md5(m1, m2):
    """md5 over the sorted (name, {mate records}) of two FASTQ files: independent of which mate is
    called mate1 (PassThroughSamCheck.stranded names them by arrival order)"""
    def load(p):
        l = open(p).read().splitlines()
        return {l[i][1:]: (l[i + 1], l[i + 3]) for i in range(0, len(l), 4)}
    a, b = load(m1), load(m2)
    rows = sorted(n + "\t" + "\t".join("\t".join(x) for x in sorted((a[n], b[n]))) for n in a)
    return md5("\n".join(rows).encode()), len(rows)


def main():
    os.makedirs(out, exist_ok=True)
    g = {"generator": "tests/golden/make_bam_golden.py", "bundled": {}}
    for s in ("Child", "Mother", "Father"):
        e = {"bam_md5": md5(open(os.path.join(REF, s + ".bam"), "rb").read())}
        for tag, excl in (("F3328", 3328), ("all", 0)):
            _, lines = bam_oracle.view(os.path.join(REF, s + ".bam"), excl)
            e[tag] = {"records": len(lines), "sam_md5": md5(("\n".join(lines) + "\n").encode())}
        e["fastq_pairs_md5"], e["fastq_pairs"] = pairs_md5(os.path.join(REF, s + ".mate1.fastq"), os.path.join(REF, s + ".mate2.fastq"))
        g["bundled"][s] = e

    # ---- synthetic BAM
    refs = [("chr1", 1000), ("chr2", 2000)]
    text = "@HD\tVN:1.6\tSO:unsorted\n@SQ\tSN:chr1\tLN:1000\n@SQ\tSN:chr2\tLN:2000\n"
    codes = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}
    ops = {c: i for i, c in enumerate("MIDNSHP=X")}
    recs = [
        # name, flag, ref, pos0, mapq, cigar[(n,op)], nref, npos0, tlen, seq, qual(None = absent), tags[(tag,type,value)], expected tag text
        ("r1", 99, 0, 9, 60, [(5, "S"), (10, "M"), (1, "I"), (4, "M")], 0, 109, 120, "ACGTNACGTMRWSYKVHDBN"[:20], "I" * 20,
         [("NM", "C", 3), ("AS", "c", -5), ("XS", "s", -300), ("XT", "S", 65535), ("XI", "i", -70000), ("XU", "I", 4000000000),
          ("XA", "A", "Q"), ("RG", "Z", "grp 1"), ("XH", "H", "1AE301"), ("XF", "f", 0.5)],
         "NM:i:3\tAS:i:-5\tXS:i:-300\tXT:i:65535\tXI:i:-70000\tXU:i:4000000000\tXA:A:Q\tRG:Z:grp 1\tXH:H:1AE301\tXF:f:0.5"),
        ("r1", 147, 0, 109, 60, [(20, "M")], 0, 9, -120, "TTTTTGGGGGCCCCCAAAAA", "#" + "J" * 19,
         [("BC", "B", ("c", [-1, 2, 3])), ("BD", "B", ("S", [1, 65535])), ("BF", "B", ("f", [1.5, -2.0])), ("BI", "B", ("I", []))],
         "BC:B:c,-1,2,3\tBD:B:S,1,65535\tBF:B:f,1.5,-2\tBI:B:I"),
        ("unmapped/1", 77, -1, -1, 0, [], -1, -1, 0, "ACGTACG", None, [("YY", "i", 0)], "YY:i:0"),
        ("unmapped/2", 141, -1, -1, 0, [], -1, -1, 0, "", None, [("YZ", "Z", "")], "YZ:Z:"),
        ("chimeric", 2113, 1, 1999, 3, [(3, "H"), (6, "="), (2, "X"), (1, "D"), (2, "N"), (1, "P")], 0, 0, 0, "ACGTACGTA", "ABCDEFGHI",
         [("SA", "Z", "chr1,1,+,9M,3,0;")], "SA:Z:chr1,1,+,9M,3,0;"),
    ]
    # enough plain records to span many BGZF blocks
    for i in range(300):
        seq = "".join("ACGT"[(i * 7 + j * (1 + i % 3)) & 3] for j in range(37 + i % 5))
        recs.append((f"bulk{i:04d}", 1024 if i % 10 == 0 else (16 if i % 3 == 0 else 0), i % 2, i * 3, 30, [(len(seq), "M")], -1, -1, 0,
                     seq, "".join(chr(33 + (i + j) % 40) for j in range(len(seq))), [("NM", "C", i % 4)], f"NM:i:{i % 4}"))
    recs = [r[:12] for r in recs]
    sam = expected_sam(refs, recs)
    open(os.path.join(out, "synthetic.bam"), "wb").write(build_bam(text, refs, recs, block=997))
    open(os.path.join(out, "synthetic.sam"), "w").write("\n".join(sam) + "\n")
    g["synthetic"] = {"records": len(sam), "header": text}
    json.dump(g, open(os.path.join(out, "golden.json"), "w"), indent=1, sort_keys=True)
    print({s: (e["F3328"]["records"], e["all"]["records"], e["fastq_pairs"]) for s, e in g["bundled"].items()}, len(sam))


main()
