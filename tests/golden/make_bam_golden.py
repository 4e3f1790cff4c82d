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
import bam_oracle  # noqa: E402
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


def bgzf(data, block):
    o = b""
    for p in list(range(0, len(data), block)) + [None]:
        chunk = b"" if p is None else data[p:p + block]
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        comp = c.compress(chunk) + c.flush()
        bsize = 12 + 6 + len(comp) + 8
        o += b"\x1f\x8b\x08\x04" + b"\0" * 4 + b"\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1)
        o += comp + struct.pack("<II", zlib.crc32(chunk), len(chunk))
    return o


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
    body = b"BAM\x01" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", len(refs))
    for n, l in refs:
        body += struct.pack("<i", len(n) + 1) + n.encode() + b"\0" + struct.pack("<i", l)
    fmt = {"c": "<b", "C": "<B", "s": "<h", "S": "<H", "i": "<i", "I": "<I", "f": "<f"}
    sam = []
    for name, flag, ref, pos, mapq, cig, nref, npos, tlen, seq, qual, tags, tagtext in recs:
        r = struct.pack("<iiBBHHHiiii", ref, pos, len(name) + 1, mapq, 4680, len(cig), flag, len(seq), nref, npos, tlen)
        r += name.encode() + b"\0"
        r += b"".join(struct.pack("<I", (n << 4) | ops[o]) for n, o in cig)
        padded = seq + ("=" if len(seq) & 1 else "")
        r += bytes((codes[padded[i]] << 4) | codes[padded[i + 1]] for i in range(0, len(padded), 2))
        r += bytes([0xFF] * len(seq)) if qual is None else bytes(ord(c) - 33 for c in qual)
        for tag, t, v in tags:
            r += tag.encode() + t.encode()
            if t == "A":
                r += v.encode()
            elif t in "ZH":
                r += v.encode() + b"\0"
            elif t == "B":
                r += v[0].encode() + struct.pack("<I", len(v[1])) + b"".join(struct.pack(fmt[v[0]], x) for x in v[1])
            else:
                r += struct.pack(fmt[t], v)
        body += struct.pack("<i", len(r)) + r
        refname = lambda i: "*" if i < 0 else refs[i][0]
        sam.append("\t".join([name, str(flag), refname(ref), str(pos + 1), str(mapq),
                              "".join(f"{n}{o}" for n, o in cig) or "*",
                              "*" if nref < 0 else ("=" if nref == ref else refname(nref)), str(npos + 1), str(tlen),
                              seq or "*", "*" if (qual is None or not seq) else qual, tagtext]))
    open(os.path.join(out, "synthetic.bam"), "wb").write(bgzf(body, 997))
    open(os.path.join(out, "synthetic.sam"), "w").write("\n".join(sam) + "\n")
    g["synthetic"] = {"records": len(sam), "header": text}
    json.dump(g, open(os.path.join(out, "golden.json"), "w"), indent=1, sort_keys=True)
    print({s: (e["F3328"]["records"], e["all"]["records"], e["fastq_pairs"]) for s, e in g["bundled"].items()}, len(sam))


main()
