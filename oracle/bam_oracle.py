"""TEST INFRASTRUCTURE ONLY (never on the product path): restatement in plain Python of what a
`.generator` script's `samtools view -F N X.bam` prints (runRufus.sh:599) — BGZF/BAM -> SAM text,
following the SAM/BAM specification (§1.4 text fields, §4.1 BGZF, §4.2 record layout).

Parity status: samtools is not installed in the build container, so the text cannot be
diffed against samtools itself: "parity unpinned" for the optional-tag formatting. What IS pinned to
the reference's own data: the names, flags, sequences and qualities of the bundled testRun/*.bam
reproduce the bundled testRun/*.mate{1,2}.fastq after re-pairing (tests/test_tools_host.py)."""
import struct
import zlib

_CIG = "MIDNSHP=X"
_SEQ = "=ACMGRSVTWYHKDBN"
_FMT = {"c": "<b", "C": "<B", "s": "<h", "S": "<H", "i": "<i", "I": "<I", "f": "<f"}


def bgzf_inflate(raw: bytes) -> bytes:
    out, p = [], 0
    while p < len(raw):
        if raw[p:p + 3] != b"\x1f\x8b\x08" or not raw[p + 3] & 4:
            raise ValueError("not BGZF")
        xlen = struct.unpack_from("<H", raw, p + 10)[0]
        q, bsize = p + 12, None
        while q < p + 12 + xlen:
            slen = struct.unpack_from("<H", raw, q + 2)[0]
            if raw[q:q + 2] == b"BC":
                bsize = struct.unpack_from("<H", raw, q + 4)[0]
            q += 4 + slen
        out.append(zlib.decompress(raw[p + 12 + xlen:p + bsize + 1 - 8], -15))
        p += bsize + 1
    return b"".join(out)


def _g(v: float) -> str:
    return "%g" % struct.unpack("<f", struct.pack("<f", v))[0]


def view(path: str, exclude: int = 0, require: int = 0):
    """-> (header text, [SAM lines without newline]) of the records passing the flag filter"""
    d = bgzf_inflate(open(path, "rb").read())
    if d[:4] != b"BAM\x01":
        raise ValueError("not BAM")
    l_text = struct.unpack_from("<i", d, 4)[0]
    header = d[8:8 + l_text].rstrip(b"\0").decode()
    p = 8 + l_text
    n_ref = struct.unpack_from("<i", d, p)[0]
    p += 4
    refs = []
    for _ in range(n_ref):
        l = struct.unpack_from("<i", d, p)[0]
        refs.append(d[p + 4:p + 4 + l - 1].decode())
        p += 4 + l + 4
    lines = []
    while p < len(d):
        bs = struct.unpack_from("<i", d, p)[0]
        rid, pos, l_rn, mapq, _bin, n_cig, flag, l_seq, nrid, npos, tlen = struct.unpack_from("<iiBBHHHiiii", d, p + 4)
        end = p + 4 + bs
        q = p + 36
        name = d[q:q + l_rn - 1].decode()
        q += l_rn
        cig = "".join(f"{v >> 4}{_CIG[v & 15]}" for v in struct.unpack_from(f"<{n_cig}I", d, q)) or "*"
        q += 4 * n_cig
        seq = "".join(_SEQ[(d[q + (i >> 1)] >> (4 * (1 - (i & 1)))) & 15] for i in range(l_seq)) or "*"
        q += (l_seq + 1) // 2
        qual = "*" if l_seq == 0 or d[q] == 0xFF else bytes(b + 33 for b in d[q:q + l_seq]).decode()
        q += l_seq
        rname = refs[rid] if rid >= 0 else "*"
        rnext = "*" if nrid < 0 else ("=" if nrid == rid else refs[nrid])
        f = [name, str(flag), rname, str(pos + 1), str(mapq), cig, rnext, str(npos + 1), str(tlen), seq, qual]
        while q < end:
            tag, t = d[q:q + 2].decode(), chr(d[q + 2])
            q += 3
            if t == "A":
                f.append(f"{tag}:A:{chr(d[q])}")
                q += 1
            elif t in "ZH":
                z = d.index(b"\0", q)
                f.append(f"{tag}:{t}:{d[q:z].decode()}")
                q = z + 1
            elif t == "B":
                sub, cnt = chr(d[q]), struct.unpack_from("<I", d, q + 1)[0]
                q += 5
                sz = struct.calcsize(_FMT[sub])
                vals = [struct.unpack_from(_FMT[sub], d, q + i * sz)[0] for i in range(cnt)]
                q += cnt * sz
                f.append(f"{tag}:B:{sub}" + "".join("," + (_g(v) if sub == "f" else str(v)) for v in vals))
            else:
                v = struct.unpack_from(_FMT[t], d, q)[0]
                q += struct.calcsize(_FMT[t])
                f.append(f"{tag}:f:{_g(v)}" if t == "f" else f"{tag}:i:{v}")
        if not (flag & exclude) and (flag & require) == require:
            lines.append("\t".join(f))
        p = end
    return header, lines
