"""A minimal BAM writer for tests (SAM spec §4.1-4.2), independent of any decoder: records are given
field by field, `expected_sam` states the text `samtools view` prints for them (§1.4)."""
import struct
import zlib

_CODES = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}
_OPS = {c: i for i, c in enumerate("MIDNSHP=X")}
_FMT = {"c": "<b", "C": "<B", "s": "<h", "S": "<H", "i": "<i", "I": "<I", "f": "<f"}


def bgzf(data: bytes, block: int) -> bytes:
    """BGZF members of `block` inflated bytes each, plus the empty end-of-file member"""
    o = b""
    for p in list(range(0, len(data), block)) + [None]:
        chunk = b"" if p is None else data[p:p + block]
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        comp = c.compress(chunk) + c.flush()
        bsize = 12 + 6 + len(comp) + 8
        o += b"\x1f\x8b\x08\x04" + b"\0" * 4 + b"\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1)
        o += comp + struct.pack("<II", zlib.crc32(chunk), len(chunk))
    return o


def _g(v: float) -> str:
    return "%g" % struct.unpack("<f", struct.pack("<f", v))[0]


def tag_text(tag, t, v) -> str:
    if t == "A":
        return f"{tag}:A:{v}"
    if t in "ZH":
        return f"{tag}:{t}:{v}"
    if t == "B":
        sub, vals = v
        return f"{tag}:B:{sub}" + "".join("," + (_g(x) if sub == "f" else str(x)) for x in vals)
    if t == "f":
        return f"{tag}:f:{_g(v)}"
    return f"{tag}:i:{v}"


def build_bam(text, refs, recs, block=997) -> bytes:
    """recs: (name, flag, ref, pos0, mapq, [(n, op)], next_ref, next_pos0, tlen, seq, qual or None, [(tag, type, value)])"""
    body = b"BAM\x01" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", len(refs))
    for n, l in refs:
        body += struct.pack("<i", len(n) + 1) + n.encode() + b"\0" + struct.pack("<i", l)
    for name, flag, ref, pos, mapq, cig, nref, npos, tlen, seq, qual, tags in recs:
        r = struct.pack("<iiBBHHHiiii", ref, pos, len(name) + 1, mapq, 4680, len(cig), flag, len(seq), nref, npos, tlen)
        r += name.encode() + b"\0"
        r += b"".join(struct.pack("<I", (n << 4) | _OPS[o]) for n, o in cig)
        padded = seq + ("=" if len(seq) & 1 else "")
        r += bytes((_CODES[padded[i]] << 4) | _CODES[padded[i + 1]] for i in range(0, len(padded), 2))
        r += bytes([0xFF] * len(seq)) if qual is None else bytes(ord(c) - 33 for c in qual)
        for tag, t, v in tags:
            r += tag.encode() + t.encode()
            if t == "A":
                r += v.encode()
            elif t in "ZH":
                r += v.encode() + b"\0"
            elif t == "B":
                r += v[0].encode() + struct.pack("<I", len(v[1])) + b"".join(struct.pack(_FMT[v[0]], x) for x in v[1])
            else:
                r += struct.pack(_FMT[t], v)
        body += struct.pack("<i", len(r)) + r
    return bgzf(body, block)


def expected_sam(refs, recs):
    name_of = lambda i: "*" if i < 0 else refs[i][0]
    lines = []
    for name, flag, ref, pos, mapq, cig, nref, npos, tlen, seq, qual, tags in recs:
        f = [name, str(flag), name_of(ref), str(pos + 1), str(mapq), "".join(f"{n}{o}" for n, o in cig) or "*",
             "*" if nref < 0 else ("=" if nref == ref else name_of(nref)), str(npos + 1), str(tlen),
             seq or "*", "*" if (qual is None or not seq) else qual]
        f += [tag_text(*t) for t in tags]
        lines.append("\t".join(f))
    return lines
