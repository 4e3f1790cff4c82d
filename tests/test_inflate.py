"""tools/rg_inflate.hpp (the DEFLATE decoder behind tools/bin/rg_bamview) against zlib: every block type (stored,
fixed, dynamic), every compression level and strategy, flush points in the middle of a stream, sizes from empty to
beyond the 32 KiB window, wrong output sizes, truncated and bit-flipped streams. The driver
(tools/micro/inflate_check.cpp) keeps a canary byte behind the output buffer. Host only."""
import os
import random
import struct
import subprocess
import zlib

import pytest

from conftest import ROOT

EXE = os.environ.get("RG_INFLATE_CHECK", os.path.join(ROOT, "tools", "bin", "rg_inflate_check"))   # (override: a sanitizer build)


def deflate(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, mem=8, flush_at=()):
    c = zlib.compressobj(level, zlib.DEFLATED, -15, mem, strategy)
    out, last = b"", 0
    for at, mode in flush_at:
        out += c.compress(data[last:at]) + c.flush(mode)
        last = at
    return out + c.compress(data[last:]) + c.flush()


def payload(rnd, kind, n):
    if kind == "random":
        return rnd.randbytes(n)
    if kind == "dna":                      # packed-nibble-like: 16 symbols, short matches everywhere
        return bytes(rnd.choice(b"\x11\x12\x14\x18\x21\x22\x24\x28\x41\x42\x44\x48\x81\x82\x84\x88") for _ in range(n))
    if kind == "qual":                     # 40 symbols, skewed: Huffman literals
        return bytes(min(40, int(rnd.expovariate(0.15))) + 2 for _ in range(n))
    if kind == "runs":                     # distance-1 and short-distance matches of every length
        out = bytearray()
        while len(out) < n:
            out += bytes([rnd.randrange(256)]) * rnd.choice([1, 2, 3, 7, 8, 9, 30, 257, 258, 259, 1000])
            if rnd.random() < 0.3:
                out += bytes(rnd.randrange(256) for _ in range(rnd.randrange(1, 8))) * rnd.randrange(1, 40)
        return bytes(out[:n])
    if kind == "text":
        words = [b"read", b"chr1", b"\t", b"151M", b"=", b"NM:i:0", b"ACGT" * 5, b"\n", b"FFFFJJJJ"]
        out = bytearray()
        while len(out) < n:
            out += rnd.choice(words)
        return bytes(out[:n])
    return b"\0" * n


def run_cases(cases, tmp_path):
    """cases: (compressed, out_len) -> list of None (rejected) or the inflated bytes"""
    blob = b"".join(struct.pack("<II", len(c), n) + c for c, n in cases)
    (tmp_path / "cases").write_bytes(blob)
    r = subprocess.run([EXE, str(tmp_path / "cases")], capture_output=True)
    assert r.returncode == 0, r.stderr
    out, p = [], 0
    for _, n in cases:
        ok = r.stdout[p]
        p += 1
        if ok:
            out.append(r.stdout[p:p + n])
            p += n
        else:
            out.append(None)
    assert p == len(r.stdout)
    return out


@pytest.fixture(scope="module", autouse=True)
def exe():
    if not os.path.exists(EXE):
        pytest.fail("tools/bin/rg_inflate_check is missing: run `make tools`")


@pytest.mark.parametrize("kind", ["random", "dna", "qual", "runs", "text", "zeros"])
def test_every_level_strategy_and_size(kind, tmp_path):
    rnd = random.Random(zlib.crc32(kind.encode()))
    cases, want = [], []
    for n in (0, 1, 2, 3, 17, 258, 259, 300, 4096, 65280, 70000, 300000):
        data = payload(rnd, kind, n)
        for level in (0, 1, 4, 6, 9):
            for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED):
                if n > 70000 and strategy != zlib.Z_DEFAULT_STRATEGY and level != 6:
                    continue
                cases.append((deflate(data, level, strategy, mem=rnd.choice([1, 8, 9])), n))
                want.append(data)
    got = run_cases(cases, tmp_path)
    bad = [i for i, (g, w) in enumerate(zip(got, want)) if g != w]
    assert not bad, f"{len(bad)} of {len(cases)} streams differ, first: case {bad[0]} (out_len {cases[bad[0]][1]})"


def test_flush_points_make_multi_block_streams(tmp_path):
    rnd = random.Random(5)
    cases, want = [], []
    for kind in ("dna", "qual", "text", "runs"):
        for n in (100, 5000, 65280, 200000):
            data = payload(rnd, kind, n)
            cuts = sorted(rnd.sample(range(1, n), min(6, n - 1)))
            flush_at = [(c, rnd.choice([zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH, zlib.Z_BLOCK])) for c in cuts]
            for level in (0, 1, 6):
                cases.append((deflate(data, level, flush_at=flush_at), n))
                want.append(data)
    assert run_cases(cases, tmp_path) == want


def test_wrong_output_size_is_rejected(tmp_path):
    data = payload(random.Random(9), "text", 5000)
    c = deflate(data)
    got = run_cases([(c, 5000), (c, 4999), (c, 5001), (c, 0), (c, 100000), (deflate(b""), 0), (deflate(b""), 1)], tmp_path)
    assert got[0] == data and got[5] == b""
    assert got[1] is None and got[2] is None and got[3] is None and got[4] is None and got[6] is None


def test_truncated_and_corrupted_streams_never_crash_and_never_disagree_with_zlib(tmp_path):
    rnd = random.Random(13)
    cases, ref = [], []
    for kind in ("dna", "qual", "text", "runs", "random"):
        for n in (50, 3000, 70000):
            data = payload(rnd, kind, n)
            c = deflate(data, rnd.choice([1, 6, 9]))
            for _ in range(40):
                bad = bytearray(c)
                how = rnd.randrange(3)
                if how == 0 and len(bad) > 1:
                    bad = bad[:rnd.randrange(0, len(bad))]
                elif how == 1:
                    for _ in range(rnd.randrange(1, 4)):
                        bad[rnd.randrange(len(bad))] ^= 1 << rnd.randrange(8)
                else:
                    at = rnd.randrange(len(bad))
                    bad[at:at + rnd.randrange(1, 9)] = rnd.randbytes(rnd.randrange(0, 9))
                bad = bytes(bad)
                cases.append((bad, n))
                try:
                    d = zlib.decompressobj(-15)
                    out = d.decompress(bad)
                    ref.append(out if d.eof and not d.unused_data and len(out) == n else None)
                except zlib.error:
                    ref.append(None)
    got = run_cases(cases, tmp_path)
    both = [(g, r) for g, r in zip(got, ref) if g is not None and r is not None]
    assert all(g == r for g, r in both)
    # whatever zlib accepts as a complete stream of exactly n bytes, the fast decoder accepts too
    assert all(g is not None for g, r in zip(got, ref) if r is not None)


def test_crc32_by_carryless_multiplication_equals_zlib(tmp_path):
    """tools/rg_crc32.hpp (what checks every BGZF member's inflated bytes) against zlib's crc32: 60 000 pseudo-random
    slices — shorter than one 64-byte fold, not a multiple of 16, tens of kilobytes, every alignment"""
    import numpy as np
    f = tmp_path / "bytes.bin"
    f.write_bytes(np.random.default_rng(5).integers(0, 256, 300_000, dtype=np.uint8).tobytes())
    r = subprocess.run([EXE, "--crc", str(f)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "60000 cases, 0 bad", r.stdout + r.stderr
