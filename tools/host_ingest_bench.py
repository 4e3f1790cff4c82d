#!/usr/bin/env python
"""Throughput of the HOST side of the path (no GPU involved): BAM -> SAM text (`tools/bin/rg_bamview`, the
`samtools view -F 3328` of a `.generator`) -> FASTQ (`PassThroughSamCheck`, `PassThroughSamCheck.stranded`), next to
the reference's own PassThrough binaries (oracle/_ref) on the same SAM text; the FASTQ scan of `jellyfish count`
(tools/rg_cli.hpp:read_sequences) is timed by tools/micro/parse_bench.cpp. Measurement driver, not product code.

    python tools/host_ingest_bench.py [pairs] > profiles/r2_host_ingest.txt

Writes a synthetic coordinate-sorted paired-end BAM (150 bp reads, mates ~300 records apart, BGZF blocks of
64 KiB, zlib level 1) to a temp directory and times each stage reading from the page cache and writing to /dev/null
or to temp files."""
import os
import shutil
import struct
import subprocess
import sys
import tempfile
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tools", "bin")
REF = os.path.join(ROOT, "oracle", "_ref")


def make_bam(path, pairs, seed=5):
    rng = np.random.default_rng(seed)
    n = 2 * pairs
    pos1 = np.sort(rng.integers(0, 200_000_000, pairs)).astype(np.int32)
    pos = np.concatenate([pos1, pos1 + 320])
    mate = np.concatenate([np.zeros(pairs, np.uint8), np.ones(pairs, np.uint8)])
    pair_id = np.concatenate([np.arange(pairs), np.arange(pairs)])
    order = np.argsort(pos, kind="stable")
    pos, mate, pair_id = pos[order], mate[order], pair_id[order]
    L, NL = 150, 10                                       # read length, name length + NUL ("r%08d\0")
    rec = np.zeros((n, 4 + 32 + NL + 4 + (L + 1) // 2 + L + 4), np.uint8)   # + one tag, NM:i:0 (the reference tools need a field after QUAL)
    size = rec.shape[1] - 4
    fixed = np.zeros(n, dtype=[("bs", "<i4"), ("ref", "<i4"), ("pos", "<i4"), ("lname", "u1"), ("mapq", "u1"), ("bin", "<u2"),
                               ("ncig", "<u2"), ("flag", "<u2"), ("lseq", "<i4"), ("nref", "<i4"), ("npos", "<i4"), ("tlen", "<i4")])
    fixed["bs"], fixed["ref"], fixed["pos"], fixed["lname"], fixed["mapq"], fixed["bin"] = size, 0, pos, NL, 60, 4680
    fixed["ncig"], fixed["lseq"], fixed["nref"] = 1, L, 0
    fixed["flag"] = np.where(mate == 0, 99, 147)
    fixed["npos"] = np.where(mate == 0, pos + 320, pos - 320)
    fixed["tlen"] = np.where(mate == 0, 470, -470)
    rec[:, :36] = fixed.view(np.uint8).reshape(n, 36)
    names = np.char.add("r", np.char.zfill(pair_id.astype(str), 8)).astype("S9")
    rec[:, 36:36 + 9] = names.view(np.uint8).reshape(n, 9)
    rec[:, 46:50] = np.frombuffer(struct.pack("<I", (L << 4) | 0), np.uint8)
    nib = np.array([1, 2, 4, 8], np.uint8)[rng.integers(0, 4, (n, L))]
    rec[:, 50:50 + L // 2] = (nib[:, 0::2] << 4) | nib[:, 1::2]
    rec[:, 50 + L // 2:50 + L // 2 + L] = rng.integers(2, 41, (n, L), dtype=np.uint8)
    rec[:, -4:] = np.frombuffer(b"NMC\0", np.uint8)
    text = "@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:chr1\tLN:248956422\n"
    body = b"BAM\x01" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", 1) + struct.pack("<i", 5) + b"chr1\0" + \
        struct.pack("<i", 248956422) + rec.tobytes()
    with open(path, "wb") as f:
        for p in list(range(0, len(body), 65280)) + [None]:
            chunk = b"" if p is None else body[p:p + 65280]
            c = zlib.compressobj(1, zlib.DEFLATED, -15)
            comp = c.compress(chunk) + c.flush()
            f.write(b"\x1f\x8b\x08\x04" + b"\0" * 4 + b"\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, 12 + 6 + len(comp) + 8 - 1))
            f.write(comp + struct.pack("<II", zlib.crc32(chunk), len(chunk)))
    return len(body)


def timed(cmd, cwd, reps=3):
    best = 1e9
    for _ in range(reps):
        t = time.time()
        subprocess.run(cmd, shell=True, check=True, cwd=cwd, executable="/bin/bash")
        best = min(best, time.time() - t)
    return best


def main():
    pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    d = tempfile.mkdtemp(prefix="rg_ingest_")
    try:
        raw = make_bam(os.path.join(d, "x.bam"), pairs)
        bam = os.path.getsize(os.path.join(d, "x.bam"))
        bases = 2 * pairs * 150
        print(f"# host ingest, {os.cpu_count()} cores (build container), synthetic BAM: {pairs} pairs x 2 x 150 bp = {bases / 1e6:.0f} Mbases, "
              f"{bam / 1e6:.0f} MB BAM ({raw / 1e6:.0f} MB inflated), best of 3 runs, page cache warm")
        subprocess.run(f"{BIN}/rg_bamview view -F 3328 -@ 8 x.bam > x.sam", shell=True, check=True, cwd=d)
        sam = os.path.getsize(os.path.join(d, "x.sam"))
        print(f"# SAM text: {sam / 1e6:.0f} MB")
        print("stage\tseconds\tMbases/s\tGB/s of SAM text")
        rows = []
        for t in (1, 2, 4, 8):
            rows.append((f"rg_bamview -F 3328 -@ {t} x.bam > /dev/null", timed(f"{BIN}/rg_bamview view -F 3328 -@ {t} x.bam > /dev/null", d)))
        for t in (1, 8):
            rows.append((f"  same pipeline with zlib's inflate (RG_BAMVIEW_ZLIB=1), -@ {t}",
                         timed(f"RG_BAMVIEW_ZLIB=1 {BIN}/rg_bamview view -F 3328 -@ {t} x.bam > /dev/null", d)))
        old = os.environ.get("RG_BAMVIEW_OLD")     # a build of an earlier tools/bamview.cpp, for the before/after rows
        if old:
            for t in (1, 8):
                rows.append((f"  rg_bamview before the pipeline rewrite (batch-synchronous, zlib), -@ {t}",
                             timed(f"{old} view -F 3328 -@ {t} x.bam > /dev/null", d)))
        for name, exe in (("ours", BIN), ("reference", REF)):
            if not os.path.exists(os.path.join(exe, "PassThroughSamCheck")):
                continue
            rows.append((f"PassThroughSamCheck ({name}) < x.sam > /dev/null", timed(f"{exe}/PassThroughSamCheck chr.{name} < x.sam > /dev/null", d)))
            rows.append((f"PassThroughSamCheck.stranded ({name}) < x.sam -> 2 FASTQ files",
                         timed(f"{exe}/PassThroughSamCheck.stranded chr.{name} st.{name} < x.sam", d)))
        rows.append(("rg_bamview -@ 8 | PassThroughSamCheck (ours) > /dev/null",
                     timed(f"{BIN}/rg_bamview view -F 3328 -@ 8 x.bam | {BIN}/PassThroughSamCheck chr.p > /dev/null", d)))
        rows.append(("rg_bamview -@ 8 | PassThroughSamCheck.stranded (ours) -> 2 FASTQ files",
                     timed(f"{BIN}/rg_bamview view -F 3328 -@ 8 x.bam | {BIN}/PassThroughSamCheck.stranded chr.q st.q", d)))
        # the same BAM decoded in process by `jellyfish count X.bam` (tools/rg_bam_input.hpp), up to the bytes handed to the counter
        for t in (1, 8):
            rows.append((f"BAM -> counter input in process (rg_input_to_seq -t {t} x.bam > /dev/null)", timed(f"{BIN}/rg_input_to_seq -t {t} x.bam > /dev/null", d)))
        rows.append(("  the chain up to the same bytes: rg_bamview -@ 8 | PassThroughSamCheck | rg_input_to_seq /dev/stdin",
                     timed(f"{BIN}/rg_bamview view -F 3328 -@ 8 x.bam | {BIN}/PassThroughSamCheck chr.r | {BIN}/rg_input_to_seq /dev/stdin > /dev/null", d)))
        # the paired input of RUFUS.Filter, up to the batch arrays handed to rg_filter_pairs (rg_pairs_check -n: no files written)
        subprocess.run(f"{BIN}/PassThroughSamCheck.stranded chr.s pairs < x.sam", shell=True, check=True, cwd=d)
        rows.append(("read pull input, the chain: rg_bamview -@ 8 | PassThroughSamCheck.stranded -> 2 FIFOs -> RUFUS.Filter's reader",
                     timed(f"rm -f q.mate1.fastq q.mate2.fastq; mkfifo q.mate1.fastq q.mate2.fastq; ({BIN}/rg_bamview view -F 3328 -@ 8 x.bam | "
                           f"{BIN}/PassThroughSamCheck.stranded chr.t q > /dev/null &) ; {BIN}/rg_pairs_check -n q.mate1.fastq q.mate2.fastq qq > /dev/null", d)))
        for t in (1, 8):
            rows.append((f"read pull input in process (RUFUS.Filter HashList x.bam x.bam ...: rg_pairs_check -n -t {t} x.bam)",
                         timed(f"{BIN}/rg_pairs_check -n -t {t} x.bam pp > /dev/null", d)))
        for name, s in rows:
            print(f"{name}\t{s:.2f}\t{bases / 1e6 / s:.0f}\t{sam / 1e9 / s:.2f}")
        # FASTQ-mode generator line (runRufus.sh:159-161): perl scripts/FastqToSam.pl vs tools/bin/rg_fastq_to_sam
        subprocess.run(f"{BIN}/PassThroughSamCheck chr.f < x.sam > x.fastq", shell=True, check=True, cwd=d)
        fq = os.path.getsize(os.path.join(d, "x.fastq"))
        print(f"# FASTQ -> SAM text ({fq / 1e6:.0f} MB FASTQ, {bases / 1e6:.0f} Mbases)")
        t = timed(f"{BIN}/rg_fastq_to_sam x.fastq > /dev/null", d)
        print(f"rg_fastq_to_sam x.fastq > /dev/null\t{t:.2f}\t{bases / 1e6 / t:.0f}\t-")
        pl = "/root/reference/scripts/FastqToSam.pl"
        if os.path.exists(pl) and shutil.which("perl"):
            t = timed(f"perl {pl} x.fastq > /dev/null", d)
            print(f"perl scripts/FastqToSam.pl x.fastq > /dev/null (reference)\t{t:.2f}\t{bases / 1e6 / t:.0f}\t-")
            same = subprocess.run(f"cmp <({BIN}/rg_fastq_to_sam x.fastq) <(perl {pl} x.fastq)", shell=True, cwd=d, executable="/bin/bash").returncode == 0
            print(f"# identical output: {same}")
        for th in (1, 2, 4, 6):
            t = timed(f"{BIN}/rg_input_to_seq -t {th} x.fastq > /dev/null", d)
            print(f"FASTQ file -> counter input (`jellyfish count -t {th} x.fastq`: rg_input_to_seq -t {th})\t{t:.2f}\t{bases / 1e6 / t:.0f}\t-")
        t = timed(f"{BIN}/rg_fastq_to_sam x.fastq | {BIN}/PassThroughSamCheck chr.g > /dev/null", d)
        print(f"rg_fastq_to_sam x.fastq | PassThroughSamCheck (ours) > /dev/null\t{t:.2f}\t{bases / 1e6 / t:.0f}\t-")
        same = all(open(os.path.join(d, f"st.ours.mate{m}.fastq"), "rb").read() == open(os.path.join(d, f"st.reference.mate{m}.fastq"), "rb").read()
                   for m in (1, 2)) if os.path.exists(os.path.join(d, "st.reference.mate1.fastq")) else None
        print(f"# re-paired FASTQ files of ours and the reference identical: {same}")
    finally:
        shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
