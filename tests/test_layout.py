"""Structural guarantees the task contract asks for, checked statically (CPU):
 * the oracle is test infrastructure: nothing under rufus_b200/ or tools/ imports, links or executes
   anything under oracle/; bench.py touches it only in its CPU-baseline / reference-arm code;
 * the product fails loudly without its CUDA library or without a CUDA device (no CPU fallback)."""
import os
import re
import subprocess
import sys

import pytest

from conftest import ROOT


def _sources(*dirs, exts):
    for d in dirs:
        for base, _, files in os.walk(os.path.join(ROOT, d)):
            if "bin" in base.split(os.sep) or "__pycache__" in base:
                continue
            for f in files:
                if f.endswith(exts):
                    yield os.path.join(base, f)


def test_product_code_never_touches_the_oracle():
    # (nor the recording stand-in of the C ABI that tests/abi_recorder.py builds for the host-logic tests)
    pat = re.compile(r"import\s+oracle|from\s+oracle|liboracle|bam_oracle|overlap_oracle|[\"'/]oracle[\"'/]|abi_recorder|RG_RECORD_DIR")
    bad = []
    for path in _sources("rufus_b200", "tools", "include", exts=(".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", "Makefile")):
        if os.path.basename(path) in ("overlap_bench.py", "multigpu_check.py", "debug_n2.py", "host_ingest_bench.py"):
            continue                                   # measurement / parity drivers, not product code
        for i, line in enumerate(open(path, errors="replace"), 1):
            if pat.search(line):
                bad.append(f"{os.path.relpath(path, ROOT)}:{i}: {line.strip()}")
    assert not bad, "\n".join(bad)


def test_bench_uses_the_oracle_only_for_the_cpu_baseline():
    """every mention of oracle/ in bench.py sits inside a function that times the REFERENCE'S side: cpu_reference_pass
    (cpu_baseline leg and --impl reference), cpu_reference_big (the large configs' cpu_baseline) and cli_leg (the
    reference's stage commands next to ours); have_reference_binaries only looks whether oracle/_ref was built. Nothing on
    the measured CUDA path touches it."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    allowed = ("cpu_reference_pass", "cpu_reference_big", "cli_leg", "have_reference_binaries")
    outside, pos = [], 0
    for m in re.finditer(r"^def (\w+)\(.*?(?=^def |\Z)", src, re.S | re.M):
        if m.group(1) in allowed:
            outside.append(src[pos:m.start()])
            pos = m.end()
    outside = "".join(outside) + src[pos:]
    assert "def main" in outside and "def run_big" in outside
    assert "import oracle" not in outside and '"oracle"' not in outside


def test_library_missing_fails_loudly(tmp_path):
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from rufus_b200 import api\n"
            "api.LIB_PATH = %r\n"
            "try:\n    api.lib()\nexcept api.RufusGpuError as e:\n    print('LOUD:', e)\n") % (ROOT, str(tmp_path / "nope.so"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert "LOUD:" in r.stdout and "no CPU fallback" in r.stdout


def test_no_device_fails_loudly():
    """rg_init without a usable CUDA device returns NULL with a message (the build container has none;
    on a GPU box the device is hidden through CUDA_VISIBLE_DEVICES)."""
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from rufus_b200 import api\n"
            "try:\n    api.Context(0)\n    print('UNEXPECTED: context created')\n"
            "except api.RufusGpuError as e:\n    print('LOUD:', e)\n") % ROOT
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    assert "LOUD:" in r.stdout, r.stdout + r.stderr
    # the executables do the same: message on stderr, non-zero exit for jellyfish
    jf = os.path.join(ROOT, "tools", "bin", "jellyfish")
    if os.path.exists(jf):
        r = subprocess.run([jf, "count", "-m", "25", "-s", "1M", "-C", "-o", "/dev/null", os.path.join(ROOT, "README.md")],
                           capture_output=True, text=True, env=env)
        assert r.returncode != 0 and r.stderr.strip()
