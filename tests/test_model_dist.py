"""tools/bin/ModelDist (the drop-in for the reference's copy-number model fit, src/ModelDist.cpp:375-973,
called at runRufus.sh:849-867) against what the REFERENCE binary wrote for the same histograms
(tests/golden/modeldist, made by make_modeldist_golden.py with oracle/_ref/ModelDist): the `.model`, `.dist`
and `.prob` files and the stdout log, byte for byte. Host-only FP64 work: no GPU needed. What runRufus.sh
consumes are lines 2 and 4 of the `.model` file (MutantMinCov, MutantSC): `3, 5, 17080, 28` for the bundled
Child sample."""
import gzip
import hashlib
import json
import os
import subprocess

import pytest

from conftest import GOLDEN, ROOT

EXE = os.path.join(ROOT, "tools", "bin", "ModelDist")
REF = os.path.join(ROOT, "oracle", "_ref", "ModelDist")
GOLD = json.load(open(os.path.join(GOLDEN, "modeldist", "golden.json")))["cases"]


def md5(b):
    return hashlib.md5(b).hexdigest()


def run(exe, tmp_path, histo, K, L, T):
    (tmp_path / "h").write_bytes(histo)
    r = subprocess.run([exe, "h", str(K), str(L), str(T)], cwd=tmp_path, capture_output=True, check=True)
    return {e: (tmp_path / f"h.7.7.{e}").read_bytes() for e in ("model", "dist", "prob")}, r.stdout


@pytest.fixture(scope="module")
def exe():
    if not os.path.exists(EXE):
        pytest.fail("tools/bin/ModelDist is missing: run `make tools`")
    return EXE


@pytest.mark.parametrize("name", sorted(GOLD))
def test_model_files_and_log_equal_the_reference_binary(exe, tmp_path, name):
    g = GOLD[name]
    histo = gzip.open(os.path.join(GOLDEN, "modeldist", name + ".histo.gz")).read()
    assert md5(histo) == g["histo_md5"]
    files, out = run(exe, tmp_path, histo, g["K"], g["L"], g["T"])
    assert files["model"].decode().split("\n")[:4] == g["model_head"]
    assert [l for l in out.decode().split("\n") if l.startswith("Best Model")] == [g["best_model"]]
    assert md5(files["model"]) == g["model_md5"]
    assert md5(files["dist"]) == g["dist_md5"]
    assert md5(files["prob"]) == g["prob_md5"]
    assert md5(out) == g["stdout_md5"]


def test_bundled_child_model_is_the_surveyed_one(exe, tmp_path):
    histo = gzip.open(os.path.join(GOLDEN, "modeldist", "Child.histo.gz")).read()
    files, _ = run(exe, tmp_path, histo, 25, 150, 2)
    assert files["model"] == open(os.path.join(GOLDEN, "modeldist", "Child.7.7.model"), "rb").read()
    head = files["model"].decode().split("\n")[:4]
    assert head == ["3", "5", "17080", "28"]  # MutantMinCov 5, MutantSC 28 -> MaxHashDepth 140 (runRufus.sh:862-867)


@pytest.mark.parametrize("threads", [1, 3, 11, 64])
def test_result_does_not_depend_on_the_thread_count(exe, tmp_path, threads):
    g = GOLD["s18"]
    histo = gzip.open(os.path.join(GOLDEN, "modeldist", "s18.histo.gz")).read()
    files, out = run(exe, tmp_path, histo, g["K"], g["L"], threads)
    assert (md5(files["model"]), md5(files["dist"]), md5(files["prob"]), md5(out)) == \
        (g["model_md5"], g["dist_md5"], g["prob_md5"], g["stdout_md5"])


def test_empty_histogram_is_reported_like_the_reference(exe, tmp_path):
    (tmp_path / "h").write_text("".join(f"{i}\t0\n" for i in range(40)))
    r = subprocess.run([exe, "h", "25", "150", "2"], cwd=tmp_path, capture_output=True)
    assert r.returncode == 1
    assert r.stdout.decode().endswith("ERROR there are no kmers in this file\n")
    assert r.stdout.decode().count("getting another") == 11
    if os.path.exists(REF):
        q = subprocess.run([REF, "h", "25", "150", "2"], cwd=tmp_path, capture_output=True)
        assert (q.returncode, q.stdout) == (r.returncode, r.stdout)


def test_missing_histogram(exe, tmp_path):
    r = subprocess.run([exe, "nope", "25", "150", "2"], cwd=tmp_path, capture_output=True)
    assert r.returncode == 0 and r.stdout == b"Call is histoFile HS ReadLength Threads\nError, HistoFile could not be opened"
    if os.path.exists(REF):
        q = subprocess.run([REF, "nope", "25", "150", "2"], cwd=tmp_path, capture_output=True)
        assert (q.returncode, q.stdout) == (r.returncode, r.stdout)


@pytest.mark.skipif(not os.path.exists(REF), reason="reference binary only exists in the build container")
def test_fresh_random_histogram_against_the_reference_binary(exe, tmp_path):
    """One histogram that is NOT a committed fixture: a different seed on every container build would be
    pointless (the fixture digests already pin seven), so the seed is fixed but the case is run live through
    both binaries (~20 s of reference time)."""
    import numpy as np
    rng = np.random.default_rng(77)
    n = 10001
    x = np.arange(n + 1)
    cov, sd = 37.3, 8.1
    h = 6e5 * np.exp(-(x - cov) ** 2 / (2 * sd * sd)) + 4e4 * np.exp(-(x - cov / 2) ** 2 / (2 * (sd * 0.7) ** 2)) \
        + 3e4 * np.exp(-(x - 2 * cov) ** 2 / (2 * (sd * 1.4) ** 2))
    h[1:] += 3e6 / np.arange(1, n + 1) ** 2.2
    h = rng.poisson(h)
    h[0] = 0
    histo = "".join(f"{i}\t{h[i]}\n" for i in range(n + 1)).encode()
    a, b = tmp_path / "ours", tmp_path / "ref"
    a.mkdir(), b.mkdir()
    assert run(exe, a, histo, 25, 150, 4) == run(REF, b, histo, 25, 150, 4)
