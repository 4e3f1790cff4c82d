#!/usr/bin/env python
"""Golden outputs of the REFERENCE ModelDist (oracle/_ref/ModelDist, built from /root/reference/src/ModelDist.cpp
by oracle/Makefile) on k-mer histograms: the three samples of the bundled trio (`jellyfish count -m 25 -L 2` ->
`histo -f` -> the tab conversion of runRufus.sh:830, all with the reference jellyfish; SURVEY §8c quotes the
`.model` head `3, 5, 17080, 28` for the Child) and four synthetic histograms (Gaussian copy-number peaks + a het
peak + a 1/x^2.5 error tail, Poisson noise, fixed seeds; the files themselves are committed so numpy's generator
is not part of the fixture). Stores the histograms (gz), md5 digests of the three files and of stdout, the four
head lines of every `.model`, and the Child's whole `.model` file. Run in the build container; ~3 minutes."""
import gzip, hashlib, json, os, shutil, subprocess, tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFD = os.path.join(HERE, "..", "..", "oracle", "_ref")
OUT = os.path.join(HERE, "modeldist")
os.makedirs(OUT, exist_ok=True)


def md5(b):
    return hashlib.md5(b).hexdigest()


def synth(cov, genome, het, err, seed, lo=1, sd_scale=1.3, n=10001):
    rng = np.random.default_rng(seed)
    x = np.arange(n + 1)
    g = lambda mu, sd, w: w * np.exp(-(x - mu) ** 2 / (2 * sd * sd)) / (sd * np.sqrt(2 * np.pi))
    sd = np.sqrt(cov) * sd_scale
    h = g(cov, sd, genome) + g(cov / 2, sd * 0.75, het) + g(2 * cov, sd * 1.5, genome * 0.05) + g(3 * cov, sd * 1.9, genome * 0.01)
    h[1:] += err / np.arange(1, n + 1) ** 2.5
    h = rng.poisson(h).astype(np.int64)
    h[:lo] = 0
    h[n] = rng.integers(0, 50)
    return "".join(f"{i}\t{h[i]}\n" for i in range(n + 1)).encode()


def trio_histo(sample):
    with tempfile.TemporaryDirectory() as d:
        for m in (1, 2):
            open(os.path.join(d, f"{m}.fq"), "wb").write(gzip.open(os.path.join(HERE, "testRun", f"{sample}.mate{m}.fastq.gz")).read())
        jf = os.path.join(REFD, "jellyfish")
        subprocess.run([jf, "count", "-m", "25", "-L", "2", "-s", "16M", "-t", "4", "-o", "x.Jhash", "-C", "1.fq", "2.fq"], cwd=d, check=True)
        subprocess.run([jf, "histo", "-f", "-o", "x.histo", "x.Jhash"], cwd=d, check=True)
        return open(os.path.join(d, "x.histo"), "rb").read().replace(b" ", b"\t")


CASES = {  # name: (histogram bytes, K, read length, threads argument)
    "Child": (lambda: trio_histo("Child"), 25, 150, 8),
    "Mother": (lambda: trio_histo("Mother"), 25, 150, 4),
    "Father": (lambda: trio_histo("Father"), 25, 150, 1),
    "s30": (lambda: synth(30, 3e6, 2e4, 5e6, 1), 25, 150, 8),
    "s45": (lambda: synth(45.4, 2.5e9, 3e6, 4e9, 2, lo=2), 25, 150, 8),
    "s18": (lambda: synth(18, 5e5, 0, 2e5, 3, lo=1, sd_scale=1.0), 31, 100, 3),
    "s60": (lambda: synth(61.7, 4e7, 4e6, 9e7, 4, lo=3, sd_scale=1.6), 21, 250, 16),
}
gold = {"generator": "tests/golden/make_modeldist_golden.py", "cases": {}}
for name, (make, K, L, T) in CASES.items():
    path = os.path.join(OUT, name + ".histo.gz")
    if os.path.exists(path):
        histo = gzip.open(path).read()
    else:
        histo = make()
        with gzip.GzipFile(path, "wb", mtime=0) as f:
            f.write(histo)
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "h"), "wb").write(histo)
        r = subprocess.run([os.path.join(REFD, "ModelDist"), "h", str(K), str(L), str(T)], cwd=d, check=True, capture_output=True)
        model = open(os.path.join(d, "h.7.7.model"), "rb").read()
        gold["cases"][name] = {
            "K": K, "L": L, "T": T, "histo_md5": md5(histo),
            "model_md5": md5(model), "dist_md5": md5(open(os.path.join(d, "h.7.7.dist"), "rb").read()),
            "prob_md5": md5(open(os.path.join(d, "h.7.7.prob"), "rb").read()), "stdout_md5": md5(r.stdout),
            "model_head": model.decode().split("\n")[:4],
            "best_model": [l for l in r.stdout.decode().split("\n") if l.startswith("Best Model")][0],
        }
        if name == "Child":
            open(os.path.join(OUT, "Child.7.7.model"), "wb").write(model)
    print(name, gold["cases"][name]["model_head"], gold["cases"][name]["best_model"])
json.dump(gold, open(os.path.join(OUT, "golden.json"), "w"), indent=1)
