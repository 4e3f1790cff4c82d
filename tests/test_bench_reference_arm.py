"""`bench.py --impl reference` (the arm the driver runs next to the CUDA arm) on a tiny workload, on the CPU: one JSON
line with the contract's keys, the reference's own binaries behind it, the digests of its outputs attached, and ranks
other than 0 leaving without work under torchrun's environment."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

REF = os.path.join(ROOT, "oracle", "_ref", "jellyfish")


def _run(extra_env=None, *args):
    env = dict(os.environ, **(extra_env or {}))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--pairs", "3000", "--genome",
                           "30000", "--steps", "1", "--warmup", "0", *args], capture_output=True, text=True, env=env, timeout=600)


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref not built (needs /root/reference)")
def test_reference_arm_prints_one_contract_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    o = json.loads(lines[0])
    assert o["impl"] == "reference" and o["n_gpus"] == 1 and o["steps"] == 1 and o["warmup"] == 0
    assert o["unit"] == "Gbases/s" and o["higher_is_better"] is True and o["value"] > 0
    assert o["cpu_baseline"]["kind"] == "reference" and o["cpu_baseline"]["value"] == o["value"]
    assert o["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert o["e2e"] == {"value": o["value"], "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert o["gpu_launches"] == 0
    # the whole (tiny) workload was run, and said so; its outputs' digests are attached (no committed golden for this size)
    assert "the whole workload: 3000 pairs/sample" in o["config"]["reference_sample"]
    p = o["parity"]
    assert p["ok"] is None and len(p["hashlist_md5"]) == 32 and p["hashlist_lines"] >= 0 and p["pulled_pairs"] >= 0


def test_reference_arm_other_ranks_leave_without_work():
    r = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_cpu_binding_logic_with_a_stand_in_nvml():
    """bench.bind_to_gpu_cpus against a fake pynvml (the build container has no NVML): binds to the GPU's CPUs inside the
    cpuset, keeps the affinity when fewer than four of them are allowed, never raises"""
    code = r'''
import importlib.util, os, sys, types
allowed = sorted(os.sched_getaffinity(0))
fake = types.ModuleType("pynvml")
fake.nvmlInit = lambda: None
fake.nvmlDeviceGetHandleByIndex = lambda i: i
mask = {"v": 0}
fake.nvmlDeviceGetCpuAffinity = lambda h, n: [(mask["v"] >> (64 * w)) & (2 ** 64 - 1) for w in range(n)]
sys.modules["pynvml"] = fake
sys.argv = ["bench.py"]
spec = importlib.util.spec_from_file_location("bench", sys.argv[1] if False else os.environ["BENCH_PY"])
b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
if len(allowed) >= 5:
    want = allowed[:4]
    mask["v"] = sum(1 << c for c in want)
    print(b.bind_to_gpu_cpus(0)); assert sorted(os.sched_getaffinity(0)) == want
    mask["v"] = sum(1 << c for c in want[:2])
    print(b.bind_to_gpu_cpus(0)); assert sorted(os.sched_getaffinity(0)) == want        # two CPUs: kept
def boom(h, n): raise RuntimeError("no nvml")
fake.nvmlDeviceGetCpuAffinity = boom
print(b.bind_to_gpu_cpus(0))
print("OK")
'''
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, BENCH_PY=os.path.join(ROOT, "bench.py")))
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr
    assert "kept (RuntimeError)" in r.stdout
