"""N=2 parity of the hash-prefix-sharded pipeline (NCCL all-to-all); skipped on 1-GPU boxes."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("exchange", ["p2p", "nccl", "p2p+filter-a2a", "p2p+slow-tail"])
def test_two_gpu_sharded_pipeline_matches_oracle(tmp_path, exchange):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    env = dict(os.environ, RG_TMP=str(tmp_path), RG_EXCHANGE=exchange.split("+")[0])
    if exchange.endswith("filter-a2a"):          # SURVEY 8e layout (2): sharded mutant set, reads exchanged, hit counts all-reduced
        env["RG_FILTER_EXCHANGE"] = "a2a"
    if exchange.endswith("slow-tail"):           # the step-by-step tail instead of the single-sync one
        env["RG_PIPELINE_SLOW"] = "1"
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tools", "multigpu_check.py")], env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0 and "MULTI-GPU PARITY OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_jellyfish_count_on_two_gpus_writes_the_single_gpu_jhash(tmp_path):
    """`jellyfish count` with every GPU behind it (one process, one shard per GPU, peer-to-peer scatter, tables that
    grow, several rounds): the .Jhash must be byte-identical to the one-GPU run's, header aside (it records argv)."""
    import hashlib
    import numpy as np
    import torch
    from rufus_b200 import synth
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    trio, _ = synth.make_trio(23, 400_000, 60_000)
    s1, q1, s2, q2 = trio["Child"]
    s1[:2000] = ord("A")                       # heavy hitters: one minimizer owns far more than a slab
    fq = tmp_path / "reads.fq"
    fq.write_bytes(synth.fastq_bytes(s1, q1, prefix="a") + synth.fastq_bytes(s2, q2, prefix="b"))
    jf = os.path.join(ROOT, "tools", "bin", "jellyfish")
    outs = {}
    for name, devs in (("one", "0"), ("two", "0,1")):
        out = tmp_path / f"{name}.Jhash"
        env = dict(os.environ, RG_DEVICES=devs, RG_CHUNK_MB="2")
        r = subprocess.run([jf, "count", "--disk", "-m", "25", "-L", "2", "-s", "16M", "-t", "4", "-o", str(out), "-C", str(fq)],
                           env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        raw = out.read_bytes()
        hlen = int(raw[:9])
        outs[name] = raw[9 + hlen:]             # the records, in (pos, key) order
    assert len(outs["one"]) > 1_000_000
    assert hashlib.md5(outs["one"]).hexdigest() == hashlib.md5(outs["two"]).hexdigest()
