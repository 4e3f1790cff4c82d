"""N=2 parity of the hash-prefix-sharded pipeline (NCCL all-to-all); skipped on 1-GPU boxes."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_two_gpu_sharded_pipeline_matches_oracle(tmp_path, exchange):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    env = dict(os.environ, RG_TMP=str(tmp_path), RG_EXCHANGE=exchange)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tools", "multigpu_check.py")], env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0 and "MULTI-GPU PARITY OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
