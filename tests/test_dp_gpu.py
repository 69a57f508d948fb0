"""N ranks == 1 rank on the concatenated batch, on real GPUs through the engine and NCCL (SURVEY.md section 4 (vi);
apex/parallel/distributed.py:L491-510).  Launches tools/dp_check.py under torch.distributed.run with two ranks; skipped
on a single-GPU box (the host-side exchange logic is covered on CPU by tests/test_parallel_gloo.py)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_ranks_equal_one_rank_on_concatenated_batch():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "dp_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0, r.stderr[-4000:]
    assert "dp_check ok" in r.stdout
    assert "uc2 N=2 vs 1, step 1" in r.stdout and "m3p N=2 vs 1, step 1" in r.stdout
