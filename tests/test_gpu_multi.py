"""Two (or more) GPUs in one box: the sharded frame -- peer-store path and NCCL all-gather path -- must equal the
single-GPU frame bit for bit.  Skipped on single-GPU boxes; the host-side sharding logic is covered
on CPU by tests/test_distributed_cpu.py."""
import os
import subprocess
import sys

import pytest
import torch

from helpers import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("p2p", ["1", "0"])
def test_two_rank_frame_is_bit_identical(p2p):
    env = dict(os.environ, GRRT_P2P=p2p)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29571" if p2p == "1" else "29572", os.path.join(ROOT, "tools", "mgpu_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("bit-identical to 1-GPU frame: True") == 2
    assert ("peer-store path" in r.stdout) == (p2p == "1")
