"""-m gpu, needs 2 GPUs: SPP-PPO data-parallel iteration over NCCL (tools/ppo_dp.py) against the same iteration on one GPU:
all ranks end bit-identical, post-iteration weights within 2e-5 of the single-GPU run."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_ppo_data_parallel_matches_single_gpu():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "ppo_dp.py"), "--envs", "64", "--steps", "128", "--batch", "2048"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["ranks_bit_identical"] and out["dp_vs_single_worst_relnorm"] < 2e-5
