"""GPU, >= 2 devices: the data-parallel step against the CPU oracle, torch DDP and the NCCL baseline (tests/dp_check.py
under torchrun, one rank per GPU).  Skipped on a single-GPU box; run it with `gpurun --gpus 2`."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_data_parallel_step_against_oracle_ddp_and_nccl():
    world = 2
    port = 29500 + os.getpid() % 400
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dp_check.py")]
    env = dict(os.environ)
    env.setdefault("ECGB200_SPIN_TIMEOUT_MS", "10000,120000")       # a protocol bug must end in a trap, not a hung box
    res = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    print(res.stdout[-6000:])
    print(res.stderr[-3000:])
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "dp_check passed" in res.stdout
    for tag in ("1. kernel level (one bucket)", "1. kernel level (two buckets)", "1. kernel level (two buckets, one-hop words)", "3. oracle, local BN (cnn)",
                "3. oracle, local BN (mm)", "3b. torch DDP", "4. oracle, SyncBN (cnn)", "4. oracle, SyncBN (mm)"):
        assert tag in res.stdout, tag
