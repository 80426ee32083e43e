"""Multi-GPU paths on real devices (skipped on a one-GPU box; the CPU coverage of the same
logic is tests/test_sharded_gloo.py): the view-sharded library over NCCL and over the
device-resident NVLink exchange, identical to the unsharded engine; ShardedStepper with an
engine on its own stream; the missing-peer timeout."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_view_sharded_two_gpus(gpu):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (covered by gpurun --gpus 2 runs, profiles/r02_sharded_2gpu.log)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tools", "sharded_nccl_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert "MISMATCH" not in r.stdout and "WRONG" not in r.stdout
