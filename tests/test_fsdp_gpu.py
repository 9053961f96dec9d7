"""BASELINE.json config 4 (large under FSDP FULL_SHARD + per-block activation checkpointing,
scripts/training/train_fsdp_timestamps.py:2588-2615,2665-2678,2711-2719) needs a process per GPU: the checks live in
tools/fsdp_check.py and are launched here with torchrun when at least two GPUs are visible."""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (one process per GPU)")
def test_fsdp_wrapped_model_matches_the_plain_model():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29517", str(ROOT / "tools" / "fsdp_check.py"), "parity"], capture_output=True, text=True,
                         timeout=600, env=env, cwd=ROOT)
    print(out.stdout[-3000:], out.stderr[-3000:])
    assert out.returncode == 0
    # two ranks print concurrently (lines may interleave): 3 checks x 2 ranks must all say PASS
    assert out.stdout.count(" PASS ") == 6 and "FAIL" not in out.stdout, out.stdout[-2000:]
