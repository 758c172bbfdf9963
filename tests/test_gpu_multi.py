"""The forecaster under an initialised NCCL process group (one process per GPU, launched like the driver launches
bench.py): eager and CUDA-graph launches interleaved with collectives, shard == slice of the whole batch bit for bit.
Runs on min(2, visible GPUs) ranks, so it also covers the single-GPU box."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_forecaster_under_nccl_group(cuda):
    import torch
    n = min(2, torch.cuda.device_count())
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29613", str(ROOT / "tests" / "_nccl_forecast.py")],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert f"NCCL_FORECAST_OK {n}" in r.stdout
