"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): see tests/multi_gpu_check.py"""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_two_rank_run_matches_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29631", str(ROOT / "tests" / "multi_gpu_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert "MULTI_GPU_CHECK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
