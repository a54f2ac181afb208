"""Sharded make_graph / run_mapping == single GPU, bit for bit (needs >= 2 GPUs; the world-2 gloo test in
tests/test_host.py covers the host-side sharding logic on CPU)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_equals_single_gpu():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29517",
                        os.path.join(ROOT, "tools", "multigpu_check.py")], capture_output=True, text=True, timeout=600)
    assert "MULTIGPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
