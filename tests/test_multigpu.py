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


def test_sharded_two_ranks_on_one_gpu_equal_single_rank():
    """The same check on ONE GPU: two ranks share cuda:0 and exchange over gloo (host-staged collectives), so the
    sharded code path -- shard-local kernels, int64 Gram all-reduce, embedding all-gather, sharded DataStore writes --
    is compared bit for bit with the single-rank result on a box with a single GPU."""
    env = dict(os.environ, MG_ONE_GPU="1", MG_CELLS="11500")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29519",
                        os.path.join(ROOT, "tools", "multigpu_check.py")], capture_output=True, text=True, timeout=900,
                       env=env)
    with open(os.path.join(ROOT, "gpurun_out", "multigpu_one_gpu.log") if os.path.isdir(os.path.join(ROOT, "gpurun_out"))
              else os.devnull, "w") as f:
        f.write(r.stdout[-5000:] + "\n" + r.stderr[-5000:])
    assert "MULTIGPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
