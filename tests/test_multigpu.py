"""Multi-GPU checks on real GPUs (skipped with fewer than 2): fused band exchange over peer memory, NCCL gather,
frame partition.  The host-side logic is also covered on CPU by test_parallel_gloo.py."""
import os
import subprocess
import sys

import pytest

from util import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 4, 8])
def test_band_exchange_and_frame_partition(world):
    import torch
    n = torch.cuda.device_count()
    if n < world:
        pytest.skip(f"needs {world} GPUs (run with gpurun --gpus {world})")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", str(29533 + world),
                          os.path.join(ROOT, "tests", "multigpu_band_worker.py")], env=env, capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0 and "MULTIGPU_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
