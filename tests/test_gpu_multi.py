"""GPU, more than one device: data-parallel gradient parity and ensemble sharding.  Each case
launches tests/dist_grad_parity_worker.py under torchrun with NCCL; cases needing more devices than the box has are
skipped (the round driver's 1-GPU box skips all of them; run with `gpurun --gpus N`)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("world", [2, 4, 8])
def test_n_gpu_gradients_equal_single_gpu_global_batch(world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_grad_parity_worker.py")]
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "DIST_GRAD_PARITY_OK world=%d" % world in res.stdout, res.stdout[-2000:]
