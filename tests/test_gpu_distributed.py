"""Multi-GPU functional check as a pytest module: spawns tests/dist_gpu_check.py under torch.distributed.run (NCCL, one
process per GPU) when the box has >= 2 GPUs and skips otherwise.  What the ranks assert is in dist_gpu_check.py: the sharded
loss equals the oracle and the single-GPU value and is bit-identical on every rank, each shard's gradients are bit-identical to
the slices of the full-batch gradients, gathered detections equal the full-batch detections.  The log is kept under
gpurun_out/ (copied to profiles/ when it comes from a GPU-box run)."""
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
def test_sharded_loss_and_detections_on_gpus(world):
    n = torch.cuda.device_count()
    if n < world:
        pytest.skip("needs %d GPUs, this box has %d" % (world, n))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_gpu_check.py")]
    res = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "dist_gpu_check_world%d.log" % world), "w") as f:
        f.write(res.stdout)
    assert res.returncode == 0, res.stdout[-4000:]
    assert "dist_gpu_check ok: world=%d" % world in res.stdout
