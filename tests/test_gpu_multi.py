"""Multi-GPU parity (needs >= 2 GPUs on the box): torchruns tests/multi_gpu_worker.py, whose every
assertion is against the CPU oracle -- the sharded (i,j) triangle with the in-kernel NVLink
exchange and with the ncclAllGather path, recompute and int32 matrix, n up to 100 000, and the
sharded multi-start population."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_paths_against_the_oracle(world):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs, this box has {_gpus()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    env = dict(os.environ, PYTHONFAULTHANDLER="1")  # a crash in a rank prints its Python stack
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0 and f"MULTI_GPU_OK world={world}" in r.stdout, (r.stdout[-3000:], r.stderr[-6000:])
