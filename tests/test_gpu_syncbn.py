"""GPU, >= 2 devices: fused MinkowskiSyncBatchNorm with the exchange INSIDE the statistics tail kernel (NVLink peer
memory, csrc/bn.cu tail_exchange / csrc/peer.cu) against torch.nn.SyncBatchNorm evaluated in float64.

Runs tools/syncbn_check.py under torchrun in a subprocess (one process per GPU, as train_lidog.py:227-231 runs): a
protocol bug that trapped a kernel would take only that subprocess down, not the test session.  Skipped on a
single-GPU box; the round's multi-GPU runs keep its output under profiles/."""
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
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2])
def test_syncbn_peer_exchange_matches_torch_syncbatchnorm(cuda, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, {torch.cuda.device_count()} visible")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "syncbn_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "SYNCBN OK" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])
    assert "peer memory (in-kernel)" in r.stdout, r.stdout[-2000:]  # the NVLink exchange ran, not the NCCL fallback
