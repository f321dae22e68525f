"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): tests/drivers/dist_parity.py under
torchrun, with the persistent fused GMRES kernel (the default from 2 ranks on), with the per-iteration kernels over peer
memory (ZGEMV epilogue -> every rank's work vector) and over the NCCL all-gather fallback.  All must agree with the
oracle and be bit-identical across ranks."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("peer,fused", [("1", None), ("1", "0"), ("0", None)])
def test_row_sharded_parity_two_ranks(peer, fused):
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, BEMB200_PEER_FUSED=peer)
    if fused is not None:
        env["BEMB200_GMRES_FUSED"] = fused
    port = 29600 + (os.getpid() % 200) + (1 if peer == "1" else 0) + (2 if fused else 0)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "drivers", "dist_parity.py")]
    r = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("OK") >= 2
