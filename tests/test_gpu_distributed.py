"""Partitioned (multi-GPU) path on the GPU box: the same driver with one rank (always)
and, when the box has two GPUs, two NCCL ranks launched with torchrun."""
import os
import subprocess
import sys

import numpy as np
import pytest

from cases import mms_1_mobile_steady
from oracle.newton import OracleBackend

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("dim,n", [(2, 20), (3, 8)])
def test_partitioned_driver_single_rank(dim, n):
    from festim_b200.distributed import DistributedBackend

    g, Hg, _ = mms_1_mobile_steady(dim, n)
    g.solver_backend = DistributedBackend()
    g.initialise()
    g.run()
    r, Hr, _ = mms_1_mobile_steady(dim, n)
    r.solver_backend = OracleBackend()
    r.initialise()
    r.run()
    a, b = Hg.post_processing_solution.x.array, Hr.post_processing_solution.x.array
    assert g.solver.solver.getConvergedReason() > 0
    assert g.solver.solver.getIterationNumber() == r.solver.solver.getIterationNumber()
    assert np.linalg.norm(a - b) / np.linalg.norm(b) < 1e-8


@pytest.mark.parametrize("peer", [True, False])
def test_partitioned_two_ranks(peer):
    """peer=True: whole-solve persistent CG over NVLink peer memory (default);
    peer=False: per-iteration NCCL halo exchange + all-reduce (FB_NO_PEER_CG=1)."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29517",
           os.path.join(ROOT, "tests", "dist_worker.py"), "14" if peer else "10"]
    env = dict(os.environ)
    env.pop("FB_NO_PEER_CG", None)
    if not peer:
        env["FB_NO_PEER_CG"] = "1"
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "ok=True" in out.stdout
    assert f"peer={peer}" in out.stdout
