"""Partitioned path with a non-symmetric Jacobian (traps, advection): the distributed GMRES of
festim_b200/krylov_dist.py on top of the library's owned-row SpMV and block-Jacobi preconditioner,
against the oracle.  One rank always; two NCCL ranks when the box has two GPUs.  (Sorted after the
other GPU suites on purpose: the newest path runs last.)"""
import os
import subprocess
import sys

import numpy as np
import pytest

from cases import advection_trap_mms_2d, mms_1_mobile_1_trap_steady
from oracle.newton import OracleBackend

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("builder,args", [(mms_1_mobile_1_trap_steady, (2, 12)), (mms_1_mobile_1_trap_steady, (3, 5)),
                                          (advection_trap_mms_2d, (12,))],
                         ids=["trap2d", "trap3d", "advection2d"])
def test_partitioned_gmres_single_rank(builder, args):
    from festim_b200.distributed import DistributedBackend

    g = builder(*args)[0]
    g.solver_backend = DistributedBackend()
    g.initialise()
    g.run()
    r = builder(*args)[0]
    r.solver_backend = OracleBackend()
    r.initialise()
    r.run()
    assert g.solver.gmres_iterations > 0 and g.solver.cg_iterations == 0
    assert g.solver.solver.getConvergedReason() > 0
    assert g.solver.solver.getIterationNumber() == r.solver.solver.getIterationNumber()
    a, b = g.u.x.array, r.u.x.array
    assert np.linalg.norm(a - b) / np.linalg.norm(b) < 1e-8


@pytest.mark.parametrize("case,n", [("trap", 6), ("c2", 6)])
def test_partitioned_gmres_two_ranks(case, n):
    """Two ranks: the library's own Newton-GMRES on the partitioned mesh (halo exchange and cross-rank
    reductions over NVLink peer memory inside its kernels, no NCCL call in the solve) - block-CSR
    values (trap) and the channel operator with the Chebyshev preconditioner (c2) - against the oracle."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29519",
           os.path.join(ROOT, "tests", "dist_worker.py"), str(n), case]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ))
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "ok=True" in out.stdout and "native=True" in out.stdout
