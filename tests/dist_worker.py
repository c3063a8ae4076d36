"""torchrun worker: partitioned solve of the 3D MMS diffusion problem on WORLD_SIZE
GPUs, compared with the single-process oracle on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from cases import mms_1_mobile_steady  # noqa: E402
from festim_b200.distributed import DistributedBackend  # noqa: E402
from oracle.newton import OracleBackend  # noqa: E402


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    case = sys.argv[2] if len(sys.argv) > 2 else "diffusion"

    def build():
        if case == "trap":          # non-symmetric Jacobian: partitioned GMRES
            from cases import mms_1_mobile_1_trap_steady

            return mms_1_mobile_1_trap_steady(3, n)[:2]
        if case == "c2":            # BASELINE configs[2] in small: channel operator, Chebyshev-preconditioned GMRES
            from cases import multi_isotope_traps_3d

            m, iso, _ = multi_isotope_traps_3d(n, final_time=2e-2)
            m.petsc_options = {"pc_type": "chebyshev"}
            return m, iso[0]
        return mms_1_mobile_steady(3, n)[:2]

    m, H = build()
    m.solver_backend = DistributedBackend(device=local)
    m.initialise()
    m.run()
    ok = True
    if rank == 0:
        r, Hr = build()
        r.solver_backend = OracleBackend()
        r.initialise()
        r.run()
        a, b = m.u.x.array, r.u.x.array
        rel = np.linalg.norm(a - b) / np.linalg.norm(b)
        ok = rel < 1e-8 and m.solver.solver.getIterationNumber() == r.solver.solver.getIterationNumber()
        print(f"DIST world={dist.get_world_size()} rel={rel:.3e} newton={m.solver.solver.its} "
              f"linear={m.solver.solver.linear_its} gmres={m.solver.gmres_iterations} peer={m.solver.peer} native={m.solver.native} ok={ok}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
