"""Restarted GMRES for the partitioned (one process per GPU) path when the Jacobian is not
symmetric: trapping / reaction couplings, advection, solution-dependent coefficients.

Same method as the single-GPU ``solve_gmres`` in ``csrc/linalg.cu`` (the stand-in for the
reference's ``ksp_type preonly`` + LU, problem.py:271-289): GMRES(30), left block-Jacobi
preconditioning, classical Gram-Schmidt applied twice (CGS2), so one Arnoldi step costs one halo
exchange and three small all-reduces, independent of the Krylov dimension.  The heavy operators
are the CUDA library's (``fb_spmv`` on the owned rows, ``fb_precond``), handed in as callbacks;
this driver keeps the Krylov basis as one tensor and does the orthogonalisation with tensor
products on whatever device the vectors live on, which lets the gloo CPU tests run the identical
recurrence with scipy operators (``tests/test_distributed_cpu.py``).  Round-1 status: functional,
not fused - the device-resident Givens / multi-dot kernels of the single-GPU solver do not take
part yet, and the Hessenberg column is read back once per step.
"""
import numpy as np


def gmres(apply_A, apply_M, allreduce, b, x, n_owned, rtol, atol=0.0, max_it=10000, restart=30):
    """Solve ``A x = b`` for the rows this rank owns.

    ``b``, ``x``: torch vectors in the rank's local numbering (owned dofs first, then ghosts);
    only ``[:n_owned]`` of ``b`` is read and of ``x`` is written (``x`` must start at zero).
    ``apply_A(v)`` refreshes the ghost entries of ``v`` (halo exchange) and returns the owned rows of
    ``A v`` (a vector at least ``n_owned`` long, valid until the next call); ``apply_M(r)`` applies
    the preconditioner to the owned entries likewise; ``allreduce(t)`` sums a small tensor over the
    ranks in place and returns it.  Stops when the preconditioned residual norm falls below
    ``max(rtol * ||M^-1 b||, atol)``.  Returns ``(iterations, converged)``.
    """
    import torch

    no = int(n_owned)
    n_loc = b.shape[0]
    V = torch.zeros(restart + 1, n_loc, dtype=b.dtype, device=b.device)
    r = torch.zeros(n_loc, dtype=b.dtype, device=b.device)
    w = torch.zeros(no, dtype=b.dtype, device=b.device)

    def norm(v):
        return float(np.sqrt(float(allreduce(torch.dot(v, v).reshape(1))[0])))

    total, first, threshold = 0, True, 0.0
    while True:
        # preconditioned residual of the current iterate
        if first:
            r[:no] = apply_M(b)[:no]
        else:
            r[:no] = b[:no] - apply_A(x)[:no]
            r[:no] = apply_M(r)[:no]
        beta = norm(r[:no])
        if first:
            threshold = max(rtol * beta, atol)
            first = False
        if not np.isfinite(beta):
            return total, False
        if beta <= threshold or beta == 0.0:
            return total, True
        if total >= max_it:
            return total, False
        V[0, :no] = r[:no] / beta
        H = np.zeros((restart + 1, restart))
        cs, sn = np.zeros(restart), np.zeros(restart)
        g = np.zeros(restart + 1)
        g[0] = beta
        k = 0
        for j in range(restart):
            w.copy_(apply_M(apply_A(V[j]))[:no])
            basis = V[: j + 1, :no]
            h1 = allreduce(basis @ w)
            w -= h1 @ basis
            h2 = allreduce(basis @ w)          # second pass: orthogonal to round-off
            w -= h2 @ basis
            hn = norm(w)
            col = (h1 + h2).cpu().numpy()
            H[: j + 1, j] = col
            H[j + 1, j] = hn
            for i in range(j):                 # earlier rotations, then the new one
                a, c = H[i, j], H[i + 1, j]
                H[i, j] = cs[i] * a + sn[i] * c
                H[i + 1, j] = -sn[i] * a + cs[i] * c
            rho = np.hypot(H[j, j], H[j + 1, j])
            if rho == 0.0 or not np.isfinite(rho):
                return total, False
            cs[j], sn[j] = H[j, j] / rho, H[j + 1, j] / rho
            H[j, j], H[j + 1, j] = rho, 0.0
            g[j + 1] = -sn[j] * g[j]
            g[j] = cs[j] * g[j]
            k = j + 1
            total += 1
            if abs(g[j + 1]) <= threshold or total >= max_it or hn == 0.0:
                break                          # (hn == 0: invariant subspace, the solution is in it)
            V[j + 1, :no] = w / hn
        y = np.linalg.solve(np.triu(H[:k, :k]), g[:k])
        x[:no] += torch.as_tensor(y, dtype=b.dtype, device=b.device) @ V[:k, :no]
        # the loop head recomputes the true preconditioned residual: convergence is declared on
        # it, not on the recurrence (which drifts over restarts)
