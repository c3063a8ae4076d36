"""The library's own CUDA sources, compiled for the CPU by tests/emu (one OS thread per CUDA
thread, real barriers for __syncwarp / __syncthreads / warp collectives), assemble the same
residual and Jacobian as the oracle - the warp-per-vertex channel assembly (assemble_vtx.cuh),
the thread-per-vertex and facet kernels on the block layout, and the block SpMV - checked here
without a GPU.  The `-m gpu` suite repeats the comparison on the real device."""
import warnings

import numpy as np
import pytest

import cases
import festim_b200 as F
from emu_context import EmuSolver
from oracle.assembly import apply_dirichlet
from oracle.newton import OracleBackend

CASES = [
    ("multi_isotope_traps_3d", (2,)), ("mms_1_mobile_1_trap_steady", (3, 2)),
    ("mms_1_mobile_1_trap_steady", (2, 4)), ("mms_1_mobile_1_trap_steady", (1, 12)),
    ("monoblock_coupled_3d", (4,)), ("advection_trap_mms_2d", (4,)), ("tds_1d", (20, 5.0)),
    ("complex_reaction", ()), ("mms_1_mobile_steady", (3, 3)), ("permeation_advection_2d", (5,)),
    ("heat_mms", (True, 12)), ("recombination_flux", (2, 4)),
]


def _problems(name, args):
    ret = getattr(cases, name)(*args)
    out = []
    for obj in ret if isinstance(ret, (tuple, list)) else [ret]:
        if isinstance(obj, F.CoupledTransientHeatTransferHydrogenTransport):
            obj.heat_problem.solver_backend = OracleBackend()
            obj.hydrogen_problem.solver_backend = OracleBackend()
            obj.initialise()
            n = obj.heat_problem.u.x.array.size
            obj.heat_problem.u.x.array[:] = 400.0 + 50.0 * np.arange(n) / n
            out += [obj.heat_problem, obj.hydrogen_problem]
        elif isinstance(obj, F.ProblemBase):
            obj.solver_backend = OracleBackend()
            obj.initialise()
            out.append(obj)
    return out


@pytest.mark.parametrize("name,args", CASES, ids=[f"{n}{a}" for n, a in CASES])
def test_emulated_kernels_match_the_oracle(name, args):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        problems = _problems(name, args)
    for p in problems:
        es = EmuSolver(p)
        N = es.n_dofs
        rng = np.random.default_rng(11)
        scale = 1e18 if name == "multi_isotope_traps_3d" else 1.0
        u = rng.uniform(0.5, 1.5, N) * scale
        un = rng.uniform(0.5, 1.5, N) * scale
        Fo, Jo = p.solver.assembler.assemble(u, un, want_J=True)
        dofs, vals = p.solver.bc_arrays()
        bo, Jbo = apply_dirichlet(Fo, Jo, u, dofs, vals)
        Fe, Je = es.assemble(u, un, 3)
        assert not np.isnan(es.Jvals).any() and not np.isnan(Fe).any()
        assert np.abs(Fe - bo).max() <= 1e-11 * np.abs(bo).max(), (name, "F")
        assert np.abs((Je - Jbo).tocoo().data).max(initial=0.0) <= 1e-11 * np.abs(Jbo.data).max(), (name, "J")
        # separate passes (what Newton iterations after the first do)
        F1, _ = es.assemble(u, un, 1)
        assert np.abs(F1 - Fe).max() <= 1e-13 * np.abs(bo).max()
        _, J2 = es.assemble(u, un, 2)
        assert np.abs((J2 - Je).tocoo().data).max(initial=0.0) <= 1e-13 * np.abs(Jbo.data).max()
        # SpMV on the stored layout
        x = rng.standard_normal(N)
        y = es.spmv(es.Jvals, x)
        ref = Jbo @ x
        assert np.abs(y - ref).max() <= 1e-11 * np.abs(ref).max(), (name, "spmv")
        # SpMV on the channel operator itself (no block-CSR values): same matrix, Dirichlet
        # rows / columns applied on the fly (solution-dependent facet terms live in the block-CSR values only)
        if es.spec.vx is not None and not any(t.u_dependent for t in es.spec.fterms):
            yc = es.spmv(None, x)
            assert np.abs(yc - ref).max() <= 1e-11 * np.abs(ref).max(), (name, "channel spmv")


def test_emulated_newton_gmres_matches_the_oracle():
    """fb_newton_solve with GMRES (block SpMV with the fused block-Jacobi application, fused
    Gram-Schmidt sweeps, device-resident Givens) on the emulated device: same Newton iteration
    count and solution as the oracle's Newton with an exact linear solve."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = cases.mms_1_mobile_1_trap_steady(2, 3)[0]
        m.solver_backend = OracleBackend()
        m.initialise()
    es = EmuSolver(m)
    u, its, reason, lin = es.newton(m.u.x.array.copy(), m.u_n.x.array.copy(), float(m.settings.atol),
                                    float(m.settings.rtol))
    nit = m.solver.solve()
    assert reason > 0 and its == nit and lin > 0
    assert np.linalg.norm(u - m.u.x.array) <= 1e-10 * np.linalg.norm(m.u.x.array)


def test_emulated_gmres_on_the_channel_operator():
    """C2-like problem (12 species, 66 stored pairs, 11 channels): GMRES with the fused block-Jacobi
    SpMV on the channel operator (J = null) solves the same system as on the block-CSR values."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = cases.multi_isotope_traps_3d(2)[0]
        m.solver_backend = OracleBackend()
        m.initialise()
    es = EmuSolver(m)
    N = es.n_dofs
    rng = np.random.default_rng(3)
    u = rng.uniform(0.5, 1.5, N) * 1e18
    un = rng.uniform(0.5, 1.5, N) * 1e18
    Fe, Je = es.assemble(u, un, 3)
    y_ch, its_ch, rel_ch, rc = es.linear_solve(None, Fe)
    assert rc == 0 and rel_ch < 1e-11
    y_bs, its_bs, rel_bs, rc = es.linear_solve(es.Jvals, Fe)
    assert rc == 0 and its_ch == its_bs
    exact = __import__("scipy.sparse.linalg", fromlist=["spsolve"]).spsolve(Je.tocsc(), Fe)
    assert np.linalg.norm(y_ch - exact) <= 1e-9 * np.linalg.norm(exact)


def test_emulated_chebyshev_preconditioned_gmres():
    """pc = Chebyshev on the channel operator: the first solve runs block-Jacobi GMRES and leaves the
    interval of D^-1 J (Ritz values of its Arnoldi matrix); the next one is right-preconditioned with
    the Chebyshev polynomial (fused recurrence steps in the TMA-pipelined SpMV kernel), converges in
    a fraction of the Krylov iterations and returns the same solution."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = cases.multi_isotope_traps_3d(2)[0]
        m.solver_backend = OracleBackend()
        m.initialise()
    es = EmuSolver(m)
    N = es.n_dofs
    rng = np.random.default_rng(4)
    u = rng.uniform(0.5, 1.5, N) * 1e18
    un = rng.uniform(0.5, 1.5, N) * 1e18
    Fe, Je = es.assemble(u, un, 3)
    y0, its0, rel0, rc = es.linear_solve(None, Fe, pc=1)
    assert rc == 0
    valid, lmin, lmax, steps = es.chebyshev_info()
    assert valid == 1.0 and 0.0 < lmin < lmax and steps >= 2
    y1, its1, rel1, rc = es.linear_solve(None, Fe, pc=1)
    assert rc == 0 and rel1 < 1e-11
    assert its1 * 2 < its0, (its0, its1)
    exact = __import__("scipy.sparse.linalg", fromlist=["spsolve"]).spsolve(Je.tocsc(), Fe)
    assert np.linalg.norm(y1 - exact) <= 1e-9 * np.linalg.norm(exact)
    assert np.linalg.norm(y0 - exact) <= 1e-9 * np.linalg.norm(exact)


def test_emulated_tma_assembly_matches_the_row_kernel(monkeypatch):
    """The TMA-pipelined Newton-iteration assembly (assemble_edge_tma.cuh: bulk-copy pipeline, both
    contractions as DMMA tiles) writes the same residual and channel operator as the row kernel
    (k_assemble_edge, selected with FB_NO_TMA_ASSEMBLY)."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = cases.multi_isotope_traps_3d(2)[0]
        m.solver_backend = OracleBackend()
        m.initialise()
    es = EmuSolver(m)
    N = es.n_dofs
    rng = np.random.default_rng(9)
    u = rng.uniform(0.5, 1.5, N) * 1e18
    un = rng.uniform(0.5, 1.5, N) * 1e18
    F0, J0 = es.assemble(u, un, 3)
    monkeypatch.setenv("FB_NO_TMA_ASSEMBLY", "1")
    F1, J1 = es.assemble(u, un, 3)
    assert np.abs(F1 - F0).max() <= 1e-13 * np.abs(F0).max()
    assert np.abs((J1 - J0).tocoo().data).max(initial=0.0) <= 1e-13 * np.abs(J0.data).max()
