"""GPU parity: CUDA path (through the C ABI) vs the CPU oracle on identical
meshes and parameters.  Tolerances are BASELINE.json's: fields 1e-8 relative L2,
derived quantities 1e-6 relative; assembled F / J entries 1e-10 relative to the
largest entry."""
import numpy as np
import pytest

import festim_b200 as F
from cases import mms_1_mobile_1_trap_steady, mms_1_mobile_steady
from oracle.newton import OracleBackend

pytestmark = pytest.mark.gpu


def _pair(builder, *args):
    gpu = builder(*args)
    ref = builder(*args)
    ref[0].solver_backend = OracleBackend()
    gpu[0].initialise()
    ref[0].initialise()
    return gpu, ref


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _check_assembly(gm, rm, seed=0):
    rng = np.random.default_rng(seed)
    u = rng.uniform(0.5, 1.5, gm.u.x.array.size)
    for m in (gm, rm):
        m.u.x.array[:] = u
        m.u_n.x.array[:] = 0.9 * u
    Fg, Jg = gm.solver.assemble_host(want_J=True)
    Fr, Jr = rm.solver.residual_and_jacobian(rm.u.x.array)
    assert np.abs(Fg - Fr).max() <= 1e-10 * np.abs(Fr).max()
    D = (Jg - Jr).tocsr()
    assert abs(D).max() <= 1e-10 * abs(Jr).max()
    # F-only assembly gives the same residual (lifting without a stored Jacobian)
    Fg2, _ = gm.solver.assemble_host(want_J=False)
    assert np.abs(Fg2 - Fr).max() <= 1e-10 * np.abs(Fr).max()


@pytest.mark.parametrize("dim,n", [(1, 50), (2, 12), (3, 6)])
def test_assembly_diffusion_mms(dim, n):
    (gm, _, _), (rm, _, _) = _pair(mms_1_mobile_steady, dim, n)
    _check_assembly(gm, rm)


@pytest.mark.parametrize("dim,n", [(1, 40), (2, 10), (3, 5)])
def test_assembly_trap_mms(dim, n):
    (gm, _, _), (rm, _, _) = _pair(mms_1_mobile_1_trap_steady, dim, n)
    _check_assembly(gm, rm)


@pytest.mark.parametrize("dim,n", [(1, 200), (2, 24), (3, 10)])
def test_solve_diffusion_mms(dim, n):
    (gm, Hg, _), (rm, Hr, _) = _pair(mms_1_mobile_steady, dim, n)
    gm.run()
    rm.run()
    assert gm.solver.solver.getConvergedReason() > 0
    assert gm.solver.solver.getIterationNumber() == rm.solver.solver.getIterationNumber()
    assert _rel(Hg.post_processing_solution.x.array, Hr.post_processing_solution.x.array) < 1e-8


@pytest.mark.parametrize("dim,n", [(1, 200), (2, 16), (3, 6)])
def test_solve_trap_mms(dim, n):
    (gm, sg, _), (rm, sr, _) = _pair(mms_1_mobile_1_trap_steady, dim, n)
    gm.run()
    rm.run()
    assert gm.solver.solver.getConvergedReason() > 0
    assert gm.solver.solver.getIterationNumber() == rm.solver.solver.getIterationNumber()
    for a, b in zip(sg, sr):
        assert _rel(a.post_processing_solution.x.array, b.post_processing_solution.x.array) < 1e-8


def test_spmv_and_dot_against_scipy():
    (gm, _, _), (rm, _, _) = _pair(mms_1_mobile_1_trap_steady, 3, 5)
    rng = np.random.default_rng(1)
    u = rng.uniform(0.5, 1.5, gm.u.x.array.size)
    gm.u.x.array[:] = u
    _, Jg = gm.solver.assemble_host(want_J=True)
    s = gm.solver
    ctx = s.ctx
    import ctypes as C

    x = rng.standard_normal(ctx.n_dofs)
    xd = ctx.to_device("x", x)
    ud = ctx.to_device("u", gm.u.x.array)
    Fd, Jd, yd = ctx.zeros(ctx.n_dofs), ctx.zeros(ctx.nnz), ctx.zeros(ctx.n_dofs)
    ctx._check(ctx.lib.fb_assemble(ctx.h, ctx.ptr(ud), ctx.ptr(ud), ctx.ptr(Fd), ctx.ptr(Jd), 3))
    ctx._check(ctx.lib.fb_spmv(ctx.h, ctx.ptr(Jd), ctx.ptr(xd), ctx.ptr(yd)))
    y = yd.cpu().numpy()
    yr = Jg @ x
    assert np.abs(y - yr).max() <= 1e-13 * np.abs(yr).max()
    out = C.c_double()
    ctx._check(ctx.lib.fb_dot(ctx.h, ctx.ptr(xd), ctx.ptr(yd), C.byref(out)))
    assert abs(out.value - x @ yr) <= 1e-12 * abs(x @ yr)


def test_transient_mms_1d():
    from cases import mms_1_mobile_transient

    (gm, Hg, _), (rm, Hr, _) = _pair(mms_1_mobile_transient, 500, 0.1, 20)
    gm.run()
    rm.run()
    assert gm.timesteps == rm.timesteps
    assert _rel(Hg.post_processing_solution.x.array, Hr.post_processing_solution.x.array) < 1e-8


def test_permeation_flux():
    """SurfaceFlux export through the device reduction vs the oracle (1e-6)."""
    from cases import permeation_1d

    g, r = _pair(permeation_1d, 201, 10.0, 0.5)
    g[0].run()
    r[0].run()
    fg, fr = np.array(g[2].data), np.array(r[2].data)
    assert g[2].t == r[2].t
    assert np.abs(fg - fr).max() <= 1e-6 * np.abs(fr).max()
    assert _rel(g[1].post_processing_solution.x.array, r[1].post_processing_solution.x.array) < 1e-8


def test_tds_adaptive_stepping_parity():
    """C0-style TDS (mobile + trap, T(t) ramp, adaptive dt with milestones): the
    Newton iteration counts drive dt, so the recorded time steps must be identical
    (SURVEY F5); fields 1e-8, fluxes and inventory 1e-6."""
    from cases import tds_1d

    g, r = _pair(tds_1d, 100, 45.0)
    g[0].run()
    r[0].run()
    assert g[0].timesteps == r[0].timesteps
    for a, b in zip(g[0].species, r[0].species):
        assert _rel(a.post_processing_solution.x.array, b.post_processing_solution.x.array) < 1e-8
    for eg, er in zip(g[3], r[3]):
        dg, dr = np.array(eg.data), np.array(er.data)
        assert np.abs(dg - dr).max() <= 1e-6 * max(np.abs(dr).max(), 1e-300)


@pytest.mark.parametrize("T_dependent", [False, True])
def test_heat_transfer_parity(T_dependent):
    """a16: heat problem, constant and solution-dependent conductivity."""
    from cases import heat_mms

    (gm, _), (rm, _) = _pair(heat_mms, T_dependent, 400)
    _check_assembly(gm, rm, seed=2)
    gm.u.x.array[:] = 0.0
    rm.u.x.array[:] = 0.0
    gm.run()
    rm.run()
    assert gm.solver.solver.getIterationNumber() == rm.solver.solver.getIterationNumber()
    assert _rel(gm.u.x.array, rm.u.x.array) < 1e-8


@pytest.mark.parametrize("heat", [False, True])
def test_flux_bc_parity(heat):
    """a7: facet (flux) terms."""
    from cases import flux_bc_mms

    (gm, _, _), (rm, _, _) = _pair(flux_bc_mms, heat, 300)
    _check_assembly(gm, rm, seed=3)
    gm.u.x.array[:] = 0.0
    rm.u.x.array[:] = 0.0
    gm.run()
    rm.run()
    assert _rel(gm.u.x.array, rm.u.x.array) < 1e-8


def test_coupled_heat_hydrogen_parity():
    """a17: staggered coupling, T aliased between the two problems."""
    from cases import coupled_heat_hydrogen_mms

    cg, sg, _ = coupled_heat_hydrogen_mms(300, 10)
    cr, sr, _ = coupled_heat_hydrogen_mms(300, 10)
    cr.heat_problem.solver_backend = OracleBackend()
    cr.hydrogen_problem.solver_backend = OracleBackend()
    for c in (cg, cr):
        c.initialise()
        c.run()
    assert cg.hydrogen_problem.timesteps == cr.hydrogen_problem.timesteps
    assert _rel(cg.heat_problem.u.x.array, cr.heat_problem.u.x.array) < 1e-8
    for a, b in zip(sg, sr):
        assert _rel(a.post_processing_solution.x.array, b.post_processing_solution.x.array) < 1e-8


def test_advection_trap_parity_2d():
    """a8: advection (non-symmetric J -> GMRES) with a trap, T given as a Function."""
    from cases import advection_trap_mms_2d

    (gm, sg, _), (rm, sr, _) = _pair(advection_trap_mms_2d, 24)
    _check_assembly(gm, rm, seed=4)
    gm.u.x.array[:] = 0.0
    rm.u.x.array[:] = 0.0
    gm.run()
    rm.run()
    assert gm.solver.solver.getIterationNumber() == rm.solver.solver.getIterationNumber()
    for a, b in zip(sg, sr):
        assert _rel(a.post_processing_solution.x.array, b.post_processing_solution.x.array) < 1e-8


def test_multi_isotope_three_traps_3d():
    """C2 in small: 12 species, 9 reactions, species-sparse Jacobian (66 of 144 pairs)."""
    from cases import multi_isotope_traps_3d

    (gm, _, _), (rm, _, _) = _pair(multi_isotope_traps_3d, 3)
    assert gm.solver.spec.npairs == 66
    _check_assembly(gm, rm, seed=5)
    for m in (gm, rm):
        m.u.x.array[:] = 0.0
        m.u_n.x.array[:] = 0.0
    gm.run()
    rm.run()
    assert gm.timesteps == rm.timesteps
    for a, b in zip(gm.species, rm.species):
        ref = b.post_processing_solution.x.array
        assert np.linalg.norm(a.post_processing_solution.x.array - ref) <= 1e-8 * max(np.linalg.norm(ref), 1e-300)


def test_complex_reaction_parity():
    """A + B <-> C + D (two-factor products, no Dirichlet BC): TotalVolume histories."""
    from cases import complex_reaction

    gm = complex_reaction(0.5, 350e-4, 120e-4, 3, 20)
    rm = complex_reaction(0.5, 350e-4, 120e-4, 3, 20)
    rm.solver_backend = OracleBackend()
    for m in (gm, rm):
        m.initialise()
        m.run()
    assert gm.timesteps == rm.timesteps
    for eg, er in zip(gm.exports, rm.exports):
        dg, dr = np.array(eg.data), np.array(er.data)
        assert np.abs(dg - dr).max() <= 1e-6 * max(np.abs(dr).max(), 1e-300)
    assert _rel(gm.u.x.array, rm.u.x.array) < 1e-8


@pytest.mark.parametrize("name", ["step_space", "step_time_space", "step_time"])
def test_non_homogeneous_density_parity(name):
    """Trap density as an expression of x, t (point-wise path) or of t (structured path)."""
    from cases import DENSITIES, non_homogeneous_density

    func, _ = DENSITIES[name]
    gm, tg = non_homogeneous_density(func, 60)
    rm, tr = non_homogeneous_density(func, 60)
    rm.solver_backend = OracleBackend()
    for m in (gm, rm):
        m.initialise()
        m.run()
    assert gm.timesteps == rm.timesteps
    assert np.abs(np.array(tg.data) - np.array(tr.data)).max() <= 1e-6 * np.abs(np.array(tr.data)).max()
    assert _rel(gm.u.x.array, rm.u.x.array) < 1e-8


@pytest.mark.parametrize("n", [4, 16])
def test_cg_two_level_preconditioner(n, monkeypatch):
    """Persistent CG: block-Jacobi + per-SM coarse space against block-Jacobi alone and
    against scipy's direct solve of the same assembled system.  n=4 has fewer vertices than
    SMs (empty blocks): the coarse space must switch itself off."""
    import ctypes as C

    import scipy.sparse.linalg as spl

    (gm, _, _), _ = _pair(mms_1_mobile_steady, 3, n)
    s = gm.solver
    ctx = s.ctx
    s._push_state()
    Fg, Jg = s.assemble_host(want_J=True)
    ref = spl.spsolve(Jg.tocsc(), Fg)
    u = ctx.nodal_to_device("u", gm.u.x.array, 1)
    Fd, Jd, yd = ctx.zeros(ctx.n_dofs), ctx.zeros(ctx.nnz), ctx.zeros(ctx.n_dofs)
    ctx._check(ctx.lib.fb_assemble(ctx.h, ctx.ptr(u), ctx.ptr(u), ctx.ptr(Fd), ctx.ptr(Jd), 3))
    sols, iters = [], []
    for no_coarse in (False, True):
        if no_coarse:
            monkeypatch.setenv("FB_NO_COARSE", "1")
        its, rr = C.c_int(0), C.c_double(0)
        ctx._check(ctx.lib.fb_linear_solve(ctx.h, ctx.ptr(Jd), ctx.ptr(Fd), ctx.ptr(yd), 1, 0, 1e-12, 0.0, 4000,
                                           C.byref(its), C.byref(rr)))
        sols.append(ctx.nodal_from_device(yd))
        iters.append(its.value)
        assert rr.value < 1e-11
    assert _rel(sols[0], sols[1]) < 1e-9
    assert _rel(sols[0], ref) < 1e-9
    if n == 16:
        assert iters[0] < iters[1]


@pytest.mark.parametrize("dim,n", [(2, 12), (3, 5)])
def test_recombination_flux_parity(dim, n):
    """Facet terms depending on x, T and two species (surface reactions) on triangles and
    tetrahedra: assembled F / J and the steady solution against the oracle."""
    from cases import recombination_flux

    (gm, sg, _), (rm, sr, _) = _pair(recombination_flux, dim, n)
    _check_assembly(gm, rm, seed=3)
    for m in (gm, rm):
        m.u.x.array[:] = 0.0
        m.u_n.x.array[:] = 0.0
    gm.run()
    rm.run()
    assert gm.solver.solver.getConvergedReason() > 0
    assert gm.solver.solver.getIterationNumber() == rm.solver.solver.getIterationNumber()
    for a, b in zip(sg, sr):
        assert _rel(a.post_processing_solution.x.array, b.post_processing_solution.x.array) < 1e-8


def test_device_resident_stepping_matches_the_python_loop(monkeypatch):
    """fb_run_steps (the implicit-Euler loop with the adaptive step-size rule inside the library, fields
    on the device throughout) against the same run driven step by step from Python: identical time
    steps (reference stepsize.py:144-167 incl. milestones) and the same field."""
    from cases import multi_isotope_traps_3d

    def run(python_loop):
        m, _, _ = multi_isotope_traps_3d(3, dt=1e-2, final_time=0.2)
        m.settings.stepsize = F.Stepsize(initial_value=1e-2, growth_factor=1.5, cutback_factor=0.8,
                                         target_nb_iterations=3, milestones=[0.05, 0.11])
        if python_loop:
            monkeypatch.setenv("FB_NO_RUN_STEPS", "1")
        else:
            monkeypatch.delenv("FB_NO_RUN_STEPS", raising=False)
        m.initialise()
        m.run()
        return m

    lib, py = run(False), run(True)
    assert getattr(lib.solver, "library_steps", 0) == len(lib.timesteps) > 3
    assert getattr(py.solver, "library_steps", 0) == 0
    assert lib.timesteps == py.timesteps
    assert float(lib.t) == float(py.t) and float(lib.dt.value) == float(py.dt.value)
    assert _rel(lib.u.x.array, py.u.x.array) < 1e-12
    assert _rel(lib.u_n.x.array, py.u_n.x.array) < 1e-12


def test_chebyshev_preconditioned_gmres_parity():
    """C2 in small with pc_type=chebyshev (what large channel-operator systems get by default):
    the polynomially preconditioned GMRES reproduces the oracle's fields and Newton iteration counts."""
    from cases import multi_isotope_traps_3d

    gm, _, _ = multi_isotope_traps_3d(5, final_time=3e-2)
    gm.petsc_options = {"pc_type": "chebyshev"}
    rm, _, _ = multi_isotope_traps_3d(5, final_time=3e-2)
    rm.solver_backend = OracleBackend()
    nit_g, nit_r = [], []
    for m, its in ((gm, nit_g), (rm, nit_r)):
        m.initialise()
        while float(m.t) < m.settings.final_time:
            m.iterate()
            its.append(m.solver.solver.getIterationNumber())
    import ctypes as C

    info = (C.c_double * 4)()
    gm.solver.ctx.lib.fb_chebyshev_info(gm.solver.ctx.h, info)
    assert info[0] == 1.0 and 0 < info[1] < info[2]
    assert nit_g == nit_r
    assert _rel(gm.u.x.array, rm.u.x.array) < 1e-8


def test_host_arrays_follow_the_device_fields():
    """The fields stay on the device between solves: the host arrays are refreshed when read, and
    an in-place edit of u.x.array by the user is picked up by the next solve."""
    from cases import multi_isotope_traps_3d

    a, _, _ = multi_isotope_traps_3d(3, final_time=1e9)
    b, _, _ = multi_isotope_traps_3d(3, final_time=1e9)
    for m in (a, b):
        m.initialise()
        m.iterate()
    assert _rel(a.u.x.array, b.u.x.array) < 1e-14
    # a user edit on the host between two steps
    for m in (a, b):
        m.u.x.array[:] *= 1.5
    a.iterate()
    b.u_n.x.array[:] = b.u_n.x.array      # touching u_n as well must not change anything
    b.iterate()
    assert _rel(a.u.x.array, b.u.x.array) < 1e-13
    assert np.array_equal(a.u_n.x.array, a.u.x.array)
