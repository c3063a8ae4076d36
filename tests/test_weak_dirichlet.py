"""Weakly enforced (Nitsche) Dirichlet conditions, SURVEY 8f2: the reference's tests
test/system_tests/test_misc.py:395-600 on the CPU oracle, and CUDA-vs-oracle parity."""
import numpy as np
import pytest

import festim_b200 as F
from cases import error_L2
from festim_b200 import ufl
from oracle.newton import OracleBackend


def _mms_1d(N, D_0, penalty):
    u_exact = lambda mod: (lambda x: 1 + mod.sin(2 * mod.pi * x[0]))
    mesh = F.Mesh1D(np.linspace(0, 1, N + 1))
    x = ufl.SpatialCoordinate(mesh.mesh)
    m = F.HydrogenTransportProblem()
    m.mesh = mesh
    vol = F.VolumeSubdomain1D(id=1, borders=[0, 1], material=F.Material(name="mat", D_0=D_0, E_D=0))
    left, right = F.SurfaceSubdomain1D(id=1, x=0), F.SurfaceSubdomain1D(id=2, x=1)
    m.subdomains = [vol, left, right]
    H = F.Species("H")
    m.species = [H]
    m.temperature = 500
    m.boundary_conditions = [F.FixedConcentrationBC(subdomain=s, value=u_exact(ufl), species=H,
                                                    enforce_weakly=True, penalty=penalty) for s in (left, right)]
    m.sources = [F.ParticleSource(value=-ufl.div(D_0 * ufl.grad(u_exact(ufl)(x))), volume=vol, species=H)]
    m.settings = F.Settings(atol=1e-12, rtol=1e-12, transient=False)
    return m, H, u_exact(np)


def _solve_1d(N, D_0, penalty, backend=None):
    m, H, exact = _mms_1d(N, D_0, penalty)
    if backend is not None:
        m.solver_backend = backend
    m.initialise()
    m.run()
    return error_L2(m.mesh.mesh, H.post_processing_solution.x.array, exact), m, H


@pytest.mark.parametrize("D_0", [1.0, 0.01])
def test_weak_dirichlet_convergence_rate(D_0):
    errors = [_solve_1d(N, D_0, 100, OracleBackend())[0] for N in (20, 40, 80, 160)]
    rates = np.log(np.array(errors[:-1]) / np.array(errors[1:])) / np.log(2)
    assert np.all(rates > 1.8), f"convergence rates {rates} are not close to 2"


def test_weak_dirichlet_insensitive_to_penalty():
    errors = [_solve_1d(40, 0.01, p, OracleBackend())[0] for p in (10, 100, 1000)]
    assert np.max(errors) / np.min(errors) < 2


def test_weak_dirichlet_penalty_is_dimensionless():
    errors = [_solve_1d(40, D_0, 100, OracleBackend())[0] for D_0 in (1e2, 1.0, 1e-4, 1e-10)]
    assert np.allclose(errors, errors[0], rtol=1e-8)


def _mms_2d(n, dim=2):
    u_exact = lambda mod: (lambda x: 1 + mod.sin(2 * mod.pi * x[0]) + mod.cos(2 * mod.pi * x[1]))
    raw = F.create_unit_square(n, n) if dim == 2 else F.create_unit_cube(n, n, n)
    mesh = F.Mesh(raw)
    T_expr = lambda x: 500 + 100 * x[0]
    from festim_b200 import fem

    T = fem.Function(fem.functionspace(mesh.mesh, ("Lagrange", 1)))
    T.interpolate(T_expr)
    D = 1 * ufl.exp(-0.1 / (F.k_B * T))
    m = F.HydrogenTransportProblem()
    m.mesh = mesh
    vol = F.VolumeSubdomain(id=1, material=F.Material(name="mat", D_0=1, E_D=0.1))
    left = F.SurfaceSubdomain(id=1, locator=lambda x: np.isclose(x[0], 0))
    right = F.SurfaceSubdomain(id=2, locator=lambda x: np.isclose(x[0], 1))
    m.subdomains = [vol, left, right]
    H = F.Species("H")
    m.species = [H]          # (the reference adds a second species without any condition: a
    m.temperature = T_expr   # singular block its direct solver happens to tolerate)
    m.boundary_conditions = [F.FixedConcentrationBC(subdomain=s, value=u_exact(ufl), species=H,
                                                    enforce_weakly=True, penalty=300) for s in (left, right)]
    x = ufl.SpatialCoordinate(mesh.mesh)
    m.sources = [F.ParticleSource(value=-ufl.div(D * ufl.grad(u_exact(ufl)(x))), volume=vol, species=H)]
    m.settings = F.Settings(atol=1e-10, rtol=1e-10, transient=False)
    return m, H, u_exact(np)


def test_MMS_weak_dirichlet_2d():
    m, H, exact = _mms_2d(50)
    m.solver_backend = OracleBackend()
    m.initialise()
    m.run()
    assert error_L2(m.mesh.mesh, H.post_processing_solution.x.array, exact) < 2e-3


@pytest.mark.gpu
@pytest.mark.parametrize("dim,n", [(1, 80), (2, 16), (3, 6)])
def test_weak_dirichlet_gpu_parity(dim, n):
    def build():
        return _mms_1d(n, 0.01, 100) if dim == 1 else _mms_2d(n, dim)

    (g, Hg, exact), (r, Hr, _) = build(), build()
    r.solver_backend = OracleBackend()
    for m in (g, r):
        m.initialise()
    rng = np.random.default_rng(5)
    u = rng.uniform(0.5, 1.5, g.u.x.array.size)
    for m in (g, r):
        m.u.x.array[:] = u
        m.u_n.x.array[:] = u
    Fg, Jg = g.solver.assemble_host(want_J=True)
    Fr, Jr = r.solver.residual_and_jacobian(r.u.x.array)
    assert np.abs(Fg - Fr).max() <= 1e-10 * np.abs(Fr).max()
    assert abs((Jg - Jr)).max() <= 1e-10 * abs(Jr).max()
    for m in (g, r):
        m.u.x.array[:] = 0.0
        m.run()
    a, b = Hg.post_processing_solution.x.array, Hr.post_processing_solution.x.array
    assert np.linalg.norm(a - b) / np.linalg.norm(b) < 1e-8
