"""Problem builders shared by the CPU (oracle) and GPU (parity) tests.  Each one
is the festim_b200 transcription of a reference test (cited per function)."""

import numpy as np

import festim_b200 as F
from festim_b200 import fem, ufl


def error_L2(mesh, u_nodal, u_exact_np, degree=8):
    """L2 norm of (P1 field - exact), integrated with a high-degree rule; stands
    in for reference test/system_tests/tools.py:18-45 (degree-raised spaces)."""
    from festim_b200 import quadrature

    x = mesh.coords[mesh.cells]
    J = x[:, 1:, :] - x[:, :1, :]
    vol = np.abs(np.linalg.det(J)) / {1: 1.0, 2: 2.0, 3: 6.0}[mesh.tdim]
    bary, w = quadrature.rule(mesh.tdim, degree)
    xq = np.einsum("qa,ead->eqd", bary, x)
    uq = np.einsum("qa,ea->eq", bary, u_nodal[mesh.cells])
    x3 = np.zeros((3,) + xq.shape[:2])
    x3[: mesh.gdim] = np.moveaxis(xq, 2, 0)
    return float(np.sqrt(np.sum(vol[:, None] * w[None, :] * (uq - u_exact_np(x3)) ** 2)))


def mms_1_mobile_steady(dim, n):
    """reference test/system_tests/test_1_mobile_species.py:20-80 (1D, T=500+100x),
    :216-282 (3D); the 2D variant uses the 3D solution on a square."""
    D_0, E_D = 1, 0.1
    if dim == 1:
        mesh = F.Mesh1D(np.linspace(0, 1, n + 1))
        u_exact = lambda mod: (lambda x: 1 + mod.sin(2 * mod.pi * x[0]))
        vol = F.VolumeSubdomain1D(id=1, borders=[0, 1], material=F.Material(D_0=D_0, E_D=E_D))
        left, right = F.SurfaceSubdomain1D(id=1, x=0), F.SurfaceSubdomain1D(id=2, x=1)
    else:
        raw = F.create_unit_square(n, n) if dim == 2 else F.create_unit_cube(n, n, n)
        mesh = F.Mesh(raw)
        u_exact = lambda mod: (lambda x: 1 + mod.sin(2 * mod.pi * x[0]) + mod.cos(2 * mod.pi * x[1]))
        vol = F.VolumeSubdomain(id=1, material=F.Material(D_0=D_0, E_D=E_D))
        left = F.SurfaceSubdomain(id=1, locator=lambda x: np.isclose(x[0], 0))
        right = F.SurfaceSubdomain(id=2, locator=lambda x: np.isclose(x[0], 1))
    T_expr = lambda x: 500 + 100 * x[0]
    x = ufl.SpatialCoordinate(mesh.mesh)
    T = fem.Function(fem.functionspace(mesh.mesh, ("Lagrange", 1)))
    T.interpolate(T_expr)
    D = D_0 * ufl.exp(-E_D / (F.k_B * T))
    m = F.HydrogenTransportProblem()
    m.mesh = mesh
    m.subdomains = [vol, left, right]
    H = F.Species("H")
    m.species = [H]
    m.temperature = T_expr
    m.boundary_conditions = [
        F.DirichletBC(subdomain=left, value=u_exact(ufl), species=H),
        F.DirichletBC(subdomain=right, value=u_exact(ufl), species=H),
    ]
    f = -ufl.div(D * ufl.grad(u_exact(ufl)(x)))
    m.sources = [F.ParticleSource(value=f, volume=vol, species=H)]
    m.settings = F.Settings(atol=1e-10, rtol=1e-10, transient=False)
    return m, H, u_exact(np)


def mms_1_mobile_1_trap_steady(dim, n):
    """reference test/system_tests/test_1_mobile_1_trap.py:20-120 (steady MMS with
    one trap: mobile + trapped species, Arrhenius k and p, constant density)."""
    D_0, E_D = 2, 0.1
    k_0, E_k, p_0, E_p, n_trap = 2, 1.5, 0.2, 0.1, 3
    if dim == 1:
        mesh = F.Mesh1D(np.linspace(0, 1, n + 1))
        u_ex = lambda mod: (lambda x: 1.5 + mod.sin(3 * mod.pi * x[0]))
        v_ex = lambda mod: (lambda x: mod.sin(3 * mod.pi * x[0]))
        mat = F.Material(D_0=D_0, E_D=E_D)
        vol = F.VolumeSubdomain1D(id=1, borders=[0, 1], material=mat)
        left, right = F.SurfaceSubdomain1D(id=1, x=0), F.SurfaceSubdomain1D(id=2, x=1)
    else:
        raw = F.create_unit_square(n, n) if dim == 2 else F.create_unit_cube(n, n, n)
        mesh = F.Mesh(raw)
        u_ex = lambda mod: (lambda x: 1 + mod.sin(2 * mod.pi * x[0]) + mod.cos(2 * mod.pi * x[1]))
        v_ex = lambda mod: (lambda x: 1 + mod.cos(2 * mod.pi * x[0]) + mod.sin(2 * mod.pi * x[1]))
        mat = F.Material(D_0=D_0, E_D=E_D)
        vol = F.VolumeSubdomain(id=1, material=mat)
        left = F.SurfaceSubdomain(id=1, locator=lambda x: np.isclose(x[0], 0))
        right = F.SurfaceSubdomain(id=2, locator=lambda x: np.isclose(x[0], 1))
    T_expr = lambda x: 500 + 100 * x[0]
    x = ufl.SpatialCoordinate(mesh.mesh)
    T = T_expr(x)
    D = D_0 * ufl.exp(-E_D / (F.k_B * T))
    k = k_0 * ufl.exp(-E_k / (F.k_B * T))
    p = p_0 * ufl.exp(-E_p / (F.k_B * T))
    ue, ve = u_ex(ufl)(x), v_ex(ufl)(x)
    f = -ufl.div(D * ufl.grad(ue)) + k * ue * (n_trap - ve) - p * ve
    g = p * ve - k * ue * (n_trap - ve)
    m = F.HydrogenTransportProblem()
    m.mesh = mesh
    m.subdomains = [vol, left, right]
    mobile = F.Species("mobile")
    trapped = F.Species("trapped", mobile=False)
    traps = F.ImplicitSpecies(n=n_trap, others=[trapped])
    m.species = [mobile, trapped]
    m.temperature = T_expr
    m.reactions = [F.Reaction(reactant=[mobile, traps], product=trapped, k_0=k_0, E_k=E_k,
                              p_0=p_0, E_p=E_p, volume=vol)]
    m.boundary_conditions = [
        F.DirichletBC(subdomain=left, value=u_ex(ufl), species=mobile),
        F.DirichletBC(subdomain=right, value=u_ex(ufl), species=mobile),
        F.DirichletBC(subdomain=left, value=v_ex(ufl), species=trapped),
        F.DirichletBC(subdomain=right, value=v_ex(ufl), species=trapped),
    ]
    m.sources = [F.ParticleSource(value=f, volume=vol, species=mobile),
                 F.ParticleSource(value=g, volume=vol, species=trapped)]
    m.settings = F.Settings(atol=1e-12, rtol=1e-12, transient=False)
    return m, (mobile, trapped), (u_ex(np), v_ex(np))
