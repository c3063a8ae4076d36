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


def mms_1_mobile_transient(n=9999, final_time=0.1, nsteps=50):
    """reference test/system_tests/test_1_mobile_species.py:83-145 (L2 < 5e-4)."""
    D_0, E_D = 1, 0.1
    mesh = F.Mesh1D(np.linspace(0, 1, n + 1))
    u_exact = lambda mod: (lambda x, t: 1 + mod.sin(2 * mod.pi * x[0]) + 2 * t**2)
    T_expr = lambda x: 600 + 50 * x[0]
    x = ufl.SpatialCoordinate(mesh.mesh)
    T = fem.Function(fem.functionspace(mesh.mesh, ("Lagrange", 1)))
    T.interpolate(T_expr)
    D = D_0 * ufl.exp(-E_D / (F.k_B * T))
    m = F.HydrogenTransportProblem()
    m.mesh = mesh
    vol = F.VolumeSubdomain1D(id=1, borders=[0, 1], material=F.Material(D_0=D_0, E_D=E_D))
    left, right = F.SurfaceSubdomain1D(id=1, x=0), F.SurfaceSubdomain1D(id=2, x=1)
    m.subdomains = [vol, left, right]
    H = F.Species("H")
    m.species = [H]
    m.temperature = T_expr
    m.boundary_conditions = [
        F.DirichletBC(subdomain=left, value=u_exact(ufl), species=H),
        F.DirichletBC(subdomain=right, value=u_exact(ufl), species=H),
    ]
    m.initial_conditions = [F.InitialConcentration(
        value=lambda x: 1 + ufl.sin(2 * ufl.pi * x[0]), species=H, volume=vol)]

    def f(x, t):
        return 4 * t - ufl.div(D * ufl.grad(u_exact(ufl)(x, t)))

    m.sources = [F.ParticleSource(value=f, volume=vol, species=H)]
    m.settings = F.Settings(atol=1e-10, rtol=1e-10, final_time=final_time)
    m.settings.stepsize = final_time / nsteps
    return m, H, (lambda x: u_exact(np)(x, final_time))


def permeation_1d(mesh_size=1001, final_time=50, dt=0.3):
    """reference test/test_permeation_problem.py:35-113: Sieverts BC upstream, c=0
    downstream, outgassing flux against the series solution (1 % mean error)."""
    L = 3e-04
    m = F.HydrogenTransportProblem()
    m.mesh = F.Mesh1D(np.linspace(0, L, num=mesh_size))
    mat = F.Material(D_0=1.9e-7, E_D=0.2, name="my_mat")
    vol = F.VolumeSubdomain1D(id=1, borders=[0, L], material=mat)
    left, right = F.SurfaceSubdomain1D(id=1, x=0), F.SurfaceSubdomain1D(id=2, x=L)
    m.subdomains = [vol, left, right]
    H = F.Species("H")
    m.species = [H]
    m.temperature = 500
    m.boundary_conditions = [
        F.DirichletBC(subdomain=right, value=0, species="H"),
        F.SievertsBC(subdomain=left, S_0=4.02e21, E_S=1.04, pressure=100, species="H"),
    ]
    flux = F.SurfaceFlux(field=H, surface=right)
    m.exports = [flux]
    m.settings = F.Settings(atol=1e2, rtol=1e-10, max_iterations=30, final_time=final_time)
    m.settings.stepsize = F.Stepsize(initial_value=dt)
    return m, H, flux, mat, L


def permeation_relative_error(D, permeability, computed_flux, L, times, P_up):
    """reference test/test_permeation_problem.py:11-32."""
    n_array = np.arange(1, 10000)[:, np.newaxis]
    summation = np.sum((-1) ** n_array * np.exp(-((np.pi * n_array) ** 2) * float(D) / L**2 * times),
                       axis=0)
    analytical = np.abs(P_up**0.5 * permeability / L * (2 * summation + 1))
    idx = np.where(analytical > 0.1 * np.max(analytical))
    return float(np.abs((computed_flux[idx] - analytical[idx]) / analytical[idx]).mean())


def tds_1d(n_cells=200, final_time=60.0):
    """BASELINE configs[0] (C0) in small: 1D tungsten TDS, mobile H + 1 trap,
    implantation-like source then a temperature ramp T(t), adaptive implicit Euler
    with milestones; parameters from reference
    test/system_tests/test_reaction_not_in_every_volume.py:10-29 and
    docs userguide (Settings(atol=1e10, rtol=1e-10))."""
    L = 2e-5
    m = F.HydrogenTransportProblem()
    m.mesh = F.Mesh1D(np.linspace(0, L, n_cells + 1))
    w = F.Material(D_0=4.1e-7, E_D=0.39, name="W")
    vol = F.VolumeSubdomain1D(id=1, borders=[0, L], material=w)
    left, right = F.SurfaceSubdomain1D(id=1, x=0), F.SurfaceSubdomain1D(id=2, x=L)
    m.subdomains = [vol, left, right]
    H = F.Species("H")
    m.species = [H]
    trap = F.Trap(name="trapped_H", mobile_species=H, k_0=8.9e-17 * 1e-1, E_k=0.39,
                  p_0=1e13, E_p=0.87, n=8.19e25 * 1e-3, volume=vol)
    m.traps = [trap]
    t_imp, t_rest, ramp = 20.0, 10.0, 8.0
    m.temperature = lambda t: 300.0 if t < t_imp + t_rest else 300.0 + ramp * (t - t_imp - t_rest)
    m.boundary_conditions = [
        F.FixedConcentrationBC(subdomain=left, value=lambda t: 1e20 if t < 20.0 else 0.0, species=H),
        F.FixedConcentrationBC(subdomain=right, value=0, species=H),
    ]
    flux_l = F.SurfaceFlux(field=H, surface=left)
    flux_r = F.SurfaceFlux(field=H, surface=right)
    inv = F.TotalVolume(field=H, volume=vol)
    m.exports = [flux_l, flux_r, inv]
    m.settings = F.Settings(atol=1e10, rtol=1e-10, final_time=final_time)
    m.settings.stepsize = F.Stepsize(
        initial_value=0.5, growth_factor=1.2, cutback_factor=0.9, target_nb_iterations=4,
        max_stepsize=lambda t: 2.0 if t < 30.0 else 1.0, milestones=[20.0, 30.0, final_time])
    return m, H, trap, (flux_l, flux_r, inv)
