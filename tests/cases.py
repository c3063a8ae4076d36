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


def mms_1_mobile_1_trap_steady(dim, n, reference_solution=False):
    """reference test/system_tests/test_1_mobile_1_trap.py:20-120 (steady MMS with
    one trap: mobile + trapped species, Arrhenius k and p, constant density).
    ``reference_solution=True`` takes the manufactured solutions of the reference's 2D / 3D
    versions (:193-421) instead of the smoother ones the parity tests use."""
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
        if reference_solution:
            u_ex = lambda mod: (lambda x: 1.5 + mod.sin(3 * mod.pi * x[0]) + mod.cos(3 * mod.pi * x[1]))
            v_ex = lambda mod: (lambda x: mod.sin(3 * mod.pi * x[0]) + mod.cos(3 * mod.pi * x[1]))
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


def permeation_1d(mesh_size=1001, final_time=50, dt=0.3, n_volumes=1):
    """reference test/test_permeation_problem.py:35-113: Sieverts BC upstream, c=0
    downstream, outgassing flux against the series solution (1 % mean error);
    ``n_volumes=4``, ``dt=0.4``: the four-subdomain version (:116-204)."""
    L = 3e-04
    m = F.HydrogenTransportProblem()
    m.mesh = F.Mesh1D(np.linspace(0, L, num=mesh_size))
    mat = F.Material(D_0=1.9e-7, E_D=0.2, name="my_mat")
    vols = [F.VolumeSubdomain1D(id=i + 1, borders=[i * L / n_volumes, (i + 1) * L / n_volumes], material=mat)
            for i in range(n_volumes)]
    vol = vols[0]
    left, right = F.SurfaceSubdomain1D(id=1, x=0), F.SurfaceSubdomain1D(id=2, x=L)
    m.subdomains = [*vols, left, right]
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


def heat_mms(T_dependent=False, n=1999):
    """reference test/test_heat_transfer_problem.py:67-110 (constant conductivity)
    and :113-165 (conductivity 3T + 2); L2 < 1e-7."""
    p = F.HeatTransferProblem()
    p.mesh = F.Mesh1D(vertices=np.linspace(0, 1, n + 1))
    left, right = F.SurfaceSubdomain1D(id=1, x=0), F.SurfaceSubdomain1D(id=2, x=1)
    mat = F.Material(D_0=1, E_D=None)
    if T_dependent:
        mat.thermal_conductivity = lambda T: 3 * T + 2
        exact = lambda x: 2 * x[0] ** 2 + 1
        source = lambda x: -(72 * x[0] ** 2 + 20)
    else:
        mat.thermal_conductivity = 4.0
        exact = lambda x: 2 * x[0] ** 2
        source = -4 * 4.0
    vol = F.VolumeSubdomain1D(id=1, borders=[0, 1], material=mat)
    p.subdomains = [left, right, vol]
    p.boundary_conditions = [F.FixedTemperatureBC(subdomain=left, value=exact),
                             F.FixedTemperatureBC(subdomain=right, value=exact)]
    p.sources = [F.HeatSource(value=source, volume=vol)]
    p.settings = F.Settings(atol=1e-10, rtol=1e-10, transient=False)
    return p, exact


def flux_bc_mms(heat=False, n=9999):
    """reference test/system_tests/test_fluxes.py:13-57 (particle flux) and :60-95
    (heat flux); L2 < 1e-7."""
    mesh = F.Mesh1D(np.linspace(0, 1, n + 1))
    x = ufl.SpatialCoordinate(mesh.mesh)
    u_exact = lambda x: 1 + 2 * x[0] ** 2
    left, right = F.SurfaceSubdomain1D(id=1, x=0), F.SurfaceSubdomain1D(id=2, x=1)
    if heat:
        lam = 2.3
        m = F.HeatTransferProblem()
        vol = F.VolumeSubdomain1D(id=1, borders=[0, 1],
                                  material=F.Material(D_0=1, E_D=1, thermal_conductivity=lam))
        m.boundary_conditions = [F.FixedTemperatureBC(subdomain=left, value=u_exact),
                                 F.HeatFluxBC(subdomain=right, value=4 * lam)]
        m.sources = [F.HeatSource(value=-ufl.div(lam * ufl.grad(u_exact(x))), volume=vol)]
        field = None
    else:
        D = 1 * ufl.exp(-0.1 / (F.k_B * 500))
        m = F.HydrogenTransportProblem()
        vol = F.VolumeSubdomain1D(id=1, borders=[0, 1], material=F.Material(D_0=1, E_D=0.1))
        H = F.Species("H")
        m.species = [H]
        m.temperature = 500
        m.boundary_conditions = [F.DirichletBC(subdomain=left, value=u_exact, species=H),
                                 F.ParticleFluxBC(subdomain=right, value=4 * D, species=H)]
        m.sources = [F.ParticleSource(value=-ufl.div(D * ufl.grad(u_exact(x))), volume=vol, species=H)]
        field = H
    m.mesh = mesh
    m.subdomains = [vol, left, right]
    m.settings = F.Settings(atol=1e-10, rtol=1e-10, transient=False)
    return m, field, u_exact


def coupled_heat_hydrogen_mms(n=1999, nsteps=20, traps_others_includes_mobile=True):
    """reference test/system_tests/test_coupled_heat_hydrogen.py:9-207 (all three
    L2 errors < 2e-7)."""
    density, heat_capacity, lam = 1.2, 2.6, 4.2
    D_0, E_D, k_0, E_k, p_0, E_p, n_trap = 1.2, 0.1, 2.2, 0.2, 0.5, 0.1, 5
    k_B, final_time = F.k_B, 1
    mesh = F.Mesh1D(vertices=np.linspace(0, 1, n + 1))
    mat = F.Material(D_0=D_0, E_D=E_D, thermal_conductivity=lam, density=density,
                     heat_capacity=heat_capacity)
    vol = F.VolumeSubdomain1D(id=1, borders=[0, 1], material=mat)
    left, right = F.SurfaceSubdomain1D(id=2, x=0), F.SurfaceSubdomain1D(id=3, x=1)
    mobile = F.Species("mobile", mobile=True)
    trapped = F.Species(name="trapped", mobile=False, subdomains=[vol])
    others = [mobile, trapped] if traps_others_includes_mobile else [trapped]
    traps = F.ImplicitSpecies(n=n_trap, others=others)
    exact_T = lambda x, t: 3 * x[0] ** 2 + 10 * t
    heat = F.HeatTransferProblem(
        mesh=mesh, subdomains=[vol, left, right],
        boundary_conditions=[F.FixedTemperatureBC(subdomain=left, value=exact_T),
                             F.FixedTemperatureBC(subdomain=right, value=exact_T)],
        sources=[F.HeatSource(value=density * heat_capacity * 10 - lam * 6, volume=vol)],
        initial_condition=F.InitialTemperature(value=lambda x: exact_T(x, 0), volume=vol),
        settings=F.Settings(atol=1e-10, rtol=1e-10, transient=True, stepsize=final_time / nsteps,
                            final_time=final_time))
    ex_m = lambda x, t: 2 * x[0] ** 2 + 15 * t
    ex_t = lambda x, t: 4 * x[0] ** 2 + 12 * t
    D = lambda x, t: D_0 * ufl.exp(-E_D / (k_B * exact_T(x, t)))
    k = lambda x, t: k_0 * ufl.exp(-E_k / (k_B * exact_T(x, t)))
    p = lambda x, t: p_0 * ufl.exp(-E_p / (k_B * exact_T(x, t)))

    def src_m(x, t):
        return (15 - ufl.div(D(x, t) * ufl.grad(ex_m(x, t)))
                + k(x, t) * ex_m(x, t) * (n_trap - ex_t(x, t)) - p(x, t) * ex_t(x, t))

    def src_t(x, t):
        return 12 - k(x, t) * ex_m(x, t) * (n_trap - ex_t(x, t)) + p(x, t) * ex_t(x, t)

    hyd = F.HydrogenTransportProblem(
        mesh=mesh, subdomains=[vol, left, right],
        boundary_conditions=[
            F.FixedConcentrationBC(subdomain=left, value=ex_m, species=mobile),
            F.FixedConcentrationBC(subdomain=right, value=ex_m, species=mobile),
            F.FixedConcentrationBC(subdomain=left, value=ex_t, species=trapped),
            F.FixedConcentrationBC(subdomain=right, value=ex_t, species=trapped)],
        species=[mobile, trapped],
        reactions=[F.Reaction(reactant=[mobile, traps], product=trapped, k_0=k_0, E_k=E_k,
                              p_0=p_0, E_p=E_p, volume=vol)],
        initial_conditions=[
            F.InitialConcentration(value=lambda x: ex_m(x, t=0), species=mobile, volume=vol),
            F.InitialConcentration(value=lambda x: ex_t(x, t=0), species=trapped, volume=vol)],
        sources=[F.ParticleSource(value=src_m, volume=vol, species=mobile),
                 F.ParticleSource(value=src_t, volume=vol, species=trapped)],
        settings=F.Settings(atol=1e-10, rtol=1e-10, transient=True, stepsize=final_time / nsteps,
                            final_time=final_time))
    coupled = F.CoupledTransientHeatTransferHydrogenTransport(heat_problem=heat, hydrogen_problem=hyd)
    exact = (lambda x: exact_T(x, final_time), lambda x: ex_m(x, final_time), lambda x: ex_t(x, final_time))
    return coupled, (mobile, trapped), exact


def advection_trap_mms_2d(n=200):
    """reference test/system_tests/test_advection_term_integration.py:15-142
    (mobile < 2e-3, trapped < 7e-5 on the 200 x 200 square)."""
    D_0, E_D, k_0, E_k, p_0, E_p, n_trap = 1.2, 0.1, 2.2, 0.6, 0.5, 0.1, 100
    raw = F.create_unit_square(n, n)
    x = ufl.SpatialCoordinate(raw)
    mat = F.Material(D_0=D_0, E_D=E_D)
    vol = F.VolumeSubdomain(id=1, material=mat)
    boundary = F.SurfaceSubdomain(id=1)
    mobile = F.Species("mobile", mobile=True)
    trapped = F.Species(name="trapped", mobile=True)
    empty = F.ImplicitSpecies(n=n_trap, others=[trapped])
    T = fem.Function(fem.functionspace(raw, ("Lagrange", 1)))
    T.interpolate(lambda x: 100 + 200 * x[0] + 100 * x[1])

    def velocity_func(x):
        values = np.zeros((2, x.shape[1]))
        values[0] = x[1] * (x[1] - 1)
        return values

    vel = F.VectorFunction(raw, velocity_func)
    ex_m = lambda x: 200 * x[0] ** 2 + 300 * x[1] ** 2
    ex_t = lambda x: 10 * x[0] ** 2 + 10 * x[1] ** 2
    D = D_0 * ufl.exp(-E_D / (F.k_B * T))
    k = k_0 * ufl.exp(-E_k / (F.k_B * T))
    p = p_0 * ufl.exp(-E_p / (F.k_B * T))
    u_sym = ufl.as_vector([x[1] * (x[1] - 1), 0.0])  # the exact velocity, as in the MMS source
    f = (-ufl.div(D * ufl.grad(ex_m(x))) + ufl.inner(u_sym, ufl.grad(ex_m(x)))
         + k * ex_m(x) * (n_trap - ex_t(x)) - p * ex_t(x))
    g = -ufl.div(D * ufl.grad(ex_t(x))) - k * ex_m(x) * (n_trap - ex_t(x)) + p * ex_t(x)
    bcs = [F.FixedConcentrationBC(subdomain=boundary, value=v, species=s)
           for s, v in zip([mobile, trapped], [ex_m, ex_t])]
    m = F.HydrogenTransportProblem(
        mesh=F.Mesh(raw), subdomains=[vol, boundary], boundary_conditions=bcs,
        species=[mobile, trapped], temperature=T,
        reactions=[F.Reaction(reactant=[mobile, empty], product=trapped, k_0=k_0, E_k=E_k,
                              p_0=p_0, E_p=E_p, volume=vol)],
        sources=[F.ParticleSource(value=f, volume=vol, species=mobile),
                 F.ParticleSource(value=g, volume=vol, species=trapped)],
        advection_terms=[F.AdvectionTerm(velocity=vel, subdomain=vol, species=mobile)],
        settings=F.Settings(atol=1e-10, rtol=1e-10, transient=False))
    return m, (mobile, trapped), (ex_m, ex_t)


def multi_isotope_traps_3d(n=4, dt=1e-2, final_time=3e-2, exchange=False):
    """BASELINE configs[2] (C2): 3D H/D/T transient with 3 trap types: mobile H, D, T
    (D_0 scaled by 1/sqrt(mass)), 9 trapped species, 9 trapping reactions sharing the
    trap sites of each type between the isotopes; T = 600 + 100 x (P1 field); Sieverts
    BC at x=0 for every isotope, c=0 at x=1."""
    raw = F.create_unit_cube(n, n, n)
    mat = F.Material(D_0={"H": 4.1e-7, "D": 4.1e-7 / np.sqrt(2), "T": 4.1e-7 / np.sqrt(3)},
                     E_D={"H": 0.39, "D": 0.39, "T": 0.39})
    vol = F.VolumeSubdomain(id=1, material=mat)
    left = F.SurfaceSubdomain(id=1, locator=lambda x: np.isclose(x[0], 0))
    right = F.SurfaceSubdomain(id=2, locator=lambda x: np.isclose(x[0], 1))
    iso = [F.Species(nm) for nm in ("H", "D", "T")]
    trap_par = [(8.9e-17, 0.39, 1e13, 0.87, 1e23), (8.9e-17, 0.39, 1e13, 1.0, 4e23),
                (8.9e-17, 0.39, 1e13, 1.5, 2e23)]
    trapped, reactions = [], []
    for i, (k0, Ek, p0, Ep, dens) in enumerate(trap_par):
        group = [F.Species(f"t{s.name}_{i}", mobile=False) for s in iso]
        trapped += group
        for s, ts, mass in zip(iso, group, (1.0, 2.0, 3.0)):
            empty = F.ImplicitSpecies(n=dens, others=list(group))
            reactions.append(F.Reaction(reactant=[s, empty], product=ts, k_0=k0 / np.sqrt(mass),
                                        E_k=Ek, p_0=p0, E_p=Ep, volume=vol))
    m = F.HydrogenTransportProblem()
    m.mesh = F.Mesh(raw)
    m.subdomains = [vol, left, right]
    m.species = iso + trapped
    m.reactions = reactions
    m.temperature = lambda x: 600 + 100 * x[0]
    m.boundary_conditions = []
    for s, P in zip(iso, (100.0, 50.0, 25.0)):
        m.boundary_conditions.append(F.SievertsBC(subdomain=left, S_0=4.02e21, E_S=1.04,
                                                  pressure=P, species=s))
        m.boundary_conditions.append(F.FixedConcentrationBC(subdomain=right, value=0, species=s))
    m.settings = F.Settings(atol=1e8, rtol=1e-10, final_time=final_time)
    m.settings.stepsize = F.Stepsize(initial_value=dt)
    return m, iso, trapped


def complex_reaction(stepsize=0.5, k=350e-4, p=120e-4, c_A_0=3, final_time=200):
    """reference test/test_complex_reaction.py:39-135: A + B <-> C + D, constant rates
    (E_k = E_p = 0, the constant-p branch of reaction.py:182-187), no boundary
    conditions, TotalVolume exports."""
    m = F.HydrogenTransportProblem()
    L = 1
    m.mesh = F.Mesh1D(np.linspace(0, L, num=10))
    mat = F.Material(D_0=1, E_D=0, name="mat")
    vol = F.VolumeSubdomain1D(id=1, borders=[0, L], material=mat)
    m.subdomains = [vol, F.SurfaceSubdomain1D(id=1, x=0), F.SurfaceSubdomain1D(id=2, x=L)]
    A, B, Cc, D = (F.Species(nm) for nm in "ABCD")
    m.species = [A, B, Cc, D]
    m.reactions = [F.Reaction(k_0=k, E_k=0, p_0=p, E_p=0, reactant=[A, B], product=[Cc, D], volume=vol)]
    m.temperature = 400
    m.boundary_conditions = []
    m.initial_conditions = [F.InitialConcentration(value=c_A_0, species=A, volume=vol),
                            F.InitialConcentration(value=c_A_0, species=B, volume=vol)]
    m.exports = [F.TotalVolume(s, vol) for s in (A, B, Cc, D)]
    m.settings = F.Settings(atol=1e-10, rtol=1e-10, max_iterations=30, final_time=final_time)
    m.settings.stepsize = F.Stepsize(initial_value=stepsize)
    return m


def concentration_A_exact(t, c_A_0, k, p):
    """reference test/test_complex_reaction.py:7-36."""
    A = 1 - p / k
    B = 2 * p / k * c_A_0
    Cq = -p / k * c_A_0**2
    roots = np.roots([A, B, Cq])
    a, b = (roots[0], roots[0]) if len(roots) == 1 else (roots[0], roots[1])
    Fq = ((c_A_0 - a) / (c_A_0 - b)) * np.exp(-(a - b) * (k - p) * t)
    return (a - b * Fq) / (1 - Fq)


def non_homogeneous_density(density_func, final_time=100):
    """reference test/system_tests/test_non_homogeneous_density.py:26-85: trap filling
    with a space/time dependent trap density, no detrapping (p_0 = 0)."""
    m = F.HydrogenTransportProblem()
    m.mesh = F.Mesh1D(np.linspace(0, 1, 100))
    vol = F.VolumeSubdomain1D(id=1, borders=[0, 1], material=F.Material(name="mat", D_0=1, E_D=0))
    left, right = F.SurfaceSubdomain1D(id=1, x=0), F.SurfaceSubdomain1D(id=2, x=1)
    m.subdomains = [vol, left, right]
    H = F.Species("H")
    trapped = F.Species("trapped_H", mobile=False)
    empty = F.ImplicitSpecies(n=density_func, name="empty", others=[trapped])
    m.species = [H, trapped]
    m.reactions = [F.Reaction(reactant=[H, empty], product=trapped, k_0=1, E_k=0, p_0=0, E_p=0, volume=vol)]
    m.temperature = 600
    m.boundary_conditions = [F.DirichletBC(subdomain=left, value=1, species=H),
                             F.DirichletBC(subdomain=right, value=1, species=H)]
    m.settings = F.Settings(atol=1e-10, rtol=1e-10, final_time=final_time)
    m.settings.stepsize = 1
    total = F.TotalVolume(field=trapped, volume=vol)
    m.exports = [total]
    return m, total


DENSITIES = {
    "step_space": (lambda x: ufl.conditional(ufl.gt(x[0], 0.5), 10, 0.0), 5.0),
    "simple_space": (lambda x: 10 - 10 * x[0], 5.0),
    "step_time_space": (lambda x, t: ufl.conditional(ufl.gt(t, 50), 10 - 10 * x[0], 0.0), 5.0),
    "step_time": (lambda t: 10.0 if t > 50 else 2.0, 10.0),
}


def recombination_flux(dim, n):
    """Two isotopes in 2D/3D with a T-dependent prescribed influx g(x, T) on the left face and
    surface reactions H + D <-> HD, H + H <-> H2 (recombination K_r c^2 and dissociation
    K_d P) on the right face; T = 500 + 100 y.  Exercises facet terms that depend on x, T and
    on several species (facet Jacobian blocks) away from 1D, where a facet is a point.
    Transient (three implicit Euler steps from zero): the steady problem with flux conditions
    only has a spurious negative root Newton finds from a zero guess."""
    raw = F.create_unit_square(n, n) if dim == 2 else F.create_unit_cube(n, n, n)
    mesh = F.Mesh(raw)
    vol = F.VolumeSubdomain(id=1, material=F.Material(D_0=1.5, E_D=0.05))
    left = F.SurfaceSubdomain(id=1, locator=lambda x: np.isclose(x[0], 0))
    right = F.SurfaceSubdomain(id=2, locator=lambda x: np.isclose(x[0], 1))
    m = F.HydrogenTransportProblem()
    m.mesh = mesh
    m.subdomains = [vol, left, right]
    H, D = F.Species("H"), F.Species("D")
    m.species = [H, D]
    m.temperature = lambda x: 500 + 100 * x[1]
    m.boundary_conditions = [
        F.ParticleFluxBC(subdomain=left, value=lambda x, T: 2.0 + x[1] + 1e-3 * T, species=H),
        F.ParticleFluxBC(subdomain=left, value=lambda x: 1.0 + 0.5 * x[1] ** 2, species=D),
        F.SurfaceReactionBC(reactant=[H, D], gas_pressure=0.5, k_r0=0.3, E_kr=0.02, k_d0=0.2, E_kd=0.01,
                            subdomain=right),
        F.SurfaceReactionBC(reactant=[H, H], gas_pressure=1.0, k_r0=0.5, E_kr=0.03, k_d0=0.1, E_kd=0.0,
                            subdomain=right),
        F.SurfaceReactionBC(reactant=[D, D], gas_pressure=0.0, k_r0=0.4, E_kr=0.0, k_d0=0.0, E_kd=0.0,
                            subdomain=right),
    ]
    m.settings = F.Settings(atol=1e-10, rtol=1e-10, transient=True, final_time=0.3)
    m.settings.stepsize = F.Stepsize(0.1)
    return m, (H, D), None


def monoblock_coupled_3d(n=4, nsteps=3, dt=0.5):
    """BASELINE configs[3] (C3) in small: coupled heat transfer + hydrogen transport on a cube
    tagged into three material slabs W / Cu / CuCrZr (tagged cube standing in for the
    ITER-like monoblock; ``VolumeSubdomain(locator=...)``, reference volume_subdomain.py:49-56).
    Heat: per-material conductivity / density / heat capacity, heat flux on the plasma-facing top,
    fixed temperature on the coolant side.  Hydrogen: continuous problem, mobile H with
    per-material Arrhenius diffusivity, two traps in W, one in Cu, one in CuCrZr (each reaction
    restricted to its material through ``Reaction.volume``), fixed concentration on top, zero at
    the bottom.  ``n`` must be a multiple of 4 (slab borders at z = 1/4 and 1/2).

    Reference quirk kept on both sides of the parity test: the coupled problem hands
    ``heat_problem.u`` to the hydrogen problem before the first heat solve
    (coupled_heat_hydrogen_problem.py:119-131), the DG1 diffusivity behind ``SurfaceFlux`` is
    interpolated once from that all-zero field (hydrogen_transport_problem.py:632-670; refreshed
    only for a callable T(t), :1049-1056), so the exported surface flux is exactly 0 while the
    transport itself uses the live temperature."""
    assert n % 4 == 0
    raw = F.create_unit_cube(n, n, n)
    mats = {
        "W": F.Material(D_0=1.5, E_D=0.10, thermal_conductivity=1.7, density=1.9, heat_capacity=1.3),
        "Cu": F.Material(D_0=2.4, E_D=0.09, thermal_conductivity=4.0, density=0.9, heat_capacity=3.8),
        "CuCrZr": F.Material(D_0=1.4, E_D=0.11, thermal_conductivity=3.2, density=0.9, heat_capacity=3.9),
    }
    eps = 1e-9
    W = F.VolumeSubdomain(id=1, material=mats["W"], locator=lambda x: x[2] >= 0.5 - eps)
    Cu = F.VolumeSubdomain(id=2, material=mats["Cu"],
                           locator=lambda x: np.logical_and(x[2] >= 0.25 - eps, x[2] <= 0.5 + eps))
    CuCrZr = F.VolumeSubdomain(id=3, material=mats["CuCrZr"], locator=lambda x: x[2] <= 0.25 + eps)
    top = F.SurfaceSubdomain(id=4, locator=lambda x: np.isclose(x[2], 1.0))
    coolant = F.SurfaceSubdomain(id=5, locator=lambda x: np.isclose(x[2], 0.0))
    subdomains = [W, Cu, CuCrZr, top, coolant]
    final_time = nsteps * dt
    heat = F.HeatTransferProblem(
        mesh=F.Mesh(raw), subdomains=subdomains,
        boundary_conditions=[F.HeatFluxBC(subdomain=top, value=lambda t: 40.0 + 100.0 * t),
                             F.FixedTemperatureBC(subdomain=coolant, value=420.0)],
        initial_condition=F.InitialTemperature(value=lambda x: 420.0 + 30.0 * x[2], volume=W),
        settings=F.Settings(atol=1e-10, rtol=1e-10, transient=True, stepsize=dt, final_time=final_time))
    H = F.Species("H")
    traps = [F.Species(nm, mobile=False) for nm in ("W_t1", "W_t2", "Cu_t", "CuCrZr_t")]
    par = [(W, 20.0, 0.10, 3.0, 0.20, 1.5), (W, 12.0, 0.10, 2.0, 0.30, 0.8),
           (Cu, 25.0, 0.09, 4.0, 0.15, 0.5), (CuCrZr, 18.0, 0.11, 3.5, 0.25, 0.6)]
    reactions = []
    for ts, (vol, k0, Ek, p0, Ep, dens) in zip(traps, par):
        empty = F.ImplicitSpecies(n=dens, others=[ts])
        reactions.append(F.Reaction(reactant=[H, empty], product=ts, k_0=k0, E_k=Ek, p_0=p0, E_p=Ep, volume=vol))
    hyd = F.HydrogenTransportProblem(
        mesh=F.Mesh(raw), subdomains=subdomains, species=[H] + traps, reactions=reactions,
        boundary_conditions=[F.FixedConcentrationBC(subdomain=top, value=1.0, species=H),
                             F.FixedConcentrationBC(subdomain=coolant, value=0.0, species=H)],
        settings=F.Settings(atol=1e-10, rtol=1e-10, transient=True, stepsize=dt, final_time=final_time))
    hyd.exports = [F.TotalVolume(field=H, volume=W), F.TotalVolume(field=traps[0], volume=W),
                   F.TotalVolume(field=traps[2], volume=Cu), F.SurfaceFlux(field=H, surface=coolant)]
    coupled = F.CoupledTransientHeatTransferHydrogenTransport(heat_problem=heat, hydrogen_problem=hyd)
    return coupled, [H] + traps


def permeation_advection_2d(n=8, nsteps=4, dt=2.0, L=3e-4):
    """BASELINE configs[4] (C4) in small: transient 2D permeation with advection and
    Sieverts / Henry surface conditions: square of side ``L``, mobile H, Poiseuille-like
    velocity ``(a y (L - y), 0)`` as in test/system_tests/test_advection_term_integration.py:57-66,
    ``SievertsBC`` upstream with the parameters of test/test_permeation_problem.py:60-66,
    ``HenrysBC`` downstream; the outgoing flux and the inventory are recorded."""
    raw = F.create_rectangle(((0.0, 0.0), (L, L)), (n, n))
    mat = F.Material(D_0=1.9e-7, E_D=0.2)
    vol = F.VolumeSubdomain(id=1, material=mat)
    left = F.SurfaceSubdomain(id=1, locator=lambda x: np.isclose(x[0], 0.0))
    right = F.SurfaceSubdomain(id=2, locator=lambda x: np.isclose(x[0], L))
    H = F.Species("H")
    a = 4.0 * 2e-6 / L**2          # peak velocity 2e-6 m/s: cell Peclet number below one

    def velocity_func(x):
        values = np.zeros((2, x.shape[1]))
        values[0] = a * x[1] * (L - x[1])
        return values

    vel = F.VectorFunction(raw, velocity_func)
    m = F.HydrogenTransportProblem(
        mesh=F.Mesh(raw), subdomains=[vol, left, right], species=[H], temperature=500.0,
        boundary_conditions=[
            F.SievertsBC(subdomain=left, S_0=4.02e21, E_S=1.04, pressure=100, species=H),
            F.HenrysBC(subdomain=right, H_0=6.0e17, E_H=0.9, pressure=lambda t: 1.0 + 0.5 * t, species=H)],
        advection_terms=[F.AdvectionTerm(velocity=vel, subdomain=vol, species=H)],
        settings=F.Settings(atol=1e-10, rtol=1e-10, transient=True, stepsize=dt, final_time=nsteps * dt))
    m.exports = [F.SurfaceFlux(field=H, surface=right), F.TotalVolume(field=H, volume=vol)]
    return m, H


def species_dependent_source_mms(n=9999):
    """reference test/system_tests/test_1_mobile_species.py:285-347: species A carries a source
    ``B**2`` of a second species B whose manufactured solution is linear (exact in P1), plus the
    manufactured source that makes ``A = 1 + sin(2 pi x)`` the solution (L2 < 1e-3)."""
    A_exact = lambda mod: (lambda x: 1 + mod.sin(2 * mod.pi * x[0]))
    B_exact = lambda x: 1 + x[0]
    m = F.HydrogenTransportProblem()
    m.mesh = F.Mesh1D(np.linspace(0, 1, n + 1))
    x = ufl.SpatialCoordinate(m.mesh.mesh)
    vol = F.VolumeSubdomain1D(id=1, borders=[0, 1], material=F.Material(name="mat", D_0=1, E_D=0))
    left, right = F.SurfaceSubdomain1D(id=1, x=0), F.SurfaceSubdomain1D(id=2, x=1)
    m.subdomains = [vol, left, right]
    A, B = F.Species("A"), F.Species("B")
    m.species = [A, B]
    m.temperature = 500
    m.boundary_conditions = [
        F.DirichletBC(subdomain=left, value=A_exact(ufl), species=A),
        F.DirichletBC(subdomain=right, value=A_exact(ufl), species=A),
        F.DirichletBC(subdomain=left, value=B_exact, species=B),
        F.DirichletBC(subdomain=right, value=B_exact, species=B),
    ]
    f = -ufl.div(1 * ufl.grad(A_exact(ufl)(x))) - B_exact(x) ** 2
    m.sources = [F.ParticleSource(value=lambda B: B**2, volume=vol, species=A, species_dependent_value={"B": B}),
                 F.ParticleSource(value=f, volume=vol, species=A)]
    m.settings = F.Settings(atol=1e-10, rtol=1e-10, transient=False)
    return m, (A, B), (A_exact(np), B_exact)


def coupled_non_matching_mms(n_heat=499, n_hyd=1999, nsteps=20):
    """reference test/system_tests/test_coupled_heat_hydrogen.py:210-332: coupled heat + hydrogen
    MMS with the heat problem on a coarse and the hydrogen problem on a fine mesh (L2 < 2e-7)."""
    density, heat_capacity, lam = 1.3, 2.3, 4.1
    D_0, E_D, k_B, final_time = 1.3, 0.2, F.k_B, 5
    mat = F.Material(D_0=D_0, E_D=E_D, thermal_conductivity=lam, density=density, heat_capacity=heat_capacity)
    vol = F.VolumeSubdomain1D(id=1, borders=[0, 1], material=mat)
    left, right = F.SurfaceSubdomain1D(id=2, x=0), F.SurfaceSubdomain1D(id=3, x=1)
    mobile = F.Species("mobile", mobile=True)
    exact_T = lambda x, t: 2 * x[0] ** 2 + 5 * t
    heat = F.HeatTransferProblem(
        mesh=F.Mesh1D(vertices=np.linspace(0, 1, n_heat + 1)), subdomains=[vol, left, right],
        boundary_conditions=[F.FixedTemperatureBC(subdomain=left, value=exact_T),
                             F.FixedTemperatureBC(subdomain=right, value=exact_T)],
        sources=[F.HeatSource(value=density * heat_capacity * 5 - lam * 4, volume=vol)],
        initial_condition=F.InitialTemperature(lambda x: exact_T(x, 0), volume=vol),
        settings=F.Settings(atol=1e-10, rtol=1e-10, transient=True, final_time=final_time,
                            stepsize=final_time / nsteps))
    exact_c = lambda x, t: 4 * x[0] ** 2 + 10 * t
    D = lambda x, t: D_0 * ufl.exp(-E_D / (k_B * exact_T(x, t)))
    source = lambda x, t: 10 - ufl.div(D(x, t) * ufl.grad(exact_c(x, t)))
    hyd = F.HydrogenTransportProblem(
        mesh=F.Mesh1D(vertices=np.linspace(0, 1, n_hyd + 1)), subdomains=[vol, left, right],
        boundary_conditions=[F.FixedConcentrationBC(subdomain=left, value=exact_c, species=mobile),
                             F.FixedConcentrationBC(subdomain=right, value=exact_c, species=mobile)],
        species=[mobile],
        initial_conditions=[F.InitialConcentration(value=lambda x: exact_c(x, t=0), species=mobile, volume=vol)],
        sources=[F.ParticleSource(value=source, volume=vol, species=mobile)],
        settings=F.Settings(atol=1e-10, rtol=1e-10, transient=True, final_time=final_time,
                            stepsize=final_time / nsteps))
    coupled = F.CoupledTransientHeatTransferHydrogenTransport(heat_problem=heat, hydrogen_problem=hyd)
    return coupled, mobile, (lambda x: exact_c(x, final_time))
