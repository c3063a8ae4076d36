"""More of the reference's own tests for the hot path, run against the CPU oracle (and, as
parity tests, on the GPU box): test/test_multispecies_problem.py:31-151 (two species
permeating with per-species diffusivities given as dicts, species addressed by name,
Sieverts BCs, analytical flux within 1 %)."""
import numpy as np
import pytest

import festim_b200 as F
from festim_b200 import fem
from oracle.newton import OracleBackend


def _analytical_error(D, permeability, computed_flux, L, times, P_up):
    n = np.arange(1, 10000)[:, np.newaxis]
    summation = np.sum((-1) ** n * np.exp(-((np.pi * n) ** 2) * float(D) / L ** 2 * times), axis=0)
    analytical = np.abs(P_up ** 0.5 * permeability / L * (2 * summation + 1))
    keep = np.where(analytical > 0.1 * np.max(analytical))
    return np.abs((computed_flux[keep] - analytical[keep]) / analytical[keep]).mean()


def _multispecies(final_time=10):
    L = 3e-04
    mesh = F.Mesh1D(np.linspace(0, L, num=1001))
    m = F.HydrogenTransportProblem()
    m.mesh = mesh
    mat = F.Material(D_0={"spe_1": 1.9e-7, "spe_2": 3.8e-7}, E_D={"spe_1": 0.2, "spe_2": 0.2}, name="my_mat")
    vol = F.VolumeSubdomain1D(id=1, borders=[0, L], material=mat)
    left, right = F.SurfaceSubdomain1D(id=1, x=0), F.SurfaceSubdomain1D(id=2, x=L)
    m.subdomains = [vol, left, right]
    s1, s2 = F.Species("spe_1"), F.Species("spe_2")
    m.species = [s1, s2]
    T = fem.Constant(mesh.mesh, 500.0)
    m.temperature = T
    m.boundary_conditions = [
        F.DirichletBC(subdomain=right, value=0, species="spe_1"),
        F.DirichletBC(subdomain=right, value=0, species="spe_2"),
        F.SievertsBC(subdomain=left, S_0=4.02e21, E_S=1.04, pressure=100, species="spe_1"),
        F.SievertsBC(subdomain=left, S_0=5.0e21, E_S=1.2, pressure=100, species="spe_2"),
    ]
    m.settings = F.Settings(atol=1e2, rtol=1e-10, max_iterations=30, final_time=final_time)
    m.settings.stepsize = F.Stepsize(initial_value=0.08)
    ex = [F.SurfaceFlux(field=s1, surface=right), F.SurfaceFlux(field=s2, surface=right),
          F.TotalVolume(field=s1, volume=vol), F.TotalVolume(field=s2, volume=vol)]
    m.exports = ex
    return m, mat, (s1, s2), ex, L


def _check_against_series(m, mat, species, ex, L):
    times = np.array(ex[0].t)
    for k, spe in enumerate(species):
        bc = m.boundary_conditions[2 + k]
        D = float(mat.get_D_0(spe)) * np.exp(-float(mat.get_E_D(spe)) / F.k_B / 500.0)
        S = float(bc.S_0) * np.exp(-float(bc.E_S) / F.k_B / 500.0)
        err = _analytical_error(D, D * S, np.abs(np.array(ex[k].data)), L, times, float(bc.pressure))
        assert err < 0.01


def test_multispecies_permeation_problem():
    m, mat, species, ex, L = _multispecies()
    m.solver_backend = OracleBackend()
    m.initialise()
    m.run()
    _check_against_series(m, mat, species, ex, L)
    # inventories grow monotonically towards the linear steady profile
    for tv in ex[2:]:
        assert np.all(np.diff(tv.data) >= -1e-9 * max(tv.data))


@pytest.mark.gpu
def test_multispecies_permeation_gpu_parity():
    g, mat, species, exg, L = _multispecies()
    r, _, _, exr, _ = _multispecies()
    r.solver_backend = OracleBackend()
    for m in (g, r):
        m.initialise()
        m.run()
    _check_against_series(g, mat, species, exg, L)
    for a, b in zip(exg, exr):
        assert np.allclose(a.data, b.data, rtol=1e-6, atol=1e-6 * max(np.abs(b.data)))
    assert np.linalg.norm(g.u.x.array - r.u.x.array) / np.linalg.norm(r.u.x.array) < 1e-8


# ---------------------------------------------------------------------------------------
# derived quantities on a prescribed field (reference test/test_total_volume.py:13-46,
# test/test_surface_quantity.py:12-56, and the Average/Minimum/Maximum exports)
# ---------------------------------------------------------------------------------------
def _prescribed_1d(backend):
    L, D = 4.0, 1.5
    m = F.HydrogenTransportProblem()
    m.mesh = F.Mesh1D(np.linspace(0, L, 10000))
    vol = F.VolumeSubdomain1D(id=1, borders=[0, L], material=F.Material(D_0=D, E_D=0, name="dummy"))
    left, right = F.SurfaceSubdomain1D(id=1, x=0), F.SurfaceSubdomain1D(id=2, x=L)
    m.subdomains = [vol, left, right]
    H = F.Species("H")
    m.species = [H]
    m.temperature = 400
    m.boundary_conditions = [F.DirichletBC(subdomain=left, value=1, species=H)]
    m.settings = F.Settings(atol=1e-10, rtol=1e-10, transient=False)
    if backend is not None:
        m.solver_backend = backend
    m.initialise()
    x = m.mesh.mesh.coords[:, 0]
    m.u.x.array[:] = 2 * x ** 2 + 1
    m.update_post_processing_solutions()
    return m, H, vol, right, L, D


def _check_prescribed_1d(m, H, vol, right, L, D):
    tv = F.TotalVolume(field=H, volume=vol)
    tv.compute(m)
    assert np.isclose(tv.value, 2 / 3 * L ** 3 + L, rtol=1e-2)
    sf = F.SurfaceFlux(field=H, surface=right)
    sf.compute(m)
    assert np.isclose(sf.value, -D * 4 * L, rtol=1e-2)          # -D dc/dx at x = L, outward normal +x
    av = F.AverageVolume(field=H, volume=vol)
    av.compute(m)
    assert np.isclose(av.value, (2 / 3 * L ** 3 + L) / L, rtol=1e-2)
    mx, mn = F.MaximumVolume(field=H, volume=vol), F.MinimumVolume(field=H, volume=vol)
    mx.compute(m)
    mn.compute(m)
    assert np.isclose(mx.value, 2 * L ** 2 + 1) and np.isclose(mn.value, 1.0)
    ts = F.TotalSurface(field=H, surface=right)
    ts.compute(m)
    assert np.isclose(ts.value, 2 * L ** 2 + 1)
    return tv.value, sf.value, av.value


def test_derived_quantities_on_prescribed_field_oracle():
    _check_prescribed_1d(*_prescribed_1d(OracleBackend()))


@pytest.mark.gpu
def test_derived_quantities_on_prescribed_field_gpu():
    vals_g = _check_prescribed_1d(*_prescribed_1d(None))
    vals_r = _check_prescribed_1d(*_prescribed_1d(OracleBackend()))
    assert np.allclose(vals_g, vals_r, rtol=1e-6)


# ---------------------------------------------------------------------------------------
# D_global (reference test/test_h_transport_problem.py:245-320): the diffusivity the surface
# flux uses is the DG1 interpolant of the owning cell's Arrhenius law at the vertex
# temperature - different temperatures and different materials at the two ends.  Observed
# through SurfaceFlux on u = x (du/dx = 1): flux(x=L) = -D(L), flux(x=0) = +D(0).
# ---------------------------------------------------------------------------------------
def _d_global_case(kind, backend):
    if kind == "temperatures":
        mats = [F.Material(D_0=1.5, E_D=0.1, name="m")] * 2
        temperature = lambda x: 100.0 * x[0] + 50
        expected = [1.5 * np.exp(-0.1 / (F.k_B * 50)), 1.5 * np.exp(-0.1 / (F.k_B * 450))]
    else:
        mats = [F.Material(D_0=1.0, E_D=0.1, name="L"), F.Material(D_0=2.0, E_D=0.2, name="R")]
        temperature = 500
        expected = [1.0 * np.exp(-0.1 / (F.k_B * 500)), 2.0 * np.exp(-0.2 / (F.k_B * 500))]
    m = F.HydrogenTransportProblem()
    m.mesh = F.Mesh1D(np.linspace(0, 4, num=101))
    vols = [F.VolumeSubdomain1D(id=1, borders=[0, 2], material=mats[0]),
            F.VolumeSubdomain1D(id=2, borders=[2, 4], material=mats[1])]
    left, right = F.SurfaceSubdomain1D(id=1, x=0), F.SurfaceSubdomain1D(id=2, x=4)
    m.subdomains = vols + [left, right]
    H = F.Species("H")
    m.species = [H]
    m.temperature = temperature
    m.boundary_conditions = [F.DirichletBC(subdomain=left, value=0, species=H)]
    m.settings = F.Settings(atol=1e-10, rtol=1e-10, transient=False)
    if backend is not None:
        m.solver_backend = backend
    m.initialise()
    m.u.x.array[:] = m.mesh.mesh.coords[:, 0]
    m.update_post_processing_solutions()
    out = []
    for surf in (left, right):
        sf = F.SurfaceFlux(field=H, surface=surf)
        sf.compute(m)
        out.append(sf.value)
    return out, expected


@pytest.mark.parametrize("kind", ["temperatures", "materials"])
def test_D_global_through_surface_flux_oracle(kind):
    (f_left, f_right), (D_left, D_right) = _d_global_case(kind, OracleBackend())
    assert np.isclose(f_left, D_left) and np.isclose(f_right, -D_right)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["temperatures", "materials"])
def test_D_global_through_surface_flux_gpu(kind):
    (f_left, f_right), (D_left, D_right) = _d_global_case(kind, None)
    assert np.isclose(f_left, D_left, rtol=1e-10) and np.isclose(f_right, -D_right, rtol=1e-10)


# ---------------------------------------------------------------------------------------
# reference test/system_tests/test_reaction_not_in_every_volume.py: a steady problem whose
# reaction exists in one of two volumes only (the immobile species gets the filler term
# c v dx in the other one, hydrogen_transport_problem.py:980-997)
# ---------------------------------------------------------------------------------------
def _reaction_not_everywhere(backend, with_bc=False):
    m = F.HydrogenTransportProblem()
    m.mesh = F.Mesh1D(np.linspace(0, 2e-04, num=1001))
    mat = F.Material(D_0=4.1e-07, E_D=0.37, name="my_mat")
    vol_1 = F.VolumeSubdomain1D(id=1, borders=[0, 1e-04], material=mat)
    vol_2 = F.VolumeSubdomain1D(id=2, borders=[1e-04, 2e-04], material=mat)
    left, right = F.SurfaceSubdomain1D(id=1, x=0), F.SurfaceSubdomain1D(id=2, x=2e-04)
    m.subdomains = [vol_1, vol_2, left, right]
    cm, ct = F.Species("cm"), F.Species("ct", mobile=False)
    m.species = [cm, ct]
    trap = F.ImplicitSpecies(n=8.19e25, others=[ct])
    m.reactions = [F.Reaction(reactant=[cm, trap], product=ct, k_0=8.9e-17, E_k=0.39, p_0=1e13, E_p=0.87,
                              volume=vol_1)]
    m.temperature = 600
    if with_bc:   # something to solve: fixed mobile concentration at both ends
        m.boundary_conditions = [F.DirichletBC(subdomain=left, value=1e20, species=cm),
                                 F.DirichletBC(subdomain=right, value=0, species=cm)]
    m.settings = F.Settings(atol=1e10, rtol=1e-10, transient=False)
    if backend is not None:
        m.solver_backend = backend
    m.initialise()
    m.run()
    return m, cm, ct


def test_sim_reaction_not_in_every_volume():
    _reaction_not_everywhere(OracleBackend())
    m, cm, ct = _reaction_not_everywhere(OracleBackend(), with_bc=True)
    x = m.mesh.mesh.coords[:, 0]
    trapped = ct.post_processing_solution.x.array
    # no trapping where there is no reaction (the P1 field decays over a few vertices past the
    # interface: the shared vertex belongs to both volumes)
    assert np.abs(trapped[x > 1.1e-4]).max() <= 1e-12 * trapped.max()
    assert trapped[x < 0.5e-4].min() > 0


@pytest.mark.gpu
def test_sim_reaction_not_in_every_volume_gpu():
    _reaction_not_everywhere(None)
    g, cmg, ctg = _reaction_not_everywhere(None, with_bc=True)
    r, cmr, ctr = _reaction_not_everywhere(OracleBackend(), with_bc=True)
    for a, b in ((cmg, cmr), (ctg, ctr)):
        ua, ub = a.post_processing_solution.x.array, b.post_processing_solution.x.array
        assert np.linalg.norm(ua - ub) / np.linalg.norm(ub) < 1e-8


# ---------------------------------------------------------------------------------------
# CustomQuantity (reference exports/custom_quantity.py; docs example and
# test/system_tests/test_misc.py: custom quantities on volumes and surfaces)
# ---------------------------------------------------------------------------------------
def _custom_quantity_case(dim, backend):
    from festim_b200 import ufl

    if dim == 1:
        mesh = F.Mesh1D(np.linspace(0, 2, 401))
        vol = F.VolumeSubdomain1D(id=1, borders=[0, 2], material=F.Material(D_0=1.5, E_D=0.1, name="m"))
        left, right = F.SurfaceSubdomain1D(id=1, x=0), F.SurfaceSubdomain1D(id=2, x=2)
    else:
        raw = F.create_unit_square(12, 12) if dim == 2 else F.create_unit_cube(5, 5, 5)
        mesh = F.Mesh(raw)
        vol = F.VolumeSubdomain(id=1, material=F.Material(D_0=1.5, E_D=0.1, name="m"))
        left = F.SurfaceSubdomain(id=1, locator=lambda x: np.isclose(x[0], 0))
        right = F.SurfaceSubdomain(id=2, locator=lambda x: np.isclose(x[0], 1))
    m = F.HydrogenTransportProblem()
    m.mesh = mesh
    m.subdomains = [vol, left, right]
    A, B = F.Species("A"), F.Species("B")
    m.species = [A, B]
    m.temperature = lambda x: 500 + 100 * x[0]
    m.boundary_conditions = [F.DirichletBC(subdomain=left, value=1, species=A),
                             F.DirichletBC(subdomain=left, value=2, species=B)]
    total = F.CustomQuantity(expr=lambda **kw: kw["A"] + 2 * kw["B"], subdomain=vol, title="Total quantity")
    flux = F.CustomQuantity(expr=lambda **kw: -kw["D_A"] * ufl.dot(ufl.grad(kw["A"]), kw["n"]), subdomain=right,
                            title="Surface flux")
    weighted = F.CustomQuantity(expr=lambda **kw: kw["x"][0] * kw["A"] * ufl.exp(-0.05 / (F.k_B * kw["T"])),
                                subdomain=vol, title="weighted")
    builtin_flux = F.SurfaceFlux(field=A, surface=right)
    builtin_total = F.TotalVolume(field=A, volume=vol)
    m.exports = [total, flux, weighted, builtin_flux, builtin_total]
    m.settings = F.Settings(atol=1e-10, rtol=1e-10, transient=False)
    if backend is not None:
        m.solver_backend = backend
    m.initialise()
    x = m.mesh.mesh.coords
    ns = 2
    m.u.x.array[0::ns] = 1 + x[:, 0] ** 2 + (x[:, 1] if dim > 1 else 0.0)
    m.u.x.array[1::ns] = 2 - 0.5 * x[:, 0]
    m.update_post_processing_solutions()
    for e in m.exports:
        e.compute(m)
    return [e.value for e in m.exports], m


def test_custom_quantity_oracle():
    (total, flux, weighted, bflux, btotal), m = _custom_quantity_case(1, OracleBackend())
    # int_0^2 (1 + x^2) + 2 (2 - x/2) dx = (2 + 8/3) + (8 - 2)
    assert np.isclose(total, 2 + 8 / 3 + 6, rtol=1e-5)
    # the custom surface flux -D_A grad(A).n reproduces the built-in SurfaceFlux
    assert np.isclose(flux, bflux, rtol=1e-12)
    assert np.isclose(flux, -1.5 * np.exp(-0.1 / (F.k_B * 700)) * 4.0, rtol=1e-2)
    assert weighted > 0 and np.isclose(btotal, 2 + 8 / 3, rtol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("dim", [1, 2, 3])
def test_custom_quantity_gpu_parity(dim):
    vals_g, _ = _custom_quantity_case(dim, None)
    vals_r, _ = _custom_quantity_case(dim, OracleBackend())
    assert np.allclose(vals_g, vals_r, rtol=1e-10)
    assert np.isclose(vals_g[1], vals_g[3], rtol=1e-10)   # custom flux == built-in flux


def test_define_D_global_is_the_DG1_field_behind_surface_flux():
    """reference hydrogen_transport_problem.py:632-670 and test/test_h_transport_problem.py:245-320:
    per-cell Arrhenius diffusivity of the cell's material at the cell's vertex temperatures -
    discontinuous at a material interface - and the weight of the gradient in SurfaceFlux."""
    D_0L, E_DL, D_0R, E_DR = 1.0, 0.1, 2.0, 0.2
    H = F.Species("H")
    left, right = F.SurfaceSubdomain1D(id=3, x=0.0), F.SurfaceSubdomain1D(id=4, x=4.0)
    m = F.HydrogenTransportProblem(
        mesh=F.Mesh1D(np.linspace(0, 4, num=101)),
        subdomains=[F.VolumeSubdomain1D(id=1, borders=[0, 2], material=F.Material(D_0=D_0L, E_D=E_DL)),
                    F.VolumeSubdomain1D(id=2, borders=[2, 4], material=F.Material(D_0=D_0R, E_D=E_DR)),
                    left, right],
        species=[H], temperature=lambda x: 100.0 * x[0] + 50,
        boundary_conditions=[F.FixedConcentrationBC(left, value=1.0, species=H),
                             F.FixedConcentrationBC(right, value=0.0, species=H)],
        settings=F.Settings(atol=1e-10, rtol=1e-10, transient=False))
    flux = F.SurfaceFlux(field=H, surface=right)
    m.exports = [flux]
    m.solver_backend = OracleBackend()
    m.initialise()
    D, _ = m.define_D_global(H)
    mesh = m.mesh.mesh
    vals = D.x.array.reshape(mesh.num_cells, 2)
    x = mesh.coords[mesh.cells][:, :, 0]
    in_left = x.mean(axis=1) < 2
    expected = np.where(in_left[:, None], D_0L * np.exp(-E_DL / (F.k_B * (100 * x + 50))),
                        D_0R * np.exp(-E_DR / (F.k_B * (100 * x + 50))))
    assert np.allclose(vals, expected, rtol=1e-14)
    # two values at the interface vertex x = 2, one per material
    at_2 = np.isclose(x, 2.0)
    assert not np.isclose(vals[at_2 & in_left[:, None]][0], vals[at_2 & ~in_left[:, None]][0])
    m.run()
    u = m.u.x.array
    grad = (u[-1] - u[-2]) / (mesh.coords[-1, 0] - mesh.coords[-2, 0])
    assert np.isclose(flux.data[-1], -vals[-1, -1] * grad, rtol=1e-12)


def test_map_surface_to_volume_subdomains():
    """reference subdomain/volume_subdomain.py:172-260 (test/test_subdomains.py): surfaces map to
    the volume whose cells they bound; a surface spanning two volumes is an error."""
    mesh = F.create_unit_square(4, 4)
    mat = F.Material(D_0=1, E_D=0)
    vl = F.VolumeSubdomain(id=1, material=mat, locator=lambda x: x[0] <= 0.5 + 1e-12)
    vr = F.VolumeSubdomain(id=2, material=mat, locator=lambda x: x[0] >= 0.5 - 1e-12)
    sl = F.SurfaceSubdomain(id=3, locator=lambda x: np.isclose(x[0], 0.0))
    sr = F.SurfaceSubdomain(id=4, locator=lambda x: np.isclose(x[0], 1.0))
    ft, ct = F.Mesh(mesh).define_meshtags([sl, sr], [vl, vr])
    got = F.map_surface_to_volume_subdomains(ft, ct, mesh.bfacet_cell, [vl, vr], [sl, sr])
    assert got == {sl: vl, sr: vr}
    top = F.SurfaceSubdomain(id=5, locator=lambda x: np.isclose(x[1], 1.0))
    ft2, ct2 = F.Mesh(mesh).define_meshtags([top], [vl, vr])
    with pytest.raises(AssertionError, match="connected to multiple volume subdomains"):
        F.map_surface_to_volume_subdomains(ft2, ct2, mesh.bfacet_cell, [vl, vr], [top])
