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
