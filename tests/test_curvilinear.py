"""Cylindrical and spherical coordinates: the reference's MMS tests
test/system_tests/test_cylindrical_coordinates.py:9-65 and
test_spherical_coordinates.py (L2 error < 1e-6), on the CPU oracle and - as parity
tests - through the CUDA path."""
import numpy as np
import pytest

import festim_b200 as F
from cases import error_L2
from oracle.newton import OracleBackend


def _mms(system, n, transient=False):
    mesh = F.Mesh1D(vertices=np.linspace(1, 2, n), coordinate_system=system)
    if system == "cylindrical":
        u_exact, f = (lambda x: 1 + x[0] ** 2), -4       # -(1/r) d/dr (r du/dr) = -4
    else:
        u_exact, f = (lambda x: 3 + 2 * x[0] ** 2), -12  # -(1/r^2) d/dr (r^2 du/dr) = -12
    mat = F.Material(D_0=1.0, E_D=0)
    left, right = F.SurfaceSubdomain1D(id=1, x=1), F.SurfaceSubdomain1D(id=2, x=2)
    vol = F.VolumeSubdomain1D(id=3, borders=[1, 2], material=mat)
    H = F.Species("H")
    m = F.HydrogenTransportProblem(
        mesh=mesh, species=[H], subdomains=[vol, left, right],
        boundary_conditions=[F.FixedConcentrationBC(subdomain=left, value=u_exact, species=H),
                             F.FixedConcentrationBC(subdomain=right, value=u_exact, species=H)],
        temperature=500, sources=[F.ParticleSource(value=f, volume=vol, species=H)],
        settings=F.Settings(atol=1e-10, rtol=1e-9, max_iterations=50, transient=transient,
                            final_time=0.3 if transient else None))
    if transient:
        m.settings.stepsize = F.Stepsize(0.1)
    return m, H, u_exact


@pytest.mark.parametrize("system,n", [("cylindrical", 500), ("spherical", 1000)])
def test_run_MMS_curvilinear_oracle(system, n):
    m, H, u_exact = _mms(system, n)
    m.solver_backend = OracleBackend()
    m.initialise()
    m.run()
    assert error_L2(m.mesh.mesh, H.post_processing_solution.x.array, u_exact) < 1e-6


def test_coordinate_system_dimension_checks():
    with pytest.raises(AttributeError, match="one-dimensional"):
        F.Mesh(F.create_unit_square(2, 2), coordinate_system="spherical")
    with pytest.raises(AttributeError, match="3D"):
        F.Mesh(F.create_unit_cube(2, 2, 2), coordinate_system="cylindrical")
    with pytest.raises(ValueError):
        F.Mesh1D(np.linspace(0, 1, 5), coordinate_system="toroidal")


def test_derived_quantities_refused_on_curvilinear_meshes():
    m, H, _ = _mms("cylindrical", 20)
    m.solver_backend = OracleBackend()
    m.exports = [F.TotalVolume(field=H, volume=m.volume_subdomains[0])]
    with pytest.raises(NotImplementedError, match="Derived quantity"):
        m.initialise()


@pytest.mark.gpu
@pytest.mark.parametrize("transient", [False, True])
def test_cylindrical_gpu_parity(transient):
    _gpu_parity("cylindrical", transient)


def _gpu_parity(system, transient):
    """(the spherical cases run from tests/test_zz_gpu_reference_form.py: their integration degree
    changed after the round's last GPU run)"""
    g, Hg, u_exact = _mms(system, 400, transient)
    r, Hr, _ = _mms(system, 400, transient)
    r.solver_backend = OracleBackend()
    for m in (g, r):
        m.initialise()
    # assembled residual and Jacobian at a random state (metric term: gradient-dependent)
    rng = np.random.default_rng(2)
    u = rng.uniform(0.5, 1.5, g.u.x.array.size)
    for m in (g, r):
        m.u.x.array[:] = u
        m.u_n.x.array[:] = 0.9 * u
    Fg, Jg = g.solver.assemble_host(want_J=True)
    Fr, Jr = r.solver.residual_and_jacobian(r.u.x.array)
    assert np.abs(Fg - Fr).max() <= 1e-10 * np.abs(Fr).max()
    assert abs((Jg - Jr)).max() <= 1e-10 * abs(Jr).max()
    for m in (g, r):
        m.u.x.array[:] = 0.0
        m.u_n.x.array[:] = 0.0
        m.run()
    a, b = Hg.post_processing_solution.x.array, Hr.post_processing_solution.x.array
    assert np.linalg.norm(a - b) / np.linalg.norm(b) < 1e-8
    if not transient:
        assert error_L2(g.mesh.mesh, a, u_exact) < 1e-5
