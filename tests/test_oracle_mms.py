"""The CPU oracle against the reference's own known-answer tests.
Thresholds are the reference's; each test cites the reference test it restates."""
import numpy as np
import pytest

from cases import error_L2, mms_1_mobile_1_trap_steady, mms_1_mobile_steady
from oracle.newton import OracleBackend


def _run(model):
    model.solver_backend = OracleBackend()
    model.initialise()
    model.run()
    return model


def test_1_mobile_MMS_steady_state_1d():
    """reference test/system_tests/test_1_mobile_species.py:20-80 (L2 < 1e-7)."""
    m, H, exact = mms_1_mobile_steady(1, 9999)
    _run(m)
    assert m.solver.solver.getConvergedReason() > 0
    assert error_L2(m.mesh.mesh, H.post_processing_solution.x.array, exact) < 1e-7


def test_1_mobile_MMS_3D():
    """reference test/system_tests/test_1_mobile_species.py:216-282 (L2 < 1e-2, n=20)."""
    m, H, exact = mms_1_mobile_steady(3, 20)
    _run(m)
    assert m.solver.solver.getIterationNumber() == 1
    assert error_L2(m.mesh.mesh, H.post_processing_solution.x.array, exact) < 1e-2


def test_1_mobile_MMS_2D():
    """2D analogue on the 50 x 50 square of test/system_tests/tools.py:11."""
    m, H, exact = mms_1_mobile_steady(2, 50)
    _run(m)
    assert error_L2(m.mesh.mesh, H.post_processing_solution.x.array, exact) < 5e-3


def test_1_mobile_1_trap_MMS_steady_state_1d():
    """reference test/system_tests/test_1_mobile_1_trap.py:20-113 (2e-7 / 1e-7)."""
    m, (mobile, trapped), (ue, ve) = mms_1_mobile_1_trap_steady(1, 9999)
    _run(m)
    assert m.solver.solver.getConvergedReason() > 0
    mesh = m.mesh.mesh
    assert error_L2(mesh, mobile.post_processing_solution.x.array, ue) < 2e-7
    assert error_L2(mesh, trapped.post_processing_solution.x.array, ve) < 1e-7


def test_1_mobile_MMS_transient():
    """reference test/system_tests/test_1_mobile_species.py:83-145 (L2 < 5e-4);
    a 2000-cell mesh keeps the CPU suite short, the error is time-discretisation
    dominated (50 implicit-Euler steps) and unchanged on the 9999-cell mesh."""
    from cases import mms_1_mobile_transient

    m, H, exact = mms_1_mobile_transient(n=2000)
    _run(m)
    assert len(m.timesteps) == 50
    assert error_L2(m.mesh.mesh, H.post_processing_solution.x.array, exact) < 5e-4


def test_permeation_problem():
    """reference test/test_permeation_problem.py:35-113 (mean relative error < 1 %)."""
    from cases import permeation_1d, permeation_relative_error

    import festim_b200 as F

    m, H, flux, mat, L = permeation_1d()
    _run(m)
    D = 1.9e-7 * np.exp(-0.2 / F.k_B / 500)
    S = 4.02e21 * np.exp(-1.04 / F.k_B / 500)
    err = permeation_relative_error(D, D * S, np.abs(np.array(flux.data)), L, np.array(flux.t), 100)
    assert err < 0.01


def test_tds_adaptive_stepping_runs():
    """C0-style TDS: adaptive stepping hits its milestones and the trap empties
    during the ramp (qualitative pin of the driver, reference problem.py:195-255)."""
    from cases import tds_1d

    m, H, trap, (fl, fr, inv) = tds_1d()
    _run(m)
    t = np.array(m.timesteps)
    assert np.isclose(t, 20.0).any() and np.isclose(t, 30.0).any()
    assert float(m.t) >= 60.0
    trapped = trap.trapped_concentration.post_processing_solution.x.array
    assert trapped.max() < 1e-3 * 8.19e22


@pytest.mark.parametrize("T_dependent", [False, True])
def test_heat_transfer_MMS(T_dependent):
    """reference test/test_heat_transfer_problem.py:67-110 and :113-165 (L2 < 1e-7)."""
    from cases import heat_mms

    p, exact = heat_mms(T_dependent)
    _run(p)
    assert p.solver.solver.getConvergedReason() > 0
    assert error_L2(p.mesh.mesh, p.u.x.array, exact) < 1e-7


@pytest.mark.parametrize("heat", [False, True])
def test_flux_bc_MMS_steady_state(heat):
    """reference test/system_tests/test_fluxes.py:13-57, :60-95 (L2 < 1e-7)."""
    from cases import flux_bc_mms

    m, field, exact = flux_bc_mms(heat)
    _run(m)
    arr = m.u.x.array if heat else field.post_processing_solution.x.array
    assert error_L2(m.mesh.mesh, arr, exact) < 1e-7


def test_MMS_coupled_problem():
    """reference test/system_tests/test_coupled_heat_hydrogen.py:9-207 (< 2e-7 x 3)."""
    from cases import coupled_heat_hydrogen_mms

    c, (mobile, trapped), exact = coupled_heat_hydrogen_mms()
    c.heat_problem.solver_backend = OracleBackend()
    c.hydrogen_problem.solver_backend = OracleBackend()
    c.initialise()
    c.run()
    mesh = c.heat_problem.mesh.mesh
    assert c.hydrogen_problem.temperature_fenics is c.heat_problem.u
    assert error_L2(mesh, c.heat_problem.u.x.array, exact[0]) < 2e-7
    assert error_L2(mesh, mobile.post_processing_solution.x.array, exact[1]) < 2e-7
    assert error_L2(mesh, trapped.post_processing_solution.x.array, exact[2]) < 2e-7


def test_MMS_1_species_1_trap_with_advection():
    """reference test/system_tests/test_advection_term_integration.py:15-142 on a
    100 x 100 square (the reference uses 200 x 200 with thresholds 2e-3 / 7e-5; the
    error is O(h^2), hence 4x the thresholds here to keep the CPU suite short)."""
    from cases import advection_trap_mms_2d

    m, (mobile, trapped), (em, et) = advection_trap_mms_2d(100)
    _run(m)
    mesh = m.mesh.mesh
    assert error_L2(mesh, mobile.post_processing_solution.x.array, em) < 4 * 2e-3
    assert error_L2(mesh, trapped.post_processing_solution.x.array, et) < 4 * 7e-5


@pytest.mark.parametrize("k, p, c_A_0", [(350e-4, 120e-4, 3), (200e-4, 100e-4, 2)])
def test_reaction(k, p, c_A_0):
    """reference test/test_complex_reaction.py:157-178 (relative L2 error < 1e-2)."""
    from cases import complex_reaction, concentration_A_exact

    m = complex_reaction(0.5, k, p, c_A_0)
    _run(m)
    total_A = m.exports[0]
    times, data = np.array(total_A.t), np.array(total_A.data)
    exact = concentration_A_exact(times, c_A_0, k, p)
    err = np.sqrt(np.sum((exact - data) ** 2)) / np.sqrt(np.sum(exact**2))
    assert err < 1e-2


@pytest.mark.parametrize("name", ["step_space", "simple_space", "step_time_space", "step_time"])
def test_non_homogeneous_density(name):
    """reference test/system_tests/test_non_homogeneous_density.py:16-85."""
    from cases import DENSITIES, non_homogeneous_density

    func, expected = DENSITIES[name]
    m, total = non_homogeneous_density(func)
    _run(m)
    assert np.isclose(total.value, expected)


@pytest.mark.parametrize("dim,n,tol_mobile,tol_trapped", [(2, 50, 4e-3, 2e-3), (3, 20, 3e-2, 9e-3)])
def test_1_mobile_1_trap_MMS_2D_3D(dim, n, tol_mobile, tol_trapped):
    """reference test/system_tests/test_1_mobile_1_trap.py:309-421 (2D, 50 x 50: 4e-3 / 2e-3) and
    :193-306 (3D, 20^3: 3e-2 / 9e-3), with the reference's manufactured solutions."""
    m, (mobile, trapped), (ue, ve) = mms_1_mobile_1_trap_steady(dim, n, reference_solution=True)
    m.settings.atol = m.settings.rtol = 1e-10
    _run(m)
    assert m.solver.solver.getConvergedReason() > 0
    mesh = m.mesh.mesh
    assert error_L2(mesh, mobile.post_processing_solution.x.array, ue) < tol_mobile
    assert error_L2(mesh, trapped.post_processing_solution.x.array, ve) < tol_trapped


def test_species_dependent_source_MMS_steady_state():
    """reference test/system_tests/test_1_mobile_species.py:285-347 (L2 < 1e-3)."""
    from cases import species_dependent_source_mms

    m, (A, B), (A_exact, B_exact) = species_dependent_source_mms()
    _run(m)
    mesh = m.mesh.mesh
    assert error_L2(mesh, A.post_processing_solution.x.array, A_exact) < 1e-3
    assert np.allclose(B.post_processing_solution.x.array, 1 + mesh.coords[:, 0], atol=1e-9)


def test_permeation_problem_multi_volume():
    """reference test/test_permeation_problem.py:116-204: the same permeation problem on four
    volume subdomains of one material (mean relative error < 1 %)."""
    from cases import permeation_1d, permeation_relative_error

    import festim_b200 as F

    m, H, flux, mat, L = permeation_1d(dt=0.4, n_volumes=4)
    _run(m)
    assert sorted(np.unique(m.volume_meshtags.values)) == [1, 2, 3, 4]
    D = 1.9e-7 * np.exp(-0.2 / F.k_B / 500)
    S = 4.02e21 * np.exp(-1.04 / F.k_B / 500)
    err = permeation_relative_error(D, D * S, np.abs(np.array(flux.data)), L, np.array(flux.t), 100)
    assert err < 0.01


def test_coupled_problem_non_matching_mesh():
    """reference test/system_tests/test_coupled_heat_hydrogen.py:210-332 (L2 < 2e-7): heat on 499
    cells, hydrogen on 1999, temperature transferred by point location + P1 interpolation."""
    from cases import coupled_non_matching_mms

    coupled, mobile, exact = coupled_non_matching_mms()
    coupled.heat_problem.solver_backend = OracleBackend()
    coupled.hydrogen_problem.solver_backend = OracleBackend()
    coupled.initialise()
    assert coupled.non_matching_meshes
    coupled.run()
    mesh = coupled.hydrogen_problem.mesh.mesh
    assert error_L2(mesh, mobile.post_processing_solution.x.array, exact) < 2e-7


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_interpolate_nonmatching_reproduces_linear_fields(dim):
    """P1 interpolation between unrelated meshes is exact for affine fields, including at target
    vertices that sit on source facets, vertices and the boundary."""
    import festim_b200 as F
    from festim_b200 import fem

    if dim == 1:
        src, dst = F.Mesh1D(np.linspace(0, 1, 38)).mesh, F.Mesh1D(np.linspace(0, 1, 101)).mesh
    elif dim == 2:
        src, dst = F.create_unit_square(7, 5), F.create_unit_square(14, 11)
    else:
        src, dst = F.create_unit_cube(4, 3, 5), F.create_unit_cube(8, 7, 5)
    coef = np.array([1.5, -2.0, 0.75])[:dim]
    f_in, f_out = fem.Function(fem.functionspace(src, ("P", 1))), fem.Function(fem.functionspace(dst, ("P", 1)))
    f_in.x.array[:] = 3.0 + src.coords @ coef
    fem.interpolate_nonmatching(f_out, f_in)
    assert np.allclose(f_out.x.array, 3.0 + dst.coords @ coef, rtol=0, atol=1e-12)
    # a non-affine field: interpolation error of the coarse P1 interpolant only
    f_in.x.array[:] = np.sin(src.coords[:, 0])
    fem.interpolate_nonmatching(f_out, f_in)
    assert np.abs(f_out.x.array - np.sin(dst.coords[:, 0])).max() < 0.02
    # target points outside the source mesh are an error
    far = fem.Function(fem.functionspace(F.Mesh1D(np.linspace(0, 2, 5)).mesh if dim == 1 else
                                         F.create_rectangle(((0, 0), (2, 1)), (3, 3)) if dim == 2 else
                                         F.create_box(((0, 0, 0), (2, 1, 1)), (2, 2, 2)), ("P", 1)))
    with pytest.raises(ValueError, match="outside the source mesh"):
        fem.interpolate_nonmatching(far, f_in)
