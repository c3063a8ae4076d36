"""Time-loop bookkeeping of ProblemBase (reference src/festim/problem.py:195-255),
exercised with the oracle backend at the create_solver seam."""
import numpy as np
import pytest

import festim_b200 as F
from oracle.newton import OracleBackend


def _model(final_time=10, stepsize=1, source=None):
    m = F.HydrogenTransportProblem()
    m.mesh = F.Mesh1D(vertices=np.linspace(0, 1, num=50))
    H = F.Species("H")
    m.species = [H]
    vol = F.VolumeSubdomain1D(id=1, borders=[0, 1], material=F.Material(D_0=1, E_D=0))
    m.subdomains = [vol]
    m.temperature = 500
    if source is not None:
        m.sources = [F.ParticleSource(value=source, volume=vol, species=H)]
    m.settings = F.Settings(transient=True, atol=1e-10, rtol=1e-10, final_time=final_time, stepsize=stepsize)
    m.solver_backend = OracleBackend()
    return m, H


def test_timesteps():
    """reference test/system_tests/test_misc.py:336-358."""
    m, _ = _model()
    m.initialise()
    m.run()
    assert np.allclose(m.timesteps, np.linspace(0, 10, num=10, endpoint=False))


def test_iterate_updates_time_and_solution():
    """reference test/test_h_transport_problem.py:139-186: du/dt = 2 with no boundary
    conditions, so u = 2 t everywhere after every step."""
    m, _ = _model(final_time=20, stepsize=2.0, source=2.0)
    m.initialise()
    for i in range(10):
        m.iterate()
        assert np.isclose(float(m.t), (i + 1) * 2.0)
        assert np.allclose(m.u.x.array, (i + 1) * 2.0 * 2.0)
        assert np.array_equal(m.u_n.x.array, m.u.x.array)


def test_last_step_is_not_clamped():
    """t may overshoot final_time (problem.py:206-207, 232)."""
    m, _ = _model(final_time=1.0, stepsize=0.3)
    m.initialise()
    m.run()
    assert np.isclose(float(m.t), 1.2)


def test_non_convergence_raises():
    """problem.py:238-241: a diverged Newton solve aborts with the reference's message."""
    m, H = _model(final_time=1.0, stepsize=1.0, source=lambda H: H**2 + 1e6)
    m.sources[0].value.species_dependent_value = {"H": H}
    m.settings.max_iterations = 1
    m.initialise()
    with pytest.raises(AssertionError, match="Non-linear solver did not converge"):
        m.run()


def test_volume_ids_must_be_unique_and_cells_tagged():
    m = F.HydrogenTransportProblem()
    m.mesh = F.Mesh1D(vertices=np.linspace(0, 1, num=11))
    mat = F.Material(D_0=1, E_D=0)
    m.species = [F.Species("H")]
    m.temperature = 300
    m.settings = F.Settings(atol=1, rtol=1, transient=False)
    m.solver_backend = OracleBackend()
    m.subdomains = [F.VolumeSubdomain1D(id=1, borders=[0, 0.5], material=mat),
                    F.VolumeSubdomain1D(id=1, borders=[0.5, 1], material=mat)]
    with pytest.raises(ValueError, match="Volume ids are not unique"):
        m.initialise()


def test_adaptive_stepsize_grows_with_easy_newton():
    """reference test/test_h_transport_problem.py:1147-1216 (dt grows when Newton needs
    fewer iterations than the target)."""
    m, _ = _model(final_time=10, stepsize=1, source=1.0)
    m.settings.stepsize = F.Stepsize(initial_value=0.5, growth_factor=1.5, cutback_factor=0.5,
                                     target_nb_iterations=4)
    m.initialise()
    dts = []
    for _ in range(4):
        m.iterate()
        dts.append(float(m.dt))
    assert np.allclose(dts, 0.5 * 1.5 ** np.arange(1, 5))
