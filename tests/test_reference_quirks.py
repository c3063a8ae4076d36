"""reference quirks a drop-in must replicate, one test per quirk that is not
already pinned elsewhere (Q1 tests/test_problem_bookkeeping.py + time_loop.json, Q4 and Q6 golden
vectors, Q7 residual_form.json, Q8 test_non_convergence_raises, Q9 tests/test_stepsize.py)."""
import types
import warnings

import numpy as np
import pytest

import festim_b200 as F
from festim_b200 import fem, ufl
from oracle.newton import OracleBackend


def _model(**kw):
    m = F.HydrogenTransportProblem()
    m.mesh = F.Mesh1D(np.linspace(0, 1, 11))
    H = F.Species("H")
    m.species = [H]
    vol = F.VolumeSubdomain1D(id=1, borders=[0, 1], material=F.Material(D_0=1.0, E_D=0.1))
    left, right = F.SurfaceSubdomain1D(id=1, x=0), F.SurfaceSubdomain1D(id=2, x=1)
    m.subdomains = [vol, left, right]
    m.temperature = 500.0
    m.settings = F.Settings(atol=1e-10, rtol=1e-10, final_time=0.2, stepsize=0.1)
    m.solver_backend = OracleBackend()
    for k, v in kw.items():
        setattr(m, k, v)
    return m, H, vol, left, right


def test_q2_function_temperature_is_not_time_dependent():
    """hydrogen_transport_problem.py:237-247: a fem.Function temperature (the coupled case) reports
    temperature_time_dependent == False, so D_global is never re-interpolated (:1049-1056)."""
    m, H, vol, left, right = _model()
    T = fem.Function(fem.functionspace(m.mesh.mesh, ("P", 1)))
    T.x.array[:] = 400.0
    m.temperature = T
    flux = F.SurfaceFlux(field=H, surface=right)
    m.exports = [flux]
    m.boundary_conditions = [F.FixedConcentrationBC(left, value=1.0, species=H),
                             F.FixedConcentrationBC(right, value=0.0, species=H)]
    m.initialise()
    assert m.temperature_time_dependent is False
    m.iterate()
    first = flux.data[-1]
    grad_1 = (m.u.x.array[-1] - m.u.x.array[-2]) / 0.1
    assert np.isclose(first, -1.0 * np.exp(-0.1 / F.k_B / 400.0) * grad_1, rtol=1e-10)
    T.x.array[:] = 800.0          # "the heat problem" moved on; D_global did not
    m.iterate()
    grad_2 = (m.u.x.array[-1] - m.u.x.array[-2]) / 0.1
    assert np.isclose(flux.data[-1], -1.0 * np.exp(-0.1 / F.k_B / 400.0) * grad_2, rtol=1e-10)


def test_q3_callables_of_t_only_get_a_python_float():
    """dirichlet_bc.py:385-419: value(t) is called with a float and must return a number (a
    ufl.conditional there is an error); value(x, t) receives the symbolic time."""
    m, H, vol, left, right = _model()
    seen = []

    def ramp(t):
        seen.append(t)
        return 1.0 + 2.0 * t

    m.boundary_conditions = [F.FixedConcentrationBC(left, value=ramp, species=H),
                             F.FixedConcentrationBC(right, value=lambda x, t: 0.0 * x[0] + t, species=H)]
    m.initialise()
    m.run()
    assert seen and all(isinstance(t, float) for t in seen)
    assert np.isclose(m.u.x.array[0], 1.0 + 2.0 * float(m.t)) and np.isclose(m.u.x.array[-1], float(m.t))

    bad, Hb, _, lb, _ = _model()
    bad.boundary_conditions = [F.FixedConcentrationBC(lb, value=lambda t: ufl.conditional(ufl.lt(t, 1.0), 1.0, 2.0),
                                                      species=Hb)]
    with pytest.raises(ValueError, match="should return a float or an int"):
        bad.initialise()


def test_q5_callable_tolerances_are_attribute_writes_on_the_solver():
    """problem.py:226-230: rtol(t) / atol(t) are evaluated at the time BEFORE the step and written
    onto the solver object."""
    m, H, vol, left, right = _model()
    m.settings = F.Settings(atol=lambda t: 1e-9 + t, rtol=lambda t: 1e-10 * (1 + t), final_time=0.3, stepsize=0.1)
    snes = types.SimpleNamespace(getConvergedReason=lambda: 2, getIterationNumber=lambda: 1,
                                 getLinearSolveIterations=lambda: 0)
    solver = types.SimpleNamespace(solve=lambda: 0, solver=snes, rtol=None, atol=None)
    m.solver_backend = types.SimpleNamespace(create_newton_solver=lambda problem, opts: solver)
    m.initialise()
    m.iterate()
    assert solver.atol == 1e-9 + 0.0 and solver.rtol == 1e-10
    m.iterate()
    assert solver.atol == 1e-9 + 0.1 and solver.rtol == 1e-10 * 1.1


def test_q10_surface_flux_ignores_advection_with_a_warning():
    """hydrogen_transport_problem.py:1098-1102."""
    m, H, vol, left, right = _model()
    vel = F.VectorFunction(m.mesh.mesh, np.full((11, 1), 0.3))
    m.advection_terms = [F.AdvectionTerm(velocity=vel, subdomain=vol, species=H)]
    m.boundary_conditions = [F.FixedConcentrationBC(left, value=1.0, species=H),
                             F.FixedConcentrationBC(right, value=0.0, species=H)]
    flux = F.SurfaceFlux(field=H, surface=right)
    m.exports = [flux]
    m.initialise()
    with pytest.warns(UserWarning, match="Advection terms are not currently accounted for"):
        m.iterate()
    grad = (m.u.x.array[-1] - m.u.x.array[-2]) / 0.1
    assert np.isclose(flux.data[-1], -np.exp(-0.1 / F.k_B / 500.0) * grad, rtol=1e-10)   # diffusive part only


def test_q11_initialise_twice_duplicates_trap_species():
    """hydrogen_transport_problem.py:356-362: create_species_from_traps appends."""
    m, H, vol, left, right = _model()
    m.traps = [F.Trap(name="trap", mobile_species=H, k_0=1.0, E_k=0.1, p_0=0.1, E_p=0.2, n=2.0, volume=vol)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m.initialise()
        assert len(m.species) == 2 and len(m.reactions) == 1
        m.initialise()
    assert len(m.species) == 3 and len(m.reactions) == 2


REFERENCE_PUBLIC_NAMES = """
AdvectionTerm AverageSurface AverageVolume CoupledTransientHeatTransferHydrogenTransport
CustomFieldExport CustomQuantity DerivedQuantity DirichletBC DirichletBCBase Enclosure
EnclosureConnection ExportBaseClass FixedConcentrationBC FixedTemperatureBC FluxBCBase GasPressure
GasSpecies HeatFluxBC HeatSource HeatTransferProblem HenrysBC HydrogenTransportProblem
HydrogenTransportProblemDiscontinuous HydrogenTransportProblemDiscontinuousChangeVar ImplicitSpecies
InitialConcentration InitialConditionBase InitialTemperature Interface InterfaceMethod KSPMonitor
Material MaximumSurface MaximumVolume Mesh Mesh1D MeshFromXDMF MinimumSurface MinimumVolume
OpeningBase ParticleFluxBC ParticleSource PrescribedFlowRate ProblemBase Profile1DExport Pump
Reaction ReactionRateExport Reservoir Settings SievertsBC SnesMonitor SolubilityLaw SourceBase
Species Stepsize SurfaceFlux SurfaceQuantity SurfaceReactionBC SurfaceSubdomain SurfaceSubdomain1D
TotalSurface TotalVolume Trap VTXSpeciesExport VTXTemperatureExport Value VelocityField
VolumeQuantity VolumeSubdomain VolumeSubdomain1D XDMFExport as_fenics_constant
as_fenics_interp_expr_and_function as_mapped_function convergenceTest find_species_from_name
find_surface_from_id find_volume_from_id map_surface_to_volume_subdomains plot
read_function_from_file
""".split()


def test_every_public_name_of_the_reference_resolves():
    """reference src/festim/__init__.py:1-104 (names listed here because the GPU box has no
    /root/reference): each one is either implemented or an explicit out-of-scope class."""
    import festim_b200

    out_of_scope = set(festim_b200._OUT_OF_SCOPE)
    for name in REFERENCE_PUBLIC_NAMES:
        obj = getattr(festim_b200, name)
        if name in out_of_scope:
            with pytest.raises(NotImplementedError):
                obj()
    implemented = [n for n in REFERENCE_PUBLIC_NAMES if n not in out_of_scope]
    assert len(implemented) >= 69
