"""Golden vectors produced by the REAL reference code (tools/make_golden.py imports
/root/reference/src/festim with its numerical dependencies stubbed): they pin the
host logic of festim_b200, the oracle's Newton stopping rule and the C-ABI step
size rule to the reference, bit for bit."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import festim_b200 as F
from festim_b200 import ufl
from oracle.newton import convergence_test

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return json.load(open(os.path.join(GOLD, name)))


def test_constants():
    c = load("constants.json")
    assert (F.k_B, F.R, F.k_B_SI) == (c["k_B"], c["R"], c["k_B_SI"])


def test_stepsize_modify_value_matches_reference():
    """reference src/festim/stepsize.py:144-167."""
    for case in load("stepsize_modify_value.json"):
        st = F.Stepsize(**case["config"])
        assert st.modify_value(case["value"], case["nb_iterations"], case["t"]) == case["expected"]


def test_c_abi_stepsize_matches_reference(b200_lib):
    for case in load("stepsize_modify_value.json"):
        cfg = case["config"]
        ms = sorted(cfg.get("milestones", []))
        arr = (C.c_double * max(len(ms), 1))(*ms)
        got = b200_lib.fb_stepsize_modify(
            case["value"], case["nb_iterations"], case["t"], 1, cfg["growth_factor"],
            cfg["cutback_factor"], cfg["target_nb_iterations"], cfg.get("max_stepsize", 0.0) or 0.0,
            len(ms), arr, cfg.get("milestone_tolerance", 1e-5))
        assert got == case["expected"]


def test_oracle_convergence_test_matches_reference():
    """reference src/festim/helpers.py:408-426."""
    for seq in load("convergence_test.json"):
        f0 = seq["rows"][0]["fnorm"]
        for row in seq["rows"]:
            got = convergence_test(row["it"], row["snorm"], row["fnorm"], f0, seq["rtol"],
                                   seq["atol"], seq["stol"], seq["max_it"])
            assert got == row["reason"], (seq, row)


def test_reaction_term_matches_reference():
    """reference src/festim/reaction.py:117-224 (all three detrapping-rate modes)."""
    mat = F.Material(D_0=1.0, E_D=0.1)
    vol = F.VolumeSubdomain(id=1, material=mat)
    for case in load("reaction_term.json"):
        A, B = F.Species("A"), F.Species("B", mobile=False)
        A.solution, B.solution = ufl.Const(case["c_mobile"]), ufl.Const(case["c_trapped"])
        imp = F.ImplicitSpecies(n=case["n"], others=[B])
        imp.value_fenics = ufl.Const(case["n"])
        r = F.Reaction(reactant=[A, imp], product=B, k_0=case["k_0"], E_k=case["E_k"],
                       p_0=case["p_0"], E_p=case["E_p"], volume=vol)
        got = float(ufl.evaluate(r.reaction_term(case["T"]), {}))
        assert got == pytest.approx(case["expected"], rel=1e-14, abs=0.0)


def test_diffusion_coefficient_matches_reference():
    """reference src/festim/material.py:262-318."""
    for case in load("diffusion_coefficient.json"):
        m = F.Material(D_0=case["D_0"], E_D=case["E_D"])
        got = float(ufl.evaluate(m.get_diffusion_coefficient(None, case["T"], None), {}))
        assert got == pytest.approx(case["expected"], rel=1e-15)


def test_surface_reaction_flux_matches_reference():
    """reference src/festim/boundary_conditions/surface_reaction.py:57-85 and :129-146: the
    flux K_d P - K_r c_A c_B of every reactant, A + A listed twice."""
    surf = F.SurfaceSubdomain1D(id=2, x=1.0)
    for case in load("surface_reaction_flux.json"):
        A, B = F.Species("A"), F.Species("B")
        A.solution, B.solution = ufl.Const(case["cA"]), ufl.Const(case["cB"])
        reactants = [A, A] if case["same"] else [A, B]
        bc = F.SurfaceReactionBC(reactant=reactants, gas_pressure=case["P"], k_r0=case["k_r0"],
                                 E_kr=case["E_kr"], k_d0=case["k_d0"], E_kd=case["E_kd"], subdomain=surf)
        assert len(bc.flux_bcs) == case["n_fluxes"]
        assert [fb.species.name for fb in bc.flux_bcs] == case["flux_species"]
        for fb, expected in zip(bc.flux_bcs, case["expected"]):
            fb.create_value_fenics(None, ufl.Const(case["T"]), None)
            got = float(ufl.evaluate(fb.value_fenics, {}))
            assert got == pytest.approx(expected, rel=1e-13, abs=0.0)


def test_nitsche_terms_match_reference():
    """reference src/festim/boundary_conditions/dirichlet_bc.py:170-244: numerical flux
    -D n.grad(u) + penalty D / h (u - g) and the full symmetric weak form at a point, for a
    given u, v and their gradients."""
    surf = F.SurfaceSubdomain1D(id=2, x=1.0)
    for case in load("nitsche_terms.json"):
        H = F.Species("H")
        bc = F.FixedConcentrationBC(subdomain=surf, value=case["g"], species=H, enforce_weakly=True,
                                    penalty=case["penalty"])
        bc.value_fenics = ufl.Const(case["g"])
        mesh = type("M", (), {"gdim": case["gdim"]})()
        u = ufl.SpeciesRef(H)
        n, gu, gv = np.array(case["n"]), np.array(case["grad_u"]), np.array(case["grad_v"])
        env = {"species": lambda s: case["u"], "species_grad": lambda s, i: gu[i],
               "normal": lambda i: n[i], "circumradius": lambda: case["h"]}
        flux = float(ufl.evaluate(bc.numerical_flux(u, case["D"], mesh, H), env))
        assert flux == pytest.approx(case["numerical_flux"], rel=1e-12)
        terms = bc.weak_formulation(H, u, lambda sid: ("ds", sid), case["D"], mesh)
        assert [t.kind for t in terms] == ["test", "ngradtest"]
        # integrand at the point: (term tested with v) * v + (term tested with n.grad v) * n.grad(v)
        total = float(ufl.evaluate(terms[0].expr, env)) * case["v"] + \
            float(ufl.evaluate(terms[1].expr, env)) * float(n @ gv)
        assert total == pytest.approx(case["weak_form"], rel=1e-11)


def test_solubility_laws_match_reference():
    """reference sieverts_bc.py:8-11 (S_0 exp(-E_S/k_B/T) sqrt(P)) and henrys_bc.py:8-11."""
    from festim_b200.boundary_conditions import henrys_law, sieverts_law

    for case in load("solubility_laws.json"):
        args = (ufl.Const(case["T"]), case["c0"], case["E"], case["P"])
        assert float(ufl.evaluate(sieverts_law(*args), {})) == pytest.approx(case["sieverts"], rel=1e-14)
        assert float(ufl.evaluate(henrys_law(*args), {})) == pytest.approx(case["henry"], rel=1e-14)


def _residual_form_problem(gold, transient, first_reaction_only=False):
    """The problem of tests/golden/residual_form.json in the API mirror."""
    par = gold["parameters"]
    m = F.HydrogenTransportProblem()
    m.mesh = F.Mesh1D(np.array(par["vertices"]))
    species = {"H": F.Species("H"), "D": F.Species("D"), "t1": F.Species("t1", mobile=False),
               "t2": F.Species("t2", mobile=False)}
    m.species = [species[n] for n in gold["species"]]
    vols = {i: F.VolumeSubdomain1D(id=i, borders=b, material=F.Material(D_0=par["D_0"][str(i)], E_D=par["E_D"][str(i)]))
            for i, b in ((1, [0.0, 0.5]), (2, [0.5, 1.0]))}
    surfs = {f["surface"]: F.SurfaceSubdomain1D(id=f["surface"], x=f["x"]) for f in par["fluxes"]}
    m.subdomains = [*vols.values(), *surfs.values()]
    m.reactions = [
        F.Reaction(reactant=[species[r["mobile"]], F.ImplicitSpecies(n=r["n"], others=[species[r["trapped"]]])],
                   product=species[r["trapped"]], k_0=r["k_0"], E_k=r["E_k"], p_0=r["p_0"], E_p=r["E_p"],
                   volume=vols[r["volume"]]) for r in par["reactions"][: 1 if first_reaction_only else None]]
    m.sources = [F.ParticleSource(value=s["value"], volume=vols[s["volume"]], species=species[s["species"]])
                 for s in par["sources"]]
    m.boundary_conditions = [F.ParticleFluxBC(subdomain=surfs[f["surface"]], value=f["value"], species=species[f["species"]])
                             for f in par["fluxes"]]
    m.temperature = par["T"]
    if transient:
        m.settings = F.Settings(atol=1e-10, rtol=1e-10, transient=True, final_time=1.0, stepsize=par["dt"])
    else:
        m.settings = F.Settings(atol=1e-10, rtol=1e-10, transient=False)
    return m


@pytest.mark.parametrize("label", ["transient", "steady"])
def test_oracle_residual_and_jacobian_match_the_reference_form(label):
    """reference src/festim/hydrogen_transport_problem.py:870-997, run for real (tools/make_golden.py):
    the oracle's assembled residual equals the reference's form integrated cell by cell, and its
    Jacobian (symbolic derivative) reproduces the exact directional difference of that residual."""
    from oracle.newton import OracleBackend

    gold = load("residual_form.json")
    m = _residual_form_problem(gold, label == "transient")
    m.solver_backend = OracleBackend()
    m.initialise()
    u, u_n, d = (np.array(gold[k]).ravel() for k in ("u", "u_n", "delta"))
    m.u.x.array[:] = u
    m.u_n.x.array[:] = u_n
    Fo, Jo = m.solver.residual_and_jacobian(m.u.x.array)
    Fg, Jd = np.array(gold[f"F_{label}"]).ravel(), np.array(gold[f"Jdelta_{label}"]).ravel()
    assert np.abs(Fo - Fg).max() <= 1e-13 * np.abs(Fg).max()
    assert np.abs(Jo @ d - Jd).max() <= 1e-12 * np.abs(Jd).max()


def test_oracle_advection_term_matches_the_reference_form():
    """reference hydrogen_transport_problem.py:967-978 on top of the form above: (grad(c) . vel) v dx
    for H in volume 1 with a P1 velocity (1D, so the velocity is a one-component vector field)."""
    from oracle.newton import OracleBackend

    gold = load("residual_form.json")
    m = _residual_form_problem(gold, True)
    vel = F.VectorFunction(m.mesh.mesh, np.array(gold["velocity"])[:, None])
    m.advection_terms = [F.AdvectionTerm(velocity=vel, subdomain=m.volume_subdomains[0], species=m.species[0])]
    m.solver_backend = OracleBackend()
    m.initialise()
    u, u_n, d = (np.array(gold[k]).ravel() for k in ("u", "u_n", "delta"))
    m.u.x.array[:] = u
    m.u_n.x.array[:] = u_n
    Fo, Jo = m.solver.residual_and_jacobian(m.u.x.array)
    Fg, Jd = np.array(gold["F_advection"]).ravel(), np.array(gold["Jdelta_advection"]).ravel()
    assert not np.allclose(Fg, np.array(gold["F_transient"]).ravel())       # the term is there
    assert np.abs(Fo - Fg).max() <= 1e-13 * np.abs(Fg).max()
    assert np.abs(Jo @ d - Jd).max() <= 1e-12 * np.abs(Jd).max()


def _heat_form_problem(gold, transient):
    par = gold["parameters"]
    p = F.HeatTransferProblem()
    p.mesh = F.Mesh1D(np.array(par["vertices"]))
    vols = {}
    for i, b in ((1, [0.0, 0.5]), (2, [0.5, 1.0])):
        mp = par["materials"][str(i)]
        lam = (lambda T, a=mp["lambda_a"], bb=mp["lambda_b"]: a + bb * T) if mp["lambda_b"] else mp["lambda_a"]
        vols[i] = F.VolumeSubdomain1D(id=i, borders=b, material=F.Material(
            D_0=1.0, E_D=0.0, thermal_conductivity=lam, density=mp["density"], heat_capacity=mp["heat_capacity"]))
    surf = F.SurfaceSubdomain1D(id=par["flux"]["surface"], x=par["flux"]["x"])
    p.subdomains = [*vols.values(), surf]
    p.sources = [F.HeatSource(value=par["source"]["value"], volume=vols[par["source"]["volume"]])]
    p.boundary_conditions = [F.HeatFluxBC(subdomain=surf, value=par["flux"]["value"])]
    if transient:
        p.settings = F.Settings(atol=1e-10, rtol=1e-10, transient=True, final_time=1.0, stepsize=par["dt"])
    else:
        p.settings = F.Settings(atol=1e-10, rtol=1e-10, transient=False)
    return p


@pytest.mark.parametrize("label", ["transient", "steady"])
def test_oracle_heat_residual_matches_the_reference_form(label):
    """reference src/festim/heat_transfer_problem.py:175-218 run for real: conductivity a + b T in one
    material and constant in the other, rho cp dT/dt, source, heat flux."""
    from oracle.newton import OracleBackend

    gold = load("heat_form.json")
    p = _heat_form_problem(gold, label == "transient")
    p.solver_backend = OracleBackend()
    p.initialise()
    p.u.x.array[:] = np.array(gold["T"])
    p.u_n.x.array[:] = np.array(gold["T_n"])
    Fo, Jo = p.solver.residual_and_jacobian(p.u.x.array)
    Fg, Jd = np.array(gold[f"F_{label}"]), np.array(gold[f"Jdelta_{label}"])
    assert np.abs(Fo - Fg).max() <= 1e-13 * np.abs(Fg).max()
    assert np.abs(Jo @ np.array(gold["delta"]) - Jd).max() <= 1e-12 * np.abs(Jd).max()


def test_oracle_repeated_reactant_matches_the_reference_form():
    """SURVEY Appendix B, Q7 (hydrogen_transport_problem.py:913-920): H + H <-> t2 - the species
    repeated in `reactant` receives the reaction term once per occurrence."""
    from oracle.newton import OracleBackend

    gold = load("residual_form.json")
    m = _residual_form_problem(gold, True)
    r = gold["repeated_reactant_reaction"]
    by_name = {s.name: s for s in m.species}
    m.reactions.append(F.Reaction(reactant=[by_name[n] for n in r["reactants"]], product=by_name[r["product"]],
                                  k_0=r["k_0"], E_k=r["E_k"], p_0=r["p_0"], E_p=r["E_p"],
                                  volume=m.volume_subdomains[r["volume"] - 1]))
    m.solver_backend = OracleBackend()
    m.initialise()
    u, u_n, d = (np.array(gold[k]).ravel() for k in ("u", "u_n", "delta"))
    m.u.x.array[:] = u
    m.u_n.x.array[:] = u_n
    Fo, Jo = m.solver.residual_and_jacobian(m.u.x.array)
    Fg, Jd = np.array(gold["F_repeated"]).ravel(), np.array(gold["Jdelta_repeated"]).ravel()
    assert np.abs(Fo - Fg).max() <= 1e-13 * np.abs(Fg).max()
    assert np.abs(Jo @ d - Jd).max() <= 1e-12 * np.abs(Jd).max()


@pytest.mark.parametrize("system", ["cylindrical", "spherical"])
def test_oracle_curvilinear_diffusion_matches_the_reference_form(system):
    """reference hydrogen_transport_problem.py:886-908, run for real with dual numbers:
    r^m D grad(u) . grad(v / r^m) dx.  The integrand is rational in r, so the golden file tabulates
    the residual per Gauss-Legendre rule size; the oracle must hit the table entry of the rule its
    estimated degree selects (to round-off) and lie within quadrature error of the 8-point value."""
    from oracle.newton import OracleBackend

    gold = load("curvilinear_form.json")
    par = gold["parameters"]
    m = F.HydrogenTransportProblem()
    m.mesh = F.Mesh1D(np.array(par["vertices"]), coordinate_system=system)
    H = F.Species("H")
    m.species = [H]
    vol = F.VolumeSubdomain1D(id=1, borders=[par["vertices"][0], par["vertices"][-1]],
                              material=F.Material(D_0=par["D_0"], E_D=par["E_D"]))
    m.subdomains = [vol]
    m.sources = [F.ParticleSource(value=par["source"], volume=vol, species=H)]
    m.temperature = par["T"]
    m.settings = F.Settings(atol=1e-10, rtol=1e-10, transient=False)
    m.solver_backend = OracleBackend()
    m.initialise()
    m.u.x.array[:] = np.array(gold["u"])
    Fo, _ = m.solver.residual_and_jacobian(m.u.x.array)
    table = {int(k): np.array(v) for k, v in gold[system].items()}
    errs = {k: np.abs(Fo - v).max() / np.abs(v).max() for k, v in table.items()}
    best = min(errs, key=errs.get)
    assert errs[best] <= 1e-13, errs
    # m = 1 (cylindrical) / 2 (spherical): u' (v' - m v / r) is of UFL degree 0 + max(0, 1 + 1) = 2
    # (a quotient counts as the sum of degrees), times r^m: degree 2 + m -> 2-point / 3-point rules
    assert best == {"cylindrical": 2, "spherical": 3}[system], errs
    assert errs[8] < 1e-3


class _ScriptedBackend:
    """A solver backend at the create_solver seam that converges after a scripted number of Newton
    iterations and leaves the solution untouched."""

    def __init__(self, iterations):
        self.script = iter(iterations)

    def create_newton_solver(self, problem, petsc_options):
        import types

        snes = types.SimpleNamespace(getConvergedReason=lambda: 2, getIterationNumber=lambda: next(self.script),
                                     getLinearSolveIterations=lambda: 0)
        return types.SimpleNamespace(solve=lambda: 0, solver=snes, rtol=None, atol=None)


def test_time_loop_matches_the_reference():
    """reference src/festim/problem.py:195-255, run for real with a scripted solver
    (tools/make_golden.py): recorded timesteps, the step size in force during every step, the final
    time (not clamped) and the final step size agree bit for bit."""
    for case in load("time_loop.json")[:-1]:
        cfg = case["config"]
        kw = dict(cfg["stepsize"])
        if "max_stepsize_ramp" in cfg:
            t0, lo, hi = cfg["max_stepsize_ramp"]
            kw["max_stepsize"] = lambda t, t0=t0, lo=lo, hi=hi: lo if t < t0 else hi
        m = F.HydrogenTransportProblem()
        m.mesh = F.Mesh1D(np.linspace(0, 1, 4))
        m.species = [F.Species("H")]
        m.subdomains = [F.VolumeSubdomain1D(id=1, borders=[0, 1], material=F.Material(D_0=1, E_D=0))]
        m.temperature = 400
        m.settings = F.Settings(atol=1, rtol=1, final_time=cfg["final_time"])
        m.settings.stepsize = F.Stepsize(**kw)
        m.solver_backend = _ScriptedBackend(case["iterations"])
        m.initialise()
        dts = []
        post = m.post_processing
        m.post_processing = lambda: (dts.append(float(m.dt)), post())[1]
        m.run()
        assert m.timesteps == case["timesteps"], cfg
        assert dts == case["dt_during_step"], cfg
        assert float(m.t) == case["final_t"] and float(m.dt) == case["final_dt"], cfg


def test_coupled_time_loop_matches_the_reference():
    """reference src/festim/coupled_heat_hydrogen_problem.py:133-163, run for real with two
    scripted solvers: heat step, hydrogen step, then both problems take the smaller new step size."""
    case = load("time_loop.json")[-1]
    cfg = case["coupled"]
    mesh = F.Mesh1D(np.linspace(0, 1, 4))
    mat = F.Material(D_0=1, E_D=0, thermal_conductivity=1.0, density=1.0, heat_capacity=1.0)
    vol = F.VolumeSubdomain1D(id=1, borders=[0, 1], material=mat)
    heat = F.HeatTransferProblem(mesh=mesh, subdomains=[vol],
                                 initial_condition=F.InitialTemperature(value=400.0, volume=vol),
                                 settings=F.Settings(atol=1, rtol=1, transient=True,
                                                     final_time=cfg["heat"]["final_time"]))
    heat.settings.stepsize = F.Stepsize(**cfg["heat"]["stepsize"])
    hyd = F.HydrogenTransportProblem(mesh=mesh, subdomains=[vol], species=[F.Species("H")],
                                     settings=F.Settings(atol=1, rtol=1, transient=True,
                                                         final_time=cfg["hydrogen"]["final_time"]))
    hyd.settings.stepsize = F.Stepsize(**cfg["hydrogen"]["stepsize"])
    heat.solver_backend = _ScriptedBackend(case["iterations_heat"])
    hyd.solver_backend = _ScriptedBackend(case["iterations_hydrogen"])
    coupled = F.CoupledTransientHeatTransferHydrogenTransport(heat_problem=heat, hydrogen_problem=hyd)
    coupled.initialise()
    dts = {"heat": [], "hyd": []}
    for key, p in (("heat", heat), ("hyd", hyd)):
        post = p.post_processing
        p.post_processing = (lambda p=p, key=key, post=post: (dts[key].append(float(p.dt)), post())[1])
    coupled.run()
    assert hyd.timesteps == case["timesteps"] and heat.timesteps == case["timesteps_heat"]
    assert dts["hyd"] == case["dt_during_step"] and dts["heat"] == case["dt_during_step_heat"]
    assert float(hyd.t) == case["final_t"] and float(hyd.dt) == case["final_dt"]


def test_oracle_steady_filler_matches_the_reference_form():
    """reference hydrogen_transport_problem.py:980-997: in a steady run an immobile species gets
    `c v dx(vol)` in every volume that holds no reaction at all (the check is per volume, not per
    species - with a reaction in each volume there is no filler, which F_steady above covers)."""
    from oracle.newton import OracleBackend

    gold = load("residual_form.json")
    m = _residual_form_problem(gold, False, first_reaction_only=True)
    m.solver_backend = OracleBackend()
    m.initialise()
    assert sum("filler" in itg.meta for itg in m.formulation.integrals) == 2      # t1 and t2 in volume 2
    u, d = (np.array(gold[k]).ravel() for k in ("u", "delta"))
    m.u.x.array[:] = u
    Fo, Jo = m.solver.residual_and_jacobian(m.u.x.array)
    Fg, Jd = np.array(gold["F_steady_filler"]).ravel(), np.array(gold["Jdelta_steady_filler"]).ravel()
    assert not np.allclose(Fg, np.array(gold["F_steady"]).ravel())
    assert np.abs(Fo - Fg).max() <= 1e-13 * np.abs(Fg).max()
    assert np.abs(Jo @ d - Jd).max() <= 1e-12 * np.abs(Jd).max()


def test_convergenceTest_callback_matches_reference():
    """reference src/festim/helpers.py:408-426 through the PETSc-style callback kept in
    festim_b200.helpers (same golden sequences as the oracle's and the device's stopping rule)."""
    class Reasons:
        DIVERGED_MAX_IT, CONVERGED_FNORM_ABS, CONVERGED_FNORM_RELATIVE = -5, 2, 3
        CONVERGED_SNORM_RELATIVE, ITERATING = 4, 0

    class FakeSnes:
        ConvergedReason = Reasons

        def __init__(self, tol):
            self.tol = tol

        def getTolerances(self):
            return self.tol

    for seq in load("convergence_test.json"):
        snes = FakeSnes((seq["rtol"], seq["atol"], seq["stol"], seq["max_it"]))
        for row in seq["rows"]:
            assert F.convergenceTest(snes, row["it"], (1.0, row["snorm"], row["fnorm"])) == row["reason"]


def test_out_of_scope_names_say_so():
    with pytest.raises(NotImplementedError, match="SURVEY 8f1"):
        F.HydrogenTransportProblemDiscontinuous()
    with pytest.raises(NotImplementedError, match="gas enclosures"):
        F.Enclosure(volume=1.0)
    with pytest.raises(AttributeError):
        F.NoSuchThing


def test_material_property_lookups_match_reference():
    """reference src/festim/material.py:148-260: D_0 / E_D / K_S_0 / E_K_S as a number or a dict
    keyed by species object or name - same values, same exception types and messages."""
    for case in load("material_lookups.json"):
        attr = case["getter"][4:]
        A, B, C = F.Species("A"), F.Species("B"), F.Species("missing")
        species = {"A": A, "B": B, "missing": C, None: None}[case["query"]]
        value = {"number": 1.25, "by_name": {"A": 1.5, "B": 2.5}, "by_object": {A: 3.5, B: 4.5},
                 "wrong_type": "oops"}[case["spec"]]
        kwargs = dict(D_0=1.0, E_D=0.1)
        kwargs[attr] = value
        if attr == "D_0" and isinstance(value, dict):
            kwargs["E_D"] = {k: 0.1 for k in value}
        if attr == "E_D" and isinstance(value, dict):
            kwargs["D_0"] = {k: 1.0 for k in value}
        try:
            got = {"value": float(getattr(F.Material(**kwargs), case["getter"])(species=species))}
        except Exception as exc:   # noqa: BLE001
            got = {"error": type(exc).__name__, "message": str(exc)}
        expected = {k: case[k] for k in ("value", "error", "message") if k in case}
        assert got == expected, case


def test_is_it_time_to_export_matches_reference():
    """reference src/festim/helpers.py:374-401: tolerance of the match and the
    consumption of matched times, query by query."""
    for case in load("host_rules.json")["is_it_time_to_export"]:
        work = None if case["times"] is None else list(case["times"])
        for q in case["queries"]:
            assert F.is_it_time_to_export(times=work, current_time=q["t"]) == q["hit"], (case["times"], q)
            assert work == q["remaining"]


def test_mesh1d_check_borders_matches_reference():
    """reference src/festim/mesh/mesh_1d.py:109-155: accepted subdomain layouts (single and
    multi-block meshes), same errors otherwise."""
    mat = F.Material(D_0=1.0, E_D=0.0)
    for case in load("host_rules.json")["check_borders"]:
        vertices = case["vertices"]
        mesh = F.Mesh1D(vertices=[np.array(b) for b in vertices] if isinstance(vertices[0], list)
                        else np.array(vertices))
        vols = [F.VolumeSubdomain1D(id=i + 1, borders=b, material=mat) for i, b in enumerate(case["borders"])]
        if case["ok"]:
            mesh.check_borders(vols)
        else:
            with pytest.raises(ValueError) as err:
                mesh.check_borders(vols)
            assert str(err.value) == case["message"], case
