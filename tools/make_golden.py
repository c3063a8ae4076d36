"""Generate golden vectors from the REAL reference (/root/reference/src/festim).

dolfinx / ufl / petsc4py / mpi4py / basix / scifem are absent from this image, so
they are replaced by inert stub modules; only the reference's pure-Python logic
runs: Stepsize.modify_value (stepsize.py:144-167), convergenceTest
(helpers.py:408-426), Reaction.reaction_term (reaction.py:117-224, with ufl.exp
bound to numpy) and Material.get_diffusion_coefficient (material.py:262-318).
Outputs: tests/golden/*.json (committed; the GPU box has no /root/reference).

    python tools/make_golden.py
"""
import importlib.abc
import importlib.machinery
import json
import math
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUBBED = ("dolfinx", "ufl", "basix", "petsc4py", "mpi4py", "scifem", "io4dolfinx", "adios2",
           "adios4dolfinx", "pyvista", "tqdm")


class StubMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        sub = StubMeta(name, (Stub,), {})
        setattr(cls, name, sub)
        return sub

    def __call__(cls, *a, **k):
        return type.__call__(cls)

    def __or__(cls, other):
        return U((cls,)) | other

    def __ror__(cls, other):
        return U((cls,)).__ror__(other)


class U(tuple):
    """`A | B | None` on stub classes: a tuple usable by isinstance()."""

    @staticmethod
    def _items(other):
        if isinstance(other, tuple):
            return tuple(other)
        if other is None:
            return (type(None),)
        if hasattr(other, "__args__"):
            return tuple(other.__args__)
        return (other,)

    def __or__(self, other):
        return U(tuple(self) + self._items(other))

    def __ror__(self, other):
        return U(self._items(other) + tuple(self))


class Stub(metaclass=StubMeta):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return Stub()

    def __call__(self, *a, **k):
        return Stub()

    def __bool__(self):
        return False


class Constant(float):
    """fem.Constant stand-in: a float that ignores the mesh argument."""

    def __new__(cls, mesh, value):
        return float.__new__(cls, value)

    @property
    def value(self):
        return float(self)


class Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in STUBBED:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        mod = types.ModuleType(spec.name)
        mod.__path__ = []
        cache = {}

        def getattr_(name):
            if name.startswith("__"):
                raise AttributeError(name)
            if name not in cache:
                cache[name] = StubMeta(name, (Stub,), {})
            return cache[name]

        mod.__getattr__ = getattr_
        return mod

    def exec_module(self, module):
        name = module.__name__
        if name == "dolfinx":
            module.default_scalar_type = np.float64
            module.default_real_type = np.float64
            module.__version__ = "0.10.0"
            import importlib

            module.fem = importlib.import_module("dolfinx.fem")
        if name == "dolfinx.fem":
            module.Constant = Constant
        if name == "ufl":
            module.exp = np.exp
            # enough of ufl for the boundary-condition formulas pinned below: the unknown and
            # the test function are floats carrying their gradient, geometry comes from CTX
            module.FacetNormal = lambda mesh: CTX["n"]
            module.Circumradius = lambda mesh: CTX["h"]
            module.inner = lambda a, b: float(np.dot(a, b))
            module.grad = lambda u: u.grad
        if name == "petsc4py.PETSc":
            module.IntType = np.int32


CTX = {}


class WithGrad(float):
    """A number that also knows its gradient (stands in for a P1 function at a point)."""

    def __new__(cls, value, grad):
        obj = float.__new__(cls, value)
        obj.grad = np.asarray(grad, dtype=float)
        return obj


class Immutable(np.ndarray):
    """ndarray whose in-place operators return new arrays, like (immutable) UFL expressions:
    the reference writes ``product *= concentration`` on the species' own solution object."""

    def __imul__(self, other):
        return np.multiply(np.asarray(self), other).view(Immutable)


class FakeMeasure:
    def ufl_domain(self):
        return None

    def __call__(self, subdomain_id):
        return 1.0


def main():
    sys.meta_path.insert(0, Finder())
    sys.path.insert(0, "/root/reference/src")
    import festim as F  # the real reference package, numerics stubbed out

    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    rng = np.random.default_rng(20261019)

    # --- Stepsize.modify_value -------------------------------------------------
    cases = []
    configs = [
        dict(initial_value=0.1, growth_factor=1.2, cutback_factor=0.8, target_nb_iterations=4),
        dict(initial_value=0.5, growth_factor=1.1, cutback_factor=0.9, target_nb_iterations=4,
             max_stepsize=2.0),
        dict(initial_value=1.0, growth_factor=1.5, cutback_factor=0.5, target_nb_iterations=3,
             milestones=[1.0, 2.5, 7.0, 20.0]),
        dict(initial_value=0.2, growth_factor=1.2, cutback_factor=0.9, target_nb_iterations=5,
             max_stepsize=3.0, milestones=[0.5, 4.0, 4.00001, 9.0], milestone_tolerance=1e-4),
    ]
    for cfg in configs:
        st = F.Stepsize(**cfg)
        for _ in range(60):
            value = float(rng.uniform(0.01, 4.0))
            t = float(rng.uniform(0.0, 22.0))
            if rng.random() < 0.2 and cfg.get("milestones"):
                t = float(rng.choice(cfg["milestones"])) * (1 + float(rng.uniform(-2e-5, 2e-5)))
            its = int(rng.integers(1, 9))
            cases.append({"config": cfg, "value": value, "nb_iterations": its, "t": t,
                          "expected": float(st.modify_value(value, its, t))})
    json.dump(cases, open(os.path.join(out_dir, "stepsize_modify_value.json"), "w"), indent=0)

    # --- convergenceTest -----------------------------------------------------------
    class Reasons:
        DIVERGED_MAX_IT, CONVERGED_FNORM_ABS, CONVERGED_FNORM_RELATIVE = -5, 2, 3
        CONVERGED_SNORM_RELATIVE, ITERATING = 4, 0

    class FakeSnes:
        ConvergedReason = Reasons

        def __init__(self, tol):
            self.tol = tol

        def getTolerances(self):
            return self.tol

    import festim.helpers as H

    seqs = []
    for _ in range(40):
        rtol = float(10 ** rng.uniform(-12, -6))
        atol = float(10 ** rng.uniform(-12, 2))
        stol = 1.4901161193847656e-10
        max_it = int(rng.integers(2, 8))
        snes = FakeSnes((rtol, atol, stol, max_it))
        f = float(10 ** rng.uniform(-3, 6))
        rows = []
        for it in range(max_it + 2):
            gnorm = 0.0 if it == 0 else float(10 ** rng.uniform(-13, 2))
            reason = H.convergenceTest(snes, it, (1.0, gnorm, f))
            rows.append({"it": it, "snorm": gnorm, "fnorm": f, "reason": int(reason)})
            f *= float(10 ** rng.uniform(-5, -0.5))
        seqs.append({"rtol": rtol, "atol": atol, "stol": stol, "max_it": max_it, "rows": rows})
    json.dump(seqs, open(os.path.join(out_dir, "convergence_test.json"), "w"), indent=0)

    # --- Reaction.reaction_term and Arrhenius diffusivity ---------------------------
    mat = F.Material(D_0=1.0, E_D=0.1)
    vol = F.VolumeSubdomain(id=1, material=mat)
    react = []
    for _ in range(60):
        k_0, E_k = float(10 ** rng.uniform(-20, 2)), float(rng.uniform(0.0, 1.5))
        mode = int(rng.integers(0, 3))
        p_0 = [0.0, float(10 ** rng.uniform(-3, 13)), float(10 ** rng.uniform(-3, 13))][mode]
        E_p = [0.0, 0.0, float(rng.uniform(0.1, 1.5))][mode]
        T = float(rng.uniform(300, 1200))
        cm, ct, n = (float(10 ** rng.uniform(0, 22)) for _ in range(3))
        A, B = F.Species("A"), F.Species("B", mobile=False)
        A.solution, B.solution = cm, ct
        imp = F.ImplicitSpecies(n=n, others=[B])
        imp.value_fenics = n
        r = F.Reaction(reactant=[A, imp], product=B, k_0=k_0, E_k=E_k, p_0=p_0, E_p=E_p, volume=vol)
        react.append({"k_0": k_0, "E_k": E_k, "p_0": p_0, "E_p": E_p, "T": T, "c_mobile": cm,
                      "c_trapped": ct, "n": n, "expected": float(r.reaction_term(T))})
    json.dump(react, open(os.path.join(out_dir, "reaction_term.json"), "w"), indent=0)

    diff = []
    for _ in range(40):
        D_0, E_D, T = float(10 ** rng.uniform(-9, 1)), float(rng.uniform(0, 1.5)), float(rng.uniform(250, 1500))
        m = F.Material(D_0=D_0, E_D=E_D)
        diff.append({"D_0": D_0, "E_D": E_D, "T": T,
                     "expected": float(m.get_diffusion_coefficient(None, T, None))})
    json.dump(diff, open(os.path.join(out_dir, "diffusion_coefficient.json"), "w"), indent=0)
    # --- SurfaceReactionBC flux (surface_reaction.py:57-85) ---------------------------------
    surf = F.SurfaceSubdomain(id=2)
    sreac = []
    for _ in range(40):
        k_r0, E_kr = float(10 ** rng.uniform(-30, 0)), float(rng.uniform(0, 1.5))
        k_d0, E_kd = float(10 ** rng.uniform(-10, 20)), float(rng.uniform(0, 1.5))
        T, P = float(rng.uniform(300, 1200)), float(10 ** rng.uniform(-3, 5))
        cA, cB = float(10 ** rng.uniform(0, 22)), float(10 ** rng.uniform(0, 22))
        same = bool(rng.random() < 0.3)
        A, B = F.Species("A"), F.Species("B")
        # ndarray: accepted by the value_fenics setter
        A.solution, B.solution = np.array([cA]).view(Immutable), np.array([cB]).view(Immutable)
        reactants = [A, A] if same else [A, B]
        bc = F.SurfaceReactionBC(reactant=reactants, gas_pressure=P, k_r0=k_r0, E_kr=E_kr, k_d0=k_d0,
                                 E_kd=E_kd, subdomain=surf)
        for fb in bc.flux_bcs:
            fb.create_value_fenics(None, T, None)
        sreac.append({"k_r0": k_r0, "E_kr": E_kr, "k_d0": k_d0, "E_kd": E_kd, "T": T, "P": P, "cA": cA,
                      "cB": cB, "same": same, "n_fluxes": len(bc.flux_bcs),
                      "flux_species": [fb.species.name for fb in bc.flux_bcs],
                      "expected": [float(np.ravel(fb.value_fenics)[0]) for fb in bc.flux_bcs]})
    json.dump(sreac, open(os.path.join(out_dir, "surface_reaction_flux.json"), "w"), indent=0)

    # --- Nitsche terms of a weakly enforced Dirichlet BC (dirichlet_bc.py:170-244) -----------
    nitsche = []
    for _ in range(40):
        gdim = int(rng.integers(1, 4))
        n = rng.standard_normal(gdim)
        n /= np.linalg.norm(n)
        CTX["n"], CTX["h"] = n, float(10 ** rng.uniform(-4, 0))
        u = WithGrad(rng.uniform(0.1, 10.0), rng.standard_normal(gdim) * 10)
        v = WithGrad(rng.uniform(0.0, 1.0), rng.standard_normal(gdim) * 10)
        D, g, penalty = float(10 ** rng.uniform(-10, 1)), float(rng.uniform(0.1, 10.0)), float(10 ** rng.uniform(0, 3))
        bc = F.FixedConcentrationBC(subdomain=surf, value=g, species=F.Species("H"), enforce_weakly=True,
                                    penalty=penalty)
        bc._value_fenics = g
        nitsche.append({"gdim": gdim, "n": n.tolist(), "h": CTX["h"], "u": float(u), "grad_u": u.grad.tolist(),
                        "v": float(v), "grad_v": v.grad.tolist(), "D": D, "g": g, "penalty": penalty,
                        "numerical_flux": float(bc.numerical_flux(u, D, None)),
                        "weak_form": float(bc.weak_formulation(u, v, FakeMeasure(), D))})
    json.dump(nitsche, open(os.path.join(out_dir, "nitsche_terms.json"), "w"), indent=0)

    # --- Sieverts' and Henry's laws (sieverts_bc.py:8-11, henrys_bc.py:8-11) ----------------
    from festim.boundary_conditions.henrys_bc import henrys_law
    from festim.boundary_conditions.sieverts_bc import sieverts_law

    laws = []
    for _ in range(40):
        T, c0, E, P = (float(rng.uniform(250, 1500)), float(10 ** rng.uniform(15, 24)), float(rng.uniform(0, 1.5)),
                       float(10 ** rng.uniform(-3, 6)))
        laws.append({"T": T, "c0": c0, "E": E, "P": P, "sieverts": float(sieverts_law(T, c0, E, P)),
                     "henry": float(henrys_law(T, c0, E, P))})
    json.dump(laws, open(os.path.join(out_dir, "solubility_laws.json"), "w"), indent=0)

    # --- the residual form itself: HydrogenTransportProblem.create_formulation --------------
    # (hydrogen_transport_problem.py:870-997).  The REAL method runs with every species'
    # solution / previous solution / test function replaced by numbers carrying their gradient
    # at one quadrature point and dx(id) / ds(id) replaced by the point's weight, so summing
    # `problem.formulation` over the Gauss points of every cell (and the boundary vertices)
    # gives the reference's element-by-element assembled residual on a small 1D mesh: diffusion
    # with per-material, per-species Arrhenius D, mass term, trapping reactions restricted to
    # their volumes (Arrhenius and constant detrapping), sources, particle fluxes, and the
    # steady-state `c = 0` filler for immobile species outside their reactions' volumes.  The
    # temperature is a constant, so all integrands are polynomials (degree <= 3) and 3-point
    # Gauss-Legendre is exact: any conforming P1 assembly must reproduce these numbers to
    # round-off.  F is quadratic in u, so (F(u + d) - F(u - d)) / 2 is J(u) d exactly.
    import ufl as _ufl_stub

    _ufl_stub.dot = lambda a, b: float(np.dot(a, b))

    class PointMeasure:
        def __init__(self, kind):
            self.kind = kind

        def __call__(self, subdomain_id):
            return CTX[self.kind].get(subdomain_id, 0.0)

    verts = np.array([0.0, 0.2, 0.5, 0.8, 1.0])
    par = {
        "vertices": verts.tolist(), "T": 500.0, "dt": 0.25,
        "D_0": {"1": {"H": 2.0, "D": 1.4}, "2": {"H": 0.5, "D": 0.35}},
        "E_D": {"1": {"H": 0.1, "D": 0.12}, "2": {"H": 0.2, "D": 0.22}},
        "reactions": [
            {"volume": 1, "mobile": "H", "trapped": "t1", "n": 3.0, "k_0": 1.5, "E_k": 0.2, "p_0": 0.3, "E_p": 0.1},
            {"volume": 2, "mobile": "D", "trapped": "t2", "n": 2.0, "k_0": 0.8, "E_k": 0.15, "p_0": 0.05, "E_p": 0.0},
        ],
        "sources": [{"volume": 2, "species": "H", "value": 0.7}, {"volume": 1, "species": "t1", "value": 0.2}],
        "fluxes": [{"surface": 3, "x": 1.0, "species": "H", "value": 0.4},
                   {"surface": 4, "x": 0.0, "species": "D", "value": -0.3}],
    }
    names = ["H", "D", "t1", "t2"]
    ns, nv = len(names), len(verts)
    u = rng.uniform(0.2, 1.5, (nv, ns))
    u_n = u * rng.uniform(0.8, 1.1, (nv, ns))
    delta = rng.uniform(-0.3, 0.3, (nv, ns))

    vel_nodal = rng.uniform(-0.8, 0.8, nv)          # P1 velocity of the advection variant

    # Q7 of SURVEY Appendix B: a species repeated in `reactant` gets the term once per occurrence
    repeated = {"volume": 2, "reactants": ["H", "H"], "product": "t2", "k_0": 0.6, "E_k": 0.05, "p_0": 0.07, "E_p": 0.02}

    def reference_residual(uvals, transient, advection=False, with_repeated=False, first_reaction_only=False):
        prob = F.HydrogenTransportProblem()
        species = [F.Species("H"), F.Species("D"), F.Species("t1", mobile=False), F.Species("t2", mobile=False)]
        by_name = dict(zip(names, species))
        prob.species = species
        vols = {i: F.VolumeSubdomain1D(id=i, borders=b, material=F.Material(D_0=par["D_0"][str(i)], E_D=par["E_D"][str(i)]))
                for i, b in ((1, [0.0, 0.5]), (2, [0.5, 1.0]))}
        surfs = {f["surface"]: F.SurfaceSubdomain1D(id=f["surface"], x=f["x"]) for f in par["fluxes"]}
        prob.subdomains = [*vols.values(), *surfs.values()]
        reactions = []
        for r in par["reactions"][: 1 if first_reaction_only else None]:
            imp = F.ImplicitSpecies(n=r["n"], others=[by_name[r["trapped"]]])
            imp.value_fenics = r["n"]
            reactions.append(F.Reaction(reactant=[by_name[r["mobile"]], imp], product=by_name[r["trapped"]],
                                        k_0=r["k_0"], E_k=r["E_k"], p_0=r["p_0"], E_p=r["E_p"],
                                        volume=vols[r["volume"]]))
        if with_repeated:
            r = repeated
            reactions.append(F.Reaction(reactant=[by_name[n] for n in r["reactants"]], product=by_name[r["product"]],
                                        k_0=r["k_0"], E_k=r["E_k"], p_0=r["p_0"], E_p=r["E_p"],
                                        volume=vols[r["volume"]]))
        prob.reactions = reactions
        sources = []
        for sdef in par["sources"]:
            src = F.ParticleSource(value=1.0, volume=vols[sdef["volume"]], species=by_name[sdef["species"]])
            src.value.fenics_object = sdef["value"]
            sources.append(src)
        prob.sources = sources
        bcs = []
        for f in par["fluxes"]:
            bc = F.ParticleFluxBC(subdomain=surfs[f["surface"]], value=1.0, species=by_name[f["species"]])
            bc._value_fenics = f["value"]
            bcs.append(bc)
        prob.boundary_conditions = bcs
        prob.settings = F.Settings(atol=1, rtol=1, transient=transient, final_time=1 if transient else None)
        prob.mesh = F.Mesh.__new__(F.Mesh)
        prob.mesh._mesh = None
        prob.mesh.coordinate_system = "cartesian"
        prob.temperature_fenics = Constant(None, par["T"])
        prob.dx, prob.ds = PointMeasure("dx"), PointMeasure("ds")
        prob._dt = Constant(None, par["dt"])
        prob.advection_terms = []
        vel_holder = types.SimpleNamespace(fenics_object=np.zeros(1))
        if advection:   # hydrogen_transport_problem.py:967-978: (grad(c) . vel) v dx on H in volume 1
            prob.advection_terms = [types.SimpleNamespace(species=[species[0]], velocity=vel_holder,
                                                          subdomain=vols[1])]

        def point(values, grads, prev, row_species, phi, gphi, dx, ds, vel=0.0):
            vel_holder.fenics_object = np.array([vel])
            for k, spe in enumerate(species):
                spe.solution = WithGrad(values[k], [grads[k]])
                spe.prev_solution = float(prev[k])
                on = k == row_species
                spe.test_function = WithGrad(phi if on else 0.0, [gphi if on else 0.0])
            CTX["dx"], CTX["ds"] = dx, ds
            prob.create_formulation()
            return float(prob.formulation)

        res = np.zeros((nv, ns))
        gx, gw = np.polynomial.legendre.leggauss(3)
        for c in range(nv - 1):
            a, b = verts[c], verts[c + 1]
            h = b - a
            vol_id = 1 if b <= 0.5 + 1e-12 else 2
            for xi, w in zip(gx, gw):
                lam = np.array([0.5 * (1 - xi), 0.5 * (1 + xi)])      # P1 basis at the point
                glam = np.array([-1.0 / h, 1.0 / h])
                vals = lam @ uvals[c:c + 2]
                grads = glam @ uvals[c:c + 2]
                prev = lam @ u_n[c:c + 2]
                for loc in range(2):
                    for srow in range(ns):
                        res[c + loc, srow] += point(vals, grads, prev, srow, lam[loc], glam[loc],
                                                    {vol_id: 0.5 * h * w}, {}, vel=float(lam @ vel_nodal[c:c + 2]))
        for f in par["fluxes"]:                                     # boundary vertices: ds = point measure
            vtx = int(np.argmin(np.abs(verts - f["x"])))
            cell = min(vtx, nv - 2)
            h = verts[cell + 1] - verts[cell]
            grads = np.array([-1.0 / h, 1.0 / h]) @ uvals[cell:cell + 2]
            for srow in range(ns):
                res[vtx, srow] += point(uvals[vtx], grads, u_n[vtx], srow, 1.0, 0.0, {}, {f["surface"]: 1.0})
        return res

    form = {"parameters": par, "species": names, "u": u.tolist(), "u_n": u_n.tolist(), "delta": delta.tolist()}
    for label, transient in (("transient", True), ("steady", False)):
        form[f"F_{label}"] = reference_residual(u, transient).tolist()
        form[f"Jdelta_{label}"] = (0.5 * (reference_residual(u + delta, transient)
                                          - reference_residual(u - delta, transient))).tolist()
    form["velocity"] = vel_nodal.tolist()
    form["F_advection"] = reference_residual(u, True, advection=True).tolist()
    form["Jdelta_advection"] = (0.5 * (reference_residual(u + delta, True, advection=True)
                                       - reference_residual(u - delta, True, advection=True))).tolist()
    # steady state with NO reaction in volume 2: the immobile species get `c v dx(2)` there
    # (:980-997 looks for any reaction in the volume, not for one involving the species)
    form["F_steady_filler"] = reference_residual(u, False, first_reaction_only=True).tolist()
    form["Jdelta_steady_filler"] = (0.5 * (reference_residual(u + delta, False, first_reaction_only=True)
                                           - reference_residual(u - delta, False, first_reaction_only=True))).tolist()
    form["repeated_reactant_reaction"] = repeated
    form["F_repeated"] = reference_residual(u, True, with_repeated=True).tolist()
    form["Jdelta_repeated"] = (0.5 * (reference_residual(u + delta, True, with_repeated=True)
                                      - reference_residual(u - delta, True, with_repeated=True))).tolist()
    json.dump(form, open(os.path.join(out_dir, "residual_form.json"), "w"), indent=0)

    # --- HeatTransferProblem.create_formulation (heat_transfer_problem.py:175-218), same method:
    # conductivity a + b T in volume 1 (callable of T) and constant in volume 2, rho cp dT/dt,
    # a volumetric source and a heat flux; quadratic in T, so the central difference is J d exactly.
    hpar = {"vertices": verts.tolist(), "dt": 0.25,
            "materials": {"1": {"lambda_a": 1.5, "lambda_b": 0.01, "density": 2.0, "heat_capacity": 3.0},
                          "2": {"lambda_a": 4.0, "lambda_b": 0.0, "density": 1.2, "heat_capacity": 2.5}},
            "source": {"volume": 1, "value": 0.9}, "flux": {"surface": 3, "x": 1.0, "value": 0.6}}
    Tn = rng.uniform(300.0, 400.0, nv)
    Told = Tn * rng.uniform(0.95, 1.02, nv)
    dT = rng.uniform(-20.0, 20.0, nv)

    def reference_heat_residual(Tvals, transient):
        prob = F.HeatTransferProblem()
        vols = {}
        for i, b in ((1, [0.0, 0.5]), (2, [0.5, 1.0])):
            mp = hpar["materials"][str(i)]
            lam = (lambda T, a=mp["lambda_a"], bb=mp["lambda_b"]: a + bb * T) if mp["lambda_b"] else mp["lambda_a"]
            vols[i] = F.VolumeSubdomain1D(id=i, borders=b, material=F.Material(
                D_0=1.0, E_D=0.0, thermal_conductivity=lam, density=mp["density"], heat_capacity=mp["heat_capacity"]))
        surf = F.SurfaceSubdomain1D(id=hpar["flux"]["surface"], x=hpar["flux"]["x"])
        prob.subdomains = [*vols.values(), surf]
        src = F.HeatSource(value=1.0, volume=vols[hpar["source"]["volume"]])
        src.value.fenics_object = hpar["source"]["value"]
        prob.sources = [src]
        bc = F.HeatFluxBC(subdomain=surf, value=1.0)
        bc._value_fenics = hpar["flux"]["value"]
        prob.boundary_conditions = [bc]
        prob.settings = F.Settings(atol=1, rtol=1, transient=transient, final_time=1 if transient else None)
        prob.dx, prob.ds = PointMeasure("dx"), PointMeasure("ds")
        prob._dt = Constant(None, hpar["dt"])

        def point(value, grad, prev, phi, gphi, dx, ds):
            prob.u = WithGrad(value, [grad])
            prob.u_n = float(prev)
            prob.test_function = WithGrad(phi, [gphi])
            CTX["dx"], CTX["ds"] = dx, ds
            prob.create_formulation()
            return float(prob.formulation)

        res = np.zeros(nv)
        gx, gw = np.polynomial.legendre.leggauss(3)
        for c in range(nv - 1):
            h = verts[c + 1] - verts[c]
            vol_id = 1 if verts[c + 1] <= 0.5 + 1e-12 else 2
            for xi, w in zip(gx, gw):
                lam = np.array([0.5 * (1 - xi), 0.5 * (1 + xi)])
                glam = np.array([-1.0 / h, 1.0 / h])
                for loc in range(2):
                    res[c + loc] += point(lam @ Tvals[c:c + 2], glam @ Tvals[c:c + 2], lam @ Told[c:c + 2],
                                          lam[loc], glam[loc], {vol_id: 0.5 * h * w}, {})
        vtx = int(np.argmin(np.abs(verts - hpar["flux"]["x"])))
        res[vtx] += point(Tvals[vtx], 0.0, Told[vtx], 1.0, 0.0, {}, {hpar["flux"]["surface"]: 1.0})
        return res

    heat = {"parameters": hpar, "T": Tn.tolist(), "T_n": Told.tolist(), "delta": dT.tolist()}
    for label, transient in (("transient", True), ("steady", False)):
        heat[f"F_{label}"] = reference_heat_residual(Tn, transient).tolist()
        heat[f"Jdelta_{label}"] = (0.5 * (reference_heat_residual(Tn + dT, transient)
                                          - reference_heat_residual(Tn - dT, transient))).tolist()
    json.dump(heat, open(os.path.join(out_dir, "heat_form.json"), "w"), indent=0)

    # --- cylindrical / spherical diffusion term (hydrogen_transport_problem.py:886-908) -------
    # r^m D grad(u) . grad(v / r^m): v / r^m needs the quotient rule, so the stand-ins here are
    # forward-mode dual numbers.  The integrand is rational in r, hence quadrature dependent: the
    # residual is tabulated for 1..8 Gauss-Legendre points per cell (on an interval a rule is
    # fixed by its number of points) and the test looks up the size the estimated degree asks for.
    class Dual:
        def __init__(self, value, grad):
            self.value, self.grad = float(value), np.asarray(grad, dtype=float)

        @staticmethod
        def lift(o):
            return o if isinstance(o, Dual) else Dual(o, [0.0])

        def __add__(self, o):
            o = Dual.lift(o)
            return Dual(self.value + o.value, self.grad + o.grad)

        __radd__ = __add__

        def __neg__(self):
            return Dual(-self.value, -self.grad)

        def __sub__(self, o):
            return self + (-Dual.lift(o))

        def __rsub__(self, o):
            return Dual.lift(o) + (-self)

        def __mul__(self, o):
            o = Dual.lift(o)
            return Dual(self.value * o.value, self.grad * o.value + self.value * o.grad)

        __rmul__ = __mul__

        def __truediv__(self, o):
            o = Dual.lift(o)
            return Dual(self.value / o.value, (self.grad * o.value - self.value * o.grad) / o.value**2)

        def __rtruediv__(self, o):
            return Dual.lift(o) / self

        def __pow__(self, n):
            return Dual(self.value**n, n * self.value ** (n - 1) * self.grad)

        def __float__(self):
            return self.value

    _ufl_stub.grad = lambda a: a.grad
    _ufl_stub.SpatialCoordinate = lambda mesh: [Dual(CTX["x"], [1.0])]
    rverts = 1.0 + verts
    cpar = {"vertices": rverts.tolist(), "T": 450.0, "D_0": 1.7, "E_D": 0.15, "source": 0.7}
    cu = rng.uniform(0.5, 1.5, nv)

    def reference_curvilinear(system, npts):
        prob = F.HydrogenTransportProblem()
        H = F.Species("H")
        prob.species = [H]
        vol = F.VolumeSubdomain1D(id=1, borders=[rverts[0], rverts[-1]], material=F.Material(D_0=cpar["D_0"], E_D=cpar["E_D"]))
        prob.subdomains = [vol]
        src = F.ParticleSource(value=1.0, volume=vol, species=H)
        src.value.fenics_object = cpar["source"]
        prob.sources = [src]
        prob.settings = F.Settings(atol=1, rtol=1, transient=False)
        prob.mesh = F.Mesh.__new__(F.Mesh)
        prob.mesh._mesh = None
        prob.mesh.coordinate_system = system
        prob.temperature_fenics = Constant(None, cpar["T"])
        prob.dx, prob.ds = PointMeasure("dx"), PointMeasure("ds")
        prob.advection_terms = []
        prob.reactions = []
        prob.boundary_conditions = []
        res = np.zeros(nv)
        gx, gw = np.polynomial.legendre.leggauss(npts)
        for c in range(nv - 1):
            a, b = rverts[c], rverts[c + 1]
            h = b - a
            for xi, w in zip(gx, gw):
                lam = np.array([0.5 * (1 - xi), 0.5 * (1 + xi)])
                glam = np.array([-1.0 / h, 1.0 / h])
                CTX["x"] = float(lam @ rverts[c:c + 2])
                CTX["dx"], CTX["ds"] = {1: 0.5 * h * w}, {}
                for loc in range(2):
                    H.solution = Dual(lam @ cu[c:c + 2], [glam @ cu[c:c + 2]])
                    H.prev_solution = 0.0
                    H.test_function = Dual(lam[loc], [glam[loc]])
                    prob.create_formulation()
                    res[c + loc] += float(prob.formulation)
        return res

    curv = {"parameters": cpar, "u": cu.tolist()}
    for system in ("cylindrical", "spherical"):
        curv[system] = {str(k): reference_curvilinear(system, k).tolist() for k in range(1, 9)}
    json.dump(curv, open(os.path.join(out_dir, "curvilinear_form.json"), "w"), indent=0)

    # --- the time loop: ProblemBase.run / iterate (problem.py:195-255) ----------------------
    # The REAL loop with a scripted solver (Newton iteration counts from a list): recorded
    # `timesteps`, the step size after every step and the final time pin the bookkeeping order
    # (t advanced before the solve, modify_value called with the new t, no clamping of the last
    # step - SURVEY quirk Q1 - milestones, max_stepsize as number or callable of t).
    class Box:
        def __init__(self, value):
            self.value = float(value)

        def __float__(self):
            return float(self.value)

    class ScriptedProblem(F.problem.ProblemBase):
        def update_time_dependent_values(self):
            pass

        def post_processing(self):
            self.dts.append(float(self._dt.value))

    loops = []
    loop_configs = [
        dict(final_time=10.0, stepsize=dict(initial_value=0.5)),
        dict(final_time=1.0, stepsize=dict(initial_value=0.3)),
        dict(final_time=20.0, its_high=7, stepsize=dict(initial_value=0.1, growth_factor=1.2, cutback_factor=0.8,
                                                        target_nb_iterations=4)),
        dict(final_time=50.0, its_high=5, stepsize=dict(initial_value=0.5, growth_factor=1.5, cutback_factor=0.5,
                                                        target_nb_iterations=3,
                                                        milestones=[1.0, 2.5, 7.0, 20.0, 50.0])),
        dict(final_time=30.0, stepsize=dict(initial_value=0.2, growth_factor=1.2, cutback_factor=0.9,
                                            target_nb_iterations=5, max_stepsize=3.0,
                                            milestones=[0.5, 4.0, 9.0])),
        dict(final_time=40.0, its_high=6, stepsize=dict(initial_value=0.2, growth_factor=1.3, cutback_factor=0.7,
                                                        target_nb_iterations=4, milestones=[10.0, 25.0]),
             max_stepsize_ramp=[10.0, 0.5, 5.0]),      # max_stepsize(t) = 0.5 if t < 10 else 5.0
    ]
    for cfg in loop_configs:
        its = [int(v) for v in rng.integers(1, cfg.get("its_high", 9), 20000)]
        kw = dict(cfg["stepsize"])
        if "max_stepsize_ramp" in cfg:
            t0, lo, hi = cfg["max_stepsize_ramp"]
            kw["max_stepsize"] = lambda t, t0=t0, lo=lo, hi=hi: lo if t < t0 else hi
        prob = ScriptedProblem()
        prob.settings = F.Settings(atol=1, rtol=1, final_time=cfg["final_time"])
        prob.settings.stepsize = F.Stepsize(**kw)
        prob.show_progress_bar = False
        prob.t = Box(0.0)
        prob._dt = Box(prob.settings.stepsize.initial_value)
        prob.dts = []
        script = iter(its)
        snes = types.SimpleNamespace(getConvergedReason=lambda: 2, getIterationNumber=lambda: next(script))
        prob.solver = types.SimpleNamespace(solve=lambda: None, solver=snes)
        prob.u = types.SimpleNamespace(x=types.SimpleNamespace(array=np.zeros(3)))
        prob.u_n = types.SimpleNamespace(x=types.SimpleNamespace(array=np.zeros(3)))
        prob.run()
        n = len(prob.timesteps)
        loops.append({"config": cfg, "iterations": its[:n], "timesteps": list(prob.timesteps),
                      "dt_during_step": prob.dts, "final_t": float(prob.t), "final_dt": float(prob._dt.value)})
    # the staggered loop of CoupledTransientHeatTransferHydrogenTransport (coupled_heat_hydrogen_
    # problem.py:133-163): both problems step, then both take the smaller of the two new step sizes
    def scripted(cls, cfg, its):
        class Scripted(cls):
            def update_time_dependent_values(self):
                pass

            def post_processing(self):
                self.dts.append(float(self._dt.value))

        prob = Scripted()
        prob.settings = F.Settings(atol=1, rtol=1, final_time=cfg["final_time"], transient=True)
        prob.settings.stepsize = F.Stepsize(**cfg["stepsize"])
        prob.show_progress_bar = False
        prob.t = Box(0.0)
        prob._dt = Box(prob.settings.stepsize.initial_value)
        prob.dts = []
        script = iter(its)
        snes = types.SimpleNamespace(getConvergedReason=lambda: 2, getIterationNumber=lambda: next(script))
        prob.solver = types.SimpleNamespace(solve=lambda: None, solver=snes)
        prob.u = types.SimpleNamespace(x=types.SimpleNamespace(array=np.zeros(3)))
        prob.u_n = types.SimpleNamespace(x=types.SimpleNamespace(array=np.zeros(3)))
        prob.mesh = F.Mesh.__new__(F.Mesh)
        prob.mesh._mesh = None
        return prob

    ccfg = {"heat": dict(final_time=15.0, stepsize=dict(initial_value=0.3, growth_factor=1.25, cutback_factor=0.9,
                                                        target_nb_iterations=4, milestones=[5.0])),
            "hydrogen": dict(final_time=15.0, stepsize=dict(initial_value=0.2, growth_factor=1.15, cutback_factor=0.85,
                                                            target_nb_iterations=3, max_stepsize=1.5))}
    its_heat = [int(v) for v in rng.integers(1, 6, 5000)]
    its_hyd = [int(v) for v in rng.integers(1, 4, 5000)]
    heat_p = scripted(F.HeatTransferProblem, ccfg["heat"], its_heat)
    hyd_p = scripted(F.HydrogenTransportProblem, ccfg["hydrogen"], its_hyd)
    coupled = F.CoupledTransientHeatTransferHydrogenTransport.__new__(F.CoupledTransientHeatTransferHydrogenTransport)
    coupled._heat_problem, coupled._hydrogen_problem = heat_p, hyd_p
    hyd_p.show_progress_bar = False
    # what initialise() does before the loop (:103-108): both start from the smaller initial step
    dt0 = min(heat_p.settings.stepsize.initial_value, hyd_p.settings.stepsize.initial_value)
    for pr in (heat_p, hyd_p):
        pr.settings.stepsize.initial_value = dt0
        pr._dt.value = dt0
    coupled.run()
    nsteps = len(hyd_p.timesteps)
    loops.append({"coupled": ccfg, "iterations_heat": its_heat[:nsteps], "iterations_hydrogen": its_hyd[:nsteps],
                  "timesteps": list(hyd_p.timesteps), "timesteps_heat": list(heat_p.timesteps),
                  "dt_during_step": hyd_p.dts, "dt_during_step_heat": heat_p.dts,
                  "final_t": float(hyd_p.t), "final_dt": float(hyd_p._dt.value)})
    json.dump(loops, open(os.path.join(out_dir, "time_loop.json"), "w"), indent=0)

    # --- per-species material properties (material.py:148-260): number or dict keyed by the
    # species object or its name; what a missing key / missing species / wrong type raises
    lookups = []
    getters = ["get_D_0", "get_E_D", "get_K_S_0", "get_E_K_S"]
    for getter in getters:
        attr = getter[4:]
        for spec_kind in ("number", "by_name", "by_object", "wrong_type"):
            for query in ("A", "B", "missing", None):
                A, B, C = F.Species("A"), F.Species("B"), F.Species("missing")
                species = {"A": A, "B": B, "missing": C, None: None}[query]
                value = {"number": 1.25, "by_name": {"A": 1.5, "B": 2.5}, "by_object": {A: 3.5, B: 4.5},
                         "wrong_type": "oops"}[spec_kind]
                kwargs = dict(D_0=1.0, E_D=0.1)
                kwargs[attr] = value
                if attr == "D_0" and isinstance(value, dict):
                    kwargs["E_D"] = {k: 0.1 for k in value}
                if attr == "E_D" and isinstance(value, dict):
                    kwargs["D_0"] = {k: 1.0 for k in value}
                try:
                    mat = F.Material(**kwargs)
                    got = getattr(mat, getter)(species=species)
                    outcome = {"value": float(got)}
                except Exception as exc:   # noqa: BLE001
                    outcome = {"error": type(exc).__name__, "message": str(exc)}
                lookups.append({"getter": getter, "spec": spec_kind, "query": query, **outcome})
    json.dump(lookups, open(os.path.join(out_dir, "material_lookups.json"), "w"), indent=0)

    # --- small host rules on the path -------------------------------------------------------
    # is_it_time_to_export (helpers.py:374-401): matching tolerance and consumption of the times;
    # Mesh1D.check_borders (mesh_1d.py:109-155): which subdomain layouts are accepted.
    import festim.helpers as H2

    export_cases = []
    for _ in range(12):
        times = sorted(float(v) for v in rng.uniform(0.0, 10.0, int(rng.integers(1, 6))))
        queries = []
        work = list(times)
        for _ in range(14):
            if rng.random() < 0.5 and times:
                tq = float(rng.choice(times)) * (1.0 + float(rng.choice([0.0, 4e-6, -4e-6, 3e-5, -3e-5])))
            else:
                tq = float(rng.uniform(0.0, 10.0))
            hit = bool(H2.is_it_time_to_export(times=work, current_time=tq))
            queries.append({"t": tq, "hit": hit, "remaining": list(work)})
        export_cases.append({"times": times, "queries": queries})
    export_cases.append({"times": None, "queries": [{"t": 1.0, "hit": bool(H2.is_it_time_to_export(None, 1.0)),
                                                     "remaining": None}]})
    border_cases = []
    layouts = [
        ([0.0, 1.0, 2.0, 4.0], [[0, 2], [2, 4]]), ([0.0, 1.0, 2.0, 4.0], [[0, 4]]),
        ([0.0, 1.0, 2.0, 4.0], [[0, 1.5], [2, 4]]), ([0.0, 1.0, 2.0, 4.0], [[0, 2], [2, 3]]),
        ([0.0, 1.0, 2.0, 4.0], [[1, 2], [2, 4]]), ([0.0, 1.0, 2.0, 4.0], []),
        ([0.0, 1.0, 2.0, 4.0], [[2, 4], [0, 2]]), ([0.0, 1.0, 2.0, 4.0], [[0, 2], [2, 5]]),
        ([[0.0, 1.0, 2.0], [3.0, 4.0]], [[0, 2], [3, 4]]), ([[0.0, 1.0, 2.0], [3.0, 4.0]], [[0, 2]]),
        ([[0.0, 1.0, 2.0], [3.0, 4.0]], [[0, 4]]), ([[0.0, 1.0, 2.0], [3.0, 4.0]], [[0, 1], [1, 2], [3, 4]]),
    ]
    mat1 = F.Material(D_0=1.0, E_D=0.0)
    for vertices, borders in layouts:
        try:
            mesh1d = F.Mesh1D.__new__(F.Mesh1D)      # (the constructor needs a real dolfinx mesh)
            mesh1d.vertices = ([np.array(b) for b in vertices] if isinstance(vertices[0], list)
                               else np.array(vertices))
            vols = [F.VolumeSubdomain1D(id=i + 1, borders=b, material=mat1) for i, b in enumerate(borders)]
            mesh1d.check_borders(vols)
            outcome = {"ok": True}
        except Exception as exc:   # noqa: BLE001
            outcome = {"ok": False, "error": type(exc).__name__, "message": str(exc)}
        border_cases.append({"vertices": vertices, "borders": borders, **outcome})
    json.dump({"is_it_time_to_export": export_cases, "check_borders": border_cases},
              open(os.path.join(out_dir, "host_rules.json"), "w"), indent=0)

    json.dump({"k_B": F.k_B, "R": F.R, "k_B_SI": F.k_B_SI},
              open(os.path.join(out_dir, "constants.json"), "w"))
    print("golden vectors written to", out_dir)


if __name__ == "__main__":
    main()
