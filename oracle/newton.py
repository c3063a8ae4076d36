"""Newton solve and derived quantities of the CPU oracle (TEST INFRASTRUCTURE).

Restates PETSc ``SNES newtonls`` with line search ``none`` and ``KSP preonly`` +
``PC lu`` as configured at reference ``src/festim/problem.py:271-289``, with
the user convergence test of ``src/festim/helpers.py:408-426``:
at iteration ``it`` with residual norm f and last step norm s,
``it > max_it`` -> DIVERGED_MAX_IT; ``f < atol`` -> CONVERGED_FNORM_ABS;
``f / f0 < rtol`` -> CONVERGED_FNORM_RELATIVE; ``s < stol and it > 0`` ->
CONVERGED_SNORM_RELATIVE.
"""

from __future__ import annotations

import numpy as np
import scipy.sparse.linalg as spla

from festim_b200 import fem
from festim_b200.constants import k_B
from oracle.assembly import FormAssembler, apply_dirichlet, cell_geometry, facet_measures

CONVERGED_FNORM_ABS = 2
CONVERGED_FNORM_RELATIVE = 3
CONVERGED_SNORM_RELATIVE = 4
DIVERGED_MAX_IT = -5
DIVERGED_LINEAR_SOLVE = -3
ITERATING = 0


def convergence_test(it, snorm, fnorm, f0, rtol, atol, stol, max_it):
    if it > max_it:
        return DIVERGED_MAX_IT
    if fnorm < atol:
        return CONVERGED_FNORM_ABS
    if fnorm / f0 < rtol:
        return CONVERGED_FNORM_RELATIVE
    if snorm < stol and it > 0:
        return CONVERGED_SNORM_RELATIVE
    return ITERATING


def lu_solve(J, b):
    """Exact solve standing for KSP preonly + PC lu (MUMPS / SuperLU_dist scale
    the matrix before factorising): row equilibration, SuperLU, one step of
    iterative refinement so every species equation is solved to round-off even
    when their scales differ by many orders of magnitude."""
    import scipy.sparse as sp

    J = J.tocsr()
    scale = 1.0 / np.maximum(abs(J).max(axis=1).toarray().ravel(), 1e-300)
    A = (sp.diags(scale) @ J).tocsc()
    rhs = scale * b
    lu = spla.splu(A)
    y = lu.solve(rhs)
    y += lu.solve(rhs - A @ y)
    return y


class _Snes:
    def __init__(self):
        self.reason = 0
        self.its = 0

    def getConvergedReason(self):
        return self.reason

    def getIterationNumber(self):
        return self.its


class OracleNewtonSolver:
    def __init__(self, problem, petsc_options):
        self.problem = problem
        self.opts = petsc_options
        self.atol = petsc_options["snes_atol"]
        self.rtol = petsc_options["snes_rtol"]
        self.stol = petsc_options["snes_stol"]
        self.max_it = petsc_options["snes_max_it"]
        self.solver = _Snes()
        species = getattr(problem, "species", None) or [problem._unknown]
        self.assembler = FormAssembler(problem.mesh.mesh, problem.formulation, species,
                                       problem.volume_meshtags, problem.facet_meshtags)
        self.fnorms = []
        self.last_J = None
        self.last_b = None
        self._D_T = None

    def bc_arrays(self):
        dofs, vals = [], []
        for bc in self.problem.bc_forms:
            dofs.append(bc.dofs)
            vals.append(bc.values())
        if not dofs:
            return np.zeros(0, dtype=np.int64), np.zeros(0)
        dofs = np.concatenate(dofs)
        vals = np.concatenate(vals)
        # later BCs overwrite earlier ones on shared dofs (set_bc order)
        _, idx = np.unique(dofs[::-1], return_index=True)
        idx = len(dofs) - 1 - idx
        return dofs[idx], vals[idx]

    def residual_and_jacobian(self, x, want_J=True):
        u_n = self.problem.u_n.x.array
        F, J = self.assembler.assemble(x, u_n, want_J=True)
        bc_dofs, bc_vals = self.bc_arrays()
        b, Jbc = apply_dirichlet(F, J, x, bc_dofs, bc_vals)
        return b, Jbc

    def solve(self):
        x = self.problem.u.x.array
        atol = self.atol(float(self.problem.t)) if callable(self.atol) else self.atol
        rtol = self.rtol(float(self.problem.t)) if callable(self.rtol) else self.rtol
        it, snorm, f0 = 0, 0.0, None
        self.fnorms = []
        while True:
            b, J = self.residual_and_jacobian(x)
            self.last_J, self.last_b = J, b
            fnorm = float(np.linalg.norm(b))
            self.fnorms.append(fnorm)
            if it == 0:
                f0 = fnorm
            reason = convergence_test(it, snorm, fnorm, f0 if f0 > 0 else 1.0, rtol, atol,
                                      self.stol, self.max_it)
            if f0 == 0.0 and reason == ITERATING:
                reason = CONVERGED_FNORM_ABS
            if reason != ITERATING:
                break
            if it >= self.max_it:
                reason = DIVERGED_MAX_IT
                break
            y = lu_solve(J, b)
            if not np.all(np.isfinite(y)):
                reason = DIVERGED_LINEAR_SOLVE
                break
            x -= y
            snorm = float(np.linalg.norm(y))
            it += 1
        self.solver.reason = reason
        self.solver.its = it
        return it

    # ---- derived quantities ----------------------------------------
    def set_D_global_temperature(self, T):
        self._D_T = T

    def total_volume(self, species_index, vol_id):
        p = self.problem
        mesh = p.mesh.mesh
        cells = p.volume_meshtags.find(vol_id)
        vol, _, _ = cell_geometry(mesh, cells)
        ns = self.assembler.ns
        us = p.u.x.array.reshape(-1, ns)[:, species_index][mesh.cells[cells]]
        return float(np.sum(vol * us.mean(axis=1)))

    def measure_volume(self, vol_id):
        p = self.problem
        vol, _, _ = cell_geometry(p.mesh.mesh, p.volume_meshtags.find(vol_id))
        return float(vol.sum())

    def measure_surface(self, surf_id):
        p = self.problem
        mesh = p.mesh.mesh
        return float(facet_measures(mesh, mesh.bfacets[p.facet_meshtags.find(surf_id)]).sum())

    def total_surface(self, species_index, surf_id):
        p = self.problem
        mesh = p.mesh.mesh
        fverts = mesh.bfacets[p.facet_meshtags.find(surf_id)]
        ns = self.assembler.ns
        us = p.u.x.array.reshape(-1, ns)[:, species_index][fverts]
        return float(np.sum(facet_measures(mesh, fverts) * us.mean(axis=1)))

    def integrate_custom(self, export):
        """Integral of a CustomQuantity's expression over its subdomain: numpy evaluation of the
        symbolic expression at the quadrature points of its estimated degree."""
        from festim_b200 import quadrature, ufl
        from oracle.assembly import cell_geometry, circumradius, facet_measures

        p, asm = self.problem, self.assembler
        mesh, d = p.mesh.mesh, p.mesh.mesh.tdim
        sub = export.subdomain
        surface = not hasattr(sub, "material")
        degree = ufl.estimate_degree(export.ufl_expr)
        u = p.u.x.array
        if surface:
            facets = p.facet_meshtags.find(sub.id)
            fverts, fcells = mesh.bfacets[facets], mesh.bfacet_cell[facets]
            _, G, _ = cell_geometry(mesh, fcells)
            cverts = mesh.cells[fcells]
            asm._cell_verts = cverts
            opp = np.array([[v not in fv for v in cv] for cv, fv in zip(cverts, fverts)])
            normal = -G[opp] / np.linalg.norm(G[opp], axis=1)[:, None]
            bary, w = quadrature.rule(d - 1, degree)
            env = asm._env(fverts, bary, G, u, u, normal, circumradius(mesh, fcells))
            measure = facet_measures(mesh, fverts)
        else:
            cells = p.volume_meshtags.find(sub.id)
            measure, G, _ = cell_geometry(mesh, cells)
            verts = mesh.cells[cells]
            asm._cell_verts = verts
            bary, w = quadrature.rule(d, degree)
            env = asm._env(verts, bary, G, u, u, np.zeros((len(cells), 3)), circumradius(mesh, cells))
        vq = np.broadcast_to(ufl.evaluate(export.ufl_expr, env), (len(measure), len(w)))
        return float(np.sum(measure * (vq @ w)))

    def surface_flux(self, species_index, surf_id):
        """int -D_h grad(u).n ds with D_h the DG1 interpolant of the Arrhenius
        diffusivity (reference ``exports/surface_flux.py:39-60``,
        ``hydrogen_transport_problem.py:632-670``)."""
        p = self.problem
        mesh = p.mesh.mesh
        d = mesh.tdim
        facets = p.facet_meshtags.find(surf_id)
        fverts = mesh.bfacets[facets]
        fcells = mesh.bfacet_cell[facets]
        _, G, x = cell_geometry(mesh, fcells)
        ns = self.assembler.ns
        ucell = p.u.x.array.reshape(-1, ns)[:, species_index][mesh.cells[fcells]]
        gradu = np.einsum("ea,ead->ed", ucell, G)
        # outward normal: minus the (normalised) gradient of the basis function of
        # the vertex opposite to the facet
        cverts = mesh.cells[fcells]
        opp = np.array([[v not in fv for v in cv] for cv, fv in zip(cverts, fverts)])
        gopp = G[opp]
        n = -gopp / np.linalg.norm(gopp, axis=1)[:, None]
        D0, ED = p.diffusivity_tables()
        vol_ids = [v.id for v in p.volume_subdomains]
        ctag = np.zeros(mesh.num_cells, dtype=np.int64)
        ctag[p.volume_meshtags.indices] = p.volume_meshtags.values
        iv = np.array([vol_ids.index(t) for t in ctag[fcells]])
        T = self._D_T
        Tv = T[fverts] if isinstance(T, np.ndarray) else np.full(fverts.shape, float(T))
        Dh = D0[iv, species_index][:, None] * np.exp(-ED[iv, species_index][:, None] / k_B / Tv)
        area = facet_measures(mesh, fverts)
        return float(np.sum(area * Dh.mean(axis=1) * (-np.einsum("ed,ed->e", gradu, n))))


class OracleBackend:
    """Plugs the oracle in at the ``create_solver`` seam (tests only)."""

    def create_newton_solver(self, problem, petsc_options):
        solver = OracleNewtonSolver(problem, petsc_options)
        if getattr(problem, "_D_global_T", None) is not None:
            solver.set_D_global_temperature(problem._D_global_T)
        return solver
