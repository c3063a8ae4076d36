"""Numpy restatement of dolfinx/FFCx assembly of the FESTIM residual form.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Each integral of the form
(``festim_b200.forms``; built by the mirror of reference
``hydrogen_transport_problem.py:870-997``) is evaluated the way an FFCx
``tabulate_tensor`` kernel does: loop over quadrature points of the rule of the
estimated degree, evaluate the integrand there, multiply by test (and trial)
basis values.  The Jacobian uses the symbolic derivative of the integrand
(``ufl.derivative`` analogue).  Strong Dirichlet conditions follow dolfinx's
``NonlinearProblem``: lifting with ``alpha=-1``, ``set_bc`` on the residual,
identity rows/columns in the matrix.
"""

from __future__ import annotations

import math

import numpy as np
import scipy.sparse as sp

from festim_b200 import quadrature, ufl


def cell_geometry(mesh, cells):
    """volumes (nc,), basis gradients (nc, d+1, d), vertex coords (nc, d+1, d)."""
    d = mesh.tdim
    x = mesh.coords[mesh.cells[cells]]
    J = x[:, 1:, :] - x[:, :1, :]                  # rows = edge vectors
    det = np.linalg.det(J)
    Jinv = np.linalg.inv(J)
    G = np.empty((len(cells), d + 1, d))
    G[:, 1:, :] = np.transpose(Jinv, (0, 2, 1))    # grad(phi_{k+1}) = column k of J^-1
    G[:, 0, :] = -G[:, 1:, :].sum(axis=1)
    return np.abs(det) / math.factorial(d), G, x


class FormAssembler:
    def __init__(self, mesh, form, species, volume_meshtags, facet_meshtags):
        self.mesh = mesh
        self.form = form
        self.species = list(species)
        self.ns = len(self.species)
        self.vt = volume_meshtags
        self.ft = facet_meshtags
        self.N = mesh.num_vertices * self.ns
        self.degrees = form.quadrature_degrees()

    # -- evaluation environment at points given by barycentric coords ------------
    def _env(self, verts, bary, G, u, u_n, normal=None, circumradius=None):
        """verts (ne, k) vertex ids of each entity, bary (nq, k)."""
        mesh, ns = self.mesh, self.ns
        xq = np.einsum("qa,ead->eqd", bary, mesh.coords[verts])

        def interp(nodal):
            return np.einsum("qa,ea->eq", bary, nodal[verts])

        def field(fn):
            return interp(fn.x.array.reshape(mesh.num_vertices, -1)[:, 0])

        def field_grad(fn, i):
            nodal = fn.x.array.reshape(mesh.num_vertices, -1)[:, 0]
            return np.einsum("ea,ea->e", nodal[self._cell_verts], G[:, :, i])[:, None]

        def species(spe):
            return interp(u.reshape(-1, ns)[:, self.species.index(spe)])

        def species_prev(spe):
            return interp(u_n.reshape(-1, ns)[:, self.species.index(spe)])

        def species_grad(spe, i):
            # P1: constant on the owning cell (metric terms of curvilinear coordinates)
            nodal = u.reshape(-1, ns)[:, self.species.index(spe)]
            return np.einsum("ea,ea->e", nodal[self._cell_verts], G[:, :, i])[:, None]

        return {"x": xq, "field": field, "field_grad": field_grad,
                "species": species, "species_prev": species_prev, "species_grad": species_grad,
                "normal": (lambda i: normal[:, i][:, None]),
                "circumradius": (lambda: circumradius[:, None])}

    def assemble(self, u, u_n, want_J=True):
        """Returns (F, J) without boundary conditions; J is CSR or None."""
        mesh, ns, d = self.mesh, self.ns, self.mesh.tdim
        F = np.zeros(self.N)
        rows, cols, vals = [], [], []

        def add_F(verts_i, s, contrib):
            np.add.at(F, verts_i.astype(np.int64) * ns + s, contrib)

        def add_J(verts_i, si, verts_j, sj, contrib):
            rows.append(verts_i.astype(np.int64) * ns + si)
            cols.append(verts_j.astype(np.int64) * ns + sj)
            vals.append(contrib)

        for (kind, tag), itgs in self.form.groups().items():
            degF, degJ = self.degrees[(kind, tag)]
            if kind == "dx":
                cells = self.vt.find(tag)
                if len(cells) == 0:
                    continue
                vol, G, _ = cell_geometry(mesh, cells)
                verts = mesh.cells[cells]
                self._cell_verts = verts
                for which, deg in (("F", degF), ("J", degJ)):
                    if which == "J" and (not want_J or degJ is None):
                        continue
                    bary, w = quadrature.rule(d, deg)
                    env = self._env(verts, bary, G, u, u_n)
                    for itg in itgs:
                        s = self.species.index(itg.species)
                        us = u.reshape(-1, ns)[:, s][verts]          # (nc, d+1)
                        if itg.kind == "diffusion":
                            Dq = np.broadcast_to(ufl.evaluate(itg.D, env), (len(cells), len(w)))
                            Dbar = Dq @ w
                            gradu = np.einsum("ea,ead->ed", us, G)
                            if which == "F":
                                for i in range(d + 1):
                                    add_F(verts[:, i], s, vol * Dbar * np.einsum("ed,ed->e", G[:, i], gradu))
                            else:
                                for i in range(d + 1):
                                    for j in range(d + 1):
                                        add_J(verts[:, i], s, verts[:, j], s,
                                              vol * Dbar * np.einsum("ed,ed->e", G[:, i], G[:, j]))
                                for spe2 in ufl.species_in(itg.D):
                                    s2 = self.species.index(spe2)
                                    dD = np.broadcast_to(
                                        ufl.evaluate(ufl.diff(itg.D, ("species", spe2)), env),
                                        (len(cells), len(w)))
                                    for i in range(d + 1):
                                        gi = np.einsum("ed,ed->e", G[:, i], gradu)
                                        for j in range(d + 1):
                                            add_J(verts[:, i], s, verts[:, j], s2,
                                                  vol * gi * ((dD * bary[:, j]) @ w))
                        elif itg.kind == "test":
                            if which == "F":
                                vq = np.broadcast_to(ufl.evaluate(itg.expr, env), (len(cells), len(w)))
                                for i in range(d + 1):
                                    add_F(verts[:, i], s, vol * ((vq * bary[:, i]) @ w))
                            else:
                                for spe2, dex in itg.derivatives():
                                    s2 = self.species.index(spe2)
                                    dq = np.broadcast_to(ufl.evaluate(dex, env), (len(cells), len(w)))
                                    for i in range(d + 1):
                                        for j in range(d + 1):
                                            add_J(verts[:, i], s, verts[:, j], s2,
                                                  vol * ((dq * bary[:, i] * bary[:, j]) @ w))
                                for (spe2, comp), dex in itg.grad_derivatives():
                                    s2 = self.species.index(spe2)
                                    dq = np.broadcast_to(ufl.evaluate(dex, env), (len(cells), len(w)))
                                    for i in range(d + 1):
                                        for j in range(d + 1):   # d_comp phi_j is constant on the cell
                                            add_J(verts[:, i], s, verts[:, j], s2,
                                                  vol * G[:, j, comp] * ((dq * bary[:, i]) @ w))
                        elif itg.kind == "advection":
                            velq = np.einsum("qa,ead->eqd", bary, itg.velocity.array[verts])
                            if which == "F":
                                gradu = np.einsum("ea,ead->ed", us, G)
                                adv = np.einsum("eqd,ed->eq", velq, gradu)
                                for i in range(d + 1):
                                    add_F(verts[:, i], s, vol * ((adv * bary[:, i]) @ w))
                            else:
                                for j in range(d + 1):
                                    vg = np.einsum("eqd,ed->eq", velq, G[:, j])
                                    for i in range(d + 1):
                                        add_J(verts[:, i], s, verts[:, j], s,
                                              vol * ((vg * bary[:, i]) @ w))
            else:  # ds
                facets = self.ft.find(tag)
                if len(facets) == 0:
                    continue
                fverts = mesh.bfacets[facets]                       # (nf, d)
                fcells = mesh.bfacet_cell[facets]
                _, G, _ = cell_geometry(mesh, fcells)
                cverts = mesh.cells[fcells]                          # (nf, d+1)
                self._cell_verts = cverts
                area = facet_measures(mesh, fverts)
                # outward unit normal: minus the direction of grad(phi) of the vertex opposite
                # to the facet; n . grad(phi_i) for every vertex of the cell; circumradius
                opp = np.array([[v not in fv for v in cv] for cv, fv in zip(cverts, fverts)])
                gopp = G[opp]
                normal = -gopp / np.linalg.norm(gopp, axis=1)[:, None]
                ng = np.einsum("ed,ead->ea", normal, G)               # (nf, d+1)
                hcell = circumradius(mesh, fcells)
                for which, deg in (("F", degF), ("J", degJ)):
                    if which == "J" and (not want_J or degJ is None):
                        continue
                    bary, w = quadrature.rule(d - 1, deg)
                    env = self._env(fverts, bary, G, u, u_n, normal, hcell)
                    for itg in itgs:
                        assert itg.kind in ("test", "ngradtest")
                        s = self.species.index(itg.species)
                        gradtest = itg.kind == "ngradtest"
                        # rows: facet vertices weighted by phi_i(q), or cell vertices weighted by
                        # the constant n . grad(phi_i)
                        if gradtest:
                            rverts = cverts
                            rw = [np.repeat(ng[:, i][:, None], len(w), axis=1) for i in range(d + 1)]
                        else:
                            rverts = fverts
                            rw = [np.repeat(bary[:, i][None, :], len(facets), axis=0) for i in range(d)]
                        if which == "F":
                            vq = np.broadcast_to(ufl.evaluate(itg.expr, env), (len(facets), len(w)))
                            for i in range(rverts.shape[1]):
                                add_F(rverts[:, i], s, area * ((vq * rw[i]) @ w))
                        else:
                            colsets = [(self.species.index(spe2), dex, fverts,
                                     [np.repeat(bary[:, j][None, :], len(facets), axis=0) for j in range(d)])
                                    for spe2, dex in itg.derivatives()]
                            colsets += [(self.species.index(spe2), dex, cverts,
                                      [np.repeat(G[:, j, comp][:, None], len(w), axis=1) for j in range(d + 1)])
                                     for (spe2, comp), dex in itg.grad_derivatives()]
                            for s2, dex, cv, cw in colsets:
                                dq = np.broadcast_to(ufl.evaluate(dex, env), (len(facets), len(w)))
                                for i in range(rverts.shape[1]):
                                    for j in range(cv.shape[1]):
                                        add_J(rverts[:, i], s, cv[:, j], s2, area * ((dq * rw[i] * cw[j]) @ w))
        J = None
        if want_J:
            if rows:
                J = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                                  shape=(self.N, self.N)).tocsr()
            else:
                J = sp.csr_matrix((self.N, self.N))
            J.sum_duplicates()
        return F, J


def circumradius(mesh, cells):
    """Circumradius of affine simplices (``ufl.Circumradius``)."""
    x = mesh.coords[mesh.cells[cells]]
    d = mesh.tdim
    dist = lambda p, q: np.linalg.norm(x[:, p] - x[:, q], axis=1)
    if d == 1:
        return 0.5 * dist(0, 1)
    if d == 2:
        e1, e2 = x[:, 1] - x[:, 0], x[:, 2] - x[:, 0]
        area = 0.5 * np.abs(e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0])
        return dist(1, 2) * dist(0, 2) * dist(0, 1) / (4.0 * area)
    aA, bB, cC = dist(0, 1) * dist(2, 3), dist(0, 2) * dist(1, 3), dist(0, 3) * dist(1, 2)
    vol = np.abs(np.linalg.det(x[:, 1:] - x[:, :1])) / 6.0
    return np.sqrt((aA + bB + cC) * (aA + bB - cC) * (aA - bB + cC) * (-aA + bB + cC)) / (24.0 * vol)


def facet_measures(mesh, fverts):
    d = mesh.tdim
    x = mesh.coords[fverts]
    if d == 1:
        return np.ones(len(fverts))
    if d == 2:
        return np.linalg.norm(x[:, 1] - x[:, 0], axis=1)
    return 0.5 * np.linalg.norm(np.cross(x[:, 1] - x[:, 0], x[:, 2] - x[:, 0]), axis=1)


def apply_dirichlet(F, J, x, bc_dofs, bc_vals):
    """dolfinx NonlinearProblem semantics: returns (b, J_bc)."""
    N = len(F)
    b = F.copy()
    if len(bc_dofs) == 0:
        return b, J
    delta = np.zeros(N)
    delta[bc_dofs] = bc_vals - x[bc_dofs]
    if J is not None:
        b += J @ delta                                   # apply_lifting, alpha = -1
    b[bc_dofs] = x[bc_dofs] - bc_vals                     # set_bc, alpha = -1
    Jbc = None
    if J is not None:
        free = np.ones(N)
        free[bc_dofs] = 0.0
        Dfree = sp.diags(free)
        Jbc = (Dfree @ J @ Dfree + sp.diags(1.0 - free)).tocsr()
    return b, Jbc
