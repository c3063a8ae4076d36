"""Stand-ins for the few ``dolfinx.fem`` types the FESTIM API exposes.

The reference branches on ``isinstance(value, fem.Constant | fem.Function)``
in many places (e.g. ``src/festim/hydrogen_transport_problem.py:383``,
``src/festim/helpers.py:268-276``); user scripts read ``u.x.array``.  These
classes keep those names and that behaviour on top of plain numpy arrays: a
``Function`` is a P1 nodal field (one value per mesh vertex, or ``bs`` values
for a blocked space), a ``Constant`` is a mutable scalar.
"""

from __future__ import annotations

import numpy as np

from festim_b200 import ufl as _ufl


class Constant(_ufl.Expr):
    """Mutable scalar; participates in expressions by reference."""

    def __init__(self, mesh, value):
        self.mesh = mesh
        self._value = np.float64(value)

    @property
    def value(self):
        return self._value

    @value.setter
    def value(self, v):
        self._value = np.float64(v)

    def __float__(self):
        return float(self._value)

    def _as_expr(self):
        return _ufl.ConstantRef(self)

    def __repr__(self):
        return f"Constant({float(self._value)})"


class _Vec:
    """``Function.x``: holds ``.array`` like ``dolfinx.la.Vector``.

    A solver backend that keeps the field on its device may ``defer`` the copy back: the host
    array is then refreshed the first time somebody looks at it.  ``version`` counts the host
    accesses (each may modify the array in place), so the backend knows whether its device copy
    is still the current one."""

    def __init__(self, n):
        self._array = np.zeros(n, dtype=np.float64)
        self._fetch = None
        self.version = 0

    @property
    def array(self):
        if self._fetch is not None:
            fetch, self._fetch = self._fetch, None
            fetch(self._array)
        self.version += 1
        return self._array

    @array.setter
    def array(self, value):
        self._array = np.ascontiguousarray(value, dtype=np.float64)
        self._fetch = None
        self.version += 1

    def defer(self, fetch):
        """The host copy is stale: ``fetch(out)`` fills it when it is next read."""
        self._fetch = fetch

    def scatter_forward(self):
        pass


class FunctionSpace:
    """P1 (or DG0/DG1 placeholder) space on a SimplexMesh, block size ``bs``."""

    def __init__(self, mesh, family="Lagrange", degree=1, bs=1, parent=None, sub_index=None):
        self.mesh = mesh
        self.family = family
        self.degree = degree
        self.bs = bs
        self.parent = parent
        self.sub_index = sub_index

    @property
    def num_dofs(self):
        if self.family == "DG" and self.degree == 0:
            return self.mesh.num_cells * self.bs
        if self.family == "DG":      # DG1: one value per (cell, local vertex), cell-major
            return self.mesh.num_cells * (self.mesh.tdim + 1) * self.bs
        return self.mesh.num_vertices * self.bs

    def sub(self, i):
        return FunctionSpace(self.mesh, self.family, self.degree, 1, parent=self, sub_index=i)

    def collapse(self):
        V = FunctionSpace(self.mesh, self.family, self.degree, 1)
        if self.parent is None:
            return V, np.arange(self.num_dofs, dtype=np.int32)
        dofs = np.arange(self.mesh.num_vertices, dtype=np.int32) * self.parent.bs + self.sub_index
        return V, dofs

    def tabulate_dof_coordinates(self):
        x = np.zeros((self.mesh.num_vertices, 3))
        x[:, : self.mesh.gdim] = self.mesh.coords
        return x


def functionspace(mesh, element=("Lagrange", 1)):
    mesh = getattr(mesh, "mesh", mesh) if not hasattr(mesh, "coords") else mesh
    family, degree = element[0], element[1]
    if family in ("P", "CG"):
        family = "Lagrange"
    return FunctionSpace(mesh, family, degree)


class Function(_ufl.Expr):
    """P1 nodal field.  ``x.array[v * bs + s]`` (node-major blocked layout)."""

    def __init__(self, V, name=""):
        self.function_space = V
        self.x = _Vec(V.num_dofs)
        self.name = name

    @property
    def gdim(self):
        return self.function_space.mesh.gdim

    def _as_expr(self):
        return _ufl.FieldRef(self)

    def __bool__(self):
        # a field object is truthy, as a dolfinx Function is (`if material.D:`); only
        # expressions built from it have no truth value
        return True

    def interpolate(self, f, cells0=None):
        """Nodal interpolation of a callable ``f(x)`` (x of shape (3, n), the
        dolfinx convention), an expression, or another Function."""
        V = self.function_space
        mesh = V.mesh
        if cells0 is None:
            verts = np.arange(mesh.num_vertices)
        else:
            verts = np.unique(mesh.cells[np.asarray(cells0, dtype=np.int64)].ravel())
        if isinstance(f, Function):
            vals = f.x.array.reshape(mesh.num_vertices, -1)[verts, 0]
        elif isinstance(f, Expression):
            vals = f.eval_at_vertices(mesh, verts)
        elif isinstance(f, (_ufl.Expr, int, float)):
            vals = Expression(f).eval_at_vertices(mesh, verts)
        elif callable(f):
            x3 = np.zeros((3, len(verts)))
            x3[: mesh.gdim] = mesh.coords[verts].T
            vals = np.asarray(f(x3), dtype=np.float64)
            if vals.ndim == 0:
                vals = np.full(len(verts), float(vals))
        else:
            raise TypeError(f"cannot interpolate {type(f)}")
        arr = self.x.array.reshape(mesh.num_vertices, -1)
        arr[verts, 0] = vals

    def sub(self, i):
        return _SubFunction(self, i)

    def copy(self):
        g = Function(self.function_space, self.name)
        g.x.array[:] = self.x.array
        return g

    def __repr__(self):
        return f"Function({self.name})"


class _SubFunction:
    """View of one species of a blocked Function (``u.sub(i)``)."""

    def __init__(self, parent, i):
        self.parent, self.i = parent, i

    def collapse(self):
        V = self.parent.function_space
        g = Function(FunctionSpace(V.mesh, V.family, V.degree, 1))
        g.x.array[:] = self.parent.x.array.reshape(-1, V.bs)[:, self.i]
        return g

    def interpolate(self, f, cells0=None):
        V = self.parent.function_space
        tmp = Function(FunctionSpace(V.mesh, V.family, V.degree, 1))
        tmp.x.array[:] = self.parent.x.array.reshape(-1, V.bs)[:, self.i]
        tmp.interpolate(f, cells0=cells0)
        self.parent.x.array.reshape(-1, V.bs)[:, self.i] = tmp.x.array


class Expression:
    """``fem.Expression(expr, points)``: an expression to evaluate at nodes."""

    def __init__(self, expr, X=None):
        self.expr = _ufl.as_expr(expr)

    def eval_at_vertices(self, mesh, verts=None):
        verts = np.arange(mesh.num_vertices) if verts is None else verts
        env = nodal_env(mesh, verts)
        vals = _ufl.evaluate(self.expr, env)
        return np.broadcast_to(np.asarray(vals, dtype=np.float64), (len(verts),)).copy()


def nodal_env(mesh, verts):
    """Evaluation environment at mesh vertices (P1 fields read at the node)."""
    pts = mesh.coords[verts]

    def field(fn):
        return fn.x.array.reshape(mesh.num_vertices, -1)[verts, 0]

    def field_grad(fn, i):
        raise NotImplementedError("gradient of a P1 field is not defined at nodes")

    def species(spe):
        raise NotImplementedError("species-dependent values cannot be interpolated")

    return {"x": pts, "field": field, "field_grad": field_grad, "species": species}


# ---------------------------------------------------------------------------------------------
# point location and interpolation between non-matching meshes (reference helpers.py:345-371
# ``nmm_interpolate`` -> dolfinx ``create_interpolation_data`` / ``interpolate_nonmatching``)
# ---------------------------------------------------------------------------------------------
class PointLocator:
    """Cell and barycentric coordinates of points in a P1 simplex mesh.

    Candidates are the cells around the nearest mesh vertices (k-d tree); the cell whose smallest
    barycentric coordinate is largest wins, so points on facets / vertices resolve consistently and
    points a round-off outside the mesh take the linear extension of the closest cell, which is
    what dolfinx's ``padding`` achieves."""

    def __init__(self, mesh, k=4):
        from scipy.spatial import cKDTree

        self.mesh = mesh
        self.k = min(k, mesh.num_vertices)
        self.tree = cKDTree(mesh.coords)
        x = mesh.coords[mesh.cells]                       # (nc, d+1, d)
        self.x0 = x[:, 0, :]
        self.Jinv = np.linalg.inv(np.transpose(x[:, 1:, :] - x[:, :1, :], (0, 2, 1)))   # (nc, d, d)

    def _bary(self, cells, pts):
        lam = np.einsum("nij,nj->ni", self.Jinv[cells], pts - self.x0[cells])
        return np.concatenate([1.0 - lam.sum(axis=1, keepdims=True), lam], axis=1)

    def locate(self, pts):
        pts = np.atleast_2d(np.asarray(pts, dtype=np.float64))[:, : self.mesh.gdim]
        n = pts.shape[0]
        ptr, idx = self.mesh.vertex_to_cell
        d1 = self.mesh.tdim + 1
        _, near = self.tree.query(pts, k=self.k)
        near = near.reshape(n, -1)
        best_cell = np.full(n, -1, dtype=np.int64)
        best_min = np.full(n, -np.inf)
        best_bary = np.zeros((n, d1))
        for col in range(near.shape[1]):
            v = near[:, col]
            counts = ptr[v + 1] - ptr[v]
            for j in range(int(counts.max())):
                has = np.nonzero(j < counts)[0]
                cells = idx[ptr[v[has]] + j] // d1
                bary = self._bary(cells, pts[has])
                score = bary.min(axis=1)
                better = score > best_min[has]
                sel = has[better]
                best_cell[sel], best_min[sel], best_bary[sel] = cells[better], score[better], bary[better]
            if np.all(best_min > -1e-12):
                break
        return best_cell, best_bary, best_min


def interpolate_nonmatching(f_out, f_in, padding=1e-11, locator=None):
    """``f_out`` <- P1 interpolant of ``f_in`` at the vertices of ``f_out``'s mesh.  Points a
    round-off outside ``f_in``'s mesh take the linear extension of the closest cell (the role of
    dolfinx's ``padding``, accepted for signature compatibility); points farther out raise.
    Returns the locator so that repeated transfers between the same meshes reuse it."""
    mesh_in, mesh_out = f_in.function_space.mesh, f_out.function_space.mesh
    locator = locator or PointLocator(mesh_in)
    cells, bary, worst = locator.locate(mesh_out.coords)
    if np.any(worst < -1e-6):      # more than a millionth of a cell outside (barycentric units)
        raise ValueError("interpolate_nonmatching: some points of the target mesh lie outside the source mesh")
    vals = f_in.x.array.reshape(mesh_in.num_vertices, -1)[:, 0]
    f_out.x.array.reshape(mesh_out.num_vertices, -1)[:, 0] = np.einsum("na,na->n", vals[mesh_in.cells[cells]], bary)
    return locator
