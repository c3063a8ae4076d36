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
    """``Function.x``: holds ``.array`` like ``dolfinx.la.Vector``."""

    def __init__(self, n):
        self.array = np.zeros(n, dtype=np.float64)

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
