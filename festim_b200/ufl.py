"""Minimal UFL-compatible expression language.

The reference lets users pass callables of ``x``, ``t``, ``T`` (and species
concentrations) that return UFL expressions (``ufl.exp``, ``ufl.sin``,
``ufl.conditional``, ``ufl.div(ufl.grad(...))``), see
reference ``src/festim/helpers.py:41-84`` and
``test/system_tests/test_1_mobile_species.py:60``.  UFL is not installed in
this image, so this module provides a small closed-form expression tree with
the same surface: arithmetic, math functions, conditionals, symbolic
``grad``/``div`` in the spatial coordinates, numpy evaluation at arbitrary
points, compilation to the byte-code the CUDA element kernels interpret
(``festim_b200/csrc/program.cuh``) and the UFL polynomial-degree estimate
(sum-of-degrees heuristic) that picks the quadrature rule.
"""

from __future__ import annotations

import math

import numpy as np

pi = math.pi
e = math.e


# --------------------------------------------------------------------------
# expression nodes
# --------------------------------------------------------------------------
class Expr:
    """Base class of scalar expression nodes."""

    __array_priority__ = 1000
    __array_ufunc__ = None  # numpy scalars defer to our reflected operators

    # arithmetic (as_expr(self): Constant/Function subclasses enter the tree
    # by reference, through ConstantRef / FieldRef nodes) ---------------------
    def __add__(self, o):
        return _sum(as_expr(self), as_expr(o))

    def __radd__(self, o):
        return _sum(as_expr(o), as_expr(self))

    def __sub__(self, o):
        return _sum(as_expr(self), _neg(as_expr(o)))

    def __rsub__(self, o):
        return _sum(as_expr(o), _neg(as_expr(self)))

    def __mul__(self, o):
        if isinstance(o, Vector):
            return o.__rmul__(self)
        return _prod(as_expr(self), as_expr(o))

    def __rmul__(self, o):
        return _prod(as_expr(o), as_expr(self))

    def __truediv__(self, o):
        return _div(as_expr(self), as_expr(o))

    def __rtruediv__(self, o):
        return _div(as_expr(o), as_expr(self))

    def __pow__(self, o):
        return _pow(as_expr(self), as_expr(o))

    def __rpow__(self, o):
        return _pow(as_expr(o), as_expr(self))

    def __neg__(self):
        return _neg(as_expr(self))

    def __pos__(self):
        return as_expr(self)

    def __abs__(self):
        return _math("abs", self)

    # comparisons build conditions (as in UFL, usable inside conditional())
    def __lt__(self, o):
        return Condition("lt", as_expr(self), as_expr(o))

    def __gt__(self, o):
        return Condition("gt", as_expr(self), as_expr(o))

    def __le__(self, o):
        return Condition("le", as_expr(self), as_expr(o))

    def __ge__(self, o):
        return Condition("ge", as_expr(self), as_expr(o))

    def __bool__(self):
        raise TypeError("a symbolic expression has no truth value; use conditional()")

    # `==` is structural equality and returns a bool, as in UFL (ufl/exprequals.py); the
    # condition `a == b` for conditional() is `eq(a, b)`
    def __eq__(self, o):
        return expr_equals(self, o)

    def __ne__(self, o):
        return not expr_equals(self, o)

    __hash__ = object.__hash__

    def __float__(self):
        v = evaluate(self, {})
        return float(v)


def expr_equals(a, b):
    """Same tree: same node types and operators, leaves (fields, constants, species) by identity,
    literals by value."""
    if a is b:
        return True
    try:
        a, b = as_expr(a), as_expr(b)
    except Exception:
        return False
    return _same(a, b)


def _same(a, b):
    if a is b:
        return True
    if type(a) is not type(b):
        return False
    if hasattr(a, "_as_expr"):   # fem.Function / fem.Constant held by a reference node
        return False
    if isinstance(a, (Expr, Condition)):
        va, vb = vars(a), vars(b)
        return va.keys() == vb.keys() and all(_same(va[k], vb[k]) for k in va)
    if isinstance(a, (bool, int, float, str)):
        return a == b
    return False   # any other leaf object (Function, Constant, Species): identity only


class Const(Expr):
    def __init__(self, value):
        self.value = float(value)

    def __array__(self, dtype=None, copy=None):
        # a literal converts to a number for numpy (np.isclose(value, expected) in user code),
        # like ufl's FloatValue
        return np.asarray(self.value, dtype=dtype)

    def __repr__(self):
        return repr(self.value)


# arithmetic of a literal with numpy arrays is plain numpy arithmetic (post-processing code
# mixes folded constants such as exp(-E / k_B / 500.0) with arrays of times)
def _const_numpy_ops():
    import operator

    def make(name, op, reflected):
        base = getattr(Expr, name)

        def method(self, o):
            if isinstance(o, np.ndarray):
                return op(o, self.value) if reflected else op(self.value, o)
            return base(self, o)

        return method

    for name, op in (("add", operator.add), ("sub", operator.sub), ("mul", operator.mul),
                     ("truediv", operator.truediv), ("pow", operator.pow)):
        setattr(Const, f"__{name}__", make(f"__{name}__", op, False))
        setattr(Const, f"__r{name}__", make(f"__r{name}__", op, True))


_const_numpy_ops()


class Coord(Expr):
    """Component ``i`` of the spatial coordinate."""

    def __init__(self, i):
        self.i = int(i)

    def __repr__(self):
        return f"x[{self.i}]"


class ConstantRef(Expr):
    """Reference to a mutable scalar (a ``fem.Constant``): t, dt, T(t), ..."""

    def __init__(self, constant):
        self.constant = constant

    def __repr__(self):
        return f"Constant({self.constant.value})"


class FieldRef(Expr):
    """Reference to a P1 nodal field (a ``fem.Function``)."""

    def __init__(self, function):
        self.function = function

    def __repr__(self):
        return f"Function({getattr(self.function, 'name', '')})"


class FieldGrad(Expr):
    """Component ``i`` of the (cell-wise constant) gradient of a P1 field."""

    def __init__(self, function, i):
        self.function = function
        self.i = int(i)

    def __repr__(self):
        return f"grad({getattr(self.function, 'name', '')})[{self.i}]"


class SpeciesRef(Expr):
    """Concentration of an unknown species at the evaluation point."""

    def __init__(self, species, prev=False):
        self.species = species
        self.prev = bool(prev)  # True: value at the previous time step (u_n)

    def __repr__(self):
        return f"c{'_n' if self.prev else ''}[{getattr(self.species, 'name', None)}]"


class SpeciesGrad(Expr):
    """Component ``i`` of the gradient of an unknown species (cell-wise constant for P1):
    what ``ufl.grad(u)[i]`` is inside an integrand.  Used by the metric terms of cylindrical
    and spherical coordinates (reference hydrogen_transport_problem.py:886-908)."""

    def __init__(self, species, i):
        self.species = species
        self.i = int(i)

    def __repr__(self):
        return f"d{self.i}(c[{getattr(self.species, 'name', None)}])"


class NormalComp(Expr):
    """Component ``i`` of the outward unit normal of the boundary facet being integrated
    (``ufl.FacetNormal(mesh)[i]``)."""

    def __init__(self, i):
        self.i = int(i)

    def __repr__(self):
        return f"n[{self.i}]"


class CellCircumradius(Expr):
    """Circumradius of the cell owning the integration entity (``ufl.Circumradius(mesh)``)."""

    def __repr__(self):
        return "h"


class Sum(Expr):
    def __init__(self, a, b):
        self.a, self.b = a, b

    def __repr__(self):
        return f"({self.a!r} + {self.b!r})"


class Product(Expr):
    def __init__(self, a, b):
        self.a, self.b = a, b

    def __repr__(self):
        return f"({self.a!r} * {self.b!r})"


class Division(Expr):
    def __init__(self, a, b):
        self.a, self.b = a, b

    def __repr__(self):
        return f"({self.a!r} / {self.b!r})"


class Power(Expr):
    def __init__(self, a, b):
        self.a, self.b = a, b

    def __repr__(self):
        return f"({self.a!r} ** {self.b!r})"


class Neg(Expr):
    def __init__(self, a):
        self.a = a

    def __repr__(self):
        return f"(-{self.a!r})"


class Math(Expr):
    FUNCS = ("exp", "ln", "sqrt", "sin", "cos", "tan", "tanh", "sinh", "cosh",
             "abs", "erf", "atan", "asin", "acos", "sign")

    def __init__(self, name, a):
        assert name in self.FUNCS, name
        self.name, self.a = name, a

    def __repr__(self):
        return f"{self.name}({self.a!r})"


class Condition:
    """Boolean node, only valid as the first argument of ``conditional``."""

    def __init__(self, op, a, b=None):
        self.op, self.a, self.b = op, a, b

    def __repr__(self):
        return f"{self.op}({self.a!r}, {self.b!r})"


class Conditional(Expr):
    def __init__(self, cond, a, b):
        self.cond, self.a, self.b = cond, a, b

    def __repr__(self):
        return f"conditional({self.cond!r}, {self.a!r}, {self.b!r})"


class MinMax(Expr):
    def __init__(self, kind, a, b):
        self.kind, self.a, self.b = kind, a, b

    def __repr__(self):
        return f"{self.kind}({self.a!r}, {self.b!r})"


class Vector:
    """Fixed-length list of scalar expressions (``as_vector`` / ``grad`` result)."""

    __array_ufunc__ = None

    def __init__(self, comps):
        self.comps = [as_expr(c) for c in comps]

    def __len__(self):
        return len(self.comps)

    def __getitem__(self, i):
        return self.comps[i]

    def __iter__(self):
        return iter(self.comps)

    def __mul__(self, o):
        return Vector([c * o for c in self.comps])

    def __rmul__(self, o):
        return Vector([o * c for c in self.comps])

    def __truediv__(self, o):
        return Vector([c / o for c in self.comps])

    def __neg__(self):
        return Vector([-c for c in self.comps])

    def __add__(self, o):
        return Vector([a + b for a, b in zip(self.comps, o.comps)])

    def __sub__(self, o):
        return Vector([a - b for a, b in zip(self.comps, o.comps)])

    def __repr__(self):
        return f"Vector({self.comps!r})"


class SpatialCoordinate(Vector):
    """``ufl.SpatialCoordinate(mesh)``; accepts a mesh object or a dimension."""

    def __init__(self, mesh_or_dim=3):
        gdim = mesh_or_dim if isinstance(mesh_or_dim, int) else _gdim_of(mesh_or_dim)
        super().__init__([Coord(i) for i in range(gdim)])
        self.gdim = gdim
        set_default_gdim(gdim)


def _gdim_of(mesh):
    for attr in ("gdim", "dim"):
        if hasattr(mesh, attr):
            return int(getattr(mesh, attr))
    if hasattr(mesh, "mesh"):
        return _gdim_of(mesh.mesh)
    raise TypeError(f"cannot infer the geometric dimension of {mesh!r}")


# --------------------------------------------------------------------------
# constructors with light constant folding (keeps programs short; does not
# change the degree estimate except for exact zeros / ones, which UFL also
# simplifies: ufl/algebra.py Sum/Product __new__)
# --------------------------------------------------------------------------
def as_expr(v):
    if hasattr(v, "_as_expr"):
        return v._as_expr()
    if isinstance(v, Expr):
        return v
    if isinstance(v, (bool, np.bool_)):
        raise TypeError("booleans are not expressions")
    if isinstance(v, (int, float, np.integer, np.floating)):
        return Const(v)
    raise TypeError(f"cannot convert {type(v)} to an expression")


def _is_const(x, val=None):
    return isinstance(x, Const) and (val is None or x.value == val)


def _sum(a, b):
    if _is_const(a) and _is_const(b):
        return Const(a.value + b.value)
    if _is_const(a, 0.0):
        return b
    if _is_const(b, 0.0):
        return a
    return Sum(a, b)


def _prod(a, b):
    if _is_const(a) and _is_const(b):
        return Const(a.value * b.value)
    if _is_const(a, 0.0) or _is_const(b, 0.0):
        return Const(0.0)
    if _is_const(a, 1.0):
        return b
    if _is_const(b, 1.0):
        return a
    return Product(a, b)


def _div(a, b):
    if _is_const(a) and _is_const(b):
        return Const(a.value / b.value)
    if _is_const(a, 0.0):
        return Const(0.0)
    if _is_const(b, 1.0):
        return a
    return Division(a, b)


def _pow(a, b):
    if _is_const(a) and _is_const(b):
        return Const(a.value ** b.value)
    if _is_const(b, 1.0):
        return a
    if _is_const(b, 0.0):
        return Const(1.0)
    return Power(a, b)


def _neg(a):
    if _is_const(a):
        return Const(-a.value)
    if isinstance(a, Neg):
        return a.a
    return Neg(a)


def _math(name, a):
    a = as_expr(a)
    if _is_const(a):
        return Const(float(_NP_FUNCS[name](a.value)))
    return Math(name, a)


def exp(a):
    return _math("exp", a)


def ln(a):
    return _math("ln", a)


def sqrt(a):
    return _math("sqrt", a)


def sin(a):
    return _math("sin", a)


def cos(a):
    return _math("cos", a)


def tan(a):
    return _math("tan", a)


def tanh(a):
    return _math("tanh", a)


def sinh(a):
    return _math("sinh", a)


def cosh(a):
    return _math("cosh", a)


def erf(a):
    return _math("erf", a)


def atan(a):
    return _math("atan", a)


def asin(a):
    return _math("asin", a)


def acos(a):
    return _math("acos", a)


def sign(a):
    return _math("sign", a)


def lt(a, b):
    return Condition("lt", as_expr(a), as_expr(b))


def gt(a, b):
    return Condition("gt", as_expr(a), as_expr(b))


def le(a, b):
    return Condition("le", as_expr(a), as_expr(b))


def ge(a, b):
    return Condition("ge", as_expr(a), as_expr(b))


def eq(a, b):
    return Condition("eq", as_expr(a), as_expr(b))


def ne(a, b):
    return Condition("ne", as_expr(a), as_expr(b))


def And(a, b):
    return Condition("and", a, b)


def Or(a, b):
    return Condition("or", a, b)


def Not(a):
    return Condition("not", a)


def conditional(cond, a, b):
    if not isinstance(cond, Condition):
        raise TypeError("conditional() expects a condition (lt, gt, le, ge, ...)")
    return Conditional(cond, as_expr(a), as_expr(b))


def max_value(a, b):
    return MinMax("max", as_expr(a), as_expr(b))


def min_value(a, b):
    return MinMax("min", as_expr(a), as_expr(b))


def as_vector(comps):
    return Vector(comps)


def dot(a, b):
    if isinstance(a, Vector) and isinstance(b, Vector):
        assert len(a) == len(b)
        out = Const(0.0)
        for x, y in zip(a, b):
            out = out + x * y
        return out
    return as_expr(a) * as_expr(b)


inner = dot


# --------------------------------------------------------------------------
# symbolic differentiation
# --------------------------------------------------------------------------
def diff(expr, var):
    """d expr / d var.  ``var`` is ("x", i) or ("species", species_object)."""
    ex = as_expr(expr)
    kind, key = var
    if isinstance(ex, (Const, ConstantRef, FieldGrad, NormalComp, CellCircumradius)):
        return Const(0.0)
    if isinstance(ex, Coord):
        return Const(1.0 if (kind == "x" and ex.i == key) else 0.0)
    if isinstance(ex, FieldRef):
        return FieldGrad(ex.function, key) if kind == "x" else Const(0.0)
    if isinstance(ex, SpeciesRef):
        if kind == "species":
            return Const(1.0 if (ex.species is key and not ex.prev) else 0.0)
        if kind == "species_grad":
            return Const(0.0)
        if ex.prev:
            raise NotImplementedError("spatial gradient of the previous solution inside an expression")
        return SpeciesGrad(ex.species, key)
    if isinstance(ex, SpeciesGrad):
        if kind == "species_grad":  # key = (species, component)
            return Const(1.0 if (ex.species is key[0] and ex.i == key[1]) else 0.0)
        return Const(0.0)  # P1: second derivatives vanish; no dependence on the point value
    if isinstance(ex, Sum):
        return diff(ex.a, var) + diff(ex.b, var)
    if isinstance(ex, Neg):
        return -diff(ex.a, var)
    if isinstance(ex, Product):
        return diff(ex.a, var) * ex.b + ex.a * diff(ex.b, var)
    if isinstance(ex, Division):
        da, db = diff(ex.a, var), diff(ex.b, var)
        out = da / ex.b
        if not _is_const(db, 0.0):
            out = out - ex.a * db / (ex.b * ex.b)
        return out
    if isinstance(ex, Power):
        da, db = diff(ex.a, var), diff(ex.b, var)
        out = Const(0.0)
        if not _is_const(da, 0.0):
            out = out + ex.b * ex.a ** (ex.b - 1) * da
        if not _is_const(db, 0.0):
            out = out + ex * ln(ex.a) * db
        return out
    if isinstance(ex, Math):
        da = diff(ex.a, var)
        if _is_const(da, 0.0):
            return Const(0.0)
        a = ex.a
        d = {
            "exp": lambda: ex,
            "ln": lambda: 1.0 / a,
            "sqrt": lambda: 0.5 / ex,
            "sin": lambda: cos(a),
            "cos": lambda: -sin(a),
            "tan": lambda: 1.0 + ex * ex,
            "tanh": lambda: 1.0 - ex * ex,
            "sinh": lambda: cosh(a),
            "cosh": lambda: sinh(a),
            "abs": lambda: sign(a),
            "erf": lambda: (2.0 / math.sqrt(math.pi)) * exp(-(a * a)),
            "atan": lambda: 1.0 / (1.0 + a * a),
            "asin": lambda: 1.0 / sqrt(1.0 - a * a),
            "acos": lambda: -1.0 / sqrt(1.0 - a * a),
            "sign": lambda: Const(0.0),
        }[ex.name]()
        return d * da
    if isinstance(ex, Conditional):
        return Conditional(ex.cond, diff(ex.a, var), diff(ex.b, var))
    if isinstance(ex, MinMax):
        op = "gt" if ex.kind == "max" else "lt"
        return Conditional(Condition(op, ex.a, ex.b), diff(ex.a, var), diff(ex.b, var))
    raise TypeError(f"cannot differentiate {type(ex)}")


def grad(expr, gdim=None):
    if isinstance(expr, Vector):
        raise NotImplementedError("gradient of a vector expression")
    gdim = gdim or _infer_gdim(expr)
    return Vector([diff(expr, ("x", i)) for i in range(gdim)])


def FacetNormal(mesh=None, gdim=None):
    """Outward unit normal of the exterior facet (vector expression)."""
    g = gdim or getattr(mesh, "gdim", None) or _default_gdim[0] or 3
    return Vector([NormalComp(i) for i in range(g)])


def Circumradius(mesh=None):
    return CellCircumradius()


def div(vec):
    if not isinstance(vec, Vector):
        raise TypeError("div() expects a vector expression")
    out = Const(0.0)
    for i, c in enumerate(vec):
        out = out + diff(c, ("x", i))
    return out


_default_gdim = [None]


def set_default_gdim(gdim):
    _default_gdim[0] = gdim


def _infer_gdim(expr):
    """Dimension for grad(): a field's mesh wins, else the last
    SpatialCoordinate created, else the highest coordinate index used."""
    g = [0, 0]

    def visit(ex):
        if isinstance(ex, Coord):
            g[0] = max(g[0], ex.i + 1)
        elif isinstance(ex, (FieldRef, FieldGrad)):
            g[1] = max(g[1], int(getattr(ex.function, "gdim", 0)))

    walk(as_expr(expr), visit)
    if g[1]:
        return g[1]
    if _default_gdim[0]:
        return max(g[0], _default_gdim[0])
    return max(g[0], 1)


def walk(ex, fn):
    if isinstance(ex, Vector):
        for c in ex:
            walk(c, fn)
        return
    fn(ex)
    for name in ("a", "b", "cond"):
        child = getattr(ex, name, None)
        if isinstance(child, (Expr, Condition)):
            walk(child, fn)


def depends_on(ex, cls):
    found = [False]

    def visit(node):
        if isinstance(node, cls):
            found[0] = True

    walk(as_expr(ex), visit)
    return found[0]


def species_in(ex):
    out = []

    def visit(node):
        if isinstance(node, SpeciesRef) and not node.prev and node.species not in out:
            out.append(node.species)

    walk(as_expr(ex), visit)
    return out


def species_grads_in(ex):
    """(species, component) pairs whose gradient the expression depends on."""
    out = []

    def visit(node):
        if isinstance(node, SpeciesGrad) and (node.species, node.i) not in out:
            out.append((node.species, node.i))

    walk(as_expr(ex), visit)
    return out


# --------------------------------------------------------------------------
# numpy evaluation
# --------------------------------------------------------------------------
def _np_erf(v):
    from scipy.special import erf as _erf

    return _erf(v)


_NP_FUNCS = {
    "exp": np.exp, "ln": np.log, "sqrt": np.sqrt, "sin": np.sin, "cos": np.cos,
    "tan": np.tan, "tanh": np.tanh, "sinh": np.sinh, "cosh": np.cosh, "abs": np.abs,
    "erf": _np_erf, "atan": np.arctan, "asin": np.arcsin, "acos": np.arccos,
    "sign": np.sign,
}


def evaluate(ex, env):
    """Evaluate with numpy.

    env keys: ``x`` -> array (..., gdim) of points; ``field`` -> callable
    (function) -> values at the points; ``field_grad`` -> callable
    (function, i) -> values; ``species`` -> callable (species) -> values.
    """
    ex = as_expr(ex)
    if isinstance(ex, Const):
        return ex.value
    if isinstance(ex, Coord):
        return env["x"][..., ex.i]
    if isinstance(ex, ConstantRef):
        return float(ex.constant.value)
    if isinstance(ex, FieldRef):
        return env["field"](ex.function)
    if isinstance(ex, FieldGrad):
        return env["field_grad"](ex.function, ex.i)
    if isinstance(ex, SpeciesRef):
        return env["species_prev" if ex.prev else "species"](ex.species)
    if isinstance(ex, SpeciesGrad):
        return env["species_grad"](ex.species, ex.i)
    if isinstance(ex, NormalComp):
        return env["normal"](ex.i)
    if isinstance(ex, CellCircumradius):
        return env["circumradius"]()
    if isinstance(ex, Sum):
        return evaluate(ex.a, env) + evaluate(ex.b, env)
    if isinstance(ex, Product):
        return evaluate(ex.a, env) * evaluate(ex.b, env)
    if isinstance(ex, Division):
        return evaluate(ex.a, env) / evaluate(ex.b, env)
    if isinstance(ex, Power):
        return np.power(evaluate(ex.a, env), evaluate(ex.b, env))
    if isinstance(ex, Neg):
        return -evaluate(ex.a, env)
    if isinstance(ex, Math):
        return _NP_FUNCS[ex.name](evaluate(ex.a, env))
    if isinstance(ex, Conditional):
        c = _eval_cond(ex.cond, env)
        with np.errstate(all="ignore"):
            return np.where(c, evaluate(ex.a, env), evaluate(ex.b, env))
    if isinstance(ex, MinMax):
        f = np.maximum if ex.kind == "max" else np.minimum
        return f(evaluate(ex.a, env), evaluate(ex.b, env))
    raise TypeError(f"cannot evaluate {type(ex)}")


def _eval_cond(c, env):
    if c.op in ("and", "or"):
        a, b = _eval_cond(c.a, env), _eval_cond(c.b, env)
        return np.logical_and(a, b) if c.op == "and" else np.logical_or(a, b)
    if c.op == "not":
        return np.logical_not(_eval_cond(c.a, env))
    a, b = evaluate(c.a, env), evaluate(c.b, env)
    return {"lt": np.less, "gt": np.greater, "le": np.less_equal,
            "ge": np.greater_equal, "eq": np.equal, "ne": np.not_equal}[c.op](a, b)


# --------------------------------------------------------------------------
# polynomial degree estimation (UFL estimate_total_polynomial_degree rules,
# ufl/algorithms/estimate_degrees.py)
# --------------------------------------------------------------------------
def estimate_degree(ex, field_degree=1, species_degree=1):
    ex = as_expr(ex)
    if isinstance(ex, (Const, ConstantRef)):
        return 0
    if isinstance(ex, Coord):
        return 1  # affine simplex geometry
    if isinstance(ex, FieldRef):
        return field_degree
    if isinstance(ex, FieldGrad):
        return max(field_degree - 1, 0)
    if isinstance(ex, SpeciesRef):
        return species_degree
    if isinstance(ex, SpeciesGrad):
        return max(species_degree - 1, 0)
    if isinstance(ex, (NormalComp, CellCircumradius)):
        return 0  # affine simplices: piecewise constant
    if isinstance(ex, Sum):
        return max(estimate_degree(ex.a), estimate_degree(ex.b))
    if isinstance(ex, (Product, Division)):
        return estimate_degree(ex.a) + estimate_degree(ex.b)
    if isinstance(ex, Neg):
        return estimate_degree(ex.a)
    if isinstance(ex, Power):
        p = estimate_degree(ex.a)
        if isinstance(ex.b, Const) and ex.b.value >= 0 and float(ex.b.value).is_integer():
            return p * int(ex.b.value)
        return p + 2
    if isinstance(ex, Math):
        p = estimate_degree(ex.a)
        if ex.name in ("abs", "sign"):
            return p
        return p + 2 if p else 0
    if isinstance(ex, Conditional):
        return max(estimate_degree(ex.a), estimate_degree(ex.b))
    if isinstance(ex, MinMax):
        return max(estimate_degree(ex.a), estimate_degree(ex.b))
    raise TypeError(f"cannot estimate the degree of {type(ex)}")


# --------------------------------------------------------------------------
# compilation to the device byte-code (stack machine)
# --------------------------------------------------------------------------
# op codes shared with festim_b200/csrc/program.cuh
OP_CONST, OP_X, OP_SCALAR, OP_FIELD, OP_FIELDGRAD, OP_SPECIES = 0, 1, 2, 3, 4, 5
OP_ADD, OP_MUL, OP_DIV, OP_POW, OP_NEG = 10, 11, 12, 13, 14
OP_MATH = 20  # arg = index in Math.FUNCS
OP_LT, OP_GT, OP_LE, OP_GE, OP_EQ, OP_NE, OP_AND, OP_OR, OP_NOT = 30, 31, 32, 33, 34, 35, 36, 37, 38
OP_SELECT, OP_MAX, OP_MIN = 40, 41, 42
OP_T = 6  # temperature at the point (T constant or P1 field of the problem)
OP_SPECIES_PREV = 7
OP_SPECIES_GRAD = 8   # iarg = species * 4 + component
OP_NORMAL, OP_CIRCUMRADIUS = 9, 15  # facet normal component / circumradius of the owning cell
OP_STORE, OP_LOAD = 50, 51  # common-subexpression temporaries
MAX_TEMPS = 16

MAX_STACK = 24


class ProgramBuilder:
    """Collects byte-code programs plus the tables they refer to.

    ``scalars``: list of objects with ``.value`` (uploaded every step);
    ``fields``: list of nodal P1 functions (objects with ``.x.array``);
    species are referred to by index through ``species_index``.
    """

    def __init__(self, species_index=None, temperature=None):
        self.code = []  # (op, iarg, farg)
        self.offsets = [0]
        self.scalars = []
        self.fields = []
        self.species_index = species_index or {}
        self.temperature = temperature  # object standing for T (Constant or Function)
        self._cache = {}
        self.expressions = []  # expression of every program (run-time code generation)

    def _scalar_id(self, c):
        for i, s in enumerate(self.scalars):
            if s is c:
                return i
        self.scalars.append(c)
        return len(self.scalars) - 1

    def _field_id(self, f):
        for i, s in enumerate(self.fields):
            if s is f:
                return i
        self.fields.append(f)
        return len(self.fields) - 1

    # -- common-subexpression elimination -------------------------------------
    def _skey(self, ex, memo):
        """Structural key of a node (equal keys <=> same value at a point)."""
        k = memo.get(id(ex))
        if k is not None:
            return k
        if isinstance(ex, Condition):
            k = ("cond", ex.op, self._skey(ex.a, memo), None if ex.b is None else self._skey(ex.b, memo))
        elif isinstance(ex, Const):
            k = ("c", ex.value)
        elif isinstance(ex, Coord):
            k = ("x", ex.i)
        elif isinstance(ex, ConstantRef):
            k = ("k", id(ex.constant))
        elif isinstance(ex, FieldRef):
            k = ("f", id(ex.function))
        elif isinstance(ex, FieldGrad):
            k = ("g", id(ex.function), ex.i)
        elif isinstance(ex, SpeciesRef):
            k = ("s", id(ex.species), ex.prev)
        elif isinstance(ex, SpeciesGrad):
            k = ("sg", id(ex.species), ex.i)
        elif isinstance(ex, NormalComp):
            k = ("n", ex.i)
        elif isinstance(ex, CellCircumradius):
            k = ("h",)
        elif isinstance(ex, Neg):
            k = ("neg", self._skey(ex.a, memo))
        elif isinstance(ex, Math):
            k = ("m", ex.name, self._skey(ex.a, memo))
        elif isinstance(ex, Conditional):
            k = ("sel", self._skey(ex.cond, memo), self._skey(ex.a, memo), self._skey(ex.b, memo))
        elif isinstance(ex, MinMax):
            k = (ex.kind, self._skey(ex.a, memo), self._skey(ex.b, memo))
        else:
            k = (type(ex).__name__, self._skey(ex.a, memo), self._skey(ex.b, memo))
        k = hash(k) if len(repr(k)) > 200 else k
        memo[id(ex)] = k
        return k

    def _count(self, ex, memo, counts):
        k = self._skey(ex, memo)
        counts[k] = counts.get(k, 0) + 1
        if counts[k] > 1:
            return  # children already counted once; a repeat costs one LOAD
        for name in ("cond", "a", "b"):
            child = getattr(ex, name, None)
            if isinstance(child, (Expr, Condition)):
                self._count(child, memo, counts)

    @staticmethod
    def _worth_caching(ex):
        return isinstance(ex, (Math, Division, Power, FieldRef, FieldGrad, Product, Sum, Conditional))

    def add(self, ex):
        """Compile ``ex``; returns the program id."""
        ex = as_expr(ex)
        key = id(ex)
        if key in self._cache and self._cache[key][0] is ex:
            return self._cache[key][1]
        out = []
        self._memo, self._counts, self._slots = {}, {}, {}
        self._count(ex, self._memo, self._counts)
        depth = self._emit(ex, out)
        if depth[1] > MAX_STACK:
            raise ValueError(f"expression needs a stack of {depth[1]} > {MAX_STACK}")
        self.code.extend(out)
        self.offsets.append(len(self.code))
        self.expressions.append(ex)
        pid = len(self.offsets) - 2
        self._cache[key] = (ex, pid)
        return pid

    @property
    def n_programs(self):
        return len(self.offsets) - 1

    def _emit(self, ex, out):
        """Returns (net pushes == 1, max stack depth); repeated subtrees are
        evaluated once, stored in a temporary and re-loaded."""
        if isinstance(ex, Condition):
            return self._emit_cond(ex, out)
        k = self._memo.get(id(ex))
        if k is None:
            k = self._skey(ex, self._memo)
        if self._counts.get(k, 0) > 1 and self._worth_caching(ex):
            if k in self._slots:
                out.append((OP_LOAD, self._slots[k], 0.0))
                return (1, 1)
            d = self._emit_node(ex, out)
            if len(self._slots) < MAX_TEMPS:
                self._slots[k] = len(self._slots)
                out.append((OP_STORE, self._slots[k], 0.0))
            return d
        return self._emit_node(ex, out)

    def _emit_node(self, ex, out):
        if isinstance(ex, Const):
            out.append((OP_CONST, 0, ex.value))
            return (1, 1)
        if isinstance(ex, Coord):
            out.append((OP_X, ex.i, 0.0))
            return (1, 1)
        if isinstance(ex, ConstantRef):
            if self.temperature is not None and ex.constant is self.temperature:
                out.append((OP_T, 0, 0.0))
            else:
                out.append((OP_SCALAR, self._scalar_id(ex.constant), 0.0))
            return (1, 1)
        if isinstance(ex, FieldRef):
            if self.temperature is not None and ex.function is self.temperature:
                out.append((OP_T, 0, 0.0))
            else:
                out.append((OP_FIELD, self._field_id(ex.function), 0.0))
            return (1, 1)
        if isinstance(ex, FieldGrad):
            out.append((OP_FIELDGRAD, self._field_id(ex.function) * 4 + ex.i, 0.0))
            return (1, 1)
        if isinstance(ex, SpeciesRef):
            out.append((OP_SPECIES_PREV if ex.prev else OP_SPECIES,
                        self.species_index[ex.species], 0.0))
            return (1, 1)
        if isinstance(ex, SpeciesGrad):
            out.append((OP_SPECIES_GRAD, self.species_index[ex.species] * 4 + ex.i, 0.0))
            return (1, 1)
        if isinstance(ex, NormalComp):
            out.append((OP_NORMAL, ex.i, 0.0))
            return (1, 1)
        if isinstance(ex, CellCircumradius):
            out.append((OP_CIRCUMRADIUS, 0, 0.0))
            return (1, 1)
        if isinstance(ex, Neg):
            d = self._emit(ex.a, out)
            out.append((OP_NEG, 0, 0.0))
            return (1, d[1])
        if isinstance(ex, Math):
            d = self._emit(ex.a, out)
            out.append((OP_MATH, Math.FUNCS.index(ex.name), 0.0))
            return (1, d[1])
        if isinstance(ex, (Sum, Product, Division, Power, MinMax)):
            da = self._emit(ex.a, out)
            db = self._emit(ex.b, out)
            op = {Sum: OP_ADD, Product: OP_MUL, Division: OP_DIV, Power: OP_POW}.get(type(ex))
            if op is None:
                op = OP_MAX if ex.kind == "max" else OP_MIN
            out.append((op, 0, 0.0))
            return (1, max(da[1], 1 + db[1]))
        if isinstance(ex, Conditional):
            dc = self._emit_cond(ex.cond, out)
            da = self._emit(ex.a, out)
            db = self._emit(ex.b, out)
            out.append((OP_SELECT, 0, 0.0))
            return (1, max(dc[1], 1 + da[1], 2 + db[1]))
        raise TypeError(f"cannot compile {type(ex)}")

    def _emit_cond(self, c, out):
        if c.op == "not":
            d = self._emit_cond(c.a, out)
            out.append((OP_NOT, 0, 0.0))
            return (1, d[1])
        da = self._emit(c.a, out)
        db = self._emit(c.b, out)
        op = {"lt": OP_LT, "gt": OP_GT, "le": OP_LE, "ge": OP_GE, "eq": OP_EQ,
              "ne": OP_NE, "and": OP_AND, "or": OP_OR}[c.op]
        out.append((op, 0, 0.0))
        return (1, max(da[1], 1 + db[1]))

    def tables(self):
        ops = np.array([c[0] for c in self.code], dtype=np.int32)
        iargs = np.array([c[1] for c in self.code], dtype=np.int32)
        fargs = np.array([c[2] for c in self.code], dtype=np.float64)
        return ops, iargs, fargs, np.array(self.offsets, dtype=np.int32)


# Error messages of the reference print `type(value)`; report the UFL module paths the
# reference's users (and tests) see, e.g. <class 'ufl.conditional.Conditional'>.
for _cls, _mod in ((Sum, "ufl.algebra"), (Product, "ufl.algebra"), (Division, "ufl.algebra"),
                   (Power, "ufl.algebra"), (Conditional, "ufl.conditional"), (MinMax, "ufl.conditional"),
                   (Math, "ufl.mathfunctions")):
    _cls.__module__ = _mod
del _cls, _mod
