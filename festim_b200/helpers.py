"""Input-value conversion, mirroring reference ``src/festim/helpers.py``.

``Value`` follows ``helpers.py:133-342``: floats become ``fem.Constant``;
callables of ``t`` only are called with a Python float and stored as a
``Constant``; other callables are mapped to symbolic ``x`` / ``t`` / ``T`` /
species arguments by name (``__code__.co_varnames``, ``helpers.py:67-84``) and
either kept as an expression (``up_to_ufl_expr=True``) or interpolated at the
mesh vertices into a ``fem.Function``.
"""

from __future__ import annotations

import numpy as np

from festim_b200 import fem, ufl


def as_fenics_constant(value, mesh):
    if isinstance(value, (float, int)) and not isinstance(value, bool):
        return fem.Constant(mesh, float(value))
    if isinstance(value, (np.floating, np.integer)):
        return fem.Constant(mesh, float(value))
    if isinstance(value, fem.Constant):
        return value
    raise TypeError(
        f"Value must be a float, an int or a dolfinx.Constant, not {type(value)}"
    )


def _varnames(fn):
    return fn.__code__.co_varnames


def as_mapped_function(value, function_space=None, t=None, temperature=None,
                       species_dependent_value=None, subdomain=None):
    arguments = _varnames(value)
    kwargs = {}
    if "t" in arguments:
        kwargs["t"] = t
    if "x" in arguments:
        kwargs["x"] = ufl.SpatialCoordinate(function_space.mesh)
    if "T" in arguments:
        kwargs["T"] = temperature
    for name, species in (species_dependent_value or {}).items():
        kwargs[name] = species.concentration
    return value(**kwargs)


def as_fenics_interp_expr_and_function(value, function_space, t=None, temperature=None,
                                       species_dependent_value=None, subdomain=None):
    mapped = as_mapped_function(value, function_space, t, temperature,
                                species_dependent_value, subdomain)
    expr = fem.Expression(mapped)
    fn = fem.Function(function_space)
    fn.interpolate(expr)
    return expr, fn


class Value:
    """User input value -> Constant / Function / expression (``helpers.py:133``)."""

    def __init__(self, input_value, species_dependent_value=None):
        self.input_value = input_value
        self.species_dependent_value = species_dependent_value or {}
        self.ufl_expression = None
        self.fenics_interpolation_expression = None
        self.fenics_object = None

    def __repr__(self):
        return str(self.input_value)

    @property
    def input_value(self):
        return self._input_value

    @input_value.setter
    def input_value(self, value):
        if value is None or isinstance(
            value, (float, int, fem.Constant, np.ndarray, fem.Expression, ufl.Expr, fem.Function)
        ) or callable(value):
            self._input_value = value
        else:
            raise TypeError(
                "Value must be a float, int, fem.Constant, np.ndarray, fem.Expression,"
                f" ufl.core.expr.Expr, fem.Function, or callable not {value}"
            )

    def _is_symbolic(self):
        return isinstance(self.input_value, ufl.Expr)

    @property
    def explicit_time_dependent(self):
        if self.input_value is None or self._is_symbolic():
            return False
        if callable(self.input_value):
            return "t" in _varnames(self.input_value)
        return False

    @property
    def species_dependent(self):
        if not callable(self.input_value) or self._is_symbolic():
            return False
        return bool(self.species_dependent_value)

    @property
    def temperature_dependent(self):
        if self.input_value is None or self._is_symbolic():
            return False
        if callable(self.input_value):
            return "T" in _varnames(self.input_value)
        return False

    def convert_input_value(self, function_space=None, t=None, temperature=None,
                            up_to_ufl_expr=False, subdomain=None):
        v = self.input_value
        if isinstance(v, (fem.Constant, fem.Function, ufl.Expr)):
            self.fenics_object = v
        elif isinstance(v, fem.Expression):
            self.fenics_interpolation_expression = v
        elif isinstance(v, (float, int)):
            self.fenics_object = as_fenics_constant(v, function_space.mesh)
        elif callable(v):
            args = _varnames(v)
            if ("t" in args and "x" not in args and "T" not in args
                    and not self.species_dependent_value):
                val = v(t=float(t))
                if not isinstance(val, (float, int)):
                    raise ValueError(
                        "self.value should return a float or an int, not " + f"{type(val)} "
                    )
                self.fenics_object = as_fenics_constant(val, function_space.mesh)
            elif up_to_ufl_expr:
                self.fenics_object = as_mapped_function(
                    v, function_space, t, temperature, self.species_dependent_value, subdomain
                )
            else:
                self.fenics_interpolation_expression, self.fenics_object = (
                    as_fenics_interp_expr_and_function(
                        v, function_space, t, temperature,
                        self.species_dependent_value, subdomain)
                )

    def update(self, t):
        if callable(self.input_value) and not self._is_symbolic():
            arguments = _varnames(self.input_value)
            if isinstance(self.fenics_object, fem.Constant) and "t" in arguments:
                self.fenics_object.value = float(self.input_value(t=t))
            elif isinstance(self.fenics_object, fem.Function):
                if self.fenics_interpolation_expression is not None:
                    self.fenics_object.interpolate(self.fenics_interpolation_expression)


def is_it_time_to_export(times, current_time, atol=0, rtol=1.0e-5):
    """True (and the matching time is consumed) when ``current_time`` is one of
    ``times``; always True when ``times`` is None (``helpers.py:374-401``)."""
    if times is None:
        return True
    for i, time in enumerate(times):
        if np.isclose(time, current_time, atol=atol, rtol=rtol):
            del times[i]
            return True
    return False
