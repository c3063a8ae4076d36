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


# ---------------------------------------------------------------------------------------------
# PETSc-style callbacks of the reference (helpers.py:404-454), kept for scripts that install them
# on `problem.solver.solver` or call them directly.  The solve itself applies the same stopping
# rule on the device (`fb_newton_solve`, csrc/newton.cu); these are its host-side statement.
# ---------------------------------------------------------------------------------------------
_residual0 = 0
_prev_xnorm = 0


def convergenceTest(snes, it, norms):
    """``helpers.py:408-426``: SNES convergence test ``(snes, it, (xnorm, snorm, fnorm)) -> reason``
    with the reason codes of ``snes.ConvergedReason``.  As in the reference the residual of
    iteration 0 is a module global."""
    global _residual0
    _xnorm, gnorm, f = norms
    rtol, atol, stol, max_its = snes.getTolerances()
    if it == 0:
        _residual0 = f
    if it > max_its:
        return snes.ConvergedReason.DIVERGED_MAX_IT
    elif f < atol:
        return snes.ConvergedReason.CONVERGED_FNORM_ABS
    elif f / _residual0 < rtol:
        return snes.ConvergedReason.CONVERGED_FNORM_RELATIVE
    elif gnorm < stol and it > 0:
        return snes.ConvergedReason.CONVERGED_SNORM_RELATIVE
    else:
        return snes.ConvergedReason.ITERATING


def SnesMonitor(snes, iter, rnorm):
    """``helpers.py:429-448``: one INFO line per Newton iteration (Python ``logging`` in place of
    ``dolfinx.log``)."""
    import logging

    global _prev_xnorm
    rtol, atol, stol, _max_its = snes.getTolerances()
    xnorm = snes.getSolution().norm()
    stepsize_rel = abs(xnorm - _prev_xnorm) / xnorm if iter > 0 else float("inf")
    relative_residual = float("inf") if iter == 0 else rnorm / _residual0
    logging.getLogger("festim_b200").info(
        f"SNES {iter=} ; {rnorm=:.5e} ({atol=}) ; {relative_residual=:.5e} ({rtol=}) ; "
        f"{stepsize_rel=:.5e} ({stol=:.5e})")
    _prev_xnorm = xnorm


def KSPMonitor(ksp, iter, rnorm):
    """``helpers.py:451-454``."""
    import logging

    logging.getLogger("festim_b200").debug(f"KSP {iter=} {rnorm=:.5e}")
