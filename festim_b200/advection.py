"""Advection term (reference ``src/festim/advection.py``)."""

import numpy as np

from festim_b200 import fem, ufl
from festim_b200.helpers import Value
from festim_b200.species import Species
from festim_b200.subdomain import VolumeSubdomain


class VectorFunction:
    """P1 vector field: ``array[v, i]``; what ``velocity`` must be (the
    stand-in for a ``fem.Function`` on a vector Lagrange space)."""

    def __init__(self, mesh, values=None):
        self.mesh = getattr(mesh, "mesh", mesh) if not hasattr(mesh, "coords") else mesh
        self.array = np.zeros((self.mesh.num_vertices, self.mesh.gdim))
        if values is not None:
            self.interpolate(values)

    def interpolate(self, f):
        if isinstance(f, VectorFunction):
            self.array[:] = f.array
        elif callable(f):
            out = np.asarray(f(self.mesh.x3()), dtype=np.float64)
            self.array[:] = out[: self.mesh.gdim].T
        else:
            self.array[:] = np.asarray(f, dtype=np.float64).reshape(self.array.shape)


class VelocityField(Value):
    """``advection.py:96-170``: velocity interpolated to a P1 vector field."""

    def __init__(self, input_value):
        self._input = input_value
        self.species_dependent_value = {}
        self.fenics_object = None
        self.fenics_interpolation_expression = None

    @property
    def input_value(self):
        return self._input

    @property
    def explicit_time_dependent(self):
        v = self._input
        return callable(v) and not isinstance(v, VectorFunction) and "t" in v.__code__.co_varnames

    def convert_input_value(self, function_space, t=None):
        v = self._input
        if isinstance(v, VectorFunction):
            vel = v
        elif callable(v):
            if v.__code__.co_varnames != ("t",):
                raise TypeError("velocity function can only be a function of time arg t")
            vel = v(t)
            if not isinstance(vel, VectorFunction):
                raise ValueError(
                    "A time dependent advection field should return an fem.Function"
                    f", not a {type(vel)}"
                )
        self.fenics_object = VectorFunction(function_space.mesh)
        self.fenics_object.interpolate(vel)

    def update(self, t):
        self.fenics_object.interpolate(self._input(t=float(t)))


class AdvectionTerm:
    """``advection.py:16-93``."""

    def __init__(self, velocity, subdomain, species):
        self.velocity = velocity
        self.subdomain = subdomain
        self.species = species

    @property
    def velocity(self):
        return self._velocity

    @velocity.setter
    def velocity(self, value):
        err = f"velocity must be a fem.Function, or callable not {type(value)}"
        if value is None or isinstance(value, VectorFunction):
            self._velocity = VelocityField(value)
        elif isinstance(value, (fem.Constant, fem.Expression, ufl.Expr)):
            raise TypeError(err)
        elif callable(value):
            self._velocity = VelocityField(value)
        else:
            raise TypeError(err)

    @property
    def subdomain(self):
        return self._subdomain

    @subdomain.setter
    def subdomain(self, value):
        if value is None or isinstance(value, VolumeSubdomain):
            self._subdomain = value
        else:
            raise TypeError(f"Subdomain must be a festim.Subdomain object, not {type(value)}")

    @property
    def species(self):
        return self._species

    @species.setter
    def species(self, value):
        if not isinstance(value, list):
            value = [value]
        for spe in value:
            if not isinstance(spe, Species):
                raise TypeError(
                    f"elements of species must be of type festim.Species not {type(spe)}"
                )
        self._species = value
