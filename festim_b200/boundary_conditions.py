"""Boundary conditions (reference ``src/festim/boundary_conditions/``)."""

from __future__ import annotations

import numpy as np

from festim_b200 import fem, helpers, ufl
from festim_b200.constants import k_B
from festim_b200.subdomain import SurfaceSubdomain


class DirichletBCBase:
    """``dirichlet_bc.py:17-168``: value -> Constant (float, f(t)) or nodal
    Function (anything depending on x / T), dofs = vertices of tagged facets."""

    def __init__(self, subdomain, value, enforce_weakly=False, penalty=None):
        self.subdomain = subdomain
        self.value = value
        self.enforce_weakly = enforce_weakly
        self.penalty = penalty
        self.value_fenics = None
        self.bc_expr = None

    @property
    def value_fenics(self):
        return self._value_fenics

    @value_fenics.setter
    def value_fenics(self, value):
        if value is not None and not isinstance(
            value, (fem.Function, fem.Constant, np.ndarray, ufl.Expr)
        ):
            raise TypeError(
                "Value must be a dolfinx.fem.Function, dolfinx.fem.Constant, a np.ndarray or a ufl expression, not"
                + f"{type(value)}"
            )
        self._value_fenics = value

    @property
    def time_dependent(self):
        if self.value is None or isinstance(self.value, fem.Constant):
            return False
        if callable(self.value) and not isinstance(self.value, ufl.Expr):
            return "t" in self.value.__code__.co_varnames
        return False

    @property
    def temperature_dependent(self):
        if self.value is None or isinstance(self.value, fem.Constant):
            return False
        if callable(self.value) and not isinstance(self.value, ufl.Expr):
            return "T" in self.value.__code__.co_varnames
        return False

    def numerical_flux(self, u, D, mesh, species):
        """Numerical flux leaving the domain through a weakly enforced boundary
        (``dirichlet_bc.py:170-209``): ``-D n.grad(u) + penalty D / h (u - value)``."""
        assert self.penalty is not None, "Penalty parameter must be given for weakly enforced Dirichlet BCs"
        assert self.value_fenics is not None, "value_fenics must be defined for weakly enforced Dirichlet BCs"
        n = ufl.FacetNormal(gdim=mesh.gdim)
        h = ufl.Circumradius(mesh)
        gradu = [ufl.SpeciesGrad(species, i) for i in range(mesh.gdim)]
        ndotgrad = n[0] * gradu[0]
        for i in range(1, mesh.gdim):
            ndotgrad = ndotgrad + n[i] * gradu[i]
        D = ufl.as_expr(D)
        return -D * ndotgrad + self.penalty * D / h * (u - ufl.as_expr(self.value_fenics))

    def weak_formulation(self, species, u, ds, D, mesh):
        """Symmetric Nitsche terms (``dirichlet_bc.py:211-244``) as integrals of the form
        machinery: consistency + penalty tested with v, symmetry term tested with n.grad(v)."""
        from festim_b200 import forms

        ds_bc = ds(self.subdomain.id)
        return [
            forms.TestIntegral(species, self.numerical_flux(u, D, mesh, species), ds_bc, {"nitsche": "flux"}),
            forms.NormalGradTestIntegral(species, -ufl.as_expr(D) * (u - ufl.as_expr(self.value_fenics)), ds_bc,
                                         {"nitsche": "symmetry"}),
        ]

    def define_surface_subdomain_vertices(self, facet_meshtags, mesh):
        """Vertices of the facets tagged with this BC's surface id
        (``dirichlet_bc.py:108-147``, ``locate_dofs_topological``)."""
        facets = facet_meshtags.find(self.subdomain.id)
        return np.unique(mesh.bfacets[facets].ravel()).astype(np.int32)

    def _create_value(self, function_space, t, temperature=None):
        mesh = function_space.mesh
        if isinstance(self.value, (int, float)):
            self.value_fenics = helpers.as_fenics_constant(self.value, mesh)
        elif isinstance(self.value, (fem.Constant, fem.Function)):
            self.value_fenics = self.value
        elif callable(self.value):
            arguments = self.value.__code__.co_varnames
            if "t" in arguments and "x" not in arguments and "T" not in arguments:
                val = self.value(t=float(t))
                if not isinstance(val, (float, int)):
                    raise ValueError(
                        "self.value should return a float or an int, not " + f"{type(val)} "
                    )
                self.value_fenics = helpers.as_fenics_constant(val, mesh)
            else:
                kwargs = {}
                if "t" in arguments:
                    kwargs["t"] = t
                if "x" in arguments:
                    kwargs["x"] = ufl.SpatialCoordinate(mesh)
                if "T" in arguments:
                    kwargs["T"] = temperature
                self.value_fenics = fem.Function(function_space)
                expr = self.value(**kwargs)
                assert isinstance(expr, ufl.Expr), f"{type(expr)}"
                self.bc_expr = fem.Expression(expr)
                self.value_fenics.interpolate(self.bc_expr)

    def update(self, t):
        """``dirichlet_bc.py:149-168``."""
        if callable(self.value) and not isinstance(self.value, (fem.Function, fem.Constant)):
            arguments = self.value.__code__.co_varnames
            if isinstance(self.value_fenics, fem.Constant) and "t" in arguments:
                self.value_fenics.value = self.value(t=t)
            else:
                self.value_fenics.interpolate(self.bc_expr)
        elif self.bc_expr is not None:
            self.value_fenics.interpolate(self.bc_expr)


class FixedConcentrationBC(DirichletBCBase):
    """``dirichlet_bc.py:247-467``."""

    def __init__(self, subdomain, value, species, enforce_weakly=False, penalty=None):
        self.species = species
        super().__init__(subdomain, value, enforce_weakly, penalty)

    def create_value(self, function_space, temperature, t, K_S=None):
        if K_S is not None:
            raise NotImplementedError("change-of-variable problems are out of scope")
        self._create_value(function_space, t, temperature)


DirichletBC = FixedConcentrationBC


class FixedTemperatureBC(DirichletBCBase):
    def create_value(self, function_space, t):
        self._create_value(function_space, t)


def sieverts_law(T, S_0, E_S, pressure):
    """c = S_0 exp(-E_S / k_B / T) sqrt(P)  (``sieverts_bc.py:8-11``)."""
    S = S_0 * ufl.exp(-E_S / k_B / T)
    return S * pressure**0.5


def henrys_law(T, H_0, E_H, pressure):
    """c = H_0 exp(-E_H / k_B / T) P  (``henrys_bc.py:8-11``)."""
    H = H_0 * ufl.exp(-E_H / k_B / T)
    return H * pressure


def _pressure_value_function(law, c0, E, pressure):
    """Value callable whose argument names reflect what ``pressure`` depends on
    (``sieverts_bc.py:82-135``): always T, plus x / t when the pressure is a
    callable of them."""
    if callable(pressure):
        args = tuple(sorted(pressure.__code__.co_varnames))
        table = {
            ("x",): lambda T, x=None: law(T, c0, E, pressure(x=x)),
            ("t",): lambda T, t=None: law(T, c0, E, pressure(t=t)),
            ("T",): lambda T: law(T, c0, E, pressure(T=T)),
            ("t", "x"): lambda T, x=None, t=None: law(T, c0, E, pressure(x=x, t=t)),
            ("T", "x"): lambda T, x=None: law(T, c0, E, pressure(x=x, T=T)),
            ("T", "t"): lambda T, t=None: law(T, c0, E, pressure(t=t, T=T)),
            ("T", "t", "x"): lambda T, x=None, t=None: law(T, c0, E, pressure(x=x, t=t, T=T)),
        }
        if args not in table:
            raise ValueError("pressure function not supported")
        return table[args]
    return lambda T: law(T, c0, E, pressure)


class SievertsBC(FixedConcentrationBC):
    stoichiometry = 2

    def __init__(self, subdomain, S_0, E_S, pressure, species, enforce_weakly=False, penalty=None):
        self.S_0 = S_0
        self.E_S = E_S
        self.pressure = pressure
        value = _pressure_value_function(sieverts_law, S_0, E_S, pressure)
        super().__init__(value=value, species=species, subdomain=subdomain,
                         enforce_weakly=enforce_weakly, penalty=penalty)


class HenrysBC(FixedConcentrationBC):
    stoichiometry = 1

    def __init__(self, subdomain, H_0, E_H, pressure, species, enforce_weakly=False, penalty=None):
        self.H_0 = H_0
        self.E_H = E_H
        self.pressure = pressure
        value = _pressure_value_function(henrys_law, H_0, E_H, pressure)
        super().__init__(value=value, species=species, subdomain=subdomain,
                         enforce_weakly=enforce_weakly, penalty=penalty)


class FluxBCBase:
    """``flux_bc.py:9-137``: value -> Constant (float, f(t)) or expression."""

    def __init__(self, subdomain, value):
        self.subdomain = subdomain
        self.value = value
        self.value_fenics = None
        self.bc_expr = None
        self.species_dependent_value = {}

    @property
    def value_fenics(self):
        return self._value_fenics

    @value_fenics.setter
    def value_fenics(self, value):
        if value is not None and not isinstance(value, (fem.Function, fem.Constant, np.ndarray, ufl.Expr)):
            raise TypeError(                                   # reference flux_bc.py:51-62
                "Value must be a dolfinx.fem.Function, dolfinx.fem.Constant,"
                f" np.ndarray or ufl.core.expr.Expr not {type(value)}")
        self._value_fenics = value

    @property
    def time_dependent(self):
        if self.value is None:
            if type(self).__name__.startswith("SurfaceReaction"):
                return False
            raise TypeError("Value must be given to determine if its time dependent")
        if isinstance(self.value, fem.Constant):
            return False
        if callable(self.value) and not isinstance(self.value, ufl.Expr):
            return "t" in self.value.__code__.co_varnames
        return False

    @property
    def temperature_dependent(self):
        if self.value is None or isinstance(self.value, fem.Constant):
            return False
        if callable(self.value) and not isinstance(self.value, ufl.Expr):
            return "T" in self.value.__code__.co_varnames
        return False

    def create_value_fenics(self, mesh, temperature, t):
        if isinstance(self.value, (int, float)):
            self.value_fenics = helpers.as_fenics_constant(self.value, mesh)
        elif isinstance(self.value, (fem.Constant, fem.Function, ufl.Expr)):
            self.value_fenics = self.value
        elif callable(self.value):
            arguments = self.value.__code__.co_varnames
            if ("t" in arguments and "x" not in arguments and "T" not in arguments
                    and not self.species_dependent_value):
                val = self.value(t=float(t))
                if not isinstance(val, (float, int)):
                    raise ValueError(
                        f"self.value should return a float or an int, not {type(val)} "
                    )
                self.value_fenics = helpers.as_fenics_constant(val, mesh)
            else:
                kwargs = {}
                if "t" in arguments:
                    kwargs["t"] = t
                if "x" in arguments:
                    kwargs["x"] = ufl.SpatialCoordinate(mesh)
                if "T" in arguments:
                    kwargs["T"] = temperature
                for name, species in self.species_dependent_value.items():
                    kwargs[name] = species.concentration
                self.value_fenics = self.value(**kwargs)

    def update(self, t):
        if callable(self.value) and not isinstance(self.value, (fem.Function, fem.Constant)):
            arguments = self.value.__code__.co_varnames
            if isinstance(self.value_fenics, fem.Constant) and "t" in arguments:
                self.value_fenics.value = self.value(t=t)


class ParticleFluxBC(FluxBCBase):
    """-D grad(c).n = value  (``flux_bc.py:140-252``)."""

    def __init__(self, subdomain, value, species, species_dependent_value=None):
        super().__init__(subdomain=subdomain, value=value)
        self.species = species
        self.species_dependent_value = species_dependent_value or {}


class HeatFluxBC(FluxBCBase):
    """-lambda grad(T).n = value  (``flux_bc.py:255-295``)."""


class SurfaceReactionBCpartial(ParticleFluxBC):
    """Flux of ONE reactant of a surface reaction A + B <-> C
    (``surface_reaction.py:9-85``): ``K_d P_C - K_r c_A c_B`` with Arrhenius ``K_r``,
    ``K_d``.  The value depends on the reactants' concentrations, so the lowering
    differentiates it with respect to every species (facet Jacobian blocks)."""

    def __init__(self, reactant, gas_pressure, k_r0, E_kr, k_d0, E_kd, subdomain, species):
        self.reactant = reactant
        self.gas_pressure = gas_pressure
        self.k_r0 = k_r0
        self.E_kr = E_kr
        self.k_d0 = k_d0
        self.E_kd = E_kd
        super().__init__(subdomain=subdomain, value=None, species=species)

    def create_value_fenics(self, mesh, temperature, t):
        from festim_b200 import k_B

        kr = self.k_r0 * ufl.exp(-self.E_kr / (k_B * ufl.as_expr(temperature)))
        kd = self.k_d0 * ufl.exp(-self.E_kd / (k_B * ufl.as_expr(temperature)))
        if hasattr(self.gas_pressure, "solution") and not callable(self.gas_pressure):
            raise NotImplementedError("gas enclosures (GasSpecies pressures) are out of scope (SURVEY 2/8)")
        if callable(self.gas_pressure):
            gas_pressure = self.gas_pressure(t=t)
        else:
            gas_pressure = self.gas_pressure
        product_of_reactants = self.reactant[0].concentration
        for reactant in self.reactant[1:]:
            product_of_reactants = product_of_reactants * reactant.concentration
        self.value_fenics = kd * ufl.as_expr(gas_pressure) - kr * product_of_reactants


class SurfaceReactionBC:
    """Surface reaction A + B <-> C on a boundary (``surface_reaction.py:88-151``): one
    :class:`SurfaceReactionBCpartial` flux per reactant (a reactant listed twice gets the
    flux twice: A + A <-> A2 removes two particles per event)."""

    def __init__(self, reactant, gas_pressure, k_r0, E_kr, k_d0, E_kd, subdomain):
        self.reactant = reactant
        self.gas_pressure = gas_pressure
        self.k_r0 = k_r0
        self.E_kr = E_kr
        self.k_d0 = k_d0
        self.E_kd = E_kd
        self.subdomain = subdomain
        self.flux_bcs = [
            SurfaceReactionBCpartial(reactant=self.reactant, gas_pressure=self.gas_pressure,
                                     k_r0=self.k_r0, E_kr=self.E_kr, k_d0=self.k_d0, E_kd=self.E_kd,
                                     subdomain=self.subdomain, species=species)
            for species in self.reactant
        ]

    @property
    def time_dependent(self):
        return False  # as the reference: the value is an expression of t, nothing to update


__all__ = [
    "DirichletBC", "DirichletBCBase", "FixedConcentrationBC", "FixedTemperatureBC",
    "SievertsBC", "HenrysBC", "FluxBCBase", "ParticleFluxBC", "HeatFluxBC",
    "SurfaceReactionBC", "SurfaceReactionBCpartial", "SurfaceSubdomain",
]
