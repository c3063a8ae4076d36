"""Species and implicit species (reference ``src/festim/species.py``)."""

from __future__ import annotations

from festim_b200 import fem, ufl
from festim_b200.helpers import as_fenics_constant


class Species:
    """``species.py:12-130``.  ``solution`` is the symbolic concentration that
    enters reaction / source expressions, ``post_processing_solution`` the
    collapsed nodal field refreshed after every solve."""

    def __init__(self, name=None, mobile=True, subdomains=None):
        self.name = name
        self.mobile = mobile
        self.solution = None
        self.prev_solution = None
        self.test_function = None
        self.sub_function = None
        self.sub_function_space = None
        self.collapsed_function_space = None
        self.map_sub_to_main_solution = None
        self.post_processing_solution = None
        self.subdomains = subdomains
        self.subdomain_to_solution = {}

    def __repr__(self):
        return f"Species({self.name})"

    def __str__(self):
        return f"{self.name}"

    @property
    def concentration(self):
        return self.solution

    @property
    def legacy(self):
        return not self.subdomain_to_solution


class ImplicitSpecies:
    """c = n - sum(others)  (``species.py:133-241``)."""

    def __init__(self, n, others=None, name=None):
        self.name = name
        self.n = n
        self.others = others
        self.value_fenics = None

    def __repr__(self):
        return f"ImplicitSpecies({self.name}, {self.n}, {self.others})"

    def __str__(self):
        return f"{self.name}"

    @property
    def concentration(self):
        for other in self.others:
            if other.solution is None:
                raise ValueError(
                    f"Cannot compute concentration of {self.name} "
                    + f"because {other.name} has no solution."
                )
        out = ufl.as_expr(self.value_fenics)
        for other in self.others:
            out = out - other.solution
        return out

    def create_value_fenics(self, mesh, t):
        """``species.py:189-227``."""
        if isinstance(self.n, (int, float)):
            self.value_fenics = as_fenics_constant(self.n, mesh)
        elif isinstance(self.n, (fem.Function, ufl.Expr)):
            self.value_fenics = self.n
        elif callable(self.n):
            arguments = self.n.__code__.co_varnames
            if "t" in arguments and "x" not in arguments and "T" not in arguments:
                val = self.n(t=float(t))
                if not isinstance(val, (float, int)):
                    raise ValueError(
                        f"self.value should return a float or an int, not {type(val)}"
                    )
                self.value_fenics = as_fenics_constant(val, mesh)
            else:
                kwargs = {}
                if "t" in arguments:
                    kwargs["t"] = t
                if "x" in arguments:
                    kwargs["x"] = ufl.SpatialCoordinate(mesh)
                self.value_fenics = self.n(**kwargs)

    def update_density(self, t):
        """``species.py:229-241``: only a density that is a function of t alone
        is refreshed (spatial expressions hold the live time constant)."""
        if isinstance(self.n, (fem.Function, ufl.Expr)):
            return
        if callable(self.n):
            arguments = self.n.__code__.co_varnames
            if isinstance(self.value_fenics, fem.Constant) and "t" in arguments:
                self.value_fenics.value = self.n(t=t)


def find_species_from_name(name, species):
    for spe in species:
        if spe.name == name:
            return spe
    raise ValueError(f"Species {name} not found in list of species")
