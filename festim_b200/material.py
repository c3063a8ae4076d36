"""Material properties (reference ``src/festim/material.py``)."""

from __future__ import annotations

from enum import Enum

from festim_b200 import fem, ufl
from festim_b200.constants import k_B


class SolubilityLaw(Enum):
    SIEVERT = "sievert"
    HENRY = "henry"
    NONE = "none"

    @classmethod
    def from_string(cls, s):
        try:
            return cls(s.lower())
        except ValueError:
            raise ValueError(f"Could not set solubility law {s=}")


def _lookup(table, species, what):
    """Scalar or per-species dict keyed by Species object or name
    (``material.py:148-204``)."""
    if isinstance(table, (float, int)):
        return table
    if isinstance(table, dict):
        if species is None:
            raise ValueError(f"species must be provided if {what} is a dict")
        if species in table:
            return table[species]
        if getattr(species, "name", None) in table:
            return table[species.name]
        raise ValueError(f"{species} is not in {what} keys")
    raise TypeError(f"{what} must be either a float, int or a dict")


class Material:
    """``material.py:31-374``: Arrhenius D = D_0 exp(-E_D / k_B / T), solubility
    K_S = K_S_0 exp(-E_K_S / k_B / T), thermal properties."""

    def __init__(self, D_0=None, E_D=None, K_S_0=None, E_K_S=None,
                 thermal_conductivity=None, density=None, heat_capacity=None,
                 name=None, solubility_law=SolubilityLaw.NONE, D=None):
        self.D_0 = D_0
        self.E_D = E_D
        self.K_S_0 = K_S_0
        self.E_K_S = E_K_S
        self.thermal_conductivity = thermal_conductivity
        self.density = density
        self.heat_capacity = heat_capacity
        self.name = name
        self.solubility_law = solubility_law
        self.D = D
        if self.D_0 and self.D:
            raise ValueError(
                "D_0 and D cannot be set at the same time. Please set only one of them."
            )
        if self.D_0 is None and self.D is None:
            raise ValueError("D_0 and D cannot both be None. Please set one of them.")

    @property
    def solubility_law(self):
        return self._solubility_law

    @solubility_law.setter
    def solubility_law(self, value):
        if isinstance(value, SolubilityLaw):
            self._solubility_law = value
        elif isinstance(value, str):
            self._solubility_law = SolubilityLaw.from_string(value)
        else:
            raise ValueError(f"Could not set solubility law {value=}")

    @property
    def D(self):
        return self._D

    @D.setter
    def D(self, value):
        if value is None or isinstance(value, fem.Function):
            self._D = value
        else:
            raise TypeError("D must be of type fem.Function")

    def get_D_0(self, species=None):
        return _lookup(self.D_0, species, "D_0")

    def get_E_D(self, species=None):
        return _lookup(self.E_D, species, "E_D")

    def get_K_S_0(self, species=None):
        return _lookup(self.K_S_0, species, "K_S_0")

    def get_E_K_S(self, species=None):
        return _lookup(self.E_K_S, species, "E_K_S")

    def get_diffusion_coefficient(self, mesh=None, temperature=None, species=None):
        """``material.py:262-318``."""
        if self.D:
            return self.D
        if isinstance(self.D_0, dict) != isinstance(self.E_D, dict):
            raise ValueError("D_0 and E_D must be either floats or dicts")
        if isinstance(self.D_0, dict):
            if list(self.D_0.keys()) != list(self.E_D.keys()):
                raise ValueError("D_0 and E_D have different keys")
            if species is None:
                raise ValueError("species must be provided if D_0 and E_D are dicts")
        elif not isinstance(self.D_0, (float, int)) or not isinstance(self.E_D, (float, int)):
            raise ValueError("D_0 and E_D must be either floats or dicts")
        D_0 = self.get_D_0(species)
        E_D = self.get_E_D(species)
        return D_0 * ufl.exp(-E_D / k_B / temperature)

    def get_solubility_coefficient(self, mesh=None, temperature=None, species=None):
        """``material.py:320-374``."""
        if isinstance(self.K_S_0, dict) != isinstance(self.E_K_S, dict):
            raise ValueError("K_S_0 and E_K_S must be either floats or dicts")
        if isinstance(self.K_S_0, dict):
            if list(self.K_S_0.keys()) != list(self.E_K_S.keys()):
                raise ValueError("K_S_0 and E_K_S have different keys")
            if species is None:
                raise ValueError("species must be provided if K_S_0 and E_K_S are dicts")
        elif not isinstance(self.K_S_0, (float, int)) or not isinstance(self.E_K_S, (float, int)):
            raise ValueError("K_S_0 and E_K_S must be either floats or dicts")
        return self.get_K_S_0(species) * ufl.exp(-self.get_E_K_S(species) / k_B / temperature)
