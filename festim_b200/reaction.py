"""Reactions between species (reference ``src/festim/reaction.py``)."""

from __future__ import annotations

from festim_b200 import ufl
from festim_b200.constants import k_B
from festim_b200.species import ImplicitSpecies, Species
from festim_b200.subdomain import VolumeSubdomain


class Reaction:
    """reactants <--> products with forward rate k_0 exp(-E_k / k_B T) and
    backward rate p_0 exp(-E_p / k_B T)  (``reaction.py:14-224``)."""

    def __init__(self, reactant, k_0, E_k, volume, product=None, p_0=None, E_p=None):
        self.k_0 = k_0
        self.E_k = E_k
        self.p_0 = p_0
        self.E_p = E_p
        self.reactant = reactant
        self.product = [] if product is None else product
        self.volume = volume

    @property
    def reactant(self):
        return self._reactant

    @reactant.setter
    def reactant(self, value):
        if not isinstance(value, list):
            value = [value]
        if len(value) == 0:
            raise ValueError(
                "reactant must be an entry of one or more species objects, not an empty list."
            )
        for v in value:
            if not isinstance(v, (Species, ImplicitSpecies)):
                raise TypeError(
                    "reactant must be an F.Species or F.ImplicitSpecies, "
                    + f"not {type(v)}"
                )
        self._reactant = value

    @property
    def volume(self):
        return self._volume

    @volume.setter
    def volume(self, value):
        if not isinstance(value, VolumeSubdomain):
            raise TypeError("volume must be of type festim.VolumeSubdomain")
        self._volume = value

    @property
    def products(self):
        return self.product if isinstance(self.product, list) else [self.product]

    def __repr__(self):
        return (f"Reaction({self}, {self.k_0}, {self.E_k}, {self.p_0}, {self.E_p})")

    def __str__(self):
        reactants = " + ".join(str(r) for r in self.reactant)
        products = " + ".join(str(p) for p in self.products) if isinstance(self.product, list) else self.product
        return f"{reactants} <--> {products}"

    def rate_mode(self):
        """Detrapping-rate mode by Python truthiness (``reaction.py:182-187``,
        SURVEY quirk Q6): 2 = Arrhenius, 1 = constant p_0, 0 = none."""
        if self.p_0 and self.E_p:
            return 2
        if self.p_0:
            return 1
        return 0

    def check_rates(self):
        """``reaction.py:155-176``."""
        if self.product == []:
            if self.p_0 is not None:
                raise ValueError(f"p_0 must be None, not {self.p_0} when no products are present.")
            if self.E_p is not None:
                raise ValueError(f"E_p must be None, not {self.E_p} when no products are present.")
        else:
            if self.p_0 is None:
                raise ValueError("p_0 cannot be None when reaction products are present.")
            if self.E_p is None:
                raise ValueError("E_p cannot be None when reaction products are present.")

    def reaction_term(self, temperature, reactant_concentrations=None, product_concentrations=None):
        """k(T) prod(reactants) - p(T) prod(products)  (``reaction.py:117-224``)."""
        self.check_rates()
        products = self.products
        k = self.k_0 * ufl.exp(-self.E_k / (k_B * temperature))
        mode = self.rate_mode()
        if mode == 2:
            p = self.p_0 * ufl.exp(-self.E_p / (k_B * temperature))
        elif mode == 1:
            p = self.p_0
        else:
            p = 0

        def conc(given, objs):
            if given is None:
                return [o.concentration for o in objs]
            assert len(given) == len(objs)
            return [o.concentration if g is None else g for g, o in zip(given, objs)]

        rc = conc(reactant_concentrations, self.reactant)
        pc = conc(product_concentrations, products)
        fwd = ufl.as_expr(rc[0])
        for c in rc[1:]:
            fwd = fwd * c
        if products:
            bwd = ufl.as_expr(pc[0])
            for c in pc[1:]:
                bwd = bwd * c
        else:
            bwd = 0
        return k * fwd - (p * bwd)
