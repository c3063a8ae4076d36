"""Trap convenience class (reference ``src/festim/trap.py``)."""

from festim_b200.reaction import Reaction
from festim_b200.species import ImplicitSpecies, Species


class Trap(Species):
    """Expands to an immobile species, its implicit empty-site species and the
    trapping reaction (``trap.py:108-122``)."""

    def __init__(self, name, mobile_species, k_0, E_k, p_0, E_p, n, volume):
        super().__init__(name)
        self.mobile_species = mobile_species
        self.k_0 = k_0
        self.E_k = E_k
        self.p_0 = p_0
        self.E_p = E_p
        self.n = n
        self.volume = volume
        self.trapped_concentration = None
        self.reaction = None

    def create_species_and_reaction(self):
        self.trapped_concentration = Species(name=self.name, mobile=False)
        self.empty_trap_sites = ImplicitSpecies(n=self.n, others=[self.trapped_concentration])
        self.reaction = Reaction(
            reactant=[self.mobile_species, self.empty_trap_sites],
            product=self.trapped_concentration,
            k_0=self.k_0, E_k=self.E_k, p_0=self.p_0, E_p=self.E_p,
            volume=self.volume,
        )
