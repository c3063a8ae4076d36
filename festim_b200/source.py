"""Volumetric sources (reference ``src/festim/source.py``)."""

from festim_b200.helpers import Value
from festim_b200.species import Species
from festim_b200.subdomain import VolumeSubdomain


class SourceBase:
    def __init__(self, value, volume):
        self.value = value
        self.volume = volume

    @property
    def value(self):
        return self._value

    @value.setter
    def value(self, value):
        self._value = value if isinstance(value, Value) else Value(value)

    @property
    def volume(self):
        return self._volume

    @volume.setter
    def volume(self, value):
        if not isinstance(value, VolumeSubdomain):
            raise TypeError("volume must be of type festim.VolumeSubdomain")
        self._volume = value


class ParticleSource(SourceBase):
    """``source.py:73-150``; the value may depend on species concentrations."""

    def __init__(self, value, volume, species, species_dependent_value=None):
        self.species = species
        self.species_dependent_value = species_dependent_value
        super().__init__(value, volume)

    @property
    def value(self):
        return self._value

    @value.setter
    def value(self, value):
        if isinstance(value, Value):
            self._value = value
        else:
            self._value = Value(value, species_dependent_value=self.species_dependent_value)

    @property
    def species(self):
        return self._species

    @species.setter
    def species(self, value):
        if not isinstance(value, Species):
            raise TypeError("species must be of type festim.Species")
        self._species = value


class HeatSource(SourceBase):
    def __init__(self, value, volume):
        super().__init__(value, volume)
        if self.value.temperature_dependent:
            raise ValueError("Heat source cannot be temperature dependent")
