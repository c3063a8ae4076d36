"""Physical constants (reference ``src/festim/__init__.py:10-12``)."""

R = 8.314462618  # Gas constant J.mol-1.K-1
k_B = 8.6173303e-5  # Boltzmann constant eV.K-1
k_B_SI = 1.380649e-23  # Boltzmann constant J.K-1
