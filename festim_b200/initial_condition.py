"""Initial conditions (reference ``src/festim/initial_condition.py``)."""

import numpy as np

from festim_b200 import fem, ufl
from festim_b200.species import Species


class InitialConditionBase:
    def __init__(self, value, volume):
        self.value = value
        self.volume = volume
        self.expr_fenics = None

    @property
    def volume(self):
        return self._volume

    @volume.setter
    def volume(self, value):
        from festim_b200.subdomain import VolumeSubdomain

        if not isinstance(value, VolumeSubdomain):   # reference initial_condition.py:59-64
            raise TypeError("volume must be of type festim.VolumeSubdomain")
        self._volume = value

    def _create(self, mesh, temperature=None):
        if isinstance(self.value, (int, float)):
            val = self.value
            self.expr_fenics = lambda x: np.full(x.shape[1], val)
        elif isinstance(self.value, fem.Function):
            self.expr_fenics = self.value
        elif callable(self.value):
            arguments = self.value.__code__.co_varnames
            kwargs = {}
            if "t" in arguments:
                raise ValueError("Initial condition cannot be a function of time.")
            if "x" in arguments:
                kwargs["x"] = ufl.SpatialCoordinate(mesh)
            if "T" in arguments and temperature is not None:
                kwargs["T"] = temperature
            self.expr_fenics = fem.Expression(self.value(**kwargs))


class InitialConcentration(InitialConditionBase):
    """``initial_condition.py:46-141``."""

    def __init__(self, value, volume, species):
        super().__init__(value=value, volume=volume)
        self.species = species

    @property
    def species(self):
        return self._species

    @species.setter
    def species(self, value):
        if not isinstance(value, Species):
            raise TypeError("species must be of type festim.Species")
        self._species = value

    def create_expr_fenics(self, mesh, temperature, function_space):
        self._create(mesh, temperature)


class InitialTemperature(InitialConditionBase):
    """``initial_condition.py:144-203``."""

    def create_expr_fenics(self, mesh, function_space):
        self._create(mesh)


def read_function_from_file(filename, name, timestamp, family="P", order=1, mesh=None):
    """Reference ``initial_condition.py:238-283``: the named field at ``timestamp`` from a restart
    file written by ``VTXSpeciesExport(checkpoint=True)`` (or any XDMF file of
    ``festim_b200.field_io.XDMFFile``), as a P1 function on the file's mesh or - when ``mesh`` is
    given - interpolated onto that mesh (identical meshes transfer exactly)."""
    from pathlib import Path

    from festim_b200 import field_io

    if order != 1 or family not in ("P", "CG", "Lagrange"):
        raise NotImplementedError("the B200 backend implements P1 elements only")
    path = Path(filename)
    if path.is_dir():
        path = path / "checkpoint.xdmf"
    rd = field_io.XDMFFile(path, "r")
    mesh_in = rd.read_mesh("mesh")
    u_in = fem.Function(fem.functionspace(mesh_in, ("P", 1)), name=name)
    values = rd.read_function(name, timestamp)
    if values.shape[0] != mesh_in.num_vertices:
        raise ValueError(f"{path}: {values.shape[0]} values for {mesh_in.num_vertices} vertices")
    u_in.x.array[:] = values
    if mesh is None:
        return u_in
    target = getattr(mesh, "mesh", mesh) if not hasattr(mesh, "coords") else mesh
    u = fem.Function(fem.functionspace(target, ("P", 1)), name=name)
    fem.interpolate_nonmatching(u, u_in)
    return u
