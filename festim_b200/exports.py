"""Derived quantities (reference ``src/festim/exports/``).

``SurfaceFlux`` = int -D grad(u).n ds (``surface_flux.py:39-60``) and
``TotalVolume`` = int u dx (``total_volume.py:23-33``) are on the hot path
 and are evaluated by the solver backend (device reduction
kernels); the remaining quantities are thin reductions over nodal arrays.
"""

from __future__ import annotations

import csv

import numpy as np

from festim_b200.species import Species
from festim_b200.subdomain import SurfaceSubdomain, VolumeSubdomain


class DerivedQuantity:
    def __init__(self, filename=None):
        self.filename = filename
        self.t = []
        self.data = []
        self.value = None
        self._first_time_export = True

    @property
    def filename(self):
        return self._filename

    @filename.setter
    def filename(self, value):
        if value is None:
            self._filename = None
            return
        if not isinstance(value, str):
            raise TypeError("filename must be of type str")
        if not value.endswith(".csv") and not value.endswith(".txt"):
            raise ValueError("filename must end with .csv or .txt")
        self._filename = value

    def write(self, t):
        """CSV ``t(s),<title>`` (``derived_quantity.py:47-60``)."""
        if self.filename is None:
            return
        if self._first_time_export:
            with open(self.filename, mode="w+", newline="") as f:
                csv.writer(f).writerow(["t(s)", f"{self.title}"])
            self._first_time_export = False
        with open(self.filename, mode="a", newline="") as f:
            csv.writer(f).writerow([t, self.value])


class SurfaceQuantity(DerivedQuantity):
    def __init__(self, field, surface, filename=None):
        super().__init__(filename=filename)
        self.field = field
        self.surface = surface
        self.D = None

    @property
    def surface(self):
        return self._surface

    @surface.setter
    def surface(self, value):
        if not isinstance(value, (int, SurfaceSubdomain)) or isinstance(value, bool):
            raise TypeError("surface should be an int or F.SurfaceSubdomain")
        self._surface = value

    @property
    def field(self):
        return self._field

    @field.setter
    def field(self, value):
        if not isinstance(value, (Species, str)):
            raise TypeError("field must be of type F.Species or str")
        self._field = value


class VolumeQuantity(DerivedQuantity):
    def __init__(self, field, volume, filename=None):
        super().__init__(filename=filename)
        self.field = field
        self.volume = volume

    @property
    def volume(self):
        return self._volume

    @volume.setter
    def volume(self, value):
        if not isinstance(value, (int, VolumeSubdomain)) or isinstance(value, bool):
            raise TypeError("volume should be an int or F.VolumeSubdomain")
        self._volume = value

    @property
    def field(self):
        return self._field

    @field.setter
    def field(self, value):
        if not isinstance(value, (Species, str)):
            raise TypeError("field must be of type F.Species or str")
        self._field = value


def _id(sub):
    return sub if isinstance(sub, int) else sub.id


class SurfaceFlux(SurfaceQuantity):
    @property
    def title(self):
        return f"{self.field.name} flux surface {_id(self.surface)}"

    def compute(self, problem):
        self.value = problem.solver.surface_flux(
            problem.species.index(self.field), _id(self.surface)
        )
        self.data.append(self.value)


class TotalSurface(SurfaceQuantity):
    @property
    def title(self):
        return f"Total {self.field.name} surface {_id(self.surface)}"

    def compute(self, problem):
        self.value = problem.solver.total_surface(
            problem.species.index(self.field), _id(self.surface)
        )
        self.data.append(self.value)


class TotalVolume(VolumeQuantity):
    @property
    def title(self):
        return f"Total {self.field.name} volume {_id(self.volume)}"

    def compute(self, problem):
        self.value = problem.solver.total_volume(
            problem.species.index(self.field), _id(self.volume)
        )
        self.data.append(self.value)


class AverageVolume(VolumeQuantity):
    @property
    def title(self):
        return f"Average {self.field.name} volume {_id(self.volume)}"

    def compute(self, problem):
        total = problem.solver.total_volume(problem.species.index(self.field), _id(self.volume))
        self.value = total / problem.solver.measure_volume(_id(self.volume))
        self.data.append(self.value)


class AverageSurface(SurfaceQuantity):
    @property
    def title(self):
        return f"Average {self.field.name} surface {_id(self.surface)}"

    def compute(self, problem):
        total = problem.solver.total_surface(problem.species.index(self.field), _id(self.surface))
        self.value = total / problem.solver.measure_surface(_id(self.surface))
        self.data.append(self.value)


def _vertices_of(problem, sub, volume):
    m = problem.mesh.mesh
    if volume:
        cells = problem.volume_meshtags.find(_id(sub))
        return np.unique(m.cells[cells].ravel())
    facets = problem.facet_meshtags.find(_id(sub))
    return np.unique(m.bfacets[facets].ravel())


class MaximumVolume(VolumeQuantity):
    @property
    def title(self):
        return f"Maximum {self.field.name} volume {_id(self.volume)}"

    def compute(self, problem):
        v = _vertices_of(problem, self.volume, True)
        self.value = float(np.max(self.field.post_processing_solution.x.array[v]))
        self.data.append(self.value)


class MinimumVolume(VolumeQuantity):
    @property
    def title(self):
        return f"Minimum {self.field.name} volume {_id(self.volume)}"

    def compute(self, problem):
        v = _vertices_of(problem, self.volume, True)
        self.value = float(np.min(self.field.post_processing_solution.x.array[v]))
        self.data.append(self.value)


class MaximumSurface(SurfaceQuantity):
    @property
    def title(self):
        return f"Maximum {self.field.name} surface {_id(self.surface)}"

    def compute(self, problem):
        v = _vertices_of(problem, self.surface, False)
        self.value = float(np.max(self.field.post_processing_solution.x.array[v]))
        self.data.append(self.value)


class MinimumSurface(SurfaceQuantity):
    @property
    def title(self):
        return f"Minimum {self.field.name} surface {_id(self.surface)}"

    def compute(self, problem):
        v = _vertices_of(problem, self.surface, False)
        self.value = float(np.min(self.field.post_processing_solution.x.array[v]))
        self.data.append(self.value)


class Profile1DExport:
    """Concentration profile of a 1D run (``exports/profile_1d.py``)."""

    def __init__(self, field, subdomain=None, times=None):
        self.field = field
        self.subdomain = subdomain
        self.times = times
        self.x = None
        self.data = []
        self.t = []


class CustomQuantity(DerivedQuantity):
    """Integral of a user expression over a volume or surface subdomain (reference
    ``exports/custom_quantity.py``).  ``expr(**kwargs)`` receives the species by name, ``n``
    (facet normal), ``t``, ``T``, ``x``, ``D_<species>`` and ``D``; the expression is compiled
    with the problem's other expressions and integrated on the device (``fb_integrate``)."""

    def __init__(self, expr, subdomain, title="Custom Quantity", filename=None):
        super().__init__(filename=filename)
        self.expr = expr
        self.subdomain = subdomain
        self._title = title
        self.ufl_expr = None

    @property
    def title(self):
        return self._title

    def compute(self, problem):
        if self.ufl_expr is None:
            raise ValueError("The UFL expression has not been evaluated yet.")
        self.value = problem.solver.integrate_custom(self)
        self.data.append(self.value)


# ---------------------------------------------------------------------------------------
# Field output (reference exports/vtx.py, exports/xdmf.py).  Writing ADIOS2 / XDMF+HDF5 files
# is file I/O outside the hot path: the classes exist so that scripts
# written for the reference run unchanged - arguments, file-name handling and `times`
# (which become milestones of the time stepping) behave as in the reference; nothing is
# written and a warning says so once.
# ---------------------------------------------------------------------------------------
class ExportBaseClass:
    """``vtx.py:16-54``: file name with the writer's extension, optional export times."""

    _warned = False

    def __init__(self, filename, ext, times=None):
        import warnings
        from pathlib import Path

        name = Path(filename)
        if name.suffix != ext:
            warnings.warn(f"Filename {filename} does not have {ext} extension, adding it.")
            name = name.with_suffix(ext)
        self._filename = name
        self.times = sorted(times) if times else times

    @property
    def filename(self):
        return self._filename

    def write(self, t):
        import warnings

        if not ExportBaseClass._warned:
            warnings.warn(f"{type(self).__name__}: field output (VTX / XDMF) is not written by the B200 "
                          "backend; use derived quantities or read `species.post_processing_solution`")
            ExportBaseClass._warned = True


def _species_list(value):
    fields = value if isinstance(value, list) else [value]
    for f in fields:
        if not isinstance(f, (Species, str)):
            raise TypeError("field must be of type festim.Species or a list of festim.Species or str")
    return fields


class XDMFExport(ExportBaseClass):
    """``xdmf.py:12-74``."""

    def __init__(self, filename, field):
        super().__init__(filename, ".xdmf")
        self.field = _species_list(field)


class VTXSpeciesExport(ExportBaseClass):
    """``vtx.py:83-180``."""

    def __init__(self, filename, field, subdomain=None, checkpoint=False, times=None):
        super().__init__(filename, ".bp", times)
        self.field = _species_list(field)
        self._subdomain = subdomain
        self._checkpoint = checkpoint


class VTXTemperatureExport(ExportBaseClass):
    """``vtx.py:57-80``."""

    def __init__(self, filename, times=None):
        super().__init__(filename, ".bp", times)
