"""Derived quantities (reference ``src/festim/exports/``).

``SurfaceFlux`` = int -D grad(u).n ds (``surface_flux.py:39-60``) and
``TotalVolume`` = int u dx (``total_volume.py:23-33``) are on the hot path
 and are evaluated by the solver backend (device reduction
kernels); the remaining quantities are thin reductions over nodal arrays.
"""

from __future__ import annotations

import csv

import numpy as np

from festim_b200.species import ImplicitSpecies, Species
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
        self.times = times.copy() if times is not None else None   # consumed by is_it_time_to_export
        self.x = None
        self._dofs = None
        self._sort_coords = None
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
# Field output (reference exports/vtx.py, exports/xdmf.py).  ADIOS2 and HDF5 do not
# exist in this image: `festim_b200/field_io.py` writes XDMF 3 with raw side-car data instead of
# HDF5, and a directory of VTK XML files (+ series.pvd) at the `.bp` path instead of an ADIOS2
# BP5 stream.  Arguments, file-name handling and `times` (which become milestones of the time
# stepping) behave as in the reference.
# ---------------------------------------------------------------------------------------
class ExportBaseClass:
    """``vtx.py:16-54``: file name with the writer's extension, optional export times."""

    def __init__(self, filename, ext, times=None):
        import warnings
        from pathlib import Path

        name = Path(filename)
        if name.suffix != ext:
            warnings.warn(f"Filename {filename} does not have {ext} extension, adding it.")
            name = name.with_suffix(ext)
        self._filename = name
        self.times = sorted(times) if times else times
        self.writer = None

    @property
    def filename(self):
        return self._filename


def _species_list(value, allow_str=True):
    kinds = (Species, str) if allow_str else (Species,)
    if isinstance(value, list):
        for f in value:
            if not isinstance(f, kinds):
                raise TypeError("field must be of type festim.Species or a list of festim.Species or str")
        return value
    if isinstance(value, Species):
        return [value]
    raise TypeError(f"field must be of type festim.Species or a list of festim.Species or str, got {type(value)}.")


class XDMFExport(ExportBaseClass):
    """``xdmf.py:12-78``: mesh once, then every field at every export time."""

    def __init__(self, filename, field):
        self._writer = None
        super().__init__(filename, ".xdmf")
        self.field = field
        self._mesh_written = False

    @property
    def field(self):
        return self._field

    @field.setter
    def field(self, value):
        if isinstance(value, list):
            for element in value:
                if not isinstance(element, Species):
                    raise TypeError(f"Each element in the list must be a species, got {type(element)}.")
            self._field = value
        elif isinstance(value, Species):
            self._field = [value]
        else:
            raise TypeError(
                f"field must be of type festim.Species or a list of festim.Species, got {type(value)}.")

    def define_writer(self, comm=None):
        from festim_b200.field_io import NullWriter, XDMFFile, is_io_rank

        self._writer = XDMFFile(self.filename, "w") if is_io_rank() else NullWriter()
        self._mesh_written = False

    def write(self, t):
        if self._writer is None:
            self.define_writer()
        if not self._mesh_written:
            self._writer.write_mesh(self.field[0].post_processing_solution.function_space.mesh)
            self._mesh_written = True
        for field in self.field:
            self._writer.write_function((field.name, field.post_processing_solution.x.array), t)

    def __del__(self):
        if getattr(self, "_writer", None) is not None:
            self._writer.close()


class VTXSpeciesExport(ExportBaseClass):
    """``vtx.py:83-180``.  ``checkpoint=True`` writes a restart file (mesh + every field at every
    export time) that ``read_function_from_file`` reads back - an XDMF document inside the ``.bp``
    directory instead of adios4dolfinx's ADIOS2 stream."""

    def __init__(self, filename, field, subdomain=None, checkpoint=False, times=None):
        super().__init__(filename, ".bp", times)
        self.field = field
        self._subdomain = subdomain
        self._checkpoint = checkpoint

    @property
    def field(self):
        return self._field

    @field.setter
    def field(self, value):
        self._field = _species_list(value)

    def get_functions(self):
        """``vtx.py:153-180`` for the continuous ("legacy") problem: the whole fields."""
        return [field.post_processing_solution for field in self._field]

    def write(self, t):
        if self._checkpoint:   # reference hydrogen_transport_problem.py:1069-1076
            for f in self._field:
                self.writer.write_function((f.name, f.post_processing_solution.x.array), t)
        else:
            self.writer.write(t, {f.name: f.post_processing_solution.x.array for f in self._field})


class VTXTemperatureExport(ExportBaseClass):
    """``vtx.py:57-80``."""

    def __init__(self, filename, times=None):
        super().__init__(filename, ".bp", times)


class CustomFieldExport(ExportBaseClass):
    """``vtx.py:183-344``: nodal (P1) values of a user expression of ``t``, ``x``, ``T`` and of the
    species named in ``species_dependent_value``, written at every export time."""

    def __init__(self, filename, expression, species_dependent_value=None, times=None, subdomain=None,
                 checkpoint=False):
        super().__init__(filename=filename, times=times, ext=".bp")
        self.expression = expression
        self.species_dependent_value = species_dependent_value or {}
        self.checkpoint = checkpoint
        self.subdomain = subdomain
        self.function = None
        self.dolfinx_expression = None

    @property
    def mixed_domain(self):
        return False   # continuous problem only (SURVEY 8f1 is out of scope)

    def set_dolfinx_expression(self, temperature, time):
        """``vtx.py:261-309``."""
        import inspect

        from festim_b200 import fem, ufl

        arguments = inspect.signature(self.expression).parameters
        kwargs = {}
        if "t" in arguments:
            kwargs["t"] = time
        if "x" in arguments:
            kwargs["x"] = ufl.SpatialCoordinate(self.function.function_space.mesh)
        if "T" in arguments:
            kwargs["T"] = temperature
        for arg in arguments:
            if arg in self.species_dependent_value:
                kwargs[arg] = self._get_species_function(self.species_dependent_value[arg])
            if arguments[arg].kind in (inspect.Parameter.VAR_KEYWORD, inspect.Parameter.VAR_POSITIONAL):
                continue
            assert kwargs.get(arg) is not None, f"Argument {arg} not found in species_dependent_value"
        self.dolfinx_expression = fem.Expression(self.expression(**kwargs))

    def _get_species_function(self, spe):
        if isinstance(spe, ImplicitSpecies):
            return spe.concentration
        return spe.post_processing_solution

    def update(self, species):
        """Interpolate the expression at the vertices (``hydrogen_transport_problem.py:1087-1090``).
        Symbolic concentrations (an implicit species' ``n - sum(others)``) read the collapsed
        post-processing fields."""
        from festim_b200 import fem, ufl

        mesh = self.function.function_space.mesh
        env = fem.nodal_env(mesh, np.arange(mesh.num_vertices))
        env["species"] = lambda spe: spe.post_processing_solution.x.array
        vals = ufl.evaluate(self.dolfinx_expression.expr, env)
        self.function.x.array[:] = np.broadcast_to(np.asarray(vals, dtype=np.float64), (mesh.num_vertices,))

    def write(self, t):
        self.writer.write(t, {self.function.name or "f": self.function.x.array})


class ReactionRateExport(CustomFieldExport):
    """``vtx.py:347-440``: ``k(T) prod(reactants)``, ``p(T) prod(products)`` or their difference."""

    def __init__(self, reaction, filename, direction="both", times=None, subdomain=None, checkpoint=False):
        import inspect

        from festim_b200 import ufl
        from festim_b200.constants import k_B

        if direction not in ("forward", "backward", "both"):
            raise ValueError("direction must be 'forward', 'backward' or 'both'")
        products = reaction.product if isinstance(reaction.product, list) else [reaction.product]
        reactant_names = [r.name for r in reaction.reactant]
        product_names = [p.name for p in products]

        def expression(T, **kwargs):
            k = reaction.k_0 * ufl.exp(-reaction.E_k / (k_B * T))
            if reaction.p_0 and reaction.E_p:
                p = reaction.p_0 * ufl.exp(-reaction.E_p / (k_B * T))
            elif reaction.p_0:
                p = reaction.p_0
            else:
                p = 0.0
            forward = k
            for name in reactant_names:
                forward = forward * kwargs[name]
            backward = ufl.as_expr(p)
            for name in product_names:
                backward = backward * kwargs[name]
            if direction == "forward":
                return forward
            if direction == "backward":
                return backward
            return forward - backward

        self.override_signature(expression, reactant_names, product_names)
        self.reaction = reaction
        self.direction = direction
        super().__init__(filename=filename, expression=expression,
                         species_dependent_value={spe.name: spe for spe in reaction.reactant + products},
                         times=times, subdomain=subdomain, checkpoint=checkpoint)

    def override_signature(self, expression, reactant_names, product_names):
        """``vtx.py:419-451``: the generated expression takes ``T`` and one argument per reactant and
        product, which is what ``set_dolfinx_expression`` looks up."""
        import inspect

        params = [inspect.Parameter("T", inspect.Parameter.POSITIONAL_OR_KEYWORD)]
        for name in dict.fromkeys(reactant_names + product_names):
            params.append(inspect.Parameter(name, inspect.Parameter.POSITIONAL_OR_KEYWORD))
        expression.__signature__ = inspect.Signature(params)
        assert inspect.signature(expression).parameters.keys() == {"T", *reactant_names, *product_names}, (
            "The expression for the reaction rate is automatically generated based on the reaction "
            "provided. The arguments of the expression must be T (temperature) and the names of the "
            f"reactants and products. The current expression has arguments "
            f"{inspect.signature(expression).parameters.keys()} but should have arguments T and "
            f"{reactant_names + product_names}.")
