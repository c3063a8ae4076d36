"""Problem base class: time loop and the solver seam.

Mirrors reference ``src/festim/problem.py``: ``define_meshtags_and_measures``
(84-116), ``define_boundary_conditions`` (118-130), ``get_petsc_options``
(132-168), ``create_solver`` (170-193), ``run`` (195-214) and ``iterate``
(216-255).  The seam is ``create_solver``: where the reference wraps the UFL
form in ``dolfinx.fem.petsc.NonlinearProblem`` (SNES newtonls + LU), this class
hands the form to a solver backend.  The default (and only product) backend is
``festim_b200.backend.B200Backend`` - hand-written sm_100a CUDA behind the
C ABI of ``include/festim_b200.h``; it raises if the library or a GPU is
missing (no CPU fallback).  The time loop only touches the same handful of
members as the reference: ``solver.solve()``, ``solver.solver
.getConvergedReason()``, ``solver.solver.getIterationNumber()``,
``u.x.array``, ``u_n.x.array``, ``t.value``, ``dt.value``.
"""

from __future__ import annotations

import warnings

import numpy as np

from festim_b200 import boundary_conditions as _bcs
from festim_b200 import fem, forms
from festim_b200 import mesh as _mesh
from festim_b200.subdomain import SurfaceSubdomain, VolumeSubdomain

_DEFAULT_PETSC_OPTS = {
    "snes_type": "newtonls",
    "snes_linesearch_type": "none",
    "snes_stol": np.sqrt(np.finfo(np.float64).eps) * 1e-2,
    "snes_divergence_tolerance": "PETSC_UNLIMITED",
    "ksp_type": "preonly",
    "pc_type": "lu",
    "snes_error_if_not_converged": True,
    "ksp_error_if_not_converged": True,
}


def get_default_petsc_options():
    return _DEFAULT_PETSC_OPTS.copy()


class DirichletBCForm:
    """What ``fem.dirichletbc`` returns in the reference: constrained dofs of
    one species and the object holding their values."""

    def __init__(self, species_index, vertices, value_fenics, block_size):
        self.species_index = species_index
        self.vertices = vertices
        self.value_fenics = value_fenics
        self.dofs = vertices.astype(np.int64) * block_size + species_index

    def values(self):
        v = self.value_fenics
        if isinstance(v, fem.Function):
            return v.x.array[self.vertices]
        return np.full(len(self.vertices), float(v.value))


class ProblemBase:
    def __init__(self, mesh=None, sources=None, exports=None, subdomains=None,
                 boundary_conditions=None, settings=None, petsc_options=None):
        self.mesh = mesh
        self.subdomains = subdomains or []
        self.boundary_conditions = boundary_conditions or []
        self.sources = sources or []
        self.exports = exports or []
        self.settings = settings
        self.dx = None
        self.ds = None
        self.function_space = None
        self.facet_meshtags = None
        self.volume_meshtags = None
        self.formulation = None
        self.bc_forms = []
        self.show_progress_bar = False
        self.progress_bar = None
        self.petsc_options = petsc_options
        self._timesteps = []
        self.solver = None
        #: solver backend; "b200" = the CUDA library (the only product path)
        self.solver_backend = "b200"

    @property
    def volume_subdomains(self):
        return [s for s in self.subdomains if isinstance(s, VolumeSubdomain)]

    @property
    def surface_subdomains(self):
        return [s for s in self.subdomains if isinstance(s, SurfaceSubdomain)]

    @property
    def dt(self):
        return self._dt

    @property
    def timesteps(self):
        return self._timesteps

    @property
    def facet_meshtags(self):
        return self._facet_meshtags

    @facet_meshtags.setter
    def facet_meshtags(self, value):
        if value is None or isinstance(value, _mesh.MeshTags):
            self._facet_meshtags = value
        else:
            raise TypeError("value must be of type dolfinx.mesh.MeshTags")

    @property
    def volume_meshtags(self):
        return self._volume_meshtags

    @volume_meshtags.setter
    def volume_meshtags(self, value):
        if value is None or isinstance(value, _mesh.MeshTags):
            self._volume_meshtags = value
        else:
            raise TypeError("value must be of type dolfinx.mesh.MeshTags")

    def define_meshtags_and_measures(self):
        if isinstance(self.mesh, _mesh.MeshFromXDMF):   # reference problem.py:88-91
            self.facet_meshtags = self.mesh.define_surface_meshtags()
            self.volume_meshtags = self.mesh.define_volume_meshtags()
        elif (isinstance(self.mesh, _mesh.Mesh) and self.facet_meshtags is None
                and self.volume_meshtags is None):
            self.facet_meshtags, self.volume_meshtags = self.mesh.define_meshtags(
                surface_subdomains=self.surface_subdomains,
                volume_subdomains=self.volume_subdomains,
                interfaces=getattr(self, "interfaces", None),
            )
        vol_ids = [vol.id for vol in self.volume_subdomains]
        if len(vol_ids) != len(np.unique(vol_ids)):
            raise ValueError("Volume ids are not unique")
        self.ds = forms.Measure("ds", self.facet_meshtags)
        self.dx = forms.Measure("dx", self.volume_meshtags)

    def define_boundary_conditions(self):
        self.bc_forms = []
        for bc in self.boundary_conditions:
            if isinstance(bc, _bcs.DirichletBCBase):
                form = self.create_dirichletbc_form(bc)
                if not bc.enforce_weakly:
                    self.bc_forms.append(form)

    def get_petsc_options(self):
        petsc_options = get_default_petsc_options()
        if self.petsc_options:
            petsc_options.update(self.petsc_options)
            if any(k in self.petsc_options for k in ("snes_atol", "snes_rtol", "snes_max_it")):
                warnings.warn(
                    "You have set one of the following PETSc options: snes_atol, "
                    "snes_rtol or snes_max_it. These options will be overwritten by "
                    "the values in festim.Settings (atol, rtol and max_iterations) to "
                    "ensure consistency between different versions of dolfinx. If you "
                    "want to set these options manually, please set them in "
                    "festim.Settings and not in the petsc_options dictionary."
                )
        petsc_options.update({
            "snes_atol": self.settings.atol,
            "snes_rtol": self.settings.rtol,
            "snes_max_it": self.settings.max_iterations,
        })
        return petsc_options

    def create_solver(self):
        """The seam (reference ``problem.py:170-193``)."""
        petsc_options = self.get_petsc_options()
        backend = self.solver_backend
        if backend == "b200":
            from festim_b200.backend import B200Backend

            backend = B200Backend()
        self.solver = backend.create_newton_solver(self, petsc_options)

    def _device_resident_stepping_ok(self):
        """True when nothing but t and dt changes between time steps and nobody looks at the
        fields in between: the whole loop can then run inside the library (fb_run_steps)."""
        import os

        st = self.settings.stepsize
        if os.environ.get("FB_NO_RUN_STEPS") or not hasattr(self.solver, "run_steps"):
            return False
        if getattr(self, "exports", None) or callable(self.settings.atol) or callable(self.settings.rtol):
            return False
        if callable(getattr(st, "max_stepsize", None)) or type(self).iterate is not ProblemBase.iterate:
            return False
        if getattr(self, "temperature_time_dependent", False) or isinstance(getattr(self, "temperature", None), fem.Function):
            return False
        if any(getattr(bc, "time_dependent", False) for bc in getattr(self, "boundary_conditions", [])):
            return False
        if any(src.value.explicit_time_dependent for src in getattr(self, "sources", [])):
            return False
        if any(getattr(bc, "temperature_dependent", False) and self.temperature_time_dependent
               for bc in getattr(self, "boundary_conditions", [])):
            return False
        return True

    def run(self):
        if self.settings.transient:
            if self._device_resident_stepping_ok():
                # device-resident stepping: same loop, same step-size rule, inside the library
                while self.t.value < self.settings.final_time:
                    times, its = self.solver.run_steps(self.settings.final_time)
                    self._timesteps.extend(times)
                    reason = self.solver.solver.getConvergedReason()
                    assert reason > 0, (
                        f"Non-linear solver did not converge. Reason code: {reason}. \n See https://petsc.org/release/manualpages/SNES/SNESConvergedReason/ for more information."
                    )
                    if not times:
                        break
                self.post_processing()
                return
            while self.t.value < self.settings.final_time:
                self.iterate()
        else:
            self.solver.solve()
            # the reference sets snes_error_if_not_converged / ksp_error_if_not_converged
            # (problem.py:287-288): PETSc raises from solve() instead of exporting a diverged field
            reason = self.solver.solver.getConvergedReason()
            if reason <= 0:
                raise RuntimeError(
                    f"Non-linear solver did not converge. Reason code: {reason}. \n See "
                    "https://petsc.org/release/manualpages/SNES/SNESConvergedReason/ for more information.")
            self.post_processing()

    def iterate(self):
        self._timesteps.append(float(self.t))
        if callable(self.settings.rtol):
            self.solver.rtol = self.settings.rtol(self.t.value)
        if callable(self.settings.atol):
            self.solver.atol = self.settings.atol(self.t.value)

        self.t.value += self.dt.value
        self.update_time_dependent_values()

        self.solver.solve()
        converged_reason = self.solver.solver.getConvergedReason()
        assert converged_reason > 0, (
            f"Non-linear solver did not converge. Reason code: {converged_reason}. \n See https://petsc.org/release/manualpages/SNES/SNESConvergedReason/ for more information."
        )
        nb_its = self.solver.solver.getIterationNumber()

        self.post_processing()
        # u_n <- u: on the device when the backend keeps the fields there (no host round trip)
        advance = getattr(self.solver, "copy_solution_to_previous", None)
        if advance is None or not advance():
            self.u_n.x.array[:] = self.u.x.array[:]

        if self.settings.stepsize.adaptive:
            self.dt.value = self.settings.stepsize.modify_value(
                value=self.dt.value, nb_iterations=nb_its, t=self.t.value
            )

    def update_time_dependent_values(self):
        t = float(self.t)
        for bc in self.boundary_conditions:
            if bc.time_dependent:
                bc.update(t=t)
        for source in self.sources:
            if source.value.explicit_time_dependent:
                source.value.update(t=t)
