"""Heat transfer problem (reference ``src/festim/heat_transfer_problem.py``).

Scalar P1 heat equation rho cp dT/dt = div(lambda grad T) + Q; lambda, rho, cp
may be callables of T (reference :175-218).  Shares the solver seam, the time
loop and the device kernels with the hydrogen problem: T is lowered as a single
"species" whose diffusivity may depend on the solution.
"""

from __future__ import annotations

from festim_b200 import boundary_conditions as _bcs
from festim_b200 import exports as _exports
from festim_b200 import fem, forms, ufl
from festim_b200.helpers import as_fenics_constant, is_it_time_to_export
from festim_b200.problem import DirichletBCForm, ProblemBase
from festim_b200.source import HeatSource
from festim_b200.species import Species


class HeatTransferProblem(ProblemBase):
    def __init__(self, mesh=None, subdomains=None, initial_condition=None,
                 boundary_conditions=None, sources=None, exports=None, settings=None,
                 petsc_options=None):
        super().__init__(mesh=mesh, sources=sources, exports=exports, subdomains=subdomains,
                         boundary_conditions=boundary_conditions, settings=settings,
                         petsc_options=petsc_options)
        self.initial_condition = initial_condition
        self._unknown = Species("T")

    @property
    def sources(self):
        return self._sources

    @sources.setter
    def sources(self, value):
        if not all(isinstance(source, HeatSource) for source in value):
            raise TypeError("sources must be a list of festim.HeatSource objects")
        self._sources = value

    @property
    def boundary_conditions(self):
        return self._boundary_conditions

    @boundary_conditions.setter
    def boundary_conditions(self, value):
        if not all(isinstance(bc, (_bcs.FixedTemperatureBC, _bcs.HeatFluxBC)) for bc in value):
            raise TypeError(
                "boundary_conditions must be a list of festim.FixedTemperatureBC "
                "or festim.HeatFluxBC objects"
            )
        self._boundary_conditions = value

    def initialise(self):
        self.define_function_space()
        self.define_meshtags_and_measures()
        self.t = fem.Constant(self.mesh.mesh, 0.0)
        if self.settings.transient:
            self._dt = as_fenics_constant(self.settings.stepsize.initial_value, self.mesh.mesh)
        self.define_boundary_conditions()
        self.create_source_values_fenics()
        self.create_flux_values_fenics()
        self.create_initial_conditions()
        self.create_formulation()
        self.create_solver()
        self.initialise_exports()

    def define_function_space(self):
        mesh = self.mesh.mesh
        self.function_space = fem.FunctionSpace(mesh, "Lagrange", 1, bs=1)
        self.u = fem.Function(self.function_space, name="T")
        self.u_n = fem.Function(self.function_space, name="T_n")
        spe = self._unknown
        spe.solution = ufl.SpeciesRef(spe)
        spe.prev_solution = ufl.SpeciesRef(spe, prev=True)
        spe.post_processing_solution = self.u
        self.test_function = 0

    def create_dirichletbc_form(self, bc):
        bc.create_value(function_space=self.function_space, t=self.t)
        verts = bc.define_surface_subdomain_vertices(self.facet_meshtags, self.mesh.mesh)
        return DirichletBCForm(0, verts, bc.value_fenics, 1)

    def create_source_values_fenics(self):
        for source in self.sources:
            source.value.convert_input_value(function_space=self.function_space, t=self.t,
                                             up_to_ufl_expr=True)

    def create_flux_values_fenics(self):
        for bc in self.boundary_conditions:
            if isinstance(bc, _bcs.HeatFluxBC):
                bc.create_value_fenics(mesh=self.mesh.mesh, temperature=self._unknown.solution,
                                       t=self.t)

    def create_initial_conditions(self):
        if not self.initial_condition:
            return
        if not self.settings.transient:
            raise ValueError("Initial conditions can only be defined for transient simulations")
        self.initial_condition.create_expr_fenics(mesh=self.mesh.mesh,
                                                  function_space=self.function_space)
        self.u_n.interpolate(self.initial_condition.expr_fenics)

    def create_formulation(self):
        """Reference :175-218."""
        F = forms.Form()
        T = self._unknown
        u, u_n = T.solution, T.prev_solution
        for vol in self.volume_subdomains:
            lam = vol.material.thermal_conductivity
            meta = {}
            if callable(lam):
                lam = lam(u)
            else:
                meta = {"arrhenius": (float(lam), 0.0)}
            F += forms.DiffusionIntegral(T, lam, self.dx(vol.id), meta)
            if self.settings.transient:
                rho, cp = vol.material.density, vol.material.heat_capacity
                if callable(rho) or callable(cp):
                    rho = rho(u) if callable(rho) else rho
                    cp = cp(u) if callable(cp) else cp
                    F += forms.TestIntegral(T, rho * cp * ((u - u_n) / self.dt), self.dx(vol.id))
                else:
                    F += forms.TestIntegral(T, rho * cp * ((u - u_n) / self.dt), self.dx(vol.id),
                                            {"mass": float(rho) * float(cp)})
        for source in self.sources:
            F += forms.TestIntegral(T, -ufl.as_expr(source.value.fenics_object),
                                    self.dx(source.volume.id), {"source": source})
        for bc in self.boundary_conditions:
            if isinstance(bc, _bcs.HeatFluxBC):
                F += forms.TestIntegral(T, -ufl.as_expr(bc.value_fenics), self.ds(bc.subdomain.id),
                                        {"flux": bc})
        self.formulation = F

    def initialise_exports(self):
        """Reference ``heat_transfer_problem.py:220-236``."""
        from festim_b200 import field_io

        for export in self.exports:
            if isinstance(export, _exports.XDMFExport):
                raise NotImplementedError("XDMF export is not implemented yet for heat transfer problems")
            if isinstance(export, _exports.VTXTemperatureExport):
                export.writer = (field_io.VTUSeriesWriter(export.filename, self.mesh.mesh)
                                 if field_io.is_io_rank() else field_io.NullWriter())
            if isinstance(export, _exports.DerivedQuantity):
                export.t, export.data = [], []

    def post_processing(self):
        for export in self.exports:
            if hasattr(export, "times"):
                if not is_it_time_to_export(current_time=float(self.t), times=export.times):
                    continue
            if isinstance(export, _exports.SurfaceQuantity):
                raise NotImplementedError(
                    "SurfaceQuantity export is not implemented yet for heat transfer problems"
                )
            if isinstance(export, _exports.VTXTemperatureExport):   # reference :266-267
                export.writer.write(float(self.t), {"T": self.u.x.array})
