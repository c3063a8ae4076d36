"""Hydrogen transport problem (continuous formulation).

Mirrors reference ``src/festim/hydrogen_transport_problem.py:58-1156``: the
``initialise`` pipeline (314-344), temperature handling (364-425), species
functions (672-751), strong Dirichlet BCs (769-807), initial conditions
(840-868), the weak form (870-997), per-step updates (999-1034) and
post-processing (1036-1156).  Unknowns live in one node-major blocked vector
``u.x.array[v * ns + s]``.
"""

from __future__ import annotations

import warnings

import numpy as np

from festim_b200 import boundary_conditions as _bcs
from festim_b200.mesh import CoordinateSystem
from festim_b200 import exports as _exports
from festim_b200 import fem, forms, ufl
from festim_b200 import species as _species
from festim_b200.constants import k_B
from festim_b200.helpers import as_fenics_constant, is_it_time_to_export
from festim_b200.problem import DirichletBCForm, ProblemBase
from festim_b200.source import ParticleSource


class HydrogenTransportProblem(ProblemBase):
    def __init__(self, mesh=None, subdomains=None, species=None, reactions=None,
                 temperature=None, sources=None, initial_conditions=None,
                 boundary_conditions=None, settings=None, exports=None, traps=None,
                 advection_terms=None, petsc_options=None, element_immobile="CG"):
        super().__init__(mesh=mesh, sources=sources, exports=exports, subdomains=subdomains,
                         boundary_conditions=boundary_conditions, settings=settings,
                         petsc_options=petsc_options)
        self.species = species or []
        self.temperature = temperature
        self.reactions = reactions or []
        self.initial_conditions = initial_conditions or []
        self.traps = traps or []
        self.advection_terms = advection_terms or []
        self.temperature_fenics = None
        self.element_immobile = element_immobile
        self._species_to_D_global = None

    # ---- properties ----------------------------------------------------------
    @property
    def temperature(self):
        return self._temperature

    @temperature.setter
    def temperature(self, value):
        if value is None or isinstance(value, (float, int, fem.Constant, fem.Function)) or callable(value):
            self._temperature = value
        else:
            raise TypeError("Value must be a float, int, fem.Constant, fem.Function, or callable")

    @property
    def temperature_fenics(self):
        return self._temperature_fenics

    @temperature_fenics.setter
    def temperature_fenics(self, value):
        if value is not None and not isinstance(value, (fem.Constant, fem.Function)):
            raise TypeError("Value must be a fem.Constant or fem.Function")
        self._temperature_fenics = value

    @property
    def temperature_time_dependent(self):
        """False for a ``fem.Function`` temperature even when it changes every
        step (coupled runs)."""
        if self.temperature is None or isinstance(self.temperature, (fem.Constant, fem.Function)):
            return False
        if callable(self.temperature):
            return "t" in self.temperature.__code__.co_varnames
        return False

    @property
    def species(self):
        return self._species

    @species.setter
    def species(self, value):
        for spe in value:
            if not isinstance(spe, _species.Species):
                raise TypeError(
                    f"elements of species must be of type festim.Species not {type(spe)}"
                )
        self._species = value

    @property
    def element_immobile(self):
        return self._element_immobile

    @element_immobile.setter
    def element_immobile(self, value):
        allowed_values = ["DG", "CG", "P"]
        if value not in allowed_values:
            raise ValueError(f"element_immobile should be in {allowed_values}")
        if value == "DG":
            raise NotImplementedError("DG immobile species are out of scope (SURVEY section 2)")
        self._element_immobile = "CG"

    @property
    def _unpacked_bcs(self):
        """All boundary conditions, the fluxes of surface reactions included (reference :289-298)."""
        out = []
        for bc in self.boundary_conditions:
            if isinstance(bc, _bcs.SurfaceReactionBC):
                out.extend(bc.flux_bcs)
            else:
                out.append(bc)
        return out

    # ---- initialise pipeline (reference :314-344) -----------------------------
    def initialise(self):
        self.create_species_from_traps()
        self.define_function_spaces(element_degree=self.settings.element_degree)
        self.define_meshtags_and_measures()
        self.assign_functions_to_species()

        self.t = fem.Constant(self.mesh.mesh, 0.0)
        if self.settings.transient:
            self._dt = as_fenics_constant(self.settings.stepsize.initial_value, self.mesh.mesh)

        self.create_implicit_species_value_fenics()
        self.define_temperature()
        self.define_boundary_conditions()
        self.convert_source_input_values_to_fenics_objects()
        self.convert_advection_term_to_fenics_objects()
        self.create_flux_values_fenics()
        self.create_initial_conditions()
        self.create_formulation()
        self.prepare_custom_quantities()   # their expressions are compiled with the problem's
        self.create_solver()
        self.initialise_exports()

    def create_implicit_species_value_fenics(self):
        for reaction in self.reactions:
            for reactant in reaction.reactant:
                if isinstance(reactant, _species.ImplicitSpecies):
                    reactant.create_value_fenics(mesh=self.mesh.mesh, t=self.t)

    def create_species_from_traps(self):
        # appends on every call, as the reference does
        for trap in self.traps:
            trap.create_species_and_reaction()
            self.species.append(trap.trapped_concentration)
            self.reactions.append(trap.reaction)

    def define_temperature(self):
        """Reference :364-425: float / f(t) -> Constant; anything with x -> P1 field."""
        mesh = self.mesh.mesh
        if self.temperature is None:
            raise ValueError("the temperature attribute needs to be defined")
        if isinstance(self.temperature, (float, int)):
            self.temperature_fenics = as_fenics_constant(self.temperature, mesh)
        elif isinstance(self.temperature, (fem.Constant, fem.Function)):
            self.temperature_fenics = self.temperature
        elif callable(self.temperature):
            arguments = self.temperature.__code__.co_varnames
            if "t" in arguments and "x" not in arguments:
                val = self.temperature(t=float(self.t))
                if not isinstance(val, (float, int)):
                    raise ValueError(
                        f"self.temperature should return a float or an int, not {type(val)} "
                    )
                self.temperature_fenics = as_fenics_constant(val, mesh)
            else:
                # its own P1 space, as the reference (:401-411): define_temperature does not
                # need define_function_spaces to have run
                self.temperature_fenics = fem.Function(fem.FunctionSpace(mesh, "Lagrange", 1),
                                                       name="temperature")
                kwargs = {}
                if "t" in arguments:
                    kwargs["t"] = self.t
                if "x" in arguments:
                    kwargs["x"] = ufl.SpatialCoordinate(mesh)
                self.temperature_expr = fem.Expression(self.temperature(**kwargs))
                self.temperature_fenics.interpolate(self.temperature_expr)

    def define_function_spaces(self, element_degree=1):
        mesh = self.mesh.mesh
        ns = len(self.species)
        self.function_space = fem.FunctionSpace(mesh, "Lagrange", 1, bs=ns)
        self.V_CG_1 = fem.FunctionSpace(mesh, "Lagrange", 1)
        self.V_DG_0 = fem.FunctionSpace(mesh, "DG", 0)
        self.V_DG_1 = fem.FunctionSpace(mesh, "DG", 1)
        self.u = fem.Function(self.function_space, name="u")
        self.u_n = fem.Function(self.function_space, name="u_n")

    def assign_functions_to_species(self):
        for idx, spe in enumerate(self.species):
            spe.sub_function_space = self.function_space.sub(idx)
            spe.sub_function = self.u.sub(idx)
            spe.post_processing_solution = fem.Function(self.V_CG_1, name=spe.name)
            spe.collapsed_function_space, spe.map_sub_to_main_solution = (
                self.function_space.sub(idx).collapse()
            )
            spe.solution = ufl.SpeciesRef(spe)
            spe.prev_solution = ufl.SpeciesRef(spe, prev=True)
            spe.test_function = idx

    def define_boundary_conditions(self):
        for bc in self._unpacked_bcs:
            if isinstance(bc.species, str):
                bc.species = _species.find_species_from_name(bc.species, self.species)
            if isinstance(bc, _bcs.ParticleFluxBC):
                bc.create_value_fenics(mesh=self.mesh.mesh, temperature=self.temperature_fenics,
                                       t=self.t)
        super().define_boundary_conditions()

    def create_dirichletbc_form(self, bc):
        """Reference :769-807."""
        bc.create_value(temperature=self.temperature_fenics,
                        function_space=bc.species.collapsed_function_space, t=self.t)
        verts = bc.define_surface_subdomain_vertices(self.facet_meshtags, self.mesh.mesh)
        return DirichletBCForm(self.species.index(bc.species), verts, bc.value_fenics,
                               len(self.species))

    def convert_source_input_values_to_fenics_objects(self):
        for source in self.sources:
            if isinstance(source, ParticleSource):
                source.value.convert_input_value(function_space=self.function_space, t=self.t,
                                                 temperature=self.temperature_fenics,
                                                 up_to_ufl_expr=True)

    def convert_advection_term_to_fenics_objects(self):
        for advec_term in self.advection_terms:
            advec_term.velocity.convert_input_value(function_space=self.function_space, t=self.t)

    def create_flux_values_fenics(self):
        for bc in self._unpacked_bcs:
            if isinstance(bc, _bcs.ParticleFluxBC):
                bc.create_value_fenics(mesh=self.mesh.mesh, temperature=self.temperature_fenics,
                                       t=self.t)

    def create_initial_conditions(self):
        """Reference :840-868: nodal interpolation into u_n on the tagged cells."""
        if len(self.initial_conditions) > 0 and not self.settings.transient:
            raise ValueError("Initial conditions can only be defined for transient simulations")
        for condition in self.initial_conditions:
            fs = condition.species.collapsed_function_space if callable(condition.value) else None
            condition.create_expr_fenics(mesh=self.mesh.mesh, temperature=self.temperature_fenics,
                                         function_space=fs)
            entities = self.volume_meshtags.find(condition.volume.id)
            idx = self.species.index(condition.species)
            self.u_n.sub(idx).interpolate(condition.expr_fenics, cells0=entities)

    # ---- the weak form (reference :870-997) -----------------------------------
    def create_formulation(self):
        F = forms.Form()
        T = self.temperature_fenics

        for spe in self.species:
            u, u_n = spe.solution, spe.prev_solution
            for vol in self.volume_subdomains:
                if spe.mobile:
                    D = vol.material.get_diffusion_coefficient(self.mesh.mesh, T, spe)
                    meta = {}
                    if not vol.material.D:
                        meta = {"arrhenius": (vol.material.get_D_0(spe), vol.material.get_E_D(spe))}
                    F += forms.DiffusionIntegral(spe, D, self.dx(vol.id), meta)
                    # cylindrical / spherical coordinates (reference :886-908):
                    #   r^m D grad(u).grad(v / r^m) = D grad(u).grad(v) - (m D / r) d_r u v,
                    # m = 1 (cylindrical), 2 (spherical), r = x[0]: the cartesian term above
                    # plus a metric term that depends on the gradient of the unknown
                    csys = self.mesh.coordinate_system
                    if csys != CoordinateSystem.CARTESIAN:
                        m = {CoordinateSystem.CYLINDRICAL: 1.0, CoordinateSystem.SPHERICAL: 2.0}[csys]
                        r = ufl.SpatialCoordinate(self.mesh.mesh)[0]
                        # UFL's degree estimate of the reference's expression (a
                        # quotient counts as the sum of degrees, grad lowers by one): r^m . (D u')
                        # . grad(v / r^m) -> m + 0 + (1 + m - 1) = 2 m, i.e. 2 (cylindrical) and 4
                        # (spherical); the metric term as written here estimates 2 for both
                        F += forms.TestIntegral(spe, -(m * ufl.as_expr(D) / r) * ufl.SpeciesGrad(spe, 0),
                                                self.dx(vol.id), {"metric": m, "extra_degree": 2 * int(m) - 2})
                if self.settings.transient:
                    F += forms.TestIntegral(spe, (u - u_n) / self.dt, self.dx(vol.id),
                                            {"mass": 1.0})

        for reaction in self.reactions:
            term = reaction.reaction_term(T)
            for reactant in reaction.reactant:
                if isinstance(reactant, _species.Species):
                    F += forms.TestIntegral(reactant, term, self.dx(reaction.volume.id),
                                            {"reaction": reaction, "sign": 1.0})
            for product in reaction.products:
                F += forms.TestIntegral(product, -term, self.dx(reaction.volume.id),
                                        {"reaction": reaction, "sign": -1.0})

        for source in self.sources:
            F += forms.TestIntegral(source.species, -ufl.as_expr(source.value.fenics_object),
                                    self.dx(source.volume.id), {"source": source})

        for bc in self._unpacked_bcs:  # surface reactions contribute one flux per reactant (:949-956)
            if isinstance(bc, _bcs.ParticleFluxBC):
                F += forms.TestIntegral(bc.species, -ufl.as_expr(bc.value_fenics),
                                        self.ds(bc.subdomain.id), {"flux": bc})

        for bc in self.boundary_conditions:  # weakly enforced Dirichlet conditions (:956-966)
            if isinstance(bc, _bcs.FixedConcentrationBC) and bc.enforce_weakly:
                vol = self.volume_subdomain_of_surface(bc.subdomain)
                D = vol.material.get_diffusion_coefficient(self.mesh.mesh, T, bc.species)
                for itg in bc.weak_formulation(bc.species, bc.species.solution, self.ds, D, self.mesh.mesh):
                    F += itg

        for adv_term in self.advection_terms:
            for spe in adv_term.species:
                F += forms.AdvectionIntegral(spe, adv_term.velocity.fenics_object,
                                             self.dx(adv_term.subdomain.id))

        if not self.settings.transient:
            # immobile species in volumes without any reaction: c = 0 (:980-997)
            for spe in self.species:
                if not spe.mobile:
                    not_defined = list(self.volume_subdomains)
                    for vol in self.volume_subdomains:
                        for reaction in self.reactions:
                            if vol == reaction.volume and vol in not_defined:
                                not_defined.remove(vol)
                    for vol in not_defined:
                        F += forms.TestIntegral(spe, spe.solution, self.dx(vol.id), {"filler": 1.0})

        self.formulation = F

    def volume_subdomain_of_surface(self, surface):
        """Volume subdomain the tagged boundary facets of ``surface`` belong to
        (reference :595-626: deduced from the facet -> cell connectivity and the tags)."""
        mesh = self.mesh.mesh
        facets = self.facet_meshtags.find(surface.id)
        if len(facets) == 0:
            raise ValueError(f"surface subdomain {surface.id} has no facets")
        ctag = np.zeros(mesh.num_cells, dtype=np.int64)
        ctag[self.volume_meshtags.indices] = self.volume_meshtags.values
        ids = np.unique(ctag[mesh.bfacet_cell[facets]])
        if len(ids) != 1:
            raise ValueError(f"surface subdomain {surface.id} touches several volume subdomains: {ids}")
        for vol in self.volume_subdomains:
            if vol.id == ids[0]:
                return vol
        raise ValueError(f"surface subdomain {surface.id} cannot be mapped to a volume subdomain")

    # ---- per-step updates (reference :999-1042) --------------------------------
    def update_time_dependent_values(self):
        super().update_time_dependent_values()
        t = float(self.t)
        for reaction in self.reactions:
            for reactant in reaction.reactant:
                if isinstance(reactant, _species.ImplicitSpecies):
                    reactant.update_density(t=t)

        if isinstance(self.temperature, fem.Function) or self.temperature_time_dependent:
            for bc in self._unpacked_bcs:
                if isinstance(bc, (_bcs.FixedConcentrationBC, _bcs.ParticleFluxBC)):
                    if bc.temperature_dependent:
                        bc.update(t=t)
            for source in self.sources:
                if source.value.temperature_dependent:
                    source.value.update(t=t)

        if self.temperature_time_dependent:
            if isinstance(self.temperature_fenics, fem.Constant):
                self.temperature_fenics.value = self.temperature(t=t)
            elif isinstance(self.temperature_fenics, fem.Function):
                self.temperature_fenics.interpolate(self.temperature_expr)

        for advec_term in self.advection_terms:
            if advec_term.velocity.explicit_time_dependent:
                advec_term.velocity.update(t=t)

    def update_post_processing_solutions(self):
        ns = len(self.species)
        if getattr(self.solver, "device_resident", False):
            # the mixed vector lives on the device: split it when a species field is read
            for idx, spe in enumerate(self.species):
                def fetch(out, idx=idx):
                    out[:] = self.u.x.array.reshape(-1, ns)[:, idx]

                spe.post_processing_solution.x.defer(fetch)
            return
        arr = self.u.x.array.reshape(-1, ns)
        for idx, spe in enumerate(self.species):
            spe.post_processing_solution.x.array[:] = arr[:, idx]

    # ---- exports ---------------------------------------------------------------
    def initialise_exports(self):
        """Reference :427-570 (derived quantities and 1D profiles only)."""
        for export in self.exports:
            times = getattr(export, "times", None)
            if times and not isinstance(export, _exports.DerivedQuantity):
                for time in times:
                    if time not in self.settings.stepsize.milestones:
                        warnings.warn(
                            "To ensure that the exports data at the desired times"
                            "the values in export.times are added to milestones"
                        )
                        self.settings.stepsize.milestones.append(time)
                self.settings.stepsize.milestones.sort()
            if hasattr(export, "field"):
                if isinstance(export.field, list):
                    for idx, field in enumerate(export.field):
                        if isinstance(field, str):
                            export.field[idx] = _species.find_species_from_name(field, self.species)
                elif isinstance(export.field, str):
                    export.field = _species.find_species_from_name(export.field, self.species)
            self._define_field_writer(export)
            if isinstance(export, _exports.Profile1DExport):
                export.data, export.t = [], []
            if isinstance(export, _exports.DerivedQuantity):
                if self.mesh.coordinate_system != CoordinateSystem.CARTESIAN:   # reference :486-493
                    raise NotImplementedError(
                        "Derived quantity exports are not implemented for "
                        f"{self.mesh.coordinate_system!s} meshes")
                export.t, export.data = [], []
        # D_global snapshot: the temperature the DG1 diffusivity was last
        # interpolated with (:632-670, refreshed per :1049-1056)
        self.refresh_D_global()

    def _define_field_writer(self, export):
        """Reference :442-483, :508-510: one writer per field export (``field_io``: VTK XML series
        at the ``.bp`` path, XDMF with raw side-car data)."""
        from festim_b200 import field_io

        def VTUSeriesWriter(path, mesh):
            return field_io.VTUSeriesWriter(path, mesh) if field_io.is_io_rank() else field_io.NullWriter()

        if isinstance(export, _exports.VTXTemperatureExport):
            self._temperature_as_function = self._get_temperature_field_as_function()
            export.writer = VTUSeriesWriter(export.filename, self.mesh.mesh)
        elif isinstance(export, _exports.VTXSpeciesExport):
            if export._checkpoint:
                # reference :464-469 (io4dolfinx.write_mesh): here the restart file is an XDMF
                # document with raw side-car data inside the `.bp` directory, read back by
                # `read_function_from_file`
                if field_io.is_io_rank():
                    export.writer = field_io.XDMFFile(export.filename / "checkpoint.xdmf", "w")
                    export.writer.write_mesh(self.mesh.mesh)
                else:
                    export.writer = field_io.NullWriter()
            else:
                export.writer = VTUSeriesWriter(export.filename, self.mesh.mesh)
        elif isinstance(export, _exports.CustomFieldExport):
            if export.checkpoint:
                raise NotImplementedError("checkpoint=True needs ADIOS2 (not in this image)")
            export.function = fem.Function(fem.FunctionSpace(self.mesh.mesh, "Lagrange", 1), name="f")
            export.set_dolfinx_expression(temperature=self.temperature_fenics, time=self.t)
            export.writer = VTUSeriesWriter(export.filename, self.mesh.mesh)
        elif isinstance(export, _exports.XDMFExport):
            export.define_writer()

    def _get_temperature_field_as_function(self):
        """Reference :572-593."""
        if isinstance(self.temperature_fenics, fem.Function):
            return self.temperature_fenics
        fn = fem.Function(fem.FunctionSpace(self.mesh.mesh, "Lagrange", 1), name="T")
        fn.x.array[:] = float(self.temperature_fenics.value)
        return fn

    def prepare_custom_quantities(self):
        """Reference :545-571: keyword arguments handed to ``CustomQuantity.expr``.  The
        diffusivities ``D_<species>`` are P1 nodal fields of the Arrhenius law of the material
        the quantity's subdomain belongs to, evaluated with the vertex temperatures (the
        restriction of the reference's DG1 ``D_global`` to that subdomain)."""
        self._custom_D_fields = []
        for export in self.exports:
            if not isinstance(export, _exports.CustomQuantity):
                continue
            sub = export.subdomain
            is_surface = not hasattr(sub, "material")
            vol = self.volume_subdomain_of_surface(sub) if is_surface else sub
            kwargs = {spe.name: spe.solution for spe in self.species}
            kwargs["n"] = ufl.FacetNormal(gdim=self.mesh.mesh.gdim)
            kwargs["t"] = self.t
            kwargs["T"] = self.temperature_fenics
            D_kwargs = {}
            for spe in self.species:
                if spe.mobile and not vol.material.D:
                    fn = fem.Function(fem.FunctionSpace(self.mesh.mesh, "Lagrange", 1), name=f"D_{spe.name}")
                    self._custom_D_fields.append((fn, vol.material.get_D_0(spe), vol.material.get_E_D(spe)))
                    D_kwargs[f"D_{spe.name}"] = fn
            kwargs.update(D_kwargs)
            kwargs["D"] = {name[2:]: fn for name, fn in D_kwargs.items()}
            if len(self.species) == 1 and D_kwargs:
                kwargs["D"] = next(iter(D_kwargs.values()))
            kwargs["x"] = ufl.SpatialCoordinate(self.mesh.mesh)
            export.ufl_expr = ufl.as_expr(export.expr(**kwargs))
        self._update_custom_D_fields()

    def _update_custom_D_fields(self):
        T = self.temperature_fenics
        Tn = T.x.array if isinstance(T, fem.Function) else float(T.value)
        for fn, D_0, E_D in getattr(self, "_custom_D_fields", []):
            fn.x.array[:] = float(D_0) * np.exp(-float(E_D) / (k_B * Tn))

    def define_D_global(self, species):
        """Reference :632-670: the DG1 field that holds, in every cell, the Arrhenius diffusivity of
        the cell's material at the temperatures of the cell's vertices - discontinuous across
        material interfaces - stored cell-major (``x.array[c * (tdim + 1) + a]``).  It is the field
        ``SurfaceFlux`` weighs the gradient with; the device evaluates the same values inside
        ``fb_surface_flux`` from ``diffusivity_tables()`` and the temperature snapshot of
        ``refresh_D_global`` instead of storing them.  Returns ``(D, None)`` (no separate
        interpolation expression is needed here)."""
        assert isinstance(species, _species.Species)
        if self.volume_subdomains[0].material.D:
            if len(self.volume_subdomains) > 1:
                raise NotImplementedError(
                    "Giving the diffusion coefficient as a function is currently "
                    "only supported for a single volume subdomain case")
            return self.volume_subdomains[0].material.D, None
        mesh = self.mesh.mesh
        ctag = np.zeros(mesh.num_cells, dtype=np.int64)
        ctag[self.volume_meshtags.indices] = self.volume_meshtags.values
        D_0, E_D = np.zeros(mesh.num_cells), np.zeros(mesh.num_cells)
        for vol in self.volume_subdomains:
            cells = ctag == vol.id
            D_0[cells] = vol.material.get_D_0(species=species)
            E_D[cells] = vol.material.get_E_D(species=species)
        T = self.temperature_fenics
        Tc = T.x.array[mesh.cells] if isinstance(T, fem.Function) else np.full(mesh.cells.shape, float(T.value))
        D = fem.Function(self.V_DG_1, name=f"D_{species.name}")
        D.x.array[:] = (D_0[:, None] * np.exp(-E_D[:, None] / k_B / Tc)).ravel()
        return D, None

    def refresh_D_global(self):
        self._update_custom_D_fields()
        T = self.temperature_fenics
        if isinstance(T, fem.Function):
            self._D_global_T = T.x.array.copy()
        else:
            self._D_global_T = float(T.value)
        if self.solver is not None and hasattr(self.solver, "set_D_global_temperature"):
            self.solver.set_D_global_temperature(self._D_global_T)

    def post_processing(self):
        """Reference :1044-1156."""
        self.update_post_processing_solutions()
        if self.temperature_time_dependent:
            self.refresh_D_global()
        for export in self.exports:
            if hasattr(export, "times"):
                if not is_it_time_to_export(current_time=float(self.t), times=export.times):
                    continue
            if isinstance(export, (_exports.SurfaceQuantity, _exports.VolumeQuantity)):
                if isinstance(export, (_exports.SurfaceFlux, _exports.TotalSurface,
                                       _exports.AverageSurface)) and self.advection_terms:
                    warnings.warn(
                        "Advection terms are not currently accounted for in the "
                        "evaluation of surface flux values"
                    )
                export.compute(self)
                export.t.append(float(self.t))
                if export.filename is not None:
                    export.write(t=float(self.t))
            if isinstance(export, _exports.CustomQuantity):
                export.compute(self)
                export.t.append(float(self.t))
                if export.filename is not None:
                    export.write(t=float(self.t))
                continue
            if isinstance(export, _exports.VTXSpeciesExport):
                export.write(float(self.t))
            elif isinstance(export, _exports.VTXTemperatureExport):
                if self.temperature_time_dependent:   # reference :1079-1086: a steady T is not rewritten
                    self._temperature_as_function = self._get_temperature_field_as_function()
                    export.writer.write(float(self.t), {"T": self._temperature_as_function.x.array})
            elif isinstance(export, _exports.CustomFieldExport):
                export.update(self.species)
                export.write(float(self.t))
            if isinstance(export, _exports.XDMFExport):
                export.write(float(self.t))
            if isinstance(export, _exports.Profile1DExport):
                if export.x is None:
                    coords = self.mesh.mesh.coords[:, 0]
                    export._sort_coords = np.argsort(coords)
                    export.x = coords[export._sort_coords]
                c = export.field.post_processing_solution.x.array[export._sort_coords]
                export.data.append(c.copy())
                export.t.append(float(self.t))

    def diffusivity_tables(self):
        """(D_0, E_D) per (volume subdomain, species), 0 for immobile species:
        the piecewise-constant data behind ``define_D_global`` (:656-663)."""
        nv, ns = len(self.volume_subdomains), len(self.species)
        D0 = np.zeros((nv, ns))
        ED = np.zeros((nv, ns))
        for iv, vol in enumerate(self.volume_subdomains):
            for s, spe in enumerate(self.species):
                if spe.mobile and not vol.material.D:
                    D0[iv, s] = vol.material.get_D_0(spe)
                    ED[iv, s] = vol.material.get_E_D(spe)
        return D0, ED


__all__ = ["HydrogenTransportProblem", "k_B"]
