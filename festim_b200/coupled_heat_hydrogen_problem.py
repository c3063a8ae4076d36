"""Staggered heat + hydrogen coupling (reference
``src/festim/coupled_heat_hydrogen_problem.py:100-163``): heat step, hydrogen
step with T aliasing the heat solution, both take min(dt_H, dt_T)."""

from festim_b200.heat_transfer_problem import HeatTransferProblem
from festim_b200.hydrogen_transport_problem import HydrogenTransportProblem


class CoupledTransientHeatTransferHydrogenTransport:
    def __init__(self, heat_problem, hydrogen_problem):
        self.heat_problem = heat_problem
        self.hydrogen_problem = hydrogen_problem
        if not self.heat_problem.settings.transient or not self.hydrogen_problem.settings.transient:
            raise TypeError("Both the heat and hydrogen problems must be set to transient")

    @property
    def heat_problem(self):
        return self._heat_problem

    @heat_problem.setter
    def heat_problem(self, value):
        if not isinstance(value, HeatTransferProblem):
            raise TypeError("heat_problem must be a festim.HeatTransferProblem object")
        value.show_progress_bar = False
        self._heat_problem = value

    @property
    def hydrogen_problem(self):
        return self._hydrogen_problem

    @hydrogen_problem.setter
    def hydrogen_problem(self, value):
        if not isinstance(value, HydrogenTransportProblem):
            raise TypeError("hydrogen_problem must be a festim.HydrogenTransportProblem object")
        self._hydrogen_problem = value

    @property
    def non_matching_meshes(self):
        return self.heat_problem.mesh.mesh is not self.hydrogen_problem.mesh.mesh

    def initialise(self):
        heat, hyd = self.heat_problem, self.hydrogen_problem
        dt0 = min(heat.settings.stepsize.initial_value, hyd.settings.stepsize.initial_value)
        heat.settings.stepsize.initial_value = dt0
        hyd.settings.stepsize.initial_value = dt0
        if heat.settings.final_time != hyd.settings.final_time:
            raise ValueError(
                "Final time values in the heat transfer and hydrogen transport "
                "model must be the same"
            )
        heat.initialise()
        if self.non_matching_meshes:
            # reference :124-129: T on the hydrogen mesh is a separate field, refreshed from the
            # heat solution by point location + P1 interpolation before every hydrogen step
            from festim_b200 import fem

            T_func = fem.Function(fem.functionspace(hyd.mesh.mesh, ("P", 1)))
            self._locator = fem.interpolate_nonmatching(T_func, heat.u)
            hyd.temperature = T_func
        else:
            hyd.temperature = heat.u  # aliased: the hydrogen kernels see every new T
        hyd.initialise()

    def iterate(self):
        heat, hyd = self.heat_problem, self.hydrogen_problem
        heat.iterate()
        if self.non_matching_meshes:   # reference :136-139
            from festim_b200 import fem

            fem.interpolate_nonmatching(hyd.temperature_fenics, heat.u, locator=self._locator)
        hyd.iterate()
        next_dt = min(float(hyd.dt), float(heat.dt))
        heat.dt.value = next_dt
        hyd.dt.value = next_dt

    def run(self):
        hyd = self.hydrogen_problem
        while hyd.t.value < hyd.settings.final_time:
            self.iterate()
