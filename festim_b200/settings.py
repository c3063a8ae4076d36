"""Simulation settings (reference ``src/festim/settings.py:35-68``)."""

from festim_b200.stepsize import Stepsize


class Settings:
    def __init__(self, atol, rtol, max_iterations=30, transient=True, final_time=None,
                 element_degree=1, stepsize=None, convergence_criterion="residual"):
        self.atol = atol
        self.rtol = rtol
        self.max_iterations = max_iterations
        self.transient = transient
        self.final_time = final_time
        if element_degree != 1:
            raise NotImplementedError("the B200 backend implements P1 elements only")
        self.element_degree = element_degree
        self.stepsize = stepsize
        self.convergence_criterion = convergence_criterion

    @property
    def stepsize(self):
        return self._stepsize

    @stepsize.setter
    def stepsize(self, value):
        if value is None:
            self._stepsize = None
        elif isinstance(value, (float, int)):
            self._stepsize = Stepsize(initial_value=value)
        elif isinstance(value, Stepsize):
            self._stepsize = value
        else:
            raise TypeError("stepsize must be an of type int, float or festim.Stepsize")
