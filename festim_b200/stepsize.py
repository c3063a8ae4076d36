"""Adaptive time-step control (reference ``src/festim/stepsize.py``).

Semantics restated from ``stepsize.py:131-195``: grow
when Newton needed fewer iterations than the target, cut back when it needed
more, cap at ``max_stepsize(t)``, never step over the next milestone.  The
same rule runs inside the library for device-resident stepping
(``fb_stepsize_modify`` in ``include/festim_b200.h``).
"""

import numpy as np


class Stepsize:
    def __init__(self, initial_value, growth_factor=None, cutback_factor=None,
                 target_nb_iterations=None, max_stepsize=None, milestones=None,
                 milestone_tolerance=1e-5):
        self.initial_value = initial_value
        self.growth_factor = growth_factor
        self.cutback_factor = cutback_factor
        self.target_nb_iterations = target_nb_iterations
        self.max_stepsize = max_stepsize
        self.milestones = milestones or []
        self.milestone_tolerance = milestone_tolerance
        if milestones and (growth_factor is None or cutback_factor is None):
            raise ValueError(
                "Milestones are only relevant if the stepsize is adaptive. "
                "Please provide growth and cutback factors."
            )

    @property
    def milestones(self):
        return self._milestones

    @milestones.setter
    def milestones(self, value):
        self._milestones = sorted(value) if value else value

    @property
    def milestone_tolerance(self):
        return self._milestone_tolerance

    @milestone_tolerance.setter
    def milestone_tolerance(self, value):
        if value <= 0:
            raise ValueError("milestone tolerance should be greater than zero")
        self._milestone_tolerance = value

    @property
    def adaptive(self):
        return self.growth_factor or self.cutback_factor or self.target_nb_iterations

    @property
    def growth_factor(self):
        return self._growth_factor

    @growth_factor.setter
    def growth_factor(self, value):
        if value is not None and value < 1:
            raise ValueError("growth factor should be greater than one")
        self._growth_factor = value

    @property
    def cutback_factor(self):
        return self._cutback_factor

    @cutback_factor.setter
    def cutback_factor(self, value):
        if value is not None and value > 1:
            raise ValueError("cutback factor should be smaller than one")
        self._cutback_factor = value

    @property
    def max_stepsize(self):
        return self._max_stepsize

    @max_stepsize.setter
    def max_stepsize(self, value):
        if isinstance(value, float) and value < self.initial_value:
            raise ValueError("maximum stepsize cannot be less than initial stepsize")
        self._max_stepsize = value

    def get_max_stepsize(self, t):
        ms = self._max_stepsize
        return ms(t) if callable(ms) else ms

    def is_adapt(self, t):
        return True

    def next_milestone(self, current_time):
        if self.milestones is None:
            return None
        later = [m for m in self.milestones if current_time < m]
        return later[0] if later else None

    def modify_value(self, value, nb_iterations, t=None):
        if not self.is_adapt(t):
            return value
        new = value
        if nb_iterations < self.target_nb_iterations:
            new = value * self.growth_factor
        elif nb_iterations > self.target_nb_iterations:
            new = value * self.cutback_factor
        cap = self.get_max_stepsize(t)
        if cap and new > cap:
            new = cap
        milestone = self.next_milestone(t)
        if milestone is not None:
            remaining = milestone - t
            at_milestone = np.isclose(t, milestone, atol=0, rtol=self.milestone_tolerance)
            if new > remaining and not at_milestone:
                new = remaining
        return new
