"""Step-size control: the behaviours pinned by reference test/test_stepsize.py:7-227,
restated against festim_b200.Stepsize (same constructor, setters and error messages)."""
import numpy as np
import pytest

import festim_b200 as F


@pytest.mark.parametrize("factor, target, direction", [
    (10, 5, "grow"), (1.2, 2, "grow"), (1, 1, "grow"),
    (0.8, 5, "shrink"), (0.5, 2, "shrink"), (1, 1, "shrink")])
def test_adaptive_stepsize_grows_and_shrinks(factor, target, direction):
    st = F.Stepsize(initial_value=2)
    st.target_nb_iterations = target
    if direction == "grow":
        st.growth_factor = factor
        new = st.modify_value(value=2, nb_iterations=target - 1)
    else:
        st.cutback_factor = factor
        new = st.modify_value(value=2, nb_iterations=target + 1)
    assert np.isclose(new, 2 * factor)


@pytest.mark.parametrize("nb_its, target", [(1, 4), (5, 4), (4, 4)])
def test_max_stepsize_caps(nb_its, target):
    st = F.Stepsize(initial_value=1)
    st.max_stepsize = 4
    st.growth_factor, st.cutback_factor, st.target_nb_iterations = 1.1, 0.9, target
    assert st.modify_value(value=10, nb_iterations=nb_its) == 4


def test_unchanged_at_target_and_custom_not_adaptive():
    st = F.Stepsize(initial_value=2)
    st.target_nb_iterations = 5
    assert np.isclose(st.modify_value(value=2, nb_iterations=5), 2)

    class Custom(F.Stepsize):
        def is_adapt(self, t):
            return False

    assert np.isclose(Custom(initial_value=2).modify_value(value=2, nb_iterations=5), 2)


def test_setters_validate():
    st = F.Stepsize(1)
    with pytest.raises(ValueError, match="growth factor should be greater than one"):
        st.growth_factor = 0.5
    st.growth_factor = None
    assert st.growth_factor is None
    with pytest.raises(ValueError, match="cutback factor should be smaller than one"):
        st.cutback_factor = 1.5
    st.cutback_factor = None
    assert st.cutback_factor is None
    with pytest.raises(ValueError, match="maximum stepsize cannot be less than initial stepsize"):
        st.max_stepsize = 0.5
    st.max_stepsize = None
    assert st.max_stepsize is None
    with pytest.raises(ValueError, match="milestone tolerance should be greater than zero"):
        st.milestone_tolerance = 0.0


@pytest.mark.parametrize("milestones, current_time, expected", [
    ([1.0, 25.4], 20.1, 25.4), ([9.8], 10.0, None), ([2.0, 0.5, 20.0], 0.0, 0.5),
    ([3.4, 9.5, 4.4], 4.4, 9.5), ([15.3, 1.2, 0.7, 1.4], 15.3, None)])
def test_next_milestone(milestones, current_time, expected):
    st = F.Stepsize(initial_value=0.5)
    st.milestones = milestones
    assert st.next_milestone(current_time=current_time) == expected


def test_overshoot_milestone():
    st = F.Stepsize(initial_value=0.1)
    st.growth_factor, st.target_nb_iterations, st.milestones = 1, 4, [1.3]
    assert st.modify_value(value=100000, nb_iterations=1, t=1) == 1.3 - 1


@pytest.mark.parametrize("t, expected", [(0, 10), (10, 10), (100, 10), (1000, None), (1001, None)])
def test_get_max_stepsize_callable(t, expected):
    st = F.Stepsize(initial_value=2)
    st.max_stepsize = lambda t: 10 if t < 1000 else None
    assert st.get_max_stepsize(t) == expected


def test_milestones_need_adaptivity():
    with pytest.raises(ValueError, match=r"Milestones are only relevant if the stepsize is adaptive."):
        F.Stepsize(initial_value=2, milestones=[1, 2, 3])


def test_milestone_tolerance():
    """reference issue #933: with a tight tolerance the step lands on the milestone."""
    st = F.Stepsize(initial_value=1, growth_factor=1.2, cutback_factor=0.9, target_nb_iterations=4,
                    milestones=[0.05, 0.1, 0.2, 0.5, 1], milestone_tolerance=1e-10)
    new_dt = st.modify_value(value=1, nb_iterations=4, t=0.499999)
    assert np.isclose(0.499999 + new_dt, 0.5)


def test_settings_stepsize_types():
    """reference src/festim/settings.py:55-68."""
    s = F.Settings(atol=1, rtol=1, stepsize=0.5)
    assert isinstance(s.stepsize, F.Stepsize) and s.stepsize.initial_value == 0.5
    with pytest.raises(TypeError):
        F.Settings(atol=1, rtol=1, stepsize="fast")


def test_is_it_time_to_export_consumes_times():
    """reference src/festim/helpers.py:374-401."""
    times = [1.0, 2.0]
    assert F.is_it_time_to_export(times=None, current_time=0.3)
    assert not F.is_it_time_to_export(times=times, current_time=1.5)
    assert F.is_it_time_to_export(times=times, current_time=1.0 + 1e-7)
    assert times == [2.0]
