"""Host rule that picks the number of operator applications per Krylov iteration of the Chebyshev-preconditioned
GMRES from the spectrum interval of D^-1 J and the reduction a solve still needs (csrc/linalg.cu:
cheb_choose_steps, through the test hook fb_chebyshev_steps)."""
import ctypes as C
import math

import pytest

from festim_b200.backend import load_library


def _steps(lmin, lmax, reduction):
    lib = load_library()
    m = C.c_int(0)
    k = lib.fb_chebyshev_steps(lmin, lmax, reduction, C.byref(m))
    return k, m.value


def _bound(lmin, lmax, k, m):
    sigma = (lmax + lmin) / (lmax - lmin)
    return math.cosh(k * math.acosh(sigma)) ** (-m)


def test_the_c2_steady_state_choice():
    # interval and reduction measured on BASELINE configs[2] at full size
    assert _steps(0.4599, 2.7484, 6.0e-8) == (7, 3)


@pytest.mark.parametrize("interval", [(0.46, 2.75), (0.9, 1.2), (0.05, 3.0), (0.3, 12.0)])
@pytest.mark.parametrize("reduction", [0.3, 1e-3, 1e-6, 6e-8, 1e-10, 1e-12, 1e-15])
def test_choice_reaches_the_reduction_within_the_modelled_miss(interval, reduction):
    lmin, lmax = interval
    k, m = _steps(lmin, lmax, reduction)
    assert 2 <= k <= 12 and 3 <= m <= 8
    if (k, m) != (12, 8):   # (12, 8) is the cap: the interval is too wide for the target
        assert _bound(lmin, lmax, k, m) <= 2.5 * reduction
    # never more than the work of the most cautious admissible choice
    cautious = min((mm * (kk + 0.7) for mm in range(3, 9) for kk in range(2, 13)
                    if _bound(lmin, lmax, kk, mm) <= reduction), default=None)
    if cautious is not None:
        assert m * (k + 0.7) <= cautious + 1e-9


def test_tighter_targets_never_cost_less():
    costs = []
    for red in [1e-2, 1e-4, 1e-6, 1e-8, 1e-10, 1e-12]:
        k, m = _steps(0.46, 2.75, red)
        costs.append(m * (k + 0.7))
    assert all(b >= a - 1.3 * 12.7 * 0.5 for a, b in zip(costs, costs[1:]))   # up to the modelled half miss
    assert costs[-1] > costs[0]


def test_bad_input_is_refused():
    lib = load_library()
    m = C.c_int(0)
    assert lib.fb_chebyshev_steps(0.0, 1.0, 1e-3, C.byref(m)) < 0
    assert lib.fb_chebyshev_steps(2.0, 1.0, 1e-3, C.byref(m)) < 0
