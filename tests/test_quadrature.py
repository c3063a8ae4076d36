"""Every quadrature rule integrates all monomials of its nominal degree exactly
(the property the reference relies on from basix)."""
import itertools
import math

import numpy as np
import pytest

from festim_b200 import quadrature


def exact(p):
    d = len(p)
    return math.prod(math.factorial(a) for a in p) * math.factorial(d) / math.factorial(sum(p) + d)


@pytest.mark.parametrize("tdim", [1, 2, 3])
@pytest.mark.parametrize("degree", list(range(0, 13)))
def test_rule_exactness(tdim, degree):
    P, W = quadrature.rule(tdim, degree)
    assert P.shape == (len(W), tdim + 1)
    assert np.all(W > 0) and abs(W.sum() - 1) < 1e-14
    assert np.allclose(P.sum(axis=1), 1.0, atol=1e-14)
    for p in itertools.product(range(degree + 1), repeat=tdim):
        if sum(p) <= degree:
            val = np.sum(W * np.prod(P[:, 1:] ** np.array(p), axis=1))
            assert abs(val - exact(p)) < 2e-15 + 1e-13 * exact(p), (p, val)


def test_interval_point_count_matches_basix():
    # basix Gauss-Jacobi: m = (degree + 2) // 2 points
    for degree in range(0, 12):
        assert len(quadrature.rule(1, degree)[1]) == (degree + 2) // 2
