"""Quadrature rules on the reference interval / triangle / tetrahedron.

The reference gets its rules from basix through FFCx: Gauss-Jacobi on
intervals (``(degree + 2) // 2`` points) and Xiao-Gimbutas tables on
triangles/tetrahedra up to degree 30.  The XG tables
are not available offline, so simplices use: the centroid rule (degree <= 1),
the classical symmetric rules of ``_quadrature_tables`` (degree 2..6, polished
to machine precision by ``tools/gen_quadrature.py``) and the collapsed
Gauss-Jacobi (Stroud conical product) rule above that.  Every rule integrates
all polynomials of its nominal degree exactly (``tests/test_quadrature.py``),
so polynomial integrands agree with the reference to round-off; for
non-polynomial integrands (Arrhenius factor of a P1 temperature) the
difference is a quadrature error of the same order as the reference's own.

Points are returned in barycentric coordinates ``(nq, d + 1)`` with
``lambda_0 = 1 - sum(xi)`` first, weights sum to one.
"""

from __future__ import annotations

import functools

import numpy as np
from scipy.special import roots_jacobi

from festim_b200._quadrature_tables import TABLES

_CELL = {1: "interval", 2: "tri", 3: "tet"}


def _gauss_jacobi_01(m, alpha):
    """m-point Gauss-Jacobi rule on [0, 1] for the weight (1 - x)^alpha."""
    x, w = roots_jacobi(m, alpha, 0.0)
    return 0.5 * (x + 1.0), w * 0.5 ** (alpha + 1)


@functools.lru_cache(maxsize=None)
def rule(tdim, degree):
    """(barycentric points (nq, tdim+1), weights (nq,)) exact to ``degree``."""
    degree = max(int(degree), 0)
    if tdim == 0:
        return np.ones((1, 1)), np.ones(1)
    if tdim == 1:
        m = (degree + 2) // 2
        x, w = _gauss_jacobi_01(m, 0.0)
        return np.stack([1.0 - x, x], axis=1), w / w.sum()
    name = _CELL[tdim]
    if degree <= 1:
        return np.full((1, tdim + 1), 1.0 / (tdim + 1)), np.ones(1)
    if degree == 2:
        if tdim == 2:
            a, b = 1.0 / 6.0, 2.0 / 3.0
            pts = np.array([[b, a, a], [a, b, a], [a, a, b]])
        else:
            a = (5.0 - np.sqrt(5.0)) / 20.0
            b = 1.0 - 3.0 * a
            pts = np.full((4, 4), a)
            np.fill_diagonal(pts, b)
        return pts, np.full(tdim + 1, 1.0 / (tdim + 1))
    # degree 3 shares the degree-4 triangle rule, degree 4 the degree-5 tet rule
    key = (name, degree)
    if tdim == 2 and degree == 3:
        key = (name, 4)
    if tdim == 3 and degree == 4:
        key = (name, 5)
    if tdim == 3 and degree == 8 and (name, 8) not in TABLES:
        key = (name, 9)  # fewer points than the collapsed degree-8 rule
    if tdim == 2 and degree == 7:
        key = (name, 8)  # 16 points: smaller than any degree-7 rule found
    if key in TABLES:
        pts, wts = TABLES[key]
        return np.array(pts, dtype=np.float64), np.array(wts, dtype=np.float64)
    # collapsed Gauss-Jacobi
    m = (degree + 2) // 2
    if tdim == 2:
        x0, w0 = _gauss_jacobi_01(m, 1.0)
        x1, w1 = _gauss_jacobi_01(m, 0.0)
        X0, X1 = np.meshgrid(x0, x1, indexing="ij")
        W = np.outer(w0, w1).ravel()
        xi = X0.ravel()
        eta = (X1 * (1.0 - X0)).ravel()
        pts = np.stack([1.0 - xi - eta, xi, eta], axis=1)
    else:
        x0, w0 = _gauss_jacobi_01(m, 2.0)
        x1, w1 = _gauss_jacobi_01(m, 1.0)
        x2, w2 = _gauss_jacobi_01(m, 0.0)
        X0, X1, X2 = np.meshgrid(x0, x1, x2, indexing="ij")
        W = (w0[:, None, None] * w1[None, :, None] * w2[None, None, :]).ravel()
        xi = X0.ravel()
        eta = (X1 * (1.0 - X0)).ravel()
        zeta = (X2 * (1.0 - X0) * (1.0 - X1)).ravel()
        pts = np.stack([1.0 - xi - eta - zeta, xi, eta, zeta], axis=1)
    return pts, W / W.sum()
