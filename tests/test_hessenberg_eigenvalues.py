"""The Hessenberg QR iteration that turns the Arnoldi matrix of a GMRES cycle into the interval of
the Chebyshev preconditioner (csrc/linalg.cu: hessenberg_eigenvalues) against numpy."""
import ctypes as C

import numpy as np
import pytest

from festim_b200.backend import load_library


@pytest.mark.parametrize("n", [1, 2, 3, 7, 12, 30])
@pytest.mark.parametrize("kind", ["random", "spd_like", "rotation"])
def test_hessenberg_eigenvalues_match_numpy(n, kind):
    lib = load_library()
    rng = np.random.default_rng(n * 7 + len(kind))
    if kind == "random":
        A = rng.standard_normal((n, n))
    elif kind == "spd_like":   # what a block-Jacobi preconditioned mass-dominated Jacobian gives
        Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
        A = Q @ np.diag(rng.uniform(0.5, 2.5, n)) @ Q.T + 1e-3 * rng.standard_normal((n, n))
    else:                      # complex pairs
        A = rng.standard_normal((n, n)) - rng.standard_normal((n, n)).T + 0.1 * np.eye(n)
    H = np.triu(A, -1).copy()
    wr, wi = np.zeros(n), np.zeros(n)
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    assert lib.fb_hessenberg_eigenvalues(n, p(np.ascontiguousarray(H)), p(wr), p(wi)) == 0
    got = np.sort_complex(wr + 1j * wi)
    ref = np.sort_complex(np.linalg.eigvals(H))
    assert np.abs(got - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max())
