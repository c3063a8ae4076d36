"""Node elimination for simplex quadrature (the Xiao-Gimbutas strategy, without symmetry).

Start from the collapsed Gauss-Jacobi rule of festim_b200/quadrature.py (exact, positive,
interior, but (m+1)^d points), repeatedly drop the least significant point and restore
exactness with damped Gauss-Newton steps (minimum-norm least squares on points AND
weights) in an orthonormal polynomial basis.  Stops when no point can be removed with all
weights positive and all points inside.  Output: a table entry for _quadrature_tables.py.

usage: python tools/reduce_quadrature.py tet 9 [target_points [tries [previous_output]]]
"""
import itertools, os, sys
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from festim_b200 import quadrature  # noqa: E402


def exponents(d, deg):
    return np.array([e for e in itertools.product(range(deg + 1), repeat=d) if sum(e) <= deg])


def jacobi(n, alpha, x):
    """Jacobi polynomials P_0..P_n^(alpha,0)(x) by recurrence (x may be complex)."""
    out = [np.ones_like(x)]
    if n >= 1:
        out.append(0.5 * (alpha + (alpha + 2.0) * x))
    for m in range(1, n):
        ab = alpha  # beta = 0
        c1 = 2.0 * (m + 1) * (m + ab + 1) * (2 * m + ab)
        c2 = (2 * m + ab + 1) * (alpha * alpha)
        c3 = (2 * m + ab) * (2 * m + ab + 1) * (2 * m + ab + 2)
        c4 = 2.0 * (m + alpha) * m * (2 * m + ab + 2)
        out.append(((c2 + c3 * x) * out[m] - c4 * out[m - 1]) / c1)
    return out


class Basis:
    """Orthonormal polynomials on the reference simplex (weights sum to 1): the
    Koornwinder-Dubiner basis (orthogonal by construction, so the final Gram/Cholesky
    normalisation is perfectly conditioned); gradients by complex step (exact for
    polynomials)."""

    def __init__(self, d, deg):
        self.d, self.deg = d, deg
        self.E = exponents(d, deg)
        P, W = quadrature.rule(d, 2 * deg + 2)
        W = W / W.sum()
        self.Linv = np.eye(len(self.E))
        V = self.eval(P[:, 1:])
        G = V.T @ (V * W[:, None])
        self.Linv = np.linalg.inv(np.linalg.cholesky(G))
        self.target = self.Linv @ (V * W[:, None]).sum(0)

    def raw(self, X):
        deg = self.deg
        cols = []
        if self.d == 2:
            x, y = X[:, 0], X[:, 1]
            a = 2.0 * x / (1.0 - y) - 1.0
            b = 2.0 * y - 1.0
            Pa = jacobi(deg, 0.0, a)
            for i, j in self.E:
                cols.append(Pa[i] * jacobi(j, 2.0 * i + 1.0, b)[j] * (1.0 - y) ** i)
        else:
            x, y, z = X[:, 0], X[:, 1], X[:, 2]
            a = 2.0 * x / (1.0 - y - z) - 1.0
            b = 2.0 * y / (1.0 - z) - 1.0
            c = 2.0 * z - 1.0
            Pa = jacobi(deg, 0.0, a)
            Pb = {i: jacobi(deg - i, 2.0 * i + 1.0, b) for i in range(deg + 1)}
            Pc = {m: jacobi(deg - m, 2.0 * m + 2.0, c) for m in range(deg + 1)}
            for i, j, k in self.E:
                cols.append(Pa[i] * Pb[i][j] * ((1.0 - y - z)) ** i * (1.0 - z) ** j * Pc[i + j][k])
        return np.stack(cols, axis=1)

    def eval(self, X):
        return self.raw(X) @ self.Linv.T

    def grad(self, X):
        h = 1e-30
        out = []
        for k in range(self.d):
            Xc = X.astype(complex)
            Xc[:, k] += 1j * h
            out.append((self.raw(Xc).imag / h) @ self.Linv.T)
        return out


def solve(B, X, w, tol=3e-16, iters=60):
    d = B.d
    n = len(w)
    best = None
    for it in range(iters):
        Phi = B.eval(X)
        r = Phi.T @ w - B.target
        err = np.abs(r).max()
        if err < tol * 50:
            best = (X.copy(), w.copy(), err)
            if err < tol * 4:
                break
        G = B.grad(X)
        J = np.concatenate([Phi.T] + [g.T * w[None, :] for g in G], axis=1)   # (nb, n(1+d))
        step, *_ = np.linalg.lstsq(J, -r, rcond=1e-13)
        dw, dX = step[:n], step[n:].reshape(d, n).T
        # damping: keep points inside and weights positive
        t = 1.0
        for _ in range(30):
            Xn, wn = X + t * dX, w + t * dw
            lam0 = 1.0 - Xn.sum(1)
            if (wn > 0).all() and (Xn > 1e-4).all() and (lam0 > 1e-4).all():
                break
            t *= 0.5
        else:
            return None
        X, w = Xn, wn
    return best


def reduce(d, deg, target=0, tries=12, seed=0, start=None):
    B = Basis(d, deg)
    if start is not None:
        P, W = start
    else:
        P, W = quadrature.rule(d, deg)
    X, w = P[:, 1:].copy(), W / W.sum()
    rng = np.random.default_rng(seed)
    while len(w) > target:
        Phi = B.eval(X)
        sig = w * (Phi ** 2).sum(1)
        order = np.argsort(sig)
        done = False
        for c in order[:tries]:
            keep = np.ones(len(w), bool)
            keep[c] = False
            res = solve(B, X[keep], w[keep] / w[keep].sum())
            if res is not None:
                X, w, err = res
                print(f"# {len(w)} points, residual {err:.1e}", file=sys.stderr, flush=True)
                done = True
                break
        if not done:
            break
    return X, w


if __name__ == "__main__":
    name, deg = sys.argv[1], int(sys.argv[2])
    target = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    tries = int(sys.argv[4]) if len(sys.argv) > 4 else 12
    d = 2 if name == "tri" else 3
    start = None
    if len(sys.argv) > 5:  # continue from a rule printed by an earlier run
        txt = open(sys.argv[5]).read()
        body = txt[txt.index("): (") + 3:].rstrip().rstrip(",")
        pts, wts = eval(body)
        start = (np.array(pts), np.array(wts))
    X, w = reduce(d, deg, target, tries=tries, start=start)
    P = np.concatenate([1.0 - X.sum(1, keepdims=True), X], axis=1)
    # verify with raw monomials against the big rule
    Pr, Wr = quadrature.rule(d, deg + 2)
    Wr = Wr / Wr.sum()
    err = 0.0
    for e in itertools.product(range(deg + 1), repeat=d + 1):
        if sum(e) <= deg:
            err = max(err, abs((w * np.prod(P ** np.array(e), axis=1)).sum() - (Wr * np.prod(Pr ** np.array(e), axis=1)).sum()))
    print(f"# {name} degree {deg}: {len(w)} points, node elimination (tools/reduce_quadrature.py), "
          f"min weight {w.min():.2e}, min barycentric {P.min():.2e}, max monomial error {err:.1e}")
    print(f"({name!r}, {deg}): (")
    print("    " + repr(P.tolist()) + ",")
    print("    " + repr(w.tolist()) + "),")
