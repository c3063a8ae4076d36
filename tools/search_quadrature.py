"""Search compact fully-symmetric simplex quadrature rules (degrees 7-10).

A fully symmetric rule integrates a polynomial exactly iff it integrates its
symmetrisation, so only the invariants of the simplex symmetry group have to be
matched: products of the barycentric power sums p2, p3 (triangle) and p2, p3, p4
(tetrahedron).  That turns the moment system into ~10-20 equations in as many
unknowns, which a multi-start Levenberg-Marquardt solves in milliseconds.  The
exact moments come from the (much larger) collapsed Gauss-Jacobi rule of
festim_b200/quadrature.py.  Output: table entries for _quadrature_tables.py.

usage: python tools/search_quadrature.py tet 9 S4,S31,S31,S31,S211,S211,S211,S211 [tries] [seed]
"""
import itertools, os, sys
import numpy as np
from scipy.optimize import least_squares

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

SIZE = {"S3": 1, "S21": 3, "S111": 6, "S4": 1, "S31": 4, "S22": 6, "S211": 12, "S1111": 24}
NPAR = {"S3": 0, "S21": 1, "S111": 2, "S4": 0, "S31": 1, "S22": 1, "S211": 2, "S1111": 3}


def rep(kind, p):
    if kind == "S3": return np.array([1 / 3] * 3)
    if kind == "S21": return np.array([p[0], p[0], 1 - 2 * p[0]])
    if kind == "S111": return np.array([p[0], p[1], 1 - p[0] - p[1]])
    if kind == "S4": return np.array([0.25] * 4)
    if kind == "S31": return np.array([p[0]] * 3 + [1 - 3 * p[0]])
    if kind == "S22": return np.array([p[0], p[0], 0.5 - p[0], 0.5 - p[0]])
    if kind == "S211": return np.array([p[0], p[0], p[1], 1 - 2 * p[0] - p[1]])
    if kind == "S1111": return np.array([p[0], p[1], p[2], 1 - p[0] - p[1] - p[2]])
    raise ValueError(kind)


def invariants(d, deg):
    gens = (2, 3) if d == 2 else (2, 3, 4)
    out = []
    for e in itertools.product(*[range(deg // g + 1) for g in gens]):
        if sum(a * g for a, g in zip(e, gens)) <= deg:
            out.append(e)
    return gens, out


def inv_values(lam, gens, exps):
    ps = [np.sum(lam ** g, axis=-1) for g in gens]
    return np.stack([np.prod([p ** a for p, a in zip(ps, e)], axis=0) for e in exps], -1)


def exact_moments(d, deg):
    from festim_b200 import quadrature
    P, W = quadrature.rule(d, deg + 2)
    W = W / W.sum()
    gens, exps = invariants(d, deg)
    return gens, exps, (inv_values(P, gens, exps) * W[:, None]).sum(0)


def unpack(structure, x):
    k = 0
    for kind in structure:
        n = NPAR[kind]
        yield kind, x[k:k + n], x[k + n]
        k += n + 1


def residual(x, structure, gens, exps, target):
    acc = np.zeros(len(exps))
    for kind, p, w in unpack(structure, x):
        acc += SIZE[kind] * w * inv_values(rep(kind, p), gens, exps)
    r = acc - target
    return np.concatenate([r, np.zeros(max(0, len(x) - len(r)))])  # 'lm' wants m >= n


def expand(structure, x):
    pts, wts = [], []
    for kind, p, w in unpack(structure, x):
        orbit = sorted(set(itertools.permutations(np.round(rep(kind, p), 17))))
        assert len(orbit) == SIZE[kind], (kind, p, len(orbit))
        pts += orbit
        wts += [w] * len(orbit)
    return np.array(pts), np.array(wts)


def search(d, deg, structure, tries=20000, seed=0):
    gens, exps, target = exact_moments(d, deg)
    rng = np.random.default_rng(seed)
    nx = sum(NPAR[k] + 1 for k in structure)
    npts = sum(SIZE[k] for k in structure)
    for t in range(tries):
        x = np.empty(nx)
        k = 0
        for kind in structure:
            n = NPAR[kind]
            x[k:k + n] = rng.uniform(0.01, 0.32 if kind != "S22" else 0.24, n)
            x[k + n] = rng.uniform(0.2, 1.8) / npts
            k += n + 1
        sol = least_squares(residual, x, args=(structure, gens, exps, target), method="lm",
                            xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=400)
        if np.abs(sol.fun).max() > 1e-13:
            continue
        ok = True
        for kind, p, w in unpack(structure, sol.x):
            lam = rep(kind, p)
            if w <= 0 or (lam <= 1e-4).any() or len(set(np.round(lam, 9))) != len(set(np.round(rep(kind, np.array([0.11, 0.17, 0.23])[:NPAR[kind]]), 9))):
                ok = False
        if not ok:
            continue
        sol = least_squares(residual, sol.x, args=(structure, gens, exps, target), method="lm",
                            xtol=3e-16, ftol=3e-16, gtol=3e-16, max_nfev=100)
        return sol.x, t
    return None, tries


if __name__ == "__main__":
    name, deg, structure = sys.argv[1], int(sys.argv[2]), sys.argv[3].split(",")
    tries = int(sys.argv[4]) if len(sys.argv) > 4 else 20000
    seed = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    d = 2 if name == "tri" else 3
    x, t = search(d, deg, structure, tries, seed)
    if x is None:
        print(f"# {name} degree {deg} {structure}: no rule in {tries} starts")
        sys.exit(1)
    P, W = expand(structure, x)
    # verify against ALL monomials of the degree with the reference rule
    from festim_b200 import quadrature
    Pr, Wr = quadrature.rule(d, deg + 2)
    Wr = Wr / Wr.sum()
    err = 0.0
    for e in itertools.product(range(deg + 1), repeat=d + 1):
        if sum(e) <= deg:
            err = max(err, abs((W * np.prod(P ** np.array(e), axis=1)).sum() - (Wr * np.prod(Pr ** np.array(e), axis=1)).sum()))
    print(f"# {name} degree {deg}: {len(W)} points, orbits {structure}, start {t}, max monomial error {err:.1e}")
    print(f"({name!r}, {deg}): (")
    print("    " + repr(P.tolist()) + ",")
    print("    " + repr(W.tolist()) + "),")
