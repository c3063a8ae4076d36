"""Derive compact fully-symmetric simplex quadrature rules numerically.

basix (the reference's quadrature provider) ships Xiao-Gimbutas tables that are
not available offline (SURVEY "Hard parts" 2).  This script polishes classical
symmetric rules (Strang-Fix / Dunavant / Keast orbit structures) to machine
precision by solving the monomial moment equations, and prints the barycentric
tables pasted into festim_b200/quadrature.py.
"""
import itertools, math, sys
import numpy as np
from scipy.optimize import least_squares

def monomials(d, deg):
    return [p for p in itertools.product(range(deg + 1), repeat=d) if sum(p) <= deg]

def exact(p):
    d = len(p)
    return math.prod(math.factorial(a) for a in p) * math.factorial(d) / math.factorial(sum(p) + d)

def orbit_points(kind, params, d):
    """barycentric points of a symmetric orbit."""
    if d == 2:
        if kind == "S3": base = [(1/3, 1/3, 1/3)]
        elif kind == "S21": a = params[0]; base = set(itertools.permutations((a, a, 1 - 2*a)))
        elif kind == "S111": a, b = params; base = set(itertools.permutations((a, b, 1 - a - b)))
    else:
        if kind == "S4": base = [(0.25,)*4]
        elif kind == "S31": a = params[0]; base = set(itertools.permutations((a, a, a, 1 - 3*a)))
        elif kind == "S22": a = params[0]; base = set(itertools.permutations((a, a, 0.5 - a, 0.5 - a)))
        elif kind == "S211": a, b = params; base = set(itertools.permutations((a, a, b, 1 - 2*a - b)))
    return np.array(sorted(base))

NPAR = {"S3": 0, "S21": 1, "S111": 2, "S4": 0, "S31": 1, "S22": 1, "S211": 2}

def build(structure, x, d):
    pts, wts = [], []
    k = 0
    for kind in structure:
        n = NPAR[kind]
        par = x[k:k+n]; w = x[k+n]; k += n + 1
        P = orbit_points(kind, par, d)
        pts.append(P); wts.append(np.full(len(P), w))
    return np.concatenate(pts), np.concatenate(wts)

def residual(x, structure, d, deg):
    P, W = build(structure, x, d)
    r = []
    for p in monomials(d, deg):
        val = np.sum(W * np.prod(P[:, 1:] ** np.array(p), axis=1))
        r.append(val - exact(p))
    return np.array(r)

def solve(structure, d, deg, x0=None, tries=200, seed=0):
    rng = np.random.default_rng(seed)
    nx = sum(NPAR[k] + 1 for k in structure)
    for t in range(tries):
        x = np.array(x0, float) if (x0 is not None and t == 0) else rng.uniform(0.02, 0.45, nx)
        sol = least_squares(residual, x, args=(structure, d, deg), xtol=3e-16, ftol=1e-30, gtol=1e-30, max_nfev=2000)
        P, W = build(structure, sol.x, d)
        if np.abs(sol.fun).max() < 1e-14 and (W > 0).all() and (P > 0).all():
            # polish with a few Gauss-Newton steps in extended precision is unnecessary: residual < 1e-14
            return sol.x, P, W
    return None

if __name__ == "__main__":
    jobs = [
        ("tri", 2, 3, ["S21", "S21"], None),
        ("tri", 2, 4, ["S21", "S21"], [0.445948490915965, 0.223381589678011, 0.091576213509771, 0.109951743655322]),
        ("tri", 2, 5, ["S3", "S21", "S21"], [0.225, 0.470142064105115, 0.132394152788506, 0.101286507323456, 0.125939180544827]),
        ("tri", 2, 6, ["S21", "S21", "S111"], [0.249286745170910, 0.116786275726379, 0.063089014491502, 0.050844906370207, 0.053145049844817, 0.310352451033784, 0.082851075618374]),
        ("tet", 3, 3, ["S22"], None),
        ("tet", 3, 3, ["S31", "S31"], None),
        ("tet", 3, 4, ["S31", "S31", "S22"], None),
        ("tet", 3, 5, ["S31", "S31", "S22"], None),
        ("tet", 3, 6, ["S31", "S31", "S31", "S211"], [0.214602871259152, 0.039922750258168, 0.040673958534611, 0.010077211055321, 0.322337890142276, 0.055357181543654, 0.063661001875018, 0.269672331458316, 0.048214285714286]),
    ]
    np.set_printoptions(precision=17, linewidth=200)
    for name, d, deg, structure, x0 in jobs:
        out = solve(structure, d, deg, x0)
        if out is None:
            print(f"# {name} degree {deg} {structure}: no solution"); continue
        x, P, W = out
        err = np.abs(residual(x, structure, d, deg)).max()
        err_next = np.abs(residual(x, structure, d, deg + 1)).max()
        print(f"# {name} degree {deg}: {len(W)} points, structure {structure}, max moment error {err:.1e} (degree {deg+1}: {err_next:.1e})")
        print(f"({name!r}, {deg}): (")
        print("    " + repr(P.tolist()) + ",")
        print("    " + repr(W.tolist()) + "),")
