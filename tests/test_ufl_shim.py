"""Expression shim: numpy evaluation, symbolic derivatives, UFL degree estimates
 and the byte-code the CUDA kernels interpret (checked here
with a reference interpreter in Python; the device interpreter is checked against
the oracle by the GPU assembly parity tests)."""
import math

import numpy as np
import pytest

import festim_b200 as F
from festim_b200 import fem, ufl


def run_program(pb, pid, x, T=0.0, species=None, scalars=None):
    ops, iargs, fargs, offs = pb.tables()
    st, tmp = [], {}
    for pc in range(offs[pid], offs[pid + 1]):
        op, ia, fa = int(ops[pc]), int(iargs[pc]), float(fargs[pc])
        if op == ufl.OP_CONST: st.append(fa)
        elif op == ufl.OP_X: st.append(x[ia])
        elif op == ufl.OP_SCALAR: st.append(float(pb.scalars[ia].value))
        elif op == ufl.OP_T: st.append(T)
        elif op == ufl.OP_SPECIES: st.append(species[ia])
        elif op == ufl.OP_LOAD: st.append(tmp[ia])
        elif op == ufl.OP_STORE: tmp[ia] = st[-1]
        elif op == ufl.OP_NEG: st[-1] = -st[-1]
        elif op == ufl.OP_MATH: st[-1] = float(ufl._NP_FUNCS[ufl.Math.FUNCS[ia]](st[-1]))
        elif op == ufl.OP_NOT: st[-1] = float(st[-1] == 0.0)
        elif op == ufl.OP_SELECT:
            b = st.pop(); a = st.pop(); c = st.pop(); st.append(a if c != 0.0 else b)
        else:
            b = st.pop(); a = st.pop()
            st.append({
                ufl.OP_ADD: lambda: a + b, ufl.OP_MUL: lambda: a * b, ufl.OP_DIV: lambda: a / b,
                ufl.OP_POW: lambda: a ** b, ufl.OP_LT: lambda: float(a < b), ufl.OP_GT: lambda: float(a > b),
                ufl.OP_LE: lambda: float(a <= b), ufl.OP_GE: lambda: float(a >= b),
                ufl.OP_EQ: lambda: float(a == b), ufl.OP_NE: lambda: float(a != b),
                ufl.OP_AND: lambda: float(a != 0 and b != 0), ufl.OP_OR: lambda: float(a != 0 or b != 0),
                ufl.OP_MAX: lambda: max(a, b), ufl.OP_MIN: lambda: min(a, b)}[op]())
    assert len(st) == 1
    return st[0]


def exprs(x, t):
    T = 500 + 100 * x[0]
    D = 1.3 * ufl.exp(-0.1 / (F.k_B * T))
    u = 1 + ufl.sin(2 * ufl.pi * x[0]) + ufl.cos(2 * ufl.pi * x[1]) * x[2] ** 2
    return [
        -ufl.div(D * ufl.grad(u)),
        ufl.conditional(ufl.lt(x[0], 0.5), 2.0 * x[1], ufl.sqrt(1 + x[2])) + t * 3,
        ufl.max_value(x[0], x[1]) / (1 + abs(x[2] - 0.3)) + ufl.tanh(x[0]) ** 3,
        4 * t - ufl.erf(x[0]) * ufl.ln(2 + x[1]) + 2 ** x[2],
    ]


def test_bytecode_matches_numpy_evaluation():
    x = ufl.SpatialCoordinate(3)
    t = fem.Constant(None, 0.7)
    rng = np.random.default_rng(3)
    pts = rng.uniform(0.05, 0.95, (20, 3))
    pb = ufl.ProgramBuilder()
    for ex in exprs(x, t):
        pid = pb.add(ex)
        ref = np.broadcast_to(ufl.evaluate(ex, {"x": pts}), (len(pts),))
        for p, r in zip(pts, ref):
            assert run_program(pb, pid, p) == pytest.approx(r, rel=1e-13, abs=1e-13)


def test_cse_shrinks_repeated_subtrees():
    x = ufl.SpatialCoordinate(3)
    D = ufl.exp(-0.1 / (F.k_B * (500 + 100 * x[0])))
    f = -ufl.div(D * ufl.grad(1 + ufl.sin(2 * ufl.pi * x[0]) + ufl.cos(2 * ufl.pi * x[1])))
    pb = ufl.ProgramBuilder()
    pb.add(f)
    ops = pb.tables()[0]
    n_exp = sum(1 for o, i in zip(ops, pb.tables()[1]) if o == ufl.OP_MATH and ufl.Math.FUNCS[i] == "exp")
    assert n_exp == 1


def test_symbolic_derivatives_against_finite_differences():
    x = ufl.SpatialCoordinate(3)
    t = fem.Constant(None, 0.3)
    pts = np.random.default_rng(5).uniform(0.1, 0.9, (10, 3))
    h = 1e-6
    for ex in exprs(x, t)[2:]:
        for i in range(3):
            d = ufl.evaluate(ufl.diff(ex, ("x", i)), {"x": pts})
            e = np.zeros(3); e[i] = h
            fd = (ufl.evaluate(ex, {"x": pts + e}) - ufl.evaluate(ex, {"x": pts - e})) / (2 * h)
            assert np.allclose(d, fd, rtol=1e-5, atol=1e-7)


def test_species_derivative_of_reaction_term():
    A, B = F.Species("A"), F.Species("B", mobile=False)
    A.solution, B.solution = ufl.SpeciesRef(A), ufl.SpeciesRef(B)
    imp = F.ImplicitSpecies(n=3.0, others=[B])
    imp.value_fenics = fem.Constant(None, 3.0)
    vol = F.VolumeSubdomain(id=1, material=F.Material(D_0=1, E_D=0))
    r = F.Reaction(reactant=[A, imp], product=B, k_0=2.0, E_k=0.2, p_0=0.5, E_p=0.1, volume=vol)
    term = r.reaction_term(600.0)
    env = {"species": lambda s: {A: 1.7, B: 0.4}[s]}
    k = 2.0 * math.exp(-0.2 / (F.k_B * 600))
    p = 0.5 * math.exp(-0.1 / (F.k_B * 600))
    assert ufl.evaluate(ufl.diff(term, ("species", A)), env) == pytest.approx(k * (3.0 - 0.4))
    assert ufl.evaluate(ufl.diff(term, ("species", B)), env) == pytest.approx(-k * 1.7 - p)


def test_degree_estimates_follow_ufl_rules():
    """SURVEY Appendix A.3 table."""
    x = ufl.SpatialCoordinate(3)
    Tc = fem.Constant(None, 500.0)
    mesh = F.create_unit_cube(1, 1, 1)
    Tf = fem.Function(fem.functionspace(mesh, ("Lagrange", 1)))
    A, B = F.Species("A"), F.Species("B")
    cA, cB = ufl.SpeciesRef(A), ufl.SpeciesRef(B)
    arr = lambda T: 2.0 * ufl.exp(-0.3 / (F.k_B * T))
    assert ufl.estimate_degree(arr(Tc)) == 0
    assert ufl.estimate_degree(arr(Tf)) == 3
    assert ufl.estimate_degree(arr(Tf) * cA * (5.0 - cB)) + 1 == 6   # reaction, T field
    assert ufl.estimate_degree(arr(Tc) * cA * (5.0 - cB)) + 1 == 3   # reaction, T constant
    assert ufl.estimate_degree((cA - ufl.SpeciesRef(A, prev=True)) / fem.Constant(None, 0.1)) + 1 == 2
    assert ufl.estimate_degree(x[0] ** 3) == 3 and ufl.estimate_degree(x[0] ** 0.5) == 3
    assert ufl.estimate_degree(ufl.sin(x[0])) == 3 and ufl.estimate_degree(ufl.sin(Tc)) == 0
    assert ufl.estimate_degree(ufl.conditional(ufl.lt(x[0], 1), x[0] ** 2, x[1])) == 2
    assert ufl.estimate_degree(ufl.grad(Tf * 2.0, 3)[0]) == 0


def test_structural_equality_like_ufl():
    """`==` on expressions is UFL's structural equality returning a bool (the reference's unit
    tests compare `reaction.reaction_term(T) == expected` this way); leaves compare by identity,
    literals by value; the condition for conditional() is `eq(a, b)`."""
    from festim_b200 import fem, ufl
    from festim_b200.mesh import create_unit_interval

    V = fem.functionspace(create_unit_interval(4), ("Lagrange", 1))
    f, g = fem.Function(V), fem.Function(V)
    c1, c2 = fem.Constant(None, 2.0), fem.Constant(None, 2.0)
    assert (2.0 * f * ufl.exp(-g / 3.0) - c1 * g) == (2.0 * f * ufl.exp(-g / 3.0) - c1 * g)
    assert (f * g) != (g * f)
    assert (f + 1.0) != (f + 2.0)
    assert (f * c1) != (f * c2)            # two constants with the same value stay different leaves
    assert f == f and f != g and not (f == 3) and not (f == "f") and f != None   # noqa: E711
    assert ufl.conditional(ufl.lt(f, 1.0), f, g) == ufl.conditional(ufl.lt(f, 1.0), f, g)
    assert ufl.conditional(ufl.lt(f, 1.0), f, g) != ufl.conditional(ufl.gt(f, 1.0), f, g)
    assert isinstance(ufl.eq(f, g), ufl.Condition)
    assert len({f: 1, g: 2}) == 2 and f in [g, f]
