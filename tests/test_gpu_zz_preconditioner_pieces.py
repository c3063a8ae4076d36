"""GPU: the pieces of the polynomial preconditioner through the C ABI, against numpy on the matrix the same
library assembles - block-Jacobi inversion in shared-memory slabs (`fb_bjacobi_setup`), its application from
the 4-byte inverse blocks (`fb_precond`), and one step of the Chebyshev recurrence on D^-1 J fused into the
channel SpMV (`fb_chebyshev_step`).  The inverse blocks are rounded to single-precision mantissas times a
power-of-two row scale, so they are compared with the exact inverse at 2^-23 relative to the row maximum;
everything that APPLIES them must agree with numpy applying the same rounded blocks to 1e-12."""
import ctypes as C

import numpy as np
import pytest

from cases import multi_isotope_traps_3d

pytestmark = pytest.mark.gpu


def _setup(n=4):
    m, _, _ = multi_isotope_traps_3d(n, final_time=3e-2)
    m.initialise()
    rng = np.random.default_rng(5)
    N = m.u.x.array.size
    m.u.x.array[:] = rng.uniform(0.5, 1.5, N) * 1e18
    m.u_n.x.array[:] = 0.9 * m.u.x.array
    _, J = m.solver.assemble_host(want_J=True)      # block-CSR values expanded from the channel operator
    s = m.solver
    ctx = s.ctx
    assert s.spec.vx is not None
    ns = s.spec.ns
    ud, und = ctx.nodal_to_device("test_u", m.u.x.array, ns), ctx.nodal_to_device("test_un", m.u_n.x.array, ns)
    Fd = ctx.zeros(N)
    ctx._check(ctx.lib.fb_assemble(ctx.h, ctx.ptr(ud), ctx.ptr(und), ctx.ptr(Fd), None, 3))   # channel operator
    ctx._check(ctx.lib.fb_bjacobi_setup(ctx.h, None))
    return m, ctx, J.tocsr(), N, rng


def _rounded_block_inverse(J, ns):
    """numpy model of k_bjacobi_setup_slab's result: inverse of every diagonal block, rows rounded to
    float mantissas times a power-of-two scale."""
    nv = J.shape[0] // ns
    out = np.zeros((nv, ns, ns))
    exact = np.zeros((nv, ns, ns))
    for v in range(nv):
        blk = J[v * ns:(v + 1) * ns, v * ns:(v + 1) * ns].toarray()
        inv = np.linalg.inv(blk)
        exact[v] = inv
        mx = np.abs(inv).max(axis=1)
        ex = np.where(mx > 0, np.frexp(mx)[1], 0)
        scale = np.ldexp(1.0, ex)[:, None]
        out[v] = (inv / scale).astype(np.float32).astype(np.float64) * scale
    return out, exact


def test_block_jacobi_application_from_the_single_precision_blocks():
    m, ctx, J, N, rng = _setup()
    ns = m.solver.spec.ns
    rounded, exact = _rounded_block_inverse(J, ns)
    r = rng.standard_normal(N)
    rd, zd = ctx.nodal_to_device("test_r", r, ns), ctx.zeros(N)
    ctx._check(ctx.lib.fb_precond(ctx.h, ctx.ptr(rd), ctx.ptr(zd)))
    z = ctx.nodal_from_device(zd, ns).reshape(-1, ns)
    z_exact = np.einsum("vsc,vc->vs", exact, r.reshape(-1, ns))
    # against the exact inverse: single-precision rounding of the blocks
    assert np.abs(z - z_exact).max() <= 3e-6 * np.abs(z_exact).max()
    # the device inversion differs from numpy's by double round-off before the rounding to float, so a few
    # mantissas may fall on the other side: compare with the rounded model at one float ulp of the row
    z_model = np.einsum("vsc,vc->vs", rounded, r.reshape(-1, ns))
    assert np.abs(z - z_model).max() <= 2e-7 * np.abs(z_exact).max()


@pytest.mark.parametrize("n", [4, 12])
def test_one_step_of_the_chebyshev_recurrence(n):
    m, ctx, J, N, rng = _setup(n)
    ns = m.solver.spec.ns
    # the operator the kernel applies: D^-1 with the device's own blocks (through fb_precond on unit vectors
    # would be N solves; instead apply fb_precond to J d computed by numpy)
    d = rng.standard_normal(N)
    r0 = rng.standard_normal(N)
    z0 = rng.standard_normal(N)
    ca, cb = 0.3, 0.7
    dd, rd, zd, od = (ctx.nodal_to_device("test_d", d, ns), ctx.nodal_to_device("test_r", r0, ns),
                      ctx.nodal_to_device("test_z", z0, ns), ctx.zeros(N))
    ctx._check(ctx.lib.fb_chebyshev_step(ctx.h, ctx.ptr(dd), ctx.ptr(rd), ctx.ptr(zd), ctx.ptr(od), ca, cb))
    Jd = ctx.nodal_to_device("test_Jd", J @ d, ns)
    Bd = ctx.zeros(N)
    ctx._check(ctx.lib.fb_precond(ctx.h, ctx.ptr(Jd), ctx.ptr(Bd)))      # D^-1 (J d) with the same blocks
    Bd = ctx.nodal_from_device(Bd, ns)
    r1 = r0 - Bd
    d1 = ca * d + cb * r1
    z1 = z0 + d1
    scale = np.abs(Bd).max()
    assert np.abs(ctx.nodal_from_device(rd, ns) - r1).max() <= 1e-11 * scale
    assert np.abs(ctx.nodal_from_device(od, ns) - d1).max() <= 1e-11 * scale
    assert np.abs(ctx.nodal_from_device(zd, ns) - z1).max() <= 1e-11 * scale


def test_channel_spmv_with_the_register_plan_matches_the_block_csr_product():
    """y = J x on the channel operator (TMA + DMMA kernel with the per-lane epilogue plan) against scipy on
    the expanded block-CSR values, on a mesh with more block rows than one chunk per SM."""
    m, ctx, J, N, rng = _setup(n=12)
    x = rng.standard_normal(N)
    ns = m.solver.spec.ns
    xd, yd = ctx.nodal_to_device("test_x", x, ns), ctx.zeros(N)
    ctx._check(ctx.lib.fb_spmv(ctx.h, None, ctx.ptr(xd), ctx.ptr(yd)))
    ref = J @ x
    assert np.abs(ctx.nodal_from_device(yd, ns) - ref).max() <= 1e-12 * np.abs(ref).max()


def test_stored_channels_are_reused_only_at_the_same_solution(monkeypatch):
    """The first assembly of a time step happens at the solution whose solution-dependent channels the last
    residual check left in memory: the library streams them instead of contracting the edge tensors again.
    Same solution -> same residual as a full assembly; a changed solution -> channels recomputed (residual and
    operator equal to those of a context that never reuses)."""
    m, ctx, J, N, rng = _setup(n=6)
    ns = m.solver.spec.ns
    u1 = m.u.x.array.copy()
    un = m.u_n.x.array.copy()
    u2 = u1.copy()
    u2[::7] *= 2.0
    x = rng.standard_normal(N)
    xd = ctx.nodal_to_device("test_x", x, ns)

    def assemble(uh):
        ud, und = ctx.nodal_to_device("test_u", uh, ns), ctx.nodal_to_device("test_un", un, ns)
        Fd, yd = ctx.zeros(N), ctx.zeros(N)
        ctx._check(ctx.lib.fb_assemble(ctx.h, ctx.ptr(ud), ctx.ptr(und), ctx.ptr(Fd), None, 3))
        ctx._check(ctx.lib.fb_spmv(ctx.h, None, ctx.ptr(xd), ctx.ptr(yd)))
        return ctx.nodal_from_device(Fd, ns), ctx.nodal_from_device(yd, ns)

    F_first, y_first = assemble(u1)          # _setup assembled at u1 already: this call reuses
    F_again, y_again = assemble(u1)
    F_moved, y_moved = assemble(u2)          # must notice the change
    monkeypatch.setenv("FB_NO_DYN_REUSE", "1")
    F_ref1, y_ref1 = assemble(u1)
    F_ref2, y_ref2 = assemble(u2)
    for a, b in ((F_first, F_ref1), (F_again, F_ref1), (F_moved, F_ref2)):
        assert np.abs(a - b).max() <= 1e-13 * np.abs(b).max()
    for a, b in ((y_first, y_ref1), (y_again, y_ref1), (y_moved, y_ref2)):
        assert np.abs(a - b).max() <= 1e-13 * np.abs(b).max()
    # the operator does depend on the solution by far more than the tolerance above: stale channels would show
    assert np.abs(y_ref2 - y_ref1).max() > 1e-10 * np.abs(y_ref1).max()
    assert np.abs(F_ref2 - F_ref1).max() > 1e-6 * np.abs(F_ref1).max()
