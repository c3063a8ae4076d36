"""numpy model of the channel assembly (festim_b200/csrc/assemble_edge.cuh): the per-cell
record, the channels and the coefficient tables of ``KernelSpec.vx`` are evaluated exactly as
the device kernels do (cell-vectorised instead of warp-per-vertex) and scattered into a sparse
matrix.  Test infrastructure: it checks the *lowering* (tables) against the oracle on the CPU,
so that the CUDA kernels only have to be a transcription of this file.
"""
import itertools

import numpy as np
import scipy.sparse as sp

from festim_b200 import fem, lowering as L
from festim_b200.constants import k_B


def _geometry(mesh):
    X = mesh.coords[mesh.cells]                       # (nc, d1, d)
    d = mesh.tdim
    Jm = np.transpose(X[:, 1:, :] - X[:, :1, :], (0, 2, 1))   # columns = edges
    det = np.linalg.det(Jm)
    inv = np.linalg.inv(Jm)                           # rows = grad of phi_1..d
    G = np.concatenate([-inv.sum(axis=1, keepdims=True), inv], axis=1)   # (nc, d1, d)
    fact = {1: 1.0, 2: 2.0, 3: 6.0}[d]
    return np.abs(det) / fact, G


def cell_records(spec, T_nodal_or_const, velocities=()):
    """(n_cells, rec) array: what k_cell_records writes."""
    from festim_b200 import quadrature

    mesh, vx = spec.mesh, spec.vx
    d1 = mesh.tdim + 1
    vol, G = _geometry(mesh)
    rec = np.zeros((mesh.num_cells, vx["rec"]))
    rec[:, 0] = vol
    sym = [(a, b) for a in range(d1) for b in range(a, d1)]
    sym3 = [(a, b, c) for a in range(d1) for b in range(a, d1) for c in range(b, d1)]
    mscale = 1.0 / (d1 * (d1 + 1))
    for tag, vid in enumerate(spec.vol_ids):
        sel = np.nonzero(spec.cell_tag == tag)[0]
        cells = mesh.cells[sel]
        dF, dJ = spec.degrees.get(("dx", vid), (1, None))
        rules = [quadrature.rule(mesh.tdim, dF), quadrature.rule(mesh.tdim, dF if dJ is None else dJ)]
        for kind, a, r, off in spec.vx["cq"][tag]:
            bary, w = rules[r]
            if np.ndim(T_nodal_or_const) == 0:
                Tq = np.full((len(sel), len(w)), float(T_nodal_or_const))
            else:
                Tq = T_nodal_or_const[cells] @ bary.T          # (nsel, nq)
            if kind == L.CQ_KD:
                assert a < len(spec.diff_energies), "emulator: Arrhenius diffusivities only"
                dbar = (np.exp(-spec.diff_energies[a] / (k_B * Tq)) * w).sum(axis=1)
                for k, (i, j) in enumerate(sym):
                    rec[sel, off + k] = vol[sel] * dbar * np.einsum("ck,ck->c", G[sel, i], G[sel, j])
            elif kind in (L.CQ_M1, L.CQ_M3):
                rho = np.exp(-spec.rslot_E[a] / (k_B * Tq)) * w       # (nsel, nq)
                idx = sym if kind == L.CQ_M1 else sym3
                for k, ids in enumerate(idx):
                    phi = np.prod(bary[:, list(ids)], axis=1)
                    rec[sel, off + k] = vol[sel] * (rho @ phi)
            elif kind == L.CQ_ADV:
                vel = velocities[a].reshape(-1, mesh.tdim)[cells]     # (nsel, d1, d)
                vg = np.einsum("cbk,cjk->cbj", vel, G[sel])           # vel_b . G_j
                for i in range(d1):
                    wgt = np.ones(d1)
                    wgt[i] = 2.0
                    val = np.einsum("b,cbj->cj", wgt, vg) * (vol[sel] * mscale)[:, None]
                    for j in range(d1):
                        rec[sel, off + i * d1 + j] = val[:, j]
    return rec


def _sym_index(d1):
    idx = {}
    for k, (a, b) in enumerate((a, b) for a in range(d1) for b in range(a, d1)):
        idx[(a, b)] = idx[(b, a)] = k
    return idx


def _sym3_index(d1):
    idx = {}
    for k, t in enumerate((a, b, c) for a in range(d1) for b in range(a, d1) for c in range(b, d1)):
        for perm in itertools.permutations(t):
            idx[perm] = k
    return idx


def neighbour_table(spec, u, un, scalars):
    """(n_vertices, nbw): [u | u - u_n | affine factors | 1] per vertex."""
    ns, vx = spec.ns, spec.vx
    U = u.reshape(-1, ns)
    Un = un.reshape(-1, ns) if spec.transient else np.zeros_like(U)
    nb = np.zeros((U.shape[0], vx["nbw"]))
    nb[:, :ns] = U
    nb[:, ns:2 * ns] = U - Un
    st = spec.struct or dict(fac_rows=[], dev_factors=[])
    for col, f in enumerate(st["dev_factors"]):
        kind, scal, a0, coefs = st["fac_rows"][f]
        val = np.full(U.shape[0], scalars[scal] if kind else a0)
        for spe, c in coefs:
            val = val + c * U[:, spe]
        nb[:, 2 * ns + col] = val
    nb[:, -1] = 1.0
    return nb


def channels(spec, rec, nb):
    """(n_cells, nch, d1, d1) channel matrices of every cell (zero where the cell's tag is not
    the channel's)."""
    mesh = spec.mesh
    d1 = mesh.tdim + 1
    s2, s3 = _sym_index(d1), _sym3_index(d1)
    ch = spec.vx["ch"]
    out = np.zeros((mesh.num_cells, len(ch), d1, d1))
    mscale = 1.0 / (d1 * (d1 + 1))
    for c, (tag, ctype, off, fac, use, ts, _, _) in enumerate(ch):
        sel = np.nonzero(spec.cell_tag == tag)[0]
        cells = mesh.cells[sel]
        for i in range(d1):
            for j in range(d1):
                if ctype == L.CH_MASS:
                    out[sel, c, i, j] = rec[sel, 0] * mscale * (2.0 if i == j else 1.0)
                elif ctype == L.CH_PAIR:
                    out[sel, c, i, j] = rec[sel, off + s2[(i, j)]]
                elif ctype == L.CH_FULL:
                    out[sel, c, i, j] = rec[sel, off + i * d1 + j]
                else:
                    for b in range(d1):
                        out[sel, c, i, j] += rec[sel, off + s3[(i, j, b)]] * nb[cells[:, b], fac]
    return out


def assemble(spec, u, un, T, inv_dt=0.0, scalars=(), velocities=(), load=None):
    """F (n_dofs) and J (csr) without Dirichlet conditions, in the mesh numbering."""
    mesh, vx, ns = spec.mesh, spec.vx, spec.ns
    d1 = mesh.tdim + 1
    N = mesh.num_vertices * ns
    rec = cell_records(spec, T, velocities)
    nb = neighbour_table(spec, u, un, scalars)
    F = np.zeros(N) if load is None else load.copy()
    rows, cols, vals = [], [], []
    pair_row = np.repeat(np.arange(ns), spec.pair_nc)
    cells = mesh.cells
    C = channels(spec, rec, nb)
    for s in range(ns):
        for chan, gcol, coef, flag in vx["ft"][s]:
            cf = coef * (inv_dt if flag & L.FLAG_INV_DT else 1.0)
            g = nb[cells, gcol]                                   # (nc, d1)
            contrib = cf * np.einsum("cij,cj->ci", C[:, chan], g)
            np.add.at(F, cells * ns + s, contrib)
    for p in range(spec.npairs):
        s, cs = pair_row[p], spec.pair_cs[p]
        for (chan, flag), coef in vx["jt"][p].items():
            cf = coef * (inv_dt if flag & L.FLAG_INV_DT else 1.0)
            for i in range(d1):
                for j in range(d1):
                    rows.append(cells[:, i] * ns + s)
                    cols.append(cells[:, j] * ns + cs)
                    vals.append(cf * C[:, chan, i, j])
    if rows:
        J = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(N, N))
    else:
        J = sp.csr_matrix((N, N))
    return F, J
