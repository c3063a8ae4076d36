"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo processes check
the partition, the owner->ghost halo exchange and that owned-row SpMV + halo +
all-reduce reproduce the global operator (no CUDA involved)."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import festim_b200 as F
from festim_b200.distributed import HaloExchange, LocalPartition, allreduce_sum, partition_rcb


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, dim, ns, p2p=False):
    if p2p:
        os.environ["FB_HALO_P2P"] = "1"
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mesh = F.create_unit_square(9, 7) if dim == 2 else F.create_unit_cube(5, 4, 6)
        part = partition_rcb(mesh.coords, world)
        lp = LocalPartition(mesh, part, rank, world)
        # 1. ownership is a partition of the vertices
        owned = [None] * world
        dist.all_gather_object(owned, lp.owned)
        allv = np.concatenate(owned)
        assert len(allv) == mesh.num_vertices and len(np.unique(allv)) == mesh.num_vertices
        sizes = [len(o) for o in owned]
        assert max(sizes) - min(sizes) <= 1
        # 2. halo exchange delivers the owner's values to every ghost
        halo = HaloExchange(lp, ns, torch, "cpu")
        f = lambda g: np.sin(g[:, None] * 0.37 + np.arange(ns)[None, :])
        vec = torch.zeros(len(lp.l2g) * ns, dtype=torch.float64)
        vec.view(-1, ns)[: lp.n_owned] = torch.from_numpy(f(lp.owned))
        halo(vec)
        assert np.array_equal(vec.view(-1, ns).numpy(), f(lp.l2g))
        # 3. owned rows of a global operator only reference owned + ghost vertices, and the
        #    local SpMV after the halo equals the global one
        rowptr, colidx = mesh.graph
        rng = np.random.default_rng(7)
        A = sp.csr_matrix((rng.standard_normal(len(colidx)), colidx, rowptr),
                          shape=(mesh.num_vertices,) * 2)
        x = rng.standard_normal(mesh.num_vertices)
        rows = A[lp.owned]
        assert np.all(lp.g2l[rows.indices] >= 0)
        xl = torch.zeros(len(lp.l2g), dtype=torch.float64)
        xl[: lp.n_owned] = torch.from_numpy(x[lp.owned])
        HaloExchange(lp, 1, torch, "cpu")(xl)
        y_local = rows[:, lp.l2g] @ xl.numpy()
        assert np.allclose(y_local, (A @ x)[lp.owned], rtol=1e-14, atol=1e-14)
        # 4. all-reduced owned dot products equal the global ones
        tot = allreduce_sum([float(x[lp.owned] @ x[lp.owned])], torch, "cpu")[0]
        assert abs(tot - x @ x) < 1e-12 * (x @ x)
        # 5. every local cell touches an owned vertex; local facets touch an owned vertex
        assert (lp.mesh.cells < lp.n_owned).any(axis=1).all()
        assert (lp.mesh.bfacets < lp.n_owned).any(axis=1).all()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dim,ns,p2p", [(2, 1, False), (3, 2, False), (3, 2, True)])
def test_partition_halo_spmv_world2(dim, ns, p2p):
    mp.spawn(_worker, args=(2, _free_port(), dim, ns, p2p), nprocs=2, join=True)


def test_halo_all_gather_world3():
    mp.spawn(_worker, args=(3, _free_port(), 3, 1, False), nprocs=3, join=True)


def test_rcb_balanced_any_count():
    mesh = F.create_unit_cube(6, 5, 4)
    for nparts in (1, 2, 3, 4, 5, 8):
        part = partition_rcb(mesh.coords, nparts)
        counts = np.bincount(part, minlength=nparts)
        assert counts.sum() == mesh.num_vertices and counts.max() - counts.min() <= 1


@pytest.mark.parametrize("world", [2, 3])
def test_sm_block_ordering_keeps_halo_lists_consistent(world):
    """Second-level (per-SM) RCB ordering of the owned vertices: block ranges cover the
    owned vertices, and every send list still addresses the neighbour's ghost segment in
    the same global-id order (what the push-style peer halo relies on)."""
    from festim_b200.mesh import create_unit_cube

    mesh = create_unit_cube(7, 6, 5)
    part = partition_rcb(mesh.coords, world)
    lps = [LocalPartition(mesh, part, r, world, sm_blocks=8) for r in range(world)]
    for p in lps:
        assert p.block_ranges is not None and p.block_ranges[0] == 0 and p.block_ranges[-1] == p.n_owned
        assert sorted(p.l2g[: p.n_owned]) == sorted(np.nonzero(part == p.rank)[0])
        # every block is a contiguous index range whose vertices are spatially compact
        plain = LocalPartition(mesh, part, p.rank, world)
        assert np.array_equal(p.ghosts, plain.ghosts)
        for q, idx in p.send.items():
            a, b = lps[q].recv[p.rank]
            assert np.array_equal(p.l2g[idx], lps[q].l2g[a:b])


def _gmres_worker(rank, world, port, ns):
    """Partitioned GMRES (festim_b200/krylov_dist.py) on a non-symmetric block operator: the
    recurrence, halo placement and all-reduces are the ones the GPU ranks run; here the owned-row
    SpMV and the block-Jacobi preconditioner are scipy / numpy."""
    from festim_b200.krylov_dist import gmres

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mesh = F.create_unit_cube(6, 5, 4)
        nv = mesh.num_vertices
        part = partition_rcb(mesh.coords, world)
        lp = LocalPartition(mesh, part, rank, world)
        rowptr, colidx = mesh.graph
        rng = np.random.default_rng(11)     # same stream on every rank: the same global operator
        Ag = sp.csr_matrix((np.ones(len(colidx)), colidx, rowptr), shape=(nv, nv))
        # node-block operator: non-symmetric off-diagonal couplings, dominant diagonal blocks
        B = sp.kron(Ag, np.ones((ns, ns))).tocsr()
        B.data = rng.uniform(-1.0, 1.0, B.nnz)
        B = B + sp.diags(np.asarray(abs(B).sum(axis=1)).ravel() * 0.6 + 1.0)
        B = B.tocsr()
        assert abs(B - B.T).max() > 0.1
        rhs = rng.standard_normal(nv * ns)
        exact = sp.linalg.spsolve(B.tocsc(), rhs)

        ldofs = lp.local_dofs(ns)
        n_own = lp.n_owned * ns
        rows = B[ldofs[:n_own]][:, ldofs]                      # owned rows, local columns
        dinv = [np.linalg.inv(rows[v * ns:(v + 1) * ns, v * ns:(v + 1) * ns].toarray())
                for v in range(lp.n_owned)]
        halo = HaloExchange(lp, ns, torch, "cpu")
        calls = {"A": 0, "reduce": 0}

        def apply_A(v):
            calls["A"] += 1
            halo(v)
            return torch.from_numpy(rows @ v.numpy())

        def apply_M(r):
            z = np.concatenate([d @ r.numpy()[v * ns:(v + 1) * ns] for v, d in enumerate(dinv)])
            return torch.from_numpy(z)

        def allreduce(t):
            calls["reduce"] += 1
            dist.all_reduce(t)
            return t

        b = torch.zeros(len(ldofs), dtype=torch.float64)
        b[:n_own] = torch.from_numpy(rhs[ldofs[:n_own]])
        x = torch.zeros_like(b)
        its, ok = gmres(apply_A, apply_M, allreduce, b, x, n_own, rtol=1e-12, max_it=500, restart=12)
        assert ok and 12 < its < 200, (its, ok)                 # more than one restart cycle
        assert np.allclose(x.numpy()[:n_own], exact[ldofs[:n_own]], rtol=1e-9, atol=1e-11)
        # one halo'd SpMV per Arnoldi step (+1 per restart), three reductions per step (+1 per restart)
        assert calls["A"] <= its + its // 12 + 2 and calls["reduce"] <= 3 * its + its // 12 + 3
        every = [None] * world
        dist.all_gather_object(every, its)
        assert len(set(every)) == 1, "all ranks must take the same decisions"
        # iteration cap is honoured and reported as not converged
        x2 = torch.zeros_like(b)
        its2, ok2 = gmres(apply_A, apply_M, allreduce, b, x2, n_own, rtol=1e-14, max_it=5, restart=12)
        assert its2 == 5 and not ok2
        # zero right-hand side: converged at once, x untouched
        x3 = torch.zeros_like(b)
        its3, ok3 = gmres(apply_A, apply_M, allreduce, torch.zeros_like(b), x3, n_own, rtol=1e-10)
        assert its3 == 0 and ok3 and not x3.any()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,ns", [(1, 2), (2, 1), (2, 3), (3, 2)])
def test_partitioned_gmres(world, ns):
    mp.spawn(_gmres_worker, args=(world, _free_port(), ns), nprocs=world, join=True)
