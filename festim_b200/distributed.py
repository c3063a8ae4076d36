"""Mesh partitioning and the multi-GPU Newton-CG driver (one process per GPU).

Replaces what the reference inherits from dolfinx/PETSc when run under ``mpirun``
(SURVEY 2.3, 8e): mesh partition, ``IndexMap`` ghost scatter, ``MPI_Allreduce`` in
norms and dot products.  Design:

* vertex-owned decomposition by recursive coordinate bisection (NVSwitch is
  uniform, so the partition need not be topology-aware);
* every rank numbers its owned vertices first and its ghosts after, grouped by
  owner, and keeps **all cells touching an owned vertex**; the gather assembly
  writes owned rows only, so there is no reverse (ADD) scatter of partial sums -
  unlike dolfinx's ``ghostUpdate(ADD, REVERSE)`` - only the forward owner->ghost
  exchange of ``u`` before assembly and of the CG vector before each SpMV;
* the Krylov recurrence is the single-reduction CG of the single-GPU path, split
  at its two communication points (``fb_dcg_*`` in ``include/festim_b200.h``):
  one halo exchange and one all-reduce of three doubles per iteration;
* collectives go through ``torch.distributed`` (NCCL on GPUs, gloo in the CPU
  tests of the host logic); all arithmetic runs in the CUDA library.

Non-symmetric Jacobians (traps, reactions, advection) take the partitioned GMRES of
``festim_b200/krylov_dist.py`` (library SpMV / preconditioner, orthogonalisation as
tensor products - functional, not yet a fused kernel).

Round-1 limits: the host keeps the global problem on every rank (only device work
is partitioned); derived quantities are evaluated on the gathered global field.
"""

from __future__ import annotations

import ctypes as C
import os
import types

import numpy as np

from festim_b200 import fem
from festim_b200.mesh import MeshTags, SimplexMesh


# --------------------------------------------------------------------------
# partitioning (pure numpy: also exercised by the gloo CPU tests)
# --------------------------------------------------------------------------
def partition_rcb(coords, nparts):
    """Recursive coordinate bisection; returns the part id of every vertex.
    Splits are proportional, so any ``nparts`` gives balanced parts."""
    part = np.zeros(len(coords), dtype=np.int32)

    def split(idx, first, count):
        if count == 1:
            part[idx] = first
            return
        left = count // 2
        ext = coords[idx].max(axis=0) - coords[idx].min(axis=0)
        axis = int(np.argmax(ext))
        order = idx[np.argsort(coords[idx, axis], kind="stable")]
        cut = (len(order) * left) // count
        split(order[:cut], first, left)
        split(order[cut:], first + left, count - left)

    split(np.arange(len(coords)), 0, nparts)
    return part


class LocalPartition:
    """Local view of rank ``rank``: numbering, cells, facets, halo lists."""

    def __init__(self, mesh, part, rank, nranks, sm_blocks=None):
        self.rank, self.nranks = rank, nranks
        cells = mesh.cells
        owned_mask = part == rank
        self.owned = np.nonzero(owned_mask)[0].astype(np.int64)
        # optional second level: the owned vertices ordered as ``sm_blocks`` compact RCB
        # parts (one per SM), so that contiguous ranges of the local numbering are compact
        # subdomains the persistent CG can keep in shared memory.  Ghost order, send lists
        # and receive segments are defined through global ids and do not change.
        self.block_ranges = None
        if sm_blocks and mesh.tdim > 1 and len(self.owned) >= 4 * sm_blocks:
            sub = partition_rcb(mesh.coords[self.owned], sm_blocks)
            self.owned = self.owned[np.argsort(sub, kind="stable")]
            counts = np.bincount(sub, minlength=sm_blocks)
            self.block_ranges = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        cell_local = owned_mask[cells].any(axis=1)
        self.local_cells = np.nonzero(cell_local)[0]
        verts = np.unique(cells[self.local_cells].ravel())
        ghosts = verts[~owned_mask[verts]]
        order = np.lexsort((ghosts, part[ghosts]))       # by owner, then global id
        self.ghosts = ghosts[order].astype(np.int64)
        self.ghost_owner = part[self.ghosts]
        self.l2g = np.concatenate([self.owned, self.ghosts])
        self.n_owned = len(self.owned)
        g2l = np.full(mesh.num_vertices, -1, dtype=np.int64)
        g2l[self.l2g] = np.arange(len(self.l2g))
        self.g2l = g2l
        # receive segments: contiguous ranges of the ghost block, one per owner
        self.recv = {}
        for q in np.unique(self.ghost_owner):
            sel = np.nonzero(self.ghost_owner == q)[0]
            self.recv[int(q)] = (self.n_owned + int(sel[0]), self.n_owned + int(sel[-1]) + 1)
        # send lists: my owned vertices that are ghosts of rank q, in q's ghost order
        self.send = {}
        for q in range(nranks):
            if q == rank:
                continue
            q_mask = part == q
            q_cells = q_mask[cells].any(axis=1)
            qv = np.unique(cells[q_cells].ravel())
            mine = qv[owned_mask[qv]]                        # sorted by global id
            if len(mine):
                self.send[q] = g2l[mine]
        # local mesh (exterior facets: the global ones that touch an owned vertex)
        lcells = g2l[cells[self.local_cells]]
        self.mesh = SimplexMesh(mesh.coords[self.l2g], lcells)
        gcell_to_local = np.full(mesh.num_cells, -1, dtype=np.int64)
        gcell_to_local[self.local_cells] = np.arange(len(self.local_cells))
        bf, bfc = mesh.bfacets, mesh.bfacet_cell
        keep = owned_mask[bf].any(axis=1) & (gcell_to_local[bfc] >= 0)
        self.local_bfacets = np.nonzero(keep)[0]
        self.mesh._bfacets = g2l[bf[keep]].astype(np.int32)
        self.mesh._bfacet_cell = gcell_to_local[bfc[keep]].astype(np.int32)

    def local_dofs(self, ns):
        return (self.l2g[:, None] * ns + np.arange(ns)[None, :]).ravel()


class HaloExchange:
    """Owner -> ghost update of a node-major blocked vector over torch.distributed.

    Default: ONE padded all-gather of every rank's boundary values (the owned vertices
    some other rank holds as ghosts) followed by a local gather into the ghost block -
    the halos are a few tens of kB, so the exchange is latency bound and one collective
    beats 2 x (number of neighbours) point-to-point operations.  ``FB_HALO_P2P=1``
    selects grouped NCCL send/recv instead (less data, more launches).
    """

    def __init__(self, part: LocalPartition, ns, torch, device):
        import torch.distributed as dist

        self.p, self.ns, self.torch, self.device = part, ns, torch, device
        self.p2p = bool(os.environ.get("FB_HALO_P2P")) or part.nranks == 1
        self.send_idx = {q: torch.as_tensor(idx, dtype=torch.long, device=device)
                         for q, idx in part.send.items()}
        self.bytes_per_exchange = 8 * ns * sum(len(v) for v in part.send.values())
        if self.p2p:
            return
        # boundary set of this rank, by global id; everybody learns everybody's list
        mine = np.unique(np.concatenate(list(part.send.values()))) if part.send else np.zeros(0, np.int64)
        mine_g = part.l2g[mine]
        order = np.argsort(mine_g)
        mine, mine_g = mine[order], mine_g[order]
        lists = [None] * part.nranks
        dist.all_gather_object(lists, mine_g)
        self.pad = max(max(len(x) for x in lists), 1)
        self.bnd_idx = torch.as_tensor(mine, dtype=torch.long, device=device)
        src = np.zeros(len(part.ghosts), dtype=np.int64)
        for q in range(part.nranks):
            sel = np.nonzero(part.ghost_owner == q)[0]
            if len(sel):
                pos = np.searchsorted(lists[q], part.ghosts[sel])
                assert np.array_equal(np.asarray(lists[q])[pos], part.ghosts[sel])
                src[sel] = q * self.pad + pos
        self.ghost_src = torch.as_tensor(src, dtype=torch.long, device=device)
        self.sendbuf = torch.zeros(self.pad, ns, dtype=torch.float64, device=device)
        self.recvbuf = torch.zeros(self.pad * part.nranks, ns, dtype=torch.float64, device=device)

    def __call__(self, vec):
        import torch.distributed as dist

        if self.p.nranks == 1:
            return
        v2 = vec.view(-1, self.ns)
        if not self.p2p:
            n = self.bnd_idx.numel()
            if n:
                self.torch.index_select(v2, 0, self.bnd_idx, out=self.sendbuf[:n])
            try:
                dist.all_gather_into_tensor(self.recvbuf, self.sendbuf)
            except (RuntimeError, NotImplementedError):
                chunks = list(self.recvbuf.view(self.p.nranks, self.pad, self.ns).unbind(0))
                dist.all_gather(chunks, self.sendbuf)
            if self.ghost_src.numel():
                self.torch.index_select(self.recvbuf, 0, self.ghost_src, out=v2[self.p.n_owned:])
            return
        ops, keep = [], []
        for q, idx in self.send_idx.items():
            buf = v2.index_select(0, idx).contiguous()
            keep.append(buf)
            ops.append(dist.P2POp(dist.isend, buf, q))
        for q, (a, b) in self.p.recv.items():
            ops.append(dist.P2POp(dist.irecv, v2[a:b], q))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()


def allreduce_sum(values, torch, device):
    import torch.distributed as dist

    t = torch.as_tensor(np.asarray(values, dtype=np.float64), device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t)
    return t.cpu().numpy()


# --------------------------------------------------------------------------
# the solver object at the create_solver seam, partitioned
# --------------------------------------------------------------------------
class _Snes:
    def __init__(self):
        self.reason, self.its, self.linear_its, self.fnorm = 0, 0, 0, 0.0

    def getConvergedReason(self):
        return self.reason

    def getIterationNumber(self):
        return self.its

    def getLinearSolveIterations(self):
        return self.linear_its


class DistributedNewtonSolver:
    def __init__(self, problem, petsc_options, device=0, lazy_gather=False):
        self.lazy_gather = lazy_gather
        import torch
        import torch.distributed as dist

        from festim_b200.backend import Context, _np_ptr
        from festim_b200.lowering import KernelSpec

        self.torch, self.dist = torch, dist
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.nranks = dist.get_world_size() if dist.is_initialized() else 1
        self.problem = problem
        gmesh = problem.mesh.mesh
        species = getattr(problem, "species", None) or [problem._unknown]
        self.ns = ns = len(species)
        self.part_of = partition_rcb(gmesh.coords, self.nranks)
        sm_blocks = None
        if self.nranks > 1 and torch.cuda.is_available() and not os.environ.get("FB_NO_SM_PARTITION"):
            sm_blocks = torch.cuda.get_device_properties(device).multi_processor_count
        self.part = lp = LocalPartition(gmesh, self.part_of, self.rank, self.nranks, sm_blocks)
        # local tags
        ctag = np.zeros(gmesh.num_cells, dtype=np.int32)
        ctag[problem.volume_meshtags.indices] = problem.volume_meshtags.values
        ftag = np.zeros(gmesh.bfacets.shape[0], dtype=np.int32)
        ftag[problem.facet_meshtags.indices] = problem.facet_meshtags.values
        vt = MeshTags(lp.mesh, gmesh.tdim, np.arange(len(lp.local_cells)), ctag[lp.local_cells])
        ft = MeshTags(lp.mesh, gmesh.tdim - 1, np.arange(len(lp.local_bfacets)), ftag[lp.local_bfacets])
        view = _ProblemView(problem, lp.mesh, vt, ft)
        self.spec = KernelSpec(view)
        # a dedicated (capturable) stream: library launches, torch copies and NCCL all use it
        torch.cuda.set_device(device)
        self.stream = torch.cuda.Stream(device=device)
        with torch.cuda.stream(self.stream):
            self.ctx = ctx = Context(self.spec, device, renumber=False)  # owned-first numbering is fixed
        ctx._check(ctx.lib.fb_set_owned(ctx.h, lp.n_owned))
        if lp.block_ranges is not None:
            ctx.block_ranges = lp.block_ranges
            ctx._set_sm_partition(*ctx.dmesh.graph)
        self.halo = HaloExchange(lp, ns, torch, ctx.device)
        self.atol = petsc_options["snes_atol"]
        self.rtol = petsc_options["snes_rtol"]
        self.stol = float(petsc_options["snes_stol"])
        self.max_it = int(petsc_options["snes_max_it"])
        snes_rtol = self.rtol if isinstance(self.rtol, (int, float)) else 1e-10
        self.ksp_rtol = float(petsc_options.get("ksp_rtol", min(1e-12, max(1e-2 * snes_rtol, 1e-15))))
        self.ksp_max_it = int(petsc_options.get("ksp_max_it", 20000))
        self.solver = _Snes()
        # Dirichlet dofs restricted to local vertices (owned and ghost columns)
        self._bc_forms = list(problem.bc_forms)
        gd = np.concatenate([bc.dofs for bc in self._bc_forms]) if self._bc_forms else np.zeros(0, np.int64)
        uniq, first_rev = np.unique(gd[::-1], return_index=True)
        pick = len(gd) - 1 - first_rev
        lv = lp.g2l[uniq // ns]
        loc = lv >= 0
        self._bc_pick = pick[loc]
        self._bc_dofs = np.ascontiguousarray(lv[loc] * ns + uniq[loc] % ns, dtype=np.int64)
        ctx._check(ctx.lib.fb_set_dirichlet(ctx.h, len(self._bc_dofs), _np_ptr(self._bc_dofs, C.c_int64)))
        self.device_resident = True
        self.ldofs = lp.local_dofs(ns)
        self.owned_gdofs = self.ldofs[: lp.n_owned * ns]
        self._signature = None
        n = ctx.n_dofs
        z = ctx.zeros
        self.F, self.J, self.y = z(n), z(ctx.nnz), z(n)
        self.r, self.uv, self.w, self.p, self.s, self.tmp = z(n), z(n), z(n), z(n), z(n), z(n)
        self.out3 = z(3)
        self.state = z(8)
        self._graph = None
        # CUDA-graph replay of the iteration: on by default for a single rank; with NCCL
        # traffic inside the capture it is opt-in (FB_DIST_CUDA_GRAPH=1) until validated
        self.use_graph = (self.nranks == 1 and not os.environ.get("FB_NO_CUDA_GRAPH")) or \
            bool(os.environ.get("FB_DIST_CUDA_GRAPH"))
        self.cg_iterations = 0
        self.gmres_iterations = 0
        self.launches_python = 0
        # peer memory (NVLink): the whole-solve persistent CG for symmetric Jacobians, and for the
        # others the library's own Newton-GMRES with halo exchange and cross-rank reductions done by
        # its kernels (fb_newton_solve on a partitioned mesh: no NCCL call inside the solve)
        from festim_b200.backend import PC

        pc = str(petsc_options.get("pc_type", "lu")).lower()
        self.pc = PC.get(pc, 2)
        self.peer = False
        self.native = False
        if self.nranks > 1 and not os.environ.get("FB_NO_PEER_CG"):
            self._setup_peer()
            self.native = self.peer and not self._is_symmetric() and not os.environ.get("FB_NO_NATIVE_GMRES")

    def _setup_peer(self):
        """Exchange CUDA IPC handles of the per-rank exchange areas and build the send
        table (owned vertex -> ghost ordinal on the neighbour) of the push-style halo."""
        from festim_b200.backend import _np_ptr

        ctx, lib, lp, dist = self.ctx, self.ctx.lib, self.part, self.dist
        if self.nranks > 8:
            return
        handle = (C.c_ubyte * 64)()
        ctx._check(lib.fb_peer_alloc(ctx.h, self.rank, self.nranks, handle))
        infos = [None] * self.nranks
        dist.all_gather_object(infos, (bytes(handle), {int(q): seg for q, seg in lp.recv.items()}, int(lp.n_owned)))
        for q, (h, _, _) in enumerate(infos):
            if q != self.rank:
                buf = (C.c_ubyte * 64).from_buffer_copy(h)
                ctx._check(lib.fb_peer_open(ctx.h, q, buf))
        v, rk, dst = [], [], []
        for q, idx in lp.send.items():
            a, b = infos[q][1][self.rank]          # my segment of q's ghost block
            assert b - a == len(idx), "send list and the neighbour's ghost segment disagree"
            v.append(np.asarray(idx, dtype=np.int64))
            rk.append(np.full(len(idx), q, dtype=np.int64))
            dst.append(np.arange(a, b, dtype=np.int64) - infos[q][2])   # ghost ordinal on rank q
        if v:
            v, rk, dst = np.concatenate(v), np.concatenate(rk), np.concatenate(dst)
            order = np.argsort(v, kind="stable")
            v, rk, dst = (np.ascontiguousarray(x[order], dtype=np.int32) for x in (v, rk, dst))
        else:
            v = rk = dst = np.zeros(0, dtype=np.int32)
        ctx._check(lib.fb_peer_set_send(ctx.h, len(v), _np_ptr(v, C.c_int32), _np_ptr(rk, C.c_int32),
                                        _np_ptr(dst, C.c_int32)))
        dist.barrier()                              # every rank has mapped every exchange area
        self.peer = True

    # ---- data movement -----------------------------------------------------------
    def _push_state(self):
        from festim_b200.backend import _np_ptr

        ctx, spec, lp, lib = self.ctx, self.spec, self.part, self.ctx.lib
        scal = spec.scalar_values()
        T = spec.temperature
        # change detection: memcmp against the last uploaded copy (as the single-GPU solver)
        parts = [scal]
        if isinstance(T, fem.Function):
            parts.append(T.x.array)
        elif T is not None:
            parts.append(np.float64(T.value))
        parts += [f.x.array for f in spec.fields]
        parts += [vel.array for _, _, vel in spec.adv]
        bc_vals = np.concatenate([bc.values() for bc in self._bc_forms])[self._bc_pick] \
            if self._bc_forms else np.zeros(0)
        bc_dev = ctx.to_device("bc_vals", bc_vals) if len(bc_vals) else None
        ctx._check(lib.fb_set_dirichlet_values(ctx.h, ctx.ptr(bc_dev)))
        last = self._signature
        if last is not None and len(last) == len(parts) and all(
                np.array_equal(a, b) for a, b in zip(last, parts)):
            return
        sig = [np.array(a, copy=True) for a in parts]
        self._signature = sig
        ctx._check(lib.fb_set_scalars(ctx.h, len(scal), _np_ptr(scal, C.c_double)))
        if isinstance(T, fem.Function):
            ctx._check(lib.fb_set_temperature(ctx.h, 0.0, ctx.ptr(ctx.to_device("T", T.x.array[lp.l2g]))))
        elif T is not None:
            ctx._check(lib.fb_set_temperature(ctx.h, float(T.value), C.c_void_p()))
        for k, f in enumerate(spec.fields):
            ctx._check(lib.fb_set_field(ctx.h, k, ctx.ptr(ctx.to_device(f"field{k}", f.x.array[lp.l2g]))))
        for k, (_, _, vel) in enumerate(spec.adv):
            ctx._check(lib.fb_set_velocity(ctx.h, k, ctx.ptr(ctx.to_device(f"vel{k}", vel.array[lp.l2g]))))
        ctx._check(lib.fb_update_load(ctx.h))

    def _dot(self, a, b):
        out = C.c_double()
        self.ctx._check(self.ctx.lib.fb_dot(self.ctx.h, self.ctx.ptr(a), self.ctx.ptr(b), C.byref(out)))
        return out.value

    def _norm(self, a):
        return float(np.sqrt(allreduce_sum([self._dot(a, a)], self.torch, self.ctx.device)[0]))

    # ---- partitioned CG (block-Jacobi preconditioned, single reduction) -------------
    def _cg_iteration(self, J, x):
        """One CG iteration with device-resident scalars (capturable in a CUDA graph)."""
        ctx, lib, ptr = self.ctx, self.ctx.lib, self.ctx.ptr
        r, u, w, p, s = self.r, self.uv, self.w, self.p, self.s
        self.halo(u)
        lib.fb_dcg_spmv_dots(ctx.h, ptr(J), ptr(u), ptr(r), ptr(w), ptr(self.out3))
        if self.nranks > 1:
            self.dist.all_reduce(self.out3)
        lib.fb_dcg_scalars(ctx.h, ptr(self.out3), ptr(self.state), int(self._max_it))
        lib.fb_dcg_update_dev(ctx.h, ptr(self.state), ptr(x), ptr(r), ptr(u), ptr(w), ptr(p), ptr(s))

    def _gmres(self, J, b, x, rtol, max_it):
        """Non-symmetric Jacobians (traps, reactions, advection): GMRES(30) + block-Jacobi with the
        library's owned-row SpMV and preconditioner, one halo exchange and three all-reduces per
        Arnoldi step (festim_b200/krylov_dist.py)."""
        from festim_b200.krylov_dist import gmres

        ctx, lib, ptr = self.ctx, self.ctx.lib, self.ctx.ptr

        def apply_A(v):
            self.halo(v)
            ctx._check(lib.fb_spmv(ctx.h, ptr(J), ptr(v), ptr(self.w)))
            return self.w

        def apply_M(r):
            ctx._check(lib.fb_precond(ctx.h, ptr(r), ptr(self.tmp)))
            return self.tmp

        def allreduce(t):
            if self.nranks > 1:
                self.dist.all_reduce(t)
            return t

        x.zero_()
        its, ok = gmres(apply_A, apply_M, allreduce, b, x, self.part.n_owned * self.ns, rtol, 0.0, max_it)
        self.gmres_iterations += its
        return its, ok

    def linear_solve(self, J, b, x, rtol, max_it):
        torch, ctx, lib, ptr = self.torch, self.ctx, self.ctx.lib, self.ctx.ptr
        ctx._check(lib.fb_bjacobi_setup(ctx.h, ptr(J)))
        if not self._is_symmetric():
            return self._gmres(J, b, x, rtol, max_it)
        ctx._check(lib.fb_precond(ctx.h, ptr(b), ptr(self.tmp)))
        thr2 = (rtol * self._norm(self.tmp)) ** 2
        if self.peer:
            its, conv = C.c_int(0), C.c_int(0)
            ctx._check(lib.fb_dcg_persistent(ctx.h, ptr(J), ptr(b), ptr(x), float(thr2), float(rtol), int(max_it),
                                             C.byref(its), C.byref(conv)))
            self.cg_iterations += its.value
            return its.value, bool(conv.value)
        r, u, w, p, s = self.r, self.uv, self.w, self.p, self.s
        ctx._check(lib.fb_dcg_init(ctx.h, ptr(b), ptr(x), ptr(r), ptr(u), ptr(p), ptr(s)))
        self._max_it = max_it
        st = torch.zeros(8, dtype=torch.float64)
        st[5] = thr2
        self.state.copy_(st.to(self.state.device))
        key = (J.data_ptr(), x.data_ptr(), max_it)
        if self.use_graph and (self._graph is None or self._graph[0] != key):
            try:
                self._cg_iteration(J, x)               # eager warm-up (NCCL channels, allocator)
                torch.cuda.current_stream().synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self.stream):
                    self._cg_iteration(J, x)
                self._graph = (key, g)
                # the warm-up iteration advanced the recurrence: restart it
                ctx._check(lib.fb_dcg_init(ctx.h, ptr(b), ptr(x), ptr(r), ptr(u), ptr(p), ptr(s)))
                self.state.copy_(st.to(self.state.device))
            except Exception as exc:  # capture not supported here: eager iterations
                import warnings

                warnings.warn(f"CUDA-graph capture of the CG iteration failed ({exc}); running eagerly")
                self.use_graph = False
                self._graph = None
                ctx._check(lib.fb_dcg_init(ctx.h, ptr(b), ptr(x), ptr(r), ptr(u), ptr(p), ptr(s)))
                self.state.copy_(st.to(self.state.device))
        chunk = 16
        while True:
            for _ in range(chunk):
                if self._graph is not None:
                    self._graph[1].replay()
                else:
                    self._cg_iteration(J, x)
            sth = self.state.cpu().numpy()
            if sth[4] != 0.0:
                break
        it = int(sth[3])
        self.cg_iterations += it
        return it, bool(sth[4] == 1.0)

    # ---- Newton (reference problem.py:271-289, helpers.py:408-426) --------------------
    def _sync_fields(self):
        """Local device copies of u and u_n (owned + ghost entries): uploaded only when the host
        arrays were touched since the last solve."""
        ctx, p = self.ctx, self.problem
        ux, unx = p.u.x, p.u_n.x
        if getattr(self, "_u_loc", None) is None or ux.version != self._u_version:
            self._u_loc = ctx.to_device("u", ux.array[self.ldofs])
            self._u_version = ux.version
        if not self.spec.transient:
            return self._u_loc, self._u_loc
        if getattr(self, "_un_loc", None) is None or unx.version != self._un_version:
            self._un_loc = ctx.to_device("u_n", unx.array[self.ldofs])
            self._un_version = unx.version
        return self._u_loc, self._un_loc

    def copy_solution_to_previous(self):
        """u_n <- u on the device and on the host copy (no upload at the next solve)."""
        p = self.problem
        if not self.spec.transient or getattr(self, "_u_loc", None) is None or p.u.x.version != self._u_version:
            return False
        if getattr(self, "_un_loc", None) is None:
            self._un_loc = self.ctx.to_device("u_n", p.u_n.x._array[self.ldofs])
        with self.torch.cuda.stream(self.stream):
            self._un_loc.copy_(self._u_loc)
        un = self._un_loc
        if self.lazy_gather:
            p.u_n.x.defer(lambda out: self._gather(un, out))
        else:
            p.u_n.x._array[:] = p.u.x._array
        self._un_version = p.u_n.x.version
        return True

    # ---- derived quantities (device reductions over the cells this rank owns + one all-reduce) ------
    def total_volume(self, species_index, vol_id):
        """int u_s dx over the volume (reference exports/total_volume.py:23-33): every cell is
        integrated by the rank that owns its first vertex."""
        torch = self.torch
        key = ("tv", vol_id)
        cache = self.__dict__.setdefault("_dq_cache", {})
        if key not in cache:
            lp, mesh = self.part, self.part.mesh
            gcells = self.problem.mesh.mesh.cells[lp.local_cells]
            mine = self.part_of[gcells[:, 0]] == self.rank
            tag = self.spec.cell_tag == self.spec.vol_ids.index(vol_id)
            sel = np.nonzero(mine & tag)[0]
            x = mesh.coords[mesh.cells[sel]]
            J = x[:, 1:, :] - x[:, :1, :]
            fact = {1: 1.0, 2: 2.0, 3: 6.0}[mesh.tdim]
            w = np.abs(np.linalg.det(J)) / fact / (mesh.tdim + 1)
            cache[key] = (torch.as_tensor(mesh.cells[sel].astype(np.int64), device=self.ctx.device),
                          torch.as_tensor(w, device=self.ctx.device))
        cells, w = cache[key]
        with torch.cuda.stream(self.stream):
            u = self._sync_fields()[0].view(-1, self.ns)[:, species_index]
            val = (u[cells].sum(dim=1) * w).sum().reshape(1)
            if self.nranks > 1:
                self.dist.all_reduce(val)
            out = float(val.cpu()[0])
        return out

    def measure_volume(self, vol_id):
        mesh = self.problem.mesh.mesh
        cells = self.problem.volume_meshtags.find(vol_id)
        x = mesh.coords[mesh.cells[cells]]
        J = x[:, 1:, :] - x[:, :1, :]
        fact = {1: 1.0, 2: 2.0, 3: 6.0}[mesh.tdim]
        return float(np.sum(np.abs(np.linalg.det(J))) / fact)

    def solve(self):
        ctx, p = self.ctx, self.problem
        with self.torch.cuda.stream(self.stream):
            self._push_state()
            u, un = self._sync_fields()
            t = float(p.t) if hasattr(p, "t") else 0.0
            atol = self.atol(t) if callable(self.atol) else self.atol
            rtol = self.rtol(t) if callable(self.rtol) else self.rtol
            it = self.newton_device(u, un, atol, rtol)
            self.u_dev = u
            self.stream.synchronize()
        if self.lazy_gather:
            # the global host field is assembled when somebody reads it: reading `u.x.array` is then
            # a collective (every rank must do it, as every rank runs the same script under mpirun)
            p.u.x.defer(lambda out: self._gather(u, out))
        else:
            self._gather(u)
        self._u_version = p.u.x.version
        return it

    def newton_device(self, u, un, atol, rtol):
        """Newton iteration on device-resident local vectors (no host copies of fields)."""
        ctx, lib, ptr = self.ctx, self.ctx.lib, self.ctx.ptr
        if self.native:
            # the library's Newton-GMRES, partitioned: halos and reductions cross NVLink inside its kernels
            n_its, reason, lin = C.c_int(0), C.c_int(0), C.c_int(0)
            fnorm = C.c_double(0.0)
            ctx._check(lib.fb_newton_solve(ctx.h, ptr(u), ptr(un), float(atol), float(rtol), self.stol, self.max_it,
                                           2, self.pc, self.ksp_rtol, self.ksp_max_it, C.byref(n_its),
                                           C.byref(reason), C.byref(fnorm), C.byref(lin)))
            self.solver.its, self.solver.reason = n_its.value, reason.value
            self.solver.linear_its, self.solver.fnorm = lin.value, fnorm.value
            return n_its.value
        it, snorm, f0, lin_total, reason = 0, 0.0, 1.0, 0, 0
        while True:
            self.halo(u)
            if un is not u:
                self.halo(un)
            ctx._check(lib.fb_assemble(ctx.h, ptr(u), ptr(un), ptr(self.F), ptr(self.J), 3 if it == 0 else 1))
            fn = self._norm(self.F)
            if it == 0:
                f0 = fn
            if not np.isfinite(fn):
                reason = -4
                break
            if f0 == 0.0 or fn < atol:
                reason = 2
                break
            if fn / f0 < rtol:
                reason = 3
                break
            if snorm < self.stol and it > 0:
                reason = 4
                break
            if it >= self.max_it:
                reason = -5
                break
            if it > 0:
                ctx._check(lib.fb_assemble(ctx.h, ptr(u), ptr(un), C.c_void_p(), ptr(self.J), 2))
            its, ok = self.linear_solve(self.J, self.F, self.y, self.ksp_rtol, self.ksp_max_it)
            lin_total += its
            if not ok:
                reason = -3
                break
            ctx._check(lib.fb_axpy(ctx.h, -1.0, ptr(self.y), ptr(u)))
            snorm = self._norm(self.y)
            it += 1
        self.solver.its, self.solver.reason, self.solver.linear_its, self.solver.fnorm = it, reason, lin_total, fn
        return it

    def _is_symmetric(self):
        spec = self.spec
        # as fb_set_spec decides for one GPU (csrc/api.cu): diffusion + mass only; reactions
        # (structured or point-wise), advection and solution-dependent coefficients are not
        return spec.struct is None and not spec.adv and not spec.diff_udep and all(
            not t.derivs for t in spec.vterms + spec.fterms)

    def _gather(self, u, out=None):
        """Owned values of every rank -> the global host vector (`out`, default the problem's u)
        on every rank: one padded NCCL all-gather, a device-side scatter into global dof order and
        one copy through a pinned buffer (no per-rank host indexing).  Collective."""
        torch, dist, p = self.torch, self.dist, self.problem
        if out is None:
            out = p.u.x._array
        n_own = self.part.n_owned * self.ns
        if self.nranks == 1:
            out[self.owned_gdofs] = u[:n_own].cpu().numpy()
            return
        if not hasattr(self, "_gather_perm"):
            counts = [None] * self.nranks
            dist.all_gather_object(counts, self.owned_gdofs)
            pad = max(len(c) for c in counts)
            n_glob = p.u.x._array.size
            perm = np.full(pad * self.nranks, n_glob, dtype=np.int64)   # padding -> a dummy slot
            for q, c in enumerate(counts):
                perm[q * pad: q * pad + len(c)] = c
            self._gather_pad = pad
            self._gather_perm = torch.as_tensor(perm, device=u.device)
            self._gather_send = torch.zeros(pad, dtype=torch.float64, device=u.device)
            self._gather_recv = torch.empty(pad * self.nranks, dtype=torch.float64, device=u.device)
            self._gather_glob = torch.empty(n_glob + 1, dtype=torch.float64, device=u.device)
            self._gather_host = torch.empty(n_glob + 1, dtype=torch.float64).pin_memory()
        with torch.cuda.stream(self.stream):
            self._gather_send[:n_own] = u[:n_own]
            dist.all_gather_into_tensor(self._gather_recv, self._gather_send)
            self._gather_glob.index_copy_(0, self._gather_perm, self._gather_recv)
            self._gather_host.copy_(self._gather_glob, non_blocking=True)
        self.stream.synchronize()
        out[:] = self._gather_host.numpy()[:-1]


class _ProblemView:
    """The problem with its mesh and tags replaced by the local partition."""

    def __init__(self, problem, local_mesh, vt, ft):
        self._problem = problem
        self.mesh = types.SimpleNamespace(mesh=local_mesh)
        self.volume_meshtags = vt
        self.facet_meshtags = ft

    def __getattr__(self, name):
        return getattr(self._problem, name)


class DistributedBackend:
    """``lazy_gather=True`` leaves the fields on the devices after a solve and assembles the global
    host array only when ``u.x.array`` is read - which every rank must then do together.  The
    default gathers after every solve, so that any rank may look at the field on its own."""

    def __init__(self, device=0, lazy_gather=False):
        self.device = device
        self.lazy_gather = lazy_gather

    def create_newton_solver(self, problem, petsc_options):
        return DistributedNewtonSolver(problem, petsc_options, self.device, self.lazy_gather)
