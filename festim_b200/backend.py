"""ctypes binding of ``libfestim_b200.so`` and the solver object that sits at the
``create_solver`` seam (reference ``src/festim/problem.py:170-193``).

``B200NewtonSolver`` exposes exactly what the reference time loop uses from
``dolfinx.fem.petsc.NonlinearProblem``: ``solve()``, ``solver
.getConvergedReason()``, ``solver.getIterationNumber()``, writable ``rtol`` /
``atol``.  PyTorch provides device buffers and the stream only.  There is no
CPU fallback: a missing library or GPU raises.
"""

from __future__ import annotations

import ctypes as C
import os
import warnings

import numpy as np

from festim_b200 import fem
from festim_b200.lowering import KernelSpec

_LIB = None
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libfestim_b200.so")

FB_ASSEMBLE_F, FB_ASSEMBLE_J = 1, 2
KSP = {"auto": 0, "preonly": 0, "cg": 1, "gmres": 2, "fgmres": 2, "tridiag": 3}
# "lu" (the reference default, an exact solve) maps to the library's own choice of preconditioner
PC = {"jacobi": 0, "bjacobi": 0, "pbjacobi": 0, "chebyshev": 1, "lu": 2, "auto": 2}

_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_f64p = C.POINTER(C.c_double)


def load_library():
    """Load the CUDA library; raises if it has not been built (no fallback)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `make` or `python -c 'import "
            "__graft_entry__ as g; g.build()'`. festim_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    vp, i, i64, d = C.c_void_p, C.c_int, C.c_int64, C.c_double
    sig = {
        "fb_create": (i, [i, vp, C.POINTER(vp)]),
        "fb_destroy": (None, [vp]),
        "fb_last_error": (C.c_char_p, [vp]),
        "fb_launch_count": (i64, [vp]),
        "fb_set_mesh": (i, [vp, i, i64, i64, _f64p, _i32p, _i32p, i64, _i32p, _i32p, _i32p]),
        "fb_set_pattern": (i, [vp, _i32p, _i32p, i64, _i32p, _i32p]),
        "fb_set_spec": (i, [vp, vp, C.c_size_t]),
        "fb_jit_load": (i, [vp, vp, C.c_size_t]),
        "fb_set_dirichlet": (i, [vp, i64, _i64p]),
        "fb_num_dofs": (i64, [vp]),
        "fb_jacobian_nnz": (i64, [vp]),
        "fb_set_scalars": (i, [vp, i, _f64p]),
        "fb_set_temperature": (i, [vp, d, vp]),
        "fb_set_field": (i, [vp, i, vp]),
        "fb_set_velocity": (i, [vp, i, vp]),
        "fb_set_dirichlet_values": (i, [vp, vp]),
        "fb_update_load": (i, [vp]),
        "fb_assemble": (i, [vp, vp, vp, vp, vp, i]),
        "fb_spmv": (i, [vp, vp, vp, vp]),
        "fb_dot": (i, [vp, vp, vp, _f64p]),
        "fb_axpy": (i, [vp, d, vp, vp]),
        "fb_linear_solve": (i, [vp, vp, vp, vp, i, i, d, d, i, C.POINTER(i), _f64p]),
        "fb_newton_solve": (i, [vp, vp, vp, d, d, d, i, i, i, d, i, C.POINTER(i), C.POINTER(i),
                                _f64p, C.POINTER(i)]),
        "fb_run_steps": (i, [vp, vp, vp, _f64p, _f64p, d, i, d, d, d, i, i, i, d, i, i, d, d, i, d, i, _f64p, d,
                             C.POINTER(i), _f64p, C.POINTER(i), C.POINTER(i), C.POINTER(i)]),
        "fb_last_timings": (i, [vp, _f64p]),
        "fb_stepsize_modify": (d, [d, i, d, i, d, d, i, d, i, _f64p, d]),
        "fb_hessenberg_eigenvalues": (i, [i, _f64p, _f64p, _f64p]),
        "fb_chebyshev_info": (i, [vp, _f64p]),
        "fb_chebyshev_step": (i, [vp, vp, vp, vp, vp, d, d]),
        "fb_chebyshev_steps": (i, [d, d, d, C.POINTER(i)]),
        "fb_total_volume": (i, [vp, vp, i, i, _f64p]),
        "fb_total_surface": (i, [vp, vp, i, i, _f64p]),
        "fb_surface_flux": (i, [vp, vp, i, i, d, vp, _f64p]),
        "fb_integrate": (i, [vp, vp, i, i, i, i, i, _f64p]),
        "fb_set_owned": (i, [vp, i64]),
        "fb_set_sm_partition": (i, [vp, i, _i32p, _i32p, _i32p, C.POINTER(C.c_uint16), i64]),
        "fb_bjacobi_setup": (i, [vp, vp]),
        "fb_precond": (i, [vp, vp, vp]),
        "fb_dcg_init": (i, [vp, vp, vp, vp, vp, vp, vp]),
        "fb_peer_alloc": (i, [vp, i, i, vp]),
        "fb_peer_open": (i, [vp, i, vp]),
        "fb_peer_set_send": (i, [vp, i64, vp, vp, vp]),
        "fb_dcg_persistent": (i, [vp, vp, vp, vp, C.c_double, C.c_double, i, vp, vp]),
        "fb_dcg_spmv_dots": (i, [vp, vp, vp, vp, vp, vp]),
        "fb_dcg_update": (i, [vp, d, d, vp, vp, vp, vp, vp, vp]),
        "fb_dcg_scalars": (i, [vp, vp, vp, i]),
        "fb_dcg_update_dev": (i, [vp, vp, vp, vp, vp, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


EXPORTED_SYMBOLS = [
    "fb_create", "fb_destroy", "fb_last_error", "fb_launch_count", "fb_set_mesh", "fb_set_pattern",
    "fb_set_spec", "fb_jit_load", "fb_set_dirichlet", "fb_num_dofs", "fb_jacobian_nnz", "fb_set_scalars",
    "fb_set_temperature", "fb_set_field", "fb_set_velocity", "fb_set_dirichlet_values",
    "fb_update_load", "fb_assemble", "fb_spmv", "fb_dot", "fb_axpy", "fb_linear_solve",
    "fb_newton_solve", "fb_run_steps", "fb_last_timings", "fb_stepsize_modify", "fb_hessenberg_eigenvalues", "fb_chebyshev_info", "fb_chebyshev_step", "fb_chebyshev_steps", "fb_total_volume",
    "fb_total_surface", "fb_surface_flux", "fb_integrate", "fb_set_owned", "fb_set_sm_partition",
    "fb_bjacobi_setup", "fb_precond",
    "fb_peer_alloc", "fb_peer_open", "fb_peer_set_send", "fb_dcg_persistent",
    "fb_dcg_init", "fb_dcg_spmv_dots", "fb_dcg_update", "fb_dcg_scalars", "fb_dcg_update_dev",
]


def _np_ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


class Context:
    """Owns an ``fb_ctx`` and the torch device buffers handed to it."""

    def __init__(self, spec: KernelSpec, device=0, renumber=True):
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("festim_b200 needs a CUDA device (no CPU fallback)")
        self.torch = torch
        self.lib = load_library()
        self.spec = spec
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        self.stream = torch.cuda.current_stream(self.device)
        self.h = C.c_void_p()
        rc = self.lib.fb_create(device, C.c_void_p(self.stream.cuda_stream), C.byref(self.h))
        if rc:
            raise RuntimeError(f"fb_create failed with status {rc}")
        self._keep = {}
        mesh = spec.mesh
        self.nv = mesh.num_vertices
        self.sm_count = torch.cuda.get_device_properties(self.device).multi_processor_count
        # Vertex renumbering by recursive coordinate bisection into one compact part
        # per SM: contiguous vertex ranges are then compact subdomains, which is what
        # the persistent CG stages in shared memory (operator rows + a small halo of
        # the iterate).  dolfinx reorders dofs too; the user-facing arrays keep the
        # mesh numbering (perm is applied on every upload / download).
        self.perm = self.inv = None
        self._perm_dev = None
        self.block_ranges = None
        if renumber and mesh.tdim > 1 and mesh.num_vertices >= 4 * self.sm_count:
            from festim_b200.distributed import partition_rcb

            part = partition_rcb(mesh.coords, self.sm_count)
            self.perm = np.argsort(part, kind="stable").astype(np.int64)   # new -> old
            self.inv = np.empty_like(self.perm)
            self.inv[self.perm] = np.arange(len(self.perm))
            counts = np.bincount(part, minlength=self.sm_count)
            self.block_ranges = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
            from festim_b200.mesh import SimplexMesh

            dmesh = SimplexMesh(mesh.coords[self.perm], self.inv[mesh.cells])
            dmesh._bfacets = self.inv[mesh.bfacets].astype(np.int32)
            dmesh._bfacet_cell = mesh.bfacet_cell
        else:
            dmesh = mesh
        self.dmesh = dmesh
        coords = np.ascontiguousarray(dmesh.coords, dtype=np.float64)
        cells = np.ascontiguousarray(dmesh.cells, dtype=np.int32)
        bf = np.ascontiguousarray(dmesh.bfacets, dtype=np.int32)
        bfc = np.ascontiguousarray(dmesh.bfacet_cell, dtype=np.int32)
        self._check(self.lib.fb_set_mesh(
            self.h, mesh.tdim, mesh.num_vertices, mesh.num_cells, _np_ptr(coords, C.c_double),
            _np_ptr(cells, C.c_int32), _np_ptr(spec.cell_tag, C.c_int32), bf.shape[0],
            _np_ptr(bf, C.c_int32), _np_ptr(bfc, C.c_int32), _np_ptr(spec.bfacet_tag, C.c_int32)))
        rowptr, colidx = dmesh.graph
        v2c_ptr, v2c_idx = dmesh.vertex_to_cell
        self._check(self.lib.fb_set_pattern(
            self.h, _np_ptr(rowptr, C.c_int32), _np_ptr(colidx, C.c_int32), len(colidx),
            _np_ptr(v2c_ptr, C.c_int32), _np_ptr(v2c_idx, C.c_int32)))
        blob = spec.blob()
        self._check(self.lib.fb_set_spec(self.h, blob, len(blob)))
        self.jit = False
        from festim_b200 import jit

        if jit.needed(spec):
            cubin = jit.compile_cubin(jit.generate_source(spec))
            self._check(self.lib.fb_jit_load(self.h, cubin, len(cubin)))
            self.jit = True
        self.n_dofs = int(self.lib.fb_num_dofs(self.h))
        self.nnz = int(self.lib.fb_jacobian_nnz(self.h))
        if self.block_ranges is not None:
            self._set_sm_partition(rowptr, colidx)

    def _set_sm_partition(self, rowptr, colidx):
        """Per-SM block: halo vertex list and 16-bit local column of every block entry."""
        nb = self.sm_count
        rng = self.block_ranges
        halo_ptr = np.zeros(nb + 1, dtype=np.int32)
        halos, lcols = [], []
        for b in range(nb):
            a, e = int(rng[b]), int(rng[b + 1])
            cols = colidx[rowptr[a]:rowptr[e]].astype(np.int64)
            inside = (cols >= a) & (cols < e)
            halo = np.unique(cols[~inside])
            if (e - a) + len(halo) > 65535:
                return  # local index does not fit 16 bits: keep the streaming path
            lc = np.where(inside, cols - a, 0)
            lc[~inside] = (e - a) + np.searchsorted(halo, cols[~inside])
            halos.append(halo.astype(np.int32))
            lcols.append(lc.astype(np.uint16))
            halo_ptr[b + 1] = halo_ptr[b] + len(halo)
        halo_idx = np.ascontiguousarray(np.concatenate(halos)) if halos else np.zeros(0, np.int32)
        lcol = np.concatenate(lcols)
        # partitioned meshes: the blocks cover the owned vertices; ghost rows are never swept
        lcol = np.ascontiguousarray(np.concatenate([lcol, np.zeros(len(colidx) - len(lcol), np.uint16)]))
        self._check(self.lib.fb_set_sm_partition(
            self.h, nb, _np_ptr(rng, C.c_int32), _np_ptr(halo_ptr, C.c_int32),
            _np_ptr(halo_idx, C.c_int32), _np_ptr(lcol, C.c_uint16), len(lcol)))

    # -- numbering ----------------------------------------------------------------
    def nodal_to_device(self, name, host_array, width=1):
        """Upload a per-vertex array (``width`` values per vertex) in device numbering;
        the renumbering is a device-side gather (the host copy stays a plain memcpy)."""
        if self.perm is None:
            return self.to_device(name, host_array)
        raw = self.to_device(name + "/user", host_array)
        if self._perm_dev is None:
            self._perm_dev = self.torch.as_tensor(self.perm, dtype=self.torch.long, device=self.device)
        slot = self._keep.get(name)
        if slot is None or slot[0].numel() != raw.numel():
            slot = self._keep[name] = (self.torch.empty_like(raw), None)
        self.torch.index_select(raw.view(self.nv, width), 0, self._perm_dev, out=slot[0].view(self.nv, width))
        return slot[0]

    def nodal_from_device(self, tensor, width=1, out=None):
        """Device vector -> host array in the user's numbering (scatter on the device, then
        one copy through a pinned staging buffer into ``out``, or into a new array)."""
        torch = self.torch
        if self.perm is not None:
            if self._perm_dev is None:
                self._perm_dev = torch.as_tensor(self.perm, dtype=torch.long, device=self.device)
            user = torch.empty_like(tensor)
            user.view(self.nv, width).index_copy_(0, self._perm_dev, tensor.view(self.nv, width))
            tensor = user
        key = ("download", tensor.numel())
        pinned = self._keep.get(key)
        if pinned is None:
            pinned = self._keep[key] = torch.empty(tensor.numel(), dtype=torch.float64).pin_memory()
        pinned.copy_(tensor, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        if out is None:
            return pinned.numpy().copy()
        out[:] = pinned.numpy()
        return out

    def dofs_to_device_numbering(self, dofs, ns):
        dofs = np.asarray(dofs, dtype=np.int64)
        if self.inv is None:
            return dofs
        return self.inv[dofs // ns] * ns + dofs % ns

    def dofs_to_user_numbering(self, dofs, ns):
        dofs = np.asarray(dofs, dtype=np.int64)
        if self.perm is None:
            return dofs
        return self.perm[dofs // ns] * ns + dofs % ns

    def _check(self, rc):
        if rc < 0:
            msg = self.lib.fb_last_error(self.h)
            raise RuntimeError(f"libfestim_b200: {msg.decode() if msg else rc}")
        return rc

    def close(self):
        if self.h:
            self.lib.fb_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- buffers ---------------------------------------------------------------
    def zeros(self, n):
        return self.torch.zeros(n, dtype=self.torch.float64, device=self.device)

    def to_device(self, name, host_array):
        """Copy a host array into a persistent device tensor (pinned staging)."""
        torch = self.torch
        arr = np.ascontiguousarray(host_array, dtype=np.float64).ravel()
        slot = self._keep.get(name)
        if slot is None or slot[0].numel() != arr.size:
            pinned = torch.empty(arr.size, dtype=torch.float64).pin_memory()
            dev = torch.empty(arr.size, dtype=torch.float64, device=self.device)
            slot = self._keep[name] = (dev, pinned)
        dev, pinned = slot
        pinned.numpy()[:] = arr
        dev.copy_(pinned, non_blocking=True)
        return dev

    @staticmethod
    def ptr(t):
        return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p()

    @property
    def launches(self):
        return int(self.lib.fb_launch_count(self.h))


def jacobian_index(spec, graph):
    """(row dof, column dof) of every stored Jacobian scalar in storage order: block CSR of
    the vertex graph, the ``npairs`` stored species pairs of block ``e`` at ``e * npairs + p``
    (``p`` = position of (row species, column species) in ``pair_row`` / ``pair_cs``)."""
    rowptr, colidx = graph
    ns, npairs = spec.ns, spec.npairs
    deg = np.diff(rowptr).astype(np.int64)
    v_of = np.repeat(np.arange(len(deg), dtype=np.int64), deg)          # row vertex of every block
    pair_row = np.repeat(np.arange(ns, dtype=np.int64), spec.pair_nc)
    pair_cs = spec.pair_cs.astype(np.int64)
    R = (v_of[:, None] * ns + pair_row[None, :]).ravel()
    Cc = (colidx.astype(np.int64)[:, None] * ns + pair_cs[None, :]).ravel()
    return R, Cc


class _Snes:
    def __init__(self):
        self.reason = 0
        self.its = 0
        self.linear_its = 0
        self.fnorm = 0.0

    def getConvergedReason(self):
        return self.reason

    def getIterationNumber(self):
        return self.its

    def getLinearSolveIterations(self):
        return self.linear_its

    def getFunctionNorm(self):
        return self.fnorm


class B200NewtonSolver:
    def __init__(self, problem, petsc_options, device=0):
        self.problem = problem
        self.spec = KernelSpec(problem)
        self.ctx = Context(self.spec, device)
        self.atol = petsc_options["snes_atol"]
        self.rtol = petsc_options["snes_rtol"]
        self.stol = float(petsc_options["snes_stol"])
        self.max_it = int(petsc_options["snes_max_it"])
        ksp = str(petsc_options.get("ksp_type", "preonly")).lower()
        pc = str(petsc_options.get("pc_type", "lu")).lower()
        known = {"snes_type", "snes_linesearch_type", "snes_stol", "snes_divergence_tolerance",
                 "ksp_type", "pc_type", "pc_factor_mat_solver_type", "snes_error_if_not_converged",
                 "ksp_error_if_not_converged", "snes_atol", "snes_rtol", "snes_max_it",
                 "ksp_rtol", "ksp_max_it", "ksp_atol"}
        for key in petsc_options:
            if key not in known:
                warnings.warn(f"PETSc option {key} is ignored by the B200 backend")
        if ksp not in KSP:
            warnings.warn(f"ksp_type {ksp} is not implemented, using gmres")
            ksp = "gmres"
        self.ksp = KSP[ksp]
        if pc not in PC:
            warnings.warn(f"pc_type {pc} is not implemented, using the default preconditioner")
        self.pc = PC.get(pc, 2)
        # the reference's default is an exact LU solve: use a Krylov tolerance far
        # below the Newton tolerance so Newton iteration counts are unchanged
        snes_rtol = self.rtol if isinstance(self.rtol, (int, float)) else 1e-10
        exact_like = min(1e-12, max(1e-2 * float(snes_rtol), 1e-15))
        self.ksp_rtol = float(petsc_options.get("ksp_rtol", exact_like if ksp == "preonly" else 1e-10))
        self.ksp_max_it = int(petsc_options.get("ksp_max_it", 20000))
        self.solver = _Snes()
        self.timings = np.zeros(4)
        self._D_T = None
        self._D_T_dev = None

        ctx, spec = self.ctx, self.spec
        # Dirichlet dofs (fixed), later BCs win on shared dofs
        self._bc_forms = list(problem.bc_forms)
        dofs = np.concatenate([bc.dofs for bc in self._bc_forms]) if self._bc_forms else \
            np.zeros(0, dtype=np.int64)
        uniq, first_rev = np.unique(dofs[::-1], return_index=True)
        self._bc_pick = len(dofs) - 1 - first_rev
        self._bc_dofs = np.ascontiguousarray(ctx.dofs_to_device_numbering(uniq, spec.ns), dtype=np.int64)
        ctx._check(ctx.lib.fb_set_dirichlet(ctx.h, len(self._bc_dofs),
                                            _np_ptr(self._bc_dofs, C.c_int64)))
        self.u_dev = None
        self.un_dev = None
        self._u_version = self._un_version = -1   # host versions the device copies correspond to
        self.device_resident = True               # fields stay on the device between solves
        self._signature = None

    # -- per-step data -----------------------------------------------------------
    def _push_state(self):
        """Upload everything the residual depends on besides u (reference
        ``update_time_dependent_values``) and refresh the cached coefficients when
        any of it changed."""
        ctx, spec, p = self.ctx, self.spec, self.problem
        lib = ctx.lib
        scal = spec.scalar_values()
        T = spec.temperature
        Tdev = None
        # change detection on everything the cached coefficients depend on: a memcmp against
        # the last uploaded copy (users mutate ``x.array`` in place, so identity is no proof)
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
            Tdev = ctx.nodal_to_device("T", T.x.array)
            ctx._check(lib.fb_set_temperature(ctx.h, 0.0, ctx.ptr(Tdev)))
        elif T is not None:
            ctx._check(lib.fb_set_temperature(ctx.h, float(T.value), C.c_void_p()))
        for k, f in enumerate(spec.fields):
            ctx._check(lib.fb_set_field(ctx.h, k, ctx.ptr(ctx.nodal_to_device(f"field{k}", f.x.array))))
        for k, (_, _, vel) in enumerate(spec.adv):
            ctx._check(lib.fb_set_velocity(ctx.h, k, ctx.ptr(ctx.nodal_to_device(f"vel{k}", vel.array, spec.tdim))))
        ctx._check(lib.fb_update_load(ctx.h))

    def _sync_fields(self):
        """Device copies of u and u_n: uploaded only when the host arrays were touched since the
        last solve (users set initial conditions or edit ``u.x.array`` in place)."""
        ctx, p, ns = self.ctx, self.problem, self.spec.ns
        ux, unx = p.u.x, p.u_n.x
        if self.u_dev is None or ux.version != self._u_version:
            self.u_dev = ctx.nodal_to_device("u", ux.array, ns)
            self._u_version = ux.version
        if not self.spec.transient:
            return self.u_dev, self.u_dev
        if self.un_dev is None or unx.version != self._un_version:
            self.un_dev = ctx.nodal_to_device("u_n", unx.array, ns)
            self._un_version = unx.version
        return self.u_dev, self.un_dev

    def copy_solution_to_previous(self):
        """u_n <- u on the device (reference problem.py:248); False when the host copy of u is the
        current one, so that the caller does the plain array copy."""
        p = self.problem
        if not self.spec.transient or self.u_dev is None or p.u.x.version != self._u_version:
            return False
        if self.un_dev is None:
            self.un_dev = self.ctx.nodal_to_device("u_n", p.u_n.x._array, self.spec.ns)
        self.un_dev.copy_(self.u_dev)
        un_dev, ns, ctx = self.un_dev, self.spec.ns, self.ctx
        p.u_n.x.defer(lambda out: ctx.nodal_from_device(un_dev, ns, out=out))
        self._un_version = p.u_n.x.version
        return True

    def solve(self):
        ctx, p = self.ctx, self.problem
        self._push_state()
        ns = self.spec.ns
        u_dev, un_dev = self._sync_fields()
        t = float(p.t) if hasattr(p, "t") else 0.0
        atol = self.atol(t) if callable(self.atol) else self.atol
        rtol = self.rtol(t) if callable(self.rtol) else self.rtol
        n_its, reason, lin = C.c_int(0), C.c_int(0), C.c_int(0)
        fnorm = C.c_double(0.0)
        rc = ctx.lib.fb_newton_solve(
            ctx.h, ctx.ptr(u_dev), ctx.ptr(un_dev), float(atol), float(rtol), self.stol,
            self.max_it, self.ksp, self.pc, self.ksp_rtol, self.ksp_max_it,
            C.byref(n_its), C.byref(reason), C.byref(fnorm), C.byref(lin))
        ctx._check(rc)
        # the solution stays on the device; the host array is refreshed when somebody reads it
        p.u.x.defer(lambda out: ctx.nodal_from_device(u_dev, ns, out=out))
        self._u_version = p.u.x.version
        self.solver.its = n_its.value
        self.solver.reason = reason.value
        self.solver.linear_its = lin.value
        self.solver.fnorm = fnorm.value
        tm = (C.c_double * 4)()
        ctx.lib.fb_last_timings(ctx.h, tm)
        self.timings = np.array(list(tm))
        return n_its.value

    def run_steps(self, t_stop, max_steps=4096):
        """Device-resident time stepping until ``t_stop`` (fb_run_steps): the implicit-Euler loop of
        reference problem.py:195-255 with the step-size rule of stepsize.py:144-167 inside the library.
        Only for problems where nothing but t and dt changes between steps (the caller checks).
        Returns (times at the start of every step, Newton iterations per step)."""
        ctx, p = self.ctx, self.problem
        self._push_state()
        u_dev, un_dev = self._sync_fields()
        st = p.settings.stepsize
        ms = np.ascontiguousarray(sorted(st.milestones or []), dtype=np.float64)
        t, dt = C.c_double(float(p.t)), C.c_double(float(p.dt.value))
        n, lin, reason = C.c_int(0), C.c_int(0), C.c_int(0)
        times = np.zeros(max_steps)
        its = np.zeros(max_steps, dtype=np.int32)
        mx = st.max_stepsize if isinstance(st.max_stepsize, (int, float)) and st.max_stepsize else 0.0
        rc = ctx.lib.fb_run_steps(
            ctx.h, ctx.ptr(u_dev), ctx.ptr(un_dev), C.byref(t), C.byref(dt), float(t_stop), int(max_steps),
            float(self.atol), float(self.rtol), self.stol, self.max_it, self.ksp, self.pc, self.ksp_rtol,
            self.ksp_max_it, int(bool(st.adaptive)), float(st.growth_factor or 1.0), float(st.cutback_factor or 1.0),
            int(st.target_nb_iterations or 0), float(mx), len(ms), _np_ptr(ms, C.c_double) if len(ms) else None,
            float(st.milestone_tolerance),
            C.byref(n), _np_ptr(times, C.c_double), its.ctypes.data_as(C.POINTER(C.c_int)), C.byref(lin), C.byref(reason))
        ctx._check(rc)
        ns = self.spec.ns
        p.t.value = t.value
        p.dt.value = dt.value
        p.u.x.defer(lambda out: ctx.nodal_from_device(u_dev, ns, out=out))
        p.u_n.x.defer(lambda out: ctx.nodal_from_device(un_dev, ns, out=out))
        self._u_version, self._un_version = p.u.x.version, p.u_n.x.version
        self.library_steps = getattr(self, "library_steps", 0) + n.value
        self.solver.its = int(its[n.value - 1]) if n.value else 0
        self.solver.reason = reason.value
        self.solver.linear_its = lin.value
        return times[: n.value].tolist(), its[: n.value].tolist()

    # -- low-level access used by the parity tests and the benchmark -----------------
    def jacobian_index(self):
        """(rows, cols) of every stored Jacobian scalar, in storage order (the
        layout documented in include/festim_b200.h)."""
        R, Cc = jacobian_index(self.spec, self.ctx.dmesh.graph)
        ns = self.spec.ns
        return self.ctx.dofs_to_user_numbering(R, ns), self.ctx.dofs_to_user_numbering(Cc, ns)

    def assemble_host(self, want_J=True):
        """Assemble F (and J as scipy CSR) at the problem's current u, u_n."""
        import scipy.sparse as sp

        ctx, p = self.ctx, self.problem
        self._push_state()
        ns = self.spec.ns
        u_dev = ctx.nodal_to_device("u", p.u.x.array, ns)
        un_dev = ctx.nodal_to_device("u_n", p.u_n.x.array, ns) if self.spec.transient else u_dev
        F = ctx.zeros(ctx.n_dofs)
        J = ctx.zeros(ctx.nnz) if want_J else None
        flags = FB_ASSEMBLE_F | (FB_ASSEMBLE_J if want_J else 0)
        ctx._check(ctx.lib.fb_assemble(ctx.h, ctx.ptr(u_dev), ctx.ptr(un_dev), ctx.ptr(F),
                                       ctx.ptr(J), flags))
        Fh = ctx.nodal_from_device(F, ns)
        if not want_J:
            return Fh, None
        R, Cc = self.jacobian_index()
        Jh = sp.csr_matrix((J.cpu().numpy(), (R, Cc)), shape=(ctx.n_dofs, ctx.n_dofs))
        return Fh, Jh

    # -- derived quantities (device reductions) -------------------------------------
    def set_D_global_temperature(self, T):
        self._D_T = T
        self._D_T_dev = None

    def _current_u(self):
        return self._sync_fields()[0]

    def total_volume(self, species_index, vol_id):
        out = C.c_double(0.0)
        ctx = self.ctx
        ctx._check(ctx.lib.fb_total_volume(ctx.h, ctx.ptr(self._current_u()), species_index,
                                           self.spec.vol_ids.index(vol_id), C.byref(out)))
        return out.value

    def integrate_custom(self, export):
        """Integral of a CustomQuantity's expression over its subdomain (device kernel)."""
        kind, tag, prog, off, nq, _ = self.spec.custom[id(export)]
        self._push_state()   # fields the expression reads (e.g. the diffusivities) may have changed
        out = C.c_double(0.0)
        ctx = self.ctx
        ctx._check(ctx.lib.fb_integrate(ctx.h, ctx.ptr(self._current_u()), kind, tag, prog, off, nq, C.byref(out)))
        return out.value

    def total_surface(self, species_index, surf_id):
        out = C.c_double(0.0)
        ctx = self.ctx
        ctx._check(ctx.lib.fb_total_surface(ctx.h, ctx.ptr(self._current_u()), species_index,
                                            self.spec.surf_ids.index(surf_id), C.byref(out)))
        return out.value

    def surface_flux(self, species_index, surf_id):
        out = C.c_double(0.0)
        ctx = self.ctx
        spec = self.spec
        if any(int(spec.diff_slot[tag, species_index]) >= len(spec.diff_energies) for tag in range(len(spec.vol_ids))):
            # the reference weights the gradient with material.D itself (define_D_global,
            # hydrogen_transport_problem.py:497-502); the device kernel only knows Arrhenius laws
            raise NotImplementedError("SurfaceFlux with a user-defined Material.D is not implemented on the "
                                      "B200 backend (use a CustomQuantity with D_<species>)")
        Tn, Tc = C.c_void_p(), 0.0
        if isinstance(self._D_T, np.ndarray):
            if self._D_T_dev is None:
                self._D_T_dev = ctx.nodal_to_device("D_T", self._D_T)
            Tn = ctx.ptr(self._D_T_dev)
        else:
            Tc = float(self._D_T)
        ctx._check(ctx.lib.fb_surface_flux(ctx.h, ctx.ptr(self._current_u()), species_index,
                                           self.spec.surf_ids.index(surf_id), Tc, Tn, C.byref(out)))
        return out.value

    def measure_volume(self, vol_id):
        mesh = self.spec.mesh
        cells = self.problem.volume_meshtags.find(vol_id)
        x = mesh.coords[mesh.cells[cells]]
        J = x[:, 1:, :] - x[:, :1, :]
        fact = {1: 1.0, 2: 2.0, 3: 6.0}[mesh.tdim]
        return float(np.sum(np.abs(np.linalg.det(J))) / fact)

    def measure_surface(self, surf_id):
        mesh = self.spec.mesh
        fv = mesh.bfacets[self.problem.facet_meshtags.find(surf_id)]
        x = mesh.coords[fv]
        if mesh.tdim == 1:
            return float(len(fv))
        if mesh.tdim == 2:
            return float(np.linalg.norm(x[:, 1] - x[:, 0], axis=1).sum())
        return float(0.5 * np.linalg.norm(np.cross(x[:, 1] - x[:, 0], x[:, 2] - x[:, 0]), axis=1).sum())


class B200Backend:
    def __init__(self, device=0):
        self.device = device

    def create_newton_solver(self, problem, petsc_options):
        solver = B200NewtonSolver(problem, petsc_options, self.device)
        if getattr(problem, "_D_global_T", None) is not None:
            solver.set_D_global_temperature(problem._D_global_T)
        return solver
