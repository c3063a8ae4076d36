"""Drives the CPU emulation build of the library (tests/emu) through the product's C ABI with
numpy arrays standing in for device buffers: the same call sequence as
festim_b200.backend.Context / B200NewtonSolver, without torch or a GPU.  Test infrastructure."""
import ctypes as C
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emu"))

from festim_b200 import fem
from festim_b200.backend import jacobian_index
from festim_b200.lowering import KernelSpec

_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        import build_emu

        path = build_emu.build()
        L = C.CDLL(path)
        for name in ("fb_create", "fb_set_mesh", "fb_set_pattern", "fb_set_spec", "fb_set_dirichlet",
                     "fb_set_scalars", "fb_set_temperature", "fb_set_field", "fb_set_velocity",
                     "fb_set_dirichlet_values", "fb_update_load", "fb_assemble", "fb_spmv"):
            getattr(L, name).restype = C.c_int
        L.fb_last_error.restype = C.c_char_p
        L.fb_num_dofs.restype = C.c_int64
        L.fb_jacobian_nnz.restype = C.c_int64
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class EmuSolver:
    """Assembly and SpMV of one problem on the emulated device."""

    def __init__(self, problem):
        self.problem = problem
        self.spec = spec = KernelSpec(problem)
        self.lib = L = lib()
        self.h = C.c_void_p()
        assert L.fb_create(0, None, C.byref(self.h)) == 0
        mesh = spec.mesh
        self.keep = []

        def arr(a, dt):
            a = np.ascontiguousarray(a, dtype=dt)
            self.keep.append(a)
            return a

        coords, cells = arr(mesh.coords, np.float64), arr(mesh.cells, np.int32)
        bf, bfc = arr(mesh.bfacets, np.int32), arr(mesh.bfacet_cell, np.int32)
        ct, bt = arr(spec.cell_tag, np.int32), arr(spec.bfacet_tag, np.int32)
        self.check(L.fb_set_mesh(self.h, mesh.tdim, C.c_int64(mesh.num_vertices), C.c_int64(mesh.num_cells),
                                 _p(coords), _p(cells), _p(ct), C.c_int64(bf.shape[0]), _p(bf), _p(bfc), _p(bt)))
        rowptr, colidx = mesh.graph
        v2c_ptr, v2c_idx = mesh.vertex_to_cell
        rowptr, colidx = arr(rowptr, np.int32), arr(colidx, np.int32)
        self.graph = (rowptr, colidx)
        self.check(L.fb_set_pattern(self.h, _p(rowptr), _p(colidx), C.c_int64(len(colidx)),
                                    _p(arr(v2c_ptr, np.int32)), _p(arr(v2c_idx, np.int32))))
        blob = spec.blob()
        self.check(L.fb_set_spec(self.h, blob, C.c_size_t(len(blob))))
        self.n_dofs = int(L.fb_num_dofs(self.h))
        self.nnz = int(L.fb_jacobian_nnz(self.h))
        self.bc_forms = list(problem.bc_forms)
        dofs = np.concatenate([bc.dofs for bc in self.bc_forms]) if self.bc_forms else np.zeros(0, np.int64)
        uniq, first_rev = np.unique(dofs[::-1], return_index=True)
        self.bc_pick = len(dofs) - 1 - first_rev
        self.bc_dofs = arr(uniq, np.int64)
        self.check(L.fb_set_dirichlet(self.h, C.c_int64(len(self.bc_dofs)), _p(self.bc_dofs)))

    def check(self, rc):
        if rc < 0:
            raise RuntimeError(self.lib.fb_last_error(self.h).decode())
        return rc

    def push_state(self):
        L, spec = self.lib, self.spec
        self.state = []

        def arr(a):
            a = np.ascontiguousarray(a, dtype=np.float64)
            self.state.append(a)
            return a

        scal = arr(spec.scalar_values())
        self.check(L.fb_set_scalars(self.h, len(scal), _p(scal)))
        T = spec.temperature
        if isinstance(T, fem.Function):
            self.check(L.fb_set_temperature(self.h, C.c_double(0.0), _p(arr(T.x.array))))
        elif T is not None:
            self.check(L.fb_set_temperature(self.h, C.c_double(float(T.value)), None))
        for k, f in enumerate(spec.fields):
            self.check(L.fb_set_field(self.h, k, _p(arr(f.x.array))))
        for k, (_, _, vel) in enumerate(spec.adv):
            self.check(L.fb_set_velocity(self.h, k, _p(arr(vel.array))))
        bc_vals = np.concatenate([bc.values() for bc in self.bc_forms])[self.bc_pick] \
            if self.bc_forms else np.zeros(0)
        self.bc_vals = arr(bc_vals)
        self.check(L.fb_set_dirichlet_values(self.h, _p(self.bc_vals) if len(bc_vals) else None))
        self.check(L.fb_update_load(self.h))

    def assemble(self, u, un, flags=3):
        """F and J (scipy CSR) with the strong Dirichlet conditions applied, as fb_assemble returns them."""
        self.push_state()
        u = np.ascontiguousarray(u, dtype=np.float64)
        un = np.ascontiguousarray(un, dtype=np.float64)
        F = np.full(self.n_dofs, np.nan)
        J = np.full(self.nnz, np.nan)
        self.check(self.lib.fb_assemble(self.h, _p(u), _p(un), _p(F) if flags & 1 else None,
                                        _p(J) if flags & 2 else None, flags))
        self.Jvals = J
        if not flags & 2:
            return F, None
        R, Cc = jacobian_index(self.spec, self.graph)
        return F, sp.csr_matrix((J, (R, Cc)), shape=(self.n_dofs, self.n_dofs))

    def spmv(self, Jvals, x):
        y = np.full(self.n_dofs, np.nan)
        x = np.ascontiguousarray(x, dtype=np.float64)
        self.check(self.lib.fb_spmv(self.h, _p(Jvals) if Jvals is not None else None, _p(x), _p(y)))
        return y


def _linear_solve(self, Jvals, b, ksp=2, rtol=1e-12, max_it=2000, pc=0):
    """fb_linear_solve on the emulated device (GMRES / tridiagonal; the persistent CG is a
    cooperative launch and is not emulated)."""
    y = np.zeros(self.n_dofs)
    its, rel = C.c_int(0), C.c_double(0.0)
    b = np.ascontiguousarray(b, dtype=np.float64)
    rc = self.lib.fb_linear_solve(self.h, _p(Jvals) if Jvals is not None else None, _p(b), _p(y), ksp, pc, C.c_double(rtol), C.c_double(0.0),
                                  max_it, C.byref(its), C.byref(rel))
    self.check(rc)
    return y, its.value, rel.value, rc


def _newton(self, u, un, atol, rtol, stol=1.49e-10, max_it=30, ksp=2, ksp_rtol=1e-12):
    self.push_state()
    u = np.ascontiguousarray(u, dtype=np.float64).copy()
    un = np.ascontiguousarray(un, dtype=np.float64)
    n_its, reason, lin = C.c_int(0), C.c_int(0), C.c_int(0)
    fnorm = C.c_double(0.0)
    rc = self.lib.fb_newton_solve(self.h, _p(u), _p(un), C.c_double(atol), C.c_double(rtol), C.c_double(stol),
                                  max_it, ksp, 0, C.c_double(ksp_rtol), 20000, C.byref(n_its), C.byref(reason),
                                  C.byref(fnorm), C.byref(lin))
    self.check(rc)
    return u, n_its.value, reason.value, lin.value


def _chebyshev_info(self):
    out = (C.c_double * 4)()
    self.lib.fb_chebyshev_info.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    self.lib.fb_chebyshev_info(self.h, out)
    return list(out)


EmuSolver.linear_solve = _linear_solve
EmuSolver.chebyshev_info = _chebyshev_info
EmuSolver.newton = _newton
