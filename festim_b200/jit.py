"""Run-time specialisation of the element kernels (NVRTC).

The reference JIT-compiles C ``tabulate_tensor`` kernels from its UFL form with
FFCx every time a problem is created (``NonlinearProblem``, reference
``src/festim/problem.py:173-182``; explicit JIT flags at
``hydrogen_transport_problem.py:2025-2040``).  The B200 analogue: the hand-written
kernel bodies of ``csrc/assemble_kernels.cuh`` are compiled once more with NVRTC
for sm_100a, with ``fb_prog_eval`` - the evaluation of the problem's closed-form
expressions (sources, fluxes, reaction rates and their species derivatives) -
generated as straight-line fp64 code with common subexpressions shared, instead
of interpreting byte-code.  The kernels, the data layout and the C ABI are
unchanged; only the expression evaluator is specialised.

``FB_NO_JIT=1`` disables the specialisation (the ahead-of-time kernels then
interpret the byte-code; both paths are parity-tested).
"""

from __future__ import annotations

import os

from festim_b200 import ufl

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")

_MATH_C = {"exp": "exp", "ln": "log", "sqrt": "sqrt", "sin": "sin", "cos": "cos", "tan": "tan",
           "tanh": "tanh", "sinh": "sinh", "cosh": "cosh", "abs": "fabs", "erf": "erf",
           "atan": "atan", "asin": "asin", "acos": "acos"}


class _Gen:
    """Emits one program as C statements with hash-consed temporaries."""

    def __init__(self, pb):
        self.pb = pb
        self.lines = []
        self.memo = {}
        self.keymemo = {}
        self.n = 0

    def tmp(self, code):
        name = f"t{self.n}"
        self.n += 1
        self.lines.append(f"      const double {name} = {code};")
        return name

    def emit(self, ex):
        pb = self.pb
        k = pb._skey(ex, self.keymemo)
        if k in self.memo:
            return self.memo[k]
        if isinstance(ex, ufl.Condition):
            if ex.op == "not":
                r = self.tmp(f"({self.emit(ex.a)} == 0.0) ? 1.0 : 0.0")
            elif ex.op in ("and", "or"):
                op = "&&" if ex.op == "and" else "||"
                r = self.tmp(f"(({self.emit(ex.a)} != 0.0) {op} ({self.emit(ex.b)} != 0.0)) ? 1.0 : 0.0")
            else:
                op = {"lt": "<", "gt": ">", "le": "<=", "ge": ">=", "eq": "==", "ne": "!="}[ex.op]
                r = self.tmp(f"({self.emit(ex.a)} {op} {self.emit(ex.b)}) ? 1.0 : 0.0")
        elif isinstance(ex, ufl.Const):
            r = repr(float(ex.value))
            if r in ("inf", "-inf", "nan"):
                r = {"inf": "(1.0/0.0)", "-inf": "(-1.0/0.0)", "nan": "(0.0/0.0)"}[r]
            else:
                r = f"({r})"
        elif isinstance(ex, ufl.Coord):
            r = f"P.x[{ex.i}]"
        elif isinstance(ex, ufl.ConstantRef):
            if pb.temperature is not None and ex.constant is pb.temperature:
                r = "P.T"
            else:
                r = f"S.scalars[{pb._scalar_id(ex.constant)}]"
        elif isinstance(ex, ufl.FieldRef):
            if pb.temperature is not None and ex.function is pb.temperature:
                r = "P.T"
            else:
                r = self.tmp(f"fb_field_value(S, {pb._field_id(ex.function)}, P)")
        elif isinstance(ex, ufl.FieldGrad):
            r = self.tmp(f"fb_field_grad(S, {pb._field_id(ex.function)}, {ex.i}, P)")
        elif isinstance(ex, ufl.SpeciesRef):
            arr = "P.cprev" if ex.prev else "P.c"
            r = f"{arr}[{pb.species_index[ex.species]}]"
        elif isinstance(ex, ufl.SpeciesGrad):
            r = f"(P.cgrad ? P.cgrad[{pb.species_index[ex.species] * 3 + ex.i}] : 0.0)"
        elif isinstance(ex, ufl.NormalComp):
            r = f"P.n[{ex.i}]"
        elif isinstance(ex, ufl.CellCircumradius):
            r = "P.h"
        elif isinstance(ex, ufl.Neg):
            r = self.tmp(f"-{self.emit(ex.a)}")
        elif isinstance(ex, ufl.Math):
            a = self.emit(ex.a)
            if ex.name == "sign":
                r = self.tmp(f"(double)(({a} > 0.0) - ({a} < 0.0))")
            else:
                r = self.tmp(f"{_MATH_C[ex.name]}({a})")
        elif isinstance(ex, ufl.Sum):
            r = self.tmp(f"{self.emit(ex.a)} + {self.emit(ex.b)}")
        elif isinstance(ex, ufl.Product):
            r = self.tmp(f"{self.emit(ex.a)} * {self.emit(ex.b)}")
        elif isinstance(ex, ufl.Division):
            r = self.tmp(f"{self.emit(ex.a)} / {self.emit(ex.b)}")
        elif isinstance(ex, ufl.Power):
            r = self.tmp(f"pow({self.emit(ex.a)}, {self.emit(ex.b)})")
        elif isinstance(ex, ufl.Conditional):
            r = self.tmp(f"({self.emit(ex.cond)} != 0.0) ? {self.emit(ex.a)} : {self.emit(ex.b)}")
        elif isinstance(ex, ufl.MinMax):
            fn = "fmax" if ex.kind == "max" else "fmin"
            r = self.tmp(f"{fn}({self.emit(ex.a)}, {self.emit(ex.b)})")
        else:
            raise TypeError(f"cannot generate code for {type(ex)}")
        self.memo[k] = r
        return r


def generate_source(spec):
    """CUDA translation unit specialising the kernels for ``spec``."""
    pb = spec.programs
    cases = []
    for pid, ex in enumerate(pb.expressions):
        g = _Gen(pb)
        res = g.emit(ex)
        body = "\n".join(g.lines)
        cases.append(f"    case {pid}: {{\n{body}\n      return {res};\n    }}")
    switch = "\n".join(cases) if cases else "    default: break;"
    dim = spec.tdim
    return f'''// generated by festim_b200/jit.py -- do not edit
#include "fb_device.cuh"

__device__ __forceinline__ double fb_field_value(const DevSpec& S, int k, const PointCtx& P) {{
  const double* f = S.fields[k];
  double v = 0.0;
  for (int a = 0; a < P.nverts; ++a) v += P.bary[a] * f[P.verts[a]];
  return v;
}}
__device__ __forceinline__ double fb_field_grad(const DevSpec& S, int k, int comp, const PointCtx& P) {{
  const double* f = S.fields[k];
  double v = 0.0;
  for (int a = 0; a < P.ncverts; ++a) v += f[P.cverts[a]] * P.G[a * P.tdim + comp];
  return v;
}}

__device__ __forceinline__ double fb_prog_eval(const DevSpec& S, int pid, const PointCtx& P) {{
  switch (pid) {{
{switch}
  }}
  return 0.0;
}}

#include "assemble_kernels.cuh"

extern "C" __global__ void jit_cell_coeffs(DevSpec S, double* __restrict__ dbar) {{
  fb_cell_coeffs_body<{dim}>(S, dbar);
}}
extern "C" __global__ void jit_load_cells(DevSpec S, double* __restrict__ load) {{
  fb_load_cells_body<{dim}>(S, load);
}}
extern "C" __global__ void jit_facets(DevSpec S, int mode, const double* __restrict__ u,
                                      const double* __restrict__ un, double* __restrict__ F,
                                      double* __restrict__ J) {{
  fb_facets_body<{dim}>(S, mode, u, un, F, J);
}}
extern "C" __global__ void __launch_bounds__(128)
jit_assemble_nodes(DevSpec S, const double* __restrict__ u, const double* __restrict__ un,
                   double* __restrict__ F, double* __restrict__ J, int flags) {{
  fb_assemble_nodes_body<{dim}>(S, u, un, F, J, flags);
}}
'''


_CACHE = {}


def compile_cubin(source, arch="sm_100a"):
    """NVRTC: CUDA source -> cubin for ``arch`` (memoised per source text)."""
    key = (arch, source)
    if key not in _CACHE:
        _CACHE[key] = _compile(source, arch)
    return _CACHE[key]


def _compile(source, arch):
    """Raises with the compiler log on failure."""
    from cuda.bindings import nvrtc

    def check(res):
        err = res[0]
        if err != nvrtc.nvrtcResult.NVRTC_SUCCESS:
            raise RuntimeError(f"NVRTC error {err}")
        return res[1:] if len(res) > 1 else None

    (prog,) = check(nvrtc.nvrtcCreateProgram(source.encode(), b"festim_b200_jit.cu", 0, [], []))
    opts = [f"--gpu-architecture={arch}".encode(), b"--std=c++17", b"-lineinfo",
            f"--include-path={CSRC}".encode(), b"--fmad=true"]
    res = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
    if res[0] != nvrtc.nvrtcResult.NVRTC_SUCCESS:
        (n,) = check(nvrtc.nvrtcGetProgramLogSize(prog))
        log = b" " * n
        nvrtc.nvrtcGetProgramLog(prog, log)
        raise RuntimeError("NVRTC compilation failed:\n" + log.decode(errors="replace"))
    (n,) = check(nvrtc.nvrtcGetCUBINSize(prog))
    cubin = b" " * n
    check(nvrtc.nvrtcGetCUBIN(prog, cubin))
    nvrtc.nvrtcDestroyProgram(prog)
    return cubin


def needed(spec):
    """Specialise only when there is an expression to compile."""
    return spec.programs.n_programs > 0 and not os.environ.get("FB_NO_JIT")
