"""Host side of the product path without a GPU: every parity case lowers to a flat kernel spec
(`festim_b200/lowering.py`), the blob is well formed, lowering is deterministic, and the run-time
specialisation (`festim_b200/jit.py`: generated CUDA source -> NVRTC -> sm_100a cubin, the analogue
of the FFCx JIT behind the reference's NonlinearProblem, problem.py:173-182) compiles offline.  The
numerical parity of what these specs compute is the `-m gpu` suite's job."""
import warnings

import numpy as np
import pytest

import cases
import festim_b200 as F
from festim_b200 import jit
from festim_b200.lowering import KernelSpec
from oracle.newton import OracleBackend

CASES = [
    ("mms_1_mobile_steady", (1, 40)), ("mms_1_mobile_steady", (2, 6)), ("mms_1_mobile_steady", (3, 3)),
    ("mms_1_mobile_1_trap_steady", (1, 40)), ("mms_1_mobile_1_trap_steady", (2, 6)),
    ("mms_1_mobile_1_trap_steady", (3, 3)), ("mms_1_mobile_transient", (60, 0.1, 3)),
    ("permeation_1d", (51, 1, 0.3)), ("tds_1d", (40, 5.0)), ("heat_mms", (False, 40)),
    ("heat_mms", (True, 40)), ("flux_bc_mms", (False, 40)), ("flux_bc_mms", (True, 40)),
    ("coupled_heat_hydrogen_mms", (40, 2)), ("advection_trap_mms_2d", (6,)),
    ("multi_isotope_traps_3d", (3,)), ("complex_reaction", ()), ("recombination_flux", (2, 5)),
    ("monoblock_coupled_3d", (4,)), ("permeation_advection_2d", (6,)),
]


def _problems(name, args):
    ret = getattr(cases, name)(*args)
    out = []
    for obj in ret if isinstance(ret, (tuple, list)) else [ret]:
        if isinstance(obj, F.CoupledTransientHeatTransferHydrogenTransport):
            obj.heat_problem.solver_backend = OracleBackend()
            obj.hydrogen_problem.solver_backend = OracleBackend()
            obj.initialise()
            out += [obj.heat_problem, obj.hydrogen_problem]
        elif isinstance(obj, F.ProblemBase):
            obj.solver_backend = OracleBackend()   # the seam; lowering does not depend on it
            obj.initialise()
            out.append(obj)
    assert out
    return out


def _sections(blob):
    """Inverse of lowering.pack_sections: {section id: array}."""
    head = np.frombuffer(blob, dtype=np.int64, count=2)
    n = int(head[1])
    table = np.frombuffer(blob, dtype=np.int64, count=2 + 4 * n)[2:].reshape(n, 4)
    out = {}
    for sid, code, size, offset in table:
        dtype = np.int32 if code == 0 else np.float64
        assert offset % 8 == 0 and offset + size * np.dtype(dtype).itemsize <= len(blob)
        out[int(sid)] = np.frombuffer(blob, dtype=dtype, count=int(size), offset=int(offset))
    assert len(out) == n
    return out


@pytest.mark.parametrize("name,args", CASES, ids=[f"{n}{a}" for n, a in CASES])
def test_case_lowers(name, args):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        problems = _problems(name, args)
        again = _problems(name, args)
    for p, q in zip(problems, again):
        spec = KernelSpec(p)
        blob = spec.blob()
        assert blob == KernelSpec(q).blob(), "lowering must be deterministic"
        sec = _sections(blob)
        assert all(np.isfinite(a).all() for a in sec.values() if a.dtype == np.float64)
        assert spec.tdim == p.mesh.mesh.tdim and spec.ns == max(len(getattr(p, "species", [])), 1)
        assert len(spec.scalar_values()) == len(spec.programs.scalars)
        if jit.needed(spec):
            src = jit.generate_source(spec)
            assert src.count("case ") == spec.programs.n_programs


@pytest.mark.parametrize("name,args", [("mms_1_mobile_1_trap_steady", (3, 3)), ("recombination_flux", (2, 5)),
                                       ("advection_trap_mms_2d", (6,))],
                         ids=["trap3d", "recombination2d", "advection2d"])
def test_specialised_kernels_compile_for_sm_100a(name, args):
    """NVRTC needs no GPU: the generated element kernels build to a cubin for sm_100a."""
    pytest.importorskip("cuda.bindings.nvrtc")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec = KernelSpec(_problems(name, args)[-1])
    assert jit.needed(spec)
    cubin = jit.compile_cubin(jit.generate_source(spec))
    assert cubin[:4] == b"\x7fELF" and len(cubin) > 10_000
