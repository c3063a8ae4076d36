"""The channel tables of the vertex-block assembly (KernelSpec.vx, festim_b200/lowering.py)
evaluated by the numpy model of the device algorithm (tests/vtx_emulator.py) reproduce the
oracle's residual and Jacobian (reference form hydrogen_transport_problem.py:870-997,
reaction.py:179-224) to round-off - on the CPU, before any CUDA runs."""
import warnings

import numpy as np
import pytest

import cases
import festim_b200 as F
import vtx_emulator as E
from festim_b200 import fem
from festim_b200.lowering import KernelSpec
from oracle.newton import OracleBackend

CASES = [
    ("multi_isotope_traps_3d", (3,)), ("mms_1_mobile_1_trap_steady", (3, 3)),
    ("mms_1_mobile_1_trap_steady", (2, 5)), ("mms_1_mobile_1_trap_steady", (1, 30)),
    ("monoblock_coupled_3d", (4,)), ("advection_trap_mms_2d", (6,)), ("tds_1d", (40, 5.0)),
    ("complex_reaction", ()), ("mms_1_mobile_steady", (3, 3)), ("mms_1_mobile_steady", (2, 6)),
    ("permeation_advection_2d", (6,)), ("mms_1_mobile_transient", (60, 0.1, 3)),
    ("coupled_heat_hydrogen_mms", (40, 2)),
]


def _problems(name, args):
    ret = getattr(cases, name)(*args)
    out = []
    for obj in ret if isinstance(ret, (tuple, list)) else [ret]:
        if isinstance(obj, F.CoupledTransientHeatTransferHydrogenTransport):
            obj.heat_problem.solver_backend = OracleBackend()
            obj.hydrogen_problem.solver_backend = OracleBackend()
            obj.initialise()
            obj.heat_problem.u.x.array[:] = 400.0 + 50.0 * np.arange(obj.heat_problem.u.x.array.size) \
                / obj.heat_problem.u.x.array.size
            out += [obj.heat_problem, obj.hydrogen_problem]
        elif isinstance(obj, F.ProblemBase):
            obj.solver_backend = OracleBackend()
            obj.initialise()
            out.append(obj)
    return out


@pytest.mark.parametrize("name,args", CASES, ids=[f"{n}{a}" for n, a in CASES])
def test_channel_tables_reproduce_the_oracle(name, args):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        problems = _problems(name, args)
    checked = 0
    for p in problems:
        spec = KernelSpec(p)
        if spec.vx is None or spec.diff_progs:
            continue
        checked += 1
        asm = p.solver.assembler
        N = p.u.x.array.size
        rng = np.random.default_rng(5)
        scale = 1.0
        T = spec.temperature
        Tval = T.x.array if isinstance(T, fem.Function) else (float(T.value) if T is not None else 300.0)
        scal = spec.scalar_values()
        inv_dt = 1.0 / scal[spec.dt_scalar] if spec.dt_scalar >= 0 else 0.0
        vel = [v.array for _, _, v in spec.adv]
        zero = np.zeros(N)
        F0, _ = asm.assemble(zero, zero, want_J=False)
        Fe0, _ = E.assemble(spec, zero, zero, Tval, inv_dt, scal, vel)
        load = F0 - Fe0
        u = rng.uniform(0.5, 1.5, N) * scale
        un = rng.uniform(0.5, 1.5, N) * scale
        Fo, Jo = asm.assemble(u, un, want_J=True)
        Fe, Je = E.assemble(spec, u, un, Tval, inv_dt, scal, vel, load=load)
        assert np.abs(Fo - Fe).max() <= 1e-11 * max(np.abs(Fo).max(), 1e-300), name
        D = (Jo - Je).tocoo()
        assert np.abs(D.data).max(initial=0.0) <= 1e-11 * np.abs(Jo.data).max(), name
    assert checked or name in ("coupled_heat_hydrogen_mms",)
