"""GPU parity for small versions of BASELINE.json configs[3] and configs[4] (the CUDA path through
the C ABI against the oracle, BASELINE tolerances: fields 1e-8 relative L2, derived quantities
1e-6)."""
import warnings

import numpy as np
import pytest

from cases import monoblock_coupled_3d, permeation_advection_2d
from oracle.newton import OracleBackend

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _same_exports(gpu_exports, ref_exports):
    for eg, er in zip(gpu_exports, ref_exports):
        dg, dr = np.array(eg.data), np.array(er.data)
        assert dg.shape == dr.shape
        assert np.abs(dg - dr).max() <= 1e-6 * max(np.abs(dr).max(), 1e-300), eg.title


def test_c3_monoblock_coupled_parity():
    """C3 in small: three materials, traps restricted to their material, heat flux + fixed
    temperature, staggered coupling with the temperature field shared between the problems."""
    cg, sg = monoblock_coupled_3d(4)
    cr, sr = monoblock_coupled_3d(4)
    cr.heat_problem.solver_backend = OracleBackend()
    cr.hydrogen_problem.solver_backend = OracleBackend()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for c in (cg, cr):
            c.initialise()
            c.run()
    assert cg.hydrogen_problem.timesteps == cr.hydrogen_problem.timesteps
    assert _rel(cg.heat_problem.u.x.array, cr.heat_problem.u.x.array) < 1e-8
    for a, b in zip(sg, sr):
        assert _rel(a.post_processing_solution.x.array, b.post_processing_solution.x.array) < 1e-8, a.name
    _same_exports(cg.hydrogen_problem.exports, cr.hydrogen_problem.exports)


def test_c4_permeation_advection_parity():
    """C4 in small: advection (GMRES), Sieverts upstream, time-dependent Henry downstream."""
    gm, Hg = permeation_advection_2d(8)
    rm, Hr = permeation_advection_2d(8)
    rm.solver_backend = OracleBackend()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for m in (gm, rm):
            m.initialise()
            m.run()
    assert gm.timesteps == rm.timesteps
    assert _rel(Hg.post_processing_solution.x.array, Hr.post_processing_solution.x.array) < 1e-8
    _same_exports(gm.exports, rm.exports)
