"""The CPU oracle against the reference's own known-answer tests.
Thresholds are the reference's; each test cites the reference test it restates."""
import numpy as np
import pytest

from cases import error_L2, mms_1_mobile_1_trap_steady, mms_1_mobile_steady
from oracle.newton import OracleBackend


def _run(model):
    model.solver_backend = OracleBackend()
    model.initialise()
    model.run()
    return model


def test_1_mobile_MMS_steady_state_1d():
    """reference test/system_tests/test_1_mobile_species.py:20-80 (L2 < 1e-7)."""
    m, H, exact = mms_1_mobile_steady(1, 9999)
    _run(m)
    assert m.solver.solver.getConvergedReason() > 0
    assert error_L2(m.mesh.mesh, H.post_processing_solution.x.array, exact) < 1e-7


def test_1_mobile_MMS_3D():
    """reference test/system_tests/test_1_mobile_species.py:216-282 (L2 < 1e-2, n=20)."""
    m, H, exact = mms_1_mobile_steady(3, 20)
    _run(m)
    assert m.solver.solver.getIterationNumber() == 1
    assert error_L2(m.mesh.mesh, H.post_processing_solution.x.array, exact) < 1e-2


def test_1_mobile_MMS_2D():
    """2D analogue on the 50 x 50 square of test/system_tests/tools.py:11."""
    m, H, exact = mms_1_mobile_steady(2, 50)
    _run(m)
    assert error_L2(m.mesh.mesh, H.post_processing_solution.x.array, exact) < 5e-3


def test_1_mobile_1_trap_MMS_steady_state_1d():
    """reference test/system_tests/test_1_mobile_1_trap.py:20-113 (2e-7 / 1e-7)."""
    m, (mobile, trapped), (ue, ve) = mms_1_mobile_1_trap_steady(1, 9999)
    _run(m)
    assert m.solver.solver.getConvergedReason() > 0
    mesh = m.mesh.mesh
    assert error_L2(mesh, mobile.post_processing_solution.x.array, ue) < 2e-7
    assert error_L2(mesh, trapped.post_processing_solution.x.array, ve) < 1e-7
