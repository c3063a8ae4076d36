"""Small versions of BASELINE.json configs[3] and configs[4] (C3 coupled heat +
hydrogen on a three-material block, C4 2D permeation with advection and Sieverts / Henry surface
conditions) on the oracle: the cases run, and the physics that makes them what they are shows in
the result.  Their GPU parity tests are in tests/test_gpu_zz_baseline_configs.py."""
import warnings

import numpy as np

from cases import monoblock_coupled_3d, permeation_advection_2d
from oracle.newton import OracleBackend


def test_c3_monoblock_coupled_small():
    coupled, species = monoblock_coupled_3d(4)
    coupled.heat_problem.solver_backend = OracleBackend()
    coupled.hydrogen_problem.solver_backend = OracleBackend()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        coupled.initialise()
        coupled.run()
    heat, hyd = coupled.heat_problem, coupled.hydrogen_problem
    assert hyd.timesteps == heat.timesteps and len(hyd.timesteps) == 3
    z = hyd.mesh.mesh.coords[:, 2]
    T = heat.u.x.array
    assert np.allclose(T[np.isclose(z, 0.0)], 420.0) and T.max() > 450.0      # heated from the top
    H, w1, w2, cu, ccz = (s.post_processing_solution.x.array for s in species)
    assert np.allclose(H[np.isclose(z, 1.0)], 1.0) and np.allclose(H[np.isclose(z, 0.0)], 0.0)
    # every trap fills where its material is: its peak lies in its own slab (outside, only the
    # consistent mass matrix of the continuous P1 field couples it to the interface nodes)
    assert w1.max() > 0.1 and z[np.argmax(w1)] >= 0.5 and w2.max() > 0.05 and z[np.argmax(w2)] >= 0.5
    assert cu.max() > 1e-3 and 0.25 <= z[np.argmax(cu)] <= 0.5
    assert ccz.max() > 1e-4 and z[np.argmax(ccz)] <= 0.25
    inv_H, inv_w1, inv_cu, flux = (np.array(e.data) for e in hyd.exports)
    assert np.all(np.diff(inv_H) > 0) and np.all(np.diff(inv_w1) > 0) and np.all(np.diff(inv_cu) > 0)
    assert np.all(flux == 0.0)      # the reference's stale-D_global quirk, see the case's docstring


def test_c4_permeation_advection_small():
    m, H = permeation_advection_2d(8)
    m.solver_backend = OracleBackend()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m.initialise()
        m.run()
    x, y = m.mesh.mesh.coords.T
    L = x.max()
    c = H.post_processing_solution.x.array
    k_B = 8.6173303e-5
    assert np.allclose(c[np.isclose(x, 0.0)], 4.02e21 * np.exp(-1.04 / k_B / 500.0) * np.sqrt(100.0))  # Sieverts
    t_end = float(m.t)
    assert np.allclose(c[np.isclose(x, L)], 6.0e17 * np.exp(-0.9 / k_B / 500.0) * (1.0 + 0.5 * t_end))  # Henry
    flux, inventory = (np.array(e.data) for e in m.exports)
    assert np.all(np.diff(flux) > 0) and np.all(np.diff(inventory) > 0)
    # the flow is fastest on the centre line: at mid-depth the concentration there runs ahead
    mid = np.isclose(x, L / 2)
    centre = c[mid & np.isclose(y, L / 2)][0]
    wall = c[mid & np.isclose(y, 0.0)][0]
    assert centre > wall
