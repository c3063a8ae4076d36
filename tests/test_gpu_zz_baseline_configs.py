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


def test_species_dependent_source_parity():
    """reference test/system_tests/test_1_mobile_species.py:285-347: a source on A that depends on
    the concentration of B (a solution-dependent volume term with an off-diagonal Jacobian block)."""
    from cases import error_L2, species_dependent_source_mms

    gm, (Ag, Bg), (A_exact, _) = species_dependent_source_mms(2000)
    rm, (Ar, Br), _ = species_dependent_source_mms(2000)
    rm.solver_backend = OracleBackend()
    for m in (gm, rm):
        m.initialise()
        m.run()
    assert gm.solver.solver.getIterationNumber() == rm.solver.solver.getIterationNumber()
    assert _rel(gm.u.x.array, rm.u.x.array) < 1e-8
    assert error_L2(gm.mesh.mesh, Ag.post_processing_solution.x.array, A_exact) < 1e-3


def test_permeation_multi_volume_parity():
    """reference test/test_permeation_problem.py:116-204 (four volume subdomains, Sieverts
    upstream): outgassing-flux history against the oracle."""
    from cases import permeation_1d

    gm, Hg, fg, _, _ = permeation_1d(mesh_size=1001, final_time=8, dt=0.4, n_volumes=4)
    rm, Hr, fr, _, _ = permeation_1d(mesh_size=1001, final_time=8, dt=0.4, n_volumes=4)
    rm.solver_backend = OracleBackend()
    for m in (gm, rm):
        m.initialise()
        m.run()
    assert gm.timesteps == rm.timesteps
    assert _rel(Hg.post_processing_solution.x.array, Hr.post_processing_solution.x.array) < 1e-8
    _same_exports(gm.exports, rm.exports)


def test_mesh_from_xdmf_and_field_files_through_the_cuda_path(tmp_path):
    """SURVEY 8f3 / 8f4 on the product path: a tagged mesh read from XDMF files, solved by the CUDA
    library, fields written as XDMF and as a VTK series; the files hold the device solution and the
    solution matches the oracle on the same files."""
    import festim_b200 as F
    from festim_b200 import field_io

    base = F.create_unit_cube(4, 3, 3)
    ctag = np.where(base.coords[base.cells].mean(axis=1)[:, 0] < 0.5, 1, 2).astype(np.int32)
    fx = base.coords[base.bfacets][:, :, 0]
    ftag = np.where(np.all(fx == 0, axis=1), 10, np.where(np.all(fx == 1, axis=1), 20, 0)).astype(np.int32)
    tagged = np.nonzero(ftag)[0]
    with field_io.XDMFFile(tmp_path / "volumes.xdmf", "w") as f:
        f.write_meshtags(F.MeshTags(base, 3, np.arange(base.num_cells), ctag), name="Grid")
    with field_io.XDMFFile(tmp_path / "facets.xdmf", "w") as f:
        f.write_meshtags(F.MeshTags(base, 2, tagged, ftag[tagged]), name="Grid")

    def build(tag):
        p = F.HydrogenTransportProblem()
        p.mesh = F.MeshFromXDMF(volume_file=tmp_path / "volumes.xdmf", facet_file=tmp_path / "facets.xdmf")
        H = F.Species("H")
        p.species = [H]
        v1 = F.VolumeSubdomain(id=1, material=F.Material(D_0=2.0, E_D=0.1))
        v2 = F.VolumeSubdomain(id=2, material=F.Material(D_0=0.5, E_D=0.05))
        s1, s2 = F.SurfaceSubdomain(id=10), F.SurfaceSubdomain(id=20)
        p.subdomains = [v1, v2, s1, s2]
        p.boundary_conditions = [F.FixedConcentrationBC(s1, value=1.0, species=H),
                                 F.FixedConcentrationBC(s2, value=0.0, species=H)]
        p.temperature = lambda x: 400.0 + 100.0 * x[0]
        p.exports = [F.XDMFExport(tmp_path / f"{tag}.xdmf", field=H),
                     F.VTXSpeciesExport(tmp_path / f"{tag}.bp", field=H),
                     F.SurfaceFlux(field=H, surface=s2)]
        p.settings = F.Settings(transient=False, atol=1e-12, rtol=1e-12)
        return p, H

    gm, Hg = build("gpu")
    rm, Hr = build("ref")
    rm.solver_backend = OracleBackend()
    for m in (gm, rm):
        m.initialise()
        m.run()
    assert _rel(Hg.post_processing_solution.x.array, Hr.post_processing_solution.x.array) < 1e-8
    _same_exports(gm.exports[2:], rm.exports[2:])
    _, _, data, _ = field_io.read_vtu(tmp_path / "gpu.bp" / "step_000000.vtu")
    assert np.array_equal(data["H"], Hg.post_processing_solution.x.array)
    gm.exports[0]._writer.close()
    rd = field_io.XDMFFile(tmp_path / "gpu.xdmf", "r")
    import xml.etree.ElementTree as ET

    item = ET.parse(tmp_path / "gpu.xdmf").getroot().find(".//Attribute/DataItem")
    assert np.array_equal(rd._read_item(item).ravel(), Hg.post_processing_solution.x.array)


def test_coupled_non_matching_meshes_parity():
    """reference test/system_tests/test_coupled_heat_hydrogen.py:210-332: the temperature field of
    the hydrogen problem is a separate P1 field refreshed from the heat solution before every step;
    the device copy must follow the in-place updates."""
    from cases import coupled_non_matching_mms, error_L2

    cg, mg, exact = coupled_non_matching_mms(199, 499, 10)
    cr, mr, _ = coupled_non_matching_mms(199, 499, 10)
    cr.heat_problem.solver_backend = OracleBackend()
    cr.hydrogen_problem.solver_backend = OracleBackend()
    for c in (cg, cr):
        c.initialise()
        c.run()
    assert cg.hydrogen_problem.timesteps == cr.hydrogen_problem.timesteps
    assert _rel(cg.heat_problem.u.x.array, cr.heat_problem.u.x.array) < 1e-8
    assert _rel(mg.post_processing_solution.x.array, mr.post_processing_solution.x.array) < 1e-8
    assert error_L2(cg.hydrogen_problem.mesh.mesh, mg.post_processing_solution.x.array, exact) < 1e-5


def test_c2_parity_beyond_one_block_per_sm():
    """C2 at n=12 (2197 vertices, 26364 dofs): more vertices than 4 x 148, so the per-SM RCB
    renumbering, the chunk plan with many runs and the channel-operator kernels (TMA / DMMA SpMV,
    Chebyshev-preconditioned GMRES when forced) run on a partition, not on a toy mesh.  Fields and
    the inventory export against the oracle, two time steps."""
    import festim_b200 as F
    from cases import multi_isotope_traps_3d

    def build(cheb):
        m, iso, _ = multi_isotope_traps_3d(12, final_time=2e-2)
        m.exports = [F.TotalVolume(field=iso[0], volume=m.subdomains[0])]
        if cheb:
            m.petsc_options = {"pc_type": "chebyshev"}
        return m

    ref = build(False)
    ref.solver_backend = OracleBackend()
    ref.initialise()
    ref.run()
    for cheb in (False, True):
        g = build(cheb)
        g.initialise()
        assert g.solver.ctx.perm is not None          # renumbered into per-SM parts
        g.run()
        assert g.timesteps == ref.timesteps
        assert _rel(g.u.x.array, ref.u.x.array) < 1e-8
        _same_exports(g.exports, ref.exports)


def test_c1_parity_at_150k_vertices():
    """C1 at n=53 (157 464 vertices: the persistent two-level CG with the operator in shared memory
    at its real size, > 48 KB per block) against the oracle's discrete system.  The problem is
    linear, so the oracle's Newton step is one solve of its assembled system; a sparse LU of a 3D
    system of this size takes minutes, so the oracle's matrix and right-hand side are solved here
    with scipy's CG to 1e-13 instead (test infrastructure, same discrete equations)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    from cases import error_L2, mms_1_mobile_steady

    g, Hg, exact = mms_1_mobile_steady(3, 53)
    g.initialise()
    g.run()
    assert g.solver.solver.getIterationNumber() == 1 and g.solver.solver.getConvergedReason() > 0
    r, Hr, _ = mms_1_mobile_steady(3, 53)
    r.solver_backend = OracleBackend()
    r.initialise()
    x0 = r.u.x.array.copy()
    b, J = r.solver.residual_and_jacobian(x0)
    d = J.diagonal()
    M = spla.LinearOperator(J.shape, matvec=lambda v: v / d)
    y, info = spla.cg(sp.csr_matrix(J), b, rtol=1e-13, atol=0.0, maxiter=5000, M=M)
    assert info == 0
    u_ref = x0 - y
    assert _rel(g.u.x.array, u_ref) < 1e-8
    assert error_L2(g.mesh.mesh, Hg.post_processing_solution.x.array, exact) < 1e-2
