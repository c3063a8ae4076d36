"""Field output and mesh input: XDMF and VTK-XML writers behind the reference's
export classes (``exports/xdmf.py``, ``exports/vtx.py``), ``MeshFromXDMF`` (``mesh/
mesh_from_xdmf.py``), ``Profile1DExport``, ``CustomFieldExport`` / ``ReactionRateExport``.  Problems
solve through the oracle backend at the create_solver seam; the files are read back and compared
with the nodal arrays they were written from."""
import os

import numpy as np
import pytest

import festim_b200 as F
from festim_b200 import field_io
from festim_b200.constants import k_B
from oracle.newton import OracleBackend


def _problem_1d(exports, transient=True, temperature=500.0):
    m = F.HydrogenTransportProblem()
    m.mesh = F.Mesh1D(vertices=np.linspace(0, 1, num=21))
    H, trapped = F.Species("H"), F.Species("trapped", mobile=False)
    empty = F.ImplicitSpecies(n=2.0, others=[trapped], name="empty")
    m.species = [H, trapped]
    vol = F.VolumeSubdomain1D(id=1, borders=[0, 1], material=F.Material(D_0=1, E_D=0.1))
    left, right = F.SurfaceSubdomain1D(id=1, x=0), F.SurfaceSubdomain1D(id=2, x=1)
    m.subdomains = [vol, left, right]
    m.reactions = [F.Reaction(reactant=[H, empty], product=trapped, k_0=3.0, E_k=0.2, p_0=0.5, E_p=0.1,
                              volume=vol)]
    m.boundary_conditions = [F.FixedConcentrationBC(subdomain=left, value=1.0, species=H),
                             F.FixedConcentrationBC(subdomain=right, value=0.0, species=H)]
    m.temperature = temperature
    m.exports = exports
    m.settings = F.Settings(transient=transient, atol=1e-10, rtol=1e-10, final_time=0.3, stepsize=0.1)
    m.solver_backend = OracleBackend()
    return m, H, trapped, empty


def test_xdmf_export_round_trip(tmp_path):
    """reference test/test_xdmf.py (writer defined, file written); here the file is also read back."""
    name = tmp_path / "out" / "fields.xdmf"
    m, H, trapped, _ = _problem_1d([])
    export = F.XDMFExport(name, field=[H, trapped])
    profile = F.Profile1DExport(field=H)
    m.exports = [export, profile]
    m.initialise()
    m.run()
    export._writer.close()
    assert os.path.exists(name) and os.path.exists(name.with_suffix(".bin"))

    rd = field_io.XDMFFile(name, "r")
    mesh = rd.read_mesh("mesh")
    assert mesh.tdim == 1 and mesh.num_cells == 20
    assert np.allclose(mesh.coords[:, 0], m.mesh.mesh.coords[:, 0])
    # one temporal collection per species, one grid per step, last step == final nodal values
    import xml.etree.ElementTree as ET

    root = ET.parse(name).getroot()
    colls = {g.get("Name"): g for g in root.iter("Grid") if g.get("GridType") == "Collection"}
    assert set(colls) == {"H", "trapped"}
    for spe in (H, trapped):
        grids = colls[spe.name].findall("Grid")
        assert len(grids) == len(profile.t) == 3
        times = [float(g.find("Time").get("Value")) for g in grids]
        assert np.allclose(times, profile.t)
        last = rd._read_item(grids[-1].find("Attribute/DataItem")).ravel()
        assert np.array_equal(last, spe.post_processing_solution.x.array)
    # the 1D profile is the same field sorted by x
    assert np.array_equal(profile.x, np.sort(m.mesh.mesh.coords[:, 0]))
    assert np.array_equal(profile.data[-1], H.post_processing_solution.x.array[np.argsort(m.mesh.mesh.coords[:, 0])])


def test_xdmf_export_field_validation(tmp_path):
    """reference test/test_xdmf.py: only Species (or lists of them) are accepted."""
    with pytest.raises(TypeError):
        F.XDMFExport(tmp_path / "a.xdmf", field=1)
    with pytest.raises(TypeError):
        F.XDMFExport(tmp_path / "a.xdmf", field=[F.Species("H"), 1])
    with pytest.warns(UserWarning, match="does not have .xdmf extension"):
        e = F.XDMFExport(tmp_path / "a.txt", field=F.Species("H"))
    assert str(e.filename).endswith("a.xdmf")


@pytest.mark.parametrize("heavy", ["binary", "xml"])
@pytest.mark.parametrize("dim", [2, 3])
def test_mesh_from_xdmf(tmp_path, heavy, dim):
    """A tagged mesh written as XDMF (dolfinx layout) comes back through MeshFromXDMF with the same
    cells, cell tags and exterior-facet tags, also when the file carries points no cell uses and
    the cells in another order."""
    base = F.create_unit_square(6, 5) if dim == 2 else F.create_unit_cube(3, 4, 2)
    rng = np.random.default_rng(7)
    # shuffle vertex numbers, add two stray points (meshio keeps points of dropped blocks)
    perm = rng.permutation(base.num_vertices + 2)
    coords = np.zeros((base.num_vertices + 2, dim))
    coords[perm[: base.num_vertices]] = base.coords
    coords[perm[base.num_vertices:]] = 9.0
    cells = perm[base.cells]
    file_mesh = F.SimplexMesh.__new__(F.SimplexMesh)
    file_mesh.coords, file_mesh.cells = coords, cells.astype(np.int32)
    file_mesh.gdim = file_mesh.tdim = dim
    file_mesh._bfacets = perm[base.bfacets].astype(np.int32)

    ctag = np.where(base.coords[base.cells].mean(axis=1)[:, 0] < 0.5, 1, 2).astype(np.int32)
    fx = base.coords[base.bfacets][:, :, 0]
    ftag = np.where(np.all(fx == 0, axis=1), 10, np.where(np.all(fx == 1, axis=1), 20, 0)).astype(np.int32)
    tagged = np.nonzero(ftag)[0]

    vfile, ffile = tmp_path / "volumes.xdmf", tmp_path / "facets.xdmf"
    with field_io.XDMFFile(vfile, "w", heavy=heavy) as f:
        grid_perm = rng.permutation(base.num_cells)
        file_mesh.cells = file_mesh.cells[grid_perm]
        f.write_meshtags(F.MeshTags(file_mesh, dim, np.arange(base.num_cells), ctag[grid_perm]), name="Grid")
    with field_io.XDMFFile(ffile, "w", heavy=heavy) as f:
        f.write_meshtags(F.MeshTags(file_mesh, dim - 1, tagged, ftag[tagged]), name="Grid")

    mesh = F.MeshFromXDMF(volume_file=vfile, facet_file=ffile)
    m = mesh.mesh
    assert m.num_vertices == base.num_vertices and m.num_cells == base.num_cells
    assert np.isclose(m.cell_volumes().sum(), 1.0)
    vt, ft = mesh.define_volume_meshtags(), mesh.define_surface_meshtags()
    cx = m.coords[m.cells].mean(axis=1)[:, 0]
    assert np.array_equal(vt.values, np.where(cx < 0.5, 1, 2))
    bx = m.coords[m.bfacets][:, :, 0]
    assert np.array_equal(ft.values, np.where(np.all(bx == 0, axis=1), 10, np.where(np.all(bx == 1, axis=1), 20, 0)))

    # and a problem on it solves: steady diffusion between the two tagged faces, u = 1 - x
    p = F.HydrogenTransportProblem()
    p.mesh = mesh
    H = F.Species("H")
    p.species = [H]
    mat = F.Material(D_0=2.0, E_D=0.0)
    v1, v2 = F.VolumeSubdomain(id=1, material=mat), F.VolumeSubdomain(id=2, material=mat)
    s1, s2 = F.SurfaceSubdomain(id=10), F.SurfaceSubdomain(id=20)
    p.subdomains = [v1, v2, s1, s2]
    p.boundary_conditions = [F.FixedConcentrationBC(s1, value=1.0, species=H),
                             F.FixedConcentrationBC(s2, value=0.0, species=H)]
    p.temperature = 400.0
    flux = F.SurfaceFlux(field=H, surface=s2)
    p.exports = [flux]
    p.settings = F.Settings(transient=False, atol=1e-12, rtol=1e-12)
    p.solver_backend = OracleBackend()
    p.initialise()
    p.run()
    assert np.allclose(H.post_processing_solution.x.array, 1.0 - m.coords[:, 0], atol=1e-10)
    assert np.isclose(flux.data[-1], 2.0, rtol=1e-9)


def test_xdmf_reader_errors(tmp_path):
    with pytest.raises(FileNotFoundError):
        field_io.XDMFFile(tmp_path / "missing.xdmf", "r")
    mesh = F.create_unit_square(2, 2)
    with field_io.XDMFFile(tmp_path / "m.xdmf", "w", heavy="xml") as f:
        f.write_mesh(mesh, name="Grid")
    rd = field_io.XDMFFile(tmp_path / "m.xdmf", "r")
    with pytest.raises(ValueError, match="no grid named"):
        rd.read_mesh("nope")
    with pytest.raises(ValueError, match="no <Attribute>"):
        rd.read_meshtags(rd.read_mesh("Grid"), "Grid")
    with pytest.raises(RuntimeError, match="not open for writing"):
        rd.write_mesh(mesh)


def test_vtx_species_export(tmp_path):
    """reference test/test_vtx.py: `.bp` path exists after the run, export times become milestones,
    file-name / field validation."""
    m, H, trapped, _ = _problem_1d([])
    name = tmp_path / "species.bp"
    export = F.VTXSpeciesExport(name, field=["H", trapped], times=[0.1, 0.3])
    m.exports = [export]
    with pytest.warns(UserWarning, match="added to milestones"):
        m.initialise()
    assert export.field[0] is H
    assert 0.3 in m.settings.stepsize.milestones
    m.run()
    assert os.path.isdir(name) and os.path.exists(name / "series.pvd")
    steps = sorted(p.name for p in name.glob("step_*.vtu"))
    assert steps == ["step_000000.vtu", "step_000001.vtu"]
    xyz, cells, data, t = field_io.read_vtu(name / steps[-1])
    assert np.isclose(t, 0.3)
    assert np.allclose(xyz[:, 0], m.mesh.mesh.coords[:, 0]) and np.array_equal(cells, m.mesh.mesh.cells)
    assert set(data) == {"H", "trapped"}

    with pytest.raises(TypeError):
        F.VTXSpeciesExport(tmp_path / "a.bp", field=[1])
    with pytest.raises(TypeError):
        F.VTXSpeciesExport(tmp_path / "a.bp", field=1)
    with pytest.warns(UserWarning, match="does not have .bp extension"):
        F.VTXTemperatureExport(tmp_path / "a.txt")


def test_vtx_last_step_holds_final_fields(tmp_path):
    m, H, trapped, _ = _problem_1d([], transient=False)
    export = F.VTXSpeciesExport(tmp_path / "s.bp", field=[H, trapped])
    m.exports = [export]
    m.initialise()
    m.run()
    _, _, data, _ = field_io.read_vtu(tmp_path / "s.bp" / "step_000000.vtu")
    assert np.array_equal(data["H"], H.post_processing_solution.x.array)
    assert np.array_equal(data["trapped"], trapped.post_processing_solution.x.array)


def test_temperature_export(tmp_path):
    """A time-dependent temperature is written at every export time (reference
    hydrogen_transport_problem.py:1079-1086); a steady one only creates the (empty) series."""
    m, *_ = _problem_1d([], temperature=lambda x, t: 300.0 + 50.0 * x[0] + 10.0 * t)
    m.exports = [F.VTXTemperatureExport(tmp_path / "T.bp")]
    m.initialise()
    m.run()
    files = sorted((tmp_path / "T.bp").glob("step_*.vtu"))
    assert len(files) == 3
    xyz, _, data, t = field_io.read_vtu(files[-1])
    assert np.allclose(data["T"], 300.0 + 50.0 * xyz[:, 0] + 10.0 * t)

    m2, *_ = _problem_1d([F.VTXTemperatureExport(tmp_path / "T0.bp")])
    m2.initialise()
    m2.run()
    assert os.path.exists(tmp_path / "T0.bp" / "series.pvd")
    assert not list((tmp_path / "T0.bp").glob("step_*.vtu"))


def test_heat_problem_temperature_export(tmp_path):
    """reference heat_transfer_problem.py:229-236, 266-267."""
    p = F.HeatTransferProblem()
    p.mesh = F.Mesh1D(vertices=np.linspace(0, 2, num=11))
    mat = F.Material(D_0=1, E_D=0, thermal_conductivity=3.0)
    vol = F.VolumeSubdomain1D(id=1, borders=[0, 2], material=mat)
    left, right = F.SurfaceSubdomain1D(id=1, x=0), F.SurfaceSubdomain1D(id=2, x=2)
    p.subdomains = [vol, left, right]
    p.boundary_conditions = [F.FixedTemperatureBC(subdomain=left, value=300.0),
                             F.FixedTemperatureBC(subdomain=right, value=500.0)]
    p.exports = [F.VTXTemperatureExport(tmp_path / "heat.bp")]
    p.settings = F.Settings(transient=False, atol=1e-10, rtol=1e-10)
    p.solver_backend = OracleBackend()
    p.initialise()
    p.run()
    xyz, _, data, _ = field_io.read_vtu(tmp_path / "heat.bp" / "step_000000.vtu")
    assert np.allclose(data["T"], 300.0 + 100.0 * xyz[:, 0])
    p.exports = [F.XDMFExport(tmp_path / "heat.xdmf", field=F.Species("H"))]
    with pytest.raises(NotImplementedError, match="XDMF export is not implemented"):
        p.initialise_exports()


def test_custom_field_and_reaction_rate_exports(tmp_path):
    """reference exports/vtx.py:183-440: nodal values of a user expression of t, x, T and species
    (explicit and implicit), and of a reaction's forward / backward / net rate."""
    m, H, trapped, empty = _problem_1d([])
    T = 500.0
    custom = F.CustomFieldExport(
        tmp_path / "custom.bp",
        expression=lambda t, x, T, c, e: 2.0 * c * e + x[0] + t + 0.0 * T,
        species_dependent_value={"c": H, "e": empty})
    reaction = m.reactions[0]
    rates = {d: F.ReactionRateExport(reaction, tmp_path / f"rate_{d}.bp", direction=d)
             for d in ("forward", "backward", "both")}
    m.exports = [custom, *rates.values()]
    m.initialise()
    m.run()
    x = m.mesh.mesh.coords[:, 0]
    cH, cT = H.post_processing_solution.x.array, trapped.post_processing_solution.x.array
    _, _, data, t = field_io.read_vtu(sorted((tmp_path / "custom.bp").glob("step_*.vtu"))[-1])
    assert np.allclose(data["f"], 2.0 * cH * (2.0 - cT) + x + t, rtol=1e-14)
    fwd = 3.0 * np.exp(-0.2 / (k_B * T)) * cH * (2.0 - cT)
    bwd = 0.5 * np.exp(-0.1 / (k_B * T)) * cT
    expected = {"forward": fwd, "backward": bwd, "both": fwd - bwd}
    for d, export in rates.items():
        _, _, data, _ = field_io.read_vtu(sorted((tmp_path / f"rate_{d}.bp").glob("step_*.vtu"))[-1])
        assert np.allclose(data["f"], expected[d], rtol=1e-13, atol=1e-300), d
    assert np.abs(expected["both"]).max() > 0

    bad = F.CustomFieldExport(tmp_path / "bad.bp", expression=lambda unknown: unknown)
    m2, *_ = _problem_1d([bad])
    with pytest.raises(AssertionError, match="not found in species_dependent_value"):
        m2.initialise()


def test_checkpoint_and_restart(tmp_path):
    """reference hydrogen_transport_problem.py:464-469, 1069-1076 (write) and
    initial_condition.py:238-283 / test/test_initial_condition.py:144-270 (read): a run restarted
    from the checkpoint of t = 0.2 reaches t = 0.3 with the fields of the uninterrupted run."""
    m, H, trapped, _ = _problem_1d([])
    m.exports = [F.VTXSpeciesExport(tmp_path / "restart.bp", field=[H, trapped], checkpoint=True)]
    m.initialise()
    m.run()
    m.exports[0].writer.close()
    assert (tmp_path / "restart.bp" / "checkpoint.xdmf").exists()
    rd = field_io.XDMFFile(tmp_path / "restart.bp" / "checkpoint.xdmf", "r")
    assert np.allclose(rd.function_times("H"), [0.1, 0.2, 0.3]) and np.allclose(rd.function_times("trapped"), [0.1, 0.2, 0.3])
    final = F.read_function_from_file(tmp_path / "restart.bp", name="H", timestamp=0.3)
    assert np.array_equal(final.x.array, H.post_processing_solution.x.array)

    r, Hr, tr, _ = _problem_1d([])
    r.settings.final_time = 0.1
    vol = r.volume_subdomains[0]
    r.initial_conditions = [
        F.InitialConcentration(value=F.read_function_from_file(tmp_path / "restart.bp", "H", 0.2, mesh=r.mesh.mesh),
                               species=Hr, volume=vol),
        F.InitialConcentration(value=F.read_function_from_file(tmp_path / "restart.bp", "trapped", 0.2),
                               species=tr, volume=vol)]
    r.initialise()
    r.run()
    assert np.allclose(Hr.post_processing_solution.x.array, H.post_processing_solution.x.array, rtol=1e-8, atol=1e-9)
    assert np.allclose(tr.post_processing_solution.x.array, trapped.post_processing_solution.x.array, rtol=1e-8, atol=1e-9)

    with pytest.raises(ValueError, match="has no time"):
        F.read_function_from_file(tmp_path / "restart.bp", "H", 0.25)
    with pytest.raises(ValueError, match="no function named"):
        F.read_function_from_file(tmp_path / "restart.bp", "D", 0.2)
    with pytest.raises(NotImplementedError):
        F.read_function_from_file(tmp_path / "restart.bp", "H", 0.2, order=2)


def test_read_function_onto_another_mesh(tmp_path):
    """reference test/test_initial_condition.py:144-190: a P1 field written to a file comes back as
    the initial condition of a problem on its own mesh object."""
    from festim_b200 import fem

    src = F.create_unit_square(6, 6)
    u_ref = fem.Function(fem.functionspace(src, ("P", 1)), name="my_function")
    u_ref.interpolate(lambda x: x[1] ** 2 + 2 * x[0] ** 2)
    with field_io.XDMFFile(tmp_path / "initial_condition.xdmf", "w") as f:
        f.write_mesh(src)
        f.write_function(u_ref, 0.2)
    dst = F.create_unit_square(6, 6)
    p = F.HydrogenTransportProblem()
    Hs = F.Species("H")
    p.species = [Hs]
    p.mesh = F.Mesh(dst)
    vol = F.VolumeSubdomain(id=1, material=F.Material(D_0=1, E_D=0.1, name="dummy_mat"))
    p.subdomains = [vol]
    value = F.read_function_from_file(tmp_path / "initial_condition.xdmf", name="my_function", timestamp=0.2, mesh=dst)
    p.initial_conditions = [F.InitialConcentration(value=value, species=Hs, volume=vol)]
    p.temperature = 300
    p.settings = F.Settings(atol=1e-10, rtol=1e-10, final_time=1)
    p.settings.stepsize = F.Stepsize(0.1)
    p.solver_backend = OracleBackend()
    p.initialise()
    np.testing.assert_allclose(u_ref.x.array, p.u_n.x.array, atol=1e-14)


def test_xdmf_index_of_a_long_series_is_complete_and_cheap(tmp_path):
    """The XML index is rewritten as a whole, so it is flushed geometrically (plus at close): 400
    time levels cost a few dozen rewrites, and the closed file lists every one of them."""
    mesh = F.create_unit_square(3, 3)
    w = field_io.XDMFFile(tmp_path / "long.xdmf", "w")
    w.write_mesh(mesh)
    for k in range(400):
        w.write_function(("c", np.full(mesh.num_vertices, float(k))), 0.1 * k)
    before_close = field_io.XDMFFile(tmp_path / "long.xdmf", "r").function_times("c")
    assert len(before_close) >= 360            # never more than a tenth behind
    w.close()
    assert w.n_flushes < 70
    rd = field_io.XDMFFile(tmp_path / "long.xdmf", "r")
    assert np.allclose(rd.function_times("c"), 0.1 * np.arange(400))
    assert np.array_equal(rd.read_function("c", 0.1 * 399), np.full(mesh.num_vertices, 399.0))
    assert np.array_equal(rd.read_function("c", 0.1 * 17), np.full(mesh.num_vertices, 17.0))


def test_vtu_series_index_is_complete_after_close(tmp_path):
    import xml.etree.ElementTree as ET

    mesh = F.create_unit_square(2, 2)
    w = field_io.VTUSeriesWriter(tmp_path / "series.bp", mesh)
    for k in range(120):
        w.write(0.5 * k, {"c": np.full(mesh.num_vertices, float(k))})
    w.close()
    sets = ET.parse(tmp_path / "series.bp" / "series.pvd").getroot().findall("Collection/DataSet")
    assert len(sets) == 120 and float(sets[-1].get("timestep")) == 0.5 * 119
    assert sets[-1].get("file") == "step_000119.vtu" and (tmp_path / "series.bp" / "step_000119.vtu").exists()


def test_reads_meshio_style_xdmf(tmp_path):
    """The layout meshio writes with data_format="XML" (the usual gmsh -> meshio -> FESTIM route):
    `DataType` / `Precision` attributes, a 2D mesh stored as XYZ with z = 0, tags as a cell-centred
    attribute with its own name, line elements for the boundary in a second file."""
    (tmp_path / "mesh.xdmf").write_text("""<Xdmf Version="3.0"><Domain><Grid Name="Grid">
<Geometry GeometryType="XYZ"><DataItem DataType="Float" Dimensions="4 3" Format="XML" Precision="8">
0 0 0
1 0 0
1 1 0
0 1 0
</DataItem></Geometry>
<Topology NodesPerElement="3" NumberOfElements="2" TopologyType="Triangle">
<DataItem DataType="Int" Dimensions="2 3" Format="XML" Precision="8">
0 1 2
0 2 3
</DataItem></Topology>
<Attribute AttributeType="Scalar" Center="Cell" Name="f">
<DataItem DataType="Int" Dimensions="2" Format="XML" Precision="8">
6
7
</DataItem></Attribute></Grid></Domain></Xdmf>""")
    (tmp_path / "mf.xdmf").write_text("""<Xdmf Version="3.0"><Domain><Grid Name="Grid">
<Geometry GeometryType="XYZ"><DataItem DataType="Float" Dimensions="4 3" Format="XML" Precision="8">
0 0 0
1 0 0
1 1 0
0 1 0
</DataItem></Geometry>
<Topology NodesPerElement="2" NumberOfElements="3" TopologyType="Polyline">
<DataItem DataType="Int" Dimensions="3 2" Format="XML" Precision="8">
0 1
2 1
0 2
</DataItem></Topology>
<Attribute AttributeType="Scalar" Center="Cell" Name="f">
<DataItem DataType="Int" Dimensions="3" Format="XML" Precision="8">
11
12
99
</DataItem></Attribute></Grid></Domain></Xdmf>""")
    mesh = F.MeshFromXDMF(volume_file=tmp_path / "mesh.xdmf", facet_file=tmp_path / "mf.xdmf")
    assert mesh.mesh.tdim == 2 and mesh.mesh.gdim == 2 and mesh.mesh.num_cells == 2
    assert list(mesh.define_volume_meshtags().values) == [6, 7]
    ft = mesh.define_surface_meshtags()
    tags = {tuple(sorted(mesh.mesh.bfacets[i])): int(v) for i, v in zip(ft.indices, ft.values)}
    # (0, 2) is the interior diagonal: its tag 99 is dropped, untagged boundary edges get 0
    assert tags == {(0, 1): 11, (1, 2): 12, (2, 3): 0, (0, 3): 0}
