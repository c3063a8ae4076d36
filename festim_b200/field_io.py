"""Field output and mesh input without HDF5 / ADIOS2.

The reference writes fields with ``dolfinx.io.XDMFFile`` (XDMF + HDF5, ``exports/xdmf.py:61-78``)
and ``dolfinx.io.VTXWriter`` (ADIOS2 BP5, ``exports/vtx.py``, ``hydrogen_transport_problem.py:
442-483``) and reads meshes with ``XDMFFile.read_mesh / read_meshtags`` (``mesh/
mesh_from_xdmf.py:45-75``).  Neither h5py nor adios2 exists in this image, so:

* ``XDMFFile`` writes the same XDMF 3 document dolfinx writes (one ``Uniform`` mesh grid, one
  ``Temporal`` collection per function, cell-centred ``Attribute`` for mesh tags), with the heavy
  data in a raw little-endian side-car file (``Format="Binary"`` + ``Seek``, read by ParaView's
  Xdmf3 reader) or inline (``Format="XML"``).  It also reads such files, and files whose heavy data is
  ``Format="HDF"`` when h5py is importable (e.g. ``meshio`` output; ``meshio`` can also write
  ``data_format="XML"``).
* ``VTUSeriesWriter`` stands in for the VTX writer: the ``.bp`` path becomes a directory (as with
  ADIOS2) holding one VTK XML unstructured grid per export time and a ``series.pvd`` index.

Host-side file I/O only: nothing here is on the solver path.
"""
import base64
import os
import xml.etree.ElementTree as ET
from pathlib import Path

import numpy as np

_XDMF_TOPOLOGY = {1: ("PolyLine", 2), 2: ("Triangle", 3), 3: ("Tetrahedron", 4)}
_XDMF_NODES = {"polyline": 2, "polyline_2": 2, "edge": 2, "triangle": 3, "tetrahedron": 4, "tet": 4,
               "polyvertex": 1}
_VTK_CELL = {1: 3, 2: 5, 3: 10}   # VTK_LINE, VTK_TRIANGLE, VTK_TETRA
_XI = "https://www.w3.org/2001/XInclude"


def is_io_rank():
    """Partitioned runs keep the gathered fields on every rank; rank 0 alone writes files."""
    import sys

    torch = sys.modules.get("torch")
    if torch is None:
        return True
    dist = torch.distributed
    return not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0


class NullWriter:
    """Writer of the ranks that do not write."""

    def write_mesh(self, *a, **k):
        pass

    write_meshtags = write_function = write = close = write_mesh


def _mesh_of(obj):
    return getattr(obj, "mesh", obj) if not hasattr(obj, "coords") else obj


# =========================================================================================
# XDMF
# =========================================================================================
class XDMFFile:
    """Subset of ``dolfinx.io.XDMFFile``: ``write_mesh``, ``write_meshtags``, ``write_function``,
    ``read_mesh``, ``read_meshtags``, ``close``.  A leading communicator argument is accepted and
    ignored.  ``heavy`` selects where arrays go: ``"binary"`` (side-car ``<stem>.bin``) or ``"xml"``."""

    def __init__(self, *args, heavy="binary"):
        args = [a for a in args if isinstance(a, (str, Path))]
        if not args:
            raise TypeError("XDMFFile(filename, mode)")
        self.filename = Path(args[0])
        self.mode = args[1] if len(args) > 1 else "r"
        if self.mode not in ("r", "w"):
            raise ValueError("mode must be 'r' or 'w'")
        if heavy not in ("binary", "xml"):
            raise ValueError("heavy must be 'binary' or 'xml'")
        self.heavy = heavy
        self._collections = {}
        self._nbytes = 0
        self._closed = False
        self._writes = self._next_flush = self.n_flushes = 0
        if self.mode == "w":
            import atexit
            import weakref

            ref = weakref.ref(self)
            atexit.register(lambda: ref() is not None and ref().close())
            ET.register_namespace("xi", _XI)
            self._root = ET.Element("Xdmf", {"Version": "3.0"})
            self._domain = ET.SubElement(self._root, "Domain")
            self._bin = self.filename.with_suffix(".bin")
            if self.heavy == "binary":
                self._bin.parent.mkdir(parents=True, exist_ok=True)
                open(self._bin, "wb").close()
        else:
            if not self.filename.exists():
                raise FileNotFoundError(str(self.filename))
            self._root = ET.parse(self.filename).getroot()
            self._domain = self._root.find("Domain")
            if self._domain is None:
                raise ValueError(f"{self.filename}: no <Domain>")

    # ---- heavy data -------------------------------------------------------------------
    def _data_item(self, parent, array):
        array = np.ascontiguousarray(array)
        if array.dtype.kind in "iu":
            array = array.astype("<i8" if array.dtype.itemsize > 4 else "<i4")
            ntype = "Int"
        else:
            array = array.astype("<f8")
            ntype = "Float"
        dims = " ".join(str(n) for n in array.shape)
        attrib = {"Dimensions": dims, "NumberType": ntype, "Precision": str(array.dtype.itemsize)}
        if self.heavy == "binary":
            attrib.update({"Format": "Binary", "Endian": "Little", "Seek": str(self._nbytes)})
            item = ET.SubElement(parent, "DataItem", attrib)
            item.text = self._bin.name
            with open(self._bin, "ab") as f:
                f.write(array.tobytes())
            self._nbytes += array.nbytes
        else:
            attrib["Format"] = "XML"
            item = ET.SubElement(parent, "DataItem", attrib)
            rows = array.reshape(array.shape[0], -1) if array.ndim > 1 else array.reshape(-1, 1)
            fmt = "%d" if ntype == "Int" else "%.17g"
            item.text = "\n" + "\n".join(" ".join(fmt % v for v in row) for row in rows) + "\n"
        return item

    def _read_item(self, item):
        dims = tuple(int(n) for n in item.get("Dimensions").split())
        ntype = item.get("NumberType", item.get("DataType", "Float"))
        prec = int(item.get("Precision", "4" if ntype in ("Int", "UInt") else "8"))
        kind = {"Int": "i", "UInt": "u", "Float": "f"}.get(ntype)
        if kind is None:
            raise ValueError(f"{self.filename}: DataItem of type {ntype}")
        fmt = item.get("Format", "XML")
        text = (item.text or "").strip()
        if fmt == "XML":
            flat = np.array(text.split(), dtype=np.float64 if kind == "f" else np.int64)
            return flat.reshape(dims)
        if fmt == "Binary":
            end = "<" if item.get("Endian", "Little") in ("Little", "Native") else ">"
            dtype = np.dtype(f"{end}{kind}{prec}")
            count = int(np.prod(dims))
            path = self.filename.parent / text
            return np.fromfile(path, dtype=dtype, count=count, offset=int(item.get("Seek", "0"))).reshape(dims)
        if fmt == "HDF":
            try:
                import h5py
            except ImportError as e:   # pragma: no cover - h5py is absent from this image
                raise ImportError(
                    f"{self.filename} keeps its arrays in HDF5 and h5py is not installed; convert the "
                    "mesh with data_format='XML' or 'Binary' (meshio) or install h5py") from e
            h5name, _, dataset = text.partition(":")   # pragma: no cover
            with h5py.File(self.filename.parent / h5name, "r") as h5:   # pragma: no cover
                return np.asarray(h5[dataset]).reshape(dims)
        raise ValueError(f"{self.filename}: DataItem Format={fmt}")

    # ---- writing ------------------------------------------------------------------------
    def _require(self, mode):
        if self.mode != mode or self._closed:
            raise RuntimeError(f"XDMFFile {self.filename} is not open for {'writing' if mode == 'w' else 'reading'}")

    def _write_topology(self, grid, tdim, connectivity):
        name, nodes = _XDMF_TOPOLOGY[tdim] if tdim > 0 else ("PolyVertex", 1)
        topo = ET.SubElement(grid, "Topology", {"TopologyType": name,
                                                "NumberOfElements": str(connectivity.shape[0]),
                                                "NodesPerElement": str(nodes)})
        self._data_item(topo, connectivity.astype(np.int64))

    def _write_geometry(self, grid, mesh):
        gdim = max(mesh.gdim, 2)   # XDMF has no 1D geometry: pad y = 0, as dolfinx does
        xyz = np.zeros((mesh.num_vertices, gdim))
        xyz[:, : mesh.gdim] = mesh.coords
        geo = ET.SubElement(grid, "Geometry", {"GeometryType": "XY" if gdim == 2 else "XYZ"})
        self._data_item(geo, xyz)

    def write_mesh(self, mesh, name="mesh"):
        self._require("w")
        mesh = _mesh_of(mesh)
        grid = ET.SubElement(self._domain, "Grid", {"Name": name, "GridType": "Uniform"})
        self._write_topology(grid, mesh.tdim, mesh.cells)
        self._write_geometry(grid, mesh)
        self._mesh = mesh
        self._flush()

    def write_meshtags(self, tags, name="mesh_tags"):
        """Cell tags (``tags.dim == tdim``) or exterior-facet tags as a grid with its own topology
        and a cell-centred attribute of the same name (dolfinx layout)."""
        self._require("w")
        mesh = tags.mesh
        if tags.dim == mesh.tdim:
            conn = mesh.cells[tags.indices]
        elif tags.dim == mesh.tdim - 1:
            conn = mesh.bfacets[tags.indices]
        else:
            raise NotImplementedError("mesh tags on cells or exterior facets only")
        grid = ET.SubElement(self._domain, "Grid", {"Name": name, "GridType": "Uniform"})
        self._write_topology(grid, tags.dim, conn)
        self._write_geometry(grid, mesh)
        att = ET.SubElement(grid, "Attribute", {"Name": name, "AttributeType": "Scalar", "Center": "Cell"})
        self._data_item(att, np.asarray(tags.values, dtype=np.int32).reshape(-1, 1))
        self._flush()

    def write_function(self, function, t=0.0, name=None):
        """One time level of a P1 nodal field (``fem.Function`` or ``(name, values)``)."""
        self._require("w")
        if isinstance(function, tuple):
            name, values = function
        else:
            name = name or function.name or "f"
            values = function.x.array
        if not hasattr(self, "_mesh"):
            raise RuntimeError("write_mesh must be called before write_function")
        values = np.asarray(values, dtype=np.float64).reshape(self._mesh.num_vertices, -1)
        coll = self._collections.get(name)
        if coll is None:
            coll = ET.SubElement(self._domain, "Grid", {"Name": name, "GridType": "Collection",
                                                       "CollectionType": "Temporal"})
            self._collections[name] = coll
        grid = ET.SubElement(coll, "Grid", {"Name": name, "GridType": "Uniform"})
        ET.SubElement(grid, f"{{{_XI}}}include", {
            "xpointer": "xpointer(/Xdmf/Domain/Grid[@GridType='Uniform'][1]/*[self::Topology or self::Geometry])"})
        ET.SubElement(grid, "Time", {"Value": repr(float(t))})
        att = ET.SubElement(grid, "Attribute", {"Name": name, "AttributeType": "Scalar", "Center": "Node"})
        self._data_item(att, values)
        # the index is rewritten as a whole: do it after each of the first writes, then whenever the
        # number of writes has grown by a tenth, and at close / interpreter exit - a long transient
        # costs O(n log n) instead of O(n^2), the side-car data is on disk at once either way
        self._writes += 1
        if self._writes <= 16 or self._writes >= self._next_flush:
            self._flush()
            self._next_flush = self._writes + max(1, self._writes // 10)

    def _flush(self):
        self.n_flushes += 1
        ET.indent(self._root)
        self.filename.parent.mkdir(parents=True, exist_ok=True)
        with open(self.filename, "wb") as f:
            f.write(b'<?xml version="1.0"?>\n<!DOCTYPE Xdmf SYSTEM "Xdmf.dtd" []>\n')
            f.write(ET.tostring(self._root))
            f.write(b"\n")

    def close(self):
        if self.mode == "w" and not self._closed:
            self._flush()
        self._closed = True

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- reading ------------------------------------------------------------------------
    def _grid(self, name):
        for grid in self._domain.iter("Grid"):
            if grid.get("Name") == name and grid.get("GridType", "Uniform") == "Uniform":
                return grid
        names = [g.get("Name") for g in self._domain.iter("Grid")]
        raise ValueError(f"{self.filename}: no grid named {name!r} (found {names})")

    def _topology(self, grid):
        topo = grid.find("Topology")
        if topo is None:
            raise ValueError(f"{self.filename}: grid {grid.get('Name')!r} has no <Topology>")
        kind = (topo.get("TopologyType") or topo.get("Type") or "").lower()
        if kind not in _XDMF_NODES:
            raise NotImplementedError(f"{self.filename}: topology {kind!r} (simplices of degree 1 only)")
        conn = self._read_item(topo.find("DataItem")).astype(np.int64)
        return conn.reshape(-1, _XDMF_NODES[kind])

    def _geometry(self, grid):
        geo = grid.find("Geometry")
        if geo is None:   # dolfinx mesh-tag grids include the mesh geometry by xpointer
            geo = next(self._domain.iter("Geometry"), None)
        if geo is None:
            raise ValueError(f"{self.filename}: no <Geometry>")
        return np.asarray(self._read_item(geo.find("DataItem")), dtype=np.float64)

    def read_mesh(self, name="Grid"):
        """``SimplexMesh`` of the named grid.  Unused trailing coordinate columns (a 2D mesh stored
        as XYZ with z = 0, a 1D mesh with y = 0) are dropped so that gdim == tdim."""
        from festim_b200.mesh import SimplexMesh

        self._require("r")
        grid = self._grid(name)
        cells = self._topology(grid)
        xyz = self._geometry(grid)
        tdim = cells.shape[1] - 1
        if tdim < 1:
            raise ValueError(f"{self.filename}: grid {name!r} holds vertices, not cells")
        if xyz.shape[1] < tdim:
            raise ValueError(f"{self.filename}: {tdim}D cells in {xyz.shape[1]}D geometry")
        if np.any(xyz[:, tdim:] != xyz[:1, tdim:]):
            raise ValueError(f"{self.filename}: manifold meshes (tdim < gdim) are not supported")
        # points no cell uses (meshio keeps those of lower-dimensional blocks) would be unknowns
        # without equations: drop them and remember the numbering of the file for read_meshtags
        used = np.zeros(xyz.shape[0], dtype=bool)
        used[cells.ravel()] = True
        file_to_mesh = np.full(xyz.shape[0], -1, dtype=np.int64)
        file_to_mesh[used] = np.arange(int(used.sum()))
        mesh = SimplexMesh(xyz[used][:, :tdim], file_to_mesh[cells].astype(np.int32))
        mesh.file_vertex_map = file_to_mesh
        return mesh

    def function_times(self, name):
        """Time values stored for the named function."""
        self._require("r")
        return [float(g.find("Time").get("Value")) for g in self._function_grids(name)]

    def _function_grids(self, name):
        for coll in self._domain.iter("Grid"):
            if coll.get("Name") == name and coll.get("GridType") == "Collection":
                return coll.findall("Grid")
        names = [g.get("Name") for g in self._domain.iter("Grid") if g.get("GridType") == "Collection"]
        raise ValueError(f"{self.filename}: no function named {name!r} (found {names})")

    def read_function(self, name, t=None):
        """Nodal values of the named function at time ``t`` (the last one stored when None)."""
        self._require("r")
        grids = self._function_grids(name)
        if t is None:
            grid = grids[-1]
        else:
            times = np.array([float(g.find("Time").get("Value")) for g in grids])
            hit = np.nonzero(np.isclose(times, float(t), rtol=1e-9, atol=1e-12))[0]
            if len(hit) == 0:
                raise ValueError(f"{self.filename}: function {name!r} has no time {t} (stored: {times.tolist()})")
            grid = grids[int(hit[-1])]
        return np.asarray(self._read_item(grid.find("Attribute/DataItem")), dtype=np.float64).reshape(-1)

    def read_meshtags(self, mesh, name="Grid"):
        """Tags of the named grid mapped onto ``mesh``: entities are matched by their vertex sets.
        Facet tags are kept for exterior facets (the continuous problem integrates over no others);
        tagged entities that match nothing raise."""
        from festim_b200.mesh import MeshTags

        self._require("r")
        mesh = _mesh_of(mesh)
        grid = self._grid(name)
        conn = self._topology(grid)
        att = grid.find(f"Attribute[@Name='{name}']")
        if att is None:
            att = grid.find("Attribute")
        if att is None:
            raise ValueError(f"{self.filename}: grid {name!r} has no <Attribute> holding the tags")
        values = np.asarray(self._read_item(att.find("DataItem"))).reshape(-1).astype(np.int64)
        if len(values) != conn.shape[0]:
            raise ValueError(f"{self.filename}: {len(values)} tag values for {conn.shape[0]} entities")
        fmap = getattr(mesh, "file_vertex_map", None)
        if fmap is not None:
            if conn.size and conn.max() >= len(fmap):
                raise ValueError(f"{self.filename}: vertex numbers beyond those of the mesh file")
            conn = fmap[conn]
            keep = (conn >= 0).all(axis=1)   # entities on points outside every cell
            conn, values = conn[keep], values[keep]
        dim = conn.shape[1] - 1
        if dim == mesh.tdim:
            target = mesh.cells
        elif dim == mesh.tdim - 1:
            target = mesh.bfacets
        else:
            raise ValueError(f"{self.filename}: tags on entities of dimension {dim} for a {mesh.tdim}D mesh")
        key_t = _entity_keys(target, mesh.num_vertices)
        key_f = _entity_keys(conn, mesh.num_vertices)
        order = np.argsort(key_t, kind="stable")
        pos = np.searchsorted(key_t[order], key_f)
        pos = np.minimum(pos, len(order) - 1)
        found = key_t[order][pos] == key_f
        if dim == mesh.tdim and not found.all():
            raise ValueError(f"{self.filename}: {int((~found).sum())} tagged cells are not cells of the mesh")
        idx = order[pos[found]].astype(np.int32)
        val = values[found].astype(np.int32)
        sort = np.argsort(idx, kind="stable")
        return MeshTags(mesh, dim, idx[sort], val[sort])


def _entity_keys(conn, nv):
    """One sortable row per entity (sorted vertex tuple packed into a structured scalar)."""
    c = np.sort(np.asarray(conn, dtype=np.int64), axis=1)
    c = np.ascontiguousarray(c)
    return c.view(np.dtype((np.void, c.dtype.itemsize * c.shape[1]))).ravel()


# =========================================================================================
# VTK XML unstructured grids (stand-in for the ADIOS2 VTX writer)
# =========================================================================================
def _vtk_array(parent, name, array, ncomp=None):
    array = np.ascontiguousarray(array)
    vtk_type = {"f8": "Float64", "i8": "Int64", "i4": "Int32", "u1": "UInt8"}[array.dtype.str[1:]]
    attrib = {"type": vtk_type, "Name": name, "format": "binary"}
    if ncomp is not None:
        attrib["NumberOfComponents"] = str(ncomp)
    el = ET.SubElement(parent, "DataArray", attrib)
    raw = array.astype(array.dtype.newbyteorder("<")).tobytes()
    el.text = base64.b64encode(np.uint64(len(raw)).tobytes() + raw).decode("ascii")
    return el


def write_vtu(path, mesh, point_data, t=None):
    """One ``.vtu`` file: the P1 mesh and nodal fields ``{name: array[num_vertices]}``."""
    mesh = _mesh_of(mesh)
    nv, nc, d1 = mesh.num_vertices, mesh.num_cells, mesh.tdim + 1
    root = ET.Element("VTKFile", {"type": "UnstructuredGrid", "version": "1.0",
                                  "byte_order": "LittleEndian", "header_type": "UInt64"})
    ug = ET.SubElement(root, "UnstructuredGrid")
    if t is not None:
        fd = ET.SubElement(ug, "FieldData")
        el = _vtk_array(fd, "TimeValue", np.array([float(t)]))
        el.set("NumberOfTuples", "1")
    piece = ET.SubElement(ug, "Piece", {"NumberOfPoints": str(nv), "NumberOfCells": str(nc)})
    pd = ET.SubElement(piece, "PointData", {"Scalars": next(iter(point_data), "")})
    for name, values in point_data.items():
        values = np.asarray(values, dtype=np.float64).reshape(-1)
        if values.shape[0] != nv:
            raise ValueError(f"field {name!r}: {values.shape[0]} values for {nv} vertices")
        _vtk_array(pd, name, values)
    pts = ET.SubElement(piece, "Points")
    xyz = np.zeros((nv, 3))
    xyz[:, : mesh.gdim] = mesh.coords
    _vtk_array(pts, "Points", xyz, ncomp=3)
    cells = ET.SubElement(piece, "Cells")
    _vtk_array(cells, "connectivity", mesh.cells.astype(np.int64).reshape(-1))
    _vtk_array(cells, "offsets", np.arange(d1, d1 * nc + 1, d1, dtype=np.int64))
    _vtk_array(cells, "types", np.full(nc, _VTK_CELL[mesh.tdim], dtype=np.uint8))
    with open(path, "wb") as f:
        f.write(b'<?xml version="1.0"?>\n')
        f.write(ET.tostring(root))
        f.write(b"\n")


def read_vtu(path):
    """Inverse of ``write_vtu`` (for tests and restarts): ``(coords, cells, point_data, t)``."""
    root = ET.parse(path).getroot()

    def arr(el):
        dtype = {"Float64": "<f8", "Int64": "<i8", "Int32": "<i4", "UInt8": "u1"}[el.get("type")]
        raw = base64.b64decode(el.text.strip())
        n = int(np.frombuffer(raw[:8], dtype="<u8")[0])
        return np.frombuffer(raw[8: 8 + n], dtype=dtype)

    piece = root.find("UnstructuredGrid/Piece")
    xyz = arr(piece.find("Points/DataArray")).reshape(-1, 3)
    conn = arr(piece.find("Cells/DataArray[@Name='connectivity']"))
    offs = arr(piece.find("Cells/DataArray[@Name='offsets']"))
    cells = conn.reshape(len(offs), -1)
    data = {el.get("Name"): arr(el).copy() for el in piece.find("PointData")}
    tv = root.find("UnstructuredGrid/FieldData/DataArray[@Name='TimeValue']")
    t = float(arr(tv)[0]) if tv is not None else None
    return xyz, cells, data, t


class VTUSeriesWriter:
    """Time series at ``path`` (a directory, like an ADIOS2 ``.bp``): ``step_000000.vtu`` ... and
    ``series.pvd``, which ParaView opens as one animated data set."""

    def __init__(self, path, mesh):
        self.path = Path(path)
        self.mesh = _mesh_of(mesh)
        self.path.mkdir(parents=True, exist_ok=True)
        for old in self.path.glob("step_*.vtu"):
            os.remove(old)
        self.times = []
        self._next_index = 0
        self._write_index()
        import atexit
        import weakref

        ref = weakref.ref(self)
        atexit.register(lambda: ref() is not None and ref().close())

    def write(self, t, point_data):
        name = f"step_{len(self.times):06d}.vtu"
        write_vtu(self.path / name, self.mesh, point_data, t=t)
        self.times.append((float(t), name))
        n = len(self.times)
        if n <= 16 or n >= self._next_index:     # same schedule as the XDMF index
            self._write_index()
            self._next_index = n + max(1, n // 10)

    def _write_index(self):
        self._indexed = len(self.times)
        root = ET.Element("VTKFile", {"type": "Collection", "version": "0.1", "byte_order": "LittleEndian"})
        coll = ET.SubElement(root, "Collection")
        for t, name in self.times:
            ET.SubElement(coll, "DataSet", {"timestep": repr(t), "group": "", "part": "0", "file": name})
        ET.indent(root)
        with open(self.path / "series.pvd", "wb") as f:
            f.write(b'<?xml version="1.0"?>\n')
            f.write(ET.tostring(root))
            f.write(b"\n")

    def close(self):
        if getattr(self, "_indexed", 0) != len(self.times):
            self._write_index()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
