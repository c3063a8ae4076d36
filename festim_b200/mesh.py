"""Simplex meshes, tags and the FESTIM mesh wrappers.

Mirrors reference ``src/festim/mesh/mesh.py:43-219`` (``Mesh``,
``define_meshtags``) and ``src/festim/mesh/mesh_1d.py:60-151`` (``Mesh1D``).
``SimplexMesh`` replaces ``dolfinx.mesh.Mesh`` as the input format of the
path: vertex coordinates, cell connectivity and (derived once) exterior
facets, the vertex graph and the vertex->cell adjacency the gather kernels
walk.  The builders reproduce the dolfinx layouts described in SURVEY A.10
(x-fastest vertices, 6 Kuhn tetrahedra per hexahedron sharing the v0-v7
diagonal, "right" diagonal for squares).
"""

from __future__ import annotations

from enum import Enum

import numpy as np
import scipy.sparse as sp


class SimplexMesh:
    """P1 interval / triangle / tetrahedron mesh (tdim == gdim)."""

    def __init__(self, coords, cells):
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        if coords.ndim == 1:
            coords = coords[:, None]
        self.coords = coords
        self.cells = np.ascontiguousarray(cells, dtype=np.int32)
        self.gdim = coords.shape[1]
        self.tdim = self.cells.shape[1] - 1
        if self.gdim != self.tdim:
            raise ValueError("only meshes with tdim == gdim are supported")
        self._bfacets = None
        self._bfacet_cell = None
        self._graph = None
        self._v2c = None

    # sizes ------------------------------------------------------------------
    @property
    def num_vertices(self):
        return self.coords.shape[0]

    @property
    def num_cells(self):
        return self.cells.shape[0]

    @property
    def dim(self):
        return self.tdim

    @property
    def geometry(self):
        """``mesh.geometry.x`` (coordinates padded to three columns) and ``.dim``, as user
        scripts written against dolfinx read them."""
        import types

        x = np.zeros((self.num_vertices, 3))
        x[:, : self.gdim] = self.coords
        return types.SimpleNamespace(x=x, dim=self.gdim)

    # exterior facets ----------------------------------------------------------
    def _build_facets(self):
        d = self.tdim
        nc = self.num_cells
        if d == 1:
            verts = self.cells.ravel()
            counts = np.bincount(verts, minlength=self.num_vertices)
            bverts = np.nonzero(counts == 1)[0].astype(np.int32)
            # owning cell of each boundary vertex
            order = np.argsort(verts, kind="stable")
            first = np.searchsorted(verts[order], bverts)
            self._bfacets = bverts[:, None]
            self._bfacet_cell = (order[first] // 2).astype(np.int32)
            return
        # all facets: drop local vertex k of each cell
        loc = [[j for j in range(d + 1) if j != k] for k in range(d + 1)]
        fac = np.concatenate([self.cells[:, l] for l in loc], axis=0).astype(np.int64)
        fac.sort(axis=1)
        nv = np.int64(self.num_vertices)
        if d == 2:
            key = fac[:, 0] * nv + fac[:, 1]
        else:
            if nv >= (1 << 21):
                raise ValueError("3D meshes above 2^21 vertices need a wider facet key")
            key = (fac[:, 0] << 42) | (fac[:, 1] << 21) | fac[:, 2]
        order = np.argsort(key, kind="stable")
        ks = key[order]
        first = np.ones(len(ks), dtype=bool)
        first[1:] = ks[1:] != ks[:-1]
        last = np.ones(len(ks), dtype=bool)
        last[:-1] = ks[1:] != ks[:-1]
        single = order[first & last]
        single.sort()
        self._bfacets = fac[single].astype(np.int32)
        self._bfacet_cell = (single % nc).astype(np.int32)

    @property
    def bfacets(self):
        """Vertices of the exterior facets, shape (N_f, tdim)."""
        if self._bfacets is None:
            self._build_facets()
        return self._bfacets

    @property
    def bfacet_cell(self):
        if self._bfacet_cell is None:
            self._build_facets()
        return self._bfacet_cell

    # vertex graph (sparsity pattern of any P1 operator) ----------------------
    @property
    def graph(self):
        """CSR (rowptr int32, colidx int32, sorted, diagonal included)."""
        if self._graph is None:
            d1 = self.tdim + 1
            rows = np.repeat(self.cells, d1, axis=1).ravel()
            cols = np.tile(self.cells, (1, d1)).ravel()
            n = self.num_vertices
            A = sp.coo_matrix((np.ones(len(rows), dtype=np.int8), (rows, cols)), shape=(n, n)).tocsr()
            A.sum_duplicates()
            A.sort_indices()
            self._graph = (A.indptr.astype(np.int32), A.indices.astype(np.int32))
        return self._graph

    @property
    def vertex_to_cell(self):
        """CSR (ptr int32, packed int32 = cell * (tdim+1) + local index)."""
        if self._v2c is None:
            verts = self.cells.ravel().astype(np.int64)
            order = np.argsort(verts, kind="stable")  # flat index = cell*(d+1)+local
            ptr = np.zeros(self.num_vertices + 1, dtype=np.int64)
            ptr[1:] = np.bincount(verts, minlength=self.num_vertices)
            ptr = np.cumsum(ptr)
            assert len(verts) < 2**31
            self._v2c = (ptr.astype(np.int32), order.astype(np.int32))
        return self._v2c

    # geometry helpers -----------------------------------------------------------
    def cell_volumes(self):
        x = self.coords[self.cells]
        J = x[:, 1:, :] - x[:, :1, :]
        fact = {1: 1.0, 2: 2.0, 3: 6.0}[self.tdim]
        return np.abs(np.linalg.det(J)) / fact

    def x3(self, verts=None):
        """Coordinates in the dolfinx locator convention: shape (3, n)."""
        c = self.coords if verts is None else self.coords[verts]
        out = np.zeros((3, c.shape[0]))
        out[: self.gdim] = c.T
        return out


# --------------------------------------------------------------------------
# builders (dolfinx.mesh.create_* analogues; an optional leading communicator
# argument is accepted and ignored so reference scripts port unchanged)
# --------------------------------------------------------------------------
def _strip_comm(args):
    if args and not isinstance(args[0], (int, np.integer, list, tuple, np.ndarray)):
        return args[1:]
    return args


def create_interval(*args):
    n, (a, b) = _strip_comm(args)
    x = np.linspace(a, b, n + 1)
    cells = np.stack([np.arange(n), np.arange(1, n + 1)], axis=1)
    return SimplexMesh(x[:, None], cells)


def create_unit_interval(*args):
    (n,) = _strip_comm(args)
    return create_interval(n, (0.0, 1.0))


def create_rectangle(*args):
    (p0, p1), (nx, ny) = _strip_comm(args)[:2]
    xs = np.linspace(p0[0], p1[0], nx + 1)
    ys = np.linspace(p0[1], p1[1], ny + 1)
    X, Y = np.meshgrid(xs, ys, indexing="xy")  # x fastest
    coords = np.stack([X.ravel(), Y.ravel()], axis=1)
    ix, iy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    v0 = (iy * (nx + 1) + ix).ravel()
    v1, v2, v3 = v0 + 1, v0 + nx + 1, v0 + nx + 2
    cells = np.empty((2 * nx * ny, 3), dtype=np.int64)
    cells[0::2] = np.stack([v0, v1, v3], axis=1)
    cells[1::2] = np.stack([v0, v2, v3], axis=1)
    return SimplexMesh(coords, cells)


def create_unit_square(*args):
    nx, ny = _strip_comm(args)[:2]
    return create_rectangle(((0.0, 0.0), (1.0, 1.0)), (nx, ny))


def create_box(*args):
    (p0, p1), (nx, ny, nz) = _strip_comm(args)[:2]
    xs = np.linspace(p0[0], p1[0], nx + 1)
    ys = np.linspace(p0[1], p1[1], ny + 1)
    zs = np.linspace(p0[2], p1[2], nz + 1)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")  # x fastest
    coords = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    iz, iy, ix = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    sx, sy, sz = 1, nx + 1, (nx + 1) * (ny + 1)
    v0 = (iz * sz + iy * sy + ix * sx).ravel().astype(np.int64)
    v1, v2, v3 = v0 + sx, v0 + sy, v0 + sx + sy
    v4, v5, v6, v7 = v0 + sz, v1 + sz, v2 + sz, v3 + sz
    tets = [(v0, v1, v3, v7), (v0, v1, v7, v5), (v0, v5, v7, v4),
            (v0, v3, v2, v7), (v0, v6, v4, v7), (v0, v2, v6, v7)]
    cells = np.empty((6 * len(v0), 4), dtype=np.int64)
    for k, t in enumerate(tets):
        cells[k::6] = np.stack(t, axis=1)
    return SimplexMesh(coords, cells)


def create_unit_cube(*args):
    nx, ny, nz = _strip_comm(args)[:3]
    return create_box(((0.0, 0.0, 0.0), (1.0, 1.0, 1.0)), (nx, ny, nz))


# --------------------------------------------------------------------------
# entity location and tags
# --------------------------------------------------------------------------
class MeshTags:
    """(indices, values) over cells (dim == tdim) or exterior facets."""

    def __init__(self, mesh, dim, indices, values):
        self.mesh = mesh
        self.dim = dim
        self.indices = np.asarray(indices, dtype=np.int32)
        self.values = np.asarray(values, dtype=np.int32)

    def find(self, value):
        return self.indices[self.values == value]


def meshtags(mesh, dim, indices, values):
    return MeshTags(mesh, dim, indices, values)


def _marker_on_vertices(mesh, marker):
    m = np.asarray(marker(mesh.x3()))
    if m.shape != (mesh.num_vertices,):
        m = np.broadcast_to(m, (mesh.num_vertices,))
    return m.astype(bool)


def locate_entities(mesh, dim, marker):
    """Entities whose vertices all satisfy ``marker`` (dolfinx semantics).
    For dim == tdim-1 only exterior facets are indexed."""
    m = _marker_on_vertices(mesh, marker)
    if dim == mesh.tdim:
        return np.nonzero(m[mesh.cells].all(axis=1))[0].astype(np.int32)
    if dim == mesh.tdim - 1:
        return np.nonzero(m[mesh.bfacets].all(axis=1))[0].astype(np.int32)
    if dim == 0:
        return np.nonzero(m)[0].astype(np.int32)
    raise NotImplementedError(f"entities of dimension {dim}")


def locate_entities_boundary(mesh, dim, marker):
    if dim != mesh.tdim - 1:
        raise NotImplementedError("only boundary facets can be located")
    return locate_entities(mesh, dim, marker)


def exterior_facet_indices(mesh):
    return np.arange(mesh.bfacets.shape[0], dtype=np.int32)


# --------------------------------------------------------------------------
# FESTIM wrappers
# --------------------------------------------------------------------------
class CoordinateSystem(Enum):
    CARTESIAN = 10
    CYLINDRICAL = 20
    SPHERICAL = 30

    @classmethod
    def from_string(cls, s):
        try:
            return cls[s.upper()]
        except KeyError:
            raise ValueError("Invalid CoordinateSystem value")

    def __str__(self):
        return self.name.lower()


class Mesh:
    """Mesh wrapper, reference ``src/festim/mesh/mesh.py:43-128``."""

    def __init__(self, mesh=None, coordinate_system="cartesian"):
        self.mesh = mesh
        self.coordinate_system = coordinate_system

    @property
    def mesh(self):
        return self._mesh

    @mesh.setter
    def mesh(self, value):
        if value is None or isinstance(value, SimplexMesh):
            self._mesh = value
        else:
            raise TypeError("Mesh must be of type festim_b200.mesh.SimplexMesh")

    @property
    def coordinate_system(self):
        return self._coordinate_system

    @coordinate_system.setter
    def coordinate_system(self, value):
        if isinstance(value, CoordinateSystem):
            self._coordinate_system = value
        elif isinstance(value, str):
            self._coordinate_system = CoordinateSystem.from_string(value)
        else:
            raise ValueError("coordinate_system must be of type str or CoordinateSystem")
        if self._mesh is not None:
            self.check_mesh_dim_coords()

    def check_mesh_dim_coords(self):
        """Reference ``mesh.py:221-233``."""
        if self.coordinate_system == CoordinateSystem.SPHERICAL and self.vdim != 1:
            raise AttributeError("spherical coordinates can be used for one-dimensional domains only")
        if self.coordinate_system == CoordinateSystem.CYLINDRICAL and self.vdim == 3:
            raise AttributeError("cylindrical coordinates cannot be used for 3D domains")

    @property
    def vdim(self):
        if self._mesh is None:
            raise RuntimeError("Mesh is not defined")
        return self._mesh.tdim

    @property
    def fdim(self):
        if self._mesh is None:
            raise RuntimeError("Mesh is not defined")
        return self._mesh.tdim - 1

    def define_meshtags(self, surface_subdomains, volume_subdomains, interfaces=None):
        """Reference ``mesh.py:130-219``: all cells/facets start at 0, surfaces then
        volumes overwrite in list order, every cell must end up non-zero."""
        m = self._mesh
        tags_volumes = np.zeros(m.num_cells, dtype=np.int32)
        tags_facets = np.zeros(m.bfacets.shape[0], dtype=np.int32)
        for surf in surface_subdomains:
            try:
                entities = surf.locate_boundary_facet_indices(m)
                tags_facets[entities] = surf.id
            except ValueError:
                if len(surface_subdomains) > 1:
                    raise ValueError(
                        "Surface subdomain must have a locator attribute if"
                        " several subdomains are defined"
                    )
                tags_facets[:] = surf.id
        for vol in volume_subdomains:
            try:
                entities = vol.locate_subdomain_entities(m)
                tags_volumes[entities] = vol.id
            except ValueError:
                if len(volume_subdomains) > 1:
                    raise ValueError(
                        "Volume subdomain must have a locator if"
                        " several subdomains are defined"
                    )
                tags_volumes[:] = vol.id
        assert np.all(tags_volumes != 0), (
            "All cells must be tagged with a non-zero value"
            "This error can be caused by a volume subdomain not being defined"
            "correctly, or by a mesh that is too coarse to capture the geometry"
            "of the volume subdomains"
        )
        if interfaces:
            raise NotImplementedError("interfaces belong to the discontinuous problem (SURVEY 8f1)")
        volume_meshtags = MeshTags(m, m.tdim, np.arange(m.num_cells), tags_volumes)
        facet_meshtags = MeshTags(m, m.tdim - 1, np.arange(len(tags_facets)), tags_facets)
        return facet_meshtags, volume_meshtags


def _is_nested(value):
    return len(value) > 0 and all(isinstance(v, (list, tuple, np.ndarray)) for v in value)


class Mesh1D(Mesh):
    """1D mesh from vertex positions, reference ``mesh_1d.py:60-151``."""

    def __init__(self, vertices, **kwargs):
        self.vertices = vertices
        super().__init__(self.generate_mesh(), **kwargs)

    @property
    def vertices(self):
        return self._vertices

    @vertices.setter
    def vertices(self, value):
        blocks = value if _is_nested(value) else [value]
        self._vertex_blocks = [np.sort(np.unique(b)).astype(np.float64) for b in blocks]
        for block in self._vertex_blocks:
            if block.size < 2:
                raise ValueError("Each block of vertices must have at least 2 vertices")
        self._vertex_blocks.sort(key=lambda block: block[0])
        for block, nxt in zip(self._vertex_blocks[:-1], self._vertex_blocks[1:]):
            if nxt[0] < block[-1]:
                raise ValueError("Blocks of vertices must not overlap")
        self._vertices = np.concatenate(self._vertex_blocks)

    @property
    def vertex_blocks(self):
        return self._vertex_blocks

    def generate_mesh(self):
        cells_per_block = []
        offset = 0
        for block in self.vertex_blocks:
            idx = np.arange(offset, offset + block.shape[0])
            cells_per_block.append(np.stack((idx[:-1], idx[1:]), axis=-1))
            offset += block.shape[0]
        return SimplexMesh(self.vertices[:, None], np.concatenate(cells_per_block))

    def check_borders(self, volume_subdomains):
        if len(volume_subdomains) == 0:
            raise ValueError("No volume subdomains defined")
        remaining = list(volume_subdomains)
        for block in self.vertex_blocks:
            in_block = [v for v in remaining
                        if block[0] <= min(v.borders) and max(v.borders) <= block[-1]]
            for v in in_block:
                remaining.remove(v)
            borders = np.sort([b for v in in_block for b in v.borders])
            if len(borders) == 0:
                raise ValueError("borders dont match domain borders")
            for start, end in zip(borders[1:-2:2], borders[2:-1:2]):
                if start != end:
                    raise ValueError("Subdomain borders don't match to each other")
            if borders[0] != block[0] or borders[-1] != block[-1]:
                raise ValueError("borders dont match domain borders")
        if remaining:
            raise ValueError("borders dont match domain borders")

    def define_meshtags(self, surface_subdomains, volume_subdomains, interfaces=None):
        self.check_borders(volume_subdomains)
        return super().define_meshtags(surface_subdomains, volume_subdomains, interfaces)


class MeshFromXDMF(Mesh):
    """Mesh and tags read from XDMF files, reference ``mesh/mesh_from_xdmf.py:8-75``.
    The files are read by ``festim_b200.field_io.XDMFFile`` (inline, raw binary or - with h5py -
    HDF5 heavy data).  Facet tags are kept on exterior facets."""

    def __init__(self, volume_file, facet_file, mesh_name="Grid", surface_meshtags_name="Grid",
                 volume_meshtags_name="Grid"):
        from festim_b200.field_io import XDMFFile

        self.volume_file = volume_file
        self.facet_file = facet_file
        self.mesh_name = mesh_name
        self.surface_meshtags_name = surface_meshtags_name
        self.volume_meshtags_name = volume_meshtags_name
        mesh = XDMFFile(self.volume_file, "r").read_mesh(name=f"{self.mesh_name}")
        super().__init__(mesh=mesh)

    def define_surface_meshtags(self):
        """Exterior facets the file does not tag get 0 (integrated over by no ``ds(id)``)."""
        from festim_b200.field_io import XDMFFile

        read = XDMFFile(self.facet_file, "r").read_meshtags(self.mesh, name=f"{self.surface_meshtags_name}")
        values = np.zeros(self.mesh.bfacets.shape[0], dtype=np.int32)
        values[read.indices] = read.values
        return MeshTags(self.mesh, self.mesh.tdim - 1, np.arange(len(values)), values)

    def define_volume_meshtags(self):
        from festim_b200.field_io import XDMFFile

        read = XDMFFile(self.volume_file, "r").read_meshtags(self.mesh, name=f"{self.volume_meshtags_name}")
        values = np.zeros(self.mesh.num_cells, dtype=np.int32)
        values[read.indices] = read.values
        return MeshTags(self.mesh, self.mesh.tdim, np.arange(len(values)), values)
