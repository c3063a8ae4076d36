"""Surface and volume subdomains (reference ``src/festim/subdomain/``).

Locators receive vertex coordinates in the dolfinx convention, an array of
shape (3, n); an entity belongs to the subdomain when the locator holds at all
of its vertices.
"""

from __future__ import annotations

import numpy as np

from festim_b200 import mesh as _mesh


class SurfaceSubdomain:
    """``surface_subdomain.py:6-57``."""

    def __init__(self, id, locator=None):
        self.id = id
        self.locator = locator

    def locate_boundary_facet_indices(self, mesh):
        if self.locator is None:
            raise ValueError("No locator function provided for locating boundary facets.")
        return _mesh.locate_entities_boundary(mesh, mesh.tdim - 1, self.locator)


class SurfaceSubdomain1D(SurfaceSubdomain):
    """``surface_subdomain.py:60-88``: the facet at ``x``."""

    def __init__(self, id, x):
        super().__init__(id, locator=lambda x_: np.isclose(x_[0], x))
        self.x = x


def find_surface_from_id(id, surfaces):
    for surf in surfaces:
        if surf.id == id:
            return surf
    raise ValueError(f"id {id} not found in list of surfaces")


class VolumeSubdomain:
    """``volume_subdomain.py:22-115`` (continuous-problem part)."""

    def __init__(self, id, material, locator=None, name=None):
        assert id != 0, "Volume subdomain id cannot be 0"
        self.id = id
        self.material = material
        self.locator = locator
        self.name = name

    @property
    def name(self):
        return self._name

    @name.setter
    def name(self, value):
        if value is None or isinstance(value, str):
            self._name = value
        else:
            raise TypeError("Name must be a string")

    def locate_subdomain_entities(self, mesh):
        if self.locator is None:
            raise ValueError("No locator function provided for locating cells.")
        return _mesh.locate_entities(mesh, mesh.tdim, self.locator)


class VolumeSubdomain1D(VolumeSubdomain):
    """``volume_subdomain.py:118-149``: cells between two borders."""

    def __init__(self, id, borders, material):
        super().__init__(
            id, material,
            locator=lambda x: np.logical_and(x[0] >= borders[0], x[0] <= borders[1]),
        )
        self.borders = borders


def find_volume_from_id(id, volumes):
    for vol in volumes:
        if vol.id == id:
            return vol
    raise ValueError(f"id {id} not found in list of volumes")


def map_surface_to_volume_subdomains(ft, ct, facet_to_cell, volume_subdomains, surface_subdomains, comm=None):
    """Reference ``subdomain/volume_subdomain.py:172-260``: which volume subdomain every tagged
    surface subdomain belongs to, from the facet tags, the cell tags and the facet -> cell
    connectivity.  Here facets are the exterior facets of a ``SimplexMesh`` and ``facet_to_cell`` is
    its ``bfacet_cell`` array (one cell per exterior facet).  A surface that touches cells of two
    volume subdomains is an error, as in the reference."""
    import numpy as np

    cells = np.asarray(facet_to_cell)[np.asarray(ft.indices)]
    ctag = dict(zip(np.asarray(ct.indices).tolist(), np.asarray(ct.values).tolist()))
    pairs = sorted({(int(s), ctag[int(c)]) for s, c in zip(np.asarray(ft.values), cells) if int(c) in ctag})
    surfaces = {s.id: s for s in surface_subdomains}
    volumes = {v.id: v for v in volume_subdomains}
    out = {}
    for s_tag, v_tag in pairs:
        s_sub, v_sub = surfaces.get(s_tag), volumes.get(v_tag)
        if s_sub and v_sub:
            if s_sub in out:
                assert out[s_sub] == v_sub, (
                    f"Surface subdomain {s_sub.id} is connected to multiple volume subdomains: "
                    f"{out[s_sub].id} and {v_sub.id}")
            out[s_sub] = v_sub
    return out
