"""Residual form container: the stand-in for the ``ufl.Form`` the reference
builds in ``create_formulation`` (``hydrogen_transport_problem.py:870-997``,
``heat_transfer_problem.py:175-218``).

A ``Form`` is a list of integrals over tagged cells (``dx(id)``) or tagged
exterior facets (``ds(id)``):

* ``DiffusionIntegral``   ``+ int D grad(u_s) . grad(v_s) dx(id)``
* ``TestIntegral``        ``+ int expr v_s dx(id) | ds(id)`` where ``expr`` is a
  closed-form expression of x, t, T, the species concentrations and the
  previous-step concentrations
* ``AdvectionIntegral``   ``+ int (grad(u_s) . vel) v_s dx(id)``

The Jacobian is the symbolic derivative of these integrands with respect to
the species concentrations (the analogue of ``ufl.derivative``), and the
quadrature degree of each (measure, id) group is the maximum of the UFL
degree estimates of its integrands.  ``meta`` records
the structure an integral was built from (Arrhenius constants, reaction
object) so the lowering can choose a specialised device kernel; the CPU oracle
ignores it and evaluates ``expr``.
"""

from __future__ import annotations

from festim_b200 import ufl


class Measure:
    def __init__(self, kind, meshtags):
        self.kind = kind
        self.meshtags = meshtags

    def __call__(self, subdomain_id):
        return (self.kind, int(subdomain_id))


class DiffusionIntegral:
    kind = "diffusion"

    def __init__(self, species, D, measure, meta=None):
        self.species = species
        self.D = ufl.as_expr(D)
        self.measure = measure
        self.meta = meta or {}

    def degree_F(self):
        return ufl.estimate_degree(self.D)

    def degree_J(self):
        deg = ufl.estimate_degree(self.D)
        for spe in ufl.species_in(self.D):
            deg = max(deg, ufl.estimate_degree(ufl.diff(self.D, ("species", spe))) + 1)
        return deg


class TestIntegral:
    kind = "test"

    def __init__(self, species, expr, measure, meta=None):
        self.species = species
        self.expr = ufl.as_expr(expr)
        self.measure = measure
        self.meta = meta or {}
        self._derivs = None

    def derivatives(self):
        """[(species, d expr / d c_species)] for every species expr depends on."""
        if self._derivs is None:
            self._derivs = [
                (spe, ufl.diff(self.expr, ("species", spe))) for spe in ufl.species_in(self.expr)
            ]
        return self._derivs

    def grad_derivatives(self):
        """[((species, component), d expr / d (d_component c_species))]: dependence on the
        cell gradient of an unknown."""
        return [(key, ufl.diff(self.expr, ("species_grad", key))) for key in ufl.species_grads_in(self.expr)]

    # meta["extra_degree"]: the term was rewritten from an expression UFL estimates higher
    # (spherical diffusion, see create_formulation); the integration degree follows the original
    def degree_F(self):
        return ufl.estimate_degree(self.expr) + 1 + int(self.meta.get("extra_degree", 0))

    def degree_J(self):
        extra = int(self.meta.get("extra_degree", 0))
        degs = [ufl.estimate_degree(d) + 2 + extra for _, d in self.derivatives()]
        degs += [ufl.estimate_degree(d) + 1 + extra for _, d in self.grad_derivatives()]  # trial gradient: degree 0
        return max(degs) if degs else None


class NormalGradTestIntegral(TestIntegral):
    """``expr * (n . grad(v)) ds``: the symmetry term of Nitsche's method.  The test
    function enters through its normal derivative, so every vertex of the cell behind the
    facet gets a row (reference ``dirichlet_bc.py:239-241``)."""

    kind = "ngradtest"

    def degree_F(self):
        return ufl.estimate_degree(self.expr)          # grad(v) is constant on the cell

    def degree_J(self):
        degs = [ufl.estimate_degree(d) + 1 for _, d in self.derivatives()]
        degs += [ufl.estimate_degree(d) for _, d in self.grad_derivatives()]
        return max(degs) if degs else None


class AdvectionIntegral:
    kind = "advection"

    def __init__(self, species, velocity, measure, meta=None):
        self.species = species
        self.velocity = velocity
        self.measure = measure
        self.meta = meta or {}

    def degree_F(self):
        return 2

    def degree_J(self):
        return 2


class Form:
    def __init__(self, integrals=None):
        self.integrals = list(integrals or [])

    def __iadd__(self, other):
        if isinstance(other, Form):
            self.integrals.extend(other.integrals)
        elif other == 0:
            pass
        else:
            self.integrals.append(other)
        return self

    def __add__(self, other):
        out = Form(self.integrals)
        out += other
        return out

    __radd__ = __add__

    def groups(self):
        """{(measure kind, id): [integrals]} in insertion order."""
        out = {}
        for itg in self.integrals:
            out.setdefault(itg.measure, []).append(itg)
        return out

    def quadrature_degrees(self):
        """{(kind, id): (degree for F, degree for J or None)}."""
        out = {}
        for key, itgs in self.groups().items():
            dF = max(i.degree_F() for i in itgs)
            dJs = [d for d in (i.degree_J() for i in itgs) if d is not None]
            out[key] = (dF, max(dJs) if dJs else None)
        return out
