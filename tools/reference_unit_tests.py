"""Run the reference's OWN unit tests (unmodified, read from /root/reference/test) against the
API mirror: `festim` is aliased to `festim_b200`, `ufl` / `dolfinx.fem` / `dolfinx.mesh` to the
stand-ins, everything solves through the CPU oracle.  A compatibility report, not part of
`tests/` (the GPU box has no /root/reference):

    python tools/reference_unit_tests.py [test_species.py test_reaction.py ...]
    python tools/reference_unit_tests.py --system [test_fluxes.py ...]   (test/system_tests)

For the system tests, only `tools.error_L2` is replaced (the reference interpolates into a
degree-raised dolfinx space; here the error is integrated with a degree-8 rule).
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/test"
DEFAULT = ["test_species.py", "test_reaction.py", "test_settings.py", "test_stepsize.py", "test_material.py",
           "test_helpers.py", "test_source.py", "test_subdomains.py", "test_flux_bc.py", "test_dirichlet_bc.py",
           "test_sievertsbc.py", "test_henrysbc.py", "test_mesh.py", "test_advection.py",
           "test_h_transport_problem.py", "test_heat_transfer_problem.py", "test_total_volume.py",
           "test_surface_quantity.py", "test_volume_quantity.py", "test_average_volume.py",
           "test_average_surface.py", "test_maximum_volume.py", "test_minimum_volume.py",
           "test_maximum_surface.py", "test_minimum_surface.py", "test_total_surface.py",
           "test_permeation_problem.py", "test_multispecies_problem.py", "test_complex_reaction.py",
           "test_coupled_heat_hydrogen_problem.py", "test_initial_condition.py", "test_derived_quantities.py",
           "test_vtx.py", "test_xdmf.py", "test_profile.py"]

CONFTEST = r'''
import sys, types
sys.path.insert(0, %(root)r)
import numpy as np
import festim_b200 as F
from festim_b200 import ufl as _ufl, fem as _fem, mesh as _mesh
from oracle.newton import OracleBackend

# every problem solves through the CPU oracle
_orig_init = F.problem.ProblemBase.__init__
def _init(self, *a, **k):
    _orig_init(self, *a, **k)
    self.solver_backend = OracleBackend()
F.problem.ProblemBase.__init__ = _init

class _Missing:
    """Placeholder for reference classes outside the scope of this repository: importing
    and naming them works (module-level parametrize lists), using them fails the test."""
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        if name in ("initialise", "run", "iterate", "create_formulation", "create_solver"):
            raise NotImplementedError(f"{type(self).__name__} is out of scope (DESIGN.md section 7)")
        # module-level set-up code of a test file (test_reaction.py builds a discontinuous problem
        # before its first test) must not take the file's other tests down with it
        return _Missing()

    def __call__(self, *a, **k):
        return _Missing()

    def __iter__(self):
        return iter(())

def _missing(name):
    if name.startswith("__"):
        raise AttributeError(name)
    cls = type(name, (_Missing,), {})
    setattr(F, name, cls)
    return cls

F.__getattr__ = _missing
sys.modules["festim"] = F
for name in ("exports", "boundary_conditions", "helpers", "stepsize", "species", "reaction", "material",
             "subdomain", "source", "settings", "problem", "mesh", "hydrogen_transport_problem",
             "heat_transfer_problem", "initial_condition", "advection", "trap", "constants"):
    if hasattr(F, name):
        sys.modules["festim." + name] = getattr(F, name)
sys.modules["ufl"] = _ufl
core = types.ModuleType("ufl.core"); expr = types.ModuleType("ufl.core.expr"); expr.Expr = _ufl.Expr
core.expr = expr; _ufl.core = core
sys.modules["ufl.core"] = core; sys.modules["ufl.core.expr"] = expr
dolfinx = types.ModuleType("dolfinx")
dolfinx.fem = _fem
dmesh = types.ModuleType("dolfinx.mesh")
dmesh.create_unit_square = lambda comm, nx, ny, *a, **k: _mesh.create_unit_square(nx, ny)
dmesh.create_unit_cube = lambda comm, nx, ny, nz, *a, **k: _mesh.create_unit_cube(nx, ny, nz)
dmesh.CellType = types.SimpleNamespace(triangle="triangle", tetrahedron="tetrahedron", interval="interval",
                                       quadrilateral="quadrilateral", hexahedron="hexahedron")
dmesh.create_unit_interval = lambda comm, nx, *a, **k: F.Mesh1D(np.linspace(0, 1, nx + 1)).mesh
dmesh.Mesh = _mesh.SimplexMesh
dmesh.MeshTags = _mesh.MeshTags
dmesh.meshtags = lambda mesh, dim, idx, vals: _mesh.MeshTags(mesh, dim, np.asarray(idx), np.asarray(vals))
dmesh.locate_entities = getattr(_mesh, "locate_entities", None)
dolfinx.mesh = dmesh
dolfinx.default_scalar_type = np.float64
sys.modules["dolfinx"] = dolfinx
sys.modules["dolfinx.fem"] = _fem
sys.modules["dolfinx.mesh"] = dmesh
_io = types.ModuleType("dolfinx.io"); _io.XDMFFile = _Missing; _io.VTXWriter = _Missing
sys.modules["dolfinx.io"] = _io; dolfinx.io = _io
for stub in ("io4dolfinx", "adios4dolfinx", "basix", "basix.ufl", "scifem", "petsc4py", "petsc4py.PETSc",
             "dolfinx.fem.petsc", "dolfinx.nls", "dolfinx.nls.petsc"):
    sys.modules.setdefault(stub, types.ModuleType(stub))
def _evaluate_function(u, points):
    """scifem.evaluate_function: P1 interpolation of a fem.Function at points (brute-force
    search of the containing cell)."""
    mesh = u.function_space.mesh
    dg1 = getattr(u.function_space, "family", "") == "DG" and u.function_space.degree == 1
    vals = None if dg1 else np.asarray(u.x.array).reshape(mesh.num_vertices, -1)
    cellvals = np.asarray(u.x.array).reshape(mesh.num_cells, mesh.tdim + 1) if dg1 else None
    pts = np.asarray(points, dtype=float).reshape(-1, np.asarray(points).shape[-1])[:, : mesh.gdim]
    X = mesh.coords[mesh.cells]                                   # (nc, d+1, d)
    out = []
    for p in pts:
        A = np.transpose(X[:, 1:] - X[:, :1], (0, 2, 1))          # (nc, d, d)
        lam = np.linalg.solve(A, (p - X[:, 0])[:, :, None])[:, :, 0]
        bary = np.concatenate([1 - lam.sum(1, keepdims=True), lam], axis=1)
        c = int(np.argmax(bary.min(axis=1)))
        out.append(bary[c] @ (cellvals[c] if dg1 else vals[mesh.cells[c]]))
    return np.array(out)

sys.modules["scifem"].evaluate_function = _evaluate_function
mpi4py = types.ModuleType("mpi4py"); mpi4py.__path__ = []
MPI = types.ModuleType("mpi4py.MPI"); MPI.COMM_WORLD = None; MPI.SUM = "sum"
mpi4py.MPI = MPI
sys.modules["mpi4py"] = mpi4py; sys.modules["mpi4py.MPI"] = MPI
# sub-modules of ufl the tests import class names from
_ufl.__path__ = []
for sub, names in {"conditional": {"Conditional": _ufl.Conditional}, "indexed": {"Indexed": _ufl.Expr},
                   "algebra": {"Sum": _ufl.Sum, "Product": _ufl.Product, "Division": _ufl.Division},
                   "classes": {"Expr": _ufl.Expr}}.items():
    m = types.ModuleType("ufl." + sub)
    m.__dict__.update(names)
    if not hasattr(_ufl, sub):          # ufl.conditional is also a function: keep it
        setattr(_ufl, sub, m)
    sys.modules["ufl." + sub] = m
'''


SYSTEM_DEFAULT = ["test_1_mobile_species.py", "test_1_mobile_1_trap.py", "test_fluxes.py",
                  "test_advection_term_integration.py", "test_coupled_heat_hydrogen.py", "test_heat_transfer.py",
                  "test_misc.py", "test_non_homogeneous_density.py", "test_reaction_not_in_every_volume.py",
                  "test_surface_reactions.py", "test_cylindrical_coordinates.py", "test_spherical_coordinates.py"]

SYSTEM_TOOLS = r'''
import numpy as np
import ufl
from dolfinx.mesh import create_unit_cube, create_unit_square
from festim import Mesh1D
import sys
sys.path.insert(0, %(root)r + "/tests")
from cases import error_L2 as _error_L2

test_mesh_1d = Mesh1D(np.linspace(0, 1, 10000))
test_mesh_2d = create_unit_square(None, 50, 50)
test_mesh_3d = create_unit_cube(None, 20, 20, 20)
x_1d = ufl.SpatialCoordinate(test_mesh_1d.mesh)
x_2d = ufl.SpatialCoordinate(test_mesh_2d)
x_3d = ufl.SpatialCoordinate(test_mesh_3d)


def error_L2(u_computed, u_exact, degree_raise=3):
    mesh = u_computed.function_space.mesh
    if isinstance(u_exact, ufl.Expr):
        expr = u_exact
        u_exact = lambda x: ufl.evaluate(expr, {"x": np.moveaxis(np.asarray(x), 0, -1)})
    return _error_L2(mesh, np.asarray(u_computed.x.array), u_exact)
'''


def main():
    system = "--system" in sys.argv
    if system:
        sys.argv.remove("--system")
    files = sys.argv[1:] or (SYSTEM_DEFAULT if system else DEFAULT)
    top = tempfile.mkdtemp(prefix="reftests_")
    tmp = os.path.join(top, "test")          # a package, so that `from .utils import ...` works
    os.makedirs(tmp)
    open(os.path.join(tmp, "__init__.py"), "w").close()
    with open(os.path.join(top, "conftest.py"), "w") as fh:
        fh.write(CONFTEST % {"root": ROOT})
    ref = os.path.join(REF, "system_tests") if system else REF
    for f in files + (["test_multi_mat_penalty.py"] if system and "test_misc.py" in files else []) + \
            ([] if system else ["utils.py", "markers.py"]):
        src = os.path.join(ref, f)
        if os.path.exists(src):
            shutil.copy(src, tmp)
    if system:
        with open(os.path.join(tmp, "tools.py"), "w") as fh:
            fh.write(SYSTEM_TOOLS % {"root": ROOT})
    cmd = [sys.executable, "-m", "pytest", tmp, "--rootdir", top, "-q", "--no-header", "-p", "no:cacheprovider",
           "--continue-on-collection-errors", "--tb=" + os.environ.get("TB", "line")] + (["-x"] if os.environ.get("X") else [])
    out = subprocess.run(cmd, cwd=top, capture_output=True, text=True)
    print(out.stdout[-int(os.environ.get("TAIL", 6000)):]); print(out.stderr[-1500:])
    shutil.rmtree(top, ignore_errors=True)


if __name__ == "__main__":
    main()
