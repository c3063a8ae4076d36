"""Host-side profile of the public solve() call on the C1 workload (where the end-to-end
time beyond the device-timed step goes)."""
import cProfile, os, pstats, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from bench import build_workload  # noqa: E402
from festim_b200.backend import B200Backend  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 55
model, H, _ = build_workload("c1", n)
model.solver_backend = B200Backend(device=0)
model.initialise()
solver = model.solver
u0 = model.u.x.array.copy()


def api_step():
    model.u.x.array[:] = u0
    solver._signature = None
    solver.solve()


for _ in range(3):
    api_step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    api_step()
torch.cuda.synchronize()
print(f"solve(): {(time.perf_counter() - t0) * 100:.3f} ms per call; device step timings (ms): {solver.timings}")
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    api_step()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
