"""Host-side profile of model.iterate() on the C2 workload: where the end-to-end time beyond the
device-timed step goes (cProfile, cumulative)."""
import cProfile
import os
import pstats
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from bench import build_c2  # noqa: E402
from festim_b200.backend import B200Backend  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    model = build_c2(n)
    model.solver_backend = B200Backend(device=0)
    model.initialise()
for _ in range(5):
    model.iterate()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    model.iterate()
torch.cuda.synchronize()
print(f"iterate(): {(time.perf_counter() - t0) * 100:.3f} ms per call; device phases (ms): {[round(float(v), 3) for v in model.solver.timings]}")
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    model.iterate()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
