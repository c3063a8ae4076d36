import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
t0 = time.time()
def log(*a):
    print(f"[{time.time()-t0:7.2f}s]", *a, file=sys.stderr, flush=True)
import numpy as np
import torch
log("torch imported", torch.cuda.is_available())
from cases import mms_1_mobile_steady
import ctypes as C
dim, n = int(sys.argv[1]), int(sys.argv[2])
maxit = int(sys.argv[3])
m, H, _ = mms_1_mobile_steady(dim, n)
m.initialise()
s = m.solver; ctx = s.ctx; lib = ctx.lib
log("initialised", ctx.n_dofs)
s._push_state()
print("renumbered:", ctx.perm is not None, file=sys.stderr)
u = ctx.zeros(ctx.n_dofs); F = ctx.zeros(ctx.n_dofs); J = ctx.zeros(ctx.nnz); y = ctx.zeros(ctx.n_dofs)
ctx._check(lib.fb_assemble(ctx.h, ctx.ptr(u), ctx.ptr(u), ctx.ptr(F), ctx.ptr(J), 3))
torch.cuda.synchronize(); log("assembled")
for mode in sys.argv[4].split(","):
    os.environ.pop("FB_CG_MULTIKERNEL", None); os.environ.pop("FB_NO_SMEM_OPERATOR", None)
    if mode == "multi": os.environ["FB_CG_MULTIKERNEL"] = "1"
    os.environ.pop("FB_NO_SM_PARTITION", None)
    if mode == "persistent-nosmem": os.environ["FB_NO_SMEM_OPERATOR"] = "1"
    if mode == "persistent-range": os.environ["FB_NO_SM_PARTITION"] = "1"
    os.environ.pop("FB_NO_COARSE", None)
    if mode == "persistent-nocoarse": os.environ["FB_NO_COARSE"] = "1"
    its = C.c_int(0); rr = C.c_double(0)
    log("solve", mode)
    t1 = time.time()
    rc = lib.fb_linear_solve(ctx.h, ctx.ptr(J), ctx.ptr(F), ctx.ptr(y), 1, 0, 1e-12, 0.0, maxit, C.byref(its), C.byref(rr))
    torch.cuda.synchronize()
    log(mode, "rc", rc, "its", its.value, "relres", rr.value, "time", time.time() - t1, lib.fb_last_error(ctx.h))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        lib.fb_linear_solve(ctx.h, ctx.ptr(J), ctx.ptr(F), ctx.ptr(y), 1, 0, 1e-12, 0.0, maxit, C.byref(its), C.byref(rr))
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    log(mode, f"avg {ms:.3f} ms per solve, {1e3*ms/max(its.value,1):.2f} us per iteration")
