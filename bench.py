#!/usr/bin/env python
"""Benchmark of the FESTIM hot path on B200: Newton-step DOFs/s.

Default workload (N=1): BASELINE.json configs[2] (C2), the largest configuration that fits one
GPU - 3D H/D/T transient with 3 trap types on an n=94 Kuhn cube (4 983 504 cells, 12 species,
10 288 500 dofs).  A "step" is one implicit-Euler time step of that problem: Newton iterations
(assemble F and the Jacobian operator, preconditioned GMRES, update, residual check), timed on
the device; the metric is dofs x Newton iterations / time.  `--workload c1` keeps the round-1
configuration (3D MMS steady diffusion, n=55).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                  [--workload c2|c1] [--n 94] [--parity-n 12] [--ref-n 8]

The same run checks the answer (parity_rel_l2: GPU vs the CPU oracle on the same problem at
--parity-n, through the same solver path) and times the oracle on the host (cpu_baseline).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
import warnings

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "newton_step_dofs_per_s"
UNIT = "DOFs/s"


def proc_grid(world):
    return {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}.get(world, (world, 1, 1))


def c2_counts(n):
    return 6 * n ** 3, (n + 1) ** 3, 12 * (n + 1) ** 3


def workload_name(name, n):
    if name == "c2":
        cells, nv, dofs = c2_counts(n)
        return (f"C2: 3D H/D/T transient, 3 trap types (12 species, 9 reactions), P1 tets n={n}, "
                f"{cells} cells, {dofs} dofs, implicit Euler dt=1e-2")
    return f"C1: 3D unit-cube MMS steady diffusion (reference test_1_mobile_MMS_3D), P1 tets n={n}"


def build_c2(n, chebyshev=False):
    from cases import multi_isotope_traps_3d

    import festim_b200 as F

    m, iso, trapped = multi_isotope_traps_3d(n, final_time=1e9)
    if chebyshev:
        m.petsc_options = {"pc_type": "chebyshev"}
    # what a user reads every step: the hydrogen inventory (a device reduction, 8 bytes back)
    m.exports = [F.TotalVolume(field=iso[0], volume=m.subdomains[0])]
    return m


def build_c1(n, grid=(1, 1, 1)):
    """C1, or its weak-scaling extension: the same problem on the box
    [0,px] x [0,py] x [0,pz] with n cells per unit length (one C1 per GPU)."""
    from cases import mms_1_mobile_steady

    if grid == (1, 1, 1):
        return mms_1_mobile_steady(3, n)[0]
    import numpy as np

    import festim_b200 as F
    from festim_b200 import fem, ufl

    px, py, pz = grid
    raw = F.create_box(((0.0, 0.0, 0.0), (float(px), float(py), float(pz))), (n * px, n * py, n * pz))
    u_exact = lambda mod: (lambda x: 1 + mod.sin(2 * mod.pi * x[0]) + mod.cos(2 * mod.pi * x[1]))
    T_expr = lambda x: 500 + 100 * x[0]
    x = ufl.SpatialCoordinate(raw)
    T = fem.Function(fem.functionspace(raw, ("Lagrange", 1)))
    T.interpolate(T_expr)
    D = 1 * ufl.exp(-0.1 / (F.k_B * T))
    m = F.HydrogenTransportProblem()
    m.mesh = F.Mesh(raw)
    vol = F.VolumeSubdomain(id=1, material=F.Material(D_0=1, E_D=0.1))
    left = F.SurfaceSubdomain(id=1, locator=lambda x: np.isclose(x[0], 0))
    right = F.SurfaceSubdomain(id=2, locator=lambda x: np.isclose(x[0], float(px)))
    m.subdomains = [vol, left, right]
    H = F.Species("H")
    m.species = [H]
    m.temperature = T_expr
    m.boundary_conditions = [F.DirichletBC(subdomain=left, value=u_exact(ufl), species=H),
                             F.DirichletBC(subdomain=right, value=u_exact(ufl), species=H)]
    m.sources = [F.ParticleSource(value=-ufl.div(D * ufl.grad(u_exact(ufl)(x))), volume=vol, species=H)]
    m.settings = F.Settings(atol=1e-10, rtol=1e-10, transient=False)
    return m


class ClockSampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.stop_flag, self.max_mhz = [], set(), False, None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", "0", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for nme, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(nme)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        import statistics

        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json, copy bandwidth)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# CPU side: the oracle (restated reference path) on a bounded sample of the workload
# ---------------------------------------------------------------------------------------------
def oracle_steps(workload, n, nsteps):
    """Run `nsteps` steps of the workload at mesh size n through the oracle; returns
    (dofs, seconds per step list, Newton iterations per step list, final field)."""
    from oracle.newton import OracleBackend

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = build_c2(n) if workload == "c2" else build_c1(n)
        m.solver_backend = OracleBackend()
        m.initialise()
    secs, its = [], []
    for _ in range(nsteps):
        t0 = time.perf_counter()
        if workload == "c2":
            m.iterate()
        else:
            m.u.x.array[:] = 0.0
            m.solver.solve()
        secs.append(time.perf_counter() - t0)
        its.append(max(m.solver.solver.getIterationNumber(), 1))
    return m.u.x.array.size, secs, its, m.u.x.array.copy()


_REF = {}


def _ref_init(workload, n):
    from oracle.newton import OracleBackend

    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = "1"
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = build_c2(n) if workload == "c2" else build_c1(n)
        m.solver_backend = OracleBackend()
        m.initialise()
    _REF["m"], _REF["workload"] = m, workload


def _ref_step(_):
    m = _REF["m"]
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if _REF["workload"] == "c2":
            m.iterate()
        else:
            m.u.x.array[:] = 0.0
            m.solver.solve()
    return m.u.x.array.size, max(m.solver.solver.getIterationNumber(), 1), time.perf_counter() - t0


def run_reference(args, config):
    """`--impl reference`: the restated reference path (oracle: numpy assembly + SuperLU Newton) on
    the host cores.  The port is single-threaded, so one independent replica of a bounded sample
    of the workload runs per core; every timed step advances all replicas by one step."""
    import multiprocessing as mp

    cores = max(1, min(os.cpu_count() or 1, 64))
    n = args.ref_n
    with mp.get_context("fork").Pool(cores, initializer=_ref_init, initargs=(args.workload, n)) as pool:
        for _ in range(args.warmup):
            pool.map(_ref_step, range(cores))
        t0 = time.perf_counter()
        work = 0
        for _ in range(args.steps):
            for ndofs, its, _ in pool.map(_ref_step, range(cores)):
                work += ndofs * its
        wall = time.perf_counter() - t0
    v = work / wall
    sample = (f"{cores} independent replicas (one per host core) of the same problem on an n={n} mesh "
              f"({ndofs} dofs each), every step = one {'time step' if args.workload == 'c2' else 'steady solve'} "
              f"of every replica: numpy assembly + SuperLU Newton solve (restated reference path, "
              f"NOT dolfinx+PETSc, which is absent from this image)")
    cfg = dict(config)
    cfg["parallelism"] = f"{cores} host cores, one replica of the sample each"
    cfg.pop("l2", None)
    cfg["reference_sample"] = {"n": n, "dofs_per_replica": int(ndofs), "replicas": cores}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall * 1e3 / args.steps,
        "higher_is_better": True, "scaling": "strong" if args.workload == "c2" else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ---------------------------------------------------------------------------------------------
# N = 1
# ---------------------------------------------------------------------------------------------
def gpu_parity(workload, n, nsteps, ref_field):
    """The same problem at the oracle's size through the same solver path (channel operator,
    Chebyshev-preconditioned GMRES): relative L2 distance of the fields."""
    import numpy as np

    from festim_b200.backend import B200Backend

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = build_c2(n, chebyshev=True) if workload == "c2" else build_c1(n)
        m.solver_backend = B200Backend()
        m.initialise()
    its = []
    for _ in range(nsteps):
        if workload == "c2":
            m.iterate()
        else:
            m.u.x.array[:] = 0.0
            m.solver.solve()
        its.append(max(m.solver.solver.getIterationNumber(), 1))
    rel = float(np.linalg.norm(m.u.x.array - ref_field) / np.linalg.norm(ref_field))
    info = (C.c_double * 4)()
    m.solver.ctx.lib.fb_chebyshev_info(m.solver.ctx.h, info)
    return rel, its, bool(info[0]), bool(m.solver.spec.vx is not None)


def run_single(args, config):
    import numpy as np
    import torch

    import __graft_entry__ as g
    from festim_b200.backend import B200Backend

    g.build()
    torch.cuda.set_device(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = build_c2(args.n) if args.workload == "c2" else build_c1(args.n)
        model.solver_backend = B200Backend(device=0)
        t0 = time.perf_counter()
        model.initialise()
        setup_s = time.perf_counter() - t0
    solver = model.solver
    ctx, lib, spec = solver.ctx, solver.ctx.lib, solver.spec
    N = ctx.n_dofs
    nnzb = len(ctx.dmesh.graph[1])
    transient = args.workload == "c2"
    flush = torch.empty(512 * 1024 * 1024 // 8, dtype=torch.float64, device=ctx.device)

    # ---- device-resident steps ("value") -------------------------------------------------
    solver._push_state()
    u_dev, un_dev = ctx.zeros(N), ctx.zeros(N)
    n_its, reason, lin = C.c_int(0), C.c_int(0), C.c_int(0)
    fnorm = C.c_double(0.0)
    tot = {"newton": 0, "linear": 0}

    def device_step():
        if transient:
            un_dev.copy_(u_dev)
        else:
            u_dev.zero_()
            ctx._check(lib.fb_update_load(ctx.h))
        rc = lib.fb_newton_solve(ctx.h, ctx.ptr(u_dev), ctx.ptr(un_dev if transient else u_dev),
                                 float(solver.atol), float(solver.rtol), solver.stol, solver.max_it,
                                 solver.ksp, solver.pc, solver.ksp_rtol, solver.ksp_max_it,
                                 C.byref(n_its), C.byref(reason), C.byref(fnorm), C.byref(lin))
        ctx._check(rc)
        assert reason.value > 0, f"Newton did not converge: reason {reason.value}"
        tot["newton"] += max(n_its.value, 1)
        tot["linear"] += lin.value

    for _ in range(args.warmup):
        device_step()
    sampler = ClockSampler()
    sampler.start()
    torch.cuda.synchronize()
    tot["newton"] = tot["linear"] = 0
    l0 = ctx.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    phase = np.zeros(4)
    for k in range(args.steps):
        flush.fill_(float(k))
        ev[k][0].record()
        device_step()
        ev[k][1].record()
        tm = (C.c_double * 4)()
        lib.fb_last_timings(ctx.h, tm)
        phase += np.array(list(tm))
    torch.cuda.synchronize()
    launches = ctx.launches - l0
    ms = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    newton_per_step = tot["newton"] / args.steps
    linear_per_step = tot["linear"] / args.steps
    phase /= args.steps
    info = (C.c_double * 4)()
    lib.fb_chebyshev_info(ctx.h, info)

    # ---- end to end through the public API ("e2e"): host arrays in, host arrays out ------
    e2e_its = []

    def api_step():
        if transient:
            model.iterate()
        else:
            model.u.x.array[:] = 0.0
            solver._signature = None  # a fresh steady solve integrates its sources again
            solver.solve()
        e2e_its.append(max(solver.solver.getIterationNumber(), 1))
        # the step's result on the host: the inventory export for the transient (the fields stay on
        # the device between steps, as the state of a time integrator does), the field for a steady solve
        return float(model.exports[0].data[-1]) if transient else float(model.u.x.array[0])

    for _ in range(max(args.warmup, 1)):   # the API path has its own fields: same warm-up as the device steps
        api_step()
    torch.cuda.synchronize()
    e2e_its.clear()
    t0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(float(k))
        api_step()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / args.steps
    # the same with the whole field read back and re-uploaded every step (what round 1 measured)
    e2e_field_s = None
    if transient:
        t0 = time.perf_counter()
        for k in range(args.steps):
            api_step()
            float(model.u.x.array[0])
        torch.cuda.synchronize()
        e2e_field_s = (time.perf_counter() - t0) / args.steps
    sampler.stop_flag = True
    sampler.join(timeout=2)
    nbc = len(solver._bc_dofs)
    h2d = (8 * nbc + 8 * len(spec.scalars)) if transient else (8 * N + 8 * nbc)
    d2h = 8 * len(model.exports) + 64 if transient else 8 * N
    e2e_val = N * (sum(e2e_its) / len(e2e_its)) / e2e_s

    # ---- roofline of the dominant kernels, timed alone with CUDA events ---------------------
    peak, peak_src = peaks()
    rng = np.random.default_rng(0)
    x = ctx.to_device("bench_x", rng.standard_normal(N))
    y, Fv = ctx.zeros(N), ctx.zeros(N)
    channel = spec.vx is not None and bool(getattr(spec, "vx", None))

    def timed(fn, reps):
        for _ in range(3):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    mesh = spec.mesh
    d = mesh.tdim
    roofline = None
    if channel:
        vx = spec.vx
        nsp, ndp = (vx["nstat"] + 1) // 2 * 2, (vx["ndyn"] + 1) // 2 * 2
        ntens = len(vx["tensors"])
        spmv_ms = timed(lambda: ctx._check(lib.fb_spmv(ctx.h, None, ctx.ptr(x), ctx.ptr(y))), 30)
        # full Newton-iteration assembly: alternate between two solutions, otherwise the library streams the
        # solution-dependent channels it already holds for that solution instead of contracting the tensors again
        u_alt = [u_dev, u_dev * (1.0 + 1e-9)]
        tog = [0]

        def asm_full():
            tog[0] ^= 1
            ctx._check(lib.fb_assemble(ctx.h, ctx.ptr(u_alt[tog[0]]), ctx.ptr(un_dev), ctx.ptr(Fv), None, 3))
        asm_ms = timed(asm_full, 10)
        asm_reuse_ms = timed(lambda: ctx._check(lib.fb_assemble(ctx.h, ctx.ptr(u_dev), ctx.ptr(un_dev), ctx.ptr(Fv), None, 3)), 10)
        # algorithmic bytes of OUR layout: every array touched once
        B_spmv = 8 * (nsp + ndp) * nnzb + 4 * nnzb + 4 * (mesh.num_vertices + 1) + 16 * N
        # assembly: static channel values + the edge tensor in, dynamic channel values out, u, u_n, F, load
        tz_entries = getattr(solver, "_tz_entries", None)
        B_asm = 8 * nsp * nnzb + 8 * ndp * nnzb + 16 * ntens * nnzb + 4 * nnzb + 8 * N * (4 if transient else 3)
        # the same operator in the species-sparse block-CSR layout of SURVEY 8d (8 |P| nnzb values)
        B_spmv_bsr = 8 * spec.npairs * nnzb + 4 * nnzb + 4 * (mesh.num_vertices + 1) + 16 * N
        B_asm_bsr = (mesh.num_cells * (4 * (d + 1) + 4) + mesh.num_vertices * (8 * d + 8 * spec.T_mode + 16 * spec.ns)
                     + 8 * spec.npairs * nnzb + 8 * N)
        traffic = None
        prof = os.path.join(ROOT, "profiles", "r2_ncu_spmv_cheb_step.json")
        if os.path.exists(prof):
            try:
                traffic = json.load(open(prof)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        ach = B_spmv / (spmv_ms * 1e-3) / 1e9
        # the launch the time step is made of: one step of the Chebyshev recurrence (same kernel, MODE 2) -
        # operator, block-Jacobi inverse blocks (4-byte mantissas + row scales) and the three vectors of the
        # recurrence in one pass
        ctx._check(lib.fb_bjacobi_setup(ctx.h, None))
        cr, cz, cd = ctx.zeros(N), ctx.zeros(N), ctx.zeros(N)
        bufs = [x, cd]

        def cheb_step():
            ctx._check(lib.fb_chebyshev_step(ctx.h, ctx.ptr(bufs[0]), ctx.ptr(cr), ctx.ptr(cz), ctx.ptr(bufs[1]), 0.25, 0.05))
            bufs.reverse()
        # systems the TMA / DMMA kernel does not take (one species: C1) have no Chebyshev-step launch:
        # their roofline entry is the plain operator application
        has_cheb = lib.fb_chebyshev_step(ctx.h, ctx.ptr(x), ctx.ptr(cr), ctx.ptr(cz), ctx.ptr(cd), 0.25, 0.05) == 0
        cheb_ms = timed(cheb_step, 30) if has_cheb else spmv_ms
        B_cheb = (B_spmv + (40 + 4 * spec.ns) * N) if has_cheb else B_spmv   # - y, + r, z, row scales in, + r, d, z out, + 4 ns^2 per vertex
        ach2 = B_cheb / (cheb_ms * 1e-3) / 1e9
        if not has_cheb:
            traffic = None   # the ncu capture in profiles/ is of the C2 kernel
        roofline = {
            "kernel": ("k_spmv_ch_tma<.,.,2> (one step of the Chebyshev recurrence on D^-1 J: channel-operator SpMV "
                       "through the TMA bulk-copy pipeline and fp64 DMMA, fused with the block-Jacobi application and "
                       "the vector updates): the launch behind ~7 of every 8 operator applications of a time step")
                      if has_cheb else "k_spmv_ch (y = J x on the channel operator, warp per block row)",
            "bound": "hbm", "achieved": ach2, "peak": peak, "unit": "GB/s", "frac": ach2 / peak,
            "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes": B_cheb, "ms": cheb_ms,
            "note": "algorithmic bytes of the channel-operator layout (8 bytes per channel and edge: "
                    f"{nsp}+{ndp} channels instead of {spec.npairs} species pairs) + 4 ns^2 + 8 ns bytes of inverse "
                    "block per vertex + 7 vector passes (x, r, z, scales in; r, d, z out); timed alone after the steps (burst), inputs larger than L2",
            "operator_only": {"kernel": "k_spmv_ch_tma<.,.,0> (y = J x)", "algorithmic_bytes": B_spmv, "ms": spmv_ms,
                              "achieved": ach, "frac": ach / peak},
            "block_csr_equivalent": {
                "what": "the same products counted with SURVEY 8d's block-CSR bytes (8 |P| nnzb): what a "
                        "block-CSR SpMV / assembly would have had to move in the same time",
                "spmv_bytes": B_spmv_bsr, "spmv_gbs": B_spmv_bsr / (spmv_ms * 1e-3) / 1e9,
                "spmv_frac_of_peak": B_spmv_bsr / (spmv_ms * 1e-3) / 1e9 / peak,
                "assembly_bytes": B_asm_bsr, "assembly_gbs": B_asm_bsr / (asm_ms * 1e-3) / 1e9,
                "assembly_frac_of_peak": B_asm_bsr / (asm_ms * 1e-3) / 1e9 / peak},
            "assembly": {"kernel": "k_assemble_edge_tma (F + solution-dependent channel values, per Newton iteration)",
                         "algorithmic_bytes": B_asm, "ms": asm_ms, "achieved": B_asm / (asm_ms * 1e-3) / 1e9,
                         "frac": B_asm / (asm_ms * 1e-3) / 1e9 / peak,
                         "ms_same_solution": asm_reuse_ms,
                         "note": "ms_same_solution: the first assembly of a time step, at the solution whose channels "
                                 "the residual check of the previous step left in memory (streamed, not recomputed)"}}
    else:
        J = ctx.zeros(ctx.nnz)
        ctx._check(lib.fb_assemble(ctx.h, ctx.ptr(u_dev), ctx.ptr(u_dev), ctx.ptr(Fv), ctx.ptr(J), 3))
        spmv_ms = timed(lambda: ctx._check(lib.fb_spmv(ctx.h, ctx.ptr(J), ctx.ptr(x), ctx.ptr(y))), 30)
        B_spmv = 8 * spec.npairs * nnzb + 4 * nnzb + 4 * (mesh.num_vertices + 1) + 16 * N
        ach = B_spmv / (spmv_ms * 1e-3) / 1e9
        roofline = {"kernel": "k_spmv", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "traffic": None, "peak_source": peak_src,
                    "algorithmic_bytes": B_spmv, "ms": spmv_ms}

    out = {
        "metric": METRIC, "value": N * newton_per_step / (ms * 1e-3), "unit": UNIT, "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong" if transient else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config,
        "newton_iterations_per_step": newton_per_step, "linear_iterations_per_step": linear_per_step,
        "phase_ms": {"assemble_F_later_iterations": phase[0], "assemble_first_iteration": phase[1],
                     "linear_solve": phase[2], "update": phase[3]},
        "solver": {"ksp": "gmres(30)", "pc": "chebyshev polynomial of the block-Jacobi-scaled operator"
                   if info[0] else "block-Jacobi",
                   "chebyshev_interval": [info[1], info[2]], "operator_applications_per_krylov_iteration": info[3],
                   "operator": "channel operator" if channel else "block-CSR", "host_setup_s": setup_s},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s * 1e3,
                "what": "model.iterate(): Dirichlet values + scalars up, Newton status + the TotalVolume export down; "
                        "u and u_n stay on the device between steps" if transient else "solver.solve(): u up, u down"},
        "gpu_launches": int(launches), "clocks": sampler.summary(), "roofline": roofline,
    }
    if e2e_field_s is not None:
        out["e2e_with_field_readback"] = {
            "value": N * (sum(e2e_its) / len(e2e_its)) / e2e_field_s, "unit": UNIT, "ms_per_step": e2e_field_s * 1e3,
            "h2d_bytes_per_step": 8 * N + 8 * nbc, "d2h_bytes_per_step": 8 * N,
            "what": "as e2e, but u.x.array is read on the host after every step (whole field down, and up again "
                    "at the next step)"}
    # ---- the answer: parity with the oracle on a bounded sample, which is also the CPU baseline --
    nsteps = 2 if transient else 1
    ndofs, secs, its, ref_field = oracle_steps(args.workload, args.parity_n, nsteps)
    rel, gits, cheb, chan = gpu_parity(args.workload, args.parity_n, nsteps, ref_field)
    out["parity_rel_l2"] = rel
    out["parity"] = {"n": args.parity_n, "dofs": int(ndofs), "steps": nsteps, "newton_iterations_gpu": gits,
                     "newton_iterations_oracle": its, "chebyshev_path": cheb, "channel_operator": chan,
                     "tolerance": 1e-8}
    out["cpu_baseline"] = {
        "value": ndofs * sum(its) / sum(secs), "unit": UNIT, "cores": 1, "kind": "port",
        "sample": f"same problem on an n={args.parity_n} mesh ({ndofs} dofs), {nsteps} "
                  f"{'time steps' if transient else 'steady solve'}: numpy assembly + SuperLU Newton solve, "
                  f"{sum(secs):.1f} s; restated reference path, not dolfinx+PETSc"}
    print(json.dumps(out))
    assert rel < 1e-8, f"parity with the oracle lost: rel L2 {rel:.3e}"
    assert gits == its, f"Newton iteration counts differ: GPU {gits}, oracle {its}"


# ---------------------------------------------------------------------------------------------
# N > 1: the mesh partitioned over the ranks
# ---------------------------------------------------------------------------------------------
def run_partitioned(args, rank, world, local_rank, config):
    """N > 1: the mesh is partitioned over the ranks (one C1-sized block per GPU, weak scaling);
    halo exchange + dot products over NVLink peer memory, max-over-ranks device timing."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from festim_b200.distributed import DistributedBackend

    grid = proc_grid(world)
    model = build_c1(args.n, grid)
    model.solver_backend = DistributedBackend(device=local_rank)
    model.initialise()
    solver = model.solver
    ctx = solver.ctx
    N_global = model.u.x.array.size
    stream = solver.stream          # library launches, torch copies and NCCL share this stream
    with torch.cuda.stream(stream):
        flush = torch.empty(512 * 1024 * 1024 // 8, dtype=torch.float64, device=ctx.device)
        solver._push_state()
        u = ctx.zeros(ctx.n_dofs)
    stream.synchronize()

    def device_step():
        with torch.cuda.stream(stream):
            u.zero_()
            ctx._check(ctx.lib.fb_update_load(ctx.h))
            solver.newton_device(u, u, float(solver.atol), float(solver.rtol))
        assert solver.solver.reason > 0, solver.solver.reason

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        device_step()
    sampler = ClockSampler()
    if rank == 0:
        sampler.start()
    barrier()
    l0 = ctx.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    for k in range(args.steps):
        with torch.cuda.stream(stream):
            flush.fill_(float(k))
        ev[k][0].record(stream)
        device_step()
        ev[k][1].record(stream)
    barrier()
    launches = ctx.launches - l0
    ms = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    its, lin = max(solver.solver.its, 1), solver.solver.linear_its

    def api_step():
        model.u.x.array[:] = 0.0
        solver._signature = None
        solver.solve()

    api_step()
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        with torch.cuda.stream(stream):
            flush.fill_(float(k))
        api_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.steps
    t = torch.tensor([ms, e2e_s], dtype=torch.float64, device=ctx.device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_s = float(t[0]), float(t[1])
    sampler.stop_flag = True
    if rank == 0:
        config = dict(config)
        how = ("whole-solve persistent CG per rank, halo pushed and dot products exchanged through "
               "NVLink peer memory (no NCCL inside the solve)") if solver.peer else \
            "halo exchange + all-reduce over NCCL per iteration"
        config.update({"parallelism": f"mesh partitioned over {world} GPUs (RCB, grid {grid}), " + how,
                       "cells": 6 * args.n ** 3 * world, "dofs": int(N_global),
                       "halo_bytes_per_exchange_rank0": solver.halo.bytes_per_exchange})
        print(json.dumps({
            "metric": METRIC, "value": N_global / (ms * 1e-3 / its), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config, "newton_iterations_per_step": its, "linear_iterations_per_step": lin,
            "e2e": {"value": N_global / (e2e_s / its), "unit": UNIT,
                    "h2d_bytes_per_step": 8 * ctx.n_dofs + 8 * len(solver._bc_dofs),
                    "d2h_bytes_per_step": 8 * int(N_global), "ms_per_step": e2e_s * 1e3},
            "gpu_launches": int(launches), "clocks": sampler.summary(),
            "roofline": None, "owned_vertices_rank0": int(solver.part.n_owned),
            "peer_memory_cg": bool(solver.peer),
            "note": "weak scaling: one C1-sized block of the domain per GPU, so the global problem and "
                    "its CG iteration count grow with N"}))
    dist.barrier()
    dist.destroy_process_group()


def run_partitioned_c2(args, rank, world, local_rank, config):
    """N > 1, strong scaling: the C2 mesh partitioned over the ranks (vertex-owned RCB parts).  Every
    rank runs the library's own Newton-GMRES on its part; halo exchange and the cross-rank
    completion of the reductions go through NVLink peer memory inside the library's kernels (no NCCL
    call in the solve).  Rank 0 then repeats the same steps on one GPU and compares the fields."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from festim_b200.backend import B200Backend
    from festim_b200.distributed import DistributedBackend

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = build_c2(args.n)
        model.solver_backend = DistributedBackend(device=local_rank, lazy_gather=True)
        t0 = time.perf_counter()
        model.initialise()
        setup_s = time.perf_counter() - t0
    solver = model.solver
    ctx = solver.ctx
    assert solver.native, "the partitioned C2 bench needs the native peer-memory path"
    N_global = model.u.x.array.size
    stream = solver.stream
    with torch.cuda.stream(stream):
        flush = torch.empty(512 * 1024 * 1024 // 8, dtype=torch.float64, device=ctx.device)
        solver._push_state()
        u = ctx.to_device("u", model.u.x.array[solver.ldofs])
        un = ctx.to_device("u_n", model.u_n.x.array[solver.ldofs])
    stream.synchronize()
    tot = {"newton": 0, "linear": 0}

    def device_step():
        with torch.cuda.stream(stream):
            un.copy_(u)
            solver.newton_device(u, un, float(solver.atol), float(solver.rtol))
        assert solver.solver.reason > 0, solver.solver.reason
        tot["newton"] += max(solver.solver.its, 1)
        tot["linear"] += solver.solver.linear_its

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        device_step()
    sampler = ClockSampler()
    if rank == 0:
        sampler.start()
    barrier()
    tot["newton"] = tot["linear"] = 0
    l0 = ctx.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    for k in range(args.steps):
        with torch.cuda.stream(stream):
            flush.fill_(float(k))
        ev[k][0].record(stream)
        device_step()
        ev[k][1].record(stream)
    barrier()
    launches = ctx.launches - l0
    ms = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    newton_per_step, linear_per_step = tot["newton"] / args.steps, tot["linear"] / args.steps
    nsteps_done = args.warmup + args.steps
    # the distributed field after warmup + steps time steps, in the global numbering on every rank
    with torch.cuda.stream(stream):
        solver._gather(u)
    stream.synchronize()
    dist_field = model.u.x.array.copy()

    # ---- end to end through the public API: host arrays in (per-rank slices), global field out ----
    e2e_its = []

    def api_step():
        model.iterate()
        e2e_its.append(max(solver.solver.getIterationNumber(), 1))
        return float(model.exports[0].data[-1])   # the inventory export: a device reduction + all-reduce

    for _ in range(max(args.warmup, 1)):
        api_step()
    barrier()
    e2e_its.clear()
    t0 = time.perf_counter()
    for k in range(args.steps):
        with torch.cuda.stream(stream):
            flush.fill_(float(k))
        api_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.steps
    t = torch.tensor([ms, e2e_s], dtype=torch.float64, device=ctx.device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_s = float(t[0]), float(t[1])
    sampler.stop_flag = True
    if rank == 0:
        # parity of the partitioned solve: the same time steps on one GPU (this rank's)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = build_c2(args.n)
            ref.solver_backend = B200Backend(device=local_rank)
            ref.initialise()
        ref_its = []
        for _ in range(nsteps_done):
            ref.iterate()
            ref_its.append(ref.solver.solver.getIterationNumber())
        rel = float(np.linalg.norm(dist_field - ref.u.x.array) / np.linalg.norm(ref.u.x.array))
        config = dict(config)
        config.update({"parallelism": f"mesh partitioned over {world} GPUs (vertex-owned RCB parts); the library's "
                                      "Newton-GMRES per rank, halo exchange and cross-rank reductions as tagged words "
                                      "in NVLink peer memory (no NCCL inside the solve)",
                       "owned_vertices_rank0": int(solver.part.n_owned),
                       "ghost_vertices_rank0": int(len(solver.part.ghosts)),
                       "halo_bytes_per_exchange_rank0": solver.halo.bytes_per_exchange})
        info = (C.c_double * 4)()
        ctx.lib.fb_chebyshev_info(ctx.h, info)
        e2e_val = N_global * (sum(e2e_its) / len(e2e_its)) / e2e_s
        print(json.dumps({
            "metric": METRIC, "value": N_global * newton_per_step / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config, "newton_iterations_per_step": newton_per_step,
            "linear_iterations_per_step": linear_per_step,
            "solver": {"ksp": "gmres(30)", "pc": "chebyshev polynomial of the block-Jacobi-scaled operator"
                       if info[0] else "block-Jacobi", "chebyshev_interval": [info[1], info[2]],
                       "host_setup_s": setup_s},
            "e2e": {"value": e2e_val, "unit": UNIT,
                    "h2d_bytes_per_step": 8 * len(solver._bc_dofs) + 8 * len(solver.spec.scalars),
                    "d2h_bytes_per_step": 8 * len(model.exports) + 64, "ms_per_step": e2e_s * 1e3,
                    "what": "model.iterate() on every rank: Dirichlet values + scalars up, Newton status + the "
                            "TotalVolume export (device reduction + all-reduce) down; fields stay on the devices"},
            "gpu_launches": int(launches), "clocks": sampler.summary(), "roofline": None,
            "parity_rel_l2": rel,
            "parity": {"against": f"the same {nsteps_done} time steps on one GPU (rank 0), full size",
                       "tolerance": 1e-8, "newton_iterations_one_gpu_last_step": ref_its[-1]}}))
        assert rel < 1e-8, f"partitioned solve differs from the single-GPU solve: rel L2 {rel:.3e}"
    dist.barrier()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default=None)
    ap.add_argument("--n", "--mesh-n", dest="n", type=int, default=None)
    ap.add_argument("--parity-n", type=int, default=None)
    ap.add_argument("--ref-n", type=int, default=None)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload is None:
        args.workload = "c2"
    if args.workload not in ("c1", "c2"):
        raise SystemExit(f"unknown workload {args.workload}")
    if args.n is None:
        args.n = 94 if args.workload == "c2" else 55
    if args.parity_n is None:
        args.parity_n = 12 if args.workload == "c2" else 24
    if args.ref_n is None:
        args.ref_n = 8 if args.workload == "c2" else 24
    config = {"workload": workload_name(args.workload, args.n), "n": args.n,
              "l2": "L2 flushed (512 MiB write) between timed steps; every streamed array is larger than L2"
              if args.workload == "c2" else "L2 flushed (512 MiB write) between timed steps",
              "parallelism": "1 GPU" if world == 1 else f"{world} GPUs"}
    if args.workload == "c2":
        cells, nv, dofs = c2_counts(args.n)
        config.update({"cells": cells, "dofs": dofs, "species": 12})
    else:
        config.update({"cells": 6 * args.n ** 3, "dofs": (args.n + 1) ** 3})

    if args.impl == "reference":
        if rank != 0:
            return
        return run_reference(args, config)

    import torch

    import __graft_entry__ as g

    if world > 1:
        import torch.distributed as dist

        g.build()
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        torch.cuda.set_device(local_rank)
        if args.workload == "c2":
            return run_partitioned_c2(args, rank, world, local_rank, config)
        return run_partitioned(args, rank, world, local_rank, config)
    return run_single(args, config)


if __name__ == "__main__":
    main()
