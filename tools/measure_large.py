"""HBM-roofline measurement on a larger-than-L2 configuration (BASELINE configs[2],
C2: 3D H/D/T transient, 3 trap types, 12 species).  Prints one JSON object with the
achieved bytes/s of the SpMV and assembly kernels against the measured HBM peak.

    python tools/measure_large.py --n 94        (5.0 M cells, 10.3 M dofs)
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=94)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--solve", type=int, default=1)
    args = ap.parse_args()
    import numpy as np
    import torch

    import __graft_entry__ as g
    from cases import multi_isotope_traps_3d

    g.build()
    t0 = time.time()
    m, iso, trapped = multi_isotope_traps_3d(args.n)
    m.initialise()
    setup_s = time.time() - t0
    s = m.solver
    ctx, lib, spec = s.ctx, s.ctx.lib, s.spec
    N, nnz = ctx.n_dofs, ctx.nnz
    nnzb = len(ctx.dmesh.graph[1])
    mesh = spec.mesh
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] \
        if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    B_spmv = 8 * spec.npairs * nnzb + 4 * nnzb + 4 * (mesh.num_vertices + 1) + 16 * N
    d = mesh.tdim
    B_asm = (mesh.num_cells * (4 * (d + 1) + 4) + mesh.num_vertices * (8 * d + 8 + 16 * spec.ns)
             + 8 * spec.npairs * nnzb + 8 * N)
    s._push_state()
    rng = np.random.default_rng(0)
    u = ctx.to_device("u", rng.uniform(0.5, 1.5, N) * 1e18)
    un = ctx.to_device("un", rng.uniform(0.5, 1.5, N) * 1e18)
    F, J, x, y = ctx.zeros(N), ctx.zeros(nnz), ctx.to_device("x", rng.standard_normal(N)), ctx.zeros(N)

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    ms_load = timed(lambda: ctx._check(lib.fb_update_load(ctx.h)), 2)
    ms_asm = timed(lambda: ctx._check(lib.fb_assemble(ctx.h, ctx.ptr(u), ctx.ptr(un), ctx.ptr(F), ctx.ptr(J), 3)), args.reps)
    ms_asm_F = timed(lambda: ctx._check(lib.fb_assemble(ctx.h, ctx.ptr(u), ctx.ptr(un), ctx.ptr(F), ctx.ptr(J), 1)), args.reps)
    ms_spmv = timed(lambda: ctx._check(lib.fb_spmv(ctx.h, ctx.ptr(J), ctx.ptr(x), ctx.ptr(y))), 20)
    y_bsr = y.clone()
    # alternate between two solutions: at an unchanged solution the library streams the stored channels instead
    u_alt, tog = [u, u * (1.0 + 1e-9)], [0]

    def asm_full():
        tog[0] ^= 1
        ctx._check(lib.fb_assemble(ctx.h, ctx.ptr(u_alt[tog[0]]), ctx.ptr(un), ctx.ptr(F), None, 3))
    ms_asm_ch = timed(asm_full, args.reps)
    ms_asm_ch_same = timed(lambda: ctx._check(lib.fb_assemble(ctx.h, ctx.ptr(u), ctx.ptr(un), ctx.ptr(F), None, 3)), args.reps)
    ms_spmv_ch = timed(lambda: ctx._check(lib.fb_spmv(ctx.h, None, ctx.ptr(x), ctx.ptr(y))), 20)
    spmv_diff = float((y - y_bsr).abs().max() / y_bsr.abs().max())
    # one step of the Chebyshev recurrence (MODE 2 of the same kernel) and the block-Jacobi set-up
    ms_bj = timed(lambda: ctx._check(lib.fb_bjacobi_setup(ctx.h, None)), 5)
    cr, cz, cd = ctx.zeros(N), ctx.zeros(N), ctx.zeros(N)
    bufs = [x, cd]

    def cheb_step():
        ctx._check(lib.fb_chebyshev_step(ctx.h, ctx.ptr(bufs[0]), ctx.ptr(cr), ctx.ptr(cz), ctx.ptr(bufs[1]), 0.25, 0.05))
        bufs.reverse()
    ms_cheb = timed(cheb_step, 20)
    vx = spec.vx
    nsp, ndp = (vx["nstat"] + 1) // 2 * 2, (vx["ndyn"] + 1) // 2 * 2
    B_spmv_ch = 8 * (nsp + ndp) * nnzb + 4 * nnzb + 4 * (mesh.num_vertices + 1) + 16 * N
    out = {
        "config": f"C2: 3D H/D/T transient, 3 trap types, n={args.n}", "cells": mesh.num_cells,
        "dofs": N, "species": spec.ns, "species_pairs": spec.npairs, "nnz_blocks": nnzb,
        "jacobian_bytes": 8 * nnz, "host_setup_s": setup_s, "jit": ctx.jit, "hbm_peak_gbs": peak,
        "spmv": {"ms": ms_spmv, "algorithmic_bytes": B_spmv, "gbs": B_spmv / ms_spmv / 1e6,
                 "frac": B_spmv / ms_spmv / 1e6 / peak},
        "assembly_F_and_J": {"ms": ms_asm, "algorithmic_bytes": B_asm, "gbs": B_asm / ms_asm / 1e6,
                             "frac": B_asm / ms_asm / 1e6 / peak},
        "assembly_F_only": {"ms": ms_asm_F},
        "channel_operator": {"static_channels": vx["nstat"], "dynamic_channels": vx["ndyn"],
                             "edge_tensors": len(vx["tensors"]),
                             "assembly_F_and_channels_ms": ms_asm_ch, "assembly_same_solution_ms": ms_asm_ch_same,
                             "spmv_ms": ms_spmv_ch, "spmv_algorithmic_bytes": B_spmv_ch,
                             "spmv_gbs": B_spmv_ch / ms_spmv_ch / 1e6, "spmv_frac": B_spmv_ch / ms_spmv_ch / 1e6 / peak,
                             "spmv_rel_diff_vs_block_csr": spmv_diff,
                             "chebyshev_step_ms": ms_cheb, "bjacobi_setup_ms": ms_bj}, "load_and_coefficients": {"ms": ms_load},
    }
    if args.solve:
        m.u.x.array[:] = 0.0
        m.u_n.x.array[:] = 0.0
        m.iterate()                      # first step: Krylov workspace allocation (3 GB), warm-up
        first_ms = [float(v) for v in s.timings]
        first_lin = s.solver.linear_its
        m.iterate()                      # Chebyshev interval known from the first step's Ritz values
        second_ms = [float(v) for v in s.timings]
        second_lin = s.solver.linear_its
        t0 = time.time()
        m.iterate()
        info = (C.c_double * 4)()
        lib.fb_chebyshev_info(ctx.h, info)
        out["chebyshev"] = {"valid": info[0], "lmin": info[1], "lmax": info[2], "applications": info[3],
                            "first_step_linear_its": first_lin, "second_step_linear_its": second_lin,
                            "second_step_phase_ms": second_ms}
        out["time_step"] = {"wall_s": time.time() - t0, "newton_its": s.solver.its,
                            "first_step_phase_ms": first_ms,
                            "linear_its": s.solver.linear_its,
                            "phase_ms": dict(zip(["assemble_F", "assemble_J", "linear_solve", "update"],
                                                 [float(v) for v in s.timings]))}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
