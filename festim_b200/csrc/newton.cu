// Newton driver, adaptive step-size rule and derived-quantity reductions.
//
// fb_newton_solve restates PETSc SNES "newtonls" with line search "none" as the
// reference configures it (src/festim/problem.py:279-289) together with FESTIM's
// own stopping rule (src/festim/helpers.py:408-426).
#include <cmath>
#include <cstdlib>

#include "fb_internal.cuh"
#include "reduce.cuh"

static int ensure_newton_buffers(fb_ctx* ctx) {
  int rc = fbi_ensure_workspace(ctx);
  if (rc) return rc;
  if (!ctx->d_F) {
    if ((rc = fb_alloc(ctx, &ctx->d_F, (size_t)ctx->S.n_dofs))) return rc;
    if ((rc = fb_alloc(ctx, &ctx->d_y, (size_t)ctx->S.n_dofs))) return rc;
  }

  return 0;
}

__global__ void k_axpy_n(long long n, double a, const double* __restrict__ x, double* __restrict__ y) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    y[i] += a * x[i];
}

__global__ void k_sub(long long n, const double* __restrict__ y, double* __restrict__ u) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    u[i] -= y[i];
}

static int festim_convergence_test(int it, double snorm, double fnorm, double f0, double rtol,
                                   double atol, double stol, int max_it) {
  // reference src/festim/helpers.py:408-426
  if (it > max_it) return FB_DIVERGED_MAX_IT;
  if (fnorm < atol) return FB_CONVERGED_FNORM_ABS;
  if (fnorm / f0 < rtol) return FB_CONVERGED_FNORM_RELATIVE;
  if (snorm < stol && it > 0) return FB_CONVERGED_SNORM_RELATIVE;
  return 0;
}

// Phase timing without a host synchronisation of its own: start / stop events go into a pool in stream order
// and are turned into milliseconds by resolve_phase_timings() after the last synchronisation of the solve.
struct PhaseTimer {
  fb_ctx* ctx;
  size_t idx;
  PhaseTimer(fb_ctx* c, int s) : ctx(c), idx(2 * c->ev_slot.size()) {
    while (ctx->ev_pool.size() < idx + 2) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      ctx->ev_pool.push_back(e);
    }
    ctx->ev_slot.push_back(s);
    cudaEventRecord(ctx->ev_pool[idx], ctx->stream);
  }
  void stop() { cudaEventRecord(ctx->ev_pool[idx + 1], ctx->stream); }
};

static void resolve_phase_timings(fb_ctx* ctx) {
  for (size_t q = 0; q < ctx->ev_slot.size(); ++q) {
    float ms = 0.f;
    if (cudaEventSynchronize(ctx->ev_pool[2 * q + 1]) == cudaSuccess &&
        cudaEventElapsedTime(&ms, ctx->ev_pool[2 * q], ctx->ev_pool[2 * q + 1]) == cudaSuccess)
      ctx->timings[ctx->ev_slot[q]] += ms;
  }
  ctx->ev_slot.clear();
  (void)cudaGetLastError();   // an event of a phase left by an early return was never recorded: not an error of the solve
}

extern "C" int fb_newton_solve(fb_ctx* ctx, double* u, const double* u_n, double atol, double rtol,
                               double stol, int max_it, int ksp, int pc, double ksp_rtol,
                               int ksp_max_it, int* n_its, int* reason, double* fnorm_out,
                               int* total_linear_its) {
  if (!ctx || !ctx->have_spec) return ctx ? fb_fail(ctx, "fb_set_spec first") : -1;
  cudaSetDevice(ctx->device);
  if (ctx->n_bc > 0 && !ctx->S.bc_vals) return fb_fail(ctx, "Dirichlet values not set");
  int rc = ensure_newton_buffers(ctx);
  if (rc) return rc;
  const long long N = ctx->S.n_dofs;
  for (int i = 0; i < 4; ++i) ctx->timings[i] = 0.0;
  ctx->ev_slot.clear();
  if (!u_n) u_n = u;
  if (ksp == FB_KSP_AUTO) {
    if (ctx->tridiag) ksp = FB_KSP_TRIDIAG;
    else ksp = ctx->symmetric ? FB_KSP_CG : FB_KSP_GMRES;
  }
  // channel operator: the residual pass leaves the solution-dependent channel values behind, so
  // the Jacobian costs nothing extra and the Krylov solver reads the operator in place (J = null)
  const bool chan = ctx->op_channels && ksp == FB_KSP_GMRES;
  if (!chan && !ctx->d_J && (rc = fb_alloc(ctx, &ctx->d_J, (size_t)ctx->nnzb * ctx->S.npairs))) return rc;
  double* Jv = chan ? nullptr : ctx->d_J;
  int it = 0, lin_total = 0, why = 0;
  double snorm = 0.0, f0 = 1.0, fn = 0.0;
  const bool guess_ok = ctx->S.transient && ctx->S.inv_dt > 0.0 && ksp == FB_KSP_GMRES && !getenv("FB_NO_KSP_GUESS");
  if (guess_ok && !ctx->d_yprev) {
    if ((rc = fb_alloc(ctx, &ctx->d_yprev, (size_t)3 * N))) return rc;
    ctx->yprev_n = 0;
  }
  // partitioned mesh: ghost entries of the fields come from their owners over peer memory
  if (u_n != u && (rc = fbi_halo(ctx, (double*)u_n))) return rc;
  while (true) {
    if ((rc = fbi_halo(ctx, u))) return rc;
    {
      PhaseTimer t(ctx, it == 0 ? 1 : 0);
      rc = fbi_assemble(ctx, u, u_n, ctx->d_F, Jv, (it == 0 && !chan) ? (FB_ASSEMBLE_F | FB_ASSEMBLE_J) : FB_ASSEMBLE_F);
      if (rc) return rc;
      t.stop();
    }
    if ((rc = fbi_norm2(ctx, ctx->d_F, &fn))) return rc;
    if (!(fn == fn) || std::isinf(fn)) { why = FB_DIVERGED_FNORM_NAN; break; }
    if (it == 0) f0 = fn;
    if (f0 == 0.0) { why = FB_CONVERGED_FNORM_ABS; break; }
    why = festim_convergence_test(it, snorm, fn, f0, rtol, atol, stol, max_it);
    if (why) break;
    if (it >= max_it) { why = FB_DIVERGED_MAX_IT; break; }
    if (it > 0 && !chan) {
      PhaseTimer t(ctx, 1);
      rc = fbi_assemble(ctx, u, u_n, nullptr, Jv, FB_ASSEMBLE_J);
      if (rc) return rc;
      t.stop();
    }
    int lin_its = 0;
    double relres = 0.0;
    {
      PhaseTimer t(ctx, 2);
      // transient runs: the first Newton update of a step is close to that of the previous step
      // (scaled by the ratio of the time steps): start the Krylov solve from it
      // (polynomial extrapolation through the last 1-3 updates while the time step is unchanged)
      const double dt_now = 1.0 / ctx->S.inv_dt;
      const bool guess = guess_ok && it == 0 && ctx->yprev_n > 0 && ctx->yprev_dt > 0.0;
      if (guess) {
        const bool same_dt = std::fabs(dt_now - ctx->yprev_dt) <= 1e-12 * dt_now;
        const int order = same_dt ? ctx->yprev_n : 1;
        static const double w[3][3] = {{1.0, 0.0, 0.0}, {2.0, -1.0, 0.0}, {3.0, -3.0, 1.0}};
        FB_CHECK(ctx, cudaMemsetAsync(ctx->d_y, 0, sizeof(double) * N, ctx->stream));
        long long g = (ctx->S.n_owned_dofs + 255) / 256;
        if (g > ctx->sm_count * 8) g = ctx->sm_count * 8;
        for (int q = 0; q < order; ++q) {
          const double wq = same_dt ? w[order - 1][q] : dt_now / ctx->yprev_dt;
          FB_LAUNCH(ctx, k_axpy_n, (unsigned)g, 256, 0, ctx->S.n_owned_dofs, wq,
                    ctx->d_yprev + (size_t)((ctx->yprev_head + 3 - q) % 3) * N, ctx->d_y);
        }
      }
      ctx->use_guess = guess;
      rc = fbi_linear_solve(ctx, Jv, ctx->d_F, ctx->d_y, ksp, pc, ksp_rtol, 0.0, ksp_max_it, &lin_its, &relres);
      ctx->use_guess = false;
      if (guess_ok && it == 0 && rc == 0) {
        if (ctx->yprev_n > 0 && std::fabs(dt_now - ctx->yprev_dt) > 1e-12 * dt_now) ctx->yprev_n = 0;   // new step size: restart the history
        ctx->yprev_head = (ctx->yprev_head + 1) % 3;
        FB_CHECK(ctx, cudaMemcpyAsync(ctx->d_yprev + (size_t)ctx->yprev_head * N, ctx->d_y, sizeof(double) * N,
                                      cudaMemcpyDeviceToDevice, ctx->stream));
        if (ctx->yprev_n < 3) ctx->yprev_n++;
        ctx->yprev_dt = dt_now;
      }
      t.stop();
    }
    lin_total += lin_its;
    if (getenv("FB_VERBOSE"))
      fprintf(stderr, "[fb] newton it=%d fnorm=%.6e (f0=%.3e) linear its=%d relres=%.3e rc=%d\n", it, fn, f0,
              lin_its, relres, rc);
    if (rc < 0) return rc;
    if (rc > 0) { why = FB_DIVERGED_LINEAR_SOLVE; break; }
    {
      PhaseTimer t(ctx, 3);
      const long long No = ctx->S.n_owned_dofs;
      long long g = (No + 255) / 256;
      if (g > ctx->sm_count * 8) g = ctx->sm_count * 8;
      FB_LAUNCH(ctx, k_sub, (unsigned)g, 256, 0, No, ctx->d_y, u);
      if ((rc = fbi_norm2(ctx, ctx->d_y, &snorm))) return rc;
      t.stop();
    }
    ++it;
  }
  FB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  FB_CHECK(ctx, cudaGetLastError());
  resolve_phase_timings(ctx);
  if (ctx->peer_ready) {   // a peer that never answered leaves the abort flag behind
    unsigned aborted = 0;
    FB_CHECK(ctx, cudaMemcpy(&aborted, ctx->d_ticket + 128, sizeof(unsigned), cudaMemcpyDeviceToHost));
    if (aborted) {
      FB_CHECK(ctx, cudaMemset(ctx->d_ticket + 128, 0, sizeof(unsigned)));
      return fb_fail(ctx, "partitioned solve: a peer rank never answered (exchange timed out)");
    }
  }
  if (n_its) *n_its = it;
  if (reason) *reason = why;
  if (fnorm_out) *fnorm_out = fn;
  if (total_linear_its) *total_linear_its = lin_total;
  return 0;
}

extern "C" int fb_last_timings(fb_ctx* ctx, double* out) {
  if (!ctx) return -1;
  for (int i = 0; i < 4; ++i) out[i] = ctx->timings[i];
  return 0;
}

// reference src/festim/stepsize.py:144-167
extern "C" double fb_stepsize_modify(double value, int nb_iterations, double t, int adaptive,
                                     double growth_factor, double cutback_factor,
                                     int target_nb_iterations, double max_stepsize, int n_milestones,
                                     const double* milestones, double milestone_tolerance) {
  if (!adaptive) return value;
  double updated = value;
  if (nb_iterations < target_nb_iterations) updated = value * growth_factor;
  else if (nb_iterations > target_nb_iterations) updated = value * cutback_factor;
  if (max_stepsize > 0.0 && updated > max_stepsize) updated = max_stepsize;
  for (int i = 0; i < n_milestones; ++i) {
    if (t < milestones[i]) {
      const double to_go = milestones[i] - t;
      const bool close = std::fabs(t - milestones[i]) <= milestone_tolerance * std::fabs(milestones[i]);
      if (updated > to_go && !close) updated = to_go;
      break;
    }
  }
  return updated;
}

// Device-resident implicit-Euler time stepping (reference src/festim/problem.py:195-255 with the
// step-size rule of src/festim/stepsize.py:144-167): the whole loop "t += dt; Newton solve;
// u_n <- u; adapt dt" runs inside the library, the fields never leave the device and no Python runs
// between steps.  Valid while nothing but t and dt changes from step to step (time-independent
// boundary values, sources and temperature, no per-step exports): the caller passes the time at
// which it has to look again (next milestone / export time / final time) as t_stop.  Like the
// reference, the last step is not clamped to t_stop unless a milestone says so.
extern "C" int fb_run_steps(fb_ctx* ctx, double* u, double* u_n, double* t, double* dt, double t_stop,
                            int max_steps, double atol, double rtol, double stol, int max_it, int ksp, int pc,
                            double ksp_rtol, int ksp_max_it, int adaptive, double growth_factor,
                            double cutback_factor, int target_nb_iterations, double max_stepsize,
                            int n_milestones, const double* milestones, double milestone_tolerance,
                            int* n_steps, double* times, int* newton_its, int* total_linear_its, int* reason) {
  if (!ctx || !ctx->have_spec) return ctx ? fb_fail(ctx, "fb_set_spec first") : -1;
  if (!ctx->S.transient || ctx->dt_scalar < 0) return fb_fail(ctx, "fb_run_steps needs a transient problem");
  if (!u || !u_n || u == u_n || !t || !dt) return fb_fail(ctx, "fb_run_steps: bad arguments");
  cudaSetDevice(ctx->device);
  int rc = fbi_ensure_workspace(ctx);
  if (rc) return rc;
  int steps = 0, lin_total = 0, why = 1;
  while (*t < t_stop && steps < max_steps) {
    if (times) times[steps] = *t;
    *t += *dt;
    // the time step is a scalar of the compiled expressions and the 1/dt of the mass terms
    ctx->S.inv_dt = 1.0 / *dt;
    ctx->h_pinned[16] = *dt;
    FB_CHECK(ctx, cudaMemcpyAsync(ctx->d_scalars + ctx->dt_scalar, ctx->h_pinned + 16, sizeof(double),
                                  cudaMemcpyHostToDevice, ctx->stream));
    int its = 0, lin = 0;
    double fn = 0.0;
    rc = fb_newton_solve(ctx, u, u_n, atol, rtol, stol, max_it, ksp, pc, ksp_rtol, ksp_max_it, &its, &why, &fn, &lin);
    if (rc) return rc;
    if (newton_its) newton_its[steps] = its;
    lin_total += lin;
    ++steps;
    if (why <= 0) break;   // the caller raises, as the reference does (problem.py:238-241)
    FB_CHECK(ctx, cudaMemcpyAsync(u_n, u, sizeof(double) * ctx->S.n_dofs, cudaMemcpyDeviceToDevice, ctx->stream));
    *dt = fb_stepsize_modify(*dt, its, *t, adaptive, growth_factor, cutback_factor, target_nb_iterations,
                             max_stepsize, n_milestones, milestones, milestone_tolerance);
  }
  FB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  if (n_steps) *n_steps = steps;
  if (total_linear_its) *total_linear_its = lin_total;
  if (reason) *reason = why;
  return 0;
}

// ---------------------------------------------------------------------------
// derived quantities
// ---------------------------------------------------------------------------
template <int DIM>
__device__ __forceinline__ double cell_volume(const DevSpec& S, long long c, int* a) {
  double X[FB_MAX_D1][3];
  for (int j = 0; j <= DIM; ++j) {
    a[j] = S.cells[c * (DIM + 1) + j];
    for (int k = 0; k < DIM; ++k) X[j][k] = S.coords[(long long)a[j] * DIM + k];
  }
  if (DIM == 1) return fabs(X[1][0] - X[0][0]);
  if (DIM == 2) {
    return 0.5 * fabs((X[1][0] - X[0][0]) * (X[2][1] - X[0][1]) - (X[1][1] - X[0][1]) * (X[2][0] - X[0][0]));
  }
  double e[3][3];
  for (int k = 0; k < 3; ++k)
    for (int c2 = 0; c2 < 3; ++c2) e[k][c2] = X[k + 1][c2] - X[0][c2];
  const double det = e[0][0] * (e[1][1] * e[2][2] - e[1][2] * e[2][1]) -
                     e[0][1] * (e[1][0] * e[2][2] - e[1][2] * e[2][0]) +
                     e[0][2] * (e[1][0] * e[2][1] - e[1][1] * e[2][0]);
  return fabs(det) / 6.0;
}

// int u dx over cells tagged vtag: sum |K| mean(u_vertices)   (degree-1 exact)
template <int DIM>
__global__ void __launch_bounds__(FB_TPB)
k_total_volume(DevSpec S, const double* __restrict__ u, int species, int vtag, double* partials,
               unsigned* ticket, double* out) {
  double v[1] = {0.0}, tot[1];
  for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < S.n_cells;
       c += (long long)gridDim.x * blockDim.x) {
    if (S.cell_tag[c] != vtag) continue;
    int a[FB_MAX_D1];
    const double vol = cell_volume<DIM>(S, c, a);
    double m = 0.0;
    for (int j = 0; j <= DIM; ++j) m += u[(long long)a[j] * S.ns + species];
    v[0] += vol * m / (DIM + 1);
  }
  if (grid_reduce<1>(v, partials, ticket, tot) && threadIdx.x == 0) out[0] = tot[0];
}

// mode 0: int u ds ; mode 1: int -D_h grad(u).n ds with D_h the DG1 interpolant of
// D0 exp(-ED/(kB T_D)) (reference hydrogen_transport_problem.py:632-670)
template <int DIM>
__global__ void __launch_bounds__(FB_TPB)
k_surface(DevSpec S, const double* __restrict__ u, int species, int stag, int mode, double TD_const,
          const double* __restrict__ TD_nodal, double* partials, unsigned* ticket, double* out) {
  double v[1] = {0.0}, tot[1];
  for (long long f = blockIdx.x * (long long)blockDim.x + threadIdx.x; f < S.n_bfacets;
       f += (long long)gridDim.x * blockDim.x) {
    if (S.bfacet_tag[f] != stag) continue;
    int fv[3];
    double Y[3][3];
    for (int j = 0; j < DIM; ++j) {
      fv[j] = S.bfacets[f * DIM + j];
      for (int k = 0; k < DIM; ++k) Y[j][k] = S.coords[(long long)fv[j] * DIM + k];
    }
    double area = 1.0;
    if (DIM == 2) area = sqrt((Y[1][0] - Y[0][0]) * (Y[1][0] - Y[0][0]) + (Y[1][1] - Y[0][1]) * (Y[1][1] - Y[0][1]));
    if (DIM == 3) {
      const double ux = Y[1][0] - Y[0][0], uy = Y[1][1] - Y[0][1], uz = Y[1][2] - Y[0][2];
      const double vx = Y[2][0] - Y[0][0], vy = Y[2][1] - Y[0][1], vz = Y[2][2] - Y[0][2];
      const double cx = uy * vz - uz * vy, cy = uz * vx - ux * vz, cz = ux * vy - uy * vx;
      area = 0.5 * sqrt(cx * cx + cy * cy + cz * cz);
    }
    if (mode == 0) {
      double m = 0.0;
      for (int j = 0; j < DIM; ++j) m += u[(long long)fv[j] * S.ns + species];
      v[0] += area * m / DIM;
      continue;
    }
    const long long c = S.bfacet_cell[f];
    const int tag = S.cell_tag[c];
    int a[FB_MAX_D1];
    double X[FB_MAX_D1][3];
    int opp = 0;
    for (int j = 0; j <= DIM; ++j) {
      a[j] = S.cells[c * (DIM + 1) + j];
      for (int k = 0; k < DIM; ++k) X[j][k] = S.coords[(long long)a[j] * DIM + k];
      bool onf = false;
      for (int q = 0; q < DIM; ++q) onf |= (fv[q] == a[j]);
      if (!onf) opp = j;
    }
    // basis gradients: reuse the closed forms through finite differences of the
    // affine basis: solve via the same formulas as the assembly kernels
    double G[FB_MAX_D1][3];
    if (DIM == 1) {
      const double J = X[1][0] - X[0][0];
      G[1][0] = 1.0 / J; G[0][0] = -1.0 / J;
    } else if (DIM == 2) {
      const double e1x = X[1][0] - X[0][0], e1y = X[1][1] - X[0][1];
      const double e2x = X[2][0] - X[0][0], e2y = X[2][1] - X[0][1];
      const double id = 1.0 / (e1x * e2y - e1y * e2x);
      G[1][0] = e2y * id; G[1][1] = -e2x * id; G[2][0] = -e1y * id; G[2][1] = e1x * id;
      G[0][0] = -(G[1][0] + G[2][0]); G[0][1] = -(G[1][1] + G[2][1]);
    } else {
      double e[3][3];
      for (int k = 0; k < 3; ++k)
        for (int c2 = 0; c2 < 3; ++c2) e[k][c2] = X[k + 1][c2] - X[0][c2];
      double c0[3] = {e[1][1] * e[2][2] - e[1][2] * e[2][1], e[1][2] * e[2][0] - e[1][0] * e[2][2], e[1][0] * e[2][1] - e[1][1] * e[2][0]};
      double c1[3] = {e[2][1] * e[0][2] - e[2][2] * e[0][1], e[2][2] * e[0][0] - e[2][0] * e[0][2], e[2][0] * e[0][1] - e[2][1] * e[0][0]};
      double c2v[3] = {e[0][1] * e[1][2] - e[0][2] * e[1][1], e[0][2] * e[1][0] - e[0][0] * e[1][2], e[0][0] * e[1][1] - e[0][1] * e[1][0]};
      const double id = 1.0 / (e[0][0] * c0[0] + e[0][1] * c0[1] + e[0][2] * c0[2]);
      for (int k = 0; k < 3; ++k) {
        G[1][k] = c0[k] * id; G[2][k] = c1[k] * id; G[3][k] = c2v[k] * id;
        G[0][k] = -(G[1][k] + G[2][k] + G[3][k]);
      }
    }
    double gu[3] = {0, 0, 0}, gn = 0.0;
    for (int k = 0; k < DIM; ++k) {
      for (int j = 0; j <= DIM; ++j) gu[k] += u[(long long)a[j] * S.ns + species] * G[j][k];
      gn += G[opp][k] * G[opp][k];
    }
    gn = sqrt(gn);
    double dudn = 0.0;  // outward normal = -G[opp]/|G[opp]|
    for (int k = 0; k < DIM; ++k) dudn -= gu[k] * G[opp][k] / gn;
    const int ds = S.diff_slot[tag * S.ns + species];
    double Dm = 0.0;
    if (ds >= 0 && ds < S.n_diffE) {
      const double D0 = S.diff_D0[tag * S.ns + species], E = S.diff_energies[ds];
      for (int j = 0; j < DIM; ++j) {
        const double T = TD_nodal ? TD_nodal[fv[j]] : TD_const;
        Dm += D0 * exp(-E / (FB_KB * T));
      }
      Dm /= DIM;
    }
    v[0] += area * Dm * (-dudn);
  }
  if (grid_reduce<1>(v, partials, ticket, tot) && threadIdx.x == 0) out[0] = tot[0];
}

static int fetch_scalar(fb_ctx* ctx, double* out) {
  FB_CHECK(ctx, cudaMemcpyAsync(ctx->h_pinned, ctx->d_sc + 61, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  FB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  FB_CHECK(ctx, cudaGetLastError());
  *out = ctx->h_pinned[0];
  return 0;
}

static unsigned red_grid(fb_ctx* ctx, long long n) {
  long long g = (n + FB_TPB - 1) / FB_TPB;
  if (g > ctx->sm_count * 8) g = ctx->sm_count * 8;
  return (unsigned)(g < 1 ? 1 : g);
}

extern "C" int fb_total_volume(fb_ctx* ctx, const double* u, int species, int vtag, double* out) {
  if (!ctx || !ctx->have_spec) return ctx ? fb_fail(ctx, "fb_set_spec first") : -1;
  cudaSetDevice(ctx->device);
  int rc = fbi_ensure_workspace(ctx);
  if (rc) return rc;
  const DevSpec& S = ctx->S;
  if (species < 0 || species >= S.ns) return fb_fail(ctx, "species out of range");
  const unsigned g = red_grid(ctx, S.n_cells);
  switch (S.tdim) {
    case 1: FB_LAUNCH(ctx, k_total_volume<1>, g, FB_TPB, 0, S, u, species, vtag, ctx->d_red, ctx->d_ticket, ctx->d_sc + 61); break;
    case 2: FB_LAUNCH(ctx, k_total_volume<2>, g, FB_TPB, 0, S, u, species, vtag, ctx->d_red, ctx->d_ticket, ctx->d_sc + 61); break;
    default: FB_LAUNCH(ctx, k_total_volume<3>, g, FB_TPB, 0, S, u, species, vtag, ctx->d_red, ctx->d_ticket, ctx->d_sc + 61); break;
  }
  return fetch_scalar(ctx, out);
}

static int surface_quantity(fb_ctx* ctx, const double* u, int species, int stag, int mode, double Tc,
                            const double* Tn, double* out) {
  if (!ctx || !ctx->have_spec) return ctx ? fb_fail(ctx, "fb_set_spec first") : -1;
  cudaSetDevice(ctx->device);
  int rc = fbi_ensure_workspace(ctx);
  if (rc) return rc;
  const DevSpec& S = ctx->S;
  if (species < 0 || species >= S.ns) return fb_fail(ctx, "species out of range");
  const unsigned g = red_grid(ctx, S.n_bfacets);
  switch (S.tdim) {
    case 1: FB_LAUNCH(ctx, k_surface<1>, g, FB_TPB, 0, S, u, species, stag, mode, Tc, Tn, ctx->d_red, ctx->d_ticket, ctx->d_sc + 61); break;
    case 2: FB_LAUNCH(ctx, k_surface<2>, g, FB_TPB, 0, S, u, species, stag, mode, Tc, Tn, ctx->d_red, ctx->d_ticket, ctx->d_sc + 61); break;
    default: FB_LAUNCH(ctx, k_surface<3>, g, FB_TPB, 0, S, u, species, stag, mode, Tc, Tn, ctx->d_red, ctx->d_ticket, ctx->d_sc + 61); break;
  }
  return fetch_scalar(ctx, out);
}

extern "C" int fb_total_surface(fb_ctx* ctx, const double* u, int species, int stag, double* out) {
  return surface_quantity(ctx, u, species, stag, 0, 0.0, nullptr, out);
}

extern "C" int fb_surface_flux(fb_ctx* ctx, const double* u, int species, int stag, double T_D_const,
                               const double* T_D_nodal, double* out) {
  return surface_quantity(ctx, u, species, stag, 1, T_D_const, T_D_nodal, out);
}
