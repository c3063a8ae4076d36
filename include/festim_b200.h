/*
 * festim_b200.h -- C ABI of libfestim_b200.so, the B200 (sm_100a) solver backend
 * for FESTIM's hot path: residual/Jacobian assembly + Newton-Krylov solve of
 * coupled diffusion-trapping-reaction systems on P1 interval/triangle/tet meshes.
 *
 * FESTIM has no plugin API; the replaceable seam is
 *   ProblemBase.create_solver()            reference src/festim/problem.py:170-193
 *   self.solver.solve()                    reference src/festim/problem.py:213,237
 *   solver.solver.getConvergedReason()     reference src/festim/problem.py:238
 *   solver.solver.getIterationNumber()     reference src/festim/problem.py:242
 * Behind that seam the reference crosses into dolfinx (C++ assembly of FFCx
 * kernels) and PETSc (SNES newtonls + KSP).  The entry points below are what a
 * ctypes binding placed at that seam calls instead.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; fb_last_error(ctx) gives
 *     the message (owned by ctx).  Non-convergence is NOT an error: it is reported
 *     through `reason` (PETSc SNESConvergedReason numbering).
 *   - pointers marked [host] are read during the call and copied; pointers marked
 *     [dev] are device pointers owned by the caller (e.g. torch tensors) that must
 *     stay valid until the next call that replaces them or fb_destroy.
 *   - all reals are fp64, local indices int32, sizes int64.
 *   - unknown layout: node-major blocked, u[v * ns + s].
 *   - Jacobian layout ("species-sparse BSR"): node-graph CSR (rowptr, colidx) of
 *     blocks; block row v stores, for each row species s, deg(v) * nc[s] values
 *     ordered [k][j]: value(v,s,k,j) at
 *       rowptr[v]*npairs + deg(v)*pp[s] + k*nc[s] + j ,
 *     the entry d F_(v,s) / d u_(colidx[rowptr[v]+k], cs[pp[s]+j]).
 *   - one host thread per ctx; all work is enqueued on the stream given to
 *     fb_create.
 */
#ifndef FESTIM_B200_H
#define FESTIM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fb_ctx fb_ctx;

/* assemble flags */
#define FB_ASSEMBLE_F 1
#define FB_ASSEMBLE_J 2

/* linear solver kinds (maps PETSc ksp_type / pc_type, reference problem.py:279-289) */
#define FB_KSP_AUTO 0     /* preonly+lu: exact block-tridiagonal LU in 1D, Krylov(tight) else */
#define FB_KSP_CG 1
#define FB_KSP_GMRES 2
#define FB_KSP_TRIDIAG 3
#define FB_PC_JACOBI 0    /* point-block Jacobi (ns x ns diagonal blocks) */
#define FB_PC_CHEBYSHEV 1 /* GMRES right-preconditioned with a Chebyshev polynomial of the block-Jacobi-scaled operator
                             (interval from the Ritz values of a first block-Jacobi cycle); channel operator only,
                             block-Jacobi otherwise */
#define FB_PC_AUTO 2      /* what pc_type "lu" maps to: Chebyshev for large systems held as channel operator, else Jacobi */

/* SNES converged reasons used (PETSc numbering) */
#define FB_CONVERGED_FNORM_ABS 2
#define FB_CONVERGED_FNORM_RELATIVE 3
#define FB_CONVERGED_SNORM_RELATIVE 4
#define FB_DIVERGED_LINEAR_SOLVE -3
#define FB_DIVERGED_FNORM_NAN -4
#define FB_DIVERGED_MAX_IT -5

/* -- lifetime ------------------------------------------------------------------ */
int fb_create(int device, void* cuda_stream, fb_ctx** out);
void fb_destroy(fb_ctx* ctx);
const char* fb_last_error(fb_ctx* ctx);
/* number of kernels this ctx has launched so far (bench.py "gpu_launches") */
int64_t fb_launch_count(fb_ctx* ctx);

/* -- problem definition (replaces dolfinx mesh/functionspace/form compilation,
 *    reference hydrogen_transport_problem.py:672-728, problem.py:173-182) ---------- */
int fb_set_mesh(fb_ctx* ctx, int tdim, int64_t n_vertices, int64_t n_cells,
                const double* coords /*[host] n_vertices*tdim*/,
                const int32_t* cells /*[host] n_cells*(tdim+1)*/,
                const int32_t* cell_tag /*[host] n_cells, compact 0..n_vtags-1*/,
                int64_t n_bfacets, const int32_t* bfacets /*[host] n_bfacets*tdim*/,
                const int32_t* bfacet_cell /*[host]*/,
                const int32_t* bfacet_tag /*[host] compact, -1 = untagged*/);
/* sparsity pattern of the vertex graph and vertex->cell adjacency (packed
 * cell*(tdim+1)+local), the analogue of dolfinx's create_sparsity_pattern */
int fb_set_pattern(fb_ctx* ctx, const int32_t* rowptr /*[host] n_vertices+1*/,
                   const int32_t* colidx /*[host] nnzb*/, int64_t nnzb,
                   const int32_t* v2c_ptr /*[host] n_vertices+1*/,
                   const int32_t* v2c_idx /*[host] n_cells*(tdim+1)*/);
/* flat kernel spec produced by festim_b200/lowering.py (layout documented there) */
int fb_set_spec(fb_ctx* ctx, const void* blob /*[host]*/, size_t nbytes);
/* optional: cubin of the element kernels specialised for this spec at run time
 * (NVRTC, festim_b200/jit.py) - the analogue of the FFCx JIT the reference triggers
 * in NonlinearProblem (problem.py:173-182; hydrogen_transport_problem.py:2025-2040).
 * Without it the ahead-of-time kernels interpret the spec's byte-code programs. */
int fb_jit_load(fb_ctx* ctx, const void* cubin /*[host]*/, size_t nbytes);
/* strong Dirichlet dofs (reference problem.py:118-130, fem.dirichletbc) */
int fb_set_dirichlet(fb_ctx* ctx, int64_t n_bc, const int64_t* bc_dofs /*[host]*/);
int64_t fb_num_dofs(fb_ctx* ctx);
int64_t fb_jacobian_nnz(fb_ctx* ctx); /* number of stored scalars = nnzb*npairs */

/* -- per-step data (reference update_time_dependent_values,
 *    hydrogen_transport_problem.py:999-1034) -------------------------------------- */
int fb_set_scalars(fb_ctx* ctx, int n, const double* scalars /*[host]*/);
int fb_set_temperature(fb_ctx* ctx, double T_const, const double* T_nodal /*[dev] or NULL*/);
int fb_set_field(fb_ctx* ctx, int k, const double* nodal /*[dev] n_vertices*/);
int fb_set_velocity(fb_ctx* ctx, int k, const double* nodal /*[dev] n_vertices*tdim*/);
int fb_set_dirichlet_values(fb_ctx* ctx, const double* bc_vals /*[dev] n_bc*/);
/* (re)compute the solution-independent load vector (sources, fluxes); call after
 * any of the setters above changed and before fb_assemble / fb_newton_solve */
int fb_update_load(fb_ctx* ctx);

/* -- the hot path ------------------------------------------------------------------ */
/* residual F(u) and/or Jacobian J(u) with strong Dirichlet conditions applied the
 * dolfinx NonlinearProblem way: lifting (alpha=-1), F_B = u_B - g, identity rows and
 * columns in J (replaces assemble_vector/apply_lifting/set_bc/assemble_matrix). */
/* Systems assembled through channels (mass-action reactions, diffusion, mass, advection: every
 * BASELINE config) keep the Jacobian as the channel operator  J = sum_c A^c (x) B_c  inside the
 * context (8 bytes per channel and edge instead of 8 bytes per species pair and edge): every
 * fb_assemble refreshes it, and Jvals may then be NULL in fb_assemble (no block-CSR values are
 * written), fb_spmv and fb_linear_solve (they use the operator of the last fb_assemble). */
int fb_assemble(fb_ctx* ctx, const double* u /*[dev]*/, const double* u_n /*[dev]*/,
                double* F /*[dev] N or NULL*/, double* Jvals /*[dev] nnz or NULL*/, int flags);
/* y = J x (species-sparse block-CSR values, or the channel operator when Jvals is NULL) */
int fb_spmv(fb_ctx* ctx, const double* Jvals /*[dev]*/, const double* x /*[dev]*/,
            double* y /*[dev]*/);
/* fused vector kernels (warp-shuffle reductions); results land in out[] on host */
int fb_dot(fb_ctx* ctx, const double* x, const double* y, double* out /*[host]*/);
int fb_axpy(fb_ctx* ctx, double a, const double* x, double* y); /* y += a x */
/* solve J y = b; returns iterations and final relative residual */
int fb_linear_solve(fb_ctx* ctx, const double* Jvals /*[dev]*/, const double* b /*[dev]*/,
                    double* y /*[dev]*/, int ksp, int pc, double rtol, double atol, int max_it,
                    int* iters /*[host]*/, double* relres /*[host]*/);
/* full Newton solve (SNES newtonls, line search none) with the FESTIM stopping rule
 * (reference helpers.py:408-426).  u is updated in place. */
int fb_newton_solve(fb_ctx* ctx, double* u /*[dev]*/, const double* u_n /*[dev]*/, double atol,
                    double rtol, double stol, int max_it, int ksp, int pc, double ksp_rtol,
                    int ksp_max_it, int* n_its, int* reason, double* fnorm,
                    int* total_linear_its);
/* Device-resident implicit-Euler stepping (reference problem.py:195-255, stepsize.py:144-167):
 * "t += dt; Newton solve; u_n <- u; adapt dt" repeated inside the library until t >= t_stop, a
 * solve fails or max_steps is reached.  Valid while only t and dt change between steps; u, u_n
 * stay on the device.  times[k], newton_its[k] describe step k; *reason is the last solve's. */
int fb_run_steps(fb_ctx* ctx, double* u /*[dev]*/, double* u_n /*[dev]*/, double* t /*[host] in/out*/,
                 double* dt /*[host] in/out*/, double t_stop, int max_steps, double atol, double rtol,
                 double stol, int max_it, int ksp, int pc, double ksp_rtol, int ksp_max_it, int adaptive,
                 double growth_factor, double cutback_factor, int target_nb_iterations,
                 double max_stepsize /*<=0: none*/, int n_milestones, const double* milestones /*[host]*/,
                 double milestone_tolerance, int* n_steps, double* times /*[host] max_steps or NULL*/,
                 int* newton_its /*[host] max_steps or NULL*/, int* total_linear_its, int* reason);
/* timing of the last fb_newton_solve, CUDA events on ctx's stream, milliseconds:
 * out[0]=assembly F, out[1]=assembly J(+F), out[2]=linear solve, out[3]=other */
int fb_last_timings(fb_ctx* ctx, double* out /*[host] 4*/);

/* state of the Chebyshev preconditioner: out = {valid, lambda_min, lambda_max, operator applications per use} */
int fb_chebyshev_info(fb_ctx* ctx, double* out /*[host] 4*/);
/* test hook: operator applications per Krylov iteration the solver picks for a spectrum interval of D^-1 J and the
 * residual reduction a solve still needs (least expected cost under the Chebyshev bound, >= 3 Krylov iterations,
 * <= 12 applications); *krylov_iterations = the iterations it expects.  Returns the applications, < 0 on bad input. */
int fb_chebyshev_steps(double lambda_min, double lambda_max, double reduction, int* krylov_iterations /*[host]*/);
/* one step of the Chebyshev recurrence on B = D^-1 J (channel operator, after fb_bjacobi_setup(ctx, NULL)),
 * the single kernel launch the polynomial preconditioner repeats (PETSc analogue: one KSPCHEBYSHEV iteration
 * inside PCKSP, reference options seam src/festim/problem.py:132-168):
 *   r <- r - B d_in ;  d_out = ca d_in + cb r ;  z <- z + d_out        (all [dev], n_dofs doubles) */
int fb_chebyshev_step(fb_ctx* ctx, const double* d_in, double* r, double* z, double* d_out, double ca, double cb);
/* test hook: eigenvalues of an n x n real upper Hessenberg matrix (row-major), the routine that
 * turns the Arnoldi matrix of a GMRES cycle into the interval of the Chebyshev preconditioner */
int fb_hessenberg_eigenvalues(int n, const double* a /*[host]*/, double* wr /*[host]*/, double* wi /*[host]*/);

/* adaptive step size (reference stepsize.py:144-167); milestones sorted ascending */
double fb_stepsize_modify(double value, int nb_iterations, double t, int adaptive,
                          double growth_factor, double cutback_factor, int target_nb_iterations,
                          double max_stepsize /* <=0: none */, int n_milestones,
                          const double* milestones, double milestone_tolerance);

/* -- derived quantities (reference exports/total_volume.py:23-33,
 *    exports/surface_flux.py:39-60) ------------------------------------------------- */
int fb_total_volume(fb_ctx* ctx, const double* u /*[dev]*/, int species, int vtag,
                    double* out /*[host]*/);
int fb_total_surface(fb_ctx* ctx, const double* u /*[dev]*/, int species, int stag,
                     double* out /*[host]*/);
/* D_h = DG1 interpolant of D0[vtag]*exp(-ED[vtag]/kB/T_D(vertex)); T_D is the
 * temperature snapshot of the last D_global refresh */
int fb_surface_flux(fb_ctx* ctx, const double* u /*[dev]*/, int species, int stag,
                    double T_D_const, const double* T_D_nodal /*[dev] or NULL*/,
                    double* out /*[host]*/);

/* user-defined derived quantity (reference exports/custom_quantity.py:117-132): integral over
 * one volume (kind 0) or surface (kind 1) subdomain of program `prog` of the spec - an
 * expression of x, t, T, the species, their cell gradients and the facet normal - with the
 * quadrature rule at `quad_offset` (nq points) of the spec's pool */
int fb_integrate(fb_ctx* ctx, const double* u /*[dev]*/, int kind, int tag, int prog, int quad_offset, int nq,
                 double* out /*[host]*/);

/* -- partitioned meshes (one process per GPU; replaces dolfinx IndexMap ghost updates and
 *    PETSc's MPI reductions).  The first n_owned vertices of the local
 *    numbering are owned, the rest are ghosts refreshed by the caller (NCCL) before
 *    fb_assemble / fb_spmv / fb_dcg_spmv_dots; kernels write owned rows only and reductions
 *    cover owned entries only. ---------------------------------------------------------- */
int fb_set_owned(fb_ctx* ctx, int64_t n_owned_vertices);
/* single GPU: compact per-SM vertex blocks (contiguous ranges after renumbering) with their
 * halo lists and 16-bit local columns, so the persistent CG keeps operator rows and the
 * gathered part of the iterate in shared memory */
int fb_set_sm_partition(fb_ctx* ctx, int nblocks, const int32_t* block_ranges /*[host] nblocks+1*/,
                        const int32_t* halo_ptr /*[host] nblocks+1*/, const int32_t* halo_idx /*[host]*/,
                        const uint16_t* local_col /*[host] nnzb*/, int64_t n_local_col);
int fb_bjacobi_setup(fb_ctx* ctx, const double* Jvals /*[dev]*/);
int fb_precond(fb_ctx* ctx, const double* r /*[dev]*/, double* z /*[dev]*/); /* z = M^-1 r */
/* single-reduction CG split at its two communication points */
int fb_dcg_init(fb_ctx* ctx, const double* b, double* x, double* r, double* u, double* p, double* s);
int fb_dcg_spmv_dots(fb_ctx* ctx, const double* Jvals, const double* u, const double* r, double* w,
                     double* out3 /*[dev] local (r,u), (w,u), (u,u)*/);
int fb_dcg_update(fb_ctx* ctx, double alpha, double beta, double* x, double* r, double* u,
                  const double* w, double* p, double* s);
/* the same with the recurrence scalars kept on the device (state[7]: alpha, beta, gamma,
 * iterations, done, threshold^2, last ||M^-1 r||^2), so a whole iteration can be captured in
 * a CUDA graph: no host round trip between the all-reduce and the update */
int fb_dcg_scalars(fb_ctx* ctx, const double* reduced3 /*[dev]*/, double* state /*[dev]*/, int max_it);
int fb_dcg_update_dev(fb_ctx* ctx, const double* state /*[dev]*/, double* x, double* r, double* u,
                      const double* w, double* p, double* s);

/* -- partitioned persistent CG over peer memory (NVLink / NVSwitch, ranks of one node).
 *    Replaces, for the whole Krylov solve, what PETSc does per iteration with
 *    VecScatter ghost updates and MPI_Allreduce (reference: KSPSolve inside
 *    NonlinearProblem.solve, src/festim/problem.py:262-289): one cooperative launch per
 *    rank; boundary entries of the iterate are pushed into the neighbours' memory and
 *    the dot products meet in every rank's exchange area, as self-validating
 *    {tag, half double} words (no fences, flags or NCCL calls inside the solve).
 *    fb_peer_alloc creates this rank's exchange area and returns its CUDA IPC handle
 *    (64 bytes) for the host to distribute; fb_peer_open maps a peer's area;
 *    fb_peer_set_send lists, sorted by vertex, every owned vertex another rank holds as
 *    ghost (dst = its ordinal in that rank's ghost block). --------------------------------------------- */
int fb_peer_alloc(fb_ctx* ctx, int rank, int nranks /*<= 8*/, unsigned char* ipc_handle /*[host] 64 bytes*/);
int fb_peer_open(fb_ctx* ctx, int peer_rank, const unsigned char* ipc_handle /*[host] 64 bytes*/);
int fb_peer_set_send(fb_ctx* ctx, int64_t n, const int32_t* vertex /*[host]*/, const int32_t* rank /*[host]*/,
                     const int32_t* dst_vertex /*[host]*/);
/* solves J x = b on the owned rows; every rank must call it with the same arguments.
 * Preconditioner M: point-block Jacobi (fb_bjacobi_setup) plus, when the rank's owned
 * vertices are grouped in per-SM blocks (fb_set_sm_partition), a rank-local coarse space.
 * Stops at ||M^-1 r|| <= rtol ||M^-1 b|| (thr2 = that bound squared for the Jacobi-only M). */
int fb_dcg_persistent(fb_ctx* ctx, const double* Jvals /*[dev]*/, const double* b /*[dev]*/, double* x /*[dev]*/,
                      double thr2, double rtol, int max_it, int* iters /*[host]*/, int* converged /*[host]*/);

#ifdef __cplusplus
}
#endif
#endif /* FESTIM_B200_H */
