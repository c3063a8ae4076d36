// Internal declarations of libfestim_b200 (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/festim_b200.h"

#include "fb_device.cuh"

struct fb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  long long launches = 0;
  int sm_count = 148;
  DevSpec S{};
  bool have_mesh = false, have_pattern = false, have_spec = false;
  long long nnzb = 0, n_bc = 0;
  bool tridiag = false;
  // per-SM compact partition (fb_set_sm_partition)
  std::vector<int> h_rowptr, h_colidx;
  const unsigned short* d_halo_blk = nullptr;   // block id of every halo vertex
  // neighbour blocks of every block (CSR) and the block entries that couple to each of them
  const int* d_cn_ptr = nullptr; const int* d_cn_blk = nullptr;
  const int* d_cn_eptr = nullptr; const int* d_cn_ent = nullptr;
  double* d_coarse = nullptr;                   // E (nb*nb), E^-1 (nb*nb), rc (nb)
  unsigned short* d_halo_blk_eff = nullptr;     // halo block ids with Dirichlet vertices masked out
  int part_nhalo = 0;
  bool coarse_masks_dirty = true;
  bool coarse_ready = false;                    // E^-1 matches the matrix being solved
  int part_blocks = 0;
  bool part_owned_only = false;  // the blocks cover the owned vertices of a partitioned mesh only
  const int* d_blk_rng = nullptr;
  const int* d_halo_ptr = nullptr;
  const int* d_halo_idx = nullptr;
  const unsigned short* d_lcol = nullptr;
  long long part_max_entries = 0, part_max_bentries = 0, part_max_su = 0, part_max_rows = 0, part_min_rows = 0;
  // peer-memory exchange area of the partitioned persistent CG (fb_peer_*)
  double* d_xchg = nullptr;            // u64 words: [0,128) reduction slots, [128,..) two per ghost dof
  long long xchg_u_doubles = 0;
  void* peer_base[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int peer_rank = 0, peer_n = 0;
  unsigned peer_seq = 1;               // next unused tag of the LL words
  unsigned xr_seq = 1, halo_seq = 0x40000001u;   // tags of the cross-rank reductions / halo exchanges of the Krylov kernels
  long long halo_words = 0;            // u64 words of one parity of the halo area
  bool peer_ready = false;             // every peer area mapped and the send table set
  const int* d_send_v = nullptr; const int* d_send_rank = nullptr; const int* d_send_dst = nullptr;
  int n_send = 0;
  int n_programs = 0;                  // compiled expressions of the spec (fb_integrate checks its argument)
  bool symmetric = false;
  int n_sp_ent = 0;            // entries of the channel SpMV epilogue table (vx_sp)
  // lane plan of the TMA channel SpMV's row epilogue (spmv_channels.cuh: fb_spmv_ep_plan); epl 0 = none
  int ep_epl = 0;
  // block-Jacobi fill table of the channel operator, grouped by channel: ptr [nch+1], pos | flag << 16 [n], coef [n]
  const int* d_bj_ptr = nullptr; const int* d_bj_pos = nullptr; const double* d_bj_coef = nullptr;
  const double* d_ep_coef = nullptr; const int* d_ep_pos = nullptr; const int* d_ep_lane = nullptr;
  bool op_channels = false;
  // Chebyshev polynomial preconditioner: interval of D^-1 J from Ritz values, operator applications
  bool cheb_valid = false, cheb_refused = false;
  double cheb_lmin = 0.0, cheb_lmax = 0.0, cheb_inv_dt = 0.0;
  int cheb_steps = 0;    // Krylov solvers use the channel operator (J == nullptr) instead of block-CSR values
  bool facet_u_terms = false;  // any solution-dependent facet term
  bool generic_u_vterms = false;  // any solution-dependent volume term evaluated point-wise
  bool have_jit = false;
  void* jit_module = nullptr;
  void* jit_fn[4] = {nullptr, nullptr, nullptr, nullptr};
  int dt_scalar = -1;
  std::vector<void*> owned;  // device allocations freed in fb_destroy
  // mutable device buffers
  double* d_scalars = nullptr;
  const double** d_fields = nullptr;
  const double** d_velocities = nullptr;
  int* d_bc_map = nullptr;
  double* d_dbar = nullptr;
  double* d_load = nullptr;
  std::vector<const double*> h_fields, h_velocities;
  // solver workspace
  double* d_F = nullptr;
  double* d_J = nullptr;
  double* d_y = nullptr;
  double* d_yprev = nullptr;   // first Newton updates of the last three time steps (ring): initial guess of the next one
  int yprev_n = 0, yprev_head = 0;
  bool use_guess = false;
  double guess_guard = 0.0;
  double yprev_dt = 0.0;
  double* d_work = nullptr;  // Krylov vectors
  long long work_doubles = 0;
  double* d_dinv = nullptr;  // block-Jacobi inverse blocks [n_vertices][ns][ns]
  // the same blocks as single-precision mantissas with one power-of-two scale per row (k_bjacobi_setup_slab):
  // inverse[v][s][c] = dscale[v][s] * dinv32[v][s][c] holds exactly - d_dinv carries these rounded values
  // the solution the stored solution-dependent channel values (et_dyn) were computed from: an assembly at the
  // same solution (the first of a time step, after the residual check of the last one) streams them instead
  // of contracting the edge tensors again (k_assemble_edge_tma, reuse mode)
  double* d_u_saved = nullptr;
  bool dyn_saved = false;
  int* d_mismatch = nullptr;
  std::vector<int> fac_scalar_idx;       // scalars the affine reaction factors read
  std::vector<double> fac_scalar_last;
  float* d_dinv32 = nullptr;
  double* d_dscale = nullptr;
  bool dinv32_valid = false;
  double* d_red = nullptr;   // reduction partials
  long long red_doubles = 0;
  double* d_sc = nullptr;    // device scalars of the Krylov loops
  unsigned* d_ticket = nullptr;
  int* d_flags = nullptr;    // done flag, iteration counter
  double* h_pinned = nullptr;
  double* d_tri = nullptr;   // dense block-tridiagonal workspace
  double timings[4] = {0, 0, 0, 0};
  cudaEvent_t ev[8];
  // phase timing of fb_newton_solve without host synchronisation: event pairs recorded in stream order,
  // turned into milliseconds after the solve's last (already needed) synchronisation
  std::vector<cudaEvent_t> ev_pool;
  std::vector<int> ev_slot;
};

#define FB_CHECK(ctx, call)                                                                 \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess) {                                                                \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                      \
      return -2;                                                                            \
    }                                                                                       \
  } while (0)

#define FB_LAUNCH(ctx, kernel, grid, block, smem, ...)                                      \
  do {                                                                                      \
    kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                        \
    (ctx)->launches++;                                                                      \
  } while (0)

static inline int fb_fail(fb_ctx* ctx, const char* msg) {
  ctx->err = msg;
  return -1;
}

template <typename T>
int fb_alloc(fb_ctx* ctx, T** ptr, size_t count) {
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, (count ? count : 1) * sizeof(T));
  if (e != cudaSuccess) {
    ctx->err = std::string("cudaMalloc: ") + cudaGetErrorString(e);
    return -2;
  }
  ctx->owned.push_back(p);
  *ptr = (T*)p;
  return 0;
}

template <typename T>
int fb_upload(fb_ctx* ctx, const T** dst, const T* host, size_t count) {
  T* p = nullptr;
  int rc = fb_alloc(ctx, &p, count);
  if (rc) return rc;
  if (count) {
    cudaError_t e = cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      ctx->err = std::string("upload: ") + cudaGetErrorString(e);
      return -2;
    }
  }
  *dst = p;
  return 0;
}

// entry points implemented across translation units
int fbi_update_coeffs(fb_ctx* ctx);
int fbi_update_vinfo(fb_ctx* ctx);
int fbi_et_symbolic(fb_ctx* ctx);
int fbi_expand(fb_ctx* ctx, double* J);
int fbi_assemble(fb_ctx* ctx, const double* u, const double* un, double* F, double* J, int flags);
int fbi_spmv(fb_ctx* ctx, const double* J, const double* x, double* y);
int fbi_norm2(fb_ctx* ctx, const double* x, double* out_host);
int fbi_linear_solve(fb_ctx* ctx, const double* J, const double* b, double* y, int ksp, int pc,
                     double rtol, double atol, int max_it, int* iters, double* relres);
int fbi_ensure_workspace(fb_ctx* ctx);
int fbi_halo(fb_ctx* ctx, double* x);
