// Residual / Jacobian assembly kernels (fp64, P1 simplices).
//
// Replaces the FFCx tabulate_tensor kernels + dolfinx assemble_vector /
// assemble_matrix / apply_lifting / set_bc that the reference reaches through
// NonlinearProblem (reference src/festim/problem.py:170-182, 237), for the form
// of src/festim/hydrogen_transport_problem.py:870-997.
//
// Design (B200-first, not a dolfinx port):
//  * GATHER instead of scatter.  One thread owns one mesh vertex (= one block row
//    of J and ns entries of F) and walks the cells around it through the
//    vertex->cell adjacency.  Every J entry is written exactly once by its owner:
//    no fp64 atomics, no colouring, no memset pass, bitwise reproducible, and the
//    same kernel assembles only the owned rows of a partitioned mesh (no reverse
//    scatter across GPUs).
//  * Arrhenius factors never run inside the Newton loop: the cell means
//    sum_q w_q exp(-E/(kB T_q)) for the F and the J quadrature rule are cached
//    per temperature update by k_cell_coeffs (one exp per distinct activation
//    energy and point), so a Newton iteration of a diffusion problem is pure
//    geometry + loads.
//  * Strong Dirichlet rows/columns are handled entry by entry while gathering
//    (lifting with the live Jacobian entry, identity row), so F and J leave the
//    kernel already in the dolfinx NonlinearProblem form.
//  * Solution-independent integrands (sources, prescribed fluxes) are integrated
//    once per update into a load vector by cell/facet-parallel kernels.
#include <cuda.h>
#include <dlfcn.h>

#include <cstdint>
#include <cstdlib>

#include "fb_internal.cuh"
#include "program.cuh"
#include "assemble_kernels.cuh"
#include "assemble_edge.cuh"
#include "assemble_edge_tma.cuh"
#include "reduce.cuh"

// ---------------------------------------------------------------------------
// kernels: thin wrappers around the bodies of assemble_kernels.cuh
// ---------------------------------------------------------------------------
template <int DIM>
__global__ void k_cell_coeffs(DevSpec S, double* __restrict__ dbar) { fb_cell_coeffs_body<DIM>(S, dbar); }

template <int DIM>
__global__ void k_load_cells(DevSpec S, double* __restrict__ load) { fb_load_cells_body<DIM>(S, load); }

template <int DIM>
__global__ void k_facets(DevSpec S, int mode, const double* __restrict__ u, const double* __restrict__ un,
                         double* __restrict__ F, double* __restrict__ J) {
  fb_facets_body<DIM>(S, mode, u, un, F, J);
}

template <int DIM>
__global__ void __launch_bounds__(128)
k_assemble_nodes(DevSpec S, const double* __restrict__ u, const double* __restrict__ un,
                 double* __restrict__ F, double* __restrict__ J, int flags) {
  fb_assemble_nodes_body<DIM>(S, u, un, F, J, flags);
}

template <int DIM>
__global__ void __launch_bounds__(128)
k_cell_records(DevSpec S, double* __restrict__ out) { fb_cell_records_body<DIM>(S, out); }

__global__ void k_vinfo_rows(DevSpec S, unsigned* __restrict__ vinfo) { fb_vinfo_rows_body(S, vinfo); }
__global__ void k_vinfo_cols(DevSpec S, unsigned* __restrict__ vinfo) { fb_vinfo_cols_body(S, vinfo); }

template <int DIM>
__global__ void k_et_symbolic(DevSpec S, int mode, int* __restrict__ width, const long long* __restrict__ tzptr,
                              unsigned char* __restrict__ tzslot, int* __restrict__ err) {
  fb_et_symbolic_body<DIM>(S, mode, width, tzptr, tzslot, err);
}

template <int DIM>
__global__ void __launch_bounds__(128)
k_et_setup(DevSpec S, double* __restrict__ stat, double* __restrict__ tvw, double* __restrict__ tz) {
  fb_et_setup_body<DIM>(S, stat, tvw, tz);
}

template <int G>
__global__ void __launch_bounds__(256)
k_assemble_edge(DevSpec S, const double* __restrict__ u, const double* __restrict__ un, double* __restrict__ F,
                double* __restrict__ dyn, int flags) {
  extern __shared__ double fb_et_smem[];
  fb_assemble_edge_body<G>(S, u, un, F, dyn, flags, fb_et_smem);
}

template <int G>
__global__ void __launch_bounds__(256)
k_et_expand(DevSpec S, double* __restrict__ J) {
  extern __shared__ double fb_et_smem[];
  fb_et_expand_body<G>(S, J, fb_et_smem);
}

// ---------------------------------------------------------------------------
// host side.  When a run-time specialised module is loaded (fb_jit_load) its
// kernels - same bodies, expressions compiled natively instead of interpreted -
// take the place of the ahead-of-time ones.
// ---------------------------------------------------------------------------
enum { JIT_CELL_COEFFS = 0, JIT_LOAD_CELLS, JIT_FACETS, JIT_ASSEMBLE_NODES, JIT_COUNT };

// The driver API is resolved at run time (dlopen) so the library still loads on a
// machine without a GPU driver (build check, symbol tests).
namespace {
struct DriverApi {
  CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
  CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                           CUstream, void**, void**) = nullptr;
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
  bool ok = false;
};
DriverApi& driver() {
  static DriverApi d;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (h) {
      d.ModuleLoadData = (decltype(d.ModuleLoadData))dlsym(h, "cuModuleLoadData");
      d.ModuleGetFunction = (decltype(d.ModuleGetFunction))dlsym(h, "cuModuleGetFunction");
      d.LaunchKernel = (decltype(d.LaunchKernel))dlsym(h, "cuLaunchKernel");
      d.GetErrorString = (decltype(d.GetErrorString))dlsym(h, "cuGetErrorString");
      d.ok = d.ModuleLoadData && d.ModuleGetFunction && d.LaunchKernel && d.GetErrorString;
    }
  }
  return d;
}
}  // namespace

static int jit_launch(fb_ctx* ctx, int which, unsigned grid, unsigned block, void** args) {
  CUresult r = driver().LaunchKernel((CUfunction)ctx->jit_fn[which], grid, 1, 1, block, 1, 1, 0,
                              (CUstream)ctx->stream, args, nullptr);
  ctx->launches++;
  if (r != CUDA_SUCCESS) {
    const char* msg = nullptr;
    driver().GetErrorString(r, &msg);
    ctx->err = std::string("cuLaunchKernel (jit): ") + (msg ? msg : "?");
    return -2;
  }
  return 0;
}

extern "C" int fb_jit_load(fb_ctx* ctx, const void* cubin, size_t nbytes) {
  if (!ctx || !ctx->have_spec) return ctx ? fb_fail(ctx, "fb_set_spec first") : -1;
  (void)nbytes;
  cudaSetDevice(ctx->device);
  cudaFree(0);  // make sure the primary context is current for the driver API
  if (!driver().ok) return fb_fail(ctx, "CUDA driver API (libcuda.so.1) not available");
  CUmodule mod;
  CUresult r = driver().ModuleLoadData(&mod, cubin);
  if (r != CUDA_SUCCESS) {
    const char* msg = nullptr;
    driver().GetErrorString(r, &msg);
    ctx->err = std::string("cuModuleLoadData: ") + (msg ? msg : "?");
    return -2;
  }
  static const char* names[JIT_COUNT] = {"jit_cell_coeffs", "jit_load_cells", "jit_facets", "jit_assemble_nodes"};
  for (int i = 0; i < JIT_COUNT; ++i) {
    CUfunction f;
    if (driver().ModuleGetFunction(&f, mod, names[i]) != CUDA_SUCCESS) {
      ctx->err = std::string("jit module lacks ") + names[i];
      return -2;
    }
    ctx->jit_fn[i] = (void*)f;
  }
  ctx->jit_module = (void*)mod;
  ctx->have_jit = true;
  return 0;
}

template <int DIM>
static int update_coeffs_dim(fb_ctx* ctx) {
  DevSpec& S = ctx->S;
  const unsigned B = 128;
  const unsigned gc = (unsigned)((S.n_cells + B - 1) / B);
  if (S.n_slots > 0) {
    if (ctx->have_jit) {
      void* args[] = {(void*)&S, (void*)&ctx->d_dbar};
      int rc = jit_launch(ctx, JIT_CELL_COEFFS, gc, B, args);
      if (rc) return rc;
    } else {
      FB_LAUNCH(ctx, k_cell_coeffs<DIM>, gc, B, 0, S, ctx->d_dbar);
    }
  }
  if (S.vx_on) {
    FB_LAUNCH(ctx, k_cell_records<DIM>, gc, B, 0, S, (double*)S.cellrec);
    FB_LAUNCH(ctx, k_et_setup<DIM>, (unsigned)((S.n_owned + B - 1) / B), B, 0, S, (double*)S.et_stat, (double*)S.et_tvw,
              (double*)S.et_tz);
  }
  FB_CHECK(ctx, cudaMemsetAsync(ctx->d_load, 0, sizeof(double) * S.n_dofs, ctx->stream));
  if (S.n_vterms > 0) {
    if (ctx->have_jit) {
      void* args[] = {(void*)&S, (void*)&ctx->d_load};
      int rc = jit_launch(ctx, JIT_LOAD_CELLS, gc, B, args);
      if (rc) return rc;
    } else {
      FB_LAUNCH(ctx, k_load_cells<DIM>, gc, B, 0, S, ctx->d_load);
    }
  }
  if (S.n_fterms > 0 && S.n_bfacets > 0) {
    const unsigned gf = (unsigned)((S.n_bfacets + B - 1) / B);
    int mode = 0;
    const double* nul = nullptr;
    double* nulw = nullptr;
    if (ctx->have_jit) {
      void* args[] = {(void*)&S, (void*)&mode, (void*)&nul, (void*)&nul, (void*)&ctx->d_load, (void*)&nulw};
      int rc = jit_launch(ctx, JIT_FACETS, gf, B, args);
      if (rc) return rc;
    } else {
      FB_LAUNCH(ctx, k_facets<DIM>, gf, B, 0, S, 0, nul, nul, ctx->d_load, nulw);
    }
  }
  FB_CHECK(ctx, cudaGetLastError());
  return 0;
}

int fbi_update_vinfo(fb_ctx* ctx) {
  const DevSpec& S = ctx->S;
  if (!S.vx_on) return 0;
  const unsigned g = (unsigned)((S.n_vertices + 255) / 256);
  FB_LAUNCH(ctx, k_vinfo_rows, g, 256, 0, S, (unsigned*)S.vinfo);
  FB_LAUNCH(ctx, k_vinfo_cols, g, 256, 0, S, (unsigned*)S.vinfo);
  FB_CHECK(ctx, cudaGetLastError());
  return 0;
}

int fbi_update_coeffs(fb_ctx* ctx) {
  ctx->dyn_saved = false;    // edge tensors rebuilt: the stored solution-dependent channels are stale
  ctx->cheb_valid = false;   // temperature / coefficients moved: re-estimate the spectrum interval
  ctx->cheb_refused = false;
  ctx->yprev_n = 0;          // the initial guess of the Krylov solve comes from another problem
  switch (ctx->S.tdim) {
    case 1: return update_coeffs_dim<1>(ctx);
    case 2: return update_coeffs_dim<2>(ctx);
    case 3: return update_coeffs_dim<3>(ctx);
  }
  return fb_fail(ctx, "bad tdim");
}

// lanes per vertex of the row kernels: the smallest power of two that covers the longest row
static int et_group(const DevSpec& S) {
  int G = 4;
  while (G < 32 && G < S.maxdeg) G *= 2;
  if (getenv("FB_ET_G")) G = atoi(getenv("FB_ET_G"));
  return G;
}

// which = 0: k_assemble_edge, 1: k_et_expand
static int et_launch(fb_ctx* ctx, int which, const double* u, const double* un, double* F, double* J, int flags) {
  DevSpec& S = ctx->S;
  const EtLayout L = fb_et_layout(S, which == 0);
  const int G = et_group(S);
  int smem_max = 0;
  cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device);
  int tpb = 256;
  if (getenv("FB_ET_TPB")) tpb = atoi(getenv("FB_ET_TPB"));
  while (tpb > G && (size_t)(L.grp0 + (size_t)(tpb / G) * L.per_grp) * 8 > (size_t)smem_max) tpb /= 2;
  const int gpb = tpb / G;
  const size_t smem = (size_t)(L.grp0 + (size_t)gpb * L.per_grp) * 8;
  if (gpb < 1 || smem > (size_t)smem_max) return fb_fail(ctx, "channel assembly: a block row does not fit in shared memory");
  const unsigned grid = (unsigned)((S.n_owned + gpb - 1) / gpb);
  double* dyn = (double*)S.et_dyn;
#define FB_ET_CASE(GG)                                                                                          \
  if (which == 0) {                                                                                             \
    FB_CHECK(ctx, cudaFuncSetAttribute((const void*)k_assemble_edge<GG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k_assemble_edge<GG><<<grid, tpb, smem, ctx->stream>>>(S, u, un, F, dyn, flags);                             \
  } else {                                                                                                      \
    FB_CHECK(ctx, cudaFuncSetAttribute((const void*)k_et_expand<GG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k_et_expand<GG><<<grid, tpb, smem, ctx->stream>>>(S, J);                                                     \
  }
  switch (G) {
    case 4: FB_ET_CASE(4) break;
    case 8: FB_ET_CASE(8) break;
    case 16: FB_ET_CASE(16) break;
    default: FB_ET_CASE(32) break;
  }
#undef FB_ET_CASE
  ctx->launches++;
  FB_CHECK(ctx, cudaGetLastError());
  return 0;
}

// TMA-pipelined Newton-iteration assembly (assemble_edge_tma.cuh): every bulk copy 16-byte aligned
static bool asm_tma_ok(fb_ctx* ctx, const double* u, const double* un) {
  const DevSpec& S = ctx->S;
  return S.vx_on && (S.ns % 2 == 0) && S.ns <= 32 && (((uintptr_t)u & 15) == 0) && (((uintptr_t)un & 15) == 0) &&
         S.ck_nx_max > 0 && S.vx_nts <= 11 && S.maxdeg <= 16 && S.vx_nch <= 16 && S.vx_nbw <= 64 &&
         !getenv("FB_NO_TMA_ASSEMBLY");   // rows of up to 16 edges: both contractions run as DMMA tiles
}

// *mismatch = 1 if the two vectors differ anywhere (caller zeroes it)
__global__ void k_vec_differs(long long n, const double* __restrict__ a, const double* __restrict__ b, int* mismatch) {
  bool diff = false;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x)
    diff |= !(a[k] == b[k]) && !(a[k] != a[k] && b[k] != b[k]);   // NaNs in the same place count as equal
  if (diff) *mismatch = 1;   // every writer stores the same value
}

static int asm_tma_launch(fb_ctx* ctx, const double* u, const double* un, double* F, int flags, bool* done) {
  DevSpec& S = ctx->S;
  *done = false;
  AsmTmaArgs a{};
  a.u = u; a.un = un; a.F = F; a.dyn = (double*)S.et_dyn; a.flags = flags;
  // reuse mode: the solution-dependent channels in memory were computed from this very solution (checked
  // on the device, entry by entry): stream them instead of the edge tensors
  a.mismatch = nullptr;
  if (S.vx_ndyn > 0 && S.vx_nts > 0 && !getenv("FB_NO_DYN_REUSE")) {
    if (!ctx->d_u_saved) {
      int rc0 = fb_alloc(ctx, &ctx->d_u_saved, (size_t)S.n_dofs);
      if (!rc0) rc0 = fb_alloc(ctx, &ctx->d_mismatch, 4);
      if (rc0) return rc0;
    }
    if (ctx->dyn_saved) {
      FB_CHECK(ctx, cudaMemsetAsync(ctx->d_mismatch, 0, sizeof(int), ctx->stream));
      long long g = (S.n_dofs + 4 * FB_TPB - 1) / (4 * FB_TPB);
      if (g > ctx->sm_count * 8) g = ctx->sm_count * 8;
      FB_LAUNCH(ctx, k_vec_differs, (unsigned)(g < 1 ? 1 : g), FB_TPB, 0, S.n_dofs, u, (const double*)ctx->d_u_saved, ctx->d_mismatch);
      a.mismatch = ctx->d_mismatch;
    }
  }
  int smem_max = 0;
  cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device);
  int nst = 3 * FB_CHUNK_GROUPS;
  if (getenv("FB_ASM_STAGES")) nst = atoi(getenv("FB_ASM_STAGES"));
  if (nst > 8) nst = 8;
  size_t smem = fb_asm_tma_layout(S, a, nst);
  while (nst > FB_CHUNK_GROUPS + 1 && smem > (size_t)smem_max) smem = fb_asm_tma_layout(S, a, --nst);
  if (smem > (size_t)smem_max) return 0;   // a chunk does not fit: the caller takes the row kernel
  const long long nchunks = (S.n_owned + FB_CHUNK - 1) / FB_CHUNK;
  long long grid = ctx->sm_count;
  if (grid > nchunks) grid = nchunks;
  if (grid < 1) grid = 1;
  FB_CHECK(ctx, cudaFuncSetAttribute((const void*)k_assemble_edge_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_assemble_edge_tma<<<(unsigned)grid, 32 * FB_CHUNK * FB_CHUNK_GROUPS + 64, smem, ctx->stream>>>(S, a);
  ctx->launches++;
  FB_CHECK(ctx, cudaGetLastError());
  if (ctx->d_u_saved) {   // et_dyn now belongs to this solution
    FB_CHECK(ctx, cudaMemcpyAsync(ctx->d_u_saved, u, sizeof(double) * (size_t)S.n_dofs, cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->dyn_saved = true;
  }
  *done = true;
  return 0;
}

int fbi_expand(fb_ctx* ctx, double* J) {
  if (!ctx->S.vx_on) return fb_fail(ctx, "no channel operator to expand");
  return et_launch(ctx, 1, nullptr, nullptr, nullptr, J, 0);
}

// structure of the edge tensor (once per mesh + spec): widths, prefix sum on the host, slots
template <int DIM>
static int et_symbolic_dim(fb_ctx* ctx) {
  DevSpec& S = ctx->S;
  const long long nv = S.n_vertices;
  int* d_w = nullptr;
  int* d_err = nullptr;
  int rc;
  const size_t nw = ((size_t)nv + FB_CHUNK + 16) & ~(size_t)15;   // whole chunks of widths are bulk-copied
  if ((rc = fb_alloc(ctx, &d_w, nw + 4))) return rc;
  d_err = d_w + nw;
  FB_CHECK(ctx, cudaMemsetAsync(d_w, 0, sizeof(int) * (nw + 4), ctx->stream));
  const unsigned g = (unsigned)((nv + 127) / 128);
  const long long keep_owned = S.n_owned;
  S.n_owned = nv;
  FB_LAUNCH(ctx, k_et_symbolic<DIM>, g, 128, 0, S, 0, d_w, (const long long*)nullptr, (unsigned char*)nullptr, d_err);
  std::vector<int> w(nw + 4);
  FB_CHECK(ctx, cudaMemcpyAsync(w.data(), d_w, sizeof(int) * (nw + 4), cudaMemcpyDeviceToHost, ctx->stream));
  FB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  if (w[nw]) { S.n_owned = keep_owned; return fb_fail(ctx, "edge tensor: more than 64 common neighbours on an edge"); }
  // slabs padded to multiples of 16 entries: every chunk's extent is a 16-byte aligned bulk copy
  std::vector<long long> ptr((size_t)nv + 1 + FB_CHUNK + 8, 0);
  for (long long v = 0; v < nv; ++v)
    ptr[v + 1] = ptr[v] + (((long long)w[v] * (ctx->h_rowptr[v + 1] - ctx->h_rowptr[v]) + 15) & ~15LL);
  for (size_t v = (size_t)nv + 1; v < ptr.size(); ++v) ptr[v] = ptr[nv];
  S.et_tz_total = ptr[nv];
  S.et_tz_chunk_max = 0;
  for (long long c0 = 0; c0 < nv; c0 += FB_CHUNK) {
    const long long c1 = c0 + FB_CHUNK < nv ? c0 + FB_CHUNK : nv;
    if (ptr[c1] - ptr[c0] > S.et_tz_chunk_max) S.et_tz_chunk_max = ptr[c1] - ptr[c0];
  }
  S.et_tzw = d_w;
  if ((rc = fb_upload(ctx, &S.et_tzptr, ptr.data(), ptr.size()))) return rc;
  unsigned char* slots = nullptr;
  if ((rc = fb_alloc(ctx, &slots, (size_t)S.et_tz_total + 1))) return rc;
  FB_CHECK(ctx, cudaMemsetAsync(slots, 0xff, (size_t)S.et_tz_total + 1, ctx->stream));
  FB_LAUNCH(ctx, k_et_symbolic<DIM>, g, 128, 0, S, 1, d_w, S.et_tzptr, slots, d_err);
  S.et_tzslot = slots;
  double* tz = nullptr;
  if ((rc = fb_alloc(ctx, &tz, (size_t)S.et_tz_total * S.vx_nts + 1))) return rc;
  S.et_tz = tz;
  S.n_owned = keep_owned;
  FB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  FB_CHECK(ctx, cudaGetLastError());
  return 0;
}

int fbi_et_symbolic(fb_ctx* ctx) {
  switch (ctx->S.tdim) {
    case 1: return et_symbolic_dim<1>(ctx);
    case 2: return et_symbolic_dim<2>(ctx);
    case 3: return et_symbolic_dim<3>(ctx);
  }
  return fb_fail(ctx, "bad tdim");
}

template <int DIM>
static int assemble_dim(fb_ctx* ctx, const double* u, const double* un, double* F, double* J, int flags) {
  DevSpec& S = ctx->S;
  const unsigned B = 128;
  const unsigned gn = (unsigned)((S.n_vertices + B - 1) / B);
  if (S.vx_on) {
    // channel assembly (assemble_edge.cuh): dynamic channel values + residual, then - when the
    // caller wants the block-CSR values - the expansion of the channel operator
    bool done = false;
    int rc = 0;
    if (asm_tma_ok(ctx, u, un) && (rc = asm_tma_launch(ctx, u, un, F, flags, &done))) return rc;
    if (!done) ctx->dyn_saved = false;   // the row kernel rewrites et_dyn without recording the solution
    if (!done && (rc = et_launch(ctx, 0, u, un, F, nullptr, flags))) return rc;
    if ((flags & FB_ASSEMBLE_J) && J && (rc = et_launch(ctx, 1, nullptr, nullptr, nullptr, J, 0))) return rc;
  } else if (ctx->have_jit) {
    void* args[] = {(void*)&S, (void*)&u, (void*)&un, (void*)&F, (void*)&J, (void*)&flags};
    int rc = jit_launch(ctx, JIT_ASSEMBLE_NODES, gn, B, args);
    if (rc) return rc;
  } else {
    FB_LAUNCH(ctx, k_assemble_nodes<DIM>, gn, B, 0, S, u, un, F, J, flags);
  }
  if (S.n_fterms > 0 && S.n_bfacets > 0 && ctx->facet_u_terms) {
    const unsigned gf = (unsigned)((S.n_bfacets + B - 1) / B);
    int mode = 1;
    double* Fp = (flags & FB_ASSEMBLE_F) ? F : nullptr;
    double* Jp = (flags & FB_ASSEMBLE_J) ? J : nullptr;
    if (ctx->have_jit) {
      void* args[] = {(void*)&S, (void*)&mode, (void*)&u, (void*)&un, (void*)&Fp, (void*)&Jp};
      int rc = jit_launch(ctx, JIT_FACETS, gf, B, args);
      if (rc) return rc;
    } else {
      FB_LAUNCH(ctx, k_facets<DIM>, gf, B, 0, S, 1, u, un, Fp, Jp);
    }
  }
  FB_CHECK(ctx, cudaGetLastError());
  return 0;
}

int fbi_assemble(fb_ctx* ctx, const double* u, const double* un, double* F, double* J, int flags) {
  if (!(ctx->have_mesh && ctx->have_pattern && ctx->have_spec)) return fb_fail(ctx, "context not fully defined");
  if ((flags & FB_ASSEMBLE_F) && !F) return fb_fail(ctx, "F is null");
  if ((flags & FB_ASSEMBLE_J) && !J && !ctx->S.vx_on) return fb_fail(ctx, "J is null");
  if (!un) un = u;
  switch (ctx->S.tdim) {
    case 1: return assemble_dim<1>(ctx, u, un, F, J, flags);
    case 2: return assemble_dim<2>(ctx, u, un, F, J, flags);
    case 3: return assemble_dim<3>(ctx, u, un, F, J, flags);
  }
  return fb_fail(ctx, "bad tdim");
}

// ---------------------------------------------------------------------------
// User-defined derived quantity (reference exports/custom_quantity.py:117-132): the integral of
// a compiled expression of x, t, T, the species, their cell gradients and the facet normal
// over one volume (kind 0) or surface (kind 1) subdomain.  The expression is one of the
// problem's programs (byte-code: an export, not a hot kernel), the rule is given by its
// offset in the quadrature pool.
// ---------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(FB_TPB)
k_integrate(DevSpec S, const double* __restrict__ u, int kind, int tag, int prog, int qoff, int nq,
            double* partials, unsigned* ticket, double* out) {
  constexpr int D1 = DIM + 1;
  const int ns = S.ns;
  const long long n_ent = kind == 0 ? S.n_cells : S.n_bfacets;
  double v[1] = {0.0}, tot[1];
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n_ent;
       e += (long long)gridDim.x * blockDim.x) {
    if ((kind == 0 ? S.cell_tag[e] : S.bfacet_tag[e]) != tag) continue;
    const long long c = kind == 0 ? e : S.bfacet_cell[e];
    int a[FB_MAX_D1], ev[FB_MAX_D1];
    double X[FB_MAX_D1][3], G[FB_MAX_D1][3], Gflat[FB_MAX_D1 * 3], Y[FB_MAX_D1][3];
    fb_load_cell<DIM>(S, c, a, X);
    const double vol = fb_geometry<DIM>(X, G);
    for (int j = 0; j < D1; ++j)
      for (int k = 0; k < DIM; ++k) Gflat[j * DIM + k] = G[j][k];
    const int nev = kind == 0 ? D1 : DIM;   // vertices of the integration entity
    double measure = vol, nrm[3] = {0.0, 0.0, 0.0};
    if (kind == 0) {
      for (int j = 0; j < D1; ++j) {
        ev[j] = a[j];
        for (int k = 0; k < DIM; ++k) Y[j][k] = X[j][k];
      }
    } else {
      for (int j = 0; j < DIM; ++j) {
        ev[j] = S.bfacets[e * DIM + j];
        for (int k = 0; k < DIM; ++k) Y[j][k] = S.coords[(long long)ev[j] * DIM + k];
      }
      measure = fb_facet_measure<DIM>(Y);
      int opp = 0;
      for (int j = 0; j < D1; ++j) {
        bool onf = false;
        for (int k = 0; k < DIM; ++k) onf |= (a[j] == ev[k]);
        if (!onf) opp = j;
      }
      double l = 0.0;
      for (int k = 0; k < DIM; ++k) l += G[opp][k] * G[opp][k];
      l = sqrt(l);
      for (int k = 0; k < DIM; ++k) nrm[k] = -G[opp][k] / l;
    }
    double cgr[FB_MAX_NS * 3], cq[FB_MAX_NS], zero[FB_MAX_NS];
    for (int s = 0; s < ns; ++s) {
      zero[s] = 0.0;
      for (int k = 0; k < 3; ++k) {
        double gk = 0.0;
        if (k < DIM)
          for (int j = 0; j < D1; ++j) gk += u[(long long)a[j] * ns + s] * G[j][k];
        cgr[s * 3 + k] = gk;
      }
    }
    const double hcell = fb_circumradius<DIM>(X);
    double acc = 0.0;
    for (int q = 0; q < nq; ++q) {
      const double* bary = S.quad_bary + (long long)(qoff + q) * D1;
      PointCtx P;
      for (int k = 0; k < DIM; ++k) {
        double x = 0.0;
        for (int j = 0; j < nev; ++j) x += bary[j] * Y[j][k];
        P.x[k] = x;
      }
      P.T = fb_temperature(S, ev, bary, nev);
      for (int s = 0; s < ns; ++s) {
        double cv = 0.0;
        for (int j = 0; j < nev; ++j) cv += bary[j] * u[(long long)ev[j] * ns + s];
        cq[s] = cv;
      }
      P.c = cq; P.cprev = zero; P.verts = ev; P.bary = bary; P.nverts = nev;
      P.cverts = a; P.G = Gflat; P.ncverts = D1; P.tdim = DIM;
      P.cgrad = cgr; P.n[0] = nrm[0]; P.n[1] = nrm[1]; P.n[2] = nrm[2]; P.h = hcell;
      acc += S.quad_w[qoff + q] * fb_prog_eval(S, prog, P);
    }
    v[0] += acc * measure;
  }
  if (grid_reduce<1>(v, partials, ticket, tot) && threadIdx.x == 0) out[0] = tot[0];
}

extern "C" int fb_integrate(fb_ctx* ctx, const double* u, int kind, int tag, int prog, int quad_offset, int nq,
                            double* out) {
  if (!ctx || !ctx->have_spec) return ctx ? fb_fail(ctx, "fb_set_spec first") : -1;
  if (!u || !out) return fb_fail(ctx, "null argument");
  cudaSetDevice(ctx->device);
  int rc = fbi_ensure_workspace(ctx);
  if (rc) return rc;
  const DevSpec& S = ctx->S;
  if (kind < 0 || kind > 1 || prog < 0 || prog >= ctx->n_programs || nq < 1) return fb_fail(ctx, "bad integrand");
  const long long n_ent = kind == 0 ? S.n_cells : S.n_bfacets;
  long long g = (n_ent + FB_TPB - 1) / FB_TPB;
  if (g > ctx->sm_count * 8) g = ctx->sm_count * 8;
  if (g < 1) g = 1;
  switch (S.tdim) {
    case 1: FB_LAUNCH(ctx, k_integrate<1>, (unsigned)g, FB_TPB, 0, S, u, kind, tag, prog, quad_offset, nq, ctx->d_red, ctx->d_ticket, ctx->d_sc + 61); break;
    case 2: FB_LAUNCH(ctx, k_integrate<2>, (unsigned)g, FB_TPB, 0, S, u, kind, tag, prog, quad_offset, nq, ctx->d_red, ctx->d_ticket, ctx->d_sc + 61); break;
    default: FB_LAUNCH(ctx, k_integrate<3>, (unsigned)g, FB_TPB, 0, S, u, kind, tag, prog, quad_offset, nq, ctx->d_red, ctx->d_ticket, ctx->d_sc + 61); break;
  }
  FB_CHECK(ctx, cudaMemcpyAsync(ctx->h_pinned, ctx->d_sc + 61, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  FB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  FB_CHECK(ctx, cudaGetLastError());
  *out = ctx->h_pinned[0];
  return 0;
}
