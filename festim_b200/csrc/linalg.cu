// Linear algebra on the species-sparse BSR Jacobian: SpMV, fused vector kernels
// with warp-shuffle reductions, point-block Jacobi, single-reduction CG,
// restarted GMRES (CGS2) and an exact block-tridiagonal solver for 1D meshes.
//
// Replaces PETSc KSP/PC/Vec behind SNES (reference src/festim/problem.py:271-289:
// ksp preonly + pc lu by default).  Scalars of the Krylov recurrences stay on the
// device: every reduction finishes in the last block of the kernel that produced
// the partial sums (ticket counter), which also advances the recurrence, so an
// iteration is two launches for CG and the host only polls a "done" flag.
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "fb_internal.cuh"
#include "reduce.cuh"
#include "spmv_channels.cuh"
#include "xrank.cuh"

// ---------------------------------------------------------------------------
// SpMV: G lanes per scalar row (v, s); values of the row are contiguous
// ---------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ double spmv_row(const DevSpec& S, const double* __restrict__ J,
                                           const double* __restrict__ x, long long row, int lane) {
  const int ns = S.ns;
  const long long v = row / ns;
  const int s = (int)(row - v * ns);
  const int r0 = S.rowptr[v], deg = S.rowptr[v + 1] - r0;
  const int nc = S.pair_nc[s], pp = S.pair_pp[s];
  // block layout: entry (k, j) of the scalar row at (r0 + k) * npairs + pp + j
  const int np = S.npairs;
  const double* __restrict__ vals = J + (long long)r0 * np + pp;
  const int n = deg * nc;
  double sum = 0.0;
  if (nc == 1) {
    const int cs = S.pair_cs[pp];
    for (int e = lane; e < n; e += G) sum += vals[(long long)e * np] * x[(long long)S.colidx[r0 + e] * ns + cs];
  } else {
    for (int e = lane; e < n; e += G) {
      const int k = e / nc, j = e - k * nc;
      sum += vals[(long long)k * np + j] * x[(long long)S.colidx[r0 + k] * ns + S.pair_cs[pp + j]];
    }
  }
  // reduce over the G lanes of this row only: neighbouring groups of the warp may
  // have left already (ragged tail), so the shuffle mask names just this group
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) & ~(G - 1)));
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(gmask, sum, o);
  return sum;
}

template <int G>
__global__ void __launch_bounds__(FB_TPB)
k_spmv(DevSpec S, const double* __restrict__ J, const double* __restrict__ x, double* __restrict__ y) {
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / G;
  const int lane = threadIdx.x % G;
  if (row >= S.n_owned_dofs) return;
  const double sum = spmv_row<G>(S, J, x, row, lane);
  if (lane == 0) y[row] = sum;
}

// Multi-species SpMV on the block layout: one warp per vertex (= block row).  The ns values
// of every neighbour are staged once in shared memory (coalesced 8*ns-byte segments) and
// reused by all stored pairs; lane l owns the species pairs l, l+32, ... of every block, so
// the npairs values of a block are one contiguous coalesced load per pass, no index
// arithmetic or division sits in the loop over the neighbours, and 4 x PPL independent
// streaming loads (ld.global.cs) are in flight per lane.  PRECOND fuses the point-block
// Jacobi application  y = D_v^-1 (J x)_v  (left preconditioning of GMRES) into the epilogue.
__device__ __forceinline__ int fb_fastdiv_u(unsigned n, unsigned magic) {
  return magic ? (int)__umulhi(n, magic) : (int)n;
}

template <int PPL, bool PRECOND>
__global__ void __launch_bounds__(256)
k_spmv_blk(DevSpec S, const double* __restrict__ J, const double* __restrict__ x, double* __restrict__ y,
           const double* __restrict__ dinv, int xs_stride, const int* __restrict__ flags) {
  extern __shared__ double xs_all[];
  if (flags && flags[0]) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const long long v = (long long)blockIdx.x * wpb + warp;
  if (v >= S.n_owned) return;
  const int ns = S.ns, np = S.npairs;
  double* xs = xs_all + (size_t)warp * xs_stride;
  double* part = xs + S.maxdeg * ns;            // [np] partial sums, later [ns*ns] products
  const int r0 = S.rowptr[v], deg = S.rowptr[v + 1] - r0;
  for (int t = lane; t < deg * ns; t += 32) {
    const int k = fb_fastdiv_u((unsigned)t, S.vx_m_ns);
    xs[t] = x[(long long)S.colidx[r0 + k] * ns + (t - k * ns)];
  }
  int cs[PPL];
  double acc[PPL];
#pragma unroll
  for (int q = 0; q < PPL; ++q) {
    const int p = lane + 32 * q;
    cs[q] = p < np ? S.pair_cs[p] : -1;
    acc[q] = 0.0;
  }
  __syncwarp();
  const double* __restrict__ Jr = J + (long long)r0 * np + lane;
#pragma unroll 4
  for (int k = 0; k < deg; ++k) {
#pragma unroll
    for (int q = 0; q < PPL; ++q)
      if (cs[q] >= 0) acc[q] += __ldcs(Jr + (long long)k * np + 32 * q) * xs[k * ns + cs[q]];
  }
#pragma unroll
  for (int q = 0; q < PPL; ++q)
    if (cs[q] >= 0) part[lane + 32 * q] = acc[q];
  __syncwarp();
  double sum = 0.0;
  if (lane < ns) {
    const int p0 = S.pair_pp[lane], p1 = S.pair_pp[lane + 1];
    for (int p = p0; p < p1; ++p) sum += part[p];
  }
  if (!PRECOND) {
    if (lane < ns) y[v * ns + lane] = sum;
    return;
  }
  if (ns == 1) {
    if (lane == 0) y[v] = dinv[v] * sum;
    return;
  }
  __syncwarp();
  double* ys = xs;   // the staged iterate is dead: reuse
  if (lane < ns) ys[lane] = sum;
  __syncwarp();
  const double* __restrict__ dv = dinv + v * (long long)(ns * ns);
  for (int t = lane; t < ns * ns; t += 32) {
    const int r = fb_fastdiv_u((unsigned)t, S.vx_m_ns);
    part[t] = __ldcs(dv + t) * ys[t - r * ns];
  }
  __syncwarp();
  if (lane < ns) {
    double z = 0.0;
    for (int c = 0; c < ns; ++c) z += part[lane * ns + c];
    y[v * ns + lane] = z;
  }
}

// SpMV on the channel operator  J = sum_c A^c (x) B_c  (assemble_edge.cuh): one warp per vertex.
// Lane l owns the products (channel, column species) l, l+32, ...: t[(c, cs)] = sum_w A^c_vw
// x_(w,cs) streams 8 bytes per channel and edge - the values of an edge are one short contiguous
// segment shared by the lanes - and the constant tables B_c are applied once per row in the
// epilogue.  Strong Dirichlet conditions are applied on the fly (eliminated columns read as
// zero, identity rows copy x), so the operator equals the block-CSR matrix k_et_expand writes.
template <int PPL, bool PRECOND>
__global__ void __launch_bounds__(256)
k_spmv_ch(DevSpec S, const double* __restrict__ x, double* __restrict__ y, const double* __restrict__ dinv,
          int xs_stride, const int* __restrict__ flags) {
  extern __shared__ double xs_all[];
  if (flags && flags[0]) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const long long v = (long long)blockIdx.x * wpb + warp;
  if (v >= S.n_owned) return;
  const int ns = S.ns, nc = S.vx_ncombo;
  double* xs = xs_all + (size_t)warp * xs_stride;
  double* part = xs + S.maxdeg * ns;            // [ncombo] products, later [ns*ns]
  const int r0 = S.rowptr[v], deg = S.rowptr[v + 1] - r0;
  const unsigned vi = S.vinfo[v];
  const unsigned rowbc = vi & 0xffffu;
  const bool colbc = (vi >> 16) & 1u;
  for (int t = lane; t < deg * ns; t += 32) {
    const int k = fb_fastdiv_u((unsigned)t, S.vx_m_ns), s = t - k * ns;
    const int w = S.colidx[r0 + k];
    double val = x[(long long)w * ns + s];
    if (colbc && ((S.vinfo[w] >> s) & 1u)) val = 0.0;
    xs[t] = val;
  }
  const double* base[PPL];
  int stride[PPL], cs[PPL];
  double acc[PPL];
#pragma unroll
  for (int q = 0; q < PPL; ++q) {
    const int p = lane + 32 * q;
    cs[q] = -1; acc[q] = 0.0; base[q] = nullptr; stride[q] = 0;
    if (p < nc) {
      const int chan = S.vx_sp_combo[2 * p];
      cs[q] = S.vx_sp_combo[2 * p + 1];
      if (chan < S.vx_nstat) { base[q] = S.et_stat + (long long)r0 * S.vx_nsp + chan; stride[q] = S.vx_nsp; }
      else { base[q] = S.et_dyn + (long long)r0 * S.vx_ndp + (chan - S.vx_nstat); stride[q] = S.vx_ndp; }
    }
  }
  __syncwarp();
#pragma unroll 4
  for (int k = 0; k < deg; ++k) {
#pragma unroll
    for (int q = 0; q < PPL; ++q)
      if (cs[q] >= 0) acc[q] += __ldcs(base[q] + (long long)k * stride[q]) * xs[k * ns + cs[q]];
  }
#pragma unroll
  for (int q = 0; q < PPL; ++q)
    if (cs[q] >= 0) part[lane + 32 * q] = acc[q];
  __syncwarp();
  double sum = 0.0;
  if (lane < ns) {
    if ((rowbc >> lane) & 1u) sum = x[v * ns + lane];
    else
      for (int e = S.vx_sp_ptr[lane]; e < S.vx_sp_ptr[lane + 1]; ++e) {
        double coef = S.vx_sp_coef[e];
        if (S.vx_sp[2 * e + 1] & 1) coef *= S.inv_dt;
        sum += coef * part[S.vx_sp[2 * e]];
      }
  }
  if (!PRECOND) {
    if (lane < ns) y[v * ns + lane] = sum;
    return;
  }
  if (ns == 1) {
    if (lane == 0) y[v] = dinv[v] * sum;
    return;
  }
  __syncwarp();
  double* ys = xs;   // the staged iterate is dead: reuse
  if (lane < ns) ys[lane] = sum;
  __syncwarp();
  const double* __restrict__ dv = dinv + v * (long long)(ns * ns);
  for (int t = lane; t < ns * ns; t += 32) {
    const int r = fb_fastdiv_u((unsigned)t, S.vx_m_ns);
    part[t] = __ldcs(dv + t) * ys[t - r * ns];
  }
  __syncwarp();
  if (lane < ns) {
    double z = 0.0;
    for (int c = 0; c < ns; ++c) z += part[lane * ns + c];
    y[v * ns + lane] = z;
  }
}

template <bool PRECOND>
static int spmv_ch_launch(fb_ctx* ctx, const double* x, double* y, const int* flags) {
  const DevSpec& S = ctx->S;
  const int wpb = 8;
  const int tail = S.vx_ncombo > S.ns * S.ns ? S.vx_ncombo : S.ns * S.ns;
  const int stride = (S.maxdeg * S.ns + tail + 1) & ~1;
  const size_t smem = (size_t)wpb * stride * sizeof(double);
  const unsigned grid = (unsigned)((S.n_owned + wpb - 1) / wpb);
  const int ppl = (S.vx_ncombo + 31) / 32;
#define FB_SPMV_CASE(N)                                                                                   \
  {                                                                                                       \
    if (smem > 48 * 1024)                                                                                 \
      FB_CHECK(ctx, cudaFuncSetAttribute((const void*)k_spmv_ch<N, PRECOND>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         (int)smem));                                                     \
    k_spmv_ch<N, PRECOND><<<grid, wpb * 32, smem, ctx->stream>>>(S, x, y, ctx->d_dinv, stride, flags);      \
  }
  if (ppl <= 1) FB_SPMV_CASE(1)
  else if (ppl == 2) FB_SPMV_CASE(2)
  else if (ppl == 3) FB_SPMV_CASE(3)
  else if (ppl == 4) FB_SPMV_CASE(4)
  else if (ppl <= 8) FB_SPMV_CASE(8)
  else return fb_fail(ctx, "channel SpMV: more than 256 (channel, species) products");
#undef FB_SPMV_CASE
  ctx->launches++;
  FB_CHECK(ctx, cudaGetLastError());
  return 0;
}

// TMA-pipelined channel SpMV (spmv_channels.cuh): usable when every bulk copy is 16-byte aligned
static bool spmv_ch_tma_ok(fb_ctx* ctx, const double* x) {
  const DevSpec& S = ctx->S;
  return S.vx_on && (S.ns % 2 == 0) && S.ns <= 16 && (((uintptr_t)x & 15) == 0) && S.vx_nch <= 16 && S.ck_nx_max > 0 &&
         ctx->ep_epl > 0 && !getenv("FB_NO_TMA_SPMV");
}

template <int MODE>
static int spmv_ch_tma_launch(fb_ctx* ctx, SpmvChArgs a) {
  const DevSpec& S = ctx->S;
  a.ep_coef = ctx->d_ep_coef; a.ep_pos = ctx->d_ep_pos; a.ep_lane = ctx->d_ep_lane;
  int smem_max = 0;
  cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device);
  // the ring must hold the chunks the consumer groups work on plus at least one in flight per group
  // as deep a ring as shared memory holds (up to 6): the copies of a chunk are issued when the chunk
  // nst places ahead of it has been consumed, and must hide the HBM latency behind nst - 1 chunks of work
  int nst = 6;
  if (getenv("FB_SPMV_STAGES")) nst = atoi(getenv("FB_SPMV_STAGES"));
  if (nst > 8) nst = 8;
  size_t smem = fb_spmv_ch_layout(S, a, MODE, nst);
  while (nst > FB_CHUNK_GROUPS + 1 && smem > (size_t)smem_max) smem = fb_spmv_ch_layout(S, a, MODE, --nst);
  if (smem > (size_t)smem_max) return fb_fail(ctx, "channel SpMV: a chunk does not fit in shared memory");
  const long long nchunks = (S.n_owned + FB_CHUNK - 1) / FB_CHUNK;
  long long grid = ctx->sm_count;
  if (grid > nchunks) grid = nchunks;
  if (grid < 1) grid = 1;
  const int mt = (S.vx_nch + 7) / 8, nt = (S.ns + 7) / 8;
#define FB_SPMV_CASE_E(M, N, E)                                                                                 \
  {                                                                                                             \
    FB_CHECK(ctx, cudaFuncSetAttribute((const void*)k_spmv_ch_tma<M, N, MODE, E>,                               \
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                \
    k_spmv_ch_tma<M, N, MODE, E><<<(unsigned)grid, 32 * FB_CHUNK * FB_CHUNK_GROUPS + 64, smem, ctx->stream>>>(S, a); \
  }
  // the plan's entries per lane select the instantiation (registers)
#define FB_SPMV_CASE(M, N)                                                                                      \
  {                                                                                                             \
    if (ctx->ep_epl == 2) FB_SPMV_CASE_E(M, N, 2)                                                               \
    else if (ctx->ep_epl == 4) FB_SPMV_CASE_E(M, N, 4)                                                          \
    else if (ctx->ep_epl == 6) FB_SPMV_CASE_E(M, N, 6)                                                          \
    else FB_SPMV_CASE_E(M, N, 8)                                                                                \
  }
  if (mt == 1 && nt == 1) FB_SPMV_CASE(1, 1)
  else if (mt == 1) FB_SPMV_CASE(1, 2)
  else if (nt == 1) FB_SPMV_CASE(2, 1)
  else FB_SPMV_CASE(2, 2)
#undef FB_SPMV_CASE_E
#undef FB_SPMV_CASE
  ctx->launches++;
  FB_CHECK(ctx, cudaGetLastError());
  return 0;
}

// y = J x (precond false) or y = D^-1 J x on the channel operator
static int spmv_ch(fb_ctx* ctx, const double* x, double* y, bool precond, const int* flags) {
  if (int rc = fbi_halo(ctx, (double*)x)) return rc;
  if (spmv_ch_tma_ok(ctx, x) && (!precond || ctx->dinv32_valid)) {
    SpmvChArgs a{};
    a.x = x; a.y = y; a.dinv32 = ctx->d_dinv32; a.dscale = ctx->d_dscale; a.flags = flags;
    return precond ? spmv_ch_tma_launch<1>(ctx, a) : spmv_ch_tma_launch<0>(ctx, a);
  }
  return precond ? spmv_ch_launch<true>(ctx, x, y, flags) : spmv_ch_launch<false>(ctx, x, y, flags);
}

static int spmv_group(const DevSpec& S) {
  const double len = (S.tdim == 1 ? 3.0 : (S.tdim == 2 ? 7.0 : 15.0)) * S.npairs / S.ns;
  if (len <= 6) return 4;
  if (len <= 20) return 8;
  if (len <= 48) return 16;
  return 32;
}

template <bool PRECOND>
static int spmv_blk_launch(fb_ctx* ctx, const double* J, const double* x, double* y, const int* flags) {
  const DevSpec& S = ctx->S;
  const int wpb = 8;
  const int tail = S.npairs > S.ns * S.ns ? S.npairs : S.ns * S.ns;
  const int stride = (S.maxdeg * S.ns + tail + 1) & ~1;
  const size_t smem = (size_t)wpb * stride * sizeof(double);
  const unsigned grid = (unsigned)((S.n_owned + wpb - 1) / wpb);
  const int ppl = (S.npairs + 31) / 32;
#define FB_SPMV_CASE(N)                                                                                   \
  {                                                                                                       \
    if (smem > 48 * 1024)                                                                                 \
      FB_CHECK(ctx, cudaFuncSetAttribute((const void*)k_spmv_blk<N, PRECOND>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         (int)smem));                                                     \
    k_spmv_blk<N, PRECOND><<<grid, wpb * 32, smem, ctx->stream>>>(S, J, x, y, ctx->d_dinv, stride, flags);  \
  }
  if (ppl <= 1) FB_SPMV_CASE(1)
  else if (ppl == 2) FB_SPMV_CASE(2)
  else if (ppl == 3) FB_SPMV_CASE(3)
  else if (ppl == 4) FB_SPMV_CASE(4)
  else FB_SPMV_CASE(8)
#undef FB_SPMV_CASE
  ctx->launches++;
  FB_CHECK(ctx, cudaGetLastError());
  return 0;
}

static bool spmv_blk_ok(const DevSpec& S) {
  const int tail = S.npairs > S.ns * S.ns ? S.npairs : S.ns * S.ns;
  return S.ns > 1 && (size_t)8 * (S.maxdeg * S.ns + tail + 2) * 8 <= 200 * 1024 && !getenv("FB_SPMV_ROWWISE");
}

int fbi_spmv(fb_ctx* ctx, const double* J, const double* x, double* y) {
  const DevSpec& S = ctx->S;
  if (!J) {
    if (!S.vx_on) return fb_fail(ctx, "no channel operator: pass the block-CSR values");
    return spmv_ch(ctx, x, y, false, nullptr);
  }
  if (int rc = fbi_halo(ctx, (double*)x)) return rc;
  if (spmv_blk_ok(S)) return spmv_blk_launch<false>(ctx, J, x, y, nullptr);
  const int G = spmv_group(S);
  const long long threads = S.n_dofs * G;
  const unsigned grid = (unsigned)((threads + FB_TPB - 1) / FB_TPB);
  switch (G) {
    case 4: FB_LAUNCH(ctx, k_spmv<4>, grid, FB_TPB, 0, S, J, x, y); break;
    case 8: FB_LAUNCH(ctx, k_spmv<8>, grid, FB_TPB, 0, S, J, x, y); break;
    case 16: FB_LAUNCH(ctx, k_spmv<16>, grid, FB_TPB, 0, S, J, x, y); break;
    default: FB_LAUNCH(ctx, k_spmv<32>, grid, FB_TPB, 0, S, J, x, y); break;
  }
  FB_CHECK(ctx, cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// fused vector kernels
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(FB_TPB)
k_dot(long long n, const double* __restrict__ x, const double* __restrict__ y,
      double* partials, unsigned* ticket, double* out, XRank X) {
  double v[1] = {0.0}, tot[1];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    v[0] += x[i] * y[i];
  if (grid_reduce<1>(v, partials, ticket, tot)) {
    xrank_sum<1>(X, tot);
    if (threadIdx.x == 0) out[0] = tot[0];
  }
}

__global__ void k_axpy(long long n, double a, const double* __restrict__ x, double* __restrict__ y) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    y[i] += a * x[i];
}

// cross-rank context of the next reduction / halo exchange (fb_peer_*); single rank: inert
static XRank xrank_make(fb_ctx* ctx, unsigned seq) {
  XRank X{};
  X.n = ctx->peer_ready ? ctx->peer_n : 1;
  X.rank = ctx->peer_rank;
  X.seq = seq;
  for (int q = 0; q < 8; ++q) X.base[q] = (unsigned long long*)ctx->peer_base[q];
  X.bar = ctx->d_ticket ? ctx->d_ticket + 64 : nullptr;
  return X;
}
static XRank xr_next(fb_ctx* ctx) { return xrank_make(ctx, ctx->peer_ready ? ctx->xr_seq++ : 0u); }

// owner -> ghost exchange of a node-major vector over peer memory (no-op on a single rank)
int fbi_halo(fb_ctx* ctx, double* x) {
  if (!ctx->peer_ready) return 0;
  const DevSpec& S = ctx->S;
  int rc = fbi_ensure_workspace(ctx);
  if (rc) return rc;
  const XRank X = xrank_make(ctx, ctx->halo_seq++);
  const long long n_ghost_dofs = S.n_dofs - S.n_owned_dofs;
  const long long nsend = (long long)ctx->n_send * S.ns;
  const long long nthr = nsend > n_ghost_dofs ? nsend : n_ghost_dofs;
  if (nthr > 0)
    FB_LAUNCH(ctx, k_halo_exchange, (unsigned)((nthr + 255) / 256), 256, 0, X, S.ns, ctx->n_send, ctx->d_send_v,
              ctx->d_send_rank, ctx->d_send_dst, S.n_owned_dofs, n_ghost_dofs, x);
  FB_CHECK(ctx, cudaGetLastError());
  return 0;
}

static unsigned vec_grid(fb_ctx* ctx, long long n) {
  long long g = (n + FB_TPB - 1) / FB_TPB;
  const long long cap = (long long)ctx->sm_count * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned)g;
}

static int dot_to_host(fb_ctx* ctx, const double* x, const double* y, double* out) {
  const DevSpec& S = ctx->S;
  FB_LAUNCH(ctx, k_dot, vec_grid(ctx, S.n_owned_dofs), FB_TPB, 0, S.n_owned_dofs, x, y, ctx->d_red, ctx->d_ticket,
            ctx->d_sc + 60, xr_next(ctx));
  FB_CHECK(ctx, cudaMemcpyAsync(ctx->h_pinned, ctx->d_sc + 60, sizeof(double), cudaMemcpyDeviceToHost,
                                ctx->stream));
  FB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  *out = ctx->h_pinned[0];
  return 0;
}

int fbi_norm2(fb_ctx* ctx, const double* x, double* out) {
  double d;
  int rc = dot_to_host(ctx, x, x, &d);
  if (rc) return rc;
  *out = sqrt(d);
  return 0;
}

extern "C" int fb_dot(fb_ctx* ctx, const double* x, const double* y, double* out) {
  int rc = fbi_ensure_workspace(ctx);
  if (rc) return rc;
  return dot_to_host(ctx, x, y, out);
}

extern "C" int fb_axpy(fb_ctx* ctx, double a, const double* x, double* y) {
  FB_LAUNCH(ctx, k_axpy, vec_grid(ctx, ctx->S.n_owned_dofs), FB_TPB, 0, ctx->S.n_owned_dofs, a, x, y);
  FB_CHECK(ctx, cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// point-block Jacobi: invert the ns x ns diagonal block of every vertex
// ---------------------------------------------------------------------------
__global__ void k_bjacobi_setup(DevSpec S, const double* __restrict__ J, double* __restrict__ dinv,
                                int* __restrict__ flags) {
  const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (v >= S.n_owned) return;
  const int ns = S.ns;
  const int r0 = S.rowptr[v], deg = S.rowptr[v + 1] - r0;
  int kself = 0;
  for (int k = 0; k < deg; ++k)
    if (S.colidx[r0 + k] == v) kself = k;
  const long long base = (long long)r0 * S.npairs;
  double A[FB_MAX_NS][FB_MAX_NS], B[FB_MAX_NS][FB_MAX_NS];
  if (!J) {
    // channel operator: diagonal block from the channel values of the diagonal edge, strong
    // Dirichlet rows / columns applied as k_et_expand does
    const unsigned vi = S.vinfo[v] & 0xffffu;
    const int* __restrict__ jtp = S.vx_itab + S.vx_o_jtptr;
    const int* __restrict__ jt = S.vx_itab + S.vx_o_jt;
    const double* __restrict__ jtc = S.vx_dtab + S.vx_o_jtc;
    const double* __restrict__ st = S.et_stat + (long long)(r0 + kself) * S.vx_nsp;
    const double* __restrict__ dy = S.et_dyn + (long long)(r0 + kself) * S.vx_ndp;
    for (int s = 0; s < ns; ++s)
      for (int c = 0; c < ns; ++c) { A[s][c] = 0.0; B[s][c] = (s == c) ? 1.0 : 0.0; }
    for (int p = 0; p < S.npairs; ++p) {
      const int s = S.pair_row[p], c = S.pair_cs[p];
      double val = 0.0;
      for (int e = jtp[p]; e < jtp[p + 1]; ++e) {
        const int chan = jt[2 * e];
        double coef = jtc[e];
        if (jt[2 * e + 1] & 1) coef *= S.inv_dt;
        val += coef * (chan < S.vx_nstat ? st[chan] : dy[chan - S.vx_nstat]);
      }
      if ((vi >> c) & 1u) val = 0.0;
      if ((vi >> s) & 1u) val = (c == s) ? 1.0 : 0.0;
      A[s][c] = val;
    }
    if (ns == 1) {
      if (A[0][0] == 0.0) flags[2] = 1;
      dinv[v] = 1.0 / A[0][0];
      return;
    }
  } else {
    if (ns == 1) {
      const double d = J[base + kself];
      if (d == 0.0) flags[2] = 1;
      dinv[v] = 1.0 / d;
      return;
    }
    for (int s = 0; s < ns; ++s)
      for (int c = 0; c < ns; ++c) {
        const int jp = S.pair_idx[s * ns + c];
        A[s][c] = jp < 0 ? 0.0
                         : J[base + (long long)kself * S.npairs + S.pair_pp[s] + jp];
        B[s][c] = (s == c) ? 1.0 : 0.0;
      }
  }
  // Gauss-Jordan with partial pivoting
  for (int col = 0; col < ns; ++col) {
    int piv = col;
    double best = fabs(A[col][col]);
    for (int r = col + 1; r < ns; ++r)
      if (fabs(A[r][col]) > best) { best = fabs(A[r][col]); piv = r; }
    if (best == 0.0) { flags[2] = 1; best = 1.0; }
    if (piv != col)
      for (int c = 0; c < ns; ++c) {
        double t = A[col][c]; A[col][c] = A[piv][c]; A[piv][c] = t;
        t = B[col][c]; B[col][c] = B[piv][c]; B[piv][c] = t;
      }
    const double ip = 1.0 / A[col][col];
    for (int c = 0; c < ns; ++c) { A[col][c] *= ip; B[col][c] *= ip; }
    for (int r = 0; r < ns; ++r) {
      if (r == col) continue;
      const double f = A[r][col];
      if (f == 0.0) continue;
      for (int c = 0; c < ns; ++c) { A[r][c] -= f * A[col][c]; B[r][c] -= f * B[col][c]; }
    }
  }
  for (int s = 0; s < ns; ++s)
    for (int c = 0; c < ns; ++c) dinv[(v * ns + s) * ns + c] = B[s][c];
}

// The same inversion, G lanes per vertex with the block in shared memory: in-place Gauss-Jordan
// with row pivoting (row interchanges undone as column interchanges at the end).  The
// thread-per-vertex kernel above keeps two ns x ns arrays per thread in local memory, which
// costs 10 ms for 857 k blocks of 12 x 12; this one stays on chip.
template <int G>
__global__ void __launch_bounds__(256)
k_bjacobi_setup_coop(DevSpec S, const double* __restrict__ J, double* __restrict__ dinv, int* __restrict__ flags) {
  extern __shared__ double bj_smem[];
  const int g = threadIdx.x & (G - 1), grp = threadIdx.x / G, gpb = blockDim.x / G;
  const long long v = (long long)blockIdx.x * gpb + grp;
  if (v >= S.n_owned) return;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) & ~(G - 1)));
  const int ns = S.ns, nn = ns * ns;
  double* A = bj_smem + (size_t)grp * (nn + 2 * ns + 2);
  double* prow = A + nn;
  double* fcol = prow + ns;
  const int r0 = S.rowptr[v], deg = S.rowptr[v + 1] - r0;
  int kself = 0;
  for (int k = g; k < deg; k += G)
    if (S.colidx[r0 + k] == v) kself = k;
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) { const int t = __shfl_xor_sync(gmask, kself, o); kself = t > kself ? t : kself; }
  for (int t = g; t < nn; t += G) A[t] = 0.0;
  __syncwarp(gmask);
  if (!J) {
    const unsigned vi = S.vinfo[v] & 0xffffu;
    const int* __restrict__ jtp = S.vx_itab + S.vx_o_jtptr;
    const int* __restrict__ jt = S.vx_itab + S.vx_o_jt;
    const double* __restrict__ jtc = S.vx_dtab + S.vx_o_jtc;
    const double* __restrict__ st = S.et_stat + (long long)(r0 + kself) * S.vx_nsp;
    const double* __restrict__ dy = S.et_dyn + (long long)(r0 + kself) * S.vx_ndp;
    for (int p = g; p < S.npairs; p += G) {
      const int s = S.pair_row[p], c = S.pair_cs[p];
      double val = 0.0;
      for (int e = jtp[p]; e < jtp[p + 1]; ++e) {
        const int chan = jt[2 * e];
        double coef = jtc[e];
        if (jt[2 * e + 1] & 1) coef *= S.inv_dt;
        val += coef * (chan < S.vx_nstat ? st[chan] : dy[chan - S.vx_nstat]);
      }
      if ((vi >> c) & 1u) val = 0.0;
      if ((vi >> s) & 1u) val = (c == s) ? 1.0 : 0.0;
      A[s * ns + c] = val;
    }
  } else {
    const long long base = (long long)(r0 + kself) * S.npairs;
    for (int p = g; p < S.npairs; p += G) A[S.pair_row[p] * ns + S.pair_cs[p]] = J[base + p];
  }
  __syncwarp(gmask);
  unsigned long long perm = 0ull;   // 4 bits per column: the row it was swapped with
  for (int col = 0; col < ns; ++col) {
    // pivot: largest |A[r][col]|, r >= col (ties -> smallest r)
    double best = -1.0;
    int piv = col;
    for (int r = col + g; r < ns; r += G) {
      const double a = fabs(A[r * ns + col]);
      if (a > best) { best = a; piv = r; }
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(gmask, best, o);
      const int op = __shfl_xor_sync(gmask, piv, o);
      if (ob > best || (ob == best && op < piv)) { best = ob; piv = op; }
    }
    if (best == 0.0) { if (g == 0) flags[2] = 1; best = 1.0; }
    perm |= (unsigned long long)piv << (4 * col);
    if (piv != col)
      for (int c = g; c < ns; c += G) { const double t = A[col * ns + c]; A[col * ns + c] = A[piv * ns + c]; A[piv * ns + c] = t; }
    __syncwarp(gmask);
    const double ip = 1.0 / (A[col * ns + col] == 0.0 ? 1.0 : A[col * ns + col]);
    __syncwarp(gmask);
    for (int c = g; c < ns; c += G) prow[c] = (c == col ? 1.0 : A[col * ns + c]) * ip;
    for (int r = g; r < ns; r += G) fcol[r] = A[r * ns + col];
    __syncwarp(gmask);
    for (int t = g; t < nn; t += G) {
      const int r = t / ns, c = t - r * ns;
      if (r == col) A[t] = prow[c];
      else A[t] = (c == col ? 0.0 : A[t]) - fcol[r] * prow[c];
    }
    __syncwarp(gmask);
  }
  for (int col = ns - 1; col >= 0; --col) {   // undo the row interchanges on the columns
    const int piv = (int)((perm >> (4 * col)) & 0xfull);
    if (piv != col) {
      for (int r = g; r < ns; r += G) { const double t = A[r * ns + col]; A[r * ns + col] = A[r * ns + piv]; A[r * ns + piv] = t; }
      __syncwarp(gmask);
    }
  }
  double* __restrict__ out = dinv + v * (long long)nn;
  for (int t = g; t < nn; t += G) out[t] = A[t];
}

// 2^e as a double (|e| <= 1000)
__device__ __forceinline__ double fb_pow2(int e) {
#ifndef FB_EMU
  return __hiloint2double((e + 1023) << 20, 0);
#else
  return ldexp(1.0, e);
#endif
}

// Thread-per-vertex inversion with every matrix in its own shared-memory slab: element e of the
// matrix of thread t at bj_smem[e * (TPB + 1) + t], so the 32 threads of a warp always hit 32
// different banks whatever rows their pivots select, and the inverse blocks leave through a coalesced
// transposed copy.  Same arithmetic (Gauss-Jordan, partial pivoting, ties -> smallest row) as the
// cooperative kernel above, which spent ~4300 warp instructions per two vertices on shuffles and
// warp barriers (C2: 3.8 ms).  NS > 0: compile-time block size (pivot row in registers).
template <int TPB, int NS>
__global__ void __launch_bounds__(TPB) k_bjacobi_setup_slab(DevSpec S, const double* __restrict__ J, double* __restrict__ dinv,
                                                            int* __restrict__ flags, const int* __restrict__ bj_ptr,
                                                            const int* __restrict__ bj_pos, const double* __restrict__ bj_coef,
                                                            float* __restrict__ dinv32, double* __restrict__ dscale) {
  extern __shared__ double bj_smem[];
  constexpr int LD = TPB + 1;
  const int ns = NS > 0 ? NS : S.ns, nn = ns * ns;
  const long long v0 = (long long)blockIdx.x * TPB;
  const long long v = v0 + threadIdx.x;
  double* A = bj_smem + threadIdx.x;
  short* sexp = (short*)(bj_smem + (size_t)nn * LD);   // [ns][LD] exponents of the row scales
#define BJ(r, c) A[((r) * ns + (c)) * LD]
  if (v < S.n_owned) {
    const int r0 = S.rowptr[v], deg = S.rowptr[v + 1] - r0;
    int kself = 0;
    for (int k = 0; k < deg; ++k)
      if (S.colidx[r0 + k] == v) kself = k;
    for (int t = 0; t < nn; ++t) A[t * LD] = 0.0;
    if (!J) {
      // entries grouped by channel: one load of the channel value, uniform table reads
      const double* __restrict__ st = S.et_stat + (long long)(r0 + kself) * S.vx_nsp;
      const double* __restrict__ dy = S.et_dyn + (long long)(r0 + kself) * S.vx_ndp;
      const int nch = S.vx_nch, nstat = S.vx_nstat;
      double cv[16];   // all channel values of the diagonal edge in flight at once (nch <= 16)
#pragma unroll
      for (int ch = 0; ch < 16; ++ch) cv[ch] = ch < nch ? (ch < nstat ? st[ch] : dy[ch - nstat]) : 0.0;
#pragma unroll
      for (int ch = 0; ch < 16; ++ch) {
        if (ch >= nch) break;
        const double val = cv[ch];
        const int e1 = __ldg(bj_ptr + ch + 1);
#pragma unroll 4
        for (int e = __ldg(bj_ptr + ch); e < e1; ++e) {
          const int w = __ldg(bj_pos + e);
          const double coef = __ldg(bj_coef + e) * ((w >> 16) ? S.inv_dt : 1.0);
          A[(w & 0xffff) * LD] += coef * val;
        }
      }
      const unsigned vi = S.vinfo[v] & 0xffffu;
      if (vi) {   // Dirichlet dofs: eliminated columns, identity rows
        for (int s = 0; s < ns; ++s)
          for (int c = 0; c < ns; ++c) {
            if ((vi >> c) & 1u) BJ(s, c) = 0.0;
            if ((vi >> s) & 1u) BJ(s, c) = (c == s) ? 1.0 : 0.0;
          }
      }
    } else {
      const long long base = (long long)(r0 + kself) * S.npairs;
      for (int p = 0; p < S.npairs; ++p) BJ(S.pair_row[p], S.pair_cs[p]) = J[base + p];
    }
    unsigned long long perm = 0ull;   // 4 bits per column: the row it was swapped with
#pragma unroll 1
    for (int col = 0; col < ns; ++col) {
      double best = -1.0;
      int piv = col;
      for (int r = col; r < ns; ++r) {
        const double a = fabs(BJ(r, col));
        if (a > best) { best = a; piv = r; }
      }
      if (best == 0.0) flags[2] = 1;
      perm |= (unsigned long long)piv << (4 * col);
      if (NS > 0) {
        // pivot row (after the interchange) scaled into registers, the row it displaces moves to row piv
        double prow[NS > 0 ? NS : 1];
        const double pd = BJ(piv, col);
        const double ip = 1.0 / (pd == 0.0 ? 1.0 : pd);
#pragma unroll
        for (int c = 0; c < NS; ++c) {
          const double pv = BJ(piv, c), cv = BJ(col, c);
          if (piv != col) BJ(piv, c) = cv;
          prow[c] = (c == col ? 1.0 : pv) * ip;
        }
#pragma unroll
        for (int r = 0; r < NS; ++r) {
          if (r == col) {
#pragma unroll
            for (int c = 0; c < NS; ++c) BJ(r, c) = prow[c];
          } else {
            // new row = row - f * (scaled pivot row), except column col: (0 - f / pivot), fixed up afterwards
            const double f = BJ(r, col);
#pragma unroll
            for (int c = 0; c < NS; ++c) BJ(r, c) = BJ(r, c) - f * prow[c];
            BJ(r, col) = 0.0 - f * ip;
          }
        }
      } else {
        if (piv != col)
          for (int c = 0; c < ns; ++c) { const double t = BJ(col, c); BJ(col, c) = BJ(piv, c); BJ(piv, c) = t; }
        const double pd = BJ(col, col);
        const double ip = 1.0 / (pd == 0.0 ? 1.0 : pd);
        for (int c = 0; c < ns; ++c) BJ(col, c) = (c == col ? 1.0 : BJ(col, c)) * ip;
        for (int r = 0; r < ns; ++r) {
          if (r == col) continue;
          const double f = BJ(r, col);
          for (int c = 0; c < ns; ++c) BJ(r, c) = (c == col ? 0.0 : BJ(r, c)) - f * BJ(col, c);
        }
      }
    }
    for (int col = ns - 1; col >= 0; --col) {   // undo the row interchanges on the columns
      const int piv = (int)((perm >> (4 * col)) & 0xfull);
      if (piv != col)
        for (int r = 0; r < ns; ++r) { const double t = BJ(r, col); BJ(r, col) = BJ(r, piv); BJ(r, piv) = t; }
    }
    // every row rounded to single-precision mantissas times a power of two (the scale of the row):
    // the channel SpMV streams the 4-byte copy, every other kernel reads the same values in double
    for (int s = 0; s < ns; ++s) {
      double mx = 0.0;
      for (int c = 0; c < ns; ++c) mx = fmax(mx, fabs(BJ(s, c)));
      int ex = 0;
      if (mx > 0.0) (void)frexp(mx, &ex);
      ex = ex < -1000 ? -1000 : (ex > 1000 ? 1000 : ex);
      const double up = fb_pow2(ex), dn = fb_pow2(-ex);
      for (int c = 0; c < ns; ++c) BJ(s, c) = (double)(float)(BJ(s, c) * dn) * up;
      sexp[s * LD + threadIdx.x] = (short)ex;
    }
  }
#undef BJ
  __syncthreads();
  const long long left = S.n_owned - v0;
  const int nvalid = (int)(left < TPB ? left : TPB);
  double* __restrict__ out = dinv + v0 * nn;
  float* __restrict__ out32 = dinv32 + v0 * nn;
  for (int i = threadIdx.x; i < nvalid * nn; i += TPB) {
    const int m = i / nn, e = i - m * nn;
    const double val = bj_smem[e * LD + m];
    const int ex = sexp[(e / ns) * LD + m];
    out[i] = val;
    out32[i] = (float)(val * fb_pow2(-ex));   // exact
  }
  double* __restrict__ outs = dscale + v0 * ns;
  for (int i = threadIdx.x; i < nvalid * ns; i += TPB) {
    const int m = i / ns, s = i - m * ns;
    outs[i] = fb_pow2((int)sexp[s * LD + m]);
  }
}

static int bjacobi_setup(fb_ctx* ctx, const double* J) {
  const DevSpec& S = ctx->S;
  ctx->dinv32_valid = false;
  if (S.ns == 1 || S.ns > 16 || getenv("FB_BJ_THREAD")) {
    FB_LAUNCH(ctx, k_bjacobi_setup, (unsigned)((S.n_owned + FB_TPB - 1) / FB_TPB), FB_TPB, 0, S, J, ctx->d_dinv, ctx->d_flags);
    return 0;
  }
  if (!getenv("FB_BJ_COOP") && (J || ctx->d_bj_ptr)) {
    constexpr int TPB = 64;
    const size_t smem = (size_t)S.ns * S.ns * (TPB + 1) * sizeof(double) + (size_t)S.ns * (TPB + 1) * sizeof(short);
    const unsigned grid = (unsigned)((S.n_owned + TPB - 1) / TPB);
    ctx->dinv32_valid = true;
#define FB_BJ_CASE(N)                                                                                                   \
    {                                                                                                                   \
      FB_CHECK(ctx, cudaFuncSetAttribute((const void*)k_bjacobi_setup_slab<TPB, N>,                                     \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                      \
      FB_LAUNCH(ctx, (k_bjacobi_setup_slab<TPB, N>), grid, TPB, smem, S, J, ctx->d_dinv, ctx->d_flags, ctx->d_bj_ptr,    \
                ctx->d_bj_pos, ctx->d_bj_coef, ctx->d_dinv32, ctx->d_dscale);                 \
    }
    if (S.ns == 12) FB_BJ_CASE(12)
    else if (S.ns == 4) FB_BJ_CASE(4)
    else if (S.ns == 6) FB_BJ_CASE(6)
    else FB_BJ_CASE(0)
#undef FB_BJ_CASE
    return 0;
  }
  const int G = 16, gpb = 256 / G;
  const size_t smem = (size_t)gpb * (S.ns * S.ns + 2 * S.ns + 2) * sizeof(double);
  FB_LAUNCH(ctx, k_bjacobi_setup_coop<16>, (unsigned)((S.n_owned + gpb - 1) / gpb), 256, smem, S, J, ctx->d_dinv, ctx->d_flags);
  return 0;
}

__device__ __forceinline__ void bjacobi_apply(int ns, const double* __restrict__ dinv, long long v,
                                              const double* rin, double* zout) {
  if (ns == 1) { zout[0] = dinv[v] * rin[0]; return; }
  for (int s = 0; s < ns; ++s) {
    double acc = 0.0;
    for (int c = 0; c < ns; ++c) acc += dinv[(v * ns + s) * ns + c] * rin[c];
    zout[s] = acc;
  }
}

__global__ void k_precond(DevSpec S, const double* __restrict__ dinv, const double* __restrict__ r,
                          double* __restrict__ z, const int* __restrict__ flags) {
  if (flags && flags[0]) return;
  const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (v >= S.n_owned) return;
  double rin[FB_MAX_NS], zo[FB_MAX_NS];
  for (int s = 0; s < S.ns; ++s) rin[s] = r[v * S.ns + s];
  bjacobi_apply(S.ns, dinv, v, rin, zo);
  for (int s = 0; s < S.ns; ++s) z[v * S.ns + s] = zo[s];
}

// z = D^-1 r from the single-precision mantissas + row scales of k_bjacobi_setup_slab (same values as
// the double blocks, 5/8 of their bytes less): one thread per dof, the ns threads of a vertex read
// consecutive rows of the block (coalesced) and share the vertex' entries of r through L1.
__global__ void k_precond32(long long n_owned, int ns, const float* __restrict__ dinv32, const double* __restrict__ dscale,
                            const double* __restrict__ r, double* __restrict__ z, const int* __restrict__ flags) {
  if (flags && flags[0]) return;
  const int vpb = blockDim.x / ns;
  const int vl = threadIdx.x / ns, s = threadIdx.x - vl * ns;
  const long long v = (long long)blockIdx.x * vpb + vl;
  if (vl >= vpb || v >= n_owned) return;
  const long long i = v * ns + s;
  const float* __restrict__ row = dinv32 + i * ns;
  const double* __restrict__ rv = r + v * ns;
  double a0 = 0.0, a1 = 0.0;
  int c = 0;
  if ((ns & 3) == 0) {
    for (; c < ns; c += 4) {
      const float4 d4 = *(const float4*)(row + c);
      a0 += (double)d4.x * rv[c] + (double)d4.z * rv[c + 2];
      a1 += (double)d4.y * rv[c + 1] + (double)d4.w * rv[c + 3];
    }
  } else {
    for (; c < ns; ++c) a0 += (double)row[c] * rv[c];
  }
  z[i] = dscale[i] * (a0 + a1);
}

static int precond_launch(fb_ctx* ctx, const double* r, double* z, const int* flags) {
  const DevSpec& S = ctx->S;
  if (ctx->dinv32_valid && S.ns <= 32) {
    const int vpb = 256 / S.ns;
    FB_LAUNCH(ctx, k_precond32, (unsigned)((S.n_owned + vpb - 1) / vpb), 256, 0, S.n_owned, S.ns, ctx->d_dinv32, ctx->d_dscale, r, z, flags);
  } else {
    FB_LAUNCH(ctx, k_precond, (unsigned)((S.n_owned + FB_TPB - 1) / FB_TPB), FB_TPB, 0, S, ctx->d_dinv, r, z, flags);
  }
  return 0;
}

// ---------------------------------------------------------------------------
// CG, Chronopoulos-Gear single-reduction form, block-Jacobi preconditioned.
//   sc[0]=alpha sc[1]=beta sc[2]=gamma sc[3]=rr sc[4]=threshold^2 sc[5]=rr0
//   flags[0]=done flags[1]=iterations flags[2]=breakdown
// ---------------------------------------------------------------------------
__global__ void k_cg_init(DevSpec S, const double* __restrict__ dinv, const double* __restrict__ b,
                          double* x, double* r, double* u, double* p, double* s_) {
  const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (v >= S.n_owned) return;
  double rin[FB_MAX_NS], zo[FB_MAX_NS];
  for (int s = 0; s < S.ns; ++s) {
    const long long i = v * S.ns + s;
    rin[s] = b[i];
    r[i] = rin[s]; x[i] = 0.0; p[i] = 0.0; s_[i] = 0.0;
  }
  bjacobi_apply(S.ns, dinv, v, rin, zo);
  for (int s = 0; s < S.ns; ++s) u[v * S.ns + s] = zo[s];
}

__global__ void __launch_bounds__(FB_TPB)
k_cg_update(DevSpec S, const double* __restrict__ dinv, const double* __restrict__ sc,
            const int* __restrict__ flags, double* __restrict__ x, double* __restrict__ r,
            double* __restrict__ u, const double* __restrict__ w, double* __restrict__ p,
            double* __restrict__ s_) {
  if (flags[0]) return;
  const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (v >= S.n_owned) return;
  const double alpha = sc[0], beta = sc[1];
  double rin[FB_MAX_NS], zo[FB_MAX_NS];
  for (int s = 0; s < S.ns; ++s) {
    const long long i = v * S.ns + s;
    const double pi = u[i] + beta * p[i];
    const double si = w[i] + beta * s_[i];
    p[i] = pi; s_[i] = si;
    x[i] += alpha * pi;
    rin[s] = r[i] - alpha * si;
    r[i] = rin[s];
  }
  bjacobi_apply(S.ns, dinv, v, rin, zo);
  for (int s = 0; s < S.ns; ++s) u[v * S.ns + s] = zo[s];
}

template <int G>
__global__ void __launch_bounds__(FB_TPB)
k_cg_spmv_dots(DevSpec S, const double* __restrict__ J, const double* __restrict__ u,
               const double* __restrict__ r, double* __restrict__ w, double* partials,
               unsigned* ticket, double* sc, int* flags) {
  if (flags[0]) return;
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / G;
  const int lane = threadIdx.x % G;
  double v[3] = {0.0, 0.0, 0.0}, tot[3];
  if (row < S.n_owned_dofs) {  // uniform within each G-lane group
    const double sum = spmv_row<G>(S, J, u, row, lane);
    if (lane == 0) {
      w[row] = sum;
      const double ri = r[row], ui = u[row];
      v[0] = ri * ui; v[1] = sum * ui; v[2] = ui * ui;  // (r,u), (w,u), ||M^-1 r||^2
    }
  }
  if (grid_reduce<3>(v, partials, ticket, tot) && threadIdx.x == 0) {
    const double gamma_new = tot[0], delta = tot[1], rr = tot[2];
    const int it = flags[1];
    sc[3] = rr;
    if (it == 0) sc[5] = rr;
    if (rr <= sc[4] || !(rr == rr)) {
      flags[0] = 1;
      if (!(rr == rr)) flags[2] = 1;
    } else {
      double alpha, beta;
      if (it == 0) { beta = 0.0; alpha = gamma_new / delta; }
      else {
        beta = gamma_new / sc[2];
        alpha = gamma_new / (delta - beta * gamma_new / sc[0]);
      }
      if (!(alpha == alpha) || isinf(alpha)) { flags[0] = 1; flags[2] = 1; }
      sc[0] = alpha; sc[1] = beta; sc[2] = gamma_new;
      flags[1] = it + 1;
    }
  }
}

// ---------------------------------------------------------------------------
// Persistent CG: the whole Chronopoulos-Gear iteration in ONE cooperative kernel.
// Each block owns a contiguous range of vertices (and their dof rows) for the
// life of the solve; two grid barriers per iteration (after the vector update,
// after SpMV + partial dots); every block reduces the per-block partial sums in
// the same order, so alpha/beta are bitwise identical everywhere and no scalar
// is broadcast.  When the operator fits, the rows a block owns are staged ONCE
// into shared memory (values + column indices, up to 227 KB per SM = 33 MB per
// B200) and every iteration after that reads only the gathered vector entries
// from L2: for the L2-resident configs the SpMV leaves the L2/HBM roofline
// behind and the iteration cost is the two grid barriers.
// ---------------------------------------------------------------------------
struct CgArgs {
  const double* J; const double* dinv; const double* b;
  double* x; double* r; double* u; double* w; double* p; double* s;
  double* partials; unsigned* bar; double* sc; int* flags;
  double thr2; int max_it; int rows_in_smem;  // rows_in_smem: 1 = stage owned rows in smem
  unsigned bar_base;                            // value of the arrival counter at launch
  // compact per-SM partition (nullptr: contiguous equal ranges, gathers from L2)
  const int* blk_rng; const int* halo_ptr; const int* halo_idx; const unsigned short* lcol;
  long long su_doubles, bentries, rows_cap;     // smem capacities: iterate cache, block entries, rows
  // two-level preconditioner: coarse space = one constant per block (nullptr: Jacobi only)
  const double* Einv; double* rc; const unsigned short* halo_blk;
  const int* cn_ptr; const int* cn_blk; double rel2;  // rel2: threshold relative to ||M^-1 b||^2
  long long smem_entries;                       // capacity (values) of the staging area
};

// Grid barrier for a cooperative (co-resident) launch.  bar[0] is a monotonically
// increasing arrival counter (release-add), bar[32] the generation the last arriver
// publishes (release store) and everybody else polls (acquire load); the caller keeps
// its generation number in a register, so a barrier costs one atomic round trip plus
// one poll round trip.  A bounded spin turns a scheduling problem into an error code
// instead of a hang (bar[64] = abort flag).  Returns false on abort.
__device__ __forceinline__ bool grid_barrier(unsigned* bar, unsigned nblocks, unsigned& gen,
                                             unsigned base_count) {
  __shared__ int ok_s;
  __syncthreads();
  ++gen;
  if (threadIdx.x == 0) {
    int ok = 1;
    const unsigned old = atom_add_release_u32(bar, 1u);
    if (old - base_count == gen * nblocks - 1u) {
      st_release_u32(bar + 32, gen);
    } else {
      unsigned spins = 0;
      while (ld_acquire_u32(bar + 32) != gen) {
        __nanosleep(64);  // keep the pollers off the L2 slice the workers still need
        if ((++spins & 0xfffu) == 0u) {
          if (spins > (1u << 24)) { atomicExch(bar + 64, 1u); ok = 0; break; }
          if (ld_acquire_u32(bar + 64)) { ok = 0; break; }
        }
      }
    }
    ok_s = ok;
  }
  __syncthreads();
  return ok_s != 0;
}

// ---------------------------------------------------------------------------
// Two-level preconditioner of the persistent CG: M^-1 = D^-1 + Z E^-1 Z^T with Z the
// piecewise-constant prolongation from the per-SM blocks (one coarse unknown per SM) and
// E = Z^T A Z (nb x nb, SPD).  The restriction Z^T r is a block sum each SM already owns,
// the coarse solve is one row of E^-1 per block, the prolongation adds one number to the
// block's rows: the coarse correction rides on the two grid barriers the iteration has
// anyway and removes the ~nb smoothest modes that make Jacobi-CG slow.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
k_coarse_build(DevSpec S, const double* __restrict__ J, const int* __restrict__ blk_rng,
               const unsigned short* __restrict__ lcol, const int* __restrict__ cn_ptr,
               const int* __restrict__ cn_blk, const int* __restrict__ cn_eptr,
               const int* __restrict__ cn_ent, double* __restrict__ E) {
  // E[b][b]: entries of block b's rows whose column is owned (local column < rows of the
  // block); E[b][c]: the precomputed entry list of the neighbour pair, one warp per pair.
  // Fixed summation orders: the preconditioner is bit-reproducible.
  __shared__ double part[32];
  const int nb = gridDim.x, b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  const int v0 = blk_rng[b], nvloc = blk_rng[b + 1] - v0;
  const int e0 = S.rowptr[v0], e1 = S.rowptr[v0 + nvloc];
  (void)e0; (void)e1;
  double acc = 0.0;
  for (int rr = tid; rr < nvloc; rr += blockDim.x) {
    if (S.bc_map[v0 + rr] >= 0) continue;  // Dirichlet rows are not in the coarse space
    for (int e = S.rowptr[v0 + rr]; e < S.rowptr[v0 + rr + 1]; ++e)
      if ((int)lcol[e] < nvloc) acc += J[e];
  }
  acc = warp_sum(acc);
  if (lane == 0) part[wid] = acc;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int q = 0; q < nw; ++q) s += part[q];
    E[(size_t)b * nb + b] = s > 0.0 ? s : 1.0;  // block of Dirichlet rows only: decoupled dummy
  }
  for (int k = cn_ptr[b] + wid; k < cn_ptr[b + 1]; k += nw) {
    double a2 = 0.0;
    for (int q = cn_eptr[k] + lane; q < cn_eptr[k + 1]; q += 32) a2 += J[cn_ent[q]];
    a2 = warp_sum(a2);
    if (lane == 0) E[(size_t)b * nb + cn_blk[k]] = a2;
  }
}

// block id of every halo vertex, FB_COARSE_NONE for Dirichlet vertices (their y slot is 0)
#define FB_COARSE_NONE 159
__global__ void k_coarse_halo_blocks(int n, const int* __restrict__ halo_idx, const unsigned short* __restrict__ blk,
                                     const int* __restrict__ bc_map, unsigned short* __restrict__ out) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h < n) out[h] = bc_map[halo_idx[h]] >= 0 ? (unsigned short)FB_COARSE_NONE : blk[h];
}

// In-place Gauss-Jordan inversion of the coarse operator (SPD: no pivoting) by ONE thread
// block.  A single SM moves 128 B/clk through shared memory, so a matrix kept there costs
// 3 accesses x nb^2 x nb steps = 0.4 ms; here every thread keeps a GJ_T x GJ_T tile in
// registers and only the pivot row and column travel through (double-buffered) shared
// memory: one __syncthreads and GJ_T^2 DFMA per thread and step.  The pivot loop is
// unrolled GJ_T times so that the position of the pivot inside a tile is a compile-time
// constant (register tiles cannot be indexed dynamically).
#define GJ_T 6
#define GJ_THREADS 640   // 25 x 25 tiles of 6 x 6 cover 150 x 150
#define GJ_MAX 150
__device__ __forceinline__ double gj_rcp(double d) {
  double x = (double)(1.0f / (float)d);  // 24 bits, then two Newton steps in fp64
  if (!(fabs(x) > 0.0) || isinf(x)) return 1.0 / d;
  x = fma(x, fma(-d, x, 1.0), x);
  x = fma(x, fma(-d, x, 1.0), x);
  return x;
}

__global__ void __launch_bounds__(GJ_THREADS, 1)
k_coarse_invert(int nb, const double* __restrict__ E, double* __restrict__ Einv) {
  __shared__ double prow[2][GJ_MAX + GJ_T], fcol[2][GJ_MAX + GJ_T];
  const int nt = (nb + GJ_T - 1) / GJ_T;  // tiles per dimension
  const int ty = threadIdx.x / nt, tx = threadIdx.x - ty * nt;
  const int r0 = ty * GJ_T, c0 = tx * GJ_T;
  const bool active = ty < nt;
  double a[GJ_T][GJ_T];
#pragma unroll
  for (int i = 0; i < GJ_T; ++i)
#pragma unroll
    for (int j = 0; j < GJ_T; ++j) {
      const int r = r0 + i, c = c0 + j;
      a[i][j] = (active && r < nb && c < nb) ? E[(size_t)r * nb + c] : (r == c ? 1.0 : 0.0);  // identity padding
    }
  int buf = 0;
  for (int tb = 0; tb < nt; ++tb) {  // tile holding the pivot
    const bool has_row = active && ty == tb, has_col = active && tx == tb;
#pragma unroll
    for (int k = 0; k < GJ_T; ++k) {  // pivot (tb * GJ_T + k): local index k in its tile
      const int col = tb * GJ_T + k;
      if (col >= nb) break;
      if (has_row) {
#pragma unroll
        for (int j = 0; j < GJ_T; ++j) prow[buf][c0 + j] = a[k][j];
      }
      if (has_col) {
#pragma unroll
        for (int i = 0; i < GJ_T; ++i) {
          fcol[buf][r0 + i] = a[i][k];
          a[i][k] = 0.0;  // the pivot column restarts from zero: a = -f * ip below
        }
      }
      __syncthreads();  // one barrier per step: the other buffer is not rewritten before the next one
      if (active) {
        const double ip = gj_rcp(prow[buf][col]);
        double f[GJ_T];
#pragma unroll
        for (int i = 0; i < GJ_T; ++i) f[i] = fcol[buf][r0 + i];
#pragma unroll
        for (int j = 0; j < GJ_T; ++j) {
          double pr = prow[buf][c0 + j] * ip;
          if (j == k && has_col) pr = ip;
#pragma unroll
          for (int i = 0; i < GJ_T; ++i) a[i][j] = fma(-f[i], pr, a[i][j]);
          if (has_row) a[k][j] = pr;  // the pivot row itself
        }
      }
      buf ^= 1;
    }
  }
  if (active) {
#pragma unroll
    for (int i = 0; i < GJ_T; ++i)
#pragma unroll
      for (int j = 0; j < GJ_T; ++j) {
        const int r = r0 + i, c = c0 + j;
        if (r < nb && c < nb) Einv[(size_t)r * nb + c] = a[i][j];
      }
  }
}

template <int G>
__global__ void __launch_bounds__(1024, 1)
k_cg_persistent(DevSpec S, CgArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double red[3][32];
  __shared__ double tot[3];
  const int ns = S.ns;
  const unsigned nb = gridDim.x;
  const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = T >> 5;
  // vertex / row ownership: a contiguous range per block
  const bool part = A.blk_rng != nullptr;
  const long long vper = (S.n_vertices + nb - 1) / nb;
  const long long v0 = part ? A.blk_rng[blockIdx.x] : (long long)blockIdx.x * vper;
  const long long v1 = part ? A.blk_rng[blockIdx.x + 1]
                            : ((v0 + vper < S.n_vertices) ? v0 + vper : S.n_vertices);
  const int nvloc = v1 > v0 ? (int)(v1 - v0) : 0;
  const int nrows = nvloc * ns;
  const long long row0 = v0 * ns;
  // ---- shared-memory staging ------------------------------------------------------
  // vals[cap] f64 | su[su_doubles] f64 | cols: int[cap] (range mode) or u16[bentries]
  // (partition mode) | rstart[rows] | rlen[rows] | rbst[rows]
  const long long rows_cap = part ? A.rows_cap : vper * ns;
  double* svals = (double*)smem_raw;
  double* su = svals + A.smem_entries;
  int* scols = (int*)(su + A.su_doubles);
  unsigned short* slcol = (unsigned short*)scols;
  int* rstart = part ? (int*)(slcol + ((A.bentries + 1) & ~1LL)) : scols + A.smem_entries;
  int* rlen = rstart + rows_cap;
  int* rbst = rlen + rows_cap;
  unsigned char* rfree = (unsigned char*)(rbst + rows_cap);  // 1: row is in the coarse space
  bool staged = false;
  int nhalo = 0;
  const int* hidx = nullptr;
  if (A.rows_in_smem && nrows > 0) {
    const long long e0 = (long long)S.rowptr[v0] * S.npairs;
    const long long e1 = (long long)S.rowptr[v1] * S.npairs;
    staged = (e1 - e0) <= A.smem_entries;
    if (part) {
      nhalo = A.halo_ptr[blockIdx.x + 1] - A.halo_ptr[blockIdx.x];
      hidx = A.halo_idx + A.halo_ptr[blockIdx.x];
      staged = staged && (long long)(nvloc + nhalo) * ns <= A.su_doubles &&
               (long long)(S.rowptr[v1] - S.rowptr[v0]) <= A.bentries;
    }
    if (staged) {
      for (long long e = tid; e < e1 - e0; e += T) svals[e] = A.J[e0 + e];
      if (part) {
        const long long b0 = S.rowptr[v0], nbe = S.rowptr[v1] - b0;
        for (long long e = tid; e < nbe; e += T) slcol[e] = A.lcol[b0 + e];
      }
      for (int rr = tid; rr < nrows; rr += T) {
        const int vl = rr / ns, sp = rr - vl * ns;
        const long long v = v0 + vl;
        const int r0 = S.rowptr[v], deg = S.rowptr[v + 1] - r0;
        const int nc = S.pair_nc[sp], pp = S.pair_pp[sp];
        const int start = (int)((long long)r0 * S.npairs + pp - e0);   // block layout: entry (k, j) at start + k * npairs + j
        rstart[rr] = start;
        rlen[rr] = deg * nc;
        rbst[rr] = r0 - S.rowptr[v0];
        if (!part)
          for (int e = 0; e < deg * nc; ++e) {
            const int k = e / nc, j = e - k * nc;
            scols[start + k * S.npairs + j] = S.colidx[r0 + k] * ns + S.pair_cs[pp + j];
          }
      }
    }
  }
  const bool cached = staged && part;  // iterate gathered from shared memory
  // ---- coarse space (one constant per block): this block needs y = E^-1 rc at itself and
  // at its neighbour blocks; warp k keeps row k of that set of E^-1 in registers ----------
  __shared__ double ysm[160], rcs[160];
  const bool coarse = A.Einv != nullptr && cached && ns == 1;
  if (coarse) {
    for (int rr = tid; rr < nrows; rr += T) rfree[rr] = S.bc_map[row0 + rr] < 0;
    if (tid == 0) ysm[FB_COARSE_NONE] = 0.0;
  }
  const int cn0 = coarse ? A.cn_ptr[blockIdx.x] : 0;
  const int nneed = coarse ? A.cn_ptr[blockIdx.x + 1] - cn0 + 1 : 0;  // self + neighbours
  const unsigned short* hblk = coarse ? A.halo_blk + A.halo_ptr[blockIdx.x] : nullptr;
  int ctarget = -1;
  double erow[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  if (coarse && wid < nneed) {
    ctarget = wid == 0 ? (int)blockIdx.x : A.cn_blk[cn0 + wid - 1];
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      const unsigned c = lane + 32 * q;
      erow[q] = c < nb ? A.Einv[(size_t)ctarget * nb + c] : 0.0;
    }
  }
  double rsum = 0.0;  // this thread's share of the block sum of r (restriction)
  // ---- init ------------------------------------------------------------------------
  for (long long v = v0 + tid; v < v1; v += T) {
    double rin[FB_MAX_NS], zo[FB_MAX_NS];
    for (int sp = 0; sp < ns; ++sp) {
      const long long i = v * ns + sp;
      rin[sp] = A.b[i];
      A.r[i] = rin[sp]; A.x[i] = 0.0; A.p[i] = 0.0; A.s[i] = 0.0;
    }
    if (coarse && rfree[v - v0]) rsum += rin[0];
    bjacobi_apply(ns, A.dinv, v, rin, zo);
    for (int sp = 0; sp < ns; ++sp) {
      A.u[v * ns + sp] = zo[sp];
      if (cached) su[(v - v0) * ns + sp] = zo[sp];
    }
  }
  if (coarse) {  // (rfree[rr] is always read by the thread that staged it)
    rsum = warp_sum(rsum);
    if (lane == 0) red[0][wid] = rsum;
    __syncthreads();
    if (wid == 0) {
      double sacc = lane < nw ? red[0][lane] : 0.0;
      sacc = warp_sum(sacc);
      if (lane == 0) A.rc[blockIdx.x] = sacc;
    }
  }
  unsigned gen = 0;
  const unsigned base_count = A.bar_base;
  if (!grid_barrier(A.bar, nb, gen, base_count)) return;
  double alpha = 0.0, beta = 0.0, gamma_old = 0.0, thr2 = A.thr2;
  int it = 0, done = 0, breakdown = 0;
  const int ngroups = T / G, grp = tid / G, gl = tid % G;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
  const double* __restrict__ uvec = A.u;
  long long tph[5] = {0, 0, 0, 0, 0};  // cycles per phase (block 0), reported in sc[8..12]
  while (true) {
    long long c0 = clock64();
    // ---- coarse correction: u = D^-1 r + Z E^-1 Z^T r -------------------------------------
    double hval0 = 0.0;
    int hb0 = FB_COARSE_NONE;
    if (coarse && tid < nhalo) {  // first halo value per thread in flight during the coarse solve
      hval0 = __ldcg(uvec + hidx[tid]);
      hb0 = hblk[tid];
    }
    if (coarse) {
      // one warp fetches the restricted residual (148 blocks polling the same 10 sectors
      // from every warp would serialise in one L2 slice)
      if (wid == 0) {
        double rv[5];  // loads first, stores after (in-order warp: one round trip, not five)
#pragma unroll
        for (int q = 0; q < 5; ++q) {
          const unsigned c = lane + 32 * q;
          rv[q] = c < nb ? __ldcg(A.rc + c) : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 5; ++q) {
          const unsigned c = lane + 32 * q;
          if (c < nb) rcs[c] = rv[q];
        }
      }
      __syncthreads();
      for (int k = wid; k < nneed; k += nw) {
        double d = 0.0;
        int tb = ctarget;
        if (k == wid) {
#pragma unroll
          for (int q = 0; q < 5; ++q) {
            const unsigned c = lane + 32 * q;
            if (c < nb) d += erow[q] * rcs[c];
          }
        } else {  // more neighbours than warps: rows from L2
          tb = A.cn_blk[cn0 + k - 1];
#pragma unroll
          for (int q = 0; q < 5; ++q) {
            const unsigned c = lane + 32 * q;
            if (c < nb) d += A.Einv[(size_t)tb * nb + c] * rcs[c];
          }
        }
        d = warp_sum(d);
        if (lane == 0) ysm[tb] = d;
      }
      __syncthreads();
      const double yself = ysm[blockIdx.x];
      for (int rr = tid; rr < nrows; rr += T)
        if (rfree[rr]) su[rr] += yself;
    }
    // ---- halo of the iterate into shared memory ---------------------------------------
    if (cached) {
      if (coarse) {
        if (tid < nhalo) su[nrows + tid] = hval0 + ysm[hb0];
        for (int h = tid + T; h < nhalo; h += T) su[nrows + h] = __ldcg(uvec + hidx[h]) + ysm[hblk[h]];
      } else {
        for (int h = tid; h < nhalo * ns; h += T) {
          const int hv = h / ns, sp = h - hv * ns;
          su[nrows + h] = __ldcg(uvec + (long long)hidx[hv] * ns + sp);
        }
      }
      __syncthreads();
    }
    // ---- phase B: w = A u on owned rows + partial dots -----------------------------------
    double v3[3] = {0.0, 0.0, 0.0};
    for (int rr = grp; rr < nrows; rr += ngroups) {
      double sum = 0.0;
      const long long row = row0 + rr;
      double ri, ui;
      if (cached) {
        ri = __ldcg(A.r + row);
        ui = su[rr];
        const int a0 = rstart[rr], n = rlen[rr], bs = rbst[rr];
        if (ns == 1) {
          for (int e = gl; e < n; e += G) sum += svals[a0 + e] * su[slcol[bs + e]];
        } else {
          const int vl = rr / ns, sp = rr - vl * ns;
          const int nc = S.pair_nc[sp], pp = S.pair_pp[sp];
          for (int e = gl; e < n; e += G) {
            const int k = e / nc, j = e - k * nc;
            sum += svals[a0 + k * S.npairs + j] * su[(int)slcol[bs + k] * ns + S.pair_cs[pp + j]];
          }
        }
      } else if (staged) {
        ri = __ldcg(A.r + row); ui = __ldcg(uvec + row);  // issued before the gathers
        const int a0 = rstart[rr], n = rlen[rr];
        const int ncr = ns == 1 ? 1 : S.pair_nc[rr - (rr / ns) * ns];
        auto eofs = [&](int e) { return ns == 1 ? e : (e / ncr) * S.npairs + (e - (e / ncr) * ncr); };
        // all gathers of a batch are in flight before the first use (L2 latency once per batch)
        for (int eb = gl; eb < n; eb += 8 * G) {
          double uv[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int e = eb + q * G;
            uv[q] = e < n ? __ldcg(uvec + scols[a0 + eofs(e)]) : 0.0;
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int e = eb + q * G;
            if (e < n) sum += svals[a0 + eofs(e)] * uv[q];
          }
        }
      } else {
        ri = __ldcg(A.r + row); ui = __ldcg(uvec + row);
        const int vl = rr / ns, sp = rr - vl * ns;
        const long long v = v0 + vl;
        const int r0 = S.rowptr[v], deg = S.rowptr[v + 1] - r0;
        const int nc = S.pair_nc[sp], pp = S.pair_pp[sp];
        const double* __restrict__ vals = A.J + (long long)r0 * S.npairs + pp;   // entry (k, j) at k * npairs + j
        const int n = deg * nc;
        if (nc == 1) {
          const int cs = S.pair_cs[pp];
          for (int e = gl; e < n; e += G)
            sum += vals[(long long)e * S.npairs] * __ldcg(uvec + (long long)S.colidx[r0 + e] * ns + cs);
        } else {
          for (int e = gl; e < n; e += G) {
            const int k = e / nc, j = e - k * nc;
            sum += vals[(long long)k * S.npairs + j] * __ldcg(uvec + (long long)S.colidx[r0 + k] * ns + S.pair_cs[pp + j]);
          }
        }
      }
#pragma unroll
      for (int o = G / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(gmask, sum, o);
      if (gl == 0) {
        A.w[row] = sum;
        v3[0] += ri * ui; v3[1] += sum * ui; v3[2] += ui * ui;
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double sred = warp_sum(v3[k]);
      if (lane == 0) red[k][wid] = sred;
    }
    __syncthreads();
    if (wid < 3) {
      double sacc = lane < nw ? red[wid][lane] : 0.0;
      sacc = warp_sum(sacc);
      if (lane == 0) A.partials[(size_t)wid * nb + blockIdx.x] = sacc;
    }
    long long c1 = clock64();
    if (!grid_barrier(A.bar, nb, gen, base_count)) return;
    long long c2 = clock64();
    // ---- every block: identical reduction of the partials --------------------------------
    if (wid < 3) {
      double pv[5];  // loads first, adds after: one L2 round trip instead of five
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const unsigned bq = lane + 32 * j;
        pv[j] = bq < nb ? __ldcg(A.partials + (size_t)wid * nb + bq) : 0.0;
      }
      double sacc = (((pv[0] + pv[1]) + pv[2]) + pv[3]) + pv[4];
      for (unsigned bq = lane + 160; bq < nb; bq += 32) sacc += __ldcg(A.partials + (size_t)wid * nb + bq);
      sacc = warp_sum(sacc);
      if (lane == 0) tot[wid] = sacc;
    }
    __syncthreads();
    const double gamma_new = tot[0], delta = tot[1], uu = tot[2];
    if (it == 0 && A.Einv != nullptr) thr2 = A.rel2 * uu;  // relative to ||M^-1 b|| of this M
    if (uu <= thr2 || !(uu == uu)) { done = 1; breakdown = !(uu == uu); }
    else if (it >= A.max_it) { done = 2; }
    else {
      if (it == 0) { beta = 0.0; alpha = gamma_new / delta; }
      else { beta = gamma_new / gamma_old; alpha = gamma_new / (delta - beta * gamma_new / alpha); }
      if (!(alpha == alpha) || isinf(alpha)) { done = 1; breakdown = 1; }
      gamma_old = gamma_new;
    }
    if (done) {
      if (blockIdx.x == 0 && tid == 0) {
        A.flags[0] = (done == 1); A.flags[1] = it; A.flags[2] = breakdown; A.sc[3] = uu;
        for (int k = 0; k < 5; ++k) A.sc[8 + k] = (double)tph[k];
      }
      break;
    }
    ++it;
    long long c3 = clock64();
    // ---- phase A: vector update on owned vertices ------------------------------------------
    rsum = 0.0;
    for (long long v = v0 + tid; v < v1; v += T) {
      double rin[FB_MAX_NS], zo[FB_MAX_NS];
      for (int sp = 0; sp < ns; ++sp) {
        const long long i = v * ns + sp;
        const double ui = cached ? su[(v - v0) * ns + sp] : __ldcg(A.u + i);
        const double wi = __ldcg(A.w + i), ri = __ldcg(A.r + i);
        const double p_old = __ldcg(A.p + i), s_old = __ldcg(A.s + i), x_old = __ldcg(A.x + i);
        const double pi = ui + beta * p_old;
        const double si = wi + beta * s_old;
        A.p[i] = pi; A.s[i] = si;
        A.x[i] = x_old + alpha * pi;
        rin[sp] = ri - alpha * si;
        A.r[i] = rin[sp];
      }
      if (coarse && rfree[v - v0]) rsum += rin[0];
      bjacobi_apply(ns, A.dinv, v, rin, zo);
      for (int sp = 0; sp < ns; ++sp) {
        A.u[v * ns + sp] = zo[sp];
        if (cached) su[(v - v0) * ns + sp] = zo[sp];
      }
    }
    if (coarse) {
      rsum = warp_sum(rsum);
      if (lane == 0) red[0][wid] = rsum;
      __syncthreads();
      if (wid == 0) {
        double sacc = lane < nw ? red[0][lane] : 0.0;
        sacc = warp_sum(sacc);
        if (lane == 0) A.rc[blockIdx.x] = sacc;
      }
    }
    long long c4 = clock64();
    if (!grid_barrier(A.bar, nb, gen, base_count)) return;  // also protects tot[]
    long long c5 = clock64();
    tph[0] += c1 - c0; tph[1] += c2 - c1; tph[2] += c3 - c2; tph[3] += c4 - c3; tph[4] += c5 - c4;
  }
}

// ---------------------------------------------------------------------------
// GMRES(m), right preconditioned, CGS2 orthogonalisation.
//   gm layout in sc: H[(m+1)*m] at 128, cs[m], sn[m], g[m+1], hd1[m+1], hd2[m+1], misc
// ---------------------------------------------------------------------------
#define GM_M 30
#define GM_H 128
#define GM_CS (GM_H + (GM_M + 1) * GM_M)
#define GM_SN (GM_CS + GM_M)
#define GM_G (GM_SN + GM_M)
#define GM_HD1 (GM_G + GM_M + 1)
#define GM_HD2 (GM_HD1 + GM_M + 1)
#define GM_MISC (GM_HD2 + GM_M + 1)  // [0]=norm2 of w, [1]=scale, [2]=residual, [3]=threshold, [4]=beta0, [5]=refined
#define GM_Y (GM_MISC + 8)
#define GM_END (GM_Y + GM_M)
#define GM_HRAW (GM_END + 8)                 // unrotated Hessenberg columns (Ritz values for the Chebyshev bounds)
#define GM_SC_DOUBLES (GM_HRAW + (GM_M + 1) * GM_M + 8)
#define GM_NVEC (GM_M + 8)                   // V[GM_M+1], w, z, r, and d0, d1, cr, cz of the Chebyshev recurrence

// Classical Gram-Schmidt with one refinement pass when needed (PETSc's default for GMRES:
// KSPGMRESClassicalGramSchmidtOrthogonalization, refine "if needed"), as two or three sweeps:
//   k_gs_dots    h1[i] = <V_i, w> (i < nv),  n0 = <w, w>
//   k_gs_update  w -= sum_i h1[i] V_i  and, in the same sweep,  h2[i] = <V_i, w>,  n1 = <w, w>;
//                the last block decides: refine iff n1 < n0 / 2
//   k_gs_refine  (only then) w -= sum_i h2[i] V_i,  n2 = <w, w>
// Every thread keeps its NVB accumulators in registers and reads w once per sweep; the NVB
// basis loads of an element are independent (one long burst of loads in flight per thread).
template <int NVB>
__global__ void __launch_bounds__(FB_TPB)
k_gs_dots(long long n, long long ld, int nv, const double* __restrict__ V, const double* __restrict__ w,
          double* partials, unsigned* ticket, double* hd, double* n0, const int* __restrict__ flags, XRank X) {
  if (flags[0]) return;
  double acc[NVB + 1], tot[NVB + 1];
#pragma unroll
  for (int i = 0; i <= NVB; ++i) acc[i] = 0.0;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n;
       k += (long long)gridDim.x * blockDim.x) {
    const double wk = w[k];
    double vv[NVB];
#pragma unroll
    for (int i = 0; i < NVB; ++i) vv[i] = i < nv ? V[(size_t)i * ld + k] : 0.0;
#pragma unroll
    for (int i = 0; i < NVB; ++i) acc[i] += vv[i] * wk;
    acc[NVB] += wk * wk;
  }
  if (grid_reduce<NVB + 1>(acc, partials, ticket, tot)) {
    xrank_sum<NVB + 1>(X, tot);
    if (threadIdx.x == 0) {
      for (int i = 0; i < nv; ++i) hd[i] = tot[i];
      n0[0] = tot[NVB];
    }
  }
}

template <int NVB>
__global__ void __launch_bounds__(FB_TPB)
k_gs_update(long long n, long long ld, int nv, const double* __restrict__ V, const double* __restrict__ h1,
            double* __restrict__ w, double* partials, unsigned* ticket, double* h2, double* misc,
            const int* __restrict__ flags, XRank X) {
  if (flags[0]) return;
  double acc[NVB + 1], tot[NVB + 1], h[NVB];
#pragma unroll
  for (int i = 0; i < NVB; ++i) { acc[i] = 0.0; h[i] = i < nv ? h1[i] : 0.0; }
  acc[NVB] = 0.0;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n;
       k += (long long)gridDim.x * blockDim.x) {
    double wk = w[k];
    double vv[NVB];
#pragma unroll
    for (int i = 0; i < NVB; ++i) vv[i] = i < nv ? V[(size_t)i * ld + k] : 0.0;
#pragma unroll
    for (int i = 0; i < NVB; ++i) wk -= h[i] * vv[i];
    w[k] = wk;
#pragma unroll
    for (int i = 0; i < NVB; ++i) acc[i] += vv[i] * wk;
    acc[NVB] += wk * wk;
  }
  if (grid_reduce<NVB + 1>(acc, partials, ticket, tot)) {
    xrank_sum<NVB + 1>(X, tot);
    if (threadIdx.x == 0) {
      for (int i = 0; i < nv; ++i) h2[i] = tot[i];
      // misc[0] = <w, w> before (k_gs_dots) -> norm^2 of the orthogonalised vector; misc[5] = refine?
      const bool refine = tot[NVB] < 0.5 * misc[0];
      misc[0] = tot[NVB];
      misc[5] = refine ? 1.0 : 0.0;
    }
  }
}

template <int NVB>
__global__ void __launch_bounds__(FB_TPB)
k_gs_refine(long long n, long long ld, int nv, const double* __restrict__ V, const double* __restrict__ h2,
            double* __restrict__ w, double* partials, unsigned* ticket, double* misc,
            const int* __restrict__ flags, XRank X) {
  if (flags[0] || misc[5] == 0.0) return;
  double acc[1] = {0.0}, tot[1], h[NVB];
#pragma unroll
  for (int i = 0; i < NVB; ++i) h[i] = i < nv ? h2[i] : 0.0;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n;
       k += (long long)gridDim.x * blockDim.x) {
    double wk = w[k];
    double vv[NVB];
#pragma unroll
    for (int i = 0; i < NVB; ++i) vv[i] = i < nv ? V[(size_t)i * ld + k] : 0.0;
#pragma unroll
    for (int i = 0; i < NVB; ++i) wk -= h[i] * vv[i];
    w[k] = wk;
    acc[0] += wk * wk;
  }
  if (grid_reduce<1>(acc, partials, ticket, tot)) {
    xrank_sum<1>(X, tot);
    if (threadIdx.x == 0) misc[0] = tot[0];
  }
}

__global__ void k_gmres_givens(int j, double* sc, int* flags) {
  if (flags[0]) return;
  double* H = sc + GM_H;
  double* cs = sc + GM_CS; double* sn = sc + GM_SN; double* g = sc + GM_G;
  double* col = H + (size_t)j * (GM_M + 1);
  const bool refined = sc[GM_MISC + 5] != 0.0;   // second Gram-Schmidt pass applied (k_gs_refine)
  for (int i = 0; i <= j; ++i) col[i] = sc[GM_HD1 + i] + (refined ? sc[GM_HD2 + i] : 0.0);
  const double hn = sqrt(sc[GM_MISC + 0]);
  col[j + 1] = hn;
  for (int i = 0; i <= j + 1; ++i) sc[GM_HRAW + (size_t)j * (GM_M + 1) + i] = col[i];
  sc[GM_MISC + 1] = hn > 0.0 ? 1.0 / hn : 0.0;
  for (int i = 0; i < j; ++i) {
    const double t = cs[i] * col[i] + sn[i] * col[i + 1];
    col[i + 1] = -sn[i] * col[i] + cs[i] * col[i + 1];
    col[i] = t;
  }
  const double a = col[j], b = col[j + 1];
  const double rho = sqrt(a * a + b * b);
  const double c = rho > 0.0 ? a / rho : 1.0, s = rho > 0.0 ? b / rho : 0.0;
  cs[j] = c; sn[j] = s;
  col[j] = rho; col[j + 1] = 0.0;
  g[j + 1] = -s * g[j];
  g[j] = c * g[j];
  const double res = fabs(g[j + 1]);
  sc[GM_MISC + 2] = res;
  flags[1] += 1;   // total iterations
  flags[3] = j + 1;  // columns in this cycle
  if (res <= sc[GM_MISC + 3] || !(res == res) || hn == 0.0) {
    flags[0] = 1;
    if (!(res == res)) flags[2] = 1;
  }
}

__global__ void k_scale_to(long long n, const double* __restrict__ w, const double* __restrict__ scale,
                           double* __restrict__ out, const int* __restrict__ flags) {
  if (flags && flags[0]) return;
  const double s = scale[0];
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n;
       k += (long long)gridDim.x * blockDim.x)
    out[k] = w[k] * s;
}

__global__ void k_gmres_solve_y(double* sc, const int* flags) {
  const int nc = flags[3];
  const double* H = sc + GM_H; const double* g = sc + GM_G; double* y = sc + GM_Y;
  for (int i = nc - 1; i >= 0; --i) {
    double acc = g[i];
    for (int k = i + 1; k < nc; ++k) acc -= H[(size_t)k * (GM_M + 1) + i] * y[k];
    y[i] = acc / H[(size_t)i * (GM_M + 1) + i];
  }
  for (int i = nc; i < GM_M; ++i) y[i] = 0.0;
}

// t = sum_i y[i] V_i (i < ncols read from flags[3])
__global__ void k_combine(long long n, long long ld, const double* __restrict__ V, const double* __restrict__ y,
                          const int* __restrict__ flags, double* __restrict__ t) {
  const int nc = flags[3];
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n;
       k += (long long)gridDim.x * blockDim.x) {
    double acc = 0.0;
    for (int i = 0; i < nc; ++i) acc += y[i] * V[(size_t)i * ld + k];
    t[k] = acc;
  }
}

__global__ void k_residual(long long n, const double* __restrict__ b, const double* __restrict__ Ax,
                           double* __restrict__ r) {
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n;
       k += (long long)gridDim.x * blockDim.x)
    r[k] = b[k] - Ax[k];
}

__global__ void k_gmres_begin(double* sc, int* flags, double thr, double guess_guard) {
  // sc[GM_MISC+0] holds ||r||^2 from the preceding reduction
  const double beta = sqrt(sc[GM_MISC + 0]);
  // an initial guess that leaves a larger residual than the zero vector would is dropped (flags[4])
  flags[4] = (guess_guard > 0.0 && beta > guess_guard) ? 1 : 0;
  sc[GM_MISC + 4] = beta;
  sc[GM_MISC + 1] = beta > 0.0 ? 1.0 / beta : 0.0;
  sc[GM_MISC + 3] = thr;
  for (int i = 0; i <= GM_M; ++i) sc[GM_G + i] = 0.0;
  sc[GM_G] = beta;
  flags[3] = 0;
  flags[0] = (beta <= thr) ? 1 : 0;
}

// ---------------------------------------------------------------------------
// exact block-tridiagonal LU for interval meshes (the reference's default
// "preonly + lu" is an exact solve; in 1D the node graph is tridiagonal)
// ---------------------------------------------------------------------------
__global__ void k_tri_extract(DevSpec S, const double* __restrict__ J, double* __restrict__ T) {
  // T[v][3][ns][ns]: sub, diag, super dense blocks
  const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (v >= S.n_vertices) return;
  const int ns = S.ns, r0 = S.rowptr[v], deg = S.rowptr[v + 1] - r0;
  double* out = T + v * 3 * ns * ns;
  for (int i = 0; i < 3 * ns * ns; ++i) out[i] = 0.0;
  const long long base = (long long)r0 * S.npairs;
  for (int k = 0; k < deg; ++k) {
    const int w = S.colidx[r0 + k];
    const int pos = w - (int)v + 1;
    if (pos < 0 || pos > 2) continue;
    for (int s = 0; s < ns; ++s)
      for (int j = 0; j < S.pair_nc[s]; ++j)
        out[(pos * ns + s) * ns + S.pair_cs[S.pair_pp[s] + j]] =
            J[base + (long long)k * S.npairs + S.pair_pp[s] + j];
  }
}

// dense ns x ns solve helpers (single thread)
__device__ void dense_lu_solve(int ns, double* A, double* B, int nrhs) {
  // solves A X = B in place (B: ns x nrhs row-major), partial pivoting
  for (int col = 0; col < ns; ++col) {
    int piv = col;
    double best = fabs(A[col * ns + col]);
    for (int r = col + 1; r < ns; ++r)
      if (fabs(A[r * ns + col]) > best) { best = fabs(A[r * ns + col]); piv = r; }
    if (piv != col) {
      for (int c = 0; c < ns; ++c) { double t = A[col * ns + c]; A[col * ns + c] = A[piv * ns + c]; A[piv * ns + c] = t; }
      for (int c = 0; c < nrhs; ++c) { double t = B[col * nrhs + c]; B[col * nrhs + c] = B[piv * nrhs + c]; B[piv * nrhs + c] = t; }
    }
    const double ip = 1.0 / A[col * ns + col];
    for (int r = col + 1; r < ns; ++r) {
      const double f = A[r * ns + col] * ip;
      if (f == 0.0) continue;
      for (int c = col; c < ns; ++c) A[r * ns + c] -= f * A[col * ns + c];
      for (int c = 0; c < nrhs; ++c) B[r * nrhs + c] -= f * B[col * nrhs + c];
    }
  }
  for (int r = ns - 1; r >= 0; --r) {
    for (int c = 0; c < nrhs; ++c) {
      double acc = B[r * nrhs + c];
      for (int k = r + 1; k < ns; ++k) acc -= A[r * ns + k] * B[k * nrhs + c];
      B[r * nrhs + c] = acc / A[r * ns + r];
    }
  }
}

__global__ void k_tri_solve(DevSpec S, double* __restrict__ T, const double* __restrict__ b,
                            double* __restrict__ y, double* __restrict__ Cp /*[nv][ns][ns]*/) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int ns = S.ns;
  const long long nv = S.n_vertices;
  double A[FB_MAX_NS * FB_MAX_NS], R[FB_MAX_NS * (FB_MAX_NS + 1)];
  // forward: C'_v = (B_v - A_v C'_{v-1})^-1 C_v ; d'_v = (..)^-1 (d_v - A_v d'_{v-1})
  for (long long v = 0; v < nv; ++v) {
    const double* sub = T + (v * 3 + 0) * ns * ns;
    const double* dia = T + (v * 3 + 1) * ns * ns;
    const double* sup = T + (v * 3 + 2) * ns * ns;
    for (int i = 0; i < ns * ns; ++i) A[i] = dia[i];
    for (int r = 0; r < ns; ++r) {
      for (int c = 0; c < ns; ++c) R[r * (ns + 1) + c] = sup[r * ns + c];
      R[r * (ns + 1) + ns] = b[v * ns + r];
    }
    if (v > 0) {
      const double* Cprev = Cp + (v - 1) * ns * ns;
      for (int r = 0; r < ns; ++r) {
        double accd = 0.0;
        for (int k = 0; k < ns; ++k) {
          const double a = sub[r * ns + k];
          if (a == 0.0) continue;
          for (int c = 0; c < ns; ++c) A[r * ns + c] -= a * Cprev[k * ns + c];
          accd += a * y[(v - 1) * ns + k];
        }
        R[r * (ns + 1) + ns] -= accd;
      }
    }
    dense_lu_solve(ns, A, R, ns + 1);
    for (int r = 0; r < ns; ++r) {
      for (int c = 0; c < ns; ++c) Cp[v * ns * ns + r * ns + c] = R[r * (ns + 1) + c];
      y[v * ns + r] = R[r * (ns + 1) + ns];
    }
  }
  for (long long v = nv - 2; v >= 0; --v) {
    const double* C = Cp + v * ns * ns;
    for (int r = 0; r < ns; ++r) {
      double acc = y[v * ns + r];
      for (int c = 0; c < ns; ++c) acc -= C[r * ns + c] * y[(v + 1) * ns + c];
      y[v * ns + r] = acc;
    }
  }
}

// ---------------------------------------------------------------------------
// host drivers
// ---------------------------------------------------------------------------
int fbi_ensure_workspace(fb_ctx* ctx) {
  const DevSpec& S = ctx->S;
  if (ctx->d_sc) return 0;
  int rc;
  if ((rc = fb_alloc(ctx, &ctx->d_sc, GM_SC_DOUBLES + 64))) return rc;
  // partial sums: (GM_M + 2) vectors x capped vector grid, or 3 x the SpMV grid
  size_t red = (size_t)FB_RED_SLOTS * ctx->sm_count * 8;
  const size_t spmv_blocks = (size_t)((S.n_dofs * 32 + FB_TPB - 1) / FB_TPB);
  if (red < 3 * spmv_blocks + 16) red = 3 * spmv_blocks + 16;
  ctx->red_doubles = (long long)red;
  if ((rc = fb_alloc(ctx, &ctx->d_red, red))) return rc;
  if ((rc = fb_alloc(ctx, &ctx->d_ticket, 256))) return rc;
  if ((rc = fb_alloc(ctx, &ctx->d_flags, 8))) return rc;
  if ((rc = fb_alloc(ctx, &ctx->d_dinv, (size_t)S.n_vertices * S.ns * S.ns))) return rc;
  if ((rc = fb_alloc(ctx, &ctx->d_dinv32, (size_t)S.n_vertices * S.ns * S.ns + 8))) return rc;
  if ((rc = fb_alloc(ctx, &ctx->d_dscale, (size_t)S.n_vertices * S.ns + 8))) return rc;
  FB_CHECK(ctx, cudaMemsetAsync(ctx->d_ticket, 0, 256 * sizeof(unsigned), ctx->stream));
  FB_CHECK(ctx, cudaMemsetAsync(ctx->d_flags, 0, 32, ctx->stream));
  FB_CHECK(ctx, cudaMallocHost((void**)&ctx->h_pinned, 64 * sizeof(double)));
  return 0;
}

static int ensure_vectors(fb_ctx* ctx, int nvec) {
  const long long need = (long long)nvec * ctx->S.n_dofs;
  if (ctx->work_doubles >= need) return 0;
  int rc = fb_alloc(ctx, &ctx->d_work, (size_t)need);  // old block stays owned until destroy
  if (rc) return rc;
  ctx->work_doubles = need;
  return 0;
}

static int read_flags(fb_ctx* ctx, int* out4) {
  FB_CHECK(ctx, cudaMemcpyAsync(ctx->h_pinned + 8, ctx->d_flags, 4 * sizeof(int), cudaMemcpyDeviceToHost,
                                ctx->stream));
  FB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  const int* f = (const int*)(ctx->h_pinned + 8);
  for (int i = 0; i < 4; ++i) out4[i] = f[i];
  return 0;
}

// E = Z^T J Z and its inverse for the two-level preconditioner (once per matrix)
static int coarse_setup(fb_ctx* ctx, const double* J) {
  ctx->coarse_ready = false;
  if (!ctx->d_coarse || ctx->part_min_rows <= 0 || ctx->part_blocks > GJ_MAX || getenv("FB_NO_COARSE") || getenv("FB_NO_SM_PARTITION") ||
      getenv("FB_NO_SMEM_OPERATOR") || getenv("FB_CG_MULTIKERNEL"))
    return 0;
  const int nb = ctx->part_blocks;
  if (ctx->coarse_masks_dirty) {
    const int nh = ctx->part_nhalo;
    if (nh > 0)
      FB_LAUNCH(ctx, k_coarse_halo_blocks, (unsigned)((nh + 255) / 256), 256, 0, nh, ctx->d_halo_idx, ctx->d_halo_blk,
                (const int*)ctx->d_bc_map, ctx->d_halo_blk_eff);
    ctx->coarse_masks_dirty = false;
  }
  double* E = ctx->d_coarse;
  double* Einv = E + (size_t)nb * nb;
  FB_LAUNCH(ctx, k_coarse_build, nb, 1024, 0, ctx->S, J, ctx->d_blk_rng, ctx->d_lcol, ctx->d_cn_ptr, ctx->d_cn_blk,
            ctx->d_cn_eptr, ctx->d_cn_ent, E);
  FB_LAUNCH(ctx, k_coarse_invert, 1, GJ_THREADS, 0, nb, (const double*)E, Einv);
  ctx->coarse_ready = true;
  return 0;
}

static int solve_cg_persistent(fb_ctx* ctx, const double* J, const double* b, double* y, double thr2, double rel2,
                               int max_it, int* iters) {
  const DevSpec& S = ctx->S;
  int rc = ensure_vectors(ctx, 5);
  if (rc) return rc;
  const long long N = S.n_dofs;
  CgArgs A;
  A.J = J; A.dinv = ctx->d_dinv; A.b = b; A.x = y;
  A.r = ctx->d_work; A.u = A.r + N; A.w = A.u + N; A.p = A.w + N; A.s = A.p + N;
  A.partials = ctx->d_red; A.bar = ctx->d_ticket + 64; A.sc = ctx->d_sc; A.flags = ctx->d_flags;
  A.thr2 = thr2; A.max_it = max_it;
  // lanes per row: ~8 entries per lane keeps many independent gathers in flight
  const double rowlen = (double)ctx->nnzb * S.npairs / (double)N;
  int G = rowlen <= 20 ? 2 : rowlen <= 40 ? 4 : rowlen <= 80 ? 8 : rowlen <= 160 ? 16 : 32;
  if (getenv("FB_CG_G")) G = atoi(getenv("FB_CG_G"));
  const int T = 1024;
  const unsigned nb = (unsigned)ctx->sm_count;
  size_t smem_max = 0;
  {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device);
    smem_max = (size_t)v - 3584;  // static shared (reduction scratch, coarse solution)
  }
  A.blk_rng = nullptr; A.halo_ptr = nullptr; A.halo_idx = nullptr; A.lcol = nullptr;
  A.su_doubles = 0; A.bentries = 0; A.rows_cap = 0;
  A.Einv = nullptr; A.rc = nullptr; A.halo_blk = nullptr; A.cn_ptr = nullptr; A.cn_blk = nullptr; A.rel2 = rel2;
  size_t smem = 0;
  const bool no_smem = getenv("FB_NO_SMEM_OPERATOR") != nullptr;
  if (ctx->part_blocks == (int)nb && !ctx->part_owned_only && !getenv("FB_NO_SM_PARTITION")) {
    // compact per-SM partition: operator rows + iterate cache (owned + halo) on chip
    A.blk_rng = ctx->d_blk_rng; A.halo_ptr = ctx->d_halo_ptr; A.halo_idx = ctx->d_halo_idx;
    A.lcol = ctx->d_lcol;
    const long long cap = (ctx->part_max_entries + 1) & ~1LL;
    const long long su = (ctx->part_max_su + 1) & ~1LL;
    const long long be = (ctx->part_max_bentries + 1) & ~1LL;
    const long long rows = ctx->part_max_rows;
    const size_t need = (size_t)(8 * cap + 8 * su + 2 * be + 13 * rows + 32);
    if (need <= smem_max && !no_smem) {
      A.rows_in_smem = 1; A.smem_entries = cap; A.su_doubles = su; A.bentries = be; A.rows_cap = rows;
      smem = need;
      if (ctx->coarse_ready) {  // every block has its rows on chip: two-level preconditioner
        A.Einv = ctx->d_coarse + (size_t)nb * nb; A.rc = ctx->d_coarse + (size_t)2 * nb * nb;
        A.halo_blk = ctx->d_halo_blk_eff; A.cn_ptr = ctx->d_cn_ptr; A.cn_blk = ctx->d_cn_blk;
      }
    } else {
      A.rows_in_smem = 0; A.smem_entries = 0;
    }
  } else {
    // staging capacity: 12 bytes per stored value + 12 bytes per owned row
    const long long vper = (S.n_vertices + nb - 1) / nb;
    const long long rows_per_block = vper * S.ns;
    long long cap = ((long long)smem_max - 12 * (rows_per_block + 2)) / 12;
    cap &= ~1LL;
    const long long avg_entries = ctx->nnzb * S.npairs / nb + 1;
    A.rows_in_smem = (cap >= avg_entries) && !no_smem;
    A.smem_entries = A.rows_in_smem ? cap : 0;
    smem = A.rows_in_smem ? (size_t)(12 * cap + 12 * (rows_per_block + 2)) : 0;
  }
  FB_CHECK(ctx, cudaMemsetAsync(ctx->d_flags, 0, 32, ctx->stream));
  FB_CHECK(ctx, cudaMemsetAsync(A.bar, 0, 130 * sizeof(unsigned), ctx->stream));
  A.bar_base = 0u;
  void* args[] = {(void*)&S, (void*)&A};
  const void* fn = nullptr;
  switch (G) {
    case 2: fn = (const void*)k_cg_persistent<2>; break;
    case 4: fn = (const void*)k_cg_persistent<4>; break;
    case 8: fn = (const void*)k_cg_persistent<8>; break;
    case 16: fn = (const void*)k_cg_persistent<16>; break;
    default: fn = (const void*)k_cg_persistent<32>; break;
  }
  FB_CHECK(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  FB_CHECK(ctx, cudaLaunchCooperativeKernel(fn, dim3(nb), dim3(T), args, smem, ctx->stream));
  ctx->launches++;
  int fl[4];
  if ((rc = read_flags(ctx, fl))) return rc;
  unsigned hb[3] = {0, 0, 0};
  FB_CHECK(ctx, cudaMemcpy(&hb[0], A.bar, 4, cudaMemcpyDeviceToHost));
  FB_CHECK(ctx, cudaMemcpy(&hb[1], A.bar + 32, 4, cudaMemcpyDeviceToHost));
  FB_CHECK(ctx, cudaMemcpy(&hb[2], A.bar + 64, 4, cudaMemcpyDeviceToHost));
  if (hb[2]) {
    char msg[160];
    snprintf(msg, sizeof(msg), "persistent CG: grid barrier timed out (count=%u gen=%u blocks=%u)", hb[0], hb[1], nb);
    FB_CHECK(ctx, cudaMemset(A.bar, 0, 130 * sizeof(unsigned)));
    return fb_fail(ctx, msg);
  }
  *iters = fl[1];
  if (getenv("FB_VERBOSE") && fl[1] > 0) {
    double ph[5];
    FB_CHECK(ctx, cudaMemcpy(ph, ctx->d_sc + 8, sizeof(ph), cudaMemcpyDeviceToHost));
    fprintf(stderr, "[fb] persistent CG G=%d smem=%d coarse=%d its=%d: cycles/iteration spmv+dots %.0f, barrier %.0f, reduce %.0f, "
            "update %.0f, barrier %.0f\n", G, A.rows_in_smem, A.Einv != nullptr, fl[1], ph[0] / fl[1], ph[1] / fl[1], ph[2] / fl[1],
            ph[3] / fl[1], ph[4] / fl[1]);
  }
  if (fl[2]) return 1;
  return fl[0] ? 0 : 2;
}

static int solve_cg(fb_ctx* ctx, const double* J, const double* b, double* y, double thr2, int max_it,
                    int* iters) {
  const DevSpec& S = ctx->S;
  int rc = ensure_vectors(ctx, 5);
  if (rc) return rc;
  const long long N = S.n_dofs;
  double* r = ctx->d_work; double* u = r + N; double* w = u + N; double* p = w + N; double* s_ = p + N;
  const unsigned gn = (unsigned)((S.n_vertices + FB_TPB - 1) / FB_TPB);
  const int G = spmv_group(S);
  const unsigned gs = (unsigned)((N * G + FB_TPB - 1) / FB_TPB);
  if ((long long)gs * 3 > ctx->red_doubles) return fb_fail(ctx, "reduction buffer too small");
  FB_CHECK(ctx, cudaMemsetAsync(ctx->d_flags, 0, 32, ctx->stream));
  double hsc[8] = {0, 0, 0, 0, thr2, 0, 0, 0};
  FB_CHECK(ctx, cudaMemcpyAsync(ctx->d_sc, hsc, sizeof(hsc), cudaMemcpyHostToDevice, ctx->stream));
  FB_LAUNCH(ctx, k_cg_init, gn, FB_TPB, 0, S, ctx->d_dinv, b, y, r, u, p, s_);
  int fl[4] = {0, 0, 0, 0};
  int launched = 0;
  const int chunk = 32;
  while (launched <= max_it) {
    for (int k = 0; k < chunk; ++k) {
      switch (G) {
        case 4: FB_LAUNCH(ctx, k_cg_spmv_dots<4>, gs, FB_TPB, 0, S, J, u, r, w, ctx->d_red, ctx->d_ticket, ctx->d_sc, ctx->d_flags); break;
        case 8: FB_LAUNCH(ctx, k_cg_spmv_dots<8>, gs, FB_TPB, 0, S, J, u, r, w, ctx->d_red, ctx->d_ticket, ctx->d_sc, ctx->d_flags); break;
        case 16: FB_LAUNCH(ctx, k_cg_spmv_dots<16>, gs, FB_TPB, 0, S, J, u, r, w, ctx->d_red, ctx->d_ticket, ctx->d_sc, ctx->d_flags); break;
        default: FB_LAUNCH(ctx, k_cg_spmv_dots<32>, gs, FB_TPB, 0, S, J, u, r, w, ctx->d_red, ctx->d_ticket, ctx->d_sc, ctx->d_flags); break;
      }
      FB_LAUNCH(ctx, k_cg_update, gn, FB_TPB, 0, S, ctx->d_dinv, ctx->d_sc, ctx->d_flags, y, r, u, w, p, s_);
    }
    launched += chunk;
    FB_CHECK(ctx, cudaGetLastError());
    if ((rc = read_flags(ctx, fl))) return rc;
    if (fl[0]) break;
  }
  *iters = fl[1];
  if (fl[2]) return 1;      // breakdown
  return fl[0] ? 0 : 2;     // 2: not converged within max_it
}

// Eigenvalues of a real upper Hessenberg matrix (Francis double-shift QR with deflation, the
// EISPACK hqr iteration), a[i * lda + j] row-major, destroyed.  Used on the <= 30 x 30 Arnoldi
// matrix of a GMRES cycle: its Ritz values give the interval of the Chebyshev preconditioner.
static int hessenberg_eigenvalues(int n, double* a, int lda, double* wr, double* wi) {
#define HA(i, j) a[(size_t)((i) - 1) * lda + ((j) - 1)]
  auto sign = [](double x, double y) { return y >= 0.0 ? std::fabs(x) : -std::fabs(x); };
  double anorm = 0.0;
  for (int i = 1; i <= n; ++i)
    for (int j = (i - 1 > 1 ? i - 1 : 1); j <= n; ++j) anorm += std::fabs(HA(i, j));
  int nn = n;
  double t = 0.0, p = 0.0, q = 0.0, r = 0.0, x, y, z, w, u, v, sft;
  while (nn >= 1) {
    int its = 0, l;
    do {
      for (l = nn; l >= 2; --l) {
        double sl = std::fabs(HA(l - 1, l - 1)) + std::fabs(HA(l, l));
        if (sl == 0.0) sl = anorm;
        if (std::fabs(HA(l, l - 1)) + sl == sl) { HA(l, l - 1) = 0.0; break; }
      }
      x = HA(nn, nn);
      if (l == nn) {   // one root
        wr[nn - 1] = x + t; wi[nn - 1] = 0.0; --nn;
      } else {
        y = HA(nn - 1, nn - 1);
        w = HA(nn, nn - 1) * HA(nn - 1, nn);
        if (l == nn - 1) {   // two roots
          p = 0.5 * (y - x);
          q = p * p + w;
          z = std::sqrt(std::fabs(q));
          x += t;
          if (q >= 0.0) {
            z = p + sign(z, p);
            wr[nn - 2] = wr[nn - 1] = x + z;
            if (z != 0.0) wr[nn - 1] = x - w / z;
            wi[nn - 2] = wi[nn - 1] = 0.0;
          } else {
            wr[nn - 2] = wr[nn - 1] = x + p;
            wi[nn - 1] = z; wi[nn - 2] = -z;
          }
          nn -= 2;
        } else {
          if (its == 60) return -1;
          if (its == 10 || its == 20 || its == 40) {   // exceptional shift
            t += x;
            for (int i = 1; i <= nn; ++i) HA(i, i) -= x;
            sft = std::fabs(HA(nn, nn - 1)) + std::fabs(HA(nn - 1, nn - 2));
            y = x = 0.75 * sft;
            w = -0.4375 * sft * sft;
          }
          ++its;
          int m;
          for (m = nn - 2; m >= l; --m) {
            z = HA(m, m);
            r = x - z;
            sft = y - z;
            p = (r * sft - w) / HA(m + 1, m) + HA(m, m + 1);
            q = HA(m + 1, m + 1) - z - r - sft;
            r = HA(m + 2, m + 1);
            sft = std::fabs(p) + std::fabs(q) + std::fabs(r);
            p /= sft; q /= sft; r /= sft;
            if (m == l) break;
            u = std::fabs(HA(m, m - 1)) * (std::fabs(q) + std::fabs(r));
            v = std::fabs(p) * (std::fabs(HA(m - 1, m - 1)) + std::fabs(z) + std::fabs(HA(m + 1, m + 1)));
            if (u + v == v) break;
          }
          for (int i = m + 2; i <= nn; ++i) {
            HA(i, i - 2) = 0.0;
            if (i != m + 2) HA(i, i - 3) = 0.0;
          }
          for (int k = m; k <= nn - 1; ++k) {
            if (k != m) {
              p = HA(k, k - 1);
              q = HA(k + 1, k - 1);
              r = 0.0;
              if (k != nn - 1) r = HA(k + 2, k - 1);
              if ((x = std::fabs(p) + std::fabs(q) + std::fabs(r)) != 0.0) { p /= x; q /= x; r /= x; }
            }
            const double sg = sign(std::sqrt(p * p + q * q + r * r), p);
            if (sg != 0.0) {
              if (k == m) {
                if (l != m) HA(k, k - 1) = -HA(k, k - 1);
              } else {
                HA(k, k - 1) = -sg * x;
              }
              p += sg;
              x = p / sg; y = q / sg; z = r / sg;
              q /= p; r /= p;
              for (int j = k; j <= nn; ++j) {
                p = HA(k, j) + q * HA(k + 1, j);
                if (k != nn - 1) { p += r * HA(k + 2, j); HA(k + 2, j) -= p * z; }
                HA(k + 1, j) -= p * y;
                HA(k, j) -= p * x;
              }
              const int mmin = nn < k + 3 ? nn : k + 3;
              for (int i = l; i <= mmin; ++i) {
                p = x * HA(i, k) + y * HA(i, k + 1);
                if (k != nn - 1) { p += z * HA(i, k + 2); HA(i, k + 2) -= p * r; }
                HA(i, k + 1) -= p * q;
                HA(i, k) -= p;
              }
            }
          }
        }
      }
    } while (l < nn - 1);
  }
#undef HA
  return 0;
}

// test hook: eigenvalues of an n x n upper Hessenberg matrix (row-major, n <= 64)
extern "C" int fb_hessenberg_eigenvalues(int n, const double* a, double* wr, double* wi) {
  if (n < 1 || n > 64 || !a || !wr || !wi) return -1;
  std::vector<double> work(a, a + (size_t)n * n);
  return hessenberg_eigenvalues(n, work.data(), n, wr, wi);
}

// state of the Chebyshev preconditioner: out = {valid, lmin, lmax, operator applications per use}
extern "C" int fb_chebyshev_info(fb_ctx* ctx, double* out) {
  if (!ctx || !out) return -1;
  out[0] = ctx->cheb_valid ? 1.0 : 0.0; out[1] = ctx->cheb_lmin; out[2] = ctx->cheb_lmax; out[3] = ctx->cheb_steps;
  return 0;
}

// Interval of the Chebyshev preconditioner from the Ritz values of the last plain GMRES cycle
// (nc columns of the unrotated Hessenberg matrix, left on the device by k_gmres_givens).
static int cheb_estimate(fb_ctx* ctx, int nc) {
  if (nc < 6) return 0;
  std::vector<double> raw((size_t)(GM_M + 1) * GM_M);
  FB_CHECK(ctx, cudaMemcpyAsync(raw.data(), ctx->d_sc + GM_HRAW, raw.size() * sizeof(double), cudaMemcpyDeviceToHost,
                                ctx->stream));
  FB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  std::vector<double> H((size_t)nc * nc, 0.0), wr(nc), wi(nc);
  for (int j = 0; j < nc; ++j)
    for (int i = 0; i < nc && i <= j + 1; ++i) H[(size_t)i * nc + j] = raw[(size_t)j * (GM_M + 1) + i];
  if (hessenberg_eigenvalues(nc, H.data(), nc, wr.data(), wi.data())) return 0;
  double lo = 1e300, hi = 0.0, im = 0.0;
  for (int i = 0; i < nc; ++i) {
    if (!(wr[i] == wr[i])) return 0;
    const double mod = std::sqrt(wr[i] * wr[i] + wi[i] * wi[i]);
    if (wr[i] < lo) lo = wr[i];
    if (mod > hi) hi = mod;
    if (std::fabs(wi[i]) > im) im = std::fabs(wi[i]);
  }
  // a polynomial preconditioner needs the spectrum in the right half plane and close to the real axis
  if (!(lo > 0.0) || im > 0.3 * hi) { ctx->cheb_valid = false; ctx->cheb_refused = true; return 0; }
  ctx->cheb_lmax = 1.1 * hi;                 // beyond lmax the polynomial grows: keep a margin
  ctx->cheb_lmin = 0.9 * lo;
  if (ctx->cheb_lmin < ctx->cheb_lmax / 100.0) ctx->cheb_lmin = ctx->cheb_lmax / 100.0;
  const double sigma = (ctx->cheb_lmax + ctx->cheb_lmin) / (ctx->cheb_lmax - ctx->cheb_lmin);
  int k = (int)std::ceil(std::acosh(20.0) / std::acosh(sigma));   // |residual polynomial| <= 1/20 on the interval
  if (k < 2) k = 2;
  if (k > 8) k = 8;
  if (getenv("FB_CHEB_DEGREE")) k = atoi(getenv("FB_CHEB_DEGREE"));
  ctx->cheb_steps = k;
  ctx->cheb_valid = true;
  ctx->cheb_inv_dt = ctx->S.inv_dt;
  if (getenv("FB_VERBOSE"))
    fprintf(stderr, "[fb] Chebyshev interval from %d Ritz values: [%.4g, %.4g] (max |Im| %.3g), %d operator applications\n", nc,
            ctx->cheb_lmin, ctx->cheb_lmax, im, k);
  return 0;
}

__global__ void k_cheb_init(long long n, const double* __restrict__ v, double inv_theta, double* __restrict__ d,
                            double* __restrict__ z, double* __restrict__ r, const int* __restrict__ flags) {
  if (flags && flags[0]) return;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
    const double vk = v[k], dk = vk * inv_theta;
    d[k] = dk; z[k] = dk; r[k] = vk;
  }
}

// Operator applications per Krylov iteration for a solve that has to reduce the (preconditioned) residual
// by `reduction`: with the interval known, m Krylov iterations with k applications each reduce it by about
// 1 / T_k(sigma)^m.  Measured on C2 (needed reduction 6e-8): k = 5 -> 5 iterations (as the bound says), k = 7 ->
// 3 (bound: 9.9e-8, 4 in one solve of six), k = 10 -> 2 (bound: 1.2e-7, 3 in one solve of three) - GMRES runs
// up to a factor 2 ahead of the bound, with a rising chance of one more iteration; but k = 11 and 12 (bound^2 well
// below the target) still took 3 iterations in half of the solves: eigenvalues outside the Ritz interval are not
// damped by the polynomial and cost a Krylov iteration each.  The rule takes the (m, k), m >= 3, k <= 12, of the least
// expected cost, an orthogonalisation step counted as 0.7 applications; *m_out = the iterations expected.
static int cheb_choose_steps_interval(double lmin, double lmax, int fallback_k, double reduction, int* m_out) {
  const double sigma = (lmax + lmin) / (lmax - lmin);
  const double a = std::acosh(sigma);
  if (!(reduction < 0.5)) reduction = 0.5;
  if (reduction < 1e-16) reduction = 1e-16;
  int best_k = fallback_k, best_m = 0;
  double best_cost = 1e300;
  for (int m = 3; m <= 8; ++m)
    for (int k = 2; k <= 12; ++k) {   // longer recurrences would trust the interval too much
      // bound on the reduction of m iterations, relative to what is needed; the chance that one more
      // iteration is needed rises steeply once the bound alone does not reach the target
      const double ratio = std::exp(-m * std::log(std::cosh(k * a))) / reduction;
      double miss;
      if (ratio <= 1.0) miss = 0.0;
      else if (ratio <= 1.7) miss = 0.2;
      else if (ratio <= 2.5) miss = 0.5;
      else continue;
      const double cost = (m + miss) * (k + 0.7);
      if (cost < best_cost) { best_cost = cost; best_k = k; best_m = m; }
    }
  if (best_m == 0) { best_k = 12; best_m = 8; }
  *m_out = best_m;
  return best_k;
}

static int cheb_choose_steps(const fb_ctx* ctx, double reduction, int* m_out) {
  return cheb_choose_steps_interval(ctx->cheb_lmin, ctx->cheb_lmax, ctx->cheb_steps, reduction, m_out);
}

extern "C" int fb_chebyshev_steps(double lambda_min, double lambda_max, double reduction, int* krylov_iterations) {
  if (!(lambda_min > 0.0) || !(lambda_max > lambda_min) || !krylov_iterations) return -1;
  return cheb_choose_steps_interval(lambda_min, lambda_max, 5, reduction, krylov_iterations);
}

// z = p(B) v with p the Chebyshev polynomial of the interval (cheb_steps - 1 applications of
// B = D^-1 J), and - when w is given - w = B z from one more application (w = v - r).
static int cheb_apply(fb_ctx* ctx, const double* v, double* z, double* w, double* d0, double* d1, double* r,
                      const int* flags) {
  const long long No = ctx->S.n_owned_dofs;
  const double theta = 0.5 * (ctx->cheb_lmax + ctx->cheb_lmin), delta = 0.5 * (ctx->cheb_lmax - ctx->cheb_lmin);
  const double sigma = theta / delta;
  double rho = 1.0 / sigma;
  FB_LAUNCH(ctx, k_cheb_init, vec_grid(ctx, No), FB_TPB, 0, No, v, 1.0 / theta, d0, z, r, flags);
  const int m = ctx->cheb_steps - 1;
  const int steps = w ? m + 1 : m;
  double* din = d0;
  double* dout = d1;
  for (int i = 0; i < steps; ++i) {
    const double rho_n = 1.0 / (2.0 * sigma - rho);
    SpmvChArgs a{};
    a.x = din; a.dinv32 = ctx->d_dinv32; a.dscale = ctx->d_dscale; a.flags = flags;
    a.r = r; a.z = z; a.dout = dout; a.vin = v; a.wout = w;
    a.ca = rho_n * rho; a.cb = 2.0 * rho_n / delta;
    a.last = (w && i == m) ? 1 : 0;
    int rc = fbi_halo(ctx, din);
    if (rc) return rc;
    rc = spmv_ch_tma_launch<2>(ctx, a);
    if (rc) return rc;
    rho = rho_n;
    double* tmp = din; din = dout; dout = tmp;
  }
  return 0;
}

static int solve_gmres(fb_ctx* ctx, const double* J, const double* b, double* y, double thr, int max_it,
                       int* iters, bool want_cheb, bool use_guess) {
  const DevSpec& S = ctx->S;
  int rc = ensure_vectors(ctx, GM_NVEC);
  if (rc) return rc;
  const long long N = S.n_dofs, No = S.n_owned_dofs;   // vector stride / entries this rank owns
  double* V = ctx->d_work;                 // (GM_M+1) vectors
  double* w = V + (size_t)(GM_M + 1) * N;
  double* z = w + N;
  double* r = z + N;
  double* cd0 = r + N;                     // Chebyshev recurrence: direction (ping-pong), residual, sum
  double* cd1 = cd0 + N;
  double* cr = cd1 + N;
  double* cz = cr + N;
  // flexible storage: Z_j = p(B) V_j is what the Chebyshev recurrence leaves in its sum vector anyway -
  // kept per Krylov vector (behind the three vectors of the refinement loop) the update y += Z yk needs no
  // second application of the polynomial (4 operator applications less per solve)
  double* Zf = ctx->work_doubles >= (long long)(GM_NVEC + 3 + GM_M) * N ? V + (size_t)(GM_NVEC + 3) * N : nullptr;
  const unsigned gv = vec_grid(ctx, No);
  double* sc = ctx->d_sc;
  // use_guess: y holds an initial guess (KSPSetInitialGuessNonzero), e.g. the Newton update of the
  // previous time step; the stopping threshold is absolute, so a guess only saves iterations
  if (!use_guess) FB_CHECK(ctx, cudaMemsetAsync(y, 0, sizeof(double) * N, ctx->stream));
  FB_CHECK(ctx, cudaMemsetAsync(ctx->d_flags, 0, 32, ctx->stream));
  int total = 0;
  int fl[4] = {0, 0, 0, 0};
  bool first = !use_guess, guess_checked = false;
  const bool cheb_possible = want_cheb && !J && spmv_ch_tma_ok(ctx, V) && ctx->dinv32_valid && !ctx->cheb_refused;
  while (total < max_it) {
    // right-preconditioned with the Chebyshev polynomial p(B) of B = D^-1 J once its interval is
    // known (Ritz values of a plain cycle); the Arnoldi residual is the block-Jacobi one either way
    const bool cheb = cheb_possible && ctx->cheb_valid;
    // r = D^-1 (b - A y)   (left preconditioning: the Arnoldi residual is the
    // preconditioned one, PETSc's default norm for GMRES)
    if (first) {
      precond_launch(ctx, b, r, (const int*)nullptr);
    } else {
      if ((rc = fbi_spmv(ctx, J, y, w))) return rc;
      FB_LAUNCH(ctx, k_residual, gv, FB_TPB, 0, No, b, w, w);
      precond_launch(ctx, w, r, (const int*)nullptr);
    }
    first = false;
    FB_LAUNCH(ctx, k_dot, gv, FB_TPB, 0, No, r, r, ctx->d_red, ctx->d_ticket, sc + GM_MISC, xr_next(ctx));
    FB_LAUNCH(ctx, k_gmres_begin, 1, 1, 0, sc, ctx->d_flags, thr, (use_guess && total == 0 && !first) ? ctx->guess_guard : 0.0);
    if (use_guess && total == 0 && !first && !guess_checked) {
      guess_checked = true;
      int f5[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      FB_CHECK(ctx, cudaMemcpyAsync(ctx->h_pinned + 24, ctx->d_flags, 8 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      FB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
      memcpy(f5, ctx->h_pinned + 24, sizeof(f5));
      if (f5[4]) {   // worse than starting from zero: do that
        FB_CHECK(ctx, cudaMemsetAsync(y, 0, sizeof(double) * N, ctx->stream));
        FB_CHECK(ctx, cudaMemsetAsync(ctx->d_flags, 0, 32, ctx->stream));
        first = true;
        continue;
      }
    }
    int poll = cheb ? 4 : 10;
    if (cheb && !getenv("FB_CHEB_DEGREE")) {
      // degree of the polynomial from the reduction this solve still needs (|r0| from the device, one word)
      FB_CHECK(ctx, cudaMemcpyAsync(ctx->h_pinned + 40, sc + GM_MISC, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
      FB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
      const double r0 = std::sqrt(ctx->h_pinned[40] > 0.0 ? ctx->h_pinned[40] : 0.0);
      int m_exp = 4;
      ctx->cheb_steps = cheb_choose_steps(ctx, r0 > 0.0 ? thr / r0 : 1.0, &m_exp);
      poll = m_exp;   // first look at the flags when the solve is expected to be done
      if (getenv("FB_VERBOSE"))
        fprintf(stderr, "[fb] Chebyshev: reduction %.2e needed -> %d applications x %d Krylov iterations expected\n",
                r0 > 0.0 ? thr / r0 : 1.0, ctx->cheb_steps, m_exp);
    }
    FB_LAUNCH(ctx, k_scale_to, gv, FB_TPB, 0, No, r, sc + GM_MISC + 1, V, (const int*)nullptr);
    for (int j = 0; j < GM_M; ++j) {
      // w = D^-1 J [p(B)] V_j: SpMV with the block-Jacobi application fused into its epilogue
      if (cheb) {
        if ((rc = cheb_apply(ctx, V + (size_t)j * N, Zf ? Zf + (size_t)j * N : cz, w, cd0, cd1, cr, ctx->d_flags))) return rc;
      } else if (!J) {
        if ((rc = spmv_ch(ctx, V + (size_t)j * N, w, true, ctx->d_flags))) return rc;
      } else if (spmv_blk_ok(S)) {
        if ((rc = fbi_halo(ctx, V + (size_t)j * N))) return rc;
        if ((rc = spmv_blk_launch<true>(ctx, J, V + (size_t)j * N, w, ctx->d_flags))) return rc;
      } else {
        if ((rc = fbi_spmv(ctx, J, V + (size_t)j * N, z))) return rc;
        precond_launch(ctx, z, w, ctx->d_flags);
      }
      const int nv = j + 1;
#define FB_GS_CASE(NVB)                                                                                        \
  {                                                                                                            \
    FB_LAUNCH(ctx, k_gs_dots<NVB>, gv, FB_TPB, 0, No, N, nv, V, w, ctx->d_red, ctx->d_ticket, sc + GM_HD1,     \
              sc + GM_MISC, ctx->d_flags, xr_next(ctx));                                                       \
    FB_LAUNCH(ctx, k_gs_update<NVB>, gv, FB_TPB, 0, No, N, nv, V, sc + GM_HD1, w, ctx->d_red, ctx->d_ticket,   \
              sc + GM_HD2, sc + GM_MISC, ctx->d_flags, xr_next(ctx));                                          \
    FB_LAUNCH(ctx, k_gs_refine<NVB>, gv, FB_TPB, 0, No, N, nv, V, sc + GM_HD2, w, ctx->d_red, ctx->d_ticket,   \
              sc + GM_MISC, ctx->d_flags, xr_next(ctx));                                                       \
  }
      if (nv <= 8) FB_GS_CASE(8)
      else if (nv <= 16) FB_GS_CASE(16)
      else if (nv <= 24) FB_GS_CASE(24)
      else FB_GS_CASE(32)
#undef FB_GS_CASE
      FB_LAUNCH(ctx, k_gmres_givens, 1, 1, 0, j, sc, ctx->d_flags);
      FB_LAUNCH(ctx, k_scale_to, gv, FB_TPB, 0, No, w, sc + GM_MISC + 1, V + (size_t)(j + 1) * N, ctx->d_flags);
      if (j % poll == poll - 1 || j == GM_M - 1 || (cheb && j >= poll)) {
        FB_CHECK(ctx, cudaGetLastError());
        if ((rc = read_flags(ctx, fl))) return rc;
        if (fl[0]) break;
      }
    }
    total = fl[1];
    // y += [p(B)] V yk
    if (fl[3] > 0) {
      FB_LAUNCH(ctx, k_gmres_solve_y, 1, 1, 0, sc, ctx->d_flags);
      FB_LAUNCH(ctx, k_combine, gv, FB_TPB, 0, No, N, (cheb && Zf) ? Zf : V, sc + GM_Y, ctx->d_flags, w);
      if (cheb && !Zf) {
        if ((rc = cheb_apply(ctx, w, cz, nullptr, cd0, cd1, cr, (const int*)nullptr))) return rc;
        FB_LAUNCH(ctx, k_axpy, gv, FB_TPB, 0, No, 1.0, cz, y);
      } else {
        FB_LAUNCH(ctx, k_axpy, gv, FB_TPB, 0, No, 1.0, w, y);
      }
    }
    if (cheb_possible && !cheb && !ctx->cheb_valid && (rc = cheb_estimate(ctx, fl[3]))) return rc;
    if (cheb && !fl[0] && !fl[2]) {
      // a full cycle without convergence: the interval is off (the operator changed); back to
      // block-Jacobi for the rest of this solve, which also re-estimates the interval
      ctx->cheb_valid = false;
    }
    if (fl[0] || fl[2]) break;
  }
  *iters = total;
  FB_CHECK(ctx, cudaGetLastError());
  if (fl[2]) return 1;
  return fl[0] ? 0 : 2;
}

static int solve_tridiag(fb_ctx* ctx, const double* J, const double* b, double* y) {
  const DevSpec& S = ctx->S;
  int rc;
  if (!ctx->d_tri) {
    if ((rc = fb_alloc(ctx, &ctx->d_tri, (size_t)S.n_vertices * 4 * S.ns * S.ns))) return rc;
  }
  const unsigned gn = (unsigned)((S.n_vertices + FB_TPB - 1) / FB_TPB);
  FB_LAUNCH(ctx, k_tri_extract, gn, FB_TPB, 0, S, J, ctx->d_tri);
  FB_LAUNCH(ctx, k_tri_solve, 1, 32, 0, S, ctx->d_tri, b, y, ctx->d_tri + (size_t)S.n_vertices * 3 * S.ns * S.ns);
  FB_CHECK(ctx, cudaGetLastError());
  return 0;
}

int fbi_linear_solve(fb_ctx* ctx, const double* J, const double* b, double* y, int ksp, int pc,
                     double rtol, double atol, int max_it, int* iters, double* relres) {
  const DevSpec& S = ctx->S;
  int rc = fbi_ensure_workspace(ctx);
  if (rc) return rc;
  *iters = 0;
  *relres = 0.0;
  if (ksp == FB_KSP_AUTO) ksp = ctx->tridiag ? FB_KSP_TRIDIAG : (ctx->symmetric ? FB_KSP_CG : FB_KSP_GMRES);
  if (!J && !S.vx_on) return fb_fail(ctx, "no channel operator: pass the block-CSR values");
  if (!J && ksp != FB_KSP_GMRES) return fb_fail(ctx, "the channel operator is solved with GMRES");
  if (ctx->peer_ready && ksp != FB_KSP_GMRES)
    return fb_fail(ctx, "partitioned meshes: fb_linear_solve runs GMRES (the partitioned CG is fb_dcg_persistent)");
  if (ksp == FB_KSP_TRIDIAG) {
    if (!ctx->tridiag) return fb_fail(ctx, "tridiagonal solver needs an interval mesh ordered along x");
    return solve_tridiag(ctx, J, b, y);
  }
  // pc: block-Jacobi (0), Chebyshev polynomial of the block-Jacobi-scaled operator (1), or the
  // default (2): Chebyshev for large systems held as channel operator, block-Jacobi otherwise
  const bool want_cheb = ksp == FB_KSP_GMRES && !getenv("FB_NO_CHEBYSHEV") &&
                         (pc == FB_PC_CHEBYSHEV || (pc == FB_PC_AUTO && S.n_dofs >= 200000));
  if (ctx->cheb_valid && ctx->cheb_inv_dt > 0.0 &&
      (S.inv_dt > 2.0 * ctx->cheb_inv_dt || S.inv_dt < 0.5 * ctx->cheb_inv_dt))
    ctx->cheb_valid = false;   // the time step moved: the spectrum of D^-1 J with it
  double bnorm;
  if ((rc = fbi_norm2(ctx, b, &bnorm))) return rc;
  if (bnorm == 0.0) {
    FB_CHECK(ctx, cudaMemsetAsync(y, 0, sizeof(double) * S.n_dofs, ctx->stream));
    return 0;
  }
  const unsigned gn = (unsigned)((S.n_vertices + FB_TPB - 1) / FB_TPB);
  if ((rc = bjacobi_setup(ctx, J))) return rc;
  if (ksp == FB_KSP_CG && (rc = coarse_setup(ctx, J))) return rc;
  // stopping rule on the preconditioned residual, ||M^-1 r|| <= max(rtol ||M^-1 b||, atol)
  // (PETSc default for left-preconditioned CG/GMRES); block-Jacobi makes it row-scaled,
  // which matters when species equations differ by many orders of magnitude
  const int base = (ksp == FB_KSP_CG) ? 5 : GM_NVEC;  // Krylov vectors, then Ay, t, corr
  const int flex = (ksp == FB_KSP_GMRES && !J && !getenv("FB_NO_FLEX")) ? GM_M : 0;   // Z_j = p(B) V_j of the polynomial preconditioner
  if ((rc = ensure_vectors(ctx, base + 3 + flex))) return rc;
  precond_launch(ctx, b, ctx->d_work, (const int*)nullptr);
  double pbnorm;
  if ((rc = fbi_norm2(ctx, ctx->d_work, &pbnorm))) return rc;
  double thr = rtol * pbnorm;
  if (atol > thr) thr = atol;
  ctx->guess_guard = pbnorm;   // the residual of the zero vector
  // Krylov solve + iterative refinement on the true residual: the Newton test is on
  // the unpreconditioned ||F||, so the true relative residual must also reach rtol
  // (the reference's LU solve leaves ~1e-15); each round solves J d = b - J y.
  double* Ay = ctx->d_work + (size_t)base * S.n_dofs;
  double* t = Ay + S.n_dofs;
  double* corr = t + S.n_dofs;
  int status = 0;
  double rn = 0.0;
  for (int round = 0; round < 3; ++round) {
    int its = 0;
    const double* rhs = b;
    double* sol = y;
    double thr_r = thr, rel = thr / pbnorm;
    if (round > 0) {
      // correction equation J d = b - J y, solved to 1e-3 of its own preconditioned norm
      rhs = t; sol = corr;
      precond_launch(ctx, t, corr, (const int*)nullptr);
      double pn;
      if ((rc = fbi_norm2(ctx, corr, &pn))) return rc;
      thr_r = 1e-3 * pn;
      rel = 1e-3;
    }
    if (ksp == FB_KSP_CG) {
      if (getenv("FB_CG_MULTIKERNEL")) {
        status = solve_cg(ctx, J, rhs, sol, thr_r * thr_r, max_it, &its);
      } else {
        status = solve_cg_persistent(ctx, J, rhs, sol, thr_r * thr_r, rel * rel, max_it, &its);
        if (status == 1 && ctx->coarse_ready) {  // singular coarse operator: one-level fallback
          ctx->coarse_ready = false;
          *iters += its;
          status = solve_cg_persistent(ctx, J, rhs, sol, thr_r * thr_r, rel * rel, max_it, &its);
        }
      }
    }
    else status = solve_gmres(ctx, J, rhs, sol, thr_r, max_it, &its, want_cheb, round == 0 && ctx->use_guess);
    *iters += its;
    if (status < 0) return status;
    if (round > 0) FB_LAUNCH(ctx, k_axpy, vec_grid(ctx, S.n_owned_dofs), FB_TPB, 0, S.n_owned_dofs, 1.0, corr, y);
    if ((rc = fbi_spmv(ctx, J, y, Ay))) return rc;
    FB_LAUNCH(ctx, k_residual, vec_grid(ctx, S.n_owned_dofs), FB_TPB, 0, S.n_owned_dofs, b, Ay, t);
    if ((rc = fbi_norm2(ctx, t, &rn))) return rc;
    *relres = rn / bnorm;
    if (status != 0 || *relres <= rtol || !(rn == rn)) break;
  }
  if (status == 1) { ctx->err = "Krylov breakdown"; return 1; }
  if (status == 2) { ctx->err = "Krylov solver did not converge"; return 2; }
  return 0;
}

// ---------------------------------------------------------------------------
// Building blocks of the partitioned (multi-GPU) CG.  Every rank owns the vertices
// [0, n_owned) of its local numbering (fb_set_owned); the host driver
// (festim_b200/distributed.py) refreshes the ghost entries of u through NCCL before
// fb_dcg_spmv_dots and all-reduces the three partial dot products it returns.  The
// recurrence is the same single-reduction (Chronopoulos-Gear) form the single-GPU
// kernels use, so one all-reduce of three doubles per iteration is all that crosses
// NVLink besides the halo.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(FB_TPB)
k_dcg_update(DevSpec S, const double* __restrict__ dinv, double alpha, double beta, double* __restrict__ x,
             double* __restrict__ r, double* __restrict__ u, const double* __restrict__ w,
             double* __restrict__ p, double* __restrict__ s_) {
  const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (v >= S.n_owned) return;
  double rin[FB_MAX_NS], zo[FB_MAX_NS];
  for (int s = 0; s < S.ns; ++s) {
    const long long i = v * S.ns + s;
    const double pi = u[i] + beta * p[i];
    const double si = w[i] + beta * s_[i];
    p[i] = pi; s_[i] = si;
    x[i] += alpha * pi;
    rin[s] = r[i] - alpha * si;
    r[i] = rin[s];
  }
  bjacobi_apply(S.ns, dinv, v, rin, zo);
  for (int s = 0; s < S.ns; ++s) u[v * S.ns + s] = zo[s];
}

// ---------------------------------------------------------------------------
// Partitioned persistent CG over peer memory (NVLink / NVSwitch).  One cooperative
// launch per rank runs the whole solve.  Everything that crosses NVLink is written as
// self-validating 8-byte words {tag : 32 | half a double : 32} straight into the
// consumer's memory (the protocol NCCL calls LL): no fence, no flag, no round trip -
// the reader spins until both halves carry the tag of the iteration it is in.
//   * halo: in the vector update every block pushes the entries of u that other ranks
//     hold as ghosts into their halo words; the neighbour's SpMV reads a ghost column
//     from its halo words.  A push of version k+1 cannot overtake a reader of version
//     k, because pushes happen after the reduction of iteration k, which every rank
//     enters only after its SpMV k.
//   * reduction: the last block to arrive at the local grid barrier sums the rank's
//     partial dot products, writes the three totals into a slot of EVERY rank's
//     exchange area and collects the slots of the other ranks before it opens the
//     local barrier.  Every rank adds the per-rank totals in rank order, so all ranks
//     take bit-identical decisions (convergence, breakdown) and leave together.
// Per iteration: one local grid barrier and one cross-rank rendezvous; no NCCL call and
// no host round trip inside the solve.
// ---------------------------------------------------------------------------
#define FB_XR 0       // u64 words: reduction slots [2 parities][8 ranks][3 values][2 halves]
struct PeerArgs {
  int rank, nranks;
  unsigned seq;               // tag of version 0 of this solve (tags never repeat across solves)
  double* base[8];            // exchange area of every rank (own one included)
  const int* send_v; const int* send_rank; const int* send_dst; int n_send;  // sorted by send_v
};

// Local grid barrier + cross-rank exchange of the three dot products, run by warp 0 of
// every block.  On return tot3 (shared) holds the sums over all ranks.
__device__ __forceinline__ bool xgrid_reduce(unsigned* bar, unsigned nblocks, unsigned& gen, unsigned base_count,
                                             const PeerArgs& P, const double* partials, unsigned tag, unsigned parity,
                                             double* tot3, unsigned long long* dbg) {
  __shared__ int xok_s;
  __syncthreads();
  ++gen;
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    unsigned old = 0;
    if (lane == 0) old = atom_add_release_u32(bar, 1u);
    old = __shfl_sync(0xffffffffu, old, 0);
    int ok = 1;
    if (old - base_count == gen * nblocks - 1u) {  // last block of this rank
      const long long d0 = clock64();
      // all loads in flight before the first add: a warp executes in order, so a
      // load-add loop would pay one L2 round trip per term (15 of them)
      double t[3], pv[3][5];
#pragma unroll
      for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          const unsigned b = lane + 32 * j;
          pv[k][j] = b < nblocks ? __ldcg(partials + (size_t)k * nblocks + b) : 0.0;
        }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double sacc = (((pv[k][0] + pv[k][1]) + pv[k][2]) + pv[k][3]) + pv[k][4];
        for (unsigned b = lane + 160; b < nblocks; b += 32) sacc += __ldcg(partials + (size_t)k * nblocks + b);
        t[k] = warp_sum(sacc);
      }
      const long long d1 = clock64();
      // lane -> (destination rank q, value k): slot [parity][my rank][k] of rank q
      if (lane < 3 * P.nranks) {
        const int q = lane / 3, k = lane - 3 * q;
        unsigned long long* dst = (unsigned long long*)P.base[q] + FB_XR + ((parity * 8 + P.rank) * 3 + k) * 2;
        ll_store(dst, tag, k == 0 ? t[0] : (k == 1 ? t[1] : t[2]));
      }
      // collect: slot [parity][q][k] of my own area
      if (lane < 3 * P.nranks) {
        const int q = lane / 3, k = lane - 3 * q;
        const unsigned long long* src =
            (const unsigned long long*)P.base[P.rank] + FB_XR + ((parity * 8 + q) * 3 + k) * 2;
        double v;
        ok = ll_load(src, tag, bar, v);
      }
      ok = __all_sync(0xffffffffu, ok);
      if (lane == 0 && dbg) {  // FB_VERBOSE: where the last block spends the rendezvous
        const long long d2 = clock64();
        atomicAdd(dbg, (unsigned long long)(d1 - d0));
        atomicAdd(dbg + 1, (unsigned long long)(d2 - d1));
      }
      if (lane == 0) st_release_u32(bar + 32, gen);  // opens the local barrier (on failure bar[64] is set)
    } else if (lane == 0) {
      unsigned spins = 0;
      while (ld_acquire_u32(bar + 32) != gen) {
        __nanosleep(64);
        if ((++spins & 0xfffu) == 0u) {
          if (spins > (1u << 24)) { atomicExch(bar + 64, 1u); ok = 0; break; }
          if (ld_acquire_u32(bar + 64)) { ok = 0; break; }
        }
      }
      if (ok && ld_acquire_u32(bar + 64)) ok = 0;
    }
    ok = __shfl_sync(0xffffffffu, ok, 0);
    // every block adds the per-rank totals in rank order (the words are complete: the last
    // block saw them before it opened the barrier)
    {
      double val = 0.0;  // lane -> (rank q, value k): every word is fetched at once
      if (ok && lane < 3 * P.nranks) {
        const int q = lane / 3, k = lane - 3 * q;
        const unsigned long long* src =
            (const unsigned long long*)P.base[P.rank] + FB_XR + ((parity * 8 + q) * 3 + k) * 2;
        const unsigned long long lo = ll_word(src), hi = ll_word(src + 1);
        val = __longlong_as_double((long long)((hi << 32) | (lo & 0xffffffffull)));
      }
      double sacc = 0.0;
      for (int q = 0; q < P.nranks; ++q) sacc += __shfl_sync(0xffffffffu, val, q * 3 + (lane % 3));  // rank order
      if (ok && lane < 3) tot3[lane] = sacc;
    }
    if (lane == 0) xok_s = ok;
  }
  __syncthreads();
  return xok_s != 0;
}

template <int G>
__global__ void __launch_bounds__(1024, 1)
k_dcg_persistent(DevSpec S, CgArgs A, PeerArgs P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double red[3][32];
  __shared__ double tot[3];
  __shared__ int srange[2];
  const int ns = S.ns;
  const unsigned nb = gridDim.x;
  const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = T >> 5;
  const long long n_own = S.n_owned;
  // owned rows of this block: a compact RCB part (fb_set_sm_partition over the owned
  // vertices) or, without one, a contiguous equal range
  const bool part = A.blk_rng != nullptr;
  const long long vper = (n_own + nb - 1) / nb;
  const long long v0 = part ? A.blk_rng[blockIdx.x] : ((long long)blockIdx.x * vper < n_own ? (long long)blockIdx.x * vper : n_own);
  const long long v1 = part ? A.blk_rng[blockIdx.x + 1] : ((v0 + vper < n_own) ? v0 + vper : n_own);
  const int nvloc = (int)(v1 - v0);
  const int nrows = nvloc * ns;
  const long long row0 = v0 * ns;
  double* const uvec = A.u;                      // owned entries of the iterate
  const long long n_own_dofs = n_own * ns;
  const unsigned long long* const halo = (const unsigned long long*)P.base[P.rank] + FB_XH;  // ghost entries
  // shared memory: vals[cap] f64 | su[su_doubles] f64 | cols: int[cap] (range mode) or
  // u16[bentries] (partition mode) | rstart[rows] | rlen[rows] | rbst[rows]
  const long long rows_cap = part ? A.rows_cap : vper * ns;
  double* svals = (double*)smem_raw;
  double* su = svals + A.smem_entries;
  int* scols = (int*)(su + A.su_doubles);
  unsigned short* slcol = (unsigned short*)scols;
  int* rstart = part ? (int*)(slcol + ((A.bentries + 1) & ~1LL)) : scols + A.smem_entries;
  int* rlen = rstart + rows_cap;
  int* rbst = rlen + rows_cap;
  unsigned char* rfree = (unsigned char*)(rbst + rows_cap);  // 1: row is in the coarse space
  bool staged = false;
  int nhalo = 0;
  const int* hidx = nullptr;
  if (A.rows_in_smem && nrows > 0) {
    const long long e0 = (long long)S.rowptr[v0] * S.npairs, e1 = (long long)S.rowptr[v1] * S.npairs;
    staged = (e1 - e0) <= A.smem_entries;
    if (part) {
      nhalo = A.halo_ptr[blockIdx.x + 1] - A.halo_ptr[blockIdx.x];
      hidx = A.halo_idx + A.halo_ptr[blockIdx.x];
      staged = staged && (long long)(nvloc + nhalo) * ns <= A.su_doubles &&
               (long long)(S.rowptr[v1] - S.rowptr[v0]) <= A.bentries;
    }
    if (staged) {
      for (long long e = tid; e < e1 - e0; e += T) svals[e] = A.J[e0 + e];
      if (part) {
        const long long b0 = S.rowptr[v0], nbe = S.rowptr[v1] - b0;
        for (long long e = tid; e < nbe; e += T) slcol[e] = A.lcol[b0 + e];
      }
      for (int rr = tid; rr < nrows; rr += T) {
        const int vl = rr / ns, sp = rr - vl * ns;
        const long long v = v0 + vl;
        const int r0 = S.rowptr[v], deg = S.rowptr[v + 1] - r0;
        const int nc = S.pair_nc[sp], pp = S.pair_pp[sp];
        const int start = (int)((long long)r0 * S.npairs + pp - e0);   // block layout: entry (k, j) at start + k * npairs + j
        rstart[rr] = start;
        rlen[rr] = deg * nc;
        rbst[rr] = r0 - S.rowptr[v0];
        if (!part)
          for (int e = 0; e < deg * nc; ++e) {
            const int k = e / nc, j = e - k * nc;
            scols[start + k * S.npairs + j] = S.colidx[r0 + k] * ns + S.pair_cs[pp + j];
          }
      }
    }
  }
  const bool cached = staged && part;  // iterate gathered from shared memory
  // ---- rank-local coarse space (one constant per block of THIS rank; couplings to other
  // ranks are left to the Jacobi part): same construction as in k_cg_persistent ----------
  __shared__ double ysm[160], rcs[160];
  const bool coarse = A.Einv != nullptr && cached && ns == 1;
  const int cn0 = coarse ? A.cn_ptr[blockIdx.x] : 0;
  const int nneed = coarse ? A.cn_ptr[blockIdx.x + 1] - cn0 + 1 : 0;  // self + neighbours
  const unsigned short* hblk = coarse ? A.halo_blk + A.halo_ptr[blockIdx.x] : nullptr;
  int ctarget = -1;
  double erow[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  if (coarse) {
    for (int rr = tid; rr < nrows; rr += T) rfree[rr] = S.bc_map[row0 + rr] < 0;
    if (tid == 0) ysm[FB_COARSE_NONE] = 0.0;
    if (wid < nneed) {
      ctarget = wid == 0 ? (int)blockIdx.x : A.cn_blk[cn0 + wid - 1];
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        const unsigned c = lane + 32 * q;
        erow[q] = c < nb ? A.Einv[(size_t)ctarget * nb + c] : 0.0;
      }
    }
  }
  double rsum = 0.0;
  // this block's slice of the send table (entries sorted by owned vertex)
  if (tid < 2) {
    const long long key = tid == 0 ? v0 : v1;
    int lo = 0, hi = P.n_send;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (P.send_v[mid] < key) lo = mid + 1; else hi = mid; }
    srange[tid] = lo;
  }
  // ---- init ------------------------------------------------------------------------
  for (long long v = v0 + tid; v < v1; v += T) {
    double rin[FB_MAX_NS], zo[FB_MAX_NS];
    for (int sp = 0; sp < ns; ++sp) {
      const long long i = v * ns + sp;
      rin[sp] = A.b[i];
      A.r[i] = rin[sp]; A.x[i] = 0.0; A.p[i] = 0.0; A.s[i] = 0.0;
    }
    if (coarse && rfree[v - v0]) rsum += rin[0];
    bjacobi_apply(ns, A.dinv, v, rin, zo);
    for (int sp = 0; sp < ns; ++sp) {
      uvec[v * ns + sp] = zo[sp];
      if (cached) su[(v - v0) * ns + sp] = zo[sp];
    }
  }
  if (coarse) {  // restriction: block sum of r
    rsum = warp_sum(rsum);
    if (lane == 0) red[0][wid] = rsum;
    __syncthreads();
    if (wid == 0) {
      double sacc = lane < nw ? red[0][lane] : 0.0;
      sacc = warp_sum(sacc);
      if (lane == 0) A.rc[blockIdx.x] = sacc;
    }
  }
  __syncthreads();
  const int s0 = srange[0], s1 = srange[1];
  // with the coarse space the entries other ranks need are complete only after the coarse
  // solve at the top of the loop: they are pushed there
  for (int q = s0 + tid; q < (coarse ? s0 : s1); q += T) {
    unsigned long long* dst = (unsigned long long*)P.base[P.send_rank[q]] + FB_XH + (long long)P.send_dst[q] * ns * 2;
    const long long src = (long long)P.send_v[q] * ns;
    for (int sp = 0; sp < ns; ++sp) ll_store(dst + 2 * sp, P.seq, uvec[src + sp]);
  }
  unsigned gen = 0;
  const unsigned base_count = A.bar_base;
  if (!grid_barrier(A.bar, nb, gen, base_count)) return;
  double alpha = 0.0, beta = 0.0, gamma_old = 0.0, thr2 = A.thr2;
  int it = 0, done = 0, breakdown = 0;
  const int ngroups = T / G, grp = tid / G, gl = tid % G;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
  long long tph[5] = {0, 0, 0, 0, 0};  // cycles per phase (block 0), reported in sc[8..12]
  while (true) {
    const long long c0 = clock64();
    const unsigned utag = P.seq + (unsigned)it;  // version of u this SpMV reads
    int failed = 0;
    // ---- coarse correction u = D^-1 r + Z E^-1 Z^T r, then the boundary entries go out ----
    if (coarse) {
      if (wid == 0) {
        double rv[5];  // loads first, stores after (in-order warp: one round trip, not five)
#pragma unroll
        for (int q = 0; q < 5; ++q) {
          const unsigned c = lane + 32 * q;
          rv[q] = c < nb ? __ldcg(A.rc + c) : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 5; ++q) {
          const unsigned c = lane + 32 * q;
          if (c < nb) rcs[c] = rv[q];
        }
      }
      __syncthreads();
      for (int k = wid; k < nneed; k += nw) {
        double d = 0.0;
        int tb = ctarget;
        if (k == wid) {
#pragma unroll
          for (int q = 0; q < 5; ++q) {
            const unsigned c = lane + 32 * q;
            if (c < nb) d += erow[q] * rcs[c];
          }
        } else {
          tb = A.cn_blk[cn0 + k - 1];
#pragma unroll
          for (int q = 0; q < 5; ++q) {
            const unsigned c = lane + 32 * q;
            if (c < nb) d += A.Einv[(size_t)tb * nb + c] * rcs[c];
          }
        }
        d = warp_sum(d);
        if (lane == 0) ysm[tb] = d;
      }
      __syncthreads();
      const double yself = ysm[blockIdx.x];
      for (int rr = tid; rr < nrows; rr += T)
        if (rfree[rr]) su[rr] += yself;
      __syncthreads();
      for (int q = s0 + tid; q < s1; q += T) {
        unsigned long long* dst = (unsigned long long*)P.base[P.send_rank[q]] + FB_XH + (long long)P.send_dst[q] * 2;
        ll_store(dst, utag, su[P.send_v[q] - v0]);
      }
    }
    // ---- halo of the iterate into shared memory: other blocks' entries from L2, other
    // ranks' entries from the halo words they pushed ------------------------------------------
    if (cached) {
      for (int h = tid; h < nhalo * ns; h += T) {
        const int hv = h / ns, sp = h - hv * ns;
        const long long c = (long long)hidx[hv] * ns + sp;
        double val;
        if (c < n_own_dofs) val = __ldcg(uvec + c) + (coarse ? ysm[hblk[hv]] : 0.0);
        else if (!ll_load(halo + 2 * (c - n_own_dofs), utag, A.bar, val)) failed = 1;
        su[nrows + h] = val;
      }
      __syncthreads();
    }
    // ---- phase B: w = A u on owned rows + partial dots -----------------------------------
    double v3[3] = {0.0, 0.0, 0.0};
    for (int rr = grp; rr < nrows; rr += ngroups) {
      double sum = 0.0;
      const long long row = row0 + rr;
      double ri, ui;
      if (cached) {
        ri = __ldcg(A.r + row);
        ui = su[rr];
        const int a0 = rstart[rr], n = rlen[rr], bs = rbst[rr];
        if (ns == 1) {
          for (int e = gl; e < n; e += G) sum += svals[a0 + e] * su[slcol[bs + e]];
        } else {
          const int vl = rr / ns, sp = rr - vl * ns;
          const int nc = S.pair_nc[sp], pp = S.pair_pp[sp];
          for (int e = gl; e < n; e += G) {
            const int k = e / nc, j = e - k * nc;
            sum += svals[a0 + k * S.npairs + j] * su[(int)slcol[bs + k] * ns + S.pair_cs[pp + j]];
          }
        }
      } else if (staged) {
        ri = __ldcg(A.r + row); ui = __ldcg(uvec + row);
        const int a0 = rstart[rr], n = rlen[rr];
        const int ncr = ns == 1 ? 1 : S.pair_nc[rr - (rr / ns) * ns];
        auto eofs = [&](int e) { return ns == 1 ? e : (e / ncr) * S.npairs + (e - (e / ncr) * ncr); };
        for (int eb = gl; eb < n; eb += 8 * G) {
          double uv[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int e = eb + q * G;
            uv[q] = 0.0;
            if (e < n) {
              const int c = scols[a0 + eofs(e)];
              if (c < n_own_dofs) uv[q] = __ldcg(uvec + c);
              else if (!ll_load(halo + 2 * (c - n_own_dofs), utag, A.bar, uv[q])) failed = 1;
            }
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int e = eb + q * G;
            if (e < n) sum += svals[a0 + eofs(e)] * uv[q];
          }
        }
      } else {
        ri = __ldcg(A.r + row); ui = __ldcg(uvec + row);
        const int vl = rr / ns, sp = rr - vl * ns;
        const long long v = v0 + vl;
        const int r0 = S.rowptr[v], deg = S.rowptr[v + 1] - r0;
        const int nc = S.pair_nc[sp], pp = S.pair_pp[sp];
        const double* __restrict__ vals = A.J + (long long)r0 * S.npairs + pp;   // entry (k, j) at k * npairs + j
        const int n = deg * nc;
        for (int e = gl; e < n; e += G) {
          const int k = e / nc, j = e - k * nc;
          const long long c = (long long)S.colidx[r0 + k] * ns + S.pair_cs[pp + j];
          double uc;
          if (c < n_own_dofs) uc = __ldcg(uvec + c);
          else if (!ll_load(halo + 2 * (c - n_own_dofs), utag, A.bar, uc)) failed = 1;
          sum += vals[(long long)k * S.npairs + j] * uc;
        }
      }
#pragma unroll
      for (int o = G / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(gmask, sum, o);
      if (gl == 0) {
        A.w[row] = sum;
        v3[0] += ri * ui; v3[1] += sum * ui; v3[2] += ui * ui;
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double sred = warp_sum(v3[k]);
      if (lane == 0) red[k][wid] = sred;
    }
    __syncthreads();
    if (wid < 3) {
      double sacc = lane < nw ? red[wid][lane] : 0.0;
      sacc = warp_sum(sacc);
      if (lane == 0) A.partials[(size_t)wid * nb + blockIdx.x] = sacc;
    }
    const long long c1 = clock64();
    if (failed) atomicExch(A.bar + 64, 1u);  // a halo word never arrived: everybody leaves at the rendezvous
    if (!xgrid_reduce(A.bar, nb, gen, base_count, P, A.partials, utag, (unsigned)it & 1u, tot,
                      (unsigned long long*)(A.sc + 16))) return;
    const long long c2 = clock64();
    const double gamma_new = tot[0], delta = tot[1], uu = tot[2];
    if (it == 0 && A.Einv != nullptr) thr2 = A.rel2 * uu;  // relative to ||M^-1 b|| of this M
    if (uu <= thr2 || !(uu == uu)) { done = 1; breakdown = !(uu == uu); }
    else if (it >= A.max_it) { done = 2; }
    else {
      if (it == 0) { beta = 0.0; alpha = gamma_new / delta; }
      else { beta = gamma_new / gamma_old; alpha = gamma_new / (delta - beta * gamma_new / alpha); }
      if (!(alpha == alpha) || isinf(alpha)) { done = 1; breakdown = 1; }
      gamma_old = gamma_new;
    }
    if (done) {
      if (blockIdx.x == 0 && tid == 0) {
        A.flags[0] = (done == 1); A.flags[1] = it; A.flags[2] = breakdown; A.sc[3] = uu;
        for (int k = 0; k < 5; ++k) A.sc[8 + k] = (double)tph[k];
      }
      break;
    }
    ++it;
    const long long c3 = clock64();
    // ---- phase A: vector update on owned vertices, then push the boundary entries ----------
    rsum = 0.0;
    for (long long v = v0 + tid; v < v1; v += T) {
      double rin[FB_MAX_NS], zo[FB_MAX_NS];
      for (int sp = 0; sp < ns; ++sp) {
        const long long i = v * ns + sp;
        const double ui = cached ? su[(v - v0) * ns + sp] : __ldcg(uvec + i);
        const double wi = __ldcg(A.w + i), ri = __ldcg(A.r + i);
        const double p_old = __ldcg(A.p + i), s_old = __ldcg(A.s + i), x_old = __ldcg(A.x + i);
        const double pi = ui + beta * p_old;
        const double si = wi + beta * s_old;
        A.p[i] = pi; A.s[i] = si;
        A.x[i] = x_old + alpha * pi;
        rin[sp] = ri - alpha * si;
        A.r[i] = rin[sp];
      }
      if (coarse && rfree[v - v0]) rsum += rin[0];
      bjacobi_apply(ns, A.dinv, v, rin, zo);
      for (int sp = 0; sp < ns; ++sp) {
        uvec[v * ns + sp] = zo[sp];
        if (cached) su[(v - v0) * ns + sp] = zo[sp];
      }
    }
    if (coarse) {
      rsum = warp_sum(rsum);
      if (lane == 0) red[0][wid] = rsum;
      __syncthreads();
      if (wid == 0) {
        double sacc = lane < nw ? red[0][lane] : 0.0;
        sacc = warp_sum(sacc);
        if (lane == 0) A.rc[blockIdx.x] = sacc;
      }
    } else {
      __syncthreads();
      for (int q = s0 + tid; q < s1; q += T) {
        unsigned long long* dst =
            (unsigned long long*)P.base[P.send_rank[q]] + FB_XH + (long long)P.send_dst[q] * ns * 2;
        const long long src = (long long)P.send_v[q] * ns;
        for (int sp = 0; sp < ns; ++sp) ll_store(dst + 2 * sp, P.seq + (unsigned)it, uvec[src + sp]);
      }
    }
    const long long c4 = clock64();
    if (!grid_barrier(A.bar, nb, gen, base_count)) return;  // local: u of the other blocks; also protects tot[]
    const long long c5 = clock64();
    tph[0] += c1 - c0; tph[1] += c2 - c1; tph[2] += c3 - c2; tph[3] += c4 - c3; tph[4] += c5 - c4;
  }
}

template <int G>
__global__ void __launch_bounds__(FB_TPB)
k_dcg_spmv_dots(DevSpec S, const double* __restrict__ J, const double* __restrict__ u,
                const double* __restrict__ r, double* __restrict__ w, double* partials, unsigned* ticket,
                double* out3) {
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / G;
  const int lane = threadIdx.x % G;
  double v[3] = {0.0, 0.0, 0.0}, tot[3];
  if (row < S.n_owned_dofs) {
    const double sum = spmv_row<G>(S, J, u, row, lane);
    if (lane == 0) {
      w[row] = sum;
      const double ri = r[row], ui = u[row];
      v[0] = ri * ui; v[1] = sum * ui; v[2] = ui * ui;
    }
  }
  if (grid_reduce<3>(v, partials, ticket, tot) && threadIdx.x == 0) {
    out3[0] = tot[0]; out3[1] = tot[1]; out3[2] = tot[2];
  }
}

extern "C" int fb_bjacobi_setup(fb_ctx* ctx, const double* J) {
  if (!ctx || !ctx->have_spec) return ctx ? fb_fail(ctx, "fb_set_spec first") : -1;
  cudaSetDevice(ctx->device);
  int rc = fbi_ensure_workspace(ctx);
  if (rc) return rc;
  const DevSpec& S = ctx->S;
  if (int rc = bjacobi_setup(ctx, J)) return rc;
  FB_CHECK(ctx, cudaGetLastError());
  return 0;
}

extern "C" int fb_precond(fb_ctx* ctx, const double* r, double* z) {
  if (!ctx || !ctx->d_dinv) return ctx ? fb_fail(ctx, "fb_bjacobi_setup first") : -1;
  const DevSpec& S = ctx->S;
  precond_launch(ctx, r, z, (const int*)nullptr);
  FB_CHECK(ctx, cudaGetLastError());
  return 0;
}

extern "C" int fb_dcg_init(fb_ctx* ctx, const double* b, double* x, double* r, double* u, double* p, double* s) {
  if (!ctx || !ctx->d_dinv) return ctx ? fb_fail(ctx, "fb_bjacobi_setup first") : -1;
  const DevSpec& S = ctx->S;
  FB_LAUNCH(ctx, k_cg_init, (unsigned)((S.n_owned + FB_TPB - 1) / FB_TPB), FB_TPB, 0, S, ctx->d_dinv, b, x, r, u, p, s);
  FB_CHECK(ctx, cudaGetLastError());
  return 0;
}

extern "C" int fb_dcg_spmv_dots(fb_ctx* ctx, const double* J, const double* u, const double* r, double* w,
                                double* out3_dev) {
  if (!ctx || !ctx->d_dinv) return ctx ? fb_fail(ctx, "fb_bjacobi_setup first") : -1;
  const DevSpec& S = ctx->S;
  const int G = spmv_group(S);
  const unsigned gs = (unsigned)((S.n_owned_dofs * G + FB_TPB - 1) / FB_TPB);
  if ((long long)gs * 3 > ctx->red_doubles) return fb_fail(ctx, "reduction buffer too small");
  switch (G) {
    case 4: FB_LAUNCH(ctx, k_dcg_spmv_dots<4>, gs, FB_TPB, 0, S, J, u, r, w, ctx->d_red, ctx->d_ticket, out3_dev); break;
    case 8: FB_LAUNCH(ctx, k_dcg_spmv_dots<8>, gs, FB_TPB, 0, S, J, u, r, w, ctx->d_red, ctx->d_ticket, out3_dev); break;
    case 16: FB_LAUNCH(ctx, k_dcg_spmv_dots<16>, gs, FB_TPB, 0, S, J, u, r, w, ctx->d_red, ctx->d_ticket, out3_dev); break;
    default: FB_LAUNCH(ctx, k_dcg_spmv_dots<32>, gs, FB_TPB, 0, S, J, u, r, w, ctx->d_red, ctx->d_ticket, out3_dev); break;
  }
  FB_CHECK(ctx, cudaGetLastError());
  return 0;
}

extern "C" int fb_dcg_update(fb_ctx* ctx, double alpha, double beta, double* x, double* r, double* u,
                             const double* w, double* p, double* s) {
  if (!ctx || !ctx->d_dinv) return ctx ? fb_fail(ctx, "fb_bjacobi_setup first") : -1;
  const DevSpec& S = ctx->S;
  FB_LAUNCH(ctx, k_dcg_update, (unsigned)((S.n_owned + FB_TPB - 1) / FB_TPB), FB_TPB, 0, S, ctx->d_dinv, alpha, beta, x, r, u, w, p, s);
  FB_CHECK(ctx, cudaGetLastError());
  return 0;
}

// Device-resident recurrence scalars for the partitioned CG, so an iteration
// (halo exchange, SpMV + dots, all-reduce, scalars, update) has no host round trip and
// can be captured once in a CUDA graph and replayed.
//   state: [0]=alpha [1]=beta [2]=gamma_old [3]=iterations [4]=done [5]=threshold^2 [6]=uu
__global__ void k_dcg_scalars(const double* __restrict__ red3, double* __restrict__ st, int max_it) {
  if (st[4] != 0.0) return;
  const double gamma = red3[0], delta = red3[1], uu = red3[2];
  st[6] = uu;
  const int it = (int)st[3];
  if (uu <= st[5] || !(uu == uu) || it >= max_it) { st[4] = (uu <= st[5]) ? 1.0 : 2.0; return; }
  double alpha, beta;
  if (it == 0) { beta = 0.0; alpha = gamma / delta; }
  else { beta = gamma / st[2]; alpha = gamma / (delta - beta * gamma / st[0]); }
  st[0] = alpha; st[1] = beta; st[2] = gamma; st[3] = (double)(it + 1);
}

__global__ void __launch_bounds__(FB_TPB)
k_dcg_update_dev(DevSpec S, const double* __restrict__ dinv, const double* __restrict__ st,
                 double* __restrict__ x, double* __restrict__ r, double* __restrict__ u,
                 const double* __restrict__ w, double* __restrict__ p, double* __restrict__ s_) {
  if (st[4] != 0.0) return;
  const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (v >= S.n_owned) return;
  const double alpha = st[0], beta = st[1];
  double rin[FB_MAX_NS], zo[FB_MAX_NS];
  for (int s = 0; s < S.ns; ++s) {
    const long long i = v * S.ns + s;
    const double pi = u[i] + beta * p[i];
    const double si = w[i] + beta * s_[i];
    p[i] = pi; s_[i] = si;
    x[i] += alpha * pi;
    rin[s] = r[i] - alpha * si;
    r[i] = rin[s];
  }
  bjacobi_apply(S.ns, dinv, v, rin, zo);
  for (int s = 0; s < S.ns; ++s) u[v * S.ns + s] = zo[s];
}

extern "C" int fb_dcg_scalars(fb_ctx* ctx, const double* red3_dev, double* state_dev, int max_it) {
  if (!ctx) return -1;
  FB_LAUNCH(ctx, k_dcg_scalars, 1, 1, 0, red3_dev, state_dev, max_it);
  return 0;
}

extern "C" int fb_dcg_update_dev(fb_ctx* ctx, const double* state_dev, double* x, double* r, double* u,
                                 const double* w, double* p, double* s) {
  if (!ctx || !ctx->d_dinv) return ctx ? fb_fail(ctx, "fb_bjacobi_setup first") : -1;
  const DevSpec& S = ctx->S;
  FB_LAUNCH(ctx, k_dcg_update_dev, (unsigned)((S.n_owned + FB_TPB - 1) / FB_TPB), FB_TPB, 0, S, ctx->d_dinv,
            state_dev, x, r, u, w, p, s);
  return 0;
}

// ---------------------------------------------------------------------------
// Peer-memory setup and launch of the partitioned persistent CG (k_dcg_persistent).
// ---------------------------------------------------------------------------
extern "C" int fb_peer_alloc(fb_ctx* ctx, int rank, int nranks, unsigned char* handle64) {
  if (!ctx || !ctx->have_spec) return ctx ? fb_fail(ctx, "fb_set_spec first") : -1;
  if (nranks < 1 || nranks > 8 || rank < 0 || rank >= nranks) return fb_fail(ctx, "1..8 ranks on one node");
  cudaSetDevice(ctx->device);
  if (!ctx->d_xchg) {
    ctx->halo_words = 2 * (ctx->S.n_dofs - ctx->S.n_owned * (long long)ctx->S.ns);   // two words per ghost dof
    ctx->xchg_u_doubles = 2 * ctx->halo_words;                                       // two parities
    const size_t n = (size_t)FB_XH + (size_t)ctx->xchg_u_doubles + 2;
    int rc = fb_alloc(ctx, &ctx->d_xchg, n);  // its own cudaMalloc block: exportable as a whole
    if (rc) return rc;
    FB_CHECK(ctx, cudaMemset(ctx->d_xchg, 0, n * sizeof(double)));
  }
  ctx->peer_rank = rank;
  ctx->peer_n = nranks;
  for (int q = 0; q < 8; ++q) ctx->peer_base[q] = nullptr;
  ctx->peer_base[rank] = ctx->d_xchg;
  if (handle64) {
    cudaIpcMemHandle_t h;
    FB_CHECK(ctx, cudaIpcGetMemHandle(&h, ctx->d_xchg));
    static_assert(sizeof(h) == 64, "IPC handle size");
    memcpy(handle64, &h, 64);
  }
  return 0;
}

extern "C" int fb_peer_open(fb_ctx* ctx, int peer, const unsigned char* handle64) {
  if (!ctx || !ctx->d_xchg) return ctx ? fb_fail(ctx, "fb_peer_alloc first") : -1;
  if (peer < 0 || peer >= ctx->peer_n) return fb_fail(ctx, "peer rank out of range");
  if (peer == ctx->peer_rank) return 0;
  cudaSetDevice(ctx->device);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  FB_CHECK(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  ctx->peer_base[peer] = p;
  return 0;
}

// send table: owned vertex v (local numbering) is ghost vertex dst of rank `rank`; sorted by v
extern "C" int fb_peer_set_send(fb_ctx* ctx, int64_t n, const int32_t* v, const int32_t* rank, const int32_t* dst) {
  if (!ctx || !ctx->d_xchg) return ctx ? fb_fail(ctx, "fb_peer_alloc first") : -1;
  cudaSetDevice(ctx->device);
  for (int64_t i = 0; i < n; ++i) {
    if (v[i] < 0 || v[i] >= ctx->S.n_owned) return fb_fail(ctx, "send vertex is not owned");
    if (rank[i] < 0 || rank[i] >= ctx->peer_n || rank[i] == ctx->peer_rank) return fb_fail(ctx, "bad destination rank");
    if (i && v[i] < v[i - 1]) return fb_fail(ctx, "send table must be sorted by vertex");
  }
  int rc;
  if ((rc = fb_upload(ctx, &ctx->d_send_v, (const int*)v, (size_t)n))) return rc;
  if ((rc = fb_upload(ctx, &ctx->d_send_rank, (const int*)rank, (size_t)n))) return rc;
  if ((rc = fb_upload(ctx, &ctx->d_send_dst, (const int*)dst, (size_t)n))) return rc;
  ctx->n_send = (int)n;
  ctx->peer_ready = ctx->peer_n > 1;   // every peer area is mapped (fb_peer_open) before the send table is set
  return 0;
}

extern "C" int fb_dcg_persistent(fb_ctx* ctx, const double* J, const double* b, double* x, double thr2,
                                 double rtol, int max_it, int* iters, int* converged) {
  if (!ctx || !ctx->d_dinv) return ctx ? fb_fail(ctx, "fb_bjacobi_setup first") : -1;
  if (!ctx->d_xchg) return fb_fail(ctx, "fb_peer_alloc first");
  for (int q = 0; q < ctx->peer_n; ++q)
    if (!ctx->peer_base[q]) return fb_fail(ctx, "fb_peer_open for every rank first");
  cudaSetDevice(ctx->device);
  const DevSpec& S = ctx->S;
  int rc = fbi_ensure_workspace(ctx);
  if (rc) return rc;
  if ((rc = ensure_vectors(ctx, 5))) return rc;
  const long long N = S.n_dofs;
  CgArgs A;
  memset(&A, 0, sizeof(A));
  A.J = J; A.dinv = ctx->d_dinv; A.b = b; A.x = x;
  A.r = ctx->d_work; A.u = A.r + N; A.w = A.u + N; A.p = A.w + N; A.s = A.p + N;
  A.partials = ctx->d_red; A.bar = ctx->d_ticket + 64; A.sc = ctx->d_sc; A.flags = ctx->d_flags;
  A.thr2 = thr2; A.max_it = max_it;
  PeerArgs P;
  memset(&P, 0, sizeof(P));
  P.rank = ctx->peer_rank; P.nranks = ctx->peer_n; P.seq = ctx->peer_seq;
  for (int q = 0; q < 8; ++q) P.base[q] = (double*)ctx->peer_base[q];
  P.send_v = ctx->d_send_v; P.send_rank = ctx->d_send_rank; P.send_dst = ctx->d_send_dst; P.n_send = ctx->n_send;
  const long long n_own_dofs = S.n_owned * (long long)S.ns;
  const double rowlen = (double)ctx->nnzb * S.npairs / (double)N;
  int G = rowlen <= 20 ? 2 : rowlen <= 40 ? 4 : rowlen <= 80 ? 8 : rowlen <= 160 ? 16 : 32;
  if (getenv("FB_CG_G")) G = atoi(getenv("FB_CG_G"));
  const int T = 1024;
  const unsigned nb = (unsigned)ctx->sm_count;
  int v = 0;
  cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device);
  const size_t smem_max = (size_t)v - 4096;  // static shared: reductions, coarse solution
  const long long vper = (S.n_owned + nb - 1) / nb;
  const long long rows_per_block = vper * S.ns;
  size_t smem = 0;
  const bool no_smem = getenv("FB_NO_SMEM_OPERATOR") != nullptr;
  if (ctx->part_blocks == (int)nb && ctx->part_owned_only && !getenv("FB_NO_SM_PARTITION")) {
    // compact per-SM parts of the owned vertices: operator rows + iterate cache on chip
    A.blk_rng = ctx->d_blk_rng; A.halo_ptr = ctx->d_halo_ptr; A.halo_idx = ctx->d_halo_idx; A.lcol = ctx->d_lcol;
    const long long capp = (ctx->part_max_entries + 1) & ~1LL;
    const long long su = (ctx->part_max_su + 1) & ~1LL;
    const long long be = (ctx->part_max_bentries + 1) & ~1LL;
    const long long rows = ctx->part_max_rows;
    const size_t need = (size_t)(8 * capp + 8 * su + 2 * be + 13 * rows + 32);
    if (need <= smem_max && !no_smem) {
      A.rows_in_smem = 1; A.smem_entries = capp; A.su_doubles = su; A.bentries = be; A.rows_cap = rows;
      smem = need;
      if ((rc = coarse_setup(ctx, J))) return rc;
      if (ctx->coarse_ready) {  // rank-local two-level preconditioner
        A.Einv = ctx->d_coarse + (size_t)nb * nb; A.rc = ctx->d_coarse + (size_t)2 * nb * nb;
        A.halo_blk = ctx->d_halo_blk_eff; A.cn_ptr = ctx->d_cn_ptr; A.cn_blk = ctx->d_cn_blk;
        A.rel2 = rtol * rtol;
      }
    }
  } else {
    // contiguous equal ranges: operator rows in shared memory when they fit, gathers from L2
    const bool rows_fit = 12 * (rows_per_block + 2) <= (long long)smem_max / 2;
    long long cap = rows_fit ? ((long long)smem_max - 12 * (rows_per_block + 2)) / 12 : 0;
    cap &= ~1LL;
    const long long avg_entries = (long long)((double)ctx->nnzb * S.npairs * ((double)n_own_dofs / (double)N) / nb) + 1;
    A.rows_in_smem = rows_fit && cap >= avg_entries && !no_smem;
    A.smem_entries = A.rows_in_smem ? cap : 0;
    smem = A.rows_in_smem ? (size_t)(12 * A.smem_entries + 12 * (rows_per_block + 2)) : 0;
  }
  FB_CHECK(ctx, cudaMemsetAsync(ctx->d_flags, 0, 32, ctx->stream));
  FB_CHECK(ctx, cudaMemsetAsync(A.bar, 0, 130 * sizeof(unsigned), ctx->stream));
  FB_CHECK(ctx, cudaMemsetAsync(ctx->d_sc + 16, 0, 2 * sizeof(double), ctx->stream));
  A.bar_base = 0u;
  void* args[] = {(void*)&S, (void*)&A, (void*)&P};
  const void* fn = nullptr;
  switch (G) {
    case 2: fn = (const void*)k_dcg_persistent<2>; break;
    case 4: fn = (const void*)k_dcg_persistent<4>; break;
    case 8: fn = (const void*)k_dcg_persistent<8>; break;
    case 16: fn = (const void*)k_dcg_persistent<16>; break;
    default: fn = (const void*)k_dcg_persistent<32>; break;
  }
  FB_CHECK(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  FB_CHECK(ctx, cudaLaunchCooperativeKernel(fn, dim3(nb), dim3(T), args, smem, ctx->stream));
  ctx->launches++;
  int fl[4];
  if ((rc = read_flags(ctx, fl))) return rc;
  unsigned aborted = 0;
  FB_CHECK(ctx, cudaMemcpy(&aborted, A.bar + 64, 4, cudaMemcpyDeviceToHost));
  if (aborted) return fb_fail(ctx, "partitioned persistent CG: rendezvous timed out (a rank did not arrive)");
  *iters = fl[1];
  *converged = fl[0] && !fl[2];
  ctx->peer_seq += (unsigned)fl[1] + 2u;  // same on every rank: the iteration count is
  if (getenv("FB_VERBOSE") && fl[1] > 0 && ctx->peer_rank == 0) {
    double ph[5];
    unsigned long long dbg[2];
    FB_CHECK(ctx, cudaMemcpy(ph, ctx->d_sc + 8, sizeof(ph), cudaMemcpyDeviceToHost));
    FB_CHECK(ctx, cudaMemcpy(dbg, ctx->d_sc + 16, sizeof(dbg), cudaMemcpyDeviceToHost));
    fprintf(stderr, "[fb] peer CG last block per rendezvous: sum of partials %.0f cycles, exchange (store + wait for "
            "all ranks) %.0f cycles\n", (double)dbg[0] / fl[1], (double)dbg[1] / fl[1]);
    fprintf(stderr, "[fb] peer CG ranks=%d G=%d smem=%d parts=%d coarse=%d its=%d: cycles/iteration spmv+dots %.0f, rendezvous %.0f, "
            "scalars %.0f, update+push %.0f, local barrier %.0f\n", ctx->peer_n, G, A.rows_in_smem, A.blk_rng != nullptr, A.Einv != nullptr, fl[1], ph[0] / fl[1],
            ph[1] / fl[1], ph[2] / fl[1], ph[3] / fl[1], ph[4] / fl[1]);
  }
  return 0;
}

extern "C" int fb_chebyshev_step(fb_ctx* ctx, const double* d_in, double* r, double* z, double* d_out, double ca, double cb) {
  if (!ctx || !ctx->have_spec) return ctx ? fb_fail(ctx, "fb_set_spec first") : -1;
  cudaSetDevice(ctx->device);
  if (!ctx->d_dinv || !ctx->dinv32_valid) return fb_fail(ctx, "fb_chebyshev_step: fb_bjacobi_setup(ctx, NULL) on the channel operator first");
  if (!spmv_ch_tma_ok(ctx, d_in)) return fb_fail(ctx, "fb_chebyshev_step: the TMA channel SpMV does not apply to this problem");
  SpmvChArgs a{};
  a.x = d_in; a.dinv32 = ctx->d_dinv32; a.dscale = ctx->d_dscale; a.flags = nullptr;
  a.r = r; a.z = z; a.dout = d_out; a.vin = nullptr; a.wout = nullptr;
  a.ca = ca; a.cb = cb; a.last = 0;
  if (int rc = fbi_halo(ctx, (double*)d_in)) return rc;
  return spmv_ch_tma_launch<2>(ctx, a);
}
