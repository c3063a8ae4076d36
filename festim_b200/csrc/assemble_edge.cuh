// Channel assembly of residual and Jacobian for diffusion-trapping-reaction systems (fp64, P1
// simplices).  Replaces the FFCx tabulate_tensor kernels + dolfinx assemble_vector /
// assemble_matrix / apply_lifting / set_bc behind NonlinearProblem (reference
// src/festim/problem.py:170-182) for the form of src/festim/hydrogen_transport_problem.py:870-997
// whenever every solution-dependent volume term is a product of P1 functions with a cell-wise
// coefficient (diffusion, mass, advection, mass-action reactions of reaction.py:179-224 with
// affine factors, species.py:167-176).  festim_b200/lowering.py (_vertex_tables) describes the
// algebra; tests/vtx_emulator.py is its numpy model.
//
// The Jacobian of such a system is a short sum of Kronecker products
//     J = sum_c  A^c (x) B_c ,      A^c an N_v x N_v matrix on the vertex graph ("channel"),
//                                   B_c a constant ns x ns table of coefficients,
// and the residual is F_s(v) = load + sum_e coef_e sum_w A^{c_e}_vw g_e(w) with g a column of the
// per-vertex table [u | u - u_n | affine reaction factors | 1].  Channels are either
//   static   (mass, Arrhenius diffusion, first-order rate moments, advection): functions of the
//            mesh and the temperature, assembled ONCE per temperature update (k_et_setup), or
//   dynamic  (second-order rate moments contracted with an affine factor f of the species):
//            A^c_vw = sum_z T_vwz f(z)  with the edge tensor T_vwz = sum_{K > v,w,z} |K| sum_q w_q
//            rho(T_q) phi_v phi_w phi_z, also assembled once per temperature update.
// A Newton iteration therefore never touches a cell: k_assemble_edge streams the edge tensor and
// the static channel values of a block row, forms the dynamic channel values (written to the
// channel operator, 8 bytes per channel and edge instead of 8 bytes per species pair and edge)
// and the residual, applying the strong Dirichlet rows / lifting on the way (dolfinx
// NonlinearProblem semantics).  Nothing is read-modify-written in global memory,
// there are no atomics, and the summation order is fixed (bitwise reproducible).
// k_et_expand turns the channel operator into the species-sparse block-CSR values of the C ABI.
#pragma once
#include "fb_device.cuh"

enum { FB_CQ_KD = 0, FB_CQ_M1 = 1, FB_CQ_M3 = 2, FB_CQ_ADV = 3 };
enum { FB_CH_MASS = 0, FB_CH_PAIR = 1, FB_CH_TRI = 2, FB_CH_FULL = 3 };
#define FB_ET_WREG 8   // triples per edge kept in registers

// position of (a, b) / (a, b, c) in the packed symmetric tables (a <= b <= c after sorting)
template <int D1>
__device__ __forceinline__ int fb_sym2(int a, int b) {
  if (a > b) { const int t = a; a = b; b = t; }
  return a * D1 - (a * (a - 1)) / 2 + (b - a);
}
template <int D1>
__device__ __forceinline__ int fb_sym3(int a, int b, int c) {
  if (a > b) { const int t = a; a = b; b = t; }
  if (b > c) { const int t = b; b = c; c = t; }
  if (a > b) { const int t = a; a = b; b = t; }
  int idx = 0;
#pragma unroll
  for (int x = 0; x < D1; ++x)
    if (x < a) idx += ((D1 - x) * (D1 - x + 1)) / 2;   // triples starting with x
#pragma unroll
  for (int y = 0; y < D1; ++y)
    if (y >= a && y < b) idx += D1 - y;                // (a, y, *)
  return idx + (c - b);
}

// n / d through the magic number ceil(2^32 / d) (exact while n * d < 2^32); magic 0 means d == 1
__device__ __forceinline__ int fb_fastdiv(unsigned n, unsigned magic) {
  return magic ? (int)__umulhi(n, magic) : (int)n;
}

// ---------------------------------------------------------------------------
// per-cell record (per temperature / velocity update)
// ---------------------------------------------------------------------------
template <int DIM>
__device__ __forceinline__ void fb_cell_records_body(const DevSpec& S, double* __restrict__ out) {
  constexpr int D1 = DIM + 1;
  constexpr int NSYM = D1 * (D1 + 1) / 2, NSYM3 = D1 * (D1 + 1) * (D1 + 2) / 6;
  const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (c >= S.n_cells) return;
  int a[FB_MAX_D1];
  double X[FB_MAX_D1][3], G[FB_MAX_D1][3];
  fb_load_cell<DIM>(S, c, a, X);
  const double vol = fb_geometry<DIM>(X, G);
  const int tag = S.cell_tag[c];
  double* __restrict__ rec = out + c * (long long)S.vx_rec;
  rec[0] = vol;
  double Tn[FB_MAX_D1];
#pragma unroll
  for (int j = 0; j < D1; ++j) Tn[j] = S.T_mode ? S.T_nodal[a[j]] : S.T_const;
  for (int q = S.vx_cq_ptr[tag]; q < S.vx_cq_ptr[tag + 1]; ++q) {
    const int kind = S.vx_cq[4 * q], idx = S.vx_cq[4 * q + 1], rule = S.vx_cq[4 * q + 2], off = S.vx_cq[4 * q + 3];
    if (kind == FB_CQ_KD) {
      const double kd = vol * S.dbar[((long long)rule * S.n_slots + idx) * S.n_cells + c];
      int k = 0;
#pragma unroll
      for (int p = 0; p < D1; ++p)
#pragma unroll
        for (int r = p; r < D1; ++r) {
          double d = 0.0;
#pragma unroll
          for (int x = 0; x < DIM; ++x) d += G[p][x] * G[r][x];
          rec[off + k++] = kd * d;
        }
    } else if (kind == FB_CQ_M1 || kind == FB_CQ_M3) {
      const int qoff = S.quad_v[tag * 4 + 2 * rule], nq = S.quad_v[tag * 4 + 2 * rule + 1];
      const double E = S.rslot_E[idx];
      double acc[NSYM3];
#pragma unroll
      for (int k = 0; k < NSYM3; ++k) acc[k] = 0.0;
      for (int qi = 0; qi < nq; ++qi) {
        const double* __restrict__ bary = S.quad_bary + (long long)(qoff + qi) * D1;
        double b[FB_MAX_D1], T = 0.0;
#pragma unroll
        for (int j = 0; j < D1; ++j) { b[j] = bary[j]; T += b[j] * Tn[j]; }
        const double w = S.quad_w[qoff + qi] * exp(-E / (FB_KB * T));
        int k = 0;
        if (kind == FB_CQ_M1) {
#pragma unroll
          for (int p = 0; p < D1; ++p)
#pragma unroll
            for (int r = p; r < D1; ++r) acc[k++] += w * b[p] * b[r];
        } else {
#pragma unroll
          for (int p = 0; p < D1; ++p)
#pragma unroll
            for (int r = p; r < D1; ++r)
#pragma unroll
              for (int t = r; t < D1; ++t) acc[k++] += w * b[p] * b[r] * b[t];
        }
      }
      const int n = kind == FB_CQ_M1 ? NSYM : NSYM3;
#pragma unroll
      for (int k = 0; k < NSYM3; ++k)
        if (k < n) rec[off + k] = vol * acc[k];
    } else {  // advection: |K|/((d+1)(d+2)) sum_b (vel_b . G_j) (1 + delta_bi)
      const double* __restrict__ vel = S.velocities[idx];
      const double sc = vol / (double)(D1 * (D1 + 1));
      double vg[FB_MAX_D1][FB_MAX_D1];   // [b][j]
#pragma unroll
      for (int b = 0; b < D1; ++b)
#pragma unroll
        for (int j = 0; j < D1; ++j) {
          double d = 0.0;
#pragma unroll
          for (int x = 0; x < DIM; ++x) d += vel[(long long)a[b] * DIM + x] * G[j][x];
          vg[b][j] = d;
        }
#pragma unroll
      for (int i = 0; i < D1; ++i)
#pragma unroll
        for (int j = 0; j < D1; ++j) {
          double s = 0.0;
#pragma unroll
          for (int b = 0; b < D1; ++b) s += vg[b][j] * (b == i ? 2.0 : 1.0);
          rec[off + i * D1 + j] = s * sc;
        }
    }
  }
}

// ---------------------------------------------------------------------------
// Dirichlet information per vertex (once per fb_set_dirichlet)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void fb_vinfo_rows_body(const DevSpec& S, unsigned* __restrict__ vinfo) {
  const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (v >= S.n_vertices) return;
  unsigned m = 0;
  for (int s = 0; s < S.ns; ++s)
    if (S.bc_map[v * S.ns + s] >= 0) m |= 1u << s;
  vinfo[v] = m;
}
__device__ __forceinline__ void fb_vinfo_cols_body(const DevSpec& S, unsigned* __restrict__ vinfo) {
  const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (v >= S.n_vertices) return;
  unsigned any = 0;
  for (int e = S.rowptr[v]; e < S.rowptr[v + 1]; ++e) any |= vinfo[S.colidx[e]] & 0xffffu;
  if (any) atomicOr(&vinfo[v], 1u << 16);
}

// ---------------------------------------------------------------------------
// Structure of the edge tensor (once per mesh): for every edge (v, w_k) of an owned row the
// row slots of the vertices z that share a cell with both, ELL per row:
// tzslot[tzptr[v] + j * deg + k], j < W_v = the largest such count in the row.
// mode 0: width[v] = W_v;  mode 1: fill the slots (array preset to 0xff = padding).
// ---------------------------------------------------------------------------
template <int DIM>
__device__ __forceinline__ void fb_et_symbolic_body(const DevSpec& S, int mode, int* __restrict__ width,
                                                    const long long* __restrict__ tzptr,
                                                    unsigned char* __restrict__ tzslot, int* __restrict__ err) {
  constexpr int D1 = DIM + 1;
  const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (v >= S.n_owned) return;
  const int r0 = S.rowptr[v], deg = S.rowptr[v + 1] - r0;
  const int e0 = S.v2c_ptr[v], e1 = S.v2c_ptr[v + 1];
  int W = 0;
  for (int k = 0; k < deg; ++k) {
    unsigned char lst[64];
    int cnt = 0;
    for (int e = e0; e < e1; ++e) {
      const int packed = S.v2c_idx[e];
      const int i = packed - (packed / D1) * D1;
      const unsigned sw = S.v2c_slot[e];
      int jf = -1;
#pragma unroll
      for (int j = 0; j < D1; ++j)
        if (j != i && (int)((sw >> (8 * j)) & 0xffu) == k) jf = j;
      if (jf < 0) continue;
#pragma unroll
      for (int b = 0; b < D1; ++b) {
        if (b == i || b == jf) continue;
        const unsigned char zs = (unsigned char)((sw >> (8 * b)) & 0xffu);
        bool seen = false;
        for (int q = 0; q < cnt; ++q) seen |= (lst[q] == zs);
        if (!seen) {
          if (cnt < 64) lst[cnt++] = zs; else *err = 1;
        }
      }
    }
    if (cnt > W) W = cnt;
    if (mode == 1) {
      const long long base = tzptr[v];
      for (int q = 0; q < cnt; ++q) tzslot[base + (long long)q * deg + k] = lst[q];
    }
  }
  if (mode == 0) width[v] = W;
}

// ---------------------------------------------------------------------------
// Static channel values and edge tensors of the owned rows from the cell records (per
// temperature / velocity update).  One thread per vertex walks its cells in adjacency order:
// every sum has a fixed order.
// ---------------------------------------------------------------------------
template <int DIM>
__device__ __forceinline__ void fb_et_setup_body(const DevSpec& S, double* __restrict__ stat, double* __restrict__ tvw,
                                                 double* __restrict__ tz) {
  constexpr int D1 = DIM + 1;
  const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (v >= S.n_owned) return;
  const int r0 = S.rowptr[v], deg = S.rowptr[v + 1] - r0;
  const int NSP = S.vx_nsp, nstat = S.vx_nstat, nts = S.vx_nts;
  const long long tz0 = nts ? S.et_tzptr[v] : 0;
  const int W = nts ? S.et_tzw[v] : 0;
  const int slab = nts ? (int)(S.et_tzptr[v + 1] - tz0) : 0;
  for (int t = 0; t < deg * NSP; ++t) stat[(long long)r0 * NSP + t] = 0.0;
  for (int ts = 0; ts < nts; ++ts) {
    for (int k = 0; k < 2 * deg; ++k) tvw[2 * (ts * S.nnzb + r0) + k] = 0.0;
    for (int t = 0; t < slab; ++t) tz[ts * S.et_tz_total + tz0 + t] = 0.0;
  }
  const int* __restrict__ ch = S.vx_itab + S.vx_o_ch;
  const double mscale = 1.0 / (double)(D1 * (D1 + 1));
  for (int e = S.v2c_ptr[v]; e < S.v2c_ptr[v + 1]; ++e) {
    const int packed = S.v2c_idx[e];
    const long long cell = packed / D1;
    const int i = packed - (int)cell * D1;
    const unsigned sw = S.v2c_slot[e];
    const int tag = S.cell_tag[cell];
    const double* __restrict__ rec = S.cellrec + cell * S.vx_rec;
    int slot[D1];
#pragma unroll
    for (int b = 0; b < D1; ++b) slot[b] = (int)((sw >> (8 * b)) & 0xffu);
    for (int c = 0; c < nstat; ++c) {
      if (ch[8 * c] != tag) continue;
      const int ctype = ch[8 * c + 1], off = ch[8 * c + 2];
#pragma unroll
      for (int j = 0; j < D1; ++j) {
        double val;
        if (ctype == FB_CH_MASS) val = rec[0] * mscale * (i == j ? 2.0 : 1.0);
        else if (ctype == FB_CH_PAIR) val = rec[off + fb_sym2<D1>(i, j)];
        else val = rec[off + i * D1 + j];
        stat[(long long)(r0 + slot[j]) * NSP + c] += val;
      }
    }
    for (int c = nstat; c < S.vx_nch; ++c) {
      // the first dynamic channel of every tensor fills it
      const int ts = ch[8 * c + 5];
      if (ch[8 * c] != tag) continue;
      bool first = true;
      for (int c2 = nstat; c2 < c; ++c2) first &= (ch[8 * c2 + 5] != ts);
      if (!first) continue;
      const int off = ch[8 * c + 2];
      double* __restrict__ Tvw = tvw + 2 * (ts * S.nnzb + r0);
      double* __restrict__ Tz = tz + ts * S.et_tz_total + tz0;
      const unsigned char* __restrict__ zs = S.et_tzslot + tz0;
#pragma unroll
      for (int j = 0; j < D1; ++j) {
        const int k = slot[j];
        if (j == i) { Tvw[2 * k] += rec[off + fb_sym3<D1>(i, i, i)]; continue; }
        Tvw[2 * k] += rec[off + fb_sym3<D1>(i, i, j)];
        Tvw[2 * k + 1] += rec[off + fb_sym3<D1>(i, j, j)];
#pragma unroll
        for (int b = 0; b < D1; ++b) {
          if (b == i || b == j) continue;
          const unsigned char want = (unsigned char)slot[b];
          for (int q = 0; q < W; ++q)
            if (zs[(long long)q * deg + k] == want) { Tz[(long long)q * deg + k] += rec[off + fb_sym3<D1>(i, j, b)]; break; }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// shared-memory layout of k_assemble_edge / k_et_expand (doubles); computed on the host too
// ---------------------------------------------------------------------------
struct EtLayout {
  int grp0;        // start (in doubles) of the per-group areas, after the block-shared tables
  int per_grp;     // doubles per group (one group of G lanes = one vertex)
  int nbwp, nchp;  // padded (odd) strides of the neighbour table and of the channel rows
  int o_nb, o_A, o_res, o_facc, o_ints;
};

__host__ __device__ inline EtLayout fb_et_layout(const DevSpec& S, bool with_nb) {
  EtLayout L;
  const int md = S.maxdeg;
  const int dt = (S.vx_dtab_n + 1) & ~1;
  const int it_doubles = (S.vx_itab_n + 1) / 2;
  L.grp0 = (dt + it_doubles + 1) & ~1;
  L.nbwp = S.vx_nbw | 1;
  L.nchp = S.vx_nch | 1;
  int o = 0;
  L.o_nb = o; o += with_nb ? md * L.nbwp : 0;
  L.o_A = o; o += md * L.nchp;
  const int nres = S.vx_nft > S.vx_njt ? S.vx_nft : S.vx_njt;
  L.o_res = o; o += with_nb ? (nres > 0 ? nres : 1) : 0;
  L.o_facc = o; o += with_nb ? FB_MAX_NS : 0;
  o = (o + 1) & ~1;
  L.o_ints = o;      // ints: cols[md], nbinfo[md]
  o += md + 1;
  L.per_grp = (o + 1) & ~1;
  return L;
}

template <int G>
__device__ __forceinline__ unsigned fb_group_mask() {
  return (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) & ~(G - 1)));
}
template <int G>
__device__ __forceinline__ double fb_group_sum(double v, unsigned gmask) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
  return v;
}
template <int G>
__device__ __forceinline__ int fb_group_max(int v, unsigned gmask) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) { const int t = __shfl_xor_sync(gmask, v, o); v = t > v ? t : v; }
  return v;
}

// ---------------------------------------------------------------------------
// the Newton-iteration kernel: G lanes per owned vertex (= one block row)
// ---------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ void fb_assemble_edge_body(const DevSpec& S, const double* __restrict__ u,
                                                      const double* __restrict__ un, double* __restrict__ F,
                                                      double* __restrict__ dyn, int flags, double* smem) {
  const EtLayout L = fb_et_layout(S, true);
  const int g = threadIdx.x & (G - 1), grp = threadIdx.x / G, gpb = blockDim.x / G;
  // ---- block-shared tables -------------------------------------------------------------
  double* dtab = smem;
  int* itab = (int*)(smem + ((S.vx_dtab_n + 1) & ~1));
  for (int t = threadIdx.x; t < S.vx_dtab_n; t += blockDim.x) dtab[t] = S.vx_dtab[t];
  for (int t = threadIdx.x; t < S.vx_itab_n; t += blockDim.x) itab[t] = S.vx_itab[t];
  __syncthreads();
  const long long v = (long long)blockIdx.x * gpb + grp;
  if (v >= S.n_owned) return;
  const unsigned gmask = fb_group_mask<G>();
  const int ns = S.ns, NBW = L.nbwp, NCHP = L.nchp, NF = S.n_factors;
  const int nstat = S.vx_nstat, ndyn = S.vx_ndyn, NSP = S.vx_nsp, NDP = S.vx_ndp;
  const bool doF = (flags & FB_ASSEMBLE_F) != 0;
  double* W_ = smem + L.grp0 + (size_t)grp * L.per_grp;
  double* nb = W_ + L.o_nb;
  double* A = W_ + L.o_A;
  double* res = W_ + L.o_res;
  double* facc = W_ + L.o_facc;
  int* cols = (int*)(W_ + L.o_ints);
  unsigned* nbinfo = (unsigned*)(cols + S.maxdeg);

  const int r0 = S.rowptr[v], deg = S.rowptr[v + 1] - r0;
  const unsigned vi = S.vinfo[v];
  const unsigned rowbc = vi & 0xffffu;
  const bool colbc = (vi >> 16) & 1u;

  // ---- phase 0: neighbour table [u | u - u_n | factors | 1] ------------------------------
  int kself = 0;
  for (int k = g; k < deg; k += G) {
    const int w = S.colidx[r0 + k];
    cols[k] = w;
    nbinfo[k] = colbc ? S.vinfo[w] : 0u;
    if (w == (int)v) kself = k;
  }
  kself = fb_group_max<G>(kself, gmask);
  __syncwarp(gmask);
  for (int t = g; t < deg * ns; t += G) {
    const int k = fb_fastdiv((unsigned)t, S.vx_m_ns), s = t - k * ns;
    const long long dof = (long long)cols[k] * ns + s;
    const double uu = u[dof];
    nb[k * NBW + s] = uu;
    nb[k * NBW + ns + s] = S.transient ? uu - un[dof] : 0.0;
  }
  for (int k = g; k < deg; k += G) nb[k * NBW + 2 * ns + NF] = 1.0;
  __syncwarp(gmask);
  for (int t = g; t < deg * NF; t += G) {
    const int k = fb_fastdiv((unsigned)t, S.vx_m_nf), f = t - k * NF;
    const int* h = S.fac_hdr + 4 * f;
    double val = h[0] ? S.scalars[h[1]] : S.fac_a0[f];
    for (int q = 0; q < h[3]; ++q) val += S.fac_coef[h[2] + q] * nb[k * NBW + S.fac_species[h[2] + q]];
    nb[k * NBW + 2 * ns + f] = val;
  }
  // ---- static channel values of the row -------------------------------------------------
  if (doF && nstat > 0) {
    const double* __restrict__ src = S.et_stat + (long long)r0 * NSP;
    for (int t = g; t < deg * NSP; t += G) {
      const int k = fb_fastdiv((unsigned)t, S.vx_m_nsp), c = t - k * NSP;
      if (c < nstat) A[k * NCHP + c] = src[t];
    }
  }
  __syncwarp(gmask);
  // ---- dynamic channel values: A^c_vw = sum_z T_vwz f_c(z) -------------------------------
  if (ndyn > 0) {
    const long long tz0 = S.et_tzptr[v];
    const int Wd = S.et_tzw[v];
    const unsigned char* __restrict__ zs = S.et_tzslot + tz0;
    for (int ts = 0; ts < S.vx_nts; ++ts) {
      const double* __restrict__ Tvw = S.et_tvw + 2 * (ts * S.nnzb + r0);
      const double* __restrict__ Tz = S.et_tz + ts * S.et_tz_total + tz0;
      const int d0 = itab[S.vx_o_dynptr + ts], d1 = itab[S.vx_o_dynptr + ts + 1];
      for (int kb = 0; kb < deg; kb += G) {
        const int k = kb + g;
        const bool act = k < deg && k != kself;
        double tvk = 0.0, twk = 0.0, tj[FB_ET_WREG];
        int zj[FB_ET_WREG];
        if (act) { tvk = Tvw[2 * k]; twk = Tvw[2 * k + 1]; }
#pragma unroll
        for (int j = 0; j < FB_ET_WREG; ++j) {
          tj[j] = 0.0; zj[j] = 0;
          if (act && j < Wd) {
            const int zz = zs[(long long)j * deg + k];
            if (zz != 0xff) { zj[j] = zz; tj[j] = Tz[(long long)j * deg + k]; }
          }
        }
        if (act) {
          for (int dd = d0; dd < d1; ++dd) {
            const int c = itab[S.vx_o_dyn + 2 * dd], fc = itab[S.vx_o_dyn + 2 * dd + 1];
            double a = tvk * nb[kself * NBW + fc] + twk * nb[k * NBW + fc];
#pragma unroll
            for (int j = 0; j < FB_ET_WREG; ++j) a += tj[j] * nb[zj[j] * NBW + fc];
            for (int j = FB_ET_WREG; j < Wd; ++j) {
              const int zz = zs[(long long)j * deg + k];
              if (zz != 0xff) a += Tz[(long long)j * deg + k] * nb[zz * NBW + fc];
            }
            A[k * NCHP + c] = a;
          }
        }
      }
      for (int dd = d0; dd < d1; ++dd) {   // diagonal edge: sum over all neighbours z of T_vvz f(z)
        const int c = itab[S.vx_o_dyn + 2 * dd], fc = itab[S.vx_o_dyn + 2 * dd + 1];
        double p = 0.0;
        for (int k = g; k < deg; k += G) p += Tvw[2 * k] * nb[k * NBW + fc];
        p = fb_group_sum<G>(p, gmask);
        if (g == 0) A[kself * NCHP + c] = p;
      }
    }
    __syncwarp(gmask);
    double* __restrict__ dst = dyn + (long long)r0 * NDP;
    for (int t = g; t < deg * NDP; t += G) {
      const int k = fb_fastdiv((unsigned)t, S.vx_m_ndp), c = t - k * NDP;
      dst[t] = c < ndyn ? A[k * NCHP + nstat + c] : 0.0;
    }
  }
  if (!doF) return;
  // ---- residual ------------------------------------------------------------------------------
  const int nft = S.vx_nft;
  for (int t = g; t < nft; t += G) {
    const int chan = itab[S.vx_o_ft + 3 * t], gcol = itab[S.vx_o_ft + 3 * t + 1], fl = itab[S.vx_o_ft + 3 * t + 2];
    double acc = 0.0;
    for (int k = 0; k < deg; ++k) acc += A[k * NCHP + chan] * nb[k * NBW + gcol];
    res[t] = acc * dtab[S.vx_o_ftc + t] * ((fl & 1) ? S.inv_dt : 1.0);
  }
  __syncwarp(gmask);
  for (int s = g; s < ns; s += G) {
    double acc = S.load[v * ns + s];
    const int a0 = itab[S.vx_o_ftptr + s], a1 = itab[S.vx_o_ftptr + s + 1];
    for (int t = a0; t < a1; ++t) acc += res[t];
    facc[s] = acc;
  }
  if (colbc) {   // lifting b += J[:, B] (g - x_B) with the live entries of the Dirichlet columns
    __syncwarp(gmask);
    const int njt = S.vx_njt;
    for (int t = g; t < njt; t += G) {
      const int chan = itab[S.vx_o_jt + 2 * t], fl = itab[S.vx_o_jt + 2 * t + 1];
      const int cs = S.pair_cs[itab[S.vx_o_jtpair + t]];
      double acc = 0.0;
      for (int k = 0; k < deg; ++k)
        if ((nbinfo[k] >> cs) & 1u)
          acc += A[k * NCHP + chan] * (S.bc_vals[S.bc_map[(long long)cols[k] * ns + cs]] - nb[k * NBW + cs]);
      res[t] = acc * dtab[S.vx_o_jtc + t] * ((fl & 1) ? S.inv_dt : 1.0);
    }
    __syncwarp(gmask);
    for (int s = g; s < ns; s += G) {
      double acc = facc[s];
      const int a0 = itab[S.vx_o_jtptr + S.pair_pp[s]], a1 = itab[S.vx_o_jtptr + S.pair_pp[s + 1]];
      for (int t = a0; t < a1; ++t) acc += res[t];
      facc[s] = acc;
    }
  }
  for (int s = g; s < ns; s += G) {
    const long long dof = v * ns + s;
    if ((rowbc >> s) & 1u) F[dof] = nb[kself * NBW + s] - S.bc_vals[S.bc_map[dof]];
    else F[dof] = facc[s];
  }
}

// ---------------------------------------------------------------------------
// channel operator -> species-sparse block-CSR values (the J of the C ABI), with the strong
// Dirichlet rows / columns applied (identity rows, eliminated columns)
// ---------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ void fb_et_expand_body(const DevSpec& S, double* __restrict__ J, double* smem) {
  const EtLayout L = fb_et_layout(S, false);
  const int g = threadIdx.x & (G - 1), grp = threadIdx.x / G, gpb = blockDim.x / G;
  double* dtab = smem;
  int* itab = (int*)(smem + ((S.vx_dtab_n + 1) & ~1));
  for (int t = threadIdx.x; t < S.vx_dtab_n; t += blockDim.x) dtab[t] = S.vx_dtab[t];
  for (int t = threadIdx.x; t < S.vx_itab_n; t += blockDim.x) itab[t] = S.vx_itab[t];
  __syncthreads();
  const long long v = (long long)blockIdx.x * gpb + grp;
  if (v >= S.n_owned) return;
  const unsigned gmask = fb_group_mask<G>();
  const int NCHP = L.nchp, nstat = S.vx_nstat, ndyn = S.vx_ndyn, NSP = S.vx_nsp, NDP = S.vx_ndp, np = S.npairs;
  double* W_ = smem + L.grp0 + (size_t)grp * L.per_grp;
  double* A = W_ + L.o_A;
  int* cols = (int*)(W_ + L.o_ints);
  unsigned* nbinfo = (unsigned*)(cols + S.maxdeg);
  const int r0 = S.rowptr[v], deg = S.rowptr[v + 1] - r0;
  const unsigned vi = S.vinfo[v];
  const unsigned rowbc = vi & 0xffffu;
  const bool anybc = vi != 0u;
  int kself = 0;
  if (anybc) {
    for (int k = g; k < deg; k += G) {
      const int w = S.colidx[r0 + k];
      nbinfo[k] = S.vinfo[w];
      if (w == (int)v) kself = k;
    }
    kself = fb_group_max<G>(kself, gmask);
  }
  if (nstat > 0) {
    const double* __restrict__ src = S.et_stat + (long long)r0 * NSP;
    for (int t = g; t < deg * NSP; t += G) {
      const int k = fb_fastdiv((unsigned)t, S.vx_m_nsp), c = t - k * NSP;
      if (c < nstat) A[k * NCHP + c] = src[t];
    }
  }
  if (ndyn > 0) {
    const double* __restrict__ src = S.et_dyn + (long long)r0 * NDP;
    for (int t = g; t < deg * NDP; t += G) {
      const int k = fb_fastdiv((unsigned)t, S.vx_m_ndp), c = t - k * NDP;
      if (c < ndyn) A[k * NCHP + nstat + c] = src[t];
    }
  }
  __syncwarp(gmask);
  const long long rowbase = (long long)r0 * np;
  for (int t = g; t < deg * np; t += G) {
    const int k = fb_fastdiv((unsigned)t, S.vx_m_np), p = t - k * np;
    double val = 0.0;
    for (int e = itab[S.vx_o_jtptr + p]; e < itab[S.vx_o_jtptr + p + 1]; ++e) {
      double coef = dtab[S.vx_o_jtc + e];
      if (itab[S.vx_o_jt + 2 * e + 1] & 1) coef *= S.inv_dt;
      val += coef * A[k * NCHP + itab[S.vx_o_jt + 2 * e]];
    }
    if (anybc) {
      const int s = S.pair_row[p], cs = S.pair_cs[p];
      if ((nbinfo[k] >> cs) & 1u) val = 0.0;
      if ((rowbc >> s) & 1u) val = (k == kself && cs == s) ? 1.0 : 0.0;
    }
    J[rowbase + t] = val;
  }
}
