// Vertex-block assembly of residual and Jacobian for diffusion-trapping-reaction systems
// (fp64, P1 simplices).  Replaces the FFCx tabulate_tensor kernels + dolfinx
// assemble_vector / assemble_matrix / apply_lifting / set_bc behind NonlinearProblem
// (reference src/festim/problem.py:170-182) for the form of
// src/festim/hydrogen_transport_problem.py:870-997 whenever every solution-dependent volume
// term is a product of P1 functions with a cell-wise coefficient (diffusion, mass,
// advection, mass-action reactions of reaction.py:179-224 with affine factors,
// species.py:167-176).  festim_b200/lowering.py (_vertex_tables) describes the algebra;
// tests/vtx_emulator.py is the numpy model of this file.
//
// Two kernels:
//  * k_cell_records (per temperature / velocity update, never inside Newton): one record of
//    vx_rec doubles per cell - |K|, diffusion pair matrices |K| mean(exp(-E/kT)) G_a.G_b,
//    rate-weighted moments |K| sum_q w_q rho_m phi_a phi_b [phi_c] (one exp per activation
//    energy and quadrature point), advection matrices.  All temperature dependence of the
//    assembly lives here.
//  * k_assemble_vtx (every Newton iteration): one warp per vertex = one block row of J.
//      phase 0  values of u, u - u_n and the affine reaction factors at the row's
//               neighbours -> shared memory (each loaded once per row);
//      phase 1  lanes = (incident cell, local column j): the cell records of a chunk of
//               cells are staged in shared memory with coalesced loads, every lane
//               evaluates the <= vx_nch channel values C^c_ij of its (cell, j);
//      phase 2  lanes = neighbours: each neighbour slot gathers the channel values of the
//               cells that contain the edge (bit masks built with 32-bit shared atomics),
//               so the "edge channels" A^c_vw are summed without fp64 atomics, in a fixed
//               order (bitwise reproducible);
//      phase 3  residual: lanes = table entries, F_s = sum_e coef_e sum_w A^c_vw g_e(w);
//      phase 4  Jacobian: every stored entry is a short linear combination of the edge
//               channels (table per species pair); the block row is built in shared
//               memory and written to HBM once, coalesced, with the strong Dirichlet
//               rows / columns applied on the way (lifting with the live entries, identity
//               rows: dolfinx NonlinearProblem semantics).
// J is written exactly once per assembly and nothing is read-modify-written in global
// memory; the only per-cell traffic is the record (read by the D1 rows that share it,
// L2-resident between them).
#pragma once
#include "fb_device.cuh"

enum { FB_CQ_KD = 0, FB_CQ_M1 = 1, FB_CQ_M3 = 2, FB_CQ_ADV = 3 };
enum { FB_CH_MASS = 0, FB_CH_PAIR = 1, FB_CH_TRI = 2, FB_CH_FULL = 3 };
enum { FB_XL_INVDT = 1 << 16, FB_XL_STORE = 1 << 17 };

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

// ---------------------------------------------------------------------------
// per-cell record
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
// shared-memory layout of k_assemble_vtx (doubles unless noted); computed on the host too
// ---------------------------------------------------------------------------
struct VtxLayout {
  int tab_d;       // block-shared double table (dtab), then the int table
  int tab_i;       // ints (offset in ints from the start of the int area)
  int warp0;       // start (in doubles) of the per-warp areas
  int per_warp;    // doubles per warp
  // offsets inside a warp area (doubles)
  int o_nb, o_A, o_R, o_fres, o_flift, o_ints;
  int cpc;         // cells per chunk
};

template <int DIM>
__host__ __device__ inline VtxLayout fb_vtx_layout(const DevSpec& S) {
  constexpr int D1 = DIM + 1;
  VtxLayout L;
  const int md = S.maxdeg;
  L.cpc = 32 / D1;
  L.tab_d = 0;
  const int dt = (S.vx_dtab_n + 1) & ~1;
  L.tab_i = 0;
  const int it_doubles = (S.vx_itab_n + 1) / 2;
  L.warp0 = (dt + it_doubles + 1) & ~1;
  int o = 0;
  L.o_A = o; o += md * S.vx_nchp;                   // 16-byte aligned rows (nchp even)
  L.o_R = o;
  {
    const int stage = L.cpc * S.vx_recp + 1 + 32 * S.vx_nchp;   // records (+1 keeps vbuf even) + channel values
    const int row = md * S.vx_rbp;
    o += ((stage > row ? stage : row) + 1) & ~1;
  }
  L.o_nb = o; o += md * S.vx_nbw;
  L.o_fres = o; o += S.vx_nft_max > 0 ? S.vx_nft_max : 1;
  L.o_flift = o; o += FB_MAX_NS;
  o = (o + 1) & ~1;
  L.o_ints = o;                                      // ints: cols[md], nbinfo[md], smask[md], cell[cpc], info[cpc]
  o += (3 * md + 2 * L.cpc + 1) / 2 + 1;
  L.per_warp = (o + 1) & ~1;
  return L;
}

// n / d through the magic number ceil(2^32 / d) (exact while n * d < 2^32); magic 0 means d == 1
__device__ __forceinline__ int fb_fastdiv(unsigned n, unsigned magic) {
  return magic ? (int)__umulhi(n, magic) : (int)n;
}

// ---------------------------------------------------------------------------
// the assembly kernel body: one warp per owned vertex
// ---------------------------------------------------------------------------
template <int DIM>
__device__ __forceinline__ void fb_assemble_vtx_body(const DevSpec& S, const double* __restrict__ u,
                                                     const double* __restrict__ un, double* __restrict__ F,
                                                     double* __restrict__ J, int flags, double* smem) {
  constexpr int D1 = DIM + 1;
  const VtxLayout L = fb_vtx_layout<DIM>(S);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  // ---- block-shared tables -------------------------------------------------------------
  double* dtab = smem + L.tab_d;
  int* itab = (int*)(smem + ((S.vx_dtab_n + 1) & ~1));
  for (int t = threadIdx.x; t < S.vx_dtab_n; t += blockDim.x) dtab[t] = S.vx_dtab[t];
  for (int t = threadIdx.x; t < S.vx_itab_n; t += blockDim.x) itab[t] = S.vx_itab[t];
  __syncthreads();
  const long long v = (long long)blockIdx.x * wpb + warp;
  if (v >= S.n_owned) return;
  const unsigned FULL = 0xffffffffu;
  const int ns = S.ns, npairs = S.npairs, NBW = S.vx_nbw, NCHP = S.vx_nchp, RECP = S.vx_recp, RBP = S.vx_rbp;
  const int NF = S.n_factors, REC = S.vx_rec;
  const bool doF = (flags & FB_ASSEMBLE_F) != 0, doJ = (flags & FB_ASSEMBLE_J) != 0;
  double* W = smem + L.warp0 + (size_t)warp * L.per_warp;
  double* A = W + L.o_A;
  double* recs = W + L.o_R;
  double* vbuf = recs + ((L.cpc * RECP + 1) & ~1);
  double* rowbuf = W + L.o_R;
  double* nb = W + L.o_nb;
  double* fres = W + L.o_fres;
  double* flift = W + L.o_flift;
  int* cols = (int*)(W + L.o_ints);
  unsigned* nbinfo = (unsigned*)(cols + S.maxdeg);
  unsigned* smask = nbinfo + S.maxdeg;
  int* cpk = (int*)(smask + S.maxdeg);            // chunk cells: packed cell * D1 + local index
  unsigned* cinfo = (unsigned*)(cpk + L.cpc);     // chunk cells: slot word

  const int r0 = S.rowptr[v], deg = S.rowptr[v + 1] - r0;
  const long long rowbase = (long long)r0 * npairs;
  const unsigned vi = S.vinfo[v];
  const unsigned rowbc = vi & 0xffffu;
  const bool colbc = (vi >> 16) & 1u;
  const bool needJ = doJ || (doF && colbc);
  const int mode_use = (doF ? 1 : 0) | (needJ ? 2 : 0);
  const int e0 = S.v2c_ptr[v], ncell = S.v2c_ptr[v + 1] - e0;

  // ---- phase 0: neighbour table [u | u - u_n | factors | 1] ------------------------------
  int kself = 0;
  for (int k = lane; k < deg; k += 32) {
    const int w = S.colidx[r0 + k];
    cols[k] = w;
    nbinfo[k] = S.vinfo[w];
    if (w == (int)v) kself = k;
  }
  kself = __reduce_max_sync(FULL, kself);
  if (lane < FB_MAX_NS) flift[lane] = 0.0;
  __syncwarp();
  for (int t = lane; t < deg * ns; t += 32) {
    const int k = fb_fastdiv((unsigned)t, S.vx_m_ns), s = t - k * ns;
    const long long dof = (long long)cols[k] * ns + s;
    const double uu = u[dof];
    nb[k * NBW + s] = uu;
    nb[k * NBW + ns + s] = S.transient ? uu - un[dof] : 0.0;
  }
  for (int k = lane; k < deg; k += 32) nb[k * NBW + 2 * ns + NF] = 1.0;
  __syncwarp();
  for (int t = lane; t < deg * NF; t += 32) {
    const int k = fb_fastdiv((unsigned)t, S.vx_m_nf), f = t - k * NF;
    const int* h = S.fac_hdr + 4 * f;
    double val = h[0] ? S.scalars[h[1]] : S.fac_a0[f];
    for (int q = 0; q < h[3]; ++q) val += S.fac_coef[h[2] + q] * nb[k * NBW + S.fac_species[h[2] + q]];
    nb[k * NBW + 2 * ns + f] = val;
  }
  // ---- which volume tags meet at this vertex ----------------------------------------------
  unsigned tagmask = 1u;
  if (S.n_vtags > 1) {
    tagmask = 0u;
    for (int e = lane; e < ncell; e += 32) tagmask |= 1u << S.cell_tag[S.v2c_idx[e0 + e] / D1];
    tagmask = __reduce_or_sync(FULL, tagmask);
  }
  __syncwarp();

  double Facc = 0.0;   // lane s < ns
  const double mscale = 1.0 / (double)(D1 * (D1 + 1));
  const int el = lane / D1, jl = lane - el * D1;   // phase-1 role: (chunk cell, local column)
  bool first_tag = true;
  while (tagmask) {
    const int tag = __ffs(tagmask) - 1;
    tagmask &= tagmask - 1;
    const int ch0 = itab[S.vx_o_chptr + tag], nch = itab[S.vx_o_chptr + tag + 1] - ch0;
    const int* chs = itab + S.vx_o_ch + 4 * ch0;
    for (int t = lane; t < deg * NCHP; t += 32) A[t] = 0.0;
    double dacc = 0.0;   // lane c < nch: diagonal (w = v) edge channel
    for (int eb = 0; eb < ncell; eb += L.cpc) {
      const int ncc = min(L.cpc, ncell - eb);
      __syncwarp();
      for (int k = lane; k < deg; k += 32) smask[k] = 0u;
      if (lane < ncc) {
        cpk[lane] = S.v2c_idx[e0 + eb + lane];       // cell * D1 + local index of v
        cinfo[lane] = S.v2c_slot[e0 + eb + lane];    // row slot of every cell vertex, 8 bits each
      }
      __syncwarp();
      // stage the records of the chunk (coalesced rows of REC doubles)
      for (int c2 = 0; c2 < ncc; ++c2) {
        const double* __restrict__ src = S.cellrec + (long long)(cpk[c2] / D1) * REC;
        for (int o = lane; o < REC; o += 32) recs[c2 * RECP + o] = src[o];
      }
      __syncwarp();
      // ---- phase 1 --------------------------------------------------------------------------
      bool valid = el < ncc;
      int i = 0;
      unsigned sw = 0;
      if (valid) {
        const int packed = cpk[el];
        i = packed - (packed / D1) * D1;
        sw = cinfo[el];
        if (S.n_vtags > 1) valid = S.cell_tag[packed / D1] == tag;
      }
      const double* rec = recs + el * RECP;
      int slotb[D1];
#pragma unroll
      for (int b = 0; b < D1; ++b) slotb[b] = (int)((sw >> (8 * b)) & 0xffu);
      int myslot = 0;
#pragma unroll
      for (int b = 0; b < D1; ++b)
        if (b == jl) myslot = slotb[b];
      int s3[D1];
#pragma unroll
      for (int b = 0; b < D1; ++b) s3[b] = fb_sym3<D1>(i, jl, b);
      const int s2 = fb_sym2<D1>(i, jl);
      for (int c = 0; c < nch; ++c) {
        const int ctype = chs[4 * c], off = chs[4 * c + 1], fac = chs[4 * c + 2], use = chs[4 * c + 3];
        double val = 0.0;
        if (valid && (use & mode_use)) {
          if (ctype == FB_CH_MASS) val = rec[0] * mscale * (i == jl ? 2.0 : 1.0);
          else if (ctype == FB_CH_PAIR) val = rec[off + s2];
          else if (ctype == FB_CH_FULL) val = rec[off + i * D1 + jl];
          else {
#pragma unroll
            for (int b = 0; b < D1; ++b) val += rec[off + s3[b]] * nb[slotb[b] * NBW + fac];
          }
        }
        vbuf[lane * NCHP + c] = val;
      }
      for (int c = nch; c < NCHP; ++c) vbuf[lane * NCHP + c] = 0.0;
      if (valid && jl != i) atomicOr(&smask[myslot], 1u << lane);
      const unsigned dmask = __ballot_sync(FULL, valid && jl == i);
      __syncwarp();
      // ---- phase 2 --------------------------------------------------------------------------
      for (int k = lane; k < deg; k += 32) {
        unsigned m = smask[k];
        while (m) {
          const int b = __ffs(m) - 1;
          m &= m - 1;
          const double2* __restrict__ src = (const double2*)(vbuf + b * NCHP);
          double2* dst = (double2*)(A + k * NCHP);
          for (int c2 = 0; c2 < NCHP / 2; ++c2) {
            double2 a2 = dst[c2];
            const double2 s2v = src[c2];
            a2.x += s2v.x; a2.y += s2v.y;
            dst[c2] = a2;
          }
        }
      }
      if (lane < nch) {
        unsigned m = dmask;
        while (m) {
          const int b = __ffs(m) - 1;
          m &= m - 1;
          dacc += vbuf[b * NCHP + lane];
        }
      }
    }
    __syncwarp();
    if (lane < nch) A[kself * NCHP + lane] += dacc;
    __syncwarp();
    // ---- phase 3: residual --------------------------------------------------------------------
    if (doF) {
      const int f0 = itab[S.vx_o_ftptr + tag * ns], f1 = itab[S.vx_o_ftptr + tag * ns + ns];
      for (int t = f0 + lane; t < f1; t += 32) {
        const int chan = itab[S.vx_o_ft + 3 * t], gcol = itab[S.vx_o_ft + 3 * t + 1], fl = itab[S.vx_o_ft + 3 * t + 2];
        double acc = 0.0;
        for (int k = 0; k < deg; ++k) acc += A[k * NCHP + chan] * nb[k * NBW + gcol];
        fres[t - f0] = acc * dtab[S.vx_o_ftc + t] * ((fl & 1) ? S.inv_dt : 1.0);
      }
      __syncwarp();
      if (lane < ns && !((rowbc >> lane) & 1u)) {
        const int a0 = itab[S.vx_o_ftptr + tag * ns + lane], a1 = itab[S.vx_o_ftptr + tag * ns + lane + 1];
        for (int t = a0; t < a1; ++t) Facc += fres[t - f0];
      }
      __syncwarp();
    }
    // ---- phase 4: Jacobian block row ----------------------------------------------------------
    if (needJ) {
      const int LPK = S.vx_lpk, KPP = 32 / LPK;
      const int kk = lane / LPK, h = lane - kk * LPK;
      const int xoff = itab[S.vx_o_xlhdr + 2 * tag], Lx = itab[S.vx_o_xlhdr + 2 * tag + 1];
      const int* xl = itab + S.vx_o_xl + xoff + h * Lx;
      const double* xc = dtab + S.vx_o_xlc + xoff + h * Lx;
      for (int kb = 0; kb < deg; kb += KPP) {
        const int k = kb + kk;
        const bool act = k < deg;
        const double* Ak = A + (act ? k : 0) * NCHP;
        double* rk = rowbuf + (act ? k : 0) * RBP;
        double acc = 0.0;
        for (int q = 0; q < Lx; ++q) {
          const int ent = xl[q];
          double coef = xc[q];
          if (ent & FB_XL_INVDT) coef *= S.inv_dt;
          acc += coef * Ak[(ent >> 8) & 0xff];
          if (ent & FB_XL_STORE) {
            if (act) rk[ent & 0xff] = acc;
            acc = 0.0;
          }
        }
      }
      __syncwarp();
      const int nrow = deg * npairs;
      const bool anybc = rowbc != 0u || colbc;
      for (int t = lane; t < nrow; t += 32) {
        const int k = fb_fastdiv((unsigned)t, S.vx_m_np), p = t - k * npairs;
        double val = rowbuf[k * RBP + p];
        bool identity = false;
        if (anybc) {
          const int s = S.pair_row[p], cs = S.pair_cs[p];
          if ((nbinfo[k] >> cs) & 1u) {   // Dirichlet column: lifting with the live entry, then eliminate
            if (doF && !((rowbc >> s) & 1u)) {
              const long long cdof = (long long)cols[k] * ns + cs;
              atomicAdd(&flift[s], val * (S.bc_vals[S.bc_map[cdof]] - nb[k * NBW + cs]));
            }
            val = 0.0;
          }
          if ((rowbc >> s) & 1u) { val = (k == kself && cs == s) ? 1.0 : 0.0; identity = true; }
        }
        if (doJ) {
          if (first_tag || identity) J[rowbase + t] = val;
          else J[rowbase + t] += val;
        }
      }
      __syncwarp();
    }
    first_tag = false;
  }
  __syncwarp();
  if (doF && lane < ns) {
    const long long dof = v * ns + lane;
    if ((rowbc >> lane) & 1u) F[dof] = u[dof] - S.bc_vals[S.bc_map[dof]];
    else F[dof] = Facc + flift[lane] + S.load[dof];
  }
}
