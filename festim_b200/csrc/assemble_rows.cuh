// Warp-per-vertex gather assembly for multi-species diffusion-trapping-reaction
// systems (fp64, P1 simplices).
//
// The thread-per-vertex kernel (assemble_kernels.cuh) evaluates reaction rates point by
// point and read-modify-writes the Jacobian row in global memory; with 12 species the
// row is 990 doubles and that dominates everything.  Here a warp owns one vertex:
//   * the whole block row (deg x |P| values) is accumulated in SHARED memory and
//     written to HBM exactly once, coalesced, in the final storage order;
//   * reaction terms  k0 exp(-Ek/kT) f1 f2 - p0 exp(-Ep/kT) g1  (every factor affine in
//     the species, reference reaction.py:179-224 / species.py:167-176) are assembled from
//     rate-weighted moments of the basis,  M^m_ab = |K| sum_q w_q rho_m(x_q) phi_i phi_a
//     phi_b,  one exp per distinct activation energy and quadrature point; the sum over
//     quadrature points is the one the reference's generated kernel performs, reordered
//     (P1 fields are linear in the barycentric coordinates), so results agree to
//     round-off;
//   * lanes work on quadrature points (rates), moment entries, species pairs (Jacobian)
//     and species (residual) in turn; no atomics, deterministic.
// Strong Dirichlet rows/columns are applied to the row in shared memory before the
// flush (lifting with the live entries, identity row), as in the scalar kernel.
#pragma once
#include "fb_device.cuh"

template <int D1>
__device__ __forceinline__ int fb_sym(int a, int b) {
  if (a > b) { const int t = a; a = b; b = t; }
  return a * D1 - (a * (a - 1)) / 2 + (b - a);
}

// per-warp shared-memory need (doubles) of fb_assemble_rows_body
template <int DIM>
__host__ __device__ inline int fb_rows_per_warp(const DevSpec& S) {
  constexpr int D1 = DIM + 1;
  const int n = S.maxdeg * S.npairs + 2 * D1 * S.ns + S.n_factors * D1 + S.n_rslots * S.nq_max +
                2 * (S.n_rslots * D1 * D1 + S.n_rslots * D1 + S.n_combos * D1) + S.ns + 2;
  return (n + 1) & ~1;
}

template <int DIM>
__device__ __forceinline__ void fb_assemble_rows_body(const DevSpec& S, const double* __restrict__ u,
                                                      const double* __restrict__ un,
                                                      double* __restrict__ F, double* __restrict__ J,
                                                      int flags, double* smem, int per_warp) {
  constexpr int D1 = DIM + 1;
  constexpr int NSYM = D1 * (D1 + 1) / 2;
  constexpr int DD = D1 * D1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const long long v = (long long)blockIdx.x * wpb + warp;
  if (v >= S.n_owned) return;
  const int ns = S.ns, npairs = S.npairs, NR = S.n_rslots, NF = S.n_factors, NC = S.n_combos;
  const bool doF = (flags & FB_ASSEMBLE_F) != 0, doJ = (flags & FB_ASSEMBLE_J) != 0;
  const int r0 = S.rowptr[v], deg = S.rowptr[v + 1] - r0;
  const long long rowbase = (long long)r0 * npairs;
  const int nrow = deg * npairs;
  // per-warp shared memory
  double* rowbuf = smem + (size_t)warp * per_warp;
  double* uc = rowbuf + (size_t)S.maxdeg * npairs;   // [D1][ns]
  double* unc = uc + D1 * ns;
  double* fac = unc + D1 * ns;                        // [NF][D1]
  double* rho = fac + NF * D1;                        // [NR][nq_max]
  double* MJ = rho + NR * S.nq_max;                   // [NR][D1][D1]  moments, J rule
  double* MF = MJ + NR * DD;                          //               moments, F rule
  double* M1J = MF + NR * DD;                         // [NR][D1]      first-order moments
  double* M1F = M1J + NR * D1;
  double* VJ = M1F + NR * D1;                         // [NC][D1]      V = M fac  per (slot, factor)
  double* VF = VJ + NC * D1;
  double* flift = VF + NC * D1;                       // [ns]
  for (int t = lane; t < nrow; t += 32) rowbuf[t] = 0.0;
  if (lane < ns) flift[lane] = 0.0;
  // Dirichlet rows of this vertex, any Dirichlet column in the row
  const unsigned rowbc = __ballot_sync(0xffffffffu, lane < ns && S.bc_map[v * ns + lane] >= 0);
  bool anybc = false;
  for (int t = lane; t < deg * ns; t += 32)
    anybc |= S.bc_map[(long long)S.colidx[r0 + t / ns] * ns + t % ns] >= 0;
  const bool colbc = __any_sync(0xffffffffu, anybc);
  const bool needJ = doJ || (doF && colbc);
  double Facc = 0.0;  // lane s < ns: residual of species s
  const double mscale = 1.0 / (double)(D1 * (D1 + 1));
  __syncwarp();

  for (int e = S.v2c_ptr[v]; e < S.v2c_ptr[v + 1]; ++e) {
    const int packed = S.v2c_idx[e];
    const long long c = packed / D1;
    const int i = packed - (int)c * D1;
    int a[FB_MAX_D1], slot[FB_MAX_D1];
    double gg[FB_MAX_D1], vol;
    {
      double X[FB_MAX_D1][3], G[FB_MAX_D1][3];
      fb_load_cell<DIM>(S, c, a, X);
      vol = fb_geometry<DIM>(X, G);
#pragma unroll
      for (int j = 0; j < D1; ++j) {
        double d = 0.0;
#pragma unroll
        for (int k = 0; k < DIM; ++k) d += G[i][k] * G[j][k];
        gg[j] = d;
      }
    }
    const int tag = S.cell_tag[c];
    if (S.v2c_slot != nullptr) {
      const unsigned ps = S.v2c_slot[e];
#pragma unroll
      for (int j = 0; j < D1; ++j) slot[j] = (ps >> (8 * j)) & 0xffu;
    } else {
#pragma unroll
      for (int j = 0; j < D1; ++j) slot[j] = fb_find_slot(S, (int)v, a[j]);
    }
    __syncwarp();
    for (int t = lane; t < D1 * ns; t += 32) {
      const long long dof = (long long)a[t / ns] * ns + t % ns;
      uc[t] = u[dof];
      unc[t] = S.transient ? un[dof] : 0.0;
    }
    __syncwarp();
    for (int t = lane; t < NF * D1; t += 32) {
      const int f = t / D1, b = t - f * D1;
      const int* h = S.fac_hdr + 4 * f;
      double val = h[0] ? S.scalars[h[1]] : S.fac_a0[f];
      for (int k = 0; k < h[3]; ++k) val += S.fac_coef[h[2] + k] * uc[b * ns + S.fac_species[h[2] + k]];
      fac[t] = val;
    }
    // ---- rate-weighted moments (J rule, and the F rule when it differs) ----------------
    const int* wo = S.wtab_off + 4 * tag;
    const bool same_rule = wo[0] == wo[2];
    for (int which = 1; which >= 0; --which) {
      if (which == 0 && (!doF || same_rule)) continue;
      if (which == 1 && !(needJ || (doF && same_rule))) continue;
      const int woff = wo[2 * which], nq = wo[2 * which + 1];
      const double* W = S.wtab + woff + (size_t)i * nq * NSYM;          // W[i][q][sym]
      const double* W0 = S.wtab + woff + (size_t)D1 * nq * NSYM + i * NSYM;
      double* M = which ? MJ : MF;
      double* M1 = which ? M1J : M1F;
      double* V = which ? VJ : VF;
      if (S.T_mode) {
        const int qoff = S.quad_v[tag * 4 + 2 * which];
        __syncwarp();
        for (int q = lane; q < nq; q += 32) {
          const double* bary = S.quad_bary + (long long)(qoff + q) * D1;
          double T = 0.0;
#pragma unroll
          for (int j = 0; j < D1; ++j) T += bary[j] * S.T_nodal[a[j]];
          const double ikT = 1.0 / (FB_KB * T);
          for (int m = 0; m < NR; ++m) rho[m * S.nq_max + q] = exp(-S.rslot_E[m] * ikT);
        }
        __syncwarp();
        for (int t = lane; t < NR * DD; t += 32) {
          const int m = t / DD, ab = t - m * DD;
          const int sym = fb_sym<D1>(ab / D1, ab % D1);
          const double* rm = rho + m * S.nq_max;
          double acc = 0.0;
#pragma unroll 4
          for (int q = 0; q < nq; ++q) acc += rm[q] * W[q * NSYM + sym];
          M[t] = acc * vol;
        }
      } else {
        for (int t = lane; t < NR * DD; t += 32) {
          const int m = t / DD, ab = t - m * DD;
          M[t] = exp(-S.rslot_E[m] / (FB_KB * S.T_const)) * W0[fb_sym<D1>(ab / D1, ab % D1)] * vol;
        }
      }
      __syncwarp();
      for (int t = lane; t < NR * D1; t += 32) {
        double acc = 0.0;
#pragma unroll
        for (int b = 0; b < D1; ++b) acc += M[t * D1 + b];
        M1[t] = acc;
      }
      for (int t = lane; t < NC * D1; t += 32) {
        const int cidx = t / D1, j = t - cidx * D1;
        const double* Mm = M + S.combo[2 * cidx] * DD + j * D1;
        const double* fv = fac + S.combo[2 * cidx + 1] * D1;
        double acc = 0.0;
#pragma unroll
        for (int b = 0; b < D1; ++b) acc += fv[b] * Mm[b];
        V[t] = acc;
      }
    }
    __syncwarp();
    const double* M1f = same_rule ? M1J : M1F;
    const double* Vf = same_rule ? VJ : VF;

    // ---- Jacobian: lanes over species pairs ------------------------------------------
    if (needJ) {
      for (int p = lane; p < npairs; p += 32) {
        const int s = S.pair_row[p];
        if (rowbc >> s & 1u) continue;
        const int jidx = p - S.pair_pp[s];
        const int sc = S.pair_cs[p];
        double acc[FB_MAX_D1];
#pragma unroll
        for (int j = 0; j < D1; ++j) acc[j] = 0.0;
        for (int k = S.jc_ptr[tag * npairs + p]; k < S.jc_ptr[tag * npairs + p + 1]; ++k) {
          const int cidx = S.jc[2 * k + 1];
          const double coef = S.jc_coef[k];
          const double* Vc = cidx >= 0 ? VJ + cidx * D1 : M1J + (-cidx - 1) * D1;
#pragma unroll
          for (int j = 0; j < D1; ++j) acc[j] += coef * Vc[j];
        }
        if (s == sc) {
          const int ds = S.diff_slot[tag * ns + s];
          double kJ = 0.0;
          if (ds >= 0) kJ = S.diff_D0[tag * ns + s] * vol * S.dbar[((long long)S.n_slots + ds) * S.n_cells + c];
          const double mm = ((S.transient ? S.mass_dt[tag * ns + s] * S.inv_dt : 0.0) + S.mass_c[tag * ns + s]) *
                            vol * mscale;
#pragma unroll
          for (int j = 0; j < D1; ++j) acc[j] += kJ * gg[j] + mm * ((i == j) ? 2.0 : 1.0);
          for (int t = 0; t < S.n_adv; ++t) {
            if (S.adv[2 * t] != tag || S.adv[2 * t + 1] != s) continue;
            const double* vel = S.velocities[t];
            double X[FB_MAX_D1][3], G[FB_MAX_D1][3];   // recomputed: keeps G out of the hot loop's registers
            int a2[FB_MAX_D1];
            fb_load_cell<DIM>(S, c, a2, X);
            fb_geometry<DIM>(X, G);
#pragma unroll
            for (int j = 0; j < D1; ++j) {
              double accv = 0.0;
#pragma unroll
              for (int b = 0; b < D1; ++b) {
                double vg = 0.0;
#pragma unroll
                for (int k = 0; k < DIM; ++k) vg += vel[(long long)a[b] * DIM + k] * G[j][k];
                accv += vg * (b == i ? 2.0 : 1.0);
              }
              acc[j] += accv * vol * mscale;
            }
          }
        }
        double* dst = rowbuf + deg * S.pair_pp[s] + jidx;
        const int nc = S.pair_nc[s];
#pragma unroll
        for (int j = 0; j < D1; ++j) dst[slot[j] * nc] += acc[j];
      }
    }
    // ---- residual: lanes over species ---------------------------------------------------
    if (doF && lane < ns && !(rowbc >> lane & 1u)) {
      const int s = lane;
      const int ds = S.diff_slot[tag * ns + s];
      double kF = 0.0;
      if (ds >= 0) kF = S.diff_D0[tag * ns + s] * vol * S.dbar[(long long)ds * S.n_cells + c];
      const double m = S.transient ? S.mass_dt[tag * ns + s] * S.inv_dt * vol * mscale : 0.0;
      const double mc = S.mass_c[tag * ns + s] * vol * mscale;
#pragma unroll
      for (int j = 0; j < D1; ++j) {
        const double Mij = (i == j) ? 2.0 : 1.0;
        const double uj = uc[j * ns + s];
        Facc += (kF * gg[j] + mc * Mij) * uj;
        if (S.transient) Facc += m * Mij * (uj - unc[j * ns + s]);
      }
      for (int t = 0; t < S.n_adv; ++t) {
        if (S.adv[2 * t] != tag || S.adv[2 * t + 1] != s) continue;
        const double* vel = S.velocities[t];
        double X[FB_MAX_D1][3], G[FB_MAX_D1][3];
        int a2[FB_MAX_D1];
        fb_load_cell<DIM>(S, c, a2, X);
        fb_geometry<DIM>(X, G);
#pragma unroll
        for (int j = 0; j < D1; ++j) {
          double accv = 0.0;
#pragma unroll
          for (int b = 0; b < D1; ++b) {
            double vg = 0.0;
#pragma unroll
            for (int k = 0; k < DIM; ++k) vg += vel[(long long)a[b] * DIM + k] * G[j][k];
            accv += vg * (b == i ? 2.0 : 1.0);
          }
          Facc += accv * vol * mscale * uc[j * ns + s];
        }
      }
      for (int k = S.fc_ptr[tag * ns + s]; k < S.fc_ptr[tag * ns + s + 1]; ++k) {
        const int im = S.fc_mono[k];
        const int* mo = S.mono + 6 * im;
        double val = 0.0;
        if (mo[2] == 2) {
          const double* Vc = Vf + mo[5] * D1;
          const double* f1 = fac + mo[3] * D1;
#pragma unroll
          for (int p1 = 0; p1 < D1; ++p1) val += f1[p1] * Vc[p1];
        } else {
          const double* M1 = M1f + mo[1] * D1;
#pragma unroll
          for (int p1 = 0; p1 < D1; ++p1) val += (mo[2] == 1 ? fac[mo[3] * D1 + p1] : 1.0) * M1[p1];
        }
        Facc += S.fc_sign[k] * S.mono_coef[im] * val;
      }
    }
  }
  __syncwarp();

  // ---- Dirichlet columns: lifting with the live entries, then eliminate ---------------
  if (colbc) {
    for (int t = lane; t < nrow; t += 32) {
      int s = 0;
      while (deg * S.pair_pp[s + 1] <= t) ++s;
      const int rem = t - deg * S.pair_pp[s];
      const int nc = S.pair_nc[s];
      const int k = rem / nc, jidx = rem - k * nc;
      const long long cdof = (long long)S.colidx[r0 + k] * ns + S.pair_cs[S.pair_pp[s] + jidx];
      const int b = S.bc_map[cdof];
      if (b >= 0) {
        if (doF && !(rowbc >> s & 1u)) atomicAdd(&flift[s], rowbuf[t] * (S.bc_vals[b] - u[cdof]));
        rowbuf[t] = 0.0;
      }
    }
    __syncwarp();
  }
  // ---- Dirichlet rows: identity --------------------------------------------------------
  if (rowbc) {
    const int kself = fb_find_slot(S, (int)v, (int)v);
    for (int t = lane; t < nrow; t += 32) {
      int s = 0;
      while (deg * S.pair_pp[s + 1] <= t) ++s;
      if (!(rowbc >> s & 1u)) continue;
      const int rem = t - deg * S.pair_pp[s];
      const int nc = S.pair_nc[s];
      const int k = rem / nc, jidx = rem - k * nc;
      rowbuf[t] = (k == kself && S.pair_cs[S.pair_pp[s] + jidx] == s) ? 1.0 : 0.0;
    }
    __syncwarp();
  }
  // ---- flush: the row goes to HBM once, coalesced --------------------------------------
  if (doJ)
    for (int t = lane; t < nrow; t += 32) J[rowbase + t] = rowbuf[t];
  if (doF && lane < ns) {
    const long long dof = v * ns + lane;
    if (rowbc >> lane & 1u) F[dof] = u[dof] - S.bc_vals[S.bc_map[dof]];
    else F[dof] = Facc + flift[lane] + S.load[dof];
  }
}
