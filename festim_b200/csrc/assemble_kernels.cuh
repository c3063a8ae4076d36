// Element-kernel bodies (fp64, P1 simplices).  Included by assemble.cu (nvcc, with
// the byte-code interpreter of program.cuh as fb_prog_eval) and by the NVRTC
// translation unit festim_b200/jit.py generates (with fb_prog_eval compiled from
// the problem's own expressions, the analogue of FFCx's generated kernels).
#pragma once
#include "fb_device.cuh"

// ---------------------------------------------------------------------------
// element geometry: basis gradients G[a][k] and volume from vertex coordinates
// ---------------------------------------------------------------------------
template <int DIM>
__device__ __forceinline__ double fb_geometry(const double (*X)[3], double (*G)[3]) {
  if (DIM == 1) {
    const double J = X[1][0] - X[0][0];
    G[1][0] = 1.0 / J;
    G[0][0] = -1.0 / J;
    return fabs(J);
  } else if (DIM == 2) {
    const double e1x = X[1][0] - X[0][0], e1y = X[1][1] - X[0][1];
    const double e2x = X[2][0] - X[0][0], e2y = X[2][1] - X[0][1];
    const double det = e1x * e2y - e1y * e2x, id = 1.0 / det;
    G[1][0] = e2y * id;  G[1][1] = -e2x * id;
    G[2][0] = -e1y * id; G[2][1] = e1x * id;
    G[0][0] = -(G[1][0] + G[2][0]);
    G[0][1] = -(G[1][1] + G[2][1]);
    return 0.5 * fabs(det);
  } else {
    double e[3][3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int c = 0; c < 3; ++c) e[k][c] = X[k + 1][c] - X[0][c];
    double c0[3] = {e[1][1] * e[2][2] - e[1][2] * e[2][1], e[1][2] * e[2][0] - e[1][0] * e[2][2],
                    e[1][0] * e[2][1] - e[1][1] * e[2][0]};  // e2 x e3
    double c1[3] = {e[2][1] * e[0][2] - e[2][2] * e[0][1], e[2][2] * e[0][0] - e[2][0] * e[0][2],
                    e[2][0] * e[0][1] - e[2][1] * e[0][0]};  // e3 x e1
    double c2[3] = {e[0][1] * e[1][2] - e[0][2] * e[1][1], e[0][2] * e[1][0] - e[0][0] * e[1][2],
                    e[0][0] * e[1][1] - e[0][1] * e[1][0]};  // e1 x e2
    const double det = e[0][0] * c0[0] + e[0][1] * c0[1] + e[0][2] * c0[2], id = 1.0 / det;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      G[1][c] = c0[c] * id;
      G[2][c] = c1[c] * id;
      G[3][c] = c2[c] * id;
      G[0][c] = -(G[1][c] + G[2][c] + G[3][c]);
    }
    return fabs(det) * (1.0 / 6.0);
  }
}

template <int DIM>
__device__ __forceinline__ void fb_load_cell(const DevSpec& S, long long c, int* a, double (*X)[3]) {
#pragma unroll
  for (int j = 0; j <= DIM; ++j) {
    a[j] = S.cells[c * (DIM + 1) + j];
#pragma unroll
    for (int k = 0; k < DIM; ++k) X[j][k] = S.coords[(long long)a[j] * DIM + k];
  }
}

// circumradius of a simplex (what ufl.Circumradius evaluates to on affine cells)
template <int DIM>
__device__ __forceinline__ double fb_circumradius(const double (*X)[3]) {
  auto dist = [&](int p, int q) {
    double d = 0.0;
    for (int k = 0; k < DIM; ++k) d += (X[p][k] - X[q][k]) * (X[p][k] - X[q][k]);
    return sqrt(d);
  };
  if (DIM == 1) return 0.5 * dist(0, 1);
  if (DIM == 2) {
    const double a = dist(1, 2), b = dist(0, 2), c = dist(0, 1);
    const double area = 0.5 * fabs((X[1][0] - X[0][0]) * (X[2][1] - X[0][1]) - (X[2][0] - X[0][0]) * (X[1][1] - X[0][1]));
    return a * b * c / (4.0 * area);
  }
  // tetrahedron: edges from vertex 0 (a, b, c) times their opposite edges (A, B, C)
  const double aA = dist(0, 1) * dist(2, 3), bB = dist(0, 2) * dist(1, 3), cC = dist(0, 3) * dist(1, 2);
  const double ux = X[1][0] - X[0][0], uy = X[1][1] - X[0][1], uz = X[1][2] - X[0][2];
  const double vx = X[2][0] - X[0][0], vy = X[2][1] - X[0][1], vz = X[2][2] - X[0][2];
  const double wx = X[3][0] - X[0][0], wy = X[3][1] - X[0][1], wz = X[3][2] - X[0][2];
  const double vol = fabs(ux * (vy * wz - vz * wy) - uy * (vx * wz - vz * wx) + uz * (vx * wy - vy * wx)) / 6.0;
  return sqrt((aA + bB + cC) * (aA + bB - cC) * (aA - bB + cC) * (-aA + bB + cC)) / (24.0 * vol);
}

__device__ __forceinline__ double fb_temperature(const DevSpec& S, const int* verts, const double* bary,
                                                 int n) {
  if (!S.T_mode) return S.T_const;
  double T = 0.0;
  for (int a = 0; a < n; ++a) T += bary[a] * S.T_nodal[verts[a]];
  return T;
}

// ---------------------------------------------------------------------------
// per-cell coefficient cache: dbar[which][slot][cell] = sum_q w_q f_slot(x_q)
// f = exp(-E/(kB T)) for Arrhenius slots, a program value for generic ones
// ---------------------------------------------------------------------------
template <int DIM>
__device__ __forceinline__ void fb_cell_coeffs_body(const DevSpec& S, double* __restrict__ dbar) {
  const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (c >= S.n_cells) return;
  int a[FB_MAX_D1];
  double X[FB_MAX_D1][3], G[FB_MAX_D1][3];
  // uniform temperature and Arrhenius slots only: the cell means are the same number
  // everywhere, no geometry or nodal data needed
  const bool uniform = !S.T_mode && S.n_slots == S.n_diffE;
  const bool need_G = S.n_slots > S.n_diffE;  // generic D(x, T, ...) programs need the geometry
  if (!uniform) {
    if (need_G) {
      fb_load_cell<DIM>(S, c, a, X);
    } else {  // Arrhenius slots of a nodal temperature: vertex ids only, no coordinates
#pragma unroll
      for (int j = 0; j <= DIM; ++j) a[j] = S.cells[c * (DIM + 1) + j];
    }
  }
  const int tag = S.cell_tag[c];
  if (need_G) fb_geometry<DIM>(X, G);
  double Gflat[FB_MAX_D1 * 3];
  if (need_G)
    for (int j = 0; j <= DIM; ++j)
      for (int k = 0; k < DIM; ++k) Gflat[j * DIM + k] = G[j][k];
  double zero[FB_MAX_NS];
  for (int s = 0; s < S.ns; ++s) zero[s] = 0.0;
  for (int which = 0; which < 2; ++which) {
    const int off = S.quad_v[tag * 4 + 2 * which], nq = S.quad_v[tag * 4 + 2 * which + 1];
    for (int slot = 0; slot < S.n_slots; ++slot) {
      double acc = 0.0;
      if (slot < S.n_diffE && !S.T_mode) {
        acc = exp(-S.diff_energies[slot] / (FB_KB * S.T_const));
      } else {
        for (int q = 0; q < nq; ++q) {
          const double* bary = S.quad_bary + (long long)(off + q) * (DIM + 1);
          const double T = fb_temperature(S, a, bary, DIM + 1);
          if (slot < S.n_diffE) {
            acc += S.quad_w[off + q] * exp(-S.diff_energies[slot] / (FB_KB * T));
          } else {
            PointCtx P;
        P.cgrad = nullptr; P.n[0] = P.n[1] = P.n[2] = 0.0; P.h = 0.0;
            for (int k = 0; k < DIM; ++k) {
              double x = 0.0;
              for (int j = 0; j <= DIM; ++j) x += bary[j] * X[j][k];
              P.x[k] = x;
            }
            P.T = T; P.c = zero; P.cprev = zero; P.verts = a; P.bary = bary; P.nverts = DIM + 1;
            P.cverts = a; P.G = Gflat; P.ncverts = DIM + 1; P.tdim = DIM;
            acc += S.quad_w[off + q] * fb_prog_eval(S, S.diff_progs[slot - S.n_diffE], P);
          }
        }
      }
      dbar[((long long)which * S.n_slots + slot) * S.n_cells + c] = acc;
    }
  }
}

// ---------------------------------------------------------------------------
// load vector: solution-independent point terms, cell-parallel with fp64 atomics
// (runs once per time/temperature update, not per Newton iteration)
// ---------------------------------------------------------------------------
template <int DIM>
__device__ __forceinline__ void fb_load_cells_body(const DevSpec& S, double* __restrict__ load) {
  const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (c >= S.n_cells) return;
  const int tag = S.cell_tag[c];
  bool any = false;
  for (int t = 0; t < S.n_vterms; ++t)
    any |= (S.vterm_hdr[t * 8] == tag && !S.vterm_hdr[t * 8 + 6]);
  if (!any) return;
  int a[FB_MAX_D1];
  double X[FB_MAX_D1][3], G[FB_MAX_D1][3], Gflat[FB_MAX_D1 * 3];
  fb_load_cell<DIM>(S, c, a, X);
  const double vol = fb_geometry<DIM>(X, G);
  for (int j = 0; j <= DIM; ++j)
    for (int k = 0; k < DIM; ++k) Gflat[j * DIM + k] = G[j][k];
  double zero[FB_MAX_NS];
  for (int s = 0; s < S.ns; ++s) zero[s] = 0.0;
  const int off = S.quad_v[tag * 4], nq = S.quad_v[tag * 4 + 1];
  for (int t = 0; t < S.n_vterms; ++t) {
    const int* h = S.vterm_hdr + t * 8;
    if (h[0] != tag || h[6]) continue;
    double acc[FB_MAX_D1];
    for (int j = 0; j <= DIM; ++j) acc[j] = 0.0;
    for (int q = 0; q < nq; ++q) {
      const double* bary = S.quad_bary + (long long)(off + q) * (DIM + 1);
      PointCtx P;
        P.cgrad = nullptr; P.n[0] = P.n[1] = P.n[2] = 0.0; P.h = 0.0;
      for (int k = 0; k < DIM; ++k) {
        double x = 0.0;
        for (int j = 0; j <= DIM; ++j) x += bary[j] * X[j][k];
        P.x[k] = x;
      }
      P.T = fb_temperature(S, a, bary, DIM + 1);
      P.c = zero; P.cprev = zero; P.verts = a; P.bary = bary; P.nverts = DIM + 1;
      P.cverts = a; P.G = Gflat; P.ncverts = DIM + 1; P.tdim = DIM;
      const double val = S.quad_w[off + q] * fb_prog_eval(S, h[1], P);
      for (int j = 0; j <= DIM; ++j) acc[j] += val * bary[j];
    }
    for (int r = 0; r < h[3]; ++r) {
      const int s = S.row_species[h[2] + r];
      const double coef = S.row_coef[h[2] + r] * vol;
      for (int j = 0; j <= DIM; ++j) atomicAdd(&load[(long long)a[j] * S.ns + s], coef * acc[j]);
    }
  }
}

// measure of a boundary facet (1 for a point, length, area)
template <int DIM>
__device__ __forceinline__ double fb_facet_measure(const double (*Y)[3]) {
  if (DIM == 1) return 1.0;
  if (DIM == 2) {
    const double dx = Y[1][0] - Y[0][0], dy = Y[1][1] - Y[0][1];
    return sqrt(dx * dx + dy * dy);
  }
  const double ux = Y[1][0] - Y[0][0], uy = Y[1][1] - Y[0][1], uz = Y[1][2] - Y[0][2];
  const double vx = Y[2][0] - Y[0][0], vy = Y[2][1] - Y[0][1], vz = Y[2][2] - Y[0][2];
  const double cx = uy * vz - uz * vy, cy = uz * vx - ux * vz, cz = ux * vy - uy * vx;
  return 0.5 * sqrt(cx * cx + cy * cy + cz * cz);
}

__device__ __forceinline__ int fb_find_slot(const DevSpec& S, int row, int col) {
  int lo = S.rowptr[row], hi = S.rowptr[row + 1] - 1;
  const int base = lo;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (S.colidx[mid] < col) lo = mid + 1; else hi = mid;
  }
  return lo - base;
}

__device__ __forceinline__ long long fb_entry(const DevSpec& S, int row, int s, int k, int j) {
  // block layout: the npairs values of block (row, k) are contiguous
  return ((long long)S.rowptr[row] + k) * S.npairs + S.pair_pp[s] + j;
}

// Facet terms.  mode 0: solution-independent terms -> load vector.
// mode 1: solution-dependent terms -> F (and J when J != nullptr), honouring
// Dirichlet rows/columns; runs after the gather kernel.
template <int DIM>
__device__ __forceinline__ void fb_facets_body(const DevSpec& S, int mode, const double* __restrict__ u,
                                               const double* __restrict__ un, double* __restrict__ F,
                                               double* __restrict__ J) {
  const long long f = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (f >= S.n_bfacets) return;
  const int tag = S.bfacet_tag[f];
  if (tag < 0) return;
  bool any = false;
  for (int t = 0; t < S.n_fterms; ++t)
    any |= (S.fterm_hdr[t * 8] == tag && (S.fterm_hdr[t * 8 + 6] != 0) == (mode == 1));
  if (!any) return;
  const int ns = S.ns;
  int fv[3];
  double Y[3][3];
  for (int j = 0; j < DIM; ++j) {
    fv[j] = S.bfacets[f * DIM + j];
    for (int k = 0; k < DIM; ++k) Y[j][k] = S.coords[(long long)fv[j] * DIM + k];
  }
  const double area = fb_facet_measure<DIM>(Y);
  const long long c = S.bfacet_cell[f];
  int a[FB_MAX_D1];
  double X[FB_MAX_D1][3], G[FB_MAX_D1][3], Gflat[FB_MAX_D1 * 3];
  fb_load_cell<DIM>(S, c, a, X);
  fb_geometry<DIM>(X, G);
  for (int j = 0; j <= DIM; ++j)
    for (int k = 0; k < DIM; ++k) Gflat[j * DIM + k] = G[j][k];
  const int offF = S.quad_s[tag * 4], nqF = S.quad_s[tag * 4 + 1];
  const int offJ = S.quad_s[tag * 4 + 2], nqJ = S.quad_s[tag * 4 + 3];
  // outward unit normal (minus the direction of the gradient of the basis function of the
  // vertex opposite to the facet), circumradius of the cell, n . grad(phi_i) of every cell
  // vertex, and - for solution-dependent terms - the cell gradient of every species
  double nrm[3] = {0.0, 0.0, 0.0}, ng[FB_MAX_D1], cgr[FB_MAX_NS * 3];
  {
    int opp = 0;
    for (int j = 0; j <= DIM; ++j) {
      bool onf = false;
      for (int k = 0; k < DIM; ++k) onf |= (a[j] == fv[k]);
      if (!onf) opp = j;
    }
    double l = 0.0;
    for (int k = 0; k < DIM; ++k) l += G[opp][k] * G[opp][k];
    l = sqrt(l);
    for (int k = 0; k < DIM; ++k) nrm[k] = -G[opp][k] / l;
    for (int j = 0; j <= DIM; ++j) {
      double d = 0.0;
      for (int k = 0; k < DIM; ++k) d += nrm[k] * G[j][k];
      ng[j] = d;
    }
  }
  const double hcell = fb_circumradius<DIM>(X);
  for (int s = 0; s < ns; ++s)
    for (int k = 0; k < 3; ++k) {
      double gk = 0.0;
      if (mode == 1 && k < DIM)
        for (int j = 0; j <= DIM; ++j) gk += u[(long long)a[j] * ns + s] * G[j][k];
      cgr[s * 3 + k] = gk;
    }
  double cq[FB_MAX_NS], cpq[FB_MAX_NS];
  for (int t = 0; t < S.n_fterms; ++t) {
    const int* h = S.fterm_hdr + t * 8;
    if (h[0] != tag || (h[6] != 0) != (mode == 1)) continue;
    // rows: the facet vertices weighted by phi_i (test function v), or every vertex of the
    // cell weighted by n . grad(phi_i) (test function entering through n . grad(v))
    const bool gradtest = h[7] == 1;
    const int nrow = gradtest ? DIM + 1 : DIM;
    // value
    double acc[FB_MAX_D1] = {0.0, 0.0, 0.0, 0.0};
    for (int q = 0; q < (F != nullptr ? nqF : 0); ++q) {
      const double* bary = S.quad_bary + (long long)(offF + q) * (DIM + 1);  // uniform pool stride; DIM entries used
      PointCtx P;
      for (int k = 0; k < DIM; ++k) {
        double x = 0.0;
        for (int j = 0; j < DIM; ++j) x += bary[j] * Y[j][k];
        P.x[k] = x;
      }
      P.T = fb_temperature(S, fv, bary, DIM);
      for (int s = 0; s < ns; ++s) {
        double v = 0.0, vp = 0.0;
        if (mode == 1)
          for (int j = 0; j < DIM; ++j) {
            v += bary[j] * u[(long long)fv[j] * ns + s];
            vp += bary[j] * un[(long long)fv[j] * ns + s];
          }
        cq[s] = v; cpq[s] = vp;
      }
      P.c = cq; P.cprev = cpq; P.verts = fv; P.bary = bary; P.nverts = DIM;
      P.cverts = a; P.G = Gflat; P.ncverts = DIM + 1; P.tdim = DIM;
      P.cgrad = cgr; P.n[0] = nrm[0]; P.n[1] = nrm[1]; P.n[2] = nrm[2]; P.h = hcell;
      const double val = S.quad_w[offF + q] * fb_prog_eval(S, h[1], P);
      for (int i = 0; i < nrow; ++i) acc[i] += val * (gradtest ? ng[i] : bary[i]);
    }
    if (F != nullptr)
      for (int r = 0; r < h[3]; ++r) {
        const int s = S.row_species[h[2] + r];
        const double coef = S.row_coef[h[2] + r] * area;
        for (int i = 0; i < nrow; ++i) {
          const long long dof = (long long)(gradtest ? a[i] : fv[i]) * ns + s;
          if (mode == 1 && S.bc_map[dof] >= 0) continue;
          atomicAdd(&F[dof], coef * acc[i]);
        }
      }
    if (mode != 1 || h[5] == 0) continue;
    // derivatives: J entries and Dirichlet lifting.  Columns: the facet vertices weighted by
    // phi_j (dependence on the value of a species at the point) or every vertex of the cell
    // weighted by d_k phi_j (dependence on component k of its cell gradient)
    for (int dd = 0; dd < h[5]; ++dd) {
      const int code = S.deriv_species[h[4] + dd];
      const int dep = code % ns, gcomp = code / ns - 1;
      const int ncol = gcomp < 0 ? DIM : DIM + 1;
      double M[FB_MAX_D1][FB_MAX_D1];
      for (int i = 0; i < FB_MAX_D1; ++i)
        for (int j = 0; j < FB_MAX_D1; ++j) M[i][j] = 0.0;
      for (int q = 0; q < nqJ; ++q) {
        const double* bary = S.quad_bary + (long long)(offJ + q) * (DIM + 1);
        PointCtx P;
        for (int k = 0; k < DIM; ++k) {
          double x = 0.0;
          for (int j = 0; j < DIM; ++j) x += bary[j] * Y[j][k];
          P.x[k] = x;
        }
        P.T = fb_temperature(S, fv, bary, DIM);
        for (int s = 0; s < ns; ++s) {
          double v = 0.0, vp = 0.0;
          for (int j = 0; j < DIM; ++j) {
            v += bary[j] * u[(long long)fv[j] * ns + s];
            vp += bary[j] * un[(long long)fv[j] * ns + s];
          }
          cq[s] = v; cpq[s] = vp;
        }
        P.c = cq; P.cprev = cpq; P.verts = fv; P.bary = bary; P.nverts = DIM;
        P.cverts = a; P.G = Gflat; P.ncverts = DIM + 1; P.tdim = DIM;
        P.cgrad = cgr; P.n[0] = nrm[0]; P.n[1] = nrm[1]; P.n[2] = nrm[2]; P.h = hcell;
        const double dv = S.quad_w[offJ + q] * fb_prog_eval(S, S.deriv_prog[h[4] + dd], P);
        for (int i = 0; i < nrow; ++i) {
          const double wi = dv * (gradtest ? ng[i] : bary[i]);
          for (int j = 0; j < ncol; ++j) M[i][j] += wi * (gcomp < 0 ? bary[j] : G[j][gcomp]);
        }
      }
      for (int r = 0; r < h[3]; ++r) {
        const int s = S.row_species[h[2] + r];
        const double coef = S.row_coef[h[2] + r] * area;
        const int jp = S.pair_idx[s * ns + dep];
        for (int i = 0; i < nrow; ++i) {
          const int vi = gradtest ? a[i] : fv[i];
          const long long rdof = (long long)vi * ns + s;
          if (S.bc_map[rdof] >= 0) continue;
          for (int j = 0; j < ncol; ++j) {
            const int vj = gcomp < 0 ? fv[j] : a[j];
            const long long cdof = (long long)vj * ns + dep;
            const int bcj = S.bc_map[cdof];
            const double aij = coef * M[i][j];
            if (bcj >= 0) {
              if (F != nullptr) atomicAdd(&F[rdof], aij * (S.bc_vals[bcj] - u[cdof]));
            } else if (J != nullptr) {
              atomicAdd(&J[fb_entry(S, vi, s, fb_find_slot(S, vi, vj), jp)], aij);
            }
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// the gather kernel: one thread per vertex
// ---------------------------------------------------------------------------
template <int DIM>
__device__ __forceinline__ void fb_assemble_nodes_body(const DevSpec& S, const double* __restrict__ u,
                                                       const double* __restrict__ un,
                                                       double* __restrict__ F, double* __restrict__ J,
                                                       int flags) {
  constexpr int D1 = DIM + 1;
  const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (v >= S.n_owned) return;  // ghost rows belong to another rank
  const int ns = S.ns;
  const bool doF = (flags & FB_ASSEMBLE_F) != 0, doJ = (flags & FB_ASSEMBLE_J) != 0;
  const int r0 = S.rowptr[v], deg = S.rowptr[v + 1] - r0;
  const long long rowbase = (long long)r0 * S.npairs;

  double Facc[FB_MAX_NS];
  unsigned rowbc = 0;
  for (int s = 0; s < ns; ++s) {
    Facc[s] = 0.0;
    if (S.bc_map[v * ns + s] >= 0) rowbc |= 1u << s;
  }
  if (doJ)
    for (long long e = 0; e < (long long)deg * S.npairs; ++e) J[rowbase + e] = 0.0;

  const double mscale = 1.0 / (double)(D1 * (D1 + 1));
  // The chain  v2c_idx -> cells -> coords  is three dependent global loads per cell and a
  // thread issues in order, so it is software-pipelined by hand: the cell index is fetched
  // two iterations ahead and the cell's vertex ids one iteration ahead; only the coordinate
  // loads of the current cell remain on the critical path.
  const int e_begin = S.v2c_ptr[v], e_end = S.v2c_ptr[v + 1];
  int pk_cur = e_begin < e_end ? S.v2c_idx[e_begin] : 0;
  int pk_nxt = e_begin + 1 < e_end ? S.v2c_idx[e_begin + 1] : pk_cur;
  int a_cur[FB_MAX_D1];
#pragma unroll
  for (int j = 0; j < D1; ++j) a_cur[j] = S.cells[(long long)(pk_cur / D1) * D1 + j];
  for (int e = e_begin; e < e_end; ++e) {
    const int packed = pk_cur;
    const long long c = packed / D1;
    const int i = packed - (int)c * D1;
    int a[FB_MAX_D1];
    double X[FB_MAX_D1][3], G[FB_MAX_D1][3];
#pragma unroll
    for (int j = 0; j < D1; ++j) {
      a[j] = a_cur[j];
#pragma unroll
      for (int k = 0; k < DIM; ++k) X[j][k] = S.coords[(long long)a[j] * DIM + k];
    }
    {  // prefetch for the next two iterations
      const int pk_far = e + 2 < e_end ? S.v2c_idx[e + 2] : pk_nxt;
      const long long c_nxt = pk_nxt / D1;
#pragma unroll
      for (int j = 0; j < D1; ++j) a_cur[j] = S.cells[c_nxt * D1 + j];
      pk_cur = pk_nxt;
      pk_nxt = pk_far;
    }
    const double vol = fb_geometry<DIM>(X, G);
    const int tag = S.cell_tag[c];
    int slot[FB_MAX_D1];
    if (S.v2c_slot != nullptr) {
      const unsigned ps = S.v2c_slot[e];
#pragma unroll
      for (int j = 0; j < D1; ++j) slot[j] = (ps >> (8 * j)) & 0xffu;
    } else {
#pragma unroll
      for (int j = 0; j < D1; ++j) slot[j] = fb_find_slot(S, (int)v, a[j]);
    }

    // ---- diffusion + mass + advection (species-diagonal, closed form) ----
    for (int s = 0; s < ns; ++s) {
      if (rowbc >> s & 1u) continue;
      const int ds = S.diff_slot[tag * ns + s];
      double kF = 0.0, kJ = 0.0;
      if (ds >= 0) {
        const double D0 = S.diff_D0[tag * ns + s] * vol;
        kF = D0 * S.dbar[(long long)ds * S.n_cells + c];
        kJ = D0 * S.dbar[((long long)S.n_slots + ds) * S.n_cells + c];
      }
      if (S.n_udiff > 0 && S.diff_uprog[(tag * ns + s) * 3] >= 0) {
        // solution-dependent diffusivity D(u): F uses mean(D) with the F rule; J adds
        // mean(D) G_i.G_j with the J rule plus (dD/du_dep)(q) phi_j(q) G_i.grad(u_s)
        const int* up = S.diff_uprog + (tag * ns + s) * 3;
        double ucl[FB_MAX_D1][FB_MAX_NS], uncl[FB_MAX_D1][FB_MAX_NS], Gf[FB_MAX_D1 * 3];
        for (int j = 0; j < D1; ++j) {
          for (int k = 0; k < DIM; ++k) Gf[j * DIM + k] = G[j][k];
          for (int s2 = 0; s2 < ns; ++s2) {
            ucl[j][s2] = u[(long long)a[j] * ns + s2];
            uncl[j][s2] = S.transient ? un[(long long)a[j] * ns + s2] : 0.0;
          }
        }
        double giu = 0.0;  // G_i . grad(u_s)
        for (int k = 0; k < DIM; ++k) {
          double gk = 0.0;
          for (int j = 0; j < D1; ++j) gk += ucl[j][s] * G[j][k];
          giu += G[i][k] * gk;
        }
        double cqs[FB_MAX_NS], cpqs[FB_MAX_NS];
        for (int pass = 0; pass < 2; ++pass) {
          if (pass == 0 && !doF) continue;
          const int off = S.quad_v[tag * 4 + 2 * pass], nq = S.quad_v[tag * 4 + 2 * pass + 1];
          double Dm = 0.0;
          for (int q = 0; q < nq; ++q) {
            const double* bary = S.quad_bary + (long long)(off + q) * D1;
            PointCtx P;
        P.cgrad = nullptr; P.n[0] = P.n[1] = P.n[2] = 0.0; P.h = 0.0;
            for (int k = 0; k < DIM; ++k) {
              double x = 0.0;
              for (int j = 0; j < D1; ++j) x += bary[j] * X[j][k];
              P.x[k] = x;
            }
            P.T = fb_temperature(S, a, bary, D1);
            for (int s2 = 0; s2 < ns; ++s2) {
              double cv = 0.0, cp = 0.0;
              for (int j = 0; j < D1; ++j) { cv += bary[j] * ucl[j][s2]; cp += bary[j] * uncl[j][s2]; }
              cqs[s2] = cv; cpqs[s2] = cp;
            }
            P.c = cqs; P.cprev = cpqs; P.verts = a; P.bary = bary; P.nverts = D1;
            P.cverts = a; P.G = Gf; P.ncverts = D1; P.tdim = DIM;
            const double wq = S.quad_w[off + q];
            Dm += wq * fb_prog_eval(S, up[0], P);
            if (pass == 1) {
              for (int dd = 0; dd < up[2]; ++dd) {
                const int dep = S.deriv_species[up[1] + dd];
                const double dv = wq * vol * giu * fb_prog_eval(S, S.deriv_prog[up[1] + dd], P);
                const int jp2 = S.pair_idx[s * ns + dep];
                for (int j = 0; j < D1; ++j) {
                  const double aij = dv * bary[j];
                  const int bcj = S.bc_map[(long long)a[j] * ns + dep];
                  if (bcj >= 0) {
                    if (doF) Facc[s] += aij * (S.bc_vals[bcj] - ucl[j][dep]);
                  } else if (doJ) {
                    J[rowbase + (long long)slot[j] * S.npairs + S.pair_pp[s] + jp2] += aij;
                  }
                }
              }
            }
          }
          if (pass == 0) kF += vol * Dm; else kJ += vol * Dm;
        }
      }
      const double m = S.transient ? S.mass_dt[tag * ns + s] * S.inv_dt * vol * mscale : 0.0;
      const double mc = S.mass_c[tag * ns + s] * vol * mscale;
      // advection: sum_q w_q phi_i(q) vel_q . G_j = sum_a (vel_a . G_j) int(phi_i phi_a)
      double advj[FB_MAX_D1];
#pragma unroll
      for (int j = 0; j < D1; ++j) advj[j] = 0.0;
      for (int t = 0; t < S.n_adv; ++t) {
        if (S.adv[2 * t] != tag || S.adv[2 * t + 1] != s) continue;
        const double* vel = S.velocities[t];
#pragma unroll
        for (int j = 0; j < D1; ++j) {
          double acc = 0.0;
#pragma unroll
          for (int b = 0; b < D1; ++b) {
            double vg = 0.0;
#pragma unroll
            for (int k = 0; k < DIM; ++k) vg += vel[(long long)a[b] * DIM + k] * G[j][k];
            acc += vg * (b == i ? 2.0 : 1.0);
          }
          advj[j] += acc * vol * mscale;
        }
      }
      const int jp = S.pair_idx[s * ns + s];
#pragma unroll
      for (int j = 0; j < D1; ++j) {
        double gg = 0.0;
#pragma unroll
        for (int k = 0; k < DIM; ++k) gg += G[i][k] * G[j][k];
        const double Mij = (i == j) ? 2.0 : 1.0;
        const long long cdof = (long long)a[j] * ns + s;
        const double uj = u[cdof];
        const double aJ = kJ * gg + (m + mc) * Mij + advj[j];
        if (doF) {
          double f = (kF * gg + mc * Mij + advj[j]) * uj;
          if (S.transient) f += m * Mij * (uj - un[cdof]);
          Facc[s] += f;
        }
        const int bcj = S.bc_map[cdof];
        if (bcj >= 0) {
          if (doF) Facc[s] += aJ * (S.bc_vals[bcj] - uj);
        } else if (doJ) {
          J[rowbase + (long long)slot[j] * S.npairs + S.pair_pp[s] + jp] += aJ;
        }
      }
    }

    // ---- solution-dependent point terms (reactions, nonlinear sources) ----
    if (S.n_vterms == 0) continue;
    bool any = false;
    for (int t = 0; t < S.n_vterms; ++t) any |= (S.vterm_hdr[t * 8] == tag && S.vterm_hdr[t * 8 + 6]);
    if (!any) continue;
    double uc[FB_MAX_D1][FB_MAX_NS], unc[FB_MAX_D1][FB_MAX_NS];
    bool lift = false;
    for (int j = 0; j < D1; ++j)
      for (int s = 0; s < ns; ++s) {
        const long long dof = (long long)a[j] * ns + s;
        uc[j][s] = u[dof];
        unc[j][s] = S.transient ? un[dof] : 0.0;
        const int b = S.bc_map[dof];
        if (b >= 0 && S.bc_vals[b] != uc[j][s]) lift = true;
      }
    double Gflat[FB_MAX_D1 * 3];
    for (int j = 0; j < D1; ++j)
      for (int k = 0; k < DIM; ++k) Gflat[j * DIM + k] = G[j][k];
    double cq[FB_MAX_NS], cpq[FB_MAX_NS];
    double cgr[FB_MAX_NS * 3];  // cell gradient of every species (constant on a P1 cell)
    for (int s = 0; s < ns; ++s)
      for (int k = 0; k < 3; ++k) {
        double gk = 0.0;
        if (k < DIM)
          for (int j = 0; j < D1; ++j) gk += uc[j][s] * G[j][k];
        cgr[s * 3 + k] = gk;
      }
    for (int pass = 0; pass < 2; ++pass) {
      // pass 0: residual with the F rule; pass 1: derivatives with the J rule
      if (pass == 0 && !doF) continue;
      if (pass == 1 && !(doJ || (doF && lift))) continue;
      const int off = S.quad_v[tag * 4 + 2 * pass], nq = S.quad_v[tag * 4 + 2 * pass + 1];
      for (int q = 0; q < nq; ++q) {
        const double* bary = S.quad_bary + (long long)(off + q) * D1;
        const double wq = S.quad_w[off + q] * vol * bary[i];
        PointCtx P;
        P.cgrad = nullptr; P.n[0] = P.n[1] = P.n[2] = 0.0; P.h = 0.0;
        for (int k = 0; k < DIM; ++k) {
          double x = 0.0;
          for (int j = 0; j < D1; ++j) x += bary[j] * X[j][k];
          P.x[k] = x;
        }
        P.T = fb_temperature(S, a, bary, D1);
        for (int s = 0; s < ns; ++s) {
          double cv = 0.0, cp = 0.0;
          for (int j = 0; j < D1; ++j) { cv += bary[j] * uc[j][s]; cp += bary[j] * unc[j][s]; }
          cq[s] = cv; cpq[s] = cp;
        }
        P.c = cq; P.cprev = cpq; P.verts = a; P.bary = bary; P.nverts = D1;
        P.cverts = a; P.G = Gflat; P.ncverts = D1; P.tdim = DIM; P.cgrad = cgr;
        for (int t = 0; t < S.n_vterms; ++t) {
          const int* h = S.vterm_hdr + t * 8;
          if (h[0] != tag || !h[6]) continue;
          if (pass == 0) {
            const double val = wq * fb_prog_eval(S, h[1], P);
            for (int r = 0; r < h[3]; ++r) {
              const int s = S.row_species[h[2] + r];
              if (!(rowbc >> s & 1u)) Facc[s] += S.row_coef[h[2] + r] * val;
            }
          } else {
            for (int dd = 0; dd < h[5]; ++dd) {
              // code = species (derivative w.r.t. the value at the point, weight phi_j) or
              // species + ns * (1 + k) (w.r.t. component k of its cell gradient, weight d_k phi_j)
              const int code = S.deriv_species[h[4] + dd];
              const int dep = code % ns, gcomp = code / ns - 1;
              const double dv = wq * fb_prog_eval(S, S.deriv_prog[h[4] + dd], P);
              for (int r = 0; r < h[3]; ++r) {
                const int s = S.row_species[h[2] + r];
                if (rowbc >> s & 1u) continue;
                const double coef = S.row_coef[h[2] + r] * dv;
                const int jp = S.pair_idx[s * ns + dep];
                for (int j = 0; j < D1; ++j) {
                  const double aij = coef * (gcomp < 0 ? bary[j] : G[j][gcomp]);
                  const int bcj = S.bc_map[(long long)a[j] * ns + dep];
                  if (bcj >= 0) {
                    if (doF) Facc[s] += aij * (S.bc_vals[bcj] - uc[j][dep]);
                  } else if (doJ) {
                    J[rowbase + (long long)slot[j] * S.npairs + S.pair_pp[s] + jp] += aij;
                  }
                }
              }
            }
          }
        }
      }
    }
  }

  // ---- finalise the rows of this vertex ----
  const int kself = fb_find_slot(S, (int)v, (int)v);
  for (int s = 0; s < ns; ++s) {
    const long long dof = v * ns + s;
    if (rowbc >> s & 1u) {
      if (doF) F[dof] = u[dof] - S.bc_vals[S.bc_map[dof]];
      if (doJ)
        J[rowbase + (long long)kself * S.npairs + S.pair_pp[s] + S.pair_idx[s * ns + s]] = 1.0;
    } else if (doF) {
      F[dof] = Facc[s] + S.load[dof];
    }
  }
}

