// Newton-iteration assembly on the channel operator, TMA-pipelined (sm_100a): the kernel of
// assemble_edge.cuh (k_assemble_edge: residual + solution-dependent channel values of every
// block row, strong Dirichlet rows / lifting applied) restructured like the channel SpMV
// (spmv_channels.cuh).  One persistent block per SM walks chunks of FB_CHUNK consecutive block
// rows; a producer warp brings everything a chunk reads from HBM into a ring of shared-memory
// stages with 1D bulk copies (cp.async.bulk) completing on one mbarrier per stage - the static
// channel values, the edge tensors (T_vvw, T_vww, the T_vwz slabs and their slot bytes), u and
// u_n at the chunk's distinct neighbour vertices (one copy per run of consecutive ids), the load
// vector, local column indices, Dirichlet words and row pointers - while FB_CHUNK_GROUPS groups of
// consumer warps (one warp per row) work on earlier chunks: affine reaction factors at the row's
// neighbours, dynamic channel values  A^c_vw = sum_z T_vwz f_c(z)  (written to the channel
// operator, coalesced), then the residual  F_s = load + sum_e coef_e sum_w A^{c_e}_vw g_e(w).
// Replaces the FFCx kernels + dolfinx assemble_vector / apply_lifting / set_bc behind
// NonlinearProblem (reference src/festim/problem.py:170-182) for the form of
// src/festim/hydrogen_transport_problem.py:870-997.
#pragma once
#include "assemble_edge.cuh"
#include "spmv_channels.cuh"

struct AsmTmaArgs {
  const double* u;
  const double* un;
  double* F;
  double* dyn;
  int flags;
  int nst;
  const int* mismatch;   // reuse mode when non-null and zero: et_dyn already holds the channels of this solution
  // shared-memory layout (bytes)
  int o_tab_i, o_fac_d, o_fac_i, o_scr, scr_bytes, o_stage0, stage_bytes, o_bar;
  int s_tvw, s_tz, s_tzs, s_us, s_uns, s_load, s_lcol, s_vinfo, s_rowinfo, s_rowptr, s_tzptr, s_tzw;
  int tz_stride;   // doubles per tensor inside a stage
  int w_fv, w_adyn, w_res, w_tf, rts;   // offsets (doubles) inside a warp's scratch; row stride of the R tile
};

__host__ inline size_t fb_asm_tma_layout(const DevSpec& S, AsmTmaArgs& a, int nst) {
  auto up16 = [](size_t b) { return (b + 15) & ~(size_t)15; };
  const int md = S.maxdeg, ns = S.ns, CH = FB_CHUNK, nts = S.vx_nts;
  size_t o = 0;
  o += up16((size_t)S.vx_dtab_n * 8);
  a.o_tab_i = (int)o; o += up16((size_t)S.vx_itab_n * 4);
  // affine reaction factors: constant term and coefficients (doubles), then offset / count / species (ints)
  a.o_fac_d = (int)o; o += up16((size_t)(S.n_factors + S.n_fac_terms + 1) * 8);
  a.o_fac_i = (int)o; o += up16((size_t)(2 * S.n_factors + S.n_fac_terms + 1) * 4);
  int w = 0;
  a.w_fv = w; w += md * (S.n_factors > 0 ? S.n_factors : 1);
  a.w_adyn = w; w += md * (S.vx_ndp > 0 ? S.vx_ndp : 1);
  const int nres = S.vx_nft > S.vx_njt ? S.vx_nft : S.vx_njt;
  a.w_res = w; w += nres > 0 ? nres : 1;
  // expanded tensor row (16 x 17), later the residual products R[channel][table column] (16 x rts)
  a.rts = 8 * ((S.vx_nbw + 7) / 8) + 1;
  a.w_tf = w; w += 16 * (a.rts > 17 ? a.rts : 17);
  a.o_scr = (int)o; a.scr_bytes = (int)up16((size_t)w * 8); o += (size_t)CH * FB_CHUNK_GROUPS * a.scr_bytes;
  size_t s = 0;
  s += up16((size_t)CH * md * S.vx_nsp * 8);                       // static channel values
  a.s_tvw = (int)s; s += up16((size_t)nts * CH * md * 16);
  a.tz_stride = (int)((S.et_tz_chunk_max + 1) & ~1LL);
  {   // the T_vwz slabs - or, in reuse mode, the stored solution-dependent channel values of the chunk
    const size_t tzb = (size_t)nts * a.tz_stride * 8, dyb = (size_t)CH * md * S.vx_ndp * 8;
    a.s_tz = (int)s; s += up16(tzb > dyb ? tzb : dyb);
  }
  a.s_tzs = (int)s; s += up16((size_t)S.et_tz_chunk_max + 16);
  a.s_us = (int)s; s += up16((size_t)(S.ck_nx_max + 2) * ns * 8);
  a.s_uns = (int)s; if (S.transient) s += up16((size_t)(S.ck_nx_max + 2) * ns * 8);
  a.s_load = (int)s; s += up16((size_t)CH * ns * 8);
  a.s_lcol = (int)s; s += up16((size_t)(CH * md + 8) * 2);
  a.s_vinfo = (int)s; s += up16((size_t)CH * 4);
  a.s_rowinfo = (int)s; s += up16((size_t)CH * 4);
  a.s_rowptr = (int)s; s += up16((size_t)(CH + 4) * 4);
  a.s_tzptr = (int)s; s += up16((size_t)(CH + 4) * 8);
  a.s_tzw = (int)s; s += up16((size_t)CH * 4);
  a.stage_bytes = (int)s;
  a.o_stage0 = (int)o; o += (size_t)nst * s;
  a.o_bar = (int)o; o += 16 * 8;
  a.nst = nst;
  return o;
}

__global__ void __launch_bounds__(32 * FB_CHUNK * FB_CHUNK_GROUPS + 64, 1)
k_assemble_edge_tma(DevSpec S, AsmTmaArgs a) {
  extern __shared__ __align__(16) unsigned char fb_as_smem[];
  constexpr int CH = FB_CHUNK, NG = FB_CHUNK_GROUPS;
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ns = S.ns, NSP = S.vx_nsp, NDP = S.vx_ndp, NST = a.nst, nts = S.vx_nts, NF = S.n_factors;
  const int nstat = S.vx_nstat, ndyn = S.vx_ndyn;
  const bool doF = (a.flags & FB_ASSEMBLE_F) != 0;
  const bool transient = S.transient != 0;
  const bool reuse = a.mismatch && a.mismatch[0] == 0 && ndyn > 0;
  double* dtab = (double*)fb_as_smem;
  int* itab = (int*)(fb_as_smem + a.o_tab_i);
  for (int t = threadIdx.x; t < S.vx_dtab_n; t += blockDim.x) dtab[t] = S.vx_dtab[t];
  for (int t = threadIdx.x; t < S.vx_itab_n; t += blockDim.x) itab[t] = S.vx_itab[t];
  double* fac_d = (double*)(fb_as_smem + a.o_fac_d);   // [NF] constant terms, then the coefficients
  int* fac_i = (int*)(fb_as_smem + a.o_fac_i);         // [NF][2] offset, count; then the species
  for (int t = threadIdx.x; t < NF; t += blockDim.x) {
    const int* h = S.fac_hdr + 4 * t;
    fac_d[t] = h[0] ? S.scalars[h[1]] : S.fac_a0[t];
    fac_i[2 * t] = h[2]; fac_i[2 * t + 1] = h[3];
  }
  for (int t = threadIdx.x; t < S.n_fac_terms; t += blockDim.x) {
    fac_d[NF + t] = S.fac_coef[t];
    fac_i[2 * NF + t] = S.fac_species[t];
  }
  unsigned long long* full = (unsigned long long*)(fb_as_smem + a.o_bar);
  unsigned long long* empty = full + 8;
  if (threadIdx.x == 0)
    for (int q = 0; q < NST; ++q) { fb_mbar_init(&full[q], 2); fb_mbar_init(&empty[q], CH); }   // full: one arrival per producer warp
  for (size_t t = threadIdx.x; t < (size_t)NST * a.stage_bytes / 16; t += blockDim.x)
    ((double2*)(fb_as_smem + a.o_stage0))[t] = make_double2(0.0, 0.0);
  fb_fence_proxy_async();
  fb_mbar_fence_init();
  __syncthreads();
  const int nchunks = (int)((S.n_owned + CH - 1) / CH);
  const int nmine = (int)blockIdx.x < nchunks ? (nchunks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp >= CH * NG) {
    // ================= two producer warps: warp 0 the row data, warp 1 the runs of u / u_n ================
    const bool px = warp > CH * NG;
    int pf_e0 = 0, pf_ne = 0, pf_nx = 0, pf_lc = 0, pf_q0 = 0, pf_q1 = 0, pf_run[3] = {0, 0, 0};
    long long pf_t0 = 0, pf_t1 = 0;
    auto prefetch = [&](int i) {
      if (i >= nmine) return;
      const long long c = blockIdx.x + i * (long long)gridDim.x;
      const long long c0 = c * CH;
      if (!px) {
        const int nrow = (int)((S.n_owned - c0) < CH ? (S.n_owned - c0) : CH);
        pf_e0 = S.rowptr[c0]; pf_ne = S.rowptr[c0 + nrow]; pf_lc = S.ck_lcolptr[c];
        pf_t0 = nts ? S.et_tzptr[c0] : 0; pf_t1 = nts ? S.et_tzptr[c0 + nrow] : 0;
      } else {
        pf_nx = S.ck_nx[c];
        const int* rr = S.ck_slots + ((size_t)c * 32 + lane) * 3;
        pf_run[0] = rr[0]; pf_run[1] = rr[1]; pf_run[2] = rr[2];
        if (S.ck_rmax > 32) { pf_q0 = S.ck_runptr[c]; pf_q1 = S.ck_runptr[c + 1]; }
      }
    };
    prefetch(0);
    int sidx = 0;
    unsigned eph = 1;   // parity of the stage's previous use: the first NST waits pass at once
    for (int i = 0; i < nmine; ++i, sidx = (sidx + 1 == NST ? 0 : sidx + 1)) {
      if (i && sidx == 0) eph ^= 1u;
      fb_mbar_wait(&empty[sidx], eph);
      const long long c = blockIdx.x + i * (long long)gridDim.x;
      const long long c0 = c * CH;
      const int nrow = (int)((S.n_owned - c0) < CH ? (S.n_owned - c0) : CH);
      unsigned char* st = fb_as_smem + a.o_stage0 + (size_t)sidx * a.stage_bytes;
      unsigned long long* b = &full[sidx];
      const int e0 = pf_e0, ne = pf_ne - pf_e0, nx = pf_nx, lc0 = pf_lc, q0 = pf_q0, q1 = pf_q1;
      const long long t0 = pf_t0;
      const int nt = (int)(pf_t1 - pf_t0);   // entries of the chunk's T_vwz slabs (multiple of 16)
      const int r0v = pf_run[0], r1v = pf_run[1], r2v = pf_run[2];
      prefetch(i + 1);
      const int nlc = (ne + 7) & ~7;
      if (px) {
        // u (and u_n) at the chunk's distinct neighbours, one copy per run of vertex ids
        if (lane == 0) fb_mbar_expect_tx(b, (unsigned)(nx * ns * 8 * (transient ? 2 : 1)));
        __syncwarp();
        if (r1v > 0) {
          fb_bulk_g2s(st + a.s_us + (size_t)r2v * ns * 8, a.u + (long long)r0v * ns, (unsigned)(r1v * ns * 8), b);
          if (transient) fb_bulk_g2s(st + a.s_uns + (size_t)r2v * ns * 8, a.un + (long long)r0v * ns, (unsigned)(r1v * ns * 8), b);
        }
        for (int q = q0 + 32 + lane; q < q1; q += 32) {
          const int v0 = S.ck_runs[3 * q], len = S.ck_runs[3 * q + 1], lo = S.ck_runs[3 * q + 2];
          fb_bulk_g2s(st + a.s_us + (size_t)lo * ns * 8, a.u + (long long)v0 * ns, (unsigned)(len * ns * 8), b);
          if (transient) fb_bulk_g2s(st + a.s_uns + (size_t)lo * ns * 8, a.un + (long long)v0 * ns, (unsigned)(len * ns * 8), b);
        }
        continue;
      }
      if (lane == 0) {
        unsigned bytes = (unsigned)(nlc * 2 + 2 * CH * 4 + (CH + 4) * 4);
        if (nts && !reuse) bytes += (unsigned)(nts * (2 * ne * 8 + nt * 8) + nt + CH * 4 + (CH + 4) * 8);
        if (reuse) bytes += (unsigned)(ne * NDP * 8);
        if (doF) bytes += (unsigned)(ne * NSP * 8 + nrow * ns * 8);
        fb_mbar_expect_tx(b, bytes);
      }
      __syncwarp();
      if (reuse && lane == 6) fb_bulk_g2s(st + a.s_tz, a.dyn + (long long)e0 * NDP, (unsigned)(ne * NDP * 8), b);
      if (lane == 0 && doF && NSP) fb_bulk_g2s(st, S.et_stat + (long long)e0 * NSP, (unsigned)(ne * NSP * 8), b);
      if (lane == 1 && doF) fb_bulk_g2s(st + a.s_load, S.load + c0 * ns, (unsigned)(nrow * ns * 8), b);
      if (lane == 2) fb_bulk_g2s(st + a.s_lcol, S.ck_lcol + lc0, (unsigned)(nlc * 2), b);
      if (lane == 3) fb_bulk_g2s(st + a.s_vinfo, S.vinfo + c0, (unsigned)(CH * 4), b);
      if (lane == 4) fb_bulk_g2s(st + a.s_rowinfo, S.ck_rowinfo + c0, (unsigned)(CH * 4), b);
      if (lane == 5) fb_bulk_g2s(st + a.s_rowptr, S.rowptr + c0, (unsigned)((CH + 4) * 4), b);
      if (lane == 6 && nts && !reuse) fb_bulk_g2s(st + a.s_tzptr, S.et_tzptr + c0, (unsigned)((CH + 4) * 8), b);
      if (lane == 7 && nts && !reuse) fb_bulk_g2s(st + a.s_tzw, S.et_tzw + c0, (unsigned)(CH * 4), b);
      if (lane == 8 && nts && nt && !reuse) fb_bulk_g2s(st + a.s_tzs, S.et_tzslot + t0, (unsigned)nt, b);
      if (lane >= 9 && lane < 9 + 2 * nts && !reuse) {   // per tensor: the (T_vvw, T_vww) pairs and the T_vwz slabs
        const int ts = (lane - 9) >> 1;
        if (((lane - 9) & 1) == 0)
          fb_bulk_g2s(st + a.s_tvw + (size_t)ts * CH * S.maxdeg * 16, S.et_tvw + 2 * (ts * S.nnzb + e0), (unsigned)(ne * 16), b);
        else if (nt)
          fb_bulk_g2s(st + a.s_tz + (size_t)ts * a.tz_stride * 8, S.et_tz + ts * S.et_tz_total + t0, (unsigned)(nt * 8), b);
      }
    }
    return;
  }

  // ================= consumer warps ========================================================================
  double* scr = (double*)(fb_as_smem + a.o_scr + (size_t)warp * a.scr_bytes);
  double* fv = scr + a.w_fv;
  double* adyn = scr + a.w_adyn;
  double* res = scr + a.w_res;
  const int grp = warp / CH, wrow = warp - grp * CH;
  const int col_one = 2 * ns + NF;
  int sidx = grp % NST;
  unsigned fph = (unsigned)((grp / NST) & 1);
  long long c0 = ((long long)blockIdx.x + (long long)grp * gridDim.x) * CH;
  const long long cstep = (long long)NG * gridDim.x * CH;
  for (int i = grp; i < nmine; i += NG, c0 += cstep) {
    const int nrow = (int)blockIdx.x + i * (int)gridDim.x == nchunks - 1
                         ? (int)(S.n_owned - (long long)(nchunks - 1) * CH) : CH;   // only the last chunk is short
    unsigned char* st = fb_as_smem + a.o_stage0 + (size_t)sidx * a.stage_bytes;
    fb_mbar_wait(&full[sidx], fph);
    if (wrow < nrow) {
      const int j = wrow;
      const long long v = c0 + j;
      const double* stat_s = (const double*)st;
      const double* us = (const double*)(st + a.s_us);
      const double* uns = (const double*)(st + a.s_uns);
      const int* rowptr_s = (const int*)(st + a.s_rowptr);
      const int e0 = rowptr_s[0];
      const int rp0 = rowptr_s[j] - e0, deg = rowptr_s[j + 1] - rowptr_s[j];
      const unsigned short* lc = (const unsigned short*)(st + a.s_lcol) + rp0;
      const unsigned vi = ((const unsigned*)(st + a.s_vinfo))[j];
      const int lself = (int)((const unsigned*)(st + a.s_rowinfo))[j];
      const unsigned rowbc = vi & 0xffffu;
      const bool colbc = (vi >> 16) & 1u;
      const double* sr = stat_s + (size_t)rp0 * NSP;
      // column `col` of the per-vertex table [u | u - u_n | affine factors | 1] at staged vertex lv;
      // kf = the row's edge index of that vertex (factor values are kept per edge)
      auto nbval = [&](int col, int lv, int kf) -> double {
        if (col < ns) return us[(size_t)lv * ns + col];
        if (col < 2 * ns) return transient ? us[(size_t)lv * ns + col - ns] - uns[(size_t)lv * ns + col - ns] : 0.0;
        if (col < col_one) return fv[kf * NF + (col - 2 * ns)];
        return 1.0;
      };
      // ---- affine reaction factors at the row's neighbours ------------------------------------------
      int kself = 0;
      for (int k = lane; k < deg; k += 32)
        if ((int)lc[k] == lself) kself = k;
      kself = __reduce_max_sync(FULL, kself);
      for (int t = lane; t < deg * NF; t += 32) {
        const int k = fb_fastdiv((unsigned)t, S.vx_m_nf), f = t - k * NF;
        double val = fac_d[f];
        const double* uk = us + (size_t)lc[k] * ns;
        const int q0 = fac_i[2 * f], q1 = q0 + fac_i[2 * f + 1];
        for (int q = q0; q < q1; ++q) val += fac_d[NF + q] * uk[fac_i[2 * NF + q]];
        fv[t] = val;
      }
      __syncwarp();
      // Rows of up to 16 edges (every row of a structured simplex mesh) do both contractions on the
      // fp64 tensor cores (DMMA m8n8k4: lane l holds A[l >> 2][l & 3], B[l & 3][l >> 2],
      // C[l >> 2][2 (l & 3) + {0, 1}]); the launcher only selects this kernel when maxdeg <= 16.
      const int krow = lane & 3, bcol = lane >> 2;
      double* tf = scr + a.w_tf;
      // ---- dynamic channel values: A^c[k] = sum_z T[k][z] f_c(z), T the row's tensor expanded to 16 x 16 --
      // solution-dependent channel values of the row: computed below, or (reuse mode) as they arrived
      const double* adp = reuse ? (const double*)(st + a.s_tz) + (size_t)rp0 * NDP : adyn;
      if (ndyn > 0 && !reuse) {
        const long long* tzptr_s = (const long long*)(st + a.s_tzptr);
        const int tzr = (int)(tzptr_s[j] - tzptr_s[0]);
        const int Wd = ((const int*)(st + a.s_tzw))[j];
        const unsigned char* zs = (const unsigned char*)(st + a.s_tzs) + tzr;
        for (int ts = 0; ts < nts; ++ts) {
          const double* Tvw = (const double*)(st + a.s_tvw) + 2 * ((size_t)ts * CH * S.maxdeg + rp0);
          const double* Tz = (const double*)(st + a.s_tz) + (size_t)ts * a.tz_stride + tzr;
          for (int t = lane; t < 16 * 17; t += 32) tf[t] = 0.0;
          __syncwarp();
          if (lane < deg) {
            const int k = lane;
            const double tvk = Tvw[2 * k];
            tf[kself * 17 + k] = tvk;                       // diagonal edge: T_vvz, z = k
            if (k != kself) {
              tf[k * 17 + kself] = tvk;                     // z = v
              tf[k * 17 + k] = Tvw[2 * k + 1];              // z = w
              for (int q = 0; q < Wd; ++q) {
                const int zz = zs[q * deg + k];
                if (zz != 0xff) tf[k * 17 + zz] = Tz[q * deg + k];
              }
            }
          }
          __syncwarp();
          const int d0 = itab[S.vx_o_dynptr + ts], d1 = itab[S.vx_o_dynptr + ts + 1];
          for (int n0 = 0; n0 < d1 - d0; n0 += 8) {
            // B operand: factor of column n0 + bcol at the neighbours z = 4 u + krow
            const bool cok = n0 + bcol < d1 - d0;
            const int fc = cok ? itab[S.vx_o_dyn + 2 * (d0 + n0 + bcol) + 1] : 0;
            double bv[4], acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
            for (int u4 = 0; u4 < 4; ++u4) {
              const int z = 4 * u4 + krow;
              bv[u4] = (cok && z < deg) ? nbval(fc, lc[z], z) : 0.0;
            }
#pragma unroll
            for (int u4 = 0; u4 < 4; ++u4)
#pragma unroll
              for (int mt = 0; mt < 2; ++mt) fb_dmma(acc[mt][0], acc[mt][1], tf[(8 * mt + bcol) * 17 + 4 * u4 + krow], bv[u4]);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
              for (int h2 = 0; h2 < 2; ++h2) {
                const int k = 8 * mt + bcol, cc = n0 + 2 * krow + h2;
                if (k < deg && cc < d1 - d0) adyn[k * NDP + (itab[S.vx_o_dyn + 2 * (d0 + cc)] - nstat)] = acc[mt][h2];
              }
          }
          __syncwarp();
        }
        for (int t = lane; t < deg; t += 32)
          for (int cpad = ndyn; cpad < NDP; ++cpad) adyn[t * NDP + cpad] = 0.0;
        __syncwarp();
        double* __restrict__ dst = a.dyn + (long long)(e0 + rp0) * NDP;
        for (int t = lane; t < deg * NDP; t += 32) dst[t] = adyn[t];
      }
      if (doF) {
        // ---- residual: R[c][g] = sum_k A^c[k] nb[k][g] (table column g), then F_s = load + sum_e coef_e R[c_e][g_e]
        const int RTS = a.rts, ntg = (S.vx_nbw + 7) / 8, mtc = (S.vx_nch + 7) / 8;
        double av[4][2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const int ch = 8 * mt + bcol;
          const bool cok = ch < S.vx_nch;
#pragma unroll
          for (int u4 = 0; u4 < 4; ++u4) {
            const int k = 4 * u4 + krow;
            double v = 0.0;
            if (cok && k < deg) v = ch < nstat ? sr[k * NSP + ch] : adp[k * NDP + ch - nstat];
            av[u4][mt] = v;
          }
        }
        __syncwarp();   // tf is free: it becomes the R tile
        if (ns <= 16) {
          // B operand by tiles of the table [u | u - u_n | factors | 1] laid out on 8-column boundaries: tiles
          // 0-1 = u (16 columns, the ones >= ns unused), 2-3 = u - u_n, 4.. = factors and the column of ones -
          // every tile loads the same way for all lanes (no per-element case analysis); results go to the
          // table's own column numbering.  Rows k >= deg of the A operand are zero, so their B entries are free.
          int ub[4];
#pragma unroll
          for (int u4 = 0; u4 < 4; ++u4) {
            const int k = 4 * u4 + krow;
            ub[u4] = k < deg ? (int)lc[k] * ns : 0;
          }
          const int ntf = (NF + 1 + 7) >> 3;
          for (int t5 = 0; t5 < 4 + ntf; ++t5) {
            const int sub = t5 < 4 ? (t5 & 1) : t5 - 4;
            if (t5 < 4 && 8 * sub >= ns) continue;
            double bv[4], acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
            if (t5 < 2) {
#pragma unroll
              for (int u4 = 0; u4 < 4; ++u4) bv[u4] = us[ub[u4] + 8 * sub + bcol];
            } else if (t5 < 4) {
#pragma unroll
              for (int u4 = 0; u4 < 4; ++u4)
                bv[u4] = transient ? us[ub[u4] + 8 * sub + bcol] - uns[ub[u4] + 8 * sub + bcol] : 0.0;
            } else {
              const int f = 8 * sub + bcol;
#pragma unroll
              for (int u4 = 0; u4 < 4; ++u4) {
                const int k = 4 * u4 + krow;
                bv[u4] = f < NF ? (k < deg ? fv[k * NF + f] : 0.0) : (f == NF ? 1.0 : 0.0);
              }
            }
#pragma unroll
            for (int u4 = 0; u4 < 4; ++u4) {
              fb_dmma(acc[0][0], acc[0][1], av[u4][0], bv[u4]);
              if (mtc > 1) fb_dmma(acc[1][0], acc[1][1], av[u4][1], bv[u4]);
            }
            // columns 2 krow, 2 krow + 1 of the tile -> table column (entries outside the table are dropped)
            const int c0 = 8 * sub + 2 * krow;
            const int lim = t5 < 4 ? ns : NF + 1;
            const int g0 = (t5 < 2 ? 0 : (t5 < 4 ? ns : 2 * ns)) + c0;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
              if (c0 < lim) tf[(8 * mt + bcol) * RTS + g0] = acc[mt][0];
              if (c0 + 1 < lim) tf[(8 * mt + bcol) * RTS + g0 + 1] = acc[mt][1];
            }
          }
        } else {
        for (int nt = 0; nt < ntg; ++nt) {
          const int gcol = 8 * nt + bcol;
          double bv[4], acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
          for (int u4 = 0; u4 < 4; ++u4) {
            const int k = 4 * u4 + krow;
            bv[u4] = (gcol < S.vx_nbw && k < deg) ? nbval(gcol, lc[k], k) : 0.0;
          }
#pragma unroll
          for (int u4 = 0; u4 < 4; ++u4) {
            fb_dmma(acc[0][0], acc[0][1], av[u4][0], bv[u4]);
            if (mtc > 1) fb_dmma(acc[1][0], acc[1][1], av[u4][1], bv[u4]);
          }
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            tf[(8 * mt + bcol) * RTS + 8 * nt + 2 * krow] = acc[mt][0];
            tf[(8 * mt + bcol) * RTS + 8 * nt + 2 * krow + 1] = acc[mt][1];
          }
        }
        }
        __syncwarp();
        const int nft = S.vx_nft;
        for (int t = lane; t < nft; t += 32) {
          const int chan = itab[S.vx_o_ft + 3 * t], gcol = itab[S.vx_o_ft + 3 * t + 1], fl = itab[S.vx_o_ft + 3 * t + 2];
          res[t] = tf[chan * RTS + gcol] * dtab[S.vx_o_ftc + t] * ((fl & 1) ? S.inv_dt : 1.0);
        }
        __syncwarp();
        double facc = 0.0;
        if (lane < ns) {
          facc = ((const double*)(st + a.s_load))[j * ns + lane];
          const int a0 = itab[S.vx_o_ftptr + lane], a1 = itab[S.vx_o_ftptr + lane + 1];
          for (int t = a0; t < a1; ++t) facc += res[t];
        }
        if (colbc) {   // lifting b += J[:, B] (g - x_B) with the live entries of the Dirichlet columns (rare rows)
          __syncwarp();
          const int njt = S.vx_njt;
          for (int t = lane; t < njt; t += 32) {
            const int chan = itab[S.vx_o_jt + 2 * t], fl = itab[S.vx_o_jt + 2 * t + 1];
            const int cs = S.pair_cs[itab[S.vx_o_jtpair + t]];
            double acc = 0.0;
            for (int k = 0; k < deg; ++k) {
              const int w = S.colidx[e0 + rp0 + k];
              if ((S.vinfo[w] >> cs) & 1u) {
                const double av = chan < nstat ? sr[k * NSP + chan] : adp[k * NDP + chan - nstat];
                acc += av * (S.bc_vals[S.bc_map[(long long)w * ns + cs]] - us[(size_t)lc[k] * ns + cs]);
              }
            }
            res[t] = acc * dtab[S.vx_o_jtc + t] * ((fl & 1) ? S.inv_dt : 1.0);
          }
          __syncwarp();
          if (lane < ns) {
            const int a0 = itab[S.vx_o_jtptr + S.pair_pp[lane]], a1 = itab[S.vx_o_jtptr + S.pair_pp[lane + 1]];
            for (int t = a0; t < a1; ++t) facc += res[t];
          }
        }
        if (lane < ns) {
          const long long dof = v * ns + lane;
          if ((rowbc >> lane) & 1u) a.F[dof] = us[(size_t)lself * ns + lane] - S.bc_vals[S.bc_map[dof]];
          else a.F[dof] = facc;
        }
      }
    }
    __syncwarp();
    if (lane == 0) fb_mbar_arrive(&empty[sidx]);
    sidx += NG;
    if (sidx >= NST) { sidx -= NST; fph ^= 1u; }
  }
}
