// SpMV on the channel operator, TMA-pipelined (sm_100a).
//
//   J = sum_c A^c (x) B_c   (assemble_edge.cuh);   y_v = sum_w sum_c A^c_vw (B_c x_w)
//
// Replaces PETSc MatMult inside KSP (reference src/festim/problem.py:271-289) for systems whose
// Jacobian is held as channel values.  One persistent block per SM walks chunks of FB_CHUNK
// consecutive block rows.  Everything a chunk reads from HBM arrives through the TMA unit as
// large 1D bulk copies (cp.async.bulk) into a ring of shared-memory stages, completing on one
// mbarrier per stage: the static and the solution-dependent channel values of the chunk's rows
// (contiguous), the ns-blocks of x at the chunk's distinct neighbour vertices (one copy per run of
// consecutive vertex ids, from the chunk plan of fb_set_pattern: neighbours shared by the rows of
// a chunk are fetched once), the 16-bit local column indices, the Dirichlet words, and for the
// fused epilogues the inverse diagonal blocks and the vectors of the recurrence.  Warp 0 issues
// the copies of chunk i + NST - 1 while all warps multiply chunk i, so no register or LSU issue
// slot is spent on streaming and the copy queue never drains.  Lane l owns the products
// (channel, column species) l, l+32, ...: t[(c, cs)] = sum_w A^c_vw x_(w,cs); the constant tables
// B_c are applied once per row.  Results of a chunk leave through one coalesced write, which is
// also where the fused epilogues run:
//   MODE 0   y = J x
//   MODE 1   y = D^-1 J x                                   (point-block Jacobi, GMRES operator)
//   MODE 2   one step of the Chebyshev recurrence on B = D^-1 J with the iterate x = d_in:
//              r <- r - B d_in ;  d_out = ca d_in + cb r ;  z <- z + d_out
//            and on the last step  w = v_in - r  (= B z, the polynomially preconditioned vector)
// Strong Dirichlet conditions are applied on the fly (eliminated columns read as zero, identity
// rows copy x), so the operator equals the block-CSR matrix k_et_expand writes.
#pragma once
#include "fb_device.cuh"

#ifndef FB_EMU
__device__ __forceinline__ unsigned fb_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fb_mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fb_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fb_mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fb_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb_smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded: a byte count that never arrives (a bug) traps instead of hanging the GPU
__device__ __forceinline__ void fb_mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned addr = fb_smem_u32(bar);
  for (unsigned spin = 0; spin < (1u << 26); ++spin) {
    unsigned done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}
// 1D bulk copy global -> shared through the TMA unit; bytes and both addresses multiples of 16
__device__ __forceinline__ void fb_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   fb_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(fb_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fb_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fb_mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(fb_smem_u32(bar)) : "memory");
}
#else   // host emulation of the execution model (tests/emu): functional mbarriers, synchronous copies
#include <cstring>
#include <map>
struct FbEmuBar { int init = 0, pending = 0; long long tx = 0; unsigned phase = 0; };
inline std::map<void*, FbEmuBar>& fb_emu_bars() { static std::map<void*, FbEmuBar> m; return m; }
inline void fb_emu_check(FbEmuBar& b) { if (b.pending == 0 && b.tx == 0) { b.phase ^= 1u; b.pending = b.init; } }
inline void fb_mbar_init(unsigned long long* bar, unsigned count) {
  std::lock_guard<std::mutex> lk(g_atomic_mutex);
  FbEmuBar b; b.init = b.pending = (int)count;
  fb_emu_bars()[bar] = b;
}
inline void fb_mbar_fence_init() {}
inline void fb_fence_proxy_async() {}
inline void fb_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  std::lock_guard<std::mutex> lk(g_atomic_mutex);
  FbEmuBar& b = fb_emu_bars()[bar];
  b.tx += bytes; b.pending -= 1;
  fb_emu_check(b);
}
inline void fb_mbar_arrive(unsigned long long* bar) {
  std::lock_guard<std::mutex> lk(g_atomic_mutex);
  FbEmuBar& b = fb_emu_bars()[bar];
  b.pending -= 1;
  fb_emu_check(b);
}
inline void fb_mbar_wait(unsigned long long* bar, unsigned parity) {
  for (;;) {
    {
      std::lock_guard<std::mutex> lk(g_atomic_mutex);
      if (fb_emu_bars()[bar].phase != parity) return;
    }
    std::this_thread::yield();
  }
}
inline void fb_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  memcpy(dst, src, bytes);
  std::lock_guard<std::mutex> lk(g_atomic_mutex);
  FbEmuBar& b = fb_emu_bars()[bar];
  b.tx -= bytes;
  fb_emu_check(b);
}
#endif

// D (8x8) += A (8x4) * B (4x8) in fp64 on the tensor cores (DMMA).  Fragments: lane l holds
// A[l >> 2][l & 3], B[l & 3][l >> 2] and C[l >> 2][2 (l & 3) + {0, 1}].
#ifndef FB_EMU
__device__ __forceinline__ void fb_dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
#else
inline void fb_dmma(double& c0, double& c1, double a, double b) {
  const int l = threadIdx.x & 31, row = l >> 2, col = 2 * (l & 3);
  for (int k = 0; k < 4; ++k) {
    const double ak = __shfl_sync(0xffffffffu, a, 4 * row + k);
    const double b0 = __shfl_sync(0xffffffffu, b, k + 4 * col);
    const double b1 = __shfl_sync(0xffffffffu, b, k + 4 * (col + 1));
    c0 += ak * b0;
    c1 += ak * b1;
  }
}
#endif

#define FB_SP_EPL 6   // table entries per lane kept in registers (two lanes per row species)

struct SpmvChArgs {
  const double* x;
  double* y;
  const double* dinv;
  const int* flags;
  // Chebyshev step (MODE 2)
  double* r;
  double* z;
  double* dout;
  const double* vin;
  double* wout;
  double ca, cb;
  int last;
  // shared-memory layout (bytes): block-shared tables, per-warp scratch, chunk results, stages
  int n_sp_ent, nst;
  int o_part, part_bytes, o_ybuf, o_stage0, stage_bytes, o_bar;
  int s_dyn, s_xs, s_dinv, s_r, s_z, s_vin, s_lcol, s_vinfo, s_rowinfo, s_rowptr;   // offsets inside a stage
};

// fills the layout for MODE; returns the bytes of dynamic shared memory with nst stages
__host__ inline size_t fb_spmv_ch_layout(const DevSpec& S, SpmvChArgs& a, int mode, int nst) {
  auto up16 = [](size_t b) { return (b + 15) & ~(size_t)15; };
  const int md = S.maxdeg, ns = S.ns, CH = FB_CHUNK;
  size_t o = 0;
  o += up16((size_t)a.n_sp_ent * 8);                       // t_coef
  o += up16((size_t)(ns + 1 + a.n_sp_ent) * 4);            // t_ptr, t_combo
  // per consumer warp: the (channel x species) product tile of its row, then the row's result
  a.o_part = (int)o; a.part_bytes = (int)up16((size_t)(256 + ns) * 8); o += (size_t)CH * FB_CHUNK_GROUPS * a.part_bytes;
  a.o_ybuf = 0;
  size_t s = 0;
  s += up16((size_t)CH * md * S.vx_nsp * 8);               // static channel values
  a.s_dyn = (int)s; s += up16((size_t)CH * md * S.vx_ndp * 8);
  a.s_xs = (int)s; s += up16((size_t)(S.ck_nx_max + 2) * ns * 8);   // + tile padding read past the last vertex
  a.s_dinv = (int)s; if (mode >= 1) s += up16((size_t)CH * ns * ns * 8);
  a.s_r = (int)s; if (mode == 2) s += up16((size_t)CH * ns * 8);
  a.s_z = (int)s; if (mode == 2) s += up16((size_t)CH * ns * 8);
  a.s_vin = (int)s; if (mode == 2) s += up16((size_t)CH * ns * 8);
  a.s_lcol = (int)s; s += up16((size_t)(CH * md + 8) * 2);
  a.s_vinfo = (int)s; s += up16((size_t)CH * 4);
  a.s_rowinfo = (int)s; s += up16((size_t)CH * 4);
  a.s_rowptr = (int)s; s += up16((size_t)(CH + 4) * 4);
  a.stage_bytes = (int)s;
  a.o_stage0 = (int)o; o += (size_t)nst * s;
  a.o_bar = (int)o; o += 16 * 8;   // full[nst], empty[nst]
  a.nst = nst;
  return o;
}

template <int MT, int NT, int MODE>
__global__ void __launch_bounds__(32 * FB_CHUNK * FB_CHUNK_GROUPS + 32, 1)
k_spmv_ch_tma(DevSpec S, SpmvChArgs a) {
  extern __shared__ __align__(16) unsigned char fb_sp_smem[];
  if (a.flags && a.flags[0]) return;
  constexpr int CH = FB_CHUNK;
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ns = S.ns, NSP = S.vx_nsp, NDP = S.vx_ndp, NST = a.nst;
  // ---- block-shared epilogue tables: sp_coef (inv_dt folded), sp_ptr, positions in the product tile
  double* t_coef = (double*)fb_sp_smem;
  int* t_ptr = (int*)(fb_sp_smem + (((size_t)a.n_sp_ent * 8 + 15) & ~(size_t)15));
  int* t_combo = t_ptr + ns + 1;
  for (int t = threadIdx.x; t < a.n_sp_ent; t += blockDim.x) {
    t_coef[t] = S.vx_sp_coef[t] * ((S.vx_sp[2 * t + 1] & 1) ? S.inv_dt : 1.0);
    const int combo = S.vx_sp[2 * t];   // (channel, column species) -> position in the row's product tile
    t_combo[t] = S.vx_sp_combo[2 * combo] * (8 * NT) + S.vx_sp_combo[2 * combo + 1];
  }
  for (int t = threadIdx.x; t <= ns; t += blockDim.x) t_ptr[t] = S.vx_sp_ptr[t];
  unsigned long long* full = (unsigned long long*)(fb_sp_smem + a.o_bar);
  unsigned long long* empty = full + 8;
  if (threadIdx.x == 0)
    for (int q = 0; q < NST; ++q) { fb_mbar_init(&full[q], 1); fb_mbar_init(&empty[q], CH); }
  // tile padding of the DMMA operands is read from the stages: it must be finite
  for (size_t t = threadIdx.x; t < (size_t)NST * a.stage_bytes / 16; t += blockDim.x)
    ((double2*)(fb_sp_smem + a.o_stage0))[t] = make_double2(0.0, 0.0);
  fb_fence_proxy_async();
  fb_mbar_fence_init();
  __syncthreads();
  const long long nchunks = (S.n_owned + CH - 1) / CH;
  const long long nmine = blockIdx.x < nchunks ? (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  constexpr int NG = FB_CHUNK_GROUPS;
  if (warp == CH * NG) {
    // ================= producer warp: every operand of a chunk through the TMA unit ==================
    // descriptors of chunk i+1 are fetched while the copies of chunk i are issued
    int pf_e0, pf_ne, pf_nx, pf_lc, pf_q0 = 0, pf_q1 = 0, pf_run[3];
    auto prefetch = [&](long long i) {
      if (i >= nmine) return;
      const long long c = blockIdx.x + i * (long long)gridDim.x;
      const long long c0 = c * CH;
      const int nrow = (int)((S.n_owned - c0) < CH ? (S.n_owned - c0) : CH);
      pf_e0 = S.rowptr[c0]; pf_ne = S.rowptr[c0 + nrow]; pf_nx = S.ck_nx[c]; pf_lc = S.ck_lcolptr[c];
      const int* rr = S.ck_slots + ((size_t)c * 32 + lane) * 3;
      pf_run[0] = rr[0]; pf_run[1] = rr[1]; pf_run[2] = rr[2];
      if (S.ck_rmax > 32) { pf_q0 = S.ck_runptr[c]; pf_q1 = S.ck_runptr[c + 1]; }
    };
    prefetch(0);
    for (long long i = 0; i < nmine; ++i) {
      const int sidx = (int)(i % NST);
      if (i >= NST) fb_mbar_wait(&empty[sidx], (unsigned)(((i / NST) - 1) & 1));   // all rows of its last chunk done
      const long long c = blockIdx.x + i * (long long)gridDim.x;
      const long long c0 = c * CH;
      const int nrow = (int)((S.n_owned - c0) < CH ? (S.n_owned - c0) : CH);
      unsigned char* st = fb_sp_smem + a.o_stage0 + (size_t)sidx * a.stage_bytes;
      unsigned long long* b = &full[sidx];
      const int e0 = pf_e0, ne = pf_ne - pf_e0, nx = pf_nx, lc0 = pf_lc, q0 = pf_q0, q1 = pf_q1;
      const int r0v = pf_run[0], r1v = pf_run[1], r2v = pf_run[2];
      prefetch(i + 1);
      const int nlc = (ne + 7) & ~7;
      if (lane == 0) {
        unsigned bytes = (unsigned)(ne * (NSP + NDP) * 8 + nx * ns * 8 + nlc * 2 + 2 * CH * 4 + (CH + 4) * 4);
        if (MODE >= 1) bytes += (unsigned)(nrow * ns * ns * 8);
        if (MODE == 2) bytes += (unsigned)((a.last ? 3 : 2) * nrow * ns * 8);
        fb_mbar_expect_tx(b, bytes);
      }
      __syncwarp();
      if (lane == 0 && NSP) fb_bulk_g2s(st, S.et_stat + (long long)e0 * NSP, (unsigned)(ne * NSP * 8), b);
      if (lane == 1 && NDP) fb_bulk_g2s(st + a.s_dyn, S.et_dyn + (long long)e0 * NDP, (unsigned)(ne * NDP * 8), b);
      if (lane == 2) fb_bulk_g2s(st + a.s_lcol, S.ck_lcol + lc0, (unsigned)(nlc * 2), b);
      if (lane == 3) fb_bulk_g2s(st + a.s_vinfo, S.vinfo + c0, (unsigned)(CH * 4), b);
      if (lane == 4) fb_bulk_g2s(st + a.s_rowinfo, S.ck_rowinfo + c0, (unsigned)(CH * 4), b);
      if (lane == 9) fb_bulk_g2s(st + a.s_rowptr, S.rowptr + c0, (unsigned)((CH + 4) * 4), b);
      if (MODE >= 1 && lane == 5)
        fb_bulk_g2s(st + a.s_dinv, a.dinv + c0 * ns * ns, (unsigned)(nrow * ns * ns * 8), b);
      if (MODE == 2) {
        if (lane == 6) fb_bulk_g2s(st + a.s_r, a.r + c0 * ns, (unsigned)(nrow * ns * 8), b);
        if (lane == 7) fb_bulk_g2s(st + a.s_z, a.z + c0 * ns, (unsigned)(nrow * ns * 8), b);
        if (lane == 8 && a.last) fb_bulk_g2s(st + a.s_vin, a.vin + c0 * ns, (unsigned)(nrow * ns * 8), b);
      }
      // x at the chunk's distinct neighbours, one copy per run of vertex ids
      if (r1v > 0)
        fb_bulk_g2s(st + a.s_xs + (size_t)r2v * ns * 8, a.x + (long long)r0v * ns, (unsigned)(r1v * ns * 8), b);
      for (int q = q0 + 32 + lane; q < q1; q += 32) {
        const int v0 = S.ck_runs[3 * q], len = S.ck_runs[3 * q + 1], lo = S.ck_runs[3 * q + 2];
        fb_bulk_g2s(st + a.s_xs + (size_t)lo * ns * 8, a.x + (long long)v0 * ns, (unsigned)(len * ns * 8), b);
      }
    }
    return;
  }

  // ================= consumer warps: group g = warp / CH takes the chunks i = g, g + NG, ...;
  // its warp j multiplies row j of each of them ======================================================
  double* part = (double*)(fb_sp_smem + a.o_part + (size_t)warp * a.part_bytes);
  double* yrow = part + 256;
  const int grp = warp / CH, wrow = warp - grp * CH;
  // DMMA fragments of this lane: channels 8 mt + (lane >> 2) (A operand), column species
  // 8 nt + (lane >> 2) (B operand); channel values of an edge are strided by nsp / ndp
  const int nch = S.vx_nch;
  int aoff[MT], astr[MT];
  bool adyn[MT];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const int ch = 8 * mt + (lane >> 2);
    adyn[mt] = ch >= S.vx_nstat && ch < nch;   // padding channels read the static block (values unused)
    aoff[mt] = adyn[mt] ? ch - S.vx_nstat : (ch < nch ? ch : 0);
    astr[mt] = adyn[mt] ? NDP : NSP;
  }
  const int bcol = lane >> 2, kq = lane & 3;
  // the table entries of row species s are added by lanes s and s + 16
  const int srow = lane & 15;
  int e_lo = 0, e_hi = 0;
  if (srow < ns) {
    const int p0 = t_ptr[srow], p1 = t_ptr[srow + 1], mid = (p0 + p1 + 1) >> 1;
    e_lo = lane < 16 ? p0 : mid; e_hi = lane < 16 ? mid : p1;
  }

  for (long long i = grp; i < nmine; i += NG) {
    const int sidx = (int)(i % NST);
    const long long c = blockIdx.x + i * (long long)gridDim.x;
    const long long c0 = c * CH;
    const int nrow = (int)((S.n_owned - c0) < CH ? (S.n_owned - c0) : CH);
    unsigned char* st = fb_sp_smem + a.o_stage0 + (size_t)sidx * a.stage_bytes;
    fb_mbar_wait(&full[sidx], (unsigned)((i / NST) & 1));
    if (wrow < nrow) {
      const int j = wrow;
      const double* stat_s = (const double*)st;
      const double* dyn_s = (const double*)(st + a.s_dyn);
      const double* xs = (const double*)(st + a.s_xs);
      const unsigned short* lcol_s = (const unsigned short*)(st + a.s_lcol);
      const int* rowptr_s = (const int*)(st + a.s_rowptr);
      const int e0 = rowptr_s[0];
      const int rp0 = rowptr_s[j] - e0, deg = rowptr_s[j + 1] - rowptr_s[j];
      const unsigned vi = ((const unsigned*)(st + a.s_vinfo))[j];
      const int lself = (int)((const unsigned*)(st + a.s_rowinfo))[j];
      const unsigned rowbc = vi & 0xffffu;
      const bool colbc = (vi >> 16) & 1u;
      const double* sr = stat_s + (size_t)rp0 * NSP;
      const double* dr = dyn_s + (size_t)rp0 * NDP;
      const unsigned short* lc = lcol_s + rp0;
      // T[c][cs] = sum_k A^c[k] x[k][cs] on the fp64 tensor cores: K = edges in steps of 4
      double acc[MT][NT][2];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
      const double* ap[MT];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) ap[mt] = (adyn[mt] ? dr : sr) + aoff[mt];
      // Edges beyond the row (k >= deg) are cancelled through the B operand alone: the A operand
      // then reads finite values of the next row.  Channels >= nch and species >= ns are tile
      // padding: their products land in entries of T no table refers to.
      const double* xb = xs + 8 * 0 + bcol;
      for (int kb = 0; kb < deg; kb += 16) {
        // all operand loads of four K steps first, then the sixteen DMMAs
        double av[4][MT], bv[4][NT];
        if (!colbc) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int k = kb + 4 * u + kq;
            const bool val = k < deg;
            const double* xk = xb + (size_t)lc[val ? k : 0] * ns;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) av[u][mt] = ap[mt][k * astr[mt]];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) { const double xv = xk[8 * nt]; bv[u][nt] = val ? xv : 0.0; }
          }
        } else {   // eliminated (Dirichlet) columns read as zero
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int k = kb + 4 * u + kq;
            const bool val = k < deg;
            const unsigned msk = val ? S.vinfo[S.colidx[e0 + rp0 + k]] : 0u;
            const double* xk = xb + (size_t)lc[val ? k : 0] * ns;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) av[u][mt] = ap[mt][k * astr[mt]];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
              const double xv = xk[8 * nt];
              bv[u][nt] = (val && !((msk >> (8 * nt + bcol)) & 1u)) ? xv : 0.0;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) fb_dmma(acc[mt][nt][0], acc[mt][nt][1], av[u][mt], bv[u][nt]);
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          double2 v2; v2.x = acc[mt][nt][0]; v2.y = acc[mt][nt][1];
          *(double2*)(part + (8 * mt + (lane >> 2)) * (8 * NT) + 8 * nt + 2 * (lane & 3)) = v2;
        }
      __syncwarp();
      // y_s = sum_e coef_e T[combo_e]: two lanes per row species, one shuffle
      // (all table loads are independent of the products: issued ahead of the DMMA results)
      double sum = 0.0, sum2 = 0.0;
#pragma unroll
      for (int u = 0; u < FB_SP_EPL; u += 2) {
        const int ea = e_lo + u, eb = e_lo + u + 1;
        const double ca = ea < e_hi ? t_coef[ea] : 0.0, cb = eb < e_hi ? t_coef[eb] : 0.0;
        sum += ca * part[ea < e_hi ? t_combo[ea] : 0];
        sum2 += cb * part[eb < e_hi ? t_combo[eb] : 0];
      }
      for (int e = e_lo + FB_SP_EPL; e < e_hi; ++e) sum += t_coef[e] * part[t_combo[e]];
      sum += sum2;
      sum += __shfl_xor_sync(FULL, sum, 16);
      const long long dof = (c0 + j) * ns + lane;
      if (lane < ns) {
        if ((rowbc >> lane) & 1u) sum = xs[(size_t)lself * ns + lane];
        if (MODE == 0) a.y[dof] = sum;
        else yrow[lane] = sum;
      }
      if (MODE >= 1) {
        __syncwarp();
        if (lane < ns) {
          double p;
          const int t = j * ns + lane;
          if (ns == 1) p = ((const double*)(st + a.s_dinv))[t] * yrow[0];
          else {
            const double2* dv = (const double2*)(st + a.s_dinv) + (size_t)t * (ns / 2);   // row s of the inverse block
            double pa = 0.0, pb = 0.0;
            for (int c2 = 0; c2 < ns; c2 += 2) {   // ns is even (16-byte rows)
              const double2 d2 = dv[c2 >> 1];
              pa += d2.x * yrow[c2];
              pb += d2.y * yrow[c2 + 1];
            }
            p = pa + pb;
          }
          if (MODE == 1) a.y[dof] = p;
          else {
            const double rn = ((const double*)(st + a.s_r))[t] - p;
            if (a.last) a.wout[dof] = ((const double*)(st + a.s_vin))[t] - rn;
            else {
              const double dn = a.ca * xs[(size_t)lself * ns + lane] + a.cb * rn;
              a.r[dof] = rn;
              a.dout[dof] = dn;
              a.z[dof] = ((const double*)(st + a.s_z))[t] + dn;
            }
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) fb_mbar_arrive(&empty[sidx]);   // this warp is done with the stage
  }
}
