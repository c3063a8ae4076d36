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
#include <vector>

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

// Lane plan of the row epilogue y_s = sum_e coef_e T[combo_e]: the entries of row species s are dealt
// to an aligned group of 1, 2, 4 or 8 lanes (at most EPL entries per lane, kept in registers for the
// whole launch); the group's partial sums meet in xor shuffles and its first lane owns y_s.
#define FB_SP_EPL_MAX 8
struct SpmvEpPlan {
  int epl = 0;                      // entries per lane (0: no plan - too many entries for 32 lanes)
  std::vector<double> coef;         // [32 * epl]
  std::vector<int> pos;             // [32 * epl]  tile position | inv_dt flag << 16
  std::vector<int> lane;            // [32]  species | log2(group) << 8 | leader << 12   (species 0xff: idle lane)
};
// sp_ptr [ns+1], sp [n*2] (combo, flag), sp_coef [n], sp_combo [ncombo*2] (channel, column species)
inline SpmvEpPlan fb_spmv_ep_plan(int ns, const int* sp_ptr, const int* sp, const double* sp_coef, const int* sp_combo) {
  SpmvEpPlan P;
  if (ns > 16) return P;
  for (int epl = 2; epl <= FB_SP_EPL_MAX; epl += 2) {   // the kernel is instantiated for 2, 4, 6, 8
    int lanes = 0;
    std::vector<int> g(ns);
    for (int s = 0; s < ns; ++s) {
      const int n = sp_ptr[s + 1] - sp_ptr[s];
      int need = (n + epl - 1) / epl, gs = 1;
      if (need < 1) need = 1;
      while (gs < need) gs *= 2;
      g[s] = gs; lanes += gs;
    }
    if (lanes > 32) continue;
    P.epl = epl;
    P.coef.assign((size_t)32 * epl, 0.0); P.pos.assign((size_t)32 * epl, 0); P.lane.assign(32, 0xff);
    // the product tile is stored fragment-major (DMMA accumulator (mt, nt) of lane l at double2 unit
    // (mt NT + nt) 32 + l): position of T[c][cs] in doubles
    const int NT = (ns + 7) / 8;
    auto tpos = [&](int c, int cs) {
      const int mt = c >> 3, r = c & 7, nt = cs >> 3, col = cs & 7;
      return ((mt * NT + nt) * 32 + 4 * r + (col >> 1)) * 2 + (col & 1);
    };
    std::vector<std::vector<std::pair<double, int>>> ent(32);
    int at = 0;
    for (int gs = 8; gs >= 1; gs /= 2)   // larger groups first: every group starts at a multiple of its size
      for (int s = 0; s < ns; ++s) {
        if (g[s] != gs) continue;
        int lg = 0;
        while ((1 << lg) < gs) ++lg;
        const int n = sp_ptr[s + 1] - sp_ptr[s];
        for (int q = 0; q < gs; ++q) P.lane[at + q] = s | (lg << 8) | ((q == 0) << 12);
        for (int t = 0; t < n; ++t) {   // dealt round-robin over the group's lanes
          const int e = sp_ptr[s] + t, combo = sp[2 * e];
          ent[at + t % gs].push_back({sp_coef[e], tpos(sp_combo[2 * combo], sp_combo[2 * combo + 1]) | ((sp[2 * e + 1] & 1) << 16)});
        }
        at += gs;
      }
    // order of a lane's entries: the 16 lanes of a half warp read 8-byte words in the same instruction,
    // word w from bank pair w mod 16 - greedily pick, per instruction, entries of unused bank pairs
    for (int u = 0; u < epl; ++u)
      for (int h = 0; h < 2; ++h) {
        int used[16] = {0};
        for (int l = 16 * h; l < 16 * h + 16; ++l) {
          auto& E = ent[l];
          if (E.empty()) continue;
          size_t best = 0;
          for (size_t q = 1; q < E.size(); ++q)
            if (used[(E[q].second & 0xffff) & 15] < used[(E[best].second & 0xffff) & 15]) best = q;
          P.coef[(size_t)l * epl + u] = E[best].first;
          P.pos[(size_t)l * epl + u] = E[best].second;
          used[(E[best].second & 0xffff) & 15]++;
          E.erase(E.begin() + best);
        }
        for (int l = 16 * h; l < 16 * h + 16; ++l) {   // idle slots (coefficient 0) read a word of the least used bank pair
          if (P.coef[(size_t)l * epl + u] != 0.0 || P.pos[(size_t)l * epl + u] != 0) continue;
          int w = 0;
          for (int q = 1; q < 16; ++q)
            if (used[q] < used[w]) w = q;
          P.pos[(size_t)l * epl + u] = w;
          used[w]++;
        }
      }
    return P;
  }
  return P;
}

struct SpmvChArgs {
  const double* x;
  double* y;
  const float* dinv32;      // inverse diagonal blocks: single-precision mantissas ...
  const double* dscale;     // ... times one power-of-two scale per row (k_bjacobi_setup_slab)
  const int* flags;
  // Chebyshev step (MODE 2)
  double* r;
  double* z;
  double* dout;
  const double* vin;
  double* wout;
  double ca, cb;
  int last;
  // lane plan of the row epilogue (fb_spmv_ep_plan)
  const double* ep_coef;
  const int* ep_pos;
  const int* ep_lane;
  // shared-memory layout (bytes): per-warp scratch, stages
  int nst, zslot;
  int o_part, part_bytes, o_ybuf, o_stage0, stage_bytes, o_bar;
  int s_dyn, s_xs, s_dinv, s_dscale, s_r, s_z, s_vin, s_lcol, s_vinfo, s_rowinfo, s_rowptr;   // offsets inside a stage
};

// fills the layout for MODE; returns the bytes of dynamic shared memory with nst stages
__host__ inline size_t fb_spmv_ch_layout(const DevSpec& S, SpmvChArgs& a, int mode, int nst) {
  auto up16 = [](size_t b) { return (b + 15) & ~(size_t)15; };
  const int md = S.maxdeg, ns = S.ns, CH = FB_CHUNK;
  size_t o = 0;
  // per consumer warp: the (channel x species) product tile of its row, then the row's result
  a.o_part = (int)o; a.part_bytes = (int)up16((size_t)(256 + ns) * 8); o += (size_t)CH * FB_CHUNK_GROUPS * a.part_bytes;
  a.o_ybuf = 0;
  size_t s = 0;
  s += up16((size_t)CH * md * S.vx_nsp * 8);               // static channel values
  a.s_dyn = (int)s; s += up16((size_t)CH * md * S.vx_ndp * 8);
  // + a slot of zeros for the edges of the K padding (zslot) and the tile padding read past a vertex
  a.zslot = S.ck_nx_max + 2;
  a.s_xs = (int)s; s += up16((size_t)(S.ck_nx_max + 4) * ns * 8);
  a.s_dinv = (int)s; if (mode >= 1) s += up16((size_t)CH * ns * ns * 4);
  a.s_dscale = (int)s; if (mode >= 1) s += up16((size_t)CH * ns * 8);
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

template <int MT, int NT, int MODE, int EPL>
__global__ void __launch_bounds__(32 * FB_CHUNK * FB_CHUNK_GROUPS + 64, 1)
k_spmv_ch_tma(DevSpec S, SpmvChArgs a) {
  extern __shared__ __align__(16) unsigned char fb_sp_smem[];
  if (a.flags && a.flags[0]) return;
  constexpr int CH = FB_CHUNK;
  constexpr int NG = FB_CHUNK_GROUPS;
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ns = S.ns, NSP = S.vx_nsp, NDP = S.vx_ndp, NST = a.nst;
  unsigned long long* full = (unsigned long long*)(fb_sp_smem + a.o_bar);
  unsigned long long* empty = full + 8;
  if (threadIdx.x == 0)
    for (int q = 0; q < NST; ++q) { fb_mbar_init(&full[q], 2); fb_mbar_init(&empty[q], CH); }   // full: one arrival per producer warp
  // tile padding of the DMMA operands and the zero slot are read from the stages: they must be finite / zero
  for (size_t t = threadIdx.x; t < (size_t)NST * a.stage_bytes / 16; t += blockDim.x)
    ((double2*)(fb_sp_smem + a.o_stage0))[t] = make_double2(0.0, 0.0);
  fb_fence_proxy_async();
  fb_mbar_fence_init();
  __syncthreads();
  const int nchunks = (int)((S.n_owned + CH - 1) / CH);
  const int nmine = (int)blockIdx.x < nchunks ? (nchunks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp >= CH * NG) {
    // ================= two producer warps: every operand of a chunk through the TMA unit ==============
    // (a single warp issuing all ~22 copies of a chunk was the slowest warp of the block: warp 0 takes the
    // row data, warp 1 the runs of x).  Descriptors of chunk i+1 are fetched while the copies of chunk i are issued.
    const bool px = warp > CH * NG;
    int pf_e0 = 0, pf_ne = 0, pf_nx = 0, pf_lc = 0, pf_q0 = 0, pf_q1 = 0, pf_run[3] = {0, 0, 0};
    auto prefetch = [&](int i) {
      if (i >= nmine) return;
      const long long c = blockIdx.x + i * (long long)gridDim.x;
      const long long c0 = c * CH;
      if (!px) {
        const int nrow = (int)((S.n_owned - c0) < CH ? (S.n_owned - c0) : CH);
        pf_e0 = S.rowptr[c0]; pf_ne = S.rowptr[c0 + nrow]; pf_lc = S.ck_lcolptr[c];
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
    for (int i = 0; i < nmine; ++i) {
      fb_mbar_wait(&empty[sidx], eph);   // all rows of the stage's last chunk done
      const long long c = blockIdx.x + i * (long long)gridDim.x;
      const long long c0 = c * CH;
      const int nrow = (int)((S.n_owned - c0) < CH ? (S.n_owned - c0) : CH);
      unsigned char* st = fb_sp_smem + a.o_stage0 + (size_t)sidx * a.stage_bytes;
      unsigned long long* b = &full[sidx];
      const int e0 = pf_e0, ne = pf_ne - pf_e0, nx = pf_nx, lc0 = pf_lc, q0 = pf_q0, q1 = pf_q1;
      const int r0v = pf_run[0], r1v = pf_run[1], r2v = pf_run[2];
      prefetch(i + 1);
      if (!px) {
        const int nlc = (ne + 7) & ~7;
        if (lane == 0) {
          unsigned bytes = (unsigned)(ne * (NSP + NDP) * 8 + nlc * 2 + 2 * CH * 4 + (CH + 4) * 4);
          if (MODE >= 1) bytes += (unsigned)(nrow * ns * ns * 4 + nrow * ns * 8);
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
          fb_bulk_g2s(st + a.s_dinv, a.dinv32 + c0 * ns * ns, (unsigned)(nrow * ns * ns * 4), b);
        if (MODE >= 1 && lane == 10) fb_bulk_g2s(st + a.s_dscale, a.dscale + c0 * ns, (unsigned)(nrow * ns * 8), b);
        if (MODE == 2) {
          if (lane == 6) fb_bulk_g2s(st + a.s_r, a.r + c0 * ns, (unsigned)(nrow * ns * 8), b);
          if (lane == 7) fb_bulk_g2s(st + a.s_z, a.z + c0 * ns, (unsigned)(nrow * ns * 8), b);
          if (lane == 8 && a.last) fb_bulk_g2s(st + a.s_vin, a.vin + c0 * ns, (unsigned)(nrow * ns * 8), b);
        }
      } else {
        // x at the chunk's distinct neighbours, one copy per run of vertex ids
        if (lane == 0) fb_mbar_expect_tx(b, (unsigned)(nx * ns * 8));
        __syncwarp();
        if (r1v > 0)
          fb_bulk_g2s(st + a.s_xs + (size_t)r2v * ns * 8, a.x + (long long)r0v * ns, (unsigned)(r1v * ns * 8), b);
        for (int q = q0 + 32 + lane; q < q1; q += 32) {
          const int v0 = S.ck_runs[3 * q], len = S.ck_runs[3 * q + 1], lo = S.ck_runs[3 * q + 2];
          fb_bulk_g2s(st + a.s_xs + (size_t)lo * ns * 8, a.x + (long long)v0 * ns, (unsigned)(len * ns * 8), b);
        }
      }
      if (++sidx == NST) { sidx = 0; eph ^= 1u; }
    }
    return;
  }

  // ================= consumer warps: group g = warp / CH takes the chunks i = g, g + NG, ...;
  // its warp j multiplies row j of each of them ======================================================
  double* part = (double*)(fb_sp_smem + a.o_part + (size_t)warp * a.part_bytes);
  double* yrow = part + 256;
  const int grp = warp / CH, j = warp - grp * CH;
  // DMMA fragments of this lane: channels 8 mt + (lane >> 2) (A operand), column species
  // 8 nt + (lane >> 2) (B operand); channel values of an edge are strided by nsp / ndp
  const int nch = S.vx_nch;
  const int bcol = lane >> 2, kq = lane & 3;
  int aoff[MT], astr[MT];   // byte offset of (edge 2 kq, this lane's channel) in the row's values; bytes per edge
  bool adyn[MT];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const int ch = 8 * mt + (lane >> 2);
    adyn[mt] = ch >= S.vx_nstat && ch < nch;   // padding channels read the static block (values unused)
    const int str = adyn[mt] ? NDP : NSP;
    aoff[mt] = ((adyn[mt] ? ch - S.vx_nstat : (ch < nch ? ch : 0)) + 2 * kq * str) * 8 + (adyn[mt] ? a.s_dyn : 0);
    astr[mt] = str * 8;
  }
  // epilogue plan of this lane, in registers for the whole launch
  double ecoef[EPL];
  int epos[EPL];
#pragma unroll
  for (int u = 0; u < EPL; ++u) {
    const int w = a.ep_pos[lane * EPL + u];
    ecoef[u] = a.ep_coef[lane * EPL + u] * (((w >> 16) & 1) ? S.inv_dt : 1.0);
    epos[u] = (w & 0xffff) * 8;
  }
  const int linfo = a.ep_lane[lane];
  const int esp = linfo & 0xff, elg = (linfo >> 8) & 0xf;
  const bool eleader = ((linfo >> 12) & 1) && esp < ns;
  // D^-1 y: lane (s, h) = (lane >> 1, lane & 1) adds half of the double2 pairs of row s (rows of
  // 96 bytes: the eight lanes of a quarter warp then read eight different groups of four banks)
  const int np2 = ns >> 1, hp = (np2 + 1) >> 1;
  const int dq0 = (lane & 1) ? hp : 0, dq1 = (lane & 1) ? np2 : hp;
  const int srow = lane >> 1;
  const bool sown = !(lane & 1) && srow < ns;   // the lane that finishes dof s of the row
  const int zoff = a.zslot * ns;

  int sidx = grp % NST;
  unsigned fph = (unsigned)((grp / NST) & 1);
  long long c0 = ((long long)blockIdx.x + (long long)grp * gridDim.x) * CH;
  const long long cstep = (long long)NG * gridDim.x * CH;
  long long dof = (c0 + j) * ns;   // first dof of this warp's row
  const long long dofstep = cstep * ns;
  for (int i = grp; i < nmine; i += NG, c0 += cstep, dof += dofstep) {
    const int nrow = (int)blockIdx.x + i * (int)gridDim.x == nchunks - 1
                         ? (int)(S.n_owned - (long long)(nchunks - 1) * CH) : CH;   // only the last chunk is short
    unsigned char* st = fb_sp_smem + a.o_stage0 + (size_t)sidx * a.stage_bytes;
    fb_mbar_wait(&full[sidx], fph);
    if (j < nrow) {
      const double* xs = (const double*)(st + a.s_xs);
      const int* rowptr_s = (const int*)(st + a.s_rowptr);
      const int rpj = rowptr_s[j], rpj1 = rowptr_s[j + 1], e0 = rowptr_s[0];
      const unsigned vi = ((const unsigned*)(st + a.s_vinfo))[j];
      const int lself = (int)((const unsigned*)(st + a.s_rowinfo))[j];
      const int rp0 = rpj - e0, deg = rpj1 - rpj;
      const unsigned rowbc = vi & 0xffffu;
      const bool colbc = (vi >> 16) & 1u;
      const unsigned short* lc = (const unsigned short*)(st + a.s_lcol) + rp0;
      // T[c][cs] = sum_k A^c[k] x[k][cs] on the fp64 tensor cores: a K step takes the edges
      // kb + {0, 2, 4, 6} + u (with six channel values per edge their rows are 96 bytes apart: the
      // four 32-byte segments a half warp reads fall into four different groups of eight banks)
      double acc[MT][NT][2];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
      const unsigned char* ap[MT];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) ap[mt] = st + aoff[mt] + rp0 * astr[mt];
      // Edges beyond the row (k >= deg) multiply the zero slot of xs: the A operand then reads finite
      // values of the next row.  Channels >= nch and species >= ns are tile padding: their products
      // land in entries of T no table refers to.
      const double* xb = xs + bcol;
      for (int kb = 0; kb < deg; kb += 8) {
        // the operand loads of two K steps (8 edges) first, then their DMMAs
        double av[2][MT], bv[2][NT];
        if (!colbc) {
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int k = kb + 2 * kq + u;
            const int lk = lc[k];
            const double* xk = xb + (k < deg ? lk * ns : zoff);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) av[u][mt] = *(const double*)(ap[mt] + u * astr[mt]);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) bv[u][nt] = xk[8 * nt];
          }
        } else {   // eliminated (Dirichlet) columns read as zero
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int k = kb + 2 * kq + u;
            const bool val = k < deg;
            const unsigned msk = val ? S.vinfo[S.colidx[rpj + k]] : 0u;
            const double* xk = xb + (val ? lc[k] * ns : zoff);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) av[u][mt] = *(const double*)(ap[mt] + u * astr[mt]);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
              const double xv = xk[8 * nt];
              bv[u][nt] = ((msk >> (8 * nt + bcol)) & 1u) ? 0.0 : xv;
            }
          }
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) ap[mt] += 8 * astr[mt];
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) fb_dmma(acc[mt][nt][0], acc[mt][nt][1], av[u][mt], bv[u][nt]);
      }
      // vectors of the fused epilogue: issued ahead of the tile store (no later load can pass it)
      const int t = j * ns + srow;
      double xself = 0.0, rr = 0.0, zz = 0.0, vv = 0.0, dsc = 0.0;
      if (sown) {
        xself = xs[lself * ns + srow];
        if (MODE >= 1) dsc = ((const double*)(st + a.s_dscale))[t];
        if (MODE == 2) {
          rr = ((const double*)(st + a.s_r))[t];
          if (a.last) vv = ((const double*)(st + a.s_vin))[t];
          else zz = ((const double*)(st + a.s_z))[t];
        }
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          double2 v2; v2.x = acc[mt][nt][0]; v2.y = acc[mt][nt][1];
          *(double2*)(part + ((mt * NT + nt) * 32 + lane) * 2) = v2;   // fragment-major: no bank conflicts
        }
      __syncwarp();
      // y_s = sum_e coef_e T[combo_e]: the lanes of a species' group add their entries, xor shuffles
      // join them (lanes of smaller groups ignore what they receive)
      double sum = 0.0, sum2 = 0.0;
#pragma unroll
      for (int u = 0; u < EPL; u += 2) {
        sum += ecoef[u] * *(const double*)((const unsigned char*)part + epos[u]);
        if (u + 1 < EPL) sum2 += ecoef[u + 1] * *(const double*)((const unsigned char*)part + epos[u + 1]);
      }
      sum += sum2;
      {
        double o1 = __shfl_xor_sync(FULL, sum, 1);
        if (elg >= 1) sum += o1;
        o1 = __shfl_xor_sync(FULL, sum, 2);
        if (elg >= 2) sum += o1;
        o1 = __shfl_xor_sync(FULL, sum, 4);
        if (elg >= 3) sum += o1;
      }
      if (eleader) {
        if ((rowbc >> esp) & 1u) sum = xs[lself * ns + esp];   // identity row of a Dirichlet dof
        if (MODE == 0) a.y[dof + esp] = sum;
        else yrow[esp] = sum;
      }
      if (MODE >= 1) {
        __syncwarp();
        double pa = 0.0, pb = 0.0;
        if (srow < ns) {
          // row s of the inverse block: 4-byte mantissas (rows of 48 bytes at ns = 12: the 16 lanes of a
          // half warp read 16 different bank pairs), the row's scale is applied to the sum
          const float2* dv = (const float2*)(st + a.s_dinv) + (size_t)(j * ns + srow) * np2;
          const double2* y2 = (const double2*)yrow;
#pragma unroll
          for (int q = 0; q < 2 * NT; ++q)
            if (dq0 + q < dq1) {
              const float2 d2 = dv[dq0 + q];
              const double2 yv = y2[dq0 + q];
              pa += (double)d2.x * yv.x;
              pb += (double)d2.y * yv.y;
            }
        }
        double p = pa + pb;
        p += __shfl_xor_sync(FULL, p, 1);
        p *= dsc;
        if (sown) {
          if (MODE == 1) a.y[dof + srow] = p;
          else {
            const double rn = rr - p;
            if (a.last) a.wout[dof + srow] = vv - rn;
            else {
              const double dn = a.ca * xself + a.cb * rn;
              a.r[dof + srow] = rn;
              a.dout[dof + srow] = dn;
              a.z[dof + srow] = zz + dn;
            }
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) fb_mbar_arrive(&empty[sidx]);   // this warp is done with the stage
    sidx += NG;
    if (sidx >= NST) { sidx -= NST; fph ^= 1u; }
  }
}
