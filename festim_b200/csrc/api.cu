// C ABI of libfestim_b200.so: context lifetime, problem definition, per-step data.
// See include/festim_b200.h for the contract and INTEGRATION.md for the binding.
#include <algorithm>
#include <cmath>
#include <utility>
#include <cstring>

#include "fb_internal.cuh"
#include "spmv_channels.cuh"

// spec section ids / int slots, shared with festim_b200/lowering.py
enum {
  S_INTS = 1, S_DIFF_D0, S_DIFF_SLOT, S_DIFF_ENERGIES, S_MASS_DT, S_MASS_C, S_PAIR_NC, S_PAIR_PP,
  S_PAIR_CS, S_PAIR_IDX, S_PROG_OPS, S_PROG_IARGS, S_PROG_FARGS, S_PROG_OFFSETS, S_VTERM_HDR,
  S_FTERM_HDR, S_ROW_SPECIES, S_ROW_COEF, S_DERIV_SPECIES, S_DERIV_PROG, S_QUAD_V, S_QUAD_S,
  S_QUAD_BARY, S_QUAD_W, S_ADV, S_DIFF_PROGS, S_DIFF_UPROG, S_RSLOT_E, S_FAC_HDR, S_FAC_A0,
  S_FAC_SPECIES, S_FAC_COEF, S_MONO, S_MONO_COEF, S_FC_PTR, S_FC_MONO, S_FC_SIGN, S_JC_PTR, S_JC,
  S_JC_COEF, S_WTAB, S_WTAB_OFF, S_STRUCT_INTS, S_COMBO, S_PAIR_ROW,
  S_VX_INTS, S_VX_CQ_PTR, S_VX_CQ, S_VX_CH_PTR, S_VX_CH, S_VX_JT_PTR, S_VX_JT, S_VX_JT_COEF,
  S_VX_FT_PTR, S_VX_FT, S_VX_FT_COEF, S_VX_SP_COMBO, S_VX_SP_PTR, S_VX_SP, S_VX_SP_COEF, S_MAX
};
enum {
  I_TDIM = 0, I_NS, I_NVTAGS, I_NSTAGS, I_TRANSIENT, I_TMODE, I_NPROGRAMS, I_NSCALARS, I_NFIELDS,
  I_NVTERMS, I_NFTERMS, I_NADV, I_NPAIRS, I_NDIFFE, I_DT_SCALAR, I_NSLOTS
};
static const long long FB_MAGIC = 0xFB2000001LL;

// element -> row-slot map of the gather kernels: for every (vertex v, incident cell) entry of
// the vertex->cell adjacency, the position of each cell vertex in v's sorted column list
__global__ void k_fill_slots(int d1, long long n_vertices, const int* __restrict__ cells,
                             const int* __restrict__ rowptr, const int* __restrict__ colidx,
                             const int* __restrict__ v2c_ptr, const int* __restrict__ v2c_idx,
                             unsigned* __restrict__ out) {
  const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (v >= n_vertices) return;
  const int r0 = rowptr[v], r1 = rowptr[v + 1];
  for (int e = v2c_ptr[v]; e < v2c_ptr[v + 1]; ++e) {
    const long long c = v2c_idx[e] / d1;
    unsigned packed = 0;
    for (int j = 0; j < d1; ++j) {
      const int col = cells[c * d1 + j];
      int lo = r0, hi = r1 - 1;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (colidx[mid] < col) lo = mid + 1; else hi = mid;
      }
      packed |= (unsigned)((lo - r0) & 0xff) << (8 * j);
    }
    out[e] = packed;
  }
}


extern "C" int fb_create(int device, void* cuda_stream, fb_ctx** out) {
  if (!out) return -1;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return -3;  // no GPU: fail loudly
  if (device < 0 || device >= count) return -3;
  if (cudaSetDevice(device) != cudaSuccess) return -3;
  fb_ctx* ctx = new fb_ctx();
  ctx->device = device;
  ctx->stream = (cudaStream_t)cuda_stream;
  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  for (int i = 0; i < 8; ++i) cudaEventCreate(&ctx->ev[i]);
  *out = ctx;
  return 0;
}

extern "C" void fb_destroy(fb_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (void* p : ctx->owned) cudaFree(p);
  if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
  for (int i = 0; i < 8; ++i) cudaEventDestroy(ctx->ev[i]);
  for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
  delete ctx;
}

extern "C" const char* fb_last_error(fb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" int64_t fb_launch_count(fb_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" int64_t fb_num_dofs(fb_ctx* ctx) { return ctx ? ctx->S.n_dofs : 0; }
extern "C" int64_t fb_jacobian_nnz(fb_ctx* ctx) { return ctx ? ctx->nnzb * ctx->S.npairs : 0; }

extern "C" int fb_set_mesh(fb_ctx* ctx, int tdim, int64_t nv, int64_t nc, const double* coords,
                           const int32_t* cells, const int32_t* cell_tag, int64_t nf,
                           const int32_t* bfacets, const int32_t* bfacet_cell,
                           const int32_t* bfacet_tag) {
  if (!ctx) return -1;
  if (tdim < 1 || tdim > 3) return fb_fail(ctx, "tdim must be 1, 2 or 3");
  if (nv <= 0 || nc <= 0) return fb_fail(ctx, "empty mesh");
  if (nv * (long long)FB_MAX_NS >= (1LL << 31)) return fb_fail(ctx, "mesh too large for int32 local indices");
  cudaSetDevice(ctx->device);
  DevSpec& S = ctx->S;
  S.tdim = tdim; S.n_vertices = nv; S.n_cells = nc; S.n_bfacets = nf; S.n_owned = nv;
  int rc;
  if ((rc = fb_upload(ctx, &S.coords, coords, (size_t)nv * tdim))) return rc;
  if ((rc = fb_upload(ctx, &S.cells, cells, (size_t)nc * (tdim + 1)))) return rc;
  if ((rc = fb_upload(ctx, &S.cell_tag, cell_tag, (size_t)nc))) return rc;
  if ((rc = fb_upload(ctx, &S.bfacets, bfacets, (size_t)nf * tdim))) return rc;
  if ((rc = fb_upload(ctx, &S.bfacet_cell, bfacet_cell, (size_t)nf))) return rc;
  if ((rc = fb_upload(ctx, &S.bfacet_tag, bfacet_tag, (size_t)nf))) return rc;
  ctx->have_mesh = true;
  return 0;
}

extern "C" int fb_set_pattern(fb_ctx* ctx, const int32_t* rowptr, const int32_t* colidx, int64_t nnzb,
                              const int32_t* v2c_ptr, const int32_t* v2c_idx) {
  if (!ctx || !ctx->have_mesh) return ctx ? fb_fail(ctx, "fb_set_mesh first") : -1;
  cudaSetDevice(ctx->device);
  DevSpec& S = ctx->S;
  int rc;
  {   // padded: the row kernels bulk-copy FB_CHUNK + 4 row pointers at a time
    int* rp = nullptr;
    if ((rc = fb_alloc(ctx, &rp, (size_t)S.n_vertices + 1 + FB_CHUNK + 8))) return rc;
    FB_CHECK(ctx, cudaMemsetAsync(rp, 0, sizeof(int) * ((size_t)S.n_vertices + 1 + FB_CHUNK + 8), ctx->stream));
    FB_CHECK(ctx, cudaMemcpyAsync(rp, rowptr, sizeof(int) * ((size_t)S.n_vertices + 1), cudaMemcpyHostToDevice, ctx->stream));
    FB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    S.rowptr = rp;
  }
  if ((rc = fb_upload(ctx, &S.colidx, colidx, (size_t)nnzb))) return rc;
  if ((rc = fb_upload(ctx, &S.v2c_ptr, v2c_ptr, (size_t)S.n_vertices + 1))) return rc;
  if ((rc = fb_upload(ctx, &S.v2c_idx, v2c_idx, (size_t)S.n_cells * (S.tdim + 1)))) return rc;
  ctx->nnzb = nnzb;
  ctx->h_rowptr.assign(rowptr, rowptr + S.n_vertices + 1);
  ctx->h_colidx.assign(colidx, colidx + nnzb);
  // block-tridiagonal in index order?  (1D meshes ordered along x)
  bool tri = true;
  for (long long v = 0; v < S.n_vertices && tri; ++v)
    for (int k = rowptr[v]; k < rowptr[v + 1]; ++k)
      if (std::llabs((long long)colidx[k] - v) > 1) { tri = false; break; }
  ctx->tridiag = tri;
  int maxdeg = 0;
  for (long long v = 0; v < S.n_vertices; ++v)
    if (rowptr[v + 1] - rowptr[v] > maxdeg) maxdeg = rowptr[v + 1] - rowptr[v];
  S.v2c_slot = nullptr;
  if (maxdeg <= 256) {
    unsigned* slots = nullptr;
    if ((rc = fb_alloc(ctx, &slots, (size_t)S.n_cells * (S.tdim + 1)))) return rc;
    FB_LAUNCH(ctx, k_fill_slots, (unsigned)((S.n_vertices + 127) / 128), 128, 0, S.tdim + 1, S.n_vertices,
              S.cells, S.rowptr, S.colidx, S.v2c_ptr, S.v2c_idx, slots);
    FB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    S.v2c_slot = slots;
  }
  // chunk plan of the TMA-pipelined row kernels (spmv_channels.cuh)
  {
    const long long nv = S.n_vertices;
    const long long nchunks = (nv + FB_CHUNK - 1) / FB_CHUNK;
    std::vector<int> runptr(nchunks + 1, 0), runs, nx(nchunks, 0), lcolptr(nchunks + 1, 0);
    int rmax = 1;
    std::vector<unsigned short> lcol;
    std::vector<unsigned> rowinfo(((size_t)nv + FB_CHUNK + 15) & ~(size_t)15, 0u);
    std::vector<int> uniq;
    int nx_max = 0;
    bool fits16 = true;
    for (long long c = 0; c < nchunks; ++c) {
      const long long c0 = c * FB_CHUNK, c1 = std::min(nv, c0 + FB_CHUNK);
      const int e0 = rowptr[c0], e1 = rowptr[c1];
      uniq.assign(colidx + e0, colidx + e1);
      std::sort(uniq.begin(), uniq.end());
      uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
      // runs of consecutive vertex ids; the staged list is the concatenation of the runs
      for (size_t i = 0; i < uniq.size();) {
        size_t j = i + 1;
        while (j < uniq.size() && uniq[j] == uniq[j - 1] + 1) ++j;
        runs.push_back(uniq[i]); runs.push_back((int)(j - i)); runs.push_back((int)i);
        i = j;
      }
      runptr[c + 1] = (int)runs.size() / 3;
      rmax = std::max(rmax, runptr[c + 1] - runptr[c]);
      nx[c] = (int)uniq.size();
      nx_max = std::max(nx_max, nx[c]);
      if (uniq.size() > 65535) fits16 = false;
      lcolptr[c] = (int)lcol.size();
      for (int e = e0; e < e1; ++e)
        lcol.push_back((unsigned short)(std::lower_bound(uniq.begin(), uniq.end(), colidx[e]) - uniq.begin()));
      while (lcol.size() % 8) lcol.push_back(0);
      for (long long v = c0; v < c1; ++v) {
        const auto it = std::lower_bound(uniq.begin(), uniq.end(), (int)v);
        rowinfo[v] = (it != uniq.end() && *it == (int)v) ? (unsigned)(it - uniq.begin()) : 0u;
      }
    }
    lcolptr[nchunks] = (int)lcol.size();
    // few long runs per chunk (compact numbering): TMA-friendly; otherwise no plan and the row
    // kernels fall back to per-row gathers
    S.ck_nx_max = (fits16 && (double)runs.size() / 3 <= 40.0 * (double)nchunks) ? nx_max : 0;
    S.ck_rmax = rmax;
    // the first 32 runs of every chunk in fixed slots (fetched one chunk ahead by the producer
    // warp, one per lane); the rare chunk with more reads the rest from the CSR list
    std::vector<int> slots;   // [chunk][32][3], unused slots have length 0
    if (S.ck_nx_max > 0) {
      slots.assign((size_t)nchunks * 32 * 3, 0);
      for (long long c = 0; c < nchunks; ++c)
        for (int q = runptr[c]; q < runptr[c + 1] && q < runptr[c] + 32; ++q)
          for (int w = 0; w < 3; ++w) slots[((size_t)c * 32 + (q - runptr[c])) * 3 + w] = runs[3 * (size_t)q + w];
    }
    if ((rc = fb_upload(ctx, &S.ck_slots, slots.data(), slots.size()))) return rc;
    if ((rc = fb_upload(ctx, &S.ck_runptr, runptr.data(), runptr.size()))) return rc;
    if ((rc = fb_upload(ctx, &S.ck_runs, runs.data(), runs.size()))) return rc;
    if ((rc = fb_upload(ctx, &S.ck_nx, nx.data(), nx.size()))) return rc;
    if ((rc = fb_upload(ctx, &S.ck_lcolptr, lcolptr.data(), lcolptr.size()))) return rc;
    if ((rc = fb_upload(ctx, &S.ck_lcol, lcol.data(), lcol.size()))) return rc;
    if ((rc = fb_upload(ctx, &S.ck_rowinfo, rowinfo.data(), rowinfo.size()))) return rc;
  }
  ctx->have_pattern = true;
  return 0;
}

namespace {
struct Section { long long id, code, count, offset; };

template <typename T>
int upload_section(fb_ctx* ctx, const char* blob, size_t nbytes, const Section* secs, int id, int code,
                   const T** dst, long long* count_out = nullptr) {
  const Section& s = secs[id];
  if (s.id != id) { *dst = nullptr; if (count_out) *count_out = 0; return fb_upload(ctx, dst, (const T*)nullptr, 0); }
  if (s.code != code) return fb_fail(ctx, "spec section has the wrong dtype");
  if ((size_t)(s.offset + s.count * (long long)sizeof(T)) > nbytes) return fb_fail(ctx, "spec section out of bounds");
  if (count_out) *count_out = s.count;
  return fb_upload(ctx, dst, (const T*)(blob + s.offset), (size_t)s.count);
}
}  // namespace

extern "C" int fb_set_spec(fb_ctx* ctx, const void* blob_, size_t nbytes) {
  if (!ctx || !ctx->have_pattern) return ctx ? fb_fail(ctx, "fb_set_pattern first") : -1;
  cudaSetDevice(ctx->device);
  ctx->dyn_saved = false;
  const char* blob = (const char*)blob_;
  if (nbytes < 16) return fb_fail(ctx, "spec blob too small");
  long long hdr[2];
  memcpy(hdr, blob, 16);
  if (hdr[0] != FB_MAGIC) return fb_fail(ctx, "bad spec magic");
  const long long nsec = hdr[1];
  if (nsec < 1 || nsec > 64 || (size_t)(16 + nsec * 32) > nbytes) return fb_fail(ctx, "bad spec header");
  Section secs[S_MAX];
  for (int i = 0; i < S_MAX; ++i) secs[i] = Section{0, 0, 0, 0};
  for (long long i = 0; i < nsec; ++i) {
    Section s;
    memcpy(&s, blob + 16 + i * 32, 32);
    if (s.id <= 0 || s.id >= S_MAX) return fb_fail(ctx, "unknown spec section");
    secs[s.id] = s;
  }
  if (secs[S_INTS].id != S_INTS || secs[S_INTS].count < 16) return fb_fail(ctx, "spec has no int table");
  int ints[16];
  memcpy(ints, blob + secs[S_INTS].offset, sizeof(ints));
  DevSpec& S = ctx->S;
  if (ints[I_TDIM] != S.tdim) return fb_fail(ctx, "spec tdim differs from the mesh");
  S.ns = ints[I_NS];
  if (S.ns < 1 || S.ns > FB_MAX_NS) return fb_fail(ctx, "unsupported number of species");
  S.n_vtags = ints[I_NVTAGS]; S.n_stags = ints[I_NSTAGS];
  S.transient = ints[I_TRANSIENT]; S.T_mode = ints[I_TMODE];
  S.n_vterms = ints[I_NVTERMS]; S.n_fterms = ints[I_NFTERMS]; S.n_adv = ints[I_NADV];
  S.npairs = ints[I_NPAIRS]; S.n_diffE = ints[I_NDIFFE]; S.n_slots = ints[I_NSLOTS];
  S.n_fields = ints[I_NFIELDS];
  ctx->dt_scalar = ints[I_DT_SCALAR];
  ctx->n_programs = ints[I_NPROGRAMS];
  // J is symmetric positive definite when nothing but diffusion and mass contributes:
  // no advection and no point term with a solution derivative (reactions couple species
  // with different coefficients in each direction)
  ctx->symmetric = (S.n_adv == 0);
  for (int pass = 0; pass < 2; ++pass) {
    const Section& hs = secs[pass == 0 ? S_VTERM_HDR : S_FTERM_HDR];
    const int nterms = pass == 0 ? S.n_vterms : S.n_fterms;
    if (hs.id == 0) continue;
    const int* h = (const int*)(blob + hs.offset);
    for (int t = 0; t < nterms; ++t) {
      if (h[t * 8 + 5] > 0) ctx->symmetric = false;
      if (pass == 1 && h[t * 8 + 6]) ctx->facet_u_terms = true;
      if (pass == 0 && h[t * 8 + 6]) ctx->generic_u_vterms = true;
    }
  }
  S.n_dofs = S.n_vertices * S.ns;
  if (S.n_owned <= 0 || S.n_owned > S.n_vertices) S.n_owned = S.n_vertices;
  S.n_owned_dofs = S.n_owned * S.ns;
  const int n_scalars = ints[I_NSCALARS];
  int rc;
#define UP(id, code, field) if ((rc = upload_section(ctx, blob, nbytes, secs, id, code, &S.field))) return rc
  UP(S_DIFF_D0, 1, diff_D0); UP(S_DIFF_SLOT, 0, diff_slot); UP(S_DIFF_ENERGIES, 1, diff_energies);
  UP(S_DIFF_PROGS, 0, diff_progs); UP(S_MASS_DT, 1, mass_dt); UP(S_MASS_C, 1, mass_c);
  UP(S_PAIR_NC, 0, pair_nc); UP(S_PAIR_PP, 0, pair_pp); UP(S_PAIR_CS, 0, pair_cs); UP(S_PAIR_IDX, 0, pair_idx);
  UP(S_PROG_OPS, 0, prog_ops); UP(S_PROG_IARGS, 0, prog_iargs); UP(S_PROG_FARGS, 1, prog_fargs);
  UP(S_PROG_OFFSETS, 0, prog_offsets);
  UP(S_VTERM_HDR, 0, vterm_hdr); UP(S_FTERM_HDR, 0, fterm_hdr);
  UP(S_ROW_SPECIES, 0, row_species); UP(S_ROW_COEF, 1, row_coef);
  UP(S_DERIV_SPECIES, 0, deriv_species); UP(S_DERIV_PROG, 0, deriv_prog);
  UP(S_QUAD_V, 0, quad_v); UP(S_QUAD_S, 0, quad_s); UP(S_QUAD_BARY, 1, quad_bary); UP(S_QUAD_W, 1, quad_w);
  UP(S_ADV, 0, adv); UP(S_DIFF_UPROG, 0, diff_uprog);
  S.n_rslots = S.n_factors = S.n_monos = S.n_fac_terms = 0;
  if (secs[S_STRUCT_INTS].id == S_STRUCT_INTS) {
    const int* si = (const int*)(blob + secs[S_STRUCT_INTS].offset);
    S.n_rslots = si[0]; S.n_factors = si[1]; S.n_monos = si[2];
    S.n_fac_terms = (int)secs[S_FAC_COEF].count;
    UP(S_RSLOT_E, 1, rslot_E); UP(S_FAC_HDR, 0, fac_hdr); UP(S_FAC_A0, 1, fac_a0);
    {
      const int* fh = (const int*)(blob + secs[S_FAC_HDR].offset);
      ctx->fac_scalar_idx.clear();
      for (int f = 0; f < S.n_factors; ++f)
        if (fh[4 * f]) ctx->fac_scalar_idx.push_back(fh[4 * f + 1]);
      ctx->fac_scalar_last.assign(ctx->fac_scalar_idx.size(), 0.0);
    }
    UP(S_FAC_SPECIES, 0, fac_species); UP(S_FAC_COEF, 1, fac_coef);
    ctx->symmetric = false;
  }
  S.maxdeg = 0;
  for (long long v = 0; v < S.n_vertices; ++v) {
    const int dg = ctx->h_rowptr[v + 1] - ctx->h_rowptr[v];
    if (dg > S.maxdeg) S.maxdeg = dg;
  }
  auto magic = [](int d) { return d <= 1 ? 0u : (unsigned)(((1ull << 32) + d - 1) / d); };
  S.vx_m_ns = magic(S.ns); S.vx_m_np = magic(S.npairs); S.vx_m_nf = magic(S.n_factors);
  // ---- channel assembly tables (assemble_edge.cuh) ---------------------------------------------
  S.vx_on = 0;
  if (secs[S_VX_INTS].id == S_VX_INTS && S.v2c_slot != nullptr && S.maxdeg < 255 && !getenv("FB_NO_VTX")) {
    const int* vi = (const int*)(blob + secs[S_VX_INTS].offset);
    S.vx_nsym = vi[1]; S.vx_nsym3 = vi[2]; S.vx_rec = vi[3]; S.vx_nbw = vi[4]; S.vx_nch = vi[5];
    if (vi[6] != S.n_factors) return fb_fail(ctx, "channel tables: factor count mismatch");
    S.vx_nstat = vi[7]; S.vx_ndyn = vi[8]; S.vx_nts = vi[9]; S.vx_ncombo = vi[10];
    UP(S_PAIR_ROW, 0, pair_row); UP(S_VX_CQ_PTR, 0, vx_cq_ptr); UP(S_VX_CQ, 0, vx_cq);
    UP(S_VX_SP_COMBO, 0, vx_sp_combo); UP(S_VX_SP_PTR, 0, vx_sp_ptr); UP(S_VX_SP, 0, vx_sp); UP(S_VX_SP_COEF, 1, vx_sp_coef);
    auto sec_i = [&](int id) { return (const int*)(blob + secs[id].offset); };
    auto sec_d = [&](int id) { return (const double*)(blob + secs[id].offset); };
    const int ns = S.ns, np = S.npairs, nch = S.vx_nch;
    const int* ch = sec_i(S_VX_CH);
    const int* jtp = sec_i(S_VX_JT_PTR); const int* jt = sec_i(S_VX_JT); const double* jtc = sec_d(S_VX_JT_COEF);
    const int* ftp = sec_i(S_VX_FT_PTR); const int* ft = sec_i(S_VX_FT); const double* ftc = sec_d(S_VX_FT_COEF);
    if (secs[S_VX_CH].count != 8LL * nch && !(nch == 1 && secs[S_VX_CH].count == 0))
      return fb_fail(ctx, "channel tables: channel list size mismatch");
    if (nch > 255 || np > 255) return fb_fail(ctx, "channel tables: too many channels / species pairs");
    S.vx_nft = ftp[ns]; S.vx_njt = jtp[np];
    ctx->n_sp_ent = (int)(secs[S_VX_SP].count / 2);
    {
      const SpmvEpPlan P = fb_spmv_ep_plan(ns, sec_i(S_VX_SP_PTR), sec_i(S_VX_SP), sec_d(S_VX_SP_COEF), sec_i(S_VX_SP_COMBO));
      ctx->ep_epl = P.epl;
      if (P.epl) {
        int rc2;
        if ((rc2 = fb_upload(ctx, &ctx->d_ep_coef, P.coef.data(), P.coef.size()))) return rc2;
        if ((rc2 = fb_upload(ctx, &ctx->d_ep_pos, P.pos.data(), P.pos.size()))) return rc2;
        if ((rc2 = fb_upload(ctx, &ctx->d_ep_lane, P.lane.data(), P.lane.size()))) return rc2;
      }
    }
    std::vector<int> itab;
    std::vector<double> dtab;
    S.vx_o_ch = (int)itab.size(); itab.insert(itab.end(), ch, ch + secs[S_VX_CH].count);
    S.vx_o_ftptr = (int)itab.size(); itab.insert(itab.end(), ftp, ftp + ns + 1);
    S.vx_o_ft = (int)itab.size(); itab.insert(itab.end(), ft, ft + 3 * S.vx_nft);
    S.vx_o_ftc = (int)dtab.size(); dtab.insert(dtab.end(), ftc, ftc + S.vx_nft);
    S.vx_o_jtptr = (int)itab.size(); itab.insert(itab.end(), jtp, jtp + np + 1);
    S.vx_o_jt = (int)itab.size(); itab.insert(itab.end(), jt, jt + 2 * S.vx_njt);
    S.vx_o_jtc = (int)dtab.size(); dtab.insert(dtab.end(), jtc, jtc + S.vx_njt);
    S.vx_o_jtpair = (int)itab.size();
    for (int p = 0; p < np; ++p)
      for (int e = jtp[p]; e < jtp[p + 1]; ++e) itab.push_back(p);
    // solution-dependent channels grouped by the edge tensor they contract: (channel, factor column)
    S.vx_o_dynptr = (int)itab.size();
    {
      std::vector<int> ptr(S.vx_nts + 1, 0), lst;
      for (int ts = 0; ts < S.vx_nts; ++ts) {
        for (int c = S.vx_nstat; c < nch; ++c)
          if (ch[8 * c + 5] == ts) { lst.push_back(c); lst.push_back(ch[8 * c + 3]); }
        ptr[ts + 1] = (int)lst.size() / 2;
      }
      if ((int)lst.size() / 2 != S.vx_ndyn) return fb_fail(ctx, "channel tables: dynamic channels without an edge tensor");
      itab.insert(itab.end(), ptr.begin(), ptr.end());
      S.vx_o_dyn = (int)itab.size();
      itab.insert(itab.end(), lst.begin(), lst.end());
    }
    {   // diagonal-block entries grouped by channel (k_bjacobi_setup_slab): A[s][c] += coef * value(channel)
      const int* prow = sec_i(S_PAIR_ROW); const int* pcs = sec_i(S_PAIR_CS);
      std::vector<int> bptr(nch + 1, 0), bpos;
      std::vector<double> bcoef;
      for (int c = 0; c < nch; ++c) {
        for (int p = 0; p < np; ++p)
          for (int e = jtp[p]; e < jtp[p + 1]; ++e)
            if (jt[2 * e] == c) { bpos.push_back((prow[p] * ns + pcs[p]) | ((jt[2 * e + 1] & 1) << 16)); bcoef.push_back(jtc[e]); }
        bptr[c + 1] = (int)bpos.size();
      }
      if ((rc = fb_upload(ctx, &ctx->d_bj_ptr, bptr.data(), bptr.size()))) return rc;
      if ((rc = fb_upload(ctx, &ctx->d_bj_pos, bpos.data(), bpos.size()))) return rc;
      if ((rc = fb_upload(ctx, &ctx->d_bj_coef, bcoef.data(), bcoef.size()))) return rc;
    }
    S.vx_itab_n = (int)itab.size(); S.vx_dtab_n = (int)dtab.size();
    if ((rc = fb_upload(ctx, &S.vx_itab, itab.data(), itab.size()))) return rc;
    if ((rc = fb_upload(ctx, &S.vx_dtab, dtab.data(), dtab.size()))) return rc;
    S.vx_nsp = (S.vx_nstat + 1) & ~1;
    S.vx_ndp = (S.vx_ndyn + 1) & ~1;
    S.vx_m_nsp = magic(S.vx_nsp); S.vx_m_ndp = magic(S.vx_ndp);
    S.nnzb = ctx->nnzb;
    double* p = nullptr;
    if ((rc = fb_alloc(ctx, &p, (size_t)S.n_cells * S.vx_rec))) return rc;
    S.cellrec = p;
    if ((rc = fb_alloc(ctx, &p, (size_t)ctx->nnzb * S.vx_nsp))) return rc;
    S.et_stat = p;
    if ((rc = fb_alloc(ctx, &p, (size_t)ctx->nnzb * S.vx_ndp))) return rc;
    FB_CHECK(ctx, cudaMemsetAsync(p, 0, sizeof(double) * (size_t)ctx->nnzb * S.vx_ndp, ctx->stream));
    S.et_dyn = p;
    if ((rc = fb_alloc(ctx, &p, (size_t)2 * ctx->nnzb * S.vx_nts))) return rc;
    S.et_tvw = p;
    unsigned* vinfo = nullptr;
    const size_t nvi = ((size_t)S.n_vertices + FB_CHUNK + 15) & ~(size_t)15;   // whole chunks are bulk-copied
    if ((rc = fb_alloc(ctx, &vinfo, nvi))) return rc;
    FB_CHECK(ctx, cudaMemsetAsync(vinfo, 0, sizeof(unsigned) * nvi, ctx->stream));
    S.vinfo = vinfo;
    S.et_tz_total = 0; S.et_tz_chunk_max = 0; S.et_tz = nullptr; S.et_tzptr = nullptr; S.et_tzslot = nullptr; S.et_tzw = nullptr;
    if (S.vx_nts > 0 && (rc = fbi_et_symbolic(ctx))) return rc;
    S.vx_on = 1;
    // the Krylov solvers work on the channel operator itself when it is the smaller format
    ctx->op_channels = (S.vx_nsp + S.vx_ndp < np) && !ctx->tridiag && !ctx->facet_u_terms &&
                       !getenv("FB_NO_CHANNEL_OP");
  }
  S.n_udiff = 0;
  if (secs[S_DIFF_UPROG].id == S_DIFF_UPROG) {
    const int* up = (const int*)(blob + secs[S_DIFF_UPROG].offset);
    for (long long i = 0; i < secs[S_DIFF_UPROG].count / 3; ++i)
      if (up[3 * i] >= 0) { S.n_udiff++; ctx->symmetric = false; }
  }
#undef UP
  // mutable per-step buffers
  if ((rc = fb_alloc(ctx, &ctx->d_scalars, (size_t)n_scalars + 1))) return rc;
  if ((rc = fb_alloc(ctx, &ctx->d_fields, (size_t)S.n_fields + 1))) return rc;
  if ((rc = fb_alloc(ctx, &ctx->d_velocities, (size_t)S.n_adv + 1))) return rc;
  if ((rc = fb_alloc(ctx, &ctx->d_bc_map, (size_t)S.n_dofs))) return rc;
  if ((rc = fb_alloc(ctx, &ctx->d_dbar, (size_t)2 * (S.n_slots ? S.n_slots : 1) * S.n_cells))) return rc;
  if ((rc = fb_alloc(ctx, &ctx->d_load, (size_t)S.n_dofs))) return rc;
  FB_CHECK(ctx, cudaMemsetAsync(ctx->d_bc_map, 0xFF, sizeof(int) * S.n_dofs, ctx->stream));
  FB_CHECK(ctx, cudaMemsetAsync(ctx->d_load, 0, sizeof(double) * S.n_dofs, ctx->stream));
  FB_CHECK(ctx, cudaMemsetAsync(ctx->d_scalars, 0, sizeof(double) * (n_scalars + 1), ctx->stream));
  ctx->h_fields.assign(S.n_fields + 1, nullptr);
  ctx->h_velocities.assign(S.n_adv + 1, nullptr);
  S.scalars = ctx->d_scalars; S.fields = ctx->d_fields; S.velocities = ctx->d_velocities;
  S.bc_map = ctx->d_bc_map; S.bc_vals = nullptr; S.dbar = ctx->d_dbar; S.load = ctx->d_load;
  S.T_const = 300.0; S.T_nodal = nullptr; S.inv_dt = 0.0;
  ctx->have_spec = true;
  return 0;
}

__global__ void k_fill_bc_map(long long n, const long long* __restrict__ dofs, int* __restrict__ map) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) map[dofs[i]] = (int)i;
}

extern "C" int fb_set_dirichlet(fb_ctx* ctx, int64_t n_bc, const int64_t* bc_dofs) {
  if (!ctx || !ctx->have_spec) return ctx ? fb_fail(ctx, "fb_set_spec first") : -1;
  cudaSetDevice(ctx->device);
  DevSpec& S = ctx->S;
  for (int64_t i = 0; i < n_bc; ++i)
    if (bc_dofs[i] < 0 || bc_dofs[i] >= S.n_dofs) return fb_fail(ctx, "Dirichlet dof out of range");
  FB_CHECK(ctx, cudaMemsetAsync(ctx->d_bc_map, 0xFF, sizeof(int) * S.n_dofs, ctx->stream));
  ctx->n_bc = n_bc;
  ctx->coarse_masks_dirty = true;
  if (n_bc == 0) return fbi_update_vinfo(ctx);
  const long long* d = nullptr;
  int rc = fb_upload(ctx, &d, (const long long*)bc_dofs, (size_t)n_bc);
  if (rc) return rc;
  FB_LAUNCH(ctx, k_fill_bc_map, (unsigned)((n_bc + 255) / 256), 256, 0, (long long)n_bc, d, ctx->d_bc_map);
  if ((rc = fbi_update_vinfo(ctx))) return rc;
  FB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int fb_set_scalars(fb_ctx* ctx, int n, const double* scalars) {
  if (!ctx || !ctx->have_spec) return ctx ? fb_fail(ctx, "fb_set_spec first") : -1;
  cudaSetDevice(ctx->device);
  if (n > 0) {
    FB_CHECK(ctx, cudaMemcpyAsync(ctx->d_scalars, scalars, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    FB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));  // scalars is caller stack/heap memory
  }
  if (ctx->dt_scalar >= 0 && ctx->dt_scalar < n) ctx->S.inv_dt = 1.0 / scalars[ctx->dt_scalar];
  // the stored solution-dependent channels depend on the scalars the affine factors read
  for (size_t q = 0; q < ctx->fac_scalar_idx.size(); ++q) {
    const int k = ctx->fac_scalar_idx[q];
    const double v = k < n ? scalars[k] : 0.0;
    if (!(v == ctx->fac_scalar_last[q])) ctx->dyn_saved = false;
    ctx->fac_scalar_last[q] = v;
  }
  return 0;
}

extern "C" int fb_set_temperature(fb_ctx* ctx, double T_const, const double* T_nodal) {
  if (!ctx || !ctx->have_spec) return ctx ? fb_fail(ctx, "fb_set_spec first") : -1;
  if (ctx->S.T_mode && !T_nodal) return fb_fail(ctx, "spec expects a nodal temperature");
  ctx->S.T_const = T_const;
  ctx->S.T_nodal = T_nodal;
  return 0;
}

static int push_ptrs(fb_ctx* ctx, const double** dev, const std::vector<const double*>& host) {
  FB_CHECK(ctx, cudaMemcpyAsync(dev, host.data(), sizeof(double*) * host.size(), cudaMemcpyHostToDevice, ctx->stream));
  FB_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int fb_set_field(fb_ctx* ctx, int k, const double* nodal) {
  if (!ctx || !ctx->have_spec) return ctx ? fb_fail(ctx, "fb_set_spec first") : -1;
  if (k < 0 || k >= ctx->S.n_fields) return fb_fail(ctx, "field index out of range");
  ctx->h_fields[k] = nodal;
  return push_ptrs(ctx, ctx->d_fields, ctx->h_fields);
}

extern "C" int fb_set_velocity(fb_ctx* ctx, int k, const double* nodal) {
  if (!ctx || !ctx->have_spec) return ctx ? fb_fail(ctx, "fb_set_spec first") : -1;
  if (k < 0 || k >= ctx->S.n_adv) return fb_fail(ctx, "velocity index out of range");
  ctx->h_velocities[k] = nodal;
  return push_ptrs(ctx, ctx->d_velocities, ctx->h_velocities);
}

extern "C" int fb_set_dirichlet_values(fb_ctx* ctx, const double* bc_vals) {
  if (!ctx || !ctx->have_spec) return ctx ? fb_fail(ctx, "fb_set_spec first") : -1;
  ctx->S.bc_vals = bc_vals;
  return 0;
}

extern "C" int fb_update_load(fb_ctx* ctx) {
  if (!ctx || !ctx->have_spec) return ctx ? fb_fail(ctx, "fb_set_spec first") : -1;
  cudaSetDevice(ctx->device);
  return fbi_update_coeffs(ctx);
}

extern "C" int fb_assemble(fb_ctx* ctx, const double* u, const double* u_n, double* F, double* Jvals,
                           int flags) {
  if (!ctx) return -1;
  cudaSetDevice(ctx->device);
  if (ctx->n_bc > 0 && !ctx->S.bc_vals) return fb_fail(ctx, "Dirichlet values not set");
  return fbi_assemble(ctx, u, u_n, F, Jvals, flags);
}

extern "C" int fb_spmv(fb_ctx* ctx, const double* Jvals, const double* x, double* y) {
  if (!ctx || !ctx->have_spec) return ctx ? fb_fail(ctx, "fb_set_spec first") : -1;
  cudaSetDevice(ctx->device);
  return fbi_spmv(ctx, Jvals, x, y);
}

extern "C" int fb_linear_solve(fb_ctx* ctx, const double* Jvals, const double* b, double* y, int ksp,
                               int pc, double rtol, double atol, int max_it, int* iters, double* relres) {
  if (!ctx || !ctx->have_spec) return ctx ? fb_fail(ctx, "fb_set_spec first") : -1;
  cudaSetDevice(ctx->device);
  return fbi_linear_solve(ctx, Jvals, b, y, ksp, pc, rtol, atol, max_it, iters, relres);
}

// vertices [0, n_owned) are owned by this rank; the rest are ghosts whose values the
// caller refreshes (halo exchange) before fb_assemble / fb_spmv.  Kernels then only
// produce owned rows and reductions only cover owned entries.
extern "C" int fb_set_owned(fb_ctx* ctx, int64_t n_owned_vertices) {
  if (!ctx || !ctx->have_mesh) return ctx ? fb_fail(ctx, "fb_set_mesh first") : -1;
  if (n_owned_vertices <= 0 || n_owned_vertices > ctx->S.n_vertices) return fb_fail(ctx, "bad owned count");
  ctx->S.n_owned = n_owned_vertices;
  ctx->S.n_owned_dofs = n_owned_vertices * (ctx->S.ns > 0 ? ctx->S.ns : 1);
  return 0;
}

// Compact per-SM partition for the persistent CG (vertices are numbered so that block b
// owns the contiguous range [rng[b], rng[b+1])): halo vertex list of every block and the
// 16-bit local column (owned first, then halo) of every block entry of its rows.
extern "C" int fb_set_sm_partition(fb_ctx* ctx, int nblocks, const int32_t* rng, const int32_t* halo_ptr,
                                   const int32_t* halo_idx, const uint16_t* lcol, int64_t n_lcol) {
  if (!ctx || !ctx->have_spec) return ctx ? fb_fail(ctx, "fb_set_spec first") : -1;
  cudaSetDevice(ctx->device);
  if (nblocks != ctx->sm_count) return fb_fail(ctx, "one block per SM expected");
  if (n_lcol != ctx->nnzb) return fb_fail(ctx, "lcol must have one entry per block entry");
  const DevSpec& S = ctx->S;
  int rc;
  if ((rc = fb_upload(ctx, &ctx->d_blk_rng, rng, (size_t)nblocks + 1))) return rc;
  if ((rc = fb_upload(ctx, &ctx->d_halo_ptr, halo_ptr, (size_t)nblocks + 1))) return rc;
  if ((rc = fb_upload(ctx, &ctx->d_halo_idx, halo_idx, (size_t)halo_ptr[nblocks]))) return rc;
  if ((rc = fb_upload(ctx, &ctx->d_lcol, (const unsigned short*)lcol, (size_t)n_lcol))) return rc;
  ctx->part_max_entries = ctx->part_max_bentries = ctx->part_max_su = ctx->part_max_rows = 0;
  ctx->part_min_rows = 1LL << 60;
  for (int b = 0; b < nblocks; ++b) {
    const long long nown = rng[b + 1] - rng[b], nhalo = halo_ptr[b + 1] - halo_ptr[b];
    const long long be = ctx->h_rowptr[rng[b + 1]] - ctx->h_rowptr[rng[b]];
    if (be * S.npairs > ctx->part_max_entries) ctx->part_max_entries = be * S.npairs;
    if (be > ctx->part_max_bentries) ctx->part_max_bentries = be;
    if ((nown + nhalo) * S.ns > ctx->part_max_su) ctx->part_max_su = (nown + nhalo) * S.ns;
    if (nown * S.ns > ctx->part_max_rows) ctx->part_max_rows = nown * S.ns;
    if (nown * S.ns < ctx->part_min_rows) ctx->part_min_rows = nown * S.ns;
  }
  // coarse space of the two-level preconditioner (one constant per block): block id of
  // every halo vertex, and per block the entries coupling to each neighbour block
  ctx->part_owned_only = rng[nblocks] != S.n_vertices;
  if (rng[nblocks] != S.n_vertices && rng[nblocks] != S.n_owned)
    return fb_fail(ctx, "blocks must cover all vertices or exactly the owned ones");
  if (S.ns == 1 && nblocks <= 160) {
    auto block_of = [&](int w) {
      int lo = 0, hi = nblocks;  // rng[lo] <= w < rng[hi]
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (rng[mid] <= w) lo = mid; else hi = mid; }
      return lo;
    };
    std::vector<unsigned short> hb((size_t)halo_ptr[nblocks] + 1);
    const int n_cov = rng[nblocks];  // vertices beyond are ghosts of a partitioned mesh: not in the coarse space
    for (int h = 0; h < halo_ptr[nblocks]; ++h)
      hb[h] = halo_idx[h] < n_cov ? (unsigned short)block_of(halo_idx[h]) : (unsigned short)159;
    std::vector<int> cn_ptr(nblocks + 1, 0), cn_blk, cn_eptr(1, 0), cn_ent;
    std::vector<std::pair<int, int>> off;  // (neighbour block, entry)
    for (int b = 0; b < nblocks; ++b) {
      off.clear();
      for (int e = ctx->h_rowptr[rng[b]]; e < ctx->h_rowptr[rng[b + 1]]; ++e) {
        const int w = ctx->h_colidx[e];
        if ((w < rng[b] || w >= rng[b + 1]) && w < n_cov) off.emplace_back(block_of(w), e);
      }
      std::stable_sort(off.begin(), off.end(),
                       [](const std::pair<int, int>& x, const std::pair<int, int>& y) { return x.first < y.first; });
      for (size_t k = 0; k < off.size(); ++k) {
        if (k == 0 || off[k].first != off[k - 1].first) {
          if (k) cn_eptr.push_back((int)cn_ent.size());
          cn_blk.push_back(off[k].first);
        }
        cn_ent.push_back(off[k].second);
      }
      if (!off.empty()) cn_eptr.push_back((int)cn_ent.size());
      cn_ptr[b + 1] = (int)cn_blk.size();
    }
    if ((rc = fb_upload(ctx, &ctx->d_halo_blk, hb.data(), hb.size()))) return rc;
    if ((rc = fb_alloc(ctx, &ctx->d_halo_blk_eff, hb.size()))) return rc;
    ctx->part_nhalo = halo_ptr[nblocks];
    ctx->coarse_masks_dirty = true;
    if ((rc = fb_upload(ctx, &ctx->d_cn_ptr, cn_ptr.data(), cn_ptr.size()))) return rc;
    if ((rc = fb_upload(ctx, &ctx->d_cn_blk, cn_blk.data(), cn_blk.size()))) return rc;
    if ((rc = fb_upload(ctx, &ctx->d_cn_eptr, cn_eptr.data(), cn_eptr.size()))) return rc;
    if ((rc = fb_upload(ctx, &ctx->d_cn_ent, cn_ent.data(), cn_ent.size()))) return rc;
    const size_t nco = (size_t)2 * nblocks * nblocks + 2 * nblocks + 8;
    if ((rc = fb_alloc(ctx, &ctx->d_coarse, nco))) return rc;
    FB_CHECK(ctx, cudaMemset(ctx->d_coarse, 0, nco * sizeof(double)));  // E keeps its structural zeros
  }
  ctx->part_blocks = nblocks;
  return 0;
}
