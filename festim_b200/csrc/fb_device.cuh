// Device-side declarations shared by the ahead-of-time build (nvcc) and the
// run-time specialisation of the element kernels (NVRTC, festim_b200/jit.py).
// No host headers here.
#pragma once

#ifndef FB_ASSEMBLE_F
#define FB_ASSEMBLE_F 1
#define FB_ASSEMBLE_J 2
#endif
#define FB_MAX_NS 16
#define FB_MAX_D1 4
#define FB_KB 8.6173303e-5  // Boltzmann constant eV/K, reference src/festim/__init__.py:11
#define FB_STACK 24
#define FB_CHUNK 16           // block rows staged together by the TMA-pipelined row kernels (multiple of 4)
#define FB_CHUNK_GROUPS 1     // consumer groups of FB_CHUNK warps working on alternate chunks

// Everything a kernel needs, passed by value (lives in the constant bank).
struct DevSpec {
  int tdim, ns, n_vtags, n_stags, transient, T_mode, npairs, n_slots, n_diffE;
  int n_vterms, n_fterms, n_adv, n_fields;
  long long n_vertices, n_cells, n_bfacets, n_dofs;
  long long n_owned;   // vertices [0, n_owned) are owned by this rank, the rest are ghosts
  long long n_owned_dofs;
  // mesh + pattern
  const double* coords;
  const int* cells;
  const int* cell_tag;
  const int* bfacets;
  const int* bfacet_cell;
  const int* bfacet_tag;
  const int* rowptr;
  const int* colidx;
  const int* v2c_ptr;
  const int* v2c_idx;
  const unsigned* v2c_slot;   // per (vertex, incident cell): row slot of each cell vertex, 8 bits each
  // tables (spec)
  const double* diff_D0;      // [n_vtags*ns]
  const int* diff_slot;       // [n_vtags*ns] index into cell-coefficient slots, -1 none
  const double* diff_energies;  // [n_diffE]
  const int* diff_progs;      // [n_slots-n_diffE] program ids of generic diffusivities
  const int* diff_uprog;      // [n_vtags*ns*3] solution-dependent diffusivity: prog, derivs off, count (-1 none)
  int n_udiff;                // number of (tag, species) with a solution-dependent diffusivity
  const double* mass_dt;      // [n_vtags*ns]
  const double* mass_c;
  const int* pair_nc;         // [ns]
  const int* pair_pp;         // [ns+1]
  const int* pair_cs;         // [npairs]
  const int* pair_idx;        // [ns*ns]
  const int* prog_ops;
  const int* prog_iargs;
  const double* prog_fargs;
  const int* prog_offsets;
  const int* vterm_hdr;       // [n_vterms*8] tag, value_prog, rows_off, n_rows, derivs_off, n_derivs, u_dep
  const int* fterm_hdr;
  const int* row_species;
  const double* row_coef;
  const int* deriv_species;
  const int* deriv_prog;
  const int* quad_v;          // [n_vtags*4] offF, nqF, offJ, nqJ
  const int* quad_s;
  const double* quad_bary;
  const double* quad_w;
  const int* adv;             // [n_adv*2] tag, species
  // affine factors and rate energies of the structured reactions (lowering.py)
  int n_rslots, n_factors, n_monos, maxdeg, n_fac_terms;
  const int* pair_row;        // [npairs] row species of every stored pair
  const double* rslot_E;      // [n_rslots] activation energies of the rate slots
  const int* fac_hdr;         // [n_factors*4] a0 kind (0 const, 1 scalar), scalar id, coef offset, count
  const double* fac_a0;
  const int* fac_species;
  const double* fac_coef;
  // channel assembly (assemble_edge.cuh): per-cell record, channels, coefficient tables, edge
  // tensors and the channel operator (static / solution-dependent channel values per edge)
  int vx_on, vx_nsym, vx_nsym3, vx_rec, vx_nbw, vx_nch, vx_nstat, vx_ndyn, vx_nts, vx_nsp, vx_ndp;
  int vx_ncombo, vx_nft, vx_njt;
  int vx_itab_n, vx_dtab_n;   // sizes of the block-shared tables copied to shared memory
  int vx_o_ch, vx_o_ftptr, vx_o_ft, vx_o_jtptr, vx_o_jt, vx_o_jtpair, vx_o_dynptr, vx_o_dyn;   // offsets into itab
  int vx_o_ftc, vx_o_jtc;                                                                       // offsets into dtab
  unsigned vx_m_ns, vx_m_nf, vx_m_np, vx_m_nsp, vx_m_ndp;   // ceil(2^32 / d) for d = ns, n_factors, npairs, nsp, ndp
  const int* vx_cq_ptr;       // [n_vtags+1] cell-record quantities per tag
  const int* vx_cq;           // [n*4] kind, index, rule, offset in the record
  const int* vx_itab;         // channel / residual / Jacobian tables (ints)
  const double* vx_dtab;      // their coefficients
  const int* vx_sp_combo;     // [ncombo*2] (channel, column species) products the SpMV accumulates
  const int* vx_sp_ptr;       // [ns+1]
  const int* vx_sp;           // [n*2] (combo, flag) per row species
  const double* vx_sp_coef;
  const double* cellrec;      // [n_cells][vx_rec]
  const unsigned* vinfo;      // per vertex: Dirichlet species mask (bits 0-15), any Dirichlet column (bit 16)
  long long nnzb, et_tz_total;
  const double* et_stat;      // [nnzb][vx_nsp] static channel values per edge (per temperature update)
  const double* et_dyn;       // [nnzb][vx_ndp] solution-dependent channel values per edge (per assembly)
  const double* et_tvw;       // [nts][nnzb][2]  T_vvw and T_vww: moments sum_K M3[v,v,w], M3[v,w,w]
  const double* et_tz;        // [nts][et_tz_total]  T_vwz, z a common neighbour of v and w
  long long et_tz_chunk_max;  // largest T_vwz extent of a chunk of FB_CHUNK rows
  const int* et_tzw;          // [n_vertices, padded] W_v: triples per edge of the row
  const long long* et_tzptr;  // [n_vertices+1, padded] start of the row's (W_v x deg) slab, padded to multiples of 16
  const unsigned char* et_tzslot;   // [et_tz_total] row slot of z (0xff: padding)
  // chunk plan (fb_set_pattern): FB_CHUNK consecutive block rows are staged together - the
  // distinct neighbour vertices of a chunk as runs of consecutive ids (one bulk copy each)
  int ck_nx_max;              // largest number of staged vertices of a chunk
  int ck_rmax;                // most runs of a chunk
  const int* ck_slots;        // [nchunks][32][3] the first 32 runs: first vertex, length (0: unused), offset in the staged list
  const int* ck_runptr;       // [nchunks+1] all runs (CSR), same triples
  const int* ck_runs;
  const int* ck_nx;           // [nchunks] staged vertices
  const int* ck_lcolptr;      // [nchunks+1] start of the chunk's edge list in ck_lcol (multiples of 8)
  const unsigned short* ck_lcol;   // per edge: position of its column vertex in the staged list
  const unsigned* ck_rowinfo; // [n_vertices, padded] per row: position of the row's own vertex in the staged list
  // per-step data
  const double* scalars;
  double T_const;
  const double* T_nodal;
  const double* const* fields;      // device array of device pointers
  const double* const* velocities;
  const int* bc_map;          // [n_dofs] index into bc_vals or -1
  const double* bc_vals;
  const double* dbar;         // [2][n_slots][n_cells] cell means of diffusivity factors
  const double* load;         // [n_dofs] solution-independent part of the residual
  double inv_dt;
};

// Evaluation point: physical coordinates, temperature, species values and the
// entity (cell or facet) it lies in, for lazy P1-field interpolation.
struct PointCtx {
  double x[3];
  double T;
  const double* c;       // [ns] species at the point
  const double* cprev;   // [ns] previous-step species at the point
  const double* cgrad;   // [ns][3] gradient of the species in the owning cell (P1: constant), or nullptr
  double n[3];           // outward unit normal of the boundary facet (facet terms)
  double h;              // circumradius of the owning cell
  const int* verts;      // vertices of the integration entity
  const double* bary;    // barycentric coordinates w.r.t. verts
  int nverts;
  const int* cverts;     // vertices of the owning cell
  const double* G;       // [ncverts][tdim] basis gradients of the owning cell
  int ncverts;
  int tdim;
};

