"""Lowering: FESTIM problem -> flat per-cell kernel spec for the CUDA library.

This replaces UFL/FFCx form compilation at the solver seam (reference
``src/festim/problem.py:170-182``: ``NonlinearProblem(F, u, bcs, ...)``).  The
residual form built by ``create_formulation`` is split into the term classes
the device kernels implement (``festim_b200/csrc/assemble.cu``):

* Arrhenius diffusion ``D_0 exp(-E_D/(k_B T)) grad u . grad v`` per (volume tag,
  species): constants ``D_0`` and an index into the table of distinct activation
  energies (cell averages of the Arrhenius factor are cached per temperature
  update, for the F and the J quadrature rule separately);
* mass terms ``m (u - u_n)/dt v`` and ``m u v`` (exact closed form);
* advection ``(grad u . vel) v`` with a P1 vector velocity;
* point terms: any other integrand ``expr v`` (reactions, sources, fluxes) as a
  byte-code program for the value plus one program per species derivative
  (the symbolic ``ufl.derivative`` analogue), applied to a list of
  (row species, coefficient) - e.g. ``+R`` on every reactant, ``-R`` on every
  product of a reaction;
* which species pairs are structurally non-zero in a node block (the
  species-sparse BSR layout documented in ``include/festim_b200.h``).

Blob layout (little endian): int64 magic, int64 n_sections, then per section
(int64 id, int64 dtype code [0=int32, 1=float64], int64 count, int64 byte
offset), then the 8-byte aligned payloads.
"""

from __future__ import annotations

import numpy as np

from festim_b200 import fem, quadrature, ufl

MAGIC = 0xFB200_0001

# section ids (shared with csrc/spec.h)
(S_INTS, S_DIFF_D0, S_DIFF_SLOT, S_DIFF_ENERGIES, S_MASS_DT, S_MASS_C, S_PAIR_NC, S_PAIR_PP,
 S_PAIR_CS, S_PAIR_IDX, S_PROG_OPS, S_PROG_IARGS, S_PROG_FARGS, S_PROG_OFFSETS, S_VTERM_HDR,
 S_FTERM_HDR, S_ROW_SPECIES, S_ROW_COEF, S_DERIV_SPECIES, S_DERIV_PROG, S_QUAD_V, S_QUAD_S,
 S_QUAD_BARY, S_QUAD_W, S_ADV, S_DIFF_PROGS, S_DIFF_UPROG, S_RSLOT_E, S_FAC_HDR, S_FAC_A0,
 S_FAC_SPECIES, S_FAC_COEF, S_MONO, S_MONO_COEF, S_FC_PTR, S_FC_MONO, S_FC_SIGN, S_JC_PTR, S_JC,
 S_JC_COEF, S_WTAB, S_WTAB_OFF, S_STRUCT_INTS, S_COMBO, S_PAIR_ROW,
 S_VX_INTS, S_VX_CQ_PTR, S_VX_CQ, S_VX_CH_PTR, S_VX_CH, S_VX_JT_PTR, S_VX_JT, S_VX_JT_COEF,
 S_VX_FT_PTR, S_VX_FT, S_VX_FT_COEF, S_VX_SP_COMBO, S_VX_SP_PTR, S_VX_SP, S_VX_SP_COEF) = range(1, 61)

# cell-record quantity kinds / channel types / flags of the vertex-block assembly
# (csrc/assemble_vtx.cuh)
CQ_KD, CQ_M1, CQ_M3, CQ_ADV = 0, 1, 2, 3
CH_MASS, CH_PAIR, CH_TRI, CH_FULL = 0, 1, 2, 3
USE_F, USE_J = 1, 2
FLAG_INV_DT = 1

# S_INTS slots
(I_TDIM, I_NS, I_NVTAGS, I_NSTAGS, I_TRANSIENT, I_TMODE, I_NPROGRAMS, I_NSCALARS, I_NFIELDS,
 I_NVTERMS, I_NFTERMS, I_NADV, I_NPAIRS, I_NDIFFE, I_DT_SCALAR, I_NSLOTS) = range(16)

MAX_NS = 16


class PointTerm:
    def __init__(self, tag, value_prog, rows, derivs, u_dependent, test_kind=0):
        self.test_kind = test_kind  # 0: v, 1: n . grad(v) (facet terms only)
        self.tag = tag
        self.value_prog = value_prog
        self.rows = rows          # [(species index, coefficient)]
        self.derivs = derivs      # [(dep species index, program id)]
        self.u_dependent = u_dependent


class KernelSpec:
    """Host-side description; ``blob()`` serialises what the library needs."""

    def __init__(self, problem):
        self.problem = problem
        mesh = problem.mesh.mesh
        self.mesh = mesh
        self.tdim = mesh.tdim
        species = getattr(problem, "species", None) or [problem._unknown]
        self.species = list(species)
        ns = self.ns = len(self.species)
        if ns > MAX_NS:
            raise NotImplementedError(f"more than {MAX_NS} species")
        self.vol_ids = [v.id for v in problem.volume_subdomains]
        self.surf_ids = [s.id for s in problem.surface_subdomains]
        nvt, nst = len(self.vol_ids), len(self.surf_ids)

        # compact tags (vectorised: configs reach 10^7 cells)
        vt = problem.volume_meshtags
        ctag = np.zeros(mesh.num_cells, dtype=np.int64)
        ctag[vt.indices] = vt.values
        self.cell_tag = self._compact(ctag, self.vol_ids, required=True)
        ft = problem.facet_meshtags
        ftag = np.zeros(mesh.bfacets.shape[0], dtype=np.int64)
        ftag[ft.indices] = ft.values
        self.bfacet_tag = self._compact(ftag, self.surf_ids, required=False)

        T = getattr(problem, "temperature_fenics", None)
        self.temperature = T
        self.T_mode = 1 if isinstance(T, fem.Function) else 0
        sidx = {spe: i for i, spe in enumerate(self.species)}
        self.programs = ufl.ProgramBuilder(species_index=sidx, temperature=T)
        self.transient = bool(problem.settings.transient)
        self.dt_scalar = -1
        if self.transient:
            self.dt_scalar = self.programs._scalar_id(problem.dt)

        self.diff_D0 = np.zeros((nvt, ns))
        # cell-coefficient slots: distinct activation energies first, then
        # generic (program) diffusivities; diff_slot < 0 = no diffusion
        self.diff_slot = np.full((nvt, ns), -1, dtype=np.int32)
        self._generic_D = []   # (tag, s, program id), slots assigned after the energies
        self.diff_energies = []
        self.diff_udep = {}    # (tag, s) -> (value program, [(dep species, derivative program)])
        self.mass_dt = np.zeros((nvt, ns))
        self.mass_c = np.zeros((nvt, ns))
        self.vterms = []
        self.fterms = []
        self.struct_reactions = {}   # (id(reaction), tag) -> (reaction, tag, [(row species, sign)])
        self.adv = []          # (vtag index, species index, VectorFunction)
        pairs = {(s, s) for s in range(ns)}

        degrees = problem.formulation.quadrature_degrees()
        self.degrees = degrees
        reaction_terms = {}
        for itg in problem.formulation.integrals:
            kind, sid = itg.measure
            s = sidx[itg.species]
            if kind == "dx":
                tag = self.vol_ids.index(sid)
            else:
                tag = self.surf_ids.index(sid)
            if itg.kind == "diffusion":
                if "arrhenius" in itg.meta:
                    D0, E = itg.meta["arrhenius"]
                    self.diff_D0[tag, s] = D0
                    self.diff_slot[tag, s] = self._energy_index(float(E))
                elif ufl.species_in(itg.D):
                    # solution-dependent diffusivity (e.g. lambda(T) of the heat problem):
                    # value program + derivative programs, evaluated while gathering
                    derivs = [(sidx[spe], self.programs.add(ufl.diff(itg.D, ("species", spe))))
                              for spe in ufl.species_in(itg.D)]
                    self.diff_udep[(tag, s)] = (self.programs.add(itg.D), derivs)
                    for dep, _ in derivs:
                        pairs.add((s, dep))
                else:
                    self.diff_D0[tag, s] = 1.0
                    self._generic_D.append((tag, s, self.programs.add(itg.D)))
            elif itg.kind == "advection":
                self.adv.append((tag, s, itg.velocity))
            elif "mass" in itg.meta and kind == "dx":
                self.mass_dt[tag, s] += float(itg.meta["mass"])
            elif "filler" in itg.meta and kind == "dx":
                self.mass_c[tag, s] += float(itg.meta["filler"])
            else:
                terms = self.vterms if kind == "dx" else self.fterms
                reaction = itg.meta.get("reaction")
                if reaction is not None and kind == "dx" and self.structured_ok(reaction):
                    self.struct_reactions.setdefault((id(reaction), tag), (reaction, tag, []))[2].append(
                        (s, float(itg.meta["sign"])))
                    for dep in self._reaction_species(reaction, sidx):
                        pairs.add((s, dep))
                elif reaction is not None:
                    key = (id(reaction), kind, tag)
                    if key not in reaction_terms:
                        # value program: the positive reaction term, evaluated once
                        # per point and applied with +-1 to every reactant / product
                        pos = itg.expr if itg.meta["sign"] > 0 else -itg.expr
                        term = self._point_term(tag, pos, [], sidx)
                        reaction_terms[key] = term
                        terms.append(term)
                    reaction_terms[key].rows.append((s, float(itg.meta["sign"])))
                    for dep, _ in reaction_terms[key].derivs:
                        pairs.add((s, dep % ns))
                else:
                    term = self._point_term(tag, itg.expr, [(s, 1.0)], sidx)
                    if itg.kind == "ngradtest":
                        if kind == "dx":
                            raise NotImplementedError("n . grad(v) test functions exist on facets only")
                        term.test_kind = 1
                        term.u_dependent = True   # evaluated by the facet kernel with the cell data
                    terms.append(term)
                    for dep, _ in term.derivs:
                        pairs.add((s, dep % ns))

        self.diff_progs = []
        for tag, s, pid in self._generic_D:
            if pid not in self.diff_progs:
                self.diff_progs.append(pid)
            self.diff_slot[tag, s] = len(self.diff_energies) + self.diff_progs.index(pid)
        self.n_slots = len(self.diff_energies) + len(self.diff_progs)

        # species pairs -> layout tables
        self.pair_nc = np.zeros(ns, dtype=np.int32)
        self.pair_pp = np.zeros(ns + 1, dtype=np.int32)
        self.pair_idx = np.full((ns, ns), -1, dtype=np.int32)
        cs = []
        for s in range(ns):
            cols = sorted(c for (r, c) in pairs if r == s)
            self.pair_nc[s] = len(cols)
            self.pair_pp[s + 1] = self.pair_pp[s] + len(cols)
            for j, c in enumerate(cols):
                self.pair_idx[s, c] = j
            cs.extend(cols)
        self.pair_cs = np.array(cs, dtype=np.int32)
        self.npairs = int(self.pair_pp[-1])
        self.struct = None
        if self.struct_reactions:
            self._build_structured(sidx)
        self._vertex_tables()

        # quadrature pools
        self._qb, self._qw = [], []
        self.quad_v = np.zeros((nvt, 4), dtype=np.int32)
        self.quad_s = np.zeros((max(nst, 1), 4), dtype=np.int32)
        for i, vid in enumerate(self.vol_ids):
            dF, dJ = degrees.get(("dx", vid), (1, None))
            self.quad_v[i] = self._add_rule(self.tdim, dF) + self._add_rule(
                self.tdim, dF if dJ is None else dJ)
        for i, sid in enumerate(self.surf_ids):
            dF, dJ = degrees.get(("ds", sid), (1, None))
            self.quad_s[i] = self._add_rule(self.tdim - 1, dF) + self._add_rule(
                self.tdim - 1, dF if dJ is None else dJ)
        # user-defined derived quantities (CustomQuantity): one program and one rule each
        self.custom = {}
        for export in getattr(problem, "exports", None) or []:
            if type(export).__name__ != "CustomQuantity" or getattr(export, "ufl_expr", None) is None:
                continue
            sub = export.subdomain
            surface = not hasattr(sub, "material")
            tag = (self.surf_ids if surface else self.vol_ids).index(sub.id)
            degree = ufl.estimate_degree(export.ufl_expr)
            off, nq = self._add_rule(self.tdim - 1 if surface else self.tdim, degree)
            self.custom[id(export)] = (int(surface), tag, self.programs.add(export.ufl_expr), off, nq, degree)

    # ------------------------------------------------------------------
    # structured reactions: k0 exp(-Ek/kT) f1 [f2] - p(T) g1 [g2] with every factor an
    # affine function of the species (a Species, or an ImplicitSpecies n - sum(others)
    # with n a constant); assembled from rate-weighted basis moments
    # (csrc/assemble_rows.cuh) instead of point-wise program evaluation
    @staticmethod
    def structured_ok(reaction):
        import os

        from festim_b200.species import ImplicitSpecies, Species

        if os.environ.get("FB_NO_STRUCTURED"):
            return False
        if len(reaction.reactant) > 2 or len(reaction.products) > 2:
            return False
        for r in reaction.reactant:
            if isinstance(r, ImplicitSpecies):
                if not isinstance(r.value_fenics, fem.Constant):
                    return False
            elif not isinstance(r, Species):
                return False
        return all(isinstance(x, (int, float)) for x in (reaction.k_0, reaction.E_k)) and \
            all(x is None or isinstance(x, (int, float)) for x in (reaction.p_0, reaction.E_p))

    @staticmethod
    def _reaction_species(reaction, sidx):
        from festim_b200.species import ImplicitSpecies

        out = []
        for r in list(reaction.reactant) + list(reaction.products):
            members = r.others if isinstance(r, ImplicitSpecies) else [r]
            for m in members:
                if sidx[m] not in out:
                    out.append(sidx[m])
        return out

    def _build_structured(self, sidx):
        """Tables for the moment-based reaction assembly."""
        from festim_b200.species import ImplicitSpecies

        self.rslot_E = []
        factors = {}      # key -> id
        fac_rows = []     # (a0_kind, a0_scalar, a0_const, [(species, coef)])
        monos = []        # (tag, slot, [factor ids], coef, [(row, sign)])

        def slot_of(E):
            E = float(E or 0.0)
            if E not in self.rslot_E:
                self.rslot_E.append(E)
            return self.rslot_E.index(E)

        def factor_of(obj):
            if isinstance(obj, ImplicitSpecies) and isinstance(obj.n, (int, float)):
                # a plain number never changes: equal densities over the same species share a factor
                key = ("impc", float(obj.n), tuple(sidx[o] for o in obj.others))
                if key not in factors:
                    factors[key] = len(fac_rows)
                    fac_rows.append((0, -1, float(obj.n), [(sidx[o], -1.0) for o in obj.others]))
            elif isinstance(obj, ImplicitSpecies):
                key = ("imp", id(obj.value_fenics), tuple(sidx[o] for o in obj.others))
                if key not in factors:
                    factors[key] = len(fac_rows)
                    fac_rows.append((1, self.programs._scalar_id(obj.value_fenics), 0.0,
                                     [(sidx[o], -1.0) for o in obj.others]))
            else:
                key = ("spe", sidx[obj])
                if key not in factors:
                    factors[key] = len(fac_rows)
                    fac_rows.append((0, -1, 0.0, [(sidx[obj], 1.0)]))
            return factors[key]

        for reaction, tag, rows in self.struct_reactions.values():
            reaction.check_rates()
            monos.append((tag, slot_of(reaction.E_k), [factor_of(r) for r in reaction.reactant],
                          float(reaction.k_0), rows))
            mode = reaction.rate_mode()
            if mode and reaction.products:
                monos.append((tag, slot_of(reaction.E_p if mode == 2 else 0.0),
                              [factor_of(pr) for pr in reaction.products], -float(reaction.p_0), rows))
        # a factor that is one species with coefficient one is read from the solution itself;
        # the others get a column of their own in the per-neighbour value table
        dev = [i for i, (kind, _, a0, coefs) in enumerate(fac_rows)
               if not (kind == 0 and a0 == 0.0 and len(coefs) == 1 and coefs[0][1] == 1.0)]
        self.struct = dict(monos=monos, fac_rows=fac_rows, dev_factors=dev)

    def _structured_sections(self):
        """Affine factors and activation energies of the structured reactions."""
        st = self.struct
        rows = [st["fac_rows"][i] for i in st["dev_factors"]]
        fac_hdr = np.zeros((max(len(rows), 1), 4), dtype=np.int32)
        fac_a0 = np.zeros(max(len(rows), 1))
        fsp, fco = [], []
        for i, (kind, scal, a0, coefs) in enumerate(rows):
            fac_hdr[i] = [kind, scal, len(fsp), len(coefs)]
            fac_a0[i] = a0
            for spe, c in coefs:
                fsp.append(spe)
                fco.append(c)
        ints = np.zeros(8, dtype=np.int32)
        ints[:3] = [len(self.rslot_E), len(rows), len(st["monos"])]
        return {
            S_STRUCT_INTS: ints, S_RSLOT_E: np.array(self.rslot_E, dtype=np.float64),
            S_FAC_HDR: fac_hdr, S_FAC_A0: fac_a0, S_FAC_SPECIES: np.array(fsp, dtype=np.int32),
            S_FAC_COEF: np.array(fco, dtype=np.float64),
        }

    # ------------------------------------------------------------------
    # vertex-block assembly (csrc/assemble_vtx.cuh).  Every volume term whose integrand is a
    # product of P1 functions with a cell-wise coefficient is assembled from *channels*:
    # per-cell (D1 x D1) matrices  C^c_ij  that do not depend on the species pair,
    #     mass        |K| (1 + delta_ij) / ((d+1)(d+2))
    #     diffusion   |K| mean_q(exp(-E/kT)) grad(phi_i).grad(phi_j)          (cached per cell)
    #     first-order rate moment   |K| sum_q w_q rho_m phi_i phi_j            (cached per cell)
    #     second-order rate moment contracted with an affine factor f of the species,
    #                 sum_b (|K| sum_q w_q rho_m phi_i phi_j phi_b) f_b        (moment cached)
    #     advection   |K| sum_q w_q phi_i vel_q.grad(phi_j)                    (cached per cell)
    # A warp sums the channels of the cells around its vertex per neighbour ("edge channels"
    # A^c_vw), then every stored Jacobian entry is a short linear combination
    #     J[(v,s),(w,s')] = sum_e coef_e A^{chan_e}_vw          (table per tag and species pair)
    # and the residual is  F[(v,s)] = load + sum_e coef_e sum_w A^{chan_e}_vw g_e(w)  with g one
    # of u_s', (u - u_n)_s', an affine factor, or 1.  Same sums over quadrature points as the
    # reference's generated kernel (reaction.py:179-224 expanded in the P1 basis), reordered.
    def _vertex_tables(self):
        import os

        self.vx = None
        if os.environ.get("FB_NO_VTX"):
            return
        if any(t.u_dependent for t in self.vterms) or self.diff_udep:
            return   # generic point-wise terms: thread-per-vertex kernel
        nvt, ns, d1 = len(self.vol_ids), self.ns, self.tdim + 1
        if nvt > 32:
            return
        nsym, nsym3 = d1 * (d1 + 1) // 2, d1 * (d1 + 1) * (d1 + 2) // 6
        st = self.struct or dict(monos=[], fac_rows=[], dev_factors=[])
        fac_rows, dev = st["fac_rows"], st["dev_factors"]
        NF = len(dev)
        g_u, g_du, g_one = (lambda s: s), (lambda s: ns + s), 2 * ns + NF

        def g_fac(f):   # column of factor f in the per-neighbour value table [u | u - u_n | factors | 1]
            return 2 * ns + dev.index(f) if f in dev else g_u(fac_rows[f][3][0][0])

        qsize = {CQ_KD: nsym, CQ_M1: nsym, CQ_M3: nsym3, CQ_ADV: d1 * d1}
        cq_all = []
        rec_max = 1
        channels, ch = {}, []                        # global channel list: [tag, type, offset, factor column, use]
        jt = [dict() for _ in range(self.npairs)]    # pair -> {(chan, flag): coef}
        ft = [[] for _ in range(ns)]                 # species -> [(chan, gcol, coef, flag)]

        def addJ(row, col, chan, coef, flag=0):
            p = int(self.pair_pp[row]) + int(self.pair_idx[row, col])
            jt[p][(chan, flag)] = jt[p].get((chan, flag), 0.0) + coef

        for tag, vid in enumerate(self.vol_ids):
            dF, dJ = self.degrees.get(("dx", vid), (1, None))
            same = dJ is None or dJ == dF
            rule = (lambda which: 1) if same else (lambda which: which)
            quantities, cq, size = {}, [], [1]          # record[0] = |K|

            def quantity(kind, a, r):
                key = (kind, a, r)
                if key not in quantities:
                    quantities[key] = size[0]
                    cq.append([kind, a, r, size[0]])
                    size[0] += qsize[kind]
                return quantities[key]

            def channel(ctype, off, fac, use):
                key = (tag, ctype, off, fac)
                if key not in channels:
                    channels[key] = len(ch)
                    ch.append([tag, ctype, off, fac, 0])
                ch[channels[key]][4] |= use
                return channels[key]

            for s in range(ns):
                ds = int(self.diff_slot[tag, s])
                if ds >= 0:
                    D0 = float(self.diff_D0[tag, s])
                    cF = channel(CH_PAIR, quantity(CQ_KD, ds, rule(0)), -1, USE_F)
                    ft[s].append((cF, g_u(s), D0, 0))
                    cJ = channel(CH_PAIR, quantity(CQ_KD, ds, rule(1)), -1, USE_J)
                    addJ(s, s, cJ, D0)
                if self.transient and self.mass_dt[tag, s] != 0.0:
                    c = channel(CH_MASS, 0, -1, USE_F | USE_J)
                    ft[s].append((c, g_du(s), float(self.mass_dt[tag, s]), FLAG_INV_DT))
                    addJ(s, s, c, float(self.mass_dt[tag, s]), FLAG_INV_DT)
                if self.mass_c[tag, s] != 0.0:
                    c = channel(CH_MASS, 0, -1, USE_F | USE_J)
                    ft[s].append((c, g_u(s), float(self.mass_c[tag, s]), 0))
                    addJ(s, s, c, float(self.mass_c[tag, s]))
            for a, (atag, s, _) in enumerate(self.adv):
                if atag == tag:
                    c = channel(CH_FULL, quantity(CQ_ADV, a, 0), -1, USE_F | USE_J)
                    ft[s].append((c, g_u(s), 1.0, 0))
                    addJ(s, s, c, 1.0)
            for mtag, slot, facs, coef, rows in st["monos"]:
                if mtag != tag:
                    continue

                def moment(which, other):
                    r = rule(which)
                    use = USE_F if which == 0 else USE_J
                    if other >= 0:
                        return channel(CH_TRI, quantity(CQ_M3, slot, r), g_fac(other), use)
                    return channel(CH_PAIR, quantity(CQ_M1, slot, r), -1, use)

                for row, sign in rows:
                    if len(facs) == 2:
                        ft[row].append((moment(0, facs[1]), g_fac(facs[0]), sign * coef, 0))
                    elif len(facs) == 1:
                        ft[row].append((moment(0, -1), g_fac(facs[0]), sign * coef, 0))
                    else:
                        ft[row].append((moment(0, -1), g_one, sign * coef, 0))
                    for pos, f in enumerate(facs):
                        other = facs[1 - pos] if len(facs) == 2 else -1
                        chan = moment(1, other)
                        for spe, a in fac_rows[f][3]:
                            addJ(row, spe, chan, sign * coef * a)
            cq_all.append(cq)
            rec_max = max(rec_max, size[0])
        # static channels (functions of the mesh and the temperature only) first, then the
        # solution-dependent ones (second-order moments contracted with an affine factor); the
        # latter read their moments from an edge tensor per distinct (tag, moment) quantity
        order = [i for i, c in enumerate(ch) if c[1] != CH_TRI] + [i for i, c in enumerate(ch) if c[1] == CH_TRI]
        new_id = {old: new for new, old in enumerate(order)}
        nstat = sum(1 for c in ch if c[1] != CH_TRI)
        tensors = []
        ch_out = []
        for old in order:
            tag, ctype, off, fac, use = ch[old]
            ts = -1
            if ctype == CH_TRI:
                if (tag, off) not in tensors:
                    tensors.append((tag, off))
                ts = tensors.index((tag, off))
            ch_out.append([tag, ctype, off, fac, use, ts, 0, 0])
        jt = [{(new_id[c], fl): v for (c, fl), v in d.items()} for d in jt]
        ft = [[(new_id[c], g, v, fl) for c, g, v, fl in lst] for lst in ft]
        # SpMV on the channel operator: y_s = sum_e coef_e t[combo_e],  t[(chan, cs)] = sum_w A^chan_vw x_(w,cs)
        pair_row = np.repeat(np.arange(ns), self.pair_nc)
        combos, sp = {}, [[] for _ in range(ns)]
        for p in range(self.npairs):
            for (chan, flag), coef in jt[p].items():
                key = (chan, int(self.pair_cs[p]))
                combos.setdefault(key, len(combos))
                sp[int(pair_row[p])].append((combos[key], flag, coef))
        self.vx = dict(nsym=nsym, nsym3=nsym3, rec=rec_max, nbw=g_one + 1, NF=NF,
                       nch=max(len(ch_out), 1), nstat=nstat, ndyn=len(ch_out) - nstat, tensors=tensors,
                       cq=cq_all, ch=ch_out, jt=jt, ft=ft, combos=list(combos), sp=sp)

    def _vertex_sections(self):
        vx = self.vx
        nvt, ns = len(self.vol_ids), self.ns

        def csr(lists, width):
            ptr, flat = [0], []
            for lst in lists:
                flat.extend(lst)
                ptr.append(len(flat))
            return (np.array(ptr, dtype=np.int32),
                    np.array(flat, dtype=np.int32).reshape(-1, width))

        cq_ptr, cq = csr(vx["cq"], 4)
        jt_ptr, jt, jt_coef = [0], [], []
        for p in range(self.npairs):
            for (chan, flag), coef in vx["jt"][p].items():
                jt.append([chan, flag])
                jt_coef.append(coef)
            jt_ptr.append(len(jt))
        ft_ptr, ft, ft_coef = [0], [], []
        for s in range(ns):
            for chan, gcol, coef, flag in vx["ft"][s]:
                ft.append([chan, gcol, flag])
                ft_coef.append(coef)
            ft_ptr.append(len(ft))
        sp_ptr, sp, sp_coef = [0], [], []
        for s in range(ns):
            for combo, flag, coef in vx["sp"][s]:
                sp.append([combo, flag])
                sp_coef.append(coef)
            sp_ptr.append(len(sp))
        ints = np.zeros(16, dtype=np.int32)
        ints[:11] = [1, vx["nsym"], vx["nsym3"], vx["rec"], vx["nbw"], len(vx["ch"]), vx["NF"],
                     vx["nstat"], vx["ndyn"], len(vx["tensors"]), len(vx["combos"])]
        return {
            S_VX_INTS: ints, S_VX_CQ_PTR: cq_ptr, S_VX_CQ: cq,
            S_VX_CH: np.array(vx["ch"], dtype=np.int32).reshape(-1, 8),
            S_VX_JT_PTR: np.array(jt_ptr, dtype=np.int32),
            S_VX_JT: np.array(jt, dtype=np.int32).reshape(-1, 2),
            S_VX_JT_COEF: np.array(jt_coef, dtype=np.float64),
            S_VX_FT_PTR: np.array(ft_ptr, dtype=np.int32),
            S_VX_FT: np.array(ft, dtype=np.int32).reshape(-1, 3),
            S_VX_FT_COEF: np.array(ft_coef, dtype=np.float64),
            S_VX_SP_COMBO: np.array(vx["combos"], dtype=np.int32).reshape(-1, 2),
            S_VX_SP_PTR: np.array(sp_ptr, dtype=np.int32),
            S_VX_SP: np.array(sp, dtype=np.int32).reshape(-1, 2),
            S_VX_SP_COEF: np.array(sp_coef, dtype=np.float64),
            S_PAIR_ROW: np.repeat(np.arange(self.ns, dtype=np.int32), self.pair_nc),
        }

    @staticmethod
    def _compact(tags, ids, required):
        """Map subdomain ids to their position in ``ids`` (-1: not a subdomain)."""
        out = np.full(len(tags), -1, dtype=np.int32)
        for i, sid in enumerate(ids):
            out[tags == sid] = i
        if required and len(tags) and out.min() < 0:
            raise ValueError("cells tagged with an id that is not a volume subdomain")
        return out

    def _energy_index(self, E):
        for i, e in enumerate(self.diff_energies):
            if e == E:
                return i
        self.diff_energies.append(E)
        return len(self.diff_energies) - 1

    def _point_term(self, tag, expr, rows, sidx):
        if expr is None:
            return None
        deps = ufl.species_in(expr)
        derivs = [(sidx[spe], self.programs.add(ufl.diff(expr, ("species", spe)))) for spe in deps]
        # dependence on the cell gradient of a species (metric terms of cylindrical /
        # spherical coordinates): derivative code = species + ns * (1 + component); the
        # Jacobian weight of column vertex j is d_k phi_j instead of phi_j
        gdeps = ufl.species_grads_in(expr)
        ns = len(sidx)
        derivs += [(sidx[spe] + ns * (1 + comp),
                    self.programs.add(ufl.diff(expr, ("species_grad", (spe, comp))))) for spe, comp in gdeps]
        u_dep = bool(deps) or bool(gdeps) or ufl.depends_on(expr, ufl.SpeciesRef)
        return PointTerm(tag, self.programs.add(expr), rows, derivs, u_dep)

    def _add_rule(self, tdim, degree):
        bary, w = quadrature.rule(tdim, degree)
        off = sum(len(x) for x in self._qw)
        # one pool for cell and facet rules, addressed by point offset: every point takes
        # tdim+1 doubles (facet rules are padded with a zero)
        if bary.shape[1] < self.tdim + 1:
            bary = np.concatenate([bary, np.zeros((bary.shape[0], self.tdim + 1 - bary.shape[1]))], axis=1)
        self._qb.append(bary.ravel())
        self._qw.append(w)
        return [off, len(w)]

    # ------------------------------------------------------------------
    @property
    def scalars(self):
        return self.programs.scalars

    @property
    def fields(self):
        return self.programs.fields

    def scalar_values(self):
        return np.array([float(s.value) for s in self.programs.scalars], dtype=np.float64)

    def blob(self):
        ops, iargs, fargs, offsets = self.programs.tables()
        ints = np.zeros(16, dtype=np.int32)
        ints[[I_TDIM, I_NS, I_NVTAGS, I_NSTAGS, I_TRANSIENT, I_TMODE]] = [
            self.tdim, self.ns, len(self.vol_ids), len(self.surf_ids), int(self.transient),
            self.T_mode]
        ints[[I_NPROGRAMS, I_NSCALARS, I_NFIELDS]] = [
            self.programs.n_programs, len(self.programs.scalars), len(self.programs.fields)]
        ints[[I_NVTERMS, I_NFTERMS, I_NADV, I_NPAIRS, I_NDIFFE, I_DT_SCALAR, I_NSLOTS]] = [
            len(self.vterms), len(self.fterms), len(self.adv), self.npairs,
            len(self.diff_energies), self.dt_scalar, self.n_slots]

        row_species, row_coef, deriv_species, deriv_prog = [], [], [], []

        def headers(terms):
            hdr = np.zeros((max(len(terms), 1), 8), dtype=np.int32)
            for i, t in enumerate(terms):
                hdr[i] = [t.tag, t.value_prog, len(row_species), len(t.rows),
                          len(deriv_species), len(t.derivs), int(t.u_dependent), int(t.test_kind)]
                for s, c in t.rows:
                    row_species.append(s)
                    row_coef.append(c)
                for s, p in t.derivs:
                    deriv_species.append(s)
                    deriv_prog.append(p)
            return hdr

        vhdr = headers(self.vterms)
        fhdr = headers(self.fterms)
        # solution-dependent diffusivities: per (tag, s): program, derivs offset, count
        nvt = len(self.vol_ids)
        uprog = np.full((nvt, self.ns, 3), -1, dtype=np.int32)
        for (tag, s), (pid, derivs) in self.diff_udep.items():
            uprog[tag, s] = [pid, len(deriv_species), len(derivs)]
            for dep, dp in derivs:
                deriv_species.append(dep)
                deriv_prog.append(dp)
        adv = np.array([[t, s] for t, s, _ in self.adv], dtype=np.int32).reshape(-1, 2)
        sections = {
            S_INTS: ints,
            S_DIFF_D0: self.diff_D0, S_DIFF_SLOT: self.diff_slot,
            S_DIFF_ENERGIES: np.array(self.diff_energies, dtype=np.float64),
            S_DIFF_PROGS: np.array(self.diff_progs, dtype=np.int32),
            S_MASS_DT: self.mass_dt, S_MASS_C: self.mass_c,
            S_PAIR_NC: self.pair_nc, S_PAIR_PP: self.pair_pp, S_PAIR_CS: self.pair_cs,
            S_PAIR_IDX: self.pair_idx,
            S_PROG_OPS: ops, S_PROG_IARGS: iargs, S_PROG_FARGS: fargs, S_PROG_OFFSETS: offsets,
            S_VTERM_HDR: vhdr, S_FTERM_HDR: fhdr,
            S_ROW_SPECIES: np.array(row_species, dtype=np.int32),
            S_ROW_COEF: np.array(row_coef, dtype=np.float64),
            S_DERIV_SPECIES: np.array(deriv_species, dtype=np.int32),
            S_DERIV_PROG: np.array(deriv_prog, dtype=np.int32),
            S_QUAD_V: self.quad_v, S_QUAD_S: self.quad_s,
            S_QUAD_BARY: np.concatenate(self._qb) if self._qb else np.zeros(0),
            S_QUAD_W: np.concatenate(self._qw) if self._qw else np.zeros(0),
            S_ADV: adv, S_DIFF_UPROG: uprog,
        }
        if self.struct is not None:
            sections.update(self._structured_sections())
        if self.vx is not None:
            sections.update(self._vertex_sections())
        return pack_sections(sections)


def pack_sections(sections):
    items = []
    for sid, arr in sections.items():
        arr = np.ascontiguousarray(arr)
        if arr.dtype == np.int32:
            code = 0
        elif arr.dtype == np.float64:
            code = 1
        else:
            raise TypeError(f"section {sid}: dtype {arr.dtype}")
        items.append((sid, code, arr))
    header = 2 + 4 * len(items)
    offset = header * 8
    table = [MAGIC, len(items)]
    payload = []
    for sid, code, arr in items:
        nbytes = arr.nbytes
        table += [sid, code, arr.size, offset]
        pad = (-nbytes) % 8
        payload.append(arr.tobytes() + b"\0" * pad)
        offset += nbytes + pad
    return np.array(table, dtype=np.int64).tobytes() + b"".join(payload)
