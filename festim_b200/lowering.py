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
 S_JC_COEF, S_WTAB, S_WTAB_OFF, S_STRUCT_INTS, S_COMBO, S_PAIR_ROW) = range(1, 46)

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

        ns, nvt = self.ns, len(self.vol_ids)
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
            if isinstance(obj, ImplicitSpecies):
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
        # contribution lists
        fc = [[] for _ in range(nvt * ns)]                 # (mono, sign)
        jc = [[] for _ in range(nvt * self.npairs)]        # (slot, other factor, coef)
        for im, (tag, slot, facs, coef, rows) in enumerate(monos):
            for row, sign in rows:
                fc[tag * ns + row].append((im, sign))
                for pos, f in enumerate(facs):
                    other = facs[1 - pos] if len(facs) == 2 else -1
                    for spe, a in fac_rows[f][3]:
                        p = int(self.pair_pp[row]) + int(self.pair_idx[row, spe])
                        jc[tag * self.npairs + p].append((slot, other, sign * coef * a))
        self.struct = dict(monos=monos, fac_rows=fac_rows, fc=fc, jc=jc)

    def _structured_sections(self):
        st = self.struct
        nsym = (self.tdim + 1) * (self.tdim + 2) // 2
        d1 = self.tdim + 1
        fac_hdr = np.zeros((max(len(st["fac_rows"]), 1), 4), dtype=np.int32)
        fac_a0 = np.zeros(max(len(st["fac_rows"]), 1))
        fsp, fco = [], []
        for i, (kind, scal, a0, coefs) in enumerate(st["fac_rows"]):
            fac_hdr[i] = [kind, scal, len(fsp), len(coefs)]
            fac_a0[i] = a0
            for spe, c in coefs:
                fsp.append(spe)
                fco.append(c)
        # (rate slot, factor) combinations whose moment-weighted nodal vectors
        # V[j] = sum_b fac[b] M[j][b] are shared by residual and Jacobian contributions
        combos = []

        def combo_of(slot, f):
            if (slot, f) not in combos:
                combos.append((slot, f))
            return combos.index((slot, f))

        mono = np.full((max(len(st["monos"]), 1), 6), -1, dtype=np.int32)
        mono_coef = np.zeros(max(len(st["monos"]), 1))
        for i, (tag, slot, facs, coef, rows) in enumerate(st["monos"]):
            mono[i, :3] = [tag, slot, len(facs)]
            for k, f in enumerate(facs):
                mono[i, 3 + k] = f
            if len(facs) == 2:
                mono[i, 5] = combo_of(slot, facs[1])
            mono_coef[i] = coef
        fc_ptr, fc_mono, fc_sign = [0], [], []
        for lst in st["fc"]:
            for im, sg in lst:
                fc_mono.append(im)
                fc_sign.append(sg)
            fc_ptr.append(len(fc_mono))
        jc_ptr, jc_tab, jc_coef = [0], [], []
        for lst in st["jc"]:
            for slot, other, c in lst:
                # second column: combo index, or -(slot+1) for the plain first-order moment
                jc_tab.append([slot, combo_of(slot, other) if other >= 0 else -(slot + 1)])
                jc_coef.append(c)
            jc_ptr.append(len(jc_coef))
        # W tables: W[which][i][q][ab] = w_q phi_i phi_a phi_b (a <= b), then its q-sum
        wtab, woff = [], np.zeros((len(self.vol_ids), 4), dtype=np.int32)
        sym = [(a, b) for a in range(d1) for b in range(a, d1)]
        for t, vid in enumerate(self.vol_ids):
            dF, dJ = self.degrees.get(("dx", vid), (1, None))
            for which, deg in enumerate((dF, dF if dJ is None else dJ)):
                bary, w = quadrature.rule(self.tdim, deg)
                W = np.zeros((d1, len(w), nsym))
                for i in range(d1):
                    for k, (a, b) in enumerate(sym):
                        W[i, :, k] = w * bary[:, i] * bary[:, a] * bary[:, b]
                if which == 1 and (dJ is None or dJ == dF):
                    woff[t, 2:4] = woff[t, 0:2]      # same rule: share the table
                    continue
                woff[t, 2 * which] = sum(len(x) for x in wtab)
                woff[t, 2 * which + 1] = len(w)
                wtab.append(W.ravel())
                wtab.append(W.sum(axis=1).ravel())
        ints = np.zeros(8, dtype=np.int32)
        ints[:5] = [len(self.rslot_E), len(st["fac_rows"]), len(st["monos"]), nsym, len(combos)]
        pair_row = np.repeat(np.arange(self.ns, dtype=np.int32), self.pair_nc)
        return {
            S_STRUCT_INTS: ints, S_RSLOT_E: np.array(self.rslot_E, dtype=np.float64),
            S_FAC_HDR: fac_hdr, S_FAC_A0: fac_a0, S_FAC_SPECIES: np.array(fsp, dtype=np.int32),
            S_FAC_COEF: np.array(fco, dtype=np.float64), S_MONO: mono, S_MONO_COEF: mono_coef,
            S_FC_PTR: np.array(fc_ptr, dtype=np.int32), S_FC_MONO: np.array(fc_mono, dtype=np.int32),
            S_FC_SIGN: np.array(fc_sign, dtype=np.float64),
            S_JC_PTR: np.array(jc_ptr, dtype=np.int32),
            S_JC: np.array(jc_tab, dtype=np.int32).reshape(-1, 2), S_JC_COEF: np.array(jc_coef, dtype=np.float64),
            S_WTAB: np.concatenate(wtab) if wtab else np.zeros(0), S_WTAB_OFF: woff,
            S_COMBO: np.array(combos, dtype=np.int32).reshape(-1, 2), S_PAIR_ROW: pair_row,
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
