"""Golden vectors produced by the REAL reference code (tools/make_golden.py imports
/root/reference/src/festim with its numerical dependencies stubbed): they pin the
host logic of festim_b200, the oracle's Newton stopping rule and the C-ABI step
size rule to the reference, bit for bit."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import festim_b200 as F
from festim_b200 import ufl
from oracle.newton import convergence_test

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return json.load(open(os.path.join(GOLD, name)))


def test_constants():
    c = load("constants.json")
    assert (F.k_B, F.R, F.k_B_SI) == (c["k_B"], c["R"], c["k_B_SI"])


def test_stepsize_modify_value_matches_reference():
    """reference src/festim/stepsize.py:144-167."""
    for case in load("stepsize_modify_value.json"):
        st = F.Stepsize(**case["config"])
        assert st.modify_value(case["value"], case["nb_iterations"], case["t"]) == case["expected"]


def test_c_abi_stepsize_matches_reference(b200_lib):
    for case in load("stepsize_modify_value.json"):
        cfg = case["config"]
        ms = sorted(cfg.get("milestones", []))
        arr = (C.c_double * max(len(ms), 1))(*ms)
        got = b200_lib.fb_stepsize_modify(
            case["value"], case["nb_iterations"], case["t"], 1, cfg["growth_factor"],
            cfg["cutback_factor"], cfg["target_nb_iterations"], cfg.get("max_stepsize", 0.0) or 0.0,
            len(ms), arr, cfg.get("milestone_tolerance", 1e-5))
        assert got == case["expected"]


def test_oracle_convergence_test_matches_reference():
    """reference src/festim/helpers.py:408-426."""
    for seq in load("convergence_test.json"):
        f0 = seq["rows"][0]["fnorm"]
        for row in seq["rows"]:
            got = convergence_test(row["it"], row["snorm"], row["fnorm"], f0, seq["rtol"],
                                   seq["atol"], seq["stol"], seq["max_it"])
            assert got == row["reason"], (seq, row)


def test_reaction_term_matches_reference():
    """reference src/festim/reaction.py:117-224 (all three detrapping-rate modes)."""
    mat = F.Material(D_0=1.0, E_D=0.1)
    vol = F.VolumeSubdomain(id=1, material=mat)
    for case in load("reaction_term.json"):
        A, B = F.Species("A"), F.Species("B", mobile=False)
        A.solution, B.solution = ufl.Const(case["c_mobile"]), ufl.Const(case["c_trapped"])
        imp = F.ImplicitSpecies(n=case["n"], others=[B])
        imp.value_fenics = ufl.Const(case["n"])
        r = F.Reaction(reactant=[A, imp], product=B, k_0=case["k_0"], E_k=case["E_k"],
                       p_0=case["p_0"], E_p=case["E_p"], volume=vol)
        got = float(ufl.evaluate(r.reaction_term(case["T"]), {}))
        assert got == pytest.approx(case["expected"], rel=1e-14, abs=0.0)


def test_diffusion_coefficient_matches_reference():
    """reference src/festim/material.py:262-318."""
    for case in load("diffusion_coefficient.json"):
        m = F.Material(D_0=case["D_0"], E_D=case["E_D"])
        got = float(ufl.evaluate(m.get_diffusion_coefficient(None, case["T"], None), {}))
        assert got == pytest.approx(case["expected"], rel=1e-15)
