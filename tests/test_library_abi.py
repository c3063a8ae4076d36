"""The C-ABI library builds for sm_100a, loads, and exports every symbol that
include/festim_b200.h declares (no compute calls: runs without a GPU)."""
import os
import re

import festim_b200.backend as backend

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "festim_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fb_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_exported(b200_lib):
    names = declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(b200_lib, name), name
    assert sorted(backend.EXPORTED_SYMBOLS) == names


def test_stepsize_rule_matches_host(b200_lib):
    """fb_stepsize_modify == festim_b200.Stepsize.modify_value (reference
    src/festim/stepsize.py:144-167); pure host arithmetic, no GPU needed."""
    import ctypes as C

    import numpy as np

    from festim_b200 import Stepsize

    rng = np.random.default_rng(0)
    ms = sorted(rng.uniform(0, 10, 5).tolist())
    st = Stepsize(0.1, growth_factor=1.2, cutback_factor=0.8, target_nb_iterations=4,
                  max_stepsize=1.5, milestones=ms)
    arr = (C.c_double * len(ms))(*ms)
    for _ in range(200):
        value = float(rng.uniform(0.01, 3))
        t = float(rng.uniform(0, 11))
        its = int(rng.integers(1, 8))
        ref = st.modify_value(value, its, t)
        got = b200_lib.fb_stepsize_modify(value, its, t, 1, 1.2, 0.8, 4, 1.5, len(ms), arr, 1e-5)
        assert got == ref


def test_no_gpu_fails_loudly(b200_lib):
    import ctypes as C

    import torch

    if torch.cuda.is_available():
        return
    h = C.c_void_p()
    assert b200_lib.fb_create(0, None, C.byref(h)) != 0
