"""CPU-side checks of the drop-in boundary: libotn.so loads, exports every symbol include/otn.h declares, the ctypes
table in opentn_b200/_lib.py covers exactly those symbols, and calls fail loudly (no fallback) without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "otn.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(otn_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from opentn_b200 import _lib
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/otn.h but not exported by libotn.so"
    assert sorted(_lib.SIGNATURES) == names, "opentn_b200/_lib.py SIGNATURES must list exactly the header's entry points"
    assert lib.otn_version() == 100


def test_struct_layouts_match_header():
    from opentn_b200 import _lib
    prm = _lib.TrParams()
    _lib.load().otn_tr_default_params(ctypes.byref(prm))
    # defaults of trust_region_rcopt.py:36-41, 107-109
    assert (prm.rho_trust, prm.radius_init, prm.maxradius, prm.niter) == (0.125, 0.01, 0.1, 20)
    assert (prm.tcg_maxiter, prm.tcg_abstol, prm.tcg_reltol) == (0, 1e-8, 1e-6)
    assert ctypes.sizeof(_lib.TrReport) == 9 * 8 + 8


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point must raise -- never silently compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from opentn_b200 import OtnError, StiefelFit
    from opentn_b200 import stiefel, transformations
    fit = StiefelFit(np.eye(16), model="full", N=2, d=2)
    x = np.zeros((1, 1, 8, 4))
    with pytest.raises(OtnError):
        fit.cost(x)
    with pytest.raises(OtnError):
        transformations.ortho2super(np.zeros((8, 4)))
    with pytest.raises(OtnError):
        stiefel.project(np.zeros((8, 4)), np.zeros((8, 4)))


def test_product_never_imports_oracle():
    """The product package must not reach into oracle/ (it is the checker, not the product)."""
    pkg = os.path.join(ROOT, "opentn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports the oracle"
                assert "/root/reference" not in txt, f"{f} reads the reference tree"


def test_argument_validation_messages():
    from opentn_b200 import StiefelFit
    with pytest.raises(ValueError, match="not a valid metric value"):        # stiefel.py:461
        StiefelFit(np.eye(16), model="full", N=2, d=2, metric="riemann")
    with pytest.raises(ValueError):
        StiefelFit(np.eye(10), model="full", d=2)
    fit = StiefelFit(np.eye(256), model="local", N=4, d=2)
    with pytest.raises(ValueError, match="does not fit model"):
        fit._prep(np.zeros((3, 40, 5)))
    with pytest.raises(NotImplementedError, match="complex isometries"):
        fit._prep(np.zeros((3, 40, 4)) + 1j)
