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


def _compile_c_demo(tmp_path):
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    from opentn_b200 import build
    lib = build.build(force=False)
    exe = str(tmp_path / "c_abi_demo")
    libdir = os.path.dirname(lib)
    cmd = ["gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "c_abi_demo.c"),
           "-L", libdir, "-lotn", f"-Wl,-rpath,{libdir}", "-lm", "-o", exe]
    subprocess.run(cmd, check=True, capture_output=True)
    return exe


def test_c_host_compiles_against_header(tmp_path):
    """include/otn.h is plain C (no torch / C++ types): a C host compiles and links against libotn.so with gcc."""
    exe = _compile_c_demo(tmp_path)
    assert os.path.exists(exe)


@pytest.mark.gpu
def test_c_host_runs_and_matches_golden(tmp_path, goldens, oracle):
    """The C host of examples/c_abi_demo.c on config 1: cost = the reference's printed 0.11294517951068227, gradient norm and
    <g, H g> equal the oracle's, three TR iterations reproduce the golden history."""
    import subprocess
    P = oracle
    exe = _compile_c_demo(tmp_path)
    Li = P.pspl_lindbladians(1.0)
    T = P.exp_operator_dt(P.create_2local_liouvillians(Li, 4, 2, pbc=True)[0], 1.0)
    xs0 = [np.asarray(x) for x in goldens["ref/xs0_pspl_n1_K10"]]
    np.ascontiguousarray(T.real).tofile(tmp_path / "t.bin")
    np.ascontiguousarray(np.stack(xs0)).tofile(tmp_path / "x.bin")
    out = subprocess.run([exe, str(tmp_path / "t.bin"), str(tmp_path / "x.bin"), "4", "3", "10"], check=True, capture_output=True, text=True)
    f, gnorm, ghg, f_after = (float(v) for v in out.stdout.split())
    assert abs(f - goldens["known/pspl_n1_tau1_K10_cost0"]) < 1e-13
    spec = P.CostSpec(T, "local", 4, 2)
    _, g = P.cost_rgrad(spec, xs0, "canonical")
    assert abs(gnorm - np.sqrt(sum((a * a).sum() for a in g))) < 1e-10
    want = P.canonical_dot(xs0, g, P.hvp(spec, xs0, g, "canonical"))
    assert abs(ghg - want) < 1e-9 * abs(want)
    assert abs(f_after - goldens["art/f_pauli_n1_tau1_rank10"][3]) < 1e-8
