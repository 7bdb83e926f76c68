"""N = 6 sites (D = 4096) through the CUDA path: the reference's only N=6 anchors (trust_region_higher_sites.ipynb cells 4, 16:
cost 0.53944709 and ||Riemannian gradient coefficient vector|| = 15.82062161395626 at the Trotter point, K = 2, canonical)."""
import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.slow]


@pytest.fixture(scope="module")
def n6(oracle):
    P = oracle
    Lnn = P.get_kitaev_nn_linbladian(1.0)
    T6 = P.exp_operator_dt(P.create_2local_liouvillians(Lnn, 6, 2, pbc=True)[0], 4.0)     # ~15 s of host expm
    xs = [P.super2ortho(x.real, rank=2) for x in P.get_general_trotter_local_ansatz([Lnn], 4, n=1)]
    return T6, xs


@pytest.mark.parametrize("dense", [False, True])
def test_n6_cost_and_gradient_norm(n6, goldens, dense):
    from opentn_b200 import StiefelFit
    from opentn_b200 import stiefel as S
    T6, xs = n6
    fit = StiefelFit(T6, model="local", N=6, d=2, metric="canonical", dense=dense)
    f, g = fit.cost_grad(xs)
    assert abs(f - goldens["known/n6_cost0"]) < 5e-9
    assert abs(f - 0.5394470932876375) < 1e-12                       # SURVEY 8c (reference cost code under the shim)
    # coefficient-vector norm == canonical norm of the tangent matrices (basis independent)
    gn = np.sqrt(sum(S.canonical_metric(gl, gl, x) for gl, x in zip(g, xs)))
    assert abs(gn - goldens["known/n6_gradnorm"]) / goldens["known/n6_gradnorm"] < 1e-11
    # one HVP: self-adjointness <w, H v> = <v, H w> in the canonical metric
    rng = np.random.default_rng(0)
    v = np.array([S.project(x, rng.standard_normal(x.shape)) for x in xs])
    w = np.array([S.project(x, rng.standard_normal(x.shape)) for x in xs])
    Hv, Hw = fit.hvp_cached(v), fit.hvp_cached(w)
    a = sum(S.canonical_metric(w[l], Hv[l], xs[l]) for l in range(3))
    b = sum(S.canonical_metric(v[l], Hw[l], xs[l]) for l in range(3))
    assert abs(a - b) < 1e-9 * abs(a)
    fit.close()


def test_n6_expm_on_device(n6, oracle):
    """The 4096 x 4096 Liouvillian propagator (scipy: ~15 s here, 1m28 on the reference author's machine,
    experiments/gradient.ipynb cell 15) through otn_expm."""
    import time
    from opentn_b200 import transformations as T
    T6, _ = n6
    L = oracle.create_2local_liouvillians(oracle.get_kitaev_nn_linbladian(1.0), 6, 2, pbc=True)[0]
    t0 = time.perf_counter()
    got = T.exp_operator_dt(L, 4.0, backend="cuda")
    dt = time.perf_counter() - t0
    assert np.abs(got - T6).max() / np.abs(T6).max() < 1e-12
    print(f"otn_expm 4096^2: {dt:.2f} s incl. host<->device copies")


def test_n6_rowsharded_cost_and_cotangents(n6, oracle, goldens):
    """Config 5's building block at its real size: the row-sharded D = 4096 chain (RowShardedChain, world 1 here; two ranks in
    test_gpu_rowsharded_multi.py) against the reference's N=6 cost anchor and against the handle path's gradient: the layer
    cotangents W_l are pulled back with the oracle's adjoint of the Kronecker embedding and compared with StiefelFit's egrad-based
    Riemannian gradient."""
    import torch
    from opentn_b200 import StiefelFit
    from opentn_b200.rowsharded import RowShardedChain
    P = oracle
    T6, xs = n6
    Ys = [torch.from_numpy(np.ascontiguousarray(np.real(y))).cuda() for y in P.local_layers_full(xs, 6, 2)]
    Tre = torch.from_numpy(np.ascontiguousarray(T6.real)).cuda()
    Tim = torch.from_numpy(np.ascontiguousarray(T6.imag)).cuda() if np.abs(T6.imag).max() > 0 else None
    ch = RowShardedChain(4096)
    f = ch.cost(Ys, Tre, Tim)
    assert abs(f - goldens["known/n6_cost0"]) < 5e-9
    assert abs(f - 0.5394470932876375) < 1e-11
    if Tim is None:
        f2, W = ch.cost_and_layer_cotangents(Ys, Tre)
        assert abs(f2 - f) < 1e-12
        spec = P.CostSpec(T6, "local", 6, 2)
        fit = StiefelFit(T6, model="local", N=6, d=2, metric="euclidean")
        _, g = fit.cost_grad(xs)
        for l in range(len(xs)):
            z = P._layer_vjp(spec, l, xs[l], W[l].cpu().numpy())
            rg = P.gradient_ambient2riemannian(xs[l], z, 1.0, 1.0)
            assert np.abs(rg - g[l]).max() / np.abs(g[l]).max() < 1e-9
        fit.close()


def test_n6_dense_triple_equals_factored(n6, oracle):
    """The dense D = 4096 path (embedding, chain, three-factor back-embedding `back_embed3_*`, pull-back) against the factored
    evaluation of the same unit of work (which tests/test_gpu_factored.py pins to the oracle): f, Riemannian gradient and
    Hessian-vector product at a random point and tangent, K = 4, both kernels of the back-embedding (OTN_BACK_EMBED_V1 = the
    gather kernel)."""
    import os
    from opentn_b200 import StiefelFit
    T6, _ = n6
    rng = np.random.default_rng(4)
    xs = np.array([oracle.random_stiefel(16, 4, rng) for _ in range(3)])
    et = np.array([oracle.project(x, rng.standard_normal(x.shape)) for x in xs])
    ref = StiefelFit(T6, model="local", N=6, d=2, metric="canonical", factored=True)
    f0, g0, h0 = ref.triple(xs, et)
    ref.close()
    for v1 in (False, True):
        if v1:
            os.environ["OTN_BACK_EMBED_V1"] = "1"
        try:
            fit = StiefelFit(T6, model="local", N=6, d=2, metric="canonical", dense=True)
            f, g, h = fit.triple(xs, et)
            fc, gc = fit.cost_grad(xs)              # gradient-only pass (non-tangent back-embedding)
            fit.close()
        finally:
            os.environ.pop("OTN_BACK_EMBED_V1", None)
        assert abs(f - f0) < 1e-11 * abs(f0) and abs(fc - f0) < 1e-11 * abs(f0)
        assert np.abs(g - g0).max() / np.abs(g0).max() < 1e-10
        assert np.abs(gc - g0).max() / np.abs(g0).max() < 1e-10
        assert np.abs(h - h0).max() / np.abs(h0).max() < 1e-10


def test_n6_split_k_pairs_match_unsplit(n6, oracle):
    """D = 4096 products run as split-K CTA pairs (thread-block clusters of two, accumulators handed over through distributed
    shared memory; csrc/gemm_dmma.cuh).  OTN_SPLITK=0 switches them off: same cost / gradient / HVP to rounding."""
    import os
    from opentn_b200 import StiefelFit
    T6, _ = n6
    rng = np.random.default_rng(5)
    xs = np.array([oracle.random_stiefel(16, 4, rng) for _ in range(3)])
    et = np.array([oracle.project(x, rng.standard_normal(x.shape)) for x in xs])
    res = []
    for split in ("1", "0"):
        os.environ["OTN_SPLITK"] = split
        try:
            fit = StiefelFit(T6, model="local", N=6, d=2, metric="canonical", dense=True)
            res.append(fit.triple(xs, et))
            fit.close()
        finally:
            os.environ.pop("OTN_SPLITK", None)
    (f1, g1, h1), (f0, g0, h0) = res
    assert abs(f1 - f0) < 1e-13 * abs(f0)
    assert np.abs(g1 - g0).max() / np.abs(g0).max() < 1e-12 and np.abs(h1 - h0).max() / np.abs(h0).max() < 1e-12
    assert not (np.array_equal(g1, g0) and np.array_equal(h1, h0))        # the switch really selects another kernel
