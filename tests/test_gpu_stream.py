"""Streamed submissions (otn_triple_submit / otn_triple_wait): results identical to the synchronous call and to the oracle,
with several batches in flight, batches of one chunk (workspace slices of neighbouring calls coincide) and of several chunks."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
N, d = 4, 2


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


def _pinned(shape):
    import torch
    return torch.empty(shape, dtype=torch.float64).pin_memory()


def _batch(P, B, K, seed, local=False):
    rng = np.random.default_rng(seed)
    n, p = (4 * K, 4) if local else (16 * K, 16)
    m = 3
    g = rng.standard_normal((B, m, n, p))
    q, r = np.linalg.qr(g)
    x = q * np.sign(np.diagonal(r, axis1=-2, axis2=-1))[..., None, :]
    z = rng.standard_normal(x.shape)
    s = np.swapaxes(x, -1, -2) @ z
    eta = z - x @ (0.5 * (s + np.swapaxes(s, -1, -2)))
    return np.ascontiguousarray(x), np.ascontiguousarray(eta)


@pytest.mark.parametrize("B", [3, 300, 700])
def test_streamed_equals_synchronous_and_oracle(oracle, B):
    """6 submissions over 3 rotating sets of pinned host arrays, never more than 3 in flight; every result is compared with the
    synchronous otn_triple on the same inputs (bitwise) and instance 0 / B-1 of each with the oracle."""
    import torch
    from opentn_b200 import StiefelFit
    P = oracle
    K = 4
    T = P.kitaev_fullsites_target()
    spec = P.CostSpec(T, "full")
    fit = StiefelFit(T, model="full", N=N, d=d, metric="canonical", max_batch=B)
    nset, nsub = 3, 6
    shape = (B, 3, 16 * K, 16)
    xs = [_pinned(shape) for _ in range(nset)]
    es = [_pinned(shape) for _ in range(nset)]
    outs = [(_pinned((B,)), _pinned(shape), _pinned(shape)) for _ in range(nset)]
    inputs = [_batch(P, B, K, 100 + i) for i in range(nsub)]
    results = [None] * nsub
    tickets = [None] * nsub

    def harvest(i):
        fit.triple_wait(tickets[i])
        results[i] = tuple(o.numpy().copy() for o in outs[i % nset])

    for i in range(nsub):
        if i >= nset:
            harvest(i - nset)               # frees host set i % nset
        xs[i % nset].copy_(torch.from_numpy(inputs[i][0]))
        es[i % nset].copy_(torch.from_numpy(inputs[i][1]))
        for o in outs[i % nset]:
            o.fill_(float("nan"))
        tickets[i] = fit.triple_submit(xs[i % nset], es[i % nset], outs[i % nset])
    for i in range(nsub - nset, nsub):
        harvest(i)

    for i in range(nsub):
        f, g, h = fit.triple(inputs[i][0], inputs[i][1])          # synchronous host path
        fs, gs, hs = results[i]
        assert np.array_equal(fs, f) and np.array_equal(gs, g) and np.array_equal(hs, h), f"submission {i}"
        for b in (0, B - 1):
            fo, go, ho = P.triple(spec, list(inputs[i][0][b]), list(inputs[i][1][b]), "canonical")
            assert abs(fs[b] - fo) / fo < 1e-12
            assert rel(gs[b], np.array(go)) < 1e-10 and rel(hs[b], np.array(ho)) < 1e-10
    fit.close()


def test_other_calls_drain_outstanding_submissions(oracle):
    """A synchronous call on the same fit while a submission is in flight waits for it (shared workspaces) and both results
    are right; numpy (pageable) host arrays are accepted too; waiting twice for a ticket is an error."""
    from opentn_b200 import OtnError, StiefelFit
    P = oracle
    K, B = 2, 600
    T = P.kitaev_fullsites_target()
    fit = StiefelFit(T, model="full", N=N, d=d, metric="euclidean", max_batch=B)
    x, e = _batch(P, B, K, 7)
    x2, _ = _batch(P, B, K, 8)
    out = (np.full(B, np.nan), np.full(x.shape, np.nan), np.full(x.shape, np.nan))
    t = fit.triple_submit(x, e, out)
    f2 = fit.cost(x2)                        # drains the pipeline first
    fit.triple_wait(t)
    with pytest.raises(OtnError):
        fit.triple_wait(t)
    f, g, h = fit.triple(x, e)
    assert np.array_equal(out[0], f) and np.array_equal(out[1], g) and np.array_equal(out[2], h)
    assert np.array_equal(f2, fit.cost(x2))
    with pytest.raises(ValueError):
        import torch
        fit.triple_submit(torch.from_numpy(x).cuda(), torch.from_numpy(e).cuda(), out)
    fit.close()


def test_streamed_two_site_model(oracle):
    """2-site (factored) model through the streamed path."""
    from opentn_b200 import StiefelFit
    P = oracle
    K, B = 4, 520
    Lk = P.create_2local_liouvillians(P.get_kitaev_nn_linbladian(1.0), N, d, pbc=True)[0]
    T = P.exp_operator_dt(Lk, 0.5)
    fit = StiefelFit(T, model="local", N=N, d=d, metric="canonical", max_batch=B)
    x, e = _batch(P, B, K, 21, local=True)
    outs = [(np.empty(B), np.empty(x.shape), np.empty(x.shape)) for _ in range(2)]
    t0 = fit.triple_submit(x, e, outs[0])
    t1 = fit.triple_submit(x, e, outs[1])
    fit.triple_wait()                        # all outstanding
    f, g, h = fit.triple(x, e)
    for o in outs:
        assert np.array_equal(o[0], f) and np.array_equal(o[1], g) and np.array_equal(o[2], h)
    spec = P.CostSpec(T, "local", N, d)
    fo, go, ho = P.triple(spec, list(x[5]), list(e[5]), "canonical")
    assert abs(f[5] - fo) / fo < 1e-12 and rel(g[5], np.array(go)) < 1e-10 and rel(h[5], np.array(ho)) < 1e-10
    fit.close()


def test_side_stream_is_bitwise_neutral(oracle, monkeypatch):
    """The antisymmetric block of every chain product runs on a side stream (fork / join around the non-GEMM kernels).  Same
    kernels, same data: U1 and a 3-iteration batched trust region must be bitwise identical with the side stream switched off."""
    from opentn_b200 import StiefelFit
    P = oracle
    K, B = 4, 40
    T = P.kitaev_fullsites_target()
    x, e = _batch(P, B, K, 33)
    res = []
    for off in (False, True):
        if off:
            monkeypatch.setenv("OTN_NO_SIDE_STREAM", "1")
        fit = StiefelFit(T, model="full", N=N, d=d, metric="canonical", max_batch=B, dense=False)
        f, g, h = fit.triple(x, e)
        xt, rep = fit.tr_run(x, niter=3)
        res.append((f, g, h, xt, rep["f_iter"], rep["rho"]))
        fit.close()
    for a, b in zip(*res):
        assert np.array_equal(np.asarray(a), np.asarray(b))
    fo, go, ho = P.triple(P.CostSpec(T, "full"), list(x[7]), list(e[7]), "canonical")
    assert abs(res[0][0][7] - fo) / fo < 1e-12 and rel(res[0][1][7], np.array(go)) < 1e-10 and rel(res[0][2][7], np.array(ho)) < 1e-10


@pytest.mark.parametrize("model,dense,m,K,B", [("full", False, 3, 16, 300), ("full", False, 5, 4, 64), ("full", True, 3, 4, 40),
                                               ("local", False, 4, 4, 64), ("local", True, 3, 2, 8)])
def test_back_stream_is_bitwise_neutral(oracle, monkeypatch, model, dense, m, K, B):
    """The cotangent pull-backs (back_layer) run on a back stream behind the GEMM that produced their input while the chain
    carries on (otn_problem::s_back).  Same kernels, same data, ordered by events: U1, cost+grad, cached HVPs and a batched
    trust region must be bitwise identical with the back stream switched off, several times in a row (races show up as flicker)."""
    import torch
    from opentn_b200 import StiefelFit
    P = oracle
    if model == "full":
        T = P.kitaev_fullsites_target()
        n, p = 16 * K, 16
    else:
        T = P.exp_operator_dt(P.create_2local_liouvillians(P.get_kitaev_nn_linbladian(1.0), N, d, pbc=True)[0], 0.5)
        n, p = 4 * K, 4
    rng = np.random.default_rng(77)
    x = np.array([[P.random_stiefel(n, p, rng) for _ in range(m)] for _ in range(B)])
    e = np.array([[P.project(a, rng.standard_normal((n, p))) for a in xb] for xb in x])
    res = []
    for off in (False, True):
        if off:
            monkeypatch.setenv("OTN_NO_BACK_STREAM", "1")
        fit = StiefelFit(T, model=model, N=N, d=d, metric="canonical", max_batch=B, dense=dense)
        runs = []
        for _ in range(1 if off else 4):
            f, g, h = fit.triple(x, e)
            f2, g2 = fit.cost_grad(x)
            h2 = fit.hvp_cached(e)
            # device-resident inputs: one otn_triple on the handle's own streams (host arrays go through the chunk pipeline)
            fd, gd, hd = fit.triple(torch.from_numpy(x).cuda(), torch.from_numpy(e).cuda())
            runs.append((f, g, h, f2, g2, h2, fd.cpu().numpy(), gd.cpu().numpy(), hd.cpu().numpy()))
        np.testing.assert_allclose(runs[0][6], runs[0][0], rtol=1e-13)
        np.testing.assert_allclose(runs[0][8], runs[0][2], rtol=0, atol=1e-13 * np.abs(runs[0][2]).max())
        for r in runs[1:]:
            for a, b in zip(runs[0], r):
                assert np.array_equal(a, b)
        xt, rep = fit.tr_run(x, niter=2)
        res.append(runs[0] + (xt, rep["f_iter"], rep["rho"]))
        fit.close()
    for a, b in zip(*res):
        assert np.array_equal(np.asarray(a), np.asarray(b))
    spec = P.CostSpec(T, "full") if model == "full" else P.CostSpec(T, "local", N, d)
    fo, go, ho = P.triple(spec, list(x[3]), list(e[3]), "canonical")
    assert abs(res[0][0][3] - fo) / fo < 1e-12 and rel(res[0][1][3], np.array(go)) < 1e-10 and rel(res[0][2][3], np.array(ho)) < 1e-10


def test_streamed_device_arrays(oracle):
    """otn_triple_submit with device-resident arrays: no copies, chunks on the two compute streams, several submissions in flight
    on rotating output sets; bitwise equal to the blocking call, and mixed host / device arrays are refused."""
    import torch
    from opentn_b200 import StiefelFit
    P = oracle
    K, B = 4, 1024
    T = P.kitaev_fullsites_target()
    x, e = _batch(P, B, K, 91)
    xd, ed = torch.from_numpy(x).cuda(), torch.from_numpy(e).cuda()
    xr, er = torch.flip(xd, dims=[0]).contiguous(), torch.flip(ed, dims=[0]).contiguous()
    fit = StiefelFit(T, model="full", N=N, d=d, metric="canonical", max_batch=B)
    f, g, h = fit.triple(xd, ed)
    outs = [tuple(torch.zeros_like(t) for t in (f, g, h)) for _ in range(2)]
    torch.cuda.synchronize()                         # torch's stream is not the handle's: its kernels must have finished
    tickets = []
    for i in range(6):                               # 6 submissions, at most 2 in flight, inputs alternate between x and flipped x
        if i >= 2:
            fit.triple_wait(tickets[i - 2])
            ref = (f, g, h) if i % 2 == 0 else tuple(torch.flip(t, dims=[0]) for t in (f, g, h))
            for a, b in zip(outs[i % 2], ref):
                assert torch.equal(a, b)
            for a in outs[i % 2]:
                a.zero_()
            torch.cuda.synchronize()
        tickets.append(fit.triple_submit(xd if i % 2 == 0 else xr, ed if i % 2 == 0 else er, outs[i % 2]))
    fit.triple_wait()
    assert torch.equal(outs[0][0], f) and torch.equal(outs[0][1], g) and torch.equal(outs[0][2], h)
    assert torch.equal(torch.flip(outs[1][1], dims=[0]), g) and torch.equal(torch.flip(outs[1][2], dims=[0]), h)
    fo, go, ho = P.triple(P.CostSpec(T, "full"), list(x[600]), list(e[600]), "canonical")
    assert abs(float(f[600]) - fo) / fo < 1e-12 and rel(g[600].cpu().numpy(), np.array(go)) < 1e-10 and rel(h[600].cpu().numpy(), np.array(ho)) < 1e-10
    with pytest.raises(ValueError, match="all host or all device"):
        fit.triple_submit(xd, ed, (np.empty(B), np.empty(x.shape), np.empty(x.shape)))
    fit.close()
