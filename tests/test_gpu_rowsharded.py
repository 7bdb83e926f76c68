"""Row-sharded composition (config 5 building block): parity of otn_gemm_rows-based chains against torch fp64 (this IS a
floating-point GEMM kernel, so a plain torch reference applies) and against the batched handle path."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_row_window_products():
    import torch
    from opentn_b200.rowsharded import RowShardedChain
    torch.manual_seed(0)
    D = 512
    A = torch.randn(D, D, dtype=torch.float64, device="cuda")
    B = torch.randn(D, D, dtype=torch.float64, device="cuda")
    ch = RowShardedChain(D)
    for ta, tb, ref in ((False, False, A @ B), (False, True, A @ B.T), (True, False, A.T @ B)):
        got = ch.matmul(A, B, ta=ta, tb=tb)
        assert torch.allclose(got, ref, rtol=0, atol=1e-11)
    # explicit row windows (what the other ranks of a sharded run compute)
    for world, rank in ((2, 1), (4, 2), (8, 7)):
        ch.world, ch.rows, ch.row0 = 1, D // world, rank * (D // world)
        out = torch.zeros(D, D, dtype=torch.float64, device="cuda")
        ch._rows(A, B, out, True, False)
        ref = A.T @ B
        sl = slice(ch.row0, ch.row0 + ch.rows)
        assert torch.allclose(out[sl], ref[sl], rtol=0, atol=1e-11)
        assert float(out.abs().sum() - out[sl].abs().sum()) == 0.0       # nothing written outside the window


def test_complex_and_chain_cost(oracle):
    import torch
    from opentn_b200.rowsharded import RowShardedChain
    torch.manual_seed(1)
    D = 256
    ch = RowShardedChain(D)
    A = torch.randn(D, D, dtype=torch.complex128, device="cuda")
    B = torch.randn(D, D, dtype=torch.complex128, device="cuda")
    Cr, Ci = ch.matmul_complex((A.real.contiguous(), A.imag.contiguous()), (B.real.contiguous(), B.imag.contiguous()))
    ref = A @ B
    assert torch.allclose(Cr, ref.real, atol=1e-11) and torch.allclose(Ci, ref.imag, atol=1e-11)
    # chain + cost + cotangents against the oracle on a real N=4 problem
    P = oracle
    rng = np.random.default_rng(3)
    xs = [P.random_stiefel(12, 4, rng) for _ in range(3)]
    T = P.kitaev_fullsites_target()
    Ys = [torch.from_numpy(np.ascontiguousarray(y)).cuda() for y in P.local_layers_full(xs, 4, 2)]
    Tre = torch.from_numpy(np.ascontiguousarray(T.real)).cuda()
    spec = P.CostSpec(T, "local", 4, 2)
    assert abs(ch.cost(Ys, Tre) - spec(xs)) < 1e-12
    f, W = ch.cost_and_layer_cotangents(Ys, Tre)
    assert abs(f - spec(xs)) < 1e-12
    # pull the cotangents back with the oracle's adjoint and compare with its egrad
    Z = P.egrad(spec, xs)
    for l in range(3):
        z = P._layer_vjp(spec, l, xs[l], W[l].cpu().numpy())
        assert np.abs(z - Z[l]).max() / np.abs(Z[l]).max() < 1e-10
    M = ch.compose(Ys)
    np.testing.assert_allclose(M.cpu().numpy(), P.model_stiefel_local(xs, 4, 2), atol=1e-13)
