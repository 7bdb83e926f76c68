"""Pins the complex-isometry oracle (oracle/complex_port.py): closed-form sweeps vs torch autodiff of an independent transcription
of the complex cost, reduction to the real oracle on real inputs, and the geometric identities of the complex Stiefel manifold."""
import numpy as np
import pytest
import torch

from oracle import complex_port as CP
from oracle import port as P


def _problem(m=3, K=2, dim=4, seed=0):
    rng = np.random.default_rng(seed)
    xs = [CP.random_stiefel(dim * K, dim, rng) for _ in range(m)]
    etas = [CP.project(x, rng.standard_normal(x.shape) + 1j * rng.standard_normal(x.shape)) for x in xs]
    T = CP.model([CP.random_stiefel(dim * K, dim, rng) for _ in range(m)]) + 0.05 * (rng.standard_normal((dim * dim,) * 2) + 1j * rng.standard_normal((dim * dim,) * 2))
    return T, xs, etas


def _torch_cost(T):
    Tt = torch.from_numpy(T)

    def f(re, im):                      # re, im: [m, n, p] real leaves
        M = None
        for l in range(re.shape[0]):
            x = torch.complex(re[l], im[l])
            n, dim = x.shape
            K = n // dim
            E = x.reshape(dim, K, dim)
            Y = torch.einsum("akc,bkd->abcd", E, E.conj()).reshape(dim * dim, dim * dim)
            M = Y if M is None else Y @ M
        return torch.linalg.norm(M - Tt)
    return f


@pytest.mark.parametrize("m,K,dim", [(3, 2, 4), (2, 3, 2), (1, 2, 4)])
def test_egrad_and_hvp_vs_torch_autodiff(m, K, dim):
    T, xs, etas = _problem(m, K, dim, seed=m * 10 + K)
    f = _torch_cost(T)
    re = torch.tensor(np.array([x.real for x in xs])); im = torch.tensor(np.array([x.imag for x in xs]))
    er = torch.tensor(np.array([e.real for e in etas])); ei = torch.tensor(np.array([e.imag for e in etas]))
    grad = torch.func.grad(f, argnums=(0, 1))
    (gr, gi), (hr, hi) = torch.func.jvp(grad, (re, im), (er, ei))
    fo, Z, Zd = CP.egrad_jvp(T, xs, etas)
    assert abs(fo - float(f(re, im))) < 1e-13
    Zt = gr.numpy() + 1j * gi.numpy()
    Zdt = hr.numpy() + 1j * hi.numpy()
    for l in range(m):
        assert np.abs(Z[l] - Zt[l]).max() / np.abs(Zt[l]).max() < 1e-12
        assert np.abs(Zd[l] - Zdt[l]).max() / np.abs(Zdt[l]).max() < 1e-11
    # cost-only sweep agrees
    f2, Z2 = CP.egrad_jvp(T, xs)
    assert f2 == fo and all(np.array_equal(a, b) for a, b in zip(Z, Z2))


def test_rgrad_hvp_vs_autodiff_of_the_riemannian_gradient():
    """H eta = P_X(connection(D rgrad[eta], rgrad, eta)): D rgrad[eta] from torch jvp of (egrad -> rgrad map) as a whole."""
    T, xs, etas = _problem(2, 2, 4, seed=3)
    f = _torch_cost(T)
    a0, a1 = 1.0, 0.5

    def rgrad(re, im):
        gr, gi = torch.func.grad(f, argnums=(0, 1))(re, im)
        x = torch.complex(re, im); z = torch.complex(gr, gi)
        xh = x.conj().transpose(-1, -2)
        nu = z / a0 + 0.5 * (1 / a1 - 2 / a0) * x @ (xh @ z) - 0.5 / a1 * x @ (z.conj().transpose(-1, -2) @ x)
        return nu.real, nu.imag
    re = torch.tensor(np.array([x.real for x in xs])); im = torch.tensor(np.array([x.imag for x in xs]))
    er = torch.tensor(np.array([e.real for e in etas])); ei = torch.tensor(np.array([e.imag for e in etas]))
    (nr, ni), (dr, di) = torch.func.jvp(rgrad, (re, im), (er, ei))
    fo, g, h = CP.triple(T, xs, etas, "canonical")
    for l in range(2):
        nu = nr[l].numpy() + 1j * ni[l].numpy()
        dnu = dr[l].numpy() + 1j * di[l].numpy()
        assert np.abs(g[l] - nu).max() < 1e-12
        want = CP.project(xs[l], CP.riemannian_connection(dnu, nu, etas[l], xs[l], a0, a1))
        assert np.abs(h[l] - want).max() / np.abs(want).max() < 1e-11
        # tangency of the results
        for v in (g[l], h[l]):
            A = xs[l].conj().T @ v
            assert np.abs(A + A.conj().T).max() < 1e-12


def test_reduces_to_real_oracle():
    rng = np.random.default_rng(1)
    xs = [P.random_stiefel(8, 4, rng) for _ in range(3)]
    etas = [P.project(x, rng.standard_normal(x.shape)) for x in xs]
    T = P.exp_operator_dt(P.lindbladian2super([P.get_kitaev_nn_linbladian(1.0)]), 1.0)      # N = 2 full-sites, complex dtype
    spec = P.CostSpec(T, "full")
    for metric in ("canonical", "euclidean", (2.0, 0.7)):
        fo, go, ho = P.triple(spec, xs, etas, metric)
        fc, gc, hc = CP.triple(T, [x.astype(complex) for x in xs], [e.astype(complex) for e in etas], metric)
        assert abs(fo - fc) < 1e-14
        for a, b in zip(go + ho, gc + hc):
            assert np.abs(b.imag).max() < 1e-14 and np.abs(a - b.real).max() < 1e-13
    assert abs(P.canonical_dot(xs, etas, go) - CP.canonical_dot(xs, etas, go)) < 1e-14


def test_complex_coefficient_vectors_carry_the_canonical_metric():
    rng = np.random.default_rng(2)
    x = CP.random_stiefel(10, 4, rng)
    a = CP.project(x, rng.standard_normal(x.shape) + 1j * rng.standard_normal(x.shape))
    b = CP.project(x, rng.standard_normal(x.shape) + 1j * rng.standard_normal(x.shape))
    va, vb = CP.parametrization_from_tangent(x, a), CP.parametrization_from_tangent(x, b)
    assert va.shape == (4 * 4 + 2 * 6 * 4,)                                       # p^2 + 2 (n - p) p real degrees of freedom
    assert abs(va @ vb - CP.canonical_dot([x], [a], [b])) < 1e-13
    np.testing.assert_allclose(CP.tangent_from_parametrization(x, va), a, atol=1e-13)
    y = CP.retract([x], [0.1 * a])[0]
    np.testing.assert_allclose(y.conj().T @ y, np.eye(4), atol=1e-13)
