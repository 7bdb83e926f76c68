"""The Riemannian epilogue as the streaming kernel `riem_stream_kernel` (csrc/stiefel.cuh) evaluates it: every output is a
combination  H eta = zdraw*a + zraw*Cz + eta*Ce + x*Cx,  rgrad = zraw*b + x*Gn  with p x p coefficient matrices computed from the
five Grams x^T x, x^T zraw, x^T zdraw, x^T eta, eta^T zraw.  This CPU test pins that algebra (a NumPy transcription of the
kernel's stage 2) to the oracle's step-by-step formulation of stiefel.py:398-409, :468-475 and :23-30."""
import numpy as np
import pytest

from oracle import port as P


def coefficient_form(x, zraw, eta, zdraw, f, s, a0, a1):
    """zraw = f * egrad, zdraw = unscaled tangent of it, s = <R, Mdot>  (what the chain kernels hand to the epilogue)."""
    p = x.shape[1]
    inv_f = 1.0 / f
    kappa = s * inv_f ** 3
    c1, c2, kap = 0.5 * (1 / a1 - 2 / a0), 0.5 / a1, (a0 - a1) / a0
    XX, XZr, XZdr, XE, EZr = x.T @ x, x.T @ zraw, x.T @ zdraw, x.T @ eta, eta.T @ zraw
    XZ, EZ, XZd = inv_f * XZr, inv_f * EZr, inv_f * XZdr - kappa * XZr
    Gn = c1 * XZ - c2 * XZ.T
    Gd = c1 * (EZ + XZd) - c2 * (XZd.T + EZ.T)
    XN = XZ / a0 + XX @ Gn
    EN = EZ / a0 + XE.T @ Gn
    S = XE @ XN.T
    Gz = 0.5 * (EN + EN.T) - kap * (S + S.T)
    Ce = Gn + kap * XN.T
    Cx0 = Gd + Gz + kap * Gn @ XE.T
    Cz = (kap / a0) * XE.T
    T0 = XZd / a0 + XZ @ Cz + XE @ Ce + XX @ Cx0
    Cx = Cx0 - 0.5 * (T0 + T0.T)
    heta = zdraw * (inv_f / a0) + zraw @ (inv_f * Cz - (kappa / a0) * np.eye(p)) + eta @ Ce + x @ Cx
    rgrad = zraw * (inv_f / a0) + x @ Gn
    return rgrad, heta


@pytest.mark.parametrize("metric", ["canonical", "euclidean"])
@pytest.mark.parametrize("model,m,K", [("full", 3, 4), ("local", 4, 3)])
def test_coefficient_form_equals_oracle(metric, model, m, K):
    N, d = 4, 2
    rng = np.random.default_rng(3)
    if model == "full":
        T = P.kitaev_fullsites_target()
        spec = P.CostSpec(T, "full")
        n, p = 16 * K, 16
    else:
        T = P.exp_operator_dt(P.create_2local_liouvillians(P.get_kitaev_nn_linbladian(1.0), N, d, pbc=True)[0], 0.5)
        spec = P.CostSpec(T, "local", N, d)
        n, p = 4 * K, 4
    xs = [P.random_stiefel(n, p, rng) for _ in range(m)]
    etas = [P.project(x, rng.standard_normal((n, p))) for x in xs]
    a0, a1 = P.metric_alphas(metric)
    f, Z, Zd = P.egrad_jvp(spec, xs, etas)
    fo, go, ho = P.triple(spec, xs, etas, metric)
    # undo the scaling the device chain never applies: egrad = zraw / f, d(egrad) = zdraw / f - zraw * s / f^3
    s = 0.37 * f          # any value: the split of Zd into (zdraw, s) is not unique, the epilogue must be consistent for all
    for x, eta, z, zd, g_ref, h_ref in zip(xs, etas, Z, Zd, go, ho):
        zraw = f * z
        zdraw = f * (zd + zraw * s / f ** 3)
        g, h = coefficient_form(x, zraw, eta, zdraw, f, s, a0, a1)
        assert np.abs(g - g_ref).max() <= 1e-13 * max(1.0, np.abs(g_ref).max())
        assert np.abs(h - h_ref).max() <= 1e-12 * max(1.0, np.abs(h_ref).max())
        # tangency of both outputs
        assert np.abs(x.T @ h + h.T @ x).max() <= 1e-12 * max(1.0, np.abs(h).max())
