"""CPU pin of the algebra two kernels added in round 2 implement (no GPU needed):

* `back_embed3_stage1/2_kernel` (csrc/assemble.cuh): the adjoint of the N = 6 Kronecker embedding contracted factor by factor
  (A = sum_{i2,j2} W Y[i2,j2], B = sum_{i0,j0} W Y[i0,j0], then three 256-term contractions per output, and the same with dual
  numbers) -- restated here in NumPy with the kernel's index tables and compared with the oracle's adjoint
  (`_layer_vjp` / `_layer_vjp_tangent`, the restatement of what jax.grad / jax.jvp generate from transformations.py:405-490).
* the upper-half product of embedded complex matrices with the mirrored store (`GemmParams::mirror`, csrc/gemm_dmma.cuh)."""
import numpy as np
import pytest


def _digit_table(P, N, d, shift):
    """tab[r] = tensored index (i0 * q^2 + i1 * q + i2) of full index r: what build_embed_tables (csrc/otn_api.cu) packs."""
    D = d ** (2 * N)
    M = np.broadcast_to(np.arange(D, dtype=np.float64)[:, None], (D, D)).copy()        # M[rt, ct] = rt
    return P.convert_supertensored2liouvillianfull(M, N, d, shift_pbc=shift)[:, 0].astype(np.int64)


@pytest.mark.parametrize("layer", [0, 1])
def test_three_factor_back_embedding_matches_oracle_adjoint(oracle, layer):
    P = oracle
    N, d, q = 6, 2, 16
    D = d ** (2 * N)
    rng = np.random.default_rng(20 + layer)
    x = P.random_stiefel(8, 4, rng)
    eta = P.project(x, rng.standard_normal(x.shape))
    W = rng.standard_normal((D, D))
    Wd = rng.standard_normal((D, D))
    c, cd = P.ortho2choi(x), P.ortho2choi(eta)
    Yl, Yld = P.choi2super(c @ c.T), P.choi2super(cd @ c.T + c @ cd.T)

    # the kernel's view: rinv[i0 + 16 i1 + 256 i2] = full index whose local digits are (i0, i1, i2)
    tab = _digit_table(P, N, d, shift=(layer % 2 == 1))
    assert sorted(tab) == list(range(D))                         # a permutation (the factored kernels rely on it too)
    i0, i1, i2 = tab // (q * q), (tab // q) % q, tab % q
    rinv = np.empty(D, dtype=np.int64)
    rinv[i0 + 16 * i1 + 256 * i2] = np.arange(D)
    idx = rinv[(np.arange(q)[:, None, None] + 16 * np.arange(q)[None, :, None] + 256 * np.arange(q)[None, None, :])]   # [i0, i1, i2]

    def stage1(Wm, Wdm):
        Wt = Wm[idx.reshape(-1)][:, idx.reshape(-1)].reshape(q, q, q, q, q, q)             # [i0, i1, i2, j0, j1, j2]
        A = np.einsum("abcdef,cf->abde", Wt, Yl)                                            # contract factor 2
        B = np.einsum("abcdef,ad->bcef", Wt, Yl)                                            # contract factor 0
        if Wdm is None:
            return A, B, None, None
        Wdt = Wdm[idx.reshape(-1)][:, idx.reshape(-1)].reshape(q, q, q, q, q, q)
        Ad = np.einsum("abcdef,cf->abde", Wdt, Yl) + np.einsum("abcdef,cf->abde", Wt, Yld)
        Bd = np.einsum("abcdef,ad->bcef", Wdt, Yl) + np.einsum("abcdef,ad->bcef", Wt, Yld)
        return A, B, Ad, Bd

    def stage2(A, B, Ad=None, Bd=None):
        f0 = np.einsum("iuJv,uv->iJ", A, Yl); f1 = np.einsum("uivJ,uv->iJ", A, Yl); f2 = np.einsum("uivJ,uv->iJ", B, Yl)
        if Ad is None:
            return f0 + f1 + f2
        g0 = np.einsum("iuJv,uv->iJ", Ad, Yl) + np.einsum("iuJv,uv->iJ", A, Yld)
        g1 = np.einsum("uivJ,uv->iJ", Ad, Yl) + np.einsum("uivJ,uv->iJ", A, Yld)
        g2 = np.einsum("uivJ,uv->iJ", Bd, Yl) + np.einsum("uivJ,uv->iJ", B, Yld)
        return g0 + g1 + g2

    A, B, Ad, Bd = stage1(W, Wd)
    Wl, Wld = stage2(A, B), stage2(A, B, Ad, Bd)

    T = np.zeros((D, D))
    local, full = P.CostSpec(T, "local", N, d), P.CostSpec(np.zeros((q, q)), "full")
    z_ref = P._layer_vjp(local, layer, x, W)
    z = P._layer_vjp(full, layer, x, Wl)
    assert np.abs(z - z_ref).max() / np.abs(z_ref).max() < 1e-12
    zd_ref = P._layer_vjp_tangent(local, layer, x, eta, W, Wd)
    zd = P._layer_vjp(full, layer, x, Wld) + P._layer_vjp(full, layer, eta, Wl)          # what back_assemble does with (Wl, Wld)
    assert np.abs(zd - zd_ref).max() / np.abs(zd_ref).max() < 1e-12


def test_upper_half_of_an_embedded_product_mirrors_to_the_full_embedding():
    """emb(A) emb(B) = emb(A B), emb(A)^T emb(B) = emb(A^H B), emb(A) emb(B)^T = emb(A B^H): computing rows [0, D) and storing
    (r, c) -> (r + D, c + D) for c < D, -(r, c) -> (r + D, c - D) otherwise reproduces the whole product; the squared norm of the
    upper half is ||.||_F^2 of the complex matrix (the partial sums the complex handle reduces)."""
    rng = np.random.default_rng(3)
    D = 12
    A = rng.standard_normal((D, D)) + 1j * rng.standard_normal((D, D))
    B = rng.standard_normal((D, D)) + 1j * rng.standard_normal((D, D))
    emb = lambda Z: np.block([[Z.real, -Z.imag], [Z.imag, Z.real]])   # noqa: E731

    def mirrored(top):
        out = np.empty((2 * D, 2 * D))
        out[:D] = top
        out[D:, D:] = top[:, :D]
        out[D:, :D] = -top[:, D:]
        return out
    for left, right, want in ((emb(A), emb(B), A @ B), (emb(A).T, emb(B), A.conj().T @ B), (emb(A), emb(B).T, A @ B.conj().T)):
        top = (left @ right)[:D]
        np.testing.assert_allclose(mirrored(top), emb(want), atol=1e-12)
        assert abs(np.sum(top ** 2) - np.linalg.norm(want) ** 2) < 1e-10


def test_row_windows_tile_the_partial_sum_array():
    """A row-sharded product writes its per-tile partial sums at `rank * tiles(D, D / world) + local tile`: the windows of all ranks
    must tile the un-sharded launch's partial array exactly (otn_gemm_rows_tiles is plain host arithmetic: no GPU needed)."""
    from opentn_b200 import _lib
    lib = _lib.load()
    for D in (256, 1024, 4096):
        whole = lib.otn_gemm_rows_tiles(D, D)
        assert whole == (D // 64) ** 2
        for world in (2, 4, 8):
            if D % (64 * world) == 0:
                assert lib.otn_gemm_rows_tiles(D, D // world) * world == whole
