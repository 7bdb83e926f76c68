"""Independent autodiff check for the oracle's closed-form derivative sweeps (test helper, CPU, float64).

A torch transcription of the reference cost closure
`lambda xi: frobenius_norm(model(xi), target)` (optimization.py:56-61,106-151; transformations.py:716-749,
405-490, 841-851), differentiated with torch.func exactly the way the reference differentiates it with JAX:
`grad` (stiefel.py:423) and `jvp(grad)` (stiefel.py:520).
"""
import numpy as np
import torch
from torch.func import grad, jvp

from oracle import port as P


def _ortho2super(x):
    dim = x.shape[1]
    K = x.shape[0] // dim
    c = x.reshape(dim, K, dim).swapaxes(1, 2).reshape(dim * dim, K)
    C = c @ c.T
    return C.reshape(dim, dim, dim, dim).swapaxes(1, 2).reshape(dim * dim, dim * dim)


def _embed(Yl, N, d, shift):
    out = Yl
    for _ in range(N // 2 - 1):
        out = torch.kron(out, Yl)
    src, dst = P.get_indices_supertensored2liouvillianfull(N)
    t = out.reshape([d] * (4 * N))
    t = torch.movedim(t, src, dst)
    if shift:
        a = list(range(N))
        b = list(range(N, 2 * N))
        idx = a[1:] + a[:1] + b[1:] + b[:1]
        t = t.permute(idx + [i + 2 * N for i in idx])
    return t.reshape(out.shape)


def make_cost(target, model, N=None, d=2):
    Tr = torch.tensor(np.real(target), dtype=torch.float64)
    Ti = torch.tensor(np.imag(target), dtype=torch.float64)

    def cost(xs):
        Ys = [_ortho2super(x) for x in xs]
        if model == "local":
            Ys = [_embed(Y, N, d, i % 2 == 1) for i, Y in enumerate(Ys)]
        M = Ys[0]
        for Y in Ys[1:]:
            M = Y @ M
        return torch.sqrt(torch.sum((M - Tr) ** 2) + torch.sum(Ti ** 2))

    return cost


def egrad_and_jvp(target, model, xs, etas, N=None, d=2):
    cost = make_cost(target, model, N, d)
    xt = tuple(torch.tensor(x, dtype=torch.float64) for x in xs)
    et = tuple(torch.tensor(e, dtype=torch.float64) for e in etas)
    g, gd = jvp(grad(cost), (xt,), (et,))
    return float(cost(xt)), [t.numpy() for t in g], [t.numpy() for t in gd]
