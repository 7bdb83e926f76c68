"""Row-sharded composition of large superoperators over the GPUs of one node (BASELINE config 5).

A single N = 6 instance has D = 4096: one chain product is 137 GFLOP (real) and is not instance-parallel.  Rank g owns the rows
[g D/G, (g+1) D/G) of every product  P_l = Y_l P_{l-1}  (it needs those rows of Y_l and all of P_{l-1}), computes them with the
same DMMA kernel as the batched path (`otn_gemm_rows`), and the row blocks are all-gathered (NCCL over NVLink) before the next
step.  The cost needs one scalar all-reduce of the partial squared norms.  The reverse sweep of the gradient shards the same
way (`W_l = G_l P_{l-1}^T` by rows of G_l; `G_{l-1} = Y_l^T G_l` by rows of the result = columns of Y_l).

Complex operands (the reference stores targets as complex128) are handled as (re, im) pairs: one complex product = two
dual-product launches,  C_re = A_re B_re + (-A_im) B_im,  C_im = A_re B_im + A_im B_re.
"""
from __future__ import annotations

import ctypes as C

from . import _lib
from ._lib import check


class RowShardedChain:
    def __init__(self, D: int, device=None, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.group = group
        if dist.is_available() and dist.is_initialized():
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        else:
            self.rank, self.world = 0, 1
        if D % (64 * self.world) != 0:
            raise ValueError("D must be a multiple of 64 * world_size")
        self.D = D
        self.rows = D // self.world
        self.row0 = self.rank * self.rows
        self.device = torch.cuda.current_device() if device is None else device
        self.lib = _lib.load()
        self.products = 0

    # ---- one row-sharded product ------------------------------------------------------------------------------------------
    def _rows(self, A1, B1, C_full, ta=False, tb=False, A2=None, B2=None):
        st = self.torch.cuda.current_stream().cuda_stream
        check(self.lib.otn_gemm_rows(A1.data_ptr(), B1.data_ptr(), None if A2 is None else A2.data_ptr(),
                                     None if B2 is None else B2.data_ptr(), C_full.data_ptr(), self.D, self.row0, self.rows,
                                     int(ta), int(tb), self.device, C.c_void_p(st)))
        self.products += 1 if A2 is None else 2

    def _gather(self, C_full):
        if self.world > 1:
            mine = C_full[self.row0:self.row0 + self.rows]
            self.dist.all_gather_into_tensor(C_full, mine.clone(), group=self.group)
        return C_full

    def matmul(self, A, B, ta=False, tb=False, gather=True):
        """op(A) @ op(B): this rank computes its row block; `gather` assembles the full matrix on every rank."""
        out = self.torch.empty((self.D, self.D), dtype=self.torch.float64, device=A.device)
        self._rows(A, B, out, ta, tb)
        return self._gather(out) if gather else out

    def matmul_complex(self, A, B, gather=True):
        """(A_re, A_im) @ (B_re, B_im) for complex operands given as pairs of real matrices."""
        Ar, Ai = A
        Br, Bi = B
        t = self.torch
        Cr = t.empty((self.D, self.D), dtype=t.float64, device=Ar.device)
        Ci = t.empty_like(Cr)
        nAi = -Ai
        self._rows(Ar, Br, Cr, A2=nAi, B2=Bi)
        self._rows(Ar, Bi, Ci, A2=Ai, B2=Br)
        if gather:
            self._gather(Cr); self._gather(Ci)
        return Cr, Ci

    # ---- chains -------------------------------------------------------------------------------------------------------------
    def compose(self, Ys):
        """compose_superops_list (transformations.py:841-851): Ys[-1] @ ... @ Ys[0], every product row-sharded."""
        P = Ys[0]
        for Y in Ys[1:]:
            P = self.matmul(Y, P)
        return P

    def compose_complex(self, Ys):
        P = Ys[0]
        for Y in Ys[1:]:
            P = self.matmul_complex(Y, P)
        return P

    def cost(self, Ys, T_re, T_im=None):
        """|| Ys[-1] ... Ys[0] - T ||_F with the last product left sharded: one scalar all-reduce instead of an all-gather."""
        t = self.torch
        P = Ys[0]
        for Y in Ys[1:-1]:
            P = self.matmul(Y, P)
        M = self.matmul(Ys[-1], P, gather=False) if len(Ys) > 1 else P
        sl = slice(self.row0, self.row0 + self.rows)
        s = t.sum((M[sl] - T_re[sl]) ** 2)
        if T_im is not None:
            s = s + t.sum(T_im[sl] ** 2)
        if self.world > 1:
            self.dist.all_reduce(s, group=self.group)
        return float(t.sqrt(s))

    def cost_and_layer_cotangents(self, Ys, T_re):
        """f and W_l = d f / d Y_l for every layer (the GEMM part of jax.grad at stiefel.py:423), row-sharded:
        G_m = R / f;  W_l = G_l P_{l-1}^T;  G_{l-1} = Y_l^T G_l."""
        t = self.torch
        m = len(Ys)
        P = [Ys[0]]
        for Y in Ys[1:]:
            P.append(self.matmul(Y, P[-1]))
        R = P[-1] - T_re
        f = float(t.linalg.norm(R))
        G = R / f
        W = [None] * m
        for l in range(m - 1, 0, -1):
            W[l] = self.matmul(G, P[l - 1], tb=True)
            G = self.matmul(Ys[l], G, ta=True)
        W[0] = G
        return f, W
