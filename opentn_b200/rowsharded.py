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


class _IpcMatrix:
    """A D x D float64 device buffer allocated by libotn (IPC-shareable), viewable as a torch tensor."""

    def __init__(self, lib, D, device):
        self.lib, self.D, self.device = lib, D, device
        self.ptr = C.c_void_p()
        self.handle = C.create_string_buffer(64)
        check(lib.otn_ipc_alloc(D * D * 8, device, C.byref(self.ptr), self.handle))
        self.__cuda_array_interface__ = {"shape": (D, D), "typestr": "<f8", "data": (self.ptr.value, False), "version": 3, "strides": None}

    def tensor(self, torch):
        return torch.as_tensor(self, device=torch.device("cuda", self.device))


class RowShardedChain:
    def __init__(self, D: int, device=None, group=None, p2p: bool = False, nbuf: int = 4):
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
        self.p2p = bool(p2p) and self.world > 1
        if self.p2p:
            self._setup_p2p(nbuf)

    # ---- fused product + all-gather over NVLink peer memory -----------------------------------------------------------------
    def _setup_p2p(self, nbuf):
        """Ring of `nbuf` IPC-shared result buffers per rank; every rank maps the buffers of all the others."""
        torch, dist = self.torch, self.dist
        self._bufs = [_IpcMatrix(self.lib, self.D, self.device) for _ in range(nbuf)]
        self._views = [b.tensor(torch) for b in self._bufs]
        mine = [bytes(b.handle.raw) for b in self._bufs]
        allh = [None] * self.world
        dist.all_gather_object(allh, mine, group=self.group)
        self._peer_ptrs = []          # per buffer: ctypes array of the other ranks' pointers
        self._opened = []
        for i in range(nbuf):
            arr = (C.c_void_p * 7)()
            k = 0
            for r in range(self.world):
                if r == self.rank:
                    continue
                p = C.c_void_p()
                check(self.lib.otn_ipc_open(allh[r][i], self.device, C.byref(p)))
                self._opened.append(p)
                arr[k] = p.value
                k += 1
            self._peer_ptrs.append(arr)
        self._next = 0
        dist.barrier(group=self.group)

    def _matmul_p2p(self, A, B, ta, tb, A2=None, B2=None):
        torch, dist = self.torch, self.dist
        i = self._next
        self._next = (self._next + 1) % len(self._bufs)
        out = self._views[i]
        st = torch.cuda.current_stream().cuda_stream
        check(self.lib.otn_gemm_rows_bcast(A.data_ptr(), B.data_ptr(), None if A2 is None else A2.data_ptr(),
                                           None if B2 is None else B2.data_ptr(), out.data_ptr(), self._peer_ptrs[i], self.world - 1,
                                           self.D, self.row0, self.rows, int(ta), int(tb), self.device, C.c_void_p(st)))
        self.products += 1 if A2 is None else 2
        torch.cuda.current_stream().synchronize()     # my tiles have landed in every peer ...
        dist.barrier(group=self.group)                # ... and so have everybody else's in mine
        return out

    def close(self):
        if getattr(self, "p2p", False):
            self.dist.barrier(group=self.group)
            for p in self._opened:
                self.lib.otn_ipc_close(p, self.device)
            self._views = []
            for b in self._bufs:
                self.lib.otn_ipc_free(b.ptr, self.device)
            self.p2p = False

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
        """op(A) @ op(B): this rank computes its row block; `gather` assembles the full matrix on every rank.
        With p2p=True the gather is fused into the product (peer stores) and the result lives in a ring buffer that is
        overwritten after `nbuf` further products."""
        if self.p2p and gather:
            return self._matmul_p2p(A, B, ta, tb)
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
