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


class _RawBuffer:
    """CUDA array interface over a raw device pointer."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 3, "strides": None}


class RowShardedChain:
    def __init__(self, D: int, device=None, group=None, p2p: bool = False, nbuf: int = 8, single: bool = False):
        """`single=True`: ignore the process group -- this rank computes every row itself (the one-GPU baseline of a scaling run)."""
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.group = group
        if not single and dist.is_available() and dist.is_initialized():
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
        self._side = None
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
        # device-side ordering: one int per rank in an IPC-shared flag array; rank r raises slot r of every peer's array after each
        # product (otn_peer_signal) and waits for all slots of its own array before the next one (otn_peer_wait)
        self._flag_ptr = C.c_void_p()
        fh = C.create_string_buffer(64)
        check(self.lib.otn_ipc_alloc(256, self.device, C.byref(self._flag_ptr), fh))
        self._flags_view = torch.as_tensor(_RawBuffer(self._flag_ptr.value, (64,), "<i4"), device=torch.device("cuda", self.device))
        self._flags_view.zero_()
        torch.cuda.synchronize()
        allf = [None] * self.world
        dist.all_gather_object(allf, bytes(fh.raw), group=self.group)
        self._peer_flags = (C.c_void_p * 7)()
        k = 0
        for r in range(self.world):
            if r == self.rank:
                continue
            p = C.c_void_p()
            check(self.lib.otn_ipc_open(allf[r], self.device, C.byref(p)))
            self._opened.append(p)
            self._peer_flags[k] = p.value
            k += 1
        self._epoch = 0
        dist.barrier(group=self.group)

    def _launch_p2p(self, A, B, ta, tb, A2=None, B2=None):
        """One product whose tiles are stored into the next ring buffer of every rank (no ordering yet)."""
        torch = self.torch
        out_index = self._next
        self._next = (self._next + 1) % len(self._bufs)
        out = self._views[out_index]
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        check(self.lib.otn_gemm_rows_ex(A.data_ptr(), B.data_ptr(), None if A2 is None else A2.data_ptr(),
                                        None if B2 is None else B2.data_ptr(), out.data_ptr(), self._peer_ptrs[out_index], self.world - 1,
                                        self.D, self.row0, self.rows, int(ta), int(tb), None, None, self.device, st))
        self.products += 1 if A2 is None else 2
        return out

    def _sync_p2p(self):
        """Device-side hand-shake after one or more `_launch_p2p`: raise my flag in every peer, wait for theirs.  Asynchronous."""
        st = C.c_void_p(self.torch.cuda.current_stream().cuda_stream)
        self._epoch += 1
        check(self.lib.otn_peer_signal(self._peer_flags, self.world - 1, self.rank, self._epoch, self.device, st))
        check(self.lib.otn_peer_wait(self._flag_ptr, self.world, self.rank, self._epoch, self.device, st))

    def _matmul_p2p(self, A, B, ta, tb, A2=None, B2=None):
        """Product + all-gather in one kernel (peer stores), ordered against the other ranks on the device: nothing here blocks
        the host.  The result lives in a ring buffer until `nbuf` further products have been issued."""
        out = self._launch_p2p(A, B, ta, tb, A2, B2)
        self._sync_p2p()
        return out

    def close(self):
        if getattr(self, "p2p", False):
            self.torch.cuda.synchronize()
            self.dist.barrier(group=self.group)
            for p in self._opened:
                self.lib.otn_ipc_close(p, self.device)
            self._views = []
            for b in self._bufs:
                self.lib.otn_ipc_free(b.ptr, self.device)
            self._flags_view = None
            self.lib.otn_ipc_free(self._flag_ptr, self.device)
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
        if self.p2p and gather:
            # The two planes are independent launches: on two streams their tiles fill the SMs together.  One plane alone is
            # rows/64 * D/64 tiles -- 512 at D = 4096 on 8 GPUs, 3.46 per SM, so a quarter of the SMs would run a 4th tile while
            # the others idle (measured: 0.71 -> see profiles/README.md); both planes together are 6.9 per SM.
            nAi = -Ai
            cur = t.cuda.current_stream()
            if self._side is None:
                self._side = t.cuda.Stream()
            self._side.wait_stream(cur)
            Cr = self._launch_p2p(Ar, Br, False, False, nAi, Bi)
            with t.cuda.stream(self._side):
                Ci = self._launch_p2p(Ar, Bi, False, False, Ai, Br)
            cur.wait_stream(self._side)
            self._sync_p2p()                      # one hand-shake for both planes
            return Cr, Ci
        Cr = t.empty((self.D, self.D), dtype=t.float64, device=Ar.device)
        Ci = t.empty_like(Cr)
        nAi = -Ai
        cur = t.cuda.current_stream()
        if self._side is None:
            self._side = t.cuda.Stream()
        self._side.wait_stream(cur)
        self._rows(Ar, Br, Cr, A2=nAi, B2=Bi)
        with t.cuda.stream(self._side):
            self._rows(Ar, Bi, Ci, A2=Ai, B2=Br)
        cur.wait_stream(self._side)
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

    def _resid_rows(self, A, B, T_re):
        """Rows of op(A) op(B) - T of this rank (left sharded, no exchange) and the sum of their squares: the residual / norm
        epilogue fused into the product (EPI_RESID), no extra pass over the matrix."""
        t = self.torch
        out = t.empty((self.D, self.D), dtype=t.float64, device=A.device)
        tiles = self.lib.otn_gemm_rows_tiles(self.D, self.rows)
        partial = t.empty(tiles, dtype=t.float64, device=A.device)
        st = C.c_void_p(t.cuda.current_stream().cuda_stream)
        check(self.lib.otn_gemm_rows_ex(A.data_ptr(), B.data_ptr(), None, None, out.data_ptr(), None, 0, self.D, self.row0, self.rows,
                                        0, 0, T_re.data_ptr(), partial.data_ptr(), self.device, st))
        self.products += 1
        return out, partial.sum()

    def cost(self, Ys, T_re, T_im=None):
        """|| Ys[-1] ... Ys[0] - T ||_F with the last product left sharded: its residual and squared norm come out of the GEMM
        epilogue, and one scalar all-reduce replaces the all-gather."""
        t = self.torch
        P = Ys[0]
        for Y in Ys[1:-1]:
            P = self.matmul(Y, P)
        sl = slice(self.row0, self.row0 + self.rows)
        if len(Ys) > 1:
            _, s = self._resid_rows(Ys[-1], P, T_re)
        else:
            s = t.sum((P[sl] - T_re[sl]) ** 2)
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
        if self.p2p and len(self._bufs) < 3 * (m - 1):
            # every forward product and every cotangent is kept until the end: 3 (m - 1) ring buffers, or the ring wraps and a later
            # product silently overwrites an earlier result that is still needed
            raise ValueError(f"cost_and_layer_cotangents with {m} layers keeps {3 * (m - 1)} products alive: construct "
                             f"RowShardedChain(..., nbuf={3 * (m - 1)}) or larger (have {len(self._bufs)})")
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
