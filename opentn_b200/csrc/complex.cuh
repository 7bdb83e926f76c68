// Complex isometries (SURVEY 8f "next" #4; the reference's "TODO: add back the conj", opentn/stiefel.py:29,36,44,184,201,214,218,320).
//
// x in C^{n x p}, x^H x = 1, crosses the C ABI as interleaved complex128 (numpy layout).  Layer superoperators
//     Y[(a,b),(c,d)] = sum_k x[(a,k),c] conj(x[(b,k),d])            (kraus2superop, transformations.py:170-180, with E_k = x[k::K,:])
// are complex D x D.  The CHAIN (compose_superops_list and what autodiff derives from it) is not re-implemented: a complex matrix
// A = Ar + i Ai is carried as the real 2D x 2D matrix  emb(A) = [[Ar, -Ai], [Ai, Ar]],  a *-homomorphism:
//     emb(AB) = emb(A) emb(B),   emb(A^H) = emb(A)^T,   ||emb(A)||_F^2 = 2 ||A||_F^2,   <emb(A), emb(B)> = 2 Re <A, B>,
// so the real FP64-DMMA chain of gemm_dmma.cuh (NN / TN / NT products, dual accumulation, residual and dot epilogues) computes the
// complex forward, reverse and tangent sweeps unchanged on 2D x 2D operands.  (Cost: 8 real D^3 products per complex product
// instead of 4 for the 4M scheme or 3 for 3M -- the price of reusing one verified chain; the kernels below are what is new.)
//
//   cassemble_emb_kernel   x [, eta] -> emb(Y) [, emb(Ydot)]
//   cback_assemble_kernel  emb(W) -> Z,  Z[(a,k),c] = sum_{b,d} (W[(a,b),(c,d)] + conj(W[(b,a),(d,c)])) u[(b,k),d]
//                          (adjoint of the assembly in the sense df = Re tr(Z^H dx); u = x, or eta for the tangent of the adjoint)
//   criem_kernel           egrad -> rgrad, connection, projection with ^T -> ^H (stiefel.py:398-409, :468-475, :23-30)
//   cretract_kernel        polar factor (X + eta) ((X + eta)^H (X + eta))^{-1/2} by coupled Newton-Schulz on the Hermitian Gram
// Closed forms: oracle/complex_port.py (pinned against torch autodiff in tests/test_oracle_complex.py).
#pragma once
#include "otn_common.cuh"
#include "stiefel.cuh"

namespace otn {

struct cplx { double re, im; };
__device__ __forceinline__ cplx cmake(double r, double i) { cplx c; c.re = r; c.im = i; return c; }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return cmake(a.re + b.re, a.im + b.im); }
__device__ __forceinline__ cplx cscale(double s, cplx a) { return cmake(s * a.re, s * a.im); }
__device__ __forceinline__ cplx cconj(cplx a) { return cmake(a.re, -a.im); }
// acc += a * b
__device__ __forceinline__ void cfma(cplx& acc, cplx a, cplx b) {
    acc.re = fma(a.re, b.re, acc.re); acc.re = fma(-a.im, b.im, acc.re);
    acc.im = fma(a.re, b.im, acc.im); acc.im = fma(a.im, b.re, acc.im);
}
// acc += a * conj(b)
__device__ __forceinline__ void cfma_conj(cplx& acc, cplx a, cplx b) {
    acc.re = fma(a.re, b.re, acc.re); acc.re = fma(a.im, b.im, acc.re);
    acc.im = fma(a.im, b.re, acc.im); acc.im = fma(-a.re, b.im, acc.im);
}
// acc += conj(a) * b
__device__ __forceinline__ void cfma_conja(cplx& acc, cplx a, cplx b) {
    acc.re = fma(a.re, b.re, acc.re); acc.re = fma(a.im, b.im, acc.re);
    acc.im = fma(a.re, b.im, acc.im); acc.im = fma(-a.im, b.re, acc.im);
}

// ---- assembly into the real embedding ----------------------------------------------------------------------------------
// grid = (dim [a], items), block 256.  emb layout: [2D][2D] row-major, D = dim^2.
template <bool TANGENT, bool PRIMAL>
__global__ void __launch_bounds__(256) cassemble_emb_kernel(const cplx* __restrict__ x, const cplx* __restrict__ eta, double* __restrict__ Y,
                                                            double* __restrict__ Yd, long long out_stride, int dim, int K,
                                                            const int* active, int m) {
    extern __shared__ __align__(16) double smraw[];
    cplx* xs = reinterpret_cast<cplx*>(smraw);
    const int item = blockIdx.y, a = blockIdx.x;
    if (active != nullptr && active[item / m] == 0) return;
    const int np = dim * K * dim, D = dim * dim, LD = 2 * D;
    cplx* es = xs + np;
    for (int i = threadIdx.x; i < np; i += 256) {
        xs[i] = x[(long long)item * np + i];
        if (TANGENT) es[i] = eta[(long long)item * np + i];
    }
    __syncthreads();
    double* Yo = Y + (long long)item * out_stride;
    double* Ydo = Yd + (long long)item * out_stride;
    for (int o = threadIdx.x; o < dim * D; o += 256) {          // o = b * D + (c * dim + d)
        const int b = o / D, cd = o - b * D, c = cd / dim, d = cd - c * dim;
        cplx s = cmake(0.0, 0.0), sd = cmake(0.0, 0.0);
        for (int k = 0; k < K; ++k) {
            const cplx xa = xs[(a * K + k) * dim + c], xb = xs[(b * K + k) * dim + d];
            if (PRIMAL) cfma_conj(s, xa, xb);
            if (TANGENT) { cfma_conj(sd, es[(a * K + k) * dim + c], xb); cfma_conj(sd, xa, es[(b * K + k) * dim + d]); }
        }
        const long long r = a * dim + b;
        if (PRIMAL) {
            Yo[r * LD + cd] = s.re; Yo[r * LD + D + cd] = -s.im;
            Yo[(r + D) * LD + cd] = s.im; Yo[(r + D) * LD + D + cd] = s.re;
        }
        if (TANGENT) {
            Ydo[r * LD + cd] = sd.re; Ydo[r * LD + D + cd] = -sd.im;
            Ydo[(r + D) * LD + cd] = sd.im; Ydo[(r + D) * LD + D + cd] = sd.re;
        }
    }
}

// ---- adjoint of the assembly ---------------------------------------------------------------------------------------------------
// grid = (dim [a], batch), block 256; out[(a,k),c] (+)= sum_{b,d} (W[(a,b),(c,d)] + conj(W[(b,a),(d,c)])) u[(b,k),d]
// W = top-left + i bottom-left block of emb(W) (instance stride w_stride doubles).  smem: u item | rows (a,b) and (b,a) of W.
__global__ void __launch_bounds__(256) cback_assemble_kernel(const double* __restrict__ Wemb, long long w_stride, const cplx* __restrict__ u,
                                                             cplx* __restrict__ out, int accumulate, int dim, int K, const int* active,
                                                             int m, int layer) {
    extern __shared__ __align__(16) double smraw[];
    const int inst = blockIdx.y, a = blockIdx.x;
    if (active != nullptr && active[inst] == 0) return;
    const int np = dim * K * dim, D = dim * dim, LD = 2 * D;
    cplx* us = reinterpret_cast<cplx*>(smraw);
    cplx* w1 = us + np;            // row (a,b): [D]
    cplx* w2 = w1 + D;             // row (b,a): [D]
    const long long item = (long long)inst * m + layer;
    for (int i = threadIdx.x; i < np; i += 256) us[i] = u[item * np + i];
    const double* W = Wemb + (long long)inst * w_stride;
    // each thread owns outputs (k, c) = o / dim, o % dim for o = tid, tid + 256, ...
    const int nout = K * dim;
    cplx acc[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q] = cmake(0.0, 0.0);
    for (int b = 0; b < dim; ++b) {
        __syncthreads();
        const long long r1 = a * dim + b, r2 = b * dim + a;
        for (int i = threadIdx.x; i < D; i += 256) {
            w1[i] = cmake(W[r1 * LD + i], W[(r1 + D) * LD + i]);
            w2[i] = cmake(W[r2 * LD + i], W[(r2 + D) * LD + i]);
        }
        __syncthreads();
        int q = 0;
        for (int o = threadIdx.x; o < nout; o += 256, ++q) {
            const int k = o / dim, c = o - k * dim;
            cplx s = acc[q];
            for (int d = 0; d < dim; ++d) {
                const cplx ws = cadd(w1[c * dim + d], cconj(w2[d * dim + c]));
                cfma(s, ws, us[(b * K + k) * dim + d]);
            }
            acc[q] = s;
        }
    }
    int q = 0;
    for (int o = threadIdx.x; o < nout; o += 256, ++q) {
        const int k = o / dim, c = o - k * dim;
        const long long og = item * np + (a * K + k) * dim + c;
        out[og] = accumulate ? cadd(out[og], acc[q]) : acc[q];
    }
}

// ---- complex Riemannian epilogue ----------------------------------------------------------------------------------------------
struct CRiemParams {
    const cplx* x; const cplx* zraw; const cplx* eta; const cplx* zdraw;
    const double* part_f; const double* part_s; int tiles;
    double* f_out; cplx* rgrad; cplx* egrad; cplx* heta;
    int n, p, m;
    double alpha0, alpha1;
    const int* active;
};

// G = A^H B  (p x p)
__device__ __forceinline__ void cgram(const cplx* A, const cplx* B, cplx* G, int n, int p) {
    for (int o = threadIdx.x; o < p * p; o += ST_THREADS) {
        const int i = o / p, j = o - i * p;
        cplx s = cmake(0.0, 0.0);
        for (int r = 0; r < n; ++r) cfma_conja(s, A[r * p + i], B[r * p + j]);
        G[o] = s;
    }
}
// Out = alpha Out + A op(G),  op = conjugate transpose if GH
template <bool GH>
__device__ __forceinline__ void caddmul(cplx* Out, double alpha, const cplx* A, const cplx* G, int n, int p) {
    for (int o = threadIdx.x; o < n * p; o += ST_THREADS) {
        const int r = o / p, j = o - r * p;
        cplx s = cscale(alpha, Out[o]);
        for (int i = 0; i < p; ++i) {
            if (GH) cfma_conj(s, A[r * p + i], G[j * p + i]); else cfma(s, A[r * p + i], G[i * p + j]);
        }
        Out[o] = s;
    }
}

// smem: X | E | Z -> nu | Zd -> Dnu -> z | 8 p x p complex work matrices.  Partial sums come from the embedded chain:
// sum part_f = ||R||^2, sum part_s = Re <R, Mdot> (the chain products compute and reduce the upper half of each embedding only).
template <bool HVP>
__global__ void __launch_bounds__(ST_THREADS) criem_kernel(const CRiemParams P) {
    extern __shared__ __align__(16) double smraw[];
    const int item = blockIdx.x, inst = item / P.m, layer = item - inst * P.m;
    if (P.active != nullptr && P.active[inst] == 0) return;
    const int n = P.n, p = P.p, np = n * p, pp = p * p;
    cplx* X = reinterpret_cast<cplx*>(smraw);
    cplx* E = X + np;
    cplx* Z = E + (HVP ? np : 0);
    cplx* Zd = Z + np;
    cplx* G = Zd + (HVP ? np : 0);
    __shared__ double sc[2];
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int t = 0; t < P.tiles; ++t) s += P.part_f[(long long)inst * P.tiles + t];
        const double f = sqrt(s);
        sc[0] = f;
        if (HVP) {
            double d = 0.0;
            for (int t = 0; t < P.tiles; ++t) d += P.part_s[(long long)inst * P.tiles + t];
            sc[1] = d;            // Re <R, Mdot> = f * fdot
        }
        if (layer == 0 && P.f_out != nullptr) P.f_out[inst] = f;
    }
    __syncthreads();
    const double f = sc[0], inv_f = 1.0 / f;
    const long long base = (long long)item * np;
    for (int i = threadIdx.x; i < np; i += ST_THREADS) {
        X[i] = P.x[base + i];
        const cplx zr = P.zraw[base + i];
        Z[i] = cscale(inv_f, zr);
        if (HVP) {
            E[i] = P.eta[base + i];
            const double kap3 = sc[1] * inv_f * inv_f * inv_f;
            const cplx zd = P.zdraw[base + i];
            Zd[i] = cmake(zd.re * inv_f - zr.re * kap3, zd.im * inv_f - zr.im * kap3);
        }
    }
    __syncthreads();
    if (!HVP && P.egrad != nullptr)
        for (int i = threadIdx.x; i < np; i += ST_THREADS) P.egrad[base + i] = Z[i];
    const double a0 = P.alpha0, a1 = P.alpha1;
    const double c1 = 0.5 * ((1.0 / a1) - (2.0 / a0)), c2 = 0.5 * (1.0 / a1), kap = (a0 - a1) / a0;
    cplx *XZ = G, *XZd = G + pp, *XE = G + 2 * pp, *EZ = G + 3 * pp, *Gn = G + 4 * pp, *Gd = G + 5 * pp, *T0 = G + 6 * pp, *T1 = G + 7 * pp;
    cgram(X, Z, XZ, n, p);
    if (HVP) { cgram(X, Zd, XZd, n, p); cgram(X, E, XE, n, p); cgram(E, Z, EZ, n, p); }
    __syncthreads();
    for (int o = threadIdx.x; o < pp; o += ST_THREADS) {
        const int i = o / p, j = o - i * p, ot = j * p + i;
        // Gn = c1 X^H Z - c2 Z^H X = c1 XZ - c2 XZ^H
        Gn[o] = cmake(c1 * XZ[o].re - c2 * XZ[ot].re, c1 * XZ[o].im + c2 * XZ[ot].im);
        if (HVP) Gd[o] = cmake(c1 * (EZ[o].re + XZd[o].re) - c2 * (XZd[ot].re + EZ[ot].re),
                               c1 * (EZ[o].im + XZd[o].im) + c2 * (XZd[ot].im + EZ[ot].im));
    }
    __syncthreads();
    if (HVP) {                                  // Dnu = Zd/a0 + eta Gn + X Gd
        caddmul<false>(Zd, 1.0 / a0, E, Gn, n, p);
        __syncthreads();
        caddmul<false>(Zd, 1.0, X, Gd, n, p);
    }
    caddmul<false>(Z, 1.0 / a0, X, Gn, n, p);    // nu = Z/a0 + X Gn
    __syncthreads();
    if (P.rgrad != nullptr)
        for (int i = threadIdx.x; i < np; i += ST_THREADS) P.rgrad[base + i] = Z[i];
    if (!HVP) return;
    cplx* Nu = Z;
    cplx* EN = XZ;      // eta^H nu
    cplx* XN = XZd;     // X^H nu
    cgram(E, Nu, EN, n, p);
    cgram(X, Nu, XN, n, p);
    __syncthreads();
    // Gz = (EN + EN^H)/2 - kap (XE XN^H + XN XE^H)
    for (int o = threadIdx.x; o < pp; o += ST_THREADS) {
        const int i = o / p, j = o - i * p, ot = j * p + i;
        cplx s = cmake(0.0, 0.0);
        for (int k = 0; k < p; ++k) { cfma_conj(s, XE[i * p + k], XN[j * p + k]); cfma_conj(s, XN[i * p + k], XE[j * p + k]); }
        T0[o] = cmake(0.5 * (EN[o].re + EN[ot].re) - kap * s.re, 0.5 * (EN[o].im - EN[ot].im) - kap * s.im);
    }
    __syncthreads();
    caddmul<false>(Zd, 1.0, X, T0, n, p);        // z = Dnu + X Gz + kap (eta XN^H + nu XE^H)
    __syncthreads();
    if (kap != 0.0) {
        for (int o = threadIdx.x; o < pp; o += ST_THREADS) { T0[o] = cscale(kap, XN[o]); T1[o] = cscale(kap, XE[o]); }
        __syncthreads();
        caddmul<true>(Zd, 1.0, E, T0, n, p);
        __syncthreads();
        caddmul<true>(Zd, 1.0, Nu, T1, n, p);
        __syncthreads();
    }
    cgram(X, Zd, T0, n, p);                      // H eta = z - X herm(X^H z)
    __syncthreads();
    for (int o = threadIdx.x; o < pp; o += ST_THREADS) {
        const int i = o / p, j = o - i * p, ot = j * p + i;
        T1[o] = cmake(-0.5 * (T0[o].re + T0[ot].re), -0.5 * (T0[o].im - T0[ot].im));
    }
    __syncthreads();
    caddmul<false>(Zd, 1.0, X, T1, n, p);
    __syncthreads();
    for (int i = threadIdx.x; i < np; i += ST_THREADS) P.heta[base + i] = Zd[i];
}

// ---- complex polar retraction ---------------------------------------------------------------------------------------------------
// out = A (A^H A)^{-1/2}, A = X + eta; inverse square root of the Hermitian positive-definite Gram by the coupled Newton-Schulz
// iteration (T = (3I - Z Y)/2, Y <- Y T, Z <- T Z) on the Gram scaled into (0, 1]; status bit OTN_STATUS_NOT_SPD (2) when it does
// not converge (singular or non-finite Gram).
__global__ void __launch_bounds__(ST_THREADS) cretract_kernel(const cplx* __restrict__ x, const cplx* __restrict__ eta, cplx* __restrict__ out,
                                                              int n, int p, int* status) {
    extern __shared__ __align__(16) double smraw[];
    const int item = blockIdx.x, np = n * p, pp = p * p;
    cplx* A = reinterpret_cast<cplx*>(smraw);
    cplx* Y = A + np;
    cplx* Zm = Y + pp;
    cplx* T = Zm + pp;
    __shared__ double red[ST_WARPS];
    __shared__ double errs;
    const long long base = (long long)item * np;
    for (int i = threadIdx.x; i < np; i += ST_THREADS) {
        const cplx a = x[base + i], e = eta ? eta[base + i] : cmake(0.0, 0.0);
        A[i] = cadd(a, e);
    }
    __syncthreads();
    cgram(A, A, Y, n, p);
    for (int o = threadIdx.x; o < pp; o += ST_THREADS) Zm[o] = cmake((o / p == o % p) ? 1.0 : 0.0, 0.0);
    __syncthreads();
    // scale the Gram so that its spectrum lies in (0, 1]: c = ||S||_F >= lambda_max, (S / c)^{-1/2} = sqrt(c) S^{-1/2}.  Inside that
    // interval the iteration converges to the PRINCIPAL root; unscaled, a Gram far from the identity can converge to a root with
    // flipped eigenvalue signs (a wrong unitary factor) without any sign of failure.
    {
        double s2 = 0.0;
        for (int o = threadIdx.x; o < pp; o += ST_THREADS) { s2 = fma(Y[o].re, Y[o].re, s2); s2 = fma(Y[o].im, Y[o].im, s2); }
        s2 = warp_sum(s2);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s2;
        __syncthreads();
        if (threadIdx.x == 0) { double t = 0.0; for (int w = 0; w < ST_WARPS; ++w) t += red[w]; errs = sqrt(t); }
        __syncthreads();
    }
    const double cnorm = errs;
    __syncthreads();
    for (int o = threadIdx.x; o < pp; o += ST_THREADS) Y[o] = cscale(1.0 / cnorm, Y[o]);
    __syncthreads();
    bool ok = false;
    for (int it = 0; it < 80; ++it) {
        double err = 0.0;
        for (int o = threadIdx.x; o < pp; o += ST_THREADS) {
            const int i = o / p, j = o - i * p;
            cplx zy = cmake(0.0, 0.0);
            for (int k = 0; k < p; ++k) cfma(zy, Zm[i * p + k], Y[k * p + j]);
            const double id = (i == j) ? 1.0 : 0.0;
            T[o] = cmake(0.5 * (3.0 * id - zy.re), -0.5 * zy.im);
            err = fma(id - zy.re, id - zy.re, err); err = fma(zy.im, zy.im, err);
        }
        err = warp_sum(err);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = err;
        __syncthreads();
        if (threadIdx.x == 0) { double t = 0.0; for (int w = 0; w < ST_WARPS; ++w) t += red[w]; errs = t; }
        __syncthreads();
        const double tot = errs;
        if (tot < 1e-30) { ok = true; break; }
        if (!(tot < 1e6)) break;                                   // diverging (or NaN)
        cplx yn = cmake(0.0, 0.0), zn = cmake(0.0, 0.0);           // pp <= 256 <= ST_THREADS: one entry per thread
        const int o = threadIdx.x;
        if (o < pp) {
            const int i = o / p, j = o - i * p;
            for (int k = 0; k < p; ++k) { cfma(yn, Y[i * p + k], T[k * p + j]); cfma(zn, T[i * p + k], Zm[k * p + j]); }
        }
        __syncthreads();
        if (o < pp) { Y[o] = yn; Zm[o] = zn; }
        __syncthreads();
    }
    if (!ok && threadIdx.x == 0 && status != nullptr) atomicOr(&status[item], 2);
    __syncthreads();
    const double back = 1.0 / sqrt(cnorm);
    for (int o = threadIdx.x; o < np; o += ST_THREADS) {
        const int r = o / p, j = o - r * p;
        cplx s = cmake(0.0, 0.0);
        for (int i = 0; i < p; ++i) cfma(s, A[r * p + i], Zm[i * p + j]);
        out[base + o] = cscale(back, s);
    }
}

}  // namespace otn
