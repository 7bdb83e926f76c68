// Small batched Stiefel-manifold kernels (sites F, G-epilogue, H, J, K of the hot path).  One CTA per isometry
// ("item" = instance * m + layer); everything for one (n,p) isometry lives in shared memory; HBM-bound:
// algorithmic traffic = (#inputs + #outputs) * n*p*8 bytes per item.
//
//   riem_kernel<HVP>   egrad -> rgrad                      gradient_ambient2riemannian, opentn/stiefel.py:398-409
//                      (+ HVP) d(rgrad)[eta] -> connection riemannian_connection, stiefel.py:468-475
//                      -> tangent projection               what parametrization_from_tangent (:266-300) keeps of z
//   project_kernel     Z - X sym(X^T Z)                    project, stiefel.py:23-30
//   retract_kernel     polar factor of X + eta             polar_decomposition_rectangular, stiefel.py:323-330
//                      as (X+eta) (A^T A)^{-1/2}, A = X+eta, inverse square root by a parallel cyclic Jacobi eigensolver
//   cdot_kernel        sum_l tr(a_l^T (I - x_l x_l^T/2) b_l)  canonical_metric, stiefel.py:76-83  (== the Euclidean dot of the
//                      (A-upper, B) coefficient vectors used by truncated_cg, trust_region_rcopt.py:111-131)
#pragma once
#include "otn_common.cuh"

namespace otn {

constexpr int ST_THREADS = 512;   // 16 warps per isometry: these kernels keep one CTA per SM (shared memory), so latency hiding must come from within
constexpr int ST_WARPS = ST_THREADS / 32;
constexpr int ST_MAXP = 16;  // p <= 16 (dim = d^N <= 16 or d^2 = 4)

// G[i][j] = sum_r A[r][i] * B[r][j]   (p x p), fixed summation order => deterministic
__device__ __forceinline__ void gram(const double* A, const double* B, double* G, int n, int p) {
    for (int o = threadIdx.x; o < p * p; o += ST_THREADS) {
        const int i = o / p, j = o - i * p;
        double s = 0.0;
        for (int r = 0; r < n; ++r) s = fma(A[r * p + i], B[r * p + j], s);
        G[o] = s;
    }
}

// Out[r][j] = alpha * Out[r][j] + sum_i A[r][i] * G[i][j]   (GT: use G[j][i])
template <bool GT>
__device__ __forceinline__ void addmul(double* Out, double alpha, const double* A, const double* G, int n, int p) {
    for (int o = threadIdx.x; o < n * p; o += ST_THREADS) {
        const int r = o / p, j = o - r * p;
        double s = alpha * Out[o];
        for (int i = 0; i < p; ++i) s = fma(A[r * p + i], GT ? G[j * p + i] : G[i * p + j], s);
        Out[o] = s;
    }
}

struct RiemParams {
    const double* x;       // [items][n*p]
    const double* zraw;    // [items][n*p]  unscaled Euclidean gradient:  egrad = zraw / f
    const double* eta;     // HVP: tangent direction
    const double* zdraw;   // HVP: unscaled derivative part, see below
    const double* part_f;  // [batch][tiles] partial sums of ||Re R||^2
    const double* part_s;  // HVP: [batch][tiles] partial sums of <Re R, Mdot>
    const double* im2;     // [n_targets] ||Im T||^2
    const int* tindex;     // [batch] target id (nullptr = 0)
    int tiles;
    double* f_out;         // [batch] (written by layer 0)
    double* rgrad;         // [items][n*p]  (nullable in HVP mode)
    double* egrad;         // optional
    double* heta;          // HVP out
    int n, p, m;
    double alpha0, alpha1;
    const int* active;
};

// smem layout: X | E(eta) | Z->nu | Zd->Dnu->z | grams (8 * p*p)
template <bool HVP>
__global__ void __launch_bounds__(ST_THREADS) riem_kernel(const RiemParams P) {
    extern __shared__ __align__(16) double sm[];
    const int item = blockIdx.x, inst = item / P.m, layer = item - inst * P.m;
    if (P.active != nullptr && P.active[inst] == 0) return;
    const int n = P.n, p = P.p, np = n * p, pp = p * p;
    double* X = sm;
    double* E = X + np;
    double* Z = E + (HVP ? np : 0);
    double* Zd = Z + np;
    double* G = Zd + (HVP ? np : 0);
    __shared__ double sc[2];
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int t = 0; t < P.tiles; ++t) s += P.part_f[(long long)inst * P.tiles + t];
        s += P.im2[P.tindex ? P.tindex[inst] : 0];
        const double f = sqrt(s);
        sc[0] = f;
        if (HVP) {
            double d = 0.0;
            for (int t = 0; t < P.tiles; ++t) d += P.part_s[(long long)inst * P.tiles + t];
            sc[1] = d;  // <Re R, Mdot> = f * fdot
        }
        if (layer == 0 && P.f_out != nullptr) P.f_out[inst] = f;
    }
    __syncthreads();
    const double f = sc[0], inv_f = 1.0 / f;
    const long long base = (long long)item * np;
    for (int i = threadIdx.x; i < np; i += ST_THREADS) {
        X[i] = P.x[base + i];
        const double zr = P.zraw[base + i];
        Z[i] = zr * inv_f;
        if (HVP) {
            E[i] = P.eta[base + i];
            // Zdot = (1/f) zdraw - (fdot / f^2) zraw,   fdot = sc[1] / f
            Zd[i] = P.zdraw[base + i] * inv_f - zr * (sc[1] * inv_f * inv_f * inv_f);
        }
    }
    __syncthreads();
    if (!HVP && P.egrad != nullptr)
        for (int i = threadIdx.x; i < np; i += ST_THREADS) P.egrad[base + i] = Z[i];
    const double a0 = P.alpha0, a1 = P.alpha1;
    const double c1 = 0.5 * ((1.0 / a1) - (2.0 / a0)), c2 = 0.5 * (1.0 / a1), kap = (a0 - a1) / a0;
    double* XZ = G;            // X^T Z
    double* XZd = G + pp;      // X^T Zd
    double* XE = G + 2 * pp;   // X^T eta
    double* EZ = G + 3 * pp;   // eta^T Z
    double* Gn = G + 4 * pp;   // c1 XZ - c2 XZ^T
    double* Gd = G + 5 * pp;
    double* T0 = G + 6 * pp;
    double* T1 = G + 7 * pp;
    gram(X, Z, XZ, n, p);
    if (HVP) { gram(X, Zd, XZd, n, p); gram(X, E, XE, n, p); gram(E, Z, EZ, n, p); }
    __syncthreads();
    for (int o = threadIdx.x; o < pp; o += ST_THREADS) {
        const int i = o / p, j = o - i * p, ot = j * p + i;
        Gn[o] = c1 * XZ[o] - c2 * XZ[ot];
        if (HVP) Gd[o] = c1 * (EZ[o] + XZd[o]) - c2 * (XZd[ot] + EZ[ot]);
    }
    __syncthreads();
    // Dnu = Zd/a0 + eta Gn + X Gd   (in place over Zd; must precede the in-place update of Z)
    if (HVP) {
        addmul<false>(Zd, 1.0 / a0, E, Gn, n, p);
        __syncthreads();
        addmul<false>(Zd, 1.0, X, Gd, n, p);
    }
    // nu = Z/a0 + X Gn   (in place over Z)
    addmul<false>(Z, 1.0 / a0, X, Gn, n, p);
    __syncthreads();
    if (P.rgrad != nullptr)
        for (int i = threadIdx.x; i < np; i += ST_THREADS) P.rgrad[base + i] = Z[i];
    if (!HVP) return;
    double* Nu = Z;
    double* EN = XZ;   // eta^T nu   (reuse)
    double* XN = XZd;  // X^T nu
    gram(E, Nu, EN, n, p);
    gram(X, Nu, XN, n, p);
    __syncthreads();
    // Gz = (EN + EN^T)/2 - kap (XE XN^T + XN XE^T)
    for (int o = threadIdx.x; o < pp; o += ST_THREADS) {
        const int i = o / p, j = o - i * p;
        double s = 0.0;
        for (int k = 0; k < p; ++k) s += XE[i * p + k] * XN[j * p + k] + XN[i * p + k] * XE[j * p + k];
        T0[o] = 0.5 * (EN[o] + EN[j * p + i]) - kap * s;
    }
    __syncthreads();
    // z = Dnu + X Gz + kap (eta XN^T + nu XE^T)
    addmul<false>(Zd, 1.0, X, T0, n, p);
    __syncthreads();
    if (kap != 0.0) {
        for (int o = threadIdx.x; o < pp; o += ST_THREADS) { T0[o] = kap * XN[o]; T1[o] = kap * XE[o]; }
        __syncthreads();
        addmul<true>(Zd, 1.0, E, T0, n, p);
        __syncthreads();
        addmul<true>(Zd, 1.0, Nu, T1, n, p);
        __syncthreads();
    }
    // H eta = z - X sym(X^T z)
    gram(X, Zd, T0, n, p);
    __syncthreads();
    for (int o = threadIdx.x; o < pp; o += ST_THREADS) {
        const int i = o / p, j = o - i * p;
        T1[o] = -0.5 * (T0[o] + T0[j * p + i]);
    }
    __syncthreads();
    addmul<false>(Zd, 1.0, X, T1, n, p);
    __syncthreads();
    for (int i = threadIdx.x; i < np; i += ST_THREADS) P.heta[base + i] = Zd[i];
}

// cost only: f[b] = sqrt(sum partial + im2)
__global__ void cost_from_partials_kernel(const double* part_f, int tiles, const double* im2, const int* tindex, double* f,
                                          int batch, const int* active, double scale = 1.0) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    if (active != nullptr && active[b] == 0) return;
    double s = 0.0;
    for (int t = 0; t < tiles; ++t) s += part_f[(long long)b * tiles + t];
    f[b] = sqrt(scale * s + im2[tindex ? tindex[b] : 0]);
}

// out = a0inv * Z + c1 X X^T Z - c2 X Z^T X   (general metric; (1,1) => project for tangent-normal split)
// mode 0: project (Z - X sym(X^T Z));  mode 1: gradient_ambient2riemannian(alpha0, alpha1)
__global__ void __launch_bounds__(ST_THREADS) project_kernel(const double* __restrict__ x, const double* __restrict__ z,
                                                             double* __restrict__ out, int n, int p, int mode, double a0, double a1) {
    extern __shared__ __align__(16) double sm[];
    const int np = n * p, pp = p * p;
    double* X = sm;
    double* Z = X + np;
    double* G = Z + np;
    double* Gn = G + pp;
    const long long base = (long long)blockIdx.x * np;
    for (int i = threadIdx.x; i < np; i += ST_THREADS) { X[i] = x[base + i]; Z[i] = z[base + i]; }
    __syncthreads();
    gram(X, Z, G, n, p);
    __syncthreads();
    double c1, c2, s0;
    if (mode == 0) { c1 = -0.5; c2 = 0.5; s0 = 1.0; }
    else { c1 = 0.5 * ((1.0 / a1) - (2.0 / a0)); c2 = 0.5 * (1.0 / a1); s0 = 1.0 / a0; }
    for (int o = threadIdx.x; o < pp; o += ST_THREADS) {
        const int i = o / p, j = o - i * p;
        Gn[o] = c1 * G[o] - c2 * G[j * p + i];
    }
    __syncthreads();
    addmul<false>(Z, s0, X, Gn, n, p);
    __syncthreads();
    for (int i = threadIdx.x; i < np; i += ST_THREADS) out[base + i] = Z[i];
}

// ---- symmetric eigen-decomposition of a p x p SPD matrix by parallel cyclic Jacobi (one warp) -----------------------
// S (in/out: diag holds eigenvalues), V (out: eigenvectors in columns).  p <= ST_MAXP.  Round-robin pairing gives
// floor(p/2) disjoint rotations per step; column and row phases are separated by __syncwarp.
__device__ inline void jacobi_eig_warp(double* S, double* V, int p, double* cs /* 2*ST_MAXP */, int* pairs /* 2*ST_MAXP */) {
    const int lane = threadIdx.x & 31;
    const int pe = (p + 1) & ~1;  // even player count (a dummy sits out when p is odd)
    const int half = pe / 2;
    for (int i = lane; i < p * p; i += 32) V[i] = (i / p == i % p) ? 1.0 : 0.0;
    __syncwarp();
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0;
        for (int i = lane; i < p * p; i += 32) {
            const int r = i / p, c = i % p;
            if (r != c) off += S[i] * S[i];
        }
        off = warp_sum(off);
        double dg = 0.0;
        for (int i = lane; i < p; i += 32) dg += S[i * p + i] * S[i * p + i];
        dg = warp_sum(dg);
        if (off <= 1e-29 * dg) break;   // off-diagonal norm below ~3e-15 relative: converged to rounding
        for (int step = 0; step < pe - 1; ++step) {
            // round-robin tournament: player pe-1 fixed, others rotate
            if (lane < half) {
                int a = (lane == 0) ? pe - 1 : (step + lane) % (pe - 1);
                int b = (step + pe - 1 - lane) % (pe - 1);
                int i = min(a, b), j = max(a, b);
                double c = 1.0, s = 0.0;
                if (j < p) {
                    const double apq = S[i * p + j];
                    if (apq != 0.0) {
                        const double theta = (S[j * p + j] - S[i * p + i]) / (2.0 * apq);
                        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                        c = 1.0 / sqrt(t * t + 1.0);
                        s = t * c;
                    }
                } else { i = j = 0; }
                pairs[2 * lane] = i; pairs[2 * lane + 1] = j;
                cs[2 * lane] = c; cs[2 * lane + 1] = s;
            }
            __syncwarp();
            // column phase: S <- S J, V <- V J
            for (int w = lane; w < half * p; w += 32) {
                const int q = w / p, r = w - q * p, i = pairs[2 * q], j = pairs[2 * q + 1];
                if (i == j) continue;
                const double c = cs[2 * q], s = cs[2 * q + 1];
                const double si = S[r * p + i], sj = S[r * p + j];
                S[r * p + i] = c * si - s * sj; S[r * p + j] = s * si + c * sj;
                const double vi = V[r * p + i], vj = V[r * p + j];
                V[r * p + i] = c * vi - s * vj; V[r * p + j] = s * vi + c * vj;
            }
            __syncwarp();
            // row phase: S <- J^T S
            for (int w = lane; w < half * p; w += 32) {
                const int q = w / p, r = w - q * p, i = pairs[2 * q], j = pairs[2 * q + 1];
                if (i == j) continue;
                const double c = cs[2 * q], s = cs[2 * q + 1];
                const double si = S[i * p + r], sj = S[j * p + r];
                S[i * p + r] = c * si - s * sj; S[j * p + r] = s * si + c * sj;
            }
            __syncwarp();
        }
    }
}

// Q = G^{-1/2} for a p x p SPD matrix G (dense, row-major), all ST_THREADS threads.
//   * ||G - I||_F < 0.5 (every trust-region step: G = I + eta^T eta with ||eta|| <= radius <= 0.1): coupled Newton-Schulz
//     iteration  T = (3I - Z Y)/2, Y <- Y T, Z <- T Z  (Y -> G^{1/2}, Z -> G^{-1/2}), quadratically convergent, p x p products only;
//   * otherwise: symmetric eigendecomposition by the one-warp Jacobi solver above.
// Work arrays Y, Z, T: p*p doubles each.  Returns (in every thread) 0 = ok, 2 = G not positive definite.
__device__ inline int inv_sqrt_spd(const double* G, double* Q, double* Y, double* Z, double* T, int p, double* cs, int* pairs,
                                   double* red /* >= ST_WARPS + 1 doubles */) {
    const int pp = p * p;
    double dev = 0.0;
    for (int o = threadIdx.x; o < pp; o += ST_THREADS) { const double e = G[o] - ((o / p == o % p) ? 1.0 : 0.0); dev = fma(e, e, dev); }
    dev = warp_sum(dev);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dev;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0.0; for (int w = 0; w < ST_THREADS / 32; ++w) t += red[w]; red[ST_WARPS] = t; }
    __syncthreads();
    const bool newton = red[ST_WARPS] < 0.25;
    __syncthreads();
    if (newton) {
        for (int o = threadIdx.x; o < pp; o += ST_THREADS) { Y[o] = G[o]; Z[o] = (o / p == o % p) ? 1.0 : 0.0; }
        __syncthreads();
        for (int it = 0; it < 30; ++it) {
            double err = 0.0;
            for (int o = threadIdx.x; o < pp; o += ST_THREADS) {          // T = (3I - Z Y)/2 ; err = ||I - Z Y||_F^2
                const int i = o / p, j = o - i * p;
                double zy = 0.0;
                for (int k = 0; k < p; ++k) zy = fma(Z[i * p + k], Y[k * p + j], zy);
                const double id = (i == j) ? 1.0 : 0.0;
                T[o] = 0.5 * (3.0 * id - zy);
                err = fma(id - zy, id - zy, err);
            }
            err = warp_sum(err);
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = err;
            __syncthreads();
            double tot = 0.0;
            for (int w = 0; w < ST_THREADS / 32; ++w) tot += red[w];
            if (tot < 1e-30) break;                                        // uniform: every thread sees the same sum
            double yn[4], zn[4];                                           // pp <= 256 <= ST_THREADS: at most one element per thread
            int cnt = 0;
            for (int o = threadIdx.x; o < pp; o += ST_THREADS, ++cnt) {
                const int i = o / p, j = o - i * p;
                double a = 0.0, b = 0.0;
                for (int k = 0; k < p; ++k) { a = fma(Y[i * p + k], T[k * p + j], a); b = fma(T[i * p + k], Z[k * p + j], b); }
                yn[cnt] = a; zn[cnt] = b;
            }
            __syncthreads();
            cnt = 0;
            for (int o = threadIdx.x; o < pp; o += ST_THREADS, ++cnt) { Y[o] = yn[cnt]; Z[o] = zn[cnt]; }
            __syncthreads();
        }
        __syncthreads();
        for (int o = threadIdx.x; o < pp; o += ST_THREADS) Q[o] = 0.5 * (Z[o] + Z[(o % p) * p + o / p]);   // symmetrise
        __syncthreads();
        return 0;
    }
    for (int o = threadIdx.x; o < pp; o += ST_THREADS) Y[o] = G[o];
    __syncthreads();
    if (threadIdx.x < 32) jacobi_eig_warp(Y, Z, p, cs, pairs);
    __syncthreads();
    for (int o = threadIdx.x; o < pp; o += ST_THREADS) {
        const int i = o / p, j = o - i * p;
        double acc = 0.0;
        for (int k = 0; k < p; ++k) acc += Z[i * p + k] * Z[j * p + k] / sqrt(Y[k * p + k]);
        Q[o] = acc;
    }
    int bad = 0;
    for (int k = 0; k < p; ++k) if (!(Y[k * p + k] > 0.0)) bad = 2;
    __syncthreads();
    return bad;
}

// retract: out = A (A^T A)^{-1/2},  A = X + eta.   status bit 2 if A^T A is not positive definite.
__global__ void __launch_bounds__(ST_THREADS) retract_kernel(const double* __restrict__ x, const double* __restrict__ eta,
                                                             double* __restrict__ out, int n, int p, int m, const int* active,
                                                             int* status) {
    extern __shared__ __align__(16) double sm[];
    const int item = blockIdx.x;
    if (active != nullptr && active[item / m] == 0) return;
    const int np = n * p, pp = p * p;
    double* A = sm;
    double* S = A + np;
    double* V = S + pp;
    double* Q = V + pp;
    double* Yw = Q + pp;
    double* Tw = Yw + pp;
    __shared__ double cs[2 * ST_MAXP];
    __shared__ int pairs[2 * ST_MAXP];
    __shared__ double red[ST_WARPS + 1];
    const long long base = (long long)item * np;
    for (int i = threadIdx.x; i < np; i += ST_THREADS) A[i] = x[base + i] + (eta ? eta[base + i] : 0.0);
    __syncthreads();
    gram(A, A, S, n, p);
    __syncthreads();
    const int bad = inv_sqrt_spd(S, Q, Yw, V, Tw, p, cs, pairs, red);
    if (threadIdx.x == 0 && status != nullptr && bad) atomicOr(&status[item / m], bad);
    __syncthreads();
    for (int o = threadIdx.x; o < np; o += ST_THREADS) {
        const int r = o / p, j = o - r * p;
        double s = 0.0;
        for (int i = 0; i < p; ++i) s = fma(A[r * p + i], Q[i * p + j], s);
        out[base + o] = s;
    }
}

// canonical dot per instance: out[b] = sum_l [ <a_l, b_l> - 1/2 <x_l^T a_l, x_l^T b_l> ];  one CTA per instance
__global__ void __launch_bounds__(ST_THREADS) cdot_kernel(const double* __restrict__ x, const double* __restrict__ a,
                                                          const double* __restrict__ bb, double* __restrict__ out, int n, int p,
                                                          int m, const int* active) {
    extern __shared__ __align__(16) double sm[];
    const int inst = blockIdx.x;
    if (active != nullptr && active[inst] == 0) return;
    const int np = n * p, pp = p * p;
    double* X = sm;
    double* A = X + np;
    double* B = A + np;
    double* GA = B + np;
    double* GB = GA + pp;
    __shared__ double red[ST_THREADS / 32];
    double total = 0.0;  // only meaningful in thread 0
    for (int l = 0; l < m; ++l) {
        const long long base = ((long long)inst * m + l) * np;
        for (int i = threadIdx.x; i < np; i += ST_THREADS) { X[i] = x[base + i]; A[i] = a[base + i]; B[i] = bb[base + i]; }
        __syncthreads();
        gram(X, A, GA, n, p);
        gram(X, B, GB, n, p);
        double s = 0.0;
        for (int i = threadIdx.x; i < np; i += ST_THREADS) s = fma(A[i], B[i], s);
        __syncthreads();
        for (int o = threadIdx.x; o < pp; o += ST_THREADS) s = fma(-0.5 * GA[o], GB[o], s);
        s = warp_sum(s);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0)
            for (int w = 0; w < ST_THREADS / 32; ++w) total += red[w];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[inst] = total;
}

}  // namespace otn

// ==================================================================================================================
// Tensor-core (DMMA.8x8x4) version of the Riemannian epilogue for p in {8, 16} (full-sites model).  The SIMT version above
// is a chain of 256-long dependent FMA loops on 8 warps per SM (latency bound: 3 ms per 3072 isometries); here every
// p x p Gram and every (n x p)(p x p) product is issued as DMMA on padded (LD = p + 4) shared-memory operands.
// ==================================================================================================================
namespace otn {

// G_q = A_q^T B_q for q < NG: (P/8)^2 blocks of 8x8 per Gram, blocks dealt to the 8 warps; fixed summation order.
template <int P, int NG>
__device__ __forceinline__ void gram_blocks_mma(const double* const (&A)[NG], const double* const (&B)[NG], double* const (&G)[NG],
                                                int n) {
    constexpr int LD = P + 4, PB = P / 8, NB = NG * PB * PB;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    for (int bid = warp; bid < NB; bid += ST_THREADS / 32) {
        const int q = bid / (PB * PB), ij = bid - q * PB * PB, i = ij / PB, j = ij - i * PB;
        const double* a = A[q] + i * 8 + g;
        const double* b = B[q] + j * 8 + g;
        double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;   // two interleaved accumulators (ILP)
        int r0 = 0;
        for (; r0 + 8 <= n; r0 += 8) {
            dmma884(c0, c1, a[(r0 + t) * LD], b[(r0 + t) * LD]);
            dmma884(d0, d1, a[(r0 + 4 + t) * LD], b[(r0 + 4 + t) * LD]);
        }
        for (; r0 < n; r0 += 4) dmma884(c0, c1, a[(r0 + t) * LD], b[(r0 + t) * LD]);
        double* out = G[q] + (i * 8 + g) * LD + j * 8 + 2 * t;
        out[0] = c0 + d0; out[1] = c1 + d1;
    }
}

// Out = alpha Out + sum_{q<NT} A_q * op(G_q)   (n x P), op = transpose if GT[q]; rows dealt to warps in blocks of 8.
template <int P, int NT>
__device__ __forceinline__ void addmul_mma(double* Out, double alpha, const double* const (&A)[NT], const double* const (&G)[NT],
                                           const bool (&GT)[NT], int n) {
    constexpr int LD = P + 4, PB = P / 8;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    for (int rb = warp; rb < n / 8; rb += ST_THREADS / 32) {
        double acc[PB][2];
#pragma unroll
        for (int j = 0; j < PB; ++j) {
            const double* o = Out + (rb * 8 + g) * LD + j * 8 + 2 * t;
            acc[j][0] = alpha * o[0]; acc[j][1] = alpha * o[1];
        }
#pragma unroll
        for (int q = 0; q < NT; ++q) {
#pragma unroll
            for (int i0 = 0; i0 < P; i0 += 4) {
                const double a = A[q][(rb * 8 + g) * LD + i0 + t];
#pragma unroll
                for (int j = 0; j < PB; ++j) {
                    const double b = GT[q] ? G[q][(j * 8 + g) * LD + i0 + t] : G[q][(i0 + t) * LD + j * 8 + g];
                    dmma884(acc[j][0], acc[j][1], a, b);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < PB; ++j) {
            double* o = Out + (rb * 8 + g) * LD + j * 8 + 2 * t;
            o[0] = acc[j][0]; o[1] = acc[j][1];
        }
    }
}

template <int P, bool HVP>
__global__ void __launch_bounds__(ST_THREADS) riem_mma_kernel(const RiemParams Pm) {
    extern __shared__ __align__(16) double sm[];
    constexpr int LD = P + 4, GS = P * LD;
    const int item = blockIdx.x, inst = item / Pm.m, layer = item - inst * Pm.m;
    if (Pm.active != nullptr && Pm.active[inst] == 0) return;
    const int n = Pm.n, np = n * P, nl = n * LD;
    double* X = sm;
    double* E = X + nl;
    double* Z = E + (HVP ? nl : 0);
    double* Zd = Z + nl;
    double* G = Zd + (HVP ? nl : 0);     // 8 Gram-sized slots
    __shared__ double sc[2];
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int t = 0; t < Pm.tiles; ++t) s += Pm.part_f[(long long)inst * Pm.tiles + t];
        s += Pm.im2[Pm.tindex ? Pm.tindex[inst] : 0];
        const double f = sqrt(s);
        sc[0] = f;
        if (HVP) {
            double d = 0.0;
            for (int t = 0; t < Pm.tiles; ++t) d += Pm.part_s[(long long)inst * Pm.tiles + t];
            sc[1] = d;
        }
        if (layer == 0 && Pm.f_out != nullptr) Pm.f_out[inst] = f;
    }
    __syncthreads();
    const double f = sc[0], inv_f = 1.0 / f;
    const long long base = (long long)item * np;
    for (int i = threadIdx.x; i < np; i += ST_THREADS) {
        const int r = i / P, c = i - r * P, o = r * LD + c;
        X[o] = Pm.x[base + i];
        const double zr = Pm.zraw[base + i];
        const double z = zr * inv_f;
        Z[o] = z;
        if (!HVP && Pm.egrad != nullptr) Pm.egrad[base + i] = z;
        if (HVP) {
            E[o] = Pm.eta[base + i];
            Zd[o] = Pm.zdraw[base + i] * inv_f - zr * (sc[1] * inv_f * inv_f * inv_f);
        }
    }
    __syncthreads();
    const double a0 = Pm.alpha0, a1 = Pm.alpha1;
    const double c1 = 0.5 * ((1.0 / a1) - (2.0 / a0)), c2 = 0.5 * (1.0 / a1), kap = (a0 - a1) / a0;
    double *XZ = G, *XZd = G + GS, *XE = G + 2 * GS, *EZ = G + 3 * GS, *Gn = G + 4 * GS, *Gd = G + 5 * GS, *T0 = G + 6 * GS, *T1 = G + 7 * GS;
    if (HVP) {
        const double* const As[4] = {X, X, X, E};
        const double* const Bs[4] = {Z, Zd, E, Z};
        double* const Gs[4] = {XZ, XZd, XE, EZ};
        gram_blocks_mma<P, 4>(As, Bs, Gs, n);
    } else {
        const double* const As[1] = {X};
        const double* const Bs[1] = {Z};
        double* const Gs[1] = {XZ};
        gram_blocks_mma<P, 1>(As, Bs, Gs, n);
    }
    __syncthreads();
    for (int o = threadIdx.x; o < P * P; o += ST_THREADS) {
        const int i = o / P, j = o - i * P, ij = i * LD + j, ji = j * LD + i;
        Gn[ij] = c1 * XZ[ij] - c2 * XZ[ji];
        if (HVP) Gd[ij] = c1 * (EZ[ij] + XZd[ij]) - c2 * (XZd[ji] + EZ[ji]);
    }
    __syncthreads();
    if (HVP) {   // Dnu = Zd/a0 + eta Gn + X Gd
        const double* const As[2] = {E, X};
        const double* const Gs[2] = {Gn, Gd};
        const bool GT[2] = {false, false};
        addmul_mma<P, 2>(Zd, 1.0 / a0, As, Gs, GT, n);
    }
    {            // nu = Z/a0 + X Gn
        const double* const As[1] = {X};
        const double* const Gs[1] = {Gn};
        const bool GT[1] = {false};
        addmul_mma<P, 1>(Z, 1.0 / a0, As, Gs, GT, n);
    }
    __syncthreads();
    if (Pm.rgrad != nullptr)
        for (int i = threadIdx.x; i < np; i += ST_THREADS) Pm.rgrad[base + i] = Z[(i / P) * LD + (i % P)];
    if (!HVP) return;
    double *Nu = Z, *EN = XZ, *XN = XZd;
    {
        const double* const As[2] = {E, X};
        const double* const Bs[2] = {Nu, Nu};
        double* const Gs[2] = {EN, XN};
        gram_blocks_mma<P, 2>(As, Bs, Gs, n);
    }
    __syncthreads();
    for (int o = threadIdx.x; o < P * P; o += ST_THREADS) {   // Gz = (EN + EN^T)/2 - kap (XE XN^T + XN XE^T); T0 = kap XN; T1 = kap XE
        const int i = o / P, j = o - i * P;
        double s = 0.0;
        for (int k = 0; k < P; ++k) s += XE[i * LD + k] * XN[j * LD + k] + XN[i * LD + k] * XE[j * LD + k];
        Gn[i * LD + j] = 0.5 * (EN[i * LD + j] + EN[j * LD + i]) - kap * s;
        T0[i * LD + j] = kap * XN[i * LD + j];
        T1[i * LD + j] = kap * XE[i * LD + j];
    }
    __syncthreads();
    {            // z = Dnu + X Gz + eta (kap XN)^T + nu (kap XE)^T
        const double* const As[3] = {X, E, Nu};
        const double* const Gs[3] = {Gn, T0, T1};
        const bool GT[3] = {false, true, true};
        addmul_mma<P, 3>(Zd, 1.0, As, Gs, GT, n);
    }
    __syncthreads();
    {            // H eta = z - X sym(X^T z)
        const double* const As[1] = {X};
        const double* const Bs[1] = {Zd};
        double* const Gs[1] = {T0};
        gram_blocks_mma<P, 1>(As, Bs, Gs, n);
    }
    __syncthreads();
    for (int o = threadIdx.x; o < P * P; o += ST_THREADS) {
        const int i = o / P, j = o - i * P;
        T1[i * LD + j] = -0.5 * (T0[i * LD + j] + T0[j * LD + i]);
    }
    __syncthreads();
    {
        const double* const As[1] = {X};
        const double* const Gs[1] = {T1};
        const bool GT[1] = {false};
        addmul_mma<P, 1>(Zd, 1.0, As, Gs, GT, n);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < np; i += ST_THREADS) Pm.heta[base + i] = Zd[(i / P) * LD + (i % P)];
}

}  // namespace otn


// ==================================================================================================================
// Streaming version of the Riemannian epilogue for p = 16 (full-sites model at N = 4).  riem_mma_kernel above stages x, eta,
// zraw, zdraw of one isometry in shared memory (184 KB at K = 16: one CTA per SM, every phase latency-exposed, K <= 20).  Here
// nothing n x p is staged.  Every output is a combination of the four inputs with p x p coefficient matrices,
//     rgrad = zraw * (1/(f a0)) + x Gn
//     H eta = zdraw * (1/(f a0)) + zraw Cz' + eta Ce + x Cx
// (derivation pinned on the CPU by tests/test_riem_coefficient_form.py against the oracle's step-by-step form of
// stiefel.py:398-409, :468-475, :23-30), and the coefficients only need the five Grams x^T x, x^T zraw, x^T zdraw, x^T eta,
// eta^T zraw.  Stage 1: DMMA Grams with fragments loaded straight from global memory (a fragment row is 64 contiguous bytes),
// rows dealt round-robin to 4 warp groups, the two column halves to 2; partial Grams meet in shared memory and are summed in a
// fixed order.  Stage 2: the 16 x 16 algebra, one thread per entry.  Stage 3: the outputs, 8 rows per warp and step, inputs
// re-read (L2 hits) as DMMA A fragments, coefficient matrices from shared memory.  53 KB of shared memory -> 4 CTAs per SM,
// whose stages overlap; no limit on the Kraus rank.
// ==================================================================================================================
namespace otn {

constexpr int RS_THREADS = 256;
template <bool HVP>
struct RiemStreamCfg {
    static constexpr int P = 16, LD = 20, NQ = HVP ? 5 : 1, KS = 4;
    static constexpr int GRAMS = NQ * P * LD;            // reduced Grams (padded rows)
    static constexpr int PART = KS * NQ * P * P;         // partial Grams; the derived matrices reuse this space afterwards
    static constexpr size_t SMEM_BYTES = (size_t)(GRAMS + PART) * sizeof(double);
};

// (3 CTAs per SM: 80 registers.  Capping them at 64 for a fourth CTA spills and measured slower: 0.29 vs 0.23 ms per step.)
template <bool HVP>
__global__ void __launch_bounds__(RS_THREADS, 3) riem_stream_kernel(const RiemParams Pm) {
    using Cfg = RiemStreamCfg<HVP>;
    constexpr int P = Cfg::P, LD = Cfg::LD, NQ = Cfg::NQ, GS = P * LD;
    enum { QXZ = 0, QXX = 1, QXZD = 2, QXE = 3, QEZ = 4 };
    extern __shared__ __align__(16) double sm[];
    const int item = blockIdx.x, inst = item / Pm.m, layer = item - inst * Pm.m;
    if (Pm.active != nullptr && Pm.active[inst] == 0) return;
    const int n = Pm.n, np = n * P;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    double* Gm = sm;                       // [NQ][P][LD]
    double* part = sm + Cfg::GRAMS;        // [KS][NQ][P][P]
    __shared__ double sc[2];
    if (tid == 0) {
        double s = 0.0;
        for (int tt = 0; tt < Pm.tiles; ++tt) s += Pm.part_f[(long long)inst * Pm.tiles + tt];
        s += Pm.im2[Pm.tindex ? Pm.tindex[inst] : 0];
        const double f = sqrt(s);
        sc[0] = f;
        if (HVP) {
            double dsum = 0.0;
            for (int tt = 0; tt < Pm.tiles; ++tt) dsum += Pm.part_s[(long long)inst * Pm.tiles + tt];
            sc[1] = dsum;
        }
        if (layer == 0 && Pm.f_out != nullptr) Pm.f_out[inst] = f;
    }
    const long long base = (long long)item * np;
    const double* __restrict__ xg = Pm.x + base;
    const double* __restrict__ zg = Pm.zraw + base;
    const double* __restrict__ eg = HVP ? Pm.eta + base : nullptr;
    const double* __restrict__ zdg = HVP ? Pm.zdraw + base : nullptr;

    // ---- stage 1: partial Grams -----------------------------------------------------------------------------------------
    {
        const int ks = warp & 3, jh = warp >> 2;
        double acc[NQ][2][2];
#pragma unroll
        for (int q = 0; q < NQ; ++q)
#pragma unroll
            for (int i = 0; i < 2; ++i) acc[q][i][0] = acc[q][i][1] = 0.0;
        const int ksteps = n >> 2;
#pragma unroll 4
        for (int s = ks; s < ksteps; s += Cfg::KS) {
            const int ro = (s * 4 + t) * P;
            // The row index of an MMA output can be permuted freely: A-fragment row g of block i stands for Gram row 2g+i (not
            // 8i+g), so both blocks come from ONE 16-byte load (ncu of the 8-byte version: L1/TEX 88 % busy, one wavefront per
            // 128-byte line touched by every load instruction).
            const double2 xa = *reinterpret_cast<const double2*>(xg + ro + 2 * g);
            const double xa0 = xa.x, xa1 = xa.y;
            const double zb = zg[ro + jh * 8 + g];
            dmma884(acc[QXZ][0][0], acc[QXZ][0][1], xa0, zb);
            dmma884(acc[QXZ][1][0], acc[QXZ][1][1], xa1, zb);
            if (HVP) {
                const double2 ea = *reinterpret_cast<const double2*>(eg + ro + 2 * g);
                const double ea0 = ea.x, ea1 = ea.y;
                const double zdb = zdg[ro + jh * 8 + g];
                // B fragments of the x^T x / x^T eta Grams: column jh*8+g of the same row (separate 8-byte loads: the column
                // index of the B operand is not permuted)
                const double xb = xg[ro + jh * 8 + g], eb = eg[ro + jh * 8 + g];
                dmma884(acc[QXX][0][0], acc[QXX][0][1], xa0, xb);
                dmma884(acc[QXX][1][0], acc[QXX][1][1], xa1, xb);
                dmma884(acc[QXZD][0][0], acc[QXZD][0][1], xa0, zdb);
                dmma884(acc[QXZD][1][0], acc[QXZD][1][1], xa1, zdb);
                dmma884(acc[QXE][0][0], acc[QXE][0][1], xa0, eb);
                dmma884(acc[QXE][1][0], acc[QXE][1][1], xa1, eb);
                dmma884(acc[QEZ][0][0], acc[QEZ][0][1], ea0, zb);
                dmma884(acc[QEZ][1][0], acc[QEZ][1][1], ea1, zb);
            }
        }
#pragma unroll
        for (int q = 0; q < NQ; ++q)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                double* o = part + ((ks * NQ + q) * P + 2 * g + i) * P + jh * 8 + 2 * t;
                o[0] = acc[q][i][0]; o[1] = acc[q][i][1];
            }
    }
    __syncthreads();
    for (int o = tid; o < NQ * P * P; o += RS_THREADS) {
        const int q = o / (P * P), rc = o - q * P * P, r = rc / P, c = rc - r * P;
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < Cfg::KS; ++k) s += part[(k * NQ + q) * P * P + rc];
        Gm[q * GS + r * LD + c] = s;
    }
    __syncthreads();

    // ---- stage 2: coefficient matrices (thread = entry (i, j)); derived matrices live where the partials were ---------------
    const double f = sc[0], inv_f = 1.0 / f;
    const double a0 = Pm.alpha0, a1 = Pm.alpha1;
    const double c1 = 0.5 * ((1.0 / a1) - (2.0 / a0)), c2 = 0.5 * (1.0 / a1), kap = (a0 - a1) / a0;
    const int i = tid / P, j = tid - i * P, ij = i * LD + j, ji = j * LD + i;
    double* XZ = Gm + QXZ * GS;
    double* Gn = part;
    double *Cx = nullptr, *Czp = nullptr, *Ce = nullptr;
    double kappa = 0.0;
    if (!HVP) {
        XZ[ij] *= inv_f;
        __syncthreads();
        Gn[ij] = c1 * XZ[ij] - c2 * XZ[ji];
    } else {
        double *XX = Gm + QXX * GS, *XZd = Gm + QXZD * GS, *XE = Gm + QXE * GS, *EZ = Gm + QEZ * GS;
        double *Gd = part + GS, *XN = part + 2 * GS, *EN = part + 3 * GS, *Cz = part + 4 * GS, *Cx0 = part + 6 * GS, *T0 = part + 7 * GS;
        Ce = part + 5 * GS; Cx = Cx0; Czp = Cz;
        kappa = sc[1] * inv_f * inv_f * inv_f;
        {
            const double xzr = XZ[ij];
            XZ[ij] = inv_f * xzr;
            XZd[ij] = inv_f * XZd[ij] - kappa * xzr;
            EZ[ij] *= inv_f;
        }
        __syncthreads();
        Gn[ij] = c1 * XZ[ij] - c2 * XZ[ji];
        Gd[ij] = c1 * (EZ[ij] + XZd[ij]) - c2 * (XZd[ji] + EZ[ji]);
        __syncthreads();
        {   // XN = X^T nu = XZ/a0 + XX Gn ;  EN = eta^T nu = EZ/a0 + XE^T Gn
            double sx = 0.0, se = 0.0;
#pragma unroll
            for (int k = 0; k < P; ++k) {
                const double gk = Gn[k * LD + j];
                sx = fma(XX[i * LD + k], gk, sx);
                se = fma(XE[k * LD + i], gk, se);
            }
            XN[ij] = XZ[ij] / a0 + sx;
            EN[ij] = EZ[ij] / a0 + se;
        }
        __syncthreads();
        {   // Gz = sym(EN) - kap (XE XN^T + XN XE^T);  Ce = Gn + kap XN^T;  Cz = (kap/a0) XE^T;  Cx0 = Gd + Gz + kap Gn XE^T
            double s = 0.0, w = 0.0;
#pragma unroll
            for (int k = 0; k < P; ++k) {
                s += XE[i * LD + k] * XN[j * LD + k] + XN[i * LD + k] * XE[j * LD + k];
                w = fma(Gn[i * LD + k], XE[j * LD + k], w);
            }
            const double gz = 0.5 * (EN[ij] + EN[ji]) - kap * s;
            Ce[ij] = Gn[ij] + kap * XN[ji];
            Cz[ij] = (kap / a0) * XE[ji];
            Cx0[ij] = Gd[ij] + gz + kap * w;
        }
        __syncthreads();
        {   // T0 = X^T z = XZd/a0 + XZ Cz + XE Ce + XX Cx0
            double s = XZd[ij] / a0;
#pragma unroll
            for (int k = 0; k < P; ++k)
                s += XZ[i * LD + k] * Cz[k * LD + j] + XE[i * LD + k] * Ce[k * LD + j] + XX[i * LD + k] * Cx0[k * LD + j];
            T0[ij] = s;
        }
        __syncthreads();
        {   // H eta = z - X sym(X^T z);  zraw enters as Z = zraw/f and through Zd = zdraw/f - kappa zraw
            const double cx = Cx0[ij] - 0.5 * (T0[ij] + T0[ji]);
            const double cz = inv_f * Cz[ij] - ((i == j) ? kappa / a0 : 0.0);
            Cx[ij] = cx;
            Czp[ij] = cz;
        }
    }
    __syncthreads();
    // Stage 3 contracts over k = 4t + s in step s (vector loads of the inputs, below).  The coefficient matrices are its B operands:
    // row k is re-filed at row (k & 3) * 4 + (k >> 2), so that step s reads row 4s + t -- the same conflict-free bank pattern
    // (4t + g) as before the permutation.
    {
        const int pi = ((i & 3) << 2) + (i >> 2), pij = pi * LD + j;
        double* perm = part + (HVP ? 8 : 1) * GS;            // free part of the partial-Gram area
        const double v0 = Gn[ij];
        double v1 = 0.0, v2 = 0.0, v3 = 0.0;
        if (HVP) { v1 = Cx[ij]; v2 = Czp[ij]; v3 = Ce[ij]; }
        __syncthreads();
        perm[pij] = v0; Gn = perm;
        if (HVP) { perm[GS + pij] = v1; perm[2 * GS + pij] = v2; perm[3 * GS + pij] = v3; Cx = perm + GS; Czp = perm + 2 * GS; Ce = perm + 3 * GS; }
        __syncthreads();
    }

    // ---- stage 3: outputs, 8 rows per warp and step ---------------------------------------------------------------------------
    const double sz = inv_f / a0;
    for (int rb = warp; rb < (n >> 3); rb += RS_THREADS / 32) {
        const int row = rb * 8 + g;
        double accN[2][2], accH[2][2];
        double2 zr2[2];
#pragma unroll
        for (int jb = 0; jb < 2; ++jb) {
            const int col = jb * 8 + 2 * t;
            zr2[jb] = *reinterpret_cast<const double2*>(zg + row * P + col);
            accN[jb][0] = sz * zr2[jb].x; accN[jb][1] = sz * zr2[jb].y;
            if (HVP) {
                const double2 zd2 = *reinterpret_cast<const double2*>(zdg + row * P + col);
                accH[jb][0] = sz * zd2.x; accH[jb][1] = sz * zd2.y;
            }
        }
        // The contraction index of an MMA can be permuted as long as A and B agree: in step s lane t takes k = 4t + s (not 4s + t),
        // so the four steps read x[row][4t .. 4t+3] -- two 16-byte loads per array and row instead of four 8-byte ones.
        double xr4[4], zr4[4], er4[4];
        {
            const double2 a0 = *reinterpret_cast<const double2*>(xg + row * P + 4 * t), a1 = *reinterpret_cast<const double2*>(xg + row * P + 4 * t + 2);
            xr4[0] = a0.x; xr4[1] = a0.y; xr4[2] = a1.x; xr4[3] = a1.y;
            if (HVP) {
                const double2 b0 = *reinterpret_cast<const double2*>(zg + row * P + 4 * t), b1 = *reinterpret_cast<const double2*>(zg + row * P + 4 * t + 2);
                const double2 c0 = *reinterpret_cast<const double2*>(eg + row * P + 4 * t), c1 = *reinterpret_cast<const double2*>(eg + row * P + 4 * t + 2);
                zr4[0] = b0.x; zr4[1] = b0.y; zr4[2] = b1.x; zr4[3] = b1.y;
                er4[0] = c0.x; er4[1] = c0.y; er4[2] = c1.x; er4[3] = c1.y;
            }
        }
#pragma unroll
        for (int s = 0; s < 4; ++s) {
#pragma unroll
            for (int jb = 0; jb < 2; ++jb) {
                const int bo = (4 * s + t) * LD + jb * 8 + g;          // row 4t + s of the matrix, re-filed (see above)
                dmma884(accN[jb][0], accN[jb][1], xr4[s], Gn[bo]);
                if (HVP) {
                    dmma884(accH[jb][0], accH[jb][1], xr4[s], Cx[bo]);
                    dmma884(accH[jb][0], accH[jb][1], zr4[s], Czp[bo]);
                    dmma884(accH[jb][0], accH[jb][1], er4[s], Ce[bo]);
                }
            }
        }
#pragma unroll
        for (int jb = 0; jb < 2; ++jb) {
            const long long og = base + row * P + jb * 8 + 2 * t;
            if (Pm.rgrad != nullptr) *reinterpret_cast<double2*>(Pm.rgrad + og) = make_double2(accN[jb][0], accN[jb][1]);
            if (!HVP && Pm.egrad != nullptr) *reinterpret_cast<double2*>(Pm.egrad + og) = make_double2(inv_f * zr2[jb].x, inv_f * zr2[jb].y);
            if (HVP) *reinterpret_cast<double2*>(Pm.heta + og) = make_double2(accH[jb][0], accH[jb][1]);
        }
    }
}

}  // namespace otn


// Streaming canonical inner product for p = 16: out[inst] = sum_l tr(a_l^T b_l) - tr((x_l^T a_l)^T (x_l^T b_l)) / 2, same Gram scheme
// as riem_stream_kernel (fragments from global memory, 4 row groups x 2 column halves, fixed summation order), 16 KB of shared
// memory instead of three staged n x p arrays.  The fragment loads of a and b touch every entry exactly once, so the Frobenius
// part rides along.
namespace otn {

__global__ void __launch_bounds__(RS_THREADS) cdot_stream_kernel(const double* __restrict__ x, const double* __restrict__ a,
                                                                 const double* __restrict__ bb, double* __restrict__ out, int n, int m,
                                                                 const int* active) {
    constexpr int P = 16, KS = 4;
    __shared__ double part[KS * 2 * P * P];
    __shared__ double red[RS_THREADS / 32];
    const int inst = blockIdx.x;
    if (active != nullptr && active[inst] == 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int ks = warp & 3, jh = warp >> 2, np = n * P, ksteps = n >> 2;
    double v = 0.0;
    for (int l = 0; l < m; ++l) {
        const long long base = ((long long)inst * m + l) * np;
        const double* __restrict__ xg = x + base;
        const double* __restrict__ ag = a + base;
        const double* __restrict__ bg = bb + base;
        double acc[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
#pragma unroll 4
        for (int s = ks; s < ksteps; s += KS) {
            const int ro = (s * 4 + t) * P;
            const double xa0 = xg[ro + g], xa1 = xg[ro + 8 + g];
            const double av = ag[ro + jh * 8 + g], bv = bg[ro + jh * 8 + g];
            v = fma(av, bv, v);
            dmma884(acc[0][0][0], acc[0][0][1], xa0, av);
            dmma884(acc[0][1][0], acc[0][1][1], xa1, av);
            dmma884(acc[1][0][0], acc[1][0][1], xa0, bv);
            dmma884(acc[1][1][0], acc[1][1][1], xa1, bv);
        }
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                double* o = part + ((ks * 2 + q) * P + i * 8 + g) * P + jh * 8 + 2 * t;
                o[0] = acc[q][i][0]; o[1] = acc[q][i][1];
            }
        __syncthreads();
        {
            double ga = 0.0, gb = 0.0;
#pragma unroll
            for (int k = 0; k < KS; ++k) { ga += part[(k * 2 + 0) * P * P + tid]; gb += part[(k * 2 + 1) * P * P + tid]; }
            v = fma(-0.5 * ga, gb, v);
        }
        __syncthreads();
    }
    v = warp_sum(v);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < RS_THREADS / 32; ++w) s += red[w];
        out[inst] = s;
    }
}

}  // namespace otn


// ==================================================================================================================
// DMMA-Gram versions of the canonical dot product and of the polar retraction for p in {8, 16} (padded rows, LD = p + 4)
// ==================================================================================================================
namespace otn {

template <int P>
__global__ void __launch_bounds__(ST_THREADS) cdot_mma_kernel(const double* __restrict__ x, const double* __restrict__ a,
                                                              const double* __restrict__ bb, double* __restrict__ out, int n, int m,
                                                              const int* active) {
    extern __shared__ __align__(16) double sm[];
    constexpr int LD = P + 4;
    const int inst = blockIdx.x;
    if (active != nullptr && active[inst] == 0) return;
    const int np = n * P, nl = n * LD;
    double* X = sm;
    double* A = X + nl;
    double* B = A + nl;
    double* GA = B + nl;
    double* GB = GA + P * LD;
    __shared__ double red[ST_THREADS / 32];
    double total = 0.0;
    for (int l = 0; l < m; ++l) {
        const long long base = ((long long)inst * m + l) * np;
        double s = 0.0;
        for (int i = threadIdx.x; i < np; i += ST_THREADS) {
            const int o = (i / P) * LD + (i % P);
            const double av = a[base + i], bv = bb[base + i];
            X[o] = x[base + i]; A[o] = av; B[o] = bv;
            s = fma(av, bv, s);
        }
        __syncthreads();
        {
            const double* const As[2] = {X, X}; const double* const Bs[2] = {A, B}; double* const Gs[2] = {GA, GB};
            gram_blocks_mma<P, 2>(As, Bs, Gs, n);
        }
        __syncthreads();
        for (int oo = threadIdx.x; oo < P * P; oo += ST_THREADS) { const int o = (oo / P) * LD + (oo % P); s = fma(-0.5 * GA[o], GB[o], s); }
        s = warp_sum(s);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0)
            for (int w = 0; w < ST_THREADS / 32; ++w) total += red[w];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[inst] = total;
}

template <int P>
__global__ void __launch_bounds__(ST_THREADS) retract_mma_kernel(const double* __restrict__ x, const double* __restrict__ eta,
                                                                 double* __restrict__ out, int n, int m, const int* active, int* status) {
    extern __shared__ __align__(16) double sm[];
    constexpr int LD = P + 4;
    const int item = blockIdx.x;
    if (active != nullptr && active[item / m] == 0) return;
    const int np = n * P, nl = n * LD;
    double* A = sm;              // padded
    double* O = A + nl;          // padded output
    double* Gp = O + nl;         // padded Gram
    double* Q = Gp + P * LD;     // padded inverse square root
    double* S = Q + P * LD;      // dense p x p work arrays
    double* V = S + P * P;
    double* Qd = V + P * P;
    double* Yw = Qd + P * P;
    double* Tw = Yw + P * P;
    __shared__ double cs[2 * ST_MAXP];
    __shared__ int pairs[2 * ST_MAXP];
    __shared__ double red[ST_WARPS + 1];
    const long long base = (long long)item * np;
    for (int i = threadIdx.x; i < np; i += ST_THREADS) {
        const int o = (i / P) * LD + (i % P);
        A[o] = x[base + i] + (eta ? eta[base + i] : 0.0);
        O[o] = 0.0;
    }
    __syncthreads();
    {
        const double* const As[1] = {A}; const double* const Bs[1] = {A}; double* const Gs[1] = {Gp};
        gram_blocks_mma<P, 1>(As, Bs, Gs, n);
    }
    __syncthreads();
    for (int oo = threadIdx.x; oo < P * P; oo += ST_THREADS) S[oo] = Gp[(oo / P) * LD + (oo % P)];
    __syncthreads();
    const int bad = inv_sqrt_spd(S, Qd, Yw, V, Tw, P, cs, pairs, red);
    if (threadIdx.x == 0 && status != nullptr && bad) atomicOr(&status[item / m], bad);
    for (int oo = threadIdx.x; oo < P * P; oo += ST_THREADS) Q[(oo / P) * LD + (oo % P)] = Qd[oo];
    __syncthreads();
    {
        const double* const As[1] = {A}; const double* const Gs[1] = {Q}; const bool GT[1] = {false};
        addmul_mma<P, 1>(O, 0.0, As, Gs, GT, n);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < np; i += ST_THREADS) out[base + i] = O[(i / P) * LD + (i % P)];
}

}  // namespace otn

namespace otn {

// riemannian_connection (stiefel.py:468-475):  D_nu + X (eta^T nu + nu^T eta)/2 + kap (I - X X^T)(eta nu^T + nu eta^T) X
// one CTA per isometry, any p <= ST_MAXP (CUDA cores; the fused HVP epilogue riem_*_kernel is the hot-path version)
__global__ void __launch_bounds__(ST_THREADS) connection_kernel(const double* __restrict__ dnu, const double* __restrict__ nu,
                                                                const double* __restrict__ eta, const double* __restrict__ x,
                                                                double* __restrict__ out, int n, int p, double kap) {
    extern __shared__ __align__(16) double sm[];
    const int np = n * p, pp = p * p;
    double* X = sm;
    double* E = X + np;
    double* Nu = E + np;
    double* Z = Nu + np;
    double* EN = Z + np;
    double* XN = EN + pp;
    double* XE = XN + pp;
    double* T0 = XE + pp;
    double* T1 = T0 + pp;
    const long long base = (long long)blockIdx.x * np;
    for (int i = threadIdx.x; i < np; i += ST_THREADS) { X[i] = x[base + i]; E[i] = eta[base + i]; Nu[i] = nu[base + i]; Z[i] = dnu[base + i]; }
    __syncthreads();
    gram(E, Nu, EN, n, p); gram(X, Nu, XN, n, p); gram(X, E, XE, n, p);
    __syncthreads();
    for (int o = threadIdx.x; o < pp; o += ST_THREADS) {
        const int i = o / p, j = o - i * p;
        double s = 0.0;
        for (int k = 0; k < p; ++k) s += XE[i * p + k] * XN[j * p + k] + XN[i * p + k] * XE[j * p + k];
        T0[o] = 0.5 * (EN[o] + EN[j * p + i]) - kap * s;
    }
    __syncthreads();
    addmul<false>(Z, 1.0, X, T0, n, p);
    __syncthreads();
    for (int o = threadIdx.x; o < pp; o += ST_THREADS) { T0[o] = kap * XN[o]; T1[o] = kap * XE[o]; }
    __syncthreads();
    addmul<true>(Z, 1.0, E, T0, n, p);
    __syncthreads();
    addmul<true>(Z, 1.0, Nu, T1, n, p);
    __syncthreads();
    for (int i = threadIdx.x; i < np; i += ST_THREADS) out[base + i] = Z[i];
}

// || A - (B_re + i B_im) ||_F (squared if asked) for `count` pairs of length-n arrays; A real (frobenius_norm, optimization.py:56-61)
__global__ void __launch_bounds__(256) frobenius_kernel(const double* __restrict__ A, const double* __restrict__ Bre,
                                                        const double* __restrict__ Bim, long long n, int squared, double* __restrict__ out) {
    __shared__ double red[8];
    const long long base = (long long)blockIdx.x * n;
    double s = 0.0;
    for (long long i = threadIdx.x; i < n; i += 256) {
        const double r = A[base + i] - Bre[base + i];
        s = fma(r, r, s);
        if (Bim != nullptr) { const double im = Bim[base + i]; s = fma(im, im, s); }
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        out[blockIdx.x] = squared ? t : sqrt(t);
    }
}

}  // namespace otn
