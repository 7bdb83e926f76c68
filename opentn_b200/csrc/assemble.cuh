// Kraus (x) conj assembly and its adjoint / tangent forms (sites A, B of the hot path), FP64, batched over "items"
// (item = instance * m + layer).
//
//   assemble_kernel       x -> Y = sum_k E_k (x) E_k           ortho2super, opentn/transformations.py:743-749
//                                                              (= ortho2choi :716 -> unfactorize_psd :804 -> choi2super :96;
//                                                               closed form kraus2superop :170-180 with E_k = x[k::K, :])
//   embed_kernel          Y_loc -> Y_full = perm(Y_loc^{(x) N/2})  create_supertensored_from_local :405-429 +
//                                                              convert_supertensored2liouvillianfull :484-490 (pbc shift :477-480)
//   back_embed_kernel     adjoint of embed (and its directional derivative)       } what jax.grad / jax.jvp generate
//   back_assemble_kernel  adjoint of assemble (and its directional derivative)    } from the two maps above
//
// Index conventions: isometry row a*K+k, column c; superoperator row (a,b) = a*dim+b, column (c,d).
#pragma once
#include "otn_common.cuh"

namespace otn {

// ------------------------------------------------------------------------------------------------------------------
// assemble: grid = (dim, items), block = ASM_THREADS.  CTA (a, item) writes rows (a, b = 0..dim-1) of Y.
//   Y[(a,b),(c,d)]  = sum_k x[aK+k,c] x[bK+k,d]
//   Yd[(a,b),(c,d)] = sum_k e[aK+k,c] x[bK+k,d] + x[aK+k,c] e[bK+k,d]            (only if eta != nullptr)
// smem: x item (dim*K*dim doubles) [+ eta item].
// ------------------------------------------------------------------------------------------------------------------
constexpr int ASM_THREADS = 256;
constexpr int ASM_MAXK = 32;

template <bool TANGENT, bool PRIMAL>
__global__ void __launch_bounds__(ASM_THREADS) assemble_kernel(const double* __restrict__ x, const double* __restrict__ eta,
                                                               double* __restrict__ Y, double* __restrict__ Yd,
                                                               long long out_stride, int dim, int K, const int* active, int m) {
    extern __shared__ __align__(16) double sm[];
    const int item = blockIdx.y, a = blockIdx.x;
    if (active != nullptr && active[item / m] == 0) return;
    const int np = dim * K * dim;
    double* xs = sm;
    double* es = sm + np;
    const double* xg = x + (long long)item * np;
    for (int i = threadIdx.x; i < np; i += ASM_THREADS) xs[i] = xg[i];
    if (TANGENT) {
        const double* eg = eta + (long long)item * np;
        for (int i = threadIdx.x; i < np; i += ASM_THREADS) es[i] = eg[i];
    }
    __syncthreads();
    const int d2 = dim * dim;
    for (int cc = threadIdx.x; cc < d2; cc += ASM_THREADS) {
        const int c = cc / dim, dd = cc - c * dim;
        double xa[ASM_MAXK], ea[ASM_MAXK];
#pragma unroll 4
        for (int k = 0; k < K; ++k) {
            xa[k] = xs[(a * K + k) * dim + c];
            if (TANGENT) ea[k] = es[(a * K + k) * dim + c];
        }
        for (int b = 0; b < dim; ++b) {
            double s = 0.0, sd = 0.0;
#pragma unroll 4
            for (int k = 0; k < K; ++k) {
                const double xb = xs[(b * K + k) * dim + dd];
                if (PRIMAL) s = fma(xa[k], xb, s);
                if (TANGENT) {
                    sd = fma(ea[k], xb, sd);
                    sd = fma(xa[k], es[(b * K + k) * dim + dd], sd);
                }
            }
            const long long o = (long long)item * out_stride + (long long)(a * dim + b) * d2 + cc;
            if (PRIMAL) Y[o] = s;
            if (TANGENT) Yd[o] = sd;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Index tables for the local -> full embedding (built on the host, see otn_api.cu):
//   tab[shift][r]         : packed local indices (8 bits per factor f) of full index r
//   inv[shift][f][i][t]   : the D/q full indices r whose f-th local index equals i   (q = d^4)
// ------------------------------------------------------------------------------------------------------------------
struct EmbedTables {
    const unsigned* tab;   // [2][D]
    const int* inv;        // [2][nf][q][D/q]
    int D, q, nf;
};

__device__ __forceinline__ int loc_idx(unsigned packed, int f) { return (packed >> (8 * f)) & 0xff; }

// embed: grid = (D / EMB_ROWS, items); Yfull[r][c] = prod_f Yl[i_f(r)][i_f(c)]   (+ tangent: sum_f Yld_f prod_{g!=f} Yl_g)
constexpr int EMB_THREADS = 256;
constexpr int EMB_ROWS = 8;

template <bool TANGENT, bool PRIMAL>
__global__ void __launch_bounds__(EMB_THREADS) embed_kernel(const double* __restrict__ Yl, const double* __restrict__ Yld,
                                                            double* __restrict__ Yf, double* __restrict__ Yfd, EmbedTables T,
                                                            const int* active, int m) {
    extern __shared__ __align__(16) double sm[];
    const int item = blockIdx.y;
    if (active != nullptr && active[item / m] == 0) return;
    const int layer = item % m;
    const int shift = layer & 1;  // odd list index => pbc shift (optimization.py:132-136)
    const int q2 = T.q * T.q;
    double* yl = sm;
    double* yld = sm + q2;
    for (int i = threadIdx.x; i < q2; i += EMB_THREADS) {
        yl[i] = Yl[(long long)item * q2 + i];
        if (TANGENT) yld[i] = Yld[(long long)item * q2 + i];
    }
    __syncthreads();
    const unsigned* tab = T.tab + (size_t)shift * T.D;
    const int r0 = blockIdx.x * EMB_ROWS;
    const long long base = (long long)item * T.D * T.D;
    for (int c = threadIdx.x; c < T.D; c += EMB_THREADS) {
        const unsigned pc = tab[c];
        for (int rr = 0; rr < EMB_ROWS; ++rr) {
            const int r = r0 + rr;
            if (r >= T.D) break;
            const unsigned pr = tab[r];
            double prod = 1.0, dsum = 0.0;
            for (int f = 0; f < T.nf; ++f) {
                const int e = loc_idx(pr, f) * T.q + loc_idx(pc, f);
                const double v = yl[e];
                if (TANGENT) dsum = dsum * v + prod * yld[e];
                prod *= v;
            }
            if (PRIMAL) Yf[base + (long long)r * T.D + c] = prod;
            if (TANGENT) Yfd[base + (long long)r * T.D + c] = dsum;
        }
    }
}

// back_embed: one warp per output (f, i, j); grid = (ceil(nf*q*q / 8), instances), block 256, one launch per layer.
//   Wl[i,j]  = sum_f sum_{r: i_f(r)=i} sum_{c: i_f(c)=j}  W[r,c]  * prod_{g != f} Yl[i_g(r), i_g(c)]
//   Wld[i,j] = same with Wd  +  sum_f sum ... W[r,c] * d/dt prod_{g != f}(Yl + t Yld)                  (TANGENT)
// Partial results per f are written to scratch [item][nf][q*q] and summed over f in fixed order by the consumer
// (back_assemble_kernel reads `nf` slices) => deterministic.
constexpr int BEM_THREADS = 256;

template <bool TANGENT>
__global__ void __launch_bounds__(BEM_THREADS) back_embed_kernel(const double* __restrict__ W, const double* __restrict__ Wd,
                                                                 long long w_stride, long long wd_stride,
                                                                 const double* __restrict__ Yl, const double* __restrict__ Yld,
                                                                 double* __restrict__ Wl, double* __restrict__ Wld, EmbedTables T,
                                                                 const int* active, int m, int layer) {
    // launched once per layer: W / Wd point at that layer's matrix of instance 0, strides are per instance
    extern __shared__ __align__(16) double sm[];
    const int inst = blockIdx.y;
    if (active != nullptr && active[inst] == 0) return;
    const int item = inst * m + layer;
    const int shift = layer & 1;
    const int q = T.q, q2 = q * q, per = T.D / q;
    double* yl = sm;
    double* yld = sm + q2;
    for (int i = threadIdx.x; i < q2; i += BEM_THREADS) {
        yl[i] = Yl[(long long)item * q2 + i];
        if (TANGENT) yld[i] = Yld[(long long)item * q2 + i];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int out = blockIdx.x * (BEM_THREADS / 32) + warp;
    if (out >= T.nf * q2) return;
    const int f = out / q2, ij = out - f * q2, i = ij / q, j = ij - i * q;
    const unsigned* tab = T.tab + (size_t)shift * T.D;
    const int* rows = T.inv + (((size_t)shift * T.nf + f) * q + i) * per;
    const int* cols = T.inv + (((size_t)shift * T.nf + f) * q + j) * per;
    const double* Wb = W + (long long)inst * w_stride;
    const double* Wdb = TANGENT ? Wd + (long long)inst * wd_stride : nullptr;
    double acc = 0.0, accd = 0.0;
    const long long terms = (long long)per * per;
    for (long long tix = lane; tix < terms; tix += 32) {
        const int r = rows[tix / per], c = cols[tix % per];
        const unsigned pr = tab[r], pc = tab[c];
        double prod = 1.0, dsum = 0.0;
        for (int g = 0; g < T.nf; ++g) {
            if (g == f) continue;
            const int e = loc_idx(pr, g) * q + loc_idx(pc, g);
            const double v = yl[e];
            if (TANGENT) dsum = dsum * v + prod * yld[e];
            prod *= v;
        }
        const double w = Wb[(long long)r * T.D + c];
        if (!TANGENT) acc = fma(w, prod, acc);
        if (TANGENT) { accd = fma(Wdb[(long long)r * T.D + c], prod, accd); accd = fma(w, dsum, accd); }
    }
    acc = warp_sum(acc);
    accd = warp_sum(accd);
    if (lane == 0) {
        if (!TANGENT) Wl[((long long)item * T.nf + f) * q2 + ij] = acc;
        else Wld[((long long)item * T.nf + f) * q2 + ij] = accd;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// back_assemble: grid = (dim, instances), block 256, one launch per layer.  CTA (a, instance):
//   Z[(a,k),c]  = sum_{b,d} (W[(a,b),(c,d)] + W[(b,a),(d,c)]) x[(b,k),d]
//   Zd[(a,k),c] = sum_{b,d} (Wd[..] + Wd[..]^swap) x[(b,k),d] + (W[..] + W[..]^swap) eta[(b,k),d]       (TANGENT)
// W may be given as `nslices` slices to be summed (the per-factor partials of back_embed).
// smem: Wsym[dim][dim][dim+1] (x2 if TANGENT), x item (+ eta item).
// ------------------------------------------------------------------------------------------------------------------
constexpr int BAS_THREADS = 256;

template <bool TANGENT>
__global__ void __launch_bounds__(BAS_THREADS) back_assemble_kernel(const double* __restrict__ W, const double* __restrict__ Wd,
                                                                    long long w_stride, long long wd_stride, int nslices,
                                                                    const double* __restrict__ x, const double* __restrict__ eta,
                                                                    double* __restrict__ Z, double* __restrict__ Zd, int dim, int K,
                                                                    const int* active, int m, int layer) {
    // launched once per layer: W / Wd point at that layer's matrix of instance 0, strides are per instance
    extern __shared__ __align__(16) double sm[];
    const int inst = blockIdx.y, a = blockIdx.x;
    if (active != nullptr && active[inst] == 0) return;
    const int item = inst * m + layer;
    const int d2 = dim * dim, np = dim * K * dim, ldc = dim + 1, wsz = dim * dim * ldc;
    double* ws = sm;
    double* wds = ws + (TANGENT ? wsz : 0);
    double* xs = wds + wsz;
    double* es = xs + np;
    for (int i = threadIdx.x; i < np; i += BAS_THREADS) {
        xs[i] = x[(long long)item * np + i];
        if (TANGENT) es[i] = eta[(long long)item * np + i];
    }
    const double* Wb = W + (long long)inst * w_stride;
    const double* Wdb = TANGENT ? Wd + (long long)inst * wd_stride : nullptr;
    const long long slice = (long long)d2 * d2;
    // pass 1: rows (a, b), columns (c, d): coalesced along (c, d)
    for (int idx = threadIdx.x; idx < dim * d2; idx += BAS_THREADS) {
        const int b = idx / d2, cd = idx - b * d2, c = cd / dim, dd = cd - c * dim;
        const long long g = (long long)(a * dim + b) * d2 + cd;
        double v = 0.0, vd = 0.0;
        for (int s = 0; s < nslices; ++s) {
            v += Wb[s * slice + g];
            if (TANGENT) vd += Wdb[s * slice + g];
        }
        if (TANGENT) { ws[(b * dim + c) * ldc + dd] = v; wds[(b * dim + c) * ldc + dd] = vd; }
        else wds[(b * dim + c) * ldc + dd] = v;
    }
    __syncthreads();
    // pass 2: rows (b, a), columns (d, c): coalesced along (d, c), transposed into [b][c][d]
    for (int idx = threadIdx.x; idx < dim * d2; idx += BAS_THREADS) {
        const int b = idx / d2, dc = idx - b * d2, dd = dc / dim, c = dc - dd * dim;
        const long long g = (long long)(b * dim + a) * d2 + dc;
        double v = 0.0, vd = 0.0;
        for (int s = 0; s < nslices; ++s) {
            v += Wb[s * slice + g];
            if (TANGENT) vd += Wdb[s * slice + g];
        }
        if (TANGENT) { ws[(b * dim + c) * ldc + dd] += v; wds[(b * dim + c) * ldc + dd] += vd; }
        else wds[(b * dim + c) * ldc + dd] += v;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < K * dim; o += BAS_THREADS) {
        const int k = o / dim, c = o - k * dim;
        double s = 0.0;
        for (int b = 0; b < dim; ++b) {
            const double* wrow = wds + (b * dim + c) * ldc;
            const double* xrow = xs + (b * K + k) * dim;
#pragma unroll 4
            for (int dd = 0; dd < dim; ++dd) s = fma(wrow[dd], xrow[dd], s);
            if (TANGENT) {
                const double* w0 = ws + (b * dim + c) * ldc;
                const double* erow = es + (b * K + k) * dim;
#pragma unroll 4
                for (int dd = 0; dd < dim; ++dd) s = fma(w0[dd], erow[dd], s);
            }
        }
        const long long og = (long long)item * np + (a * K + k) * dim + c;
        if (TANGENT) Zd[og] = s; else Z[og] = s;
    }
}

}  // namespace otn
