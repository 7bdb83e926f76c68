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
// back_embed for three Kronecker factors of side 16 (N = 6, d = 2: D = 4096), dense layout.  back_embed_kernel above gathers
// every W[r,c] once per factor with 8-byte scattered loads from 96 CTAs (1.4 ms per layer and pass at D = 4096 -- more than a
// row-sharded chain product on 8 GPUs).  Here the adjoint is contracted factor by factor, reading W in whole rows:
//   stage 1 (WHICH = 2):  A[(i0,i1),(j0,j1)] = sum_{i2,j2} W[r(i0,i1,i2), c(j0,j1,j2)] Yl[i2,j2]
//           (WHICH = 0):  B[(i1,i2),(j1,j2)] = sum_{i0,j0} W[r(i0,i1,i2), c(j0,j1,j2)] Yl[i0,j0]
//   stage 2:  Wl_0[i0,j0] = sum_{i1,j1} A Yl[i1,j1];  Wl_1[i1,j1] = sum_{i0,j0} A Yl[i0,j0];  Wl_2[i2,j2] = sum_{i1,j1} B Yl[i1,j1]
// and the same with dual numbers (Wd, Yld) for the tangent pass.  One stage-1 CTA owns the 16 rows that share the two kept row
// digits: per column c it sums the 16 rows against Yl[., jw(c)] (coalesced row reads), then folds the 16 columns that share the
// kept column digits through shared memory.  grid = (256, 2 = {A, B}, instances).  Fixed summation order: deterministic.
// ------------------------------------------------------------------------------------------------------------------
constexpr int BE3_THREADS = 256;
struct Be3Peers { double* p[7]; int n; };      // the same scratch array on the other ranks of a row-sharded handle (n = 0: none)
constexpr size_t BE3_SMEM = (2 * 256 + 2 * 4096) * sizeof(double) + 4096 * sizeof(int);

template <bool TANGENT>
__global__ void __launch_bounds__(BE3_THREADS) back_embed3_stage1_kernel(const double* __restrict__ W, const double* __restrict__ Wd,
                                                                         long long w_stride, long long wd_stride,
                                                                         const double* __restrict__ Yl, const double* __restrict__ Yld,
                                                                         double* __restrict__ AB, Be3Peers peers, int x0, EmbedTables T,
                                                                         const int* active, int m, int layer) {
    extern __shared__ __align__(16) double sm[];
    const int inst = blockIdx.z;
    if (active != nullptr && active[inst] == 0) return;
    const int which = blockIdx.y == 0 ? 2 : 0;           // the factor contracted here
    const int ka = which == 2 ? 0 : 1, kb = which == 2 ? 1 : 2;   // the two kept factors, in order
    const int item = inst * m + layer, shift = layer & 1;
    constexpr int D = 4096;
    double* yl = sm; double* yld = sm + 256; double* s = sm + 512; double* sd = s + D;
    int* rinv = reinterpret_cast<int*>(sd + D);
    const unsigned* tab = T.tab + (size_t)shift * D;
    for (int i = threadIdx.x; i < 256; i += BE3_THREADS) {
        yl[i] = Yl[(long long)item * 256 + i];
        yld[i] = TANGENT ? Yld[(long long)item * 256 + i] : 0.0;
    }
    for (int r = threadIdx.x; r < D; r += BE3_THREADS) {
        const unsigned pk = tab[r];
        rinv[loc_idx(pk, 0) + 16 * loc_idx(pk, 1) + 256 * loc_idx(pk, 2)] = r;
    }
    __syncthreads();
    const int bx = x0 + blockIdx.x;                       // (x0: first CTA of this rank on a row-sharded handle)
    const int ia = bx >> 4, ib = bx & 15;
    const int sa = 4 * ka, sb = 4 * kb, sw = 4 * which;    // digit f of the packed index sits at bit 4 f
    const int rbase = (ia << sa) + (ib << sb);
    const double* Wb = W + (long long)inst * w_stride;
    const double* Wdb = TANGENT ? Wd + (long long)inst * wd_stride : nullptr;
    for (int c = threadIdx.x; c < D; c += BE3_THREADS) {
        const int jw = loc_idx(tab[c], which);
        double acc = 0.0, accd = 0.0;
#pragma unroll
        for (int iw = 0; iw < 16; ++iw) {                 // 16 (32 with the tangent) independent loads in flight per thread
            const long long o = (long long)rinv[rbase + (iw << sw)] * D + c;
            const double w = Wb[o], y = yl[iw * 16 + jw];
            acc = fma(w, y, acc);
            if (TANGENT) { accd = fma(Wdb[o], y, accd); accd = fma(w, yld[iw * 16 + jw], accd); }
        }
        s[c] = acc;
        if (TANGENT) sd[c] = accd;
    }
    __syncthreads();
    {
        const int ja = threadIdx.x >> 4, jb = threadIdx.x & 15;
        const int cbase = (ja << sa) + (jb << sb);
        double acc = 0.0, accd = 0.0;
        for (int jw = 0; jw < 16; ++jw) {
            const int c = rinv[cbase + (jw << sw)];
            acc += s[c];
            if (TANGENT) accd += sd[c];
        }
        // AB[inst][{A, B}][{primal, dual}][(ia,ib)][(ja,jb)]
        const long long o = (((long long)inst * 2 + blockIdx.y) * 2) * 65536 + (long long)bx * 256 + threadIdx.x;
        AB[o] = acc;
        if (TANGENT) AB[o + 65536] = accd;
        for (int q = 0; q < peers.n; ++q) {
            peers.p[q][o] = acc;
            if (TANGENT) peers.p[q][o + 65536] = accd;
        }
    }
}

// stage 2: grid = (3 factors, instances), 256 threads = the (i, j) of the local cotangent
template <bool TANGENT>
__global__ void __launch_bounds__(256) back_embed3_stage2_kernel(const double* __restrict__ AB, const double* __restrict__ Yl,
                                                                 const double* __restrict__ Yld, double* __restrict__ Wl,
                                                                 double* __restrict__ Wld, const int* active, int m, int layer) {
    const int inst = blockIdx.y, f = blockIdx.x;
    if (active != nullptr && active[inst] == 0) return;
    const int item = inst * m + layer;
    __shared__ double yl[256], yld[256];
    yl[threadIdx.x] = Yl[(long long)item * 256 + threadIdx.x];
    yld[threadIdx.x] = TANGENT ? Yld[(long long)item * 256 + threadIdx.x] : 0.0;
    __syncthreads();
    const int i = threadIdx.x >> 4, j = threadIdx.x & 15;
    // f = 0: A[(i,u),(j,v)] Yl[u,v];  f = 1: A[(u,i),(v,j)] Yl[u,v];  f = 2: B[(u,i),(v,j)] Yl[u,v]
    const double* M = AB + (((long long)inst * 2 + (f == 2 ? 1 : 0)) * 2) * 65536;
    const double* Md = M + 65536;
    double acc = 0.0, accd = 0.0;
    for (int u = 0; u < 16; ++u)
#pragma unroll 4
        for (int v = 0; v < 16; ++v) {
            const int o = (f == 0) ? ((i * 16 + u) * 256 + j * 16 + v) : ((u * 16 + i) * 256 + v * 16 + j);
            const double a = M[o], y = yl[u * 16 + v];
            acc = fma(a, y, acc);
            if (TANGENT) { accd = fma(Md[o], y, accd); accd = fma(a, yld[u * 16 + v], accd); }
        }
    const long long o = ((long long)item * 3 + f) * 256 + threadIdx.x;
    if (!TANGENT) Wl[o] = acc; else Wld[o] = accd;
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
                                                                    const int* active, int m, int layer_arg) {
    // launched once per layer: W / Wd point at that layer's matrix of instance 0, strides are per instance;
    // or once for all layers (layer_arg < 0, grid.z = m): W / Wd point at layer 0 and layers are nslices * d^4 apart
    extern __shared__ __align__(16) double sm[];
    const int inst = blockIdx.y, a = blockIdx.x;
    if (active != nullptr && active[inst] == 0) return;
    const int layer = layer_arg >= 0 ? layer_arg : (int)blockIdx.z;
    const int item = inst * m + layer;
    const int d2 = dim * dim, np = dim * K * dim, ldc = dim + 1, wsz = dim * dim * ldc;
    if (layer_arg < 0) {
        W += (long long)layer * nslices * d2 * d2;
        if (TANGENT) Wd += (long long)layer * nslices * d2 * d2;
    }
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

// ==================================================================================================================
// Tensor-core (DMMA.8x8x4) versions for dim % 8 == 0 (full-sites model: dim = d^N = 16).  The SIMT kernels above need one
// shared-memory load per FMA and ran ~9x slower than the HBM floor; here a row (a,b) of Y is vec(X_a^T X_b), a dim x dim x K
// product fed from fragments (1 LDS per 256 FMA), so the kernels become store / L2-read bound.
// smem rows are padded to dim+4 doubles (bank-conflict-free fragment loads), Kraus index padded with zero rows.
// ==================================================================================================================
namespace otn {

template <int DIM, bool TANGENT, bool PRIMAL>
__global__ void __launch_bounds__(256) assemble_mma_kernel(const double* __restrict__ x, const double* __restrict__ eta,
                                                           double* __restrict__ Y, double* __restrict__ Yd, long long out_stride,
                                                           int K, const int* active, int m) {
    extern __shared__ __align__(16) double sm[];
    constexpr int LD = DIM + 4, MI = DIM / 8;
    const int item = blockIdx.y, a = blockIdx.x;
    if (active != nullptr && active[item / m] == 0) return;
    const int Kp = (K + 3) & ~3;
    const int rows = DIM * Kp;
    double* xs = sm;
    double* es = sm + rows * LD;
    const long long np = (long long)DIM * K * DIM;
    const double* xg = x + (long long)item * np;
    const double* eg = TANGENT ? eta + (long long)item * np : nullptr;
    for (int i = threadIdx.x; i < rows * DIM; i += 256) {
        const int r = i / DIM, c = i - r * DIM, aa = r / Kp, k = r - aa * Kp;
        const bool ok = k < K;
        xs[r * LD + c] = ok ? xg[(aa * K + k) * DIM + c] : 0.0;
        if (TANGENT) es[r * LD + c] = ok ? eg[(aa * K + k) * DIM + c] : 0.0;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    for (int b = warp; b < DIM; b += 8) {
        double acc[MI][MI][2], accd[MI][MI][2];
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < MI; ++j) { acc[i][j][0] = acc[i][j][1] = 0.0; accd[i][j][0] = accd[i][j][1] = 0.0; }
        for (int kk = 0; kk < Kp; kk += 4) {
            double xa[MI], xb[MI], ea[MI], eb[MI];
#pragma unroll
            for (int i = 0; i < MI; ++i) {
                xa[i] = xs[(a * Kp + kk + t) * LD + i * 8 + g];
                xb[i] = xs[(b * Kp + kk + t) * LD + i * 8 + g];
                if (TANGENT) { ea[i] = es[(a * Kp + kk + t) * LD + i * 8 + g]; eb[i] = es[(b * Kp + kk + t) * LD + i * 8 + g]; }
            }
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < MI; ++j) {
                    if (PRIMAL) dmma884(acc[i][j][0], acc[i][j][1], xa[i], xb[j]);
                    if (TANGENT) { dmma884(accd[i][j][0], accd[i][j][1], ea[i], xb[j]); dmma884(accd[i][j][0], accd[i][j][1], xa[i], eb[j]); }
                }
        }
        const long long row = (long long)item * out_stride + (long long)(a * DIM + b) * (DIM * DIM);
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < MI; ++j) {
                const int off = (i * 8 + g) * DIM + j * 8 + 2 * t;
                if (PRIMAL) *reinterpret_cast<double2*>(Y + row + off) = make_double2(acc[i][j][0], acc[i][j][1]);
                if (TANGENT) *reinterpret_cast<double2*>(Yd + row + off) = make_double2(accd[i][j][0], accd[i][j][1]);
            }
    }
}

}  // namespace otn

// ==================================================================================================================
// Streaming back-assembly for the dense formulation (an earlier one-CTA-per-(a, instance) kernel that re-read W with a
// transposed gather ran 3-7x slower; see git history).
//   out[(a,k),c]  (+)=  2 * sum_{b,d} W[(a,b),(c,d)] * u[(b,k),d]          u = x or eta, W swap-symmetric
// One CTA per (instance, layer) streams the whole D x D cotangent once through a 3-stage cp.async ring (one stage = the
// DIM rows {(a,b)}_a of a fixed b, 32 KB at DIM = 16), so ~64 KB per SM are in flight while the previous stage is consumed
// by DMMA.  Viewed as a GEMM: M = DIM^2 rows (a,c), N = K, contraction (b,d) = DIM^2; each warp owns DIM^2/8 output rows,
// so there is no cross-warp reduction and the summation order is fixed.
// The tangent needs two cotangents with two right-hand sides (Wd*x + W*eta): two launches, the second accumulating.
// ==================================================================================================================
namespace otn {

template <int DIM>
struct BasCfg {
    static constexpr int LD = DIM + 4, STAGES = 3, D2 = DIM * DIM;
    static constexpr int STAGE_ELEMS = DIM * DIM * LD;           // [a][c][LD]
    static size_t smem_bytes(int K) { return (size_t)(STAGES * STAGE_ELEMS + DIM * ((K + 7) & ~7) * LD) * 8; }
};

template <int DIM>
__global__ void __launch_bounds__(256) back_assemble_stream_kernel(const double* __restrict__ W, long long w_stride,
                                                                   const double* __restrict__ u, double* __restrict__ out,
                                                                   int K, int accumulate, const int* active, int m, int layer) {
    using Cfg = BasCfg<DIM>;
    constexpr int LD = Cfg::LD, D2 = Cfg::D2, ST = Cfg::STAGES;
    constexpr int MI = D2 / 64;                 // 8-row blocks per warp (8 warps share D2 rows)
    extern __shared__ __align__(16) double sm[];
    const int inst = blockIdx.x;
    if (active != nullptr && active[inst] == 0) return;
    const int item = inst * m + layer;
    const int Kp = (K + 7) & ~7, NI = Kp / 8;
    double* stage = sm;
    double* us = sm + ST * Cfg::STAGE_ELEMS;    // [b][k (padded)][LD]
    const long long np = (long long)DIM * K * DIM;
    const double* Wb = W + (long long)inst * w_stride;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;

    auto load_stage = [&](int b, int s) {
        // rows (a, b) for a = 0..DIM-1: global row a*DIM + b, D2 doubles each -> smem [a][c][d]
        double* dst = stage + s * Cfg::STAGE_ELEMS;
        constexpr int CH = D2 / 2;              // 16-byte chunks per row
        for (int id = tid; id < DIM * CH; id += 256) {
            const int a = id / CH, ch = id - a * CH, c = (2 * ch) / DIM, dd = 2 * ch - c * DIM;
            cp_async16(dst + (a * DIM + c) * LD + dd, Wb + (long long)(a * DIM + b) * D2 + 2 * ch);
        }
    };
#pragma unroll
    for (int s = 0; s < ST - 1; ++s) { load_stage(s, s); cp_async_commit(); }
    for (int i = tid; i < DIM * Kp * DIM; i += 256) {
        const int r = i / DIM, c = i - r * DIM, bb = r / Kp, k = r - bb * Kp;
        us[r * LD + c] = (k < K) ? u[(long long)item * np + (bb * K + k) * DIM + c] : 0.0;
    }
    double acc[MI][4][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int b = 0; b < DIM; ++b) {
        cp_async_wait<ST - 2>();
        __syncthreads();
        if (b + ST - 1 < DIM) load_stage(b + ST - 1, (b + ST - 1) % ST);
        cp_async_commit();
        const double* As = stage + (b % ST) * Cfg::STAGE_ELEMS;
#pragma unroll
        for (int d0 = 0; d0 < DIM; d0 += 4) {
            double wa[MI];
#pragma unroll
            for (int i = 0; i < MI; ++i) wa[i] = As[(warp * (MI * 8) + i * 8 + g) * LD + d0 + t];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (j < NI) {
                    const double ub = us[(b * Kp + j * 8 + g) * LD + d0 + t];
#pragma unroll
                    for (int i = 0; i < MI; ++i) dmma884(acc[i][j][0], acc[i][j][1], wa[i], ub);
                }
        }
    }
    cp_async_wait<0>();
    // row index R = (a, c) = warp*(MI*8) + i*8 + g ; columns k = j*8 + 2t, +1
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int R = warp * (MI * 8) + i * 8 + g, a = R / DIM, c = R - a * DIM;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (j < NI) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int k = j * 8 + 2 * t + e;
                    if (k < K) {
                        const long long og = (long long)item * np + (a * K + k) * DIM + c;
                        const double v = 2.0 * acc[i][j][e];
                        out[og] = accumulate ? out[og] + v : v;
                    }
                }
            }
    }
}

}  // namespace otn
