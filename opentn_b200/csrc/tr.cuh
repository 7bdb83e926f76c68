// Batched Riemannian trust region + Steihaug truncated CG (site K of the hot path), lock-step over instances with
// per-instance masks.  Restates opentn/trust_region_rcopt.py:5-166 on (n,p) tangent matrices with the canonical inner
// product  <a,b>_x = sum_l tr(a_l^T (I - x_l x_l^T/2) b_l), which equals the Euclidean dot product of the reference's
// (A-upper, B) coefficient vectors (stiefel.py:266-300, 453-466), so every scalar the reference's tCG computes
// (rsq, dHd, alpha, beta, t, rho) is reproduced without ever forming X_perp.
#pragma once
#include "otn_common.cuh"
#include "stiefel.cuh"

namespace otn {

struct TrState {
    int gd_final = 0;   // which half of the Gd ping-pong buffer holds Ghat_0 after the tangent reverse sweep
    int tiles = 1;      // partial-sum tiles of the cached forward pass
    // device arrays (allocated by the driver)
    double *g = nullptr, *r = nullptr, *z = nullptr, *d = nullptr, *Hd = nullptr, *xn = nullptr, *Hz = nullptr;   // [B][m][np]
    double *radius = nullptr, *fx = nullptr, *fn = nullptr, *rsq = nullptr, *stoptol = nullptr, *gz = nullptr, *zHz = nullptr;
    int *tcg_active = nullptr, *on_boundary = nullptr, *n_cg = nullptr, *status = nullptr, *accepted = nullptr, *n_active = nullptr;
    int* rejected = nullptr;   // 1 - accepted: instances whose cached forward pass (of x_next) does not belong to their x
    double* rho = nullptr;
    int cap_batch = 0;
};

struct TcgParams {
    const double* x;
    double *r, *z, *d;
    double* Hz;            // H z, accumulated alongside z (z = sum_j alpha_j d_j  =>  H z = sum_j alpha_j H d_j): saves the extra
                           // `hess @ eta` product of trust_region_rcopt.py:67
    const double* Hd;
    const double* g;
    double *radius, *rsq, *stoptol;
    int *tcg_active, *on_boundary, *n_cg, *status, *n_active;
    int n, p, m, maxiter;
    double abstol, reltol;
};

// block-wide sum of `nv` per-thread values, fixed order (warp shuffle tree, then warps in index order)
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* red /* [NV][ST_WARPS] */, double (&out)[NV]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double s = warp_sum(v[i]);
        if (lane == 0) red[i * ST_WARPS + warp] = s;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double s = 0.0;
        for (int w = 0; w < ST_THREADS / 32; ++w) s += red[i * ST_WARPS + w];
        out[i] = s;
    }
    __syncthreads();
}

// Shared-memory layout of the small kernels below: PT == 0 -> rows of p doubles, Grams on the CUDA cores (any p);
// PT in {8, 16} -> rows padded to PT + 4 doubles and every Gram issued as DMMA (gram_blocks_mma), which removes the
// 256-long dependent FMA chains that made the first version of the TR step cost as much as an HVP.
template <int PT>
struct TrLay {
    static __device__ __forceinline__ int ld(int p) { return PT ? PT + 4 : p; }
};

template <int PT, int NG>
__device__ __forceinline__ void grams(const double* const (&A)[NG], const double* const (&B)[NG], double* const (&G)[NG], int n, int p) {
    if constexpr (PT != 0) {
        gram_blocks_mma<PT, NG>(A, B, G, n);
    } else {
#pragma unroll
        for (int q = 0; q < NG; ++q) gram(A[q], B[q], G[q], n, p);
    }
}

// r = g, z = 0, d = -g, rsq = <g,g>, stoptol = max(abstol, reltol sqrt(rsq))     trust_region_rcopt.py:110-114
template <int PT>
__global__ void __launch_bounds__(ST_THREADS) tcg_init_kernel(const TcgParams P) {
    extern __shared__ __align__(16) double sm[];
    const int inst = blockIdx.x;
    const int ld = TrLay<PT>::ld(P.p);
    const int np = P.n * P.p, nl = P.n * ld, pl = P.p * ld;
    double* X = sm;
    double* A = X + nl;
    double* GA = A + nl;
    __shared__ double red[ST_WARPS];
    double v[1] = {0.0};
    for (int l = 0; l < P.m; ++l) {
        const long long base = ((long long)inst * P.m + l) * np;
        for (int i = threadIdx.x; i < np; i += ST_THREADS) {
            const double gv = P.g[base + i];
            const int o = (i / P.p) * ld + (i % P.p);
            X[o] = P.x[base + i]; A[o] = gv;
            P.r[base + i] = gv; P.z[base + i] = 0.0; P.d[base + i] = -gv; P.Hz[base + i] = 0.0;
            v[0] = fma(gv, gv, v[0]);
        }
        __syncthreads();
        {
            const double* const As[1] = {X}; const double* const Bs[1] = {A}; double* const Gs[1] = {GA};
            grams<PT, 1>(As, Bs, Gs, P.n, P.p);
        }
        __syncthreads();
        for (int o = threadIdx.x; o < P.p * P.p; o += ST_THREADS) { const double q = GA[(o / P.p) * ld + (o % P.p)]; v[0] = fma(-0.5 * q, q, v[0]); }
        __syncthreads();
        (void)pl;
    }
    double out[1];
    block_sum<1>(v, red, out);
    if (threadIdx.x == 0) {
        P.rsq[inst] = out[0];
        P.stoptol[inst] = fmax(P.abstol, P.reltol * sqrt(out[0]));
        P.tcg_active[inst] = 1;
        P.on_boundary[inst] = 0;
        P.n_cg[inst] = 0;
    }
}

// one tCG iteration given Hd = H d                                              trust_region_rcopt.py:116-132, 137-166
template <int PT>
__global__ void __launch_bounds__(ST_THREADS) tcg_step_kernel(const TcgParams P) {
    extern __shared__ __align__(16) double sm[];
    const int inst = blockIdx.x;
    if (P.tcg_active[inst] == 0) return;
    if (threadIdx.x == 0) atomicAdd(P.n_active + 1, 1);   // Hessian-vector products consumed (one per active instance and step)
    const int ld = TrLay<PT>::ld(P.p);
    const int np = P.n * P.p, nl = P.n * ld, pl = P.p * ld, pp = P.p * P.p;
    double* X = sm;
    double* A = X + nl;       // d
    double* B = A + nl;       // Hd
    double* C = B + nl;       // z  (pass A) / r (pass B)
    double* GA = C + nl;
    double* GB = GA + pl;
    double* GC = GB + pl;
    __shared__ double red[4 * ST_WARPS];
    __shared__ double sc[4];
    __shared__ int flag;
    // ---- pass A: dHd, dsq, zd, zz
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    for (int l = 0; l < P.m; ++l) {
        const long long base = ((long long)inst * P.m + l) * np;
        for (int i = threadIdx.x; i < np; i += ST_THREADS) {
            const double dv = P.d[base + i], hv = P.Hd[base + i], zv = P.z[base + i];
            const int o = (i / P.p) * ld + (i % P.p);
            X[o] = P.x[base + i]; A[o] = dv; B[o] = hv; C[o] = zv;
            v[0] = fma(dv, hv, v[0]); v[1] = fma(dv, dv, v[1]); v[2] = fma(zv, dv, v[2]); v[3] = fma(zv, zv, v[3]);
        }
        __syncthreads();
        {
            const double* const As[3] = {X, X, X}; const double* const Bs[3] = {A, B, C}; double* const Gs[3] = {GA, GB, GC};
            grams<PT, 3>(As, Bs, Gs, P.n, P.p);
        }
        __syncthreads();
        for (int oo = threadIdx.x; oo < pp; oo += ST_THREADS) {
            const int o = (oo / P.p) * ld + (oo % P.p);
            v[0] = fma(-0.5 * GA[o], GB[o], v[0]); v[1] = fma(-0.5 * GA[o], GA[o], v[1]);
            v[2] = fma(-0.5 * GC[o], GA[o], v[2]); v[3] = fma(-0.5 * GC[o], GC[o], v[3]);
        }
        __syncthreads();
    }
    double o4[4];
    block_sum<4>(v, red, o4);
    if (threadIdx.x == 0) {
        const double dHd = o4[0], dsq = o4[1], zd = o4[2], zz = o4[3];
        const double radius = P.radius[inst], rsq = P.rsq[inst];
        double t = 0.0;
        int fl = 0;  // 0 = interior CG update, 1 = move to boundary and stop, 2 = failed
        if (dsq != 0.0) {  // _move_to_boundary (:137-152); dsq == 0 leaves z unchanged
            const double pq = zd / dsq, q = (zz - radius * radius) / dsq;
            const double disc = pq * pq - q;
            if (disc < 0.0) { fl = 2; atomicOr(&P.status[inst], 4); }      // solve_quadratic_equation raises (:159-160)
            else if (pq == 0.0) t = sqrt(-q);
            else {
                const double x1 = -(pq + (pq > 0.0 ? 1.0 : -1.0) * sqrt(disc));
                const double x2 = q / x1;
                t = fmax(x1, x2);
            }
        }
        const double alpha = rsq / dHd;
        if (fl == 0 && (dHd <= 0.0 || alpha > t)) fl = 1;
        if (fl == 0 && !(alpha == alpha)) { fl = 2; atomicOr(&P.status[inst], 1); }
        sc[0] = t; sc[1] = alpha;
        flag = fl;
        P.n_cg[inst] += 1;
    }
    __syncthreads();
    const int fl = flag;
    if (fl == 2) { if (threadIdx.x == 0) P.tcg_active[inst] = 0; return; }
    if (fl == 1) {  // z += t d ; on_boundary                                    (:120-122)
        const double t = sc[0];
        for (int l = 0; l < P.m; ++l) {
            const long long base = ((long long)inst * P.m + l) * np;
            for (int i = threadIdx.x; i < np; i += ST_THREADS) {
                P.z[base + i] = fma(t, P.d[base + i], P.z[base + i]);
                P.Hz[base + i] = fma(t, P.Hd[base + i], P.Hz[base + i]);
            }
        }
        if (threadIdx.x == 0) { P.tcg_active[inst] = 0; P.on_boundary[inst] = 1; }
        return;
    }
    // ---- pass B: r += alpha Hd ; z += alpha d ; rsq_next                        (:124-126)
    const double alpha = sc[1];
    double w[1] = {0.0};
    for (int l = 0; l < P.m; ++l) {
        const long long base = ((long long)inst * P.m + l) * np;
        for (int i = threadIdx.x; i < np; i += ST_THREADS) {
            const double hv = P.Hd[base + i];
            const double rv = fma(alpha, hv, P.r[base + i]);
            P.r[base + i] = rv;
            P.Hz[base + i] = fma(alpha, hv, P.Hz[base + i]);
            P.z[base + i] = fma(alpha, P.d[base + i], P.z[base + i]);
            const int o = (i / P.p) * ld + (i % P.p);
            X[o] = P.x[base + i]; C[o] = rv;
            w[0] = fma(rv, rv, w[0]);
        }
        __syncthreads();
        {
            const double* const As[1] = {X}; const double* const Bs[1] = {C}; double* const Gs[1] = {GC};
            grams<PT, 1>(As, Bs, Gs, P.n, P.p);
        }
        __syncthreads();
        for (int oo = threadIdx.x; oo < pp; oo += ST_THREADS) { const double q = GC[(oo / P.p) * ld + (oo % P.p)]; w[0] = fma(-0.5 * q, q, w[0]); }
        __syncthreads();
    }
    double o1[1];
    block_sum<1>(w, red, o1);
    if (threadIdx.x == 0) {
        const double rsq_next = o1[0];
        int stop = 0;
        if (sqrt(rsq_next) <= P.stoptol[inst]) stop = 1;                          // (:127-129)
        if (P.n_cg[inst] >= P.maxiter) stop = 1;                                  // (:115, 133-134)
        sc[2] = rsq_next / P.rsq[inst];                                           // beta (:130)
        P.rsq[inst] = rsq_next;
        flag = stop;
        if (stop) P.tcg_active[inst] = 0; else atomicAdd(P.n_active, 1);
    }
    __syncthreads();
    if (flag) return;
    const double beta = sc[2];
    for (int l = 0; l < P.m; ++l) {                                               // d = -r + beta d (:131)
        const long long base = ((long long)inst * P.m + l) * np;
        for (int i = threadIdx.x; i < np; i += ST_THREADS) P.d[base + i] = fma(beta, P.d[base + i], -P.r[base + i]);
    }
}

// ---- streaming versions for p = 16 (full-sites model at N = 4) -----------------------------------------------------------
// Same arithmetic as tcg_init_kernel / tcg_step_kernel; the Grams x^T v come from DMMA fragments loaded straight from global
// memory (scheme of riem_stream_kernel: warp = (row group ks of 4, column half jh of 2), partial Grams summed in a fixed order),
// so no n x p array is staged: 24 KB of shared memory instead of 164 KB, several CTAs per SM.  The B-fragment element of a warp
// lane, (row 4 s + t, column 8 jh + g), visits every entry of an n x 16 array exactly once over the CTA: the Frobenius parts of the
// inner products and the element-wise updates ride on those loads.
template <int NQ>
__device__ __forceinline__ void rs_store_partials(double* part, const double (&acc)[NQ][2][2], int ks, int jh, int g, int t) {
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            double* o = part + ((ks * NQ + q) * 16 + i * 8 + g) * 16 + jh * 8 + 2 * t;
            o[0] = acc[q][i][0]; o[1] = acc[q][i][1];
        }
}
template <int NQ>
__device__ __forceinline__ double rs_gram_entry(const double* part, int q, int entry) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) s += part[(k * NQ + q) * 256 + entry];
    return s;
}
template <int NV>
__device__ __forceinline__ void rs_block_sum(double (&v)[NV], double* red /* [NV][RS_THREADS / 32] */, double (&out)[NV]) {
    constexpr int NW = RS_THREADS / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const double s = warp_sum(v[i]);
        if (lane == 0) red[i * NW + warp] = s;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double s = 0.0;
        for (int w = 0; w < NW; ++w) s += red[i * NW + w];
        out[i] = s;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(RS_THREADS) tcg_init_stream_kernel(const TcgParams P) {
    __shared__ double part[4 * 256];
    __shared__ double red[RS_THREADS / 32];
    const int inst = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int ks = warp & 3, jh = warp >> 2, np = P.n * 16, ksteps = P.n >> 2;
    double v[1] = {0.0};
    for (int l = 0; l < P.m; ++l) {
        const long long base = ((long long)inst * P.m + l) * np;
        const double* __restrict__ xg = P.x + base;
        double acc[1][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}};
#pragma unroll 4
        for (int s = ks; s < ksteps; s += 4) {
            const int ro = (s * 4 + t) * 16;
            const long long e = base + ro + jh * 8 + g;
            const double xa0 = xg[ro + g], xa1 = xg[ro + 8 + g];
            const double gv = P.g[e];
            P.r[e] = gv; P.z[e] = 0.0; P.d[e] = -gv; P.Hz[e] = 0.0;
            v[0] = fma(gv, gv, v[0]);
            dmma884(acc[0][0][0], acc[0][0][1], xa0, gv);
            dmma884(acc[0][1][0], acc[0][1][1], xa1, gv);
        }
        rs_store_partials<1>(part, acc, ks, jh, g, t);
        __syncthreads();
        { const double q = rs_gram_entry<1>(part, 0, tid); v[0] = fma(-0.5 * q, q, v[0]); }
        __syncthreads();
    }
    double out[1];
    rs_block_sum<1>(v, red, out);
    if (tid == 0) {
        P.rsq[inst] = out[0];
        P.stoptol[inst] = fmax(P.abstol, P.reltol * sqrt(out[0]));
        P.tcg_active[inst] = 1;
        P.on_boundary[inst] = 0;
        P.n_cg[inst] = 0;
    }
}

__global__ void __launch_bounds__(RS_THREADS) tcg_step_stream_kernel(const TcgParams P) {
    __shared__ double part[4 * 3 * 256];
    __shared__ double red[4 * (RS_THREADS / 32)];
    __shared__ double sc[4];
    __shared__ int flag;
    const int inst = blockIdx.x;
    if (P.tcg_active[inst] == 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    if (tid == 0) atomicAdd(P.n_active + 1, 1);   // Hessian-vector products consumed (one per active instance and step)
    const int ks = warp & 3, jh = warp >> 2, np = P.n * 16, ksteps = P.n >> 2;
    // ---- pass A: dHd, dsq, zd, zz
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    for (int l = 0; l < P.m; ++l) {
        const long long base = ((long long)inst * P.m + l) * np;
        const double* __restrict__ xg = P.x + base;
        double acc[3][2][2];
#pragma unroll
        for (int q = 0; q < 3; ++q) { acc[q][0][0] = acc[q][0][1] = acc[q][1][0] = acc[q][1][1] = 0.0; }
#pragma unroll 4
        for (int s = ks; s < ksteps; s += 4) {
            const int ro = (s * 4 + t) * 16;
            const long long e = base + ro + jh * 8 + g;
            const double xa0 = xg[ro + g], xa1 = xg[ro + 8 + g];
            const double dv = P.d[e], hv = P.Hd[e], zv = P.z[e];
            v[0] = fma(dv, hv, v[0]); v[1] = fma(dv, dv, v[1]); v[2] = fma(zv, dv, v[2]); v[3] = fma(zv, zv, v[3]);
            dmma884(acc[0][0][0], acc[0][0][1], xa0, dv); dmma884(acc[0][1][0], acc[0][1][1], xa1, dv);
            dmma884(acc[1][0][0], acc[1][0][1], xa0, hv); dmma884(acc[1][1][0], acc[1][1][1], xa1, hv);
            dmma884(acc[2][0][0], acc[2][0][1], xa0, zv); dmma884(acc[2][1][0], acc[2][1][1], xa1, zv);
        }
        rs_store_partials<3>(part, acc, ks, jh, g, t);
        __syncthreads();
        {
            const double ga = rs_gram_entry<3>(part, 0, tid), gb = rs_gram_entry<3>(part, 1, tid), gc = rs_gram_entry<3>(part, 2, tid);
            v[0] = fma(-0.5 * ga, gb, v[0]); v[1] = fma(-0.5 * ga, ga, v[1]);
            v[2] = fma(-0.5 * gc, ga, v[2]); v[3] = fma(-0.5 * gc, gc, v[3]);
        }
        __syncthreads();
    }
    double o4[4];
    rs_block_sum<4>(v, red, o4);
    if (tid == 0) {
        const double dHd = o4[0], dsq = o4[1], zd = o4[2], zz = o4[3];
        const double radius = P.radius[inst], rsq = P.rsq[inst];
        double tt = 0.0;
        int fl = 0;  // 0 = interior CG update, 1 = move to boundary and stop, 2 = failed
        if (dsq != 0.0) {  // _move_to_boundary (:137-152); dsq == 0 leaves z unchanged
            const double pq = zd / dsq, q = (zz - radius * radius) / dsq;
            const double disc = pq * pq - q;
            if (disc < 0.0) { fl = 2; atomicOr(&P.status[inst], 4); }      // solve_quadratic_equation raises (:159-160)
            else if (pq == 0.0) tt = sqrt(-q);
            else {
                const double x1 = -(pq + (pq > 0.0 ? 1.0 : -1.0) * sqrt(disc));
                const double x2 = q / x1;
                tt = fmax(x1, x2);
            }
        }
        const double alpha = rsq / dHd;
        if (fl == 0 && (dHd <= 0.0 || alpha > tt)) fl = 1;
        if (fl == 0 && !(alpha == alpha)) { fl = 2; atomicOr(&P.status[inst], 1); }
        sc[0] = tt; sc[1] = alpha;
        flag = fl;
        P.n_cg[inst] += 1;
    }
    __syncthreads();
    const int fl = flag;
    if (fl == 2) { if (tid == 0) P.tcg_active[inst] = 0; return; }
    const long long ibase = (long long)inst * P.m * np;
    const int tot = P.m * np;
    if (fl == 1) {  // z += t d ; on_boundary                                    (:120-122)
        const double tt = sc[0];
        for (int i = tid; i < tot; i += RS_THREADS) {
            P.z[ibase + i] = fma(tt, P.d[ibase + i], P.z[ibase + i]);
            P.Hz[ibase + i] = fma(tt, P.Hd[ibase + i], P.Hz[ibase + i]);
        }
        if (tid == 0) { P.tcg_active[inst] = 0; P.on_boundary[inst] = 1; }
        return;
    }
    // ---- pass B: r += alpha Hd ; z += alpha d ; rsq_next                        (:124-126)
    const double alpha = sc[1];
    double w[1] = {0.0};
    for (int l = 0; l < P.m; ++l) {
        const long long base = ((long long)inst * P.m + l) * np;
        const double* __restrict__ xg = P.x + base;
        double acc[1][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}};
#pragma unroll 4
        for (int s = ks; s < ksteps; s += 4) {
            const int ro = (s * 4 + t) * 16;
            const long long e = base + ro + jh * 8 + g;
            const double xa0 = xg[ro + g], xa1 = xg[ro + 8 + g];
            const double hv = P.Hd[e];
            const double rv = fma(alpha, hv, P.r[e]);
            P.r[e] = rv;
            P.Hz[e] = fma(alpha, hv, P.Hz[e]);
            P.z[e] = fma(alpha, P.d[e], P.z[e]);
            w[0] = fma(rv, rv, w[0]);
            dmma884(acc[0][0][0], acc[0][0][1], xa0, rv);
            dmma884(acc[0][1][0], acc[0][1][1], xa1, rv);
        }
        rs_store_partials<1>(part, acc, ks, jh, g, t);
        __syncthreads();
        { const double q = rs_gram_entry<1>(part, 0, tid); w[0] = fma(-0.5 * q, q, w[0]); }
        __syncthreads();
    }
    double o1[1];
    rs_block_sum<1>(w, red, o1);
    if (tid == 0) {
        const double rsq_next = o1[0];
        int stop = 0;
        if (sqrt(rsq_next) <= P.stoptol[inst]) stop = 1;                          // (:127-129)
        if (P.n_cg[inst] >= P.maxiter) stop = 1;                                  // (:115, 133-134)
        sc[2] = rsq_next / P.rsq[inst];                                           // beta (:130)
        P.rsq[inst] = rsq_next;
        flag = stop;
        if (stop) P.tcg_active[inst] = 0; else atomicAdd(P.n_active, 1);
    }
    __syncthreads();
    if (flag) return;
    const double beta = sc[2];
    for (int i = tid; i < tot; i += RS_THREADS) P.d[ibase + i] = fma(beta, P.d[ibase + i], -P.r[ibase + i]);   // d = -r + beta d (:131)
}

struct TrUpdateParams {
    double *radius, *rho;
    const double *fx, *fn, *gz, *zHz;
    const int* on_boundary;
    int *accepted, *status;
    int* rejected;
    double rho_trust, maxradius;
    int batch;
    // history rows for this iteration (nullable)
    double *h_rho, *h_radius; int *h_ncg, *h_onb, *h_acc; const int* n_cg;
};

// rho, radius update, acceptance                                                  trust_region_rcopt.py:67-77
__global__ void tr_update_kernel(const TrUpdateParams P) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= P.batch) return;
    const double rho = (P.fn[b] - P.fx[b]) / (P.gz[b] + 0.5 * P.zHz[b]);
    double radius = P.radius[b];
    if (rho < 0.25) radius *= 0.25;
    else if (rho > 0.75 && P.on_boundary[b]) radius = fmin(2.0 * radius, P.maxradius);
    const int acc = rho > P.rho_trust ? 1 : 0;
    if (!(rho == rho)) atomicOr(&P.status[b], 1);
    P.radius[b] = radius; P.rho[b] = rho; P.accepted[b] = acc; P.rejected[b] = 1 - acc;
    if (P.h_rho) P.h_rho[b] = rho;
    if (P.h_radius) P.h_radius[b] = radius;
    if (P.h_ncg) P.h_ncg[b] = P.n_cg[b];
    if (P.h_onb) P.h_onb[b] = P.on_boundary[b];
    if (P.h_acc) P.h_acc[b] = acc;
}

// x <- x_next where accepted
__global__ void select_copy_kernel(double* __restrict__ x, const double* __restrict__ xn, const int* __restrict__ accepted, long long per) {
    const int b = blockIdx.y;
    if (!accepted[b]) return;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per; i += (long long)gridDim.x * blockDim.x)
        x[b * per + i] = xn[b * per + i];
}

__global__ void fill_kernel(double* a, double v, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = v;
}

}  // namespace otn
