// Initial-point factory on the device (SURVEY 8f "next" #2) for the 2-site model:
//   super2ortho = super2choi (transformations.py:68-94) -> rank-truncated PSD factor (factorize_psd_truncated :793-802, an SVD of
//   the symmetric Choi matrix, i.e. its eigen-decomposition) -> choi2ortho (:701-714), batched over superoperators.
// One warp per superoperator: the Choi matrix (side dim^2 <= 16) is diagonalised by the parallel cyclic Jacobi of stiefel.cuh.

namespace {

__global__ void __launch_bounds__(32) super2ortho_kernel(const double* __restrict__ superop, int dim, int K, double* __restrict__ x) {
    const int q = dim * dim, lane = threadIdx.x;
    __shared__ double C[ST_MAXP * ST_MAXP], V[ST_MAXP * ST_MAXP], cs[2 * ST_MAXP], w[ST_MAXP];
    __shared__ int pairs[2 * ST_MAXP], order[ST_MAXP];
    const double* S = superop + (size_t)blockIdx.x * q * q;
    // Choi[(a,c),(b,e)] = S[(a,b),(c,e)], symmetrised against rounding
    for (int i = lane; i < q * q; i += 32) {
        const int r = i / q, c = i % q;
        const int a = r / dim, cc = r % dim, b = c / dim, e = c % dim;
        const double v1 = S[(a * dim + b) * q + (cc * dim + e)];
        const double v2 = S[(b * dim + a) * q + (e * dim + cc)];   // transposed Choi entry [(b,e),(a,c)]
        C[i] = 0.5 * (v1 + v2);
    }
    __syncwarp();
    jacobi_eig_warp(C, V, q, cs, pairs);
    __syncwarp();
    if (lane < q) w[lane] = C[lane * q + lane];
    __syncwarp();
    // rank of each eigenvalue by decreasing magnitude (singular values of the SVD route); ties broken by index
    if (lane < q) {
        int rk = 0;
        const double me = fabs(w[lane]);
        for (int j = 0; j < q; ++j) {
            const double o = fabs(w[j]);
            if (o > me || (o == me && j < lane)) ++rk;
        }
        order[rk] = lane;
    }
    __syncwarp();
    // x[(a,k), b] = X[(a,b), k],  X[:, k] = v_k sqrt|w_k|, sign fixed so that the largest component of v_k is positive
    double* xo = x + (size_t)blockIdx.x * dim * K * dim;
    for (int k = 0; k < K; ++k) {
        const int src = order[k];
        double best = 0.0;
        for (int r = 0; r < q; ++r) { const double v = V[r * q + src]; if (fabs(v) > fabs(best)) best = v; }
        const double scale = (best < 0.0 ? -1.0 : 1.0) * sqrt(fabs(w[src]));
        for (int r = lane; r < q; r += 32) {
            const int a = r / dim, b = r % dim;
            xo[(a * K + k) * dim + b] = V[r * q + src] * scale;
        }
    }
}

// ---- the same factory for LARGE Choi matrices (full-sites ansatz: side dim^2 = 256 at N = 4) ---------------------------------------
// Only the K dominant eigenpairs of the PSD Choi matrix C are needed (K <= 16), so instead of a full 256 x 256 eigen-decomposition:
// block power iteration with a Rayleigh-Ritz step per iteration.  One CTA per superoperator; C stays in global memory (512 KB, L2),
// the K iterate vectors live in shared memory:
//     W = C V;   G = W^T W = V^T C^2 V;   G = U diag(L) U^T (one-warp Jacobi);   V <- W U L^{-1/2}
// V converges to the dominant eigenvectors of C with L -> lambda^2; directions whose L vanishes (rank(C) < K: the reference's SVD
// returns rounding-level singular values there) become zero columns, i.e. zero Kraus operators -- ortho2super(x) is unchanged.
// Stops when every L moved by less than 1e-14 L_max and every Ritz residual |C v - theta v| is below 2e-14 theta_max, or after
// `max_iter` iterations (slowly separating spectra: the error then decays like (lambda_{K+1} / lambda_K)^iterations).
__global__ void choi_from_superop_kernel(const double* __restrict__ superop, int dim, double* __restrict__ C) {
    const int q = dim * dim;
    const double* S = superop + (size_t)blockIdx.x * q * q;
    double* Co = C + (size_t)blockIdx.x * q * q;
    for (int i = threadIdx.x; i < q * q; i += blockDim.x) {
        const int r = i / q, c = i % q;
        const int a = r / dim, cc = r % dim, b = c / dim, e = c % dim;
        Co[i] = 0.5 * (S[(size_t)(a * dim + b) * q + (cc * dim + e)] + S[(size_t)(b * dim + a) * q + (e * dim + cc)]);
    }
}

__global__ void __launch_bounds__(1024) super2ortho_big_kernel(const double* __restrict__ Cg, int dim, int K, int max_iter,
                                                               double* __restrict__ x, int* __restrict__ iters_out) {
    extern __shared__ __align__(16) double sm[];
    const int q = dim * dim, tid = threadIdx.x;
    double* V = sm;                     // [q][K]
    double* W = V + q * K;              // [q][K]
    __shared__ double G[ST_MAXP * ST_MAXP], U[ST_MAXP * ST_MAXP], Lcur[ST_MAXP], Lold[ST_MAXP], cs[2 * ST_MAXP];
    __shared__ int pairs[2 * ST_MAXP], order[ST_MAXP], done;
    __shared__ double resid[ST_MAXP], theta[ST_MAXP];
    const double* C = Cg + (size_t)blockIdx.x * q * q;
    // deterministic start: a fixed pseudo-random matrix (same for every call)
    for (int i = tid; i < q * K; i += blockDim.x) {
        unsigned h = (unsigned)i * 2654435761u + 12345u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
        V[i] = (double)h * (1.0 / 4294967296.0) - 0.5;
    }
    if (tid < ST_MAXP) Lold[tid] = 0.0;
    if (tid == 0) done = 0;
    __syncthreads();
    int it = 0;
    for (; it < max_iter; ++it) {
        // W = C V  (C symmetric: column r of C read as row entries C[i][r], coalesced over r)
        for (int o = tid; o < q * 4; o += blockDim.x) {
            const int r = o % q, jg = o / q;                    // 4 column groups of K/4 (K <= 16 -> at most 4 columns each)
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
            const int j0 = jg * ((K + 3) / 4);
            for (int i = 0; i < q; ++i) {
                const double c = C[(size_t)i * q + r];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) { const int j = j0 + jj; if (jj < (K + 3) / 4 && j < K) acc[jj] = fma(c, V[i * K + j], acc[jj]); }
            }
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) { const int j = j0 + jj; if (jj < (K + 3) / 4 && j < K) W[r * K + j] = acc[jj]; }
        }
        __syncthreads();
        // residual of the current Ritz pairs: theta_j = v_j . C v_j,  res_j = |C v_j - theta_j v_j|  (columns of V are orthonormal
        // and Rayleigh-Ritz rotated by the previous iteration); eigenvalue movement alone stops too early, it is quadratic in the
        // subspace error
        if (tid < K) {
            double th = 0.0;
            for (int r = 0; r < q; ++r) th = fma(V[r * K + tid], W[r * K + tid], th);
            double rs = 0.0;
            for (int r = 0; r < q; ++r) { const double e = W[r * K + tid] - th * V[r * K + tid]; rs = fma(e, e, rs); }
            resid[tid] = sqrt(rs); theta[tid] = fabs(th);
        }
        for (int o = tid; o < K * K; o += blockDim.x) {          // G = W^T W
            const int i = o / K, j = o % K;
            double s = 0.0;
            for (int r = 0; r < q; ++r) s = fma(W[r * K + i], W[r * K + j], s);
            G[o] = s;
        }
        __syncthreads();
        if (tid < 32) jacobi_eig_warp(G, U, K, cs, pairs);
        __syncthreads();
        if (tid == 0) {
            double lmax = 0.0, move = 0.0;
            for (int k = 0; k < K; ++k) { Lcur[k] = G[k * K + k]; lmax = fmax(lmax, Lcur[k]); }
            // eigenvalue order is not stable between iterations: compare sorted spectra
            double a[ST_MAXP], b[ST_MAXP];
            for (int k = 0; k < K; ++k) { a[k] = Lcur[k]; b[k] = Lold[k]; }
            for (int i = 1; i < K; ++i) { double v = a[i]; int j = i - 1; while (j >= 0 && a[j] < v) { a[j + 1] = a[j]; --j; } a[j + 1] = v; }
            for (int i = 1; i < K; ++i) { double v = b[i]; int j = i - 1; while (j >= 0 && b[j] < v) { b[j + 1] = b[j]; --j; } b[j + 1] = v; }
            for (int k = 0; k < K; ++k) move = fmax(move, fabs(a[k] - b[k]));
            for (int k = 0; k < K; ++k) Lold[k] = Lcur[k];
            double rmax = 0.0, tmax = 0.0;
            for (int k = 0; k < K; ++k) { rmax = fmax(rmax, resid[k]); tmax = fmax(tmax, theta[k]); }
            done = (it > 2 && move <= 1e-14 * lmax && rmax <= 2e-14 * tmax) ? 1 : 0;
            G[0] = lmax;                                           // (diagonal already copied)
        }
        __syncthreads();
        const double lmax = G[0];
        for (int o = tid; o < q * K; o += blockDim.x) {           // V = W U L^{-1/2}; vanished directions -> 0
            const int r = o / K, j = o % K;
            double s = 0.0;
            if (Lcur[j] > 1e-28 * lmax) {
                for (int i = 0; i < K; ++i) s = fma(W[r * K + i], U[i * K + j], s);
                s *= rsqrt(Lcur[j]);
            }
            V[o] = s;
        }
        __syncthreads();
        if (done) break;
    }
    if (tid == 0 && iters_out) iters_out[blockIdx.x] = it;
    // one more exact normalisation pass is not needed: V^T V = 1 by construction (up to rounding of the last update).
    // eigenvalues of C: lambda = sqrt(L); order by decreasing lambda, sign convention as in super2ortho_kernel
    if (tid < K) {
        int rk = 0;
        for (int j = 0; j < K; ++j) if (Lcur[j] > Lcur[tid] || (Lcur[j] == Lcur[tid] && j < tid)) ++rk;
        order[rk] = tid;
    }
    __syncthreads();
    double* xo = x + (size_t)blockIdx.x * dim * K * dim;
    for (int k = 0; k < K; ++k) {
        const int src = order[k];
        __shared__ double bestv;
        if (tid == 0) {
            double best = 0.0;
            for (int r = 0; r < q; ++r) { const double v = V[r * K + src]; if (fabs(v) > fabs(best)) best = v; }
            bestv = best;
        }
        __syncthreads();
        const double scale = (bestv < 0.0 ? -1.0 : 1.0) * sqrt(sqrt(fmax(Lcur[src], 0.0)));
        for (int r = tid; r < q; r += blockDim.x) {
            const int a = r / dim, b = r % dim;
            xo[(a * K + k) * dim + b] = V[r * K + src] * scale;
        }
        __syncthreads();
    }
}

}  // namespace

extern "C" int otn_super2ortho(const double* superop, int count, int dim, int K, double* x, int device) {
    if (!superop || !x || count < 1 || dim < 1 || K < 1) return fail("bad argument");
    const int q = dim * dim;
    if (K > q) return fail("otn_super2ortho: rank %d exceeds the Choi side %d", K, q);
    if (q > ST_MAXP && (K > ST_MAXP || q > 1024)) return fail("otn_super2ortho: for Choi side %d (> %d) the rank must be <= %d and the side <= 1024", q, ST_MAXP, ST_MAXP);
    OTN_TRY(use_device(device));
    DeviceGuard dev_guard__(device);
    Scratch& S = g_scratch[device];
    const void* sd; void* xd;
    OTN_TRY(stage_in(superop, (size_t)count * q * q * 8, S.a, 0, &sd));
    OTN_TRY(stage_out(x, (size_t)count * dim * K * dim * 8, S.o, &xd));
    if (q > ST_MAXP) {           // full-sites Choi matrices: block power iteration for the K dominant eigenpairs
        OTN_TRY(S.b.ensure((size_t)count * q * q * 8));
        choi_from_superop_kernel<<<count, 1024>>>((const double*)sd, dim, S.b.as<double>());
        const size_t sm = (size_t)2 * q * K * 8;
        OTN_TRY(set_smem(super2ortho_big_kernel, sm));
        super2ortho_big_kernel<<<count, 1024, sm>>>(S.b.as<double>(), dim, K, 2000, (double*)xd, nullptr);
        g_launches += 2;
        CUDA_TRY(cudaGetLastError());
        OTN_TRY(finish_out(x, xd, (size_t)count * dim * K * dim * 8, 0));
        CUDA_TRY(cudaDeviceSynchronize());
        return 0;
    }
    super2ortho_kernel<<<count, 32>>>((const double*)sd, dim, K, (double*)xd);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    OTN_TRY(finish_out(x, xd, (size_t)count * dim * K * dim * 8, 0));
    CUDA_TRY(cudaDeviceSynchronize());
    return 0;
}
