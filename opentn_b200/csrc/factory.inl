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

}  // namespace

extern "C" int otn_super2ortho(const double* superop, int count, int dim, int K, double* x, int device) {
    if (!superop || !x || count < 1 || dim < 1 || K < 1) return fail("bad argument");
    const int q = dim * dim;
    if (q > ST_MAXP) return fail("otn_super2ortho: Choi side dim^2 = %d > %d (2-site model only; full-sites starts stay on the host)", q, ST_MAXP);
    if (K > q) return fail("otn_super2ortho: rank %d exceeds the Choi side %d", K, q);
    OTN_TRY(use_device(device));
    DeviceGuard dev_guard__(device);
    Scratch& S = g_scratch[device];
    const void* sd; void* xd;
    OTN_TRY(stage_in(superop, (size_t)count * q * q * 8, S.a, 0, &sd));
    OTN_TRY(stage_out(x, (size_t)count * dim * K * dim * 8, S.o, &xd));
    super2ortho_kernel<<<count, 32>>>((const double*)sd, dim, K, (double*)xd);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    OTN_TRY(finish_out(x, xd, (size_t)count * dim * K * dim * 8, 0));
    CUDA_TRY(cudaDeviceSynchronize());
    return 0;
}
