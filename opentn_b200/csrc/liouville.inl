// Liouvillian assembly on the device (SURVEY 8f "next" #1): the sums of vectorised dissipators the reference builds with
// np.kron on D x D complex matrices (transformations.py:243-250 vectorize_dissipative, :252-272 op2fullspace /
// dissipative2liouvillian_full, :324-369 odd / even / pbc sums).  One thread per output entry, no Kronecker product is ever
// materialised:  with L_full = L placed on the sites (s0, s1) of an N-site register (identity elsewhere),
//   D[L_full][(a,b),(c,d)] = L_full[a,c] conj(L_full[b,d]) - (L^+L)_full[a,c] delta_bd / 2 - delta_ac (L^+L)_full[d,b] / 2.
// HBM-bound: 16 bytes written per entry, nothing read but the q x q jump operators (q = d^2).

namespace {

constexpr int LV_MAX_TERMS = 64;
struct LvTerms {
    int n_terms, N, d, q;                  // q = d^2
    int op[LV_MAX_TERMS];                  // jump operator of the term
    int w0[LV_MAX_TERMS], w1[LV_MAX_TERMS];  // d^(N-1-s0), d^(N-1-s1): weights of the two sites in a basis index
};

// LdL[j] = L_j^+ L_j (q x q complex), one block per jump operator
__global__ void lv_gram_kernel(const double* __restrict__ Lre, const double* __restrict__ Lim, int q, double* __restrict__ Gre,
                               double* __restrict__ Gim) {
    const double* lr = Lre + (size_t)blockIdx.x * q * q;
    const double* li = Lim ? Lim + (size_t)blockIdx.x * q * q : nullptr;
    for (int e = threadIdx.x; e < q * q; e += blockDim.x) {
        const int i = e / q, j = e % q;
        double sr = 0.0, si = 0.0;
        for (int k = 0; k < q; ++k) {      // conj(L[k,i]) L[k,j]
            const double ar = lr[k * q + i], ai = li ? -li[k * q + i] : 0.0, br = lr[k * q + j], bi = li ? li[k * q + j] : 0.0;
            sr += ar * br - ai * bi;
            si += ar * bi + ai * br;
        }
        Gre[(size_t)blockIdx.x * q * q + e] = sr;
        Gim[(size_t)blockIdx.x * q * q + e] = si;
    }
}

__global__ void lv_assemble_kernel(const LvTerms t, const double* __restrict__ Lre, const double* __restrict__ Lim,
                                   const double* __restrict__ Gre, const double* __restrict__ Gim, int dim,
                                   double* __restrict__ out_re, double* __restrict__ out_im) {
    const long long D = (long long)dim * dim, total = D * D;
    const int d = t.d, q = t.q;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long row = e / D, col = e - row * D;
        const int a = (int)(row / dim), b = (int)(row % dim), c = (int)(col / dim), dd = (int)(col % dim);
        double sr = 0.0, si = 0.0;
        for (int k = 0; k < t.n_terms; ++k) {
            const int w0 = t.w0[k], w1 = t.w1[k];
            const int a0 = (a / w0) % d, a1 = (a / w1) % d, c0 = (c / w0) % d, c1 = (c / w1) % d;
            const int b0 = (b / w0) % d, b1 = (b / w1) % d, d0 = (dd / w0) % d, d1 = (dd / w1) % d;
            // indices with the two sites removed must agree for the embedded operator to be non-zero
            const bool ac_rest = (a - a0 * w0 - a1 * w1) == (c - c0 * w0 - c1 * w1);
            const bool bd_rest = (b - b0 * w0 - b1 * w1) == (dd - d0 * w0 - d1 * w1);
            const int ia = a0 * d + a1, ic = c0 * d + c1, ib = b0 * d + b1, id = d0 * d + d1;
            const double* lr = Lre + (size_t)t.op[k] * q * q;
            const double* li = Lim ? Lim + (size_t)t.op[k] * q * q : nullptr;
            const double* gr = Gre + (size_t)t.op[k] * q * q;
            const double* gi = Gim + (size_t)t.op[k] * q * q;
            if (ac_rest && bd_rest) {      // L[a,c] conj(L[b,d])
                const double xr = lr[ia * q + ic], xi = li ? li[ia * q + ic] : 0.0;
                const double yr = lr[ib * q + id], yi = li ? -li[ib * q + id] : 0.0;
                sr += xr * yr - xi * yi;
                si += xr * yi + xi * yr;
            }
            if (ac_rest && b == dd) { sr -= 0.5 * gr[ia * q + ic]; si -= 0.5 * gi[ia * q + ic]; }
            if (bd_rest && a == c) { sr -= 0.5 * gr[id * q + ib]; si -= 0.5 * gi[id * q + ib]; }
        }
        out_re[e] = sr;
        out_im[e] = si;
    }
}

}  // namespace

extern "C" int otn_liouvillian(const double* L_re, const double* L_im, int n_ops, int N, int d, const int* term_op,
                               const int* term_site0, const int* term_site1, int n_terms, double* out_re, double* out_im,
                               int device) {
    if (!L_re || !out_re || !out_im || !term_op || !term_site0 || !term_site1) return fail("null argument");
    if (n_ops < 1 || N < 2 || d < 2 || n_terms < 1) return fail("bad argument");
    if (n_terms > LV_MAX_TERMS) return fail("otn_liouvillian: at most 64 terms per call");
    double dimd = 1.0;
    for (int i = 0; i < N; ++i) dimd *= d;
    if (dimd > 4096.0) return fail("otn_liouvillian: d^N must not exceed 4096");
    const int dim = (int)dimd, q = d * d;
    LvTerms t{};
    t.n_terms = n_terms; t.N = N; t.d = d; t.q = q;
    for (int k = 0; k < n_terms; ++k) {
        const int s0 = term_site0[k], s1 = term_site1[k];
        if (term_op[k] < 0 || term_op[k] >= n_ops || s0 < 0 || s0 >= N || s1 < 0 || s1 >= N || s0 == s1)
            return fail("otn_liouvillian: bad term");
        t.op[k] = term_op[k];
        int w0 = 1, w1 = 1;
        for (int i = 0; i < N - 1 - s0; ++i) w0 *= d;
        for (int i = 0; i < N - 1 - s1; ++i) w1 *= d;
        t.w0[k] = w0; t.w1[k] = w1;
    }
    OTN_TRY(use_device(device));
    DeviceGuard dev_guard__(device);
    Scratch& S = g_scratch[device];                 // per device (see otn_expm_batch)
    DevBuf &in_re = S.a, &in_im = S.b, &gram = S.c, &o_re = S.o, &o_im = S.x2;
    const size_t lbytes = (size_t)n_ops * q * q * 8, obytes = (size_t)dim * dim * dim * dim * 8;
    const void *lre, *lim; void *ore, *oim;
    OTN_TRY(stage_in(L_re, lbytes, in_re, 0, &lre));
    OTN_TRY(stage_in(L_im, lbytes, in_im, 0, &lim));
    OTN_TRY(stage_out(out_re, obytes, o_re, &ore));
    OTN_TRY(stage_out(out_im, obytes, o_im, &oim));
    OTN_TRY(gram.ensure(2 * lbytes));
    double* gre = gram.as<double>();
    double* gim = gre + (size_t)n_ops * q * q;
    lv_gram_kernel<<<n_ops, 64>>>((const double*)lre, (const double*)lim, q, gre, gim);
    ++g_launches;
    const long long total = (long long)dim * dim * dim * dim;
    const int blocks = (int)std::min<long long>((total + 255) / 256, 148LL * 16);
    lv_assemble_kernel<<<blocks, 256>>>(t, (const double*)lre, (const double*)lim, gre, gim, dim, (double*)ore, (double*)oim);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    OTN_TRY(finish_out(out_re, ore, obytes, 0));
    OTN_TRY(finish_out(out_im, oim, obytes, 0));
    CUDA_TRY(cudaDeviceSynchronize());
    return 0;
}
