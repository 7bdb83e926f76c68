// exp(tau * A) on the device (SURVEY 8f "next" #1: the step immediately before the hot path; transformations.py:24-32 calls
// scipy/jax `expm`, which takes ~15-90 s for the 4096 x 4096 Liouvillian of N = 6).
//
// Scaling and squaring with a degree-16 Taylor polynomial evaluated by Paterson-Stockmeyer in X^4 (6 products) on the same FP64
// DMMA GEMM kernel as the chain:  X = tau A / 2^s  with  ||X||_1 <= 1/2  (truncation error 0.5^17/17! = 2e-20),
//   T16(X) = B0 + X^4 (B1 + X^4 (B2 + X^4 (B3 + c16 X^4))),   Bj = sum_{i<4} c_{4j+i} X^i,   then s squarings.
// Complex matrices are carried as (re, im) pairs; one complex product = two dual-product launches.

namespace {

__global__ void col_abs_sum_kernel(const double* __restrict__ re, const double* __restrict__ im, int D, double* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= D) return;
    double s = 0.0;
    for (int i = 0; i < D; ++i) {
        const double a = re[(long long)i * D + j], b = im ? im[(long long)i * D + j] : 0.0;
        s += im ? sqrt(a * a + b * b) : fabs(a);
    }
    out[j] = s;
}

// out_b = diag I + c1 X1_b + c2 X2_b + c3 X3_b + c4 X4_b for every instance b (null pointers skipped; `diag` = coefficient
// of the identity; `c1b`, when given, multiplies c1 per instance; s1 = instance stride of X1, `sb` of everything else)
__global__ void lincomb_kernel(double* __restrict__ out, long long n, int D, int count, long long sb, double diag, double c1,
                               const double* __restrict__ c1b, const double* __restrict__ X1, long long s1, double c2,
                               const double* __restrict__ X2, double c3, const double* __restrict__ X3, double c4,
                               const double* __restrict__ X4) {
    const long long total = n * count;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long b = t / n, i = t - b * n, o = b * sb + i;
        double v = ((i / D) == (i % D)) ? diag : 0.0;
        if (X1) v = fma(c1b ? c1 * c1b[b] : c1, X1[b * s1 + i], v);
        if (X2) v = fma(c2, X2[o], v);
        if (X3) v = fma(c3, X3[o], v);
        if (X4) v = fma(c4, X4[o], v);
        out[o] = v;
    }
}

struct CMat { double* re; double* im; };   // im == nullptr for a real matrix; instance b at re + b * stride

struct ExpmCtx {
    int D; long long n; bool cplx; int count; long long sb; cudaStream_t st; double* neg;   // neg: scratch for -Im
    int launch(const double* A1, const double* B1, const double* A2, const double* B2, double* C) {
        GemmParams p{};
        p.A1 = A1; p.B1 = B1; p.A2 = A2; p.B2 = B2; p.C = C; p.D = D; p.row0 = 0; p.rows = D; p.batch = count; p.epi = EPI_NONE;
        p.sA1 = p.sB1 = p.sA2 = p.sB2 = p.sC = sb;
        p.splitk = gemm_want_split(D);           // D >= 1024 (the N = 6 propagator): split-K CTA pairs
        ++g_launches;
        CUDA_TRY(gemm_launch(p, false, false, st));
        return 0;
    }
    // C = A * B
    int mul(CMat A, CMat B, CMat C) {
        if (!cplx) return launch(A.re, B.re, nullptr, nullptr, C.re);
        lincomb_kernel<<<1024, 256, 0, st>>>(neg, n, D, count, sb, 0.0, -1.0, nullptr, A.im, sb, 0.0, nullptr, 0.0, nullptr, 0.0, nullptr);
        CUDA_TRY(cudaGetLastError());
        OTN_TRY(launch(A.re, B.re, neg, B.im, C.re));     // Re = Ar Br - Ai Bi
        OTN_TRY(launch(A.re, B.im, A.im, B.re, C.im));    // Im = Ar Bi + Ai Br
        return 0;
    }
    // out = diag I + c1 X1 + c2 X2 + c3 X3 + c4 X4
    int comb(CMat out, double diag, double c1, CMat X1, double c2, CMat X2, double c3, CMat X3, double c4, CMat X4) {
        lincomb_kernel<<<1024, 256, 0, st>>>(out.re, n, D, count, sb, diag, c1, nullptr, X1.re, sb, c2, X2.re, c3, X3.re, c4, X4.re);
        if (cplx) lincomb_kernel<<<1024, 256, 0, st>>>(out.im, n, D, count, sb, 0.0, c1, nullptr, X1.im, sb, c2, X2.im, c3, X3.im, c4, X4.im);
        CUDA_TRY(cudaGetLastError());
        return 0;
    }
};

}  // namespace

extern "C" int otn_expm_batch(const double* A_re, const double* A_im, int D, const double* taus, int count, double* out_re,
                              double* out_im, int device) {
    if (!A_re || !out_re || !taus || D < 1 || count < 1) return fail("bad argument");
    if ((A_im != nullptr) != (out_im != nullptr)) return fail("A_im and out_im must be given together");
    OTN_TRY(use_device(device));
    DeviceGuard dev_guard__(device);
    const bool cplx = A_im != nullptr;
    const long long n = (long long)D * D;
    const size_t bytes = (size_t)n * 8;
    Scratch& S = g_scratch[device];                 // per device: a buffer cached for one GPU must never be handed to another
    DevBuf &in_re = S.a, &in_im = S.b, &ws = S.x1, &o_re = S.o, &o_im = S.x2;
    const void *are, *aim; void *ore, *oim;
    OTN_TRY(stage_in(A_re, bytes, in_re, 0, &are));
    OTN_TRY(stage_in(A_im, bytes, in_im, 0, &aim));
    OTN_TRY(stage_out(out_re, bytes * count, o_re, &ore));
    OTN_TRY(stage_out(out_im, bytes * count, o_im, &oim));
    // ||A||_1; one scaling power for the whole sweep, from the largest |tau| (over-scaling the smaller ones is harmless)
    const int nmat = 8;                                             // X, X2, X3, X4, B, P, Q, neg
    const long long sb = (cplx ? 2 : 1) * n;                        // instance stride inside one work matrix
    OTN_TRY(ws.ensure((size_t)nmat * count * sb * 8 + (size_t)(D + count) * 8));
    double* base = ws.as<double>();
    double* colsum = base + (size_t)nmat * count * sb;
    double* scales = colsum + D;
    col_abs_sum_kernel<<<(D + 127) / 128, 128>>>((const double*)are, (const double*)aim, D, colsum);
    CUDA_TRY(cudaGetLastError());
    std::vector<double> hs(D);
    CUDA_TRY(cudaMemcpy(hs.data(), colsum, (size_t)D * 8, cudaMemcpyDeviceToHost));
    double nrm = 0.0, tmax = 0.0;
    for (double v : hs) nrm = std::max(nrm, v);
    for (int b = 0; b < count; ++b) {
        if (!(taus[b] == taus[b]) || std::isinf(taus[b])) return fail("otn_expm: tau is not finite");
        tmax = std::max(tmax, std::fabs(taus[b]));
    }
    nrm *= tmax;
    if (!(nrm == nrm) || std::isinf(nrm)) return fail("otn_expm: matrix norm is not finite");
    int s = 0;
    while (std::ldexp(nrm, -s) > 0.5 && s < 60) ++s;
    std::vector<double> hscale(count);
    for (int b = 0; b < count; ++b) hscale[b] = std::ldexp(taus[b], -s);
    CUDA_TRY(cudaMemcpy(scales, hscale.data(), (size_t)count * 8, cudaMemcpyHostToDevice));

    auto mat = [&](int k) { CMat m; m.re = base + (size_t)k * count * sb; m.im = cplx ? m.re + n : nullptr; return m; };
    CMat X = mat(0), X2 = mat(1), X3 = mat(2), X4 = mat(3), B = mat(4), P = mat(5), Q = mat(6);
    ExpmCtx cx{D, n, cplx, count, sb, 0, mat(7).re};
    double c[17];
    c[0] = 1.0;
    for (int k = 1; k <= 16; ++k) c[k] = c[k - 1] / k;
    // X_b = tau_b A / 2^s (A shared by the sweep: instance stride 0)
    lincomb_kernel<<<1024, 256>>>(X.re, n, D, count, sb, 0.0, 1.0, scales, (const double*)are, 0, 0.0, nullptr, 0.0, nullptr, 0.0, nullptr);
    if (cplx) lincomb_kernel<<<1024, 256>>>(X.im, n, D, count, sb, 0.0, 1.0, scales, (const double*)aim, 0, 0.0, nullptr, 0.0, nullptr, 0.0, nullptr);
    CUDA_TRY(cudaGetLastError());
    OTN_TRY(cx.mul(X, X, X2));
    OTN_TRY(cx.mul(X2, X, X3));
    OTN_TRY(cx.mul(X2, X2, X4));
    // P = B3 + c16 X4
    OTN_TRY(cx.comb(P, c[12], c[13], X, c[14], X2, c[15], X3, c[16], X4));
    for (int j = 2; j >= 0; --j) {
        OTN_TRY(cx.mul(X4, P, Q));                                                           // Q = X4 P
        OTN_TRY(cx.comb(B, c[4 * j], c[4 * j + 1], X, c[4 * j + 2], X2, c[4 * j + 3], X3, 1.0, Q));   // P = Bj + Q
        std::swap(P, B);
    }
    for (int k = 0; k < s; ++k) { OTN_TRY(cx.mul(P, P, Q)); std::swap(P, Q); }
    CUDA_TRY(cudaMemcpy2DAsync(ore, bytes, P.re, (size_t)sb * 8, bytes, count, cudaMemcpyDeviceToDevice, 0));
    if (cplx) CUDA_TRY(cudaMemcpy2DAsync(oim, bytes, P.im, (size_t)sb * 8, bytes, count, cudaMemcpyDeviceToDevice, 0));
    OTN_TRY(finish_out(out_re, ore, bytes * count, 0));
    OTN_TRY(finish_out(out_im, oim, bytes * count, 0));
    CUDA_TRY(cudaDeviceSynchronize());
    return 0;
}

extern "C" int otn_expm(const double* A_re, const double* A_im, int D, double tau, double* out_re, double* out_im, int device) {
    return otn_expm_batch(A_re, A_im, D, &tau, 1, out_re, out_im, device);
}
