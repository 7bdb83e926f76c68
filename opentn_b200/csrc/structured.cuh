// Structured (block-diagonal) evaluation of the full-sites model.
//
// Every layer superoperator Y = sum_k E_k (x) E_k commutes with the swap S:(a,b)<->(b,a) of the doubled index, hence so do all
// chain products, tangents and (after symmetrising the target, see otn_create) all cotangents.  In the orthonormal basis
//     u^s_{ab} = (e_ab + e_ba)/sqrt2 (a<b),  u^s_{aa} = e_aa          (NS = dim(dim+1)/2 vectors, 136 at dim = 16)
//     u^a_{ab} = (e_ab - e_ba)/sqrt2 (a<b)                            (NA = dim(dim-1)/2 vectors, 120 at dim = 16)
// such matrices are block diagonal, Y = Ys (+) Ya with
//     Ys[{ab},{cd}] = n_ab n_cd sum_k (x_ac x_bd + x_ad x_bc)   (symmetric square of E_k;  n_ab = 1, n_aa = 1/sqrt2)
//     Ya[{ab},{cd}] =           sum_k (x_ac x_bd - x_ad x_bc)   (exterior square of E_k)
// so every D^3 product of the dense formulation (transformations.py:841-851) becomes one NS^3 and one NA^3 product:
// 2(136^3 + 120^3) = 8.5 MFLOP instead of 33.6 MFLOP (3.95x fewer), still on the FP64 tensor pipe.  Blocks are stored
// zero-padded to the GEMM tile (144 and 128), which keeps the padding exactly zero through every product.
// The same change of basis applied to the adjoint of the assembly gives (derivation in DESIGN.md section 3):
//     Z[(a,k),c] = sum_{b,d} ( nu_ab nu_cd Ws[{ab},{cd}] + sg_ab sg_cd Wa[{ab},{cd}] ) x[(b,k),d]
// with nu_ab = sqrt2 if a == b else 1, sg_ab = +1 (a<b), -1 (a>b), 0 (a == b).
#pragma once
#include <type_traits>
#include "assemble.cuh"
#include "otn_common.cuh"

namespace otn {

struct SymLayout {
    int dim, NS, NA;        // pair counts
    int Ds, Da;             // padded block sides (leading dimensions)
    long long off_a;        // offset of the antisymmetric block inside a slot (= Ds*Ds)
    long long slot;         // Ds*Ds + Da*Da
};

__host__ __device__ __forceinline__ int pair_s(int a, int b, int dim) { return a * dim - a * (a - 1) / 2 + (b - a); }       // a <= b
__host__ __device__ __forceinline__ int pair_a(int a, int b, int dim) { return a * dim - a * (a + 1) / 2 + (b - a - 1); }   // a <  b

// ---- assembly into blocks: one CTA per item, the NS unordered pairs {a,b} are dealt to the 8 warps ------------------
template <int DIM, bool TANGENT, bool PRIMAL>
__global__ void __launch_bounds__(256) assemble_sym_kernel(const double* __restrict__ x, const double* __restrict__ eta,
                                                           double* __restrict__ Y, double* __restrict__ Yd, SymLayout L, int K,
                                                           const int* active, int m) {
    extern __shared__ __align__(16) double sm[];
    // x / eta live unpadded in shared memory, 4-column groups XOR-swizzled by (row & SWZ): conflict-free fragment loads at
    // 64 KB instead of 80 KB, which lets two CTAs share an SM
    constexpr int LD = DIM, SWZ = DIM / 4 - 1, MI = DIM / 8, LS = DIM + 1, NS = DIM * (DIM + 1) / 2;
    const int item = blockIdx.x;
    if (active != nullptr && active[item / m] == 0) return;
    const int Kp = (K + 3) & ~3, rows = DIM * Kp;
    double* xs = sm;
    double* es = xs + rows * LD;
    double* scr = es + (TANGENT ? rows * LD : 0);        // [8 warps][2][DIM*LS]
    __shared__ unsigned char pc[NS], pd[NS];
    const long long np = (long long)DIM * K * DIM;
    for (int i = threadIdx.x; i < rows * DIM; i += 256) {
        const int r = i / DIM, c = i - r * DIM, aa = r / Kp, k = r - aa * Kp;
        const bool ok = k < K;
        const int o = r * LD + (c ^ ((r & SWZ) << 2));
        xs[o] = ok ? x[(long long)item * np + (aa * K + k) * DIM + c] : 0.0;
        if (TANGENT) es[o] = ok ? eta[(long long)item * np + (aa * K + k) * DIM + c] : 0.0;
    }
    for (int i = threadIdx.x; i < DIM * DIM; i += 256) {
        const int c = i / DIM, dd = i - c * DIM;
        if (c <= dd) { pc[pair_s(c, dd, DIM)] = (unsigned char)c; pd[pair_s(c, dd, DIM)] = (unsigned char)dd; }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    double* Bp = scr + warp * 2 * DIM * LS;
    double* Bt = Bp + DIM * LS;
    double* Ys = Y + (long long)item * L.slot;
    double* Ya = Ys + L.off_a;
    double* Yds = Yd + (long long)item * L.slot;
    double* Yda = Yds + L.off_a;
    const double rs2 = 0.70710678118654752440;
    for (int pid = warp; pid < NS; pid += 8) {
        int a = 0, rem = pid;
        while (rem >= DIM - a) { rem -= DIM - a; ++a; }
        const int b = a + rem;
        double acc[MI][MI][2], accd[MI][MI][2];
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < MI; ++j) { acc[i][j][0] = acc[i][j][1] = 0.0; accd[i][j][0] = accd[i][j][1] = 0.0; }
        for (int kk = 0; kk < Kp; kk += 4) {
            double xa[MI], xb[MI], ea[MI], eb[MI];
#pragma unroll
            for (int i = 0; i < MI; ++i) {
                // rows a*Kp+kk+t and b*Kp+kk+t: Kp and kk are multiples of 4, so (row & SWZ) == t (SWZ <= 3)
                const int col = (i * 8 + g) ^ ((t & SWZ) << 2);
                xa[i] = xs[(a * Kp + kk + t) * LD + col];
                xb[i] = xs[(b * Kp + kk + t) * LD + col];
                if (TANGENT) { ea[i] = es[(a * Kp + kk + t) * LD + col]; eb[i] = es[(b * Kp + kk + t) * LD + col]; }
            }
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < MI; ++j) {
                    if (PRIMAL) dmma884(acc[i][j][0], acc[i][j][1], xa[i], xb[j]);
                    if (TANGENT) { dmma884(accd[i][j][0], accd[i][j][1], ea[i], xb[j]); dmma884(accd[i][j][0], accd[i][j][1], xa[i], eb[j]); }
                }
        }
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < MI; ++j) {
                const int o = (i * 8 + g) * LS + j * 8 + 2 * t;
                if (PRIMAL) { Bp[o] = acc[i][j][0]; Bp[o + 1] = acc[i][j][1]; }
                if (TANGENT) { Bt[o] = accd[i][j][0]; Bt[o + 1] = accd[i][j][1]; }
            }
        __syncwarp();
        const double nab = (a == b) ? rs2 : 1.0;
        const long long rs = (long long)pair_s(a, b, DIM) * L.Ds;
        const long long ra = (a < b) ? (long long)pair_a(a, b, DIM) * L.Da : 0;
        for (int ci = lane; ci < NS; ci += 32) {
            const int c = pc[ci], dd = pd[ci];
            const double w = nab * ((c == dd) ? rs2 : 1.0);
            if (PRIMAL) {
                const double u = Bp[c * LS + dd], v = Bp[dd * LS + c];
                Ys[rs + ci] = w * (u + v);
                if (a < b && c < dd) Ya[ra + pair_a(c, dd, DIM)] = u - v;
            }
            if (TANGENT) {
                const double u = Bt[c * LS + dd], v = Bt[dd * LS + c];
                Yds[rs + ci] = w * (u + v);
                if (a < b && c < dd) Yda[ra + pair_a(c, dd, DIM)] = u - v;
            }
        }
        __syncwarp();
    }
}

// ---- assembly into blocks, second version (DIM = 16) ----------------------------------------------------------------------
// ncu of assemble_sym_kernel (profiles/ncu_step_r02.txt): 14 instructions per DMMA, tensor pipe 51 % active, 37.7 M shared-memory
// bank conflicts per launch (C fragments scattered with 8-byte stores into rows of 17 doubles, pair tables read per element).
// Same decomposition -- one CTA per item, a warp computes the 16 x 16 product B = X_a^T X_b of a pair {a,b} with DMMA and folds it
// into row {ab} of the symmetric and antisymmetric blocks -- but: a warp owns 17 CONSECUTIVE pairs (the a fragments stay in
// registers while b runs), the product goes to shared memory twice, row-major in rows of 24 doubles (16-byte stores, conflict free)
// and transposed in rows of 18 doubles (every half-warp hits 16 distinct 8-byte banks), so that B[c][d] and B[d][c] are both read along
// the output order; every per-lane index of the fold is computed once per kernel; x / eta arrive by cp.async.
struct AsmSym2Cfg {
    static constexpr int DIM = 16, NS = 136, NA = 120, LB = 24, LT = 18, SCR = DIM * LB + DIM * LT;     // doubles per warp
    static size_t smem_bytes(int Kp, bool tangent) { return ((size_t)DIM * Kp * DIM * (tangent ? 2 : 1) + (size_t)8 * SCR) * 8; }
};

template <bool TANGENT, bool PRIMAL>
__global__ void __launch_bounds__(256, 2) assemble_sym2_kernel(const double* __restrict__ x, const double* __restrict__ eta, double* __restrict__ Y,
                                                               double* __restrict__ Yd, SymLayout L, int K, const int* active, int m) {
    using Cfg = AsmSym2Cfg;
    constexpr int DIM = Cfg::DIM, NS = Cfg::NS, MI = 2, LB = Cfg::LB, LT = Cfg::LT, RND = (NS + 31) / 32, PPW = NS / 8;
    extern __shared__ __align__(16) double sm[];
    const int item = blockIdx.x;
    if (active != nullptr && active[item / m] == 0) return;
    const int Kp = (K + 3) & ~3, rows = DIM * Kp, KQ = Kp / 4;
    double* xs = sm;                                    // [a*Kp + k][16], 4-column groups XOR-swizzled by (row & 3)
    double* es = xs + rows * DIM;
    double* scr = es + (TANGENT ? rows * DIM : 0);
    const long long np = (long long)DIM * K * DIM;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    for (int id = tid; id < rows * 8; id += 256) {      // 16-byte chunks
        const int r = id >> 3, c = (id & 7) * 2, aa = r / Kp, k = r - aa * Kp;
        const int o = r * DIM + (c ^ ((r & 3) << 2));
        if (k < K) {
            const long long so = (long long)item * np + (aa * K + k) * DIM + c;
            cp_async16(xs + o, x + so);
            if (TANGENT) cp_async16(es + o, eta + so);
        } else {
            xs[o] = xs[o + 1] = 0.0;
            if (TANGENT) es[o] = es[o + 1] = 0.0;
        }
    }
    cp_async_commit();

    // fold constants of this lane: output columns ci = lane + 32 r of a block row <-> (c, d), c <= d
    int offU[RND], offV[RND], pa[RND];
    double wc[RND];
#pragma unroll
    for (int r = 0; r < RND; ++r) {
        int ci = lane + 32 * r, c = 0;
        if (ci >= NS) ci = NS - 1;                      // (masked at the store)
        int rem = ci;
        while (rem >= DIM - c) { rem -= DIM - c; ++c; }
        const int dd = c + rem;
        offU[r] = c * LB + dd;
        offV[r] = DIM * LB + c * LT + dd;
        pa[r] = (c < dd) ? pair_a(c, dd, DIM) : -1;
        wc[r] = (c == dd) ? 0.70710678118654752440 : 1.0;
    }
    double* Bw = scr + warp * Cfg::SCR;                 // [16][LB] row-major | [16][LT] transposed
    double* Ys = Y + (long long)item * L.slot;
    double* Yds = Yd + (long long)item * L.slot;
    cp_async_wait<0>();
    __syncthreads();

    // pairs pid = 17 warp .. 17 warp + 16 in the order of pair_s
    int a = 0, b;
    {
        int rem = warp * PPW;
        while (rem >= DIM - a) { rem -= DIM - a; ++a; }
        b = a + rem;
    }
    const int fcol0 = g ^ (t << 2), fcol1 = (8 + g) ^ (t << 2);       // fragment columns c = 8 i + g behind the swizzle (row & 3 == t)
    double xa[4][MI], ea[4][MI];
    int a_loaded = -1;
    auto fold = [&](const double (&acc)[MI][MI][2], double* Os, double* Oa, bool anti) {
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < MI; ++j) {
                *reinterpret_cast<double2*>(Bw + (8 * i + g) * LB + 8 * j + 2 * t) = make_double2(acc[i][j][0], acc[i][j][1]);
                Bw[DIM * LB + (8 * j + 2 * t) * LT + 8 * i + g] = acc[i][j][0];
                Bw[DIM * LB + (8 * j + 2 * t + 1) * LT + 8 * i + g] = acc[i][j][1];
            }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < RND; ++r) {
            const double u = Bw[offU[r]], v = Bw[offV[r]];
            if (r < RND - 1 || lane + 32 * r < NS) {
                Os[lane + 32 * r] = wc[r] * (u + v);
                if (anti && pa[r] >= 0) Oa[pa[r]] = u - v;
            }
        }
        __syncwarp();
    };
#pragma unroll 1
    for (int n = 0; n < PPW; ++n) {
        if (a != a_loaded) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (q < KQ) {
                    const int row = (a * Kp + 4 * q + t) * DIM;
                    xa[q][0] = xs[row + fcol0]; xa[q][1] = xs[row + fcol1];
                    if (TANGENT) { ea[q][0] = es[row + fcol0]; ea[q][1] = es[row + fcol1]; }
                }
            a_loaded = a;
        }
        double acc[MI][MI][2], accd[MI][MI][2];
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < MI; ++j) { acc[i][j][0] = acc[i][j][1] = 0.0; accd[i][j][0] = accd[i][j][1] = 0.0; }
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (q < KQ) {
                const int row = (b * Kp + 4 * q + t) * DIM;
                double xb[MI], eb[MI];
                xb[0] = xs[row + fcol0]; xb[1] = xs[row + fcol1];
                if (TANGENT) { eb[0] = es[row + fcol0]; eb[1] = es[row + fcol1]; }
                if (a == b) {                                           // weight 1/sqrt2 of a diagonal pair, folded into the b operand
                    const double h = 0.70710678118654752440;
                    xb[0] *= h; xb[1] *= h;
                    if (TANGENT) { eb[0] *= h; eb[1] *= h; }
                }
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < MI; ++j) {
                        if (PRIMAL) dmma884(acc[i][j][0], acc[i][j][1], xa[q][i], xb[j]);
                        if (TANGENT) { dmma884(accd[i][j][0], accd[i][j][1], ea[q][i], xb[j]); dmma884(accd[i][j][0], accd[i][j][1], xa[q][i], eb[j]); }
                    }
            }
        const long long rs = (long long)pair_s(a, b, DIM) * L.Ds;
        const long long ra = L.off_a + ((a < b) ? (long long)pair_a(a, b, DIM) * L.Da : 0);
        if (PRIMAL) fold(acc, Ys + rs, Ys + ra, a < b);
        if (TANGENT) fold(accd, Yds + rs, Yds + ra, a < b);
        if (++b == DIM) { ++a; b = a; }
    }
}

// ---- adjoint of the assembly from cotangent blocks -----------------------------------------------------------------------
// AS CTAs per instance, each owning DIM/AS values of the output index a; every CTA streams the rows {(a,b)}_a of Ws and Wa for
// b = 0..DIM-1 through a 3-stage cp.async ring and feeds DMMA.  Versus the first version (dense tile decoded into shared
// memory, 4.7 ms per step): (i) the A fragments are decoded straight from the staged block rows into registers
// (no dense tile, one barrier per stage instead of two); (ii) up to two right-hand sides share one pass over W
// (out1 (+)= A u1, out2 (+)= A u2): in a joint (f, rgrad, H eta) evaluation the primal cotangent W is needed against both x
// (gradient) and eta (tangent of the adjoint), so 9 streaming passes per step become 6; (iii) the right-hand sides live
// unpadded in shared memory behind an XOR swizzle, which lets two CTAs share an SM.
template <int DIM, int AS, int NRHS, int KP>
struct BasSym2Cfg {
    static constexpr int STAGES = 3, D2 = DIM * DIM, DA = DIM / AS;
    static constexpr int RAW = DA * D2;
    static constexpr size_t SMEM_BYTES = (size_t)(STAGES * RAW + NRHS * DIM * KP * DIM) * 8;
};

// KP = Kraus rank rounded up to a multiple of 8 (compile time, so that every shared-memory offset of the inner loop folds into
// an immediate; the first version recomputed the swizzled addresses per DMMA and spent 2/3 of its instructions on integer math).
template <int DIM, int AS, int NRHS, int KP>
__global__ void __launch_bounds__(256) back_assemble_sym2_kernel(const double* __restrict__ W, long long w_stride,
                                                                 const double* __restrict__ u1, const double* __restrict__ u2,
                                                                 double* __restrict__ out1, double* __restrict__ out2, int acc1, int acc2,
                                                                 SymLayout L, int K, const int* active, int m, int layer) {
    using Cfg = BasSym2Cfg<DIM, AS, NRHS, KP>;
    constexpr int D2 = Cfg::D2, ST = Cfg::STAGES, DA = Cfg::DA, NS = DIM * (DIM + 1) / 2, NA = DIM * (DIM - 1) / 2;
    constexpr int MI = DIM / 8, DQ = DIM / 4, SWZ = DIM / 4 - 1, NI = KP / 8, NJ = NRHS * NI;
    constexpr int USB = NRHS * KP * DIM;                 // doubles of the right-hand sides per b
    static_assert(DA == 8, "one a value per warp");
    extern __shared__ __align__(16) double sm[];
    const int inst = blockIdx.x / AS, a_lo = (blockIdx.x - inst * AS) * DA;
    if (active != nullptr && active[inst] == 0) return;
    const int item = inst * m + layer;
    double* raw = sm;
    double* us = raw + ST * Cfg::RAW;           // [b][rhs][k][DIM], 4-column groups XOR-swizzled by (k & SWZ)
    const long long np = (long long)DIM * K * DIM;
    const double* Ws = W + (long long)inst * w_stride;
    const double* Wa = Ws + L.off_a;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;

    // right-hand sides: asynchronous 16-byte copies (a swizzled 4-column group keeps its two halves contiguous)
    for (int i = tid; i < DIM * NRHS * KP * (DIM / 2); i += 256) {
        const int rr = i / (DIM / 2), ch = i - rr * (DIM / 2);          // rr = (b * NRHS + rhs) * KP + k
        const int k = rr % KP, br = rr / KP, rhs = br % NRHS, bb = br / NRHS;
        const int c = 2 * ch;
        double* dst = us + rr * DIM + (c ^ ((k & SWZ) << 2));
        if (k < K) {
            const double* u = (NRHS == 2 && rhs == 1) ? u2 : u1;
            cp_async16(dst, u + (long long)item * np + (bb * K + k) * DIM + c);
        } else { dst[0] = 0.0; dst[1] = 0.0; }
    }
    auto load_stage = [&](int b, int s) {
        double* dst = raw + s * Cfg::RAW;
        constexpr int CS = NS / 2, CA = NA / 2;
        for (int id = tid; id < DA * (CS + CA); id += 256) {
            const int al = id / (CS + CA), ch = id - al * (CS + CA), a = a_lo + al;
            const int lo = a < b ? a : b, hi = a < b ? b : a;
            if (ch < CS) cp_async16(dst + al * D2 + 2 * ch, Ws + (long long)pair_s(lo, hi, DIM) * L.Ds + 2 * ch);
            else if (a != b) cp_async16(dst + al * D2 + NS + 2 * (ch - CS), Wa + (long long)pair_a(lo, hi, DIM) * L.Da + 2 * (ch - CS));
        }
    };
    load_stage(0, 0);
    cp_async_commit();                                   // group 0: right-hand sides + stage 0
#pragma unroll
    for (int s = 1; s < ST - 1; ++s) { load_stage(s, s); cp_async_commit(); }

    // per-lane decode constants: fragment element (row c_i = i*8+g, column d = dq*4+t)
    int ixs[MI][DQ], ixa[MI][DQ];
    double cfd[MI][DQ], scd[MI][DQ];
    const double s2 = 1.41421356237309504880;
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int dq = 0; dq < DQ; ++dq) {
            const int c = i * 8 + g, dd = dq * 4 + t, lo = c < dd ? c : dd, hi = c < dd ? dd : c;
            ixs[i][dq] = pair_s(lo, hi, DIM);
            ixa[i][dq] = (c == dd) ? NS : NS + pair_a(lo, hi, DIM);
            cfd[i][dq] = (c == dd) ? s2 : 1.0;
            scd[i][dq] = (c == dd) ? 0.0 : ((c < dd) ? 1.0 : -1.0);
        }
    // swizzled column offset of this lane inside a right-hand-side row: ((dq*4 + t) ^ ((g & SWZ) << 2))
    int dsw[DQ];
#pragma unroll
    for (int dq = 0; dq < DQ; ++dq) dsw[dq] = g * DIM + (((dq ^ (g & SWZ)) << 2) + t);
    double acc[MI][NJ][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const int a = a_lo + warp;

    for (int b = 0; b < DIM; ++b) {
        cp_async_wait<ST - 2>();
        __syncthreads();                                   // stage b landed; stage b-1 fully consumed by every warp
        if (b + ST - 1 < DIM) load_stage(b + ST - 1, (b + ST - 1) % ST);
        cp_async_commit();
        const double* R = raw + (b % ST) * Cfg::RAW + warp * D2;
        const double* ub = us + b * USB;
        const double nab = (a == b) ? s2 : 1.0;
        const double sgo = (a == b) ? 0.0 : ((a < b) ? 1.0 : -1.0);   // sign of the antisymmetric part (0 on the diagonal: no Wa row staged)
#pragma unroll
        for (int dq = 0; dq < DQ; ++dq) {
            double wa[MI];
#pragma unroll
            for (int i = 0; i < MI; ++i) {
                const double ra = (a == b) ? 0.0 : R[ixa[i][dq]];
                wa[i] = fma(sgo * scd[i][dq], ra, nab * cfd[i][dq] * R[ixs[i][dq]]);
            }
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const double uv = ub[((j / NI) * KP + (j % NI) * 8) * DIM + dsw[dq]];
#pragma unroll
                for (int i = 0; i < MI; ++i) dmma884(acc[i][j][0], acc[i][j][1], wa[i], uv);
            }
        }
    }
    cp_async_wait<0>();
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int c = i * 8 + g;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int rhs = j / NI;
            double* out = (NRHS == 2 && rhs == 1) ? out2 : out1;
            const int accm = (NRHS == 2 && rhs == 1) ? acc2 : acc1;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int k = (j % NI) * 8 + 2 * t + e;
                if (k < K) {
                    const long long og = (long long)item * np + (a * K + k) * DIM + c;
                    out[og] = accm ? out[og] + acc[i][j][e] : acc[i][j][e];
                }
            }
        }
    }
}

// ---- adjoint of the assembly, all layers and both cotangents in ONE launch (fused-chain path) ----------------------------------------
// With the fused chain every layer cotangent W_l and its tangent Wd_l exist when the pull-backs start, so one kernel does, for
// every (instance, layer):      zraw = A(W) x,      zdraw = A(W) eta + A(Wd) x
// (A = the adjoint of the assembly above).  Versus back_assemble_sym2_kernel called 6 times per step: the rows of W AND Wd of
// one b travel through the cp.async ring together with the 16 x KP slices x_b, eta_b they are contracted with (no per-CTA copy
// of all of x / eta: 108 KB instead of 160 KB of shared memory, two CTAs per SM), one barrier serves 48 DMMA per warp instead of
// 32 / 16, and x, eta, the decode constants and the launch are paid once instead of twice.
struct BackAll {
    const double* W[16]; long long sW[16];        // per layer: cotangent blocks of instance 0 and the instance stride
    const double* Wd[16]; long long sWd[16];
};

template <int DIM, int AS, int KP>
struct BasSym3Cfg {
    static constexpr int STAGES = 3, D2 = DIM * DIM, DA = DIM / AS;
    static constexpr int RAW = DA * D2;                       // rows {(a,b)}_a of one matrix
    static constexpr int UB = KP * DIM;                       // x_b (or eta_b): [k][DIM], swizzled
    static constexpr int STAGE = 2 * RAW + 2 * UB;
    static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE * 8;
};

// GRAD: zraw = A(W) x is wanted; TAN: zdraw = A(W) eta + A(Wd) x is wanted (gradient only / HVP with cached primal / both)
template <int DIM, int AS, int KP, bool GRAD, bool TAN>
__global__ void __launch_bounds__(256, 2) back_assemble_sym3_kernel(const BackAll P, const double* __restrict__ x, const double* __restrict__ eta,
                                                                    double* __restrict__ zraw, double* __restrict__ zdraw, SymLayout L, int K,
                                                                    const int* active, int m) {
    using Cfg = BasSym3Cfg<DIM, AS, KP>;
    constexpr int D2 = Cfg::D2, ST = Cfg::STAGES, DA = Cfg::DA, NS = DIM * (DIM + 1) / 2, NA = DIM * (DIM - 1) / 2;
    constexpr int MI = DIM / 8, DQ = DIM / 4, SWZ = DIM / 4 - 1, NI = KP / 8;
    static_assert(DA == 8, "one a value per warp");
    extern __shared__ __align__(16) double sm[];
    const int inst = blockIdx.x / AS, a_lo = (blockIdx.x - inst * AS) * DA, layer = blockIdx.y;
    if (active != nullptr && active[inst] == 0) return;
    const long long item = (long long)inst * m + layer;
    const long long np = (long long)DIM * K * DIM;
    const double* Ws = P.W[layer] + (long long)inst * P.sW[layer];
    const double* Wa = Ws + L.off_a;
    const double* Wds = TAN ? P.Wd[layer] + (long long)inst * P.sWd[layer] : nullptr;
    const double* Wda = TAN ? Wds + L.off_a : nullptr;
    const double* xg = x + item * np;
    const double* eg = TAN ? eta + item * np : nullptr;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;

    auto load_stage = [&](int b, int s) {
        double* dst = sm + s * Cfg::STAGE;
        constexpr int CS = NS / 2, CA = NA / 2, CR = CS + CA;
        for (int id = tid; id < (TAN ? 2 : 1) * DA * CR; id += 256) {
            const int which = id / (DA * CR), rem = id - which * (DA * CR);
            const int al = rem / CR, ch = rem - al * CR, a = a_lo + al;
            const int lo = a < b ? a : b, hi = a < b ? b : a;
            double* row = dst + which * Cfg::RAW + al * D2;
            const double* S = which ? Wds : Ws;
            const double* A = which ? Wda : Wa;
            if (ch < CS) cp_async16(row + 2 * ch, S + (long long)pair_s(lo, hi, DIM) * L.Ds + 2 * ch);
            else if (a != b) cp_async16(row + NS + 2 * (ch - CS), A + (long long)pair_a(lo, hi, DIM) * L.Da + 2 * (ch - CS));
        }
        // x_b, eta_b: rows (b, k), 4-column groups XOR-swizzled by (k & SWZ) (a swizzled group keeps its two halves contiguous)
        double* ub = dst + 2 * Cfg::RAW;
        for (int id = tid; id < (TAN ? 2 : 1) * KP * (DIM / 2); id += 256) {
            const int which = id / (KP * (DIM / 2)), rem = id - which * (KP * (DIM / 2));
            const int k = rem / (DIM / 2), c = 2 * (rem - k * (DIM / 2));
            double* d = ub + which * Cfg::UB + k * DIM + (c ^ ((k & SWZ) << 2));
            if (k < K) cp_async16(d, (which ? eg : xg) + (long long)(b * K + k) * DIM + c);
            else { d[0] = 0.0; d[1] = 0.0; }
        }
    };
#pragma unroll
    for (int s = 0; s < ST - 1; ++s) { load_stage(s, s); cp_async_commit(); }

    // per-lane decode constants: fragment element (row c_i = i*8+g, column d = dq*4+t)
    int ixs[MI][DQ], ixa[MI][DQ];
    double cfd[MI][DQ], scd[MI][DQ];
    const double s2 = 1.41421356237309504880;
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int dq = 0; dq < DQ; ++dq) {
            const int c = i * 8 + g, dd = dq * 4 + t, lo = c < dd ? c : dd, hi = c < dd ? dd : c;
            ixs[i][dq] = pair_s(lo, hi, DIM);
            ixa[i][dq] = (c == dd) ? NS : NS + pair_a(lo, hi, DIM);
            cfd[i][dq] = (c == dd) ? s2 : 1.0;
            scd[i][dq] = (c == dd) ? 0.0 : ((c < dd) ? 1.0 : -1.0);
        }
    int dsw[DQ];
#pragma unroll
    for (int dq = 0; dq < DQ; ++dq) dsw[dq] = g * DIM + (((dq ^ (g & SWZ)) << 2) + t);
    double acc1[MI][NI][2], acc2[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) { acc1[i][j][0] = acc1[i][j][1] = 0.0; acc2[i][j][0] = acc2[i][j][1] = 0.0; }
    const int a = a_lo + warp;

    for (int b = 0; b < DIM; ++b) {
        cp_async_wait<ST - 2>();
        __syncthreads();                                   // stage b landed; stage b-1 fully consumed by every warp
        if (b + ST - 1 < DIM) load_stage(b + ST - 1, (b + ST - 1) % ST);
        cp_async_commit();
        const double* stage = sm + (b % ST) * Cfg::STAGE;
        const double* R = stage + warp * D2;
        const double* Rd = stage + Cfg::RAW + warp * D2;
        const double* ux = stage + 2 * Cfg::RAW;
        const double* ue = ux + Cfg::UB;
        const double nab = (a == b) ? s2 : 1.0;
        const double sgo = (a == b) ? 0.0 : ((a < b) ? 1.0 : -1.0);
#pragma unroll
        for (int dq = 0; dq < DQ; ++dq) {
            double wa[MI], wda[MI];
#pragma unroll
            for (int i = 0; i < MI; ++i) {
                const double cs = nab * cfd[i][dq], ca = sgo * scd[i][dq];
                const double ra = (a == b) ? 0.0 : R[ixa[i][dq]];
                wa[i] = fma(ca, ra, cs * R[ixs[i][dq]]);
                if (TAN) {
                    const double rda = (a == b) ? 0.0 : Rd[ixa[i][dq]];
                    wda[i] = fma(ca, rda, cs * Rd[ixs[i][dq]]);
                }
            }
#pragma unroll
            for (int j = 0; j < NI; ++j) {
                const double vx = ux[(j * 8) * DIM + dsw[dq]];
                const double ve = TAN ? ue[(j * 8) * DIM + dsw[dq]] : 0.0;
#pragma unroll
                for (int i = 0; i < MI; ++i) {
                    if (GRAD) dmma884(acc1[i][j][0], acc1[i][j][1], wa[i], vx);
                    if (TAN) {
                        dmma884(acc2[i][j][0], acc2[i][j][1], wa[i], ve);
                        dmma884(acc2[i][j][0], acc2[i][j][1], wda[i], vx);
                    }
                }
            }
        }
    }
    cp_async_wait<0>();
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int c = i * 8 + g;
#pragma unroll
        for (int j = 0; j < NI; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int k = j * 8 + 2 * t + e;
                if (k < K) {
                    const long long og = item * np + (a * K + k) * DIM + c;
                    if (GRAD) zraw[og] = acc1[i][j][e];
                    if (TAN) zdraw[og] = acc2[i][j][e];
                }
            }
    }
}

// ---- pull-back, fourth version: no CTA-wide barrier, no per-thread copy addressing --------------------------------------------
// ncu of back_assemble_sym3_kernel (profiles/ncu_step_r02.txt): 14 instructions per DMMA, 30 % of them IMAD -- the per-thread
// address arithmetic of the cp.async ring -- tensor pipe 52 % active, one __syncthreads per b.  Here every warp owns one `a` and
// streams ITS rows {a,b} of W / Wd (S row | A row = 2 KB per matrix) through a private 3-stage ring with 1-D bulk copies (TMA
// engine, one lane issues, completion by transaction count on a per-warp mbarrier); the x_b / eta_b tiles are shared by the 8
// warps of the CTA, copied by one warp per b (rotating) with cp.async into rows of 20 doubles (fragment loads take the minimum of
// two wavefronts) and handed over by mbarriers (full: cp.async arrivals; empty: one arrival per warp, waited for one iteration
// later).  The decode costs one FP64 instruction per fragment element: sign and sqrt2 factors that depend on (a,b) select one of
// three loop bodies per b (a < b, a > b, a == b: warp-uniform), the ones that depend on (c,d) are static per fragment or, for the
// four fragments that straddle the diagonal, an XOR of the sign bit; diagonal elements read the antisymmetric part from a
// zero pad.
template <int KP>
struct BasSym4Cfg {
    static constexpr int DIM = 16, NS = 136, NA = 120, ST = 3, WARPS = 8;
    static constexpr int ROWB = (NS + NA) * 8;            // bytes: S row | A row of one matrix
    static constexpr int WARP_STAGE = 2 * ROWB;           // W, Wd
    static constexpr int XLD = 20;                        // doubles per row of an x_b / eta_b tile (16 + 4 zero pad)
    static constexpr int XT = KP * XLD * 8;               // bytes per tile
    static constexpr int STAGE = WARPS * WARP_STAGE + 2 * XT;
    static constexpr int NBAR = WARPS * ST + 2 * ST;
    static constexpr size_t SMEM_BYTES = (size_t)ST * STAGE + NBAR * 8;
};

template <int KP, bool GRAD, bool TAN>
__global__ void __launch_bounds__(256, KP <= 16 ? 2 : 1) back_assemble_sym4_kernel(const BackAll P, const double* __restrict__ x, const double* __restrict__ eta,
                                                                                   double* __restrict__ zraw, double* __restrict__ zdraw, SymLayout L, int K,
                                                                                   const int* active, int m) {
    using Cfg = BasSym4Cfg<KP>;
    constexpr int DIM = Cfg::DIM, NS = Cfg::NS, NA = Cfg::NA, ST = Cfg::ST, MI = 2, DQ = 4, NI = KP / 8, XLD = Cfg::XLD;
    constexpr int ROWB = Cfg::ROWB, WST = Cfg::WARP_STAGE, XT = Cfg::XT, STAGE = Cfg::STAGE;
    extern __shared__ __align__(128) unsigned char smb[];
    unsigned long long* fullr = reinterpret_cast<unsigned long long*>(smb + ST * STAGE);      // [warp][stage], count 1 + tx
    unsigned long long* fullx = fullr + Cfg::WARPS * ST;                                          // [stage], 32 cp.async arrivals
    unsigned long long* emptyx = fullx + ST;                                                      // [stage], one arrival per warp
    const int inst = blockIdx.x >> 1, a_lo = (blockIdx.x & 1) * 8, layer = blockIdx.y;
    if (active != nullptr && active[inst] == 0) return;
    const long long item = (long long)inst * m + layer;
    const long long np = (long long)DIM * K * DIM;
    const double* Ws = P.W[layer] + (long long)inst * P.sW[layer];
    const double* Wds = TAN ? P.Wd[layer] + (long long)inst * P.sWd[layer] : nullptr;
    const double* xg = x + item * np;
    const double* eg = TAN ? eta + item * np : nullptr;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int a = a_lo + warp;

    if (tid == 0) {
        for (int i = 0; i < Cfg::WARPS * ST; ++i) mbar_init(&fullr[i], 1);
        for (int s = 0; s < ST; ++s) { mbar_init(&fullx[s], 32); mbar_init(&emptyx[s], Cfg::WARPS); }
        mbar_fence_init();
    }
    for (int s = 0; s < ST; ++s) {                        // tiles start as zeros: pad columns and rows k >= K stay zero for good
        double2* z = reinterpret_cast<double2*>(smb + s * STAGE + Cfg::WARPS * WST);
        for (int i = tid; i < 2 * XT / 16; i += 256) z[i] = make_double2(0.0, 0.0);
    }
    __syncthreads();

    auto issue_rows = [&](int bb) {                       // lane 0 of every warp: rows {a,bb} of W (and Wd) into the warp's ring
        const int s = bb % ST;
        unsigned char* dst = smb + s * STAGE + warp * WST;
        unsigned long long* bar = &fullr[warp * ST + s];
        const int lo = a < bb ? a : bb, hi = a < bb ? bb : a;
        const long long rs = (long long)pair_s(lo, hi, DIM) * L.Ds;
        const bool off = a != bb;
        const long long ra = off ? L.off_a + (long long)pair_a(lo, hi, DIM) * L.Da : 0;
        mbar_arrive_expect_tx(bar, (unsigned)((NS * 8 + (off ? NA * 8 : 0)) * (TAN ? 2 : 1)));
        bulk_g2s(dst, Ws + rs, NS * 8, bar);
        if (off) bulk_g2s(dst + NS * 8, Ws + ra, NA * 8, bar);
        if (TAN) {
            bulk_g2s(dst + ROWB, Wds + rs, NS * 8, bar);
            if (off) bulk_g2s(dst + ROWB + NS * 8, Wds + ra, NA * 8, bar);
        }
    };
    auto issue_x = [&](int bb) {                          // one whole warp: x_bb (and eta_bb), rows (bb,k), 16-byte chunks
        const int s = bb % ST;
        unsigned char* xt = smb + s * STAGE + Cfg::WARPS * WST;
#pragma unroll
        for (int q = 0; q < KP / 4; ++q) {
            const int k = (lane >> 3) + 4 * q;
            if (k < K) {
                const long long so = (long long)(bb * K + k) * DIM + (lane & 7) * 2;
                cp_async16(xt + k * (XLD * 8) + (lane & 7) * 16, xg + so);
                if (TAN) cp_async16(xt + XT + k * (XLD * 8) + (lane & 7) * 16, eg + so);
            }
        }
        mbar_arrive_cp_async(&fullx[s]);
    };
    if (lane == 0) { issue_rows(0); issue_rows(1); }
    if (warp < 2) issue_x(warp);

    // per-lane constants.  Fragment element (i, dq): c = i*8 + g, d = dq*4 + t.  (0, dq>=2): c < d; (1, dq<2): c > d; the other
    // four straddle the diagonal ("mixed", index q = i*2 + (dq&1)).  Addresses are 32-bit shared addresses of stage 0; stage, matrix
    // and tile offsets are immediates of the loads.
    const unsigned sm0 = (unsigned)__cvta_generic_to_shared(smb);
    const unsigned zpad = sm0 + Cfg::WARPS * WST + DIM * 8;             // a zero pad of the x tile
    unsigned pS[MI][DQ], pA[MI][DQ], pAd[4], sgm[4];
    double cfm[4];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int dq = 0; dq < DQ; ++dq) {
            const int c = i * 8 + g, dd = dq * 4 + t, lo = c < dd ? c : dd, hi = c < dd ? dd : c;
            pS[i][dq] = sm0 + warp * WST + pair_s(lo, hi, DIM) * 8;
            pA[i][dq] = (c == dd) ? zpad : sm0 + warp * WST + (NS + pair_a(lo, hi, DIM)) * 8;
            if ((i == 0) == (dq < 2)) {
                const int q = i * 2 + (dq & 1);
                pAd[q] = (c == dd) ? zpad : pA[i][dq] + ROWB;
                sgm[q] = (c > dd) ? 0x80000000u : 0u;
                cfm[q] = (c == dd) ? 1.41421356237309504880 : 1.0;
            }
        }
    const unsigned pU = sm0 + Cfg::WARPS * WST + (g * XLD + t) * 8;     // B fragments: x_b[k = j*8 + g][d = dq*4 + t]
    double acc1[MI][NI][2], acc2[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) { acc1[i][j][0] = acc1[i][j][1] = 0.0; acc2[i][j][0] = acc2[i][j][1] = 0.0; }

    struct Frag { double rs[MI], ra[MI], rds[MI], rda[MI], vx[NI], ve[NI]; };
    // MODE: +1 (a < b), -1 (a > b), 0 (a == b: symmetric part only, weight sqrt2); SO: byte offset of the stage
    auto load = [&](auto MODE_, auto SO_, auto DQ_, Frag& f) {
        constexpr int MODE = decltype(MODE_)::value, SO = decltype(SO_)::value, dq = decltype(DQ_)::value;
        static_for<MI>([&](auto I_) {
            constexpr int i = decltype(I_)::value;
            constexpr bool mixed = (i == 0) == (dq < 2);
            constexpr int q = i * 2 + (dq & 1);
            f.rs[i] = lds_f64<SO>(pS[i][dq]);
            if (TAN) f.rds[i] = lds_f64<SO + ROWB>(pS[i][dq]);
            if (MODE != 0) {
                f.ra[i] = lds_f64<SO>(pA[i][dq]);
                if (TAN) f.rda[i] = mixed ? lds_f64<SO>(pAd[q]) : lds_f64<SO + ROWB>(pA[i][dq]);
            }
        });
        static_for<NI>([&](auto J_) {
            constexpr int j = decltype(J_)::value;
            f.vx[j] = lds_f64<SO + (j * 8 * XLD + dq * 4) * 8>(pU);
            if (TAN) f.ve[j] = lds_f64<SO + XT + (j * 8 * XLD + dq * 4) * 8>(pU);
        });
    };
    auto flip = [](double v, unsigned msk) { return __hiloint2double(__double2hiint(v) ^ (int)msk, __double2loint(v)); };
    auto mma = [&](auto MODE_, auto DQ_, const Frag& f) {
        constexpr int MODE = decltype(MODE_)::value, dq = decltype(DQ_)::value;
        double wa[MI], wda[MI];
#pragma unroll
        for (int i = 0; i < MI; ++i) {
            const bool mixed = (i == 0) == (dq < 2);
            const int q = i * 2 + (dq & 1);
            if (MODE == 0) {
                const double cf = mixed ? cfm[q] * 1.41421356237309504880 : 1.41421356237309504880;
                wa[i] = cf * f.rs[i];
                if (TAN) wda[i] = cf * f.rds[i];
            } else if (!mixed) {
                const bool plus = (i == 0) == (MODE > 0);             // sign of sigma_ab sigma_cd
                wa[i] = plus ? f.rs[i] + f.ra[i] : f.rs[i] - f.ra[i];
                if (TAN) wda[i] = plus ? f.rds[i] + f.rda[i] : f.rds[i] - f.rda[i];
            } else {
                const double ra = flip(f.ra[i], sgm[q]);
                wa[i] = MODE > 0 ? fma(cfm[q], f.rs[i], ra) : fma(cfm[q], f.rs[i], -ra);
                if (TAN) { const double rda = flip(f.rda[i], sgm[q]); wda[i] = MODE > 0 ? fma(cfm[q], f.rds[i], rda) : fma(cfm[q], f.rds[i], -rda); }
            }
        }
#pragma unroll
        for (int j = 0; j < NI; ++j)
#pragma unroll
            for (int i = 0; i < MI; ++i) {
                if (GRAD) dmma884(acc1[i][j][0], acc1[i][j][1], wa[i], f.vx[j]);
                if (TAN) {
                    dmma884(acc2[i][j][0], acc2[i][j][1], wa[i], f.ve[j]);
                    dmma884(acc2[i][j][0], acc2[i][j][1], wda[i], f.vx[j]);
                }
            }
    };
    auto body = [&](auto MODE_, auto SO_) {                             // fragments of dq+1 are loaded before the DMMAs of dq
        Frag f0, f1;
        load(MODE_, SO_, std::integral_constant<int, 0>{}, f0);
        load(MODE_, SO_, std::integral_constant<int, 1>{}, f1);
        mma(MODE_, std::integral_constant<int, 0>{}, f0);
        load(MODE_, SO_, std::integral_constant<int, 2>{}, f0);
        mma(MODE_, std::integral_constant<int, 1>{}, f1);
        load(MODE_, SO_, std::integral_constant<int, 3>{}, f1);
        mma(MODE_, std::integral_constant<int, 2>{}, f0);
        mma(MODE_, std::integral_constant<int, 3>{}, f1);
    };

    unsigned ph = 0;                                                    // parity of the ring's current round
#pragma unroll 1
    for (int b0 = 0; b0 < DIM; b0 += ST, ph ^= 1) {
        static_for<ST>([&](auto S_) {
            constexpr int s = decltype(S_)::value;
            const int b = b0 + s;
            if (b < DIM) {
                if (b + 2 < DIM) {
                    if (lane == 0) issue_rows(b + 2);                 // stage (b-1) % ST: this warp left it one iteration ago
                    if (warp == (b & 7)) {
                        if (b >= 1) mbar_wait(&emptyx[(s + 2) % ST], s == 0 ? ph ^ 1 : ph);     // every warp has left iteration b-1
                        issue_x(b + 2);
                    }
                }
                mbar_wait(&fullr[warp * ST + s], ph);
                mbar_wait(&fullx[s], ph);
                using SO = std::integral_constant<int, s * STAGE>;
                if (b > a) body(std::integral_constant<int, 1>{}, SO{});
                else if (b < a) body(std::integral_constant<int, -1>{}, SO{});
                else body(std::integral_constant<int, 0>{}, SO{});
                __syncwarp();
                if (lane == 0) mbar_arrive(&emptyx[s]);
            }
        });
    }
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int c = i * 8 + g;
#pragma unroll
        for (int j = 0; j < NI; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int k = j * 8 + 2 * t + e;
                if (k < K) {
                    const long long og = item * np + (a * K + k) * DIM + c;
                    if (GRAD) zraw[og] = acc1[i][j][e];
                    if (TAN) zdraw[og] = acc2[i][j][e];
                }
            }
    }
}

// ---- blocks -> dense superoperator (otn_model in structured mode) ------------------------------------------------------
//   M[(a,b),(c,d)] = cs_ab cs_cd Ms[{ab},{cd}] + ca_ab ca_cd Ma[{ab},{cd}],  cs = 1 (a==b) else 1/sqrt2,  ca = +-1/sqrt2, 0 on a==b
__global__ void unblock_kernel(const double* __restrict__ blocks, long long in_stride, double* __restrict__ M, SymLayout L, int batch) {
    const int D = L.dim * L.dim;
    const long long tot = (long long)batch * D * D;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < tot; i += (long long)gridDim.x * blockDim.x) {
        const int inst = (int)(i / ((long long)D * D));
        const int rc = (int)(i - (long long)inst * D * D), r = rc / D, c = rc - r * D;
        const int a = r / L.dim, b = r - a * L.dim, cc = c / L.dim, dd = c - cc * L.dim;
        const int rlo = a < b ? a : b, rhi = a < b ? b : a, clo = cc < dd ? cc : dd, chi = cc < dd ? dd : cc;
        const double* Bs = blocks + (long long)inst * in_stride;
        const double h = 0.70710678118654752440;
        double v = ((a == b) ? 1.0 : h) * ((cc == dd) ? 1.0 : h) * Bs[(long long)pair_s(rlo, rhi, L.dim) * L.Ds + pair_s(clo, chi, L.dim)];
        if (a != b && cc != dd) {
            const double w = 0.5 * Bs[L.off_a + (long long)pair_a(rlo, rhi, L.dim) * L.Da + pair_a(clo, chi, L.dim)];
            v += ((a < b) == (cc < dd)) ? w : -w;
        }
        M[i] = v;
    }
}

// ---- 2-site model in the block basis ------------------------------------------------------------------------------------
// element (r, c) of the swap-symmetric matrix represented by the blocks (same formula as unblock_kernel)
__device__ __forceinline__ double sym_at(const double* __restrict__ Bs, const SymLayout& L, int r, int c) {
    const int a = r / L.dim, b = r - a * L.dim, cc = c / L.dim, dd = c - cc * L.dim;
    const int rlo = a < b ? a : b, rhi = a < b ? b : a, clo = cc < dd ? cc : dd, chi = cc < dd ? dd : cc;
    const double h = 0.70710678118654752440;
    double v = ((a == b) ? 1.0 : h) * ((cc == dd) ? 1.0 : h) * Bs[(long long)pair_s(rlo, rhi, L.dim) * L.Ds + pair_s(clo, chi, L.dim)];
    if (a != b && cc != dd) {
        const double w = 0.5 * Bs[L.off_a + (long long)pair_a(rlo, rhi, L.dim) * L.Da + pair_a(clo, chi, L.dim)];
        v += ((a < b) == (cc < dd)) ? w : -w;
    }
    return v;
}

// embed_sym: blocks of Y_full = perm(Y_loc^{(x) N/2}) without materialising Y_full:
//   Ys[{ab},{cd}] = n_ab n_cd (Y[(a,b),(c,d)] + Y[(a,b),(d,c)]),   Ya[{ab},{cd}] = Y[(a,b),(c,d)] - Y[(a,b),(d,c)]
// (Y is swap symmetric).  grid = (ceil(NS / EMB_ROWS), items); thread = one column pair {cd}.
template <bool TANGENT, bool PRIMAL>
__global__ void __launch_bounds__(EMB_THREADS) embed_sym_kernel(const double* __restrict__ Yl, const double* __restrict__ Yld,
                                                                double* __restrict__ Y, double* __restrict__ Yd, EmbedTables T,
                                                                SymLayout L, const int* active, int m) {
    extern __shared__ __align__(16) double sm[];
    const int item = blockIdx.y;
    if (active != nullptr && active[item / m] == 0) return;
    const int layer = item % m, shift = layer & 1;
    const int q2 = T.q * T.q;
    double* yl = sm;
    double* yld = sm + q2;
    for (int i = threadIdx.x; i < q2; i += EMB_THREADS) {
        yl[i] = Yl[(long long)item * q2 + i];
        if (TANGENT) yld[i] = Yld[(long long)item * q2 + i];
    }
    __syncthreads();
    const unsigned* tab = T.tab + (size_t)shift * T.D;
    const int dim = L.dim;
    auto emb = [&](unsigned pr, unsigned pc, double& prim, double& tang) {
        double prod = 1.0, dsum = 0.0;
        for (int f = 0; f < T.nf; ++f) {
            const int e = loc_idx(pr, f) * T.q + loc_idx(pc, f);
            const double v = yl[e];
            if (TANGENT) dsum = dsum * v + prod * yld[e];
            prod *= v;
        }
        prim = prod; tang = dsum;
    };
    double* Ys = Y + (long long)item * L.slot;
    double* Ya = Ys + L.off_a;
    double* Yds = Yd + (long long)item * L.slot;
    double* Yda = Yds + L.off_a;
    const double rs2 = 0.70710678118654752440;
    for (int ci = threadIdx.x; ci < L.NS; ci += EMB_THREADS) {
        int c = 0, rem = ci;
        while (rem >= dim - c) { rem -= dim - c; ++c; }
        const int dd = c + rem;
        const unsigned p1 = tab[c * dim + dd], p2 = tab[dd * dim + c];
        for (int rr = 0; rr < EMB_ROWS; ++rr) {
            const int ri = blockIdx.x * EMB_ROWS + rr;
            if (ri >= L.NS) break;
            int a = 0, rm = ri;
            while (rm >= dim - a) { rm -= dim - a; ++a; }
            const int b = a + rm;
            const unsigned pr = tab[a * dim + b];
            double v1, t1, v2, t2;
            emb(pr, p1, v1, t1);
            emb(pr, p2, v2, t2);
            const double w = ((a == b) ? rs2 : 1.0) * ((c == dd) ? rs2 : 1.0);
            if (PRIMAL) Ys[(long long)ri * L.Ds + ci] = w * (v1 + v2);
            if (TANGENT) Yds[(long long)ri * L.Ds + ci] = w * (t1 + t2);
            if (a < b && c < dd) {
                const long long o = (long long)pair_a(a, b, dim) * L.Da + pair_a(c, dd, dim);
                if (PRIMAL) Ya[o] = v1 - v2;
                if (TANGENT) Yda[o] = t1 - t2;
            }
        }
    }
}

// Packed per-full-index lookup for sym_at: bits 0-11 pair_s, 12-23 pair_a, 24 diag (a == b), 25 lt (a < b); built on the host.
__host__ __device__ __forceinline__ unsigned sym_pack(int a, int b, int dim) {
    const int lo = a < b ? a : b, hi = a < b ? b : a;
    return (unsigned)pair_s(lo, hi, dim) | ((unsigned)(a != b ? pair_a(lo, hi, dim) : 0) << 12) | ((unsigned)(a == b) << 24) | ((unsigned)(a < b) << 25);
}
__device__ __forceinline__ double sym_at_packed(const double* __restrict__ Bs, const SymLayout& L, unsigned pr, unsigned pc) {
    const double h = 0.70710678118654752440;
    const bool dr = (pr >> 24) & 1, dc = (pc >> 24) & 1;
    double v = (dr ? 1.0 : h) * (dc ? 1.0 : h) * Bs[(long long)(pr & 0xfff) * L.Ds + (pc & 0xfff)];
    if (!dr && !dc) {
        const double w = 0.5 * Bs[L.off_a + (long long)((pr >> 12) & 0xfff) * L.Da + ((pc >> 12) & 0xfff)];
        v += (((pr >> 25) & 1) == ((pc >> 25) & 1)) ? w : -w;
    }
    return v;
}

// back_embed from cotangent blocks: identical to back_embed_kernel (assemble.cuh) with W[r,c] reconstructed from the blocks.
template <bool TANGENT>
__global__ void __launch_bounds__(BEM_THREADS) back_embed_sym_kernel(const double* __restrict__ W, const double* __restrict__ Wd,
                                                                     long long w_stride, long long wd_stride,
                                                                     const double* __restrict__ Yl, const double* __restrict__ Yld,
                                                                     double* __restrict__ Wl, double* __restrict__ Wld, EmbedTables T,
                                                                     SymLayout L, const unsigned* __restrict__ symtab, const int* active,
                                                                     int m, int layer) {
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
        const unsigned sr = symtab[r], sc = symtab[c];
        const double w = sym_at_packed(Wb, L, sr, sc);
        if (!TANGENT) acc = fma(w, prod, acc);
        if (TANGENT) { accd = fma(sym_at_packed(Wdb, L, sr, sc), prod, accd); accd = fma(w, dsum, accd); }
    }
    acc = warp_sum(acc);
    accd = warp_sum(accd);
    if (lane == 0) {
        if (!TANGENT) Wl[((long long)item * T.nf + f) * q2 + ij] = acc;
        else Wld[((long long)item * T.nf + f) * q2 + ij] = accd;
    }
}

// Fast back-embed for two Kronecker factors (N = 4) from cotangent blocks: one CTA per instance, thread = tensored column
// (j0, j1).  With T[i0][i1][j0][j1] = W[r(i0,i1)][c(j0,j1)]:
//   O_0[i0][j0] = sum_{i1,j1} T * Y[i1][j1]      (half-warp shuffle reduction over j1)
//   O_1[i1][j1] = sum_{i0,j0} T * Y[i0][j0]      (register accumulation over i0, one shared-memory reduction over j0 at the end)
// Each (row r, all columns) access stays inside one 1 KB row of Ws / Wa, so the gathers are sector-efficient.
template <bool TANGENT>
__global__ void __launch_bounds__(256) back_embed2_sym_kernel(const double* __restrict__ W, const double* __restrict__ Wd,
                                                              long long w_stride, long long wd_stride,
                                                              const double* __restrict__ Yl, const double* __restrict__ Yld,
                                                              double* __restrict__ Wl, double* __restrict__ Wld, SymLayout L,
                                                              const unsigned* __restrict__ symtab, const int* __restrict__ t2f,
                                                              const int* active, int m, int layer) {
    constexpr int Q = 16, NS = 136;             // q = d^4 = 16 (d = 2); D = 256; NS + NA = 256
    extern __shared__ __align__(16) double sm[];
    double* raw = sm;                           // [i1][256]: row r(i0,i1) of Ws (NS entries) followed by the row of Wa
    double* rawd = raw + Q * 256;               // same for Wd (TANGENT)
    double* red = rawd + (TANGENT ? Q * 256 : 0);   // [j0][i1][Q+1]
    __shared__ double yl[Q * Q], yld[Q * Q];
    __shared__ unsigned srow[Q * Q];
    const int inst = blockIdx.x;
    if (active != nullptr && active[inst] == 0) return;
    const int item = inst * m + layer, shift = layer & 1;
    const int t = threadIdx.x, j0 = t >> 4, j1 = t & 15;
    const int* map = t2f + shift * 256;
    yl[t] = Yl[(long long)item * 256 + t];
    if (TANGENT) yld[t] = Yld[(long long)item * 256 + t];
    srow[t] = symtab[map[t]];                   // packed pair info of the full index of tensored index t
    __syncthreads();
    const unsigned sc = srow[t];
    const int psc = sc & 0xfff, pac = (sc >> 12) & 0xfff;
    const bool dc = (sc >> 24) & 1, ltc = (sc >> 25) & 1;
    const double* Ws = W + (long long)inst * w_stride;
    const double* Wa = Ws + L.off_a;
    const double* Wds = TANGENT ? Wd + (long long)inst * wd_stride : nullptr;
    const double* Wda = TANGENT ? Wds + L.off_a : nullptr;
    const double h = 0.70710678118654752440;
    double o1[Q];
#pragma unroll
    for (int i = 0; i < Q; ++i) o1[i] = 0.0;
    double* out0 = (TANGENT ? Wld : Wl) + ((long long)item * 2 + 0) * 256;
    double* out1 = (TANGENT ? Wld : Wl) + ((long long)item * 2 + 1) * 256;
    for (int i0 = 0; i0 < Q; ++i0) {
        // stage the 16 rows r(i0, i1): coalesced, each global row read exactly once per CTA
        double v[Q], vd[Q];
#pragma unroll
        for (int i1 = 0; i1 < Q; ++i1) {
            const unsigned sr = srow[i0 * Q + i1];
            const bool dr = (sr >> 24) & 1;
            if (t < NS) {
                v[i1] = Ws[(long long)(sr & 0xfff) * L.Ds + t];
                if (TANGENT) vd[i1] = Wds[(long long)(sr & 0xfff) * L.Ds + t];
            } else {
                v[i1] = dr ? 0.0 : Wa[(long long)((sr >> 12) & 0xfff) * L.Da + (t - NS)];
                if (TANGENT) vd[i1] = dr ? 0.0 : Wda[(long long)((sr >> 12) & 0xfff) * L.Da + (t - NS)];
            }
        }
        __syncthreads();                        // previous chunk fully consumed
#pragma unroll
        for (int i1 = 0; i1 < Q; ++i1) { raw[i1 * 256 + t] = v[i1]; if (TANGENT) rawd[i1 * 256 + t] = vd[i1]; }
        __syncthreads();
        const double y00 = yl[i0 * Q + j0], yd00 = TANGENT ? yld[i0 * Q + j0] : 0.0;
        double p = 0.0;
#pragma unroll
        for (int i1 = 0; i1 < Q; ++i1) {
            const unsigned sr = srow[i0 * Q + i1];
            const bool dr = (sr >> 24) & 1;
            const double cf = (dr ? 1.0 : h) * (dc ? 1.0 : h);
            double w = cf * raw[i1 * 256 + psc], wd = 0.0;
            if (TANGENT) wd = cf * rawd[i1 * 256 + psc];
            if (!dr && !dc) {
                const double sg = ((((sr >> 25) & 1) != 0) == ltc) ? 0.5 : -0.5;
                w = fma(sg, raw[i1 * 256 + NS + pac], w);
                if (TANGENT) wd = fma(sg, rawd[i1 * 256 + NS + pac], wd);
            }
            if (!TANGENT) {
                p = fma(w, yl[i1 * Q + j1], p);
                o1[i1] = fma(w, y00, o1[i1]);
            } else {
                p = fma(wd, yl[i1 * Q + j1], p); p = fma(w, yld[i1 * Q + j1], p);
                o1[i1] = fma(wd, y00, o1[i1]); o1[i1] = fma(w, yd00, o1[i1]);
            }
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o, 16);
        if (j1 == 0) out0[i0 * Q + j0] = p;
    }
    __syncthreads();
#pragma unroll
    for (int i1 = 0; i1 < Q; ++i1) red[(j0 * Q + i1) * (Q + 1) + j1] = o1[i1];
    __syncthreads();
    {
        const int i1 = t >> 4, jj = t & 15;
        double s = 0.0;
        for (int a = 0; a < Q; ++a) s += red[(a * Q + i1) * (Q + 1) + jj];
        out1[i1 * Q + jj] = s;
    }
}

}  // namespace otn
