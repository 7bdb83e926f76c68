// Factored ("never materialise Y_full") evaluation of the 2-site model (SURVEY 8f "next" #3).
//
// The reference builds every layer as a D x D matrix, Y_full = perm . (Y_loc (x) ... (x) Y_loc) . perm^T
// (create_supertensored_from_local transformations.py:405-429, convert_supertensored2liouvillianfull :484-490) and multiplies
// the layers (compose_superops_list :841-851): 2 D^3 flops per product.  Here a layer is never formed: it is applied to a state
// one 16 x 16 factor at a time (2 * 16 * D flops per column and factor, 8x fewer at N = 4, 85x fewer at N = 6), and because a
// layer acts on the ROW index only, every column of the model matrix M = L_m ... L_1 evolves independently:
//
//   one CTA owns a tile of NCOLS columns c and does the whole evaluation for them in shared memory --
//     forward    S <- L_l S           starting from identity columns (states before each layer go to a CTA-private scratch)
//     residual   G  = w_c (S - Ts[:, c]),   ||R||^2 partial
//     reverse    for l = m-1 .. 0:  W_l^{(f)} += contract_f(B_f, F_{f+1}),  G <- L_l^T G
//                (B_f = G with the transposed factors < f applied, F_{f+1} = P_{l-1} with the factors > f applied)
//     tangent    every application / contraction in forward mode (dual numbers) for the Hessian-vector product.
//
// Swap symmetry (move 2) halves the work again: column (c2,c1) of every matrix on the path is the row-swapped
// column (c1,c2), so only the sdim (sdim+1)/2 columns with c1 <= c2 are evolved, with weight 1 (1/2 on the diagonal) in the
// cotangent; the reduction kernel adds the swapped partner, W = H + swap(H).
//
// All 16 x 16 applications and contractions run on the FP64 tensor pipe (DMMA.8x8x4): the local superoperator is the A operand
// (fragments in registers for a whole pass over the state), state vectors are the B operand, gathered from shared memory through
// two small index tables (local digit -> row offset, rest index -> row offset) derived from the reference's permutation.
#pragma once
#include "otn_common.cuh"

namespace otn {

enum { FACT_COST = 0, FACT_GRAD = 1, FACT_HVP = 2, FACT_MODEL = 3, FACT_HVPONLY = 4 };   // HVPONLY: tangent cotangents only (tCG)
// threads per CTA: one warp per two n-tiles of a state -- 8 warps at N = 4 with 8 columns (two CTAs per SM), 2 warps with the
// 2-column tiles used for a handful of instances (4x more CTAs), 16 warps at N = 6 (one CTA per SM: 192 KB of states)
template <int NF, int NCOLS> struct FactThreads {
    static constexpr int THREADS = (NF == 3) ? 512 : 32 * NCOLS;
    static constexpr int WARPS = THREADS / 32;
};

struct FactParams {
    const double* Yl; const double* Yld;   // [B][m][256] local superoperators and tangents
    const double* Tt;                      // [n_targets][tiles][D * NCOLS] symmetrised target, tile-column layout
    const int* tindex; const int* active;
    const int* jrow;                       // [2][NF][16]      local index of factor f -> row offset
    const int* restrow;                    // [2][NF][D / 16]  rest index (factor digits zero) -> row
    const int* colidx;                     // [tiles * NCOLS]  full column index of every tile column
    const double* colw;                    // [tiles * NCOLS]  cotangent weight (1, 1/2 on the swap diagonal, 0 for padding)
    double* Pst;                           // [B][tiles][m-1][2][D * NCOLS] states entering layers 1..m-1 (and tangents)
    double* Wp;                            // [B][tiles][m][NF][2][256] per-tile local cotangents (and tangents)
    double* part_f; double* part_s; int part_stride;
    double* Mout; const int* swaprow;      // FACT_MODEL: dense [B][D][D] output, swap of a full index
    int m, tiles;
};

template <int NF> struct FactDims { static constexpr int D = (NF == 2) ? 256 : 4096; static constexpr int REST = D / 16; };

// Shared-memory layout of a state: element (row r, column c) lives at (r * NCOLS + c) ^ fact_swz(r).  The XOR mask is linear in
// the bits of r and chosen so that the 16 addresses a half-warp touches in every fragment load -- 4 rows that differ in the two
// low bits of a factor's local index (bra bits of the pair's sites, wherever the permutation puts them) x 4 consecutive columns /
// rest indices -- fall into 16 different 8-byte banks (ncu before the swizzle: 5-way conflicts, L1 pipe 84 % busy).
template <int NF, int NCOLS>
__device__ __forceinline__ int fact_swz(int r) {
    if (NF == 2 && NCOLS == 8) return ((((r >> 1) ^ (r >> 2) ^ (r >> 3)) & 1) << 3) | ((((r >> 1) ^ (r >> 3)) & 1) << 2);
    if (NF == 3) return ((r >> 4) & 3) << 2;                                                               // NCOLS == 1
    return 0;   // narrow tiles (latency regime, not shared-memory bound): plain layout
}

constexpr int FACT_YLD = 20;   // row pitch of the staged 16 x 16 local superoperators (+4: conflict-free fragment reads)

// A-operand fragments of the local superoperator (or its transpose), and of its tangent: lane holds A[8 mt + l/4][4 ks + l%4]
template <bool DUAL>
__device__ __forceinline__ void fact_frags(const double* __restrict__ Ysm, const double* __restrict__ Ydsm, bool trans,
                                           double (&a)[2][4], double (&ad)[2][4]) {
    const int lane = threadIdx.x & 31, lr = lane >> 2, lc = lane & 3;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const int r = 8 * mt + lr, c = 4 * ks + lc;
            const int idx = trans ? c * FACT_YLD + r : r * FACT_YLD + c;
            a[mt][ks] = Ysm[idx];
            ad[mt][ks] = DUAL ? Ydsm[idx] : 0.0;
        }
}

// dst = A src (and dstd = Ad src + A srcd) on one factor; A given as fragments.  May run in place: a warp reads all 16 entries
// of its vectors before it writes them.  Two n-tiles (16 vectors) per step for instruction-level parallelism.
template <int NF, int NCOLS, bool DUAL>
__device__ __forceinline__ void fact_apply(const double* src, const double* srcd, double* dst, double* dstd,
                                           const double (&a)[2][4], const double (&ad)[2][4],
                                           const int* __restrict__ jrow, const int* __restrict__ restrow) {
    constexpr int NV = FactDims<NF>::REST * NCOLS, NT = NV / 8, U = 2;
    constexpr int FACT_WARPS = FactThreads<NF, NCOLS>::WARPS;
    static_assert(NT % (FACT_WARPS * U) == 0, "n-tiles must split evenly over the warps");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, lr = lane >> 2, lc = lane & 3;
    int jb[4], sb[4], jc[2], sc[2];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) { const int j = jrow[4 * ks + lc]; jb[ks] = j * NCOLS; sb[ks] = fact_swz<NF, NCOLS>(j); }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) { const int j = jrow[8 * mt + lr]; jc[mt] = j * NCOLS; sc[mt] = fact_swz<NF, NCOLS>(j); }
#pragma unroll 1
    for (int nt0 = warp * U; nt0 < NT; nt0 += FACT_WARPS * U) {
        double b[U][4], bd[U][4];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int v = 8 * (nt0 + u) + lr;
            const int rr = restrow[v / NCOLS];
            const int base = rr * NCOLS + (v % NCOLS), sr = fact_swz<NF, NCOLS>(rr);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const int a_ = (base + jb[ks]) ^ sr ^ sb[ks];
                b[u][ks] = src[a_];
                if (DUAL) bd[u][ks] = srcd[a_];
            }
        }
        double c[U][2][2], cd[U][2][2], ce[U][2][2];      // cd: Ad b, ce: A bd (separate dependency chains)
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) { c[u][mt][0] = c[u][mt][1] = 0.0; cd[u][mt][0] = cd[u][mt][1] = 0.0; ce[u][mt][0] = ce[u][mt][1] = 0.0; }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    dmma884(c[u][mt][0], c[u][mt][1], a[mt][ks], b[u][ks]);
                    if (DUAL) {
                        dmma884(cd[u][mt][0], cd[u][mt][1], ad[mt][ks], b[u][ks]);
                        dmma884(ce[u][mt][0], ce[u][mt][1], a[mt][ks], bd[u][ks]);
                    }
                }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int v0 = 8 * (nt0 + u) + 2 * lc, v1 = v0 + 1;
            const int r0 = restrow[v0 / NCOLS], r1 = restrow[v1 / NCOLS];
            const int b0 = r0 * NCOLS + (v0 % NCOLS), b1 = r1 * NCOLS + (v1 % NCOLS);
            const int s0 = fact_swz<NF, NCOLS>(r0), s1 = fact_swz<NF, NCOLS>(r1);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                if (NCOLS % 2 == 0) {               // both columns of the fragment sit in the same row: one 16-byte store
                    const int a_ = (b0 + jc[mt]) ^ s0 ^ sc[mt];
                    *reinterpret_cast<double2*>(dst + a_) = make_double2(c[u][mt][0], c[u][mt][1]);
                    if (DUAL) *reinterpret_cast<double2*>(dstd + a_) = make_double2(cd[u][mt][0] + ce[u][mt][0], cd[u][mt][1] + ce[u][mt][1]);
                } else {
                    const int a0 = (b0 + jc[mt]) ^ s0 ^ sc[mt], a1 = (b1 + jc[mt]) ^ s1 ^ sc[mt];
                    dst[a0] = c[u][mt][0];
                    dst[a1] = c[u][mt][1];
                    if (DUAL) { dstd[a0] = cd[u][mt][0] + ce[u][mt][0]; dstd[a1] = cd[u][mt][1] + ce[u][mt][1]; }
                }
            }
        }
    }
}

// acc[i][j] += sum_v X[i, v] Z[j, v]   (and accd += Xd Z + X Zd) over the vectors of factor `f` handled by this warp
template <int NF, int NCOLS, bool DUAL, bool PRIMAL>
__device__ __forceinline__ void fact_contract(const double* X, const double* Xd, const double* Z, const double* Zd,
                                              const int* __restrict__ jrow, const int* __restrict__ restrow,
                                              double (&acc)[2][2][2], double (&accd)[2][2][2]) {
    constexpr int NV = FactDims<NF>::REST * NCOLS, FACT_WARPS = FactThreads<NF, NCOLS>::WARPS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, lr = lane >> 2, lc = lane & 3;
    const int j0 = jrow[lr], j1 = jrow[8 + lr];
    const int ro[2] = {j0 * NCOLS, j1 * NCOLS}, so[2] = {fact_swz<NF, NCOLS>(j0), fact_swz<NF, NCOLS>(j1)};
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) { acc[mt][nt][0] = acc[mt][nt][1] = 0.0; accd[mt][nt][0] = accd[mt][nt][1] = 0.0; }
    for (int kv = warp; kv < NV / 4; kv += FACT_WARPS) {
        const int v = 4 * kv + lc;
        const int rr = restrow[v / NCOLS];
        const int base = rr * NCOLS + (v % NCOLS), sr = fact_swz<NF, NCOLS>(rr);
        double x[2], z[2], xd[2], zd[2];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int a_ = (base + ro[t]) ^ sr ^ so[t];
            x[t] = X[a_];
            z[t] = Z[a_];
            if (DUAL) { xd[t] = Xd[a_]; zd[t] = Zd[a_]; }
        }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                if (PRIMAL) dmma884(acc[mt][nt][0], acc[mt][nt][1], x[mt], z[nt]);
                if (DUAL) {
                    dmma884(accd[mt][nt][0], accd[mt][nt][1], xd[mt], z[nt]);
                    dmma884(accd[mt][nt][0], accd[mt][nt][1], x[mt], zd[nt]);
                }
            }
    }
}

template <int NF, int NCOLS, int MODE>
struct FactCfg {
    static constexpr int D = FactDims<NF>::D;
    static constexpr int SB = D * NCOLS;                                   // doubles per state buffer
    static constexpr int NBUF = (MODE == FACT_HVP || MODE == FACT_HVPONLY) ? 6 : (MODE == FACT_GRAD ? 3 : 1);
    static constexpr int TAB_INTS = 2 * NF * 16 + 2 * NF * (D / 16);
    static constexpr size_t SMEM_BYTES = (size_t)NBUF * SB * 8 + 2 * 16 * FACT_YLD * 8 + (size_t)TAB_INTS * 4 + 64 * 8;
};

template <int NF, int NCOLS, int MODE>
__global__ void __launch_bounds__(FactThreads<NF, NCOLS>::THREADS, (NF == 2 ? 2 : 1)) fact_kernel(const FactParams p) {
    using Cfg = FactCfg<NF, NCOLS, MODE>;
    constexpr int FACT_THREADS = FactThreads<NF, NCOLS>::THREADS, FACT_WARPS = FactThreads<NF, NCOLS>::WARPS;
    constexpr int D = Cfg::D, SB = Cfg::SB, REST = D / 16;
    constexpr bool DUAL = MODE == FACT_HVP || MODE == FACT_HVPONLY;
    constexpr bool REV = MODE == FACT_GRAD || DUAL;
    constexpr bool PRIMALW = MODE != FACT_HVPONLY;
    const int tile = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    if (p.active != nullptr && p.active[b] == 0) return;
    extern __shared__ __align__(16) double fsm[];
    // buffers: G (= forward state S), [Gd], P, [Pd], tmp, [tmpd]
    double* G = fsm;
    double* Gd = DUAL ? G + SB : nullptr;
    double* P = REV ? G + (DUAL ? 2 : 1) * SB : nullptr;
    double* Pd = DUAL ? P + SB : nullptr;
    double* tmp = REV ? P + (DUAL ? 2 : 1) * SB : nullptr;
    double* tmpd = DUAL ? tmp + SB : nullptr;
    double* Ysm = fsm + (size_t)Cfg::NBUF * SB;
    double* Ydsm = Ysm + 16 * FACT_YLD;
    double* red = Ydsm + 16 * FACT_YLD;             // 64 doubles
    int* jrow_s = reinterpret_cast<int*>(red + 64); // [2][NF][16]
    int* rest_s = jrow_s + 2 * NF * 16;             // [2][NF][REST]
    for (int i = tid; i < 2 * NF * 16; i += FACT_THREADS) jrow_s[i] = p.jrow[i];
    for (int i = tid; i < 2 * NF * REST; i += FACT_THREADS) rest_s[i] = p.restrow[i];
    const int* cidx = p.colidx + (size_t)tile * NCOLS;
    const double* cw = p.colw + (size_t)tile * NCOLS;
    // identity columns
    for (int i = tid; i < SB; i += FACT_THREADS) {
        const int r = i / NCOLS, c = i % NCOLS;
        G[i ^ fact_swz<NF, NCOLS>(r)] = (r == cidx[c]) ? 1.0 : 0.0;
        if (DUAL) Gd[i] = 0.0;
    }
    const double* Yb = p.Yl + (size_t)b * p.m * 256;
    const double* Ydb = DUAL ? p.Yld + (size_t)b * p.m * 256 : nullptr;
    double* scratch = REV ? p.Pst + ((size_t)b * p.tiles + tile) * (size_t)(p.m - 1) * 2 * SB : nullptr;

    // ---- forward ----------------------------------------------------------------------------------------------------
    for (int l = 0; l < p.m; ++l) {
        __syncthreads();                            // previous layer done with Ysm; state complete
        for (int t = tid; t < 256; t += FACT_THREADS) {
            Ysm[(t >> 4) * FACT_YLD + (t & 15)] = Yb[(size_t)l * 256 + t];
            if (DUAL) Ydsm[(t >> 4) * FACT_YLD + (t & 15)] = Ydb[(size_t)l * 256 + t];
        }
        if (REV && l >= 1) {                        // state entering layer l
            if (l == p.m - 1) {                     // the reverse sweep starts here: keep it in shared memory
                for (int i = tid * 2; i < SB; i += FACT_THREADS * 2) {
                    *reinterpret_cast<double2*>(P + i) = *reinterpret_cast<const double2*>(G + i);
                    if (DUAL) *reinterpret_cast<double2*>(Pd + i) = *reinterpret_cast<const double2*>(Gd + i);
                }
            } else {
                double* dstp = scratch + (size_t)(l - 1) * 2 * SB;
                for (int i = tid * 2; i < SB; i += FACT_THREADS * 2) {
                    *reinterpret_cast<double2*>(dstp + i) = *reinterpret_cast<const double2*>(G + i);
                    if (DUAL) *reinterpret_cast<double2*>(dstp + SB + i) = *reinterpret_cast<const double2*>(Gd + i);
                }
            }
        }
        __syncthreads();
        const int sh = l & 1;
        double a[2][4], ad[2][4];
        fact_frags<DUAL>(Ysm, Ydsm, false, a, ad);
#pragma unroll 1
        for (int g = NF - 1; g >= 0; --g) {
            fact_apply<NF, NCOLS, DUAL>(G, Gd, G, Gd, a, ad, jrow_s + (sh * NF + g) * 16, rest_s + (sh * NF + g) * REST);
            __syncthreads();
        }
    }
    if (MODE == FACT_MODEL) {
        double* Mb = p.Mout + (size_t)b * D * D;
        for (int i = tid; i < SB; i += FACT_THREADS) {
            const int r = i / NCOLS, c = i % NCOLS;
            if (cw[c] == 0.0) continue;
            const int col = cidx[c];
            const double val = G[i ^ fact_swz<NF, NCOLS>(r)];
            Mb[(size_t)r * D + col] = val;
            Mb[(size_t)p.swaprow[r] * D + p.swaprow[col]] = val;
        }
        return;
    }
    // ---- residual ---------------------------------------------------------------------------------------------------
    {
        const double* Tt = p.Tt + ((size_t)(p.tindex ? p.tindex[b] : 0) * p.tiles + tile) * SB;
        double sf = 0.0, ss = 0.0;
        for (int i = tid; i < SB; i += FACT_THREADS) {
            const double w = cw[i % NCOLS];
            const int ip = i ^ fact_swz<NF, NCOLS>(i / NCOLS);
            const double r = G[ip] - Tt[i];
            sf = fma(2.0 * w * r, r, sf);
            if (REV) G[ip] = w * r;
            if (DUAL) { const double sd = Gd[ip]; ss = fma(2.0 * w * r, sd, ss); Gd[ip] = w * sd; }
        }
        sf = warp_sum(sf);
        if (DUAL) ss = warp_sum(ss);
        if ((tid & 31) == 0) { red[tid >> 5] = sf; red[16 + (tid >> 5)] = ss; }
        __syncthreads();
        if (tid == 0) {
            double tf = 0.0, ts = 0.0;
            for (int w = 0; w < FACT_WARPS; ++w) { tf += red[w]; ts += red[16 + w]; }
            p.part_f[(size_t)b * p.part_stride + tile] = tf;
            if (DUAL) p.part_s[(size_t)b * p.part_stride + tile] = ts;
        }
    }
    if (!REV) return;
    // ---- reverse ----------------------------------------------------------------------------------------------------
    double* wp = p.Wp + ((size_t)b * p.tiles + tile) * (size_t)p.m * NF * 512;
    const int lane = tid & 31, warp = tid >> 5, lr = lane >> 2, lc = lane & 3;
#pragma unroll 1
    for (int l = p.m - 1; l >= 0; --l) {
        __syncthreads();
        for (int t = tid; t < 256; t += FACT_THREADS) {
            Ysm[(t >> 4) * FACT_YLD + (t & 15)] = Yb[(size_t)l * 256 + t];
            if (DUAL) Ydsm[(t >> 4) * FACT_YLD + (t & 15)] = Ydb[(size_t)l * 256 + t];
        }
        if (l >= 1) {
            // P / Pd already hold the state entering layer l: left there by the forward pass (l = m-1) or prefetched with
            // cp.async during the tail of layer l+1
            cp_async_wait<0>();
        } else {
            for (int i = tid; i < SB; i += FACT_THREADS) {
                P[i ^ fact_swz<NF, NCOLS>(i / NCOLS)] = ((i / NCOLS) == cidx[i % NCOLS]) ? 1.0 : 0.0;
                if (DUAL) Pd[i] = 0.0;
            }
        }
        __syncthreads();
        const int sh = l & 1;
#pragma unroll 1
        for (int f = 0; f < NF; ++f) {
            const double* Fs = P; const double* Fds = Pd;
            if (f < NF - 1) {                       // tmp = (factors NF-1 .. f+1) applied to P
                double a[2][4], ad[2][4];
                fact_frags<DUAL>(Ysm, Ydsm, false, a, ad);
#pragma unroll 1
                for (int g = NF - 1; g > f; --g) {
                    const bool first = (g == NF - 1);
                    fact_apply<NF, NCOLS, DUAL>(first ? P : tmp, first ? Pd : tmpd, tmp, tmpd, a, ad,
                                                jrow_s + (sh * NF + g) * 16, rest_s + (sh * NF + g) * REST);
                    __syncthreads();
                }
                Fs = tmp; Fds = tmpd;
            }
            double acc[2][2][2], accd[2][2][2];
            fact_contract<NF, NCOLS, DUAL, PRIMALW>(G, Gd, Fs, Fds, jrow_s + (sh * NF + f) * 16, rest_s + (sh * NF + f) * REST, acc, accd);
            __syncthreads();                        // tmp no longer read: reuse it for the per-warp partials
            if (f == NF - 1 && l >= 2) {            // P / Pd are free now: prefetch the state entering layer l-1
                const double* srcp = scratch + (size_t)(l - 2) * 2 * SB;
                for (int i = tid * 2; i < SB; i += FACT_THREADS * 2) {
                    cp_async16(P + i, srcp + i);
                    if (DUAL) cp_async16(Pd + i, srcp + SB + i);
                }
                cp_async_commit();
            }
            double* part = tmp + (size_t)warp * (DUAL ? 512 : 256);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        // column XOR-swizzled by the row so that the 8 rows of a fragment store hit different banks
                        const int idx = (8 * mt + lr) * 16 + ((8 * nt + 2 * lc + e) ^ ((lr & 1) | ((lr & 2) << 2)));
                        if (PRIMALW) part[idx] = acc[mt][nt][e];
                        if (DUAL) part[256 + idx] = accd[mt][nt][e];
                    }
            __syncthreads();
            for (int t = tid; t < 256; t += FACT_THREADS) {
                double s = 0.0, sd = 0.0;
                const int row = t >> 4, ti = (t & ~15) | ((t & 15) ^ ((row & 1) | ((row & 2) << 2)));
#pragma unroll
                for (int w = 0; w < FACT_WARPS; ++w) {
                    if (PRIMALW) s += tmp[(size_t)w * (DUAL ? 512 : 256) + ti];
                    if (DUAL) sd += tmp[(size_t)w * 512 + 256 + ti];
                }
                double* o = wp + ((size_t)l * NF + f) * 512;
                if (PRIMALW) o[t] = s;
                if (DUAL) o[256 + t] = sd;
            }
            __syncthreads();
            if (!(l == 0 && f == NF - 1)) {         // B_{f+1} = A_f^T B_f
                double at[2][4], adt[2][4];
                fact_frags<DUAL>(Ysm, Ydsm, true, at, adt);
                fact_apply<NF, NCOLS, DUAL>(G, Gd, G, Gd, at, adt, jrow_s + (sh * NF + f) * 16, rest_s + (sh * NF + f) * REST);
                __syncthreads();
            }
        }
    }
}

// first stage of the tile reduction when there are many tiles (N = 6: 2080): out[b][z][lf] = sum of the tiles of chunk z
__global__ void fact_partial_reduce_kernel(const double* __restrict__ Wp, double* __restrict__ out, int tiles, int chunk, int mnf,
                                           bool primal, bool dual, const int* active) {
    const int lf = blockIdx.x, b = blockIdx.y, z = blockIdx.z, t = threadIdx.x, nsplit = gridDim.z;
    if (active != nullptr && active[b] == 0) return;
    const int t0 = z * chunk, t1 = min(tiles, t0 + chunk);
    double s = 0.0, sd = 0.0;
    for (int tile = t0; tile < t1; ++tile) {
        const double* src = Wp + (((size_t)b * tiles + tile) * mnf + lf) * 512;
        if (primal) s += src[t];
        if (dual) sd += src[256 + t];
    }
    double* o = out + (((size_t)b * nsplit + z) * mnf + lf) * 512;
    o[t] = s;
    o[256 + t] = sd;
}

// Wl[b][l][f] = H + swap(H), H = sum over tiles of the per-tile partials; swap: local index (a,b) -> (b,a) on rows and columns
__global__ void fact_reduce_kernel(const double* __restrict__ Wp, double* __restrict__ Wl, double* __restrict__ Wld, int tiles, int m,
                                   int nf, bool primal, bool dual, const int* active) {
    const int lf = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
    if (active != nullptr && active[b] == 0) return;
    __shared__ double H[512];
    double s = 0.0, sd = 0.0;
    for (int tile = 0; tile < tiles; ++tile) {
        const double* src = Wp + (((size_t)b * tiles + tile) * m * nf + lf) * 512;
        if (primal) s += src[t];
        if (dual) sd += src[256 + t];
    }
    H[t] = s; H[256 + t] = sd;
    __syncthreads();
    const int i = t >> 4, j = t & 15;
    const int ts = (((i & 3) << 2) | (i >> 2)) * 16 + (((j & 3) << 2) | (j >> 2));
    const size_t o = ((size_t)b * m * nf + lf) * 256 + t;
    if (primal) Wl[o] = H[t] + H[ts];
    if (dual) Wld[o] = H[256 + t] + H[256 + ts];
}

}  // namespace otn
