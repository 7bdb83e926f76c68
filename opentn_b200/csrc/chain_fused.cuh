// Fused per-instance chain: ONE CTA owns one diagonal block (144 x 144 symmetric or 128 x 128 antisymmetric at N = 4) of one
// instance and executes the WHOLE sequence of chain products of a unit of work on it -- forward products, residual, reverse
// sweep and the forward-over-reverse tangent sweep: 18 products at m = 3 -- instead of 24 batched GEMM launches whose CTAs each
// own a 48 x 48 / 64 x 64 tile of one product.
//
// Replaces: `compose_superops_list` (opentn/transformations.py:841-851) and the products jax.grad / jax.jvp generate from it
// (opentn/stiefel.py:423, 520), like gemm_dmma.cuh, which stays for every shape this kernel does not take.
//
// Why (ncu of the tiled kernel, profiles/ncu_struct_r01.txt): with a 48 x 48 tile every DMMA k-step needs 3 KB of operands from
// L2 per 36 DMMA -- 21 B/clk/SM, 6 TB/s chip-wide at the DMMA issue rate -- and a CTA lives for only 9 k-tiles, so prologue,
// epilogue and barrier drain are a fifth of its life: tensor pipe 72.6 % active.  Here the CTA tile is the whole block: every
// operand element is read once per product (7 B/clk/SM), the accumulators of the 144 x 144 result live in the registers of 12
// warps (72 x 24 per warp: 27 DMMA per 12 fragment loads), and the cp.async ring never drains: the products of a unit of work are
// ordered so that no product reads the result of its predecessor (primal and tangent sweeps interleaved), which lets the tile
// loads of product i+1 start under the last k-tiles and the epilogue of product i.
//
// Memory: intermediates (P, Pd, G, Gd) are written to global memory and read back by the same CTA a few microseconds later (L2
// hits for the most recent ones); ordering = __threadfence() after the epilogue stores + the __syncthreads() every k-tile starts
// with, and at least one whole product lies between a store and the first cp.async that reads it.
#pragma once
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "gemm_dmma.cuh"

namespace otn {

enum { CH_Y = 0, CH_P, CH_G, CH_W, CH_YD, CH_PD, CH_GD, CH_WD, CH_NARR };
enum { CHF_TA = 1, CHF_TB = 2, CHF_RESID = 4, CHF_DOT = 8, CHF_DEP = 16 };   // CHF_DEP: reads the result of the previous op
constexpr int CHAIN_MAXOPS = 96;
constexpr unsigned char CH_NONE = 0xff;

// operand = (array << 4) | layer
struct ChainOp { unsigned char a1, b1, a2, b2, c, e, flags, pad; };

struct ChainParams {
    double* base[CH_NARR];          // work arrays of the handle (already shifted to the first instance of this launch and to the block)
    long long stride[CH_NARR];      // per-instance stride (doubles)
    long long lstride[CH_NARR];     // per-layer stride (doubles)
    int D, kvalid;                  // block side (leading dimension) and populated contraction extent
    const double* T; long long sT; const int* tindex;     // targets (this block), stride per target, target id per instance
    double* part_f; double* part_s; int part_stride;      // partial[b * part_stride]: ||R||^2 and <R, Mdot> of this block
    const int* active;
    int nops;
    ChainOp ops[CHAIN_MAXOPS];
};

template <int DT, int WM, int WN, int STAGES>
struct ChainCfg {
    static constexpr int BK = 16;
    static constexpr int WARPS_M = DT / WM, WARPS_N = DT / WN;
    static constexpr int CONS_WARPS = WARPS_M * WARPS_N;          // consumer (MMA) warps: whole warpgroups
    static constexpr int PROD_WARPS = 4, PROD_THREADS = PROD_WARPS * 32;      // one producer warpgroup (cp.async)
    static constexpr int THREADS = (CONS_WARPS + PROD_WARPS) * 32;
    // registers: the producers give theirs up (setmaxnreg), the consumers take what the 64 K register file leaves
    // setmaxnreg moves registers inside the pool the CTA was LAUNCHED with (launch registers x threads; what the SM has beyond that
    // is not part of it -- asking for more blocks forever): the consumers get what the producers give up
    static constexpr int PROD_REGS = 56;
    static constexpr int LAUNCH_REGS_RAW = 65536 / THREADS / 8 * 8;
    static constexpr int LAUNCH_REGS = LAUNCH_REGS_RAW > 255 ? 248 : LAUNCH_REGS_RAW;
    static constexpr int CONS_REGS_RAW = LAUNCH_REGS + PROD_THREADS * (LAUNCH_REGS - PROD_REGS) / (CONS_WARPS * 32) / 8 * 8;
    static constexpr int CONS_REGS = CONS_REGS_RAW > 232 ? 232 : CONS_REGS_RAW;
    static constexpr int LDK = BK + 4;            // k-contiguous source:  [DT][BK + 4]
    static constexpr int LDM = DT + 4;            // m/n-contiguous source: [BK][DT + 4]
    static constexpr int OPER = DT * LDK;         // >= BK * LDM for DT >= 80
    static constexpr int STAGE = 2 * OPER;
    static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE * sizeof(double) + 2 * STAGES * sizeof(unsigned long long) + 16;
    // warp id -> warp tile (row-major in the WARPS_M x WARPS_N grid), balancing the trimmed tile counts over warp id mod 4
    // (tables packed 4 bits per warp into one constant: {0,2,3,4, 1,6,8,10, 5,7,9,11} and {0,1,2,5, 7,3,4,6})
    __host__ __device__ static constexpr int tile_of_warp(int w) {
        if (DT == 144 && WM == 72 && WN == 24) return (int)((0xb975a8614320ULL >> (4 * w)) & 15);
        if (DT == 128 && WM == 64 && WN == 32) return (int)((0x64375210ULL >> (4 * w)) & 15);
        // 16 warps of 32 x 32 on the 120-wide block: {16,16,12,12} MMA tiles on three sub-partitions, {16,16,16,9} on the fourth
        // (57 of 225 at most; with 8 warps of 64 x 32 the best deal is 60)
        if (DT == 128 && WM == 32 && WN == 32) return (int)((0xfb73aedc96418520ULL >> (4 * w)) & 15);
        return w;
    }
    static_assert(DT % WM == 0 && DT % WN == 0 && WM % 8 == 0 && WN % 8 == 0, "warp tiles must tile the block in 8 x 8 MMA tiles");
    static_assert(CONS_WARPS % 4 == 0, "setmaxnreg works on whole warpgroups");
    static_assert(DT * LDK >= BK * LDM, "operand slot too small for the transposed layout");
    static_assert(PROD_THREADS == 128 && BK == 16 && DT % 16 == 0, "producer copy loops: 16 rows x 8 chunk columns of threads");
};

// Warp-specialised: CONS_WARPS consumer warps own the accumulators and issue nothing but fragment loads and DMMA; one producer
// warpgroup streams the operand tiles of the whole op program through a STAGES-deep shared-memory ring with cp.async.  Hand-off
// by mbarriers (full[s]: the producers' copies of slot s have landed; empty[s]: every consumer warp has read slot s) -- there is no
// CTA-wide barrier anywhere, so the warps drift apart: while one warp stores its part of product i (epilogue), the others keep
// the tensor pipe busy with product i+1, whose tiles are already in the ring.  (The first version of this kernel -- every warp
// loads and computes, one __syncthreads per k-tile -- left the tensor pipe 27 % idle: ncu stall reasons barrier 10 %, short
// scoreboard 5 %, membar / lg_throttle 4 %, profiles/ncu_chain_fused_v1_r02.txt.)
template <int DT, int WM, int WN, int STAGES>
__global__ void __launch_bounds__(ChainCfg<DT, WM, WN, STAGES>::THREADS, 1) chain_fused_kernel(const __grid_constant__ ChainParams p) {
    using Cfg = ChainCfg<DT, WM, WN, STAGES>;
    constexpr int BK = Cfg::BK, MI = WM / 8, NI = WN / 8, LDK = Cfg::LDK, LDM = Cfg::LDM;
    extern __shared__ __align__(16) double smem[];
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem + STAGES * Cfg::STAGE);
    unsigned long long* empty = full + STAGES;
    volatile int* ops_done = reinterpret_cast<volatile int*>(empty + STAGES);     // consumer-warp completions, all ops

    const int b = blockIdx.x;
    if (p.active != nullptr && p.active[b] == 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int D = DT;                        // the block side is the leading dimension (chain_launch checks p.D == DT)
    const int ktiles_one = (p.kvalid + BK - 1) / BK;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], Cfg::PROD_THREADS); mbar_init(&empty[s], Cfg::CONS_WARPS); }
        *ops_done = 0;
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    auto mat = [&](unsigned char code) -> double* {
        const int arr = code >> 4, layer = code & 15;
        return p.base[arr] + (long long)b * p.stride[arr] + (long long)layer * p.lstride[arr];
    };

    if (warp >= Cfg::CONS_WARPS) {
        // ================================================ producers ================================================
        reg_dealloc<Cfg::PROD_REGS>();
        const int pt = tid - Cfg::CONS_WARPS * 32;
        const int r0 = pt >> 3, c0 = pt & 7;
        auto copy_kc = [&](double* S, const double* G) {
            double* s0 = S + r0 * LDK + 2 * c0;
            const double* g0 = G + (long long)r0 * D + 2 * c0;
#pragma unroll
            for (int j = 0; j < DT / 16; ++j) cp_async16(s0 + j * 16 * LDK, g0 + (long long)j * 16 * D);
        };
        auto copy_mc = [&](double* S, const double* G) {
            double* s0 = S + r0 * LDM + 2 * c0;
            const double* g0 = G + (long long)r0 * D + 2 * c0;
#pragma unroll
            for (int j = 0; j < DT / 16; ++j) cp_async16(s0 + j * 16, g0 + j * 16);
        };
        unsigned gt = 0;                         // tiles issued so far: slot = gt % STAGES, use = gt / STAGES
        for (int i = 0; i < p.nops; ++i) {
            const ChainOp op = p.ops[i];
            if (op.flags & CHF_DEP) {            // reads the result of op i-1: every consumer warp must have stored and fenced it
                while (*ops_done < Cfg::CONS_WARPS * i) __nanosleep(64);
                __threadfence();
            }
            const int nprod = (op.a2 != CH_NONE) ? 2 : 1;
            for (int pr = 0; pr < nprod; ++pr) {
                const double* A = mat(pr ? op.a2 : op.a1);
                const double* B = mat(pr ? op.b2 : op.b1);
                for (int kt = 0; kt < ktiles_one; ++kt, ++gt) {
                    const int slot = gt % STAGES;
                    if (gt >= STAGES) mbar_wait(&empty[slot], ((gt / STAGES) - 1) & 1);
                    const int k0 = kt * BK;
                    double* As = smem + slot * Cfg::STAGE;
                    double* Bs = As + Cfg::OPER;
                    // every copy loop is "base + j * constant": thread (r0, c0) = (pt / 8, pt % 8) owns chunk column c0 of rows
                    // r0 + 16 j (k-contiguous operand) or chunks c0 + 8 j of row r0 (m/n-contiguous operand)
                    if (!(op.flags & CHF_TA)) copy_kc(As, A + k0);            // A is [M][K] row-major: rows m, k contiguous
                    else copy_mc(As, A + (long long)k0 * D);                  // A stored [K][M]: rows k, m contiguous
                    if (!(op.flags & CHF_TB)) copy_mc(Bs, B + (long long)k0 * D);   // B is [K][N] row-major
                    else copy_kc(Bs, B + k0);                                 // B stored [N][K]
                    mbar_arrive_cp_async(&full[slot]);
                }
            }
        }
        cp_async_wait_all();
        return;
    }

    // ================================================== consumers ==================================================
    reg_alloc<Cfg::CONS_REGS>();
    // Output tiles beyond the populated extent (136 of 144, 120 of 128) are zero padding: their MMAs are skipped (10.8 % / 12 % of
    // the 8 x 8 tiles) and the padding in memory is never written (it is zero from the handle's creation).  The warps on the last
    // row / column of the warp grid therefore have less work; `tile_of_warp` deals the warp tiles to the warp ids so that the four
    // SM sub-partitions (warp id mod 4) carry nearly equal DMMA counts (75 / 75 / 72 / 67 of 289 instead of 81 each of 324).
    const int wtile = Cfg::tile_of_warp(warp);
    const int wm0 = (wtile / Cfg::WARPS_N) * WM, wn0 = (wtile % Cfg::WARPS_N) * WN;
    // (the populated extent ends inside the last 8-wide tile row / column: 136 = 144 - 8, 120 = 128 - 8; anything else keeps all tiles)
    const bool trim = p.kvalid == DT - 8;
    const int mi_cnt = (trim && wm0 + WM == DT) ? MI - 1 : MI, ni_cnt = (trim && wn0 + WN == DT) ? NI - 1 : NI;
    const int g = lane >> 2, t = lane & 3;
    double acc[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    unsigned gt = 0;
    for (int i = 0; i < p.nops; ++i) {
        const ChainOp op = p.ops[i];
        const int ktiles = ktiles_one * (op.a2 != CH_NONE ? 2 : 1);
        // MC x NC = the 8 x 8 MMA tiles of this warp that are not padding: compile-time trip counts (a run-time predicate per DMMA
        // costs an ISETP per DMMA and measured slower than not trimming at all)
        auto run = [&](auto TA_, auto TB_, auto MC_, auto NC_) {
            constexpr bool TA = decltype(TA_)::value, TB = decltype(TB_)::value;
            constexpr int MC = decltype(MC_)::value, NC = decltype(NC_)::value;
            int kt_in = 0;                                        // k-tile index inside the current product
            for (int kt = 0; kt < ktiles; ++kt, ++gt) {
                const int slot = gt % STAGES;
                mbar_wait(&full[slot], (gt / STAGES) & 1);
                const double* As = smem + slot * Cfg::STAGE;
                const double* Bs = As + Cfg::OPER;
                const int kend = p.kvalid - kt_in * BK;           // < BK only in the last k-tile of a padded block
                auto kstep = [&](int kk) {
                    double a[MC], bf[NC];
#pragma unroll
                    for (int ii = 0; ii < MC; ++ii) {
                        const int m = wm0 + ii * 8 + g;
                        a[ii] = TA ? As[(kk + t) * LDM + m] : As[m * LDK + kk + t];
                    }
#pragma unroll
                    for (int jj = 0; jj < NC; ++jj) {
                        const int n = wn0 + jj * 8 + g;
                        bf[jj] = TB ? Bs[n * LDK + kk + t] : Bs[(kk + t) * LDM + n];
                    }
#pragma unroll
                    for (int ii = 0; ii < MC; ++ii)
#pragma unroll
                        for (int jj = 0; jj < NC; ++jj) dmma884(acc[ii][jj][0], acc[ii][jj][1], a[ii], bf[jj]);
                };
                if (kend >= BK) {
#pragma unroll
                    for (int kk = 0; kk < BK; kk += 4) kstep(kk);
                } else {
#pragma unroll 1
                    for (int kk = 0; kk < kend; kk += 4) kstep(kk);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[slot]);         // this warp is done reading the slot
                if (++kt_in == ktiles_one) kt_in = 0;
            }
        };
        const bool ta = op.flags & CHF_TA, tb = op.flags & CHF_TB;
        auto run_t = [&](auto MC_, auto NC_) {
            if (!ta && !tb) run(std::false_type{}, std::false_type{}, MC_, NC_);
            else if (!ta) run(std::false_type{}, std::true_type{}, MC_, NC_);
            else run(std::true_type{}, std::false_type{}, MC_, NC_);
        };
        using IM = std::integral_constant<int, MI>; using IM1 = std::integral_constant<int, MI - 1>;
        using IN = std::integral_constant<int, NI>; using IN1 = std::integral_constant<int, NI - 1>;
        if (mi_cnt == MI && ni_cnt == NI) run_t(IM{}, IN{});
        else if (mi_cnt == MI) run_t(IM{}, IN1{});
        else if (ni_cnt == NI) run_t(IM1{}, IN{});
        else run_t(IM1{}, IN1{});

        // ---- epilogue, per warp: C = acc (- T); partial sums of this warp's tile; accumulators reset ------------------------
        double* C = mat(op.c);
        const double* E = nullptr;
        if (op.flags & CHF_RESID) E = p.T + (long long)(p.tindex ? p.tindex[b] : 0) * p.sT;
        if (op.flags & CHF_DOT) E = mat(op.e);
        double part = 0.0;
#pragma unroll
        for (int ii = 0; ii < MI; ++ii) {
            const long long row = wm0 + ii * 8 + g;
#pragma unroll
            for (int jj = 0; jj < NI; ++jj) {
                if (ii >= mi_cnt || jj >= ni_cnt) continue;        // zero padding: neither computed nor stored
                const int col = wn0 + jj * 8 + 2 * t;
                double2 v = make_double2(acc[ii][jj][0], acc[ii][jj][1]);
                acc[ii][jj][0] = acc[ii][jj][1] = 0.0;
                if (op.flags & CHF_RESID) {
                    const double2 e = *reinterpret_cast<const double2*>(E + row * D + col);
                    v.x -= e.x; v.y -= e.y;
                    part = fma(v.x, v.x, part); part = fma(v.y, v.y, part);
                } else if (op.flags & CHF_DOT) {
                    // R at exactly the entries this thread stored two ops... one op ago: own writes, no cross-warp dependency
                    const double2 e = *reinterpret_cast<const double2*>(E + row * D + col);
                    part = fma(v.x, e.x, part); part = fma(v.y, e.y, part);
                }
                *reinterpret_cast<double2*>(C + row * D + col) = v;
            }
        }
        if (op.flags & (CHF_RESID | CHF_DOT)) {       // one partial per consumer warp, summed later in a fixed order
            part = warp_sum(part);
            if (lane == 0) ((op.flags & CHF_RESID) ? p.part_f : p.part_s)[(long long)b * p.part_stride + warp] = part;
        }
        __threadfence();                              // results visible in L2 before the producers' later cp.async read them
        __syncwarp();
        if (lane == 0) atomicAdd(const_cast<int*>(ops_done), 1);
    }
}

// ---- host side: the op program of a unit of work -------------------------------------------------------------------------
enum { CHAIN_FWD_MODEL = 0, CHAIN_FWD = 1, CHAIN_REV = 2, CHAIN_FWD_REV = 3, CHAIN_TANGENT = 4, CHAIN_TRIPLE = 5 };

struct ChainProgram {
    ChainOp ops[CHAIN_MAXOPS];
    int nops = 0;
    int products = 0;          // single D^3 products executed (a dual op counts 2)
};

inline unsigned char ch(int arr, int layer) { return (unsigned char)((arr << 4) | layer); }

// Builds the op list for `mode` at chain length m >= 2.  Layout conventions are those of otn_api.cu: P_0 = Y_0, G_{m-1} = R,
// W_0 = G_0, Gd[0/1] ping-pong with Mdot in Gd[0]; Wd has one slot per layer here (the tiled path reuses a single one because it
// pulls each Wd back before the next is formed).  *gd_final receives the ping-pong slot holding Ghat_0 (= Wd_0).
inline bool chain_build(int mode, int m, ChainProgram& pr, int* gd_final) {
    pr.nops = 0; pr.products = 0;
    if (m < 2 || m > 15) return false;
    auto P = [&](int l) { return l == 0 ? ch(CH_Y, 0) : ch(CH_P, l); };
    auto Pd = [&](int l) { return l == 0 ? ch(CH_YD, 0) : ch(CH_PD, l); };
    auto push = [&](unsigned char a1, unsigned char b1, unsigned char a2, unsigned char b2, unsigned char c, unsigned char e, int flags) {
        ChainOp& o = pr.ops[pr.nops++];
        o.a1 = a1; o.b1 = b1; o.a2 = a2; o.b2 = b2; o.c = c; o.e = e; o.flags = (unsigned char)flags; o.pad = 0;
        pr.products += (a2 != CH_NONE) ? 2 : 1;
    };
    const bool fwd = mode == CHAIN_FWD_MODEL || mode == CHAIN_FWD || mode == CHAIN_FWD_REV || mode == CHAIN_TRIPLE;
    const bool rev = mode == CHAIN_REV || mode == CHAIN_FWD_REV || mode == CHAIN_TRIPLE;
    const bool tan = mode == CHAIN_TANGENT || mode == CHAIN_TRIPLE;
    // forward: P_l = Y_l P_{l-1}  |  Pd_l = Yd_l P_{l-1} + Y_l Pd_{l-1}, interleaved so that neighbours are independent
    for (int l = 1; l < m; ++l) {
        const bool last = l == m - 1;
        if (fwd) {
            const bool resid = last && mode != CHAIN_FWD_MODEL;
            push(ch(CH_Y, l), P(l - 1), CH_NONE, CH_NONE, resid ? ch(CH_G, m - 1) : ch(CH_P, l), CH_NONE, resid ? CHF_RESID : 0);
        }
        if (tan) push(ch(CH_YD, l), P(l - 1), ch(CH_Y, l), Pd(l - 1), last ? ch(CH_GD, 0) : ch(CH_PD, l), last ? ch(CH_G, m - 1) : CH_NONE,
                      last ? CHF_DOT : 0);
    }
    // reverse: G_{l-1} = Y_l^T G_l | W_l = G_l P_{l-1}^T | Ghat_{l-1} = Yd_l^T G_l + Y_l^T Ghat_l | Wd_l = Ghat_l P_{l-1}^T + G_l Pd_{l-1}^T
    int cur = 0;
    for (int l = m - 1; l >= 1; --l) {
        if (rev) {
            push(ch(CH_Y, l), ch(CH_G, l), CH_NONE, CH_NONE, ch(CH_G, l - 1), CH_NONE, CHF_TA);
            push(ch(CH_G, l), P(l - 1), CH_NONE, CH_NONE, ch(CH_W, l), CH_NONE, CHF_TB);
        }
        if (tan) {
            push(ch(CH_YD, l), ch(CH_G, l), ch(CH_Y, l), ch(CH_GD, cur), ch(CH_GD, 1 - cur), CH_NONE, CHF_TA);
            push(ch(CH_GD, cur), P(l - 1), ch(CH_G, l), Pd(l - 1), ch(CH_WD, l), CH_NONE, CHF_TB);
            cur = 1 - cur;
        }
    }
    if (gd_final) *gd_final = cur;
    // dependency flags: an op that reads the result of its predecessor cannot be prefetched under it
    for (int i = 1; i < pr.nops; ++i) {
        const ChainOp& o = pr.ops[i];
        const unsigned char w = pr.ops[i - 1].c;
        if (o.a1 == w || o.b1 == w || o.a2 == w || o.b2 == w) pr.ops[i].flags |= CHF_DEP;
        // (the DOT operand `e` is only read in the epilogue, after the predecessor has completed)
    }
    return pr.nops <= CHAIN_MAXOPS;
}

template <int DT, int WM, int WN, int STAGES>
inline cudaError_t chain_launch_cfg(const ChainParams& p, int batch, cudaStream_t st) {
    using Cfg = ChainCfg<DT, WM, WN, STAGES>;
    auto kern = chain_fused_kernel<DT, WM, WN, STAGES>;
    static std::atomic<unsigned long long> configured{0};
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ULL << (dev & 63);
    if (!(configured.load(std::memory_order_acquire) & bit)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured.fetch_or(bit, std::memory_order_release);
    }
    kern<<<(unsigned)batch, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(p);
    return cudaGetLastError();
}

inline bool chain_fused_supports(int D) { return D == 144 || D == 128; }
// consumer warps = partial sums per instance and block
inline bool chain_128_w8() { static const bool v = getenv("OTN_CHAIN_W8") != nullptr; return v; }   // env: A/B switch (8 warps of 64 x 32)
inline int chain_fused_partials(int D) { return D == 144 ? 12 : (chain_128_w8() ? 8 : 16); }

inline cudaError_t chain_launch(const ChainParams& p, int batch, cudaStream_t st) {
    if (p.D == 144) return chain_launch_cfg<144, 72, 24, 4>(p, batch, st);   // 12 consumer warps (3 per SM sub-partition) + 4 producer warps
    if (p.D == 128 && chain_128_w8()) return chain_launch_cfg<128, 64, 32, 4>(p, batch, st);   //  8 consumer warps (2 per SM sub-partition)
    if (p.D == 128) return chain_launch_cfg<128, 32, 32, 4>(p, batch, st);   // 16 consumer warps (4 per SM sub-partition) + 4 producer warps
    return cudaErrorInvalidValue;
}

}  // namespace otn
