// Batched square FP64 GEMM for the superoperator chains (sites C / E / G of the hot path):
//     C_b = op(A1_b) op(B1_b) [+ op(A2_b) op(B2_b)]        all D x D, row-major, b = instance
// on the FP64 tensor pipe (mma.sync m8n8k4 -> SASS DMMA.8x8x4; tcgen05 has no FP64 kind), operands staged through
// shared memory by a multi-stage cp.async (LDGSTS) ring, fragments read conflict-free thanks to +4 padding.
//
// Replaces: reference `compose_superops_list` (opentn/transformations.py:841-851, `jnp.dot(op, composition)`) and the
// products JAX's reverse / forward-over-reverse sweeps generate from it (opentn/stiefel.py:423,520).
//
// Fused epilogues (save one full pass over a D x D matrix each):
//   EPI_RESID : C = acc - T      and  partial[b][tile] = sum (acc - T)^2       (frobenius_norm, optimization.py:56-61)
//   EPI_DOT   : C = acc          and  partial[b][tile] = sum acc * R           (d/dt of the norm)
// Partials are reduced later in a fixed order => bitwise reproducible costs.
#pragma once
#include <atomic>
#include <cstdlib>

#include <cooperative_groups.h>

#include "otn_common.cuh"

namespace otn {

enum { EPI_NONE = 0, EPI_RESID = 1, EPI_DOT = 2 };

struct GemmParams {
    const double* A1; const double* B1;      // first product
    const double* A2; const double* B2;      // optional second product accumulated into the same tile (nullptr = none)
    double* C;
    long long sA1, sB1, sA2, sB2, sC;        // per-instance strides in doubles (0 = shared operand)
    int D;                                   // matrix side (leading dimension)
    int kvalid;                              // contraction extent actually populated (0 = D); rows/columns beyond it are zero
                                             // padding of the block-diagonal evaluation and are skipped (multiple of 4)
    int row0, rows;                          // row window of C computed by this launch (row sharding); rows % BM == 0
    int batch;
    const int* active;                       // optional per-instance mask (nullptr = all active)
    // epilogue
    int epi;
    const double* E;                         // EPI_RESID: targets (real part) ; EPI_DOT: residuals
    long long sE;                            // stride of E per instance (EPI_DOT) or per target (EPI_RESID)
    const int* e_index;                      // EPI_RESID: target id per instance (nullptr = 0)
    double* partial;                         // partial sums: partial[b * partial_stride + tile]
    int partial_stride;                      // >= tiles of this launch (several launches may share one row per instance)
    // fused all-gather of a row-sharded product: every output tile is also stored into the same position of the peers'
    // result matrices (NVLink peer-mapped pointers; instance b at the same offset b * sC as in C), and the tile's partial sum
    // into the peers' partial arrays (Ppeer[q] != nullptr: row-sharded handles, otn_rowshard_attach)
    double* Cpeer[7];
    double* Ppeer[7];
    int npeer;
    // Real embedding of complex matrices, emb(A) = [[Ar, -Ai], [Ai, Ar]] (side D = 2 * mirror): a product of embeddings is an
    // embedding, so only its upper half is computed (rows [0, mirror): 4 instead of 8 real products per complex product) and
    // every element is mirrored into the lower half -- (r, c) -> (r + mirror, c + mirror) for c < mirror, and
    // -(r, c) -> (r + mirror, c - mirror) otherwise.  0 = off.  Partial sums then cover the upper half only.
    int mirror;
    // Split-K over a pair of CTAs (thread-block cluster of 2; 64 x 64 tiles only): each CTA of the pair runs half of the k-tiles of
    // one output tile (for a dual product: one product each), the second hands its accumulators to the first through distributed
    // shared memory, which adds them (own + partner: a fixed order) and runs the epilogue.  Twice as many, half as long CTAs: for a
    // launch with only 3-4 tiles per SM (a row-sharded D = 4096 product on 8 GPUs: 512 tiles on 148 SMs) the SMs end level, and
    // the epilogues -- peer stores over NVLink -- of the first wave run under the math of the second.  0 = off.
    int splitk;
};

template <int BM, int BN, int BK, int WM, int WN, int STAGES, bool TA, bool TB>
struct GemmCfg {
    static constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN;
    static constexpr int THREADS = WARPS_M * WARPS_N * 32;
    // smem tile layouts: k-contiguous source -> [rows][BK+4]; m/n-contiguous source -> [BK][rows+4]
    static constexpr int LDA = TA ? (BM + 4) : (BK + 4);
    static constexpr int LDB = TB ? (BK + 4) : (BN + 4);
    static constexpr int A_ELEMS = TA ? BK * LDA : BM * LDA;
    static constexpr int B_ELEMS = TB ? BN * LDB : BK * LDB;
    static constexpr int STAGE_ELEMS = A_ELEMS + B_ELEMS;
    static constexpr size_t SMEM_BYTES = size_t(STAGES) * STAGE_ELEMS * sizeof(double);
};

template <int BM, int BN, int BK, int WM, int WN, int STAGES, bool TA, bool TB, int SPLIT = 1>
__global__ void __launch_bounds__(GemmCfg<BM, BN, BK, WM, WN, STAGES, TA, TB>::THREADS)
gemm_dmma_kernel(const GemmParams p) {
    using Cfg = GemmCfg<BM, BN, BK, WM, WN, STAGES, TA, TB>;
    constexpr int NT = Cfg::THREADS;
    constexpr int MI = WM / 8, NI = WN / 8;
    extern __shared__ __align__(16) double smem[];

    const int tiles_n = p.D / BN;
    const int tiles = (p.rows / BM) * tiles_n;
    const int work = SPLIT == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;      // SPLIT == 2: CTAs 2w and 2w+1 are one cluster
    const int half = SPLIT == 2 ? (int)(blockIdx.x & 1) : 0;
    const int b = work / tiles;
    const int tile = work - b * tiles;
    if (p.active != nullptr && p.active[b] == 0) return;                         // (both CTAs of a pair leave together)
    const int m0 = p.row0 + (tile / tiles_n) * BM;
    const int n0 = (tile % tiles_n) * BN;
    const int D = p.D;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm0 = (warp / Cfg::WARPS_N) * WM, wn0 = (warp % Cfg::WARPS_N) * WN;
    const int g = lane >> 2, t = lane & 3;

    const int nprod = (p.A2 != nullptr) ? 2 : 1;
    const int kvalid = p.kvalid > 0 ? p.kvalid : D;
    const int ktiles_one = (kvalid + BK - 1) / BK;
    const int ktiles_all = ktiles_one * nprod;
    // this CTA's share of the k-tiles: all of them, or one half (a dual product splits into its two products)
    const int kt_begin = (SPLIT == 2 && half == 1) ? (ktiles_all + 1) / 2 : 0;
    const int ktiles = (SPLIT == 2) ? (half == 0 ? (ktiles_all + 1) / 2 : ktiles_all) : ktiles_all;

    const double* A1 = p.A1 + (long long)b * p.sA1;
    const double* B1 = p.B1 + (long long)b * p.sB1;
    const double* A2 = nprod == 2 ? p.A2 + (long long)b * p.sA2 : nullptr;
    const double* B2 = nprod == 2 ? p.B2 + (long long)b * p.sB2 : nullptr;

    auto load_tile = [&](int kt, int stage) {
        const double* A = (kt < ktiles_one) ? A1 : A2;
        const double* B = (kt < ktiles_one) ? B1 : B2;
        const int k0 = (kt < ktiles_one ? kt : kt - ktiles_one) * BK;
        double* As = smem + stage * Cfg::STAGE_ELEMS;
        double* Bs = As + Cfg::A_ELEMS;
        if (!TA) {  // A is [M][K] row-major: rows m, k contiguous
            constexpr int CH = BK / 2;
            for (int id = tid; id < BM * CH; id += NT) {
                int r = id / CH, c = id - r * CH;
                cp_async16(As + r * Cfg::LDA + 2 * c, A + (long long)(m0 + r) * D + k0 + 2 * c);
            }
        } else {    // A is stored [K][M]: rows k, m contiguous
            constexpr int CH = BM / 2;
            for (int id = tid; id < BK * CH; id += NT) {
                int r = id / CH, c = id - r * CH;
                cp_async16(As + r * Cfg::LDA + 2 * c, A + (long long)(k0 + r) * D + m0 + 2 * c);
            }
        }
        if (!TB) {  // B is [K][N] row-major: rows k, n contiguous
            constexpr int CH = BN / 2;
            for (int id = tid; id < BK * CH; id += NT) {
                int r = id / CH, c = id - r * CH;
                cp_async16(Bs + r * Cfg::LDB + 2 * c, B + (long long)(k0 + r) * D + n0 + 2 * c);
            }
        } else {    // B is stored [N][K]: rows n, k contiguous
            constexpr int CH = BK / 2;
            for (int id = tid; id < BN * CH; id += NT) {
                int r = id / CH, c = id - r * CH;
                cp_async16(Bs + r * Cfg::LDB + 2 * c, B + (long long)(n0 + r) * D + k0 + 2 * c);
            }
        }
    };

    double acc[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (kt_begin + s < ktiles) load_tile(kt_begin + s, s);
        cp_async_commit();
    }

    for (int kt = kt_begin; kt < ktiles; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nk = kt + STAGES - 1;
            if (nk < ktiles) load_tile(nk, (nk - kt_begin) % STAGES);
            cp_async_commit();
        }
        const double* As = smem + ((kt - kt_begin) % STAGES) * Cfg::STAGE_ELEMS;
        const double* Bs = As + Cfg::A_ELEMS;
        const int kend = kvalid - (kt < ktiles_one ? kt : kt - ktiles_one) * BK;   // < BK only in the last k-tile of a padded block
        auto kstep = [&](int kk) {
            double a[MI], bf[NI];
#pragma unroll
            for (int i = 0; i < MI; ++i) {
                int m = wm0 + i * 8 + g;
                a[i] = TA ? As[(kk + t) * Cfg::LDA + m] : As[m * Cfg::LDA + kk + t];
            }
#pragma unroll
            for (int j = 0; j < NI; ++j) {
                int n = wn0 + j * 8 + g;
                bf[j] = TB ? Bs[n * Cfg::LDB + kk + t] : Bs[(kk + t) * Cfg::LDB + n];
            }
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], bf[j]);
        };
        if (kend >= BK) {
#pragma unroll
            for (int kk = 0; kk < BK; kk += 4) kstep(kk);
        } else {
#pragma unroll 1
            for (int kk = 0; kk < kend; kk += 4) kstep(kk);
        }
    }
    cp_async_wait<0>();

    if (SPLIT == 2) {
        // the second CTA of the pair parks its accumulators in its (now idle) operand ring, thread-linear so that the first CTA's
        // reads over distributed shared memory are coalesced; sum = own + partner
        static_assert(SPLIT == 1 || (size_t)MI * NI * 2 * NT * sizeof(double) <= Cfg::SMEM_BYTES, "accumulators must fit the operand ring");
        namespace cg = cooperative_groups;
        cg::cluster_group cluster = cg::this_cluster();
        __syncthreads();                                         // every warp is done reading the ring
        if (half == 1) {
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j) {
                    smem[((i * NI + j) * 2 + 0) * NT + tid] = acc[i][j][0];
                    smem[((i * NI + j) * 2 + 1) * NT + tid] = acc[i][j][1];
                }
        }
        cluster.sync();
        if (half == 0) {
            const double* other = cluster.map_shared_rank(smem, 1);
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j) {
                    acc[i][j][0] += other[((i * NI + j) * 2 + 0) * NT + tid];
                    acc[i][j][1] += other[((i * NI + j) * 2 + 1) * NT + tid];
                }
        }
        cluster.sync();                                          // the partner's shared memory stays alive until it has been read
        if (half == 1) return;
    }

    // ---- epilogue ----
    double* C = p.C + (long long)b * p.sC;
    const double* E = nullptr;
    if (p.epi == EPI_RESID) E = p.E + (long long)(p.e_index ? p.e_index[b] : 0) * p.sE;
    if (p.epi == EPI_DOT) E = p.E + (long long)b * p.sE;
    double part = 0.0;
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const long long row = m0 + wm0 + i * 8 + g;
#pragma unroll
        for (int j = 0; j < NI; ++j) {
            const int col = n0 + wn0 + j * 8 + 2 * t;
            double2 v = make_double2(acc[i][j][0], acc[i][j][1]);
            if (p.epi == EPI_RESID) {
                double2 e = *reinterpret_cast<const double2*>(E + row * D + col);
                v.x -= e.x; v.y -= e.y;
                part = fma(v.x, v.x, part); part = fma(v.y, v.y, part);
            } else if (p.epi == EPI_DOT) {
                double2 e = *reinterpret_cast<const double2*>(E + row * D + col);
                part = fma(v.x, e.x, part); part = fma(v.y, e.y, part);
            }
            *reinterpret_cast<double2*>(C + row * D + col) = v;
            if (p.mirror > 0) {
                const bool left = col < p.mirror;
                const double2 w = left ? v : make_double2(-v.x, -v.y);
                *reinterpret_cast<double2*>(C + (row + p.mirror) * D + (left ? col + p.mirror : col - p.mirror)) = w;
            }
            for (int q = 0; q < p.npeer; ++q) *reinterpret_cast<double2*>(p.Cpeer[q] + (long long)b * p.sC + row * D + col) = v;
        }
    }
    if (p.epi != EPI_NONE) {
        __shared__ double red[Cfg::THREADS / 32];
        part = warp_sum(part);
        __syncthreads();
        if (lane == 0) red[warp] = part;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int w = 0; w < Cfg::THREADS / 32; ++w) s += red[w];
            p.partial[(long long)b * p.partial_stride + tile] = s;
            for (int q = 0; q < p.npeer; ++q)
                if (p.Ppeer[q] != nullptr) p.Ppeer[q][(long long)b * p.partial_stride + tile] = s;
        }
    }
}

// ---- generic fallback on the FP64 CUDA cores (any D, incl. D < 64 or D % 64 != 0) ------------------------------------
// 32x32 output tile per 256-thread CTA, 16-deep k tiles, each thread a 2x2 register tile.
template <bool TA, bool TB>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmParams p) {
    constexpr int TS = 32, KS = 16;
    __shared__ double As[KS][TS + 1];
    __shared__ double Bs[KS][TS + 1];
    const int D = p.D;
    const int tiles_n = (D + TS - 1) / TS;
    const int tiles_m = (p.rows + TS - 1) / TS;
    const int tiles = tiles_m * tiles_n;
    const int b = blockIdx.x / tiles;
    const int tile = blockIdx.x - b * tiles;
    if (p.active != nullptr && p.active[b] == 0) return;
    const int m0 = p.row0 + (tile / tiles_n) * TS, n0 = (tile % tiles_n) * TS;
    const int mend = p.row0 + p.rows;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int nprod = p.A2 ? 2 : 1;
    double acc[2][2] = {{0, 0}, {0, 0}};
    for (int pr = 0; pr < nprod; ++pr) {
        const double* A = (pr == 0 ? p.A1 + (long long)b * p.sA1 : p.A2 + (long long)b * p.sA2);
        const double* B = (pr == 0 ? p.B1 + (long long)b * p.sB1 : p.B2 + (long long)b * p.sB2);
        for (int k0 = 0; k0 < D; k0 += KS) {
            for (int id = tid; id < KS * TS; id += 256) {
                int kk, mm;
                if (!TA) { mm = id / KS; kk = id % KS; } else { kk = id / TS; mm = id % TS; }
                int m = m0 + mm, k = k0 + kk;
                double v = 0.0;
                if (m < mend && k < D) v = TA ? A[(long long)k * D + m] : A[(long long)m * D + k];
                As[kk][mm] = v;
                int nn;
                if (!TB) { kk = id / TS; nn = id % TS; } else { nn = id / KS; kk = id % KS; }
                int n = n0 + nn; k = k0 + kk;
                v = 0.0;
                if (n < D && k < D) v = TB ? B[(long long)n * D + k] : B[(long long)k * D + n];
                Bs[kk][nn] = v;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < KS; ++kk) {
                double a0 = As[kk][ty], a1 = As[kk][ty + 16], b0 = Bs[kk][tx], b1 = Bs[kk][tx + 16];
                acc[0][0] = fma(a0, b0, acc[0][0]); acc[0][1] = fma(a0, b1, acc[0][1]);
                acc[1][0] = fma(a1, b0, acc[1][0]); acc[1][1] = fma(a1, b1, acc[1][1]);
            }
            __syncthreads();
        }
    }
    double* C = p.C + (long long)b * p.sC;
    const double* E = nullptr;
    if (p.epi == EPI_RESID) E = p.E + (long long)(p.e_index ? p.e_index[b] : 0) * p.sE;
    if (p.epi == EPI_DOT) E = p.E + (long long)b * p.sE;
    double part = 0.0;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            int m = m0 + ty + 16 * i, n = n0 + tx + 16 * j;
            if (m < mend && n < D) {
                double v = acc[i][j];
                if (p.epi == EPI_RESID) { v -= E[(long long)m * D + n]; part = fma(v, v, part); }
                else if (p.epi == EPI_DOT) { part = fma(v, E[(long long)m * D + n], part); }
                C[(long long)m * D + n] = v;
                if (p.mirror > 0) C[(long long)(m + p.mirror) * D + (n < p.mirror ? n + p.mirror : n - p.mirror)] = (n < p.mirror) ? v : -v;
            }
        }
    if (p.epi != EPI_NONE) {
        __shared__ double red[8];
        part = warp_sum(part);
        if ((tid & 31) == 0) red[tid >> 5] = part;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int w = 0; w < 8; ++w) s += red[w];
            p.partial[(long long)b * p.partial_stride + tile] = s;
        }
    }
}

// (A persistent variant -- one CTA per resident slot walking a static tile list with the cp.async ring running across tile
// boundaries -- was measured in round 1: 22.8 vs 28.6 TFLOP/s at D = 144.  The hardware CTA scheduler with 4-5 co-resident CTAs
// per SM already hides the per-tile prologue; the extra index arithmetic and register pressure cost more than they saved.)

// ---- host-side dispatch -------------------------------------------------------------------------------------------
struct GemmPlan {
    int kind;      // 0 = simt fallback, 1 = dmma 128x128, 2 = dmma 64x64, 3 = dmma 48x48
    int tiles;     // CTAs per instance (for the `partial` buffer)
};

// Split-K CTA pairs (GemmParams::splitk) for the large single-instance products (D >= 1024: the N = 6 chain, whole or row-sharded):
// measured 2 % faster on a full D = 4096 product and 25 % on its 8-GPU row shard (512 tiles on 148 SMs).  The rule depends on D
// only, so a row-sharded product and the one-GPU product sum in the same order and stay bitwise equal.
// env OTN_SPLITK: "0" = never, "1" = always (tests, A/B measurements).
inline int gemm_want_split(int D) {
    const char* env = getenv("OTN_SPLITK");
    if (env != nullptr) return env[0] == '1' ? 1 : 0;
    return D >= 1024 ? 1 : 0;
}

inline GemmPlan gemm_plan(int D, int rows, int batch) {
    GemmPlan pl;
    static const int big_tile_min_d = [] { const char* e = getenv("OTN_GEMM_BIG_TILE_MIN_D"); return e ? atoi(e) : (1 << 30); }();   // 64x64 tiles measured faster even at D = 4096 (32.7 vs 30.0 TFLOP/s)
    // 64x64 CTA tiles (4 warps of 32x32, 3 CTAs/SM) measured faster than 128x128 at D = 256 (32.4 vs 28.2 TFLOP/s); the
    // larger tile only pays for big single-instance products where L2 traffic matters (D >= 2048).
    if (D % 128 == 0 && rows % 128 == 0 && D >= big_tile_min_d) {
        pl.kind = 1; pl.tiles = (rows / 128) * (D / 128);
    } else if (D % 64 == 0 && rows % 64 == 0) {
        pl.kind = 2; pl.tiles = (rows / 64) * (D / 64);
    } else if (D % 48 == 0 && rows % 48 == 0) {   // the 144 x 144 symmetric block of the structured evaluation
        pl.kind = 3; pl.tiles = (rows / 48) * (D / 48);
    } else {
        pl.kind = 0; pl.tiles = ((rows + 31) / 32) * ((D + 31) / 32);
    }
    return pl;
}

template <int BM, int BN, int BK, int WM, int WN, int STAGES, bool TA, bool TB>
inline cudaError_t launch_dmma(const GemmParams& p, int tiles, cudaStream_t st) {
    using Cfg = GemmCfg<BM, BN, BK, WM, WN, STAGES, TA, TB>;
    auto kern = gemm_dmma_kernel<BM, BN, BK, WM, WN, STAGES, TA, TB>;
    // opt in to > 48 KB of dynamic shared memory once per device (function attributes are per device context)
    static std::atomic<unsigned long long> configured{0};
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ULL << (dev & 63);
    if (!(configured.load(std::memory_order_acquire) & bit)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured.fetch_or(bit, std::memory_order_release);
    }
    kern<<<(unsigned)((long long)p.batch * tiles), Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(p);
    return cudaGetLastError();
}

// the split-K variant: clusters of two CTAs per output tile
template <int BM, int BN, int BK, int WM, int WN, int STAGES, bool TA, bool TB>
inline cudaError_t launch_dmma_split2(const GemmParams& p, int tiles, cudaStream_t st) {
    using Cfg = GemmCfg<BM, BN, BK, WM, WN, STAGES, TA, TB>;
    auto kern = gemm_dmma_kernel<BM, BN, BK, WM, WN, STAGES, TA, TB, 2>;
    static std::atomic<unsigned long long> configured{0};
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ULL << (dev & 63);
    if (!(configured.load(std::memory_order_acquire) & bit)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured.fetch_or(bit, std::memory_order_release);
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(2LL * p.batch * tiles));
    cfg.blockDim = dim3(Cfg::THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, p);
}

template <bool TA, bool TB>
inline cudaError_t gemm_launch_t(const GemmParams& p, cudaStream_t st) {
    GemmPlan pl = gemm_plan(p.D, p.rows, p.batch);
    if (pl.kind == 1) return launch_dmma<128, 128, 16, 64, 32, 3, TA, TB>(p, pl.tiles, st);
    if (pl.kind == 2 && p.splitk) return launch_dmma_split2<64, 64, 16, 32, 32, 3, TA, TB>(p, pl.tiles, st);
    if (pl.kind == 2) return launch_dmma<64, 64, 16, 32, 32, 3, TA, TB>(p, pl.tiles, st);
    if (pl.kind == 3) return launch_dmma<48, 48, 16, 24, 24, 3, TA, TB>(p, pl.tiles, st);
    gemm_simt_kernel<TA, TB><<<(unsigned)((long long)p.batch * pl.tiles), 256, 0, st>>>(p);
    return cudaGetLastError();
}

inline cudaError_t gemm_launch(const GemmParams& p, bool ta, bool tb, cudaStream_t st) {
    if (!ta && !tb) return gemm_launch_t<false, false>(p, st);
    if (!ta && tb) return gemm_launch_t<false, true>(p, st);
    if (ta && !tb) return gemm_launch_t<true, false>(p, st);
    return cudaErrorInvalidValue;  // A^T B^T never occurs on this path
}

}  // namespace otn
