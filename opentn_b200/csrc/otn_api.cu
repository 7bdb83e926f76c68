// libotn: C ABI (include/otn.h) over the sm_100a kernels.  No torch, no cuBLAS/cuSOLVER: every arithmetic site of the hot path
// runs in the hand-written kernels of gemm_dmma.cuh / assemble.cuh / stiefel.cuh / tr.cuh.
#include "../../include/otn.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>   // header-only; a no-op unless a profiler (nsys / ncu --nvtx) is attached

#include "assemble.cuh"
#include "chain_fused.cuh"
#include "complex.cuh"
#include "factored.cuh"
#include "gemm_dmma.cuh"
#include "stiefel.cuh"
#include "structured.cuh"
#include "tr.cuh"

using namespace otn;

// ---------------------------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return -1;
}
#define CUDA_TRY(expr)                                                                                          \
    do {                                                                                                        \
        cudaError_t e__ = (expr);                                                                               \
        if (e__ != cudaSuccess) return fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)
#define OTN_TRY(expr)         \
    do {                      \
        int r__ = (expr);     \
        if (r__ != 0) return r__; \
    } while (0)

extern "C" const char* otn_last_error(void) { return g_err.c_str(); }
extern "C" int otn_version(void) { return 100; }

// Every entry point runs on its own device but must leave the caller's current device untouched (a host framework such as
// torch tracks the current device per thread; changing it behind its back misplaces the caller's next allocation).
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; if (prev != dev) cudaSetDevice(dev); }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

static bool is_device_ptr(const void* p) {
    if (p == nullptr) return false;
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// grow-only device buffer
struct DevBuf {
    void* ptr = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (ptr) cudaFree(ptr);
        ptr = nullptr; cap = 0;
        CUDA_TRY(cudaMalloc(&ptr, bytes));
        cap = bytes;
        return 0;
    }
    void release() { if (ptr) cudaFree(ptr); ptr = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(ptr); }
};

// input: returns a device pointer holding `bytes` of `src` (in place when src is already on the device)
static int stage_in(const void* src, size_t bytes, DevBuf& buf, cudaStream_t st, const void** out) {
    if (src == nullptr) { *out = nullptr; return 0; }
    if (is_device_ptr(src)) { *out = src; return 0; }
    OTN_TRY(buf.ensure(bytes));
    CUDA_TRY(cudaMemcpyAsync(buf.ptr, src, bytes, cudaMemcpyHostToDevice, st));
    *out = buf.ptr;
    return 0;
}
// output: device pointer to compute into; `finish_out` copies back if dst is a host pointer
static int stage_out(void* dst, size_t bytes, DevBuf& buf, void** out) {
    if (dst == nullptr) { *out = nullptr; return 0; }
    if (is_device_ptr(dst)) { *out = dst; return 0; }
    OTN_TRY(buf.ensure(bytes));
    *out = buf.ptr;
    return 0;
}
static int finish_out(void* dst, const void* dev, size_t bytes, cudaStream_t st) {
    if (dst == nullptr || dst == dev) return 0;
    CUDA_TRY(cudaMemcpyAsync(dst, dev, bytes, cudaMemcpyDeviceToHost, st));
    return 0;
}

template <class K>
static int set_smem(K kern, size_t bytes) {
    if (bytes > 227 * 1024)
        return fail("kernel needs %zu bytes of shared memory (> 227 KB): isometry too large for the fused Stiefel kernels "
                    "(gradient / HVP / trust region need n*p <= 6144, i.e. Kraus rank <= 24 for the full-sites N = 4 model)", bytes);
    // opt in early: the 48 KB default limit covers static + dynamic shared memory together
    if (bytes > 40 * 1024) CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// launch accounting + optional per-kernel CUDA-event timing (bench.py's live roofline measurement)
// ---------------------------------------------------------------------------------------------------------------------
enum { CAT_GEMM = 0, CAT_ASSEMBLE = 1, CAT_BACK = 2, CAT_RIEM = 3, CAT_TR = 4, CAT_OTHER = 5, CAT_COUNT = 6 };
static std::atomic<long long> g_launches{0};
struct ProfRec { int cat; cudaEvent_t e0, e1; double flops; int n; };
// Side stream of the block-diagonal chain: the antisymmetric block of every chain product is launched on `st`, the symmetric one
// on the handle's stream, so CTAs of one block fill the drain phase of the other (between two non-GEMM kernels the two chains are
// independent).  fork = the side stream waits for everything enqueued on the main stream so far; join = the main stream waits for
// the side stream.  Every non-GEMM launch site opens a ProfScope, which joins first: a non-GEMM kernel is a barrier for both chains.
struct SideStream {
    cudaStream_t st = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool enabled = false;      // structured evaluation with two diagonal blocks
    bool suspended = false;    // pipelined chunks (already overlapped on two compute streams), graph-captured tCG steps
    bool busy = false;         // side work enqueued since the last join
    // live profiling of a fork..join group: overlapping launches have no additive individual durations, so the group is timed
    ProfRec grp{}; bool grp_live = false;
    bool active() const { return enabled && !suspended && st != nullptr; }
};
struct Profiler {
    bool on = false;
    std::vector<ProfRec> recs;
    std::vector<cudaEvent_t> pool;
    SideStream side;
    cudaEvent_t get() {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e; cudaEventCreate(&e); return e;
    }
    // called with the main stream before anything that is not a chain GEMM is enqueued on it
    void join(cudaStream_t main) {
        if (!side.busy) return;
        cudaEventRecord(side.ev_join, side.st);
        cudaStreamWaitEvent(main, side.ev_join, 0);
        side.busy = false;
        if (side.grp_live) { cudaEventRecord(side.grp.e1, main); recs.push_back(side.grp); side.grp_live = false; }
    }
    // called before a chain GEMM goes to the side stream
    void fork(cudaStream_t main) {
        if (side.busy) return;
        if (on) { side.grp = ProfRec{CAT_GEMM, get(), get(), 0.0, 0}; cudaEventRecord(side.grp.e0, main); side.grp_live = true; }
        cudaEventRecord(side.ev_fork, main);
        cudaStreamWaitEvent(side.st, side.ev_fork, 0);
        side.busy = true;
    }
};
// NVTX ranges: every launch site of the handle sits inside one of these, named by the stage of the path
static const char* const kNvtxNames[CAT_COUNT] = {"otn:chain_gemm", "otn:kraus_assembly", "otn:back_assembly", "otn:riemannian_epilogue",
                                                  "otn:trust_region", "otn:other"};
struct ProfScope {
    Profiler* pr; cudaStream_t st; ProfRec rec; bool live;
    ProfScope(Profiler* p, cudaStream_t s, int cat, double flops, int nkernels = 1, bool nojoin = false) : pr(p), st(s), live(p && p->on) {
        nvtxRangePushA(kNvtxNames[cat]);
        g_launches += nkernels;
        if (p && cat != CAT_GEMM && !nojoin) p->join(s);
        if (p && cat == CAT_GEMM && p->side.grp_live) {   // part of a timed fork..join group
            p->side.grp.flops += flops; p->side.grp.n += nkernels; live = false;
        }
        if (live) { rec.cat = cat; rec.flops = flops; rec.n = nkernels; rec.e0 = pr->get(); rec.e1 = pr->get(); cudaEventRecord(rec.e0, st); }
    }
    ~ProfScope() { if (live) { cudaEventRecord(rec.e1, st); pr->recs.push_back(rec); } nvtxRangePop(); }
};
extern "C" long long otn_launch_count(void) { return g_launches.load(); }

// ---------------------------------------------------------------------------------------------------------------------
// embedding index tables (transformations.py:431-444, 473-482)
// ---------------------------------------------------------------------------------------------------------------------
static int ipow(int b, int e) { int r = 1; while (e-- > 0) r *= b; return r; }

static void build_embed_tables(int N, int d, std::vector<unsigned>& tab, std::vector<int>& inv) {
    const int nax = 2 * N, nf = N / 2, q = ipow(d, 4), D = ipow(d, 2 * N);
    // moveaxis(source_one_side -> destination_one_side) on the 2N axes of one side
    std::vector<int> src, dst;
    for (int i = 0; i < (N - 2) / 2; ++i) { src.push_back(2 + 4 * i); src.push_back(3 + 4 * i); }
    for (int i = N; i < 2 * N - 2; ++i) dst.push_back(i);
    std::vector<int> order;
    for (int a = 0; a < nax; ++a) if (std::find(src.begin(), src.end(), a) == src.end()) order.push_back(a);
    std::vector<std::pair<int, int>> ds;
    for (size_t i = 0; i < src.size(); ++i) ds.push_back({dst[i], src[i]});
    std::sort(ds.begin(), ds.end());
    for (auto& pr : ds) order.insert(order.begin() + pr.first, pr.second);
    tab.assign(2 * (size_t)D, 0u);
    for (int shift = 0; shift < 2; ++shift) {
        std::vector<int> perm(nax);  // new axis j <- old (tensored) axis perm[j]
        for (int j = 0; j < nax; ++j) {
            int jj = j;
            if (shift) {  // transpose(idx), idx = [1..N-1, 0, N+1..2N-1, N]
                jj = (j < N) ? ((j + 1) % N) : (N + ((j - N + 1) % N));
            }
            perm[j] = order[jj];
        }
        for (int r = 0; r < D; ++r) {
            int rem = r, rt = 0;
            for (int j = nax - 1; j >= 0; --j) {
                const int digit = rem % d;
                rem /= d;
                rt += digit * ipow(d, nax - 1 - perm[j]);
            }
            unsigned packed = 0;
            for (int f = 0; f < nf; ++f) packed |= (unsigned)((rt / ipow(q, nf - 1 - f)) % q) << (8 * f);
            tab[(size_t)shift * D + r] = packed;
        }
    }
    const int per = D / q;
    inv.assign((size_t)2 * nf * q * per, 0);
    for (int shift = 0; shift < 2; ++shift)
        for (int f = 0; f < nf; ++f) {
            std::vector<int> fill(q, 0);
            for (int r = 0; r < D; ++r) {
                const int i = (tab[(size_t)shift * D + r] >> (8 * f)) & 0xff;
                inv[(((size_t)shift * nf + f) * q + i) * per + fill[i]++] = r;
            }
        }
}

// ---------------------------------------------------------------------------------------------------------------------
// the handle
// ---------------------------------------------------------------------------------------------------------------------
struct otn_problem {
    int N, d, m, K, model, metric, max_batch, n_targets, device;
    int dim;        // isometry column count p (= d^N full, d^2 local)
    int n, p, np;   // isometry shape
    int D, q, nf;   // superoperator side; local superoperator side; kron factors
    long long SZ;   // doubles per stored superoperator ("slot"): D*D dense, Ds*Ds + Da*Da in structured mode
    bool structured = false;   // block-diagonal (swap-symmetric basis) evaluation, see structured.cuh
    SymLayout sym{};
    int nblk = 1; int blkD[2] = {0, 0}; long long blkOff[2] = {0, 0}; int blkTiles[2] = {0, 0}; int blkTileOff[2] = {0, 0};
    double alpha0, alpha1;
    cudaStream_t stream = nullptr;
    // constants
    double* T_re = nullptr;      // [n_targets][SZ]
    double* im2 = nullptr;       // [n_targets]
    int* tindex = nullptr;       // [max_batch]
    unsigned* tab = nullptr; int* inv = nullptr; unsigned* symtab = nullptr; int* t2f = nullptr;
    // work matrices, per instance [m][SZ] unless noted
    double *Y = nullptr, *P = nullptr, *G = nullptr, *W = nullptr, *Yd = nullptr, *Pd = nullptr;
    double *Gd = nullptr;        // [2][SZ]
    double *Wd = nullptr;        // [SZ]; [m][SZ] with the fused chain (all tangent cotangents are formed before the pull-backs run)
    bool cplx = false;           // complex isometries: x is interleaved complex128, the chain runs on the real embedding (complex.cuh)
    bool fused = false;          // per-instance fused chain kernel (chain_fused.cuh) instead of one batched GEMM launch per product
    long long wd_stride = 0;     // per-instance stride of Wd
    // local model: [m][q*q] and per-factor partials [m][nf][q*q]
    double *Yl = nullptr, *Yld = nullptr, *Wl = nullptr, *Wld = nullptr;
    double* be3 = nullptr;       // [max_batch][2][2][256*256]: stage-1 results of the 3-factor back-embedding (dense N = 6)
    // small per-instance state
    double *x = nullptr, *eta = nullptr, *zraw = nullptr, *zdraw = nullptr;   // [m][np]
    double *part_f = nullptr, *part_s = nullptr;                              // [tiles]
    int tiles = 0;
    int cached_batch = 0;        // primal cache valid for this many instances (0 = none)
    // staging for host callers
    DevBuf in_x, in_eta, out_a, out_b, out_c, out_f, out_M;
    TrState tr;                  // trust-region state (tr.cuh)
    Profiler prof;
    bool own_stream = true;
    // host-pointer pipeline (otn_triple): chunked H2D / compute / D2H overlap
    cudaStream_t s_in = nullptr, s_out = nullptr, s_comp[2] = {nullptr, nullptr}, s_side[2] = {nullptr, nullptr}, s_backp[2] = {nullptr, nullptr};
    // Back stream: the pull-backs of the layer cotangents to the isometries (back_layer) only feed the Riemannian epilogue, so they
    // run on their own stream behind the GEMM that produced their input while the chain continues (back_begin / back_sync).
    // Off while per-launch profiling is on (overlapping launches have no additive durations) and inside graph capture.
    cudaStream_t s_back = nullptr, s_back_own = nullptr;
    cudaEvent_t ev_bk_main = nullptr, ev_bk_side = nullptr, ev_bk_done = nullptr;
    bool back_busy = false;
    bool back_wd_pending = false;              // a pull-back reading the (single) Wd buffer is out on the back stream
    // staging slots of the host-array pipeline (inputs + outputs of one chunk each): more slots than compute streams, so that the
    // H2D copy of chunk c+2 does not have to wait for the kernels of chunk c (with the fused chain a 512-instance chunk computes in
    // 3.6 ms and copies in 1.8 ms: two slots left the copy engine and the SMs waiting for each other)
    enum { NSLOT = 4 };
    cudaEvent_t ev_in[NSLOT] = {}, ev_done[NSLOT] = {}, ev_out[NSLOT] = {};
    cudaEvent_t ev_cs_done[2] = {nullptr, nullptr};   // newest chunk of each compute stream
    long long pipe_seq = 0;      // chunks enqueued so far (staging slot = pipe_seq % NSLOT, compute stream = pipe_seq & 1; reuse is ordered across calls)
    // streamed submissions (otn_triple_submit / otn_triple_wait): ticket t completes when ev_ticket[t % 4] has fired
    enum { TICKETS = 4 };
    cudaEvent_t ev_ticket[TICKETS] = {nullptr, nullptr, nullptr, nullptr};
    long long next_ticket = 0;
    bool pipe_inflight = false;
    std::vector<std::pair<int, int>> pipe_unsynced[2];   // [s]: workspace slices (first instance, count) enqueued on compute stream 1-s
                                                         // since stream s last waited for it (triple_pipelined)
    DevBuf pipe_buf;
    DevBuf tr_buf;
    // factored evaluation of the 2-site model (factored.cuh)
    bool factored = false;
    int f_ncols = 0, f_tiles = 0;
    double *f_Tt = nullptr, *f_Pst = nullptr, *f_Wp = nullptr, *f_Wp2 = nullptr, *f_colw = nullptr;
    int f_nsplit = 1;            // two-stage tile reduction when there are many tiles
    int *f_jrow = nullptr, *f_restrow = nullptr, *f_colidx = nullptr, *f_swaprow = nullptr;
    EmbedTables tables() const { return EmbedTables{tab, inv, D, q, nf}; }
    // Row-sharded evaluation of the dense chain over the GPUs of one node (otn_rowshard_attach; BASELINE config 5): every rank
    // holds the full handle and runs the whole program, but each chain product computes only the rows [rank D/world, ...) of its
    // result and stores every tile (and the tile's partial sum) into the same array of all peers through IPC-mapped pointers.
    struct RowShard {
        enum { NARR = 9 };                       // arrays written across ranks: P, G, W, Pd, Gd, Wd, part_f, part_s, be3
        int rank = 0, world = 1, epoch = 0;
        double* base[NARR] = {}; size_t bytes[NARR] = {};
        double* peer[NARR][7] = {};
        int* flags = nullptr;                    // [64]: slot r = last epoch rank r has reached; slot 32 = time-out marker
        int* peer_flags[7] = {};
        bool exported = false;
        cudaStream_t side = nullptr;             // second product of an independent pair (hgemm_pair)
        cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    } rs;
};

// Streamed submissions (otn_triple_submit) leave work in flight on the pipeline streams; every other entry point of the handle
// shares the workspaces with it and therefore drains the pipeline first.
static int drain_pipeline(otn_problem* h) {
    if (!h->pipe_inflight) return 0;
    DeviceGuard dev_guard__(h->device);
    CUDA_TRY(cudaStreamSynchronize(h->s_in));
    for (int i = 0; i < 2; ++i) CUDA_TRY(cudaStreamSynchronize(h->s_comp[i]));
    CUDA_TRY(cudaStreamSynchronize(h->s_out));
    h->pipe_inflight = false;
    for (auto& v : h->pipe_unsynced) v.clear();
    return 0;
}

static int check_batch(otn_problem* h, int batch, bool drain = true) {
    if (h == nullptr) return fail("null handle");
    if (batch < 1 || batch > h->max_batch) return fail("batch %d outside [1, max_batch=%d]", batch, h->max_batch);
    if (drain) return drain_pipeline(h);
    return 0;
}

extern "C" void otn_destroy(otn_problem* h) {
    if (!h) return;
    DeviceGuard dev_guard__(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (int q = 0; q < 7; ++q) {            // row-sharded handle: unmap the peers' arrays
        for (int a = 0; a < otn_problem::RowShard::NARR; ++a) if (h->rs.peer[a][q]) cudaIpcCloseMemHandle(h->rs.peer[a][q]);
        if (h->rs.peer_flags[q]) cudaIpcCloseMemHandle(h->rs.peer_flags[q]);
    }
    if (h->rs.flags) cudaFree(h->rs.flags);
    if (h->rs.side) { cudaStreamSynchronize(h->rs.side); cudaStreamDestroy(h->rs.side); }
    for (cudaEvent_t e : {h->rs.ev_fork, h->rs.ev_join}) if (e) cudaEventDestroy(e);
    if (h->prof.side.st) { cudaStreamSynchronize(h->prof.side.st); cudaStreamDestroy(h->prof.side.st); }
    if (h->s_back_own) { cudaStreamSynchronize(h->s_back_own); cudaStreamDestroy(h->s_back_own); }
    for (int i = 0; i < 2; ++i) if (h->s_backp[i]) { cudaStreamSynchronize(h->s_backp[i]); cudaStreamDestroy(h->s_backp[i]); }
    for (cudaEvent_t e : {h->ev_bk_main, h->ev_bk_side, h->ev_bk_done}) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : {h->prof.side.ev_fork, h->prof.side.ev_join}) if (e) cudaEventDestroy(e);
    for (void* ptr : {(void*)h->T_re, (void*)h->im2, (void*)h->tindex, (void*)h->tab, (void*)h->inv, (void*)h->symtab, (void*)h->t2f, (void*)h->Y, (void*)h->P,
                      (void*)h->G, (void*)h->W, (void*)h->Yd, (void*)h->Pd, (void*)h->Gd, (void*)h->Wd, (void*)h->Yl, (void*)h->Yld,
                      (void*)h->Wl, (void*)h->Wld, (void*)h->be3, (void*)h->x, (void*)h->eta, (void*)h->zraw, (void*)h->zdraw, (void*)h->part_f,
                      (void*)h->part_s, (void*)h->f_Tt, (void*)h->f_Pst, (void*)h->f_Wp, (void*)h->f_Wp2, (void*)h->f_colw, (void*)h->f_jrow,
                      (void*)h->f_restrow, (void*)h->f_colidx, (void*)h->f_swaprow})
        if (ptr) cudaFree(ptr);
    for (DevBuf* b : {&h->in_x, &h->in_eta, &h->out_a, &h->out_b, &h->out_c, &h->out_f, &h->out_M, &h->tr_buf}) b->release();
    for (auto& r : h->prof.recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    for (auto e : h->prof.pool) cudaEventDestroy(e);
    for (int i = 0; i < otn_problem::NSLOT; ++i) for (cudaEvent_t e : {h->ev_in[i], h->ev_done[i], h->ev_out[i]}) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : h->ev_cs_done) if (e) cudaEventDestroy(e);
    if (h->s_in) cudaStreamSynchronize(h->s_in);
    for (int i = 0; i < 2; ++i) if (h->s_comp[i]) cudaStreamSynchronize(h->s_comp[i]);
    if (h->s_out) cudaStreamSynchronize(h->s_out);
    for (cudaEvent_t e : h->ev_ticket) if (e) cudaEventDestroy(e);
    if (h->s_in) cudaStreamDestroy(h->s_in);
    if (h->s_out) cudaStreamDestroy(h->s_out);
    for (int i = 0; i < 2; ++i) if (h->s_comp[i]) cudaStreamDestroy(h->s_comp[i]);
    for (int i = 0; i < 2; ++i) if (h->s_side[i]) { cudaStreamSynchronize(h->s_side[i]); cudaStreamDestroy(h->s_side[i]); }
    h->pipe_buf.release();
    if (h->stream && h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" int otn_create(otn_problem** out, int N, int d, int m, int K, int model, int metric, int max_batch, int n_targets,
                          const double* target_re, const double* target_im, int device) {
    return otn_create_ex(out, N, d, m, K, model, metric, max_batch, n_targets, target_re, target_im, device, 0);
}

extern "C" int otn_create_ex(otn_problem** out, int N, int d, int m, int K, int model, int metric, int max_batch, int n_targets,
                             const double* target_re, const double* target_im, int device, int flags) {
    if (out == nullptr) return fail("out is null");
    *out = nullptr;
    if (d < 2 || N < 1 || m < 1 || K < 1 || max_batch < 1 || n_targets < 1) return fail("bad size argument");
    if (model != OTN_MODEL_FULL && model != OTN_MODEL_LOCAL) return fail("model must be OTN_MODEL_FULL or OTN_MODEL_LOCAL");
    if (model == OTN_MODEL_LOCAL && (N % 2 != 0)) return fail("Only even number of sites allowed");  // transformations.py:414
    if (target_re == nullptr) return fail("target_re is null");
    if (K > ASM_MAXK) return fail("Kraus rank %d > %d not supported", K, ASM_MAXK);
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail("device %d not available (%d CUDA devices)", device, ndev);
    DeviceGuard dev_guard__(device);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail("libotn is built for sm_100a (B200); device %d is sm_%d%d", device, prop.major, prop.minor);

    double dD = std::pow((double)d, 2.0 * N);
    if (dD > 16384.0) return fail("D = d^(2N) too large");
    otn_problem* h = new otn_problem();
    h->N = N; h->d = d; h->m = m; h->K = K; h->model = model; h->max_batch = max_batch; h->n_targets = n_targets; h->device = device;
    h->D = ipow(d, 2 * N);
    h->cplx = (flags & OTN_FLAG_COMPLEX) != 0;
    if (h->cplx && model != OTN_MODEL_FULL) { delete h; return fail("complex isometries: full-sites model only (OTN_MODEL_FULL)"); }
    if (h->cplx && (flags & (OTN_FLAG_BLOCKS | OTN_FLAG_FACTORED))) { delete h; return fail("OTN_FLAG_COMPLEX excludes OTN_FLAG_BLOCKS / OTN_FLAG_FACTORED"); }
    const int chainD = h->cplx ? 2 * h->D : h->D;      // side of the matrices the chain multiplies (real embedding of complex ones)
    h->SZ = (long long)chainD * chainD;
    h->q = ipow(d, 4);
    h->nf = N / 2;
    h->dim = (model == OTN_MODEL_FULL) ? ipow(d, N) : d * d;
    h->nblk = 1; h->blkD[0] = chainD; h->blkOff[0] = 0;
    const int sdim = ipow(d, N);   // side of the doubled index (a,b): D = sdim^2
    // block-diagonal evaluation: 3.3x fewer chain flops, but two launches per product with fewer CTAs each -- for a handful of
    // small instances the path is launch/latency bound and the dense chain is faster (measured at D = 256, batch 1: 2.1 vs 3.2 ms
    // per U1), so AUTO picks blocks only when there is enough work to fill the GPU
    const bool blocks_possible = !h->cplx && ((model == OTN_MODEL_FULL && (h->dim == 16 || h->dim == 8)) || (model == OTN_MODEL_LOCAL && sdim >= 8 && sdim <= 64));
    const bool enough_work = (long long)max_batch * h->D >= 4096;
    h->structured = blocks_possible && !(flags & OTN_FLAG_DENSE) && ((flags & OTN_FLAG_BLOCKS) || enough_work);
    const bool fact_possible = model == OTN_MODEL_LOCAL && d == 2 && (N == 4 || N == 6);
    if ((flags & OTN_FLAG_FACTORED) && !fact_possible) {
        delete h;
        return fail("OTN_FLAG_FACTORED needs the 2-site model with d = 2 and N = 4 or 6");
    }
    if ((flags & OTN_FLAG_FACTORED) && (flags & (OTN_FLAG_DENSE | OTN_FLAG_BLOCKS))) {
        delete h;
        return fail("OTN_FLAG_FACTORED excludes OTN_FLAG_DENSE / OTN_FLAG_BLOCKS");
    }
    // AUTO picks the factored evaluation whenever it applies: measured 2.7x (N = 4, 1024 instances, m = 9) to 19x (N = 6, one
    // instance) faster than the block-diagonal chain, and 2.6x faster at batch 1 (profiles/README.md)
    if ((flags & OTN_FLAG_FACTORED) || (fact_possible && !(flags & (OTN_FLAG_DENSE | OTN_FLAG_BLOCKS)))) {
        h->factored = true;
        h->structured = false;
        // N = 4: 8-column tiles (17 per instance); with only a handful of instances 2-column tiles give 4x more, smaller CTAs
        h->f_ncols = (N == 4) ? (max_batch <= 4 ? 2 : 8) : 1;
        h->f_tiles = (sdim * (sdim + 1) / 2 + h->f_ncols - 1) / h->f_ncols;
    }
    if (h->structured) {
        SymLayout& L = h->sym;
        L.dim = sdim; L.NS = sdim * (sdim + 1) / 2; L.NA = sdim * (sdim - 1) / 2;
        auto pad = [](int n) { const int p48 = (n + 47) / 48 * 48, p64 = (n + 63) / 64 * 64; return p48 < p64 ? p48 : p64; };
        L.Ds = pad(L.NS); L.Da = pad(L.NA);
        L.off_a = (long long)L.Ds * L.Ds; L.slot = L.off_a + (long long)L.Da * L.Da;
        h->SZ = L.slot;
        h->nblk = 2; h->blkD[0] = L.Ds; h->blkD[1] = L.Da; h->blkOff[0] = 0; h->blkOff[1] = L.off_a;
    }
    // fused per-instance chain: block-diagonal evaluation with blocks the fused kernel is instantiated for (144 + 128: N = 4),
    // m >= 2, and enough instances to give every SM a CTA (below that the tiled kernels spread one product over more SMs)
    {
        const char* env = getenv("OTN_FUSED_CHAIN");     // A/B switch for measurements and tests: "0" = never, "1" = always
        const bool want = env ? (env[0] != '0') : max_batch >= 64;
        h->fused = want && h->structured && m >= 2 && m <= 15 && chain_fused_supports(h->blkD[0]) && chain_fused_supports(h->blkD[1]);
    }
    h->tiles = 0;
    for (int i = 0; i < h->nblk; ++i) {
        h->blkTiles[i] = h->fused ? chain_fused_partials(h->blkD[i]) : gemm_plan(h->blkD[i], h->cplx ? h->blkD[i] / 2 : h->blkD[i], 1).tiles;
        h->blkTileOff[i] = h->tiles;
        h->tiles += h->blkTiles[i];
    }
    if (h->factored) h->tiles = h->f_tiles;
    h->p = h->dim; h->n = h->dim * K; h->np = h->n * h->p * (h->cplx ? 2 : 1);   // doubles per isometry
    if (h->p > ST_MAXP) { const int pcols = h->p; delete h; return fail("isometry with p = %d columns > %d not supported", pcols, ST_MAXP); }
    if (model == OTN_MODEL_LOCAL && (h->nf > 4 || h->q > 255)) { delete h; return fail("local model supports N <= 8 and d <= 3"); }
    if ((long long)max_batch * m > 65535) { delete h; return fail("max_batch * m must be <= 65535 (split the batch)"); }
    *out = h;  // from here on failures go through otn_destroy by the caller? no: clean up ourselves
    auto bail = [&](int rc) { otn_destroy(h); *out = nullptr; return rc; };
#define CREATE_TRY(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { fail("%s failed: %s", #expr, cudaGetErrorString(e__)); return bail(-1); } } while (0)
    if (otn_set_metric(h, metric) != 0) return bail(-1);
    CREATE_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    if (h->nblk == 2 && getenv("OTN_NO_SIDE_STREAM") == nullptr) {     // env: A/B switch for measurements
        SideStream& sd = h->prof.side;
        CREATE_TRY(cudaStreamCreateWithFlags(&sd.st, cudaStreamNonBlocking));
        CREATE_TRY(cudaEventCreateWithFlags(&sd.ev_fork, cudaEventDisableTiming));
        CREATE_TRY(cudaEventCreateWithFlags(&sd.ev_join, cudaEventDisableTiming));
        sd.enabled = true;
    }
    if (getenv("OTN_NO_BACK_STREAM") == nullptr) {                      // env: A/B switch for measurements
        CREATE_TRY(cudaStreamCreateWithFlags(&h->s_back_own, cudaStreamNonBlocking));
        CREATE_TRY(cudaEventCreateWithFlags(&h->ev_bk_main, cudaEventDisableTiming));
        CREATE_TRY(cudaEventCreateWithFlags(&h->ev_bk_side, cudaEventDisableTiming));
        CREATE_TRY(cudaEventCreateWithFlags(&h->ev_bk_done, cudaEventDisableTiming));
        h->s_back = h->s_back_own;
    }
    const size_t B = (size_t)max_batch, SZ = (size_t)h->SZ, M = (size_t)m;
    // targets.  Every layer superoperator Y = sum_k E_k (x) E_k (and hence every chain product and tangent) is invariant under
    // the swap S: (a,b) <-> (b,a) applied to rows and columns, so only the swap-symmetric part Ts = (T + S T S)/2 of Re T
    // carries gradient; the antisymmetric rest and Im T add the constant ||Re T - Ts||^2 + ||Im T||^2 under the square root.
    // Running the reverse chain on R' = M - Ts makes every cotangent W swap-symmetric, which halves the back-assembly traffic.
    CREATE_TRY(cudaMalloc(&h->T_re, n_targets * SZ * 8));
    {
        const size_t DD = (size_t)h->D * h->D;
        std::vector<double> im2(n_targets, 0.0), tre(DD), tsym(DD), tmp(DD), tblk(SZ);
        const int dm = ipow(d, N), Dn = h->D;
        for (int t = 0; t < n_targets && h->cplx; ++t) {
            // emb(T) = [[Tr, -Ti], [Ti, Tr]]; no swap symmetrisation (complex layers are swap-Hermitian, not swap-symmetric)
            std::vector<double> emb(SZ);
            CREATE_TRY(cudaMemcpy(tre.data(), target_re + (size_t)t * DD, DD * 8, cudaMemcpyDefault));
            std::fill(tmp.begin(), tmp.end(), 0.0);
            if (target_im != nullptr) CREATE_TRY(cudaMemcpy(tmp.data(), target_im + (size_t)t * DD, DD * 8, cudaMemcpyDefault));
            const size_t Dn2 = 2 * (size_t)Dn;
            for (int r = 0; r < Dn; ++r)
                for (int c = 0; c < Dn; ++c) {
                    const double re = tre[(size_t)r * Dn + c], im = tmp[(size_t)r * Dn + c];
                    emb[(size_t)r * Dn2 + c] = re; emb[(size_t)r * Dn2 + Dn + c] = -im;
                    emb[((size_t)r + Dn) * Dn2 + c] = im; emb[((size_t)r + Dn) * Dn2 + Dn + c] = re;
                }
            CREATE_TRY(cudaMemcpy(h->T_re + (size_t)t * SZ, emb.data(), SZ * 8, cudaMemcpyHostToDevice));
            im2[t] = 0.0;
        }
        for (int t = 0; t < n_targets && !h->cplx; ++t) {
            CREATE_TRY(cudaMemcpy(tre.data(), target_re + (size_t)t * DD, DD * 8, cudaMemcpyDefault));
            double s = 0.0;
            for (int r = 0; r < Dn; ++r) {
                const int rs = (r % dm) * dm + r / dm;
                for (int c = 0; c < Dn; ++c) {
                    const int cs = (c % dm) * dm + c / dm;
                    const double v = 0.5 * (tre[(size_t)r * Dn + c] + tre[(size_t)rs * Dn + cs]);
                    const double a = tre[(size_t)r * Dn + c] - v;
                    tsym[(size_t)r * Dn + c] = v;
                    s += a * a;
                }
            }
            if (h->factored) {
                // tile-column layout of the columns (c1 <= c2): Tt[t][tile][r][c]
                const int nc = h->f_ncols, tiles = h->f_tiles;
                if (t == 0) CREATE_TRY(cudaMalloc(&h->f_Tt, (size_t)n_targets * tiles * Dn * nc * 8));
                std::vector<double> tt((size_t)tiles * Dn * nc, 0.0);
                int k = 0;
                for (int c1 = 0; c1 < dm; ++c1)
                    for (int c2 = c1; c2 < dm; ++c2, ++k) {
                        const int col = c1 * dm + c2, tile = k / nc, c = k % nc;
                        for (int r = 0; r < Dn; ++r) tt[((size_t)tile * Dn + r) * nc + c] = tsym[(size_t)r * Dn + col];
                    }
                CREATE_TRY(cudaMemcpy(h->f_Tt + (size_t)t * tt.size(), tt.data(), tt.size() * 8, cudaMemcpyHostToDevice));
            }
            if (!h->structured) {
                CREATE_TRY(cudaMemcpy(h->T_re + (size_t)t * SZ, tsym.data(), SZ * 8, cudaMemcpyHostToDevice));
            } else {
                // blocks of the symmetrised target: Ts[{ab},{cd}] = u^s_ab . T u^s_cd, Ta[{ab},{cd}] = u^a_ab . T u^a_cd
                const SymLayout& L = h->sym;
                std::fill(tblk.begin(), tblk.end(), 0.0);
                auto at = [&](int a, int b, int c, int e) { return tsym[(size_t)(a * dm + b) * Dn + (c * dm + e)]; };
                const double h2 = std::sqrt(0.5);
                for (int a = 0; a < dm; ++a) for (int b = a; b < dm; ++b)
                    for (int c = 0; c < dm; ++c) for (int e = c; e < dm; ++e) {
                        // tsym is swap symmetric: T[(b,a),(e,c)] = T[(a,b),(c,e)]
                        const double na = (a == b) ? h2 : 1.0, nc = (c == e) ? h2 : 1.0;
                        tblk[(size_t)pair_s(a, b, dm) * L.Ds + pair_s(c, e, dm)] = na * nc * (at(a, b, c, e) + at(a, b, e, c));
                        if (a < b && c < e)
                            tblk[L.off_a + (size_t)pair_a(a, b, dm) * L.Da + pair_a(c, e, dm)] = at(a, b, c, e) - at(a, b, e, c);
                    }
                CREATE_TRY(cudaMemcpy(h->T_re + (size_t)t * SZ, tblk.data(), SZ * 8, cudaMemcpyHostToDevice));
            }
            if (target_im != nullptr) {
                CREATE_TRY(cudaMemcpy(tmp.data(), target_im + (size_t)t * DD, DD * 8, cudaMemcpyDefault));
                for (size_t i = 0; i < DD; ++i) s += tmp[i] * tmp[i];
            }
            im2[t] = s;
        }
        CREATE_TRY(cudaMalloc(&h->im2, n_targets * 8));
        CREATE_TRY(cudaMemcpy(h->im2, im2.data(), n_targets * 8, cudaMemcpyHostToDevice));
    }
    CREATE_TRY(cudaMalloc(&h->tindex, B * 4));
    CREATE_TRY(cudaMemset(h->tindex, 0, B * 4));
    if (model == OTN_MODEL_LOCAL) {
        std::vector<unsigned> tab; std::vector<int> inv;
        build_embed_tables(N, d, tab, inv);
        CREATE_TRY(cudaMalloc(&h->tab, tab.size() * 4));
        CREATE_TRY(cudaMemcpy(h->tab, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
        CREATE_TRY(cudaMalloc(&h->inv, inv.size() * 4));
        CREATE_TRY(cudaMemcpy(h->inv, inv.data(), inv.size() * 4, cudaMemcpyHostToDevice));
        if (h->structured) {
            std::vector<unsigned> st(h->D);
            for (int r = 0; r < h->D; ++r) st[r] = sym_pack(r / sdim, r % sdim, sdim);
            CREATE_TRY(cudaMalloc(&h->symtab, st.size() * 4));
            CREATE_TRY(cudaMemcpy(h->symtab, st.data(), st.size() * 4, cudaMemcpyHostToDevice));
            // tensored index (loc_0, loc_1, ...) -> full index, per shift
            std::vector<int> t2f(2 * (size_t)h->D);
            for (int sh = 0; sh < 2; ++sh)
                for (int r = 0; r < h->D; ++r) {
                    int rt = 0;
                    for (int f = 0; f < h->nf; ++f) rt = rt * h->q + (int)((tab[(size_t)sh * h->D + r] >> (8 * f)) & 0xff);
                    t2f[(size_t)sh * h->D + rt] = r;
                }
            CREATE_TRY(cudaMalloc(&h->t2f, t2f.size() * 4));
            CREATE_TRY(cudaMemcpy(h->t2f, t2f.data(), t2f.size() * 4, cudaMemcpyHostToDevice));
        }
        if (h->factored) {
            const int NF = h->nf, Dn = h->D, REST = Dn / 16, nc = h->f_ncols;
            std::vector<int> jrow(2 * NF * 16, -1), restrow((size_t)2 * NF * REST, 0);
            for (int sh = 0; sh < 2; ++sh)
                for (int f = 0; f < NF; ++f) {
                    auto loc = [&](int r, int g) { return (int)((tab[(size_t)sh * Dn + r] >> (8 * g)) & 0xff); };
                    int k = 0;
                    for (int r = 0; r < Dn; ++r) {
                        bool others0 = true;
                        for (int g = 0; g < NF; ++g) if (g != f && loc(r, g) != 0) others0 = false;
                        if (others0) jrow[(sh * NF + f) * 16 + loc(r, f)] = r;
                        if (loc(r, f) == 0) restrow[(size_t)(sh * NF + f) * REST + k++] = r;
                    }
                    // the gather tables rely on the index map being a digit permutation: row = rest row + local-digit offset
                    for (int kk = 0; kk < REST; ++kk)
                        for (int j = 0; j < 16; ++j) {
                            const int r0 = restrow[(size_t)(sh * NF + f) * REST + kk], r = r0 + jrow[(sh * NF + f) * 16 + j];
                            bool ok = r >= 0 && r < Dn && loc(r, f) == j;
                            for (int g = 0; ok && g < NF; ++g) if (g != f && loc(r, g) != loc(r0, g)) ok = false;
                            if (!ok) { fail("factored evaluation: embedding is not a digit permutation"); return bail(-1); }
                        }
                }
            std::vector<int> colidx((size_t)h->f_tiles * nc, 0), swaprow(Dn);
            std::vector<double> colw((size_t)h->f_tiles * nc, 0.0);
            int k = 0;
            for (int c1 = 0; c1 < sdim; ++c1)
                for (int c2 = c1; c2 < sdim; ++c2, ++k) { colidx[k] = c1 * sdim + c2; colw[k] = (c1 == c2) ? 0.5 : 1.0; }
            for (int r = 0; r < Dn; ++r) swaprow[r] = (r % sdim) * sdim + r / sdim;
            CREATE_TRY(cudaMalloc(&h->f_jrow, jrow.size() * 4));
            CREATE_TRY(cudaMemcpy(h->f_jrow, jrow.data(), jrow.size() * 4, cudaMemcpyHostToDevice));
            CREATE_TRY(cudaMalloc(&h->f_restrow, restrow.size() * 4));
            CREATE_TRY(cudaMemcpy(h->f_restrow, restrow.data(), restrow.size() * 4, cudaMemcpyHostToDevice));
            CREATE_TRY(cudaMalloc(&h->f_colidx, colidx.size() * 4));
            CREATE_TRY(cudaMemcpy(h->f_colidx, colidx.data(), colidx.size() * 4, cudaMemcpyHostToDevice));
            CREATE_TRY(cudaMalloc(&h->f_swaprow, swaprow.size() * 4));
            CREATE_TRY(cudaMemcpy(h->f_swaprow, swaprow.data(), swaprow.size() * 4, cudaMemcpyHostToDevice));
            CREATE_TRY(cudaMalloc(&h->f_colw, colw.size() * 8));
            CREATE_TRY(cudaMemcpy(h->f_colw, colw.data(), colw.size() * 8, cudaMemcpyHostToDevice));
            const size_t SBs = (size_t)Dn * nc;
            if (m > 1) CREATE_TRY(cudaMalloc(&h->f_Pst, B * h->f_tiles * (M - 1) * 2 * SBs * 8));
            CREATE_TRY(cudaMalloc(&h->f_Wp, B * h->f_tiles * M * NF * 512 * 8));
            if (h->f_tiles > 64) {
                h->f_nsplit = (h->f_tiles + 31) / 32;
                CREATE_TRY(cudaMalloc(&h->f_Wp2, B * h->f_nsplit * M * NF * 512 * 8));
            }
        }
        const size_t q2 = (size_t)h->q * h->q;
        CREATE_TRY(cudaMalloc(&h->Yl, B * M * q2 * 8));
        CREATE_TRY(cudaMalloc(&h->Yld, B * M * q2 * 8));
        CREATE_TRY(cudaMalloc(&h->Wl, B * M * h->nf * q2 * 8));
        CREATE_TRY(cudaMalloc(&h->Wld, B * M * h->nf * q2 * 8));
        if (!h->structured && !h->factored && h->nf == 3 && h->q == 16 && getenv("OTN_BACK_EMBED_V1") == nullptr)   // env: A/B switch
            CREATE_TRY(cudaMalloc(&h->be3, B * 4 * 65536 * 8));
    }
    if (!h->factored) {        // the factored evaluation never materialises a D x D work matrix
        for (double** ptr : {&h->Y, &h->P, &h->G, &h->W, &h->Yd, &h->Pd}) {
            CREATE_TRY(cudaMalloc(ptr, B * M * SZ * 8));
            if (h->structured) CREATE_TRY(cudaMemset(*ptr, 0, B * M * SZ * 8));   // zero padding of the blocks is an invariant
        }
        CREATE_TRY(cudaMalloc(&h->Gd, B * 2 * SZ * 8));
        h->wd_stride = (h->fused ? (long long)m : 1LL) * h->SZ;
        CREATE_TRY(cudaMalloc(&h->Wd, B * (size_t)h->wd_stride * 8));
    }
    if (h->structured) { CREATE_TRY(cudaMemset(h->Gd, 0, B * 2 * SZ * 8)); CREATE_TRY(cudaMemset(h->Wd, 0, B * (size_t)h->wd_stride * 8)); }
    for (double** ptr : {&h->x, &h->eta, &h->zraw, &h->zdraw}) CREATE_TRY(cudaMalloc(ptr, B * M * h->np * 8));
    CREATE_TRY(cudaMalloc(&h->part_f, B * h->tiles * 8));
    CREATE_TRY(cudaMalloc(&h->part_s, B * h->tiles * 8));
#undef CREATE_TRY
    return 0;
}

extern "C" int otn_set_metric(otn_problem* h, int metric) {
    if (!h) return fail("null handle");
    if (metric == OTN_METRIC_EUCLIDEAN) { h->alpha0 = 1.0; h->alpha1 = 1.0; }         // stiefel.py:456-457
    else if (metric == OTN_METRIC_CANONICAL) { h->alpha0 = 1.0; h->alpha1 = 0.5; }    // stiefel.py:458-459
    else return fail("%d is not a valid metric value", metric);
    h->metric = metric;
    return 0;
}

// general two-parameter family of metrics g(D1, D2) = tr(D1^T (alpha0 (I - X X^T) + alpha1 X X^T) D2) (stiefel.py:398-426): every
// kernel of the Riemannian epilogue carries (alpha0, alpha1), the presets above are two points of the family
extern "C" int otn_set_metric_general(otn_problem* h, double alpha0, double alpha1) {
    if (!h) return fail("null handle");
    if (!(alpha0 > 0.0) || !(alpha1 > 0.0) || std::isinf(alpha0) || std::isinf(alpha1)) return fail("alpha0 and alpha1 must be positive and finite");
    OTN_TRY(drain_pipeline(h));
    h->alpha0 = alpha0; h->alpha1 = alpha1;
    h->metric = (alpha0 == 1.0 && alpha1 == 1.0) ? OTN_METRIC_EUCLIDEAN : (alpha0 == 1.0 && alpha1 == 0.5) ? OTN_METRIC_CANONICAL : -1;
    return 0;
}

extern "C" int otn_set_target_index(otn_problem* h, const int* index, int batch) {
    OTN_TRY(check_batch(h, batch));
    DeviceGuard dev_guard__(h->device);
    std::vector<int> host(batch);
    CUDA_TRY(cudaMemcpy(host.data(), index, batch * 4, cudaMemcpyDefault));
    for (int b = 0; b < batch; ++b)
        if (host[b] < 0 || host[b] >= h->n_targets) return fail("target index %d of instance %d outside [0, %d)", host[b], b, h->n_targets);
    CUDA_TRY(cudaMemcpy(h->tindex, host.data(), batch * 4, cudaMemcpyHostToDevice));
    return 0;
}

extern "C" int otn_query(const otn_problem* h, int* n, int* pp, int* D, int* m, long long* ws) {
    if (!h) return fail("null handle");
    if (n) *n = h->n;
    if (pp) *pp = h->p;
    if (D) *D = h->D;
    if (m) *m = h->m;
    if (ws) *ws = h->factored ? ((long long)h->f_tiles * (2LL * (h->m - 1) * h->D * h->f_ncols + (long long)h->m * h->nf * 512)) * 8 + 4LL * h->m * h->np * 8
                              : (6LL * h->m + 3) * h->SZ * 8 + 4LL * h->m * h->np * 8;
    return 0;
}

extern "C" void* otn_stream(const otn_problem* h) { return h ? (void*)h->stream : nullptr; }

extern "C" int otn_set_stream(otn_problem* h, void* stream) {
    if (!h) return fail("null handle");
    DeviceGuard dev_guard__(h->device);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (h->own_stream && h->stream) CUDA_TRY(cudaStreamDestroy(h->stream));
    h->stream = (cudaStream_t)stream;
    h->own_stream = false;
    return 0;
}

extern "C" int otn_profile_enable(otn_problem* h, int on) {
    if (!h) return fail("null handle");
    h->prof.on = on != 0;
    return 0;
}

// sums (and clears) the per-category event timings recorded since the last read: ms[6], launches[6], flops[6]
extern "C" int otn_profile_read(otn_problem* h, double* ms, long long* launches, double* flops) {
    if (!h) return fail("null handle");
    DeviceGuard dev_guard__(h->device);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    for (int c = 0; c < CAT_COUNT; ++c) { if (ms) ms[c] = 0; if (launches) launches[c] = 0; if (flops) flops[c] = 0; }
    for (auto& r : h->prof.recs) {
        float t = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&t, r.e0, r.e1));
        if (ms) ms[r.cat] += t;
        if (launches) launches[r.cat] += r.n;
        if (flops) flops[r.cat] += r.flops;
        h->prof.pool.push_back(r.e0); h->prof.pool.push_back(r.e1);
    }
    h->prof.recs.clear();
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// pipeline pieces (all enqueue on h->stream; x / eta are device pointers [batch][m][np])
// ---------------------------------------------------------------------------------------------------------------------
// back stream (see otn_problem::s_back)
static bool back_overlap(const otn_problem* h) {
    return h->s_back != nullptr && !h->prof.on && !h->prof.side.suspended && !h->factored && h->rs.world == 1;   // row-sharded: one stream, see rs_barrier
}
// the back stream waits for everything enqueued so far on the main stream and, if the antisymmetric chain is out on it, the side stream
static cudaStream_t back_begin(otn_problem* h) {
    cudaEventRecord(h->ev_bk_main, h->stream);
    cudaStreamWaitEvent(h->s_back, h->ev_bk_main, 0);
    if (h->prof.side.busy) {
        cudaEventRecord(h->ev_bk_side, h->prof.side.st);
        cudaStreamWaitEvent(h->s_back, h->ev_bk_side, 0);
    }
    h->back_busy = true;
    return h->s_back;
}
// the main stream (and a busy side stream) wait for the pull-backs enqueued so far
static void back_sync(otn_problem* h) {
    if (!h->back_busy) return;
    cudaEventRecord(h->ev_bk_done, h->s_back);
    cudaStreamWaitEvent(h->stream, h->ev_bk_done, 0);
    if (h->prof.side.busy) cudaStreamWaitEvent(h->prof.side.st, h->ev_bk_done, 0);
    h->back_busy = false;
    h->back_wd_pending = false;
}

// Row-sharded handles: device-side barrier between the ranks.  Every rank raises its slot in the flag array of every peer and spins
// until all slots of its own array have reached the epoch (the same scheme as otn_peer_signal / otn_peer_wait, one kernel).  The
// kernel before it on the stream has written its tiles into the peers' arrays: the system-scope fence orders those stores before
// the flag.  A rank whose peers never arrive (a host-side failure over there) gives up after 30 s and marks slot 32 instead of
// hanging the GPU; otn_rowshard_status reports it.
struct RsPeerFlags { int* p[7]; };
__global__ void rs_barrier_kernel(RsPeerFlags peers, volatile int* mine, int world, int rank, int epoch) {
    __threadfence_system();
    const int t = threadIdx.x;
    if (t < world - 1) { volatile int* f = peers.p[t] + rank; *f = epoch; }
    __threadfence_system();
    if (t < world && t != rank) {
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (mine[t] < epoch) {
            __nanosleep(100);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 30000000000ULL) { mine[32] = epoch; break; }
        }
    }
    __threadfence_system();
}
static cudaError_t rs_barrier(otn_problem* h, cudaStream_t st) {
    RsPeerFlags P{};
    for (int q = 0; q < h->rs.world - 1; ++q) P.p[q] = h->rs.peer_flags[q];
    rs_barrier_kernel<<<1, 32, 0, st>>>(P, h->rs.flags, h->rs.world, h->rs.rank, ++h->rs.epoch);
    ++g_launches;
    return cudaGetLastError();
}
// the same position inside peer q's copy of the array `ptr` points into
static double* rs_peer(const otn_problem* h, int q, const double* ptr) {
    for (int a = 0; a < otn_problem::RowShard::NARR; ++a) {
        const char* lo = reinterpret_cast<const char*>(h->rs.base[a]);
        const char* c = reinterpret_cast<const char*>(ptr);
        if (lo != nullptr && c >= lo && c < lo + h->rs.bytes[a]) return reinterpret_cast<double*>(reinterpret_cast<char*>(h->rs.peer[a][q]) + (c - lo));
    }
    return nullptr;
}

// row-sharded handle: row window of this rank, tiles (and partial sums) also stored into every peer
static cudaError_t rs_window(const otn_problem* h, GemmParams& p) {
    const int rows = p.D / h->rs.world;
    p.row0 = h->rs.rank * rows; p.rows = rows;
    if (p.epi != EPI_NONE) p.partial += h->rs.rank * gemm_plan(p.D, rows, 1).tiles;
    p.npeer = h->rs.world - 1;
    if (getenv("OTN_RS_NO_PEER_STORES") != nullptr) p.npeer = 0;      // env: timing diagnostic only (results are then incomplete)
    for (int q = 0; q < p.npeer; ++q) {
        p.Cpeer[q] = rs_peer(h, q, p.C);
        p.Ppeer[q] = (p.epi != EPI_NONE) ? rs_peer(h, q, p.partial) : nullptr;
        if (p.Cpeer[q] == nullptr || (p.epi != EPI_NONE && p.Ppeer[q] == nullptr)) return cudaErrorInvalidValue;
    }
    return cudaSuccess;
}
static double gemm_flops(const GemmParams& p) { return 2.0 * p.D * (double)p.rows * (p.kvalid > 0 ? p.kvalid : p.D) * p.batch * (p.A2 ? 2 : 1); }

// one chain product = one launch per diagonal block (1 block in the dense formulation, 2 in the structured one)
static cudaError_t hgemm(otn_problem* h, const GemmParams& p0, bool ta, bool tb) {
    for (int i = 0; i < h->nblk; ++i) {
        GemmParams p = p0;
        const long long o = h->blkOff[i];
        p.D = h->blkD[i]; p.row0 = 0; p.rows = h->blkD[i];
        p.kvalid = h->structured ? (i == 0 ? h->sym.NS : h->sym.NA) : 0;
        p.A1 += o; p.B1 += o; p.C += o;
        if (p.A2) { p.A2 += o; p.B2 += o; }
        if (p.epi != EPI_NONE) { p.E += o; p.partial += h->blkTileOff[i]; p.partial_stride = h->tiles; }
        if (h->cplx) { p.rows = p.D / 2; p.mirror = p.D / 2; }     // products of embeddings: upper half computed, lower half mirrored (gemm_dmma.cuh)
        if (!h->structured) p.splitk = gemm_want_split(p.D);       // large single-instance products: split-K CTA pairs
        if (h->rs.world > 1) {
            // Barrier BEFORE the product: no peer may still be reading (in a replicated kernel enqueued since the last product)
            // the array this product overwrites over there; barrier AFTER it: the result is complete on every rank before
            // anything reads it.  Everything of a row-sharded handle is ordered on h->stream (no back stream; the side stream of
            // hgemm_pair forks and joins between the two barriers), so stream order on each rank plus the barriers is the whole
            // ordering argument.
            cudaError_t eb = rs_window(h, p);
            if (eb == cudaSuccess) eb = rs_barrier(h, h->stream);
            if (eb != cudaSuccess) return eb;
        }
        // executed flops: zero-padded k-steps beyond kvalid are skipped by the kernel
        cudaStream_t st = h->stream;
        if (h->nblk == 2 && h->prof.side.active()) {
            h->prof.fork(h->stream);                  // (first product after a join; no-op while the side stream is busy)
            if (i == 1) st = h->prof.side.st;
        }
        {
            ProfScope ps(&h->prof, st, CAT_GEMM, gemm_flops(p));
            cudaError_t e = gemm_launch(p, ta, tb, st);
            if (e != cudaSuccess) return e;
        }
        if (h->rs.world > 1) {
            cudaError_t eb = rs_barrier(h, h->stream);
            if (eb != cudaSuccess) return eb;
        }
    }
    return cudaSuccess;
}

// Two chain products that do not depend on each other (W_l = G_l P_{l-1}^T and G_{l-1} = Y_l^T G_l of the reverse sweeps).  On a
// row-sharded handle they share ONE barrier pair and run concurrently on two streams: at 8 GPUs one product is 512 tiles of
// 64 x 64 -- 3.46 per SM, so a quarter of the SMs would carry a fourth tile while the others idle -- and the peer stores of one
// product's epilogues overlap the other's math.  Everywhere else: the two products one after the other.
static cudaError_t hgemm_pair(otn_problem* h, const GemmParams& a, bool ta_a, bool tb_a, const GemmParams& b, bool ta_b, bool tb_b) {
    if (h->rs.world == 1 || h->rs.side == nullptr || h->prof.on || h->nblk != 1 || getenv("OTN_RS_NO_PAIRS") != nullptr) {   // env: A/B switch
        cudaError_t e = hgemm(h, a, ta_a, tb_a);
        return e != cudaSuccess ? e : hgemm(h, b, ta_b, tb_b);
    }
    GemmParams p[2] = {a, b};
    for (GemmParams& q : p) {
        q.D = h->blkD[0]; q.kvalid = 0; q.splitk = gemm_want_split(q.D);
        if (q.epi != EPI_NONE) q.partial_stride = h->tiles;
        cudaError_t e = rs_window(h, q);
        if (e != cudaSuccess) return e;
    }
    cudaError_t e = rs_barrier(h, h->stream);
    if (e != cudaSuccess) return e;
    cudaEventRecord(h->rs.ev_fork, h->stream);
    cudaStreamWaitEvent(h->rs.side, h->rs.ev_fork, 0);
    g_launches += 2;
    e = gemm_launch(p[0], ta_a, tb_a, h->stream);
    if (e != cudaSuccess) return e;
    e = gemm_launch(p[1], ta_b, tb_b, h->rs.side);
    if (e != cudaSuccess) return e;
    cudaEventRecord(h->rs.ev_join, h->rs.side);
    cudaStreamWaitEvent(h->stream, h->rs.ev_join, 0);
    return rs_barrier(h, h->stream);
}
static GemmParams gp(const otn_problem* h, int batch, const int* active) {
    GemmParams p{};
    p.D = h->D; p.row0 = 0; p.rows = h->D; p.batch = batch; p.active = active; p.epi = EPI_NONE;
    return p;
}
static inline double* layer_ptr(double* base, int l, long long SZ) { return base + (long long)l * SZ; }

// fused chain (chain_fused.cuh): the whole product sequence of `mode` for every instance, one launch per diagonal block
// (symmetric block on the handle's stream, antisymmetric one on the side stream, like hgemm)
static int hchain(otn_problem* h, int batch, const int* active, int mode) {
    ChainProgram prog;
    int gd_final = 0;
    if (!chain_build(mode, h->m, prog, &gd_final)) return fail("fused chain: cannot build the op list for m = %d", h->m);
    if (mode == CHAIN_TANGENT || mode == CHAIN_TRIPLE) h->tr.gd_final = gd_final;
    const long long SZ = h->SZ, S = (long long)h->m * SZ;
    for (int i = 0; i < h->nblk; ++i) {
        ChainParams p{};
        const long long o = h->blkOff[i];
        double* bases[CH_NARR] = {h->Y, h->P, h->G, h->W, h->Yd, h->Pd, h->Gd, h->Wd};
        const long long strides[CH_NARR] = {S, S, S, S, S, S, 2 * SZ, h->wd_stride};
        for (int a = 0; a < CH_NARR; ++a) { p.base[a] = bases[a] + o; p.stride[a] = strides[a]; p.lstride[a] = SZ; }
        p.D = h->blkD[i]; p.kvalid = (i == 0 ? h->sym.NS : h->sym.NA);
        p.T = h->T_re + o; p.sT = SZ; p.tindex = h->tindex;
        p.part_f = h->part_f + h->blkTileOff[i]; p.part_s = h->part_s + h->blkTileOff[i]; p.part_stride = h->tiles;
        p.active = active; p.nops = prog.nops;
        std::memcpy(p.ops, prog.ops, sizeof(ChainOp) * prog.nops);
        cudaStream_t st = h->stream;
        if (h->prof.side.active()) {
            h->prof.fork(h->stream);
            if (i == 1) st = h->prof.side.st;
        }
        // executed flops: output tiles that are pure padding are skipped by the kernel when the populated extent ends one 8-wide
        // tile short of the block side (136 of 144, 120 of 128), and the contraction always stops at kvalid
        const double side = (p.kvalid == p.D - 8) ? (double)p.kvalid : (double)p.D;
        ProfScope ps(&h->prof, st, CAT_GEMM, 2.0 * side * side * p.kvalid * batch * prog.products);
        CUDA_TRY(chain_launch(p, batch, st));
    }
    return 0;
}

// layer superoperators Y_l (and tangents Yd_l) for all instances
static int build_layers(otn_problem* h, const double* x, const double* eta, int batch, const int* active, bool primal) {
    const int items = batch * h->m;
    const bool tangent = eta != nullptr;
    back_sync(h);
    ProfScope ps(&h->prof, h->stream, CAT_ASSEMBLE, 2.0 * h->K * (double)h->dim * h->dim * h->dim * h->dim * items * ((primal ? 1 : 0) + (tangent ? 2 : 0)),
                 h->model == OTN_MODEL_LOCAL ? 2 : 1);
    const size_t smem = (size_t)h->np * 8 * (tangent ? 2 : 1);
    double* Yout = (h->model == OTN_MODEL_FULL) ? h->Y : h->Yl;
    double* Ydout = (h->model == OTN_MODEL_FULL) ? h->Yd : h->Yld;
    const long long ostride = (h->model == OTN_MODEL_FULL) ? h->SZ : (long long)h->q * h->q;
    dim3 grid(h->dim, items);
    if (h->cplx) {
        const size_t smc = (size_t)h->np * 8 * (tangent ? 2 : 1);
        const cplx* xc = reinterpret_cast<const cplx*>(x);
        const cplx* ec = reinterpret_cast<const cplx*>(eta);
#define CASM(TG, PR)                                                                                                 \
    do {                                                                                                             \
        OTN_TRY(set_smem(cassemble_emb_kernel<TG, PR>, smc));                                                        \
        cassemble_emb_kernel<TG, PR><<<grid, 256, smc, h->stream>>>(xc, ec, h->Y, h->Yd, h->SZ, h->dim, h->K, active, h->m); \
    } while (0)
        if (tangent && primal) CASM(true, true); else if (tangent) CASM(true, false); else CASM(false, true);
#undef CASM
        CUDA_TRY(cudaGetLastError());
        return 0;
    }
    if (h->structured && h->model == OTN_MODEL_FULL) {
        const int Kp = (h->K + 3) & ~3;
        const size_t sm = ((size_t)h->dim * Kp * h->dim * (tangent ? 2 : 1) + (size_t)8 * 2 * h->dim * (h->dim + 1)) * 8;
#define ASM_SYM(DIMV, TG, PR)                                                                                        \
    do {                                                                                                             \
        OTN_TRY(set_smem(assemble_sym_kernel<DIMV, TG, PR>, sm));                                                    \
        assemble_sym_kernel<DIMV, TG, PR><<<items, 256, sm, h->stream>>>(x, eta, h->Y, h->Yd, h->sym, h->K, active, h->m); \
    } while (0)
        static const bool v1 = getenv("OTN_ASM_V1") != nullptr;      // env: A/B switch for measurements (first version of the kernel)
        if (h->dim == 16 && Kp <= 16 && !v1) {
            const size_t sm2 = AsmSym2Cfg::smem_bytes(Kp, tangent);
#define ASM_SYM2(TG, PR)                                                                                             \
    do {                                                                                                             \
        OTN_TRY(set_smem(assemble_sym2_kernel<TG, PR>, sm2));                                                        \
        assemble_sym2_kernel<TG, PR><<<items, 256, sm2, h->stream>>>(x, eta, h->Y, h->Yd, h->sym, h->K, active, h->m); \
    } while (0)
            if (tangent && primal) ASM_SYM2(true, true); else if (tangent) ASM_SYM2(true, false); else ASM_SYM2(false, true);
#undef ASM_SYM2
        } else if (h->dim == 16) { if (tangent && primal) ASM_SYM(16, true, true); else if (tangent) ASM_SYM(16, true, false); else ASM_SYM(16, false, true); }
        else { if (tangent && primal) ASM_SYM(8, true, true); else if (tangent) ASM_SYM(8, true, false); else ASM_SYM(8, false, true); }
#undef ASM_SYM
    } else if (h->dim == 16 || h->dim == 8) {
        const int Kp = (h->K + 3) & ~3;
        const size_t sm = (size_t)h->dim * Kp * (h->dim + 4) * 8 * (tangent ? 2 : 1);
#define ASM_MMA(DIMV, TG, PR)                                                                                              \
    do {                                                                                                                   \
        OTN_TRY(set_smem(assemble_mma_kernel<DIMV, TG, PR>, sm));                                                          \
        assemble_mma_kernel<DIMV, TG, PR><<<grid, 256, sm, h->stream>>>(x, eta, Yout, Ydout, ostride, h->K, active, h->m); \
    } while (0)
        if (h->dim == 16) { if (tangent && primal) ASM_MMA(16, true, true); else if (tangent) ASM_MMA(16, true, false); else ASM_MMA(16, false, true); }
        else { if (tangent && primal) ASM_MMA(8, true, true); else if (tangent) ASM_MMA(8, true, false); else ASM_MMA(8, false, true); }
#undef ASM_MMA
    } else if (tangent && primal) {
        OTN_TRY(set_smem(assemble_kernel<true, true>, smem));
        assemble_kernel<true, true><<<grid, ASM_THREADS, smem, h->stream>>>(x, eta, Yout, Ydout, ostride, h->dim, h->K, active, h->m);
    } else if (tangent) {
        OTN_TRY(set_smem(assemble_kernel<true, false>, smem));
        assemble_kernel<true, false><<<grid, ASM_THREADS, smem, h->stream>>>(x, eta, Yout, Ydout, ostride, h->dim, h->K, active, h->m);
    } else {
        OTN_TRY(set_smem(assemble_kernel<false, true>, smem));
        assemble_kernel<false, true><<<grid, ASM_THREADS, smem, h->stream>>>(x, eta, Yout, Ydout, ostride, h->dim, h->K, active, h->m);
    }
    CUDA_TRY(cudaGetLastError());
    if (h->factored) return 0;                 // layers stay local (16 x 16): factored.cuh applies them factor by factor
    if (h->model == OTN_MODEL_LOCAL && h->structured) {
        dim3 g2((h->sym.NS + EMB_ROWS - 1) / EMB_ROWS, items);
        const size_t sm2 = (size_t)h->q * h->q * 8 * 2;
        if (tangent && primal) embed_sym_kernel<true, true><<<g2, EMB_THREADS, sm2, h->stream>>>(h->Yl, h->Yld, h->Y, h->Yd, h->tables(), h->sym, active, h->m);
        else if (tangent) embed_sym_kernel<true, false><<<g2, EMB_THREADS, sm2, h->stream>>>(h->Yl, h->Yld, h->Y, h->Yd, h->tables(), h->sym, active, h->m);
        else embed_sym_kernel<false, true><<<g2, EMB_THREADS, sm2, h->stream>>>(h->Yl, h->Yld, h->Y, h->Yd, h->tables(), h->sym, active, h->m);
        CUDA_TRY(cudaGetLastError());
    } else if (h->model == OTN_MODEL_LOCAL) {
        dim3 g2((h->D + EMB_ROWS - 1) / EMB_ROWS, items);
        const size_t sm2 = (size_t)h->q * h->q * 8 * 2;
        if (tangent && primal) embed_kernel<true, true><<<g2, EMB_THREADS, sm2, h->stream>>>(h->Yl, h->Yld, h->Y, h->Yd, h->tables(), active, h->m);
        else if (tangent) embed_kernel<true, false><<<g2, EMB_THREADS, sm2, h->stream>>>(h->Yl, h->Yld, h->Y, h->Yd, h->tables(), active, h->m);
        else embed_kernel<false, true><<<g2, EMB_THREADS, sm2, h->stream>>>(h->Yl, h->Yld, h->Y, h->Yd, h->tables(), active, h->m);
        CUDA_TRY(cudaGetLastError());
    }
    return 0;
}

__global__ void resid_kernel(const double* __restrict__ Y, long long sY, const double* __restrict__ T, long long sT, const int* tindex,
                             double* __restrict__ R, long long sR, const double* __restrict__ Rdot, long long sRd, double* partial,
                             long long SZ, const int* active, double scale) {
    // m == 1 only (no GEMM to fuse into): R = Y - T and partial = ||R||^2, or partial = <R, Rdot> when Rdot != nullptr
    const int b = blockIdx.x;
    if (active != nullptr && active[b] == 0) return;
    __shared__ double red[8];
    double s = 0.0;
    if (Rdot == nullptr) {
        const double* t = T + (long long)(tindex ? tindex[b] : 0) * sT;
        for (long long i = threadIdx.x; i < SZ; i += blockDim.x) {
            const double v = Y[b * sY + i] - t[i];
            R[b * sR + i] = v;
            s = fma(v, v, s);
        }
    } else {
        for (long long i = threadIdx.x; i < SZ; i += blockDim.x) s = fma(R[b * sR + i], Rdot[b * sRd + i], s);
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < 8; ++w) tot += red[w];
        partial[b] = tot * scale;       // complex handles: the sums run over the whole embedding = twice the complex quantity
    }
}

// forward chain: P_l = Y_l P_{l-1}; the last product writes the residual R = M - Ts into G[m-1] (cost / gradient paths) or the
// model matrix M into P[m-1] (`model_only`, for otn_model)
static int forward_chain(otn_problem* h, int batch, const int* active, bool model_only, int* tiles_used) {
    const long long SZ = h->SZ, S = (long long)h->m * SZ;
    const int m = h->m;
    *tiles_used = h->tiles;
    if (m == 1) {
        if (model_only) return 0;   // M = Y_0
        ProfScope ps(&h->prof, h->stream, CAT_OTHER, 0.0);
        resid_kernel<<<batch, 256, 0, h->stream>>>(h->Y, S, h->T_re, SZ, h->tindex, h->G, S, nullptr, 0, h->part_f, SZ, active, h->cplx ? 0.5 : 1.0);
        *tiles_used = 1;
        CUDA_TRY(cudaGetLastError());
        return 0;
    }
    if (h->fused) return hchain(h, batch, active, model_only ? CHAIN_FWD_MODEL : CHAIN_FWD);
    for (int l = 1; l < m; ++l) {
        GemmParams p = gp(h, batch, active);
        p.A1 = layer_ptr(h->Y, l, SZ); p.sA1 = S;
        p.B1 = (l == 1) ? h->Y : layer_ptr(h->P, l - 1, SZ); p.sB1 = S;
        if (l < m - 1 || model_only) { p.C = layer_ptr(h->P, l, SZ); p.sC = S; }
        else {
            p.C = layer_ptr(h->G, m - 1, SZ); p.sC = S;
            p.epi = EPI_RESID; p.E = h->T_re; p.sE = SZ; p.e_index = h->tindex; p.partial = h->part_f; p.partial_stride = h->tiles;
        }
        CUDA_TRY(hgemm(h, p, false, false));
    }
    return 0;
}

// copy the model matrix of every instance (P[m-1], or Y[0] when m == 1) to a dense [batch][D][D] array
static int export_model(otn_problem* h, int batch, double* Mout) {
    const long long SZ = h->SZ, S = (long long)h->m * SZ;
    const double* src = (h->m == 1) ? h->Y : layer_ptr(h->P, h->m - 1, SZ);
    ProfScope ps(&h->prof, h->stream, CAT_OTHER, 0.0);
    if (h->structured) {
        unblock_kernel<<<148 * 8, 256, 0, h->stream>>>(src, S, Mout, h->sym, batch);
        CUDA_TRY(cudaGetLastError());
    } else {
        const size_t DD = (size_t)h->D * h->D * 8;
        CUDA_TRY(cudaMemcpy2DAsync(Mout, DD, src, (size_t)S * 8, DD, batch, cudaMemcpyDeviceToDevice, h->stream));
    }
    return 0;
}

// reverse chain on the unscaled residual: G_{l-1} = Y_l^T G_l ; W_l = G_l P_{l-1}^T ; zraw_l = back_assemble(W_l)
static int back_layer(otn_problem* h, int batch, const int* active, int l, const double* x, bool tangent, const double* eta,
                      bool primal_with_eta = false) {
    // pulls the cotangent of layer l (W_l, and for the tangent pass Wd) back to the isometry
    const bool ovl = back_overlap(h);
    const cudaStream_t bst = ovl ? back_begin(h) : h->stream;
    // kernels launched below: the back-embedding of the 2-site model, then one pull-back, or two when the tangent pass pulls Wd
    // back against x and W against eta separately (dense streaming kernel; block kernel outside the joint x / eta pass)
    const bool two_pullbacks = tangent && ((h->structured && h->model == OTN_MODEL_FULL) ? !primal_with_eta : (h->dim == 16 || h->dim == 8));
    ProfScope ps(&h->prof, bst, CAT_BACK, 4.0 * h->K * (double)h->dim * h->dim * h->dim * h->dim * batch * (tangent ? 2 : 1),
                 ((h->model == OTN_MODEL_LOCAL && !h->factored) ? (h->be3 ? 2 : 1) : 0) + (two_pullbacks ? 2 : 1), ovl);
    if (ovl && tangent) h->back_wd_pending = true;
    const long long SZ = h->SZ, S = (long long)h->m * SZ;
    const double* Wfull = (l == 0) ? h->G : layer_ptr(h->W, l, SZ);                  // W_0 = G_0
    const double* Wdfull = nullptr; long long sWd = 0;
    if (tangent) {
        if (l == 0) { Wdfull = h->Gd + (long long)h->tr.gd_final * SZ; sWd = 2 * SZ; }
        else { Wdfull = h->fused ? h->Wd + (long long)l * SZ : h->Wd; sWd = h->wd_stride; }
    }
    if (h->cplx) {
        // zraw = A(W) x ;  tangent: zdraw = A(Wd) x + A(W) eta   (A = adjoint of the complex assembly, complex.cuh)
        const size_t smc = ((size_t)h->np / 2 + 2 * (size_t)h->dim * h->dim) * 16;
        OTN_TRY(set_smem(cback_assemble_kernel, smc));
        dim3 gc(h->dim, batch);
        const cplx* xc = reinterpret_cast<const cplx*>(x);
        if (!tangent) {
            cback_assemble_kernel<<<gc, 256, smc, bst>>>(Wfull, S, xc, reinterpret_cast<cplx*>(h->zraw), 0, h->dim, h->K, active, h->m, l);
            if (primal_with_eta)
                cback_assemble_kernel<<<gc, 256, smc, bst>>>(Wfull, S, reinterpret_cast<const cplx*>(eta), reinterpret_cast<cplx*>(h->zdraw), 0, h->dim, h->K, active, h->m, l);
        } else {
            cback_assemble_kernel<<<gc, 256, smc, bst>>>(Wdfull, sWd, xc, reinterpret_cast<cplx*>(h->zdraw), primal_with_eta ? 1 : 0, h->dim, h->K, active, h->m, l);
            if (!primal_with_eta)
                cback_assemble_kernel<<<gc, 256, smc, bst>>>(Wfull, S, reinterpret_cast<const cplx*>(eta), reinterpret_cast<cplx*>(h->zdraw), 1, h->dim, h->K, active, h->m, l);
        }
        CUDA_TRY(cudaGetLastError());
        return 0;
    }
    const double* Wsup = Wfull; long long sW = S; const double* Wdsup = Wdfull; long long sWds = sWd; int nslices = 1;
    if (h->model == OTN_MODEL_LOCAL) {
        const int q2 = h->q * h->q;
        dim3 g((h->nf * q2 + 7) / 8, batch);
        const size_t sm = (size_t)q2 * 8 * 2;
        if (h->factored) {
            // Wl / Wld already hold the per-factor local cotangents (fact_reduce_kernel)
        } else if (h->structured && h->nf == 2 && h->q == 16) {
            const size_t sm2 = (size_t)(16 * 256 * (tangent ? 2 : 1) + 16 * 16 * 17) * 8;
            OTN_TRY(set_smem(back_embed2_sym_kernel<false>, sm2));
            OTN_TRY(set_smem(back_embed2_sym_kernel<true>, sm2));
            if (!tangent) back_embed2_sym_kernel<false><<<batch, 256, sm2, bst>>>(Wfull, nullptr, S, 0, h->Yl, h->Yld, h->Wl, h->Wld, h->sym, h->symtab, h->t2f, active, h->m, l);
            else back_embed2_sym_kernel<true><<<batch, 256, sm2, bst>>>(Wfull, Wdfull, S, sWd, h->Yl, h->Yld, h->Wl, h->Wld, h->sym, h->symtab, h->t2f, active, h->m, l);
        } else if (h->structured) {
            if (!tangent) back_embed_sym_kernel<false><<<g, BEM_THREADS, sm, bst>>>(Wfull, nullptr, S, 0, h->Yl, h->Yld, h->Wl, h->Wld, h->tables(), h->sym, h->symtab, active, h->m, l);
            else back_embed_sym_kernel<true><<<g, BEM_THREADS, sm, bst>>>(Wfull, Wdfull, S, sWd, h->Yl, h->Yld, h->Wl, h->Wld, h->tables(), h->sym, h->symtab, active, h->m, l);
        } else if (h->be3 != nullptr) {
            // three factors of side 16 (N = 6): factor-by-factor contraction with whole-row reads of W (assemble.cuh)
            // Row-sharded handle: the 256 stage-1 CTAs per pass are dealt to the ranks, each stores its rows of A / B into every
            // peer; barriers around stage 1 as around a chain product (every output element has one producer: still bitwise equal).
            Be3Peers peers{};
            int x0 = 0, nx = 256;
            if (h->rs.world > 1) {
                nx = 256 / h->rs.world; x0 = h->rs.rank * nx;
                peers.n = h->rs.world - 1;
                for (int q = 0; q < peers.n; ++q) {
                    peers.p[q] = rs_peer(h, q, h->be3);
                    if (peers.p[q] == nullptr) return fail("row-sharded back-embedding: scratch not mapped");
                }
                CUDA_TRY(rs_barrier(h, h->stream));
            }
            dim3 g1(nx, 2, batch), g2(3, batch);
            if (!tangent) {
                OTN_TRY(set_smem(back_embed3_stage1_kernel<false>, BE3_SMEM));
                back_embed3_stage1_kernel<false><<<g1, BE3_THREADS, BE3_SMEM, bst>>>(Wfull, nullptr, S, 0, h->Yl, h->Yld, h->be3, peers, x0, h->tables(), active, h->m, l);
            } else {
                OTN_TRY(set_smem(back_embed3_stage1_kernel<true>, BE3_SMEM));
                back_embed3_stage1_kernel<true><<<g1, BE3_THREADS, BE3_SMEM, bst>>>(Wfull, Wdfull, S, sWd, h->Yl, h->Yld, h->be3, peers, x0, h->tables(), active, h->m, l);
            }
            if (h->rs.world > 1) CUDA_TRY(rs_barrier(h, h->stream));
            if (!tangent) back_embed3_stage2_kernel<false><<<g2, 256, 0, bst>>>(h->be3, h->Yl, h->Yld, h->Wl, h->Wld, active, h->m, l);
            else back_embed3_stage2_kernel<true><<<g2, 256, 0, bst>>>(h->be3, h->Yl, h->Yld, h->Wl, h->Wld, active, h->m, l);
        } else if (!tangent) back_embed_kernel<false><<<g, BEM_THREADS, sm, bst>>>(Wfull, nullptr, S, 0, h->Yl, h->Yld, h->Wl, h->Wld, h->tables(), active, h->m, l);
        else back_embed_kernel<true><<<g, BEM_THREADS, sm, bst>>>(Wfull, Wdfull, S, sWd, h->Yl, h->Yld, h->Wl, h->Wld, h->tables(), active, h->m, l);
        CUDA_TRY(cudaGetLastError());
        const long long ls = (long long)h->nf * q2;
        Wsup = h->Wl + (long long)l * ls; sW = (long long)h->m * ls;
        Wdsup = h->Wld + (long long)l * ls; sWds = sW;
        nslices = h->nf;
    }
    const int dim = h->dim;
    dim3 grid(dim, batch);
    if (h->structured && h->model == OTN_MODEL_FULL) {
#define BAS_SYM(DIMV, ASV, KPV)                                                                                                  \
    do {                                                                                                                         \
        const size_t sm1 = BasSym2Cfg<DIMV, ASV, 1, KPV>::SMEM_BYTES, sm2 = BasSym2Cfg<DIMV, ASV, 2, KPV>::SMEM_BYTES;           \
        OTN_TRY(set_smem(back_assemble_sym2_kernel<DIMV, ASV, 1, KPV>, sm1));                                                    \
        OTN_TRY(set_smem(back_assemble_sym2_kernel<DIMV, ASV, 2, KPV>, sm2));                                                    \
        if (!tangent && primal_with_eta) {      /* zraw = A(W) x  and  zdraw = A(W) eta  in one pass over W */                   \
            back_assemble_sym2_kernel<DIMV, ASV, 2, KPV><<<batch * ASV, 256, sm2, bst>>>(Wsup, sW, x, eta, h->zraw, h->zdraw, 0, 0, h->sym, h->K, active, h->m, l); \
        } else if (!tangent) {                                                                                                   \
            back_assemble_sym2_kernel<DIMV, ASV, 1, KPV><<<batch * ASV, 256, sm1, bst>>>(Wsup, sW, x, nullptr, h->zraw, nullptr, 0, 0, h->sym, h->K, active, h->m, l); \
        } else if (primal_with_eta) {           /* the W eta part is already in zdraw: only += A(Wd) x */                        \
            back_assemble_sym2_kernel<DIMV, ASV, 1, KPV><<<batch * ASV, 256, sm1, bst>>>(Wdsup, sWds, x, nullptr, h->zdraw, nullptr, 1, 0, h->sym, h->K, active, h->m, l); \
        } else {                                                                                                                 \
            back_assemble_sym2_kernel<DIMV, ASV, 1, KPV><<<batch * ASV, 256, sm1, bst>>>(Wdsup, sWds, x, nullptr, h->zdraw, nullptr, 0, 0, h->sym, h->K, active, h->m, l); \
            back_assemble_sym2_kernel<DIMV, ASV, 1, KPV><<<batch * ASV, 256, sm1, bst>>>(Wsup, sW, eta, nullptr, h->zdraw, nullptr, 1, 0, h->sym, h->K, active, h->m, l); \
        }                                                                                                                        \
    } while (0)
#define BAS_SYM_K(DIMV, ASV)                                                                        \
    do {                                                                                            \
        const int kp = (h->K + 7) & ~7;                                                             \
        if (kp == 8) BAS_SYM(DIMV, ASV, 8); else if (kp == 16) BAS_SYM(DIMV, ASV, 16);              \
        else if (kp == 24) BAS_SYM(DIMV, ASV, 24); else BAS_SYM(DIMV, ASV, 32);                     \
    } while (0)
        if (dim == 16) BAS_SYM_K(16, 2); else BAS_SYM_K(8, 1);
#undef BAS_SYM_K
#undef BAS_SYM
    } else if (dim == 16 || dim == 8) {
        // streaming kernel, one CTA per instance; tangent = Wd*x then += W*eta
#define BAS_STREAM(DIMV)                                                                                                        \
    do {                                                                                                                        \
        const size_t sm = BasCfg<DIMV>::smem_bytes(h->K);                                                                       \
        OTN_TRY(set_smem(back_assemble_stream_kernel<DIMV>, sm));                                                               \
        if (!tangent) {                                                                                                         \
            back_assemble_stream_kernel<DIMV><<<batch, 256, sm, bst>>>(Wsup, sW, x, h->zraw, h->K, 0, active, h->m, l);   \
        } else {                                                                                                                \
            back_assemble_stream_kernel<DIMV><<<batch, 256, sm, bst>>>(Wdsup, sWds, x, h->zdraw, h->K, 0, active, h->m, l); \
            back_assemble_stream_kernel<DIMV><<<batch, 256, sm, bst>>>(Wsup, sW, eta, h->zdraw, h->K, 1, active, h->m, l);  \
        }                                                                                                                       \
    } while (0)
        if (dim == 16) BAS_STREAM(16); else BAS_STREAM(8);
#undef BAS_STREAM
    } else {
        const size_t wsz = (size_t)dim * dim * (dim + 1);
        if (!tangent) {
            const size_t sm = (wsz + h->np) * 8;
            OTN_TRY(set_smem(back_assemble_kernel<false>, sm));
            back_assemble_kernel<false><<<grid, BAS_THREADS, sm, bst>>>(Wsup, nullptr, sW, 0, nslices, x, nullptr, h->zraw, nullptr, dim, h->K, active, h->m, l);
        } else {
            const size_t sm = (2 * wsz + 2 * h->np) * 8;
            OTN_TRY(set_smem(back_assemble_kernel<true>, sm));
            back_assemble_kernel<true><<<grid, BAS_THREADS, sm, bst>>>(Wsup, Wdsup, sW, sWds, nslices, x, eta, nullptr, h->zdraw, dim, h->K, active, h->m, l);
        }
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// fused-chain path, full-sites model in blocks: zraw = A(W_l) x (`grad`) and / or zdraw = A(W_l) eta + A(Wd_l) x (`tan`) for every
// layer in one launch
static bool back_all_ok(const otn_problem* h) {
    return h->fused && h->structured && h->model == OTN_MODEL_FULL && h->dim == 16 && h->K <= 32 && h->m <= 16 && h->m >= 2 &&
           getenv("OTN_BACK_PER_LAYER") == nullptr;               // env: A/B switch for measurements
}
static int back_all_layers(otn_problem* h, int batch, const int* active, const double* x, const double* eta, bool grad, bool tan) {
    const long long SZ = h->SZ, S = (long long)h->m * SZ;
    ProfScope ps(&h->prof, h->stream, CAT_BACK, 4.0 * h->K * (double)h->dim * h->dim * h->dim * h->dim * batch * h->m * ((grad ? 1 : 0) + (tan ? 2 : 0)));
    BackAll P{};
    for (int l = 0; l < h->m; ++l) {
        if (l == 0) { P.W[l] = h->G; P.sW[l] = S; P.Wd[l] = h->Gd + (long long)h->tr.gd_final * SZ; P.sWd[l] = 2 * SZ; }
        else { P.W[l] = layer_ptr(h->W, l, SZ); P.sW[l] = S; P.Wd[l] = h->Wd + (long long)l * SZ; P.sWd[l] = h->wd_stride; }
    }
    const int kp = (h->K + 7) & ~7;
    dim3 grid(batch * 2, h->m);
#define BAS3(KPV, GR, TN)                                                                                                  \
    do {                                                                                                                   \
        const size_t sm = BasSym3Cfg<16, 2, KPV>::SMEM_BYTES;                                                              \
        OTN_TRY(set_smem(back_assemble_sym3_kernel<16, 2, KPV, GR, TN>, sm));                                              \
        back_assemble_sym3_kernel<16, 2, KPV, GR, TN><<<grid, 256, sm, h->stream>>>(P, x, eta, h->zraw, h->zdraw, h->sym, h->K, active, h->m); \
    } while (0)
#define BAS3_K(GR, TN)                                                                                                     \
    do { if (kp == 8) BAS3(8, GR, TN); else if (kp == 16) BAS3(16, GR, TN); else if (kp == 24) BAS3(24, GR, TN); else BAS3(32, GR, TN); } while (0)
#define BAS4(KPV, GR, TN)                                                                                                  \
    do {                                                                                                                   \
        const size_t sm = BasSym4Cfg<KPV>::SMEM_BYTES;                                                                     \
        OTN_TRY(set_smem(back_assemble_sym4_kernel<KPV, GR, TN>, sm));                                                     \
        back_assemble_sym4_kernel<KPV, GR, TN><<<grid, 256, sm, h->stream>>>(P, x, eta, h->zraw, h->zdraw, h->sym, h->K, active, h->m); \
    } while (0)
#define BAS4_K(GR, TN)                                                                                                     \
    do { if (kp == 8) BAS4(8, GR, TN); else if (kp == 16) BAS4(16, GR, TN); else if (kp == 24) BAS4(24, GR, TN); else BAS4(32, GR, TN); } while (0)
    static const bool v3 = getenv("OTN_BACK_V3") != nullptr;      // env: A/B switch for measurements (cp.async ring + CTA barrier version)
    if (!v3) { if (grad && tan) BAS4_K(true, true); else if (tan) BAS4_K(false, true); else BAS4_K(true, false); }
    else if (grad && tan) BAS3_K(true, true); else if (tan) BAS3_K(false, true); else BAS3_K(true, false);
#undef BAS4_K
#undef BAS4
#undef BAS3_K
#undef BAS3
    CUDA_TRY(cudaGetLastError());
    return 0;
}

static int reverse_chain(otn_problem* h, int batch, const int* active, const double* x, const double* eta_joint = nullptr) {
    const long long SZ = h->SZ, S = (long long)h->m * SZ;
    if (h->fused && h->m >= 2) {
        OTN_TRY(hchain(h, batch, active, CHAIN_REV));
        if (back_all_ok(h) && eta_joint == nullptr) return back_all_layers(h, batch, active, x, nullptr, true, false);
        for (int l = h->m - 1; l >= 0; --l) OTN_TRY(back_layer(h, batch, active, l, x, false, eta_joint, eta_joint != nullptr));
        return 0;
    }
    for (int l = h->m - 1; l >= 1; --l) {
        GemmParams p = gp(h, batch, active);                       // W_l = G_l P_{l-1}^T
        p.A1 = layer_ptr(h->G, l, SZ); p.sA1 = S;
        p.B1 = (l == 1) ? h->Y : layer_ptr(h->P, l - 1, SZ); p.sB1 = S;
        p.C = layer_ptr(h->W, l, SZ); p.sC = S;
        GemmParams r = gp(h, batch, active);                       // G_{l-1} = Y_l^T G_l
        r.A1 = layer_ptr(h->Y, l, SZ); r.sA1 = S;
        r.B1 = layer_ptr(h->G, l, SZ); r.sB1 = S;
        r.C = layer_ptr(h->G, l - 1, SZ); r.sC = S;
        const bool ovl = back_overlap(h);      // W_l goes to the back stream at once; the chain carries on underneath
        if (!ovl) {                            // (row-sharded handles: the two independent products as a concurrent pair)
            CUDA_TRY(hgemm_pair(h, p, false, true, r, true, false));
            OTN_TRY(back_layer(h, batch, active, l, x, false, eta_joint, eta_joint != nullptr));
            continue;
        }
        CUDA_TRY(hgemm(h, p, false, true));
        OTN_TRY(back_layer(h, batch, active, l, x, false, eta_joint, eta_joint != nullptr));
        CUDA_TRY(hgemm(h, r, true, false));
    }
    OTN_TRY(back_layer(h, batch, active, 0, x, false, eta_joint, eta_joint != nullptr));
    return 0;
}

// tangent pass: needs the primal cache (Y, P, G, W, zraw, part_f) of the same x
static int tangent_pass(otn_problem* h, int batch, const int* active, const double* x, const double* eta, bool layers_built,
                        bool joint = false) {
    const long long SZ = h->SZ, S = (long long)h->m * SZ;
    const int m = h->m;
    if (!layers_built) OTN_TRY(build_layers(h, x, eta, batch, active, false));
    if (h->fused && m >= 2) {
        OTN_TRY(hchain(h, batch, active, CHAIN_TANGENT));
        if (back_all_ok(h) && !joint) return back_all_layers(h, batch, active, x, eta, false, true);
        for (int l = m - 1; l >= 0; --l) OTN_TRY(back_layer(h, batch, active, l, x, true, eta, joint));
        return 0;
    }
    // forward tangent: Pd_l = Yd_l P_{l-1} + Y_l Pd_{l-1};  last one is Mdot -> Gd[0], with <R, Mdot> partials
    int cur = 0;
    if (m == 1) {
        CUDA_TRY(cudaMemcpy2DAsync(h->Gd, 2 * SZ * 8, h->Yd, S * 8, SZ * 8, batch, cudaMemcpyDeviceToDevice, h->stream));
        ProfScope ps(&h->prof, h->stream, CAT_OTHER, 0.0);
        resid_kernel<<<batch, 256, 0, h->stream>>>(nullptr, 0, nullptr, 0, nullptr, h->G, S, h->Gd, 2 * SZ, h->part_s, SZ, active, h->cplx ? 0.5 : 1.0);
        CUDA_TRY(cudaGetLastError());
    }
    for (int l = 1; l < m; ++l) {
        GemmParams p = gp(h, batch, active);
        p.A1 = layer_ptr(h->Yd, l, SZ); p.sA1 = S;
        p.B1 = (l == 1) ? h->Y : layer_ptr(h->P, l - 1, SZ); p.sB1 = S;
        p.A2 = layer_ptr(h->Y, l, SZ); p.sA2 = S;
        p.B2 = (l == 1) ? h->Yd : layer_ptr(h->Pd, l - 1, SZ); p.sB2 = S;
        if (l < m - 1) { p.C = layer_ptr(h->Pd, l, SZ); p.sC = S; }
        else {
            p.C = h->Gd; p.sC = 2 * SZ;
            p.epi = EPI_DOT; p.E = layer_ptr(h->G, m - 1, SZ); p.sE = S; p.partial = h->part_s; p.partial_stride = h->tiles;
        }
        CUDA_TRY(hgemm(h, p, false, false));
    }
    // reverse tangent
    for (int l = m - 1; l >= 1; --l) {
        const double* Ghat = h->Gd + (long long)cur * SZ;
        const double* Pprev = (l == 1) ? h->Y : layer_ptr(h->P, l - 1, SZ);
        const double* Pdprev = (l == 1) ? h->Yd : layer_ptr(h->Pd, l - 1, SZ);
        GemmParams p = gp(h, batch, active);                       // What_l = Ghat_l P_{l-1}^T + G_l Pd_{l-1}^T
        p.A1 = Ghat; p.sA1 = 2 * SZ; p.B1 = Pprev; p.sB1 = S;
        p.A2 = layer_ptr(h->G, l, SZ); p.sA2 = S; p.B2 = Pdprev; p.sB2 = S;
        p.C = h->Wd; p.sC = SZ;
        if (h->back_wd_pending) back_sync(h);  // Wd is one buffer: the pull-back of the previous layer's Wd must have read it
        GemmParams r = gp(h, batch, active);                       // Ghat_{l-1} = Yd_l^T G_l + Y_l^T Ghat_l
        r.A1 = layer_ptr(h->Yd, l, SZ); r.sA1 = S; r.B1 = layer_ptr(h->G, l, SZ); r.sB1 = S;
        r.A2 = layer_ptr(h->Y, l, SZ); r.sA2 = S; r.B2 = Ghat; r.sB2 = 2 * SZ;
        r.C = h->Gd + (long long)(1 - cur) * SZ; r.sC = 2 * SZ;
        const bool ovl = back_overlap(h);
        if (!ovl) {
            CUDA_TRY(hgemm_pair(h, p, false, true, r, true, false));
        } else {
            CUDA_TRY(hgemm(h, p, false, true));
            OTN_TRY(back_layer(h, batch, active, l, x, true, eta, joint));
            CUDA_TRY(hgemm(h, r, true, false));
        }
        if (!ovl) OTN_TRY(back_layer(h, batch, active, l, x, true, eta, joint));
        cur = 1 - cur;
    }
    h->tr.gd_final = cur;
    OTN_TRY(back_layer(h, batch, active, 0, x, true, eta, joint));
    return 0;
}

static int riem_launch(otn_problem* h, int batch, const int* active, int tiles, const double* x, const double* eta, bool hvp,
                       double* f, double* rgrad, double* egrad, double* heta) {
    back_sync(h);                              // zraw / zdraw come from the back stream
    ProfScope ps(&h->prof, h->stream, CAT_RIEM, 0.0);
    if (h->cplx) {
        CRiemParams Pc{};
        Pc.x = reinterpret_cast<const cplx*>(x); Pc.zraw = reinterpret_cast<const cplx*>(h->zraw);
        Pc.eta = reinterpret_cast<const cplx*>(eta); Pc.zdraw = reinterpret_cast<const cplx*>(h->zdraw);
        Pc.part_f = h->part_f; Pc.part_s = h->part_s; Pc.tiles = tiles; Pc.f_out = f;
        Pc.rgrad = reinterpret_cast<cplx*>(rgrad); Pc.egrad = reinterpret_cast<cplx*>(egrad); Pc.heta = reinterpret_cast<cplx*>(heta);
        Pc.n = h->n; Pc.p = h->p; Pc.m = h->m; Pc.alpha0 = h->alpha0; Pc.alpha1 = h->alpha1; Pc.active = active;
        const size_t smc = ((hvp ? 4 : 2) * (size_t)h->n * h->p + 8 * (size_t)h->p * h->p) * 16;
        if (!hvp) { OTN_TRY(set_smem(criem_kernel<false>, smc)); criem_kernel<false><<<batch * h->m, ST_THREADS, smc, h->stream>>>(Pc); }
        else { OTN_TRY(set_smem(criem_kernel<true>, smc)); criem_kernel<true><<<batch * h->m, ST_THREADS, smc, h->stream>>>(Pc); }
        CUDA_TRY(cudaGetLastError());
        return 0;
    }
    RiemParams P{};
    P.x = x; P.zraw = h->zraw; P.eta = eta; P.zdraw = h->zdraw; P.part_f = h->part_f; P.part_s = h->part_s; P.im2 = h->im2;
    P.tindex = h->tindex; P.tiles = tiles; P.f_out = f; P.rgrad = rgrad; P.egrad = egrad; P.heta = heta;
    P.n = h->n; P.p = h->p; P.m = h->m; P.alpha0 = h->alpha0; P.alpha1 = h->alpha1; P.active = active;
    const size_t pp = (size_t)h->p * h->p;
    const size_t sm_mma = ((hvp ? 4 : 2) * (size_t)h->n * (h->p + 4) + 8 * h->p * (h->p + 4)) * 8;
    // p = 16 (full-sites model at N = 4): streaming kernel, nothing n x p staged in shared memory (4 CTAs per SM, any Kraus rank)
    const bool staged_only = getenv("OTN_RIEM_STAGED") != nullptr;             // env: A/B switch for measurements and tests
    if (h->p == 16 && h->n % 8 == 0 && !staged_only) {
        if (!hvp) {
            OTN_TRY(set_smem(riem_stream_kernel<false>, RiemStreamCfg<false>::SMEM_BYTES));
            riem_stream_kernel<false><<<batch * h->m, RS_THREADS, RiemStreamCfg<false>::SMEM_BYTES, h->stream>>>(P);
        } else {
            OTN_TRY(set_smem(riem_stream_kernel<true>, RiemStreamCfg<true>::SMEM_BYTES));
            riem_stream_kernel<true><<<batch * h->m, RS_THREADS, RiemStreamCfg<true>::SMEM_BYTES, h->stream>>>(P);
        }
    }
    // large Kraus ranks (K > 20 at p = 16): the padded DMMA layout no longer fits 227 KB; the unpadded CUDA-core kernel does up to K = 24
    else if ((h->p == 16 || h->p == 8) && h->n % 8 == 0 && sm_mma <= 227 * 1024) {
        const size_t sm = sm_mma;
#define RIEM_MMA(PV)                                                                               \
    do {                                                                                           \
        if (!hvp) { OTN_TRY(set_smem(riem_mma_kernel<PV, false>, sm)); riem_mma_kernel<PV, false><<<batch * h->m, ST_THREADS, sm, h->stream>>>(P); } \
        else { OTN_TRY(set_smem(riem_mma_kernel<PV, true>, sm)); riem_mma_kernel<PV, true><<<batch * h->m, ST_THREADS, sm, h->stream>>>(P); }      \
    } while (0)
        if (h->p == 16) RIEM_MMA(16); else RIEM_MMA(8);
#undef RIEM_MMA
    } else if (!hvp) {
        const size_t sm = (2 * (size_t)h->np + 8 * pp) * 8;
        OTN_TRY(set_smem(riem_kernel<false>, sm));
        riem_kernel<false><<<batch * h->m, ST_THREADS, sm, h->stream>>>(P);
    } else {
        const size_t sm = (4 * (size_t)h->np + 8 * pp) * 8;
        OTN_TRY(set_smem(riem_kernel<true>, sm));
        riem_kernel<true><<<batch * h->m, ST_THREADS, sm, h->stream>>>(P);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// factored evaluation: pull the local cotangents Wl (and Wld) of ALL layers back to the isometries in one launch
static int fact_back_layers(otn_problem* h, int batch, const int* active, const double* x, bool tangent, const double* eta) {
    ProfScope ps(&h->prof, h->stream, CAT_BACK, 4.0 * h->K * (double)h->dim * h->dim * h->dim * h->dim * batch * h->m * (tangent ? 2 : 1));
    const int dim = h->dim;
    const long long ls = (long long)h->nf * h->q * h->q, sW = (long long)h->m * ls;
    const size_t wsz = (size_t)dim * dim * (dim + 1);
    dim3 grid(dim, batch, h->m);
    if (!tangent) {
        const size_t sm = (wsz + h->np) * 8;
        OTN_TRY(set_smem(back_assemble_kernel<false>, sm));
        back_assemble_kernel<false><<<grid, BAS_THREADS, sm, h->stream>>>(h->Wl, nullptr, sW, 0, h->nf, x, nullptr, h->zraw, nullptr, dim, h->K, active, h->m, -1);
    } else {
        const size_t sm = (2 * wsz + 2 * h->np) * 8;
        OTN_TRY(set_smem(back_assemble_kernel<true>, sm));
        back_assemble_kernel<true><<<grid, BAS_THREADS, sm, h->stream>>>(h->Wl, h->Wld, sW, sW, h->nf, x, eta, nullptr, h->zdraw, dim, h->K, active, h->m, -1);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// factored evaluation (factored.cuh): one launch does forward / residual / reverse / tangent for every column tile
static int fact_eval(otn_problem* h, int batch, const int* active, int mode, double* Mout) {
    FactParams p{};
    p.Yl = h->Yl; p.Yld = h->Yld; p.Tt = h->f_Tt; p.tindex = h->tindex; p.active = active; p.jrow = h->f_jrow; p.restrow = h->f_restrow;
    p.colidx = h->f_colidx; p.colw = h->f_colw; p.Pst = h->f_Pst; p.Wp = h->f_Wp; p.part_f = h->part_f; p.part_s = h->part_s;
    p.part_stride = h->f_tiles; p.Mout = Mout; p.swaprow = h->f_swaprow; p.m = h->m; p.tiles = h->f_tiles;
    const double cols = 0.5 * std::sqrt((double)h->D) * (std::sqrt((double)h->D) + 1.0);
    // applications + contractions of 2*16*16 flops per 16-vector: forward nf (x3 dual), reverse nf(nf-1)/2 + nf + nf (x3 dual)
    const double units = (mode == FACT_COST || mode == FACT_MODEL) ? h->nf
                         : (h->nf + 0.5 * h->nf * (h->nf - 1) + 2.0 * h->nf) * (mode == FACT_GRAD ? 1.0 : 3.0) - (mode == FACT_HVPONLY ? h->nf : 0.0);
    ProfScope ps(&h->prof, h->stream, CAT_GEMM, units * h->m * 2.0 * 16.0 * h->D * cols * batch);
    dim3 grid(h->f_tiles, batch);
#define FACT_LAUNCH(NFV, NCV, MODEV)                                                                     \
    do {                                                                                                 \
        const size_t sm = FactCfg<NFV, NCV, MODEV>::SMEM_BYTES;                                          \
        OTN_TRY(set_smem(fact_kernel<NFV, NCV, MODEV>, sm));                                             \
        fact_kernel<NFV, NCV, MODEV><<<grid, FactThreads<NFV, NCV>::THREADS, sm, h->stream>>>(p);       \
    } while (0)
#define FACT_MODE(NFV, NCV)                                                                              \
    do {                                                                                                 \
        if (mode == FACT_COST) FACT_LAUNCH(NFV, NCV, FACT_COST);                                         \
        else if (mode == FACT_GRAD) FACT_LAUNCH(NFV, NCV, FACT_GRAD);                                    \
        else if (mode == FACT_HVP) FACT_LAUNCH(NFV, NCV, FACT_HVP);                                      \
        else if (mode == FACT_HVPONLY) FACT_LAUNCH(NFV, NCV, FACT_HVPONLY);                              \
        else FACT_LAUNCH(NFV, NCV, FACT_MODEL);                                                          \
    } while (0)
    if (h->nf == 2 && h->f_ncols == 8) FACT_MODE(2, 8); else if (h->nf == 2) FACT_MODE(2, 2); else FACT_MODE(3, 1);
#undef FACT_MODE
#undef FACT_LAUNCH
    CUDA_TRY(cudaGetLastError());
    if (mode == FACT_GRAD || mode == FACT_HVP || mode == FACT_HVPONLY) {
        dim3 g2(h->m * h->nf, batch);
        const bool primal = mode != FACT_HVPONLY, dual = mode != FACT_GRAD;
        if (h->f_nsplit > 1) {
            dim3 g1(h->m * h->nf, batch, h->f_nsplit);
            fact_partial_reduce_kernel<<<g1, 256, 0, h->stream>>>(h->f_Wp, h->f_Wp2, h->f_tiles, 32, h->m * h->nf, primal, dual, active);
            fact_reduce_kernel<<<g2, 256, 0, h->stream>>>(h->f_Wp2, h->Wl, h->Wld, h->f_nsplit, h->m, h->nf, primal, dual, active);
        } else {
            fact_reduce_kernel<<<g2, 256, 0, h->stream>>>(h->f_Wp, h->Wl, h->Wld, h->f_tiles, h->m, h->nf, primal, dual, active);
        }
        CUDA_TRY(cudaGetLastError());
    }
    return 0;
}

// device-level building blocks shared with the trust-region driver
int otn_dev_cost(otn_problem* h, const double* x, int batch, const int* active, double* f) {
    int tiles;
    OTN_TRY(build_layers(h, x, nullptr, batch, active, true));
    if (h->factored) { OTN_TRY(fact_eval(h, batch, active, FACT_COST, nullptr)); tiles = h->f_tiles; }
    else OTN_TRY(forward_chain(h, batch, active, false, &tiles));
    ProfScope ps(&h->prof, h->stream, CAT_OTHER, 0.0);
    cost_from_partials_kernel<<<(batch + 127) / 128, 128, 0, h->stream>>>(h->part_f, tiles, h->im2, h->tindex, f, batch, active,
                                                                          1.0);   // (complex: the partials cover the upper half of emb(R) = ||R||^2)
    CUDA_TRY(cudaGetLastError());
    h->cached_batch = 0;
    return 0;
}
int otn_dev_cost_grad(otn_problem* h, const double* x, int batch, const int* active, double* f, double* rgrad, double* egrad) {
    int tiles;
    OTN_TRY(build_layers(h, x, nullptr, batch, active, true));
    if (h->factored) {
        OTN_TRY(fact_eval(h, batch, active, FACT_GRAD, nullptr));
        tiles = h->f_tiles;
        OTN_TRY(fact_back_layers(h, batch, active, x, false, nullptr));
    } else {
        OTN_TRY(forward_chain(h, batch, active, false, &tiles));
        OTN_TRY(reverse_chain(h, batch, active, x));
    }
    OTN_TRY(riem_launch(h, batch, active, tiles, x, nullptr, false, f, rgrad, egrad, nullptr));
    h->tr.tiles = tiles;
    return 0;
}
// cost + gradient when the forward state (layers Y, products P, residual, cost partials) of the instances with stale[b] == 0 is
// already in the work matrices (left there by otn_dev_cost at the same x): forward pass for the stale instances only, reverse
// pass and Riemannian epilogue for all.  Chain (non-factored) evaluations only.
int otn_dev_cost_grad_cached_forward(otn_problem* h, const double* x, int batch, const int* stale, double* f, double* rgrad) {
    int tiles;
    OTN_TRY(build_layers(h, x, nullptr, batch, stale, true));
    OTN_TRY(forward_chain(h, batch, stale, false, &tiles));
    OTN_TRY(reverse_chain(h, batch, nullptr, x));
    OTN_TRY(riem_launch(h, batch, nullptr, tiles, x, nullptr, false, f, rgrad, nullptr, nullptr));
    h->tr.tiles = tiles;
    return 0;
}
int otn_dev_hvp(otn_problem* h, const double* x, const double* eta, int batch, const int* active, double* heta) {
    if (h->factored) {          // primal recomputed alongside the tangent (dual numbers); zraw is the cached primal pull-back
        OTN_TRY(build_layers(h, x, eta, batch, active, true));
        OTN_TRY(fact_eval(h, batch, active, FACT_HVPONLY, nullptr));   // Wl (primal cotangents) stays as cached by cost_grad
        OTN_TRY(fact_back_layers(h, batch, active, x, true, eta));
    } else {
        OTN_TRY(tangent_pass(h, batch, active, x, eta, false));
    }
    OTN_TRY(riem_launch(h, batch, active, h->tr.tiles, x, eta, true, nullptr, nullptr, nullptr, heta));
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// public model / cost entry points
// ---------------------------------------------------------------------------------------------------------------------
static size_t xbytes(const otn_problem* h, int batch) { return (size_t)batch * h->m * h->np * 8; }

// keep a device copy of x for the cached-HVP path when the caller's x is a host array (device x is referenced in place
// only during the call, so it is copied too)
static int keep_x(otn_problem* h, const double* xdev, int batch) {
    if (xdev != h->x) CUDA_TRY(cudaMemcpyAsync(h->x, xdev, xbytes(h, batch), cudaMemcpyDeviceToDevice, h->stream));
    return 0;
}

extern "C" int otn_model(otn_problem* h, const double* x, int batch, double* M) {
    OTN_TRY(check_batch(h, batch));
    if (h->cplx) return fail("otn_model: not available for complex isometries (use otn_kraus2superop + otn_compose on (re, im) planes)");
    DeviceGuard dev_guard__(h->device);
    if (!x || !M) return fail("null argument");
    const void* xd; void* Md; int tiles;
    OTN_TRY(stage_in(x, xbytes(h, batch), h->in_x, h->stream, &xd));
    const size_t mbytes = (size_t)batch * h->D * h->D * 8;
    OTN_TRY(stage_out(M, mbytes, h->out_M, &Md));
    OTN_TRY(build_layers(h, (const double*)xd, nullptr, batch, nullptr, true));
    if (h->factored) {
        OTN_TRY(fact_eval(h, batch, nullptr, FACT_MODEL, (double*)Md));
    } else {
        OTN_TRY(forward_chain(h, batch, nullptr, true, &tiles));
        OTN_TRY(export_model(h, batch, (double*)Md));
    }
    OTN_TRY(finish_out(M, Md, mbytes, h->stream));
    h->cached_batch = 0;
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int otn_cost(otn_problem* h, const double* x, int batch, double* f) {
    OTN_TRY(check_batch(h, batch));
    DeviceGuard dev_guard__(h->device);
    if (!x || !f) return fail("null argument");
    const void* xd; void* fd;
    OTN_TRY(stage_in(x, xbytes(h, batch), h->in_x, h->stream, &xd));
    OTN_TRY(stage_out(f, (size_t)batch * 8, h->out_f, &fd));
    OTN_TRY(otn_dev_cost(h, (const double*)xd, batch, nullptr, (double*)fd));
    OTN_TRY(finish_out(f, fd, (size_t)batch * 8, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int otn_cost_grad(otn_problem* h, const double* x, int batch, double* f, double* rgrad, double* egrad) {
    OTN_TRY(check_batch(h, batch));
    DeviceGuard dev_guard__(h->device);
    if (!x) return fail("null argument");
    const void* xd; void *fd, *gd, *ed;
    OTN_TRY(stage_in(x, xbytes(h, batch), h->in_x, h->stream, &xd));
    OTN_TRY(stage_out(f, (size_t)batch * 8, h->out_f, &fd));
    OTN_TRY(stage_out(rgrad, xbytes(h, batch), h->out_a, &gd));
    OTN_TRY(stage_out(egrad, xbytes(h, batch), h->out_b, &ed));
    OTN_TRY(keep_x(h, (const double*)xd, batch));
    OTN_TRY(otn_dev_cost_grad(h, h->x, batch, nullptr, (double*)fd, (double*)gd, (double*)ed));
    h->cached_batch = batch;
    OTN_TRY(finish_out(f, fd, (size_t)batch * 8, h->stream));
    OTN_TRY(finish_out(rgrad, gd, xbytes(h, batch), h->stream));
    OTN_TRY(finish_out(egrad, ed, xbytes(h, batch), h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return 0;
}

// one U1 pass over `batch` instances whose x / eta / outputs are device arrays (work matrices of instances 0..batch-1)
static int triple_device(otn_problem* h, const double* x, const double* eta, int batch, double* f, double* rgrad, double* heta) {
    int tiles;
    // primal + tangent layers in one assembly pass, then forward / reverse / tangent chains
    OTN_TRY(build_layers(h, x, eta, batch, nullptr, true));
    if (h->factored) {
        OTN_TRY(fact_eval(h, batch, nullptr, FACT_HVP, nullptr));
        h->tr.tiles = tiles = h->f_tiles;
        OTN_TRY(fact_back_layers(h, batch, nullptr, x, false, nullptr));
        OTN_TRY(fact_back_layers(h, batch, nullptr, x, true, eta));
        OTN_TRY(riem_launch(h, batch, nullptr, tiles, x, eta, true, f, rgrad, nullptr, heta));
        return 0;
    }
    const bool joint = h->structured && h->model == OTN_MODEL_FULL;   // W_l is pulled back against x and eta in one pass
    if (h->fused && h->m >= 2) {
        // forward, residual, reverse and tangent sweeps of every instance in ONE launch per diagonal block, then the pull-backs
        h->tr.tiles = tiles = h->tiles;
        OTN_TRY(hchain(h, batch, nullptr, CHAIN_TRIPLE));
        if (back_all_ok(h)) {
            OTN_TRY(back_all_layers(h, batch, nullptr, x, eta, true, true));          // every layer, W and Wd, in one launch
        } else {
            for (int l = h->m - 1; l >= 0; --l) OTN_TRY(back_layer(h, batch, nullptr, l, x, false, joint ? eta : nullptr, joint));
            for (int l = h->m - 1; l >= 0; --l) OTN_TRY(back_layer(h, batch, nullptr, l, x, true, eta, joint));
        }
        OTN_TRY(riem_launch(h, batch, nullptr, tiles, x, eta, true, f, rgrad, nullptr, heta));
        return 0;
    }
    OTN_TRY(forward_chain(h, batch, nullptr, false, &tiles));
    h->tr.tiles = tiles;
    OTN_TRY(reverse_chain(h, batch, nullptr, x, joint ? eta : nullptr));
    OTN_TRY(tangent_pass(h, batch, nullptr, x, eta, true, joint));
    OTN_TRY(riem_launch(h, batch, nullptr, tiles, x, eta, true, f, rgrad, nullptr, heta));
    return 0;
}

// Move every per-instance array of the handle by `db` instances: a chunk of the host-staged pipeline then works on its own slice
// of the workspaces (and on its own target indices), so the kernels of neighbouring chunks may overlap.
static void shift_instance_views(otn_problem* h, long long db) {
    const long long SZ = h->SZ, m = h->m, q2 = (long long)h->q * h->q;
    auto mv = [&](double*& ptr, long long stride) { if (ptr) ptr += db * stride; };
    for (double** ptr : {&h->Y, &h->P, &h->G, &h->W, &h->Yd, &h->Pd}) mv(*ptr, m * SZ);
    mv(h->Gd, 2 * SZ); mv(h->Wd, h->wd_stride);
    mv(h->Yl, m * q2); mv(h->Yld, m * q2); mv(h->Wl, m * h->nf * q2); mv(h->Wld, m * h->nf * q2); mv(h->be3, 4 * 65536);
    for (double** ptr : {&h->x, &h->eta, &h->zraw, &h->zdraw}) mv(*ptr, m * h->np);
    mv(h->part_f, h->tiles); mv(h->part_s, h->tiles);
    if (h->factored) {
        mv(h->f_Pst, (long long)h->f_tiles * (m - 1) * 2 * h->D * h->f_ncols);
        mv(h->f_Wp, (long long)h->f_tiles * m * h->nf * 512);
        mv(h->f_Wp2, (long long)h->f_nsplit * m * h->nf * 512);
    }
    h->tindex += db;
}

// Host arrays in, host arrays out: the batch is cut into chunks and pipelined so that the PCIe copies of chunk c+1 (H2D) and
// chunk c-1 (D2H) overlap the kernels of chunk c; chunks alternate between two compute streams and work on disjoint slices of
// the workspaces, so the tail of one chunk's kernels overlaps the head of the next.  Double-buffered staging, ordering by events.
// `ticket` == nullptr: synchronous (returns when the host arrays are filled).  Otherwise the call only enqueues: the slot-reuse
// events order its chunks behind those of earlier submissions, so the head H2D of this call overlaps the kernels of the previous
// one and the tail D2H of the previous call overlaps the kernels of this one; *ticket identifies the completion event.
static int triple_pipelined(otn_problem* h, const double* x, const double* eta, int batch, double* f, double* rgrad, double* heta,
                            long long* ticket = nullptr, bool dev = false) {
    const int C = 256;
    const size_t per = (size_t)h->m * h->np;
    if (!h->s_in) {
        CUDA_TRY(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CUDA_TRY(cudaStreamCreateWithFlags(&h->s_comp[i], cudaStreamNonBlocking));
            CUDA_TRY(cudaStreamCreateWithFlags(&h->s_side[i], cudaStreamNonBlocking));
            CUDA_TRY(cudaStreamCreateWithFlags(&h->s_backp[i], cudaStreamNonBlocking));
            CUDA_TRY(cudaEventCreateWithFlags(&h->ev_cs_done[i], cudaEventDisableTiming));
        }
        for (int i = 0; i < otn_problem::NSLOT; ++i) {
            CUDA_TRY(cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&h->ev_out[i], cudaEventDisableTiming));
        }
        for (cudaEvent_t& e : h->ev_ticket) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    // staging: per slot  x | eta | rgrad | heta | f
    const int CMAX = 3 * C;
    const size_t slot_doubles = 4 * (size_t)CMAX * per + CMAX;
    // dev: every array is device resident (otn_triple_submit with device pointers): no staging and no copies, the chunks read and
    // write the caller's arrays in place; what remains is the schedule -- chunks alternating between the two compute streams on
    // disjoint workspace slices, no host synchronisation between submissions -- so that the first kernels of one submission run
    // under the last kernels of the previous one.  Ordered behind the handle's stream at submission time.
    constexpr int NSLOT = otn_problem::NSLOT;
    if (!dev) OTN_TRY(h->pipe_buf.ensure(NSLOT * slot_doubles * 8));
    if (!dev) CUDA_TRY(cudaStreamSynchronize(h->stream));
    // (copies are not the bottleneck: PCIe moves the 403 MB of a step in 4.2 ms, both directions overlapped)
    // chunk schedule: short first and last chunks (exposed head H2D / tail D2H), longer ones in between (kernel efficiency)
    std::vector<int> sizes;
    if (const char* env = getenv("OTN_PIPE_SCHEDULE")) {          // experiment hook: comma-separated chunk sizes
        int left = batch;
        for (const char* q = env; *q && left > 0;) {
            int n = atoi(q);
            if (n <= 0) break;
            n = std::min(std::min(n, CMAX), left);
            sizes.push_back(n); left -= n;
            while (*q && *q != ',') ++q;
            if (*q == ',') ++q;
        }
        while (left > 0) { const int n = std::min(C, left); sizes.push_back(n); left -= n; }
    } else if (ticket != nullptr) {
        // streamed: head and tail copies hide behind the neighbouring submissions, so few large chunks (kernel efficiency) win;
        // measured at 1024 instances (ms per step): 512/512 8.81, 4 x 256 9.07, 128/256/256/256/128 9.66, 96/288/288/256/96 9.69
        for (int left = batch; left > 0; left -= 2 * C) sizes.push_back(std::min(2 * C, left));
    } else if (batch >= 4 * C) {
        // measured at 1024 instances (ms per step): 96/288/288/256/96 10.37, 64/192/256/256/192/64 10.44, 128/384/384/128 10.68,
        // 8 x 128 10.83
        const int edge = 3 * C / 8, mid = 9 * C / 8;
        int left = batch - 2 * edge;
        sizes.push_back(edge);
        while (left > 0) { const int n = std::min(mid, left); sizes.push_back(n); left -= n; }
        sizes.push_back(edge);
    } else {
        for (int left = batch; left > 0; left -= C) sizes.push_back(std::min(C, left));
    }
    const int nchunks = (int)sizes.size();
    int b0 = 0;
    for (int c = 0; c < nchunks; ++c, ++h->pipe_seq) {
        const int slot = (int)(h->pipe_seq % NSLOT), cs = (int)(h->pipe_seq & 1), nb = sizes[c];   // staging slot, compute stream
        double* base = dev ? nullptr : h->pipe_buf.as<double>() + slot * slot_doubles;
        double *dx = base, *de = dx + (size_t)CMAX * per, *dg = de + (size_t)CMAX * per, *dh = dg + (size_t)CMAX * per, *df = dh + (size_t)CMAX * per;
        cudaStream_t comp = h->s_comp[cs];
        if (dev) {
            dx = const_cast<double*>(x) + (size_t)b0 * per; de = const_cast<double*>(eta) + (size_t)b0 * per;
            dg = rgrad + (size_t)b0 * per; dh = heta + (size_t)b0 * per; df = f + b0;
            CUDA_TRY(cudaEventRecord(h->ev_in[slot], h->stream));                            // inputs are ready in the handle's stream order
            CUDA_TRY(cudaStreamWaitEvent(comp, h->ev_in[slot], 0));
        } else {
            if (h->pipe_seq >= NSLOT) CUDA_TRY(cudaStreamWaitEvent(h->s_in, h->ev_done[slot], 0));   // kernels of chunk c-NSLOT finished reading dx/de
            CUDA_TRY(cudaMemcpyAsync(dx, x + (size_t)b0 * per, nb * per * 8, cudaMemcpyHostToDevice, h->s_in));
            CUDA_TRY(cudaMemcpyAsync(de, eta + (size_t)b0 * per, nb * per * 8, cudaMemcpyHostToDevice, h->s_in));
            CUDA_TRY(cudaEventRecord(h->ev_in[slot], h->s_in));
            CUDA_TRY(cudaStreamWaitEvent(comp, h->ev_in[slot], 0));
            if (h->pipe_seq >= NSLOT) CUDA_TRY(cudaStreamWaitEvent(comp, h->ev_out[slot], 0));   // D2H of chunk c-NSLOT drained dg/dh/df
        }
        // Chunks on the two compute streams overlap in time and must not share a workspace slice.  Inside one call the slices
        // are disjoint; across calls a chunk may meet the slice of an EARLIER chunk that ran on the other stream (any earlier
        // chunk, not only the most recent one: an odd chunk count flips the slot parity between calls).  pipe_unsynced[s] holds
        // the slices enqueued on the other stream since stream s last waited for it; a chunk that overlaps one of them waits for
        // the other stream's newest chunk (stream order covers the older ones) and clears the list.
        {
            auto& mine = h->pipe_unsynced[cs];
            bool clash = mine.size() > 16;                      // bound the list: an occasional extra wait is harmless
            for (const auto& r : mine) if (b0 < r.first + r.second && r.first < b0 + nb) { clash = true; break; }
            if (clash) {
                CUDA_TRY(cudaStreamWaitEvent(comp, h->ev_cs_done[cs ^ 1], 0));
                mine.clear();
            }
            auto& theirs = h->pipe_unsynced[cs ^ 1];
            bool known = false;
            for (const auto& r : theirs) if (r.first == b0 && r.second == nb) { known = true; break; }
            if (!known) theirs.push_back({b0, nb});
        }
        cudaStream_t user_stream = h->stream;
        h->stream = comp;
        cudaStream_t user_side = h->prof.side.st;
        if (h->prof.side.enabled) h->prof.side.st = h->s_side[cs];   // each compute stream forks to its own side stream
        cudaStream_t user_back = h->s_back;
        if (h->s_back) h->s_back = h->s_backp[cs];                   // ... and has its own back stream
        shift_instance_views(h, b0);
        const int rc = triple_device(h, dx, de, nb, df, dg, dh);
        shift_instance_views(h, -(long long)b0);
        h->s_back = user_back;
        h->prof.side.st = user_side;
        h->stream = user_stream;
        if (rc != 0) return rc;
        CUDA_TRY(cudaEventRecord(h->ev_done[slot], comp));
        CUDA_TRY(cudaEventRecord(h->ev_cs_done[cs], comp));
        CUDA_TRY(cudaStreamWaitEvent(h->s_out, h->ev_done[slot], 0));
        if (!dev) {
            if (f) CUDA_TRY(cudaMemcpyAsync(f + b0, df, (size_t)nb * 8, cudaMemcpyDeviceToHost, h->s_out));
            if (rgrad) CUDA_TRY(cudaMemcpyAsync(rgrad + (size_t)b0 * per, dg, nb * per * 8, cudaMemcpyDeviceToHost, h->s_out));
            if (heta) CUDA_TRY(cudaMemcpyAsync(heta + (size_t)b0 * per, dh, nb * per * 8, cudaMemcpyDeviceToHost, h->s_out));
        }
        CUDA_TRY(cudaEventRecord(h->ev_out[slot], h->s_out));
        b0 += nb;
    }
    h->cached_batch = 0;   // the work matrices hold the last chunk only
    if (ticket != nullptr) {
        const long long t = h->next_ticket++;
        CUDA_TRY(cudaEventRecord(h->ev_ticket[t % otn_problem::TICKETS], h->s_out));   // s_out runs the D2H copies in order
        h->pipe_inflight = true;
        *ticket = t;
        return 0;
    }
    CUDA_TRY(cudaStreamSynchronize(h->s_in));
    for (int i = 0; i < 2; ++i) CUDA_TRY(cudaStreamSynchronize(h->s_comp[i]));
    CUDA_TRY(cudaStreamSynchronize(h->s_out));
    h->pipe_inflight = false;
    for (auto& v : h->pipe_unsynced) v.clear();
    return 0;
}

extern "C" int otn_triple(otn_problem* h, const double* x, const double* eta, int batch, double* f, double* rgrad, double* heta) {
    OTN_TRY(check_batch(h, batch));
    DeviceGuard dev_guard__(h->device);
    if (!x || !eta) return fail("null argument");
    const bool all_host = !is_device_ptr(x) && !is_device_ptr(eta) && (!f || !is_device_ptr(f)) && (!rgrad || !is_device_ptr(rgrad)) &&
                          (!heta || !is_device_ptr(heta));
    if (all_host && batch >= 512 && h->rs.world == 1) return triple_pipelined(h, x, eta, batch, f, rgrad, heta);
    const void *xd, *ed; void *fd, *gd, *hd;
    OTN_TRY(stage_in(x, xbytes(h, batch), h->in_x, h->stream, &xd));
    OTN_TRY(stage_in(eta, xbytes(h, batch), h->in_eta, h->stream, &ed));
    OTN_TRY(stage_out(f, (size_t)batch * 8, h->out_f, &fd));
    OTN_TRY(stage_out(rgrad, xbytes(h, batch), h->out_a, &gd));
    OTN_TRY(stage_out(heta, xbytes(h, batch), h->out_c, &hd));
    OTN_TRY(keep_x(h, (const double*)xd, batch));
    OTN_TRY(triple_device(h, h->x, (const double*)ed, batch, (double*)fd, (double*)gd, (double*)hd));
    h->cached_batch = batch;
    OTN_TRY(finish_out(f, fd, (size_t)batch * 8, h->stream));
    OTN_TRY(finish_out(rgrad, gd, xbytes(h, batch), h->stream));
    OTN_TRY(finish_out(heta, hd, xbytes(h, batch), h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int otn_triple_submit(otn_problem* h, const double* x, const double* eta, int batch, double* f, double* rgrad, double* heta,
                                 long long* ticket) {
    OTN_TRY(check_batch(h, batch, /*drain=*/false));
    DeviceGuard dev_guard__(h->device);
    if (!x || !eta || !ticket) return fail("null argument");
    if (h->rs.world > 1) return fail("otn_triple_submit: not available on a row-sharded handle (use otn_triple)");
    const bool dx = is_device_ptr(x);
    if (dx) {
        if (!f || !rgrad || !heta) return fail("otn_triple_submit with device arrays needs f, rgrad and heta");
        if (!is_device_ptr(eta) || !is_device_ptr(f) || !is_device_ptr(rgrad) || !is_device_ptr(heta))
            return fail("otn_triple_submit: the arrays of one submission must be all host or all device resident");
    } else if (is_device_ptr(eta) || (f && is_device_ptr(f)) || (rgrad && is_device_ptr(rgrad)) || (heta && is_device_ptr(heta))) {
        return fail("otn_triple_submit: the arrays of one submission must be all host or all device resident");
    }
    // the completion event of ticket t - TICKETS is about to be re-recorded: that submission must have finished
    if (h->next_ticket >= otn_problem::TICKETS && h->ev_ticket[0] != nullptr)
        CUDA_TRY(cudaEventSynchronize(h->ev_ticket[h->next_ticket % otn_problem::TICKETS]));
    return triple_pipelined(h, x, eta, batch, f, rgrad, heta, ticket, dx);
}

extern "C" int otn_triple_wait(otn_problem* h, long long ticket) {
    if (h == nullptr) return fail("null handle");
    if (ticket < 0) return drain_pipeline(h);
    if (ticket >= h->next_ticket) return fail("ticket %lld was never issued (next is %lld)", ticket, h->next_ticket);
    if (ticket + otn_problem::TICKETS < h->next_ticket) return 0;   // its event was re-used: submit waited for it already
    DeviceGuard dev_guard__(h->device);
    CUDA_TRY(cudaEventSynchronize(h->ev_ticket[ticket % otn_problem::TICKETS]));
    if (ticket + 1 == h->next_ticket) {   // the newest submission is done, hence all of them: nothing left in flight
        CUDA_TRY(cudaStreamSynchronize(h->s_out));
        h->pipe_inflight = false;
    }
    return 0;
}

extern "C" int otn_hvp_cached(otn_problem* h, const double* eta, int batch, double* heta) {
    OTN_TRY(check_batch(h, batch));
    DeviceGuard dev_guard__(h->device);
    if (!eta || !heta) return fail("null argument");
    if (h->cached_batch < batch) return fail("no primal cache for %d instances: call otn_cost_grad or otn_triple first", batch);
    const void* ed; void* hd;
    OTN_TRY(stage_in(eta, xbytes(h, batch), h->in_eta, h->stream, &ed));
    OTN_TRY(stage_out(heta, xbytes(h, batch), h->out_c, &hd));
    OTN_TRY(otn_dev_hvp(h, h->x, (const double*)ed, batch, nullptr, (double*)hd));
    OTN_TRY(finish_out(heta, hd, xbytes(h, batch), h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// stateless representation / manifold entry points: per-thread scratch + the legacy default stream of `device`
// ---------------------------------------------------------------------------------------------------------------------
struct Scratch {
    DevBuf a, b, c, o, s, x1, x2;                   // inputs a/b/c, output o, small/int s, extras x1/x2 (work space, second output)
    void release() { for (DevBuf* d : {&a, &b, &c, &o, &s, &x1, &x2}) d->release(); }
};
static thread_local Scratch g_scratch[16];

static int use_device(int device) {
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev || device >= 16) return fail("device %d not available", device);
    return 0;
}

// Frees the calling thread's staging / work buffers of the stateless entry points on `device` (they are grow-only: an N = 6
// complex expm leaves ~2 GB behind); device < 0: all devices.
extern "C" int otn_release_scratch(int device) {
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    for (int dv = 0; dv < ndev && dv < 16; ++dv) {
        if (device >= 0 && dv != device) continue;
        DeviceGuard dev_guard__(dv);
        cudaDeviceSynchronize();
        g_scratch[dv].release();
    }
    return 0;
}

extern "C" int otn_ortho2super(const double* x, int count, int dim, int K, double* Y, int device) {
    if (!x || !Y || count < 1 || dim < 1 || K < 1) return fail("bad argument");
    if (K > ASM_MAXK) return fail("Kraus rank %d > %d not supported", K, ASM_MAXK);
    OTN_TRY(use_device(device));
    DeviceGuard dev_guard__(device);
    Scratch& S = g_scratch[device];
    const size_t np = (size_t)dim * K * dim, d4 = (size_t)dim * dim * dim * dim;
    const void* xd; void* Yd;
    OTN_TRY(stage_in(x, count * np * 8, S.a, 0, &xd));
    OTN_TRY(stage_out(Y, count * d4 * 8, S.o, &Yd));
    OTN_TRY(set_smem(assemble_kernel<false, true>, np * 8));
    for (int c0 = 0; c0 < count; c0 += 32768) {
        const int c = std::min(32768, count - c0);
        assemble_kernel<false, true><<<dim3(dim, c), ASM_THREADS, np * 8>>>((const double*)xd + c0 * np, nullptr, (double*)Yd + c0 * d4,
                                                                              nullptr, (long long)d4, dim, K, nullptr, 1);
    }
    CUDA_TRY(cudaGetLastError());
    OTN_TRY(finish_out(Y, Yd, count * d4 * 8, 0));
    CUDA_TRY(cudaDeviceSynchronize());
    return 0;
}

extern "C" int otn_embed_local(const double* Yloc, int count, int N, int d, int shift_pbc, double* Yfull, int device) {
    if (!Yloc || !Yfull || count < 1) return fail("bad argument");
    if (N % 2 != 0) return fail("Only even number of sites allowed");
    if (N / 2 > 4 || ipow(d, 4) > 255) return fail("embedding supports N <= 8 and d <= 3");
    OTN_TRY(use_device(device));
    DeviceGuard dev_guard__(device);
    Scratch& S = g_scratch[device];
    std::vector<unsigned> tab; std::vector<int> inv;
    build_embed_tables(N, d, tab, inv);
    const int D = ipow(d, 2 * N), q = ipow(d, 4);
    OTN_TRY(S.s.ensure(tab.size() * 4));
    CUDA_TRY(cudaMemcpy(S.s.ptr, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
    const size_t q2 = (size_t)q * q, SZ = (size_t)D * D;
    const void* yd; void* Yd;
    OTN_TRY(stage_in(Yloc, count * q2 * 8, S.a, 0, &yd));
    OTN_TRY(stage_out(Yfull, count * SZ * 8, S.o, &Yd));
    EmbedTables T{S.s.as<unsigned>(), nullptr, D, q, N / 2};
    // layer parity selects the table: pass m = 2 and offset the item index so that (item % 2) == shift
    // (simplest: launch per item with an explicit table pointer)
    if (shift_pbc) T.tab += D;
    for (int c = 0; c < count; ++c) {
        EmbedTables Tc = T;
        embed_kernel<false, true><<<dim3((D + EMB_ROWS - 1) / EMB_ROWS, 1), EMB_THREADS, q2 * 16>>>(
            (const double*)yd + c * q2, nullptr, (double*)Yd + c * SZ, nullptr, Tc, nullptr, 2);
    }
    CUDA_TRY(cudaGetLastError());
    OTN_TRY(finish_out(Yfull, Yd, count * SZ * 8, 0));
    CUDA_TRY(cudaDeviceSynchronize());
    return 0;
}

extern "C" int otn_compose(const double* Ys, int count, int m, int D, double* M, int device) {
    if (!Ys || !M || count < 1 || m < 1 || D < 1) return fail("bad argument");
    OTN_TRY(use_device(device));
    DeviceGuard dev_guard__(device);
    Scratch& S = g_scratch[device];
    const size_t SZ = (size_t)D * D;
    const void* yd; void* Md;
    OTN_TRY(stage_in(Ys, count * m * SZ * 8, S.a, 0, &yd));
    OTN_TRY(stage_out(M, count * SZ * 8, S.o, &Md));
    if (m == 1) { CUDA_TRY(cudaMemcpy(Md, yd, count * SZ * 8, cudaMemcpyDeviceToDevice)); }
    OTN_TRY(S.b.ensure(2 * count * SZ * 8));
    const double* prev = (const double*)yd; long long sprev = (long long)m * SZ;
    for (int l = 1; l < m; ++l) {
        GemmParams p{};
        p.D = D; p.rows = D; p.batch = count; p.epi = EPI_NONE;
        p.A1 = (const double*)yd + l * SZ; p.sA1 = (long long)m * SZ;
        p.B1 = prev; p.sB1 = sprev;
        double* outp = (l == m - 1) ? (double*)Md : S.b.as<double>() + (size_t)(l & 1) * count * SZ;
        p.C = outp; p.sC = SZ;
        p.splitk = gemm_want_split(D);
        CUDA_TRY(gemm_launch(p, false, false, 0));
        prev = outp; sprev = SZ;
    }
    OTN_TRY(finish_out(M, Md, count * SZ * 8, 0));
    CUDA_TRY(cudaDeviceSynchronize());
    return 0;
}

static int project_like(const double* x, const double* z, int count, int n, int p, int mode, double a0, double a1, double* out, int device) {
    if (!x || !z || !out || count < 1 || n < p || p < 1) return fail("bad argument");
    if (p > ST_MAXP) return fail("p = %d > %d not supported", p, ST_MAXP);
    OTN_TRY(use_device(device));
    DeviceGuard dev_guard__(device);
    Scratch& S = g_scratch[device];
    const size_t np = (size_t)n * p;
    const void *xd, *zd; void* od;
    OTN_TRY(stage_in(x, count * np * 8, S.a, 0, &xd));
    OTN_TRY(stage_in(z, count * np * 8, S.b, 0, &zd));
    OTN_TRY(stage_out(out, count * np * 8, S.o, &od));
    const size_t sm = (2 * np + 2 * (size_t)p * p) * 8;
    OTN_TRY(set_smem(project_kernel, sm));
    project_kernel<<<count, ST_THREADS, sm>>>((const double*)xd, (const double*)zd, (double*)od, n, p, mode, a0, a1);
    CUDA_TRY(cudaGetLastError());
    OTN_TRY(finish_out(out, od, count * np * 8, 0));
    CUDA_TRY(cudaDeviceSynchronize());
    return 0;
}
extern "C" int otn_project(const double* x, const double* z, int count, int n, int p, double* out, int device) {
    return project_like(x, z, count, n, p, 0, 1.0, 1.0, out, device);
}
extern "C" int otn_rgrad_map(const double* x, const double* z, int count, int n, int p, double a0, double a1, double* out, int device) {
    if (!(a0 > 0.0) || !(a1 > 0.0)) return fail("alpha0 and alpha1 must be positive");
    return project_like(x, z, count, n, p, 1, a0, a1, out, device);
}

extern "C" int otn_retract(const double* x, const double* eta, int count, int n, int p, double* out, int* status, int device) {
    if (!x || !out || count < 1 || n < p || p < 1) return fail("bad argument");
    if (p > ST_MAXP) return fail("p = %d > %d not supported", p, ST_MAXP);
    OTN_TRY(use_device(device));
    DeviceGuard dev_guard__(device);
    Scratch& S = g_scratch[device];
    const size_t np = (size_t)n * p;
    const void *xd, *ed; void *od, *sd;
    OTN_TRY(stage_in(x, count * np * 8, S.a, 0, &xd));
    OTN_TRY(stage_in(eta, count * np * 8, S.b, 0, &ed));
    OTN_TRY(stage_out(out, count * np * 8, S.o, &od));
    OTN_TRY(stage_out(status, (size_t)count * 4, S.s, &sd));
    if (sd) CUDA_TRY(cudaMemset(sd, 0, (size_t)count * 4));
    const size_t sm = (np + 5 * (size_t)p * p) * 8;
    OTN_TRY(set_smem(retract_kernel, sm));
    retract_kernel<<<count, ST_THREADS, sm>>>((const double*)xd, (const double*)ed, (double*)od, n, p, 1, nullptr, (int*)sd);
    CUDA_TRY(cudaGetLastError());
    OTN_TRY(finish_out(out, od, count * np * 8, 0));
    OTN_TRY(finish_out(status, sd, (size_t)count * 4, 0));
    CUDA_TRY(cudaDeviceSynchronize());
    return 0;
}

extern "C" int otn_connection(const double* dnu, const double* nu, const double* eta, const double* x, int count, int n, int p,
                              double alpha0, double alpha1, double* out, int device) {
    if (!dnu || !nu || !eta || !x || !out || count < 1 || n < p || p < 1) return fail("bad argument");
    if (p > ST_MAXP) return fail("p = %d > %d not supported", p, ST_MAXP);
    if (!(alpha0 > 0.0)) return fail("alpha0 must be positive");
    OTN_TRY(use_device(device));
    DeviceGuard dev_guard__(device);
    Scratch& S = g_scratch[device];
    const size_t bytes = (size_t)count * n * p * 8;
    const void *dd, *nd, *ed, *xd; void* od;
    OTN_TRY(stage_in(dnu, bytes, S.a, 0, &dd));
    OTN_TRY(stage_in(nu, bytes, S.b, 0, &nd));
    OTN_TRY(stage_in(eta, bytes, S.c, 0, &ed));
    OTN_TRY(stage_in(x, bytes, S.x1, 0, &xd));
    OTN_TRY(stage_out(out, bytes, S.o, &od));
    const size_t sm = (4 * (size_t)n * p + 5 * (size_t)p * p) * 8;
    OTN_TRY(set_smem(connection_kernel, sm));
    connection_kernel<<<count, ST_THREADS, sm>>>((const double*)dd, (const double*)nd, (const double*)ed, (const double*)xd, (double*)od,
                                                 n, p, (alpha0 - alpha1) / alpha0);
    CUDA_TRY(cudaGetLastError());
    OTN_TRY(finish_out(out, od, bytes, 0));
    CUDA_TRY(cudaDeviceSynchronize());
    return 0;
}

extern "C" int otn_frobenius(const double* A, const double* B_re, const double* B_im, int count, long long n, int squared, double* out,
                             int device) {
    if (!A || !B_re || !out || count < 1 || n < 1) return fail("bad argument");
    OTN_TRY(use_device(device));
    DeviceGuard dev_guard__(device);
    Scratch& S = g_scratch[device];
    const size_t bytes = (size_t)count * n * 8;
    const void *ad, *brd, *bid; void* od;
    OTN_TRY(stage_in(A, bytes, S.a, 0, &ad));
    OTN_TRY(stage_in(B_re, bytes, S.b, 0, &brd));
    OTN_TRY(stage_in(B_im, bytes, S.c, 0, &bid));
    OTN_TRY(stage_out(out, (size_t)count * 8, S.o, &od));
    frobenius_kernel<<<count, 256>>>((const double*)ad, (const double*)brd, (const double*)bid, n, squared, (double*)od);
    CUDA_TRY(cudaGetLastError());
    OTN_TRY(finish_out(out, od, (size_t)count * 8, 0));
    CUDA_TRY(cudaDeviceSynchronize());
    return 0;
}

extern "C" int otn_canonical_dot(const double* x, const double* a, const double* b, int batch, int m, int n, int p, double* out, int device) {
    if (!x || !a || !b || !out || batch < 1 || m < 1 || n < p || p < 1) return fail("bad argument");
    if (p > ST_MAXP) return fail("p = %d > %d not supported", p, ST_MAXP);
    OTN_TRY(use_device(device));
    DeviceGuard dev_guard__(device);
    Scratch& S = g_scratch[device];
    const size_t bytes = (size_t)batch * m * n * p * 8;
    const void *xd, *ad, *bd; void* od;
    OTN_TRY(stage_in(x, bytes, S.a, 0, &xd));
    OTN_TRY(stage_in(a, bytes, S.b, 0, &ad));
    OTN_TRY(stage_in(b, bytes, S.c, 0, &bd));
    OTN_TRY(stage_out(out, (size_t)batch * 8, S.o, &od));
    const size_t sm = (3 * (size_t)n * p + 2 * (size_t)p * p) * 8;
    OTN_TRY(set_smem(cdot_kernel, sm));
    cdot_kernel<<<batch, ST_THREADS, sm>>>((const double*)xd, (const double*)ad, (const double*)bd, (double*)od, n, p, m, nullptr);
    CUDA_TRY(cudaGetLastError());
    OTN_TRY(finish_out(out, od, (size_t)batch * 8, 0));
    CUDA_TRY(cudaDeviceSynchronize());
    return 0;
}

// Row-window product on device arrays: C[row0:row0+rows, :] = op(A1) op(B1) [+ op(A2) op(B2)] restricted to those rows of the
// result (all operands D x D row-major, device pointers, `stream` = a cudaStream_t or NULL).  Building block of the row-sharded
// single-instance N = 6 composition: rank g owns a row block and all-gathers it after every chain step.
extern "C" int otn_gemm_rows(const double* A1, const double* B1, const double* A2, const double* B2, double* C, int D, int row0,
                             int rows, int transA, int transB, int device, void* stream) {
    if (!A1 || !B1 || !C || D < 1 || rows < 1 || row0 < 0 || row0 + rows > D) return fail("bad argument");
    if ((A2 == nullptr) != (B2 == nullptr)) return fail("A2 and B2 must be given together");
    if (transA && transB) return fail("A^T B^T is not supported");
    OTN_TRY(use_device(device));
    DeviceGuard dev_guard__(device);
    for (const void* ptr : {(const void*)A1, (const void*)B1, (const void*)C})
        if (!is_device_ptr(ptr)) return fail("otn_gemm_rows takes device pointers");
    GemmParams p{};
    p.A1 = A1; p.B1 = B1; p.A2 = A2; p.B2 = B2; p.C = C; p.D = D; p.row0 = row0; p.rows = rows; p.batch = 1; p.epi = EPI_NONE;
    p.splitk = gemm_want_split(D);
    ++g_launches;
    CUDA_TRY(gemm_launch(p, transA != 0, transB != 0, (cudaStream_t)stream));
    return 0;
}

// Fused product + all-gather: like otn_gemm_rows, and every tile of the row block is also stored into `peers[q]` (the same
// D x D result buffer on the other GPUs, mapped through CUDA IPC / NVLink peer access), so that no separate collective runs
// after the product.  The caller orders the steps of a chain with its own barrier.
extern "C" int otn_gemm_rows_bcast(const double* A1, const double* B1, const double* A2, const double* B2, double* C,
                                   double* const* peers, int npeer, int D, int row0, int rows, int transA, int transB, int device,
                                   void* stream) {
    if (!A1 || !B1 || !C || D < 1 || rows < 1 || row0 < 0 || row0 + rows > D) return fail("bad argument");
    if ((A2 == nullptr) != (B2 == nullptr)) return fail("A2 and B2 must be given together");
    if (transA && transB) return fail("A^T B^T is not supported");
    if (npeer < 0 || npeer > 7 || (npeer > 0 && peers == nullptr)) return fail("npeer must be in [0, 7]");
    OTN_TRY(use_device(device));
    DeviceGuard dev_guard__(device);
    GemmParams p{};
    p.A1 = A1; p.B1 = B1; p.A2 = A2; p.B2 = B2; p.C = C; p.D = D; p.row0 = row0; p.rows = rows; p.batch = 1; p.epi = EPI_NONE;
    p.npeer = npeer;
    for (int q = 0; q < npeer; ++q) p.Cpeer[q] = peers[q];
    p.splitk = gemm_want_split(D);
    ++g_launches;
    CUDA_TRY(gemm_launch(p, transA != 0, transB != 0, (cudaStream_t)stream));
    return 0;
}

// IPC-shareable device buffers for the peer stores above
extern "C" int otn_ipc_alloc(long long bytes, int device, void** ptr, unsigned char* handle64) {
    if (!ptr || !handle64 || bytes <= 0) return fail("bad argument");
    OTN_TRY(use_device(device));
    DeviceGuard dev_guard__(device);
    CUDA_TRY(cudaMalloc(ptr, (size_t)bytes));
    cudaIpcMemHandle_t hd;
    CUDA_TRY(cudaIpcGetMemHandle(&hd, *ptr));
    static_assert(sizeof(hd) == 64, "cudaIpcMemHandle_t is 64 bytes");
    std::memcpy(handle64, &hd, 64);
    return 0;
}
extern "C" int otn_ipc_open(const unsigned char* handle64, int device, void** ptr) {
    if (!ptr || !handle64) return fail("bad argument");
    OTN_TRY(use_device(device));
    DeviceGuard dev_guard__(device);
    cudaIpcMemHandle_t hd;
    std::memcpy(&hd, handle64, 64);
    CUDA_TRY(cudaIpcOpenMemHandle(ptr, hd, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
extern "C" int otn_ipc_close(void* ptr, int device) {
    OTN_TRY(use_device(device));
    DeviceGuard dev_guard__(device);
    CUDA_TRY(cudaIpcCloseMemHandle(ptr));
    return 0;
}
extern "C" int otn_ipc_free(void* ptr, int device) {
    OTN_TRY(use_device(device));
    DeviceGuard dev_guard__(device);
    CUDA_TRY(cudaFree(ptr));
    return 0;
}


// ---------------------------------------------------------------------------------------------------------------------
// row-sharded handles (BASELINE config 5: one dense N = 6 instance over the GPUs of a node)
// ---------------------------------------------------------------------------------------------------------------------
// One process per GPU, each with an identical handle (same sizes, target, OTN_FLAG_DENSE).  Protocol:
//   otn_rowshard_export(h, mine)           -> OTN_ROWSHARD_HANDLES = 10 IPC handles of this rank (P, G, W, Pd, Gd, Wd, part_f,
//                                             part_s, back-embedding scratch, flags)
//   (exchange `mine` between the ranks: all-gather of 10 * 64 bytes)
//   otn_rowshard_attach(h, rank, world, all)   all = [world][10][64]
// From then on every entry point of the handle (otn_cost, otn_cost_grad, otn_triple, otn_hvp_cached, otn_tr_run, ...) must be
// called by ALL ranks with the same arguments: the replicated kernels (assembly, embedding, pull-backs, Riemannian epilogue, tCG)
// run everywhere on identical data, the chain products are split by rows (hgemm).  Results are bitwise equal on every rank and
// bitwise equal to the un-sharded handle (same tiles, same fixed-order partial sums).
static int rs_arrays(otn_problem* h, double** base, size_t* bytes) {
    const size_t B = (size_t)h->max_batch, SZ = (size_t)h->SZ, M = (size_t)h->m;
    double* b[otn_problem::RowShard::NARR] = {h->P, h->G, h->W, h->Pd, h->Gd, h->Wd, h->part_f, h->part_s, h->be3};
    const size_t n[otn_problem::RowShard::NARR] = {B * M * SZ * 8, B * M * SZ * 8, B * M * SZ * 8, B * M * SZ * 8, B * 2 * SZ * 8,
                                                   B * (size_t)h->wd_stride * 8, B * h->tiles * 8, B * h->tiles * 8, B * 4 * 65536 * 8};
    for (int a = 0; a < otn_problem::RowShard::NARR; ++a) { base[a] = b[a]; bytes[a] = n[a]; }
    return 0;
}
static int rs_check_handle(const otn_problem* h) {
    if (!h) return fail("null handle");
    if (h->structured || h->factored || h->fused || h->cplx)
        return fail("row sharding applies to the dense chain: create the handle with OTN_FLAG_DENSE (real isometries)");
    return 0;
}

extern "C" int otn_rowshard_export(otn_problem* h, unsigned char* handles) {
    OTN_TRY(rs_check_handle(h));
    if (!handles) return fail("null argument");
    OTN_TRY(drain_pipeline(h));
    DeviceGuard dev_guard__(h->device);
    if (h->rs.flags == nullptr) {
        CUDA_TRY(cudaMalloc(&h->rs.flags, 64 * sizeof(int)));
        CUDA_TRY(cudaMemset(h->rs.flags, 0, 64 * sizeof(int)));
    }
    OTN_TRY(rs_arrays(h, h->rs.base, h->rs.bytes));
    for (int a = 0; a <= otn_problem::RowShard::NARR; ++a) {
        cudaIpcMemHandle_t hd;
        void* ptr = (a < otn_problem::RowShard::NARR) ? (void*)h->rs.base[a] : (void*)h->rs.flags;
        std::memset(handles + (size_t)a * 64, 0, 64);
        if (ptr == nullptr) continue;                          // (the back-embedding scratch exists for three-factor layers only)
        CUDA_TRY(cudaIpcGetMemHandle(&hd, ptr));
        std::memcpy(handles + (size_t)a * 64, &hd, 64);
    }
    h->rs.exported = true;
    return 0;
}

extern "C" int otn_rowshard_detach(otn_problem* h) {
    if (!h) return fail("null handle");
    DeviceGuard dev_guard__(h->device);
    if (h->stream) CUDA_TRY(cudaStreamSynchronize(h->stream));
    for (int q = 0; q < 7; ++q) {
        for (int a = 0; a < otn_problem::RowShard::NARR; ++a)
            if (h->rs.peer[a][q]) { cudaIpcCloseMemHandle(h->rs.peer[a][q]); h->rs.peer[a][q] = nullptr; }
        if (h->rs.peer_flags[q]) { cudaIpcCloseMemHandle(h->rs.peer_flags[q]); h->rs.peer_flags[q] = nullptr; }
    }
    h->rs.world = 1; h->rs.rank = 0;
    return 0;
}

extern "C" int otn_rowshard_attach(otn_problem* h, int rank, int world, const unsigned char* all_handles) {
    OTN_TRY(rs_check_handle(h));
    if (!all_handles) return fail("null argument");
    if (world < 1 || world > 8 || rank < 0 || rank >= world) return fail("rank %d / world %d outside [0, 8]", rank, world);
    if (!h->rs.exported) return fail("otn_rowshard_attach: call otn_rowshard_export first");
    if (h->rs.world > 1) return fail("handle is already row-sharded (otn_rowshard_detach first)");
    if (h->D % (64 * world) != 0) return fail("row sharding needs D = %d to be a multiple of 64 * world = %d", h->D, 64 * world);
    if (gemm_plan(h->D, h->D / world, 1).kind != gemm_plan(h->D, h->D, 1).kind) return fail("row window changes the tile shape");
    OTN_TRY(drain_pipeline(h));
    DeviceGuard dev_guard__(h->device);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    int k = 0;
    for (int r = 0; r < world; ++r) {
        if (r == rank) continue;
        for (int a = 0; a <= otn_problem::RowShard::NARR; ++a) {
            cudaIpcMemHandle_t hd;
            std::memcpy(&hd, all_handles + ((size_t)r * (otn_problem::RowShard::NARR + 1) + a) * 64, 64);
            if (a < otn_problem::RowShard::NARR && h->rs.base[a] == nullptr) continue;
            void* ptr = nullptr;
            const cudaError_t e = cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                fail("cudaIpcOpenMemHandle (rank %d, array %d) failed: %s", r, a, cudaGetErrorString(e));
                h->rs.world = 1; otn_rowshard_detach(h);
                return -1;
            }
            if (a < otn_problem::RowShard::NARR) h->rs.peer[a][k] = (double*)ptr; else h->rs.peer_flags[k] = (int*)ptr;
        }
        ++k;
    }
    if (h->rs.side == nullptr) {
        CUDA_TRY(cudaStreamCreateWithFlags(&h->rs.side, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&h->rs.ev_fork, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&h->rs.ev_join, cudaEventDisableTiming));
    }
    // barrier epochs restart at zero: clear the flags a previous attachment may have left behind.  No peer writes into them before
    // the host-side barrier the caller places between this call and the first sharded call (see otn.h).
    CUDA_TRY(cudaMemset(h->rs.flags, 0, 64 * sizeof(int)));
    h->rs.rank = rank; h->rs.world = world; h->rs.epoch = 0;
    h->cached_batch = 0;
    return 0;
}

// 0 = healthy; 1 = a device-side barrier of this rank timed out (a peer never arrived): results since then are garbage
extern "C" int otn_rowshard_status(otn_problem* h, int* timed_out) {
    if (!h || !timed_out) return fail("null argument");
    *timed_out = 0;
    if (h->rs.flags == nullptr) return 0;
    DeviceGuard dev_guard__(h->device);
    int v = 0;
    CUDA_TRY(cudaMemcpy(&v, h->rs.flags + 32, sizeof(int), cudaMemcpyDeviceToHost));
    *timed_out = v != 0;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// target construction (section 8f "next")
// ---------------------------------------------------------------------------------------------------------------------
#include "expm.inl"
#include "liouville.inl"
#include "factory.inl"
#include "extras.inl"

// ---------------------------------------------------------------------------------------------------------------------
// trust region
// ---------------------------------------------------------------------------------------------------------------------
#include "tr_driver.inl"
