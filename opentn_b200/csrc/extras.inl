// Smaller entry points of the C ABI (included by otn_api.cu):
//   otn_kraus2superop  : complex Kraus list -> superoperator  (transformations.py:170-180)
//   otn_gather / otn_comm_* : the ONE collective of the data-parallel path, the final all-gather of costs and isometries over
//                        NCCL; NCCL is bound at run time with dlopen so that libotn.so has no link-time dependency
//   otn_build_id       : hash of the sources this binary was compiled from (opentn_b200/build.py checks it)

#include <dlfcn.h>

namespace {

// out[(a,b),(c,d)] = sum_k E_k[a,c] * conj(E_k[b,d]);  E as (re, im) planes [count][K][dim][dim] (im nullable = real Kraus
// operators), out as planes [count][dim^2][dim^2].  One CTA per (a, item): the K rows E_k[a,:] stay in shared memory, the rows
// E_k[b,:] are staged per b; every thread owns output columns (c,d) of the current row (a,b).  Writes are coalesced over (c,d).
__global__ void __launch_bounds__(256) kraus2superop_kernel(const double* __restrict__ Ere, const double* __restrict__ Eim, int K, int dim,
                                                            double* __restrict__ Ore, double* __restrict__ Oim) {
    extern __shared__ __align__(16) double sm[];
    double* ar = sm;                 // [K][dim] row a, real
    double* ai = ar + K * dim;       // imaginary
    double* br = ai + K * dim;       // [K][dim] row b
    double* bi = br + K * dim;
    const int a = blockIdx.x, item = blockIdx.y;
    const long long eoff = (long long)item * K * dim * dim;
    const long long D = (long long)dim * dim;
    for (int i = threadIdx.x; i < K * dim; i += blockDim.x) {
        const int k = i / dim, c = i - k * dim;
        ar[i] = Ere[eoff + ((long long)k * dim + a) * dim + c];
        ai[i] = Eim ? Eim[eoff + ((long long)k * dim + a) * dim + c] : 0.0;
    }
    for (int b = 0; b < dim; ++b) {
        __syncthreads();
        for (int i = threadIdx.x; i < K * dim; i += blockDim.x) {
            const int k = i / dim, c = i - k * dim;
            br[i] = Ere[eoff + ((long long)k * dim + b) * dim + c];
            bi[i] = Eim ? Eim[eoff + ((long long)k * dim + b) * dim + c] : 0.0;
        }
        __syncthreads();
        const long long row = (long long)item * D * D + ((long long)a * dim + b) * D;
        for (int cd = threadIdx.x; cd < dim * dim; cd += blockDim.x) {
            const int c = cd / dim, d = cd - c * dim;
            double sr = 0.0, si = 0.0;
            for (int k = 0; k < K; ++k) {
                const double xr = ar[k * dim + c], xi = ai[k * dim + c], yr = br[k * dim + d], yi = bi[k * dim + d];
                sr = fma(xr, yr, sr); sr = fma(xi, yi, sr);          // Re (x conj y)
                si = fma(xi, yr, si); si = fma(-xr, yi, si);         // Im (x conj y)
            }
            Ore[row + cd] = sr;
            if (Oim) Oim[row + cd] = si;
        }
    }
}

}  // namespace

extern "C" int otn_kraus2superop(const double* E_re, const double* E_im, int count, int K, int dim, double* out_re, double* out_im,
                                 int device) {
    if (!E_re || !out_re || count < 1 || K < 1 || dim < 1) return fail("bad argument");
    if (E_im != nullptr && out_im == nullptr) return fail("complex Kraus operators need out_im");
    if ((size_t)4 * K * dim * 8 > 200 * 1024) return fail("otn_kraus2superop: K * dim = %d too large for the shared-memory rows", K * dim);
    OTN_TRY(use_device(device));
    DeviceGuard dev_guard__(device);
    Scratch& S = g_scratch[device];
    const size_t ebytes = (size_t)count * K * dim * dim * 8, obytes = (size_t)count * dim * dim * dim * dim * 8;
    const void *er, *ei; void *orp, *oip;
    OTN_TRY(stage_in(E_re, ebytes, S.a, 0, &er));
    OTN_TRY(stage_in(E_im, ebytes, S.b, 0, &ei));
    OTN_TRY(stage_out(out_re, obytes, S.o, &orp));
    OTN_TRY(stage_out(out_im, obytes, S.x2, &oip));
    const size_t sm = (size_t)4 * K * dim * 8;
    OTN_TRY(set_smem(kraus2superop_kernel, sm));
    for (int c0 = 0; c0 < count; c0 += 32768) {
        const int c = std::min(32768, count - c0);
        const size_t eo = (size_t)c0 * K * dim * dim, oo = (size_t)c0 * dim * dim * dim * dim;
        kraus2superop_kernel<<<dim3(dim, c), 256, sm>>>((const double*)er + eo, ei ? (const double*)ei + eo : nullptr, K, dim,
                                                         (double*)orp + oo, oip ? (double*)oip + oo : nullptr);
        ++g_launches;
    }
    CUDA_TRY(cudaGetLastError());
    OTN_TRY(finish_out(out_re, orp, obytes, 0));
    OTN_TRY(finish_out(out_im, oip, obytes, 0));
    CUDA_TRY(cudaDeviceSynchronize());
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// NCCL, bound at run time
// ---------------------------------------------------------------------------------------------------------------------
namespace {

typedef struct ncclComm* nccl_comm_t;
struct NcclUid { char b[128]; };   // ncclUniqueId (passed by value)
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(void* id128) = nullptr;
    int (*CommInitRank)(nccl_comm_t*, int nranks, NcclUid id, int rank) = nullptr;
    int (*CommInitAll)(nccl_comm_t*, int ndev, const int* devlist) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*AllGather)(const void* send, void* recv, size_t count, int dtype, nccl_comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
enum { NCCL_DOUBLE = 8 };   // ncclFloat64

static int nccl_load() {
    if (g_nccl.lib) return 0;
    // a process that already uses NCCL (e.g. through torch) gets that copy back: same soname
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return fail("NCCL not found (dlopen libnccl.so.2): %s", dlerror());
#define NCCL_SYM(field, name)                                                          \
    do {                                                                               \
        *(void**)(&g_nccl.field) = dlsym(lib, name);                                   \
        if (!g_nccl.field) return fail("NCCL symbol %s missing", name);                \
    } while (0)
    NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
    NCCL_SYM(CommInitRank, "ncclCommInitRank");
    NCCL_SYM(CommInitAll, "ncclCommInitAll");
    NCCL_SYM(CommDestroy, "ncclCommDestroy");
    NCCL_SYM(AllGather, "ncclAllGather");
    NCCL_SYM(GroupStart, "ncclGroupStart");
    NCCL_SYM(GroupEnd, "ncclGroupEnd");
    NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef NCCL_SYM
    g_nccl.lib = lib;
    return 0;
}
#define NCCL_TRY(expr)                                                                                         \
    do {                                                                                                       \
        int r__ = (expr);                                                                                      \
        if (r__ != 0) return fail("%s failed: %s", #expr, g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "?"); \
    } while (0)

// final cost of every instance at the iterates the handle holds (h->x, left there by otn_tr_run / otn_cost_grad)
static int shard_final_cost(otn_problem* h, int batch, double* f_dev) {
    OTN_TRY(otn_dev_cost(h, h->x, batch, nullptr, f_dev));
    return 0;
}

}  // namespace

struct otn_comm {
    nccl_comm_t comm = nullptr;
    int rank = 0, nranks = 1, device = 0;
    DevBuf send, recv;
};

extern "C" int otn_comm_unique_id(unsigned char* id128) {
    if (!id128) return fail("null argument");
    OTN_TRY(nccl_load());
    NCCL_TRY(g_nccl.GetUniqueId(id128));
    return 0;
}

extern "C" int otn_comm_init(otn_comm** out, int rank, int nranks, const unsigned char* id128, int device) {
    if (!out || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail("bad argument");
    *out = nullptr;
    OTN_TRY(nccl_load());
    OTN_TRY(use_device(device));
    DeviceGuard dev_guard__(device);
    otn_comm* c = new otn_comm();
    c->rank = rank; c->nranks = nranks; c->device = device;
    NcclUid id;
    std::memcpy(id.b, id128, 128);
    const int rc = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
    if (rc != 0) { delete c; return fail("ncclCommInitRank failed: %s", g_nccl.GetErrorString(rc)); }
    *out = c;
    return 0;
}

extern "C" void otn_comm_destroy(otn_comm* c) {
    if (!c) return;
    DeviceGuard dev_guard__(c->device);
    if (c->comm) g_nccl.CommDestroy(c->comm);
    c->send.release(); c->recv.release();
    delete c;
}

// One process per GPU: every rank contributes the `batch` instances of its handle (same batch on every rank); on return f_all
// [nranks * batch] and x_all [nranks * batch][m][n][p] (host or device arrays, nullable) hold the costs at / the isometries of
// all ranks, rank-major.  This is the only collective of the data-parallel path.
extern "C" int otn_comm_gather(otn_comm* c, otn_problem* h, int batch, double* f_all, double* x_all) {
    if (!c || !h) return fail("null argument");
    OTN_TRY(check_batch(h, batch));
    if (h->device != c->device) return fail("handle lives on device %d, communicator on %d", h->device, c->device);
    DeviceGuard dev_guard__(h->device);
    const size_t per = (size_t)h->m * h->np, cnt = (size_t)batch * (per + 1);
    OTN_TRY(c->send.ensure(cnt * 8));
    OTN_TRY(c->recv.ensure(cnt * 8 * c->nranks));
    double* snd = c->send.as<double>();
    cudaStream_t st = h->stream;
    OTN_TRY(shard_final_cost(h, batch, snd));                                            // [batch] costs | [batch][per] isometries
    CUDA_TRY(cudaMemcpyAsync(snd + batch, h->x, (size_t)batch * per * 8, cudaMemcpyDeviceToDevice, st));
    { ProfScope ps(&h->prof, st, CAT_OTHER, 0.0); NCCL_TRY(g_nccl.AllGather(snd, c->recv.ptr, cnt, NCCL_DOUBLE, c->comm, st)); }
    const double* rcv = c->recv.as<double>();
    for (int r = 0; r < c->nranks; ++r) {
        if (f_all) CUDA_TRY(cudaMemcpyAsync(f_all + (size_t)r * batch, rcv + r * cnt, (size_t)batch * 8, cudaMemcpyDefault, st));
        if (x_all) CUDA_TRY(cudaMemcpyAsync(x_all + (size_t)r * batch * per, rcv + r * cnt + batch, (size_t)batch * per * 8, cudaMemcpyDefault, st));
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

// One process driving several GPUs (one handle per device): the same gather with a communicator clique built on the spot
// (ncclCommInitAll).  Every shard contributes `batch` instances; outputs as above, shard-major.
extern "C" int otn_gather(otn_problem* const* shards, int nshards, int batch, double* f_all, double* x_all) {
    if (!shards || nshards < 1 || nshards > 16) return fail("bad argument");
    OTN_TRY(nccl_load());
    std::vector<int> devs(nshards);
    for (int s = 0; s < nshards; ++s) {
        if (!shards[s]) return fail("null shard");
        OTN_TRY(check_batch(shards[s], batch));
        devs[s] = shards[s]->device;
        if (shards[s]->m != shards[0]->m || shards[s]->np != shards[0]->np) return fail("shards must share the isometry shape");
        for (int q = 0; q < s; ++q) if (devs[q] == devs[s]) return fail("otn_gather: two shards on device %d (one handle per GPU)", devs[s]);
    }
    otn_problem* h0 = shards[0];
    const size_t per = (size_t)h0->m * h0->np, cnt = (size_t)batch * (per + 1);
    std::vector<nccl_comm_t> comms(nshards, nullptr);
    NCCL_TRY(g_nccl.CommInitAll(comms.data(), nshards, devs.data()));
    std::vector<DevBuf> snd(nshards), rcv(nshards);
    int rc = 0;
    for (int s = 0; s < nshards && rc == 0; ++s) {
        DeviceGuard g(devs[s]);
        if ((rc = snd[s].ensure(cnt * 8)) != 0) break;
        if ((rc = rcv[s].ensure(cnt * 8 * nshards)) != 0) break;
        if ((rc = shard_final_cost(shards[s], batch, snd[s].as<double>())) != 0) break;
        if (cudaMemcpyAsync(snd[s].as<double>() + batch, shards[s]->x, (size_t)batch * per * 8, cudaMemcpyDeviceToDevice, shards[s]->stream) != cudaSuccess)
            rc = fail("copy of the isometries of shard %d failed", s);
    }
    if (rc == 0) {
        g_nccl.GroupStart();
        for (int s = 0; s < nshards; ++s) {
            DeviceGuard g(devs[s]);
            ++g_launches;
            const int e = g_nccl.AllGather(snd[s].ptr, rcv[s].ptr, cnt, NCCL_DOUBLE, comms[s], shards[s]->stream);
            if (e != 0 && rc == 0) rc = fail("ncclAllGather failed: %s", g_nccl.GetErrorString(e));
        }
        const int e = g_nccl.GroupEnd();
        if (e != 0 && rc == 0) rc = fail("ncclGroupEnd failed: %s", g_nccl.GetErrorString(e));
    }
    if (rc == 0) {
        DeviceGuard g(devs[0]);
        const double* r0 = rcv[0].as<double>();
        cudaStream_t st = h0->stream;
        for (int s = 0; s < nshards && rc == 0; ++s) {
            if (f_all && cudaMemcpyAsync(f_all + (size_t)s * batch, r0 + s * cnt, (size_t)batch * 8, cudaMemcpyDefault, st) != cudaSuccess) rc = fail("copy out failed");
            if (x_all && cudaMemcpyAsync(x_all + (size_t)s * batch * per, r0 + s * cnt + batch, (size_t)batch * per * 8, cudaMemcpyDefault, st) != cudaSuccess) rc = fail("copy out failed");
        }
    }
    for (int s = 0; s < nshards; ++s) {
        DeviceGuard g(devs[s]);
        cudaStreamSynchronize(shards[s]->stream);
        if (comms[s]) g_nccl.CommDestroy(comms[s]);
        snd[s].release(); rcv[s].release();
    }
    return rc;
}

// ---------------------------------------------------------------------------------------------------------------------
#ifndef OTN_SOURCE_HASH
#define OTN_SOURCE_HASH "unknown"
#endif
// the marker prefix lets build.py find the hash in the binary without loading it
extern "C" const char* otn_build_id(void) {
    static const char id[] = "OTN_SOURCE_HASH=" OTN_SOURCE_HASH;
    return id + 16;
}
