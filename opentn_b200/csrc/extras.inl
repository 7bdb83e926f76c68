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

extern "C" int otn_retract_complex(const double* x, const double* eta, int count, int n, int p, double* out, int* status, int device) {
    if (!x || !out || count < 1 || n < p || p < 1) return fail("bad argument");
    if (p > ST_MAXP) return fail("p = %d > %d not supported", p, ST_MAXP);
    OTN_TRY(use_device(device));
    DeviceGuard dev_guard__(device);
    Scratch& S = g_scratch[device];
    const size_t bytes = (size_t)count * n * p * 16;
    const void *xd, *ed; void *od, *sd;
    OTN_TRY(stage_in(x, bytes, S.a, 0, &xd));
    OTN_TRY(stage_in(eta, bytes, S.b, 0, &ed));
    OTN_TRY(stage_out(out, bytes, S.o, &od));
    OTN_TRY(stage_out(status, (size_t)count * 4, S.s, &sd));
    if (sd) CUDA_TRY(cudaMemset(sd, 0, (size_t)count * 4));
    const size_t sm = ((size_t)n * p + 3 * (size_t)p * p) * 16;
    OTN_TRY(set_smem(cretract_kernel, sm));
    cretract_kernel<<<count, ST_THREADS, sm>>>((const cplx*)xd, (const cplx*)ed, (cplx*)od, n, p, (int*)sd);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    OTN_TRY(finish_out(out, od, bytes, 0));
    OTN_TRY(finish_out(status, sd, (size_t)count * 4, 0));
    CUDA_TRY(cudaDeviceSynchronize());
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Initial-point factory, random part (SURVEY 8f "next" #2): counter-based normals, Haar-random isometries, kicked Trotter starts.
// Seed convention (restated in oracle/port.py: philox_normal): element e of instance i draws
//     (x0, x1, x2, x3) = Philox4x32-10(counter = (e_lo, e_hi, stream, 0), key = (seed_lo, seed_hi))
//     u1 = (((x0 << 32 | x1) >> 11) + 0.5) * 2^-53,  u2 likewise from (x2, x3),   z = sqrt(-2 ln u1) cos(2 pi u2)
// so every entry depends only on (seed, stream, e): any launch shape, any GPU count, any order gives the same numbers.
// ---------------------------------------------------------------------------------------------------------------------
namespace {

__device__ __forceinline__ void philox_round(unsigned& c0, unsigned& c1, unsigned& c2, unsigned& c3, unsigned k0, unsigned k1) {
    const unsigned long long p0 = 0xD2511F53ull * c0, p1 = 0xCD9E8D57ull * c2;
    const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0, n1 = (unsigned)p1, n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1, n3 = (unsigned)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}

__device__ __forceinline__ double philox_normal(unsigned long long seed, unsigned stream, unsigned long long e) {
    unsigned c0 = (unsigned)e, c1 = (unsigned)(e >> 32), c2 = stream, c3 = 0u;
    unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c0, c1, c2, c3, k0, k1);
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    const double u1 = ((double)((((unsigned long long)c0 << 32) | c1) >> 11) + 0.5) * 1.1102230246251565e-16;   // 2^-53
    const double u2 = ((double)((((unsigned long long)c2 << 32) | c3) >> 11) + 0.5) * 1.1102230246251565e-16;
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

__global__ void rng_normal_kernel(const unsigned long long* __restrict__ seeds, int count, long long per, unsigned stream,
                                  double* __restrict__ out) {
    const long long total = (long long)count * per;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long inst = i / per, e = i - inst * per;
        out[i] = philox_normal(seeds[inst], stream, (unsigned long long)e);
    }
}

// eta[inst][...] *= kick / sqrt(nrm2[inst])
__global__ void scale_instances_kernel(double* __restrict__ eta, const double* __restrict__ nrm2, double kick, long long per, int count) {
    const long long total = (long long)count * per;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
        eta[i] *= kick / sqrt(nrm2[i / per]);
}

static int rng_fill(const unsigned long long* seeds_host, int count, long long per, unsigned stream, double* out_dev, DevBuf& seedbuf) {
    OTN_TRY(seedbuf.ensure((size_t)count * 8));
    CUDA_TRY(cudaMemcpy(seedbuf.ptr, seeds_host, (size_t)count * 8, cudaMemcpyHostToDevice));
    const long long total = (long long)count * per;
    const int blocks = (int)std::min<long long>((total + 255) / 256, 148LL * 16);
    rng_normal_kernel<<<blocks, 256>>>(seedbuf.as<unsigned long long>(), count, per, stream, out_dev);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace

enum { OTN_RNG_STREAM_GENERIC = 0, OTN_RNG_STREAM_STIEFEL = 1, OTN_RNG_STREAM_KICK = 2 };

extern "C" int otn_random_normal(const unsigned long long* seeds, int count, long long per, int stream_id, double* out, int device) {
    if (!seeds || !out || count < 1 || per < 1 || stream_id < 0) return fail("bad argument");
    OTN_TRY(use_device(device));
    DeviceGuard dev_guard__(device);
    Scratch& S = g_scratch[device];
    void* od;
    OTN_TRY(stage_out(out, (size_t)count * per * 8, S.o, &od));
    OTN_TRY(rng_fill(seeds, count, per, (unsigned)stream_id, (double*)od, S.s));
    OTN_TRY(finish_out(out, od, (size_t)count * per * 8, 0));
    CUDA_TRY(cudaDeviceSynchronize());
    return 0;
}

// x[count][m][n][p]: every isometry is the polar factor G (G^T G)^{-1/2} of a standard-normal n x p matrix G -- uniformly (Haar)
// distributed on St(n, p), like the sign-fixed QR factor the reference-side benchmarks draw on the host (numpy qr of a Gaussian).
extern "C" int otn_random_stiefel(const unsigned long long* seeds, int count, int m, int n, int p, double* x, int device) {
    if (!seeds || !x || count < 1 || m < 1 || n < p || p < 1) return fail("bad argument");
    if (p > ST_MAXP) return fail("p = %d > %d not supported", p, ST_MAXP);
    OTN_TRY(use_device(device));
    DeviceGuard dev_guard__(device);
    Scratch& S = g_scratch[device];
    const size_t np = (size_t)n * p, bytes = (size_t)count * m * np * 8;
    void* od;
    OTN_TRY(stage_out(x, bytes, S.o, &od));
    OTN_TRY(S.a.ensure(bytes));
    OTN_TRY(rng_fill(seeds, count, (long long)m * np, OTN_RNG_STREAM_STIEFEL, S.a.as<double>(), S.s));
    const size_t sm = (np + 5 * (size_t)p * p) * 8;
    OTN_TRY(set_smem(retract_kernel, sm));
    retract_kernel<<<count * m, ST_THREADS, sm>>>(S.a.as<double>(), nullptr, (double*)od, n, p, 1, nullptr, nullptr);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    OTN_TRY(finish_out(x, od, bytes, 0));
    CUDA_TRY(cudaDeviceSynchronize());
    return 0;
}

// out[i] = retract(x0[i], kick * xi / ||xi||_canonical),  xi = P_{x0[i]}(G_i),  G_i standard normal from seeds[i]  (the kicked Trotter
// starts of a seed sweep; host counterpart: opentn_b200.sweep.kicked_start).  x0, out: [count][m][n][p], host or device.
extern "C" int otn_kicked_starts(const double* x0, const unsigned long long* seeds, int count, int m, int n, int p, double kick,
                                 double* out, int device) {
    if (!x0 || !seeds || !out || count < 1 || m < 1 || n < p || p < 1) return fail("bad argument");
    if (p > ST_MAXP) return fail("p = %d > %d not supported", p, ST_MAXP);
    if ((long long)count * m > 2000000000LL) return fail("too many isometries");
    OTN_TRY(use_device(device));
    DeviceGuard dev_guard__(device);
    Scratch& S = g_scratch[device];
    const size_t np = (size_t)n * p, pp = (size_t)p * p, per = (size_t)m * np, bytes = (size_t)count * per * 8;
    const void* xd; void* od;
    OTN_TRY(stage_in(x0, bytes, S.a, 0, &xd));
    OTN_TRY(stage_out(out, bytes, S.o, &od));
    OTN_TRY(S.b.ensure(bytes));                       // G
    OTN_TRY(S.x1.ensure(bytes));                      // its tangent projection
    OTN_TRY(S.c.ensure((size_t)count * 8));
    double* gauss = S.b.as<double>();
    double* eta = S.x1.as<double>();
    OTN_TRY(rng_fill(seeds, count, (long long)per, OTN_RNG_STREAM_KICK, gauss, S.s));
    const size_t sm_p = (2 * np + 2 * pp) * 8, sm_c = (3 * np + 2 * pp) * 8, sm_r = (np + 5 * pp) * 8;
    OTN_TRY(set_smem(project_kernel, sm_p));
    project_kernel<<<count * m, ST_THREADS, sm_p>>>((const double*)xd, gauss, eta, n, p, 0, 1.0, 1.0);
    OTN_TRY(set_smem(cdot_kernel, sm_c));
    cdot_kernel<<<count, ST_THREADS, sm_c>>>((const double*)xd, eta, eta, S.c.as<double>(), n, p, m, nullptr);
    const int blocks = (int)std::min<long long>(((long long)count * per + 255) / 256, 148LL * 16);
    scale_instances_kernel<<<blocks, 256>>>(eta, S.c.as<double>(), kick, (long long)per, count);
    OTN_TRY(set_smem(retract_kernel, sm_r));
    retract_kernel<<<count * m, ST_THREADS, sm_r>>>((const double*)xd, eta, (double*)od, n, p, 1, nullptr, nullptr);
    g_launches += 4;
    CUDA_TRY(cudaGetLastError());
    OTN_TRY(finish_out(out, od, bytes, 0));
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
// Row-sharded chain products (BASELINE config 5): device-side ordering between the ranks of a sharded product sequence.
// Every rank stores the tiles of its row block into the result buffer of every peer (otn_gemm_rows_bcast).  What was a host
// round trip per product (stream synchronise + barrier) is two tiny kernels on the same stream: after product k a rank raises
// flag[k] in every peer's flag array (system-scope fence first, so its tiles are visible), and before product k+1 it spins until
// all peers have raised theirs.  No host involvement between the products of a chain.
// ---------------------------------------------------------------------------------------------------------------------
namespace {

struct PeerFlagPtrs { int* p[8]; };

__global__ void peer_signal_kernel(PeerFlagPtrs peers, int npeer, int slot, int value) {
    __threadfence_system();                        // the kernel before this one on the stream wrote the tiles
    if ((int)threadIdx.x < npeer) {
        volatile int* f = peers.p[threadIdx.x] + slot;
        *f = value;
    }
    __threadfence_system();
}

__global__ void peer_wait_kernel(const volatile int* flags, int n, int skip, int value) {
    if ((int)threadIdx.x < n && (int)threadIdx.x != skip)
        while (flags[threadIdx.x] < value) __nanosleep(200);
    __threadfence_system();
}

}  // namespace

extern "C" int otn_peer_signal(int* const* peer_flags, int npeer, int slot, int value, int device, void* stream) {
    if (npeer < 0 || npeer > 7 || (npeer > 0 && !peer_flags)) return fail("npeer must be in [0, 7]");
    OTN_TRY(use_device(device));
    DeviceGuard dev_guard__(device);
    PeerFlagPtrs P{};
    for (int q = 0; q < npeer; ++q) P.p[q] = peer_flags[q];
    peer_signal_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(P, npeer, slot, value);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int otn_peer_wait(const int* flags, int n, int skip, int value, int device, void* stream) {
    if (!flags || n < 1 || n > 32) return fail("bad argument");
    OTN_TRY(use_device(device));
    DeviceGuard dev_guard__(device);
    peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(flags, n, skip, value);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// otn_gemm_rows_bcast with the fused residual epilogue of the batched path: C = op(A1) op(B1) [+ ...] - T on the row window, and
// partial[tile] = sum over the tile of (C - T)^2 (the caller adds the partials of all ranks in a fixed order: frobenius_norm,
// optimization.py:56-61, without a separate pass over a D x D matrix).  T == NULL: plain product.
extern "C" int otn_gemm_rows_ex(const double* A1, const double* B1, const double* A2, const double* B2, double* C,
                                double* const* peers, int npeer, int D, int row0, int rows, int transA, int transB,
                                const double* T, double* partial, int device, void* stream) {
    if (!A1 || !B1 || !C || D < 1 || rows < 1 || row0 < 0 || row0 + rows > D) return fail("bad argument");
    if ((A2 == nullptr) != (B2 == nullptr)) return fail("A2 and B2 must be given together");
    if (transA && transB) return fail("A^T B^T is not supported");
    if (npeer < 0 || npeer > 7 || (npeer > 0 && peers == nullptr)) return fail("npeer must be in [0, 7]");
    if ((T == nullptr) != (partial == nullptr)) return fail("T and partial must be given together");
    OTN_TRY(use_device(device));
    DeviceGuard dev_guard__(device);
    GemmParams p{};
    p.A1 = A1; p.B1 = B1; p.A2 = A2; p.B2 = B2; p.C = C; p.D = D; p.row0 = row0; p.rows = rows; p.batch = 1;
    p.npeer = npeer;
    for (int q = 0; q < npeer; ++q) p.Cpeer[q] = peers[q];
    p.epi = T ? EPI_RESID : EPI_NONE;
    if (T) { p.E = T; p.sE = 0; p.e_index = nullptr; p.partial = partial; p.partial_stride = gemm_plan(D, rows, 1).tiles; }
    p.splitk = gemm_want_split(D);
    ++g_launches;
    CUDA_TRY(gemm_launch(p, transA != 0, transB != 0, (cudaStream_t)stream));
    return 0;
}

// number of partial sums otn_gemm_rows_ex writes for a row window
extern "C" int otn_gemm_rows_tiles(int D, int rows) { return gemm_plan(D, rows, 1).tiles; }

// ---------------------------------------------------------------------------------------------------------------------
#ifndef OTN_SOURCE_HASH
#define OTN_SOURCE_HASH "unknown"
#endif
// the marker prefix lets build.py find the hash in the binary without loading it
extern "C" const char* otn_build_id(void) {
    static const char id[] = "OTN_SOURCE_HASH=" OTN_SOURCE_HASH;
    return id + 16;
}
