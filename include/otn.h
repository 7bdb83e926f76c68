/*
 * libotn -- C ABI of the B200-native (sm_100a, FP64) Stiefel trust-region fit of trotterised Kraus/LPDO channels.
 *
 * The reference (EmilianoG-byte/opentn) is pure Python and defines no FFI; its boundary for this path is a set of
 * Python callables with fixed signatures.  Each entry point below names the reference interface
 * (file:line under the reference tree) whose arithmetic it replaces.  `INTEGRATION.md` shows the ctypes stub a
 * maintainer of the reference would add; `opentn_b200/_lib.py` is that stub in this repository.
 *
 * Conventions
 *   - every array is row-major float64 (int32 for index/status arrays); no torch / numpy types cross this boundary;
 *   - every data pointer may be a HOST or a DEVICE pointer (detected with cudaPointerGetAttributes); host arrays are
 *     staged through the handle's device buffers, device arrays are used in place;
 *   - an isometry batch is x[batch][m][n][p]; model FULL: n = d^N * K, p = d^N; model LOCAL: n = d^2 * K, p = d^2
 *     (row index a*K + k, E_k = x[k::K, :], opentn/transformations.py:716-726);
 *   - superoperators are D x D with D = d^(2N), row (out, out*), column (in, in*), row-major vectorisation;
 *   - tangent vectors cross the boundary as (n,p) matrices (basis independent), not as (A-upper,B) coefficient vectors;
 *   - return value 0 = success, negative = error (message via otn_last_error(), thread local);
 *   - a handle is not thread-safe; distinct handles are independent.  All work of a call is enqueued on the handle's
 *     stream and the call returns after that stream has been synchronised.
 */
#ifndef OTN_H
#define OTN_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct otn_problem otn_problem;

enum { OTN_MODEL_FULL = 0, OTN_MODEL_LOCAL = 1 };
enum { OTN_METRIC_EUCLIDEAN = 0, OTN_METRIC_CANONICAL = 1 };

/* status bits reported per instance by otn_tr_run / otn_retract */
enum { OTN_STATUS_OK = 0, OTN_STATUS_NAN = 1, OTN_STATUS_NOT_SPD = 2, OTN_STATUS_NEG_DISCRIMINANT = 4 };

const char* otn_last_error(void);
int otn_version(void);
/* hash (sha256, hex) of the CUDA / header sources this binary was compiled from; opentn_b200/build.py rebuilds when it differs from
 * the sources on disk, so a shipped binary can always be tied to its source */
const char* otn_build_id(void);

/* ---- problem handle ------------------------------------------------------------------------------------------------
 * A handle owns: the targets exp(tau L) (n_targets of them, real parts [n_targets][D][D]; imaginary parts may be NULL
 * and only contribute the constant ||Im T||^2 under the square root), the index tables of the local->full embedding,
 * all work matrices for `max_batch` instances, and one CUDA stream.
 * Replaces the closure  f = lambda xi: frobenius_norm(model(xi), target)  the reference notebooks build
 * (experiments/trust_region_pauli_noise.ipynb cell 21) with model = model_stiefel_local (optimization.py:106) or
 * model_Ys_stiefel (optimization.py:147). */
int otn_create(otn_problem** out, int N, int d, int m, int K, int model, int metric, int max_batch, int n_targets,
               const double* target_re, const double* target_im, int device);
/* flags: OTN_FLAG_DENSE forces the dense D x D formulation of the chain (what the reference executes).  By default the
 * full-sites model is evaluated in the swap-symmetric basis, where every superoperator of the path is block diagonal
 * (136 + 120 instead of 256 at N = 4): identical results, 3.95x fewer chain flops (opentn_b200/csrc/structured.cuh).
 * OTN_FLAG_FACTORED (2-site model, d = 2, N = 4 or 6): no D x D layer or product is ever formed; the 16 x 16 local
 * superoperators are applied factor by factor to the columns of the running product, each CTA owning a tile of columns for the
 * whole forward / reverse / tangent evaluation (opentn_b200/csrc/factored.cuh): 2*16*D instead of 2*D^2 flops per column and
 * factor, and O(D * columns) instead of O(m D^2) memory per instance. */
enum { OTN_FLAG_DENSE = 1, OTN_FLAG_BLOCKS = 2, OTN_FLAG_FACTORED = 4, OTN_FLAG_COMPLEX = 8 };   /* none: factored where it applies, else blocks when
                                                                            * max_batch * D >= 4096 (enough work to fill the GPU) */
/* OTN_FLAG_COMPLEX (full-sites model): COMPLEX isometries, x^H x = 1 -- the extension the reference marks "TODO: add back the conj"
 * (stiefel.py:29,36,44,184,201,214,218,320).  Every isometry / tangent array of the handle's entry points (x, eta, rgrad, egrad,
 * heta) is then interleaved complex128 [batch][m][n][p][2]; layers are Y = sum_k E_k (x) conj(E_k) (kraus2superop,
 * transformations.py:170-180), the target is used as given (re and im planes), gradients are in the convention
 * df = Re tr(Z^H dx), and ^T becomes ^H throughout stiefel.py:23-30, :398-409, :468-475.  otn_cost / otn_cost_grad / otn_triple /
 * otn_hvp_cached / otn_triple_submit work; otn_model and otn_tr_run do not (host trust-region loop instead). */
int otn_create_ex(otn_problem** out, int N, int d, int m, int K, int model, int metric, int max_batch, int n_targets,
                  const double* target_re, const double* target_im, int device, int flags);
void otn_destroy(otn_problem* p);
/* target id of each instance (default: all 0) */
int otn_set_target_index(otn_problem* p, const int* index, int batch);
/* metric of the Riemannian gradient / connection: (alpha0, alpha1) of stiefel.py:398-409; OTN_METRIC_* presets */
int otn_set_metric(otn_problem* p, int metric);
/* any (alpha0, alpha1) > 0 of the two-parameter family gradient_stiefel_general takes (stiefel.py:411-426); gradient, connection
 * and Hessian-vector products follow it, the trust region's inner product stays the coefficient-vector dot of the reference */
int otn_set_metric_general(otn_problem* p, double alpha0, double alpha1);
/* n, p, D, m, per-instance workspace bytes */
int otn_query(const otn_problem* p, int* n, int* pp, int* D, int* m, long long* workspace_bytes);
/* the handle's cudaStream_t (as void*) so a host framework can order its own work after ours */
void* otn_stream(const otn_problem* p);
/* run on the caller's stream instead (e.g. torch.cuda.current_stream().cuda_stream); the handle no longer owns it */
int otn_set_stream(otn_problem* p, void* stream);

/* ---- measurement hooks (bench.py) ------------------------------------------------------------------------------------
 * otn_launch_count   : kernels launched by this library in this process so far
 * otn_profile_enable : when on, every kernel launch of the handle is bracketed by a CUDA-event pair on its stream
 * otn_profile_read   : sums (and clears) the recorded timings per category
 *                      [0]=chain GEMM  [1]=Kraus assembly/embedding  [2]=back-assembly  [3]=Riemannian epilogue  [4]=TR/tCG  [5]=other
 *                      ms[6], launches[6], flops[6] (algorithmic flops of the recorded launches) */
long long otn_launch_count(void);
int otn_profile_enable(otn_problem* p, int on);
int otn_profile_read(otn_problem* p, double* ms, long long* launches, double* flops);

/* ---- model / cost layer ---------------------------------------------------------------------------------------------
 * otn_model     : M[batch][D][D] = Y_m ... Y_1        model_stiefel_local / model_Ys_stiefel (optimization.py:106,147)
 * otn_cost      : f[batch] = || M - T ||_F            frobenius_norm (optimization.py:56-61)
 * otn_cost_grad : + Riemannian gradient rgrad[batch][m][n][p]   gradient_stiefel_general (stiefel.py:411-426);
 *                 egrad (nullable) receives what jax.grad(func)(xi) returns (stiefel.py:423)
 * otn_triple    : unit of work U1 = (f, rgrad, H eta) sharing the primal pass; H eta is
 *                 P_X(riemannian_connection(jvp(rgrad)[eta], rgrad, eta, x)), i.e. one matrix-vector product with the
 *                 matrix riemannian_hessian_vec builds column by column (stiefel.py:477-544), as tangent matrices
 * otn_hvp_cached: H eta at the x of the most recent otn_cost_grad / otn_triple call on this handle (what
 *                 truncated_cg's `hess @ d` needs, trust_region_rcopt.py:116) */
int otn_model(otn_problem* p, const double* x, int batch, double* M);
int otn_cost(otn_problem* p, const double* x, int batch, double* f);
int otn_cost_grad(otn_problem* p, const double* x, int batch, double* f, double* rgrad, double* egrad);
int otn_triple(otn_problem* p, const double* x, const double* eta, int batch, double* f, double* rgrad, double* heta);
int otn_hvp_cached(otn_problem* p, const double* eta, int batch, double* heta);

/* ---- streamed form of otn_triple for host arrays ("_async variant") ---------------------------------------
 * otn_triple_submit : enqueues the chunked H2D / kernels / D2H pipeline of one otn_triple call and returns at once with a
 *                     ticket; the caller may submit the next batch (other host arrays) before waiting, so the copies of one
 *                     batch overlap the kernels of its neighbours.  The host arrays (pinned memory for real overlap) must stay
 *                     valid and untouched until the ticket has been waited for.  At most 4 tickets are outstanding: a 5th
 *                     submit first waits for the oldest.  Replaces a loop of cost / gradient / `hess @ d` evaluations over
 *                     many parameter sets (the tau x rank x seed sweeps of the reference's notebooks), stiefel.py:411-544.
 *                     Device arrays (all of x, eta, f, rgrad, heta device resident): no copies; the submission is ordered behind
 *                     the handle's stream and computes in place in the caller's arrays, without the host synchronisation that
 *                     ends otn_triple, so consecutive submissions overlap across their boundary.  Outputs are valid after
 *                     otn_triple_wait (they are not ordered into the handle's stream).
 * otn_triple_wait   : blocks until the outputs of `ticket` are in the caller's arrays (ticket < 0: all of them).
 *                     Any other call on the handle waits for all outstanding submissions first. */
int otn_triple_submit(otn_problem* p, const double* x, const double* eta, int batch, double* f, double* rgrad, double* heta,
                      long long* ticket);
int otn_triple_wait(otn_problem* p, long long ticket);

/* ---- representation layer (stateless; `device` selects the GPU) -----------------------------------------------------
 * otn_ortho2super : Y[count][dim^2][dim^2] = sum_k E_k (x) E_k      ortho2super (transformations.py:743-749),
 *                   == kraus2superop of E_k = x[k::K,:] (transformations.py:170-180)
 * otn_embed_local : Yfull[count][D][D] from Yloc[count][d^4][d^4]     create_supertensored_from_local(pbc=True)
 *                   (transformations.py:405-429) + convert_supertensored2liouvillianfull(shift_pbc) (:484-490)
 * otn_compose     : M[count][D][D] = Ys[count][m-1] @ ... @ Ys[count][0]   compose_superops_list (transformations.py:841-851) */
int otn_ortho2super(const double* x, int count, int dim, int K, double* Y, int device);
/* otn_kraus2superop : out[count][dim^2][dim^2] = sum_k E_k (x) conj(E_k) for explicit, possibly COMPLEX Kraus operators
 *                   E[count][K][dim][dim] given as (re, im) planes (E_im == NULL: real; then out_im may be NULL)
 *                   kraus2superop (transformations.py:170-180, `np.kron(kraus, kraus.conj())`) */
int otn_kraus2superop(const double* E_re, const double* E_im, int count, int K, int dim, double* out_re, double* out_im, int device);
/* frees the calling thread's grow-only staging buffers of the stateless entry points on `device` (< 0: every device) */
int otn_release_scratch(int device);
int otn_embed_local(const double* Yloc, int count, int N, int d, int shift_pbc, double* Yfull, int device);
int otn_compose(const double* Ys, int count, int m, int D, double* M, int device);
/* otn_gemm_rows : rows [row0, row0+rows) of  op(A1) op(B1) [+ op(A2) op(B2)]  (D x D, device pointers, asynchronous on `stream`).
 *                 One step of compose_superops_list (transformations.py:850, `jnp.dot(op, composition)`) restricted to a row
 *                 block -- the unit of the row-sharded N = 6 composition; a complex product is two dual launches. */
int otn_gemm_rows(const double* A1, const double* B1, const double* A2, const double* B2, double* C, int D, int row0, int rows,
                  int transA, int transB, int device, void* stream);
/* otn_gemm_rows_bcast : the same product fused with its all-gather -- every output tile is also stored into peers[0..npeer)
 *                 (the result buffer of the other ranks, mapped over NVLink through CUDA IPC), so the exchange overlaps the
 *                 math tile by tile and no NCCL collective follows.  Steps of a chain are ordered by the caller's barrier.
 * otn_ipc_*     : allocate an IPC-shareable device buffer / export its 64-byte handle / map a peer's buffer / release. */
int otn_gemm_rows_bcast(const double* A1, const double* B1, const double* A2, const double* B2, double* C, double* const* peers,
                        int npeer, int D, int row0, int rows, int transA, int transB, int device, void* stream);
/* otn_gemm_rows_ex  : otn_gemm_rows_bcast with the residual epilogue of the batched path: with T != NULL the row window receives
 *                     product - T and partial[tile] the squared norm of each output tile (otn_gemm_rows_tiles(D, rows) of them;
 *                     frobenius_norm, optimization.py:56-61, fused into the last product of a composition).
 * otn_peer_signal / otn_peer_wait : device-side ordering of a row-sharded product sequence, asynchronous on `stream`: after its
 *                     product a rank writes `value` into slot `slot` of every peer's flag array (system-scope fence first), before
 *                     the next product it waits until flags[0..n) (except `skip`, its own slot) are >= value.  Replaces a host
 *                     synchronise + barrier per product. */
int otn_gemm_rows_ex(const double* A1, const double* B1, const double* A2, const double* B2, double* C, double* const* peers,
                     int npeer, int D, int row0, int rows, int transA, int transB, const double* T, double* partial, int device,
                     void* stream);
int otn_gemm_rows_tiles(int D, int rows);
int otn_peer_signal(int* const* peer_flags, int npeer, int slot, int value, int device, void* stream);
int otn_peer_wait(const int* flags, int n, int skip, int value, int device, void* stream);
int otn_ipc_alloc(long long bytes, int device, void** ptr, unsigned char* handle64);
int otn_ipc_open(const unsigned char* handle64, int device, void** ptr);
int otn_ipc_close(void* ptr, int device);
int otn_ipc_free(void* ptr, int device);
/* ---- row-sharded handles (BASELINE config 5: ONE dense N = 6 instance over the GPUs of a node) ------------------------
 * One process per GPU, every rank creates the same handle with OTN_FLAG_DENSE (real isometries).  After
 *   otn_rowshard_export(h, mine)              mine = [OTN_ROWSHARD_HANDLES][64]: IPC handles of the arrays written across ranks
 *                                             (chain products, their partial sums, back-embedding scratch) + the flag array
 *   (the host exchanges them: all-gather of 640 bytes per rank)
 *   otn_rowshard_attach(h, rank, world, all)  all = [world][OTN_ROWSHARD_HANDLES][64]
 * every chain product of every entry point of the handle (otn_cost, otn_cost_grad, otn_triple, otn_hvp_cached, otn_model,
 * otn_tr_run) computes only the rows [rank D/world, (rank+1) D/world) of its result and stores each tile -- and the tile's
 * residual-norm / dot partial -- into the same array of every peer over NVLink (product fused with its all-gather); ranks are
 * ordered by device-side barriers, never by the host.  ALL ranks must make the same calls with the same arguments; the small
 * replicated kernels (assembly, embedding, pull-backs, Riemannian epilogue, tCG updates) run on every rank.  Results are bitwise
 * equal on all ranks and to the un-sharded handle.  This shards what jax.grad / jax.jvp spend their time on at D = 4096
 * (stiefel.py:423, 520 over compose_superops_list, transformations.py:841-851), which the reference could not finish
 * (experiments/trust_region_higher_sites.ipynb).  Needs D % (64 world) == 0, world <= 8.
 * The host places one barrier between the ranks after otn_rowshard_attach (every rank has mapped its peers and cleared its flags
 * before the first product stores into them) and one before otn_rowshard_detach / otn_destroy (no rank frees arrays a peer still
 * has mapped); StiefelFit.row_shard() / close() do both.
 * Anything that changes the launch sequence -- otn_profile_enable, the OTN_* environment switches -- must be set alike on all
 * ranks: the device-side barriers count launches.
 * otn_rowshard_status : *timed_out = 1 when a device-side barrier of this rank gave up after 30 s (a peer never arrived). */
#define OTN_ROWSHARD_HANDLES 10
int otn_rowshard_export(otn_problem* p, unsigned char* handles);
int otn_rowshard_attach(otn_problem* p, int rank, int world, const unsigned char* all_handles);
int otn_rowshard_detach(otn_problem* p);
int otn_rowshard_status(otn_problem* p, int* timed_out);

/* ---- target construction (SURVEY 8f "next" #1) -----------------------------------------------------------------------------
 * otn_expm : out = exp(tau * A), A a D x D real matrix (A_im == out_im == NULL) or complex as (re, im) pairs.
 *            exp_operator_dt (transformations.py:24-32; scipy / jax `expm`): scaling and squaring with a degree-16 Taylor
 *            polynomial (Paterson-Stockmeyer), every product on the chain's DMMA GEMM kernel. */
int otn_expm(const double* A_re, const double* A_im, int D, double tau, double* out_re, double* out_im, int device);
/* otn_super2ortho : x[i] = super2ortho(superop[i], rank=K) (transformations.py:728-741: super2choi :68-94, rank-truncated PSD
 *                   factor :793-802, choi2ortho :701-714) for `count` real dim^2 x dim^2 superoperators; dim^2 <= 16 (the 2-site
 *                   model).  Columns of the factor carry a sign convention (largest component positive) where the reference's
 *                   SVD leaves the sign open: ortho2super(x) and every cost are identical, x itself only up to those signs.
 *                   x: [count][dim*K][dim]. */
int otn_super2ortho(const double* superop, int count, int dim, int K, double* x, int device);
/* ---- initial-point factory, random part (SURVEY 8f "next" #2) ----------------------------------------------------------------------
 * Counter-based generator (Philox4x32-10; convention in csrc/extras.inl, restated in oracle/port.py): every number depends only
 * on (seed of the instance, stream, element index), so a sweep draws the same starts however it is sharded.  `seeds` are HOST arrays.
 * otn_random_normal  : out[count][per] standard normals (stream_id >= 0 selects an independent stream)
 * otn_random_stiefel : x[count][m][n][p], each isometry the polar factor of a standard-normal matrix (Haar-distributed; the device
 *                      counterpart of the host-side `qf(G)` starts of the reference-side studies)
 * otn_kicked_starts  : out = retract(x0, kick * xi / |xi|), xi = project(x0, G) in the canonical norm summed over the m isometries of
 *                      an instance -- "Trotter start + random tangent kick" (tau x rank x seed sweeps); project = stiefel.py:23-30,
 *                      retract = polar_decomposition_rectangular, stiefel.py:323-330 */
int otn_random_normal(const unsigned long long* seeds, int count, long long per, int stream_id, double* out, int device);
int otn_random_stiefel(const unsigned long long* seeds, int count, int m, int n, int p, double* x, int device);
int otn_kicked_starts(const double* x0, const unsigned long long* seeds, int count, int m, int n, int p, double kick, double* out,
                      int device);
/* otn_expm_batch : out[b] = exp(taus[b] * A) for b < count (tau sweeps of one generator: the Trotter layers exp(tau L),
 *                  exp(tau/2 L_odd), exp(tau L_even) of create_trotter_layers, transformations.py:387-402, for every tau of a
 *                  study in one batched launch sequence).  `taus` is a HOST array; out is [count][D][D] per plane. */
int otn_expm_batch(const double* A_re, const double* A_im, int D, const double* taus, int count, double* out_re, double* out_im,
                   int device);
/* otn_liouvillian : out = sum_t D[L_{term_op[t]} acting on the sites (term_site0[t], term_site1[t]) of an N-site register],
 *                   D[L] = L (x) conj(L) - (L^+L (x) 1)/2 - (1 (x) (L^+L)^T)/2  (vectorize_dissipative transformations.py:243-250
 *                   of op2fullspace :252-260; the odd / even / periodic sums of local_lindbladian_to_full_liouvillians :337-369
 *                   are term lists).  L: [n_ops][d^2][d^2] as (re, im) planes (L_im nullable); term arrays are HOST arrays;
 *                   out: D x D planes, D = d^(2N); n_terms <= 64, d^N <= 4096. */
int otn_liouvillian(const double* L_re, const double* L_im, int n_ops, int N, int d, const int* term_op, const int* term_site0,
                    const int* term_site1, int n_terms, double* out_re, double* out_im, int device);

/* ---- manifold layer (stateless) ---------------------------------------------------------------------------------------
 * otn_project       : Z - X sym(X^T Z)                                   project (stiefel.py:23-30)
 * otn_rgrad_map     : Z/a0 + (1/a1 - 2/a0)/2 X X^T Z - 1/(2 a1) X Z^T X   gradient_ambient2riemannian (stiefel.py:398-409)
 * otn_retract       : polar factor of X + eta                            polar_decomposition_rectangular (stiefel.py:323-330);
 *                     status (nullable, int[count]) gets OTN_STATUS_NOT_SPD where (X+eta)^T(X+eta) is not positive definite
 * otn_canonical_dot : out[batch] = sum_l tr(a_l^T (I - x_l x_l^T / 2) b_l)   canonical_metric (stiefel.py:76-83) summed over
 *                     the m isometries of an instance (= Euclidean dot of the coefficient vectors of stiefel.py:453-466) */
int otn_project(const double* x, const double* z, int count, int n, int p, double* out, int device);
int otn_rgrad_map(const double* x, const double* z, int count, int n, int p, double alpha0, double alpha1, double* out, int device);
int otn_retract(const double* x, const double* eta, int count, int n, int p, double* out, int* status, int device);
int otn_canonical_dot(const double* x, const double* a, const double* b, int batch, int m, int n, int p, double* out, int device);
/* otn_retract_complex : polar factor of X + eta for complex isometries (interleaved complex128 [count][n][p][2]): polar retraction of
 *                       stiefel.py:323-330 on the complex Stiefel manifold; status bit OTN_STATUS_NOT_SPD where the Newton-Schulz
 *                       inverse square root of (X+eta)^H (X+eta) does not converge (||.. - 1|| >= 1) */
int otn_retract_complex(const double* x, const double* eta, int count, int n, int p, double* out, int* status, int device);
/* otn_connection : D_nu + X (eta^T nu + nu^T eta)/2 + ((a0-a1)/a0)(I - X X^T)(eta nu^T + nu eta^T) X   riemannian_connection
 *                  (stiefel.py:468-475), `count` isometries at once
 * otn_frobenius  : out[c] = || A_c - (B_re_c + i B_im_c) ||_F  (squared if asked; B_im may be NULL)   frobenius_norm
 *                  (optimization.py:56-61) for `count` pairs of n-element arrays */
int otn_connection(const double* dnu, const double* nu, const double* eta, const double* x, int count, int n, int p, double alpha0,
                   double alpha1, double* out, int device);
int otn_frobenius(const double* A, const double* B_re, const double* B_im, int count, long long n, int squared, double* out, int device);

/* ---- optimiser layer --------------------------------------------------------------------------------------------------
 * Batched Riemannian trust region (Absil Alg. 10) with Steihaug truncated CG (Alg. 11), one independent run per instance,
 * all instances advanced in lock-step with per-instance masks:  riemannian_trust_region_optimize + truncated_cg
 * (trust_region_rcopt.py:5-166) with gradfunc = gradient_stiefel_vec, hessfunc = riemannian_hessian_vec (matrix-free),
 * retract = retract_stiefel. */
typedef struct {
    double rho_trust;    /* 0.125  (trust_region_rcopt.py:36) */
    double radius_init;  /* 0.01   (:37)  used when radius_inout == NULL */
    double maxradius;    /* 0.1    (:38) */
    int niter;           /* 20     (:39) */
    int tcg_maxiter;     /* <= 0 => 2 * (m * dof) (:107) */
    double tcg_abstol;   /* 1e-8   (:108) */
    double tcg_reltol;   /* 1e-6   (:109) */
} otn_tr_params;

typedef struct {
    double* f_iter;      /* [niter][batch]  cost BEFORE step k (same meaning as the reference's f_iter) ; nullable */
    double* rho;         /* [niter][batch]  nullable */
    double* radius;      /* [niter][batch]  radius AFTER the update of step k ; nullable */
    int* n_cg;           /* [niter][batch]  Hessian-vector products used by tCG in step k ; nullable */
    int* on_boundary;    /* [niter][batch]  nullable */
    int* accepted;       /* [niter][batch]  nullable */
    double* x_iter;      /* [niter+1][batch][m][n][p] iterate history (save_x=True) ; nullable */
    double* radius_inout;/* [batch] initial radii in, final radii out (warm restart, :37) ; nullable */
    int* status;         /* [batch] OTN_STATUS_* bits ; nullable */
    long long hvp_total; /* out: Hessian-vector products evaluated over all instances (tCG only: the model-value product `hess @ eta`
                            of :67 is accumulated from the tCG products, H eta = sum_j alpha_j H d_j, not recomputed) */
} otn_tr_report;

void otn_tr_default_params(otn_tr_params* prm);
/* x [batch][m][n][p] in/out: on return the iterate after the last step (x_iter[-1] of the reference) */
int otn_tr_run(otn_problem* p, double* x, int batch, const otn_tr_params* prm, otn_tr_report* rep);

/* ---- the one collective of the data-parallel path: final all-gather of costs and isometries -------------------------
 * Instances are sharded over GPUs with no exchange during the fit; when every shard has finished (otn_tr_run, or any call that
 * leaves the iterates in the handle), the costs f(x) at and the isometries x of ALL shards are gathered with one ncclAllGather
 * (NCCL is bound at run time: dlopen of libnccl.so.2).  The reference has no counterpart (single process); this replaces the
 * np.save / np.load round trips of its notebooks (experiments/trust_region_pauli_noise.ipynb cells 60-61) for a sharded sweep.
 * otn_gather       : ONE process driving `nshards` GPUs, one handle per device, `batch` instances each; f_all [nshards*batch],
 *                    x_all [nshards*batch][m][n][p] (host or device, nullable), shard-major.
 * otn_comm_*       : ONE PROCESS PER GPU: rank 0 creates the 128-byte id (otn_comm_unique_id) and hands it to the other ranks by any
 *                    means, every rank calls otn_comm_init, then otn_comm_gather with its own handle (same `batch` on every rank). */
typedef struct otn_comm otn_comm;
int otn_gather(otn_problem* const* shards, int nshards, int batch, double* f_all, double* x_all);
int otn_comm_unique_id(unsigned char* id128);
int otn_comm_init(otn_comm** out, int rank, int nranks, const unsigned char* id128, int device);
int otn_comm_gather(otn_comm* c, otn_problem* p, int batch, double* f_all, double* x_all);
void otn_comm_destroy(otn_comm* c);

#ifdef __cplusplus
}
#endif
#endif /* OTN_H */
