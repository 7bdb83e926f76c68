// Host driver of the batched trust region (included by otn_api.cu).  Mirrors riemannian_trust_region_optimize
// (opentn/trust_region_rcopt.py:5-89): per iteration  grad -> tCG (hess @ d) -> retract -> f(x), f(x_next) -> rho -> radius/accept.

extern "C" void otn_tr_default_params(otn_tr_params* prm) {
    if (!prm) return;
    prm->rho_trust = 0.125; prm->radius_init = 0.01; prm->maxradius = 0.1; prm->niter = 20;
    prm->tcg_maxiter = 0; prm->tcg_abstol = 1e-8; prm->tcg_reltol = 1e-6;
}

static int tr_alloc(otn_problem* h) {
    TrState& t = h->tr;
    if (t.cap_batch >= h->max_batch) return 0;
    const size_t B = (size_t)h->max_batch, V = B * h->m * h->np;
    const size_t bytes = 7 * (V * 8 + 16) + 8 * (B * 8 + 16) + 7 * (B * 4 + 16) + 64;   // every slice is rounded up to 16 B
    OTN_TRY(h->tr_buf.ensure(bytes));
    char* base = h->tr_buf.as<char>();
    auto take = [&](size_t nbytes) { char* r = base; base += (nbytes + 15) / 16 * 16; return r; };
    t.g = (double*)take(V * 8); t.r = (double*)take(V * 8); t.z = (double*)take(V * 8); t.d = (double*)take(V * 8);
    t.Hd = (double*)take(V * 8); t.xn = (double*)take(V * 8); t.Hz = (double*)take(V * 8);
    t.radius = (double*)take(B * 8); t.fx = (double*)take(B * 8); t.fn = (double*)take(B * 8); t.rsq = (double*)take(B * 8);
    t.stoptol = (double*)take(B * 8); t.gz = (double*)take(B * 8); t.zHz = (double*)take(B * 8); t.rho = (double*)take(B * 8);
    t.tcg_active = (int*)take(B * 4); t.on_boundary = (int*)take(B * 4); t.n_cg = (int*)take(B * 4); t.status = (int*)take(B * 4);
    t.accepted = (int*)take(B * 4); t.rejected = (int*)take(B * 4); t.n_active = (int*)take(16);
    t.cap_batch = h->max_batch;
    return 0;
}

static bool tr_use_mma(const otn_problem* h) {
    // the padded layout of tcg_step (4 arrays of n x (p+4)) must fit 227 KB; otherwise the unpadded CUDA-core kernels are used
    return (h->p == 16 || h->p == 8) && h->n % 8 == 0 && (4 * (size_t)h->n * (h->p + 4) + 3 * h->p * (h->p + 4)) * 8 <= 227 * 1024;
}

static int cdot_launch(otn_problem* h, const double* x, const double* a, const double* b, int batch, double* out) {
    ProfScope ps(&h->prof, h->stream, CAT_TR, 0.0);
    if (h->p == 16 && h->n % 4 == 0 && getenv("OTN_RIEM_STAGED") == nullptr) {
        cdot_stream_kernel<<<batch, RS_THREADS, 0, h->stream>>>(x, a, b, out, h->n, h->m, nullptr);
    } else if (tr_use_mma(h)) {
        const size_t ld = h->p + 4, sm = (3 * (size_t)h->n * ld + 2 * h->p * ld) * 8;
        if (h->p == 16) { OTN_TRY(set_smem(cdot_mma_kernel<16>, sm)); cdot_mma_kernel<16><<<batch, ST_THREADS, sm, h->stream>>>(x, a, b, out, h->n, h->m, nullptr); }
        else { OTN_TRY(set_smem(cdot_mma_kernel<8>, sm)); cdot_mma_kernel<8><<<batch, ST_THREADS, sm, h->stream>>>(x, a, b, out, h->n, h->m, nullptr); }
    } else {
        const size_t sm = (3 * (size_t)h->np + 2 * (size_t)h->p * h->p) * 8;
        OTN_TRY(set_smem(cdot_kernel, sm));
        cdot_kernel<<<batch, ST_THREADS, sm, h->stream>>>(x, a, b, out, h->n, h->p, h->m, nullptr);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

static int retract_launch(otn_problem* h, const double* x, const double* z, double* xn, int batch, int* status) {
    ProfScope ps(&h->prof, h->stream, CAT_TR, 0.0);
    const size_t pp = (size_t)h->p * h->p;
    if (tr_use_mma(h)) {
        const size_t ld = h->p + 4, sm = (2 * (size_t)h->n * ld + 2 * h->p * ld + 5 * pp) * 8;
        if (h->p == 16) { OTN_TRY(set_smem(retract_mma_kernel<16>, sm)); retract_mma_kernel<16><<<batch * h->m, ST_THREADS, sm, h->stream>>>(x, z, xn, h->n, h->m, nullptr, status); }
        else { OTN_TRY(set_smem(retract_mma_kernel<8>, sm)); retract_mma_kernel<8><<<batch * h->m, ST_THREADS, sm, h->stream>>>(x, z, xn, h->n, h->m, nullptr, status); }
    } else {
        const size_t sm = ((size_t)h->np + 5 * pp) * 8;
        OTN_TRY(set_smem(retract_kernel, sm));
        retract_kernel<<<batch * h->m, ST_THREADS, sm, h->stream>>>(x, z, xn, h->n, h->p, h->m, nullptr, status);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

template <int PT>
static int tcg_launch(otn_problem* h, const TcgParams& tp, int batch, bool init) {
    const size_t ld = PT ? PT + 4 : h->p;
    const size_t sm_init = (2 * (size_t)h->n * ld + h->p * ld) * 8, sm_step = (4 * (size_t)h->n * ld + 3 * h->p * ld) * 8;
    ProfScope ps(&h->prof, h->stream, CAT_TR, 0.0);
    if (init) { OTN_TRY(set_smem(tcg_init_kernel<PT>, sm_init)); tcg_init_kernel<PT><<<batch, ST_THREADS, sm_init, h->stream>>>(tp); }
    else { OTN_TRY(set_smem(tcg_step_kernel<PT>, sm_step)); tcg_step_kernel<PT><<<batch, ST_THREADS, sm_step, h->stream>>>(tp); }
    CUDA_TRY(cudaGetLastError());
    return 0;
}
static int tcg_dispatch(otn_problem* h, const TcgParams& tp, int batch, bool init) {
    if (h->p == 16 && h->n % 4 == 0 && getenv("OTN_RIEM_STAGED") == nullptr) {      // streaming kernels (nothing n x p staged)
        ProfScope ps(&h->prof, h->stream, CAT_TR, 0.0);
        if (init) tcg_init_stream_kernel<<<batch, RS_THREADS, 0, h->stream>>>(tp);
        else tcg_step_stream_kernel<<<batch, RS_THREADS, 0, h->stream>>>(tp);
        CUDA_TRY(cudaGetLastError());
        return 0;
    }
    if (tr_use_mma(h)) return h->p == 16 ? tcg_launch<16>(h, tp, batch, init) : tcg_launch<8>(h, tp, batch, init);
    return tcg_launch<0>(h, tp, batch, init);
}

extern "C" int otn_tr_run(otn_problem* h, double* x, int batch, const otn_tr_params* prm_in, otn_tr_report* rep) {
    OTN_TRY(check_batch(h, batch));
    DeviceGuard dev_guard__(h->device);
    if (!x) return fail("null argument");
    if (h->cplx) return fail("otn_tr_run: the batched device solver is real-only; complex isometries run the host loop "
                             "(opentn_b200.trust_region_rcopt over otn_cost_grad / otn_hvp_cached / otn_retract_complex)");
    otn_tr_params prm;
    if (prm_in) prm = *prm_in; else otn_tr_default_params(&prm);
    if (!(prm.rho_trust >= 0.0 && prm.rho_trust < 0.25)) return fail("need 0 <= rho_trust < 0.25");   // trust_region_rcopt.py:47
    if (prm.niter < 0) return fail("niter must be >= 0");
    OTN_TRY(tr_alloc(h));
    TrState& t = h->tr;
    cudaStream_t st = h->stream;
    const size_t V = xbytes(h, batch);
    const long long per = (long long)h->m * h->np;
    const int dof = h->p * (h->p - 1) / 2 + h->p * (h->n - h->p);
    const int maxiter = prm.tcg_maxiter > 0 ? prm.tcg_maxiter : 2 * h->m * dof;                        // :107

    const void* xin;
    OTN_TRY(stage_in(x, V, h->in_x, st, &xin));
    if (xin != h->x) CUDA_TRY(cudaMemcpyAsync(h->x, xin, V, cudaMemcpyDeviceToDevice, st));
    if (rep && rep->radius_inout) CUDA_TRY(cudaMemcpyAsync(t.radius, rep->radius_inout, (size_t)batch * 8, cudaMemcpyDefault, st));
    else { fill_kernel<<<(batch + 127) / 128, 128, 0, st>>>(t.radius, prm.radius_init, batch); CUDA_TRY(cudaGetLastError()); }
    CUDA_TRY(cudaMemsetAsync(t.status, 0, (size_t)batch * 4, st));
    if (rep) rep->hvp_total = 0;
    long long hvps = 0;
    auto hist_d = [&](double* base, int k) -> double* { return base ? base + (size_t)k * batch : nullptr; };
    auto hist_i = [&](int* base, int k) -> int* { return base ? base + (size_t)k * batch : nullptr; };
    // history rows may live on the host: stage per-iteration rows through small device buffers
    DevBuf& hb = h->out_b;
    OTN_TRY(hb.ensure((size_t)batch * 8 * 2 + (size_t)batch * 4 * 3 + 64));
    double* d_hrho = hb.as<double>(); double* d_hrad = d_hrho + batch;
    int* d_hncg = (int*)(d_hrad + batch); int* d_honb = d_hncg + batch; int* d_hacc = d_honb + batch;

    if (rep && rep->x_iter) CUDA_TRY(cudaMemcpyAsync(rep->x_iter, h->x, V, cudaMemcpyDefault, st));     // x_iter = [x] (:52)

    TcgParams tp{};
    tp.x = h->x; tp.r = t.r; tp.z = t.z; tp.d = t.d; tp.Hd = t.Hd; tp.Hz = t.Hz; tp.g = t.g; tp.radius = t.radius; tp.rsq = t.rsq;
    tp.stoptol = t.stoptol; tp.tcg_active = t.tcg_active; tp.on_boundary = t.on_boundary; tp.n_cg = t.n_cg; tp.status = t.status;
    tp.n_active = t.n_active; tp.n = h->n; tp.p = h->p; tp.m = h->m; tp.maxiter = maxiter; tp.abstol = prm.tcg_abstol; tp.reltol = prm.tcg_reltol;

    // launch-bound regime (few small instances): sparser host checks and a CUDA graph for the tCG inner loop
    const bool small_batch = (long long)batch * h->D <= 16384;
    // (stream capture cannot begin on the legacy default stream -- which is what otn_set_stream(h, 0) selects, e.g. torch's default
    // stream handle -- nor is it worth it on the per-thread default stream: plain launches there)
    const bool capturable = st != nullptr && st != cudaStreamLegacy && st != cudaStreamPerThread;
    const bool use_graph = small_batch && capturable && !h->prof.on && h->rs.world == 1 &&   // (a row-sharded handle's barrier epochs are launch arguments)
                           getenv("OTN_NO_GRAPH") == nullptr;   // env: A/B switch for measurements
    struct SideSuspend {       // the captured tCG step stays on one stream
        SideStream& s; bool prev;
        SideSuspend(SideStream& s_, bool on) : s(s_), prev(s_.suspended) { if (on) s.suspended = true; }
        ~SideSuspend() { s.suspended = prev; }
    } side_suspend(h->prof.side, use_graph);
    struct GraphHolder { cudaGraphExec_t exec = nullptr; bool failed = false; ~GraphHolder() { if (exec) cudaGraphExecDestroy(exec); } } cg_graph;

    for (int k = 0; k < prm.niter; ++k) {
        // grad = gradfunc(x) ; fx = f(x)  (same primal pass)                                       (:59, :63)
        // From the second iteration on, the layers, the forward chain and the residual of every ACCEPTED instance are still in
        // the work matrices: f(x_next) of the previous iteration computed them, and x = x_next now.  Only rejected instances
        // (x unchanged, forward state overwritten by the trial point) redo the forward pass.
        if (k > 0 && !h->factored) OTN_TRY(otn_dev_cost_grad_cached_forward(h, h->x, batch, t.rejected, t.fx, t.g));
        else OTN_TRY(otn_dev_cost_grad(h, h->x, batch, nullptr, t.fx, t.g, nullptr));
        if (rep && rep->f_iter) CUDA_TRY(cudaMemcpyAsync(hist_d(rep->f_iter, k), t.fx, (size_t)batch * 8, cudaMemcpyDefault, st));
        // eta, on_boundary = truncated_cg(grad, hess, radius)                                      (:61)
        OTN_TRY(tcg_dispatch(h, tp, batch, true));
        // The host only needs to know when every instance has left tCG.  Finished instances are masked on the device
        // (their CTAs exit at once), so the count of still-active instances is read back -- a stream synchronisation -- only
        // every `check` steps when the batch is small enough for launch/sync latency to matter; the Hessian-vector products
        // actually consumed are counted on the device.
        const int check = small_batch ? 4 : 1;
        int counters[2] = {batch, 0};
        CUDA_TRY(cudaMemsetAsync(t.n_active, 0, 8, st));
        auto cg_step = [&]() -> int {               // one tCG iteration: H d, then the batched CG update
            OTN_TRY(otn_dev_hvp(h, h->x, t.d, batch, t.tcg_active, t.Hd));
            CUDA_TRY(cudaMemsetAsync(t.n_active, 0, 4, st));
            OTN_TRY(tcg_dispatch(h, tp, batch, false));
            return 0;
        };
        for (int j = 0; j < maxiter && counters[0] > 0; ++j) {
            if (cg_graph.exec) {
                CUDA_TRY(cudaGraphLaunch(cg_graph.exec, st));
            } else if (use_graph && !cg_graph.failed && (k > 0 || j > 0)) {
                // every pointer of the step is owned by the handle, so the ~8 launches of one iteration are captured once (after
                // a first plain execution has configured the kernels) and replayed for the rest of this call
                cudaGraph_t graph = nullptr;
                if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
                    cudaGetLastError();                       // a stream that cannot be captured: same fallback as a failed instantiate
                    cg_graph.failed = true;
                    OTN_TRY(cg_step());
                    goto step_done;
                }
                const int rc = cg_step();
                const cudaError_t ec = cudaStreamEndCapture(st, &graph);
                if (rc == 0 && ec == cudaSuccess && cudaGraphInstantiate(&cg_graph.exec, graph, 0) == cudaSuccess) {
                    cudaGraphDestroy(graph);
                    CUDA_TRY(cudaGraphLaunch(cg_graph.exec, st));
                } else {
                    if (graph) cudaGraphDestroy(graph);
                    cudaGetLastError();
                    cg_graph.exec = nullptr; cg_graph.failed = true;
                    OTN_TRY(cg_step());
                }
            } else {
                OTN_TRY(cg_step());
            }
        step_done:
            if ((j + 1) % check == 0 || j + 1 == maxiter) {
                CUDA_TRY(cudaMemcpyAsync(counters, t.n_active, 8, cudaMemcpyDeviceToHost, st));
                CUDA_TRY(cudaStreamSynchronize(st));
            }
        }
        CUDA_TRY(cudaMemcpyAsync(counters, t.n_active, 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        hvps += counters[1];
        // model decrease  <grad, eta> + 1/2 <eta, hess @ eta>  (:67).  hess @ eta is not recomputed: eta = sum_j alpha_j d_j,
        // so H eta = sum_j alpha_j (H d_j) was accumulated by tcg_step_kernel from the products tCG already paid for.
        OTN_TRY(cdot_launch(h, h->x, t.g, t.z, batch, t.gz));
        OTN_TRY(cdot_launch(h, h->x, t.z, t.Hz, batch, t.zHz));
        // x_next = retract(x, eta) ; f(x_next)                                                     (:62, :67)
        OTN_TRY(retract_launch(h, h->x, t.z, t.xn, batch, t.status));
        OTN_TRY(otn_dev_cost(h, t.xn, batch, nullptr, t.fn));
        TrUpdateParams up{};
        up.radius = t.radius; up.rho = t.rho; up.fx = t.fx; up.fn = t.fn; up.gz = t.gz; up.zHz = t.zHz; up.on_boundary = t.on_boundary;
        up.accepted = t.accepted; up.rejected = t.rejected; up.status = t.status; up.rho_trust = prm.rho_trust; up.maxradius = prm.maxradius; up.batch = batch;
        up.h_rho = d_hrho; up.h_radius = d_hrad; up.h_ncg = d_hncg; up.h_onb = d_honb; up.h_acc = d_hacc; up.n_cg = t.n_cg;
        { ProfScope ps(&h->prof, st, CAT_TR, 0.0, 2); tr_update_kernel<<<(batch + 127) / 128, 128, 0, st>>>(up); }
        CUDA_TRY(cudaGetLastError());
        select_copy_kernel<<<dim3(8, batch), 256, 0, st>>>(h->x, t.xn, t.accepted, per);                 // (:76-77)
        CUDA_TRY(cudaGetLastError());
        if (rep) {
            if (rep->rho) CUDA_TRY(cudaMemcpyAsync(hist_d(rep->rho, k), d_hrho, (size_t)batch * 8, cudaMemcpyDefault, st));
            if (rep->radius) CUDA_TRY(cudaMemcpyAsync(hist_d(rep->radius, k), d_hrad, (size_t)batch * 8, cudaMemcpyDefault, st));
            if (rep->n_cg) CUDA_TRY(cudaMemcpyAsync(hist_i(rep->n_cg, k), d_hncg, (size_t)batch * 4, cudaMemcpyDefault, st));
            if (rep->on_boundary) CUDA_TRY(cudaMemcpyAsync(hist_i(rep->on_boundary, k), d_honb, (size_t)batch * 4, cudaMemcpyDefault, st));
            if (rep->accepted) CUDA_TRY(cudaMemcpyAsync(hist_i(rep->accepted, k), d_hacc, (size_t)batch * 4, cudaMemcpyDefault, st));
            if (rep->x_iter) CUDA_TRY(cudaMemcpyAsync(rep->x_iter + (size_t)(k + 1) * batch * per, h->x, V, cudaMemcpyDefault, st));
        }
        CUDA_TRY(cudaStreamSynchronize(st));   // history staging buffers are reused next iteration
    }
    h->cached_batch = 0;
    if (x != h->x) CUDA_TRY(cudaMemcpyAsync(x, h->x, V, cudaMemcpyDefault, st));
    if (rep) {
        if (rep->radius_inout) CUDA_TRY(cudaMemcpyAsync(rep->radius_inout, t.radius, (size_t)batch * 8, cudaMemcpyDefault, st));
        if (rep->status) CUDA_TRY(cudaMemcpyAsync(rep->status, t.status, (size_t)batch * 4, cudaMemcpyDefault, st));
        rep->hvp_total = hvps;
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}
