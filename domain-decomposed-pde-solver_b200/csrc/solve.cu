// solve.cu — host drivers of SpMV and preconditioned CG (role of belosSolver,
// BelosMueLuSolver.cpp:87-139, with Belos "CG" + Ifpack2 Jacobi/Chebyshev per the north-star).
//
// No host round trip inside an iteration: scalars live on the device, every kernel evaluates
// the Belos stopping test (||r||/||r0|| <= tol, checked BEFORE an iteration) itself and becomes a
// no-op once it fires, so the host only polls every `check_every` iterations.  Multi-GPU: the halo
// exchange of the SpMV input overlaps the interior slices; dot products are all-reduced as one
// (single-reduce solver) or two (classical) tiny NCCL calls per iteration.
#include "comm.cuh"
#include "device_utils.cuh"
#include "kernels.cuh"
#include "solve.cuh"

namespace heat {

int sm_count(int device) {
    static int cached[kMaxDevices] = {};                 // per device: a process may hold contexts on several GPUs
    if (device < 0 || device >= kMaxDevices) return 148;
    if (cached[device] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || v <= 0) return 148;
        cached[device] = v;
    }
    return cached[device];
}

// y = A x with the halo exchange of x overlapped with the interior slices
int spmv_halo(heat_ctx *ctx, heat_matrix *A, double *x, double *y, CgGate gate, double *dot_out, bool with_yy) {
    const int sms = sm_count(ctx->device);
    const bool split = ctx->nranks > 1 && A->halo.n_neighbors > 0;
    DotOut d{A->partials.p, 0, 0, A->iscal.p + I_COUNTER, dot_out};
    d.with_yy = with_yy;
    if (!split) {
        const int g = spmv_grid(A->n_slices, sms);
        d.total_blocks = g;
        return launch_spmv(A, x, y, 0, A->n_slices, gate, d, g, ctx->stream);
    }
    const int g1 = spmv_grid(A->n_int_slices, sms), g2 = spmv_grid(A->n_bnd_slices, sms) > 2048 ? 2048 : spmv_grid(A->n_bnd_slices, sms);
    const int g1c = g1 > 2048 ? 2048 : g1;
    HEAT_TRY(halo_begin(ctx, A, x));
    d.part_offset = 0; d.total_blocks = g1c + g2;
    HEAT_TRY(launch_spmv(A, x, y, 0, A->n_int_slices, gate, d, g1c, ctx->stream));
    HEAT_TRY(halo_end(ctx, A));
    d.part_offset = g1c;
    HEAT_TRY(launch_spmv(A, x, y, A->n_int_slices, A->n_bnd_slices, gate, d, g2, ctx->stream));
    return 0;
}

// y = A x through the PEER-MEMORY halo path — the very launch the multi-GPU CG loop makes (halo pushed into the
// neighbours' ghost segments by peer stores, ONE SpMV launch over [interior | boundary] slices, the boundary slices
// waiting on the neighbours' epoch flags) — as a stand-alone collective: parity and measurement hook (heat_spmv
// itself goes through the NCCL halo).  The kernel is launched `repeat` times on the same input; *kernel_ms is the
// average device time of one launch, *xy_global the all-reduced sum_i x_i y_i.
int spmv_peer_once(heat_ctx *ctx, heat_matrix *A, const double *x, double *y, int repeat, double *xy_global, double *kernel_ms) {
    HEAT_CUDA(cudaSetDevice(ctx->device));
    if (ctx->nranks < 2 || !ctx->peer_enabled) HEAT_FAIL(52, "heat_spmv_peer: the peer-memory path is not available (one rank, HEAT_COMM=nccl, or no CUDA IPC)");
    HEAT_TRY(ensure_workspace(A, false, false));
    if (!A->peer) HEAT_TRY(peer_matrix_setup(ctx, A));
    if (!A->peer) HEAT_FAIL(52, "heat_spmv_peer: peer-memory set-up failed for this matrix");
    if (repeat < 1) repeat = 1;
    cudaStream_t st = ctx->stream;
    const int which = (int)(ctx->peer_halo_epoch & 1);                 // alternate the two IPC-mapped input buffers
    double *pin = which ? A->w_p2.p : A->w_p.p;
    double *S = A->scal.p;
    int *I = A->iscal.p;
    HEAT_CUDA(cudaMemsetAsync(I, 0, sizeof(int) * I_COUNT, st));
    HEAT_CUDA(cudaMemcpyAsync(pin, x, sizeof(double) * (size_t)A->n_owned, cudaMemcpyDeviceToDevice, st));
    PeerPush pp = A->peer->push[which];
    pp.epoch = ctx->peer_halo_epoch + 1;
    HEAT_TRY(launch_halo_push(pin, pp, st));
    SpmvPeer sp;
    sp.on = true; sp.n_interior = A->n_int_slices; sp.halo = A->peer->halo;
    sp.halo.epoch = ctx->peer_halo_epoch + 1;
    sp.red = peer_red_of(ctx); sp.seq_out = ctx->peer_red_seq + 1; sp.I = I;
    const int g = spmv_grid(A->n_slices, sm_count(ctx->device));
    DotOut d{A->partials.p, 0, g, I + I_COUNTER, S + S_TMP2};
    CgGate nogate{nullptr, nullptr, nullptr, 0};
    HEAT_TRY(launch_spmv_peer(A, pin, y, nogate, d, sp, g, st));       // first launch: includes the wait for the halo
    HEAT_CUDA(cudaEventRecord(ctx->ev_a, st));
    for (int q = 1; q < repeat; ++q) HEAT_TRY(launch_spmv_peer(A, pin, y, nogate, d, sp, g, st));
    HEAT_CUDA(cudaEventRecord(ctx->ev_b, st));
    // the all-reduce doubles as the cross-rank fence: once it completes here, every rank's SpMV has read its
    // ghosts, so the next call may push into the same buffers
    HEAT_TRY(comm_allreduce_sum(ctx, S + S_TMP2, 1));
    double h_xy = 0.0;
    int hI[2] = {0, 0};
    HEAT_CUDA(cudaMemcpyAsync(&h_xy, S + S_TMP2, sizeof(double), cudaMemcpyDeviceToHost, st));
    HEAT_CUDA(cudaMemcpyAsync(hI, I, sizeof(hI), cudaMemcpyDeviceToHost, st));
    HEAT_CUDA(cudaStreamSynchronize(st));
    ctx->peer_halo_epoch += 4; ctx->peer_red_seq += 4;                 // identical on every rank
    if (hI[I_STATUS] == 3) HEAT_FAIL(51, "heat_spmv_peer: peer-memory halo timed out");
    if (xy_global) *xy_global = h_xy;
    if (kernel_ms) {
        float ms = 0.f;
        if (repeat > 1) HEAT_CUDA(cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b));
        *kernel_ms = repeat > 1 ? ms / (repeat - 1) : 0.0;
    }
    return 0;
}

int ensure_workspace(heat_matrix *A, bool single_reduce, bool cheb) {
    const size_t nv = (size_t)(A->n_owned + A->n_ghost);
    if (!A->w_r.p) HEAT_TRY(A->w_r.alloc((size_t)A->n_owned));
    if (!A->w_p.p) HEAT_TRY(A->w_p.alloc(nv));
    if (!A->w_ap.p) HEAT_TRY(A->w_ap.alloc((size_t)A->n_owned));
    if (single_reduce) {
        if (!A->w_s.p) HEAT_TRY(A->w_s.alloc((size_t)A->n_owned));
        if (!A->w_u.p) HEAT_TRY(A->w_u.alloc(nv));
    }
    if (cheb) {
        if (!A->w_u.p) HEAT_TRY(A->w_u.alloc(nv));                  // z (SpMV input: needs ghosts)
        if (!A->w_w.p) HEAT_TRY(A->w_w.alloc((size_t)A->n_owned));
        if (!A->w_t.p) HEAT_TRY(A->w_t.alloc((size_t)A->n_owned));
    }
    if (!A->partials.p) HEAT_TRY(A->partials.alloc((size_t)kMaxPartials * 2));
    if (!A->scal.p) HEAT_TRY(A->scal.alloc(S_COUNT));
    if (!A->iscal.p) {
        HEAT_TRY(A->iscal.alloc(I_COUNT));
        HEAT_CUDA(cudaMemset(A->iscal.p, 0, sizeof(int) * I_COUNT));
    }
    return 0;
}

struct ChebCoef { double inv_theta, delta, s1; };

// z = p_k(D^-1 A) D^-1 r, Ifpack2 recurrences, zero starting solution
int cheb_apply(heat_ctx *ctx, heat_matrix *A, const heat_solve_opts &o, double lmax, const double *r,
               double *z, CgGate gate, int grid) {
    const double ratio = o.cheb_ratio > 0 ? o.cheb_ratio : 30.0;
    const double alpha = lmax / ratio, beta = 1.1 * lmax;
    const double delta = 2.0 / (beta - alpha), theta = 0.5 * (beta + alpha), s1 = theta * delta;
    double rho = 1.0 / s1;
    HEAT_TRY(launch_cheb_first(A->n_owned, A->dinv.p, r, 1.0 / theta, A->w_w.p, z, gate, grid, ctx->stream));
    for (int d = 1; d < o.cheb_degree; ++d) {
        const double rho_new = 1.0 / (2.0 * s1 - rho);
        HEAT_TRY(spmv_halo(ctx, A, z, A->w_t.p, gate, nullptr));
        HEAT_TRY(launch_cheb_step(A->n_owned, A->dinv.p, r, A->w_t.p, rho_new * rho, 2.0 * rho_new * delta, A->w_w.p, z,
                                  gate, grid, ctx->stream));
        rho = rho_new;
    }
    return 0;
}

// lambda_max(D^-1 A) by 10 power iterations from a fixed pseudo-random start (Ifpack2 default is a
// random start; a counter-based start keeps the solve reproducible and GPU-count invariant)
int estimate_lambda_max(heat_ctx *ctx, heat_matrix *A, double *lmax_out) {
    const int grid = vec_grid(A->n_owned, sm_count(ctx->device));
    double *x = A->w_u.p;
    HEAT_TRY(launch_fill_hash(A->n_owned, x, A->owned_contiguous ? nullptr : A->d_owned_gids.p, A->gid0, 777, ctx->stream));
    double lam = 1.0;
    CgGate nogate{nullptr, nullptr, nullptr, 0};
    for (int itp = 0; itp < 10; ++itp) {
        double *y = A->w_t.p;
        HEAT_TRY(spmv_halo(ctx, A, x, y, nogate, nullptr));
        // y2 = D^-1 A x : cheb_first with r := y, theta := 1 writes it to w_w (and to the scratch w_r)
        HEAT_TRY(launch_cheb_first(A->n_owned, A->dinv.p, y, 1.0, A->w_w.p, A->w_r.p, nogate, grid, ctx->stream));
        y = A->w_w.p;
        double *S = A->scal.p;
        HEAT_TRY(launch_dot2(A->n_owned, x, y, y, y, S + S_TMP0, S + S_TMP1, A->partials.p, A->iscal.p + I_COUNTER2, grid, ctx->stream));
        HEAT_TRY(launch_dot2(A->n_owned, x, x, x, x, S + S_TMP2, S + S_TMP2 + 1, A->partials.p, A->iscal.p + I_COUNTER2, grid, ctx->stream));
        HEAT_TRY(comm_allreduce_sum(ctx, S + S_TMP0, 3));
        double h[3];
        HEAT_CUDA(cudaMemcpyAsync(h, S + S_TMP0, sizeof(double) * 3, cudaMemcpyDeviceToHost, ctx->stream));
        HEAT_CUDA(cudaStreamSynchronize(ctx->stream));
        lam = h[0] / h[2];                                           // Rayleigh quotient x.D^-1Ax / x.x
        const double ny = sqrt(h[1]);
        if (!(ny > 0.0)) break;
        HEAT_TRY(launch_axpby(A->n_owned, 1.0 / ny, y, 0.0, x, grid, ctx->stream));
    }
    *lmax_out = lam;
    return 0;
}

// Power method of ExodusMatrixTest.cpp:56-129: q = z/||z|| ; z = A q ; lambda = q.z ; every 50
// iterations (and at the last one) ||z - lambda q|| is compared with the tolerance.  Two launches
// per iteration (scale, SpMV with fused q.z and z.z); the host is only involved at the reports.
// The reference starts from Tpetra's randomize(); here the start vector is the counter-based hash
// of the global row id, so the estimate is reproducible and GPU-count invariant.
int power_method_device(heat_ctx *ctx, heat_matrix *A, int niters, double tol, uint64_t seed, heat_power_info *info) {
    HEAT_CUDA(cudaSetDevice(ctx->device));
    if (niters < 0) HEAT_FAIL(2, "heat_power_method: negative iteration count");
    HEAT_TRY(ensure_workspace(A, false, false));
    const int64_t n = A->n_owned;
    const int vgrid = vec_grid(n, sm_count(ctx->device));
    cudaStream_t st = ctx->stream;
    double *q = A->w_p.p, *z = A->w_ap.p, *S = A->scal.p;
    int *I = A->iscal.p;
    CgGate nogate{nullptr, nullptr, nullptr, 0};
    const int reportFrequency = 50;                                     // ExodusMatrixTest.cpp:91
    HEAT_CUDA(cudaMemsetAsync(I, 0, sizeof(int) * I_COUNT, st));
    HEAT_CUDA(cudaEventRecord(ctx->ev_a, st));
    HEAT_TRY(launch_fill_hash(n, z, A->owned_contiguous ? nullptr : A->d_owned_gids.p, A->gid0, seed, st));
    HEAT_TRY(launch_dot2(n, z, z, z, z, S + S_TMP1, S + S_TMP2, A->partials.p, I + I_COUNTER2, vgrid, st));
    HEAT_TRY(comm_allreduce_sum(ctx, S + S_TMP1, 1));
    double lambda = 0.0, residual = 0.0;
    int iters = niters, converged = 0, n_reports = 0;
    for (int iter = 0; iter < niters; ++iter) {
        HEAT_TRY(launch_pm_scale(n, z, S + S_TMP1, q, vgrid, st));                        // q := z / normz
        HEAT_TRY(spmv_halo(ctx, A, q, z, nogate, S + S_TMP0, true));                      // z := A q ; q.z ; z.z
        HEAT_TRY(comm_allreduce_sum(ctx, S + S_TMP0, 2));
        if (iter % reportFrequency == 0 || iter + 1 == niters) {
            HEAT_TRY(launch_pm_resid(n, z, q, S + S_TMP0, S + S_TMP2, A->partials.p, I + I_COUNTER2, vgrid, st));
            HEAT_TRY(comm_allreduce_sum(ctx, S + S_TMP2, 1));
            double h[3];
            HEAT_CUDA(cudaMemcpyAsync(h, S + S_TMP0, sizeof(h), cudaMemcpyDeviceToHost, st));
            HEAT_CUDA(cudaStreamSynchronize(st));
            lambda = h[0];
            residual = sqrt(h[2]);
            if (info && info->report_buf && n_reports < info->report_capacity) {
                double *rep = info->report_buf + 3 * (size_t)n_reports++;
                rep[0] = (double)iter; rep[1] = lambda; rep[2] = residual;
            }
            if (residual < tol) { iters = iter; converged = 1; break; }
        }
    }
    HEAT_CUDA(cudaEventRecord(ctx->ev_b, st));
    HEAT_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f;
    HEAT_CUDA(cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b));
    if (info) {
        info->lambda = lambda; info->residual = residual; info->iters = iters; info->converged = converged;
        info->solve_ms = ms; info->report_count = n_reports;
    }
    return 0;
}

int solve_device(heat_ctx *ctx, heat_matrix *A, double *x, const double *b, const heat_solve_opts &o,
                 heat_solve_info *info, const std::function<int(int)> *on_poll) {
    HEAT_CUDA(cudaSetDevice(ctx->device));
    const bool single = o.solver == HEAT_SOLVER_CG_SINGLE_REDUCE;
    const bool is_cheb = o.prec == HEAT_PREC_CHEBYSHEV, is_ilu = o.prec == HEAT_PREC_ILU0;
    const bool cheb = is_cheb || is_ilu;            // "general" preconditioner: z = M^-1 r is a separate step
    if (o.solver != HEAT_SOLVER_CG && !single && o.solver != HEAT_SOLVER_GMRES) HEAT_FAIL(2, "heat_solve: unknown solver %d", o.solver);
    if (o.prec != HEAT_PREC_NONE && o.prec != HEAT_PREC_JACOBI && !cheb) HEAT_FAIL(2, "heat_solve: unknown preconditioner %d", o.prec);
    if (cheb && single) HEAT_FAIL(2, "heat_solve: Chebyshev / ILU(0) are implemented for HEAT_SOLVER_CG and HEAT_SOLVER_GMRES only");
    if (o.max_iters < 0 || !(o.tol >= 0.0)) HEAT_FAIL(2, "heat_solve: bad max_iters/tol");
    if (o.solver == HEAT_SOLVER_GMRES) return gmres_device(ctx, A, x, b, o, info, on_poll);
    HEAT_TRY(ensure_workspace(A, single, cheb));
    if (is_ilu) HEAT_TRY(ilu0_setup(ctx, A));
    // z = M^-1 r for the general preconditioners (ILU(0) of a symmetric matrix is L D L^T: a valid CG preconditioner)
    auto gen_prec = [&](double lmax_, const double *r_, double *z_, CgGate gate_, int grid_) -> int {
        return is_ilu ? ilu_apply(ctx, A, r_, z_) : cheb_apply(ctx, A, o, lmax_, r_, z_, gate_, grid_);
    };
    // peer-memory path (peer.cuh): classical CG with the fused Jacobi/identity preconditioner
    if (!single && !cheb && ctx->nranks > 1 && ctx->peer_enabled && !A->peer) HEAT_TRY(peer_matrix_setup(ctx, A));
    const bool peer = !single && !cheb && ctx->nranks > 1 && A->peer != nullptr;
    // Chebyshev-PCG with the polynomial folded into the CG kernels (cg.cu): over peer memory on >1 GPU, and the same
    // kernels with a degenerate (no-neighbour) state on one GPU.  HEAT_CHEB_FUSED=0 / HEAT_COMM=nccl: the plain path.
    bool cheb_fused = false;
    if (is_cheb && !single && o.cheb_degree >= 1) {
        const char *cf = getenv("HEAT_CHEB_FUSED"), *hc = getenv("HEAT_COMM");
        const bool want = !(cf && atoi(cf) == 0) && !(hc && strcmp(hc, "nccl") == 0) && spmv_peer_supported();
        if (want && (ctx->nranks == 1 || ctx->peer_enabled)) {
            HEAT_TRY(peer_matrix_setup(ctx, A, true));
            cheb_fused = A->peer != nullptr && A->peer->has_z;
            if (cheb_fused) {
                if (!A->w_r2.p) HEAT_TRY(A->w_r2.alloc((size_t)A->n_owned));
                if (!A->w_w2.p) HEAT_TRY(A->w_w2.alloc((size_t)A->n_owned));
            }
        }
    }
    const int64_t n = A->n_owned;
    const int sms = sm_count(ctx->device);
    const int vgrid = vec_grid(n, sms);
    cudaStream_t st = ctx->stream;

    // unit "preconditioner" for HEAT_PREC_NONE: a vector of ones in place of D^-1
    DevBuf<double> ones;
    const double *dinv = A->dinv.p;
    if (o.prec == HEAT_PREC_NONE) {
        HEAT_TRY(ones.alloc((size_t)n));
        HEAT_TRY(launch_fill(n, ones.p, 1.0, st));
        dinv = ones.p;
    }

    DevBuf<CgRec> Hbuf;
    const size_t nrec = (size_t)o.max_iters + 2;
    HEAT_TRY(Hbuf.alloc(nrec));
    CgRec *H = Hbuf.p;
    HEAT_CUDA(cudaMemsetAsync(H, 0, sizeof(CgRec) * nrec, st));
    HEAT_CUDA(cudaMemsetAsync(A->iscal.p, 0, sizeof(int) * I_COUNT, st));
    double hS[S_COUNT] = {0};
    hS[S_TOL2] = o.tol * o.tol;
    HEAT_CUDA(cudaMemcpyAsync(A->scal.p, hS, sizeof(hS), cudaMemcpyHostToDevice, st));
    double *S = A->scal.p;
    int *I = A->iscal.p;
    double *r = A->w_r.p, *p = A->w_p.p, *ap = A->w_ap.p;
    CgGate nogate{nullptr, nullptr, nullptr, 0};

    double lmax = o.cheb_lambda_max;
    if (is_cheb && !(lmax > 0.0)) {                  // Ifpack2 estimates lambda_max once, at set-up: cached with the matrix
        if (!(A->cheb_lmax_est > 0.0)) HEAT_TRY(estimate_lambda_max(ctx, A, &A->cheb_lmax_est));
        lmax = A->cheb_lmax_est;
    }

    HEAT_CUDA(cudaEventRecord(ctx->ev_a, st));
    // ---- r0 = b - A x0 ; z0 ; p0 (or u0, w0) ; H[0] ----
    HEAT_TRY(spmv_halo(ctx, A, x, ap, nogate, nullptr));
    if (ctx->wait_before_rhs) HEAT_CUDA(cudaStreamWaitEvent(st, ctx->wait_before_rhs, 0));   // b still in flight (heat_solve_host)
    if (single) {
        double *u = A->w_u.p, *s = A->w_s.p, *w = ap;
        HEAT_TRY(launch_cg_init(n, b, ap, dinv, r, u, H, A->partials.p, I + I_COUNTER2, vgrid, st));
        HEAT_TRY(launch_fill(n, p, 0.0, st));
        HEAT_TRY(launch_fill(n, s, 0.0, st));
        HEAT_TRY(spmv_halo(ctx, A, u, w, nogate, &H[0].delta));
        HEAT_TRY(comm_allreduce_sum(ctx, &H[0].rz, 3));
    } else if (!cheb) {
        HEAT_TRY(launch_cg_init(n, b, ap, dinv, r, p, H, A->partials.p, I + I_COUNTER2, vgrid, st));
        HEAT_TRY(comm_allreduce_sum(ctx, &H[0].rz, 3));
        if (peer) {                                   // ghosts of p0 go to the neighbours by peer stores
            PeerPush pp = A->peer->push[0];
            pp.epoch = ctx->peer_halo_epoch + 1;
            HEAT_TRY(launch_halo_push(p, pp, st));
        }
    } else {
        double *z = A->w_u.p;
        HEAT_TRY(launch_cg_init(n, b, ap, A->dinv.p, r, z, H, A->partials.p, I + I_COUNTER2, vgrid, st));   // r (z overwritten below)
        HEAT_TRY(gen_prec(lmax, r, z, nogate, vgrid));
        HEAT_TRY(launch_dot2(n, r, z, r, r, &H[0].rz, &H[0].rr, A->partials.p, I + I_COUNTER2, vgrid, st));
        HEAT_TRY(comm_allreduce_sum(ctx, &H[0].rz, 3));
        HEAT_TRY(launch_axpby(n, 1.0, z, 0.0, p, vgrid, st));
        if (cheb_fused) {                             // ghosts of p0 go to the neighbours by peer stores
            PeerPush pp = A->peer->push[0];
            pp.epoch = ctx->peer_halo_epoch + 1;
            HEAT_TRY(launch_halo_push(p, pp, st));
        }
    }

#ifdef HEAT_PEER_TRACE
    DevBuf<TraceBuf> trace;
    const char *trace_file = getenv("HEAT_PEER_TRACE_FILE");
    if (peer && trace_file) {
        HEAT_TRY(trace.alloc(1));
        std::vector<unsigned long long> init((size_t)kTraceIters * 9);
        for (size_t q = 0; q < init.size(); ++q) init[q] = (q % 3 == 0) ? ~0ull : 0ull;
        HEAT_CUDA(cudaMemcpy(trace.p, init.data(), sizeof(TraceBuf), cudaMemcpyHostToDevice));
        HEAT_TRY(trace_set_cg(trace.p)); HEAT_TRY(trace_set_spmv(trace.p));
    }
#endif
    // ---- iterations, polled every check_every ----
    // consecutive kernels sweep the vectors in opposite directions (kernels.cuh: CgGate::dir); HEAT_CG_ALTERNATE=0: all forward
    const char *alt_env = getenv("HEAT_CG_ALTERNATE");
    const bool alternate = !(alt_env && atoi(alt_env) == 0) && !single && !cheb;
    // programmatic dependent launch of the loop's kernels (device_utils.cuh); HEAT_PDL=0: ordinary launches.  Only where
    // the loop is kernels of this library back to back (one GPU, or the peer-memory path): no NCCL launch in between.
    const char *pdl_env = getenv("HEAT_PDL");
    const bool pdl = !(pdl_env && atoi(pdl_env) == 0) && !single && !cheb && (ctx->nranks == 1 || peer);
    const int check = o.check_every > 0 ? o.check_every : 32;
    int launched = 0, h_iters = 0, h_status = 0;
    while (launched < o.max_iters) {
        const int batch = (o.max_iters - launched) < check ? (o.max_iters - launched) : check;
        for (int q = 0; q < batch; ++q) {
            const int it = launched + q;
            CgGate gate{H, S, I, it, alternate ? 1 + (it & 1) : 0, pdl ? 1 : 0};
            if (single) {
                double *u = A->w_u.p, *s = A->w_s.p, *w = ap;
                HEAT_TRY(launch_cg_fused_update(n, x, r, p, s, u, w, dinv, gate, H, I, A->partials.p, I + I_COUNTER2, vgrid, st));
                HEAT_TRY(spmv_halo(ctx, A, u, w, gate, &H[it + 1].delta));
                HEAT_TRY(comm_allreduce_sum(ctx, &H[it + 1].rz, 3));
            } else if (peer) {
                // 3 launches, no NCCL: halo = peer stores fused into update_p, dots = peer inboxes
                double *pbuf[2] = {A->w_p.p, A->w_p2.p};
                double *pin = pbuf[it & 1], *pout = pbuf[(it + 1) & 1];
                const unsigned long long s1 = ctx->peer_red_seq + 2ull * (unsigned)it + 1, s2 = s1 + 1;
                const PeerRed pr = peer_red_of(ctx);
                SpmvPeer sp;
                sp.on = true; sp.n_interior = A->n_int_slices; sp.halo = A->peer->halo;
                sp.halo.epoch = ctx->peer_halo_epoch + 1 + (unsigned)it;
                sp.red = pr; sp.seq_out = s1; sp.I = I;
                const int g = spmv_grid(A->n_slices, sms);
                DotOut d{A->partials.p, 0, g, I + I_COUNTER, S + S_TMP2};
                HEAT_TRY(launch_spmv_peer(A, pin, ap, gate, d, sp, g, st));
                HEAT_TRY(launch_cg_update_xr_peer(n, x, r, pin, ap, dinv, gate, H, S, I, A->partials.p, I + I_COUNTER2, pr, s1, s2, vgrid, st));
                PeerPush pp = A->peer->push[(it + 1) & 1];
                pp.epoch = ctx->peer_halo_epoch + 2 + (unsigned)it;
                HEAT_TRY(launch_cg_update_p_peer(n, pout, pin, r, dinv, nullptr, gate, H, I, pr, s2, pp, vgrid, st));
            } else if (cheb_fused) {
                // degree k: k SpMV launches + k + 1 vector launches, no NCCL (cg.cu)
                const int k = o.cheb_degree;
                const double ratio = o.cheb_ratio > 0 ? o.cheb_ratio : 30.0;
                const double c_alpha = lmax / ratio, c_beta = 1.1 * lmax;
                const double delta = 2.0 / (c_beta - c_alpha), theta = 0.5 * (c_beta + c_alpha), s1c = theta * delta;
                double *pbuf[2] = {A->w_p.p, A->w_p2.p}, *zbuf[2] = {A->w_u.p, A->w_u2.p};
                double *rbuf[2] = {A->w_r.p, A->w_r2.p}, *wbuf[2] = {A->w_w.p, A->w_w2.p};     // out-of-place r and W (cg.cu)
                double *pin = pbuf[it & 1], *pout = pbuf[(it + 1) & 1];
                const double *r_in = rbuf[it & 1];
                double *r_new = rbuf[(it + 1) & 1];
                const unsigned long long s1 = ctx->peer_red_seq + 2ull * (unsigned)it + 1, s2 = s1 + 1;
                const unsigned long long e0 = ctx->peer_halo_epoch + 1 + (unsigned long long)it * (unsigned)k;   // epoch of this iteration's first SpMV
                const PeerRed pr = peer_red_of(ctx);
                const int g = spmv_grid(A->n_slices, sms);
                SpmvPeer sp;
                sp.on = true; sp.n_interior = A->n_int_slices; sp.halo = A->peer->halo; sp.red = pr; sp.I = I;
                sp.halo.epoch = e0; sp.seq_out = s1;
                DotOut d{A->partials.p, 0, g, I + I_COUNTER, S + S_TMP2};
                HEAT_TRY(launch_spmv_peer(A, pin, ap, gate, d, sp, g, st));
                PeerPush pz = A->peer->push[2];
                pz.epoch = e0 + 1;
                HEAT_TRY(launch_cheb_xr_first_peer(k == 1, n, x, r_in, r_new, pin, ap, A->dinv.p, 1.0 / theta, wbuf[0], zbuf[0], gate, H, S, I,
                                                   A->partials.p, I + I_COUNTER2, pr, s1, s2, pz, vgrid, st));
                double rho = 1.0 / s1c;
                for (int j = 1; j < k; ++j) {
                    const double rho_new = 1.0 / (2.0 * s1c - rho);
                    double *zin = zbuf[(j - 1) & 1], *zout = zbuf[j & 1];
                    sp.halo.epoch = e0 + (unsigned)j;
                    DotOut nod{A->partials.p, 0, g, I + I_COUNTER, nullptr};
                    HEAT_TRY(launch_spmv_peer(A, zin, A->w_t.p, gate, nod, sp, g, st));
                    PeerPush pj = A->peer->push[2 + (j & 1)];
                    pj.epoch = e0 + (unsigned)j + 1;
                    HEAT_TRY(launch_cheb_step_peer(j == k - 1, n, A->dinv.p, r_new, A->w_t.p, rho_new * rho, 2.0 * rho_new * delta,
                                                   wbuf[(j - 1) & 1], wbuf[j & 1], zin, zout, gate, S, A->partials.p, I + I_COUNTER2, pr, s2, pj,
                                                   vgrid, st));
                    rho = rho_new;
                }
                PeerPush pp = A->peer->push[(it + 1) & 1];
                pp.epoch = e0 + (unsigned)k;
                HEAT_TRY(launch_cg_update_p_peer(n, pout, pin, r_new, A->dinv.p, zbuf[(k - 1) & 1], gate, H, I, pr, s2, pp, vgrid, st));
            } else if (!cheb) {
                HEAT_TRY(spmv_halo(ctx, A, p, ap, gate, S + S_PAP0));
                HEAT_TRY(comm_allreduce_sum(ctx, S + S_PAP0, 1));
                HEAT_TRY(launch_cg_update_xr(n, x, r, p, ap, dinv, gate, H, S, I, A->partials.p, I + I_COUNTER2, vgrid, st));
                HEAT_TRY(comm_allreduce_sum(ctx, &H[it + 1].rz, 3));
                HEAT_TRY(launch_cg_update_p(n, p, r, dinv, gate, vgrid, st));
            } else {
                double *z = A->w_u.p;
                HEAT_TRY(spmv_halo(ctx, A, p, ap, gate, S + S_PAP0));
                HEAT_TRY(comm_allreduce_sum(ctx, S + S_PAP0, 1));
                HEAT_TRY(launch_cg_xr_plain(n, x, r, p, ap, gate, S, I, vgrid, st));
                HEAT_TRY(gen_prec(lmax, r, z, gate, vgrid));
                HEAT_TRY(launch_cg_dots(n, r, z, gate, H, I, A->partials.p, I + I_COUNTER2, vgrid, st));
                HEAT_TRY(comm_allreduce_sum(ctx, &H[it + 1].rz, 3));
                HEAT_TRY(launch_cg_p_plain(n, p, z, gate, vgrid, st));
            }
        }
        launched += batch;
        int hI[2];
        HEAT_CUDA(cudaMemcpyAsync(hI, I, sizeof(int) * 2, cudaMemcpyDeviceToHost, st));
        HEAT_CUDA(cudaStreamSynchronize(st));
        h_iters = hI[I_ITERS]; h_status = hI[I_STATUS];
        if (on_poll && h_status == 0) {
            HEAT_TRY((*on_poll)(h_iters));          // x holds iterate h_iters (trajectory output)
            // the callback may keep ONE rank busy for seconds (rank 0 gathers and writes the Exodus frame) while the
            // others would launch the next batch and spin on its stamps until the peer-wait budget runs out: line the
            // ranks up again first (a stream-ordered 1-double all-reduce; the NCCL path simply waits in its collectives)
            if ((peer || cheb_fused) && ctx->nranks > 1) HEAT_TRY(comm_allreduce_sum(ctx, S + S_COUNT - 1, 1));
        }
        if (h_iters < launched) break;              // the stopping test fired (or breakdown): frozen
    }
    HEAT_CUDA(cudaEventRecord(ctx->ev_b, st));
    if (peer) {                                       // identical on every rank: same launches everywhere
        ctx->peer_red_seq += 2ull * (unsigned)o.max_iters + 4;
        ctx->peer_halo_epoch += (unsigned)o.max_iters + 4;
    }
    if (cheb_fused) {
        ctx->peer_red_seq += 2ull * (unsigned)o.max_iters + 4;
        ctx->peer_halo_epoch += (unsigned long long)o.max_iters * (unsigned)o.cheb_degree + 4;
    }
#ifdef HEAT_PEER_TRACE
    if (trace.p) {
        HEAT_CUDA(cudaStreamSynchronize(st));
        std::vector<unsigned long long> h((size_t)kTraceIters * 9);
        HEAT_CUDA(cudaMemcpy(h.data(), trace.p, sizeof(TraceBuf), cudaMemcpyDeviceToHost));
        HEAT_TRY(trace_set_cg(nullptr)); HEAT_TRY(trace_set_spmv(nullptr));
        FILE *fp = fopen((std::string(trace_file) + std::to_string(ctx->rank) + ".txt").c_str(), "w");
        if (fp) {
            const int nit = h_iters < kTraceIters ? h_iters : kTraceIters;
            for (int it = 0; it < nit; ++it) {
                for (int q = 0; q < 9; ++q) fprintf(fp, "%llu ", h[(size_t)it * 9 + q]);
                fprintf(fp, "\n");
            }
            fclose(fp);
        }
    }
#endif
    CgRec h0, hk;
    HEAT_CUDA(cudaMemcpyAsync(&h0, H, sizeof(CgRec), cudaMemcpyDeviceToHost, st));
    HEAT_CUDA(cudaMemcpyAsync(&hk, H + h_iters, sizeof(CgRec), cudaMemcpyDeviceToHost, st));
    HEAT_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f;
    HEAT_CUDA(cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b));
    if (info) {
        info->iters = h_iters;
        info->r0_norm = sqrt(h0.rr);
        info->achieved_tol = h0.rr > 0.0 ? sqrt(hk.rr / h0.rr) : 0.0;
        info->converged = (hk.rr <= o.tol * o.tol * h0.rr) ? 1 : 0;
        info->solve_ms = ms;
    }
    if (h_status == 3) HEAT_FAIL(51, "heat_solve: peer-memory communication timed out at iteration %d (a rank left the solve?)", h_iters);
    if (h_status == 2) HEAT_FAIL(50, "heat_solve: CG breakdown (p.Ap <= 0) at iteration %d — matrix not SPD?", h_iters);
    return 0;
}

}  // namespace heat
