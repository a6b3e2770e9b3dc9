// gmres.cu — restarted, right-preconditioned GMRES(m): the reference's literal Krylov method
// (Belos "GMRES" with setRightPrec, BelosMueLuSolver.cpp:102-109; the reference runs it with
// "Maximum Iterations" = 1 inside a restart loop, :113-133 — see belosSolver() in
// include/ExodusIO_b200.hpp for that loop).
//
//   cycle:  r = b - A x ; beta = ||r|| ; v_0 = r / beta
//           j = 0..m-1:  w = A M^-1 v_j ; two passes of classical Gram-Schmidt against v_0..v_j (Belos'
//                        ICGS); h_{j+1,j} = ||w|| ; v_{j+1} = w / h_{j+1,j} ; Givens update of the
//                        least-squares problem ; implicit residual |g_{j+1}| / ||r_0|| tested against tol
//           x += M^-1 (V y)
//
// Vector work is fused into three kernels: multi-dot (w against up to 8 basis vectors per pass over
// w), multi-axpy (w -= V h, optionally with ||w||^2 in the same pass) and combine (u = V y).  The
// (m+1) x m Hessenberg matrix, the rotations and the triangular solve are scalar work kept on the
// host, as in Belos; that costs one small device->host copy per iteration.
#include <cmath>

#include "comm.cuh"
#include "device_utils.cuh"
#include "kernels.cuh"
#include "solve.cuh"

namespace heat {

constexpr int kDotTile = 8;

// out[t] = sum_i V[(j0+t)*stride + i] * w[i], t < cnt <= kDotTile
__global__ void __launch_bounds__(kBlock) gmres_mdot_kernel(int64_t n, const double *__restrict__ V, int64_t stride, int j0,
                                                            int cnt, const double *__restrict__ w, double *out,
                                                            double *partials, int *counter) {
    double acc[kDotTile];
#pragma unroll
    for (int t = 0; t < kDotTile; ++t) acc[t] = 0.0;
    const int64_t step = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += step) {
        const double wv = w[i];
#pragma unroll
        for (int t = 0; t < kDotTile; ++t)
            if (t < cnt) acc[t] += V[(int64_t)(j0 + t) * stride + i] * wv;
    }
    // all kDotTile slots are written (zeros beyond cnt): the caller's buffer has kDotTile entries of slack
    double *const oo[kDotTile] = {out, out + 1, out + 2, out + 3, out + 4, out + 5, out + 6, out + 7};
    grid_sum<kDotTile>(acc, partials, 0, gridDim.x, counter, oo);
}

// w -= sum_{t<cnt} h[t] V[t]; if nrm2 != nullptr also reduces ||w_new||^2
template <bool NORM>
__global__ void __launch_bounds__(kBlock) gmres_maxpy_kernel(int64_t n, const double *__restrict__ V, int64_t stride, int cnt,
                                                             const double *__restrict__ h, double *__restrict__ w,
                                                             double *nrm2, double *partials, int *counter) {
    double acc[1] = {0.0};
    const int64_t step = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += step) {
        double s = w[i];
        for (int t = 0; t < cnt; ++t) s = fma(-h[t], V[(int64_t)t * stride + i], s);
        w[i] = s;
        if (NORM) acc[0] += s * s;
    }
    if (NORM) {
        double *const o[1] = {nrm2};
        grid_sum<1>(acc, partials, 0, gridDim.x, counter, o);
    }
}

// u = sum_{t<cnt} y[t] V[t]
__global__ void __launch_bounds__(kBlock) gmres_combine_kernel(int64_t n, const double *__restrict__ V, int64_t stride, int cnt,
                                                               const double *__restrict__ y, double *__restrict__ u) {
    const int64_t step = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += step) {
        double s = 0.0;
        for (int t = 0; t < cnt; ++t) s = fma(y[t], V[(int64_t)t * stride + i], s);
        u[i] = s;
    }
}

// z = d .* v  (Jacobi) or z = v (d == nullptr)
__global__ void __launch_bounds__(kBlock) diag_scale_kernel(int64_t n, const double *__restrict__ d, const double *__restrict__ v,
                                                            double *__restrict__ z) {
    const int64_t step = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += step) z[i] = d ? d[i] * v[i] : v[i];
}

// z = M^-1 v   (z: SpMV input, has ghost room)
static int prec_apply(heat_ctx *ctx, heat_matrix *A, const heat_solve_opts &o, double lmax, const double *v, double *z, int grid) {
    CgGate nogate{nullptr, nullptr, nullptr, 0};
    switch (o.prec) {
        case HEAT_PREC_CHEBYSHEV: return cheb_apply(ctx, A, o, lmax, v, z, nogate, grid);
        case HEAT_PREC_ILU0: return ilu_apply(ctx, A, v, z);
        default:
            diag_scale_kernel<<<grid, kBlock, 0, ctx->stream>>>(A->n_owned, o.prec == HEAT_PREC_JACOBI ? A->dinv.p : nullptr, v, z);
            HEAT_LAUNCHED();
            return 0;
    }
}

int gmres_device(heat_ctx *ctx, heat_matrix *A, double *x, const double *b, const heat_solve_opts &o,
                 heat_solve_info *info, const std::function<int(int)> *on_poll) {
    const int64_t n = A->n_owned;
    const size_t nv = (size_t)(A->n_owned + A->n_ghost);
    int m = o.gmres_restart > 0 ? o.gmres_restart : 300;           // Belos "Num Blocks" (default 300)
    if (o.max_iters > 0 && m > o.max_iters) m = o.max_iters;
    if (m < 1) m = 1;
    cudaStream_t st = ctx->stream;
    const int grid = vec_grid(n, sm_count(ctx->device));
    HEAT_TRY(ensure_workspace(A, false, o.prec == HEAT_PREC_CHEBYSHEV));
    if (!A->w_u.p) HEAT_TRY(A->w_u.alloc(nv));
    if (o.prec == HEAT_PREC_ILU0) HEAT_TRY(ilu0_setup(ctx, A));
    double lmax = o.cheb_lambda_max;
    if (o.prec == HEAT_PREC_CHEBYSHEV && !(lmax > 0.0)) {
        if (!(A->cheb_lmax_est > 0.0)) HEAT_TRY(estimate_lambda_max(ctx, A, &A->cheb_lmax_est));
        lmax = A->cheb_lmax_est;
    }

    const int64_t stride = (n + 1) & ~(int64_t)1;
    if ((double)stride * (double)(m + 1) * 8.0 > 150e9)
        HEAT_FAIL(2, "GMRES: %d basis vectors of %lld rows do not fit the GPU; lower gmres_restart", m + 1, (long long)n);
    DevBuf<double> V, hdev, partials, ubuf;
    DevBuf<int> counter;
    DevBuf<CgRec> rec;
    HEAT_TRY(V.alloc((size_t)stride * (size_t)(m + 1)));
    HEAT_TRY(hdev.alloc((size_t)(2 * (m + 1) + 2 + kDotTile)));
    HEAT_TRY(partials.alloc((size_t)kDotTile * kMaxPartials));
    HEAT_TRY(counter.alloc(1));
    HEAT_TRY(rec.alloc(1));
    HEAT_TRY(ubuf.alloc((size_t)n));
    HEAT_CUDA(cudaMemsetAsync(counter.p, 0, sizeof(int), st));
    HEAT_CUDA(cudaMemsetAsync(rec.p, 0, sizeof(CgRec), st));
    double *h1 = hdev.p, *h2 = hdev.p + (m + 1), *nrm2 = hdev.p + 2 * (m + 1);
    double *r = A->w_r.p, *w = A->w_ap.p, *z = A->w_u.p, *u = ubuf.p;
    CgGate nogate{nullptr, nullptr, nullptr, 0};

    std::vector<double> H((size_t)(m + 1) * (size_t)m, 0.0), cs((size_t)m), sn((size_t)m), g((size_t)m + 1), y((size_t)m);
    std::vector<double> hh((size_t)(2 * (m + 1) + 2));
    double r0norm = -1.0, resid = 0.0;
    int iters = 0;
    bool converged = false, stagnated = false;
    HEAT_CUDA(cudaEventRecord(ctx->ev_a, st));
    bool first_cycle = true;
    while (true) {
        // ---- explicit residual of the current iterate ----
        HEAT_TRY(spmv_halo(ctx, A, x, w, nogate, nullptr));
        if (first_cycle && ctx->wait_before_rhs) HEAT_CUDA(cudaStreamWaitEvent(st, ctx->wait_before_rhs, 0));
        first_cycle = false;
        HEAT_TRY(launch_cg_init(n, b, w, A->dinv.p, r, u, rec.p, partials.p, counter.p, grid, st));   // r = b - A x ; rec.rr = r.r
        HEAT_TRY(comm_allreduce_sum(ctx, &rec.p->rz, 3));
        CgRec hrec;
        HEAT_CUDA(cudaMemcpyAsync(&hrec, rec.p, sizeof(CgRec), cudaMemcpyDeviceToHost, st));
        HEAT_CUDA(cudaStreamSynchronize(st));
        const double beta = sqrt(hrec.rr);
        if (r0norm < 0.0) r0norm = beta;
        resid = beta;
        if (!(beta > o.tol * r0norm) || iters >= o.max_iters) { converged = !(beta > o.tol * r0norm); break; }
        // v_0 = r / beta
        HEAT_TRY(launch_pm_scale(n, r, &rec.p->rr, V.p, grid, st));
        std::fill(g.begin(), g.end(), 0.0);
        g[0] = beta;
        int k = 0;                                           // columns built in this cycle
        for (int j = 0; j < m && iters < o.max_iters; ++j) {
            HEAT_TRY(prec_apply(ctx, A, o, lmax, V.p + (int64_t)j * stride, z, grid));
            HEAT_TRY(spmv_halo(ctx, A, z, w, nogate, nullptr));
            for (int pass = 0; pass < 2; ++pass) {           // iterated classical Gram-Schmidt (ICGS, 2 passes)
                double *hp = pass == 0 ? h1 : h2;
                for (int j0 = 0; j0 <= j; j0 += kDotTile) {
                    const int cnt = (j + 1 - j0) < kDotTile ? (j + 1 - j0) : kDotTile;
                    gmres_mdot_kernel<<<grid, kBlock, 0, st>>>(n, V.p, stride, j0, cnt, w, hp + j0, partials.p, counter.p);
                    HEAT_LAUNCHED();
                }
                HEAT_TRY(comm_allreduce_sum(ctx, hp, j + 1));
                if (pass == 0) gmres_maxpy_kernel<false><<<grid, kBlock, 0, st>>>(n, V.p, stride, j + 1, hp, w, nullptr, partials.p, counter.p);
                else gmres_maxpy_kernel<true><<<grid, kBlock, 0, st>>>(n, V.p, stride, j + 1, hp, w, nrm2, partials.p, counter.p);
                HEAT_LAUNCHED();
            }
            HEAT_TRY(comm_allreduce_sum(ctx, nrm2, 1));
            HEAT_CUDA(cudaMemcpyAsync(hh.data(), hdev.p, sizeof(double) * (size_t)(2 * (m + 1) + 1), cudaMemcpyDeviceToHost, st));
            HEAT_CUDA(cudaStreamSynchronize(st));
            const double hnext = sqrt(hh[(size_t)(2 * (m + 1))]);
            double *Hj = H.data() + (size_t)j * (size_t)(m + 1);
            for (int t = 0; t <= j; ++t) Hj[t] = hh[(size_t)t] + hh[(size_t)(m + 1 + t)];
            Hj[j + 1] = hnext;
            for (int t = 0; t < j; ++t) {                    // previous rotations
                const double a = cs[(size_t)t] * Hj[t] + sn[(size_t)t] * Hj[t + 1];
                Hj[t + 1] = -sn[(size_t)t] * Hj[t] + cs[(size_t)t] * Hj[t + 1];
                Hj[t] = a;
            }
            const double den = hypot(Hj[j], Hj[j + 1]);
            if (!(den > 0.0)) { stagnated = true; break; }   // A M^-1 v_j lies in span(V) and is zero: exact breakdown
            cs[(size_t)j] = Hj[j] / den; sn[(size_t)j] = Hj[j + 1] / den;
            Hj[j] = den; Hj[j + 1] = 0.0;
            g[(size_t)j + 1] = -sn[(size_t)j] * g[(size_t)j];
            g[(size_t)j] = cs[(size_t)j] * g[(size_t)j];
            ++iters; k = j + 1;
            resid = fabs(g[(size_t)j + 1]);                  // Belos' implicit residual norm
            if (!(resid > o.tol * r0norm) || !(hnext > 0.0)) break;
            HEAT_TRY(launch_pm_scale(n, w, nrm2, V.p + (int64_t)(j + 1) * stride, grid, st));      // v_{j+1}
        }
        if (k > 0) {
            for (int i = k - 1; i >= 0; --i) {               // R y = g
                double s = g[(size_t)i];
                for (int t = i + 1; t < k; ++t) s -= H[(size_t)t * (size_t)(m + 1) + (size_t)i] * y[(size_t)t];
                y[(size_t)i] = s / H[(size_t)i * (size_t)(m + 1) + (size_t)i];
            }
            HEAT_CUDA(cudaMemcpyAsync(h1, y.data(), sizeof(double) * (size_t)k, cudaMemcpyHostToDevice, st));
            gmres_combine_kernel<<<grid, kBlock, 0, st>>>(n, V.p, stride, k, h1, u);
            HEAT_LAUNCHED();
            HEAT_TRY(prec_apply(ctx, A, o, lmax, u, z, grid));
            HEAT_TRY(launch_axpby(n, 1.0, z, 1.0, x, grid, st));                                   // x += M^-1 V y
            HEAT_CUDA(cudaStreamSynchronize(st));            // y (host) must outlive the copy
        }
        if (on_poll) HEAT_TRY((*on_poll)(iters));
        if (stagnated || k == 0) break;
        if (!(resid > o.tol * r0norm)) {                     // implicit test fired: Belos reports convergence
            converged = true;
            break;
        }
        if (iters >= o.max_iters) break;
    }
    HEAT_CUDA(cudaEventRecord(ctx->ev_b, st));
    HEAT_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f;
    HEAT_CUDA(cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b));
    if (info) {
        info->iters = iters;
        info->r0_norm = r0norm < 0.0 ? 0.0 : r0norm;
        info->achieved_tol = r0norm > 0.0 ? resid / r0norm : 0.0;
        info->converged = converged ? 1 : 0;
        info->solve_ms = ms;
    }
    return 0;
}

}  // namespace heat
