// cg.cu — fused vector kernels of the preconditioned CG iteration (sm_100a, HBM-bound).
//
// Replaces what Belos + Tpetra::MultiVector do per iteration on the reference path
// (BelosMueLuSolver.cpp:114-133 -> MultiVector::update/dot/norm2 + Teuchos::reduceAll) with a
// few streaming kernels: 128-bit loads/stores, warp-shuffle + last-block deterministic
// reductions, scalars kept on the device (no host round trip inside an iteration).
//
// Classical PCG (HEAT_SOLVER_CG), iteration k (records H[k] = {gamma_k, -, rr_k, alpha_k}):
//   spmv+dot :  Ap = A p ; S[PAP] = p.Ap                                   (spmv.cu)
//   update_xr:  alpha = gamma_k / p.Ap ; x += alpha p ; r -= alpha Ap ; z = D^-1 r
//               H[k+1].rz = r.z ; H[k+1].rr = r.r ; iters = k+1
//   update_p :  beta = gamma_{k+1}/gamma_k ; p = D^-1 r + beta p
// Single-reduction PCG (HEAT_SOLVER_CG_SINGLE_REDUCE, Chronopoulos-Gear):
//   fused_update: beta, alpha from H[k], H[k-1] ; p = u + beta p ; s = w + beta s ;
//                 x += alpha p ; r -= alpha s ; u = D^-1 r ; H[k+1].{rz,rr}
//   spmv+dot    : w = A u ; H[k+1].delta = w.u     -> ONE all-reduce of {rz, delta, rr}
#include <mutex>

#include "device_utils.cuh"
#include "kernels.cuh"
#include "peer.cuh"

namespace heat {

__device__ __forceinline__ bool cg_done(const CgGate &g) {
    if (g.H == nullptr) return false;
    if (g.I[I_STATUS] != 0) return true;
    const double rr = g.H[g.it].rr, rr0 = g.H[0].rr;
    return !(rr > g.S[S_TOL2] * rr0);
}

int vec_grid(int64_t n, int sm_count) {
    int64_t blocks = (n / 2 + kBlock - 1) / kBlock;
    return grid_for(blocks, sm_count, 8);
}

// The streaming kernels loop grid-stride, so the grid should be exactly ONE resident wave: vec_grid assumes 8 blocks per
// SM, but update_xr holds 4 (64 registers) and update_p 6 — a 1184-block grid then ran as 592 + 592 or 888 + 296, and
// the partial second wave streamed with a third of the machine (visible at 1/8 of the problem: 150 us for 143 us of
// traffic).  Occupancy is asked once per kernel and device.
template <typename Kern>
static int one_wave(Kern kern, int grid) {
    struct Entry { const void *k; int dev; int cap; };
    static std::vector<Entry> cache;
    static std::mutex mu;                                // contexts on different GPUs may be driven by different threads
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return grid;
    std::lock_guard<std::mutex> lock(mu);
    for (const Entry &e : cache)
        if (e.k == (const void *)kern && e.dev == dev) return grid < e.cap ? grid : e.cap;
    int per_sm = 0, sms = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kBlock, 0) != cudaSuccess || per_sm < 1 ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) {
        cudaGetLastError();
        return grid;
    }
    cache.push_back(Entry{(const void *)kern, dev, per_sm * sms});
    return grid < per_sm * sms ? grid : per_sm * sms;
}

// -------------------------------------------------------------------------------------------------
// r = b - Ax ; z = D^-1 r ; q = z (p, or u of the single-reduce variant) ; H[0] = {r.z, 0, r.r}
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) cg_init_kernel(int64_t n, const double *__restrict__ b,
                                                         const double *__restrict__ ax,
                                                         const double *__restrict__ dinv,
                                                         double *__restrict__ r, double *__restrict__ q,
                                                         CgRec *H, double *partials, int *counter) {
    double acc[2] = {0.0, 0.0};
    const int64_t n2 = n >> 1, stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n2; i += stride) {
        double2 bv = ld_stream_f64x2(b + 2 * i), av = ld_stream_f64x2(ax + 2 * i), dv = ld_stream_f64x2(dinv + 2 * i);
        double2 rv = make_double2(bv.x - av.x, bv.y - av.y);
        double2 zv = make_double2(dv.x * rv.x, dv.y * rv.y);
        *reinterpret_cast<double2 *>(r + 2 * i) = rv;
        *reinterpret_cast<double2 *>(q + 2 * i) = zv;
        acc[0] += rv.x * zv.x + rv.y * zv.y;
        acc[1] += rv.x * rv.x + rv.y * rv.y;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const int64_t i = n - 1;
        double rv = b[i] - ax[i], zv = dinv[i] * rv;
        r[i] = rv; q[i] = zv;
        acc[0] += rv * zv; acc[1] += rv * rv;
    }
    double *const out[2] = {&H[0].rz, &H[0].rr};
    grid_sum<2>(acc, partials, 0, gridDim.x, counter, out);
}

int launch_cg_init(int64_t n, const double *b, const double *ax, const double *dinv, double *r,
                   double *q, CgRec *H, double *partials, int *counter, int grid, cudaStream_t st) {
    cg_init_kernel<<<grid, kBlock, 0, st>>>(n, b, ax, dinv, r, q, H, partials, counter);
    HEAT_LAUNCHED();
    return 0;
}

// -------------------------------------------------------------------------------------------------
// classical: x += alpha p ; r -= alpha Ap ; z = D^-1 r ; reduce r.z, r.r into H[it+1]
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) cg_update_xr_kernel(int64_t n, double *__restrict__ x,
                                                              double *__restrict__ r,
                                                              const double *__restrict__ p,
                                                              const double *__restrict__ ap,
                                                              const double *__restrict__ dinv, CgGate g,
                                                              CgRec *H, double *S, int *I,
                                                              double *partials, int *counter) {
    pdl_wait();
    pdl_launch_dependents();                 // every CTA of this grid has started: the next kernel's CTAs may move in as SMs free up
    if (cg_done(g)) return;
    const double pap = S[S_PAP0];
    if (!(pap > 0.0)) {                                   // Belos: "p.Ap <= 0" is a breakdown
        if (blockIdx.x == 0 && threadIdx.x == 0) I[I_STATUS] = 2;
        return;
    }
    const double alpha = H[g.it].rz / pap;
    double acc[2] = {0.0, 0.0};
    const int64_t n2 = n >> 1, stride = (int64_t)gridDim.x * kBlock;
    const bool back = xr_backward(g);
    for (int64_t t = (int64_t)blockIdx.x * kBlock + threadIdx.x; t < n2; t += stride) {
        const int64_t i = back ? n2 - 1 - t : t;
        double2 pv = ld_stream_f64x2(p + 2 * i), av = ld_stream_f64x2(ap + 2 * i), dv = ld_stream_f64x2(dinv + 2 * i);
        double2 xv = *reinterpret_cast<const double2 *>(x + 2 * i);
        double2 rv = *reinterpret_cast<const double2 *>(r + 2 * i);
        xv.x = fma(alpha, pv.x, xv.x); xv.y = fma(alpha, pv.y, xv.y);
        rv.x = fma(-alpha, av.x, rv.x); rv.y = fma(-alpha, av.y, rv.y);
        *reinterpret_cast<double2 *>(x + 2 * i) = xv;
        *reinterpret_cast<double2 *>(r + 2 * i) = rv;
        acc[0] += (dv.x * rv.x) * rv.x + (dv.y * rv.y) * rv.y;
        acc[1] += rv.x * rv.x + rv.y * rv.y;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const int64_t i = n - 1;
        double xv = fma(alpha, p[i], x[i]), rv = fma(-alpha, ap[i], r[i]);
        x[i] = xv; r[i] = rv;
        acc[0] += (dinv[i] * rv) * rv; acc[1] += rv * rv;
    }
    double *const out[2] = {&H[g.it + 1].rz, &H[g.it + 1].rr};
    if (grid_sum<2>(acc, partials, 0, gridDim.x, counter, out)) {
        H[g.it].alpha = alpha;
        I[I_ITERS] = g.it + 1;
    }
}

int launch_cg_update_xr(int64_t n, double *x, double *r, const double *p, const double *ap,
                        const double *dinv, CgGate gate, CgRec *H, double *S, int *I,
                        double *partials, int *counter, int grid, cudaStream_t st) {
    grid = one_wave(cg_update_xr_kernel, grid);
    HEAT_CUDA(launch_kernel(cg_update_xr_kernel, grid, kBlock, 0, st, gate.pdl != 0, n, x, r, p, ap, dinv, gate, H, S, I, partials, counter));
    HEAT_LAUNCHED();
    return 0;
}

// classical: p = D^-1 r + beta p, beta = gamma_{it+1} / gamma_it
__global__ void __launch_bounds__(kBlock) cg_update_p_kernel(int64_t n, double *__restrict__ p,
                                                             const double *__restrict__ r,
                                                             const double *__restrict__ dinv, CgGate g) {
    pdl_wait();
    pdl_launch_dependents();                 // every CTA of this grid has started: the next kernel's CTAs may move in as SMs free up
    if (cg_done(g)) return;
    CgGate nxt = g; nxt.it = g.it + 1;
    if (cg_done(nxt)) return;                              // converged: p is never used again
    const double beta = g.H[g.it + 1].rz / g.H[g.it].rz;
    const int64_t n2 = n >> 1, stride = (int64_t)gridDim.x * kBlock;
    const bool back = p_backward(g);
    for (int64_t t = (int64_t)blockIdx.x * kBlock + threadIdx.x; t < n2; t += stride) {
        const int64_t i = back ? n2 - 1 - t : t;
        double2 rv = ld_stream_f64x2(r + 2 * i), dv = ld_stream_f64x2(dinv + 2 * i);
        double2 pv = *reinterpret_cast<const double2 *>(p + 2 * i);
        pv.x = fma(beta, pv.x, dv.x * rv.x); pv.y = fma(beta, pv.y, dv.y * rv.y);
        *reinterpret_cast<double2 *>(p + 2 * i) = pv;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const int64_t i = n - 1;
        p[i] = fma(beta, p[i], dinv[i] * r[i]);
    }
}

int launch_cg_update_p(int64_t n, double *p, const double *r, const double *dinv, CgGate gate,
                       int grid, cudaStream_t st) {
    grid = one_wave(cg_update_p_kernel, grid);
    HEAT_CUDA(launch_kernel(cg_update_p_kernel, grid, kBlock, 0, st, gate.pdl != 0, n, p, (const double *)r, dinv, gate));
    HEAT_LAUNCHED();
    return 0;
}

// -------------------------------------------------------------------------------------------------
// peer-memory variants (peer.cuh): no NCCL inside the iteration
// -------------------------------------------------------------------------------------------------
// classical update_xr: alpha needs the GLOBAL p.Ap -> wait for reduction `seq_in`; the local sums
// r.z, r.r are published to every rank as reduction `seq_out` by the last block.
__global__ void __launch_bounds__(kBlock) cg_update_xr_peer_kernel(int64_t n, double *__restrict__ x,
                                                                   double *__restrict__ r,
                                                                   const double *__restrict__ p,
                                                                   const double *__restrict__ ap,
                                                                   const double *__restrict__ dinv, CgGate g,
                                                                   CgRec *H, double *S, int *I, double *partials,
                                                                   int *counter, PeerRed pr,
                                                                   unsigned long long seq_in,
                                                                   unsigned long long seq_out) {
    pdl_wait();
    pdl_launch_dependents();                 // every CTA of this grid has started: the next kernel's CTAs may move in as SMs free up
    if (cg_done(g)) return;
    if (threadIdx.x == 0) HEAT_TRACE_MIN(g.it, 1, 0);
    __shared__ double sh[2];
    if (threadIdx.x < 32) {
        double o[3];
        const bool ok = peer_red_wait_warp(pr, seq_in, o, I);
        if (threadIdx.x == 0) { sh[0] = o[0]; sh[1] = ok ? 1.0 : 0.0; }
    }
    __syncthreads();
    if (threadIdx.x == 0) HEAT_TRACE_MAX(g.it, 1, 1);
    if (sh[1] == 0.0) return;                              // communication timeout: I_STATUS = 3
    const double pap = sh[0];
    if (!(pap > 0.0)) {
        if (blockIdx.x == 0 && threadIdx.x == 0) I[I_STATUS] = 2;
        return;
    }
    const double alpha = H[g.it].rz / pap;
    double acc[2] = {0.0, 0.0};
    const int64_t n2 = n >> 1, stride = (int64_t)gridDim.x * kBlock;
    const bool back = xr_backward(g);
    for (int64_t t = (int64_t)blockIdx.x * kBlock + threadIdx.x; t < n2; t += stride) {
        const int64_t i = back ? n2 - 1 - t : t;
        double2 pv = ld_stream_f64x2(p + 2 * i), av = ld_stream_f64x2(ap + 2 * i), dv = ld_stream_f64x2(dinv + 2 * i);
        double2 xv = *reinterpret_cast<const double2 *>(x + 2 * i);
        double2 rv = *reinterpret_cast<const double2 *>(r + 2 * i);
        xv.x = fma(alpha, pv.x, xv.x); xv.y = fma(alpha, pv.y, xv.y);
        rv.x = fma(-alpha, av.x, rv.x); rv.y = fma(-alpha, av.y, rv.y);
        *reinterpret_cast<double2 *>(x + 2 * i) = xv;
        *reinterpret_cast<double2 *>(r + 2 * i) = rv;
        acc[0] += (dv.x * rv.x) * rv.x + (dv.y * rv.y) * rv.y;
        acc[1] += rv.x * rv.x + rv.y * rv.y;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const int64_t i = n - 1;
        double xv = fma(alpha, p[i], x[i]), rv = fma(-alpha, ap[i], r[i]);
        x[i] = xv; r[i] = rv;
        acc[0] += (dinv[i] * rv) * rv; acc[1] += rv * rv;
    }
    double *const out[2] = {S + S_TMP0, S + S_TMP1};
    if (grid_sum_block<2>(acc, partials, 0, gridDim.x, counter, out)) {
        if (threadIdx.x == 0) H[g.it].alpha = alpha;
        if (threadIdx.x < 32) peer_red_push_warp(pr, seq_out, S[S_TMP0], S[S_TMP1], 0.0);
    }
    if (threadIdx.x == 0) HEAT_TRACE_MAX(g.it, 1, 2);
}

int launch_cg_update_xr_peer(int64_t n, double *x, double *r, const double *p, const double *ap, const double *dinv,
                             CgGate gate, CgRec *H, double *S, int *I, double *partials, int *counter, PeerRed pr,
                             unsigned long long seq_in, unsigned long long seq_out, int grid, cudaStream_t st) {
    grid = one_wave(cg_update_xr_peer_kernel, grid);
    HEAT_CUDA(launch_kernel(cg_update_xr_peer_kernel, grid, kBlock, 0, st, gate.pdl != 0, n, x, r, p, ap, dinv, gate, H, S, I, partials, counter,
                            pr, seq_in, seq_out));
    HEAT_LAUNCHED();
    return 0;
}

// boundary entries of v = (compute(i)) go straight into the neighbours' ghost segments, then the
// epoch flags are raised by the last of the pushing blocks
template <typename F>
__device__ __forceinline__ void peer_push_phase(const PeerPush &push, F value_of) {
    if ((int)blockIdx.x >= push.n_blocks) return;
    const long long total = push.send_ptr[push.n_nbr];
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)push.n_blocks * blockDim.x) {
        int s = 0;
        while (t >= push.send_ptr[s + 1]) ++s;
        push.dst[s][t - push.send_ptr[s]] = value_of(push.send_idx[t]);
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int tk = atomicAdd(push.ticket, 1);
        if (tk == push.n_blocks - 1) {
            __threadfence_system();
            for (int s = 0; s < push.n_nbr; ++s) st_release_sys(push.flag[s], push.epoch);
            *push.ticket = 0;
        }
    }
}

// classical update_p: beta needs the GLOBAL r.z (reduction `seq_in`); block 0 records the global
// {r.z, r.r} in H[it+1]; p_out = D^-1 r + beta p_in (ping-pong buffers) and the halo push of p_out.
// z == nullptr: Jacobi fused (z = D^-1 r); else z holds M^-1 r of a general preconditioner (Chebyshev)
__global__ void __launch_bounds__(kBlock) cg_update_p_peer_kernel(int64_t n, double *p_out, const double *p_in,
                                                                  const double *__restrict__ r,
                                                                  const double *__restrict__ dinv,
                                                                  const double *__restrict__ z, CgGate g,
                                                                  CgRec *H, int *I, PeerRed pr,
                                                                  unsigned long long seq_in, PeerPush push) {
    pdl_wait();
    pdl_launch_dependents();                 // every CTA of this grid has started: the next kernel's CTAs may move in as SMs free up
    if (cg_done(g)) return;
    if (threadIdx.x == 0) HEAT_TRACE_MIN(g.it, 2, 0);
    __shared__ double sh[3];
    if (threadIdx.x < 32) {
        double o[3];
        const bool ok = peer_red_wait_warp(pr, seq_in, o, I);
        if (threadIdx.x == 0) { sh[0] = o[0]; sh[1] = o[1]; sh[2] = ok ? 1.0 : 0.0; }
    }
    __syncthreads();
    if (threadIdx.x == 0) HEAT_TRACE_MAX(g.it, 2, 1);
    if (sh[2] == 0.0) return;
    const double rz_new = sh[0], rr_new = sh[1];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        H[g.it + 1].rz = rz_new; H[g.it + 1].rr = rr_new;
        I[I_ITERS] = g.it + 1;
    }
    if (!(rr_new > g.S[S_TOL2] * g.H[0].rr)) return;        // converged: p is never used again
    const double beta = rz_new / g.H[g.it].rz;
    const int64_t n2 = n >> 1, stride = (int64_t)gridDim.x * kBlock;
    if (z) {
        peer_push_phase(push, [&](int32_t i) { return fma(beta, p_in[i], z[i]); });
        for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n2; i += stride) {
            double2 zv = ld_stream_f64x2(z + 2 * i);
            double2 pv = *reinterpret_cast<const double2 *>(p_in + 2 * i);
            pv.x = fma(beta, pv.x, zv.x); pv.y = fma(beta, pv.y, zv.y);
            *reinterpret_cast<double2 *>(p_out + 2 * i) = pv;
        }
        if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) p_out[n - 1] = fma(beta, p_in[n - 1], z[n - 1]);
        return;
    }
    peer_push_phase(push, [&](int32_t i) { return fma(beta, p_in[i], dinv[i] * r[i]); });
    const bool back = p_backward(g);
    for (int64_t t = (int64_t)blockIdx.x * kBlock + threadIdx.x; t < n2; t += stride) {
        const int64_t i = back ? n2 - 1 - t : t;
        double2 rv = ld_stream_f64x2(r + 2 * i), dv = ld_stream_f64x2(dinv + 2 * i);
        double2 pv = *reinterpret_cast<const double2 *>(p_in + 2 * i);
        pv.x = fma(beta, pv.x, dv.x * rv.x); pv.y = fma(beta, pv.y, dv.y * rv.y);
        *reinterpret_cast<double2 *>(p_out + 2 * i) = pv;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const int64_t i = n - 1;
        p_out[i] = fma(beta, p_in[i], dinv[i] * r[i]);
    }
    if (threadIdx.x == 0) HEAT_TRACE_MAX(g.it, 2, 2);
}

// -------------------------------------------------------------------------------------------------
// Chebyshev-preconditioned CG on the peer-memory path (also the fused single-GPU path, P == 1): the Ifpack2
// recurrences folded into the CG vector kernels.  Every kernel that PRODUCES an SpMV input
// (z of the polynomial, p of CG) also stores its boundary entries into the neighbours' ghost segments, and the last
// polynomial step carries the r.z / r.r reduction, so an iteration of degree k is k SpMV launches + k + 1 vector
// launches and no NCCL call (the NCCL path: 3k + 5 launches, 2 all-reduces, k send/recv groups).
//   xr_first : alpha = gamma / p.Ap (global, reduction seq_in) ; x += alpha p ; r -= alpha Ap ;
//              W = D^-1 r / theta ; Z = W                            [k == 1: + r.Z, r.r -> reduction seq_out]
//   step     : W = c1 W + c2 D^-1 (r - A Z) ; Z' = Z + W             [last step: + r.Z', r.r -> reduction seq_out]
// Z is ping-ponged between two buffers for the same reason p is: a neighbour may already deliver the boundary of
// the next polynomial iterate while this GPU still gathers the previous one.
// -------------------------------------------------------------------------------------------------
// r and W are written OUT OF PLACE (r_in -> r_out, w_in -> w_out, ping-ponged by the host like p and Z): the blocks
// that deliver the halo evaluate the boundary entries from the kernel's INPUTS while other blocks are already
// writing its outputs, so no array may be both.
template <bool LAST>
__global__ void __launch_bounds__(kBlock)
cheb_xr_first_peer_kernel(int64_t n, double *__restrict__ x, const double *__restrict__ r_in, double *__restrict__ r_out,
                          const double *__restrict__ p, const double *__restrict__ ap, const double *__restrict__ dinv,
                          double inv_theta, double *__restrict__ w, double *z_out, CgGate g, CgRec *H, double *S, int *I,
                          double *partials, int *counter, PeerRed pr, unsigned long long seq_in, unsigned long long seq_out,
                          PeerPush push) {
    if (cg_done(g)) return;
    __shared__ double sh[2];
    if (threadIdx.x < 32) {
        double o[3];
        const bool ok = peer_red_wait_warp(pr, seq_in, o, I);
        if (threadIdx.x == 0) { sh[0] = o[0]; sh[1] = ok ? 1.0 : 0.0; }
    }
    __syncthreads();
    if (sh[1] == 0.0) return;                              // communication timeout: I_STATUS = 3
    const double pap = sh[0];
    if (!(pap > 0.0)) {
        if (blockIdx.x == 0 && threadIdx.x == 0) I[I_STATUS] = 2;
        return;
    }
    const double alpha = H[g.it].rz / pap;
    if (!LAST) peer_push_phase(push, [&](int32_t i) { return dinv[i] * fma(-alpha, ap[i], r_in[i]) * inv_theta; });
    double acc[2] = {0.0, 0.0};
    const int64_t n2 = n >> 1, stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n2; i += stride) {
        double2 pv = ld_stream_f64x2(p + 2 * i), av = ld_stream_f64x2(ap + 2 * i), dv = ld_stream_f64x2(dinv + 2 * i);
        double2 xv = *reinterpret_cast<const double2 *>(x + 2 * i);
        double2 rv = ld_stream_f64x2(r_in + 2 * i);
        xv.x = fma(alpha, pv.x, xv.x); xv.y = fma(alpha, pv.y, xv.y);
        rv.x = fma(-alpha, av.x, rv.x); rv.y = fma(-alpha, av.y, rv.y);
        const double2 wv = make_double2(dv.x * rv.x * inv_theta, dv.y * rv.y * inv_theta);
        *reinterpret_cast<double2 *>(x + 2 * i) = xv;
        *reinterpret_cast<double2 *>(r_out + 2 * i) = rv;
        *reinterpret_cast<double2 *>(z_out + 2 * i) = wv;
        if (!LAST) *reinterpret_cast<double2 *>(w + 2 * i) = wv;
        if (LAST) { acc[0] += rv.x * wv.x + rv.y * wv.y; acc[1] += rv.x * rv.x + rv.y * rv.y; }
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const int64_t i = n - 1;
        const double xv = fma(alpha, p[i], x[i]), rv = fma(-alpha, ap[i], r_in[i]), wv = dinv[i] * rv * inv_theta;
        x[i] = xv; r_out[i] = rv; z_out[i] = wv;
        if (!LAST) w[i] = wv;
        if (LAST) { acc[0] += rv * wv; acc[1] += rv * rv; }
    }
    if (LAST) {
        double *const out[2] = {S + S_TMP0, S + S_TMP1};
        if (grid_sum_block<2>(acc, partials, 0, gridDim.x, counter, out)) {
            if (threadIdx.x == 0) H[g.it].alpha = alpha;
            if (threadIdx.x < 32) peer_red_push_warp(pr, seq_out, S[S_TMP0], S[S_TMP1], 0.0);
        }
    } else if (blockIdx.x == 0 && threadIdx.x == 0) {
        H[g.it].alpha = alpha;
    }
}

template <bool LAST>
__global__ void __launch_bounds__(kBlock)
cheb_step_peer_kernel(int64_t n, const double *__restrict__ dinv, const double *__restrict__ r, const double *__restrict__ az,
                      double c1, double c2, const double *__restrict__ w_in, double *__restrict__ w_out, const double *z_in,
                      double *z_out, CgGate g, double *S, double *partials, int *counter, PeerRed pr, unsigned long long seq_out,
                      PeerPush push) {
    if (cg_done(g)) return;
    if (!LAST) peer_push_phase(push, [&](int32_t i) { return z_in[i] + (c1 * w_in[i] + c2 * (dinv[i] * (r[i] - az[i]))); });
    double acc[2] = {0.0, 0.0};
    const int64_t n2 = n >> 1, stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n2; i += stride) {
        const double2 dv = ld_stream_f64x2(dinv + 2 * i), rv = ld_stream_f64x2(r + 2 * i), tv = ld_stream_f64x2(az + 2 * i);
        double2 wv = ld_stream_f64x2(w_in + 2 * i);
        double2 zv = *reinterpret_cast<const double2 *>(z_in + 2 * i);
        wv.x = c1 * wv.x + c2 * (dv.x * (rv.x - tv.x)); wv.y = c1 * wv.y + c2 * (dv.y * (rv.y - tv.y));
        zv.x += wv.x; zv.y += wv.y;
        if (!LAST) *reinterpret_cast<double2 *>(w_out + 2 * i) = wv;
        *reinterpret_cast<double2 *>(z_out + 2 * i) = zv;
        if (LAST) { acc[0] += rv.x * zv.x + rv.y * zv.y; acc[1] += rv.x * rv.x + rv.y * rv.y; }
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const int64_t i = n - 1;
        const double wv = c1 * w_in[i] + c2 * (dinv[i] * (r[i] - az[i])), zv = z_in[i] + wv;
        if (!LAST) w_out[i] = wv;
        z_out[i] = zv;
        if (LAST) { acc[0] += r[i] * zv; acc[1] += r[i] * r[i]; }
    }
    if (LAST) {
        double *const out[2] = {S + S_TMP0, S + S_TMP1};
        if (grid_sum_block<2>(acc, partials, 0, gridDim.x, counter, out) && threadIdx.x < 32)
            peer_red_push_warp(pr, seq_out, S[S_TMP0], S[S_TMP1], 0.0);
    }
}

int launch_cheb_xr_first_peer(bool last, int64_t n, double *x, const double *r_in, double *r_out, const double *p, const double *ap,
                              const double *dinv, double inv_theta, double *w, double *z_out, CgGate gate, CgRec *H, double *S, int *I,
                              double *partials, int *counter, PeerRed pr, unsigned long long seq_in, unsigned long long seq_out,
                              PeerPush push, int grid, cudaStream_t st) {
    grid = one_wave(last ? cheb_xr_first_peer_kernel<true> : cheb_xr_first_peer_kernel<false>, grid);
    if (push.n_blocks > grid) push.n_blocks = grid;
    if (last) cheb_xr_first_peer_kernel<true><<<grid, kBlock, 0, st>>>(n, x, r_in, r_out, p, ap, dinv, inv_theta, w, z_out, gate, H, S, I, partials, counter, pr, seq_in, seq_out, push);
    else cheb_xr_first_peer_kernel<false><<<grid, kBlock, 0, st>>>(n, x, r_in, r_out, p, ap, dinv, inv_theta, w, z_out, gate, H, S, I, partials, counter, pr, seq_in, seq_out, push);
    HEAT_LAUNCHED();
    return 0;
}
int launch_cheb_step_peer(bool last, int64_t n, const double *dinv, const double *r, const double *az, double c1, double c2,
                          const double *w_in, double *w_out, const double *z_in, double *z_out, CgGate gate, double *S, double *partials,
                          int *counter, PeerRed pr, unsigned long long seq_out, PeerPush push, int grid, cudaStream_t st) {
    grid = one_wave(last ? cheb_step_peer_kernel<true> : cheb_step_peer_kernel<false>, grid);
    if (push.n_blocks > grid) push.n_blocks = grid;
    if (last) cheb_step_peer_kernel<true><<<grid, kBlock, 0, st>>>(n, dinv, r, az, c1, c2, w_in, w_out, z_in, z_out, gate, S, partials, counter, pr, seq_out, push);
    else cheb_step_peer_kernel<false><<<grid, kBlock, 0, st>>>(n, dinv, r, az, c1, c2, w_in, w_out, z_in, z_out, gate, S, partials, counter, pr, seq_out, push);
    HEAT_LAUNCHED();
    return 0;
}

#ifdef HEAT_PEER_TRACE
int trace_set_cg(TraceBuf *buf) {
    HEAT_CUDA(cudaMemcpyToSymbol(g_trace, &buf, sizeof(buf)));
    return 0;
}
#endif

int launch_cg_update_p_peer(int64_t n, double *p_out, const double *p_in, const double *r, const double *dinv, const double *z,
                            CgGate gate, CgRec *H, int *I, PeerRed pr, unsigned long long seq_in, PeerPush push,
                            int grid, cudaStream_t st) {
    grid = one_wave(cg_update_p_peer_kernel, grid);
    if (push.n_blocks > grid) push.n_blocks = grid;
    HEAT_CUDA(launch_kernel(cg_update_p_peer_kernel, grid, kBlock, 0, st, gate.pdl != 0, n, p_out, p_in, r, dinv, z, gate, H, I, pr, seq_in, push));
    HEAT_LAUNCHED();
    return 0;
}

// stand-alone halo push of an existing vector (the first SpMV input of a solve)
__global__ void __launch_bounds__(kBlock) halo_push_kernel(const double *__restrict__ x, PeerPush push) {
    peer_push_phase(push, [&](int32_t i) { return x[i]; });
}
int launch_halo_push(const double *x, PeerPush push, cudaStream_t st) {
    if (push.n_nbr == 0) return 0;                 // single rank / no neighbour: nothing to deliver
    if (push.n_blocks < 1) push.n_blocks = 1;
    halo_push_kernel<<<push.n_blocks, kBlock, 0, st>>>(x, push);
    HEAT_LAUNCHED();
    return 0;
}

// -------------------------------------------------------------------------------------------------
// single-reduce (Chronopoulos-Gear): all vector work of an iteration in one pass
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) cg_fused_update_kernel(int64_t n, double *__restrict__ x,
                                                                 double *__restrict__ r,
                                                                 double *__restrict__ p,
                                                                 double *__restrict__ s,
                                                                 double *__restrict__ u,
                                                                 const double *__restrict__ w,
                                                                 const double *__restrict__ dinv,
                                                                 CgGate g, CgRec *H, int *I,
                                                                 double *partials, int *counter) {
    if (cg_done(g)) return;
    const double gamma = H[g.it].rz, delta = H[g.it].delta;
    double beta = 0.0, denom = delta;
    if (g.it > 0) {
        beta = gamma / H[g.it - 1].rz;
        denom = delta - beta * gamma / H[g.it - 1].alpha;
    }
    if (!(denom > 0.0)) {
        if (blockIdx.x == 0 && threadIdx.x == 0) I[I_STATUS] = 2;
        return;
    }
    const double alpha = gamma / denom;
    double acc[2] = {0.0, 0.0};
    const int64_t n2 = n >> 1, stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n2; i += stride) {
        double2 wv = ld_stream_f64x2(w + 2 * i), dv = ld_stream_f64x2(dinv + 2 * i);
        double2 uv = *reinterpret_cast<const double2 *>(u + 2 * i);
        double2 pv = *reinterpret_cast<const double2 *>(p + 2 * i);
        double2 sv = *reinterpret_cast<const double2 *>(s + 2 * i);
        double2 xv = *reinterpret_cast<const double2 *>(x + 2 * i);
        double2 rv = *reinterpret_cast<const double2 *>(r + 2 * i);
        pv.x = fma(beta, pv.x, uv.x); pv.y = fma(beta, pv.y, uv.y);
        sv.x = fma(beta, sv.x, wv.x); sv.y = fma(beta, sv.y, wv.y);
        xv.x = fma(alpha, pv.x, xv.x); xv.y = fma(alpha, pv.y, xv.y);
        rv.x = fma(-alpha, sv.x, rv.x); rv.y = fma(-alpha, sv.y, rv.y);
        uv.x = dv.x * rv.x; uv.y = dv.y * rv.y;
        *reinterpret_cast<double2 *>(p + 2 * i) = pv;
        *reinterpret_cast<double2 *>(s + 2 * i) = sv;
        *reinterpret_cast<double2 *>(x + 2 * i) = xv;
        *reinterpret_cast<double2 *>(r + 2 * i) = rv;
        *reinterpret_cast<double2 *>(u + 2 * i) = uv;
        acc[0] += rv.x * uv.x + rv.y * uv.y;
        acc[1] += rv.x * rv.x + rv.y * rv.y;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const int64_t i = n - 1;
        double pv = fma(beta, p[i], u[i]), sv = fma(beta, s[i], w[i]);
        double xv = fma(alpha, pv, x[i]), rv = fma(-alpha, sv, r[i]), uv = dinv[i] * rv;
        p[i] = pv; s[i] = sv; x[i] = xv; r[i] = rv; u[i] = uv;
        acc[0] += rv * uv; acc[1] += rv * rv;
    }
    double *const out[2] = {&H[g.it + 1].rz, &H[g.it + 1].rr};
    if (grid_sum<2>(acc, partials, 0, gridDim.x, counter, out)) {
        H[g.it].alpha = alpha;
        I[I_ITERS] = g.it + 1;
    }
}

int launch_cg_fused_update(int64_t n, double *x, double *r, double *p, double *s, double *u,
                           const double *w, const double *dinv, CgGate gate, CgRec *H, int *I,
                           double *partials, int *counter, int grid, cudaStream_t st) {
    cg_fused_update_kernel<<<grid, kBlock, 0, st>>>(n, x, r, p, s, u, w, dinv, gate, H, I, partials, counter);
    HEAT_LAUNCHED();
    return 0;
}

// -------------------------------------------------------------------------------------------------
// generic pieces (Chebyshev-preconditioned CG, residual checks)
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) axpby_kernel(int64_t n, double a, const double *__restrict__ x,
                                                       double b, double *__restrict__ y) {
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride)
        y[i] = (b == 0.0) ? a * x[i] : fma(a, x[i], b * y[i]);
}
int launch_axpby(int64_t n, double a, const double *x, double b, double *y, int grid, cudaStream_t st) {
    axpby_kernel<<<grid, kBlock, 0, st>>>(n, a, x, b, y);
    HEAT_LAUNCHED();
    return 0;
}

__global__ void __launch_bounds__(kBlock) dot2_kernel(int64_t n, const double *__restrict__ a,
                                                      const double *__restrict__ b,
                                                      const double *__restrict__ c,
                                                      const double *__restrict__ d, double *out_ab,
                                                      double *out_cd, double *partials, int *counter) {
    double acc[2] = {0.0, 0.0};
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) {
        acc[0] += a[i] * b[i];
        acc[1] += c[i] * d[i];
    }
    double *const out[2] = {out_ab, out_cd};
    grid_sum<2>(acc, partials, 0, gridDim.x, counter, out);
}
int launch_dot2(int64_t n, const double *a, const double *b, const double *c, const double *d,
                double *out_ab, double *out_cd, double *partials, int *counter, int grid, cudaStream_t st) {
    dot2_kernel<<<grid, kBlock, 0, st>>>(n, a, b, c, d, out_ab, out_cd, partials, counter);
    HEAT_LAUNCHED();
    return 0;
}

// -------------------------------------------------------------------------------------------------
// power method (ExodusMatrixTest.cpp:97-100): q = z / ||z||, the norm read from a device scalar
// (||z||^2 comes out of the previous SpMV's fused reduction, so no host round trip per iteration)
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) pm_scale_kernel(int64_t n, const double *__restrict__ z,
                                                          const double *__restrict__ zz, double *__restrict__ q) {
    const double inv = 1.0 / sqrt(*zz);
    const int64_t n2 = n >> 1, stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n2; i += stride) {
        const double2 v = ld_stream_f64x2(z + 2 * i);
        *reinterpret_cast<double2 *>(q + 2 * i) = make_double2(inv * v.x, inv * v.y);
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) q[n - 1] = inv * z[n - 1];
}
int launch_pm_scale(int64_t n, const double *z, const double *zz, double *q, int grid, cudaStream_t st) {
    pm_scale_kernel<<<grid, kBlock, 0, st>>>(n, z, zz, q);
    HEAT_LAUNCHED();
    return 0;
}
// ||z - lambda q||^2 (ExodusMatrixTest.cpp:104-105), lambda read from a device scalar
__global__ void __launch_bounds__(kBlock) pm_resid_kernel(int64_t n, const double *__restrict__ z,
                                                          const double *__restrict__ q, const double *__restrict__ lambda,
                                                          double *out, double *partials, int *counter) {
    const double lam = *lambda;
    double acc[1] = {0.0};
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) {
        const double d = z[i] - lam * q[i];
        acc[0] += d * d;
    }
    double *const o[1] = {out};
    grid_sum<1>(acc, partials, 0, gridDim.x, counter, o);
}
int launch_pm_resid(int64_t n, const double *z, const double *q, const double *lambda, double *out,
                    double *partials, int *counter, int grid, cudaStream_t st) {
    pm_resid_kernel<<<grid, kBlock, 0, st>>>(n, z, q, lambda, out, partials, counter);
    HEAT_LAUNCHED();
    return 0;
}

// Ifpack2 Chebyshev, zero start: W = D^-1 r / theta ; Z = W
__global__ void __launch_bounds__(kBlock) cheb_first_kernel(int64_t n, const double *__restrict__ dinv,
                                                            const double *__restrict__ r, double inv_theta,
                                                            double *__restrict__ w, double *__restrict__ z,
                                                            CgGate g) {
    if (cg_done(g)) return;
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) {
        double v = dinv[i] * r[i] * inv_theta;
        w[i] = v; z[i] = v;
    }
}
int launch_cheb_first(int64_t n, const double *dinv, const double *r, double inv_theta, double *w,
                      double *z, CgGate gate, int grid, cudaStream_t st) {
    cheb_first_kernel<<<grid, kBlock, 0, st>>>(n, dinv, r, inv_theta, w, z, gate);
    HEAT_LAUNCHED();
    return 0;
}
// W = c1 W + c2 D^-1 (r - A Z) ; Z += W
__global__ void __launch_bounds__(kBlock) cheb_step_kernel(int64_t n, const double *__restrict__ dinv,
                                                           const double *__restrict__ r,
                                                           const double *__restrict__ az, double c1, double c2,
                                                           double *__restrict__ w, double *__restrict__ z,
                                                           CgGate g) {
    if (cg_done(g)) return;
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) {
        double v = c1 * w[i] + c2 * (dinv[i] * (r[i] - az[i]));
        w[i] = v; z[i] += v;
    }
}
int launch_cheb_step(int64_t n, const double *dinv, const double *r, const double *az, double c1,
                     double c2, double *w, double *z, CgGate gate, int grid, cudaStream_t st) {
    cheb_step_kernel<<<grid, kBlock, 0, st>>>(n, dinv, r, az, c1, c2, w, z, gate);
    HEAT_LAUNCHED();
    return 0;
}

// x += alpha p ; r -= alpha Ap (no reduction; z comes from a separate preconditioner apply)
__global__ void __launch_bounds__(kBlock) cg_xr_plain_kernel(int64_t n, double *__restrict__ x,
                                                             double *__restrict__ r,
                                                             const double *__restrict__ p,
                                                             const double *__restrict__ ap, CgGate g,
                                                             double *S, int *I) {
    if (cg_done(g)) return;
    const double pap = S[S_PAP0];
    if (!(pap > 0.0)) {
        if (blockIdx.x == 0 && threadIdx.x == 0) I[I_STATUS] = 2;
        return;
    }
    const double alpha = g.H[g.it].rz / pap;
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) {
        x[i] = fma(alpha, p[i], x[i]);
        r[i] = fma(-alpha, ap[i], r[i]);
    }
}
int launch_cg_xr_plain(int64_t n, double *x, double *r, const double *p, const double *ap,
                       CgGate gate, double *S, int *I, int grid, cudaStream_t st) {
    cg_xr_plain_kernel<<<grid, kBlock, 0, st>>>(n, x, r, p, ap, gate, S, I);
    HEAT_LAUNCHED();
    return 0;
}

// H[it+1] = {r.z, -, r.r} ; iters = it+1   (gated: a frozen solve must not write records)
__global__ void __launch_bounds__(kBlock) cg_dots_kernel(int64_t n, const double *__restrict__ r,
                                                         const double *__restrict__ z, CgGate g, CgRec *H,
                                                         int *I, double *partials, int *counter) {
    if (cg_done(g)) return;
    double acc[2] = {0.0, 0.0};
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) {
        const double rv = r[i];
        acc[0] += rv * z[i];
        acc[1] += rv * rv;
    }
    double *const out[2] = {&H[g.it + 1].rz, &H[g.it + 1].rr};
    if (grid_sum<2>(acc, partials, 0, gridDim.x, counter, out)) I[I_ITERS] = g.it + 1;
}
int launch_cg_dots(int64_t n, const double *r, const double *z, CgGate gate, CgRec *H, int *I,
                   double *partials, int *counter, int grid, cudaStream_t st) {
    cg_dots_kernel<<<grid, kBlock, 0, st>>>(n, r, z, gate, H, I, partials, counter);
    HEAT_LAUNCHED();
    return 0;
}

// p = z + beta p
__global__ void __launch_bounds__(kBlock) cg_p_plain_kernel(int64_t n, double *__restrict__ p,
                                                            const double *__restrict__ z, CgGate g) {
    if (cg_done(g)) return;
    CgGate nxt = g; nxt.it = g.it + 1;
    if (cg_done(nxt)) return;
    const double beta = g.H[g.it + 1].rz / g.H[g.it].rz;
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) p[i] = fma(beta, p[i], z[i]);
}
int launch_cg_p_plain(int64_t n, double *p, const double *z, CgGate gate, int grid, cudaStream_t st) {
    cg_p_plain_kernel<<<grid, kBlock, 0, st>>>(n, p, z, gate);
    HEAT_LAUNCHED();
    return 0;
}

// -------------------------------------------------------------------------------------------------
__global__ void fill_kernel(int64_t n, double *x, double v) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] = v;
}
int launch_fill(int64_t n, double *x, double v, cudaStream_t st) {
    if (n <= 0) return 0;
    int grid = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
    fill_kernel<<<grid, 256, 0, st>>>(n, x, v);
    HEAT_LAUNCHED();
    return 0;
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// x[i] = U(-1,1) keyed on the global reduced id: GPU-count invariant input
__global__ void fill_hash_kernel(int64_t n, double *x, const int64_t *gids, int64_t gid0, uint64_t seed) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t g = (uint64_t)(gids ? gids[i] : gid0 + i);
        uint64_t h = splitmix64(g ^ (0x9E3779B97F4A7C15ull * seed));
        x[i] = (double)(h >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
    }
}
int launch_fill_hash(int64_t n, double *x, const int64_t *gids, int64_t gid0, uint64_t seed, cudaStream_t st) {
    if (n <= 0) return 0;
    int grid = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
    fill_hash_kernel<<<grid, 256, 0, st>>>(n, x, gids, gid0, seed);
    HEAT_LAUNCHED();
    return 0;
}

// halo pack: out[i] = x[idx[i]]
__global__ void gather_kernel(int64_t n, const double *__restrict__ x, const int32_t *__restrict__ idx,
                              double *__restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = x[idx[i]];
}
int launch_gather(int64_t n, const double *x, const int32_t *idx, double *out, cudaStream_t st) {
    if (n <= 0) return 0;
    int grid = (int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
    gather_kernel<<<grid, 256, 0, st>>>(n, x, idx, out);
    HEAT_LAUNCHED();
    return 0;
}

}  // namespace heat
