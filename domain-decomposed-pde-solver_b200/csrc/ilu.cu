// ilu.cu — ILU(0) preconditioner: factorisation of the local diagonal block on the device and
// level-scheduled sparse triangular solves.
//
// Role on the reference path: Ifpack2 "ILUT" used as RIGHT preconditioner of Belos GMRES
// (BelosMueLuSolver.cpp:93-109).  Ifpack2's ILUT is a rank-local factorisation (no overlap), so with
// several ranks each GPU factors the block of its own rows and columns; ghost columns are ignored.
// Its default parameter list (level-of-fill 1, drop tolerance 0) keeps about as many entries per row
// as A has; ILU(0) — exactly the pattern of A — is the deterministic member of that family that can be
// pinned by an oracle (the Trilinos version the reference ran with is not recorded).
//
// Rows are grouped into dependency levels on the host (integer work on the pattern only); every level
// is one launch, one thread per row.  Arithmetic is contraction-free (this file is compiled with
// -fmad=false) and runs in the textbook IKJ order, so the factors match the CPU oracle bit for bit.
#include <algorithm>

#include "device_utils.cuh"
#include "kernels.cuh"
#include "solve.cuh"

namespace heat {

struct IluState {
    DevBuf<double> lu;                 // L (unit diagonal, strictly lower) and U on the pattern of the local CSR
    DevBuf<int64_t> dpos;              // position of the diagonal entry of every row
    DevBuf<int32_t> lrows, urows;      // rows ordered by level (forward / backward substitution)
    std::vector<int64_t> lptr, uptr;   // level offsets into lrows / urows
};

// row i of the current level: for k in lower(i) ascending: l_ik = a_ik / u_kk ; a_ij -= l_ik u_kj for j in upper(k)
__global__ void ilu0_factor_kernel(const int32_t *__restrict__ rows, int64_t n_level, const int64_t *__restrict__ rp,
                                   const int32_t *__restrict__ col, const int64_t *__restrict__ dpos, int64_t n_owned,
                                   double *__restrict__ lu) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_level) return;
    const int32_t i = rows[t];
    const int64_t b = rp[i], e = rp[i + 1];
    for (int64_t q = b; q < e; ++q) {
        const int32_t k = col[q];
        if (k >= i) continue;                              // diagonal, upper part or a ghost column
        const double lik = lu[q] / lu[dpos[k]];
        lu[q] = lik;
        for (int64_t s = rp[k]; s < rp[k + 1]; ++s) {
            const int32_t j = col[s];
            if (j <= k || j >= n_owned) continue;          // only u_kj, j > k, inside the local block
            for (int64_t p = b; p < e; ++p)
                if (col[p] == j) { lu[p] = lu[p] - lik * lu[s]; break; }
        }
    }
}

// forward substitution, unit lower factor: y_i = v_i - sum_{k<i} l_ik y_k
__global__ void ilu_lower_kernel(const int32_t *__restrict__ rows, int64_t n_level, const int64_t *__restrict__ rp,
                                 const int32_t *__restrict__ col, const double *__restrict__ lu,
                                 const double *__restrict__ v, double *__restrict__ y) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_level) return;
    const int32_t i = rows[t];
    double s = v[i];
    for (int64_t q = rp[i]; q < rp[i + 1]; ++q) {
        const int32_t k = col[q];
        if (k < i) s = s - lu[q] * y[k];
    }
    y[i] = s;
}

// backward substitution in place: z_i = (y_i - sum_{j>i} u_ij z_j) / u_ii
__global__ void ilu_upper_kernel(const int32_t *__restrict__ rows, int64_t n_level, const int64_t *__restrict__ rp,
                                 const int32_t *__restrict__ col, const int64_t *__restrict__ dpos, int64_t n_owned,
                                 const double *__restrict__ lu, double *__restrict__ z) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_level) return;
    const int32_t i = rows[t];
    double s = z[i];
    for (int64_t q = rp[i]; q < rp[i + 1]; ++q) {
        const int32_t j = col[q];
        if (j > i && j < n_owned) s = s - lu[q] * z[j];
    }
    z[i] = s / lu[dpos[i]];
}

void ilu_free(heat_matrix *A) {
    delete A->ilu;
    A->ilu = nullptr;
}

// builds A->ilu (once per matrix)
int ilu0_setup(heat_ctx *ctx, heat_matrix *A) {
    if (A->ilu) return 0;
    const int64_t n = A->n_owned, nnz = A->nnz;
    cudaStream_t st = ctx->stream;
    HEAT_TRY(sell_to_csr(A, st));                     // matrices assembled straight into SELL have no CSR yet
    std::vector<int64_t> rp((size_t)n + 1);
    std::vector<int32_t> col((size_t)nnz);
    HEAT_CUDA(cudaStreamSynchronize(st));
    HEAT_CUDA(cudaMemcpy(rp.data(), A->row_ptr.p, sizeof(int64_t) * rp.size(), cudaMemcpyDeviceToHost));
    if (nnz) HEAT_CUDA(cudaMemcpy(col.data(), A->col.p, sizeof(int32_t) * col.size(), cudaMemcpyDeviceToHost));
    // dependency levels of the forward (lower) and backward (upper) sweeps; owned columns of a row are ascending
    std::vector<int64_t> dpos((size_t)n, -1);
    std::vector<int32_t> levL((size_t)n, 0), levU((size_t)n, 0);
    int32_t maxL = 0, maxU = 0;
    for (int64_t i = 0; i < n; ++i) {
        int32_t l = 0;
        for (int64_t q = rp[(size_t)i]; q < rp[(size_t)i + 1]; ++q) {
            const int32_t c = col[(size_t)q];
            if (c == i) dpos[(size_t)i] = q;
            else if (c < i) l = std::max(l, levL[(size_t)c] + 1);
        }
        if (dpos[(size_t)i] < 0) HEAT_FAIL(52, "ILU(0): row %lld has no diagonal entry", (long long)i);
        levL[(size_t)i] = l; maxL = std::max(maxL, l);
    }
    for (int64_t i = n - 1; i >= 0; --i) {
        int32_t l = 0;
        for (int64_t q = rp[(size_t)i]; q < rp[(size_t)i + 1]; ++q) {
            const int32_t c = col[(size_t)q];
            if (c > i && c < n) l = std::max(l, levU[(size_t)c] + 1);
        }
        levU[(size_t)i] = l; maxU = std::max(maxU, l);
    }
    IluState *S = new IluState();
    auto bucket = [&](const std::vector<int32_t> &lev, int32_t maxl, std::vector<int64_t> &ptr, std::vector<int32_t> &rows) {
        ptr.assign((size_t)maxl + 2, 0);
        for (int64_t i = 0; i < n; ++i) ptr[(size_t)lev[(size_t)i] + 1]++;
        for (size_t l = 0; l + 1 < ptr.size(); ++l) ptr[l + 1] += ptr[l];
        rows.resize((size_t)n);
        std::vector<int64_t> fill(ptr.begin(), ptr.end() - 1);
        for (int64_t i = 0; i < n; ++i) rows[(size_t)fill[(size_t)lev[(size_t)i]]++] = (int32_t)i;
    };
    std::vector<int32_t> lrows, urows;
    if (n > 0) { bucket(levL, maxL, S->lptr, lrows); bucket(levU, maxU, S->uptr, urows); }
    else { S->lptr.assign(1, 0); S->uptr.assign(1, 0); }
    int rc = S->lu.alloc((size_t)nnz);
    if (!rc) rc = S->dpos.alloc((size_t)n);
    if (!rc) rc = S->lrows.alloc((size_t)n);
    if (!rc) rc = S->urows.alloc((size_t)n);
    if (rc) { delete S; return rc; }
    cudaError_t e = cudaSuccess;
    if (nnz) e = cudaMemcpyAsync(S->lu.p, A->val.p, sizeof(double) * (size_t)nnz, cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess && n) e = cudaMemcpyAsync(S->dpos.p, dpos.data(), sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && n) e = cudaMemcpyAsync(S->lrows.p, lrows.data(), sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && n) e = cudaMemcpyAsync(S->urows.p, urows.data(), sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { delete S; HEAT_FAIL(100, "ILU(0) set-up copy: %s", cudaGetErrorString(e)); }
    A->ilu = S;
    // factorisation: rows of forward level l only depend on rows of levels < l
    for (size_t l = 0; l + 1 < S->lptr.size(); ++l) {
        const int64_t cnt = S->lptr[l + 1] - S->lptr[l];
        if (cnt == 0) continue;
        ilu0_factor_kernel<<<(unsigned)((cnt + 127) / 128), 128, 0, st>>>(S->lrows.p + S->lptr[l], cnt, A->row_ptr.p, A->col.p,
                                                                         S->dpos.p, n, S->lu.p);
        HEAT_LAUNCHED();
    }
    HEAT_CUDA(cudaStreamSynchronize(st));               // the host vectors above must outlive the copies
    return 0;
}

// z = U^-1 L^-1 v on the local block (v, z: owned entries; may not alias)
int ilu_apply(heat_ctx *ctx, heat_matrix *A, const double *v, double *z) {
    IluState *S = A->ilu;
    if (!S) HEAT_FAIL(4, "ILU(0) apply before set-up");
    cudaStream_t st = ctx->stream;
    for (size_t l = 0; l + 1 < S->lptr.size(); ++l) {
        const int64_t cnt = S->lptr[l + 1] - S->lptr[l];
        if (cnt == 0) continue;
        ilu_lower_kernel<<<(unsigned)((cnt + 127) / 128), 128, 0, st>>>(S->lrows.p + S->lptr[l], cnt, A->row_ptr.p, A->col.p,
                                                                       S->lu.p, v, z);
        HEAT_LAUNCHED();
    }
    for (size_t l = 0; l + 1 < S->uptr.size(); ++l) {
        const int64_t cnt = S->uptr[l + 1] - S->uptr[l];
        if (cnt == 0) continue;
        ilu_upper_kernel<<<(unsigned)((cnt + 127) / 128), 128, 0, st>>>(S->urows.p + S->uptr[l], cnt, A->row_ptr.p, A->col.p,
                                                                       S->dpos.p, A->n_owned, S->lu.p, z);
        HEAT_LAUNCHED();
    }
    return 0;
}

int ilu_export(const heat_matrix *A, double *lu_host, int *n_levels_lower, int *n_levels_upper) {
    if (!A->ilu) HEAT_FAIL(4, "ILU(0) factors have not been computed (solve with HEAT_PREC_ILU0 first)");
    HEAT_CUDA(cudaStreamSynchronize(A->ctx->stream));
    if (lu_host && A->nnz) HEAT_CUDA(cudaMemcpy(lu_host, A->ilu->lu.p, sizeof(double) * (size_t)A->nnz, cudaMemcpyDeviceToHost));
    if (n_levels_lower) *n_levels_lower = (int)A->ilu->lptr.size() - 1;
    if (n_levels_upper) *n_levels_upper = (int)A->ilu->uptr.size() - 1;
    return 0;
}

}  // namespace heat
