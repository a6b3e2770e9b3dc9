// spmv.cu — fp64 SELL-C SpMV for sm_100a (HBM-bound; no tensor cores: nothing here is a dense
// contraction).  Replaces Tpetra::CrsMatrix::apply (called inside Belos on the reference path;
// explicit use at ExodusMatrixTest.cpp:101).
//
// Format: SELL-C with C = 64 rows per slice = one warp, R = 2 adjacent rows per lane.  Entry k of
// the two rows of a lane is one 16-byte (val) + one 8-byte (col) load, consecutive lanes are
// consecutive in memory: every warp-level load is a fully coalesced 512 B / 256 B request.
// val/col are read exactly once -> ld.global.nc.L1::no_allocate keeps L1 for the x gather.
// Per-row accumulation runs left to right in CSR column order with fma() — bit-identical to
// the CPU oracle's oracle_spmv.  Optional fused dot sum_i y_i*x_i (p.Ap of CG) is reduced
// deterministically (device_utils.cuh: grid_sum).
//
// Algorithmic bytes per launch: 12*nnz + 16*n + 4*(n+1).
#include "device_utils.cuh"
#include "kernels.cuh"

namespace heat {

__device__ __forceinline__ bool gate_done(const CgGate &g) {
    if (g.H == nullptr) return false;
    if (g.I[I_STATUS] != 0) return true;
    const double rr = g.H[g.it].rr, rr0 = g.H[0].rr;
    return !(rr > g.S[S_TOL2] * rr0);
}

template <bool DOT>
__global__ void __launch_bounds__(kBlock)
sell_spmv_kernel(const int64_t *__restrict__ slice_ptr, const int32_t *__restrict__ col,
                 const double *__restrict__ val, const double *__restrict__ x, double *__restrict__ y,
                 int64_t n_rows, const int32_t *__restrict__ slice_list, int64_t n_list, CgGate gate,
                 DotOut dot) {
    if (gate_done(gate)) return;
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = (int64_t)gridDim.x * kWarpsPerBlock;
    double dsum = 0.0;
    for (int64_t t = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); t < n_list; t += warps_total) {
        const int64_t s = slice_list ? (int64_t)slice_list[t] : t;
        const int64_t base = slice_ptr[s];
        const int w = (int)((slice_ptr[s + 1] - base) >> 6);       // entries per row in this slice
        const double *vp = val + base + 2 * lane;
        const int32_t *cp = col + base + 2 * lane;
        double acc0 = 0.0, acc1 = 0.0;
#pragma unroll 4
        for (int k = 0; k < w; ++k) {
            const double2 v = ld_stream_f64x2(vp + (int64_t)k * kSellChunk);
            const int2 c = ld_stream_s32x2(cp + (int64_t)k * kSellChunk);
            acc0 = fma(v.x, __ldg(x + c.x), acc0);
            acc1 = fma(v.y, __ldg(x + c.y), acc1);
        }
        const int64_t row = s * kSellChunk + 2 * lane;
        if (row + 1 < n_rows) {
            *reinterpret_cast<double2 *>(y + row) = make_double2(acc0, acc1);
            if (DOT) dsum += acc0 * __ldg(x + row) + acc1 * __ldg(x + row + 1);
        } else if (row < n_rows) {
            y[row] = acc0;
            if (DOT) dsum += acc0 * __ldg(x + row);
        }
    }
    if (DOT) {
        double acc[1] = {dsum};
        double *const out[1] = {dot.out};
        grid_sum<1>(acc, dot.partials, dot.part_offset, dot.total_blocks, dot.counter, out);
    }
}

int spmv_grid(int64_t n_list, int sm_count) {
    int64_t blocks = (n_list + kWarpsPerBlock - 1) / kWarpsPerBlock;
    return grid_for(blocks, sm_count, 8);
}

int launch_spmv(const heat_matrix *A, const double *x, double *y, const int32_t *slice_list,
                int64_t n_list, CgGate gate, DotOut dot, int grid, cudaStream_t st) {
    if (n_list <= 0 && dot.out == nullptr) return 0;
    if (dot.out)
        sell_spmv_kernel<true><<<grid, kBlock, 0, st>>>(A->slice_ptr.p, A->sell_col.p, A->sell_val.p, x, y,
                                                       A->n_owned, slice_list, n_list, gate, dot);
    else
        sell_spmv_kernel<false><<<grid, kBlock, 0, st>>>(A->slice_ptr.p, A->sell_col.p, A->sell_val.p, x, y,
                                                        A->n_owned, slice_list, n_list, gate, dot);
    HEAT_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace heat
