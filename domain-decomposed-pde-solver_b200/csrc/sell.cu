// sell.cu — local CSR -> SELL-C (C = 64, 2 rows per lane) conversion, diagonal extraction and the
// interior / boundary slice split used to overlap the halo exchange with the interior SpMV.
// (Role of Tpetra::CrsMatrix::fillComplete's local-matrix + Import set-up, ExodusIO.hpp:609.)
#include <cstdlib>

#include <cub/cub.cuh>

#include "device_utils.cuh"
#include "kernels.cuh"
#include "sell_dict.cuh"

namespace heat {

// one thread per slice: width = longest row of the slice
__global__ void sell_width_kernel(const int64_t *__restrict__ row_ptr, int64_t n_rows, int64_t n_slices,
                                  int64_t *__restrict__ slice_entries) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slices) return;
    const int64_t r0 = s * kSellChunk;
    const int64_t r1 = (r0 + kSellChunk < n_rows) ? r0 + kSellChunk : n_rows;
    int64_t w = 0;
    for (int64_t r = r0; r < r1; ++r) {
        int64_t len = row_ptr[r + 1] - row_ptr[r];
        w = len > w ? len : w;
    }
    slice_entries[s] = w * kSellChunk;
}

// one warp per slice: lane owns rows 2*lane, 2*lane+1; padding = (value 0, column = own row).  Rows past
// the end of the matrix (tail of the last slice) point 64 rows back — a valid column at a CONSTANT
// col-row offset, so the tail does not add 60 distinct offsets to the last slice's table (sell_dict_kernel)
__global__ void __launch_bounds__(kBlock)
sell_fill_kernel(const int64_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                 const double *__restrict__ val, int64_t n_rows, int64_t n_owned_cols, int64_t n_slices,
                 const int64_t *__restrict__ slice_ptr, int32_t *__restrict__ scol,
                 double *__restrict__ sval, int32_t *__restrict__ slice_is_boundary) {
    const int lane = threadIdx.x & 31;
    const int64_t s = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (s >= n_slices) return;
    const int64_t base = slice_ptr[s];
    const int w = (int)((slice_ptr[s + 1] - base) >> 6);
    const int64_t row0 = s * kSellChunk + 2 * lane, row1 = row0 + 1;
    int64_t b0 = 0, l0 = 0, b1 = 0, l1 = 0;
    if (row0 < n_rows) { b0 = row_ptr[row0]; l0 = row_ptr[row0 + 1] - b0; }
    if (row1 < n_rows) { b1 = row_ptr[row1]; l1 = row_ptr[row1 + 1] - b1; }
    const int32_t pad0 = row0 < n_rows ? (int32_t)row0 : (row0 >= kSellChunk ? (int32_t)(row0 - kSellChunk) : 0);
    const int32_t pad1 = row1 < n_rows ? (int32_t)row1 : (row1 >= kSellChunk ? (int32_t)(row1 - kSellChunk) : 0);
    int ghost = 0;
    for (int k = 0; k < w; ++k) {
        int32_t c0 = pad0, c1 = pad1;
        double v0 = 0.0, v1 = 0.0;
        if (k < l0) { c0 = col[b0 + k]; v0 = val[b0 + k]; }
        if (k < l1) { c1 = col[b1 + k]; v1 = val[b1 + k]; }
        ghost |= (c0 >= n_owned_cols) | (c1 >= n_owned_cols);
        const int64_t o = base + (int64_t)k * kSellChunk + 2 * lane;
        *reinterpret_cast<int2 *>(scol + o) = make_int2(c0, c1);
        *reinterpret_cast<double2 *>(sval + o) = make_double2(v0, v1);
    }
    ghost = __any_sync(0xffffffffu, ghost);
    if (lane == 0 && slice_is_boundary) slice_is_boundary[s] = ghost;
}

// packed per-slice metadata in processing order (list == nullptr: natural order).  cptr / mode == nullptr: the column
// stream has `cbytes` bytes per entry in every slice (1: all table-indexed, 4: int32 ids) at cbytes * entry offset.
__global__ void slice_meta_kernel(const int64_t *__restrict__ slice_ptr, const int64_t *__restrict__ cptr,
                                  const uint8_t *__restrict__ mode, int cbytes, const int32_t *__restrict__ list, int64_t n_list,
                                  SliceMeta *__restrict__ meta) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_list) return;
    const int64_t s = list ? (int64_t)list[t] : t;
    const int64_t b = slice_ptr[s];
    const int32_t w = (int32_t)((slice_ptr[s + 1] - b) >> 6);
    const int m = mode ? mode[s] : (cbytes == 1 ? kColModeU8 : kColModeI32);
    const int64_t cb = cptr ? cptr[s] : b * cbytes;
    meta[t] = SliceMeta{(uint32_t)(b >> 6), (uint32_t)(cb >> 6), w | (m << 24), (int32_t)s};
}

// Compact column stream, chosen per slice.  One warp per slice collects the distinct (col - row) offsets of the
// slice's 64 x w entries (first-seen order: k ascending, then lane, then the lane's first row) in a shared-memory
// table.  On a structured mesh a slice has as many distinct offsets as the stencil has directions (15 on the Kuhn
// cube, a few more where ghost columns start), so the 4-byte column id of every entry shrinks to a 1-byte table
// index: 12 -> 9 bytes per stored entry.  A slice with more than kSellDictCap distinct offsets (any unstructured
// numbering) stores int16 deltas col - row instead, 12 -> 10 bytes, whenever they fit — every mesh whose bandwidth is
// below 32768; only if some slice fits neither does the matrix keep the int32 stream.
// PASS 1 classifies: mode[s], table length, column bytes of the slice.  PASS 2 writes the stream and the tables.
__device__ __forceinline__ bool fits_i16(int v) { return v >= -32768 && v <= 32767; }

__global__ void __launch_bounds__(kBlock)
sell_classify_kernel(const int64_t *__restrict__ slice_ptr, const int32_t *__restrict__ scol, int64_t n_slices, int force16,
                     uint8_t *__restrict__ mode, int64_t *__restrict__ cbytes, int *__restrict__ max_tab, int *__restrict__ any_mode) {
    __shared__ int32_t tab_s[kWarpsPerBlock][kSellDictCap];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t s = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
    if (s >= n_slices) return;
    int32_t *tab = tab_s[warp];
    const int64_t base = slice_ptr[s];
    const int w = (int)((slice_ptr[s + 1] - base) >> 6);
    const int row0 = (int)(s * kSellChunk) + 2 * lane;
    int T = 0;
    bool table_ok = true, i16_ok = true;
    for (int k = 0; k < w; ++k) {
        const int2 c = *reinterpret_cast<const int2 *>(scol + base + (int64_t)k * kSellChunk + 2 * lane);
        const int off[2] = {c.x - row0, c.y - (row0 + 1)};
        i16_ok = i16_ok && fits_i16(off[0]) && fits_i16(off[1]);
        int id[2];
        if (table_ok) table_ok = sell_dict_step(tab, T, off, id);
    }
    i16_ok = __all_sync(0xffffffffu, i16_ok);
    const int m = (table_ok && !(force16 && i16_ok)) ? kColModeU8 : i16_ok ? kColModeI16 : kColModeI32;
    if (lane == 0) {
        mode[s] = (uint8_t)m;
        cbytes[s] = (int64_t)w * kSellChunk * (m == kColModeU8 ? 1 : m == kColModeI16 ? 2 : 4);
        if (m == kColModeU8) atomicMax(max_tab, T);
        atomicOr(any_mode, 1 << m);
    }
}

__global__ void __launch_bounds__(kBlock)
sell_compact_kernel(const int64_t *__restrict__ slice_ptr, const int32_t *__restrict__ scol, int64_t n_slices,
                    const uint8_t *__restrict__ mode, const int64_t *__restrict__ cptr, int tpad,
                    uint8_t *__restrict__ cstream, int32_t *__restrict__ tabs) {
    __shared__ int32_t tab_s[kWarpsPerBlock][kSellDictCap];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t s = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
    if (s >= n_slices) return;
    int32_t *tab = tab_s[warp];
    const int64_t base = slice_ptr[s];
    const int w = (int)((slice_ptr[s + 1] - base) >> 6);
    const int row0 = (int)(s * kSellChunk) + 2 * lane;
    const int m = mode ? mode[s] : kColModeU8;
    uint8_t *out = cstream + (cptr ? cptr[s] : base);
    int T = 0;
    for (int k = 0; k < w; ++k) {
        const int2 c = *reinterpret_cast<const int2 *>(scol + base + (int64_t)k * kSellChunk + 2 * lane);
        const int off[2] = {c.x - row0, c.y - (row0 + 1)};
        if (m == kColModeU8) {
            int id[2];
            sell_dict_step(tab, T, off, id);
            *reinterpret_cast<uchar2 *>(out + (int64_t)k * kSellChunk + 2 * lane) = make_uchar2((unsigned char)id[0], (unsigned char)id[1]);
        } else {
            *reinterpret_cast<short2 *>(out + 2 * ((int64_t)k * kSellChunk + 2 * lane)) = make_short2((short)off[0], (short)off[1]);
        }
    }
    if (tabs) {
        __syncwarp();
        for (int t = lane; t < tpad; t += 32) tabs[s * tpad + t] = (m == kColModeU8 && t < T) ? tab[t] : 0;
    }
}

// HEAT_SPMV_CIDX: 0 = keep int32 columns only; 1 (default) = compact where possible; 2 = prefer int16 deltas over
// tables (tests / measurements of the int16 path on structured meshes)
int sell_cidx_mode() {
    const char *e = getenv("HEAT_SPMV_CIDX");
    const int v = e ? atoi(e) : 1;
    return v < 0 || v > 2 ? 1 : v;
}

// builds the compact column stream (A->sell_idx8, sell_tab, slice_cptr, slice_mode, sell_cmode); leaves the int32
// stream in charge (sell_cmode = 4) when some slice fits neither a table nor int16 deltas
static int sell_build_dict(heat_matrix *A, cudaStream_t st) {
    const int64_t ns = A->n_slices;
    A->sell_tpad = 0; A->sell_cmode = 4;
    A->sell_idx8.release(); A->sell_tab.release(); A->slice_cptr.release(); A->slice_mode.release();
    const int want = sell_cidx_mode();
    if (ns == 0 || A->sell_padded == 0 || want == 0 || !spmv_compact_supported()) return 0;
    DevBuf<int> d_flags;                              // [0] longest table, [1] bit m set: some slice has mode m
    HEAT_TRY(d_flags.alloc(2));
    HEAT_CUDA(cudaMemsetAsync(d_flags.p, 0, 2 * sizeof(int), st));
    DevBuf<int64_t> cbytes;
    HEAT_TRY(cbytes.alloc((size_t)ns));
    HEAT_TRY(A->slice_mode.alloc((size_t)ns));
    const unsigned grid = (unsigned)((ns + kWarpsPerBlock - 1) / kWarpsPerBlock);
    sell_classify_kernel<<<grid, kBlock, 0, st>>>(A->slice_ptr.p, A->sell_col.p, ns, want == 2, A->slice_mode.p, cbytes.p, d_flags.p, d_flags.p + 1);
    HEAT_LAUNCHED();
    int h[2] = {0, 0};
    HEAT_CUDA(cudaMemcpyAsync(h, d_flags.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    HEAT_CUDA(cudaStreamSynchronize(st));
    if (h[1] & (1 << kColModeI32)) { A->slice_mode.release(); return 0; }       // some slice needs int32 ids: keep sell_col
    const bool mixed = (h[1] & (1 << kColModeI16)) != 0;
    const int tpad = h[0] < 4 ? 4 : (h[0] + 3) & ~3;   // 16-byte granules for the TMA copy of a table
    int64_t total = A->sell_padded;
    if (mixed) {
        HEAT_TRY(A->slice_cptr.alloc((size_t)ns + 1));
        HEAT_CUDA(cudaMemsetAsync(A->slice_cptr.p, 0, sizeof(int64_t), st));
        size_t tb = 0;
        HEAT_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tb, cbytes.p, A->slice_cptr.p + 1, ns, st));
        DevBuf<char> tmp; HEAT_TRY(tmp.alloc(tb));
        HEAT_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, tb, cbytes.p, A->slice_cptr.p + 1, ns, st));
        HEAT_CUDA(cudaMemcpyAsync(&total, A->slice_cptr.p + ns, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        HEAT_CUDA(cudaStreamSynchronize(st));
    } else {
        A->slice_mode.release();                       // every slice is table-indexed: the stream is idx8[entry]
    }
    HEAT_TRY(A->sell_idx8.alloc((size_t)total));
    HEAT_TRY(A->sell_tab.alloc((size_t)ns * (size_t)tpad));
    sell_compact_kernel<<<grid, kBlock, 0, st>>>(A->slice_ptr.p, A->sell_col.p, ns, mixed ? A->slice_mode.p : nullptr,
                                                 mixed ? A->slice_cptr.p : nullptr, tpad, A->sell_idx8.p, A->sell_tab.p);
    HEAT_LAUNCHED();
    A->sell_tpad = tpad;
    A->sell_cmode = mixed ? 2 : 1;
    return 0;
}

__global__ void extract_diag_kernel(const int64_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                                    const double *__restrict__ val, int64_t n_rows, double *__restrict__ diag,
                                    double *__restrict__ dinv) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    double d = 0.0;
    for (int64_t q = row_ptr[r]; q < row_ptr[r + 1]; ++q)
        if (col[q] == r) d = val[q];
    diag[r] = d;
    dinv[r] = 1.0 / d;
}

int launch_extract_diag(const heat_matrix *A, cudaStream_t st) {
    if (A->n_owned == 0) return 0;
    const int64_t n = A->n_owned;
    extract_diag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(A->row_ptr.p, A->col.p, A->val.p, n,
                                                                    A->diag.p, A->dinv.p);
    HEAT_LAUNCHED();
    return 0;
}

// interior / boundary slice lists (h_flags[s] != 0: slice s references a ghost column; null: no ghosts) and the
// packed per-slice metadata in processing order
int sell_finish_lists(heat_matrix *A, const int32_t *h_flags, cudaStream_t st) {
    const int64_t ns = A->n_slices;
    A->n_int_slices = ns;
    A->n_bnd_slices = 0;
    if (h_flags && ns > 0) {
        std::vector<int32_t> li, lb;
        for (int64_t s = 0; s < ns; ++s) (h_flags[(size_t)s] ? lb : li).push_back((int32_t)s);
        A->n_int_slices = (int64_t)li.size();
        A->n_bnd_slices = (int64_t)lb.size();
        HEAT_TRY(A->slices_interior.alloc(li.size()));
        HEAT_TRY(A->slices_boundary.alloc(lb.size()));
        std::vector<int32_t> all(li);
        all.insert(all.end(), lb.begin(), lb.end());
        HEAT_TRY(A->slices_all.alloc(all.size()));
        if (!all.empty())
            HEAT_CUDA(cudaMemcpyAsync(A->slices_all.p, all.data(), sizeof(int32_t) * all.size(), cudaMemcpyHostToDevice, st));
        if (!li.empty())
            HEAT_CUDA(cudaMemcpyAsync(A->slices_interior.p, li.data(), sizeof(int32_t) * li.size(), cudaMemcpyHostToDevice, st));
        if (!lb.empty())
            HEAT_CUDA(cudaMemcpyAsync(A->slices_boundary.p, lb.data(), sizeof(int32_t) * lb.size(), cudaMemcpyHostToDevice, st));
        HEAT_CUDA(cudaStreamSynchronize(st));          // the host vectors die here
    }
    return 0;
}

// packed per-slice metadata of the SpMV, in processing order; needs the column stream (sell_cmode, slice_cptr, slice_mode)
int sell_build_meta(heat_matrix *A, cudaStream_t st) {
    const int64_t ns = A->n_slices;
    HEAT_TRY(A->slice_meta.alloc((size_t)ns));
    if (ns > 0) {
        slice_meta_kernel<<<(unsigned)((ns + 255) / 256), 256, 0, st>>>(A->slice_ptr.p, A->slice_cptr.p, A->slice_mode.p,
                                                                      A->sell_cmode == 1 ? 1 : 4, A->n_ghost > 0 ? A->slices_all.p : nullptr,
                                                                      ns, A->slice_meta.p);
        HEAT_LAUNCHED();
    }
    return 0;
}

// ---- SELL -> CSR (export / ILU only): matrices assembled straight into the SpMV format have no CSR until asked ----
__global__ void rowlen_widen_kernel(const uint8_t *__restrict__ len8, int64_t n, int64_t *__restrict__ len64) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r <= n) len64[r] = r < n ? (int64_t)len8[r] : 0;
}
__global__ void __launch_bounds__(kBlock)
sell_to_csr_kernel(const int64_t *__restrict__ slice_ptr, const double *__restrict__ sval, const int32_t *__restrict__ scol,
                   const uint8_t *__restrict__ idx8, const int32_t *__restrict__ tabs, int tpad, const uint8_t *__restrict__ rowlen,
                   int64_t n_rows, int64_t n_slices, const int64_t *__restrict__ row_ptr, int32_t *__restrict__ col,
                   double *__restrict__ val) {
    // (only matrices assembled straight into SELL come here: their stream is all table-indexed or all int32)
    const int lane = threadIdx.x & 31;
    const int64_t s = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (s >= n_slices) return;
    const int64_t base = slice_ptr[s];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int64_t row = s * kSellChunk + 2 * lane + h;
        if (row >= n_rows) continue;
        const int len = rowlen[row];
        const int64_t o = row_ptr[row];
        for (int k = 0; k < len; ++k) {
            const int64_t e = base + (int64_t)k * kSellChunk + 2 * lane + h;
            col[o + k] = idx8 ? (int32_t)row + tabs[s * tpad + idx8[e]] : scol[e];
            val[o + k] = sval[e];
        }
    }
}
int sell_to_csr(heat_matrix *A, cudaStream_t st) {
    if (A->row_ptr.p) return 0;
    if (!A->sell_rowlen.p || !A->sell_val.p) HEAT_FAIL(60, "matrix has neither a CSR nor row lengths to rebuild it from");
    const int64_t n = A->n_owned, ns = A->n_slices;
    HEAT_TRY(A->row_ptr.alloc((size_t)n + 1));
    {
        DevBuf<int64_t> len64; HEAT_TRY(len64.alloc((size_t)n + 1));
        rowlen_widen_kernel<<<(unsigned)((n + 256) / 256), 256, 0, st>>>(A->sell_rowlen.p, n, len64.p);
        HEAT_LAUNCHED();
        size_t tb = 0;
        HEAT_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, len64.p, A->row_ptr.p, n + 1, st));
        DevBuf<char> tmp; HEAT_TRY(tmp.alloc(tb));
        HEAT_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, len64.p, A->row_ptr.p, n + 1, st));
        HEAT_CUDA(cudaStreamSynchronize(st));
    }
    HEAT_TRY(A->col.alloc((size_t)A->nnz)); HEAT_TRY(A->val.alloc((size_t)A->nnz));
    if (ns > 0) {
        sell_to_csr_kernel<<<(unsigned)((ns + kWarpsPerBlock - 1) / kWarpsPerBlock), kBlock, 0, st>>>(
            A->slice_ptr.p, A->sell_val.p, A->sell_col.p, A->sell_idx8.p, A->sell_tab.p, A->sell_tpad, A->sell_rowlen.p, n, ns,
            A->row_ptr.p, A->col.p, A->val.p);
        HEAT_LAUNCHED();
    }
    HEAT_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int sell_from_csr(heat_matrix *A, cudaStream_t st) {
    const int64_t n = A->n_owned;
    A->n_slices = (n + kSellChunk - 1) / kSellChunk;
    const int64_t ns = A->n_slices;
    HEAT_TRY(A->slice_ptr.alloc((size_t)ns + 1));
    HEAT_TRY(A->diag.alloc((size_t)n));
    HEAT_TRY(A->dinv.alloc((size_t)n));
    HEAT_CUDA(cudaMemsetAsync(A->slice_ptr.p, 0, sizeof(int64_t) * (size_t)(ns + 1), st));
    if (ns > 0) {
        DevBuf<int64_t> entries;
        HEAT_TRY(entries.alloc((size_t)ns));
        sell_width_kernel<<<(unsigned)((ns + 255) / 256), 256, 0, st>>>(A->row_ptr.p, n, ns, entries.p);
        HEAT_LAUNCHED();
        size_t tmp_bytes = 0;
        HEAT_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, entries.p, A->slice_ptr.p + 1, ns, st));
        DevBuf<char> tmp;
        HEAT_TRY(tmp.alloc(tmp_bytes));
        HEAT_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, tmp_bytes, entries.p, A->slice_ptr.p + 1, ns, st));
        HEAT_CUDA(cudaMemcpyAsync(&A->sell_padded, A->slice_ptr.p + ns, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        HEAT_CUDA(cudaStreamSynchronize(st));
        HEAT_TRY(A->sell_col.alloc((size_t)A->sell_padded));
        HEAT_TRY(A->sell_val.alloc((size_t)A->sell_padded));
        DevBuf<int32_t> flags;
        const bool need_split = A->n_ghost > 0;
        if (need_split) HEAT_TRY(flags.alloc((size_t)ns));
        sell_fill_kernel<<<(unsigned)((ns + kWarpsPerBlock - 1) / kWarpsPerBlock), kBlock, 0, st>>>(
            A->row_ptr.p, A->col.p, A->val.p, n, A->n_owned, ns, A->slice_ptr.p, A->sell_col.p, A->sell_val.p,
            need_split ? flags.p : nullptr);
        HEAT_LAUNCHED();
        std::vector<int32_t> h;
        if (need_split) {
            h.resize((size_t)ns);
            HEAT_CUDA(cudaMemcpyAsync(h.data(), flags.p, sizeof(int32_t) * (size_t)ns, cudaMemcpyDeviceToHost, st));
            HEAT_CUDA(cudaStreamSynchronize(st));
        }
        HEAT_TRY(sell_finish_lists(A, need_split ? h.data() : nullptr, st));
    } else {
        HEAT_TRY(sell_finish_lists(A, nullptr, st));
    }
    HEAT_TRY(sell_build_dict(A, st));
    HEAT_TRY(sell_build_meta(A, st));
    HEAT_TRY(launch_extract_diag(A, st));
    return 0;
}

}  // namespace heat
