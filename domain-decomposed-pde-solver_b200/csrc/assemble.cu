// assemble.cu — device assembly of the reduced heat system (sm_100a).
//
// Restates IO::assemble (ExodusIO.hpp:128-723) as data-parallel kernels:
//   :216-252  Dirichlet elimination + renumbering   -> flag + exclusive scan (red[])
//   :340-386  adjacency from element cliques        -> node->element lists by radix SORT (no atomics),
//                                                      then one thread per row merges its incident
//                                                      elements into a sorted unique neighbour list
//   :591-608  values (-1 / full degree)             -> written by the row's owner thread
//   :671-687  B = sum of Dirichlet neighbour values -> same pass
// plus the north-star P1 operator on the same pattern: the row owner recomputes the element
// gradients of each incident element (owner-computes gather: ATOMIC-FREE and bit-reproducible;
// accumulation order = ascending element id, identical to the CPU oracle).
// This file is compiled with -fmad=false so the element arithmetic matches the oracle bit for bit.
//
// Two front ends share the arithmetic: explicit connectivity (any Exodus mesh) and the analytic
// Kuhn-cube connectivity (BASELINE.json configs[2..4]; per-GPU slabs, nothing materialised).
#include <cub/cub.cuh>

#include "assemble.cuh"
#include "device_utils.cuh"

namespace heat {

constexpr int kMaxNbr = 128;   // longest neighbour list a row may have (explicit-mesh path)

// ---- P1 gradients (G = det * grad phi), s = 1/(6|det|) or 1/(2|det|) ------------------------------
__device__ __forceinline__ void tet_G(const double (&p)[4][3], double (&G)[4][3], double &s) {
    const double ax = p[1][0] - p[0][0], ay = p[1][1] - p[0][1], az = p[1][2] - p[0][2];
    const double bx = p[2][0] - p[0][0], by = p[2][1] - p[0][1], bz = p[2][2] - p[0][2];
    const double cx = p[3][0] - p[0][0], cy = p[3][1] - p[0][1], cz = p[3][2] - p[0][2];
    G[1][0] = by * cz - bz * cy; G[1][1] = bz * cx - bx * cz; G[1][2] = bx * cy - by * cx;
    G[2][0] = cy * az - cz * ay; G[2][1] = cz * ax - cx * az; G[2][2] = cx * ay - cy * ax;
    G[3][0] = ay * bz - az * by; G[3][1] = az * bx - ax * bz; G[3][2] = ax * by - ay * bx;
#pragma unroll
    for (int d = 0; d < 3; ++d) G[0][d] = -((G[1][d] + G[2][d]) + G[3][d]);
    const double det = (ax * G[1][0] + ay * G[1][1]) + az * G[1][2];
    s = 1.0 / (6.0 * fabs(det));
}
__device__ __forceinline__ void tri_G(const double (&p)[4][3], double (&G)[4][3], double &s) {
    const double ax = p[1][0] - p[0][0], ay = p[1][1] - p[0][1];
    const double bx = p[2][0] - p[0][0], by = p[2][1] - p[0][1];
    G[1][0] = by;  G[1][1] = -bx; G[1][2] = 0.0;
    G[2][0] = -ay; G[2][1] = ax;  G[2][2] = 0.0;
    G[0][0] = -(G[1][0] + G[2][0]); G[0][1] = -(G[1][1] + G[2][1]); G[0][2] = 0.0;
    G[3][0] = 0.0; G[3][1] = 0.0; G[3][2] = 0.0;
    const double det = ax * by - ay * bx;
    s = 1.0 / (2.0 * fabs(det));
}

// =================================================================================================
// explicit-connectivity path
// =================================================================================================
__global__ void dof_flag_kernel(int64_t N, const double *__restrict__ bc, int32_t *__restrict__ flag) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < N) flag[g] = isnan(bc[g]) ? 1 : 0;
}
// red[g] = reduced id or -1 ; red2orig[red] = g
__global__ void red_finish_kernel(int64_t N, const int32_t *__restrict__ flag, int32_t *__restrict__ red,
                                  int64_t *__restrict__ red2orig) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= N) return;
    if (flag[g]) red2orig[red[g]] = g; else red[g] = -1;
}
__global__ void n2e_keys_kernel(int64_t total, int npe, const int32_t *__restrict__ conn,
                                int32_t *__restrict__ keys, int32_t *__restrict__ elems) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= total) return;
    keys[q] = conn[q];
    elems[q] = (int32_t)(q / npe);
}
// n2e_ptr[g] = first position in the sorted key array with key >= g   (g = 0..N)
__global__ void n2e_ptr_kernel(int64_t N, int64_t total, const int32_t *__restrict__ keys,
                               int64_t *__restrict__ n2e_ptr) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g > N) return;
    int64_t lo = 0, hi = total;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (keys[mid] < g) lo = mid + 1; else hi = mid;
    }
    n2e_ptr[g] = lo;
}

// sorted unique neighbour nodes of g over its incident elements (ExodusIO.hpp:360-376)
__device__ __forceinline__ int collect_neighbours(int64_t g, const int64_t *__restrict__ n2e_ptr,
                                                  const int32_t *__restrict__ n2e, int npe,
                                                  const int32_t *__restrict__ conn, int32_t *buf, int *overflow) {
    int m = 0;
    for (int64_t q = n2e_ptr[g]; q < n2e_ptr[g + 1]; ++q) {
        const int32_t *e = conn + (int64_t)n2e[q] * npe;
        for (int k = 0; k < npe; ++k) {
            const int32_t v = e[k];
            if (v == g) continue;
            int lo = 0, hi = m;
            while (lo < hi) { int mid = (lo + hi) >> 1; if (buf[mid] < v) lo = mid + 1; else hi = mid; }
            if (lo < m && buf[lo] == v) continue;
            if (m == kMaxNbr) { *overflow = 1; continue; }
            for (int t = m; t > lo; --t) buf[t] = buf[t - 1];
            buf[lo] = v;
            ++m;
        }
    }
    return m;
}

// pass 1 over ALL reduced rows: row length = DOF neighbours + diagonal (FIXED D3: always a diagonal)
__global__ void __launch_bounds__(128)
pattern_count_kernel(int64_t n, const int64_t *__restrict__ red2orig, const int32_t *__restrict__ red,
                     const int64_t *__restrict__ n2e_ptr, const int32_t *__restrict__ n2e, int npe,
                     const int32_t *__restrict__ conn, int64_t *__restrict__ row_len, int *overflow) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t buf[kMaxNbr];
    const int u = collect_neighbours(red2orig[i], n2e_ptr, n2e, npe, conn, buf, overflow);
    int len = 1;
    for (int t = 0; t < u; ++t) len += (red[buf[t]] >= 0);
    row_len[i] = len;
}

// pass 2 over ALL reduced rows: global column ids, ascending, diagonal in place
__global__ void __launch_bounds__(128)
pattern_fill_kernel(int64_t n, const int64_t *__restrict__ red2orig, const int32_t *__restrict__ red,
                    const int64_t *__restrict__ n2e_ptr, const int32_t *__restrict__ n2e, int npe,
                    const int32_t *__restrict__ conn, const int64_t *__restrict__ row_ptr,
                    int32_t *__restrict__ col, int *overflow) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t buf[kMaxNbr];
    const int u = collect_neighbours(red2orig[i], n2e_ptr, n2e, npe, conn, buf, overflow);
    int32_t *c = col + row_ptr[i];
    int len = 0;
    bool placed = false;
    for (int t = 0; t < u; ++t) {
        const int32_t r = red[buf[t]];
        if (r < 0) continue;
        if (!placed && r > i) { c[len++] = (int32_t)i; placed = true; }
        c[len++] = r;
    }
    if (!placed) c[len++] = (int32_t)i;
}

// values + right-hand side of the OWNED rows.  lrow l <-> global reduced row gi = owned ? owned[l] : l.
// Local row l copies the global pattern row (ascending global ids), columns mapped through g2l.
__global__ void __launch_bounds__(128)
values_kernel(int64_t n_owned, const int32_t *__restrict__ owned, const int32_t *__restrict__ g2l,
              const int64_t *__restrict__ red2orig, const int32_t *__restrict__ red,
              const int64_t *__restrict__ n2e_ptr, const int32_t *__restrict__ n2e, int npe,
              const int32_t *__restrict__ conn, const double *__restrict__ X, const double *__restrict__ Y,
              const double *__restrict__ Z, const double *__restrict__ bc,
              const int64_t *__restrict__ grow_ptr, const int32_t *__restrict__ gcol,
              const int64_t *__restrict__ lrow_ptr, int32_t *__restrict__ lcol, double *__restrict__ lval,
              double *__restrict__ b, int mode, int *overflow) {
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_owned) return;
    const int64_t gi = owned ? (int64_t)owned[l] : l;
    const int64_t g = red2orig[gi];
    const int32_t *gc = gcol + grow_ptr[gi];
    const int len = (int)(grow_ptr[gi + 1] - grow_ptr[gi]);
    int32_t *lc = lcol ? lcol + lrow_ptr[l] : nullptr;
    double *lv = lval + lrow_ptr[l];
    int dpos = 0;
    for (int t = 0; t < len; ++t) {
        const int32_t c = gc[t];
        if (c == gi) dpos = t;
        if (lc) lc[t] = g2l ? g2l[c] : c;
    }
    double bsum = 0.0;
    if (mode == HEAT_OP_GRAPH_LAPLACIAN) {
        int32_t buf[kMaxNbr];
        const int u = collect_neighbours(g, n2e_ptr, n2e, npe, conn, buf, overflow);
        for (int t = 0; t < u; ++t)
            if (red[buf[t]] < 0) bsum += bc[buf[t]];               // :671-687
        for (int t = 0; t < len; ++t) lv[t] = -1.0;                // :601
        lv[dpos] = (double)u;                                      // :606 full degree
    } else {
        for (int t = 0; t < len; ++t) lv[t] = 0.0;
        for (int64_t q = n2e_ptr[g]; q < n2e_ptr[g + 1]; ++q) {
            const int32_t *e = conn + (int64_t)n2e[q] * npe;
            double p[4][3], G[4][3], s;
            int a = -1;
            for (int k = 0; k < npe; ++k) {
                const int32_t v = e[k];
                p[k][0] = X[v]; p[k][1] = Y[v]; p[k][2] = Z ? Z[v] : 0.0;
                if (v == g && a < 0) a = k;
            }
            if (npe == 4) tet_G(p, G, s); else tri_G(p, G, s);
            for (int k = 0; k < npe; ++k) {
                const double kab = ((G[a][0] * G[k][0] + G[a][1] * G[k][1]) + G[a][2] * G[k][2]) * s;
                const int32_t j = e[k];
                if (j == g) {
                    lv[dpos] += kab;
                } else if (red[j] >= 0) {
                    const int32_t r = red[j];
                    int lo = 0, hi = len - 1;
                    while (lo < hi) { int mid = (lo + hi) >> 1; if (gc[mid] < r) lo = mid + 1; else hi = mid; }
                    lv[lo] += kab;
                } else {
                    const double t2 = kab * bc[j];
                    bsum = bsum - t2;
                }
            }
        }
    }
    b[l] = bsum;
}

int GeneralAssembler::upload(const HostMesh &m, const std::vector<double> &node_bc, cudaStream_t st) {
    N = m.num_nodes; ne = m.num_elem; npe = m.npe; has_z = !m.z.empty();
    HEAT_TRY(x.alloc((size_t)N)); HEAT_TRY(y.alloc((size_t)N)); HEAT_TRY(z.alloc((size_t)N));
    HEAT_TRY(conn.alloc((size_t)(ne * npe))); HEAT_TRY(bc.alloc((size_t)N));
    HEAT_CUDA(cudaMemcpyAsync(x.p, m.x.data(), sizeof(double) * (size_t)N, cudaMemcpyHostToDevice, st));
    HEAT_CUDA(cudaMemcpyAsync(y.p, m.y.data(), sizeof(double) * (size_t)N, cudaMemcpyHostToDevice, st));
    if (has_z) HEAT_CUDA(cudaMemcpyAsync(z.p, m.z.data(), sizeof(double) * (size_t)N, cudaMemcpyHostToDevice, st));
    HEAT_CUDA(cudaMemcpyAsync(conn.p, m.conn.data(), sizeof(int32_t) * (size_t)(ne * npe), cudaMemcpyHostToDevice, st));
    HEAT_CUDA(cudaMemcpyAsync(bc.p, node_bc.data(), sizeof(double) * (size_t)N, cudaMemcpyHostToDevice, st));
    return 0;
}

// Kuhn cube materialised on the device (same numbering as oracle_cube_mesh)
__global__ void cube_nodes_kernel(int nx, int ny, int nz, double *__restrict__ X, double *__restrict__ Y,
                                  double *__restrict__ Z, double *__restrict__ bc) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t N = (int64_t)nx * ny * nz;
    if (g >= N) return;
    const int i = (int)(g % nx), j = (int)((g / nx) % ny), k = (int)(g / ((int64_t)nx * ny));
    X[g] = -5.0 + 10.0 * (double)i / (double)(nx - 1);
    Y[g] = -5.0 + 10.0 * (double)j / (double)(ny - 1);
    Z[g] = -5.0 + 10.0 * (double)k / (double)(nz - 1);
    bc[g] = (i == 0) ? 1000.0 : (i == nx - 1) ? 100.0 : nan("");
}
__constant__ int c_perms[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
__global__ void cube_conn_kernel(int nx, int ny, int nz, int32_t *__restrict__ conn) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t ncell = (int64_t)(nx - 1) * (ny - 1) * (nz - 1);
    if (t >= ncell * 6) return;
    const int64_t cell = t / 6;
    const int p = (int)(t % 6);
    const int ci = (int)(cell % (nx - 1)), cj = (int)((cell / (nx - 1)) % (ny - 1));
    const int ck = (int)(cell / ((int64_t)(nx - 1) * (ny - 1)));
    const int64_t stride[3] = {1, nx, (int64_t)nx * ny};
    int64_t v = ci + (int64_t)nx * (cj + (int64_t)ny * ck);
    int4 e;
    e.x = (int)v; v += stride[c_perms[p][0]];
    e.y = (int)v; v += stride[c_perms[p][1]];
    e.z = (int)v; v += stride[c_perms[p][2]];
    e.w = (int)v;
    *reinterpret_cast<int4 *>(conn + t * 4) = e;
}

int GeneralAssembler::make_cube(int nx, int ny, int nz, cudaStream_t st) {
    N = (int64_t)nx * ny * nz; ne = 6ll * (nx - 1) * (ny - 1) * (nz - 1); npe = 4; has_z = true;
    if (ne >= (1ll << 31) || N >= (1ll << 31)) HEAT_FAIL(20, "explicit cube too large for int32 ids");
    HEAT_TRY(x.alloc((size_t)N)); HEAT_TRY(y.alloc((size_t)N)); HEAT_TRY(z.alloc((size_t)N));
    HEAT_TRY(conn.alloc((size_t)(ne * 4))); HEAT_TRY(bc.alloc((size_t)N));
    cube_nodes_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(nx, ny, nz, x.p, y.p, z.p, bc.p);
    HEAT_LAUNCHED();
    cube_conn_kernel<<<(unsigned)((ne + 255) / 256), 256, 0, st>>>(nx, ny, nz, conn.p);
    HEAT_LAUNCHED();
    return 0;
}

int GeneralAssembler::build_pattern(cudaStream_t st) {
    // ---- elimination (:216-252) ----
    DevBuf<int32_t> flag;
    HEAT_TRY(flag.alloc((size_t)N)); HEAT_TRY(red.alloc((size_t)N));
    dof_flag_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(N, bc.p, flag.p);
    HEAT_LAUNCHED();
    {
        size_t tb = 0;
        HEAT_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, flag.p, red.p, N, st));
        DevBuf<char> tmp; HEAT_TRY(tmp.alloc(tb));
        HEAT_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, flag.p, red.p, N, st));
        int32_t last_red = 0, last_flag = 0;
        HEAT_CUDA(cudaMemcpyAsync(&last_red, red.p + (N - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        HEAT_CUDA(cudaMemcpyAsync(&last_flag, flag.p + (N - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        HEAT_CUDA(cudaStreamSynchronize(st));
        n = (int64_t)last_red + last_flag;
    }
    HEAT_TRY(red2orig.alloc((size_t)n));
    red_finish_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(N, flag.p, red.p, red2orig.p);
    HEAT_LAUNCHED();
    flag.release();

    // ---- node -> incident elements, ascending element id: stable radix sort of (node, elem) ----
    const int64_t total = ne * npe;
    {
        DevBuf<int32_t> keys_in, elems_in, keys_out;
        HEAT_TRY(keys_in.alloc((size_t)total)); HEAT_TRY(elems_in.alloc((size_t)total));
        HEAT_TRY(keys_out.alloc((size_t)total)); HEAT_TRY(n2e.alloc((size_t)total));
        HEAT_TRY(n2e_ptr.alloc((size_t)N + 1));
        if (total > 0) {
            n2e_keys_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(total, npe, conn.p, keys_in.p, elems_in.p);
            HEAT_LAUNCHED();
            int end_bit = 1;
            while (end_bit < 31 && (1ll << end_bit) < N) ++end_bit;
            size_t tb = 0;
            HEAT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, keys_in.p, keys_out.p, elems_in.p, n2e.p, total, 0, end_bit, st));
            DevBuf<char> tmp; HEAT_TRY(tmp.alloc(tb));
            HEAT_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, keys_in.p, keys_out.p, elems_in.p, n2e.p, total, 0, end_bit, st));
        }
        n2e_ptr_kernel<<<(unsigned)((N + 1 + 255) / 256), 256, 0, st>>>(N, total, keys_out.p, n2e_ptr.p);
        HEAT_LAUNCHED();
        HEAT_CUDA(cudaStreamSynchronize(st));
    }

    // ---- pattern of all reduced rows ----
    DevBuf<int> ovf; HEAT_TRY(ovf.alloc(1));
    HEAT_CUDA(cudaMemsetAsync(ovf.p, 0, sizeof(int), st));
    DevBuf<int64_t> row_len; HEAT_TRY(row_len.alloc((size_t)n + 1));
    HEAT_TRY(grow_ptr.alloc((size_t)n + 1));
    HEAT_CUDA(cudaMemsetAsync(row_len.p, 0, sizeof(int64_t) * (size_t)(n + 1), st));
    if (n > 0) {
        pattern_count_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, red2orig.p, red.p, n2e_ptr.p, n2e.p, npe,
                                                                       conn.p, row_len.p, ovf.p);
        HEAT_LAUNCHED();
    }
    {
        size_t tb = 0;
        HEAT_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, row_len.p, grow_ptr.p, n + 1, st));
        DevBuf<char> tmp; HEAT_TRY(tmp.alloc(tb));
        HEAT_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, row_len.p, grow_ptr.p, n + 1, st));
        size_t tb2 = 0;
        DevBuf<int64_t> mx; HEAT_TRY(mx.alloc(1));
        HEAT_CUDA(cub::DeviceReduce::Max(nullptr, tb2, row_len.p, mx.p, n + 1, st));
        DevBuf<char> tmp2; HEAT_TRY(tmp2.alloc(tb2));
        HEAT_CUDA(cub::DeviceReduce::Max(tmp2.p, tb2, row_len.p, mx.p, n + 1, st));
        int64_t h_mx = 0; int h_ovf = 0;
        HEAT_CUDA(cudaMemcpyAsync(&nnz, grow_ptr.p + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        HEAT_CUDA(cudaMemcpyAsync(&h_mx, mx.p, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        HEAT_CUDA(cudaMemcpyAsync(&h_ovf, ovf.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        HEAT_CUDA(cudaStreamSynchronize(st));
        if (h_ovf) HEAT_FAIL(21, "assemble: a node has more than %d neighbours", kMaxNbr);
        max_row = (int32_t)h_mx;
    }
    HEAT_TRY(gcol.alloc((size_t)nnz));
    if (n > 0) {
        pattern_fill_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, red2orig.p, red.p, n2e_ptr.p, n2e.p, npe,
                                                                      conn.p, grow_ptr.p, gcol.p, ovf.p);
        HEAT_LAUNCHED();
    }
    return 0;
}

int GeneralAssembler::fill_values(int mode, int64_t n_owned, const int32_t *d_owned, const int32_t *d_g2l,
                                  const int64_t *d_lrow_ptr, int32_t *d_lcol, double *d_lval, double *d_b,
                                  cudaStream_t st) {
    if (mode == HEAT_OP_P1_FEM && npe != 4 && npe != 3)
        HEAT_FAIL(22, "P1_FEM needs TETRA (4 nodes) or TRI (3 nodes) elements, mesh has %d nodes per element", npe);
    if (n_owned == 0) return 0;
    DevBuf<int> ovf; HEAT_TRY(ovf.alloc(1));
    HEAT_CUDA(cudaMemsetAsync(ovf.p, 0, sizeof(int), st));
    values_kernel<<<(unsigned)((n_owned + 127) / 128), 128, 0, st>>>(
        n_owned, d_owned, d_g2l, red2orig.p, red.p, n2e_ptr.p, n2e.p, npe, conn.p, x.p, y.p, has_z ? z.p : nullptr,
        bc.p, grow_ptr.p, gcol.p, d_lrow_ptr, d_lcol, d_lval, d_b, mode, ovf.p);
    HEAT_LAUNCHED();
    return 0;
}

// =================================================================================================
// analytic Kuhn-cube path (one thread per owned row, nothing but the matrix is written)
// =================================================================================================
// the 15 stencil slots in ascending global-id order; slot 7 is the diagonal
__device__ __forceinline__ void slot_delta(int slot, int &di, int &dj, int &dk) {
    const int code = slot > 7 ? slot - 7 : 7 - slot;
    const int sgn = slot > 7 ? 1 : -1;
    di = sgn * (code & 1); dj = sgn * ((code >> 1) & 1); dk = sgn * ((code >> 2) & 1);
}

__device__ __forceinline__ void cube_row_ijk(const CubeGeom &c, int64_t l, int &i, int &j, int &k) {
    const int64_t plane = c.plane;
    k = (int)(l / plane) + c.k0;
    const int64_t rem = l % plane;
    j = (int)(rem / (c.nx - 2));
    i = (int)(rem % (c.nx - 2)) + 1;
}
// local column id of DOF node (i,j,k): owned rows first, then the lower ghost plane, then the upper
__device__ __forceinline__ int32_t cube_local_col(const CubeGeom &c, int i, int j, int k) {
    const int64_t inplane = (int64_t)j * (c.nx - 2) + (i - 1);
    if (k >= c.k0 && k < c.k1) return (int32_t)((int64_t)(k - c.k0) * c.plane + inplane);
    if (k < c.k0) return (int32_t)(c.n_owned + inplane);
    return (int32_t)(c.n_owned + c.ghost_lo + inplane);
}

__global__ void cube_count_kernel(CubeGeom c, int64_t *__restrict__ row_len) {
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= c.n_owned) return;
    int i, j, k;
    cube_row_ijk(c, l, i, j, k);
    int len = 0;
    for (int s = 0; s < 15; ++s) {
        int di, dj, dk;
        slot_delta(s, di, dj, dk);
        const int ii = i + di, jj = j + dj, kk = k + dk;
        len += (ii >= 1 && ii <= c.nx - 2 && jj >= 0 && jj < c.ny && kk >= 0 && kk < c.nz);
    }
    row_len[l] = len;
}

__global__ void __launch_bounds__(128)
cube_fill_kernel(CubeGeom c, int mode, const int64_t *__restrict__ row_ptr, int32_t *__restrict__ col,
                 double *__restrict__ val, double *__restrict__ b) {
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= c.n_owned) return;
    int i, j, k;
    cube_row_ijk(c, l, i, j, k);
    double v[15];
#pragma unroll
    for (int s = 0; s < 15; ++s) v[s] = 0.0;
    double bsum = 0.0;
    if (mode == HEAT_OP_GRAPH_LAPLACIAN) {
        int deg = 0;
        for (int s = 0; s < 15; ++s) {
            if (s == 7) continue;
            int di, dj, dk;
            slot_delta(s, di, dj, dk);
            const int ii = i + di, jj = j + dj, kk = k + dk;
            if (ii < 0 || ii >= c.nx || jj < 0 || jj >= c.ny || kk < 0 || kk >= c.nz) continue;
            ++deg;
            if (ii == 0) bsum += 1000.0; else if (ii == c.nx - 1) bsum += 100.0; else v[s] = -1.0;
        }
        v[7] = (double)deg;
    } else {
        // incident tets in ascending element id: cells by (ck, cj, ci), then the 6 Kuhn permutations
        for (int ck = k - 1; ck <= k; ++ck) {
            if (ck < 0 || ck >= c.nz - 1) continue;
            for (int cj = j - 1; cj <= j; ++cj) {
                if (cj < 0 || cj >= c.ny - 1) continue;
                for (int ci = i - 1; ci <= i; ++ci) {
                    if (ci < 0 || ci >= c.nx - 1) continue;
                    const int o[3] = {i - ci, j - cj, k - ck};          // node offset inside the cell
                    for (int pm = 0; pm < 6; ++pm) {
                        int V[4][3];
                        V[0][0] = 0; V[0][1] = 0; V[0][2] = 0;
                        for (int q = 0; q < 3; ++q) {
                            V[q + 1][0] = V[q][0]; V[q + 1][1] = V[q][1]; V[q + 1][2] = V[q][2];
                            V[q + 1][c_perms[pm][q]] += 1;
                        }
                        int a = -1;
                        for (int q = 0; q < 4; ++q)
                            if (V[q][0] == o[0] && V[q][1] == o[1] && V[q][2] == o[2]) a = q;
                        if (a < 0) continue;
                        double p[4][3], G[4][3], s;
                        for (int q = 0; q < 4; ++q) {
                            p[q][0] = -5.0 + 10.0 * (double)(ci + V[q][0]) / (double)(c.nx - 1);
                            p[q][1] = -5.0 + 10.0 * (double)(cj + V[q][1]) / (double)(c.ny - 1);
                            p[q][2] = -5.0 + 10.0 * (double)(ck + V[q][2]) / (double)(c.nz - 1);
                        }
                        tet_G(p, G, s);
                        for (int q = 0; q < 4; ++q) {
                            const double kab = ((G[a][0] * G[q][0] + G[a][1] * G[q][1]) + G[a][2] * G[q][2]) * s;
                            const int di = V[q][0] - o[0], dj = V[q][1] - o[1], dk = V[q][2] - o[2];
                            const int code = (di != 0) + 2 * (dj != 0) + 4 * (dk != 0);
                            const int slot = (di + dj + dk) > 0 ? 7 + code : 7 - code;
                            const int ii = i + di;
                            if (ii == 0) { const double t2 = kab * 1000.0; bsum = bsum - t2; }
                            else if (ii == c.nx - 1) { const double t2 = kab * 100.0; bsum = bsum - t2; }
                            else v[slot] += kab;
                        }
                    }
                }
            }
        }
    }
    int32_t *cp = col + row_ptr[l];
    double *vp = val + row_ptr[l];
    int len = 0;
    for (int s = 0; s < 15; ++s) {
        int di, dj, dk;
        slot_delta(s, di, dj, dk);
        const int ii = i + di, jj = j + dj, kk = k + dk;
        if (ii >= 1 && ii <= c.nx - 2 && jj >= 0 && jj < c.ny && kk >= 0 && kk < c.nz) {
            cp[len] = cube_local_col(c, ii, jj, kk);
            vp[len] = v[s];
            ++len;
        }
    }
    b[l] = bsum;
}

int cube_assemble(const CubeGeom &c, int mode, heat_matrix *A, double *d_b, cudaStream_t st) {
    const int64_t n = c.n_owned;
    HEAT_TRY(A->row_ptr.alloc((size_t)n + 1));
    DevBuf<int64_t> row_len; HEAT_TRY(row_len.alloc((size_t)n + 1));
    HEAT_CUDA(cudaMemsetAsync(row_len.p, 0, sizeof(int64_t) * (size_t)(n + 1), st));
    if (n > 0) {
        cube_count_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(c, row_len.p);
        HEAT_LAUNCHED();
    }
    size_t tb = 0;
    HEAT_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, row_len.p, A->row_ptr.p, n + 1, st));
    DevBuf<char> tmp; HEAT_TRY(tmp.alloc(tb));
    HEAT_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, row_len.p, A->row_ptr.p, n + 1, st));
    HEAT_CUDA(cudaMemcpyAsync(&A->nnz, A->row_ptr.p + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    HEAT_CUDA(cudaStreamSynchronize(st));
    row_len.release(); tmp.release();
    A->max_row_len = 15;
    HEAT_TRY(A->col.alloc((size_t)A->nnz));
    HEAT_TRY(A->val.alloc((size_t)A->nnz));
    if (n > 0) {
        cube_fill_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(c, mode, A->row_ptr.p, A->col.p, A->val.p, d_b);
        HEAT_LAUNCHED();
    }
    return 0;
}

}  // namespace heat
