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
#include "kernels.cuh"
#include "sell_dict.cuh"
#include "kuhn_rows_generated.cuh"

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
// One thread per row (owner computes: atomic-free, accumulation in ascending element id like the oracle).  The row's
// entries are accumulated in SHARED memory ([entry][thread], conflict-free) and written once — the first version
// read-modify-wrote them in global memory for each of the ~96 element contributions of a row, 16x the output in DRAM
// writes (ncu: 32 GB written for a 2 GB matrix at 256^3).  Node coordinates come as one 32-byte (x, y, z, 0) record
// per node — one sector per gathered node instead of three — and a tet's four node ids as one 16-byte load.
constexpr int kAccCap = 24;            // row entries (values AND column ids) held in shared memory; longer rows work in global memory
constexpr int kValThreads = 128;
// NPE: nodes per element known at compile time for the P1 operator (4: tets, 3: triangles) so that the element
// arrays stay in registers; 0: graph Laplacian (any clique size)
template <int NPE>
__global__ void __launch_bounds__(kValThreads)
values_kernel(int64_t n_owned, const int32_t *__restrict__ owned, const int32_t *__restrict__ g2l,
              const int64_t *__restrict__ red2orig, const int32_t *__restrict__ red,
              const int64_t *__restrict__ n2e_ptr, const int32_t *__restrict__ n2e, int npe,
              const int32_t *__restrict__ conn, const double4 *__restrict__ xyz, const double *__restrict__ bc,
              const int64_t *__restrict__ grow_ptr, const int32_t *__restrict__ gcol,
              const int64_t *__restrict__ lrow_ptr, int32_t *__restrict__ lcol, double *__restrict__ lval,
              double *__restrict__ b, int mode, int *overflow) {
    __shared__ double acc_s[kAccCap][kValThreads];
    __shared__ int32_t col_s[kAccCap][kValThreads];        // the row's pattern: every element contribution is binary-searched in it
    const int tid = threadIdx.x;
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_owned) return;
    const int64_t gi = owned ? (int64_t)owned[l] : l;
    const int64_t g = red2orig[gi];
    const int32_t *gc = gcol + grow_ptr[gi];
    const int len = (int)(grow_ptr[gi + 1] - grow_ptr[gi]);
    int32_t *lc = lcol ? lcol + lrow_ptr[l] : nullptr;
    double *lv = lval + lrow_ptr[l];
    int dpos = 0;
    const bool in_smem = NPE != 0 && len <= kAccCap;
    for (int t = 0; t < len; ++t) {
        const int32_t c = gc[t];
        if (c == gi) dpos = t;
        if (lc) lc[t] = g2l ? g2l[c] : c;
        if (in_smem) { col_s[t][tid] = c; acc_s[t][tid] = 0.0; }
    }
    double bsum = 0.0;
    if constexpr (NPE == 0) {
        int32_t buf[kMaxNbr];
        const int u = collect_neighbours(g, n2e_ptr, n2e, npe, conn, buf, overflow);
        for (int t = 0; t < u; ++t)
            if (red[buf[t]] < 0) bsum += bc[buf[t]];               // :671-687
        for (int t = 0; t < len; ++t) lv[t] = -1.0;                // :601
        lv[dpos] = (double)u;                                      // :606 full degree
    } else {
        if (!in_smem) { for (int t = 0; t < len; ++t) lv[t] = 0.0; }
        for (int64_t q = n2e_ptr[g]; q < n2e_ptr[g + 1]; ++q) {
            constexpr int kN = NPE == 3 ? 3 : 4;
            const int32_t *e = conn + (int64_t)n2e[q] * kN;
            int32_t ev[4];
            if (kN == 4) {
                const int4 e4 = __ldg(reinterpret_cast<const int4 *>(e));
                ev[0] = e4.x; ev[1] = e4.y; ev[2] = e4.z; ev[3] = e4.w;
            } else {
                ev[0] = e[0]; ev[1] = e[1]; ev[2] = e[2]; ev[3] = 0;
            }
            double p[4][3], G[4][3], s;
            int32_t rv[4];                                      // reduced ids of the element's nodes (-1: Dirichlet)
            int a = -1;
#pragma unroll
            for (int k = 0; k < kN; ++k) {
                const double2 *c2 = reinterpret_cast<const double2 *>(xyz + ev[k]);     // one 32-byte sector, two 16-byte loads
                const double2 xy = __ldg(c2), zw = __ldg(c2 + 1);
                p[k][0] = xy.x; p[k][1] = xy.y; p[k][2] = zw.x;
                rv[k] = (int32_t)__double_as_longlong(zw.y);
                if (ev[k] == g && a < 0) a = k;
            }
            if (kN == 4) tet_G(p, G, s); else tri_G(p, G, s);
            double Ga[3];                                       // G[a] without a dynamically indexed (local-memory) array
#pragma unroll
            for (int d = 0; d < 3; ++d) Ga[d] = a == 0 ? G[0][d] : a == 1 ? G[1][d] : a == 2 ? G[2][d] : G[3][d];
#pragma unroll
            for (int k = 0; k < kN; ++k) {
                const double kab = ((Ga[0] * G[k][0] + Ga[1] * G[k][1]) + Ga[2] * G[k][2]) * s;
                const int32_t j = ev[k];
                int slot = -1;
                if (j == g) {
                    slot = dpos;
                } else if (rv[k] >= 0) {
                    const int32_t r = rv[k];
                    int lo = 0, hi = len - 1;
                    if (in_smem) { while (lo < hi) { int mid = (lo + hi) >> 1; if (col_s[mid][tid] < r) lo = mid + 1; else hi = mid; } }
                    else { while (lo < hi) { int mid = (lo + hi) >> 1; if (gc[mid] < r) lo = mid + 1; else hi = mid; } }
                    slot = lo;
                } else {
                    const double t2 = kab * bc[j];
                    bsum = bsum - t2;
                }
                if (slot >= 0) {
                    if (in_smem) acc_s[slot][tid] += kab; else lv[slot] += kab;
                }
            }
        }
        if (in_smem)
            for (int t = 0; t < len; ++t) lv[t] = acc_s[t][tid];
    }
    b[l] = bsum;
}

// (x, y, z, reduced id) records of the nodes: one 32-byte sector per gathered node carries everything the element
// loop needs to know about it (the reduced id, -1 for Dirichlet nodes, rides in the fourth slot as an integer)
__global__ void pack_xyz_kernel(int64_t N, const double *__restrict__ X, const double *__restrict__ Y, const double *__restrict__ Z,
                                const int32_t *__restrict__ red, double4 *__restrict__ xyz) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < N) xyz[g] = make_double4(X[g], Y[g], Z ? Z[g] : 0.0, __longlong_as_double((long long)red[g]));
}

// elapsed device time between two points of a stream (both events are created and destroyed here)
struct StreamTimer {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaStream_t st;
    explicit StreamTimer(cudaStream_t s) : st(s) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, st); }
    float stop() {                                       // synchronises the stream
        float ms = 0.f;
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        return ms;
    }
    ~StreamTimer() { if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); }
};

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
        StreamTimer t_sort(st);
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
        phase_ms[0] = t_sort.stop();
    }

    // ---- pattern of all reduced rows ----
    DevBuf<int> ovf; HEAT_TRY(ovf.alloc(1));
    HEAT_CUDA(cudaMemsetAsync(ovf.p, 0, sizeof(int), st));
    DevBuf<int64_t> row_len; HEAT_TRY(row_len.alloc((size_t)n + 1));
    HEAT_TRY(grow_ptr.alloc((size_t)n + 1));
    HEAT_CUDA(cudaMemsetAsync(row_len.p, 0, sizeof(int64_t) * (size_t)(n + 1), st));
    if (n > 0) {
        StreamTimer t_count(st);
        pattern_count_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, red2orig.p, red.p, n2e_ptr.p, n2e.p, npe,
                                                                       conn.p, row_len.p, ovf.p);
        HEAT_LAUNCHED();
        phase_ms[1] = t_count.stop();
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
        StreamTimer t_fill(st);
        pattern_fill_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, red2orig.p, red.p, n2e_ptr.p, n2e.p, npe,
                                                                      conn.p, grow_ptr.p, gcol.p, ovf.p);
        HEAT_LAUNCHED();
        phase_ms[2] = t_fill.stop();
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
    if (!xyz.p) {
        HEAT_TRY(xyz.alloc((size_t)N * 4));
        pack_xyz_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(N, x.p, y.p, has_z ? z.p : nullptr, red.p, reinterpret_cast<double4 *>(xyz.p));
        HEAT_LAUNCHED();
    }
    StreamTimer t_val(st);
    const unsigned vgrid = (unsigned)((n_owned + kValThreads - 1) / kValThreads);
    const double4 *xyz4 = reinterpret_cast<const double4 *>(xyz.p);
#define HEAT_VALUES_ARGS n_owned, d_owned, d_g2l, red2orig.p, red.p, n2e_ptr.p, n2e.p, npe, conn.p, xyz4, bc.p, grow_ptr.p, gcol.p, \
                         d_lrow_ptr, d_lcol, d_lval, d_b, mode, ovf.p
    if (mode == HEAT_OP_GRAPH_LAPLACIAN) values_kernel<0><<<vgrid, kValThreads, 0, st>>>(HEAT_VALUES_ARGS);
    else if (npe == 4) values_kernel<4><<<vgrid, kValThreads, 0, st>>>(HEAT_VALUES_ARGS);
    else values_kernel<3><<<vgrid, kValThreads, 0, st>>>(HEAT_VALUES_ARGS);
#undef HEAT_VALUES_ARGS
    HEAT_LAUNCHED();
    phase_ms[3] = t_val.stop();
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
    // local row ids fit 31 bits (api.cu checks n_owned + ghosts < 2^31): 32-bit divisions
    const uint32_t plane = (uint32_t)c.plane, w = (uint32_t)(c.nx - 2), lu = (uint32_t)l;
    const uint32_t kk = lu / plane, rem = lu - kk * plane, jj = rem / w;
    k = (int)kk + c.k0;
    j = (int)jj;
    i = (int)(rem - jj * w) + 1;
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

// =================================================================================================
// analytic Kuhn-cube path, DIRECT to the SpMV format (default): one warp per 64-row SELL slice
// =================================================================================================
// What IO::assemble's insertGlobalValues + fillComplete (ExodusIO.hpp:591-609) amount to for the structured
// benchmark cubes, without ever materialising a CSR: each warp computes the 64 rows of one slice (lane l: rows
// l and l+32 of the slice, i.e. neighbouring nodes along x), stages the slice in shared memory in its final
// SELL layout [entry k][row], builds the slice's offset table and 1-byte indices there (sell_dict.cuh) and
// writes values / indices with ONE TMA bulk store each (cp.async.bulk.global.shared::cta) — fully coalesced,
// 8 + 1 bytes per stored entry, against the 12-byte CSR written row by row at a 180-byte stride (and then
// re-read, re-written as SELL and scanned twice for the dictionary) of the CSR-first path below.
// The element arithmetic is the oracle's, operation for operation and in the same order (ascending element
// id), so the values stay bit-identical: the 8 cells x 6 permutations around a node are unrolled at compile
// time (which vertex of which permutation the node is, and which stencil slot each of the tet's vertices
// lands in, are constants), node coordinates are evaluated once per row (9 divisions instead of 288).
constexpr int kCubeTileVal = 15 * kSellChunk * 8;       // 7680 B
constexpr int kCubeTileCol = 15 * kSellChunk * 4;       // 3840 B
constexpr int kCubeTileIdx = 15 * kSellChunk;           //  960 B
constexpr int kCubeTileTab = kSellDictCap * 4;          //  256 B
constexpr int kCubeTileBytes = kCubeTileVal + kCubeTileCol + kCubeTileIdx + kCubeTileTab;   // 12736 B per warp
constexpr int kCubeWarps = 8;

__host__ __device__ constexpr int kuhn_axis(int pm, int q) {        // axis of step q of permutation pm (lexicographic)
    return pm == 0 ? (q == 0 ? 0 : q == 1 ? 1 : 2) : pm == 1 ? (q == 0 ? 0 : q == 1 ? 2 : 1) :
           pm == 2 ? (q == 0 ? 1 : q == 1 ? 0 : 2) : pm == 3 ? (q == 0 ? 1 : q == 1 ? 2 : 0) :
           pm == 4 ? (q == 0 ? 2 : q == 1 ? 0 : 1) : (q == 0 ? 2 : q == 1 ? 1 : 0);
}
// coordinate d of vertex q of permutation pm, relative to the cell's first node: 1 once axis d has been stepped
__host__ __device__ constexpr int kuhn_vtx(int pm, int q, int d) {
    int v = 0;
    for (int t = 0; t < q; ++t) v += (kuhn_axis(pm, t) == d) ? 1 : 0;
    return v;
}
__host__ __device__ constexpr int kuhn_vertex_of(int pm, int ox, int oy, int oz) {   // -1: the node is not on pm's path
    for (int q = 0; q < 4; ++q)
        if (kuhn_vtx(pm, q, 0) == ox && kuhn_vtx(pm, q, 1) == oy && kuhn_vtx(pm, q, 2) == oz) return q;
    return -1;
}

// contributions of tet (cell = node - (OX,OY,OZ), permutation PM) to the row of the node.  X[t], Y[t], Z[t] are
// the coordinates of grid lines i-1+t, j-1+t, k-1+t.
template <int OX, int OY, int OZ, int PM>
__device__ __forceinline__ void kuhn_tet_row(const double (&X)[3], const double (&Y)[3], const double (&Z)[3],
                                             bool at_lo, bool at_hi, double (&v)[15], double &bsum) {
    constexpr int a = kuhn_vertex_of(PM, OX, OY, OZ);
    if constexpr (a >= 0) {
        double p[4][3], G[4][3], s;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            p[q][0] = X[1 - OX + kuhn_vtx(PM, q, 0)];
            p[q][1] = Y[1 - OY + kuhn_vtx(PM, q, 1)];
            p[q][2] = Z[1 - OZ + kuhn_vtx(PM, q, 2)];
        }
        tet_G(p, G, s);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double kab = ((G[a][0] * G[q][0] + G[a][1] * G[q][1]) + G[a][2] * G[q][2]) * s;
            const int di = kuhn_vtx(PM, q, 0) - OX, dj = kuhn_vtx(PM, q, 1) - OY, dk = kuhn_vtx(PM, q, 2) - OZ;
            const int code = (di != 0) + 2 * (dj != 0) + 4 * (dk != 0);
            const int slot = (di + dj + dk) > 0 ? 7 + code : 7 - code;
            if (di < 0 && at_lo) { const double t2 = kab * 1000.0; bsum = bsum - t2; }
            else if (di > 0 && at_hi) { const double t2 = kab * 100.0; bsum = bsum - t2; }
            else v[slot] += kab;
        }
    }
}
template <int OX, int OY, int OZ>
__device__ __forceinline__ void kuhn_cell_row(const double (&X)[3], const double (&Y)[3], const double (&Z)[3],
                                              bool at_lo, bool at_hi, double (&v)[15], double &bsum) {
    kuhn_tet_row<OX, OY, OZ, 0>(X, Y, Z, at_lo, at_hi, v, bsum);
    kuhn_tet_row<OX, OY, OZ, 1>(X, Y, Z, at_lo, at_hi, v, bsum);
    kuhn_tet_row<OX, OY, OZ, 2>(X, Y, Z, at_lo, at_hi, v, bsum);
    kuhn_tet_row<OX, OY, OZ, 3>(X, Y, Z, at_lo, at_hi, v, bsum);
    kuhn_tet_row<OX, OY, OZ, 4>(X, Y, Z, at_lo, at_hi, v, bsum);
    kuhn_tet_row<OX, OY, OZ, 5>(X, Y, Z, at_lo, at_hi, v, bsum);
}

// the 15 slot values + right-hand side of node (i,j,k)
__device__ __forceinline__ void cube_row_values(const CubeGeom &c, int mode, int i, int j, int k, double (&v)[15], double &bsum) {
#pragma unroll
    for (int s = 0; s < 15; ++s) v[s] = 0.0;
    bsum = 0.0;
    if (mode == HEAT_OP_GRAPH_LAPLACIAN) {
        int deg = 0;
#pragma unroll
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
        return;
    }
    double X[3], Y[3], Z[3];
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        X[t] = -5.0 + 10.0 * (double)(i - 1 + t) / (double)(c.nx - 1);
        Y[t] = -5.0 + 10.0 * (double)(j - 1 + t) / (double)(c.ny - 1);
        Z[t] = -5.0 + 10.0 * (double)(k - 1 + t) / (double)(c.nz - 1);
    }
    const bool at_lo = i == 1, at_hi = i == c.nx - 2;
    const bool jlo = j >= 1, jhi = j <= c.ny - 2, klo = k >= 1, khi = k <= c.nz - 2;   // cells at j-1 / j / k-1 / k exist
#ifndef HEAT_KUHN_FULL_ARITHMETIC
    // the same 24 tets with the exact-zero operands of tet_G / K_ab eliminated and common subexpressions shared
    // (tools/gen_kuhn_rows.py: ~125 fp64 operations + 15 divisions per row instead of ~1500; bit-identical values)
    kuhn_row_p1_generated(X, Y, Z, jlo, jhi, klo, khi, at_lo, at_hi, v, bsum);
    return;
#endif
    // ascending element id: cells by (ck, cj, ci); the x-cells i-1 and i always exist for 1 <= i <= nx-2
    if (klo && jlo) { kuhn_cell_row<1, 1, 1>(X, Y, Z, at_lo, at_hi, v, bsum); kuhn_cell_row<0, 1, 1>(X, Y, Z, at_lo, at_hi, v, bsum); }
    if (klo && jhi) { kuhn_cell_row<1, 0, 1>(X, Y, Z, at_lo, at_hi, v, bsum); kuhn_cell_row<0, 0, 1>(X, Y, Z, at_lo, at_hi, v, bsum); }
    if (khi && jlo) { kuhn_cell_row<1, 1, 0>(X, Y, Z, at_lo, at_hi, v, bsum); kuhn_cell_row<0, 1, 0>(X, Y, Z, at_lo, at_hi, v, bsum); }
    if (khi && jhi) { kuhn_cell_row<1, 0, 0>(X, Y, Z, at_lo, at_hi, v, bsum); kuhn_cell_row<0, 0, 0>(X, Y, Z, at_lo, at_hi, v, bsum); }
}

__device__ __forceinline__ bool cube_slot_stored(const CubeGeom &c, int s, int i, int j, int k, int &ii, int &jj, int &kk) {
    int di, dj, dk;
    slot_delta(s, di, dj, dk);
    ii = i + di; jj = j + dj; kk = k + dk;
    return ii >= 1 && ii <= c.nx - 2 && jj >= 0 && jj < c.ny && kk >= 0 && kk < c.nz;
}

// one warp per slice: slice_entries[s] = 64 * (longest row of the slice); *nnz_total += stored entries
__global__ void __launch_bounds__(kBlock)
cube_width_kernel(CubeGeom c, int64_t n_slices, int64_t *__restrict__ slice_entries, unsigned long long *nnz_total) {
    const int lane = threadIdx.x & 31;
    const int64_t s = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    int len_sum = 0;
    if (s < n_slices) {
        int wmax = 0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t l = s * kSellChunk + h * 32 + lane;
            if (l >= c.n_owned) continue;
            int i, j, k, ii, jj, kk, len = 0;
            cube_row_ijk(c, l, i, j, k);
#pragma unroll
            for (int q = 0; q < 15; ++q) len += cube_slot_stored(c, q, i, j, k, ii, jj, kk) ? 1 : 0;
            wmax = len > wmax ? len : wmax;
            len_sum += len;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
        if (lane == 0) slice_entries[s] = (int64_t)wmax * kSellChunk;
    }
    __shared__ int blk;
    if (threadIdx.x == 0) blk = 0;
    __syncthreads();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) len_sum += __shfl_xor_sync(0xffffffffu, len_sum, o);
    if (lane == 0 && len_sum) atomicAdd(&blk, len_sum);
    __syncthreads();
    if (threadIdx.x == 0 && blk) atomicAdd(nnz_total, (unsigned long long)blk);
}

struct CubeSellArgs {
    CubeGeom c; int mode;
    const int64_t *slice_ptr; int64_t n_slices;
    double *val;                     // [sell_padded]
    int32_t *col;                    // [sell_padded]           (int32 column stream; null when byte-indexed)
    uint8_t *idx8; int32_t *tab;     // [sell_padded], [n_slices][tpad]   (byte-indexed stream; null otherwise)
    int tpad;
    uint8_t *rowlen;                 // [n_owned] stored entries of every row
    double *diag, *dinv, *b;         // [n_owned]
    int32_t *is_boundary;            // [n_slices] (null on one GPU): slice references a ghost column
    int *max_tab;                    // atomicMax of the table lengths (> tpad: the caller falls back to the CSR-first path)
};

template <bool C8>
__global__ void __launch_bounds__(kCubeWarps * 32, 2) cube_sell_kernel(CubeSellArgs a) {
    extern __shared__ __align__(128) unsigned char cube_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *my = cube_smem + (size_t)warp * kCubeTileBytes;
    double *tv = reinterpret_cast<double *>(my);
    int32_t *tc = reinterpret_cast<int32_t *>(my + kCubeTileVal);
    uint8_t *ti = my + kCubeTileVal + kCubeTileCol;
    int32_t *tab = reinterpret_cast<int32_t *>(my + kCubeTileVal + kCubeTileCol + kCubeTileIdx);
    const CubeGeom &c = a.c;
    for (int64_t s = (int64_t)blockIdx.x * kCubeWarps + warp; s < a.n_slices; s += (int64_t)gridDim.x * kCubeWarps) {
        const int64_t base = a.slice_ptr[s];
        const int w = (int)((a.slice_ptr[s + 1] - base) >> 6);
        int ghost = 0;
        bool full = true;                                // every row of the slice stores all 15 stencil slots
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            const int rl = h * 32 + lane;
            const int64_t l = s * kSellChunk + rl;
            int cnt = 0;
            // padding: (value 0, column = own row); rows past the end point 64 rows back (constant offset, sell.cu)
            int32_t padcol = l < c.n_owned ? (int32_t)l : (l >= kSellChunk ? (int32_t)(l - kSellChunk) : 0);
            if (l < c.n_owned) {
                int i, j, k;
                cube_row_ijk(c, l, i, j, k);
                double v[15], bsum;
                cube_row_values(c, a.mode, i, j, k, v, bsum);
                if (i >= 2 && i <= c.nx - 3 && j >= 1 && j <= c.ny - 2 && k >= c.k0 + 1 && k <= c.k1 - 2) {
                    // all 15 neighbours are owned unknowns: column = row + the slot's constant offset, no per-slot tests
                    const int32_t wd = c.nx - 2, pl = (int32_t)c.plane;
#pragma unroll
                    for (int q = 0; q < 15; ++q) {
                        const int code = q > 7 ? q - 7 : 7 - q;
                        const int32_t off = (code & 1) + ((code >> 1) & 1) * wd + ((code >> 2) & 1) * pl;
                        tv[q * kSellChunk + rl] = v[q];
                        tc[q * kSellChunk + rl] = (int32_t)l + (q > 7 ? off : -off);
                    }
                    cnt = 15;
                } else {
#pragma unroll
                    for (int q = 0; q < 15; ++q) {
                        int ii, jj, kk;
                        if (cube_slot_stored(c, q, i, j, k, ii, jj, kk)) {
                            const int32_t cc = cube_local_col(c, ii, jj, kk);
                            ghost |= cc >= c.n_owned;
                            tv[cnt * kSellChunk + rl] = v[q];
                            tc[cnt * kSellChunk + rl] = cc;
                            ++cnt;
                        }
                    }
                }
                a.diag[l] = v[7];
                a.dinv[l] = 1.0 / v[7];
                a.b[l] = bsum;
                a.rowlen[l] = (uint8_t)cnt;
            }
            full = full && cnt == 15;
            for (; cnt < w; ++cnt) { tv[cnt * kSellChunk + rl] = 0.0; tc[cnt * kSellChunk + rl] = padcol; }
        }
        __syncwarp();
        if (C8 && __all_sync(0xffffffffu, full && !ghost)) {
            // interior slice (all but ~3 % of a cube): all 64 rows store the 15 stencil slots in slot order and every
            // column is owned, so entry k of every row has offset (column - row) of slot k — the table IS the 15
            // offsets of lane 0's first row and every index is k; no look-up at all
            if (lane < 15) tab[lane] = tc[lane * kSellChunk] - (int32_t)(s * kSellChunk);
            for (int e = lane; e < 15 * kSellChunk / 4; e += 32)                       // idx[k][0..63] = k, 4 bytes per store
                reinterpret_cast<uint32_t *>(ti)[e] = 0x01010101u * (uint32_t)(e / (kSellChunk / 4));
            __syncwarp();
            for (int t = lane; t < a.tpad; t += 32) a.tab[s * a.tpad + t] = t < 15 ? tab[t] : 0;
        } else if (C8) {
            const int row0 = (int)(s * kSellChunk) + 2 * lane;
            int T = 0;
            bool ok = true;
            for (int k = 0; k < w && ok; ++k) {
                const int2 cc = *reinterpret_cast<const int2 *>(tc + k * kSellChunk + 2 * lane);
                const int off[2] = {cc.x - row0, cc.y - (row0 + 1)};
                int id[2];
                ok = sell_dict_step(tab, T, off, id);
                if (ok) *reinterpret_cast<uchar2 *>(ti + k * kSellChunk + 2 * lane) = make_uchar2((unsigned char)id[0], (unsigned char)id[1]);
            }
            if (!ok) T = kSellDictCap + 1;
            if (lane == 0 && T > a.tpad) atomicMax(a.max_tab, T);
            __syncwarp();
            for (int t = lane; t < a.tpad; t += 32) a.tab[s * a.tpad + t] = t < T ? tab[t] : 0;
        }
        if (a.is_boundary) {
            ghost = __any_sync(0xffffffffu, ghost);
            if (lane == 0) a.is_boundary[s] = ghost;
        }
        // the staged slice goes out as bulk stores: make the generic-proxy writes visible to the async proxy first
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0 && w > 0) {
            const uint32_t sv = (uint32_t)__cvta_generic_to_shared(tv);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         ::"l"(a.val + base), "r"(sv), "r"((uint32_t)w * kSellChunk * 8) : "memory");
            if (C8) {
                const uint32_t si = (uint32_t)__cvta_generic_to_shared(ti);
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                             ::"l"(a.idx8 + base), "r"(si), "r"((uint32_t)w * kSellChunk) : "memory");
            } else {
                const uint32_t sc = (uint32_t)__cvta_generic_to_shared(tc);
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                             ::"l"(a.col + base), "r"(sc), "r"((uint32_t)w * kSellChunk * 4) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");       // the tile is reused by the next slice
        }
        __syncwarp();
    }
}

// Direct-to-SELL assembly of one GPU's slab.  Fills A's SELL arrays, slice tables, row lengths, diagonal and
// right-hand side; leaves the CSR unbuilt (sell.cu: sell_to_csr builds it on demand for exports / ILU).
// Returns 0 with *done = false when the slice tables do not fit the analytic bound (the caller then takes the
// CSR-first path) — cannot happen for slab partitions, kept as a guard.
int cube_assemble_sell(const CubeGeom &c, int mode, bool byte_index, heat_matrix *A, double *d_b, cudaStream_t st, bool *done) {
    *done = false;
    const int64_t n = c.n_owned;
    const int64_t ns = (n + kSellChunk - 1) / kSellChunk;
    if (ns == 0) return 0;
    struct EventPair {                                   // destroyed on every exit path
        cudaEvent_t a = nullptr, b = nullptr;
        ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
    } ev;
    HEAT_CUDA(cudaEventCreate(&ev.a)); HEAT_CUDA(cudaEventCreate(&ev.b));
    cudaEvent_t e0 = ev.a, e1 = ev.b;
    A->n_slices = ns;
    HEAT_TRY(A->slice_ptr.alloc((size_t)ns + 1));
    HEAT_TRY(A->diag.alloc((size_t)n)); HEAT_TRY(A->dinv.alloc((size_t)n)); HEAT_TRY(A->sell_rowlen.alloc((size_t)n));
    DevBuf<int64_t> entries; HEAT_TRY(entries.alloc((size_t)ns));
    DevBuf<unsigned long long> d_nnz; HEAT_TRY(d_nnz.alloc(1));
    DevBuf<int> d_maxtab; HEAT_TRY(d_maxtab.alloc(1));
    HEAT_CUDA(cudaMemsetAsync(d_nnz.p, 0, sizeof(unsigned long long), st));
    HEAT_CUDA(cudaMemsetAsync(d_maxtab.p, 0, sizeof(int), st));
    HEAT_CUDA(cudaMemsetAsync(A->slice_ptr.p, 0, sizeof(int64_t), st));
    cube_width_kernel<<<(unsigned)((ns + kWarpsPerBlock - 1) / kWarpsPerBlock), kBlock, 0, st>>>(c, ns, entries.p, d_nnz.p);
    HEAT_LAUNCHED();
    size_t tb = 0;
    HEAT_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tb, entries.p, A->slice_ptr.p + 1, ns, st));
    DevBuf<char> tmp; HEAT_TRY(tmp.alloc(tb));
    HEAT_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, tb, entries.p, A->slice_ptr.p + 1, ns, st));
    unsigned long long h_nnz = 0;
    HEAT_CUDA(cudaMemcpyAsync(&A->sell_padded, A->slice_ptr.p + ns, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    HEAT_CUDA(cudaMemcpyAsync(&h_nnz, d_nnz.p, sizeof(h_nnz), cudaMemcpyDeviceToHost, st));
    HEAT_CUDA(cudaStreamSynchronize(st));
    A->nnz = (int64_t)h_nnz;
    A->max_row_len = (int32_t)(A->sell_padded > 0 ? 15 : 0);
    // offsets a slice can see: the 15 stencil offsets, 4 more per ghost plane (the dk = -1 / +1 slots of the rows
    // next to it), the tail rows' -64
    const int tpad = (c.ghost_lo || c.ghost_hi) ? 24 : 16;
    CubeSellArgs a;
    a.c = c; a.mode = mode; a.slice_ptr = A->slice_ptr.p; a.n_slices = ns;
    HEAT_TRY(A->sell_val.alloc((size_t)A->sell_padded));
    a.val = A->sell_val.p; a.col = nullptr; a.idx8 = nullptr; a.tab = nullptr; a.tpad = 0;
    if (byte_index) {
        HEAT_TRY(A->sell_idx8.alloc((size_t)A->sell_padded));
        HEAT_TRY(A->sell_tab.alloc((size_t)ns * (size_t)tpad));
        a.idx8 = A->sell_idx8.p; a.tab = A->sell_tab.p; a.tpad = tpad;
    } else {
        HEAT_TRY(A->sell_col.alloc((size_t)A->sell_padded));
        a.col = A->sell_col.p;
    }
    a.rowlen = A->sell_rowlen.p; a.diag = A->diag.p; a.dinv = A->dinv.p; a.b = d_b;
    DevBuf<int32_t> flags;
    const bool need_split = A->n_ghost > 0;
    if (need_split) HEAT_TRY(flags.alloc((size_t)ns));
    a.is_boundary = need_split ? flags.p : nullptr;
    a.max_tab = d_maxtab.p;
    const size_t smem = (size_t)kCubeWarps * kCubeTileBytes;
    int dev = 0, sms = 148;
    HEAT_CUDA(cudaGetDevice(&dev));
    HEAT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int64_t blocks = (ns + kCubeWarps - 1) / kCubeWarps;
    if (blocks > (int64_t)sms * 2 * 8) blocks = (int64_t)sms * 2 * 8;       // persistent-ish: 8 waves of 2 CTAs per SM
    HEAT_CUDA(cudaEventRecord(e0, st));
    if (byte_index) {
        HEAT_CUDA(cudaFuncSetAttribute(cube_sell_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cube_sell_kernel<true><<<(unsigned)blocks, kCubeWarps * 32, smem, st>>>(a);
    } else {
        HEAT_CUDA(cudaFuncSetAttribute(cube_sell_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cube_sell_kernel<false><<<(unsigned)blocks, kCubeWarps * 32, smem, st>>>(a);
    }
    HEAT_LAUNCHED();
    HEAT_CUDA(cudaEventRecord(e1, st));
    int h_maxtab = 0;
    HEAT_CUDA(cudaMemcpyAsync(&h_maxtab, d_maxtab.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    std::vector<int32_t> h_flags;
    if (need_split) {
        h_flags.resize((size_t)ns);
        HEAT_CUDA(cudaMemcpyAsync(h_flags.data(), flags.p, sizeof(int32_t) * (size_t)ns, cudaMemcpyDeviceToHost, st));
    }
    HEAT_CUDA(cudaStreamSynchronize(st));
    float fill_ms = 0.f;
    HEAT_CUDA(cudaEventElapsedTime(&fill_ms, e0, e1));
    A->assemble_fill_ms = fill_ms;
    if (h_maxtab > tpad) {          // guard: a slice saw more distinct offsets than the analytic bound
        A->sell_val.release(); A->sell_idx8.release(); A->sell_tab.release(); A->sell_col.release(); A->sell_rowlen.release();
        A->slice_ptr.release(); A->diag.release(); A->dinv.release();
        A->n_slices = 0; A->sell_padded = 0; A->nnz = 0;
        return 0;
    }
    A->sell_tpad = byte_index ? tpad : 0;
    A->sell_cmode = byte_index ? 1 : 4;
    HEAT_TRY(sell_finish_lists(A, need_split ? h_flags.data() : nullptr, st));
    HEAT_TRY(sell_build_meta(A, st));
    *done = true;
    return 0;
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
