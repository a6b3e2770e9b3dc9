// refview.cpp — "the system as the reference would have assembled it" (pure host code).
//
// The product assembles the FIXED semantics of IO::assemble: every node outside the nodesets is an
// unknown, and an unknown whose neighbours are all Dirichlet keeps its diagonal.  The reference itself (one rank)
// deviates in two documented ways, both confirmed by running its own code (oracle/ref_shim/README.md):
//   D3  a DOF with no DOF neighbour never reaches insertGlobalValues (ExodusIO.hpp:380-386, :591): its row stays
//       empty and it gets no entry in globalIDMap — B keeps its entry (:671-687);
//   D1  `i < getMaxLocalIndex()` (:220, :440) skips the last node: if that node is a DOF it loses its row, and every
//       entry that referenced it lands in column 0 (sparseMapping[...] default-inserts 0 at :598), summed with what
//       is there by fillComplete.
// heat_reference_view_csr applies exactly that to an exported FIXED system, so that a maintainer can diff the
// product's matrix against what the Trilinos build produced.  tests/test_reference_pins.py holds it to the
// digests of the reference's own output on every mesh of its data/ directory and on random meshes.
#include <cstdint>
#include <vector>

#include "common.cuh"

extern "C" int heat_reference_view_csr(heat_ctx *ctx, int64_t n, const int64_t *row_ptr, const int32_t *col, const double *val,
                                       const double *b, const int64_t *red2orig, int64_t *n_out, int64_t *nnz_out,
                                       int64_t *row_ptr_out, int32_t *col_out, double *val_out, double *b_out,
                                       int64_t *n_map_out, int64_t *idmap_reduced_out, int64_t *idmap_original_out) {
    if (!ctx || !row_ptr || (n > 0 && (!red2orig || !b)) || !n_out || !nnz_out || !n_map_out)
        HEAT_FAIL(2, "heat_reference_view_csr: null argument");
    const HostMesh &m = ctx->mesh;
    if (!m.valid || m.is_cube) HEAT_FAIL(4, "heat_reference_view_csr: needs a mesh from heat_open / heat_mesh_set");
    if (n < 0 || (row_ptr[n] > 0 && (!col || !val))) HEAT_FAIL(2, "heat_reference_view_csr: bad matrix");
    const int64_t N = m.num_nodes;
    std::vector<char> in_set((size_t)N, 0);
    for (const auto &s : m.nodesets)
        for (int64_t g : s.second) in_set[(size_t)g] = 1;
    int64_t ndof = 0;
    for (int64_t g = 0; g < N; ++g) ndof += !in_set[(size_t)g];
    if (ndof != n) HEAT_FAIL(2, "heat_reference_view_csr: the mesh has %lld unknowns, the matrix %lld rows", (long long)ndof, (long long)n);
    for (int64_t k = 0; k < n; ++k)
        if (red2orig[k] < 0 || red2orig[k] >= N || in_set[(size_t)red2orig[k]] || (k > 0 && red2orig[k] <= red2orig[k - 1]))
            HEAT_FAIL(2, "heat_reference_view_csr: red2orig is not the ascending list of the mesh's unknowns");
    // D3: does node g have a neighbour outside every nodeset?
    std::vector<char> has_dof_nbr((size_t)N, 0);
    for (int64_t e = 0; e < m.num_elem; ++e) {
        const int32_t *c = m.conn.data() + e * m.npe;
        int nfree = 0;
        for (int a = 0; a < m.npe; ++a) nfree += !in_set[(size_t)c[a]];
        for (int a = 0; a < m.npe; ++a)
            if (nfree - (in_set[(size_t)c[a]] ? 0 : 1) > 0) has_dof_nbr[(size_t)c[a]] = 1;
    }
    const bool d1 = n > 0 && red2orig[n - 1] == N - 1;               // the last mesh node is a DOF: the reference drops it
    const int64_t n_ref = d1 ? n - 1 : n;
    // pass 1: sizes
    int64_t nnz = 0, n_map = 0;
    for (int64_t k = 0; k < n_ref; ++k) {
        if (!has_dof_nbr[(size_t)red2orig[k]]) continue;
        ++n_map;
        bool to_last = false, has_col0 = false;
        for (int64_t q = row_ptr[k]; q < row_ptr[k + 1]; ++q) {
            if (d1 && col[q] == n - 1) to_last = true;
            else { ++nnz; has_col0 |= col[q] == 0; }
        }
        if (to_last && !has_col0) ++nnz;                              // the fold creates entry (k, 0)
    }
    *n_out = n_ref; *nnz_out = nnz; *n_map_out = n_map;
    if (!row_ptr_out) return 0;                                       // size query
    if ((nnz > 0 && (!col_out || !val_out)) || (n_ref > 0 && !b_out) || (n_map > 0 && (!idmap_reduced_out || !idmap_original_out)))
        HEAT_FAIL(2, "heat_reference_view_csr: null output");
    // pass 2: fill (columns stay ascending: column 0 is the smallest)
    int64_t w = 0, wm = 0;
    row_ptr_out[0] = 0;
    for (int64_t k = 0; k < n_ref; ++k) {
        b_out[k] = b[k];
        if (has_dof_nbr[(size_t)red2orig[k]]) {
            idmap_reduced_out[wm] = k; idmap_original_out[wm] = red2orig[k]; ++wm;
            double fold = 0.0;
            bool to_last = false;
            for (int64_t q = row_ptr[k]; q < row_ptr[k + 1]; ++q)
                if (d1 && col[q] == n - 1) { fold += val[q]; to_last = true; }
            const int64_t start = w;
            bool has_col0 = false;
            for (int64_t q = row_ptr[k]; q < row_ptr[k + 1]; ++q) {
                if (d1 && col[q] == n - 1) continue;
                if (col[q] == 0) { has_col0 = true; col_out[w] = 0; val_out[w] = val[q] + fold; ++w; }
                else { col_out[w] = col[q]; val_out[w] = val[q]; ++w; }
            }
            if (to_last && !has_col0) {                               // insert (k, 0) in front of the row
                for (int64_t q = w; q > start; --q) { col_out[q] = col_out[q - 1]; val_out[q] = val_out[q - 1]; }
                col_out[start] = 0; val_out[start] = fold; ++w;
            }
        }
        row_ptr_out[k + 1] = w;
    }
    return 0;
}
