// assemble.cuh — device assembly front ends (definitions in assemble.cu).
#pragma once
#include "common.cuh"

namespace heat {

// one GPU's slab of the synthetic Kuhn cube: owns node planes k0 <= k < k1 (all i in 1..nx-2, all j)
struct CubeGeom {
    int nx, ny, nz, k0, k1;
    int64_t n_owned;    // (nx-2)*ny*(k1-k0)
    int64_t plane;      // (nx-2)*ny DOFs per k-plane
    int64_t ghost_lo;   // plane if k0 > 0 else 0     (ghosts owned by rank-1 come first)
    int64_t ghost_hi;   // plane if k1 < nz else 0
};

struct GeneralAssembler {
    int64_t N = 0, ne = 0, n = 0, nnz = 0;
    int npe = 0;
    bool has_z = false;
    int32_t max_row = 0;
    DevBuf<double> x, y, z, bc, xyz;    // xyz: packed (x, y, z, 0) per node, built by fill_values
    DevBuf<int32_t> conn, red, n2e, gcol;
    DevBuf<int64_t> red2orig, n2e_ptr, grow_ptr;
    // device time by phase (CUDA events on the assembly stream): node->element radix sort, pattern count,
    // pattern fill, values — what bench.py --assemble explicit reports
    float phase_ms[4] = {0.f, 0.f, 0.f, 0.f};

    int upload(const HostMesh &m, const std::vector<double> &node_bc, cudaStream_t st);
    int make_cube(int nx, int ny, int nz, cudaStream_t st);
    // elimination + node->element lists + global pattern (grow_ptr, gcol: global reduced ids)
    int build_pattern(cudaStream_t st);
    // values + rhs of the owned rows.  d_owned == nullptr: rows 0..n_owned-1.  d_lcol == nullptr: the
    // local pattern IS the global pattern (single rank) and only values are written.
    int fill_values(int mode, int64_t n_owned, const int32_t *d_owned, const int32_t *d_g2l,
                    const int64_t *d_lrow_ptr, int32_t *d_lcol, double *d_lval, double *d_b, cudaStream_t st);
};

// CSR-first path (one thread per row; kept as the fallback and for HEAT_CUBE_ASSEMBLY=csr)
int cube_assemble(const CubeGeom &c, int mode, heat_matrix *A, double *d_b, cudaStream_t st);
// default: one warp per slice straight into the SELL arrays (+ byte indices, tables, diagonal, row lengths)
int cube_assemble_sell(const CubeGeom &c, int mode, bool byte_index, heat_matrix *A, double *d_b, cudaStream_t st, bool *done);

}  // namespace heat
