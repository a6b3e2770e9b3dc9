// kernels.cuh — launch interfaces of the sm_100a kernels (definitions in *.cu).
#pragma once
#include "common.cuh"
#include "peer.cuh"

namespace heat {

// One record per CG iteration k: gamma_k = r_k.z_k, delta_k (single-reduce: w_k.u_k),
// rr_k = ||r_k||^2, alpha_k.  Records are zero-initialised; an unwritten record (rr == 0) reads
// as "converged", which freezes every later launch once the stopping test fires.
struct CgRec { double rz, delta, rr, alpha; };

struct CgGate {               // stopping test evaluated by every block of every CG kernel
    const CgRec *H;           // iteration records
    const double *S;          // scalars (S_TOL2)
    const int *I;             // I_STATUS
    int it;                   // iteration this launch belongs to
    // Traversal direction of the iteration's kernels (0: all forward).  Consecutive kernels of the loop sweep the
    // vectors in OPPOSITE directions — SpMV forward, update_xr backward, update_p forward, next SpMV backward, ... —
    // so each kernel starts on the ~100 MB its predecessor touched last, which are still in the 126 MB L2.  It is
    // the last-level-cache blocking a stream of whole-vector sweeps allows; it matters when a GPU's share of the
    // vectors is comparable to the L2 (strong scaling: 134 MB per vector on 8 GPUs), not at 1 GB per vector.
    //   1: even iteration (SpMV forward, update_xr backward, update_p forward)   2: odd iteration (the mirror image)
    int dir;
    int pdl;                  // host side: launch with programmatic stream serialization (device_utils.cuh)
};
__host__ __device__ inline bool spmv_backward(const CgGate &g) { return g.dir == 2; }
__host__ __device__ inline bool xr_backward(const CgGate &g) { return g.dir == 1; }
__host__ __device__ inline bool p_backward(const CgGate &g) { return g.dir == 2; }

struct DotOut {               // where a fused dot product is reduced to
    double *partials; int part_offset; int total_blocks; int *counter; double *out;
    bool with_yy = false;     // SpMV only: also reduce sum_i y_i^2 into out[1]
};

struct SellDict {             // column stream of the SELL matrix (sell.cu)
    const uint8_t *cstream;   // per slice (SliceMeta::cbase64): 1-byte table indices, int16 deltas or int32 column ids
    const int32_t *tab;       // [n_slices][tpad] distinct (col - row) offsets of each table-indexed slice
    int tpad;                 // table stride, a multiple of 4 entries (16 bytes: one TMA granule)
};

struct SpmvPeer {             // peer-memory mode of the SpMV (single launch over interior+boundary slices)
    bool on = false;
    int64_t n_interior = 0;   // leading entries of the slice list that never touch a ghost column
    PeerHalo halo;            // flags to wait for before the first boundary slice
    PeerRed red;              // where the fused dot goes (all ranks' inboxes)
    unsigned long long seq_out = 0;
    int *I = nullptr;
};

#ifdef HEAT_PEER_TRACE
int trace_set_cg(TraceBuf *buf);
int trace_set_spmv(TraceBuf *buf);
#endif
// ---- sell.cu ----
int sell_from_csr(heat_matrix *A, cudaStream_t st);
int sell_finish_lists(heat_matrix *A, const int32_t *h_flags, cudaStream_t st);   // interior/boundary slice lists
int sell_build_meta(heat_matrix *A, cudaStream_t st);  // packed SliceMeta in processing order (after the column stream is chosen)
int sell_cidx_mode();                                  // HEAT_SPMV_CIDX: 0 int32 only, 1 compact (default), 2 prefer int16
int sell_to_csr(heat_matrix *A, cudaStream_t st);      // builds row_ptr/col/val from the SELL arrays if absent
// ---- spmv.cu ----
// y = A x over entries [first, first + n_list) of A->slice_meta (processing order: interior slices, then
// boundary slices).  gate.H == nullptr disables the CG stopping test; dot.out == nullptr disables the
// fused sum_i y_i * x_i.
int launch_spmv(const heat_matrix *A, const double *x, double *y, int64_t first,
                int64_t n_list, CgGate gate, DotOut dot, int grid, cudaStream_t st);
int launch_spmv_peer(const heat_matrix *A, const double *x, double *y, CgGate gate, DotOut dot, SpmvPeer peer,
                     int grid, cudaStream_t st);
bool spmv_peer_supported();
bool spmv_compact_supported();                         // the selected SpMV variant can stream the compact column formats
int spmv_grid(int64_t n_list, int sm_count);
// ---- cg.cu ----
int launch_cg_init(int64_t n, const double *b, const double *ax, const double *dinv, double *r,
                   double *p_or_u, CgRec *H, double *partials, int *counter, int grid, cudaStream_t st);
int launch_cg_update_xr(int64_t n, double *x, double *r, const double *p, const double *ap,
                        const double *dinv, CgGate gate, CgRec *H, double *S, int *I,
                        double *partials, int *counter, int grid, cudaStream_t st);
int launch_cg_update_p(int64_t n, double *p, const double *r, const double *dinv, CgGate gate,
                       int grid, cudaStream_t st);
int launch_cg_update_xr_peer(int64_t n, double *x, double *r, const double *p, const double *ap, const double *dinv,
                             CgGate gate, CgRec *H, double *S, int *I, double *partials, int *counter, PeerRed pr,
                             unsigned long long seq_in, unsigned long long seq_out, int grid, cudaStream_t st);
int launch_cg_update_p_peer(int64_t n, double *p_out, const double *p_in, const double *r, const double *dinv, const double *z,
                            CgGate gate, CgRec *H, int *I, PeerRed pr, unsigned long long seq_in, PeerPush push,
                            int grid, cudaStream_t st);
// Chebyshev-PCG on the peer-memory path (cg.cu): `last` = this launch carries the r.z / r.r reduction
int launch_cheb_xr_first_peer(bool last, int64_t n, double *x, const double *r_in, double *r_out, const double *p, const double *ap,
                              const double *dinv, double inv_theta, double *w, double *z_out, CgGate gate, CgRec *H, double *S, int *I,
                              double *partials, int *counter, PeerRed pr, unsigned long long seq_in, unsigned long long seq_out,
                              PeerPush push, int grid, cudaStream_t st);
int launch_cheb_step_peer(bool last, int64_t n, const double *dinv, const double *r, const double *az, double c1, double c2,
                          const double *w_in, double *w_out, const double *z_in, double *z_out, CgGate gate, double *S, double *partials,
                          int *counter, PeerRed pr, unsigned long long seq_out, PeerPush push, int grid, cudaStream_t st);
int launch_halo_push(const double *x, PeerPush push, cudaStream_t st);
int launch_cg_fused_update(int64_t n, double *x, double *r, double *p, double *s, double *u,
                           const double *w, const double *dinv, CgGate gate, CgRec *H, int *I,
                           double *partials, int *counter, int grid, cudaStream_t st);
// generic pieces for the Chebyshev path / residual
int launch_axpby(int64_t n, double a, const double *x, double b, double *y, int grid, cudaStream_t st);
int launch_dot2(int64_t n, const double *a, const double *b, const double *c, const double *d,
                double *out_ab, double *out_cd, double *partials, int *counter, int grid, cudaStream_t st);
int launch_cheb_first(int64_t n, const double *dinv, const double *r, double inv_theta, double *w,
                      double *z, CgGate gate, int grid, cudaStream_t st);
int launch_cheb_step(int64_t n, const double *dinv, const double *r, const double *az, double c1,
                     double c2, double *w, double *z, CgGate gate, int grid, cudaStream_t st);
int launch_cg_xr_plain(int64_t n, double *x, double *r, const double *p, const double *ap,
                       CgGate gate, double *S, int *I, int grid, cudaStream_t st);
int launch_cg_p_plain(int64_t n, double *p, const double *z, CgGate gate, int grid, cudaStream_t st);
int launch_cg_dots(int64_t n, const double *r, const double *z, CgGate gate, CgRec *H, int *I,
                   double *partials, int *counter, int grid, cudaStream_t st);
int launch_pm_scale(int64_t n, const double *z, const double *zz, double *q, int grid, cudaStream_t st);
int launch_pm_resid(int64_t n, const double *z, const double *q, const double *lambda, double *out,
                    double *partials, int *counter, int grid, cudaStream_t st);
int launch_fill(int64_t n, double *x, double v, cudaStream_t st);
int launch_fill_hash(int64_t n, double *x, const int64_t *gids, int64_t gid0, uint64_t seed, cudaStream_t st);
int launch_gather(int64_t n, const double *x, const int32_t *idx, double *out, cudaStream_t st);
int launch_extract_diag(const heat_matrix *A, cudaStream_t st);
int vec_grid(int64_t n, int sm_count);

}  // namespace heat
