/*
 * heat_oracle.h — CPU ORACLE (TEST INFRASTRUCTURE ONLY; never linked into the product).
 *
 * A plain-C restatement of the reference's steady-state heat path:
 *   IO::assemble           /root/reference/ExodusIO.hpp:128-723
 *   belosSolver (as CG)    /root/reference/BelosMueLuSolver.cpp:87-139
 *   IO::writeSolution      /root/reference/ExodusIO.hpp:1972-2070  (field scatter only)
 *   PowerMethod::run       /root/reference/ExodusMatrixTest.cpp:56-129
 *   (IO::getMatrix's ownership rule, ExodusIO.hpp:1089-1295, is restated in oracle.py)
 *
 * PARITY STATUS.
 *   oracle_assemble / get_matrix (ExodusIO.hpp): PINNED against the reference's own code.  oracle/_ref/
 *   ref_driver compiles /root/reference/ExodusIO.hpp, unmodified, against single-rank stand-ins for its
 *   third-party headers (oracle/ref_shim/README.md) and runs it; what it computed on all 17 meshes of the
 *   reference's data/ (A, B, id map, getMatrix, decompose and writeSolution records) is committed as
 *   tests/golden/ref_pins.json and reproduced bit for bit by tests/test_reference_pins.py (this oracle has
 *   the FIXED semantics; the one documented transformation to the reference's off-by-one is pins.apply_d1).
 *   oracle_pcg / gmres / ilu0 / power method (Belos, Ifpack2, Tpetra: third party, absent, version
 *   unrecorded): "parity unpinned" against a reference binary; the reference ships no golden vectors, tests
 *   or outputs of its own.  Pinned instead against
 *   (i)   the hand-checkable 3x3 system of data/rectangle-tris-boundary.exo,
 *   (ii)  an independent numpy/scipy restatement (oracle.assemble_np) incl. sparse direct solves,
 *   (iii) the analytic P1 solution T = 550 - 90 x on data/tet-cube-heat.exo,
 *   (iv)  METIS known answers with the library the container ships.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.
 */
#ifndef HEAT_ORACLE_H
#define HEAT_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORACLE_GRAPH_LAPLACIAN = 0, ORACLE_P1_FEM = 1 };
enum { ORACLE_PREC_NONE = 0, ORACLE_PREC_JACOBI = 1, ORACLE_PREC_CHEBYSHEV = 2 };

typedef struct {
    int64_t  num_nodes;   /* N: nodes of the mesh                                    */
    int64_t  n;           /* DOF rows (nodes in no nodeset)                          */
    int64_t  nnz;
    int64_t *row_ptr;     /* [n+1]                                                   */
    int32_t *col;         /* [nnz] reduced ids, ascending inside each row            */
    double  *val;         /* [nnz]                                                   */
    double  *b;           /* [n]                                                     */
    int64_t *red2orig;    /* [n] reduced id -> 0-based original node (globalIDMap)   */
    int32_t  max_row;     /* longest row                                             */
} oracle_system;

/* ExodusIO.hpp:216-252 (elimination), :340-386 (adjacency), :591-608 (values), :671-687 (B).
 * node_bc[g] is NaN for a DOF node, else the prescribed temperature of g (= the LOWEST id of a
 * nodeset containing g, the reference's RHS rule at :676-681).  conn is 0-based, npe nodes per
 * element, every element a clique.  mode P1_FEM needs npe==4 (tets, x/y/z) or npe==3 (tris, x/y).
 * FIXED semantics: every non-nodeset node is a DOF (no D1 off-by-one); a DOF row
 * with no DOF neighbour keeps its diagonal (D3).                                              */
int oracle_assemble(int64_t num_nodes, const double *x, const double *y, const double *z,
                    int64_t num_elem, int npe, const int32_t *conn, const double *node_bc,
                    int mode, oracle_system *out);
void oracle_system_free(oracle_system *s);

/* Appendix E of SURVEY.md: structured Kuhn tet cube on [-5,5]^3; node g = i + nx*(j + ny*k);
 * element id = 6*cell + perm, cell = ci + (nx-1)*(cj + (ny-1)*ck), perms of (x,y,z) in
 * lexicographic order; node_bc = 1000 on i==0, 100 on i==nx-1 (mirrors tet-cube-heat.exo).
 * Caller allocates x,y,z,node_bc [nx*ny*nz] and conn [6*(nx-1)*(ny-1)*(nz-1)*4].               */
void oracle_cube_mesh(int nx, int ny, int nz, double *x, double *y, double *z,
                      int32_t *conn, double *node_bc);

/* y = A x, each row accumulated left-to-right in column order with fma() (bit-reproducible). */
void oracle_spmv(int64_t n, const int64_t *row_ptr, const int32_t *col, const double *val,
                 const double *x, double *y);

/* Belos-style classical PCG: status test first, relative to ||r0||.
 * prec: NONE / JACOBI (z = D^-1 r) / CHEBYSHEV (Ifpack2 recurrences; degree, lambda_max, ratio).
 * x holds x0 on entry and the solution on exit.  res_hist (may be NULL) gets ||r_k||/||r_0||
 * for k = 0..iters.  Returns iterations performed, or -1 on breakdown (p.Ap <= 0).           */
int oracle_pcg(int64_t n, const int64_t *row_ptr, const int32_t *col, const double *val,
               const double *b, double *x, int prec, int cheb_degree, double cheb_lambda_max,
               double cheb_ratio, int max_iters, double tol, double *achieved_tol,
               double *res_hist);

/* ExodusIO.hpp:1981-1989 + :2045-2055: dense nodal field; DOF nodes from x, nodeset nodes from
 * node_bc (lowest-id rule, consistent with the RHS — see D2 in SURVEY.md).                    */
void oracle_scatter_field(int64_t num_nodes, const double *node_bc, int64_t n,
                          const int64_t *red2orig, const double *x, double *field);

/* PowerMethod::run (ExodusMatrixTest.cpp:56-129): z = start vector in, last A q out.            */
int oracle_power_method(int64_t n, const int64_t *row_ptr, const int32_t *col, const double *val,
                        double *z, int niters, double tolerance, double *lambda_out,
                        double *residual_out, int *converged);

/* ILU(0) (IKJ order) on the CSR pattern + its triangular solves; restarted right-preconditioned
 * GMRES(m): the literal reference solver (BelosMueLuSolver.cpp:93-109) with ILU(0) for Ifpack2 ILUT. */
enum { ORACLE_PREC_ILU0 = 3 };
int  oracle_ilu0(int64_t n, const int64_t *row_ptr, const int32_t *col, const double *val, double *lu);
void oracle_ilu0_apply(int64_t n, const int64_t *row_ptr, const int32_t *col, const double *lu,
                       const double *v, double *z);
int  oracle_gmres(int64_t n, const int64_t *row_ptr, const int32_t *col, const double *val, const double *b,
                  double *x, int prec, int cheb_degree, double cheb_lambda_max, double cheb_ratio, int restart,
                  int max_iters, double tol, double *achieved_tol, int *converged);

/* The system oracle_assemble builds on oracle_cube_mesh(nx,ny,nz), bit for bit, without the explicit mesh
 * (closed-form Kuhn connectivity; 15-slot rows).  For the CPU baseline at sizes whose explicit mesh does
 * not fit the host.  n < 2^31 rows.                                                             */
int oracle_cube_assemble(int nx, int ny, int nz, int mode, oracle_system *out);

double oracle_pcg_loop_seconds(void);   /* wall seconds of the last oracle_pcg iteration loop (set-up excluded) */
int oracle_num_threads(void);
void oracle_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
