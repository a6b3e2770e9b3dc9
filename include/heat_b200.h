/*
 * heat_b200.h — C ABI of the B200-native steady-state heat path.
 *
 * This is the drop-in boundary for ONE path of LouisJenkinsCS/Domain-Decomposed-PDE-Solver:
 *     mesh (Exodus-II) -> Dirichlet elimination -> CSR assembly -> fp64 SpMV -> PCG -> nodal field
 * Every entry point names the reference interface it replaces (file:line in /root/reference).
 * The reference exposes a header-only C++ class (ExodusIO::IO) with Tpetra types in its
 * signatures; the C++ mirror of that class over this ABI is include/ExodusIO_b200.hpp.
 *
 * Conventions: all functions return 0 on success, non-zero on failure (the reference returns
 * bool + perror/cerr; heat_last_error() holds the message).  Plain pointers and sizes only.
 * Pointers named *_host are host memory, *_dev device memory of the context's GPU, and plain
 * `void*` data pointers accept either (resolved with cudaPointerGetAttributes).  One heat_ctx
 * per process == per GPU == per "MPI rank" of the reference; calls on a ctx are not
 * thread-safe (neither is ExodusIO::IO).  There is NO CPU fallback: without a CUDA device every
 * compute entry point fails loudly.
 */
#ifndef HEAT_B200_H
#define HEAT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HEAT_B200_VERSION 100

typedef struct heat_ctx    heat_ctx;     /* ExodusIO::IO state, ExodusIO.hpp:2081-2098          */
typedef struct heat_matrix heat_matrix;  /* Teuchos::RCP<Tpetra::CrsMatrix<>>   (local rows)     */
typedef struct heat_vector heat_vector;  /* Teuchos::RCP<Tpetra::MultiVector<>> (1 column)       */

/* operator assembled on the reference's sparsity pattern */
enum { HEAT_OP_GRAPH_LAPLACIAN = 0,      /* reference-exact, ExodusIO.hpp:116-127, :591-608      */
       HEAT_OP_P1_FEM = 1 };             /* north-star P1 stiffness, same pattern                */
/* Krylov solver (reference: Belos "GMRES", BelosMueLuSolver.cpp:106; north-star: CG)            */
enum { HEAT_SOLVER_CG = 0,               /* classical Belos-style PCG, 2 reductions / iteration  */
       HEAT_SOLVER_CG_SINGLE_REDUCE = 1, /* Chronopoulos-Gear PCG, 1 reduction / iteration       */
       HEAT_SOLVER_GMRES = 2 };          /* restarted GMRES(m), RIGHT preconditioned: the literal
                                            reference solver (BelosMueLuSolver.cpp:102-109)      */
/* preconditioner (reference: Ifpack2 "ILUT", BelosMueLuSolver.cpp:93; north-star: Jacobi/Chebyshev) */
enum { HEAT_PREC_NONE = 0, HEAT_PREC_JACOBI = 1, HEAT_PREC_CHEBYSHEV = 2,
       HEAT_PREC_ILU0 = 3 };             /* ILU(0) of the rank-local block (stands in for Ifpack2
                                            ILUT with its default list; CG or GMRES)             */
/* partitioner used by heat_assemble when nranks > 1                                             */
enum { HEAT_PART_CONTIGUOUS = 0,         /* Tpetra uniform contiguous map, ExodusIO.hpp:252      */
       HEAT_PART_METIS_KWAY = 1,         /* METIS k-way on the matrix row graph (role of Zoltan2
                                            "parmetis", ExodusIO.hpp:644-656)                    */
       HEAT_PART_SLAB = 2 };             /* synthetic cubes: slabs along the slowest index       */

const char *heat_last_error(void);
int  heat_version(void);
int  heat_device_count(void);            /* CUDA devices visible; 0 => every compute call fails  */
/* number of this library's own kernels launched by this process so far (CUB/NCCL not counted)   */
unsigned long long heat_kernel_launches(void);

/* ---- lifecycle: IO() / open / create / ~IO  (ExodusIO.hpp:85, :88-100, :103-114, :2072-2079) */
int  heat_ctx_create(int device, heat_ctx **out);
int  heat_ctx_set_stream(heat_ctx *ctx, void *cuda_stream);   /* cudaStream_t; default: own stream */
/* Output conventions of a particular reference BUILD (defaults 8 / 0 = fp64 records, defect D2 fixed):
 *   word_size 4           the reference opens and creates its files with cpu/io word size sizeof(real_t)
 *                         (ExodusIO.hpp:89-90, :104-105); with a METIS whose real_t is float — the usual build —
 *                         coordinates, distribution factors, time values and nodal results are written as float32.
 *   largest_nodeset_id!=0 a node that belongs to several nodesets shows the LARGEST id in the written field, as the
 *                         reference's writeSolution does (:1983-1989), instead of the id its right-hand side used.
 * Call before heat_decompose.  Pure host state.                                                                   */
int  heat_ctx_set_output(heat_ctx *ctx, int word_size, int largest_nodeset_id);
int  heat_open(heat_ctx *ctx, const char *path, int read_only);
int  heat_create(heat_ctx *ctx, const char *path);
int  heat_close(heat_ctx *ctx);                               /* closes files, frees everything   */

/* ---- meshes that do not come from a file ---------------------------------------------------- */
/* Explicit mesh from host memory (what open()+the reads at ExodusIO.hpp:143-192,:342-359 yield):
 * conn is 0-based [num_elem][npe]; nodeset s has id ns_ids[s] and nodes
 * ns_nodes[ns_ptr[s] .. ns_ptr[s+1]) (0-based).  z may be NULL for 2-D meshes.                   */
int  heat_mesh_set(heat_ctx *ctx, int64_t num_nodes, int num_dim, const double *x_host,
                   const double *y_host, const double *z_host, int64_t num_elem, int npe,
                   const int32_t *conn_host, int num_node_sets, const int64_t *ns_ids,
                   const int64_t *ns_ptr, const int64_t *ns_nodes_host);
/* Synthetic structured Kuhn tet cube (BASELINE.json configs[2..4]):
 * nx*ny*nz nodes on [-5,5]^3, nodeset 1000 on i==0, 100 on i==nx-1.  `explicit_mesh` != 0
 * materialises coordinates + connectivity on the device and runs the general assembly kernels;
 * 0 uses the analytic-connectivity kernels (needed for 512^3 and for per-GPU slabs).            */
int  heat_mesh_cube(heat_ctx *ctx, int nx, int ny, int nz, int explicit_mesh);
/* ex_get_ids(EX_NODE_SET) (ExodusIO.hpp:172): nodeset ids, ascending; ids_out may be NULL        */
int  heat_mesh_nodeset_ids(const heat_ctx *ctx, int *count, int64_t *ids_out);

/* ---- multi-GPU plumbing (replaces MPI_COMM_WORLD; one process per GPU) ---------------------- */
#define HEAT_COMM_ID_BYTES 128
int  heat_comm_unique_id(char id_out[HEAT_COMM_ID_BYTES]);    /* rank 0 makes it; broadcast it    */
int  heat_comm_init(heat_ctx *ctx, int rank, int nranks, const char id[HEAT_COMM_ID_BYTES]);
int  heat_comm_rank(const heat_ctx *ctx, int *rank, int *nranks);

/* ---- assemble  (IO::assemble, ExodusIO.hpp:128-723) ------------------------------------------ *
 * Builds this rank's rows of A (reduced system, Dirichlet nodes eliminated and moved to B),
 * X (zero; the reference's unseeded random X is not reproducible) and B, all on
 * the device.  Collective over the ranks of heat_comm_init.                                      */
int  heat_assemble(heat_ctx *ctx, int op_mode, int partitioner, heat_matrix **A, heat_vector **X,
                   heat_vector **B);

/* ---- getMatrix  (IO::getMatrix, ExodusIO.hpp:733-1489) + power method (ExodusMatrixTest.cpp) -- *
 * The graph Laplacian (or P1 stiffness) of the WHOLE mesh: one row per node, nodesets not applied
 * (singular, as the reference says at :729-731).  Rows are distributed by ELEMENT partition
 * (METIS_PartMeshDual, the serial counterpart of ParMETIS_V3_PartMeshKway at :919) followed by the
 * reference's node-ownership rule (:1191-1295).  Collective; needs a mesh from heat_open/mesh_set. */
int  heat_get_matrix(heat_ctx *ctx, int op_mode, heat_matrix **A);
/* the rule alone (pure host): owner[v] = the part in whose elements node v has the most distinct
 * neighbours, ties to the lowest part; epart[e] in [0, nparts).                                  */
int  heat_node_owners(int64_t num_nodes, int64_t num_elem, int npe, const int32_t *conn_host,
                      const int64_t *epart_host, int nparts, int32_t *owner_out);
/* nodeSetMap of getMatrix (:1447-1466): nodes (0-based) of nodeset `set_id` whose rows this rank
 * owns, ascending.  nodes_out may be NULL to query *count.                                       */
int  heat_matrix_owned_nodeset(const heat_matrix *A, int64_t set_id, int64_t *count, int64_t *nodes_out);
/* PowerMethod::run (ExodusMatrixTest.cpp:56-129): q = z/||z||, z = A q, lambda = q.z; the residual
 * ||A q - lambda q||_2 is evaluated every 50 iterations and at the last one and compared with
 * `tolerance`.  The reference starts from an unseeded random vector; here z0 is the counter-based
 * U(-1,1) vector of heat_vector_fill_hash(seed).  iters = index of the iteration that converged
 * (the reference's "Converged after <iter> iterations"), or niters.                              */
typedef struct {
    double lambda, residual; int iters, converged; double solve_ms;
    /* optional (in): room for the reports the reference prints every 50 iterations (:103-111);
     * report k holds {iteration, lambda, residual} in report_buf[3k..3k+2]; (out) report_count      */
    double *report_buf; int report_capacity, report_count;
} heat_power_info;
int  heat_power_method(heat_ctx *ctx, heat_matrix *A, int niters, double tolerance, uint64_t seed,
                       heat_power_info *info);

/* ---- solve  (belosSolver, BelosMueLuSolver.cpp:87-139) --------------------------------------- *
 * Stops when ||r||_2/||r0||_2 <= tol (status test before each iteration, as Belos) or after
 * max_iters iterations.  Returns 0 also when max_iters was hit; *converged tells which.          */
typedef struct {
    int    solver;            /* HEAT_SOLVER_*                                                    */
    int    prec;              /* HEAT_PREC_*                                                      */
    int    max_iters;         /* reference default 300 (BelosMueLuSolver.cpp:149)                 */
    double tol;               /* reference default 1e-14 (:151); north-star 1e-10                 */
    int    cheb_degree;       /* Ifpack2 "chebyshev: degree" (default 1)                          */
    double cheb_lambda_max;   /* "chebyshev: max eigenvalue"; <=0 => 10 power iterations          */
    double cheb_ratio;        /* "chebyshev: ratio eigenvalue" (default 30)                       */
    int    check_every;       /* host convergence poll period in iterations (0 => 32)             */
    int    gmres_restart;     /* GMRES: Belos "Num Blocks" = iterations per restart cycle (300)    */
} heat_solve_opts;
typedef struct {
    int    iters;
    int    converged;
    double achieved_tol;      /* ||r||/||r0|| (Belos achievedTol(), BelosMueLuSolver.cpp:121)     */
    double r0_norm;
    double solve_ms;          /* device time of the iteration loop (CUDA events)                  */
} heat_solve_info;
void heat_solve_opts_default(heat_solve_opts *o);
int  heat_solve(heat_ctx *ctx, heat_matrix *A, heat_vector *X, const heat_vector *B,
                const heat_solve_opts *opts, heat_solve_info *info);
/* Same solve with the reference's per-iteration field output (BelosMueLuSolver.cpp:113-133 calls
 * io.writeSolution(X, i) after every pass): ONE Krylov run; every `write_every` iterations, and at
 * the end, the current iterate is written as time step first_timestep, first_timestep+1, ...
 * through heat_write_solution (collective; needs heat_create + heat_decompose on rank 0).         */
int  heat_solve_trajectory(heat_ctx *ctx, heat_matrix *A, heat_vector *X, const heat_vector *B,
                           const heat_solve_opts *opts, int write_every, int first_timestep,
                           heat_solve_info *info, int *frames_written);
/* Same, through host buffers (the end-to-end path): b_host[n_owned] is copied in, x_host holds
 * x0 on entry and the solution on exit.                                                          */
int  heat_solve_host(heat_ctx *ctx, heat_matrix *A, const double *b_host, double *x_host,
                     const heat_solve_opts *opts, heat_solve_info *info);

/* A SEQUENCE of such solves with one matrix (a loop around belosSolver with a new B per pass: time stepping, many
 * right-hand sides).  System k takes b_hosts[k], starts from x0_hosts[k] (x0_hosts or x0_hosts[k] NULL: zero) and
 * leaves its solution in x_hosts[k]; infos (may be NULL) gets one record per system.  Two staging sets and two copy
 * streams: the inputs of system k+1 go up and the solution of system k-1 goes down WHILE system k is solved, so the
 * stream runs at the device-resident rate.  Pinned host buffers are needed for the overlap (pageable ones work, serially). */
int  heat_solve_host_batch(heat_ctx *ctx, heat_matrix *A, int count, const double *const *b_hosts,
                           const double *const *x0_hosts, double *const *x_hosts, const heat_solve_opts *opts,
                           heat_solve_info *infos);

/* y = A x  (Tpetra::CrsMatrix::apply; explicit use at ExodusMatrixTest.cpp:101) incl. halo.      */
int  heat_spmv(heat_ctx *ctx, heat_matrix *A, heat_vector *x, heat_vector *y);
/* The same product through the PEER-MEMORY halo path, i.e. the very SpMV launch the multi-GPU CG loop makes
 * (Tpetra's Import inside apply, replaced by NVLink peer stores + epoch flags, csrc/peer.cuh) — heat_spmv goes
 * through the NCCL halo.  Collective; fails with code 52 where the peer path is unavailable (one rank,
 * HEAT_COMM=nccl, no CUDA IPC).  The kernel is launched `repeat` (>= 1) times on the same input;
 * *kernel_ms (may be NULL) = average device time of launches 2..repeat, *xy_global (may be NULL) = sum_i x_i y_i
 * over all ranks.  Parity and measurement hook.                                                    */
int  heat_spmv_peer(heat_ctx *ctx, heat_matrix *A, heat_vector *x, heat_vector *y, int repeat,
                    double *xy_global, double *kernel_ms);
/* fixed number of CG iterations with no convergence exit (bench hook; same kernels as solve)    */
int  heat_cg_iterations(heat_ctx *ctx, heat_matrix *A, heat_vector *X, const heat_vector *B,
                        const heat_solve_opts *opts, int iters, heat_solve_info *info);

/* ---- output  (IO::decompose ExodusIO.hpp:1496-1969, IO::writeSolution :1972-2070) ------------ */
int  heat_decompose(heat_ctx *ctx, int partitions);           /* rank 0 only, after heat_create   */
int  heat_write_solution(heat_ctx *ctx, const heat_vector *X, int timestep);   /* collective      */
/* the file half of writeSolution alone (ExodusIO.hpp:2027-2069): one dense nodal array as time step
 * `timestep` of "Steady-State Heat Solution".  Pure host code; the first call defines the result
 * variables and rewrites the file, later calls touch only their record.                           */
int  heat_write_nodal_field(heat_ctx *ctx, const double *field_host, int64_t num_nodes, int timestep);
/* dense nodal field writeSolution would store (DOF nodes from X, nodeset nodes = their id)       */
int  heat_nodal_field(heat_ctx *ctx, const heat_vector *X, double *field_host, int64_t num_nodes);
/* the scatter half alone, for a solution held on the HOST in reduced-id order (what heat_solve_host returns on one
 * rank, heat_matrix_export_red2orig gives the order): with heat_write_nodal_field it completes the host-buffer path
 * solve -> field -> file without a device vector.  Pure host code.                                  */
int  heat_scatter_nodal_field(heat_ctx *ctx, const double *x_reduced_host, int64_t n_global, double *field_host,
                              int64_t num_nodes);
/* The system the REFERENCE's IO::assemble returns for the context's mesh (one rank), derived from the FIXED system
 * this library assembles (heat_matrix_export_csr / heat_vector_get / heat_matrix_export_red2orig): the rows of
 * unknowns without an unknown neighbour are never inserted and have no id-map entry (defect D3, ExodusIO.hpp:
 * 380-386, :591), and if the last mesh node is an unknown it is dropped and the entries that referenced it are
 * summed into column 0 (defect D1, :220, :440, :598).  For diffing against a Trilinos run; held to the reference's
 * own output by tests/test_reference_pins.py.  Call with row_ptr_out == NULL for the sizes (n_out, nnz_out,
 * n_map_out), then with arrays of n_out+1 / nnz_out / nnz_out / n_out / n_map_out / n_map_out entries.  Host only. */
int  heat_reference_view_csr(heat_ctx *ctx, int64_t n, const int64_t *row_ptr, const int32_t *col, const double *val,
                             const double *b, const int64_t *red2orig, int64_t *n_out, int64_t *nnz_out,
                             int64_t *row_ptr_out, int32_t *col_out, double *val_out, double *b_out,
                             int64_t *n_map_out, int64_t *idmap_reduced_out, int64_t *idmap_original_out);
/* METIS_PartMeshDual with the reference's arguments (ExodusIO.hpp:1615); arrays are int64        */
int  heat_decompose_partition(heat_ctx *ctx, int partitions, int64_t *objval, int64_t *epart_host,
                              int64_t *npart_host);

/* ---- inspection / parity hooks (host copies) ------------------------------------------------- */
typedef struct {
    int64_t num_nodes, num_elem, n_global, nnz_global;
    int64_t n_owned, n_ghost, nnz_local;
    int32_t max_row_len, npe, num_dim, num_node_sets;
    int32_t rank, nranks, n_neighbors, sell_chunk;
    int64_t sell_padded_nnz, n_boundary_slices, n_slices;
    double  assemble_ms;      /* device time of pattern+values+SELL (CUDA events)                 */
    int32_t peer_path;        /* 1 once the NVLink peer-memory halo/all-reduce path is set up      */
    int32_t col_index_bytes;  /* widest entry of the SpMV column stream, chosen per 64-row slice: 1 (every slice stores a
                                 1-byte index into its table of distinct col-row offsets: structured numberings), 2
                                 (slices that fit no table store int16 col-row deltas: any mesh of bandwidth < 32768),
                                 4 (int32 column ids)                                                  */
    double  assemble_fill_ms; /* the matrix-fill kernel alone (analytic cubes: cube_sell_kernel, straight into the
                                 SpMV format; explicit meshes: values_kernel), CUDA events               */
    double  asm_phase_ms[4];  /* explicit-mesh assembly by phase: node->element radix sort, pattern count,
                                 pattern fill, values (0 for the analytic cube path)                     */
    int32_t csr_resident;     /* 1 if a CSR copy is resident (always for explicit meshes; analytic cubes are
                                 assembled straight into SELL and build the CSR only for an export / ILU)  */
    int32_t reserved0;
    int64_t matrix_bytes;     /* device bytes held by the matrix arrays (CSR + SELL + tables + diagonal)  */
} heat_matrix_info;
int  heat_matrix_get_info(const heat_matrix *A, heat_matrix_info *info);
/* local CSR: row_ptr[n_owned+1] (int64), col[nnz_local] LOCAL column ids (int32), val.           */
int  heat_matrix_export_csr(const heat_matrix *A, int64_t *row_ptr_host, int32_t *col_host,
                            double *val_host);
/* maps: owned_gids[n_owned], ghost_gids[n_ghost], ghost_owner[n_ghost] (reduced global ids)      */
int  heat_matrix_export_maps(const heat_matrix *A, int64_t *owned_gids_host, int64_t *ghost_gids_host,
                             int32_t *ghost_owner_host);
/* send plan: for neighbour slot s (0..n_neighbors): rank nbr_rank[s], local rows
 * send_idx[send_ptr[s]..send_ptr[s+1]) are sent, recv_ptr[s]..recv_ptr[s+1] ghosts received.     */
int  heat_matrix_export_plan(const heat_matrix *A, int32_t *nbr_rank_host, int64_t *send_ptr_host,
                             int32_t *send_idx_host, int64_t *recv_ptr_host);
/* reduced global id -> 0-based original node (globalIDMap, ExodusIO.hpp:572-576) of owned rows   */
int  heat_matrix_export_red2orig(const heat_matrix *A, int64_t *red2orig_host);
/* ILU(0) factors (after a solve with HEAT_PREC_ILU0): lu_host[nnz_local] on the pattern of
 * heat_matrix_export_csr — strictly lower part = L (unit diagonal), the rest = U; ghost-column
 * entries keep the values of A.  Also the number of dependency levels of the two sweeps.        */
int  heat_matrix_export_ilu0(const heat_matrix *A, double *lu_host, int *n_levels_lower, int *n_levels_upper);
int  heat_matrix_free(heat_matrix *A);

int  heat_vector_create(heat_ctx *ctx, const heat_matrix *A, heat_vector **out);
int64_t heat_vector_size(const heat_vector *v);               /* owned entries                    */
void *heat_vector_device_ptr(heat_vector *v);                 /* fp64, owned entries then ghosts  */
int  heat_vector_set(heat_ctx *ctx, heat_vector *v, const void *src, int64_t count);
int  heat_vector_get(heat_ctx *ctx, const heat_vector *v, void *dst, int64_t count);
int  heat_vector_fill(heat_ctx *ctx, heat_vector *v, double value);
/* x[i] = U(-1,1) from a counter-based generator keyed on the GLOBAL reduced id:
 * identical values for any GPU count.                                                            */
int  heat_vector_fill_hash(heat_ctx *ctx, const heat_matrix *A, heat_vector *v, uint64_t seed);
int  heat_vector_free(heat_vector *v);

/* ---- host-side partition / ghost-map plan (pure CPU; also used internally by heat_assemble) -- *
 * From a global CSR pattern (reduced ids, diagonal included or not) and part[row] in [0,nranks):
 * Tpetra conventions — owned rows ascending; ghost columns grouped by owner rank ascending,
 * ascending gid inside an owner; send list p->q = q's ghosts owned by p in q's ghost order.
 * Call with NULL outputs to get the sizes.                                                       */
typedef struct { int64_t n_owned, n_ghost; int32_t n_neighbors; int64_t n_send; } heat_plan_sizes;
int  heat_plan_build(int64_t n_global, const int64_t *row_ptr, const int32_t *col, const int32_t *part,
                     int nranks, int rank, heat_plan_sizes *sizes, int64_t *owned_gids,
                     int64_t *ghost_gids, int32_t *ghost_owner, int32_t *nbr_rank, int64_t *send_ptr,
                     int64_t *send_gids, int64_t *recv_ptr);
/* partition a global row graph exactly as heat_assemble does                                     */
int  heat_partition_rows(int64_t n_global, const int64_t *row_ptr, const int32_t *col, int partitioner,
                         int nranks, int32_t *part_out);

#ifdef __cplusplus
}
#endif
#endif /* HEAT_B200_H */
