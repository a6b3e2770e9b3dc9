// api.cu — C ABI (include/heat_b200.h): context, mesh, assemble orchestration, vectors, exports.
#include <cmath>
#include <cstdarg>
#include <algorithm>
#include <limits>

#include "assemble.cuh"
#include "comm.cuh"
#include "exodus.hpp"
#include "kernels.cuh"
#include "solve.cuh"

namespace heat {

static thread_local char g_err[1024] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static unsigned long long g_launches = 0;
void count_launch() { ++g_launches; }

void contiguous_partition(int64_t n, int nranks, int32_t *part);
int metis_kway_partition(int64_t n, const int64_t *row_ptr, const int32_t *col, int nranks, int32_t *part);

static bool is_device_ptr(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// slab of rank r when nz planes are dealt to P ranks (first nz % P ranks get one more)
static void slab_range(int nz, int P, int r, int &k0, int &k1) {
    const int base = nz / P, rem = nz % P;
    k0 = r * base + std::min(r, rem);
    k1 = k0 + base + (r < rem ? 1 : 0);
}

// node_bc from nodesets: lowest id wins (RHS rule, ExodusIO.hpp:676-681); hi = highest id (:1983-1989)
void build_node_bc(heat_ctx *ctx) {
    const HostMesh &m = ctx->mesh;
    const double nanv = std::numeric_limits<double>::quiet_NaN();
    ctx->node_bc.assign((size_t)m.num_nodes, nanv);
    ctx->node_bc_hi.assign((size_t)m.num_nodes, nanv);
    for (auto it = m.nodesets.rbegin(); it != m.nodesets.rend(); ++it)      // descending id: lowest written last
        for (int64_t g : it->second) ctx->node_bc[(size_t)g] = (double)it->first;
    for (auto it = m.nodesets.begin(); it != m.nodesets.end(); ++it)        // ascending id: highest written last
        for (int64_t g : it->second) ctx->node_bc_hi[(size_t)g] = (double)it->first;
}

static int finish_matrix(heat_ctx *ctx, heat_matrix *A) {
    if (!A->sell_val.p) HEAT_TRY(sell_from_csr(A, ctx->stream));      // not assembled straight into SELL
    HEAT_TRY(ensure_workspace(A, false, false));
    HaloPlan &h = A->halo;
    if (h.n_neighbors > 0) {
        HEAT_TRY(h.d_send_idx.alloc(h.send_idx.size()));
        HEAT_TRY(h.d_send_buf.alloc(h.send_idx.size()));
        if (!h.send_idx.empty())
            HEAT_CUDA(cudaMemcpyAsync(h.d_send_idx.p, h.send_idx.data(), sizeof(int32_t) * h.send_idx.size(),
                                      cudaMemcpyHostToDevice, ctx->stream));
    }
    if (!A->owned_contiguous) {
        HEAT_TRY(A->d_owned_gids.alloc(A->owned_gids.size()));
        HEAT_CUDA(cudaMemcpyAsync(A->d_owned_gids.p, A->owned_gids.data(), sizeof(int64_t) * A->owned_gids.size(),
                                  cudaMemcpyHostToDevice, ctx->stream));
    }
    HEAT_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

static int new_vector(heat_ctx *ctx, const heat_matrix *A, heat_vector **out) {
    heat_vector *v = new heat_vector();
    v->ctx = ctx; v->n_owned = A->n_owned; v->n_ghost = A->n_ghost;
    int rc = v->d.alloc((size_t)(A->n_owned + A->n_ghost));
    if (rc) { delete v; return rc; }
    cudaError_t e = cudaMemsetAsync(v->d.p, 0, sizeof(double) * (size_t)(A->n_owned + A->n_ghost), ctx->stream);
    if (e != cudaSuccess) { delete v; HEAT_FAIL(100, "cudaMemsetAsync: %s", cudaGetErrorString(e)); }
    *out = v;
    return 0;
}

// ---- analytic cube, any rank count ----------------------------------------------------------------
static int assemble_cube_analytic(heat_ctx *ctx, int mode, heat_matrix *A, heat_vector **B) {
    const HostMesh &m = ctx->mesh;
    if (m.nx < 3 || m.ny < 2 || m.nz < 2) HEAT_FAIL(2, "cube needs nx >= 3, ny >= 2, nz >= 2");
    if (ctx->nranks > m.nz) HEAT_FAIL(2, "more ranks (%d) than k-planes (%d)", ctx->nranks, m.nz);
    CubeGeom c;
    c.nx = m.nx; c.ny = m.ny; c.nz = m.nz;
    slab_range(m.nz, ctx->nranks, ctx->rank, c.k0, c.k1);
    c.plane = (int64_t)(m.nx - 2) * m.ny;
    c.n_owned = c.plane * (c.k1 - c.k0);
    c.ghost_lo = c.k0 > 0 ? c.plane : 0;
    c.ghost_hi = c.k1 < m.nz ? c.plane : 0;
    if (c.n_owned + c.ghost_lo + c.ghost_hi >= (1ll << 31)) HEAT_FAIL(2, "local column ids exceed int32");
    A->n_global = c.plane * m.nz;
    A->n_owned = c.n_owned;
    A->n_ghost = c.ghost_lo + c.ghost_hi;
    A->owned_contiguous = true;
    A->gid0 = c.plane * c.k0;
    HEAT_TRY(new_vector(ctx, A, B));
    {
        const char *how = getenv("HEAT_CUBE_ASSEMBLY");                // "csr": the CSR-first path
        const bool byte_index = sell_cidx_mode() == 1 && spmv_compact_supported();   // else: int32 ids (or the CSR-first path decides)
        bool done = false;
        if (!(how && strcmp(how, "csr") == 0) && sell_cidx_mode() != 2)
            HEAT_TRY(cube_assemble_sell(c, mode, byte_index, A, (*B)->d.p, ctx->stream, &done));
        if (!done) HEAT_TRY(cube_assemble(c, mode, A, (*B)->d.p, ctx->stream));
    }
    // halo plan: rank-1 (lower plane) first, then rank+1 — ghosts grouped by owner ascending
    HaloPlan &h = A->halo;
    h.send_ptr.assign(1, 0); h.recv_ptr.assign(1, 0);
    auto add_nbr = [&](int q, int64_t first_local_row) {
        h.nbr_rank.push_back(q);
        for (int64_t t = 0; t < c.plane; ++t) h.send_idx.push_back((int32_t)(first_local_row + t));
        h.send_ptr.push_back((int64_t)h.send_idx.size());
        h.recv_ptr.push_back(h.recv_ptr.back() + c.plane);
    };
    if (c.k0 > 0) add_nbr(ctx->rank - 1, 0);
    if (c.k1 < m.nz) add_nbr(ctx->rank + 1, c.n_owned - c.plane);
    h.n_neighbors = (int)h.nbr_rank.size();
    // nnz_global by the closed form of SURVEY.md Appendix E
    {
        const int64_t a = m.nx - 2, b = m.ny, cc = m.nz;
        const int64_t edges = (a - 1) * b * cc + a * (b - 1) * cc + a * b * (cc - 1) + (a - 1) * (b - 1) * cc +
                              a * (b - 1) * (cc - 1) + (a - 1) * b * (cc - 1) + (a - 1) * (b - 1) * (cc - 1);
        A->nnz_global = a * b * cc + 2 * edges;
    }
    return 0;
}

// ---- explicit connectivity (Exodus meshes, explicit cubes), any rank count -------------------------
// node_bc: NaN = DOF node, else its prescribed value.  node_owner (may be null): owner rank of every
// ORIGINAL node — overrides `partitioner` (IO::getMatrix distributes rows by element partition).
static int assemble_general(heat_ctx *ctx, int mode, int partitioner, const std::vector<double> &node_bc,
                            const int32_t *node_owner, heat_matrix *A, heat_vector **B) {
    const HostMesh &m = ctx->mesh;
    GeneralAssembler ga;
    cudaStream_t st = ctx->stream;
    if (m.is_cube) {
        HEAT_TRY(ga.make_cube(m.nx, m.ny, m.nz, st));
    } else {
        if (m.num_nodes >= (1ll << 31) || m.num_elem >= (1ll << 31)) HEAT_FAIL(2, "mesh too large for int32 ids");
        HEAT_TRY(ga.upload(m, node_bc, st));
    }
    HEAT_TRY(ga.build_pattern(st));
    A->n_global = ga.n; A->nnz_global = ga.nnz; A->max_row_len = ga.max_row;
    const int P = ctx->nranks;
    if (P == 1) {
        A->n_owned = ga.n; A->n_ghost = 0; A->nnz = ga.nnz;
        A->owned_contiguous = true; A->gid0 = 0;
        HEAT_TRY(new_vector(ctx, A, B));
        HEAT_TRY(A->val.alloc((size_t)ga.nnz));
        HEAT_TRY(ga.fill_values(mode, ga.n, nullptr, nullptr, ga.grow_ptr.p, nullptr, A->val.p, (*B)->d.p, st));
        HEAT_CUDA(cudaStreamSynchronize(st));
        for (int q = 0; q < 4; ++q) A->asm_phase_ms[q] = ga.phase_ms[q];
        A->assemble_fill_ms = ga.phase_ms[3];
        A->row_ptr = std::move(ga.grow_ptr);
        A->col = std::move(ga.gcol);
        A->red2orig_owned.resize((size_t)ga.n);
        if (ga.n > 0)
            HEAT_CUDA(cudaMemcpy(A->red2orig_owned.data(), ga.red2orig.p, sizeof(int64_t) * (size_t)ga.n, cudaMemcpyDeviceToHost));
        return 0;
    }
    // ---- P > 1: every rank holds the global pattern (as every reference rank reads the whole file) ----
    std::vector<int64_t> grow((size_t)ga.n + 1);
    std::vector<int32_t> gcol((size_t)ga.nnz);
    HEAT_CUDA(cudaMemcpy(grow.data(), ga.grow_ptr.p, sizeof(int64_t) * grow.size(), cudaMemcpyDeviceToHost));
    if (ga.nnz > 0) HEAT_CUDA(cudaMemcpy(gcol.data(), ga.gcol.p, sizeof(int32_t) * gcol.size(), cudaMemcpyDeviceToHost));
    std::vector<int64_t> r2o((size_t)ga.n);
    if (ga.n > 0) HEAT_CUDA(cudaMemcpy(r2o.data(), ga.red2orig.p, sizeof(int64_t) * r2o.size(), cudaMemcpyDeviceToHost));
    std::vector<int32_t> part((size_t)ga.n);
    if (node_owner) {
        for (int64_t i = 0; i < ga.n; ++i) part[(size_t)i] = node_owner[(size_t)r2o[(size_t)i]];
    } else if (partitioner == HEAT_PART_METIS_KWAY) {
        HEAT_TRY(metis_kway_partition(ga.n, grow.data(), gcol.data(), P, part.data()));
    } else if (partitioner == HEAT_PART_SLAB && m.is_cube) {
        const int64_t nxy = (int64_t)m.nx * m.ny;
        for (int64_t i = 0; i < ga.n; ++i) {
            const int k = (int)(r2o[(size_t)i] / nxy);
            int owner = 0;
            for (int q = 0; q < P; ++q) { int a, b; slab_range(m.nz, P, q, a, b); if (k >= a && k < b) owner = q; }
            part[(size_t)i] = owner;
        }
    } else {
        contiguous_partition(ga.n, P, part.data());
    }
    heat_plan_sizes sz;
    HEAT_TRY(heat_plan_build(ga.n, grow.data(), gcol.data(), part.data(), P, ctx->rank, &sz, nullptr, nullptr, nullptr,
                             nullptr, nullptr, nullptr, nullptr));
    A->owned_gids.resize((size_t)sz.n_owned); A->ghost_gids.resize((size_t)sz.n_ghost);
    A->ghost_owner.resize((size_t)sz.n_ghost);
    HaloPlan &h = A->halo;
    h.n_neighbors = sz.n_neighbors;
    h.nbr_rank.resize((size_t)sz.n_neighbors);
    h.send_ptr.assign((size_t)sz.n_neighbors + 1, 0); h.recv_ptr.assign((size_t)sz.n_neighbors + 1, 0);
    std::vector<int64_t> send_gids((size_t)sz.n_send);
    HEAT_TRY(heat_plan_build(ga.n, grow.data(), gcol.data(), part.data(), P, ctx->rank, &sz, A->owned_gids.data(),
                             A->ghost_gids.data(), A->ghost_owner.data(), h.nbr_rank.data(), h.send_ptr.data(),
                             send_gids.data(), h.recv_ptr.data()));
    A->n_owned = sz.n_owned; A->n_ghost = sz.n_ghost;
    A->owned_contiguous = false;
    std::vector<int32_t> g2l((size_t)ga.n, -1), owned32((size_t)sz.n_owned);
    std::vector<int64_t> lrow((size_t)sz.n_owned + 1, 0);
    for (int64_t l = 0; l < sz.n_owned; ++l) {
        const int64_t gi = A->owned_gids[(size_t)l];
        g2l[(size_t)gi] = (int32_t)l; owned32[(size_t)l] = (int32_t)gi;
        lrow[(size_t)l + 1] = lrow[(size_t)l] + (grow[(size_t)gi + 1] - grow[(size_t)gi]);
    }
    for (int64_t t = 0; t < sz.n_ghost; ++t) g2l[(size_t)A->ghost_gids[(size_t)t]] = (int32_t)(sz.n_owned + t);
    h.send_idx.resize((size_t)sz.n_send);
    for (int64_t t = 0; t < sz.n_send; ++t) h.send_idx[(size_t)t] = g2l[(size_t)send_gids[(size_t)t]];
    A->nnz = lrow.back();
    A->red2orig_owned.resize((size_t)sz.n_owned);
    for (int64_t l = 0; l < sz.n_owned; ++l) A->red2orig_owned[(size_t)l] = r2o[(size_t)A->owned_gids[(size_t)l]];
    DevBuf<int32_t> d_owned, d_g2l;
    HEAT_TRY(d_owned.alloc(owned32.size())); HEAT_TRY(d_g2l.alloc(g2l.size()));
    HEAT_TRY(A->row_ptr.alloc(lrow.size())); HEAT_TRY(A->col.alloc((size_t)A->nnz)); HEAT_TRY(A->val.alloc((size_t)A->nnz));
    if (!owned32.empty()) HEAT_CUDA(cudaMemcpyAsync(d_owned.p, owned32.data(), sizeof(int32_t) * owned32.size(), cudaMemcpyHostToDevice, st));
    if (!g2l.empty()) HEAT_CUDA(cudaMemcpyAsync(d_g2l.p, g2l.data(), sizeof(int32_t) * g2l.size(), cudaMemcpyHostToDevice, st));
    HEAT_CUDA(cudaMemcpyAsync(A->row_ptr.p, lrow.data(), sizeof(int64_t) * lrow.size(), cudaMemcpyHostToDevice, st));
    HEAT_TRY(new_vector(ctx, A, B));
    HEAT_TRY(ga.fill_values(mode, sz.n_owned, d_owned.p, d_g2l.p, A->row_ptr.p, A->col.p, A->val.p, (*B)->d.p, st));
    HEAT_CUDA(cudaStreamSynchronize(st));
    for (int q = 0; q < 4; ++q) A->asm_phase_ms[q] = ga.phase_ms[q];
    A->assemble_fill_ms = ga.phase_ms[3];
    return 0;
}

}  // namespace heat

using namespace heat;

extern "C" const char *heat_last_error(void) { return heat::g_err; }
extern "C" int heat_version(void) { return HEAT_B200_VERSION; }
extern "C" unsigned long long heat_kernel_launches(void) { return heat::g_launches; }
extern "C" int heat_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int heat_ctx_create(int device, heat_ctx **out) {
    if (!out) HEAT_FAIL(2, "heat_ctx_create: null out");
    if (device == -1) {            // host-only context: file I/O and decompose work, compute calls fail loudly
        heat_ctx *c = new heat_ctx();
        c->device = -1;
        *out = c;
        return 0;
    }
    int ndev = heat_device_count();
    if (ndev <= 0) HEAT_FAIL(3, "no CUDA device: libheat_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) HEAT_FAIL(2, "device %d out of range (%d visible)", device, ndev);
    HEAT_CUDA(cudaSetDevice(device));
    heat_ctx *c = new heat_ctx();
    c->device = device;
    HEAT_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
    HEAT_CUDA(cudaEventCreate(&c->ev_a));
    HEAT_CUDA(cudaEventCreate(&c->ev_b));
    HEAT_CUDA(cudaEventCreateWithFlags(&c->ev_halo, cudaEventDisableTiming));
    HEAT_CUDA(cudaEventCreateWithFlags(&c->ev_pack, cudaEventDisableTiming));
    *out = c;
    return 0;
}

#define HEAT_NEED_GPU(ctx, what)                                                                   \
    if ((ctx)->device < 0) HEAT_FAIL(3, what ": host-only context (no CUDA device) — there is no CPU fallback")

extern "C" int heat_ctx_set_stream(heat_ctx *ctx, void *cuda_stream) {
    if (!ctx) HEAT_FAIL(2, "null ctx");
    HEAT_NEED_GPU(ctx, "heat_ctx_set_stream");
    HEAT_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
    return 0;
}

extern "C" int heat_ctx_set_output(heat_ctx *ctx, int word_size, int largest_nodeset_id) {
    if (!ctx) HEAT_FAIL(2, "null ctx");
    if (word_size != 4 && word_size != 8) HEAT_FAIL(2, "heat_ctx_set_output: word size must be 4 or 8, not %d", word_size);
    if (ctx->write_file && ctx->write_file->nc.dim_id("num_nodes") >= 0 && word_size != ctx->out_word_size)
        HEAT_FAIL(4, "heat_ctx_set_output: the output mesh is already written with word size %d", ctx->out_word_size);
    ctx->out_word_size = word_size;
    ctx->out_largest_id = largest_nodeset_id != 0;
    return 0;
}

extern "C" int heat_close(heat_ctx *ctx) {
    if (!ctx) return 0;
    if (ctx->device >= 0) {
        cudaSetDevice(ctx->device);
        cudaDeviceSynchronize();
    }
    exo_close(ctx->read_file); ctx->read_file = nullptr;
    exo_close(ctx->write_file); ctx->write_file = nullptr;
    comm_destroy(ctx);
    if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->copy_out_stream) cudaStreamDestroy(ctx->copy_out_stream);
    for (int q = 0; q < 2; ++q) {
        if (ctx->ev_in[q]) cudaEventDestroy(ctx->ev_in[q]);
        if (ctx->ev_done[q]) cudaEventDestroy(ctx->ev_done[q]);
        if (ctx->ev_out[q]) cudaEventDestroy(ctx->ev_out[q]);
    }
    if (ctx->ev_copy) cudaEventDestroy(ctx->ev_copy);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->ev_a) cudaEventDestroy(ctx->ev_a);
    if (ctx->ev_b) cudaEventDestroy(ctx->ev_b);
    if (ctx->ev_halo) cudaEventDestroy(ctx->ev_halo);
    if (ctx->ev_pack) cudaEventDestroy(ctx->ev_pack);
    delete ctx;
    return 0;
}

extern "C" int heat_mesh_set(heat_ctx *ctx, int64_t num_nodes, int num_dim, const double *x, const double *y,
                             const double *z, int64_t num_elem, int npe, const int32_t *conn, int num_node_sets,
                             const int64_t *ns_ids, const int64_t *ns_ptr, const int64_t *ns_nodes) {
    if (!ctx || !x || !y || (num_elem > 0 && !conn) || num_nodes < 0 || npe < 2) HEAT_FAIL(2, "heat_mesh_set: bad arguments");
    HostMesh &m = ctx->mesh;
    m = HostMesh();
    m.valid = true; m.num_nodes = num_nodes; m.num_elem = num_elem; m.num_dim = num_dim; m.npe = npe;
    m.x.assign(x, x + num_nodes); m.y.assign(y, y + num_nodes);
    if (z) m.z.assign(z, z + num_nodes);
    m.conn.assign(conn, conn + num_elem * npe);
    for (int64_t q = 0; q < num_elem * npe; ++q)
        if (conn[q] < 0 || conn[q] >= num_nodes) HEAT_FAIL(2, "heat_mesh_set: connectivity entry %lld out of range", (long long)conn[q]);
    for (int s = 0; s < num_node_sets; ++s) {
        auto &v = m.nodesets[ns_ids[s]];
        for (int64_t t = ns_ptr[s]; t < ns_ptr[s + 1]; ++t) {
            if (ns_nodes[t] < 0 || ns_nodes[t] >= num_nodes) HEAT_FAIL(2, "heat_mesh_set: nodeset node out of range");
            v.push_back(ns_nodes[t]);
        }
    }
    m.elem_type = npe == 4 ? "TETRA" : npe == 3 ? "TRI" : npe == 8 ? "HEX" : "UNKNOWN";
    return 0;
}

extern "C" int heat_mesh_cube(heat_ctx *ctx, int nx, int ny, int nz, int explicit_mesh) {
    if (!ctx || nx < 3 || ny < 2 || nz < 2) HEAT_FAIL(2, "heat_mesh_cube: need nx >= 3, ny >= 2, nz >= 2");
    HostMesh &m = ctx->mesh;
    m = HostMesh();
    m.valid = true; m.is_cube = true; m.cube_explicit = explicit_mesh != 0;
    m.nx = nx; m.ny = ny; m.nz = nz;
    m.num_nodes = (int64_t)nx * ny * nz;
    m.num_elem = 6ll * (nx - 1) * (ny - 1) * (nz - 1);
    m.num_dim = 3; m.npe = 4; m.elem_type = "TETRA";
    return 0;
}

extern "C" int heat_assemble(heat_ctx *ctx, int op_mode, int partitioner, heat_matrix **A_out, heat_vector **X_out,
                             heat_vector **B_out) {
    if (!ctx || !A_out || !X_out || !B_out) HEAT_FAIL(2, "heat_assemble: null argument");
    if (!ctx->mesh.valid) HEAT_FAIL(4, "heat_assemble: no mesh (call heat_open / heat_mesh_set / heat_mesh_cube first)");
    if (op_mode != HEAT_OP_GRAPH_LAPLACIAN && op_mode != HEAT_OP_P1_FEM) HEAT_FAIL(2, "heat_assemble: unknown operator %d", op_mode);
    HEAT_NEED_GPU(ctx, "heat_assemble");
    HEAT_CUDA(cudaSetDevice(ctx->device));
    const HostMesh &m = ctx->mesh;
    if (!m.is_cube) build_node_bc(ctx);
    heat_matrix *A = new heat_matrix();
    A->ctx = ctx; A->op_mode = op_mode;
    heat_vector *B = nullptr, *X = nullptr;
    HEAT_CUDA(cudaEventRecord(ctx->ev_a, ctx->stream));
    int rc = (m.is_cube && !m.cube_explicit) ? assemble_cube_analytic(ctx, op_mode, A, &B)
                                             : assemble_general(ctx, op_mode, partitioner, ctx->node_bc, nullptr, A, &B);
    if (!rc) rc = finish_matrix(ctx, A);
    if (!rc) rc = new_vector(ctx, A, &X);
    if (rc) { delete A; delete B; delete X; return rc; }
    HEAT_CUDA(cudaEventRecord(ctx->ev_b, ctx->stream));
    HEAT_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    HEAT_CUDA(cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b));
    A->assemble_ms = ms;
    ctx->n_global = A->n_global;
    ctx->owned_gids.resize((size_t)A->n_owned);
    if (A->owned_contiguous) for (int64_t l = 0; l < A->n_owned; ++l) ctx->owned_gids[(size_t)l] = A->gid0 + l;
    else ctx->owned_gids = A->owned_gids;
    *A_out = A; *X_out = X; *B_out = B;
    return 0;
}

// IO::getMatrix (ExodusIO.hpp:733-1489): the Laplacian of the WHOLE mesh (nodesets are not applied, so
// the matrix is singular, :729-731), one row per node, rows distributed by element partition + the
// node-ownership rule of :1191-1295.  Element partition: METIS_PartMeshDual with the reference's
// ncommon (the serial counterpart of ParMETIS_V3_PartMeshKway at :919, which does not exist here).
extern "C" int heat_get_matrix(heat_ctx *ctx, int op_mode, heat_matrix **A_out) {
    if (!ctx || !A_out) HEAT_FAIL(2, "heat_get_matrix: null argument");
    if (!ctx->mesh.valid) HEAT_FAIL(4, "heat_get_matrix: no mesh (readFID == -1)");
    if (ctx->mesh.is_cube) HEAT_FAIL(4, "heat_get_matrix: needs a mesh from heat_open / heat_mesh_set");
    if (op_mode != HEAT_OP_GRAPH_LAPLACIAN && op_mode != HEAT_OP_P1_FEM) HEAT_FAIL(2, "heat_get_matrix: unknown operator %d", op_mode);
    HEAT_NEED_GPU(ctx, "heat_get_matrix");
    HEAT_CUDA(cudaSetDevice(ctx->device));
    const HostMesh &m = ctx->mesh;
    std::vector<int32_t> owner;
    if (ctx->nranks > 1) {
        std::vector<int64_t> epart((size_t)m.num_elem), npart((size_t)m.num_nodes);
        int64_t obj = 0;
        HEAT_TRY(heat_decompose_partition(ctx, ctx->nranks, &obj, epart.data(), npart.data()));
        owner.resize((size_t)m.num_nodes);
        HEAT_TRY(heat_node_owners(m.num_nodes, m.num_elem, m.npe, m.conn.data(), epart.data(), ctx->nranks, owner.data()));
    }
    const std::vector<double> no_bc((size_t)m.num_nodes, std::numeric_limits<double>::quiet_NaN());
    heat_matrix *A = new heat_matrix();
    A->ctx = ctx; A->op_mode = op_mode;
    heat_vector *B = nullptr;
    HEAT_CUDA(cudaEventRecord(ctx->ev_a, ctx->stream));
    int rc = assemble_general(ctx, op_mode, HEAT_PART_CONTIGUOUS, no_bc, owner.empty() ? nullptr : owner.data(), A, &B);
    if (!rc) rc = finish_matrix(ctx, A);
    delete B;                                            // no Dirichlet nodes => the load vector is zero
    if (rc) { delete A; return rc; }
    HEAT_CUDA(cudaEventRecord(ctx->ev_b, ctx->stream));
    HEAT_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    HEAT_CUDA(cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b));
    A->assemble_ms = ms;
    *A_out = A;
    return 0;
}

// nodeSetMap of getMatrix (ExodusIO.hpp:1447-1466): the nodes of nodeset `set_id` whose rows this rank owns
extern "C" int heat_matrix_owned_nodeset(const heat_matrix *A, int64_t set_id, int64_t *count, int64_t *nodes_out) {
    if (!A || !count) HEAT_FAIL(2, "heat_matrix_owned_nodeset: null argument");
    const HostMesh &m = A->ctx->mesh;
    auto it = m.nodesets.find(set_id);
    if (it == m.nodesets.end()) HEAT_FAIL(5, "heat_matrix_owned_nodeset: no nodeset with id %lld", (long long)set_id);
    std::vector<int64_t> r2o((size_t)A->n_owned);
    HEAT_TRY(heat_matrix_export_red2orig(A, r2o.data()));
    std::vector<char> mine((size_t)m.num_nodes, 0);
    for (int64_t g : r2o) mine[(size_t)g] = 1;
    std::vector<int64_t> out;
    for (int64_t g : it->second)
        if (mine[(size_t)g]) out.push_back(g);
    std::sort(out.begin(), out.end());                               // std::set<idx_t> order
    out.erase(std::unique(out.begin(), out.end()), out.end());
    *count = (int64_t)out.size();
    if (nodes_out) std::copy(out.begin(), out.end(), nodes_out);
    return 0;
}

// ex_get_ids(readFID, EX_NODE_SET, ids) (ExodusIO.hpp:172, :1443): ascending ids of the mesh's nodesets
extern "C" int heat_mesh_nodeset_ids(const heat_ctx *ctx, int *count, int64_t *ids_out) {
    if (!ctx || !count) HEAT_FAIL(2, "heat_mesh_nodeset_ids: null argument");
    if (!ctx->mesh.valid) HEAT_FAIL(4, "heat_mesh_nodeset_ids: no mesh");
    *count = (int)ctx->mesh.nodesets.size();
    if (ids_out) {
        int w = 0;
        for (const auto &kv : ctx->mesh.nodesets) ids_out[w++] = kv.first;
    }
    return 0;
}

extern "C" int heat_power_method(heat_ctx *ctx, heat_matrix *A, int niters, double tolerance, uint64_t seed,
                                 heat_power_info *info) {
    if (!ctx || !A) HEAT_FAIL(2, "heat_power_method: null argument");
    HEAT_NEED_GPU(ctx, "heat_power_method");
    return power_method_device(ctx, A, niters, tolerance, seed, info);
}

extern "C" void heat_solve_opts_default(heat_solve_opts *o) {
    if (!o) return;
    o->solver = HEAT_SOLVER_CG; o->prec = HEAT_PREC_JACOBI;
    o->max_iters = 300; o->tol = 1e-14;                  // BelosMueLuSolver.cpp:149-151
    o->cheb_degree = 1; o->cheb_lambda_max = 0.0; o->cheb_ratio = 30.0; o->check_every = 32;
    o->gmres_restart = 300;                               // Belos "Num Blocks" default
}

extern "C" int heat_solve(heat_ctx *ctx, heat_matrix *A, heat_vector *X, const heat_vector *B,
                          const heat_solve_opts *opts, heat_solve_info *info) {
    if (!ctx || !A || !X || !B || !opts) HEAT_FAIL(2, "heat_solve: null argument");
    if (X->n_owned != A->n_owned || B->n_owned != A->n_owned) HEAT_FAIL(2, "heat_solve: vector/matrix size mismatch");
    if (X->n_ghost != A->n_ghost) HEAT_FAIL(2, "heat_solve: X holds %lld ghost entries, the matrix needs %lld (create it with heat_vector_create on this matrix)",
                                            (long long)X->n_ghost, (long long)A->n_ghost);
    return solve_device(ctx, A, X->d.p, B->d.p, *opts, info);
}

// belosSolver's per-iteration output (BelosMueLuSolver.cpp:113-133: io.writeSolution(X, i) after every
// pass of the loop) inside ONE Krylov run: the device loop is polled every `write_every` iterations and
// the iterate it holds at that moment is written as the next time step.
extern "C" int heat_solve_trajectory(heat_ctx *ctx, heat_matrix *A, heat_vector *X, const heat_vector *B,
                                     const heat_solve_opts *opts, int write_every, int first_timestep,
                                     heat_solve_info *info, int *frames_written) {
    if (!ctx || !A || !X || !B || !opts) HEAT_FAIL(2, "heat_solve_trajectory: null argument");
    if (X->n_owned != A->n_owned || B->n_owned != A->n_owned || X->n_ghost != A->n_ghost) HEAT_FAIL(2, "heat_solve_trajectory: vector/matrix size mismatch");
    if (write_every < 1 || first_timestep < 0) HEAT_FAIL(2, "heat_solve_trajectory: write_every >= 1 and first_timestep >= 0 needed");
    heat_solve_opts o = *opts;
    o.check_every = write_every;
    int frames = 0, last_iters = -1;
    const std::function<int(int)> hook = [&](int iters_done) -> int {
        if (iters_done == last_iters) return 0;            // the loop stopped exactly on a poll boundary
        last_iters = iters_done;
        return heat_write_solution(ctx, X, first_timestep + frames++);
    };
    HEAT_TRY(solve_device(ctx, A, X->d.p, B->d.p, o, info, &hook));
    if (frames == 0) HEAT_TRY(heat_write_solution(ctx, X, first_timestep + frames++));     // max_iters == 0
    if (frames_written) *frames_written = frames;
    return 0;
}

extern "C" int heat_cg_iterations(heat_ctx *ctx, heat_matrix *A, heat_vector *X, const heat_vector *B,
                                  const heat_solve_opts *opts, int iters, heat_solve_info *info) {
    if (!opts) HEAT_FAIL(2, "heat_cg_iterations: null opts");
    heat_solve_opts o = *opts;
    o.tol = 0.0; o.max_iters = iters;
    if (o.check_every <= 0 || o.check_every > iters) o.check_every = iters > 0 ? iters : 1;
    return heat_solve(ctx, A, X, B, &o, info);
}

// End-to-end entry point: host buffers in, host buffer out.  The device staging buffers live with the
// matrix (no cudaMalloc per call); x0 goes up on the compute stream, b on a second stream so that its
// transfer overlaps the r0 = b - A x0 set-up SpMV (solve_device waits for it right after that SpMV).
extern "C" int heat_solve_host(heat_ctx *ctx, heat_matrix *A, const double *b_host, double *x_host,
                               const heat_solve_opts *opts, heat_solve_info *info) {
    if (!ctx || !A || !b_host || !x_host || !opts) HEAT_FAIL(2, "heat_solve_host: null argument");
    HEAT_NEED_GPU(ctx, "heat_solve_host");
    HEAT_CUDA(cudaSetDevice(ctx->device));
    const size_t nv = (size_t)(A->n_owned + A->n_ghost), nb = sizeof(double) * (size_t)A->n_owned;
    if (!A->h_x.p) HEAT_TRY(A->h_x.alloc(nv));
    if (!A->h_b.p) HEAT_TRY(A->h_b.alloc((size_t)A->n_owned));
    if (!ctx->copy_stream) {
        HEAT_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        HEAT_CUDA(cudaEventCreateWithFlags(&ctx->ev_copy, cudaEventDisableTiming));
    }
    // the previous user of h_b (an earlier solve on ctx->stream) must be done before it is overwritten
    HEAT_CUDA(cudaEventRecord(ctx->ev_copy, ctx->stream));
    HEAT_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_copy, 0));
    HEAT_CUDA(cudaMemcpyAsync(A->h_b.p, b_host, nb, cudaMemcpyHostToDevice, ctx->copy_stream));
    HEAT_CUDA(cudaEventRecord(ctx->ev_copy, ctx->copy_stream));
    HEAT_CUDA(cudaMemcpyAsync(A->h_x.p, x_host, nb, cudaMemcpyHostToDevice, ctx->stream));
    ctx->wait_before_rhs = ctx->ev_copy;
    int rc = solve_device(ctx, A, A->h_x.p, A->h_b.p, *opts, info);
    ctx->wait_before_rhs = nullptr;
    if (rc) return rc;
    HEAT_CUDA(cudaMemcpyAsync(x_host, A->h_x.p, nb, cudaMemcpyDeviceToHost, ctx->stream));
    HEAT_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// A sequence of solves with ONE matrix through host buffers (time stepping, many right-hand sides): the role of a
// loop around belosSolver with a new B per pass.  Two staging sets and two copy streams: while system k is solved the
// inputs of system k+1 go up and the solution of system k-1 goes down (PCIe is full duplex), so a stream of solves
// runs at the device-resident rate instead of paying 24 n bytes of serial copies per system.
extern "C" int heat_solve_host_batch(heat_ctx *ctx, heat_matrix *A, int count, const double *const *b_hosts,
                                     const double *const *x0_hosts, double *const *x_hosts, const heat_solve_opts *opts,
                                     heat_solve_info *infos) {
    if (!ctx || !A || count < 0 || (count > 0 && (!b_hosts || !x_hosts)) || !opts) HEAT_FAIL(2, "heat_solve_host_batch: bad arguments");
    HEAT_NEED_GPU(ctx, "heat_solve_host_batch");
    HEAT_CUDA(cudaSetDevice(ctx->device));
    for (int k = 0; k < count; ++k)
        if (!b_hosts[k] || !x_hosts[k]) HEAT_FAIL(2, "heat_solve_host_batch: null buffer for system %d", k);
    const size_t nv = (size_t)(A->n_owned + A->n_ghost), nb = sizeof(double) * (size_t)A->n_owned;
    DevBuf<double> *hx[2] = {&A->h_x, &A->h_x2}, *hb[2] = {&A->h_b, &A->h_b2};
    for (int q = 0; q < 2; ++q) {
        if (!hx[q]->p) HEAT_TRY(hx[q]->alloc(nv));
        if (!hb[q]->p) HEAT_TRY(hb[q]->alloc((size_t)A->n_owned));
    }
    if (!ctx->copy_stream) {
        HEAT_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        HEAT_CUDA(cudaEventCreateWithFlags(&ctx->ev_copy, cudaEventDisableTiming));
    }
    if (!ctx->copy_out_stream) {
        HEAT_CUDA(cudaStreamCreateWithFlags(&ctx->copy_out_stream, cudaStreamNonBlocking));
        for (int q = 0; q < 2; ++q) {
            HEAT_CUDA(cudaEventCreateWithFlags(&ctx->ev_in[q], cudaEventDisableTiming));
            HEAT_CUDA(cudaEventCreateWithFlags(&ctx->ev_done[q], cudaEventDisableTiming));
            HEAT_CUDA(cudaEventCreateWithFlags(&ctx->ev_out[q], cudaEventDisableTiming));
        }
    }
    // whatever used the staging buffers before (an earlier heat_solve_host on ctx->stream) must be done
    HEAT_CUDA(cudaEventRecord(ctx->ev_copy, ctx->stream));
    HEAT_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_copy, 0));
    auto upload = [&](int k) -> int {                   // inputs of system k into staging set k & 1 (copy-in stream)
        const int q = k & 1;
        if (k >= 2) HEAT_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_out[q], 0));    // system k-2 has left the set
        if (x0_hosts && x0_hosts[k]) HEAT_CUDA(cudaMemcpyAsync(hx[q]->p, x0_hosts[k], nb, cudaMemcpyHostToDevice, ctx->copy_stream));
        else HEAT_CUDA(cudaMemsetAsync(hx[q]->p, 0, nb, ctx->copy_stream));
        HEAT_CUDA(cudaMemcpyAsync(hb[q]->p, b_hosts[k], nb, cudaMemcpyHostToDevice, ctx->copy_stream));
        HEAT_CUDA(cudaEventRecord(ctx->ev_in[q], ctx->copy_stream));
        return 0;
    };
    if (count > 0) HEAT_TRY(upload(0));
    for (int k = 0; k < count; ++k) {
        const int q = k & 1;
        if (k + 1 < count) HEAT_TRY(upload(k + 1));                        // goes up while system k is solved
        HEAT_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_in[q], 0));
        HEAT_TRY(solve_device(ctx, A, hx[q]->p, hb[q]->p, *opts, infos ? infos + k : nullptr));
        HEAT_CUDA(cudaEventRecord(ctx->ev_done[q], ctx->stream));
        HEAT_CUDA(cudaStreamWaitEvent(ctx->copy_out_stream, ctx->ev_done[q], 0));
        HEAT_CUDA(cudaMemcpyAsync(x_hosts[k], hx[q]->p, nb, cudaMemcpyDeviceToHost, ctx->copy_out_stream));   // goes down while k+1 is solved
        HEAT_CUDA(cudaEventRecord(ctx->ev_out[q], ctx->copy_out_stream));
    }
    HEAT_CUDA(cudaStreamSynchronize(ctx->copy_out_stream));
    HEAT_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    return 0;
}

extern "C" int heat_spmv(heat_ctx *ctx, heat_matrix *A, heat_vector *x, heat_vector *y) {
    if (!ctx || !A || !x || !y) HEAT_FAIL(2, "heat_spmv: null argument");
    if (x->n_owned != A->n_owned || y->n_owned != A->n_owned || x == y || x->n_ghost != A->n_ghost) HEAT_FAIL(2, "heat_spmv: bad vectors");
    HEAT_CUDA(cudaSetDevice(ctx->device));
    CgGate nogate{nullptr, nullptr, nullptr, 0};
    return spmv_halo(ctx, A, x->d.p, y->d.p, nogate, nullptr);
}

extern "C" int heat_spmv_peer(heat_ctx *ctx, heat_matrix *A, heat_vector *x, heat_vector *y, int repeat,
                              double *xy_global, double *kernel_ms) {
    if (!ctx || !A || !x || !y) HEAT_FAIL(2, "heat_spmv_peer: null argument");
    if (x->n_owned != A->n_owned || y->n_owned != A->n_owned || x == y) HEAT_FAIL(2, "heat_spmv_peer: bad vectors");
    return spmv_peer_once(ctx, A, x->d.p, y->d.p, repeat, xy_global, kernel_ms);
}

// ---- inspection ------------------------------------------------------------------------------------
extern "C" int heat_matrix_get_info(const heat_matrix *A, heat_matrix_info *info) {
    if (!A || !info) HEAT_FAIL(2, "heat_matrix_get_info: null argument");
    const heat_ctx *c = A->ctx;
    memset(info, 0, sizeof(*info));
    info->num_nodes = c->mesh.num_nodes; info->num_elem = c->mesh.num_elem;
    info->n_global = A->n_global; info->nnz_global = A->nnz_global;
    info->n_owned = A->n_owned; info->n_ghost = A->n_ghost; info->nnz_local = A->nnz;
    info->max_row_len = A->max_row_len; info->npe = c->mesh.npe; info->num_dim = c->mesh.num_dim;
    info->num_node_sets = (int32_t)c->mesh.nodesets.size();
    info->rank = c->rank; info->nranks = c->nranks; info->n_neighbors = A->halo.n_neighbors;
    info->sell_chunk = kSellChunk; info->sell_padded_nnz = A->sell_padded;
    info->n_boundary_slices = A->n_bnd_slices; info->n_slices = A->n_slices;
    info->assemble_ms = A->assemble_ms;
    info->peer_path = A->peer ? 1 : 0;
    info->col_index_bytes = A->sell_cmode;
    info->assemble_fill_ms = A->assemble_fill_ms;
    for (int q = 0; q < 4; ++q) info->asm_phase_ms[q] = A->asm_phase_ms[q];
    info->csr_resident = A->row_ptr.p ? 1 : 0;
    info->matrix_bytes = (int64_t)(A->row_ptr.n * 8 + A->col.n * 4 + A->val.n * 8 + A->sell_val.n * 8 + A->sell_col.n * 4 +
                                   A->sell_idx8.n + A->sell_tab.n * 4 + A->slice_ptr.n * 8 + A->slice_meta.n * 16 +
                                   A->sell_rowlen.n + A->diag.n * 8 + A->dinv.n * 8 + A->slice_cptr.n * 8 + A->slice_mode.n);
    return 0;
}

extern "C" int heat_matrix_export_csr(const heat_matrix *A, int64_t *row_ptr, int32_t *col, double *val) {
    if (!A) HEAT_FAIL(2, "null matrix");
    HEAT_CUDA(cudaSetDevice(A->ctx->device));
    HEAT_TRY(sell_to_csr(const_cast<heat_matrix *>(A), A->ctx->stream));     // assembled straight into SELL: build it now
    HEAT_CUDA(cudaStreamSynchronize(A->ctx->stream));
    if (row_ptr) HEAT_CUDA(cudaMemcpy(row_ptr, A->row_ptr.p, sizeof(int64_t) * (size_t)(A->n_owned + 1), cudaMemcpyDeviceToHost));
    if (col && A->nnz) HEAT_CUDA(cudaMemcpy(col, A->col.p, sizeof(int32_t) * (size_t)A->nnz, cudaMemcpyDeviceToHost));
    if (val && A->nnz) HEAT_CUDA(cudaMemcpy(val, A->val.p, sizeof(double) * (size_t)A->nnz, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int heat_matrix_export_maps(const heat_matrix *A, int64_t *owned, int64_t *ghost, int32_t *ghost_owner) {
    if (!A) HEAT_FAIL(2, "null matrix");
    const heat_ctx *c = A->ctx;
    if (owned) {
        if (A->owned_contiguous) for (int64_t l = 0; l < A->n_owned; ++l) owned[l] = A->gid0 + l;
        else std::copy(A->owned_gids.begin(), A->owned_gids.end(), owned);
    }
    if (A->owned_contiguous && A->n_ghost > 0 && c->mesh.is_cube && !c->mesh.cube_explicit) {
        // analytic slab: lower plane (rank-1) then upper plane (rank+1)
        const int64_t plane = (int64_t)(c->mesh.nx - 2) * c->mesh.ny;
        int64_t w = 0;
        const bool has_lo = A->gid0 > 0;
        if (has_lo) for (int64_t t = 0; t < plane; ++t, ++w) { if (ghost) ghost[w] = A->gid0 - plane + t; if (ghost_owner) ghost_owner[w] = c->rank - 1; }
        if (w < A->n_ghost) for (int64_t t = 0; t < plane; ++t, ++w) { if (ghost) ghost[w] = A->gid0 + A->n_owned + t; if (ghost_owner) ghost_owner[w] = c->rank + 1; }
    } else {
        if (ghost) std::copy(A->ghost_gids.begin(), A->ghost_gids.end(), ghost);
        if (ghost_owner) std::copy(A->ghost_owner.begin(), A->ghost_owner.end(), ghost_owner);
    }
    return 0;
}

extern "C" int heat_matrix_export_plan(const heat_matrix *A, int32_t *nbr_rank, int64_t *send_ptr, int32_t *send_idx,
                                       int64_t *recv_ptr) {
    if (!A) HEAT_FAIL(2, "null matrix");
    const HaloPlan &h = A->halo;
    if (nbr_rank) std::copy(h.nbr_rank.begin(), h.nbr_rank.end(), nbr_rank);
    if (send_ptr) { if (h.send_ptr.empty()) send_ptr[0] = 0; else std::copy(h.send_ptr.begin(), h.send_ptr.end(), send_ptr); }
    if (send_idx) std::copy(h.send_idx.begin(), h.send_idx.end(), send_idx);
    if (recv_ptr) { if (h.recv_ptr.empty()) recv_ptr[0] = 0; else std::copy(h.recv_ptr.begin(), h.recv_ptr.end(), recv_ptr); }
    return 0;
}

extern "C" int heat_matrix_export_red2orig(const heat_matrix *A, int64_t *out) {
    if (!A || !out) HEAT_FAIL(2, "null argument");
    const heat_ctx *c = A->ctx;
    if (c->mesh.is_cube && !c->mesh.cube_explicit) {
        const int64_t w = c->mesh.nx - 2;
        for (int64_t l = 0; l < A->n_owned; ++l) {
            const int64_t gi = A->gid0 + l;
            out[l] = (gi % w) + 1 + (int64_t)c->mesh.nx * (gi / w);     // node = i + nx*(j + ny*k)
        }
        return 0;
    }
    std::copy(A->red2orig_owned.begin(), A->red2orig_owned.end(), out);
    return 0;
}

extern "C" int heat_matrix_free(heat_matrix *A) {
    if (A) { cudaSetDevice(A->ctx->device); cudaStreamSynchronize(A->ctx->stream); peer_matrix_teardown(A); ilu_free(A); delete A; }
    return 0;
}

// L and U of HEAT_PREC_ILU0 on the pattern of the local CSR (parity hook)
extern "C" int heat_matrix_export_ilu0(const heat_matrix *A, double *lu_host, int *n_levels_lower, int *n_levels_upper) {
    if (!A) HEAT_FAIL(2, "null matrix");
    HEAT_CUDA(cudaSetDevice(A->ctx->device));
    return ilu_export(A, lu_host, n_levels_lower, n_levels_upper);
}

// ---- vectors -----------------------------------------------------------------------------------------
extern "C" int heat_vector_create(heat_ctx *ctx, const heat_matrix *A, heat_vector **out) {
    if (!ctx || !A || !out) HEAT_FAIL(2, "heat_vector_create: null argument");
    HEAT_CUDA(cudaSetDevice(ctx->device));
    return new_vector(ctx, A, out);
}
extern "C" int64_t heat_vector_size(const heat_vector *v) { return v ? v->n_owned : -1; }
extern "C" void *heat_vector_device_ptr(heat_vector *v) { return v ? (void *)v->d.p : nullptr; }

extern "C" int heat_vector_set(heat_ctx *ctx, heat_vector *v, const void *src, int64_t count) {
    if (!ctx || !v || !src || count < 0 || count > v->n_owned) HEAT_FAIL(2, "heat_vector_set: bad arguments");
    HEAT_CUDA(cudaSetDevice(ctx->device));
    HEAT_CUDA(cudaMemcpyAsync(v->d.p, src, sizeof(double) * (size_t)count,
                              is_device_ptr(src) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
    HEAT_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
extern "C" int heat_vector_get(heat_ctx *ctx, const heat_vector *v, void *dst, int64_t count) {
    if (!ctx || !v || !dst || count < 0 || count > v->n_owned) HEAT_FAIL(2, "heat_vector_get: bad arguments");
    HEAT_CUDA(cudaSetDevice(ctx->device));
    HEAT_CUDA(cudaMemcpyAsync(dst, v->d.p, sizeof(double) * (size_t)count,
                              is_device_ptr(dst) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream));
    HEAT_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
extern "C" int heat_vector_fill(heat_ctx *ctx, heat_vector *v, double value) {
    if (!ctx || !v) HEAT_FAIL(2, "heat_vector_fill: null argument");
    HEAT_CUDA(cudaSetDevice(ctx->device));
    return launch_fill(v->n_owned, v->d.p, value, ctx->stream);
}
extern "C" int heat_vector_fill_hash(heat_ctx *ctx, const heat_matrix *A, heat_vector *v, uint64_t seed) {
    if (!ctx || !A || !v || v->n_owned != A->n_owned) HEAT_FAIL(2, "heat_vector_fill_hash: bad arguments");
    HEAT_CUDA(cudaSetDevice(ctx->device));
    return launch_fill_hash(v->n_owned, v->d.p, A->owned_contiguous ? nullptr : A->d_owned_gids.p, A->gid0, seed, ctx->stream);
}
extern "C" int heat_vector_free(heat_vector *v) {
    if (v) { cudaSetDevice(v->ctx->device); cudaStreamSynchronize(v->ctx->stream); delete v; }
    return 0;
}
