// comm.cu — multi-GPU plumbing: one process per GPU, NCCL over NVLink 5 / NVSwitch.
//
// Replaces the reference's MPI usage on the hot path: the per-SpMV Tpetra Import
// (MPI_Isend/Irecv of ghost values) becomes a pack kernel + grouped ncclSend/ncclRecv on a
// high-priority side stream overlapped with the interior SpMV; every Teuchos::reduceAll becomes
// (part of) one ncclAllReduce of <= 3 doubles.
// NCCL is bound with dlopen at heat_comm_init time so that (a) single-GPU use needs no NCCL and
// (b) inside a torch process the already-loaded libnccl.so.2 (same soname) is the one used.
#include <dlfcn.h>
#include <nccl.h>

#include "comm.cuh"
#include "kernels.cuh"
#include "peer.cuh"
#include "solve.cuh"

namespace heat {

static_assert(NCCL_UNIQUE_ID_BYTES == HEAT_COMM_ID_BYTES, "ncclUniqueId size");

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load() {
    if (g_nccl.handle) return 0;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) HEAT_FAIL(40, "cannot load libnccl.so.2: %s", dlerror());
#define LOAD(field, name)                                                          \
    *(void **)(&g_nccl.field) = dlsym(h, name);                                    \
    if (!g_nccl.field) HEAT_FAIL(40, "libnccl: missing symbol %s", name)
    LOAD(GetUniqueId, "ncclGetUniqueId");
    LOAD(CommInitRank, "ncclCommInitRank");
    LOAD(CommDestroy, "ncclCommDestroy");
    LOAD(AllReduce, "ncclAllReduce");
    LOAD(Send, "ncclSend");
    LOAD(Recv, "ncclRecv");
    LOAD(AllGather, "ncclAllGather");
    LOAD(GroupStart, "ncclGroupStart");
    LOAD(GroupEnd, "ncclGroupEnd");
    LOAD(GetErrorString, "ncclGetErrorString");
#undef LOAD
    g_nccl.handle = h;
    return 0;
}

#define HEAT_NCCL(call)                                                                           \
    do {                                                                                          \
        ncclResult_t r__ = (call);                                                                \
        if (r__ != ncclSuccess) HEAT_FAIL(41, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, heat::g_nccl.GetErrorString(r__)); \
    } while (0)

void peer_arena_teardown(heat_ctx *ctx);
int comm_destroy(heat_ctx *ctx) {
    peer_arena_teardown(ctx);
    if (ctx->nccl_comm && g_nccl.handle) g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
    return 0;
}

int comm_allreduce_sum(heat_ctx *ctx, double *buf, int count) {
    if (ctx->nranks <= 1) return 0;
    HEAT_NCCL(g_nccl.AllReduce(buf, buf, (size_t)count, ncclDouble, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    return 0;
}

// pack boundary values of x (main stream) then exchange on the side stream; x's ghost segment
// [n_owned, n_owned+n_ghost) receives directly (ghosts are grouped by owner => contiguous recv).
int halo_begin(heat_ctx *ctx, heat_matrix *A, double *x) {
    HaloPlan &h = A->halo;
    if (ctx->nranks <= 1 || h.n_neighbors == 0) return 0;
    const int64_t n_send = h.send_ptr.back();
    HEAT_TRY(launch_gather(n_send, x, h.d_send_idx.p, h.d_send_buf.p, ctx->stream));
    HEAT_CUDA(cudaEventRecord(ctx->ev_pack, ctx->stream));
    HEAT_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_pack, 0));
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    HEAT_NCCL(g_nccl.GroupStart());
    for (int s = 0; s < h.n_neighbors; ++s) {
        const int64_t ns = h.send_ptr[s + 1] - h.send_ptr[s], nr = h.recv_ptr[s + 1] - h.recv_ptr[s];
        if (ns > 0)
            HEAT_NCCL(g_nccl.Send(h.d_send_buf.p + h.send_ptr[s], (size_t)ns, ncclDouble, h.nbr_rank[s], comm, ctx->comm_stream));
        if (nr > 0)
            HEAT_NCCL(g_nccl.Recv(x + A->n_owned + h.recv_ptr[s], (size_t)nr, ncclDouble, h.nbr_rank[s], comm, ctx->comm_stream));
    }
    HEAT_NCCL(g_nccl.GroupEnd());
    HEAT_CUDA(cudaEventRecord(ctx->ev_halo, ctx->comm_stream));
    return 0;
}

int halo_end(heat_ctx *ctx, heat_matrix *A) {
    if (ctx->nranks <= 1 || A->halo.n_neighbors == 0) return 0;
    HEAT_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_halo, 0));
    return 0;
}

// -------------------------------------------------------------------------------------------------
// peer-memory path set-up (peer.cuh): CUDA-IPC exchange of the arenas and of the SpMV-input buffers
// -------------------------------------------------------------------------------------------------
static int allgather_bytes(heat_ctx *ctx, const void *mine, size_t bytes, std::vector<unsigned char> &all) {
    const int P = ctx->nranks;
    DevBuf<unsigned char> d;
    HEAT_TRY(d.alloc(bytes * (size_t)P));
    HEAT_CUDA(cudaMemcpyAsync(d.p + bytes * (size_t)ctx->rank, mine, bytes, cudaMemcpyHostToDevice, ctx->stream));
    HEAT_NCCL(g_nccl.AllGather(d.p + bytes * (size_t)ctx->rank, d.p, bytes, ncclChar, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    all.resize(bytes * (size_t)P);
    HEAT_CUDA(cudaMemcpyAsync(all.data(), d.p, all.size(), cudaMemcpyDeviceToHost, ctx->stream));
    HEAT_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// every rank must reach the same verdict: all-reduce (min) of a local ok flag through the gather
static int all_agree(heat_ctx *ctx, int ok_local, int *ok_all) {
    std::vector<unsigned char> all;
    unsigned char mine = ok_local ? 1 : 0;
    HEAT_TRY(allgather_bytes(ctx, &mine, 1, all));
    int ok = 1;
    for (unsigned char c : all) ok &= (c != 0);
    *ok_all = ok;
    return 0;
}

int peer_arena_setup(heat_ctx *ctx) {
    ctx->peer_enabled = false;
    const char *env = getenv("HEAT_COMM");
    if (env && strcmp(env, "nccl") == 0) return 0;
    if (ctx->nranks < 2 || ctx->nranks > kPeerMaxRanks) return 0;
    int ok = 1;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (cudaMalloc((void **)&ctx->peer_arena, sizeof(unsigned long long) * kPeerArenaWords) != cudaSuccess) ok = 0;
    if (ok && cudaMemset(ctx->peer_arena, 0, sizeof(unsigned long long) * kPeerArenaWords) != cudaSuccess) ok = 0;
    if (ok && cudaIpcGetMemHandle(&mine, ctx->peer_arena) != cudaSuccess) ok = 0;
    cudaGetLastError();
    std::vector<unsigned char> all;
    HEAT_TRY(allgather_bytes(ctx, &mine, sizeof(mine), all));
    if (ok) {
        for (int q = 0; q < ctx->nranks && ok; ++q) {
            if (q == ctx->rank) { ctx->peer_arena_of[q] = ctx->peer_arena; continue; }
            cudaIpcMemHandle_t h;
            memcpy(&h, all.data() + sizeof(h) * (size_t)q, sizeof(h));
            void *mapped = nullptr;
            if (cudaIpcOpenMemHandle(&mapped, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); }
            ctx->peer_arena_of[q] = (unsigned long long *)mapped;
        }
    }
    int ok_all = 0;
    HEAT_TRY(all_agree(ctx, ok, &ok_all));
    ctx->peer_enabled = ok_all != 0;
    return 0;
}

void peer_arena_teardown(heat_ctx *ctx) {
    for (int q = 0; q < kPeerMaxRanks; ++q) {
        if (ctx->peer_arena_of[q] && ctx->peer_arena_of[q] != ctx->peer_arena) cudaIpcCloseMemHandle(ctx->peer_arena_of[q]);
        ctx->peer_arena_of[q] = nullptr;
    }
    if (ctx->peer_arena) cudaFree(ctx->peer_arena);
    ctx->peer_arena = nullptr;
    ctx->peer_enabled = false;
}

PeerRed peer_red_of(const heat_ctx *ctx) {
    PeerRed pr;
    if (ctx->nranks == 1 && ctx->peer_arena) {      // single GPU: the "all-reduce" is this GPU's own inbox
        pr.P = 1; pr.rank = 0; pr.inbox[0] = ctx->peer_arena;
        return pr;
    }
    if (!ctx->peer_enabled) return pr;
    pr.P = ctx->nranks; pr.rank = ctx->rank;
    for (int q = 0; q < ctx->nranks; ++q) pr.inbox[q] = ctx->peer_arena_of[q];
    return pr;
}

struct PeerRecord {                   // what every rank tells the others about one matrix
    cudaIpcMemHandle_t h_p[4];        // p ping-pong buffers, then the z ping-pong buffers (Chebyshev)
    long long n_owned;
    int n_nbr;
    int nbr_rank[kPeerMaxNbr];
    long long recv_ptr[kPeerMaxNbr + 1];
};

// Collective.  Allocates the ping-pong SpMV-input buffers (w_p, w_p2), exchanges their IPC handles and
// builds the two push plans.  On any failure on any rank the matrix silently stays on the NCCL path.
int peer_matrix_setup(heat_ctx *ctx, heat_matrix *A, bool need_z) {
    if (A->peer && (!need_z || A->peer->has_z)) return 0;
    if (A->peer) peer_matrix_teardown(A);                            // same decision on every rank: redo with the z buffers
    const HaloPlan &h = A->halo;
    const size_t nv = (size_t)(A->n_owned + A->n_ghost);
    const int nbuf = need_z ? 4 : 2;
    DevBuf<double> *bufs[4] = {&A->w_p, &A->w_p2, &A->w_u, &A->w_u2};
    if (ctx->nranks == 1) {
        // degenerate state: no neighbour, the reductions go through this GPU's own inbox
        if (!ctx->peer_arena) {
            HEAT_CUDA(cudaMalloc((void **)&ctx->peer_arena, sizeof(unsigned long long) * kPeerArenaWords));
            HEAT_CUDA(cudaMemset(ctx->peer_arena, 0, sizeof(unsigned long long) * kPeerArenaWords));
            ctx->peer_arena_of[0] = ctx->peer_arena;
        }
        for (int b = 0; b < nbuf; ++b)
            if (!bufs[b]->p) HEAT_TRY(bufs[b]->alloc(nv));
        PeerMatrixState *st = new PeerMatrixState();
        for (int b = 0; b < 4; ++b) { st->push[b].n_nbr = 0; st->push[b].n_blocks = 0; st->push[b].ticket = A->iscal.p + I_COUNTER + 2; }
        st->halo.n_nbr = 0;
        st->halo.flags = ctx->peer_arena + kPeerInboxWords;
        st->has_z = need_z;
        A->peer = st;
        return 0;
    }
    if (!ctx->peer_enabled) return 0;
    int ok = (h.n_neighbors <= kPeerMaxNbr) && spmv_peer_supported() && (A->n_ghost == 0 || A->slices_all.p != nullptr);
    PeerRecord mine;
    memset(&mine, 0, sizeof(mine));
    for (int b = 0; b < nbuf; ++b)
        if (!bufs[b]->p && bufs[b]->alloc(nv)) ok = 0;
    if (ok) {
        for (int b = 0; b < nbuf; ++b)
            if (cudaIpcGetMemHandle(&mine.h_p[b], bufs[b]->p) != cudaSuccess) ok = 0;
        cudaGetLastError();
        mine.n_owned = A->n_owned; mine.n_nbr = h.n_neighbors;
        for (int s = 0; s < h.n_neighbors && s < kPeerMaxNbr; ++s) { mine.nbr_rank[s] = h.nbr_rank[s]; mine.recv_ptr[s] = h.recv_ptr[s]; }
        if (h.n_neighbors <= kPeerMaxNbr) mine.recv_ptr[h.n_neighbors] = h.n_neighbors ? h.recv_ptr[h.n_neighbors] : 0;
    }
    std::vector<unsigned char> all;
    HEAT_TRY(allgather_bytes(ctx, &mine, sizeof(mine), all));
    PeerMatrixState *st = new PeerMatrixState();
    if (ok) {
        for (int b = 0; b < 4; ++b) {
            PeerPush &pp = st->push[b];
            pp.n_nbr = b < nbuf ? h.n_neighbors : 0;
            pp.send_idx = h.d_send_idx.p;
            pp.ticket = A->iscal.p + I_COUNTER + 2;                   // I[4]: push ticket
            for (int s = 0; s <= h.n_neighbors; ++s) pp.send_ptr[s] = h.n_neighbors ? h.send_ptr[s] : 0;
        }
        for (int s = 0; s < h.n_neighbors && ok; ++s) {
            const int q = h.nbr_rank[s];
            PeerRecord rec;
            memcpy(&rec, all.data() + sizeof(rec) * (size_t)q, sizeof(rec));
            int sp = -1;
            for (int t = 0; t < rec.n_nbr && t < kPeerMaxNbr; ++t) if (rec.nbr_rank[t] == ctx->rank) sp = t;
            if (sp < 0 || rec.recv_ptr[sp + 1] - rec.recv_ptr[sp] != h.send_ptr[s + 1] - h.send_ptr[s]) { ok = 0; break; }
            for (int b = 0; b < nbuf; ++b) {
                void *mapped = nullptr;
                if (cudaIpcOpenMemHandle(&mapped, rec.h_p[b], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); break; }
                st->mapped.push_back(mapped);
                st->push[b].dst[s] = (double *)mapped + rec.n_owned + rec.recv_ptr[sp];
                st->push[b].flag[s] = ctx->peer_arena_of[q] + kPeerInboxWords + sp;
            }
        }
    }
    int ok_all = 0;
    HEAT_TRY(all_agree(ctx, ok, &ok_all));
    if (!ok_all) {
        for (void *m : st->mapped) cudaIpcCloseMemHandle(m);
        delete st;
        return 0;
    }
    // Blocks that deliver the halo at the start of the producing kernel.  They finish later than the others by the time
    // their share of the push takes (a gather + a remote store per entry), which is pure tail: with 64 blocks for the
    // 522 240 entries of an interior slab the 8-GPU timeline showed update_p at 107 us against 84 us of streaming.  Up to
    // one resident wave (4 CTAs per SM) now takes part, ~1 K entries each.
    const long long total = h.n_neighbors ? h.send_ptr[h.n_neighbors] : 0;
    int nb = (int)((total + 4 * 256 - 1) / (4 * 256));
    const int nb_cap = 4 * sm_count(ctx->device);
    nb = nb < 1 ? 1 : nb > nb_cap ? nb_cap : nb;
    for (int b = 0; b < 4; ++b) st->push[b].n_blocks = nb;
    st->halo.n_nbr = h.n_neighbors;
    st->halo.flags = ctx->peer_arena + kPeerInboxWords;
    st->has_z = need_z;
    A->peer = st;
    return 0;
}

void peer_matrix_teardown(heat_matrix *A) {
    if (!A->peer) return;
    for (void *m : A->peer->mapped) cudaIpcCloseMemHandle(m);
    delete A->peer;
    A->peer = nullptr;
}

// Gather every rank's owned values on rank 0 in reduced-id order (the blocking Send/Recv gather of
// IO::writeSolution, ExodusIO.hpp:2000-2026, as grouped NCCL send/recv).
int comm_gather_reduced(heat_ctx *ctx, const std::vector<double> &xl, int64_t n_global, std::vector<double> &xg) {
    if (ctx->nranks <= 1) { xg = xl; return 0; }
    const int P = ctx->nranks, me = ctx->rank;
    const int64_t nl = (int64_t)xl.size();
    if ((int64_t)ctx->owned_gids.size() != nl) HEAT_FAIL(42, "gather: vector does not match the assembled row map");
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    cudaStream_t st = ctx->stream;
    DevBuf<int64_t> d_counts;
    HEAT_TRY(d_counts.alloc((size_t)P));
    HEAT_CUDA(cudaMemcpyAsync(d_counts.p + me, &nl, sizeof(int64_t), cudaMemcpyHostToDevice, st));
    HEAT_NCCL(g_nccl.AllGather(d_counts.p + me, d_counts.p, 1, ncclInt64, comm, st));
    std::vector<int64_t> counts((size_t)P);
    HEAT_CUDA(cudaMemcpyAsync(counts.data(), d_counts.p, sizeof(int64_t) * (size_t)P, cudaMemcpyDeviceToHost, st));
    HEAT_CUDA(cudaStreamSynchronize(st));
    DevBuf<double> d_val; DevBuf<int64_t> d_gid;
    HEAT_TRY(d_val.alloc((size_t)nl)); HEAT_TRY(d_gid.alloc((size_t)nl));
    if (nl) {
        HEAT_CUDA(cudaMemcpyAsync(d_val.p, xl.data(), sizeof(double) * (size_t)nl, cudaMemcpyHostToDevice, st));
        HEAT_CUDA(cudaMemcpyAsync(d_gid.p, ctx->owned_gids.data(), sizeof(int64_t) * (size_t)nl, cudaMemcpyHostToDevice, st));
    }
    if (me != 0) {
        HEAT_NCCL(g_nccl.GroupStart());
        if (nl) {
            HEAT_NCCL(g_nccl.Send(d_val.p, (size_t)nl, ncclDouble, 0, comm, st));
            HEAT_NCCL(g_nccl.Send(d_gid.p, (size_t)nl, ncclInt64, 0, comm, st));
        }
        HEAT_NCCL(g_nccl.GroupEnd());
        HEAT_CUDA(cudaStreamSynchronize(st));
        return 0;
    }
    int64_t total = 0;
    for (int64_t c : counts) total += c;
    if (total != n_global) HEAT_FAIL(42, "gather: ranks own %lld rows, system has %lld", (long long)total, (long long)n_global);
    DevBuf<double> d_allv; DevBuf<int64_t> d_allg;
    HEAT_TRY(d_allv.alloc((size_t)total)); HEAT_TRY(d_allg.alloc((size_t)total));
    HEAT_NCCL(g_nccl.GroupStart());
    int64_t off = counts[0];
    for (int q = 1; q < P; ++q) {
        if (counts[(size_t)q]) {
            HEAT_NCCL(g_nccl.Recv(d_allv.p + off, (size_t)counts[(size_t)q], ncclDouble, q, comm, st));
            HEAT_NCCL(g_nccl.Recv(d_allg.p + off, (size_t)counts[(size_t)q], ncclInt64, q, comm, st));
        }
        off += counts[(size_t)q];
    }
    HEAT_NCCL(g_nccl.GroupEnd());
    if (nl) {
        HEAT_CUDA(cudaMemcpyAsync(d_allv.p, d_val.p, sizeof(double) * (size_t)nl, cudaMemcpyDeviceToDevice, st));
        HEAT_CUDA(cudaMemcpyAsync(d_allg.p, d_gid.p, sizeof(int64_t) * (size_t)nl, cudaMemcpyDeviceToDevice, st));
    }
    std::vector<double> hv((size_t)total);
    std::vector<int64_t> hg((size_t)total);
    HEAT_CUDA(cudaMemcpyAsync(hv.data(), d_allv.p, sizeof(double) * (size_t)total, cudaMemcpyDeviceToHost, st));
    HEAT_CUDA(cudaMemcpyAsync(hg.data(), d_allg.p, sizeof(int64_t) * (size_t)total, cudaMemcpyDeviceToHost, st));
    HEAT_CUDA(cudaStreamSynchronize(st));
    xg.assign((size_t)n_global, 0.0);
    for (int64_t t = 0; t < total; ++t) {
        if (hg[(size_t)t] < 0 || hg[(size_t)t] >= n_global) HEAT_FAIL(42, "gather: bad global id");
        xg[(size_t)hg[(size_t)t]] = hv[(size_t)t];
    }
    return 0;
}

}  // namespace heat

extern "C" int heat_comm_unique_id(char id_out[HEAT_COMM_ID_BYTES]) {
    HEAT_TRY(heat::nccl_load());
    ncclUniqueId id;
    HEAT_NCCL(heat::g_nccl.GetUniqueId(&id));
    memcpy(id_out, id.internal, HEAT_COMM_ID_BYTES);
    return 0;
}

extern "C" int heat_comm_init(heat_ctx *ctx, int rank, int nranks, const char id_in[HEAT_COMM_ID_BYTES]) {
    if (!ctx || nranks < 1 || rank < 0 || rank >= nranks) HEAT_FAIL(2, "heat_comm_init: bad arguments");
    if (nranks == 1) { ctx->rank = 0; ctx->nranks = 1; return 0; }
    if (ctx->device < 0) HEAT_FAIL(3, "heat_comm_init: host-only context (no CUDA device)");
    HEAT_TRY(heat::nccl_load());
    HEAT_CUDA(cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(id.internal, id_in, HEAT_COMM_ID_BYTES);
    ncclComm_t comm;
    HEAT_NCCL(heat::g_nccl.CommInitRank(&comm, nranks, id, rank));
    ctx->nccl_comm = comm;
    ctx->rank = rank;
    ctx->nranks = nranks;
    if (!ctx->ev_halo) HEAT_FAIL(3, "heat_comm_init: context has no CUDA resources");
    if (!ctx->comm_stream) {
        int lo = 0, hi = 0;
        HEAT_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        HEAT_CUDA(cudaStreamCreateWithPriority(&ctx->comm_stream, cudaStreamNonBlocking, hi));
    }
    HEAT_TRY(heat::peer_arena_setup(ctx));
    return 0;
}

extern "C" int heat_comm_rank(const heat_ctx *ctx, int *rank, int *nranks) {
    if (!ctx) HEAT_FAIL(2, "heat_comm_rank: null ctx");
    if (rank) *rank = ctx->rank;
    if (nranks) *nranks = ctx->nranks;
    return 0;
}
