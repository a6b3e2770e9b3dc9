// partition.cpp — host-side row partitioning and owned/ghost/send maps (pure CPU).
//
// Reference roles: Tpetra::Map (contiguous uniform map, ExodusIO.hpp:171/:252), Zoltan2 "parmetis"
// repartition of the matrix row graph (:644-656), and the column map + Import plan that
// Tpetra::CrsMatrix::fillComplete builds (:609): owned rows first (ascending global id), then
// ghost columns grouped by owner rank ascending, ascending global id inside an owner.
// ParMETIS/Zoltan2 do not exist in this image and their output depends on rank count and build;
// the k-way METIS the CUDA toolkit ships (idx_t=int64, real_t=float) is called on the same graph
// instead — "bit-exact partition" means bit-exact against THAT library.
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <vector>

#include "common.cuh"

extern "C" {
// libmetis_static.a (CUDA 12.9 toolkit): 64-bit idx_t, 32-bit real_t; no header ships with it.
int METIS_PartGraphKway(int64_t *nvtxs, int64_t *ncon, int64_t *xadj, int64_t *adjncy, int64_t *vwgt,
                        int64_t *vsize, int64_t *adjwgt, int64_t *nparts, float *tpwgts, float *ubvec,
                        int64_t *options, int64_t *objval, int64_t *part);
int METIS_PartMeshDual(int64_t *ne, int64_t *nn, int64_t *eptr, int64_t *eind, int64_t *vwgt, int64_t *vsize,
                       int64_t *ncommon, int64_t *nparts, float *tpwgts, int64_t *options, int64_t *objval,
                       int64_t *epart, int64_t *npart);
}

namespace heat {

// Tpetra::Map(numGlobal, indexBase, comm) uniform contiguous distribution: the first
// (n % P) ranks get one extra row.
void contiguous_partition(int64_t n, int nranks, int32_t *part) {
    const int64_t base = n / nranks, rem = n % nranks;
    int64_t row = 0;
    for (int p = 0; p < nranks; ++p) {
        const int64_t cnt = base + (p < rem ? 1 : 0);
        for (int64_t q = 0; q < cnt; ++q) part[row++] = p;
    }
}

int metis_kway_partition(int64_t n, const int64_t *row_ptr, const int32_t *col, int nranks, int32_t *part) {
    if (nranks == 1 || n == 0) {
        std::fill(part, part + n, 0);
        return 0;
    }
    std::vector<int64_t> xadj((size_t)n + 1, 0), adj;
    adj.reserve((size_t)row_ptr[n]);
    for (int64_t i = 0; i < n; ++i) {
        for (int64_t q = row_ptr[i]; q < row_ptr[i + 1]; ++q)
            if (col[q] != i) adj.push_back(col[q]);
        xadj[(size_t)i + 1] = (int64_t)adj.size();
    }
    int64_t nv = n, ncon = 1, np = nranks, objval = 0;
    std::vector<int64_t> p64((size_t)n, 0);
    int rc = METIS_PartGraphKway(&nv, &ncon, xadj.data(), adj.data(), nullptr, nullptr, nullptr, &np, nullptr,
                                 nullptr, nullptr, &objval, p64.data());
    if (rc != 1) HEAT_FAIL(30, "METIS_PartGraphKway failed (rc=%d)", rc);
    for (int64_t i = 0; i < n; ++i) part[i] = (int32_t)p64[(size_t)i];
    return 0;
}

int metis_part_mesh_dual(int64_t ne, int64_t nn, int npe, const int32_t *conn, int64_t ncommon, int64_t nparts,
                         int64_t *objval, int64_t *epart, int64_t *npart) {
    std::vector<int64_t> eptr((size_t)ne + 1), eind((size_t)(ne * npe));
    for (int64_t e = 0; e <= ne; ++e) eptr[(size_t)e] = e * npe;
    for (int64_t q = 0; q < ne * npe; ++q) eind[(size_t)q] = conn[q];
    int64_t ne_ = ne, nn_ = nn, nc = ncommon, np = nparts;
    int rc = METIS_PartMeshDual(&ne_, &nn_, eptr.data(), eind.data(), nullptr, nullptr, &nc, &np, nullptr, nullptr,
                                objval, epart, npart);
    if (rc != 1) HEAT_FAIL(31, "METIS_PartMeshDual failed (rc=%d)", rc);
    return 0;
}

}  // namespace heat

// Node ownership of IO::getMatrix (ExodusIO.hpp:1089-1295).  Each rank holds the elements METIS gave
// it; a node shared by several ranks goes to the rank in whose LOCAL adjacency (cliques of its own
// elements, :1089-1097) the node has the most distinct neighbours (nodeToFreq, :1163-1168), ties to
// the lowest rank (:1268-1270).  The reference finds the sharers with one-sided MPI_Gets and
// exchanges (node, freq) lists (with the request-array overflow D9); the rule itself is a pure
// function of (connectivity, epart), evaluated here for every node at once.  Nodes in no element
// (absent from every rank's nodeIndices in the reference) are given to rank 0.
extern "C" int heat_node_owners(int64_t num_nodes, int64_t num_elem, int npe, const int32_t *conn,
                                const int64_t *epart, int nparts, int32_t *owner_out) {
    if (num_nodes < 0 || num_elem < 0 || npe < 1 || nparts < 1 || !owner_out || (num_elem > 0 && (!conn || !epart)))
        HEAT_FAIL(2, "heat_node_owners: bad arguments");
    std::vector<int64_t> ptr((size_t)num_nodes + 1, 0);
    for (int64_t q = 0; q < num_elem * npe; ++q) {
        if (conn[q] < 0 || conn[q] >= num_nodes) HEAT_FAIL(2, "heat_node_owners: connectivity entry out of range");
        ptr[(size_t)conn[q] + 1]++;
    }
    for (int64_t e = 0; e < num_elem; ++e)
        if (epart[e] < 0 || epart[e] >= nparts) HEAT_FAIL(2, "heat_node_owners: epart[%lld] = %lld out of range", (long long)e, (long long)epart[e]);
    for (int64_t v = 0; v < num_nodes; ++v) ptr[(size_t)v + 1] += ptr[(size_t)v];
    std::vector<int64_t> n2e((size_t)(num_elem * npe)), fill(ptr.begin(), ptr.end() - 1);
    for (int64_t e = 0; e < num_elem; ++e)
        for (int k = 0; k < npe; ++k) n2e[(size_t)fill[(size_t)conn[e * npe + k]]++] = e;
    std::vector<int64_t> keys;                       // (part << 40 | neighbour), distinct per node
    std::vector<int64_t> freq((size_t)nparts);
    for (int64_t v = 0; v < num_nodes; ++v) {
        keys.clear();
        for (int64_t t = ptr[(size_t)v]; t < ptr[(size_t)v + 1]; ++t) {
            const int64_t e = n2e[(size_t)t];
            for (int k = 0; k < npe; ++k) {
                const int64_t u = conn[e * npe + k];
                if (u != v) keys.push_back((epart[e] << 40) | u);
            }
        }
        std::sort(keys.begin(), keys.end());
        keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
        std::fill(freq.begin(), freq.end(), -1);     // -1: the rank does not hold the node at all
        for (int64_t t = ptr[(size_t)v]; t < ptr[(size_t)v + 1]; ++t) {
            int64_t &f = freq[(size_t)epart[(size_t)n2e[(size_t)t]]];
            if (f < 0) f = 0;
        }
        for (int64_t key : keys) freq[(size_t)(key >> 40)]++;
        int32_t best = 0;
        for (int p = 1; p < nparts; ++p)
            if (freq[(size_t)p] > freq[(size_t)best]) best = p;     // strict '>' keeps the lowest rank on ties
        owner_out[v] = best;
    }
    return 0;
}

extern "C" int heat_partition_rows(int64_t n_global, const int64_t *row_ptr, const int32_t *col, int partitioner,
                                   int nranks, int32_t *part_out) {
    if (nranks < 1 || n_global < 0 || !part_out) HEAT_FAIL(2, "heat_partition_rows: bad arguments");
    if (partitioner == HEAT_PART_METIS_KWAY) {
        if (!row_ptr || !col) HEAT_FAIL(2, "heat_partition_rows: METIS needs the row graph");
        return heat::metis_kway_partition(n_global, row_ptr, col, nranks, part_out);
    }
    if (partitioner == HEAT_PART_CONTIGUOUS) {
        heat::contiguous_partition(n_global, nranks, part_out);
        return 0;
    }
    HEAT_FAIL(2, "heat_partition_rows: partitioner %d needs a mesh (use heat_assemble)", partitioner);
}

extern "C" int heat_plan_build(int64_t n, const int64_t *row_ptr, const int32_t *col, const int32_t *part,
                               int nranks, int rank, heat_plan_sizes *sizes, int64_t *owned_gids,
                               int64_t *ghost_gids, int32_t *ghost_owner, int32_t *nbr_rank, int64_t *send_ptr,
                               int64_t *send_gids, int64_t *recv_ptr) {
    if (!row_ptr || !col || !part || rank < 0 || rank >= nranks) HEAT_FAIL(2, "heat_plan_build: bad arguments");
    // ghosts of a rank q = columns of q's rows owned by someone else, sorted by (owner, gid)
    auto ghosts_of = [&](int q, std::vector<int64_t> &gh) {
        std::vector<int64_t> keys;
        for (int64_t i = 0; i < n; ++i) {
            if (part[i] != q) continue;
            for (int64_t t = row_ptr[i]; t < row_ptr[i + 1]; ++t) {
                const int32_t c = col[t];
                if (part[c] != q) keys.push_back(((int64_t)part[c] << 40) | (int64_t)c);
            }
        }
        std::sort(keys.begin(), keys.end());
        keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
        gh.resize(keys.size());
        for (size_t t = 0; t < keys.size(); ++t) gh[t] = keys[t] & ((1ll << 40) - 1);
    };
    std::vector<int64_t> my_ghosts;
    ghosts_of(rank, my_ghosts);
    int64_t n_owned = 0;
    for (int64_t i = 0; i < n; ++i) n_owned += (part[i] == rank);
    // neighbours: ranks I receive from or send to (symmetric pattern => same set; keep the union)
    std::vector<std::vector<int64_t>> send_lists((size_t)nranks);
    for (int q = 0; q < nranks; ++q) {
        if (q == rank) continue;
        std::vector<int64_t> gq;
        ghosts_of(q, gq);
        for (int64_t g : gq)
            if (part[g] == rank) send_lists[(size_t)q].push_back(g);       // q's ghost order
    }
    std::vector<int64_t> recv_cnt((size_t)nranks, 0);
    for (int64_t g : my_ghosts) recv_cnt[(size_t)part[g]]++;
    std::vector<int32_t> nbrs;
    for (int q = 0; q < nranks; ++q)
        if (q != rank && (recv_cnt[(size_t)q] > 0 || !send_lists[(size_t)q].empty())) nbrs.push_back(q);
    int64_t n_send = 0;
    for (int32_t q : nbrs) n_send += (int64_t)send_lists[(size_t)q].size();
    if (sizes) {
        sizes->n_owned = n_owned; sizes->n_ghost = (int64_t)my_ghosts.size();
        sizes->n_neighbors = (int32_t)nbrs.size(); sizes->n_send = n_send;
    }
    if (owned_gids) {
        int64_t w = 0;
        for (int64_t i = 0; i < n; ++i)
            if (part[i] == rank) owned_gids[w++] = i;
    }
    if (ghost_gids) std::copy(my_ghosts.begin(), my_ghosts.end(), ghost_gids);
    if (ghost_owner)
        for (size_t t = 0; t < my_ghosts.size(); ++t) ghost_owner[t] = part[my_ghosts[t]];
    if (nbr_rank) std::copy(nbrs.begin(), nbrs.end(), nbr_rank);
    if (send_ptr && recv_ptr) {
        send_ptr[0] = 0; recv_ptr[0] = 0;
        int64_t w = 0;
        for (size_t s = 0; s < nbrs.size(); ++s) {
            const auto &lst = send_lists[(size_t)nbrs[s]];
            if (send_gids) std::copy(lst.begin(), lst.end(), send_gids + w);
            w += (int64_t)lst.size();
            send_ptr[s + 1] = w;
            recv_ptr[s + 1] = recv_ptr[s] + recv_cnt[(size_t)nbrs[s]];
        }
    }
    return 0;
}
