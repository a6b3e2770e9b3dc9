// ExodusIO_b200.hpp — C++ mirror of the reference's `ExodusIO::IO` (ExodusIO.hpp:83-2225) and
// `belosSolver` (BelosMueLuSolver.cpp:87) over the C ABI of include/heat_b200.h.
//
// Same method names, argument order, bool returns and call-order contract as the reference:
//     IO io; io.open(in, true); io.assemble(&A, &X, &B); io.create(out); io.decompose(parts);
//     belosSolver(A, X, B, iterations, tolerance, io, verbose);   // calls io.writeSolution
// Tpetra cannot exist here, so Teuchos::RCP<Tpetra::CrsMatrix<>> / MultiVector<> are replaced by
// ref-counted handles heat::Matrix / heat::Vector (std::shared_ptr, same ownership semantics).
// One IO per process == per GPU (the reference: one IO per MPI rank).
#pragma once
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <iostream>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

#include "heat_b200.h"

namespace heat {

struct MatrixHandle {
    heat_matrix *h = nullptr;
    ~MatrixHandle() { heat_matrix_free(h); }
};
struct VectorHandle {
    heat_vector *h = nullptr;
    ~VectorHandle() { heat_vector_free(h); }
};
using Matrix = std::shared_ptr<MatrixHandle>;   // role of Teuchos::RCP<Tpetra::CrsMatrix<>>
using Vector = std::shared_ptr<VectorHandle>;   // role of Teuchos::RCP<Tpetra::MultiVector<>>

struct Options {                 // knobs the reference hard-codes (ExodusIO.hpp:644-646, BelosMueLuSolver.cpp:93-106)
    int device = 0;
    int op_mode = HEAT_OP_GRAPH_LAPLACIAN;
    int partitioner = HEAT_PART_METIS_KWAY;
    int solver = HEAT_SOLVER_CG;
    int prec = HEAT_PREC_JACOBI;
    int cheb_degree = 3;
    double cheb_lambda_max = 0.0;
    int write_every = 0;         // 0: write the final field only; k: every k iterations (the reference: 1)
    int gmres_restart = 300;     // HEAT_SOLVER_GMRES: Belos "Num Blocks"
    int output_word_size = 8;    // 4: write float32 records like a reference built on a float real_t METIS (ExodusIO.hpp:104-105)
    bool output_largest_nodeset_id = false;   // nodes in several nodesets: largest id in the output field (:1983-1989)
    bool literal_loop = false;   // reproduce the reference's loop literally: numIterations solves with
                                 // "Maximum Iterations" = 1, field written after each (BelosMueLuSolver.cpp:102,113-133)
};

}  // namespace heat

namespace ExodusIO {

class IO {
   public:
    explicit IO(const heat::Options &opt = heat::Options()) : opt_(opt) {
        if (heat_ctx_create(opt.device, &ctx_)) std::cerr << "heat_ctx_create: " << heat_last_error() << std::endl;
        else if (heat_ctx_set_output(ctx_, opt.output_word_size, opt.output_largest_nodeset_id ? 1 : 0)) perror_("heat_ctx_set_output");
    }
    IO(const IO &) = delete;
    IO &operator=(const IO &) = delete;
    ~IO() { heat_close(ctx_); }                                   // ExodusIO.hpp:2072-2079

    // Opens the read-in Exodus file (ExodusIO.hpp:88-100)
    bool open(std::string fname, bool read_only = false) {
        if (!ctx_ || heat_open(ctx_, fname.c_str(), read_only ? 1 : 0)) { perror_("ex_open"); return false; }
        return true;
    }
    // Opens the written-out Exodus file (ExodusIO.hpp:103-114)
    bool create(std::string fname) {
        if (!ctx_ || heat_create(ctx_, fname.c_str())) { perror_("ex_create"); return false; }
        return true;
    }
    // ExodusIO.hpp:128 — A, X, B for the steady-state heat problem of the opened mesh
    bool assemble(heat::Matrix *A, heat::Vector *X, heat::Vector *B, bool verbose = false) {
        if (!ctx_) return false;
        auto a = std::make_shared<heat::MatrixHandle>();
        auto x = std::make_shared<heat::VectorHandle>();
        auto b = std::make_shared<heat::VectorHandle>();
        if (heat_assemble(ctx_, opt_.op_mode, opt_.partitioner, &a->h, &x->h, &b->h)) { perror_("assemble"); return false; }
        if (verbose) {
            heat_matrix_info mi;
            heat_matrix_get_info(a->h, &mi);
            std::cout << "# of Nodes: " << mi.num_nodes << "\n# of Elements: " << mi.num_elem << "\n# of Node Sets: "
                      << mi.num_node_sets << "\nDOF rows: " << mi.n_global << " (rank " << mi.rank << " owns " << mi.n_owned
                      << ", ghosts " << mi.n_ghost << ")\nnnz: " << mi.nnz_global << "\nassemble ms: " << mi.assemble_ms
                      << std::endl;
        }
        *A = a; *X = x; *B = b;
        return true;
    }
    // ExodusIO.hpp:733 — Laplacian of the whole mesh (singular: nodesets are NOT applied), rows
    // distributed by element partition + the node-ownership rule; nodeSetMap receives, per nodeset
    // id, the nodes whose rows this rank owns (0-based here; the reference keeps Exodus' 1-based ids)
    bool getMatrix(heat::Matrix *ret, std::map<int, std::set<int64_t>> &nodeSetMap, bool verbose = false) {
        if (!ctx_) return false;
        auto a = std::make_shared<heat::MatrixHandle>();
        if (heat_get_matrix(ctx_, opt_.op_mode, &a->h)) { perror_("getMatrix"); return false; }
        int nsets = 0;
        if (heat_mesh_nodeset_ids(ctx_, &nsets, nullptr)) { perror_("getMatrix"); return false; }
        std::vector<int64_t> ids((size_t)nsets);
        if (nsets > 0) heat_mesh_nodeset_ids(ctx_, &nsets, ids.data());
        for (int64_t id : ids) {
            int64_t cnt = 0;
            if (heat_matrix_owned_nodeset(a->h, id, &cnt, nullptr)) { perror_("getMatrix"); return false; }
            std::vector<int64_t> nodes((size_t)cnt);
            if (cnt > 0) heat_matrix_owned_nodeset(a->h, id, &cnt, nodes.data());
            nodeSetMap[(int)id].insert(nodes.begin(), nodes.end());
        }
        if (verbose) {
            heat_matrix_info mi;
            heat_matrix_get_info(a->h, &mi);
            std::cout << "Rank #" << mi.rank << ": " << mi.n_owned << " of " << mi.n_global << " rows, " << mi.n_ghost
                      << " ghost columns, " << mi.nnz_local << " entries" << std::endl;
        }
        *ret = a;
        return true;
    }
    // ExodusIO.hpp:1496 — partition with METIS and write the mesh with one element block per partition
    bool decompose(int partitions, bool verbose = false) {
        if (!ctx_ || heat_decompose(ctx_, partitions)) { perror_("decompose"); return false; }
        if (verbose) std::cout << "Decomposed into " << partitions << " partitions." << std::endl;
        return true;
    }
    // ExodusIO.hpp:1972 — nodal variable "Steady-State Heat Solution" at step timestep+1
    bool writeSolution(heat::Vector vec, const int timestep, bool verbose = false) {
        if (!ctx_ || !vec || heat_write_solution(ctx_, vec->h, timestep)) { perror_("writeSolution"); return false; }
        if (verbose) std::cout << "Wrote nodal values for timestep " << timestep << "!" << std::endl;
        return true;
    }

    heat_ctx *ctx() { return ctx_; }
    const heat::Options &options() const { return opt_; }

   private:
    void perror_(const char *what) { std::cerr << what << ": " << heat_last_error() << std::endl; }
    heat_ctx *ctx_ = nullptr;
    heat::Options opt_;
};

}  // namespace ExodusIO

// The reference's debug dump (BelosMueLuSolver.cpp:28-84): one line per owned row with GLOBAL ids,
// "row: [(col,val),(col,val),...] ~timestamp~" for a matrix and "row: [val] ~timestamp~" for a vector,
// columns sorted ascending; the per-rank files "<prefix><rank>.out" with their "[Section]" headers are
// merged by the reference's mpi_output_combiner.py, which orders lines by the microsecond timestamp.
namespace heat {
inline uint64_t getTime() {
    return (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::system_clock::now().time_since_epoch()).count();
}
inline void printCrsMatrix(const Matrix &A, std::ostream &output) {
    heat_matrix_info mi;
    heat_matrix_get_info(A->h, &mi);
    std::vector<int64_t> rp((size_t)mi.n_owned + 1), owned((size_t)mi.n_owned + 1), ghost((size_t)mi.n_ghost + 1);
    std::vector<int32_t> col((size_t)mi.nnz_local + 1), owner((size_t)mi.n_ghost + 1);
    std::vector<double> val((size_t)mi.nnz_local + 1);
    heat_matrix_export_csr(A->h, rp.data(), col.data(), val.data());
    heat_matrix_export_maps(A->h, owned.data(), ghost.data(), owner.data());
    std::vector<std::pair<int64_t, double>> entries;
    for (int64_t r = 0; r < mi.n_owned; ++r) {
        entries.clear();
        for (int64_t q = rp[(size_t)r]; q < rp[(size_t)r + 1]; ++q) {
            const int32_t c = col[(size_t)q];
            entries.emplace_back(c < mi.n_owned ? owned[(size_t)c] : ghost[(size_t)(c - mi.n_owned)], val[(size_t)q]);
        }
        std::sort(entries.begin(), entries.end());
        output << owned[(size_t)r] << ": [";
        for (size_t i = 0; i < entries.size(); ++i) output << (i ? "," : "") << "(" << entries[i].first << "," << entries[i].second << ")";
        output << "] ~" << getTime() << "~" << std::endl;
    }
}
inline void printMultiVector(heat_ctx *ctx, const Matrix &A, const Vector &X, std::ostream &output) {
    heat_matrix_info mi;
    heat_matrix_get_info(A->h, &mi);
    std::vector<int64_t> owned((size_t)mi.n_owned + 1);
    std::vector<double> x((size_t)mi.n_owned + 1);
    heat_matrix_export_maps(A->h, owned.data(), nullptr, nullptr);
    heat_vector_get(ctx, X->h, x.data(), mi.n_owned);
    for (int64_t r = 0; r < mi.n_owned; ++r) output << owned[(size_t)r] << ": [" << x[(size_t)r] << "] ~" << getTime() << "~" << std::endl;
}
}  // namespace heat

// Solves A x = b (BelosMueLuSolver.cpp:87-139).  The reference runs Belos GMRES(1)+ILUT restarted in a
// loop and writes the field after every iteration; here the Krylov loop is device-resident PCG and
// the field is written every `write_every` iterations (and at the end) without restarting it.
inline void belosSolver(const heat::Matrix A, const heat::Vector X, const heat::Vector B, size_t numIterations,
                        double tolerance, ExodusIO::IO &io, bool verbose) {
    const heat::Options &opt = io.options();
    heat_solve_opts o;
    heat_solve_opts_default(&o);
    o.solver = opt.solver; o.prec = opt.prec; o.tol = tolerance;
    o.cheb_degree = opt.cheb_degree; o.cheb_lambda_max = opt.cheb_lambda_max;
    o.gmres_restart = opt.gmres_restart;
    int rank = 0, nranks = 1;
    heat_comm_rank(io.ctx(), &rank, &nranks);
    size_t iterations = 0;
    heat_solve_info info{};
    bool converged = false;
    o.max_iters = (int)numIterations;
    if (opt.literal_loop) {
        // the reference's loop as written: every pass is a fresh solve() limited to ONE iteration, starting
        // from the current X, with its own initial residual as the reference of the relative test (so the
        // test can only fire if one iteration gains `tolerance`), and the field is written after each pass
        o.max_iters = 1;
        for (size_t i = 0; i < numIterations; ++i) {
            if (heat_solve(io.ctx(), A->h, X->h, B->h, &o, &info)) { std::cerr << heat_last_error() << std::endl; return; }
            io.writeSolution(X, (int)i, verbose);
            ++iterations;
            if (info.converged || info.achieved_tol <= tolerance) { converged = info.converged != 0; break; }
        }
        if (rank == 0)        // the reference prints this line unconditionally after the loop (D8)
            std::cout << "The Belos solve took " << iterations << " iteration(s), but did not converge. Achieved tolerance = "
                      << info.achieved_tol << "." << std::endl;
        return;
    }
    if (opt.write_every > 0) {
        // trajectory mode (the reference writes after EVERY pass of its loop, :114-117): one Krylov run,
        // the iterate is written every write_every iterations as time steps 0, 1, 2, ...
        int frames = 0;
        if (heat_solve_trajectory(io.ctx(), A->h, X->h, B->h, &o, opt.write_every, 0, &info, &frames)) {
            std::cerr << heat_last_error() << std::endl;
            return;
        }
        if (verbose && rank == 0) std::cout << "Wrote " << frames << " time steps." << std::endl;
    } else {
        if (heat_solve(io.ctx(), A->h, X->h, B->h, &o, &info)) { std::cerr << heat_last_error() << std::endl; return; }
        io.writeSolution(X, 0, verbose);
    }
    iterations = (size_t)info.iters;
    converged = info.converged != 0;
    if (rank == 0) {
        if (converged)
            std::cout << "The Belos solve took " << iterations << " iteration(s) to reach a relative residual tolerance of "
                      << info.achieved_tol << "." << std::endl;
        else
            std::cout << "The Belos solve took " << iterations << " iteration(s), but did not converge. Achieved tolerance = "
                      << info.achieved_tol << "." << std::endl;
    }
}
