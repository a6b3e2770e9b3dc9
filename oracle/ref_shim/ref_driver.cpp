// ref_driver.cpp — runs the REFERENCE's own ExodusIO.hpp (compiled where it lies, /root/reference/ExodusIO.hpp,
// unmodified) on one rank and dumps what it computed, so that the oracle restatement can be pinned against the
// reference itself.  TEST INFRASTRUCTURE (oracle/ref_shim/README.md): nothing here is product code, and none of
// the product's or the oracle's sources are linked in.
//
// Call order = main() of BelosMueLuSolver.cpp:165-211:  open(read_only) -> assemble -> create -> decompose(max(2,
// ranks)) -> writeSolution(X, i) for i = 0, 1 (the Belos solve between them cannot run here: X is a fixed,
// float-exact function of the row id instead).  A second IO object does what main() of ExodusMatrixTest.cpp:131-176
// does: getMatrix on one rank, then the reference's own PowerMethod<CrsMatrix<>>::run(A, 500, 1e-2) (compiled from
// ExodusMatrixTest.cpp, its main() renamed; the vector operations it calls are the sequential loops of the stand-in
// Tpetra, the start vector is the seeded counter-based one instead of an unseeded randomize()).
//
//   REF_SHIM_DUMP_DIR=<dir with mesh.exo.dump>  ref_driver <mesh.exo> <out_prefix> [nparts=2] [--no-getmatrix]
//
// Outputs (record container of exodus_shim.cpp):
//   <out_prefix>.assemble.dump        n, row ids, CSR of A (columns ascending), B, reduced->original id map, nodesets
//   <out_prefix>.mpi-proc-0.out       A and B printed by the reference's printCrsMatrix / printMultiVector (ref_dump.cpp)
//   <out_prefix>.timing               wall seconds of the reference's assemble / decompose / writeSolution / getMatrix calls
//   <out_prefix>.getmatrix.dump       the same for getMatrix, + the power method's lambda and its printed log
//   <out_prefix>.solution.exo.shimdump   everything create/decompose/writeSolution handed to the Exodus API
// every header ExodusIO.hpp pulls in goes first, so that the access hack below touches ExodusIO.hpp alone
#include <Tpetra_Core.hpp>
#include <Zoltan2_Adapter.hpp>
#include <parmetis.h>
#include "exodusII.h"
#include <algorithm>
#include <atomic>
#include <cassert>
#include <cstdlib>
#include <iostream>
#include <set>
#include <sstream>
#include <thread>
#include <utility>
#include <Teuchos_CommandLineProcessor.hpp>
#define private public            // the id map and the nodeset cache are private members of ExodusIO::IO
#define main reference_matrix_test_main
#include "ExodusMatrixTest.cpp"   // includes ExodusIO.hpp (which has no include guard) exactly once
#undef main
#undef private

#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

bool ref_dump_system(const Teuchos::RCP<Tpetra::CrsMatrix<>> &A, const Teuchos::RCP<Tpetra::MultiVector<>> &B, const std::string &path);   // ref_dump.cpp

namespace {

void rec(FILE *fp, const char *name, const char *dtype, const void *data, size_t count, size_t item) {
    std::fprintf(fp, "%s %s %llu\n", name, dtype, (unsigned long long)count);
    if (count) std::fwrite(data, item, count, fp);
}
void rec_i32(FILE *fp, const char *name, const std::vector<int32_t> &v) { rec(fp, name, "i32", v.data(), v.size(), 4); }
void rec_f64(FILE *fp, const char *name, const std::vector<double> &v) { rec(fp, name, "f64", v.data(), v.size(), 8); }

void dump_matrix(FILE *fp, const Teuchos::RCP<Tpetra::CrsMatrix<>> &A) {
    std::vector<int32_t> rows, rowptr{0}, cols;
    std::vector<double> vals;
    for (const auto &r : A->shim_rows()) {
        rows.push_back((int32_t)r.first);
        for (const auto &cv : r.second) { cols.push_back((int32_t)cv.first); vals.push_back(cv.second); }
        rowptr.push_back((int32_t)cols.size());
    }
    const std::vector<int32_t> n{(int32_t)A->getRowMap()->getGlobalNumElements()};
    rec_i32(fp, "n", n);
    rec_i32(fp, "A_rows", rows);
    rec_i32(fp, "A_rowptr", rowptr);
    rec_i32(fp, "A_cols", cols);
    rec_f64(fp, "A_vals", vals);
}
void dump_nodesets(FILE *fp, const std::map<int, std::set<idx_t>> &ns) {
    std::vector<int32_t> ids;
    for (const auto &s : ns) ids.push_back(s.first);
    rec_i32(fp, "ns_ids", ids);
    for (const auto &s : ns) {
        const std::vector<int32_t> nodes(s.second.begin(), s.second.end());
        rec_i32(fp, ("ns_" + std::to_string(s.first)).c_str(), nodes);
    }
}

double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

}  // namespace

int main(int argc, char **argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: ref_driver <mesh.exo> <out_prefix> [nparts] [--no-getmatrix]\n"); return 2; }
    const std::string input = argv[1], prefix = argv[2];
    int nparts = 2;
    bool get_matrix = true;
    for (int i = 3; i < argc; ++i) {
        if (!std::strcmp(argv[i], "--no-getmatrix")) get_matrix = false;
        else nparts = std::atoi(argv[i]);
    }
    {
        ExodusIO::IO io;
        if (!io.open(input, true)) return 1;
        Teuchos::RCP<Tpetra::CrsMatrix<>> A;
        Teuchos::RCP<Tpetra::MultiVector<>> X, B;
        const double t0 = now();
        if (!io.assemble(&A, &X, &B, false)) { std::fprintf(stderr, "ref_driver: assemble failed\n"); return 1; }
        const double t_assemble = now() - t0;
        {
            FILE *tf = std::fopen((prefix + ".timing").c_str(), "w");
            if (tf) { std::fprintf(tf, "assemble_s %.6f\n", t_assemble); std::fclose(tf); }
        }

        FILE *fp = std::fopen((prefix + ".assemble.dump").c_str(), "wb");
        if (!fp) return 1;
        dump_matrix(fp, A);
        rec_f64(fp, "B", B->shim_data());
        std::vector<int32_t> bg;
        for (auto g : B->getMap()->getNodeElementList()) bg.push_back((int32_t)g);
        rec_i32(fp, "B_rows", bg);
        std::vector<int32_t> mk, mv;
        for (const auto &kv : io.globalIDMap) { mk.push_back((int32_t)kv.first); mv.push_back((int32_t)kv.second); }
        rec_i32(fp, "idmap_reduced", mk);
        rec_i32(fp, "idmap_original", mv);
        dump_nodesets(fp, io.nodeSetMap);
        std::fclose(fp);

        if (!ref_dump_system(A, B, prefix + ".mpi-proc-0.out")) return 1;

        if (!io.create(prefix + ".solution.exo")) return 1;
        const double t1 = now();
        if (!io.decompose(nparts, false)) { std::fprintf(stderr, "ref_driver: decompose failed\n"); return 1; }
        const double t_decompose = now() - t1, t2 = now();
        // stand-in for the iterates of the Belos loop (BelosMueLuSolver.cpp:113-116): float-exact values keyed on the row id
        const auto gids = X->getMap()->getNodeElementList();
        for (int step = 0; step < 2; ++step) {
            for (int l = 0; l < gids.size(); ++l) X->shim_data()[(size_t)l] = 0.25 + 0.5 * (double)gids[l] + 4096.0 * step;
            if (!io.writeSolution(X, step, false)) return 1;
        }
        FILE *tf = std::fopen((prefix + ".timing").c_str(), "a");
        if (tf) { std::fprintf(tf, "decompose_s %.6f\nwrite_solution_2_steps_s %.6f\n", t_decompose, now() - t2); std::fclose(tf); }
    }   // ~IO closes both files: the shim flushes <prefix>.solution.exo.shimdump
    if (get_matrix) {
        ExodusIO::IO io;
        if (!io.open(input, true)) return 1;
        Teuchos::RCP<Tpetra::CrsMatrix<>> A;
        std::map<int, std::set<idx_t>> nodeSetMap;
        if (!io.getMatrix(&A, nodeSetMap, false)) { std::fprintf(stderr, "ref_driver: getMatrix failed\n"); return 3; }
        FILE *fp = std::fopen((prefix + ".getmatrix.dump").c_str(), "wb");
        if (!fp) return 1;
        dump_matrix(fp, A);
        dump_nodesets(fp, nodeSetMap);
        std::ostringstream log;
        log.precision(17);
        const std::vector<double> lambda{PowerMethod<Tpetra::CrsMatrix<>>::run(*A, 500, 1.0e-2, log)};    // ExodusMatrixTest.cpp:171
        rec_f64(fp, "pm_lambda", lambda);
        const std::string text = log.str();
        rec(fp, "pm_log", "str", text.data(), text.size(), 1);
        std::fclose(fp);
    }
    return 0;
}
