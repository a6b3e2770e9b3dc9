// heat_solver — CLI mirror of the reference's exec/BelosMueLuSolver (BelosMueLuSolver.cpp:141-218):
// same flags (--input --solution --iterations --tolerance --verbose/--no-verbose --outputPrefix
// --reportAfterIterations) and the same call order
//     io.open -> io.assemble -> (rank 0) io.create + io.decompose(max(2, ranks)) -> belosSolver
// plus the knobs the reference hard-codes: --operator graph|p1, --solver cg|cg1|gmres, --prec
// none|jacobi|chebyshev|ilu0, --gmres-restart M, --reference-loop (GMRES + ILU with one iteration per
// solve() call, as the reference is written), --partitions N, --write-every K, --device D.
// Single process, single GPU (multi-GPU runs are driven by one process per GPU through the C ABI;
// see INTEGRATION.md).
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "ExodusIO_b200.hpp"

static bool flag_value(const std::string &arg, const char *name, std::string &out) {
    const std::string key = std::string("--") + name + "=";
    if (arg.compare(0, key.size(), key) == 0) { out = arg.substr(key.size()); return true; }
    return false;
}

int main(int argc, char *argv[]) {
    std::string inputFile, outputPrefix = "mpi-proc-", solution = "solution.exo", v;
    bool verbose = false, dump = false;
    size_t numIterations = 300;            // BelosMueLuSolver.cpp:149
    double tolerance = 1e-14;              // :151
    int partitions = 0;
    heat::Options opt;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (flag_value(a, "input", v)) inputFile = v;
        else if (flag_value(a, "solution", v)) solution = v;
        else if (flag_value(a, "outputPrefix", v)) outputPrefix = v;
        else if (flag_value(a, "iterations", v)) numIterations = (size_t)std::strtoull(v.c_str(), nullptr, 10);
        else if (flag_value(a, "reportAfterIterations", v)) { /* parsed but unused, as in the reference (D8) */ }
        else if (flag_value(a, "tolerance", v)) tolerance = std::strtod(v.c_str(), nullptr);
        else if (flag_value(a, "partitions", v)) partitions = std::atoi(v.c_str());
        else if (flag_value(a, "device", v)) opt.device = std::atoi(v.c_str());
        else if (flag_value(a, "write-every", v)) opt.write_every = std::atoi(v.c_str());
        else if (flag_value(a, "cheb-degree", v)) opt.cheb_degree = std::atoi(v.c_str());
        else if (flag_value(a, "cheb-lambda-max", v)) opt.cheb_lambda_max = std::strtod(v.c_str(), nullptr);
        else if (flag_value(a, "operator", v)) opt.op_mode = (v == "p1") ? HEAT_OP_P1_FEM : HEAT_OP_GRAPH_LAPLACIAN;
        else if (flag_value(a, "solver", v))
            opt.solver = (v == "cg1") ? HEAT_SOLVER_CG_SINGLE_REDUCE : (v == "gmres") ? HEAT_SOLVER_GMRES : HEAT_SOLVER_CG;
        else if (flag_value(a, "prec", v))
            opt.prec = (v == "none") ? HEAT_PREC_NONE : (v == "chebyshev") ? HEAT_PREC_CHEBYSHEV
                     : (v == "ilu0" || v == "ilut") ? HEAT_PREC_ILU0 : HEAT_PREC_JACOBI;
        else if (flag_value(a, "gmres-restart", v)) opt.gmres_restart = std::atoi(v.c_str());
        else if (a == "--reference-loop") {      // the reference as shipped: GMRES(1) + ILU, one iteration per solve()
            opt.solver = HEAT_SOLVER_GMRES; opt.prec = HEAT_PREC_ILU0; opt.literal_loop = true;
        }
        else if (a == "--verbose") verbose = true;
        else if (a == "--no-verbose") verbose = false;
        else if (a == "--dump") dump = true;
        else if (a == "--real4") opt.output_word_size = 4;                       // float32 output like a float-real_t reference build
        else if (a == "--reference-output") { opt.output_word_size = 4; opt.output_largest_nodeset_id = true; }
        else { std::cerr << "unknown option '" << a << "'" << std::endl; return EXIT_FAILURE; }
    }
    if (inputFile.empty()) {
        std::cerr << "No input file was provided; use the '--input' parameter!" << std::endl;
        return EXIT_FAILURE;
    }
    const int rank = 0, ranks = 1;
    ExodusIO::IO io(opt);
    if (!io.open(inputFile, true)) {
        std::cerr << "Process #" << rank << ": Failed to open input Exodus file '" << inputFile << "'" << std::endl;
        return EXIT_FAILURE;
    }
    heat::Matrix A;
    heat::Vector X, B;
    if (!io.assemble(&A, &X, &B, verbose)) {
        std::cerr << "Process #" << rank << ": Failed to getMatrix!!" << std::endl;
        return EXIT_FAILURE;
    }
    std::ofstream output;
    if (dump) {
        // the reference's only observability format (BelosMueLuSolver.cpp:37-84, :169-215): "<prefix><rank>.out"
        output.open(outputPrefix + std::to_string(rank) + ".out");
        if (!output.good()) {
            std::cerr << "Process #" << rank << ": Failed to open output file '" << outputPrefix << rank << ".out'" << std::endl;
            return EXIT_FAILURE;
        }
        output << "[Laplacian: A]" << std::endl;
        heat::printCrsMatrix(A, output);
        output << "[RHS: B]" << std::endl;
        heat::printMultiVector(io.ctx(), A, B, output);
    }
    if (!io.create(solution)) {
        std::cerr << "Process #" << rank << ": Failed to create output file '" << solution << "'" << std::endl;
    }
    io.decompose(partitions > 0 ? partitions : std::max(2, ranks), verbose);      // :209
    belosSolver(A, X, B, numIterations, tolerance, io, verbose);
    if (dump) {
        output << "[Solution: X]" << std::endl;
        heat::printMultiVector(io.ctx(), A, X, output);
    }
    return 0;
}
