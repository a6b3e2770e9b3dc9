// heat_matrix_test — CLI mirror of the reference's exec/ExodusMatrixTest (ExodusMatrixTest.cpp:132-170):
// same flags (--input --verbose/--no-verbose), same call order io.open -> io.getMatrix -> power
// method (500 iterations, tolerance 1e-2, :163) and the same progress lines.  Single process, single
// GPU: the reference insists on >= 2 MPI ranks (:149-152) because ParMETIS needs them; the rows of a
// multi-GPU getMatrix are distributed through the C ABI, one process per GPU.
#include <cstdlib>
#include <iostream>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "ExodusIO_b200.hpp"

int main(int argc, char *argv[]) {
    std::string inputFile;
    bool verbose = false;
    heat::Options opt;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a.compare(0, 8, "--input=") == 0) inputFile = a.substr(8);
        else if (a.compare(0, 9, "--device=") == 0) opt.device = std::atoi(a.c_str() + 9);
        else if (a == "--verbose") verbose = true;
        else if (a == "--no-verbose") verbose = false;
        else { std::cerr << "unknown option '" << a << "'" << std::endl; return EXIT_FAILURE; }
    }
    const int rank = 0;
    if (inputFile.empty()) {
        std::cerr << "No input file was provided; use the '--input' parameter!" << std::endl;
        return EXIT_FAILURE;
    }
    ExodusIO::IO io(opt);
    if (!io.open(inputFile, true)) {
        std::cerr << "Process #" << rank << ": Failed to open input Exodus file '" << inputFile << "'" << std::endl;
        return EXIT_FAILURE;
    }
    heat::Matrix ret;
    std::map<int, std::set<int64_t>> nodeSetMap;
    if (!io.getMatrix(&ret, nodeSetMap, verbose)) {
        std::cerr << "Process #" << rank << ": Failed to getMatrix!!" << std::endl;
        return EXIT_FAILURE;
    }
    heat_matrix_info mi;
    heat_matrix_get_info(ret->h, &mi);
    std::cout << "Global dimensions: [" << mi.n_global << ", " << mi.n_global << "], global number of entries: "
              << mi.nnz_global << ", max entries per row: " << mi.max_row_len << std::endl;
    if (verbose)
        for (const auto &kv : nodeSetMap) std::cout << "Nodeset " << kv.first << ": " << kv.second.size() << " owned nodes" << std::endl;
    // PowerMethod<...>::run(*ret, 500, 1.0e-2, std::cout)   (ExodusMatrixTest.cpp:163)
    const int niters = 500;
    std::vector<double> rep(3 * (size_t)(niters / 50 + 2));
    heat_power_info pi{};
    pi.report_buf = rep.data(); pi.report_capacity = niters / 50 + 2;
    if (heat_power_method(io.ctx(), ret->h, niters, 1.0e-2, 12345, &pi)) {
        std::cerr << "power method: " << heat_last_error() << std::endl;
        return EXIT_FAILURE;
    }
    for (int k = 0; k < pi.report_count; ++k)
        std::cout << "Iteration " << (int)rep[3 * k] << ":" << std::endl
                  << "- lambda = " << rep[3 * k + 1] << std::endl
                  << "- ||A*q - lambda*q||_2 = " << rep[3 * k + 2] << std::endl;
    if (pi.converged) std::cout << "Converged after " << pi.iters << " iterations" << std::endl;
    else std::cout << "Failed to converge after " << niters << " iterations" << std::endl;
    std::cout << "Lambda = " << pi.lambda << std::endl;
    return 0;
}
