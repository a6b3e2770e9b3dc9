// heat_assemble_test — CLI mirror of the reference's exec/ExodusAssembleTest (ExodusAssembleTest.cpp:4-40): same
// flags (--input --verbose/--no-verbose), io.open -> io.assemble, nothing else.  The reference insists on >= 2 MPI
// ranks (:18-21); here one process drives one GPU (several: through the C ABI).
#include <cstdlib>
#include <iostream>
#include <string>

#include "ExodusIO_b200.hpp"

int main(int argc, char *argv[]) {
    std::string inputFile;
    bool verbose = false;
    heat::Options opt;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a.compare(0, 8, "--input=") == 0) inputFile = a.substr(8);
        else if (a.compare(0, 9, "--device=") == 0) opt.device = std::atoi(a.c_str() + 9);
        else if (a.compare(0, 11, "--operator=") == 0) opt.op_mode = a.substr(11) == "p1" ? HEAT_OP_P1_FEM : HEAT_OP_GRAPH_LAPLACIAN;
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
    heat::Matrix A;
    heat::Vector X, B;
    if (!io.assemble(&A, &X, &B, verbose)) {
        std::cerr << "Process #" << rank << ": Failed to getMatrix!!" << std::endl;
        return EXIT_FAILURE;
    }
    return 0;
}
