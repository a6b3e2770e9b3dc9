// ref_dump.cpp — second translation unit of oracle/_ref/ref_driver: compiles /root/reference/BelosMueLuSolver.cpp
// (unmodified, its main() renamed and never called) for its printCrsMatrix / printMultiVector, and writes the
// "[Laplacian: A]" / "[RHS: B]" sections exactly as its main() does (BelosMueLuSolver.cpp:190-199).  ExodusIO.hpp
// has no include guard, hence a TU of its own.  TEST INFRASTRUCTURE (oracle/ref_shim/README.md).
#define main reference_solver_main
#include "BelosMueLuSolver.cpp"
#undef main

bool ref_dump_system(const Teuchos::RCP<Tpetra::CrsMatrix<>> &A, const Teuchos::RCP<Tpetra::MultiVector<>> &B, const std::string &path) {
    std::ofstream output(path);
    if (!output.good()) return false;
    output << "[Laplacian: A]" << std::endl;
    printCrsMatrix(A, output);
    output << "[RHS: B]" << std::endl;
    printMultiVector(B, output);
    return true;
}
