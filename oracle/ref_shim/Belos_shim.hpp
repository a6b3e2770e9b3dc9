// Belos_shim.hpp — COMPILE-ONLY stand-ins for the Belos / Ifpack2 names BelosMueLuSolver.cpp mentions, so that the
// reference's own printCrsMatrix / printMultiVector (BelosMueLuSolver.cpp:37-84) can be compiled from that file
// and run.  Nothing here solves anything: belosSolver() (BelosMueLuSolver.cpp:87-139) is never called by the
// driver, and every entry point it would reach aborts with a message — the Krylov solve stays unpinned.
// TEST INFRASTRUCTURE (oracle/ref_shim/README.md).
#pragma once
#include <fstream>

#include "Tpetra_Core.hpp"

[[noreturn]] inline void shim_no_trilinos(const char *what) {
    std::fprintf(stderr, "ref_shim: %s — Belos / Ifpack2 do not exist in this image\n", what);
    std::abort();
}

namespace Ifpack2 {
template <class Scalar = double>
class Preconditioner : public Tpetra::Operator<> {
   public:
    void setParameters(const Teuchos::ParameterList &) {}
    void initialize() { shim_no_trilinos("Ifpack2::Preconditioner::initialize"); }
    void compute() { shim_no_trilinos("Ifpack2::Preconditioner::compute"); }
};
class Factory {
   public:
    template <class MatrixPtr>
    Teuchos::RCP<Preconditioner<>> create(const char *, const MatrixPtr &) { shim_no_trilinos("Ifpack2::Factory::create"); }
};
}  // namespace Ifpack2

namespace Belos {
enum ReturnType { Converged, Unconverged };
enum ResetType { Problem, RecycleSubspace };
template <class S, class MV, class OP>
class LinearProblem {
   public:
    LinearProblem(const Teuchos::RCP<OP> &, const Teuchos::RCP<MV> &, const Teuchos::RCP<const MV> &) {}
    void setRightPrec(const Teuchos::RCP<const OP> &) {}
    bool setProblem() { return true; }
};
template <class S, class MV, class OP>
class SolverManager {
   public:
    void setProblem(const Teuchos::RCP<LinearProblem<S, MV, OP>> &) {}
    ReturnType solve() { shim_no_trilinos("Belos::SolverManager::solve"); }
    double achievedTol() const { return 0.0; }
    void reset(ResetType) {}
};
template <class S, class MV, class OP>
class SolverFactory {
   public:
    Teuchos::RCP<SolverManager<S, MV, OP>> create(const char *, const Teuchos::RCP<Teuchos::ParameterList> &) {
        shim_no_trilinos("Belos::SolverFactory::create");
    }
};
}  // namespace Belos
