// Zoltan2_Adapter.hpp — single-rank stand-in for the Zoltan2 repartitioning of ExodusIO.hpp:644-690: with one
// rank the "parmetis" partition is the identity, so applyPartitioningSolution hands back a copy.
// TEST INFRASTRUCTURE (oracle/ref_shim/README.md).
#pragma once
#include "Tpetra_Core.hpp"

namespace Zoltan2 {
struct PartitioningSolution {};
template <class Adapter>
class PartitioningProblem {
   public:
    PartitioningProblem(Adapter *, Teuchos::ParameterList *) {}
    void solve() {}
    const PartitioningSolution &getSolution() const { return sol_; }
   private:
    PartitioningSolution sol_;
};
template <class Matrix>
class XpetraCrsMatrixAdapter {
   public:
    explicit XpetraCrsMatrixAdapter(const Teuchos::RCP<Matrix> &) {}
    void applyPartitioningSolution(const Matrix &in, Teuchos::RCP<Matrix> &out, const PartitioningSolution &) const {
        out = Teuchos::rcp(new Matrix(in));
    }
};
template <class MV>
class XpetraMultiVectorAdapter {
   public:
    explicit XpetraMultiVectorAdapter(const Teuchos::RCP<MV> &) {}
    void applyPartitioningSolution(const MV &in, Teuchos::RCP<MV> &out, const PartitioningSolution &) const {
        out = Teuchos::rcp(new MV(in));
    }
};
}  // namespace Zoltan2
