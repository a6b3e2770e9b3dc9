// Tpetra_Core.hpp — single-rank stand-in for the Tpetra classes /root/reference/ExodusIO.hpp uses (Map,
// CrsMatrix, MultiVector).  Semantics kept: contiguous / list maps with invalid() for missing ids,
// insertGlobalValues SUMS repeated (row, col) entries as fillComplete does, rows come back sorted by column.
// The BLAS-1 / apply members exist for PowerMethod::run (ExodusMatrixTest.cpp:56-129): plain sequential loops in
// index order; randomize() is the counter-based U(-1,1) generator of SURVEY.md §8d keyed on the local index
// (seed $REF_SHIM_SEED, default 12345) so that a run can be repeated by the oracle and the product.
// TEST INFRASTRUCTURE (oracle/ref_shim/README.md).
#pragma once
#include <algorithm>
#include <cstddef>
#include <cmath>
#include <cstdint>
#include <cstdlib>

#include "Teuchos_shim.hpp"

namespace Tpetra {

typedef size_t global_size_t;
namespace Details { namespace DefaultTypes {
typedef double scalar_type;
typedef int local_ordinal_type;
typedef long long global_ordinal_type;
typedef void node_type;
} }

enum LookupStatus { AllIDsPresent, IDNotPresent };

inline Teuchos::RCP<const Teuchos::Comm<int>> getDefaultComm() {
    static Teuchos::RCP<const Teuchos::Comm<int>> c(new Teuchos::Comm<int>());
    return c;
}

struct ScopeGuard { ScopeGuard(int *, char ***) {} };

template <class LO = Details::DefaultTypes::local_ordinal_type, class GO = Details::DefaultTypes::global_ordinal_type,
          class Node = Details::DefaultTypes::node_type>
class Map {
   public:
    typedef LO local_ordinal_type;
    typedef GO global_ordinal_type;
    // contiguous uniform map (one rank: everything)
    Map(global_size_t numGlobal, GO indexBase, const Teuchos::RCP<const Teuchos::Comm<int>> &comm) : comm_(comm) {
        for (global_size_t i = 0; i < numGlobal; ++i) gids_.push_back(indexBase + (GO)i);
        index();
    }
    // contiguous map from local counts (one rank: 0..numLocal-1)
    Map(global_size_t, size_t numLocal, GO indexBase, const Teuchos::RCP<const Teuchos::Comm<int>> &comm) : comm_(comm) {
        for (size_t i = 0; i < numLocal; ++i) gids_.push_back(indexBase + (GO)i);
        index();
    }
    // arbitrary list of owned global ids
    Map(global_size_t, const Teuchos::Array<GO> &indices, GO, const Teuchos::RCP<const Teuchos::Comm<int>> &comm)
        : gids_(indices.begin(), indices.end()), comm_(comm) { index(); }

    LO getMinLocalIndex() const { return 0; }
    LO getMaxLocalIndex() const { return gids_.empty() ? Teuchos::OrdinalTraits<LO>::invalid() : (LO)gids_.size() - 1; }
    GO getGlobalElement(LO l) const { return (l < 0 || (size_t)l >= gids_.size()) ? Teuchos::OrdinalTraits<GO>::invalid() : gids_[(size_t)l]; }
    LO getLocalElement(GO g) const { auto it = g2l_.find(g); return it == g2l_.end() ? Teuchos::OrdinalTraits<LO>::invalid() : it->second; }
    bool isNodeGlobalElement(GO g) const { return g2l_.count(g) != 0; }
    bool isNodeLocalElement(LO l) const { return l >= 0 && (size_t)l < gids_.size(); }
    size_t getNodeNumElements() const { return gids_.size(); }
    global_size_t getGlobalNumElements() const { return gids_.size(); }
    bool isDistributed() const { return false; }
    bool isOneToOne() const { return true; }
    Teuchos::Array<GO> getNodeElementList() const { return Teuchos::Array<GO>(gids_.begin(), gids_.end()); }
    Teuchos::RCP<const Teuchos::Comm<int>> getComm() const { return comm_; }
    LookupStatus getRemoteIndexList(const Teuchos::Array<GO> &gids, Teuchos::Array<int> &nodeIDs, Teuchos::Array<LO> &lids) const {
        LookupStatus st = AllIDsPresent;
        for (int i = 0; i < gids.size(); ++i) {
            const LO l = getLocalElement(gids[i]);
            lids[i] = l;
            nodeIDs[i] = (l == Teuchos::OrdinalTraits<LO>::invalid()) ? -1 : 0;
            if (nodeIDs[i] < 0) st = IDNotPresent;
        }
        return st;
    }
    template <class OS> void describe(OS &os, Teuchos::EVerbosityLevel) const { os << "Tpetra::Map (shim): " << gids_.size() << " ids on one rank\n"; }

   private:
    void index() { for (size_t i = 0; i < gids_.size(); ++i) g2l_[gids_[i]] = (LO)i; }
    std::vector<GO> gids_;
    std::map<GO, LO> g2l_;
    Teuchos::RCP<const Teuchos::Comm<int>> comm_;
};

template <class Scalar, class LO, class GO, class Node> class MultiVector;

template <class Scalar = Details::DefaultTypes::scalar_type, class LO = Details::DefaultTypes::local_ordinal_type,
          class GO = Details::DefaultTypes::global_ordinal_type, class Node = Details::DefaultTypes::node_type>
class Operator {
   public:
    virtual ~Operator() = default;
};

template <class Scalar = Details::DefaultTypes::scalar_type, class LO = Details::DefaultTypes::local_ordinal_type,
          class GO = Details::DefaultTypes::global_ordinal_type, class Node = Details::DefaultTypes::node_type>
class CrsMatrix : public Operator<Scalar, LO, GO, Node> {
   public:
    typedef Scalar scalar_type;
    typedef LO local_ordinal_type;
    typedef GO global_ordinal_type;
    typedef Node node_type;
    typedef Map<LO, GO, Node> map_type;
    CrsMatrix(const Teuchos::RCP<const map_type> &rowMap, size_t) : rowMap_(rowMap) {}
    void insertGlobalValues(GO row, const Teuchos::Array<GO> &cols, const Teuchos::Array<Scalar> &vals) {
        auto &r = rows_[row];
        for (int i = 0; i < cols.size(); ++i) r[cols[i]] += vals[i];      // duplicates are summed (fillComplete)
    }
    void replaceGlobalValues(GO row, const Teuchos::Array<GO> &cols, const Teuchos::Array<Scalar> &vals) {
        auto &r = rows_[row];
        for (int i = 0; i < cols.size(); ++i) r[cols[i]] = vals[i];
    }
    void fillComplete(const Teuchos::RCP<const map_type> &, const Teuchos::RCP<const map_type> &) {}
    void fillComplete() {}
    void resumeFill() {}
    size_t getNumEntriesInGlobalRow(GO row) const { auto it = rows_.find(row); return it == rows_.end() ? 0 : it->second.size(); }
    void getGlobalRowCopy(GO row, Teuchos::Array<GO> &cols, Teuchos::Array<Scalar> &vals, size_t &n) const {
        n = 0;
        auto it = rows_.find(row);
        if (it == rows_.end()) return;
        for (auto &kv : it->second) { cols[n] = kv.first; vals[n] = kv.second; ++n; }
    }
    global_size_t getGlobalNumRows() const { return rowMap_->getGlobalNumElements(); }
    Teuchos::RCP<const map_type> getRowMap() const { return rowMap_; }
    Teuchos::RCP<const map_type> getDomainMap() const { return rowMap_; }
    Teuchos::RCP<const map_type> getRangeMap() const { return rowMap_; }
    Teuchos::RCP<const map_type> getMap() const { return rowMap_; }
    template <class OS> void describe(OS &os, Teuchos::EVerbosityLevel) const { os << "Tpetra::CrsMatrix (shim): " << rows_.size() << " rows\n"; }
    const std::map<GO, std::map<GO, Scalar>> &shim_rows() const { return rows_; }      // for the driver's dump
    // y := A x, row by row, columns ascending
    void apply(const MultiVector<Scalar, LO, GO, Node> &x, MultiVector<Scalar, LO, GO, Node> &y) const {
        for (auto &v : y.shim_data()) v = Scalar(0);
        for (const auto &r : rows_) {
            Scalar acc = Scalar(0);
            for (const auto &cv : r.second) acc += cv.second * x.shim_data()[(size_t)rowMap_->getLocalElement(cv.first)];
            y.shim_data()[(size_t)rowMap_->getLocalElement(r.first)] = acc;
        }
    }

   private:
    Teuchos::RCP<const map_type> rowMap_;
    std::map<GO, std::map<GO, Scalar>> rows_;
};

template <class Scalar = Details::DefaultTypes::scalar_type, class LO = Details::DefaultTypes::local_ordinal_type,
          class GO = Details::DefaultTypes::global_ordinal_type, class Node = Details::DefaultTypes::node_type>
class MultiVector {
   public:
    typedef Map<LO, GO, Node> map_type;
    typedef Scalar mag_type;
    MultiVector(const Teuchos::RCP<const map_type> &map, size_t numVecs = 1)
        : map_(map), v_(map->getNodeNumElements() * numVecs, Scalar(0)), nvec_(numVecs) {}
    void randomize() {
        const char *e = std::getenv("REF_SHIM_SEED");
        const uint64_t seed = e ? std::strtoull(e, nullptr, 10) : 12345ull;
        for (size_t l = 0; l < v_.size(); ++l) {
            uint64_t z = ((uint64_t)l ^ (0x9E3779B97F4A7C15ull * seed)) + 0x9E3779B97F4A7C15ull;      // splitmix64
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            z ^= z >> 31;
            v_[l] = (double)(z >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
        }
    }
    Scalar dot(const MultiVector &o) const { Scalar s = 0; for (size_t i = 0; i < v_.size(); ++i) s += v_[i] * o.v_[i]; return s; }
    mag_type norm2() const { return std::sqrt(dot(*this)); }
    void scale(Scalar alpha, const MultiVector &src) { for (size_t i = 0; i < v_.size(); ++i) v_[i] = alpha * src.v_[i]; }
    // this := alpha*A + beta*B + gamma*this
    void update(Scalar alpha, const MultiVector &A, Scalar beta, const MultiVector &B, Scalar gamma) {
        for (size_t i = 0; i < v_.size(); ++i) v_[i] = alpha * A.v_[i] + beta * B.v_[i] + gamma * v_[i];
    }
    Scalar *get1dViewNonConst() { return v_.data(); }
    Teuchos::Array<Scalar> get1dView() const { return Teuchos::Array<Scalar>(v_.begin(), v_.end()); }
    Teuchos::RCP<const map_type> getMap() const { return map_; }
    size_t getNumVectors() const { return nvec_; }
    Teuchos::RCP<const MultiVector> getVector(size_t j) const {            // column j as a one-column copy
        auto *v = new MultiVector(map_, 1);
        const size_t n = map_->getNodeNumElements();
        for (size_t i = 0; i < n; ++i) v->v_[i] = v_[j * n + i];
        return Teuchos::RCP<const MultiVector>(std::shared_ptr<const MultiVector>(v));
    }
    global_size_t getGlobalLength() const { return map_->getGlobalNumElements(); }
    std::vector<Scalar> &shim_data() { return v_; }
    const std::vector<Scalar> &shim_data() const { return v_; }

   private:
    Teuchos::RCP<const map_type> map_;
    std::vector<Scalar> v_;
    size_t nvec_;
};

template <class Scalar = Details::DefaultTypes::scalar_type, class LO = Details::DefaultTypes::local_ordinal_type,
          class GO = Details::DefaultTypes::global_ordinal_type, class Node = Details::DefaultTypes::node_type>
using Vector = MultiVector<Scalar, LO, GO, Node>;

}  // namespace Tpetra
